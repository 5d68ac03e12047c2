"""Time the laser-wake configuration (BASELINE.json configs[3] shape: Vay pusher, cold 1-ppc plasma, laser through a
Silver-Mueller side, `remove` particles, moving window) on one GPU:  python tools/laser_wake_bench.py [nx ny nz steps]
Prints ms per step before and while the window moves.  Not part of bench.py (that is configs[1])."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from smilei_b200 import namelist  # noqa: E402
from smilei_b200.simulation import Simulation  # noqa: E402

nx, ny, nz, steps = (int(v) for v in (sys.argv[1:5] + ["1024", "128", "128", "200"][len(sys.argv) - 1:]))
SRC = f"""
dx, dtrans, dt = 0.2, 3., 0.19
Main(geometry="3Dcartesian", interpolation_order=2, timestep=dt, number_of_timesteps=100000,
     cell_length=[dx, dtrans, dtrans], number_of_cells=[{nx}, {ny}, {nz}], number_of_patches=[{nx // 8}, 1, 1],
     EM_boundary_conditions=[["silver-muller"]], solve_poisson=False)
MovingWindow(time_start={steps // 2}*dt, velocity_x=0.9997)
Species(name="electron", position_initialization="regular", momentum_initialization="cold", particles_per_cell=1,
        mass=1.0, charge=-1.0, charge_density=0.000494, pusher="vay", boundary_conditions=[["remove", "remove"]]*3)
LaserGaussian3D(box_side="xmin", a0=2., focus=[0., {ny}*dtrans/2., {nz}*dtrans/2.], waist=10.,
                time_envelope=tgaussian(center=2**0.5*19.80, fwhm=19.80))
"""
params = namelist.load_namelist(SRC, is_source=True)
sim = Simulation(params)
sim.create_particles()
npart = sim.n_particles()[0]
sim.run(5)
torch.cuda.synchronize()
for label, n in (("fixed box, laser on", steps // 2 - 5), ("window moving", steps - steps // 2)):
    t0 = time.perf_counter()
    sim.run(n)
    torch.cuda.synchronize()
    dt_ms = (time.perf_counter() - t0) / n * 1e3
    print(f"{label}: {dt_ms:.3f} ms/step, {npart / dt_ms / 1e6:.3f} G pushes/s ({nx}x{ny}x{nz} cells, {npart} particles, "
          f"n_moved = {sim.simWindow.n_moved})")
sim.close()
