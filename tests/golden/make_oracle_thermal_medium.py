"""tests/golden/oracle_thermal_medium_curves.npz: Ukin (per species) and Uelm after each of the first NSTEPS steps of
tst3d_v_o2_thermal_plasma_medium at FULL size (128^3 cells, 16^3 patches, 64 ppc regular, 2 x 134 M particles, the
reference's particle streams of seed 0) computed by the CPU ORACLE through the repository's driver (Simulation on
tests/oracle_patch.OraclePatch).  About 4 minutes per step on 8 cores and ~45 GB of host memory.

    python tests/golden/make_oracle_thermal_medium.py [NSTEPS=4]

tests/test_gpu_simulation.py::test_reference_validation_thermal_plasma_medium holds the first steps of the GPU run to
these values: trajectory parity at the benchmark's own size (32-bit index paths, 64 particles per cell).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

if __name__ == "__main__":
    from oracle_patch import OraclePatch
    from smilei_b200 import namelist
    from smilei_b200.simulation import Simulation
    from test_gpu_simulation import THERMAL_MEDIUM
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    t0 = time.time()
    sim = Simulation(namelist.load_namelist(THERMAL_MEDIUM, is_source=True), patch_factory=OraclePatch)
    sim.create_particles(reference_streams=True)
    uk, ue = sim.scalars()
    K, E = [np.asarray(uk, dtype=float)], [ue]
    print("created", time.time() - t0, K[-1], E[-1], flush=True)
    for _, k, e in sim.run(nsteps, scalars_every=1):
        K.append(np.asarray(k, dtype=float))
        E.append(e)
        print("step", len(E) - 1, time.time() - t0, K[-1], E[-1], flush=True)
        np.savez(os.path.join(HERE, "oracle_thermal_medium_curves.npz"), ukin=np.asarray(K), uelm=np.asarray(E))
