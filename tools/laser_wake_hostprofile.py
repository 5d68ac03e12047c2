"""Host-side profile (cProfile) of 20 steps of the laser-wake workload of bench.py on one GPU: where the host thread waits
or works between the kernels.  python tools/laser_wake_hostprofile.py   (on a B200)"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from smilei_b200 import namelist
from smilei_b200.simulation import Simulation
n=256; dt=0.19; warm=3
params = namelist.load_namelist(bench.LASER_WAKE_SRC.format(gsize=[n]*3, npx=n//8, tstart=(warm+0.25)*dt), is_source=True)
sim = Simulation(params, rank_grid=(1,1,1), rank=0, device="cuda:0")
sim.create_particles()
sim.run(warm); torch.cuda.synchronize()
pr=cProfile.Profile(); pr.enable()
t0=time.perf_counter(); sim.run(20); torch.cuda.synchronize(); t1=time.perf_counter()
pr.disable()
print("ms/step wall", (t1-t0)*1e3/20)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
