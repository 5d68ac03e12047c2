"""Throughput of the whole step against the number of particles per cell (same total particle count).

    python tools/ppc_probe.py            (on a B200)

Thermal plasma as in bench.py, created on the device at ppc = 1, 2, 4, 8, 16, 32, 64 with the box sized to 2 x 64..134 M particles;
5 warm-up steps, 5 timed steps (CUDA events), plus the time of the sort alone.  Prints one JSON line per case.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    import bench
    from smilei_b200 import namelist
    from smilei_b200.simulation import Simulation
    T, dx, dt = bench.plasma_constants()
    cases = (((1, 1, 1), 400), ((2, 1, 1), 320), ((2, 2, 1), 256), ((2, 2, 2), 256), ((4, 2, 2), 200), ((4, 4, 2), 160), ((4, 4, 4), 128))
    if len(sys.argv) > 2:                  # one case: ppc as a,b,c and cells per dimension (e.g. under ncu)
        cases = ((tuple(int(v) for v in sys.argv[1].split(",")), int(sys.argv[2])),)
    for ppc, ncell in cases:
        params = namelist.load_namelist(bench.namelist_source([ncell] * 3, 2, "boris"), is_source=True)
        sim = Simulation(params, capacity_factor=1.08)
        sim.init_thermal(ppc, density=1.0, temperature=T, seed=0)
        npart = sum(sim.patch.species_count(s) for s in range(2))
        for _ in range(5):
            sim.step()
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        for _ in range(5):
            sim.step()
        e[1].record()
        e[2].record()
        for _ in range(5):
            for s in range(2):
                sim.patch.sort(s)          # sorted already: the detection pass only
        e[3].record()
        torch.cuda.synchronize()
        ms = e[0].elapsed_time(e[1]) / 5
        print(json.dumps({"ppc": ppc[0] * ppc[1] * ppc[2], "ncell": ncell, "particles": npart, "ms_per_step": ms,
                          "pushes_per_s": npart / ms * 1e3, "ms_resort_sorted_both_species": e[2].elapsed_time(e[3]) / 5}),
              flush=True)
        sim.close()
        del sim
        torch.cuda.empty_cache()
