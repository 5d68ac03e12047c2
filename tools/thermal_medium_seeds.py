"""How far apart are two random realisations of tst3d_v_o2_thermal_plasma_medium (128^3 cells, 64 ppc regular, 500 steps),
in the reference's validation metric?     python tools/thermal_medium_seeds.py [seeds...]      (on a B200, ~70 s per seed)

Runs the benchmark on the GPU path from the reference's particle streams for random_seed 0, 1, 2 and prints, for
Ukin/avg, Uelm/avg, Utot/avg: max |seed a - seed b| for every pair and max |seed s - stored reference curve|
(validation/references/tst3d_v_o2_thermal_plasma_medium.py.txt, tolerance 1e-3 each).  The reference validates against
ONE stored realisation; if two seeds of the present sources differ from each other in Uelm/avg by more than that
tolerance, the tolerance cannot be met by any implementation that does not reproduce the stored run's random stream.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if __name__ == "__main__":
    from smilei_b200 import namelist
    from smilei_b200.simulation import Simulation
    from test_gpu_simulation import THERMAL_MEDIUM
    seeds = [int(a) for a in sys.argv[1:]] or [0, 1, 2]
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_validation_thermal_plasma_medium.npz"))
    curves, raw = {}, {}
    for seed in seeds:
        t0 = time.time()
        params = namelist.load_namelist(THERMAL_MEDIUM, is_source=True)
        params.random_seed = seed
        sim = Simulation(params)
        sim.create_particles(reference_streams=True)
        uk, ue = sim.scalars()
        K, E = [float(uk.sum())], [ue]
        for _, k, e in sim.run(500, scalars_every=10):
            K.append(float(k.sum()))
            E.append(e)
        sim.close()
        K, E = np.asarray(K), np.asarray(E)
        U = K + E
        curves[seed] = {"ukin": K / K.mean(), "uelm": E / E.mean(), "utot": U / U.mean()}
        raw[seed] = {"seconds": time.time() - t0}
        print("seed", seed, "done in", raw[seed]["seconds"], file=sys.stderr, flush=True)
    out = {"seconds_per_seed (creation on the host + 500 steps)": {str(s): raw[s]["seconds"] for s in raw}}
    for name in ("ukin", "uelm", "utot"):
        for s in curves:
            out[f"{name}: seed {s} vs stored reference"] = float(np.max(np.abs(curves[s][name] - gold[name])))
        for i, a in enumerate(seeds):
            for b in seeds[i + 1:]:
                out[f"{name}: seed {a} vs seed {b}"] = float(np.max(np.abs(curves[a][name] - curves[b][name])))
    print(json.dumps(out, indent=1))
