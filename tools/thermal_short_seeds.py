"""How far apart are two random realisations of tst3d_v_o2_thermal_plasma_short, in the reference's validation metric?

    python tools/thermal_short_seeds.py      (on a B200)

Runs the benchmark on the GPU path from the reference's particle streams for random_seed 0, 1, 2 and prints, for
Ukin/avg, Uelm/avg, Utot/avg: max |seed a - seed b| for every pair and max |seed s - stored reference curve|.
If the stored curve (validation/references/tst3d_v_o2_thermal_plasma_short.py.txt) is as far from the seed-0 run as
another seed is, it was not produced by the seed-0 stream of the present sources.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if __name__ == "__main__":
    from smilei_b200.simulation import Simulation
    from test_reference_streams import thermal_short
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_validation_thermal_plasma_short.npz"))
    curves = {}
    for seed in (0, 1, 2):
        params = thermal_short()
        params.random_seed = seed
        sim = Simulation(params)
        sim.create_particles(reference_streams=True)
        uk, ue = sim.scalars()
        K, E = [float(uk.sum())], [ue]
        for _, k, e in sim.run(2000, scalars_every=10):
            K.append(float(k.sum()))
            E.append(e)
        sim.close()
        K, E = np.asarray(K), np.asarray(E)
        U = K + E
        curves[seed] = {"ukin": K / K.mean(), "uelm": E / E.mean(), "utot": U / U.mean()}
    out = {}
    for name in ("ukin", "uelm", "utot"):
        for s in curves:
            out[f"{name}: seed {s} vs stored reference"] = float(np.max(np.abs(curves[s][name] - gold[name])))
            out[f"{name}: seed {s} vs stored reference, first 3 samples after t=0"] = \
                [float(v) for v in np.abs(curves[s][name] - gold[name])[1:4]]
        for a, b in ((0, 1), (0, 2), (1, 2)):
            out[f"{name}: seed {a} vs seed {b}"] = float(np.max(np.abs(curves[a][name] - curves[b][name])))
            out[f"{name}: seed {a} vs seed {b}, first 3 samples after t=0"] = \
                [float(v) for v in np.abs(curves[a][name] - curves[b][name])[1:4]]
    print(json.dumps(out, indent=1))
