"""tests/golden/ref_streams.npz: initial particles and patch numbering produced by the REFERENCE's own
ParticleCreator / Random / Hilbert_functions (oracle/_ref, needs /root/reference) for two patches of
benchmarks/gpu/tst3d_v_o2_thermal_plasma_short.py (seed 0, 4x4x4 patches of 8^3 cells, 8 ppc random,
protons then electrons on the patch's stream).

    python tests/golden/make_ref_streams_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

if __name__ == "__main__":
    from ref_creator import RefCreator
    from test_reference_streams import COLS, _one_patch, thermal_short
    ref = RefCreator()
    params = thermal_short()
    out = {}
    for tag, P in (("p000", (0, 0, 0)), ("p312", (3, 1, 2))):
        h = ref.hilbert((2, 2, 2), P)
        a, state = _one_patch(ref.patch, params, h, P, "random")
        out[tag + "_hindex"] = np.int64(h)
        out[tag + "_state"] = np.int64(state)
        for name in ("proton", "electron"):
            for c in COLS:
                out[f"{tag}_{name}_{c}"] = a[name][c]
    out["hilbert_4x4x4"] = np.asarray([ref.hilbert((2, 2, 2), (x, y, z)) for x in range(4) for y in range(4) for z in range(4)])
    out["hilbert_8x2x4"] = np.asarray([ref.hilbert((3, 1, 2), (x, y, z)) for x in range(8) for y in range(2) for z in range(4)])
    np.savez_compressed(os.path.join(HERE, "ref_streams.npz"), **out)
    print("wrote ref_streams.npz:", {k: np.shape(v) for k, v in list(out.items())[:6]})
