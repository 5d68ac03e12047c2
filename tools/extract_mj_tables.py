"""Write smilei_b200/data/mj_tables.npz: the two 1000-point tables of ParticleCreator::maxwellJuttner
(ln of the inverse cumulative functions, src/Particles/ParticleCreator.cpp:1074-1365), read out of the
reference build oracle/_ref/libsmilei_ref.so (needs /root/reference; `make -C oracle ref` first).

    python tools/extract_mj_tables.py

They are input DATA of the sampling method: a run that has to start from the reference's particles for a
given random_seed needs these exact 2000 doubles, they cannot be recomputed bit for bit.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if __name__ == "__main__":
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsmilei_ref.so"))
    out = {}
    for name in ("lnInvF", "lnInvH"):
        f = getattr(ref, "ref_" + name)
        f.restype = C.POINTER(C.c_double)
        out[name] = np.ctypeslib.as_array(f(), shape=(1000,)).copy()
    np.savez_compressed(os.path.join(ROOT, "smilei_b200", "data", "mj_tables.npz"), **out)
    print({k: (v[0], v[-1]) for k, v in out.items()})
