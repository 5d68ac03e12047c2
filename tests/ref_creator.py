"""ctypes access to the reference's ParticleCreator / Hilbert functions inside oracle/_ref/libsmilei_ref.so
(oracle/ref_build/ref_creator_harness.cpp).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

import oracle_lib as ol


class RefCreator:
    def __init__(self):
        self.lib = C.CDLL(ol.REF_SO)
        self.lib.ref_hilbert_index3d.restype = C.c_uint

    def hilbert(self, m, p):
        return self.lib.ref_hilbert_index3d(C.c_uint(m[0]), C.c_uint(m[1]), C.c_uint(m[2]), int(p[0]), int(p[1]), int(p[2]))

    def patch(self, state, position_init, momentum_init, box, box_min, cell, nppc, n_real, charge, temperature, mass,
              regular_number=None, positions=None):
        """One species in one patch: the loop of ParticleCreator::create (ParticleCreator.cpp:300-338) around the
        reference's per-cell static functions."""
        st = C.c_uint(state)
        cols = {k: [] for k in ("x", "y", "z", "px", "py", "pz", "w", "q")}
        ip = 0
        reg = (C.c_int * 3)(*(regular_number or [0, 0, 0]))
        for i in range(box[0]):
            for j in range(box[1]):
                for k in range(box[2]):
                    n = int(nppc[i, j, k])
                    if not n_real[i, j, k] > 0 or n <= 0:
                        continue
                    idx = (C.c_double * 3)(*[np.float64(a) * np.float64(cell[d]) + np.float64(box_min[d]) + 0.0
                                             for d, a in enumerate((i, j, k))])
                    o = {c: np.zeros(n) for c in ("x", "y", "z", "px", "py", "pz", "w")}
                    o["q"] = np.zeros(n, dtype=np.int16)
                    if position_init is None:
                        for c, v in zip("xyz", positions):
                            o[c][:] = v[ip:ip + n]
                    T = float(temperature[i, j, k])
                    self.lib.ref_create_cell(
                        C.byref(st), (position_init or "").encode(), momentum_init.encode(), C.c_uint(n), idx,
                        (C.c_double * 3)(*cell), C.c_double(mass), (C.c_double * 3)(T, T, T),
                        C.c_double(n_real[i, j, k]), C.c_double(charge[i, j, k]), reg,
                        *[o[c].ctypes.data_as(C.c_void_p) for c in ("x", "y", "z", "px", "py", "pz", "w", "q")])
                    for c in cols:
                        cols[c].append(o[c])
                    ip += n
        return {c: np.concatenate(v) for c, v in cols.items()}, st.value
