"""Generate tests/golden/*.npz from the REFERENCE build (oracle/_ref/libsmilei_ref.so).

Run in the container that has /root/reference:   python tests/golden/make_golden.py
Every array written here was produced by the reference's own classes (see
oracle/ref_build/ref_harness.cpp for the exact calls); inputs are stored next to the
outputs so that the fixtures do not depend on a random-number stream.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def one_case(ref, tag, n, order, cell, dt, pcoord, npatch, pusher, mass, charge, nparts, seed):
    g = ol.make_grid(n, order, cell, dt, pcoord, npatch)
    rng = np.random.default_rng(seed)
    F = ol.random_fields(g, rng, scale=0.2)
    P = ol.random_particles(g, rng, nparts, p_scale=0.8, charge=charge)
    out = {"n": np.array(n), "order": order, "cell": np.array(cell), "dt": dt, "pcoord": np.array(pcoord),
           "npatch": np.array(npatch), "pusher": pusher, "mass": mass}
    for k, v in F.items():
        out["in_" + k] = v
    for k, v in P.items():
        out["in_" + k] = v.copy()
    # Species::dynamics order: gather, push, BC tag, deposit (Species.cpp:591,727,757,782)
    E, B, iold, delta = ref.interp(g, order, F, P["x"], P["y"], P["z"])
    invgf = ref.push(g, pusher, mass, P["x"], P["y"], P["z"], P["px"], P["py"], P["pz"], P["q"], E, B)
    tags = ref.bc_tag(g, P["x"], P["y"], P["z"])
    J = {k: F[k].copy() for k in ("Jx", "Jy", "Jz")}
    ref.project(g, order, J, P["x"], P["y"], P["z"], P["q"], P["w"], iold, delta)
    out.update(Epart=E, Bpart=B, iold=iold, deltaold=delta, invgf=invgf, tags=tags)
    for k in ("x", "y", "z", "px", "py", "pz"):
        out["out_" + k] = P[k]
    for k in J:
        out["out_" + k] = J[k]
    # keys of the pushed particles that stayed (SpeciesV::computeParticleCellKeys)
    keys = tags.copy()
    ncells = (n[0] + 1) * (n[1] + 1) * (n[2] + 1)
    count = np.zeros(ncells, dtype=np.int32)
    ref.cell_keys(g, P["x"], P["y"], P["z"], keys=keys, count=count)
    out.update(keys=keys, count=count)
    # VectorPatch::solveMaxwell order on the input fields with the deposited J
    X = {k: v.copy() for k, v in F.items()}
    X.update({k: J[k].copy() for k in J})
    ref.save_B(g, X)
    ref.maxwell_ampere(g, X)
    ref.maxwell_faraday(g, X)
    ref.center_B(g, X)
    for k in ("Ex", "Ey", "Ez", "Bx", "By", "Bz", "Bxm", "Bym", "Bzm"):
        out["mw_" + k] = X[k]
    out["norm2"] = np.array([ref.field_norm2(g, X[k], k) for k in ("Ex", "Ey", "Ez", "Bxm", "Bym", "Bzm")])
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print("wrote", tag)


def bc_case(ref, tag, seed):
    """Boundary conditions of the `next` rows: remove_particle_inf/sup and ElectroMagnBC3D_SM."""
    n, cell, dt = (6, 5, 7), (0.1, 0.12, 0.09), 0.05
    g = ol.make_grid(n, 2, cell, dt)
    rng = np.random.default_rng(seed)
    P = ol.random_particles(g, rng, 2000, p_scale=1.0)
    mn, mx = ol.patch_bounds(g)
    for i, c in enumerate("xyz"):
        P[c] = np.ascontiguousarray(mn[i] + (rng.random(2000) * 1.4 - 0.2) * (mx[i] - mn[i]))
    out = {"n": np.array(n), "cell": np.array(cell), "dt": dt}
    for k, v in P.items():
        out["in_" + k] = v
    sides = (1, 1, 0, 0, 1, 0)
    keys, q, lost = ref.bc_apply(g, sides, P)
    out.update(sides=np.array(sides), bc_keys=keys, bc_q=q, bc_lost=lost)
    F = ol.random_fields(g, rng, names=("Ex", "Ey", "Ez", "Bx", "By", "Bz"))
    for k, v in F.items():
        out["in_" + k] = v
    p = [n[i] + 2 * 2 + 1 for i in range(3)]
    for ib in range(6):
        a0 = ib // 2
        a1, a2 = (1 if a0 == 0 else 0), (1 if a0 == 2 else 2)
        kv = [0.2, -0.3, 0.25]
        kv[a0] = 1.0 if ib % 2 == 0 else -1.0
        db1 = np.ascontiguousarray(rng.standard_normal((p[a1], p[a2] + 1)))
        db2 = np.ascontiguousarray(rng.standard_normal((p[a1] + 1, p[a2])))
        X = {k: v.copy() for k, v in F.items()}
        ref.apply_SM(g, ib, kv, (1, 0, 0, 1), X, db1, db2)
        out["sm%d_k" % ib] = np.array(kv)
        out["sm%d_db1" % ib] = db1
        out["sm%d_db2" % ib] = db2
        for c in ("Bx", "By", "Bz"):
            out["sm%d_%s" % (ib, c)] = X[c]
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print("wrote", tag)


if __name__ == "__main__":
    ref = ol.Reference()
    bc_case(ref, "bc_remove_sm", 11)
    one_case(ref, "o2_boris_e", (6, 5, 4), 2, (0.07, 0.08, 0.09), 0.035, (0, 0, 0), (1, 1, 1), 0, 1.0, -1, 600, 1)
    one_case(ref, "o2_vay_p", (5, 6, 4), 2, (0.2, 0.3, 0.3), 0.1, (1, 0, 2), (3, 1, 4), 1, 1836.0, 1, 600, 2)
    one_case(ref, "o2_hc_e", (4, 4, 7), 2, (0.1, 0.1, 0.1), 0.05, (1, 1, 0), (2, 2, 1), 2, 1.0, -1, 600, 3)
    one_case(ref, "o4_boris_e", (6, 5, 5), 4, (0.07, 0.08, 0.09), 0.035, (0, 0, 0), (1, 1, 1), 0, 1.0, -1, 500, 4)
    one_case(ref, "o4_vay_e", (5, 5, 6), 4, (0.1, 0.1, 0.1), 0.05, (0, 1, 1), (2, 2, 2), 1, 1.0, -1, 500, 5)
