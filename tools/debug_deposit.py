"""Ad-hoc: GPU deposit against the oracle on a handful of particles (development aid, not a test)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
import smilei_b200

def run(N, p_scale, seed, onecell=False, n=(8, 8, 8)):
    cell, dt, order = (0.07, 0.07, 0.07), 0.038, 2
    g = ol.make_grid(n, order, cell, dt)
    orc = ol.Oracle()
    p = smilei_b200.Patch(n, cell, dt, interp_order=order, n_species=1)
    rng = np.random.default_rng(seed)
    F = ol.random_fields(g, rng, scale=0.0)
    for k, v in F.items():
        p.field_set(k, v)
    P = ol.random_particles(g, rng, N, p_scale=p_scale, charge=-1)
    if onecell:
        for i, k in enumerate("xyz"):
            P[k][:] = (3.0 + 0.9 * (rng.random(N) - 0.5)) * cell[i]
    p.species_config(0, 1.0, "boris", N)
    p.species_set(0, P["x"], P["y"], P["z"], P["px"], P["py"], P["pz"], P["w"], P["q"])
    p.sort(0)
    S = p.species_get(0)
    p.dynamics(0)
    out = p.species_get(0)
    E, B, iold, delta = orc.interp(g, order, F, S["x"], S["y"], S["z"])
    orc.push(g, 0, 1.0, S["x"], S["y"], S["z"], S["px"], S["py"], S["pz"], S["q"], E, B)
    J = {k: np.zeros_like(F[k]) for k in ("Jx", "Jy", "Jz")}
    orc.project(g, order, J, S["x"], S["y"], S["z"], S["q"], S["w"], iold, delta)
    nx = sum((np.round(S[k] / cell[i]) != np.round(out[k] / cell[i])).astype(int) for i, k in enumerate("xyz"))
    print(f"N={N} p={p_scale} onecell={onecell}: movers by #dims", np.bincount(nx, minlength=4), "flags", p.debug_flags())
    for k in ("Jx", "Jy", "Jz"):
        G = p.field_get(k)
        d = np.abs(G - J[k])
        m = np.max(np.abs(J[k]))
        idx = np.unravel_index(np.argmax(d), d.shape)
        print("  ", k, "rel err %.3e" % (d.max() / max(m, 1e-300)), "at", idx, "gpu", G[idx], "ref", J[k][idx],
              "sum gpu %.6e ref %.6e" % (G.sum(), J[k].sum()), "nonzero gpu/ref", (G != 0).sum(), (J[k] != 0).sum())
    p.close()

for args in [(1, 0.01, 1, True), (4, 0.01, 2, True), (40, 0.01, 3, True), (400, 0.01, 3, True), (3000, 0.01, 4, False), (3000, 0.2, 5, False), (3000, 1.5, 6, False)]:
    run(*args)
