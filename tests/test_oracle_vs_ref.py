"""Pin the CPU oracle: bit-for-bit against the reference's OWN operator classes.

oracle/_ref/libsmilei_ref.so is built by oracle/ref_build/build_ref.sh from the sources
under /root/reference/src (Interpolator3D{2,4}Order, Pusher{Boris,Vay,HigueraCary},
Projector3D{2,4}Order, MA_Solver3D_norm, MF_Solver3D_Yee, ElectroMagn3D, SpeciesV,
BoundaryConditionType, Field3D, Particles).  Both sides are compiled -O2 -ffp-contract=off,
so the comparison is exact equality of every output double / int.

These tests run wherever the reference build exists (this container; the library also
travels to the GPU box).  tests/test_oracle_golden.py checks the same oracle against
committed outputs of that reference build, so the pin holds where _ref is absent.
"""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")

CASES = [  # (n, cell, dt, pcoord, npatch)
    ((8, 8, 8), (0.07, 0.07, 0.07), 0.038, (0, 0, 0), (1, 1, 1)),
    ((8, 9, 10), (0.07, 0.08, 0.09), 0.03, (1, 0, 2), (3, 1, 4)),
    ((6, 12, 7), (0.2, 3.0, 3.0), 0.19, (2, 1, 0), (4, 2, 1)),
]


@pytest.fixture(scope="module")
def libs():
    return ol.Oracle(), ol.Reference()


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_interp_bit_exact(libs, order, case):
    orc, ref = libs
    n, cell, dt, pc, npch = CASES[case]
    g = ol.make_grid(n, order, cell, dt, pc, npch)
    rng = np.random.default_rng(10 * case + order)
    F = ol.random_fields(g, rng)
    P = ol.random_particles(g, rng, 3000)
    a = orc.interp(g, order, F, P["x"], P["y"], P["z"])
    b = ref.interp(g, order, F, P["x"], P["y"], P["z"])
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


@pytest.mark.parametrize("pusher", [0, 1, 2])
@pytest.mark.parametrize("mass,charge", [(1.0, -1), (1836.0, 1)])
def test_push_bit_exact(libs, pusher, mass, charge):
    orc, ref = libs
    n, cell, dt, pc, npch = CASES[1]
    g = ol.make_grid(n, 2, cell, dt, pc, npch)
    rng = np.random.default_rng(100 + pusher)
    N = 4000
    P = ol.random_particles(g, rng, N, p_scale=2.0, charge=charge)
    E = rng.standard_normal(3 * N)
    B = rng.standard_normal(3 * N)
    Pa = {k: v.copy() for k, v in P.items()}
    Pb = {k: v.copy() for k, v in P.items()}
    ia = orc.push(g, pusher, mass, Pa["x"], Pa["y"], Pa["z"], Pa["px"], Pa["py"], Pa["pz"], Pa["q"], E, B)
    ib = ref.push(g, pusher, mass, Pb["x"], Pb["y"], Pb["z"], Pb["px"], Pb["py"], Pb["pz"], Pb["q"], E, B)
    assert np.array_equal(ia, ib)
    for k in ("x", "y", "z", "px", "py", "pz"):
        assert np.array_equal(Pa[k], Pb[k]), k
    assert not np.array_equal(Pa["px"], P["px"])


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_project_bit_exact(libs, order, case):
    """gather -> push -> deposit chain, J compared exactly (same serial summation order)."""
    orc, ref = libs
    n, cell, dt, pc, npch = CASES[case]
    g = ol.make_grid(n, order, cell, dt, pc, npch)
    rng = np.random.default_rng(1000 + 10 * case + order)
    F = ol.random_fields(g, rng, scale=0.1)
    N = 2500
    P = ol.random_particles(g, rng, N, p_scale=1.0)
    E, B, iold, delta = orc.interp(g, order, F, P["x"], P["y"], P["z"])
    orc.push(g, 0, 1.0, P["x"], P["y"], P["z"], P["px"], P["py"], P["pz"], P["q"], E, B)
    Ja = {k: F[k].copy() for k in ("Jx", "Jy", "Jz")}
    Jb = {k: F[k].copy() for k in ("Jx", "Jy", "Jz")}
    orc.project(g, order, Ja, P["x"], P["y"], P["z"], P["q"], P["w"], iold, delta)
    ref.project(g, order, Jb, P["x"], P["y"], P["z"], P["q"], P["w"], iold, delta)
    for k in Ja:
        assert np.array_equal(Ja[k], Jb[k]), k
        assert not np.array_equal(Ja[k], F[k])


def test_project_rho_bit_exact(libs):
    orc, ref = libs
    n, cell, dt, pc, npch = CASES[1]
    g = ol.make_grid(n, 2, cell, dt, pc, npch)
    rng = np.random.default_rng(77)
    F = ol.random_fields(g, rng, scale=0.1)
    N = 1500
    P = ol.random_particles(g, rng, N, p_scale=1.0)
    E, B, iold, delta = orc.interp(g, 2, F, P["x"], P["y"], P["z"])
    orc.push(g, 0, 1.0, P["x"], P["y"], P["z"], P["px"], P["py"], P["pz"], P["q"], E, B)
    Ja = {k: np.zeros(ol.field_dims(g, k)) for k in ("Jx", "Jy", "Jz", "rho")}
    Jb = {k: np.zeros(ol.field_dims(g, k)) for k in ("Jx", "Jy", "Jz", "rho")}
    orc.project_rho_o2(g, Ja, P["x"], P["y"], P["z"], P["q"], P["w"], iold, delta)
    ref.project_rho_o2(g, Jb, P["x"], P["y"], P["z"], P["q"], P["w"], iold, delta)
    for k in Ja:
        assert np.array_equal(Ja[k], Jb[k]), k


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_maxwell_bit_exact(libs, order, case):
    orc, ref = libs
    n, cell, dt, pc, npch = CASES[case]
    g = ol.make_grid(n, order, cell, dt, pc, npch)
    rng = np.random.default_rng(2000 + 10 * case + order)
    F = ol.random_fields(g, rng)
    Fa = {k: v.copy() for k, v in F.items()}
    Fb = {k: v.copy() for k, v in F.items()}
    for lib, X in ((orc, Fa), (ref, Fb)):
        lib.save_B(g, X)
        lib.maxwell_ampere(g, X)
        lib.maxwell_faraday(g, X)
        lib.center_B(g, X)
    for k in F:
        assert np.array_equal(Fa[k], Fb[k]), k
    for k in ("Ex", "Ey", "Ez", "Bx", "By", "Bz", "Bxm", "Bym", "Bzm"):
        assert not np.array_equal(Fa[k], F[k]), k


@pytest.mark.parametrize("case", range(len(CASES)))
def test_keys_and_bc_bit_exact(libs, case):
    orc, ref = libs
    n, cell, dt, pc, npch = CASES[case]
    g = ol.make_grid(n, 2, cell, dt, pc, npch)
    rng = np.random.default_rng(3000 + case)
    N = 20000
    P = ol.random_particles(g, rng, N)
    mn, mx = ol.patch_bounds(g)
    # put particles exactly on node / half-cell boundaries, where round() decides
    for i, c in enumerate("xyz"):
        P[c][:200] = mn[i] + rng.integers(0, n[i], 200) * g.cell[i]
        P[c][200:400] = mn[i] + (rng.integers(0, n[i], 200) + 0.5) * g.cell[i]
    ncells = (n[0] + 1) * (n[1] + 1) * (n[2] + 1)
    ca = np.zeros(ncells, dtype=np.int32)
    cb = np.zeros(ncells, dtype=np.int32)
    ka = orc.cell_keys(g, P["x"], P["y"], P["z"], count=ca)
    kb = ref.cell_keys(g, P["x"], P["y"], P["z"], count=cb)
    assert np.array_equal(ka, kb) and np.array_equal(ca, cb)
    assert ka.min() >= 0 and ka.max() < ncells and ca.sum() == N
    # leavers on every face
    for i, c in enumerate("xyz"):
        P[c][1000 + 100 * i:1050 + 100 * i] = mn[i] - 0.3 * g.cell[i]
        P[c][1050 + 100 * i:1100 + 100 * i] = mx[i] + 0.3 * g.cell[i]
    P["x"][5000] = mx[0]  # exactly on the upper bound leaves (>=)
    P["y"][5001] = mn[1]  # exactly on the lower bound stays (<)
    ta = orc.bc_tag(g, P["x"], P["y"], P["z"])
    tb = ref.bc_tag(g, P["x"], P["y"], P["z"])
    assert np.array_equal(ta, tb)
    assert set(np.unique(ta)) == {0, -2, -3, -4, -5, -6, -7}
    assert ta[5000] == -3 and ta[5001] == 0


@pytest.mark.parametrize("pcoord,npatch", [((0, 0, 0), (1, 1, 1)), ((0, 1, 2), (2, 3, 3)), ((1, 2, 1), (3, 3, 3))])
def test_norm2_bit_exact(libs, pcoord, npatch):
    orc, ref = libs
    g = ol.make_grid((8, 6, 7), 2, (0.1, 0.1, 0.1), 0.05, pcoord, npatch)
    rng = np.random.default_rng(5)
    for name in ("Ex", "Ey", "Ez", "Bx", "By", "Bz", "rho"):
        f = np.ascontiguousarray(rng.standard_normal(ol.field_dims(g, name)))
        assert orc.field_norm2(g, f, name) == ref.field_norm2(g, f, name)


@pytest.mark.parametrize("sides", [(1, 0, 0, 0, 0, 0), (1, 1, 0, 0, 1, 0), (1, 1, 1, 1, 1, 1), (0, 0, 0, 1, 0, 0)])
def test_remove_particle_bc(libs, sides):
    """PartBoundCond::apply with remove_particle_inf/sup on some global sides, internal_inf/sup on the others:
    keys, zeroed charges and the lost energy, bit for bit."""
    orc, ref = libs
    n, cell, dt = (6, 5, 7), (0.1, 0.12, 0.09), 0.05
    g = ol.make_grid(n, 2, cell, dt)
    rng = np.random.default_rng(77)
    P = ol.random_particles(g, rng, 5000, p_scale=1.0)
    mn, mx = ol.patch_bounds(g)
    for i, c in enumerate("xyz"):       # spread beyond the patch on every side, corners included
        P[c] = np.ascontiguousarray(mn[i] + (rng.random(5000) * 1.4 - 0.2) * (mx[i] - mn[i]))
    a = orc.bc_apply(g, sides, P)
    b = ref.bc_apply(g, sides, P)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[2] == b[2]
    assert (a[0] == -1).sum() > 0 and (a[1] == 0).sum() == (a[0] == -1).sum()


@pytest.mark.parametrize("i_boundary", range(6))
@pytest.mark.parametrize("laser", [False, True])
def test_silver_muller(libs, i_boundary, laser):
    """ElectroMagnBC3D_SM constructor + apply on every face, oblique incidence vector, with and without injected
    amplitudes, transverse sides mixed (one with a neighbour, one without): exact equality of Bx, By, Bz."""
    orc, ref = libs
    n, cell, dt = (6, 5, 7), (0.1, 0.12, 0.09), 0.05
    g = ol.make_grid(n, 2, cell, dt)
    rng = np.random.default_rng(5 + i_boundary)
    F = ol.random_fields(g, rng, names=("Ex", "Ey", "Ez", "Bx", "By", "Bz"))
    axis0 = i_boundary // 2
    axis1, axis2 = (1 if axis0 == 0 else 0), (1 if axis0 == 2 else 2)
    p = [g.n[i] + 2 * g.o[i] + 1 for i in range(3)]
    k = [0.2, -0.3, 0.25]
    k[axis0] = 1.0 if i_boundary % 2 == 0 else -1.0
    db1 = np.ascontiguousarray(rng.standard_normal((p[axis1], p[axis2] + 1))) if laser else None
    db2 = np.ascontiguousarray(rng.standard_normal((p[axis1] + 1, p[axis2]))) if laser else None
    isb = (1, 0, 0, 1)
    A = {k_: v.copy() for k_, v in F.items()}
    B = {k_: v.copy() for k_, v in F.items()}
    orc.apply_SM(g, i_boundary, k, isb, A, db1, db2)
    ref.apply_SM(g, i_boundary, k, isb, B, db1, db2)
    changed = 0
    for name in ("Bx", "By", "Bz"):
        assert np.array_equal(A[name], B[name]), name
        changed += int((A[name] != F[name]).sum())
    assert changed > 0


@pytest.mark.parametrize("order", [2, 4])
def test_moved_window_origin(libs, order):
    """A patch the moving window has advanced (Patch::initStep3 with n_moved, Patch.cpp:159-163): gather, boundary
    tags, cell keys and deposit all use the moved origin — exact equality with the reference classes."""
    orc, ref = libs
    n, cell, dt = (10, 9, 11), (0.2, 0.3, 0.25), 0.1
    g = ol.make_grid(n, order, cell, dt, (0, 1, 0), (1, 2, 1), n_moved=24)
    rng = np.random.default_rng(60 + order)
    F = ol.random_fields(g, rng)
    P = ol.random_particles(g, rng, 3000, p_scale=0.5)
    mn, mx = ol.patch_bounds(g)
    assert mn[0] == 24 * cell[0] and np.all(P["x"] >= mn[0])
    a = orc.interp(g, order, F, P["x"], P["y"], P["z"])
    b = ref.interp(g, order, F, P["x"], P["y"], P["z"])
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    E, B, iold, delta = a
    Pa = {k: v.copy() for k, v in P.items()}
    Pb = {k: v.copy() for k, v in P.items()}
    orc.push(g, 0, 1.0, Pa["x"], Pa["y"], Pa["z"], Pa["px"], Pa["py"], Pa["pz"], Pa["q"], E, B)
    ref.push(g, 0, 1.0, Pb["x"], Pb["y"], Pb["z"], Pb["px"], Pb["py"], Pb["pz"], Pb["q"], E, B)
    ta = orc.bc_tag(g, Pa["x"], Pa["y"], Pa["z"])
    tb = ref.bc_tag(g, Pb["x"], Pb["y"], Pb["z"])
    assert np.array_equal(ta, tb) and (ta < 0).sum() > 0
    ka, kb = ta.copy(), tb.copy()
    orc.cell_keys(g, Pa["x"], Pa["y"], Pa["z"], keys=ka)
    ref.cell_keys(g, Pb["x"], Pb["y"], Pb["z"], keys=kb)
    assert np.array_equal(ka, kb)
    Ja = {k: F[k].copy() for k in ("Jx", "Jy", "Jz")}
    Jb = {k: F[k].copy() for k in ("Jx", "Jy", "Jz")}
    orc.project(g, order, Ja, Pa["x"], Pa["y"], Pa["z"], Pa["q"], Pa["w"], iold, delta)
    ref.project(g, order, Jb, Pb["x"], Pb["y"], Pb["z"], Pb["q"], Pb["w"], iold, delta)
    for k in Ja:
        assert np.array_equal(Ja[k], Jb[k]), k


def _sort_case(g, rng, N, n_leave, n_arr):
    """Particles after a step: most still in the patch, n_leave tagged for exchange (negative key) anywhere in the
    list, and arrivals from the six neighbours (inside the patch, as the reference's exchange delivers them)."""
    P = ol.random_particles(g, rng, N)
    tags = np.zeros(N, dtype=np.int32)
    lv = rng.choice(N, n_leave, replace=False)
    tags[lv] = -1 - rng.integers(0, 7, n_leave)
    arr = []
    for k in range(6):
        m = int(n_arr[k])
        a = ol.random_particles(g, rng, m) if m else None
        arr.append(a)
    return P, tags, arr


def _cells_as_multisets(part, first):
    """Per cell: the sorted tuple list of its particles (all columns) - order inside a cell is not specified."""
    cols = np.stack([part[k] for k in ("x", "y", "z", "px", "py", "pz", "w")] + [part["q"].astype(float)], axis=1)
    out = []
    for c in range(len(first) - 1):
        blk = cols[first[c]:first[c + 1]]
        out.append(blk[np.lexsort(blk.T[::-1])])
    return out


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("n_leave,n_arr", [(0, (0,) * 6), (700, (0,) * 6), (300, (90, 110, 0, 40, 70, 60)),
                                           (50, (200, 150, 100, 120, 90, 80))])
def test_sort_matches_reference_cycle_sort(libs, case, n_leave, n_arr):
    """the reference's in-place cycle sort (SpeciesV::sortParticles, SpeciesV.cpp:599-762) called for
    real — leavers erased, arrivals filling their holes (:669-697), fewer and more arrivals than leavers — against
    the canonical order of this build (stable counting sort of residents followed by arrivals): identical
    first_index, identical particle count, and per cell the same multiset of particles (the order inside a cell is
    algorithm-dependent in the reference and not a specification; its cell_keys array is not moved with the
    particles, Particles.cpp:813-830, so it is not an output)."""
    orc, ref = libs
    n, cell, dt, pc, npch = CASES[case]
    g = ol.make_grid(n, 2, cell, dt, pc, npch)
    rng = np.random.default_rng(500 + 10 * case + n_leave)
    P, tags, arr = _sort_case(g, rng, 20000, n_leave, n_arr)
    R, rfirst = ref.sort(g, P, tags, arr)
    # this build's canonical sort: residents then arrivals (x then y then z, - then +), stable by key
    cols = ("x", "y", "z", "px", "py", "pz", "w", "q")
    allp = {k: np.concatenate([P[k]] + [a[k] for a in arr if a is not None]) for k in cols}
    keys = np.concatenate([tags, np.zeros(sum(n_arr), dtype=np.int32)])
    orc.cell_keys(g, np.ascontiguousarray(allp["x"]), np.ascontiguousarray(allp["y"]), np.ascontiguousarray(allp["z"]),
                  keys=keys)
    ncells = (n[0] + 1) * (n[1] + 1) * (n[2] + 1)
    first, perm = orc.counting_sort_perm(keys, ncells)
    mine = {k: allp[k][perm] for k in cols}
    assert len(R["x"]) == len(perm) == 20000 - n_leave + sum(n_arr)
    assert np.array_equal(rfirst, first)
    for a, b in zip(_cells_as_multisets(R, rfirst), _cells_as_multisets(mine, first)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("species_arrays", [False, True])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_diag_step_deposit_bit_exact(libs, order, species_arrays, case):
    """SURVEY f-3: Projector3D{2,4}Order::currentsAndDensityWrapper with diag_flag = true — into the totals, or into
    the species' own Jx_s/Jy_s/Jz_s/rho_s (Projector3D2Order.cpp:756-763) followed by ElectroMagn3D::computeTotalRhoJ
    (ElectroMagn3D.cpp:1753-1799) — run on the reference's classes, against the oracle's orc_project_rho /
    orc_compute_total_rhoJ: every double equal."""
    orc, ref = libs
    n, cell, dt, pc, npch = CASES[case]
    if order == 4:
        n = tuple(max(v, 10) for v in n)
    g = ol.make_grid(n, order, cell, dt, pc, npch)
    rng = np.random.default_rng(700 + 10 * case + order)
    F = ol.random_fields(g, rng)
    P = ol.random_particles(g, rng, 4000, p_scale=1.0)
    E, B, iold, delta = orc.interp(g, order, F, P["x"], P["y"], P["z"])
    orc.push(g, 0, 1.0, P["x"], P["y"], P["z"], P["px"], P["py"], P["pz"], P["q"], E, B)
    names = ("Jx", "Jy", "Jz", "rho")
    dims = {k: ol.field_dims(g, k) for k in names}
    tot0 = {k: np.ascontiguousarray(rng.standard_normal(dims[k])) for k in names}     # what other species left there
    Ja, Jb = {k: v.copy() for k, v in tot0.items()}, {k: v.copy() for k, v in tot0.items()}
    if species_arrays:
        Sa = {k: np.zeros(dims[k]) for k in names}
        Sb = {k: np.zeros(dims[k]) for k in names}
        orc.project_rho(g, order, Sa, P["x"], P["y"], P["z"], P["q"], P["w"], iold, delta)
        orc.compute_total_rhoJ(g, Ja, Sa)
        ref.project_rho_species(g, order, Jb, Sb, P["x"], P["y"], P["z"], P["q"], P["w"], iold, delta)
        for k in names:
            assert np.array_equal(Sa[k], Sb[k]), k
            assert np.abs(Sb[k]).max() > 0
    else:
        orc.project_rho(g, order, Ja, P["x"], P["y"], P["z"], P["q"], P["w"], iold, delta)
        ref.project_rho_species(g, order, Jb, None, P["x"], P["y"], P["z"], P["q"], P["w"], iold, delta)
    for k in names:
        assert np.array_equal(Ja[k], Jb[k]), k


def test_vectorised_reference_path_matches_scalar_operators(libs):
    """The timed CPU baseline drives the reference's VECTORISED species path (Interpolator3D2OrderV, pusher,
    computeParticleCellKeys, Projector3D2OrderV per cell, then sortParticles with the leavers coming back through the
    receive buffer — oracle/ref_build/ref_harness.cpp::ref_time_dynamics_V).  Its currents after one step must be those
    of the scalar operator classes on the same particles (1e-13: the V projector sums in another order), and a
    multi-step run with the sort must keep going (the particle count is conserved: every leaver returns)."""
    orc, ref = libs
    n, cell, dt = (8, 8, 8), (0.07, 0.07, 0.07), 0.038
    g = ol.make_grid(n, 2, cell, dt)
    rng = np.random.default_rng(77)
    F = ol.random_fields(g, rng, scale=1e-2)
    N = 8 ** 3 * 12
    P = ol.random_particles(g, rng, N, p_scale=0.3)
    keys = orc.cell_keys(g, P["x"], P["y"], P["z"])
    first, perm = orc.counting_sort_perm(keys, 9 ** 3)
    S = {k: np.ascontiguousarray(v[perm]) for k, v in P.items()}
    names = ("Jx", "Jy", "Jz")
    dims = [ol.field_dims(g, k) for k in names]
    Jv = np.zeros(sum(int(np.prod(d)) for d in dims))
    ref.time_dynamics_V(g, 0, 1.0, F, S, first, 1, 1, 1, with_sort=True, J_out=Jv)
    R = {k: v.copy() for k, v in S.items()}
    E, B, iold, delta = ref.interp(g, 2, F, R["x"], R["y"], R["z"])
    ref.push(g, 0, 1.0, R["x"], R["y"], R["z"], R["px"], R["py"], R["pz"], R["q"], E, B)
    J = {k: np.zeros(ol.field_dims(g, k)) for k in names}
    ref.project(g, 2, J, R["x"], R["y"], R["z"], R["q"], R["w"], iold, delta)
    off = 0
    for k, d in zip(names, dims):
        m = int(np.prod(d))
        a = Jv[off:off + m].reshape(d)
        off += m
        assert np.abs(a - J[k]).max() <= 1e-13 * np.abs(J[k]).max(), k
    t, _ = ref.time_dynamics_V(g, 0, 1.0, F, S, first, 2, 5, 2, with_sort=True)
    assert t > 0
