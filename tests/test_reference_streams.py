"""the product's host-side particle creation (sb200_create_particles_ref, sb200_hilbert_index3d in
smilei_b200/csrc/creator.cu) against the reference's own ParticleCreator / Hilbert_functions / Random compiled from
/root/reference (oracle/_ref), bit for bit; and against a committed fixture of that build where _ref is absent.
Host-only entry points: no GPU needed."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from smilei_b200 import capi, particles_init
from smilei_b200.namelist import load_namelist

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "ref_streams.npz")
need_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
COLS = ("x", "y", "z", "px", "py", "pz", "w", "q")

THERMAL_SHORT = """
import math as m
Te = 100./511.
Ti = 10./511.
dx = 0.5*m.sqrt(Te)
dt = 0.5*dx/m.sqrt(3.)
Main(geometry="3Dcartesian", interpolation_order=2, timestep=dt, simulation_time=%(nsteps)d*dt,
     cell_length=[dx,dx,dx], grid_length=[%(ncell)d*dx]*3, number_of_patches=%(npatch)s,
     EM_boundary_conditions=[["periodic"]], print_every=100)
Species(name="proton", position_initialization="%(posinit)s", momentum_initialization="mj", particles_per_cell=8,
        c_part_max=1.0, mass=1836.0, charge=1.0, charge_density=1., mean_velocity=[0.,0.,0.], temperature=[Ti],
        pusher="boris", boundary_conditions=[["periodic","periodic"]]*3)
Species(name="electron", position_initialization="proton", momentum_initialization="mj", particles_per_cell=8,
        c_part_max=1.0, mass=1.0, charge=-1.0, charge_density=1., mean_velocity=[0.,0.,0.], temperature=[Te],
        pusher="boris", boundary_conditions=[["periodic","periodic"]]*3)
DiagScalar(every=10)
"""


def thermal_short(ncell=32, npatch=(4, 4, 4), posinit="random", nsteps=2001):
    """benchmarks/gpu/tst3d_v_o2_thermal_plasma_short.py of the reference (its diagnostics blocks left out)."""
    return load_namelist(THERMAL_SHORT % dict(ncell=ncell, npatch=list(npatch), posinit=posinit, nsteps=nsteps),
                         is_source=True)


TST3D_THERMAL = """
import math as m
T = 10./511.
n0 = 1.
dx = 0.5*m.sqrt(T)
dt = 0.95*dx/m.sqrt(3.)
Lx = %(ncell)d.*dx
Ly = Lx
Lz = Lx
def n0_(x,y,z):
    if (0.1*Lx<x<0.9*Lx) and (0.1*Ly<y<0.9*Ly) and (0.1*Lz<z<0.9*Lz):
        return n0
    else:
        return 0.
Main(geometry="3Dcartesian", interpolation_order=%(order)d, timestep=dt, simulation_time=2.*m.pi,
     cell_length=[dx,dx,dx], grid_length=[Lx,Ly,Lz], number_of_patches=[4,4,4],
     EM_boundary_conditions=[["periodic"]], print_every=1)
Species(name="proton", position_initialization="regular", momentum_initialization="mj", particles_per_cell=8,
        c_part_max=1.0, mass=1836.0, charge=1.0, charge_density=n0_, mean_velocity=[0.,0.,0.], temperature=[T],
        pusher="boris", boundary_conditions=[["periodic","periodic"]]*3)
Species(name="electron", position_initialization="regular", momentum_initialization="mj", particles_per_cell=8,
        c_part_max=1.0, mass=1.0, charge=-1.0, charge_density=n0_, mean_velocity=[0.,0.,0.], temperature=[T],
        pusher="boris", boundary_conditions=[["periodic","periodic"]]*3)
DiagScalar(every=1)
"""


def tst3d_thermal(order):
    """benchmarks/tst3d_01_thermal_plasma.py (order 2, 32^3 cells: BASELINE.json configs[0]) and
    benchmarks/tst3d_v_o4_thermal_plasma.py (order 4, 40^3 cells: configs[2] at the reference's own size): a plasma
    slab in the central 80 % of a periodic box, 8 ppc regular, 10 keV, 163 steps; diagnostics blocks left out."""
    return load_namelist(TST3D_THERMAL % dict(order=order, ncell={2: 32, 4: 40}[order]), is_source=True)


# ------------------------------------------------------------------------------------------ Hilbert numbering
SHAPES = [(2, 2, 2), (0, 0, 0), (1, 1, 1), (3, 3, 3), (3, 1, 2), (1, 3, 0), (0, 2, 3), (2, 0, 0), (4, 2, 2), (2, 3, 3)]


@pytest.mark.parametrize("m", SHAPES)
def test_hilbert_is_a_bijection(m):
    n = [1 << v for v in m]
    h = sorted(capi.hilbert_index3d(m, (x, y, z)) for x in range(n[0]) for y in range(n[1]) for z in range(n[2]))
    assert h == list(range(n[0] * n[1] * n[2]))


def test_hilbert_cube_neighbours():
    """On a cube the curve is continuous: consecutive patches are face neighbours."""
    m = (3, 3, 3)
    pos = {capi.hilbert_index3d(m, (x, y, z)): (x, y, z) for x in range(8) for y in range(8) for z in range(8)}
    for h in range(511):
        assert sum(abs(a - b) for a, b in zip(pos[h], pos[h + 1])) == 1


def test_hilbert_rejects_outside():
    with pytest.raises(capi.SmileiB200Error):
        capi.hilbert_index3d((2, 2, 2), (4, 0, 0))


@need_ref
@pytest.mark.parametrize("m", SHAPES)
def test_hilbert_vs_reference(m):
    from ref_creator import RefCreator
    ref = RefCreator()
    n = [1 << v for v in m]
    for x in range(n[0]):
        for y in range(n[1]):
            for z in range(n[2]):
                assert capi.hilbert_index3d(m, (x, y, z)) == ref.hilbert(m, (x, y, z)), (m, x, y, z)


# ------------------------------------------------------------------------------------------ particle creation
def _one_patch(creator, params, hpatch, P, posinit):
    """Both species of the patch at patch coordinates P through `creator` (product or reference)."""
    psize = [params.global_size[d] // params.number_of_patches[d] for d in range(3)]
    box_min = [P[d] * (psize[d] * params.cell_length[d]) for d in range(3)]
    state = (params.random_seed + hpatch) & 0xffffffff or 0xffffffff
    out = {}
    for sp in params.species:
        nppc, n_real, charge, T = particles_init._cell_profiles(params, sp, psize, box_min)
        src = out.get(sp.position_initialization)
        arrays, state = creator(state, None if src is not None else sp.position_initialization, "maxwell-juettner",
                                psize, box_min, params.cell_length, nppc, n_real, charge, T, sp.mass,
                                sp.regular_number or None,
                                positions=None if src is None else [src[k] for k in "xyz"])
        out[sp.name] = arrays
    return out, state


@need_ref
@pytest.mark.parametrize("posinit", ["random", "regular", "centered"])
@pytest.mark.parametrize("P", [(0, 0, 0), (3, 1, 2)])
def test_patch_bit_exact_vs_reference(posinit, P):
    """Protons (T/m = 1e-5: Maxwell-Boltzmann table) then electrons (T = 0.196: Maxwell-Juttner table with rejection)
    on one stream, every column and the final stream state equal to the reference's."""
    from ref_creator import RefCreator
    ref = RefCreator()
    params = thermal_short(posinit=posinit)
    h = capi.hilbert_index3d((2, 2, 2), P)
    a, sa = _one_patch(capi.create_particles_ref, params, h, P, posinit)
    b, sb = _one_patch(ref.patch, params, h, P, posinit)
    assert sa == sb
    for name in ("proton", "electron"):
        assert len(a[name]["x"]) == 8 ** 3 * 8
        for c in COLS:
            assert np.array_equal(a[name][c], b[name][c]), (name, c)
    assert np.array_equal(a["proton"]["x"], a["electron"]["x"])


@need_ref
def test_regular_number_and_hot_cold_branches():
    """regular_number = [4,2,1]; temperatures on both sides of the 0.1 switch, including the analytic tails."""
    from ref_creator import RefCreator
    ref = RefCreator()
    box, cell, box_min = (2, 3, 2), (0.1, 0.2, 0.3), (1.2, 0.0, 2.4)
    nppc = np.full(box, 8, dtype=np.int32)
    n_real = np.full(box, 0.006)
    n_real[0, 1, 1] = 0.                       # an empty cell draws nothing
    charge = np.full(box, -1.)
    for T, mass in ((0.05, 1.), (0.1, 1.), (3.0, 1.), (50., 1.), (1e-3, 1836.)):
        temperature = np.full(box, T)
        a, sa = capi.create_particles_ref(12345, "regular", "mj", box, box_min, cell, nppc, n_real, charge, temperature,
                                          mass, [4, 2, 1])
        b, sb = ref.patch(12345, "regular", "maxwell-juettner", box, box_min, cell, nppc, n_real, charge, temperature,
                          mass, [4, 2, 1])
        assert sa == sb and len(a["x"]) == 8 * 11
        for c in COLS:
            assert np.array_equal(a[c], b[c]), (T, c)


@need_ref
def test_regular_64_per_cell_follows_the_reference_truncation():
    """pow(64., 1./3.) is 3.9999999999999996 in IEEE double and the reference stores it in an int
    (ParticleCreator.cpp:651-652): 64 'regular' particles sit on a 3x3x3 lattice of spacing ~0.975/4 of the cell.
    The benchmark tst3d_v_o2_thermal_plasma_medium starts from exactly that; so does the product."""
    from ref_creator import RefCreator
    ref = RefCreator()
    box, cell, box_min = (2, 1, 2), (0.22, 0.22, 0.22), (0.0, 0.44, 1.1)
    nppc = np.full(box, 64, dtype=np.int32)
    n_real = np.full(box, 0.01)
    charge = np.full(box, 1.)
    temperature = np.full(box, 0.0196)
    a, sa = capi.create_particles_ref(7, "regular", "mj", box, box_min, cell, nppc, n_real, charge, temperature, 1836.)
    b, sb = ref.patch(7, "regular", "maxwell-juettner", box, box_min, cell, nppc, n_real, charge, temperature, 1836.)
    assert sa == sb
    for c in COLS:
        assert np.array_equal(a[c], b[c]), c
    assert len(np.unique(a["x"][:64])) == 3 and len(np.unique(a["z"][:64])) == 3


def test_streams_do_not_depend_on_the_rank_layout():
    """The particles of the box are the same multiset whether one rank or eight ranks create them."""
    params = thermal_short(ncell=16, npatch=(2, 2, 2))
    whole = particles_init.create_reference_streams(params, params.species, (16, 16, 16), (0, 0, 0))
    parts = [particles_init.create_reference_streams(params, params.species, (8, 8, 8), (a, b, c))
             for a in range(2) for b in range(2) for c in range(2)]
    for name in ("proton", "electron"):
        A = np.stack([whole[name][c].astype(np.float64) for c in COLS], axis=1)
        B = np.concatenate([np.stack([p[name][c].astype(np.float64) for c in COLS], axis=1) for p in parts])
        assert A.shape == B.shape == (16 ** 3 * 8, 8)
        assert np.array_equal(A[np.lexsort(A.T)], B[np.lexsort(B.T)])


def test_against_committed_fixture():
    """tests/golden/ref_streams.npz: patches (0,0,0) and (3,1,2) of the thermal-plasma-short namelist as the
    REFERENCE build created them (tests/golden/make_ref_streams_golden.py)."""
    g = np.load(GOLDEN)
    params = thermal_short()
    for tag, P in (("p000", (0, 0, 0)), ("p312", (3, 1, 2))):
        h = capi.hilbert_index3d((2, 2, 2), P)
        assert h == int(g[tag + "_hindex"])
        a, state = _one_patch(capi.create_particles_ref, params, h, P, "random")
        assert state == int(g[tag + "_state"])
        for name in ("proton", "electron"):
            for c in COLS:
                assert np.array_equal(a[name][c], g[f"{tag}_{name}_{c}"]), (tag, name, c)
    assert np.array_equal(np.asarray([capi.hilbert_index3d((2, 2, 2), (x, y, z)) for x in range(4) for y in range(4)
                                      for z in range(4)]), g["hilbert_4x4x4"])
    assert np.array_equal(np.asarray([capi.hilbert_index3d((3, 1, 2), (x, y, z)) for x in range(8) for y in range(2)
                                      for z in range(4)]), g["hilbert_8x2x4"])
