"""CPU tests of the host logic: namelist layer, operator factories, time-loop ordering and the
exchange layer under torch.distributed/gloo with 2 processes.  The numerical back end is the
oracle-backed OraclePatch (tests/oracle_patch.py) — the same Patch interface the CUDA library
implements — so what is tested here is the PRODUCT's sequencing and message plan."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
from oracle_patch import OraclePatch
from smilei_b200 import namelist, operators
from smilei_b200.capi import SmileiB200Error
from smilei_b200.simulation import Simulation

NAMELIST = """
import math
T = 10./511.
dx = 0.5*math.sqrt(T)
Main(geometry="3Dcartesian", interpolation_order={order}, timestep=0.95*dx/math.sqrt(3.),
     simulation_time=20*0.95*dx/math.sqrt(3.), cell_length=[dx,dx,dx], grid_length=[{nx}*dx,{ny}*dx,{nz}*dx],
     number_of_patches=[1,1,1], EM_boundary_conditions=[["periodic"]], print_every=1)
Species(name="proton", position_initialization="regular", momentum_initialization="mj", particles_per_cell=8,
        mass=1836.0, charge=1.0, charge_density=1., temperature=[T], pusher="{pusher}",
        boundary_conditions=[["periodic","periodic"]]*3)
Species(name="electron", position_initialization="proton", momentum_initialization="mj", particles_per_cell=8,
        mass=1.0, charge=-1.0, charge_density=1., temperature=[T], pusher="{pusher}",
        boundary_conditions=[["periodic","periodic"]]*3)
DiagScalar(every=1)
LoadBalancing(every=20)
"""


def make_params(order=2, n=(12, 12, 12), pusher="boris"):
    return namelist.load_namelist(NAMELIST.format(order=order, nx=n[0], ny=n[1], nz=n[2], pusher=pusher), is_source=True)


def test_namelist_derived_quantities():
    p = make_params()
    T = 10. / 511.
    dx = 0.5 * T ** 0.5
    assert p.global_size == [12, 12, 12] and p.oversize == [2, 2, 2]
    assert p.timestep == 0.95 * dx / 3 ** 0.5
    assert p.n_time == int((20 * 0.95 * dx / 3 ** 0.5) / p.timestep)
    assert [s.name for s in p.species] == ["proton", "electron"]
    assert p.species[0].mass == 1836. and p.species[1].pusher == "boris"
    assert p.cell_volume == 1.0 * dx * dx * dx
    assert make_params(order=4).oversize == [4, 4, 4]
    p.check_hot_path()


@pytest.mark.skipif(not os.path.isdir("/root/reference/benchmarks"), reason="reference namelists not present")
@pytest.mark.parametrize("path,order,nspec,n_time,size", [
    ("benchmarks/tst3d_01_thermal_plasma.py", 2, 2, 163, 32),
    ("benchmarks/gpu/tst3d_v_o2_thermal_plasma_short.py", 2, 2, 2001, 32),
    ("benchmarks/tst3d_v_o4_thermal_plasma.py", 4, 2, None, 40),
])
def test_reference_namelists_load_unmodified(path, order, nspec, n_time, size):
    p = namelist.load_namelist(os.path.join("/root/reference", path))
    assert p.interpolation_order == order and len(p.species) == nspec
    assert p.global_size == [size] * 3 and p.number_of_patches == [4, 4, 4]
    if n_time is not None:
        assert p.n_time == n_time          # 163 steps for tst3d_01
    p.check_hot_path()


@pytest.mark.skipif(not os.path.isdir("/root/reference/benchmarks"), reason="reference namelists not present")
def test_laser_wake_namelist_loads_unmodified():
    """BASELINE.json configs[3]: Vay pusher, lasers through Silver-Mueller sides with oblique absorption vectors,
    `remove` particles, moving window — all on the path."""
    p = namelist.load_namelist("/root/reference/benchmarks/tst3d_s_o2_laser_wake_yee_vay.py")
    p.check_hot_path()
    assert p.global_size == [512, 40, 40] and p.number_of_patches == [64, 4, 4] and p.n_time == 1077
    assert p.EM_BCs == [["silver-muller", "silver-muller"]] * 3 and p.EM_BCs_k[3] == [1., -0.005, 0.]
    assert p.species[0].pusher == "vay" and p.species[0].boundary_conditions == [["remove", "remove"]] * 3
    assert p.has_window and float(p.window.time_start) == 102.4 and p.window.velocity_x == 0.9997
    assert len(p.laser_blocks) == 1 and p.laser_blocks[0].space_time_profile is not None


@pytest.mark.skipif(not os.path.isdir("/root/reference/benchmarks"), reason="reference namelists not present")
@pytest.mark.parametrize("path", ["benchmarks/tst3d_08_envelope_wake.py", "benchmarks/tst3d_22_em_dispersion_m4.py",
                                  "benchmarks/tst3d_13_particle_injection_x.py"])
def test_out_of_scope_namelist_is_rejected_loudly(path):
    """Envelope model / PML, another Maxwell solver, particle injectors: rejected, never silently ignored."""
    with pytest.raises(namelist.NamelistError):
        namelist.load_namelist(os.path.join("/root/reference", path)).check_hot_path()


def test_factories_follow_the_reference_surface():
    p = make_params(order=4, pusher="vay")
    sim = Simulation(p, patch_factory=OraclePatch)
    sp = sim.vecSpecies[0]
    assert isinstance(sp.Interp, operators.Interpolator3D4Order)
    assert isinstance(sp.Push, operators.PusherVay)
    assert isinstance(sp.Proj, operators.Projector3D4Order)
    assert isinstance(sim.EMfields.MaxwellAmpereSolver_, operators.MA_Solver3D_norm)
    assert isinstance(sim.EMfields.MaxwellFaradaySolver_, operators.MF_Solver3D_Yee)
    # call-order contract of the fused operators
    with pytest.raises(SmileiB200Error):
        sp.Push(sp.particles, sim.smpi, 0, 0, 0)
    with pytest.raises(SmileiB200Error):
        sim.EMfields.MaxwellFaradaySolver_(sim.EMfields)
    p.species[0].pusher = "borisnr"
    with pytest.raises(SmileiB200Error):
        operators.PusherFactory.create(p, p.species[0])


def test_single_rank_loop_conserves_particles_and_energy():
    p = make_params(n=(12, 12, 12))
    sim = Simulation(p, patch_factory=OraclePatch)
    sim.create_particles(seed=1)
    n0 = sim.n_particles()
    assert n0 == [12 ** 3 * 8] * 2
    uk0, ue0 = sim.scalars()
    assert ue0 == 0.0                      # electrons sit on the protons: rho = 0, no field at t = 0
    hist = sim.run(10, scalars_every=1)
    assert sim.n_particles() == n0
    tot0 = uk0.sum() + ue0
    tot = np.array([h[1].sum() + h[2] for h in hist])
    assert np.all(np.abs(tot / tot0 - 1) < 2e-3)      # the reference's own Utot tolerance (validate_tst3d_v_o2_*: 1e-3 on ratios)
    assert hist[-1][2] > 0


def _global_state(n, seed, nper=6):
    """Explicit arrays (the reference's numpy position/momentum initialisation) for the whole box."""
    T = 10. / 511.
    dx = 0.5 * T ** 0.5
    rng = np.random.default_rng(seed)
    N = n[0] * n[1] * n[2] * nper
    pos = [rng.random(N) * n[d] * dx for d in range(3)]
    out = {}
    for name, mass, q in (("proton", 1836., 1), ("electron", 1., -1)):
        s = (T / mass) ** 0.5 * 3.0      # hot: plenty of particles cross patch faces, edges and corners
        out[name] = dict(x=pos[0].copy(), y=pos[1].copy(), z=pos[2].copy(), px=s * rng.standard_normal(N),
                         py=s * rng.standard_normal(N), pz=s * rng.standard_normal(N), w=np.full(N, dx ** 3 / nper),
                         q=np.full(N, q, dtype=np.int16))
    return out


def _run_rank(rank, world, rank_grid, n, steps, port, ret):
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p = make_params(n=n)
    sim = Simulation(p, rank_grid=rank_grid, rank=rank, patch_factory=OraclePatch)
    state = _global_state(n, 5)
    mn, mx = sim.patch.mn, sim.patch.mx
    for sp in sim.vecSpecies:
        a = state[sp.name]
        inside = np.ones(len(a["x"]), bool)
        for d, c in enumerate("xyz"):
            inside &= (a[c] >= mn[d]) & (a[c] < mx[d])
        sim.set_particles(sp.ispec, **{k: v[inside] for k, v in a.items()})
    hist = sim.run(steps, scalars_every=1)
    parts = [sim.patch.species_get(s.ispec) for s in sim.vecSpecies]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, parts)
        sent = sim.exchanger.bytes_sent
        dist.barrier()
        dist.destroy_process_group()
    else:
        gathered, sent = [parts], 0
    if rank == 0:
        ret["hist"] = [(h[0], h[1].tolist(), h[2]) for h in hist]
        ret["parts"] = gathered
        ret["sent"] = sent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _launch(rank_grid, n, steps):
    world = int(np.prod(rank_grid))
    if world == 1:
        ret = {}
        _run_rank(0, 1, rank_grid, n, steps, 0, ret)
        return ret
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run_rank, args=(world, rank_grid, n, steps, _free_port(), ret), nprocs=world, join=True)
    return dict(ret)


def _canonical(parts_per_rank, ispec):
    cols = {k: np.concatenate([r[ispec][k] for r in parts_per_rank]) for k in ("x", "y", "z", "px", "py", "pz", "w")}
    order = np.lexsort((cols["pz"], cols["py"], cols["px"]))     # momenta are unique identifiers
    return {k: v[order] for k, v in cols.items()}


@pytest.fixture(scope="module")
def single_rank_reference():
    return _launch((1, 1, 1), (12, 12, 12), 6)


@pytest.mark.parametrize("rank_grid", [(2, 1, 1), (1, 2, 1), (1, 1, 2)])
def test_two_ranks_gloo_match_single_rank(single_rank_reference, rank_grid):
    """world_size-2 run of the product's Simulation + Exchanger == the 1-rank run: same particles
    (every one of them, wherever it migrated), same energies, up to the rounding of the different J
    summation order at the shared planes."""
    ref = single_rank_reference
    two = _launch(rank_grid, (12, 12, 12), 6)
    assert two["sent"] > 0
    for (it_a, uk_a, ue_a), (it_b, uk_b, ue_b) in zip(ref["hist"], two["hist"]):
        assert it_a == it_b
        assert np.allclose(uk_a, uk_b, rtol=1e-11, atol=0)
        assert abs(ue_a - ue_b) <= 1e-9 * abs(ue_a)
    for ispec in range(2):
        a = _canonical(ref["parts"], ispec)
        b = _canonical(two["parts"], ispec)
        assert len(a["x"]) == len(b["x"]) == 12 ** 3 * 6
        for k in a:
            assert np.allclose(a[k], b[k], rtol=0, atol=1e-11), k


def test_three_ranks_gloo_distinct_neighbours():
    """3 ranks along x: the -x and +x neighbours are different peers (separate messages, per-direction sizes)."""
    n = (18, 12, 12)
    ref = _launch((1, 1, 1), n, 4)
    three = _launch((3, 1, 1), n, 4)
    for (it_a, uk_a, ue_a), (it_b, uk_b, ue_b) in zip(ref["hist"], three["hist"]):
        assert np.allclose(uk_a, uk_b, rtol=1e-11, atol=0)
        assert abs(ue_a - ue_b) <= 1e-9 * abs(ue_a)
    for ispec in range(2):
        a = _canonical(ref["parts"], ispec)
        b = _canonical(three["parts"], ispec)
        assert len(a["x"]) == len(b["x"]) == 18 * 12 * 12 * 6
        for k in a:
            assert np.allclose(a[k], b[k], rtol=0, atol=1e-11), k


# ---------------------------------------------------------------------------------------------------------
# open box: Silver-Mueller sides, a laser, `remove` particles — ranks at a box side have no neighbour there
# ---------------------------------------------------------------------------------------------------------
OPEN_NAMELIST = """
Main(geometry="3Dcartesian", interpolation_order=2, timestep=0.04, number_of_timesteps=30,
     cell_length=[0.1, 0.25, 0.25], number_of_cells=[24, 12, 12], number_of_patches=[1, 1, 1],
     EM_boundary_conditions=[["silver-muller"], ["silver-muller"], ["periodic"]])
LaserGaussian3D(box_side="xmin", a0=1.5, omega=1.0, focus=[1.2, 1.5, 1.5], waist=0.9,
                time_envelope=tgaussian(start=0., duration=1.6, fwhm=0.6, center=0.8))
Species(name="electron", position_initialization="regular", regular_number=[1, 1, 1], momentum_initialization="cold",
        particles_per_cell=1, mass=1.0, charge=-1.0, number_density=0.02, pusher="boris",
        boundary_conditions=[["remove"], ["remove"], ["periodic"]])
"""


def _run_rank_open(rank, world, rank_grid, steps, port, ret):
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p = namelist.load_namelist(OPEN_NAMELIST, is_source=True)
    sim = Simulation(p, rank_grid=rank_grid, rank=rank, patch_factory=OraclePatch)
    rng = np.random.default_rng(3)
    N = 3000
    L = [p.global_size[d] * p.cell_length[d] for d in range(3)]
    a = {c: rng.random(N) * L[i] * (1 - 1e-12) for i, c in enumerate("xyz")}
    for c in ("px", "py", "pz"):
        a[c] = 0.4 * rng.standard_normal(N)
    a["w"] = np.full(N, 1e-3)
    a["q"] = np.full(N, -1, dtype=np.int16)
    mn, mx = sim.patch.mn, sim.patch.mx
    inside = np.ones(N, bool)
    for d, c in enumerate("xyz"):
        inside &= (a[c] >= mn[d]) & (a[c] < mx[d])
    sim.set_particles(0, **{k: np.ascontiguousarray(v[inside]) for k, v in a.items()})
    hist = sim.run(steps, scalars_every=1)
    parts = [sim.patch.species_get(0)]
    lost = sim.patch.species_lost_energy(0)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (parts, lost))
        dist.barrier()
        dist.destroy_process_group()
    else:
        gathered = [(parts, lost)]
    if rank == 0:
        ret["hist"] = [(h[0], h[1].tolist(), h[2]) for h in hist]
        ret["parts"] = [g[0] for g in gathered]
        ret["lost"] = sum(g[1] for g in gathered)


def _launch_open(rank_grid, steps):
    world = int(np.prod(rank_grid))
    if world == 1:
        ret = {}
        _run_rank_open(0, 1, rank_grid, steps, 0, ret)
        return ret
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run_rank_open, args=(world, rank_grid, steps, _free_port(), ret), nprocs=world, join=True)
    return dict(ret)


@pytest.mark.parametrize("rank_grid", [(2, 1, 1), (1, 2, 1)])
def test_open_box_two_ranks_gloo_match_single_rank(rank_grid):
    """Laser through the xmin Silver-Mueller side, open in x and y: split across two ranks along an OPEN
    dimension (each rank has a neighbour on one side only) the run reproduces the single-rank one — fields
    energy, kinetic energy, the surviving particles and the energy carried away by the removed ones."""
    one = _launch_open((1, 1, 1), 30)
    two = _launch_open(rank_grid, 30)
    assert one["lost"] > 0 and abs(one["lost"] - two["lost"]) <= 1e-9 * one["lost"]
    for (it_a, uk_a, ue_a), (it_b, uk_b, ue_b) in zip(one["hist"], two["hist"]):
        assert it_a == it_b
        assert np.allclose(uk_a, uk_b, rtol=1e-10, atol=0)
        assert abs(ue_a - ue_b) <= 1e-9 * abs(ue_a)
    a = _canonical(one["parts"], 0)
    b = _canonical(two["parts"], 0)
    assert 0 < len(a["x"]) == len(b["x"]) < 3000
    for k in a:
        assert np.allclose(a[k], b[k], rtol=0, atol=1e-10), k


# ---------------------------------------------------------------------------------------------------------
# moving window with the box split across ranks along y (one rank along x spans the window direction)
# ---------------------------------------------------------------------------------------------------------
WINDOW_NAMELIST = """
Main(geometry="3Dcartesian", interpolation_order=2, timestep=0.09, number_of_timesteps=70,
     cell_length=[0.1, 0.5, 0.5], number_of_cells=[48, 8, 8], number_of_patches=[12, 1, 1],
     EM_boundary_conditions=[["silver-muller"]])
MovingWindow(time_start=2.5, velocity_x=0.9997)
LaserGaussian3D(box_side="xmin", a0=1.0, omega=2.0, focus=[0., 2.0, 2.0], waist=1.5,
                time_envelope=tgaussian(start=0., duration=2.4, fwhm=0.8, center=1.2))
Species(name="electron", position_initialization="regular", momentum_initialization="cold",
        particles_per_cell=1, mass=1.0, charge=-1.0, charge_density=0.01, pusher="vay",
        boundary_conditions=[["remove"]])
"""


def _run_rank_window(rank, world, rank_grid, steps, port, ret):
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p = namelist.load_namelist(WINDOW_NAMELIST, is_source=True)
    sim = Simulation(p, rank_grid=rank_grid, rank=rank, patch_factory=OraclePatch)
    sim.create_particles()
    hist = sim.run(steps, scalars_every=1)
    parts = [sim.patch.species_get(0)]
    info = (parts, sim.patch.species_lost_energy(0), sim.simWindow.n_moved)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, info)
        dist.barrier()
        dist.destroy_process_group()
    else:
        gathered = [info]
    if rank == 0:
        ret["hist"] = [(h[0], h[1].tolist(), h[2]) for h in hist]
        ret["parts"] = [g[0] for g in gathered]
        ret["lost"] = sum(g[1] for g in gathered)
        ret["n_moved"] = [g[2] for g in gathered]


def _launch_window(rank_grid, steps):
    world = int(np.prod(rank_grid))
    if world == 1:
        ret = {}
        _run_rank_window(0, 1, rank_grid, steps, 0, ret)
        return ret
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run_rank_window, args=(world, rank_grid, steps, _free_port(), ret), nprocs=world, join=True)
    return dict(ret)


@pytest.mark.parametrize("rank_grid", [(1, 2, 1), (2, 1, 1)])
def test_moving_window_two_ranks_gloo_match_single_rank(rank_grid):
    """Laser + cold plasma + moving window with the box split in two, along y (every rank slides its patch by the
    same stride and creates the particles of its share of the uncovered cells) or along x (the +x rank hands its
    leftmost interior planes and the particles it leaves behind to the -x rank, only the +x rank creates
    particles): the run reproduces the single-rank one."""
    one = _launch_window((1, 1, 1), 70)
    two = _launch_window(rank_grid, 70)
    assert one["n_moved"][0] >= 10 * 4 and two["n_moved"] == [one["n_moved"][0]] * 2
    for (it_a, uk_a, ue_a), (it_b, uk_b, ue_b) in zip(one["hist"], two["hist"]):
        assert it_a == it_b
        assert np.allclose(uk_a, uk_b, rtol=1e-9, atol=1e-30)
        assert abs(ue_a - ue_b) <= 1e-9 * abs(ue_a)
    cols_a = {k: np.concatenate([r[0][k] for r in one["parts"]]) for k in ("x", "y", "z", "px", "py", "pz", "w")}
    cols_b = {k: np.concatenate([r[0][k] for r in two["parts"]]) for k in ("x", "y", "z", "px", "py", "pz", "w")}
    ia = np.lexsort((cols_a["z"], cols_a["y"], cols_a["x"]))
    ib = np.lexsort((cols_b["z"], cols_b["y"], cols_b["x"]))
    assert len(ia) == len(ib) > 0
    for k in cols_a:
        assert np.allclose(cols_a[k][ia], cols_b[k][ib], rtol=0, atol=1e-9), k


# ------------------------------------------------------------------------------------------------------------
# initial particles on the reference's per-patch streams, across ranks
def _run_rank_streams(rank, world, rank_grid, steps, port, ret):
    from test_reference_streams import thermal_short
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p = thermal_short(ncell=16, npatch=(2, 2, 2), nsteps=steps)
    sim = Simulation(p, rank_grid=rank_grid, rank=rank, patch_factory=OraclePatch)
    sim.create_particles(reference_streams=True)          # each rank draws the reference patches it holds
    hist = [(0,) + sim.scalars()] + sim.run(steps, scalars_every=1)
    parts = [sim.patch.species_get(s.ispec) for s in sim.vecSpecies]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, parts)
        dist.barrier()
        dist.destroy_process_group()
    else:
        gathered = [parts]
    if rank == 0:
        ret["hist"] = [(h[0], h[1].tolist(), h[2]) for h in hist]
        ret["parts"] = gathered


@pytest.mark.parametrize("rank_grid", [(1, 2, 1)])
def test_reference_streams_two_ranks_gloo_match_single_rank(rank_grid):
    """The namelist's particles do not depend on the rank layout: a 2-rank run that creates its particles from the
    reference's streams follows the 1-rank run (energies from step 0 on, every particle after 4 steps)."""
    ret1 = {}
    _run_rank_streams(0, 1, (1, 1, 1), 4, 0, ret1)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run_rank_streams, args=(2, rank_grid, 4, _free_port(), ret), nprocs=2, join=True)
    two = dict(ret)
    for (it_a, uk_a, ue_a), (it_b, uk_b, ue_b) in zip(ret1["hist"], two["hist"]):
        assert it_a == it_b
        assert np.allclose(uk_a, uk_b, rtol=1e-11, atol=0)
        assert abs(ue_a - ue_b) <= 1e-9 * abs(ue_a)
    for ispec in range(2):
        a = _canonical(ret1["parts"], ispec)
        b = _canonical(two["parts"], ispec)
        assert len(a["x"]) == len(b["x"]) == 16 ** 3 * 8
        for k in a:
            assert np.allclose(a[k], b[k], rtol=0, atol=1e-11), k


def _run_rank_diag(rank, world, rank_grid, n, port, ret):
    """Two ordinary steps then a diag step (diag_flag = True) with the electrons owning Jx_s .. rho_s; returns the
    interior (non-ghost) part of rho, Jx and the electrons' rho_s of this rank with its global offset."""
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p = make_params(n=n)
    sim = Simulation(p, rank_grid=rank_grid, rank=rank, patch_factory=OraclePatch)
    state = _global_state(n, 9)
    mn, mx = sim.patch.mn, sim.patch.mx
    for sp in sim.vecSpecies:
        a = state[sp.name]
        inside = np.ones(len(a["x"]), bool)
        for d, c in enumerate("xyz"):
            inside &= (a[c] >= mn[d]) & (a[c] < mx[d])
        sim.set_particles(sp.ispec, **{k: v[inside] for k, v in a.items()})
    sim.EMfields.allocateSpeciesFields(1)
    sim.run(2)
    sim.step(diag_flag=True)
    o = sim.patch.g.o[0]
    nloc = [sim.patch.g.n[d] for d in range(3)]
    out = {}
    for name in ("rho", "Jx", ("rho", 1), ("Jy", 1)):
        a = sim.patch.field_get(name)
        out[str(name)] = a[o:o + nloc[0], o:o + nloc[1], o:o + nloc[2]].copy()     # nodes [0, n) of the patch
    out["origin"] = [sim.pcoord[d] * nloc[d] for d in range(3)]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, out)
        dist.barrier()
        dist.destroy_process_group()
    else:
        gathered = [out]
    if rank == 0:
        ret["fields"] = gathered


def _assemble(pieces, n, key):
    full = np.zeros(n)
    for pc in pieces:
        a, o = pc[key], pc["origin"]
        full[o[0]:o[0] + a.shape[0], o[1]:o[1] + a.shape[1], o[2]:o[2] + a.shape[2]] = a
    return full


@pytest.mark.parametrize("rank_grid", [(2, 1, 1), (1, 1, 2)])
def test_diag_step_two_ranks_gloo_match_single_rank(rank_grid):
    """Diag step across ranks: rho is halo-summed with J (SyncVectorPatch::sumRhoJ with diag_flag), the species' own
    arrays too (sumRhoJs), and the totals contain the species arrays (computeTotalRhoJ) — the global rho, Jx and the
    electrons' rho_s / Jy_s of a 2-rank run equal the 1-rank run's on every node, shared planes and periodic wrap
    included."""
    n = (12, 12, 12)
    one = {}
    _run_rank_diag(0, 1, (1, 1, 1), n, 0, one)
    mgr = mp.Manager()
    two = mgr.dict()
    mp.spawn(_run_rank_diag, args=(2, rank_grid, n, _free_port(), two), nprocs=2, join=True)
    for key in ("rho", "Jx", "('rho', 1)", "('Jy', 1)"):
        a = _assemble(one["fields"], n, key)
        b = _assemble(two["fields"], n, key)
        assert np.abs(a).max() > 0
        assert np.max(np.abs(a - b)) <= 1e-11 * np.abs(a).max(), key
    # the totals hold BOTH species: rho of a neutral plasma is far smaller than the electrons' own rho_s
    assert np.abs(_assemble(one["fields"], n, "rho")).max() < np.abs(_assemble(one["fields"], n, "('rho', 1)")).max()


def test_moving_window_options_outside_the_path_are_rejected():
    """ MovingWindow keywords the path does not implement are refused at start-up, not ignored:
    additional shifts (SimWindow.cpp:95,100) and a stride equal to the whole box (a single patch along x)."""
    ok = namelist.load_namelist(WINDOW_NAMELIST, is_source=True)
    ok.check_hot_path()
    extra = WINDOW_NAMELIST.replace("MovingWindow(time_start=2.5, velocity_x=0.9997)",
                                    "MovingWindow(time_start=2.5, velocity_x=0.9997, number_of_additional_shifts=2, "
                                    "additional_shifts_time=3.)")
    with pytest.raises(namelist.NamelistError):
        namelist.load_namelist(extra, is_source=True).check_hot_path()
    one = WINDOW_NAMELIST.replace("number_of_patches=[12, 1, 1]", "number_of_patches=[1, 1, 1]")
    with pytest.raises(namelist.NamelistError):
        namelist.load_namelist(one, is_source=True).check_hot_path()


def test_regular_cold_cells_uniform_branch_matches_the_general_one():
    """particles_init.regular_cold_cells: a uniform plasma (every profile a number) takes a branch without per-cell
    arrays — the slab a moving window uncovers asks for it every few steps; it must return exactly what the general
    branch returns for the same plasma written as a profile."""
    from smilei_b200 import namelist, particles_init as pi
    src = '''
Main(geometry="3Dcartesian", interpolation_order=2, timestep=0.19, simulation_time=1.9, cell_length=[0.2,3.,3.],
     grid_length=[6.4,48.,48.], number_of_patches=[4,4,4], EM_boundary_conditions=[["silver-muller"]])
Species(name="electron", position_initialization="regular", momentum_initialization="cold", particles_per_cell=%s, mass=1.0,
        charge=-1.0, charge_density=%s, mean_velocity=[0.,0.,0.], temperature=[0.], pusher="vay",
        boundary_conditions=[["remove","remove"]]*3)
'''
    for ppc in ("1", "8"):
        p1 = namelist.load_namelist(src % (ppc, "0.000494"), is_source=True)
        p2 = namelist.load_namelist(src % (ppc, "lambda x,y,z: 0.000494 + 0*x"), is_source=True)
        for n, oc in (((8, 16, 16), (40, 0, 0)), ((32, 16, 16), None), ((8, 16, 16), (48, 0, 0))):
            a = pi.regular_cold_cells(p1, p1.species[0], n, (0, 0, 0), origin_cells=oc)
            b = pi.regular_cold_cells(p2, p2.species[0], n, (0, 0, 0), origin_cells=oc)
            assert a[0] == b[0] and a[4] == b[4] and a[5] == b[5]
            for x, y in zip(a[1:4], b[1:4]):
                assert x.dtype == y.dtype and np.array_equal(x, y)
    p0 = namelist.load_namelist(src % ("1", "0."), is_source=True)
    assert len(pi.regular_cold_cells(p0, p0.species[0], (8, 16, 16), (0, 0, 0))[1]) == 0
