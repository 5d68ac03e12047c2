"""The namelists the validation tests run are RESTATED in the tests (the GPU box has no /root/reference).  Where the
reference tree is present (this container), each restated namelist is held against the reference's own benchmark
file loaded directly through smilei_b200.namelist: same box, time step, patches, boundary conditions, window,
species, density / temperature values and laser amplitudes on the entry face.  CPU only, skipped without the tree."""
import os

import numpy as np
import pytest

from smilei_b200 import namelist
from smilei_b200.laser import Laser

BENCH = "/root/reference/benchmarks"
pytestmark = pytest.mark.skipif(not os.path.isdir(BENCH), reason="reference tree not present")


def _value(v, pts):
    """A namelist value (number, list, or profile of position) reduced to numbers at sample points."""
    if callable(v):
        return [float(v(*p)) for p in pts]
    if isinstance(v, (list, tuple)):
        return [_value(e, pts) for e in v]
    return v


def _species_signature(params):
    L = params.grid_length
    pts = [(0.1 * L[0], 0.2 * L[1], 0.3 * L[2]), (0.5 * L[0], 0.5 * L[1], 0.5 * L[2]), (0.9 * L[0], 0.7 * L[1], 0.1 * L[2])]
    out = []
    for s in params.species:
        out.append(dict(name=s.name, mass=s.mass, charge=_value(s.charge, pts), pusher=s.pusher,
                        ppc=_value(s.particles_per_cell, pts), pos=s.position_initialization,
                        mom=s.momentum_initialization, T=_value(s.temperature, pts), v=_value(s.mean_velocity, pts),
                        rho=_value(s.charge_density, pts), n=_value(s.number_density, pts), bc=s.boundary_conditions))
    return out


def _main_signature(params):
    w = params.window
    return dict(geometry=params.geometry, order=params.interpolation_order, cell=params.cell_length,
                size=params.global_size, dt=params.timestep, n_time=params.n_time, patches=params.number_of_patches,
                bcs=params.EM_BCs, bcs_k=params.EM_BCs_k, oversize=params.oversize, seed=params.random_seed,
                window=None if w is None else (float(w.time_start), float(w.velocity_x)),
                n_lasers=len(params.laser_blocks))        # diagnostics blocks are left out of the restatements


def _laser_amplitudes(params, times):
    """By and Bz of every laser on its whole entry face (one patch = the whole box) at the given times."""
    out = []
    for block in params.laser_blocks:
        L = Laser(block, params)
        L.init_fields(params.global_size, params.oversize, params.cell_length, [0., 0., 0.])
        out.append([[L.amplitude(c, t) for c in (0, 1)] for t in times])
    return out


def _same(a, b, path=""):
    if isinstance(a, dict):
        assert a.keys() == b.keys(), path
        for k in a:
            _same(a[k], b[k], path + "." + str(k))
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), (path, a, b)
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, "%s[%d]" % (path, i))
    elif isinstance(a, float) or isinstance(b, float):
        assert a == pytest.approx(b, rel=1e-14, abs=0.), (path, a, b)
    else:
        assert a == b, (path, a, b)


def _check(restated, ref_file, laser_times=(), laser_rel=1e-12):
    ref = namelist.load_namelist(os.path.join(BENCH, ref_file))
    _same(_main_signature(restated), _main_signature(ref), "main")
    _same(_species_signature(restated), _species_signature(ref), "species")
    if laser_times:
        scale = max(np.max(np.abs(a)) for las in _laser_amplitudes(ref, laser_times) for t in las for a in t)
        assert scale > 0.
        for got, want in zip(_laser_amplitudes(restated, laser_times), _laser_amplitudes(ref, laser_times)):
            for g_t, w_t in zip(got, want):
                for g, w in zip(g_t, w_t):
                    assert g.shape == w.shape
                    assert np.max(np.abs(g - w)) <= laser_rel * scale, np.max(np.abs(g - w)) / scale
    return ref


@pytest.mark.parametrize("case,ref_file", [
    ("vay", "tst3d_s_o2_laser_wake_yee_vay.py"), ("higueracary", "tst3d_s_o2_laser_wake_yee_higuera.py"),
    ("boris", "tst3d_v_o2_laser_wake_yee_boris.py"), ("boris_o4", "tst3d_v_o4_laser_wake_boris.py")])
def test_laser_wake_restatement(case, ref_file):
    """BASELINE.json configs[3] and its siblings: LASER_WAKE_NAMELIST of test_gpu_simulation.py uses LaserGaussian3D
    where the reference files write the beam out by hand; the amplitudes on the xmin face must agree."""
    from test_gpu_simulation import LASER_WAKE_CASES, LASER_WAKE_NAMELIST
    pusher, order, npatch_x, kvec = LASER_WAKE_CASES[case]
    restated = namelist.load_namelist(LASER_WAKE_NAMELIST.format(pusher=pusher, order=order, npatch_x=npatch_x, kvec=kvec),
                                      is_source=True)
    _check(restated, ref_file, laser_times=(5., 20., 28., 40., 60.), laser_rel=1e-9)


def test_em_propagation_restatement():
    from test_gpu_simulation import EM_PROPAGATION_NAMELIST
    restated = namelist.load_namelist(EM_PROPAGATION_NAMELIST, is_source=True)
    _check(restated, "tst3d_00_em_propagation.py", laser_times=(3., 17., 40.), laser_rel=1e-12)


def test_thermal_medium_restatement():
    from test_gpu_simulation import THERMAL_MEDIUM
    _check(namelist.load_namelist(THERMAL_MEDIUM, is_source=True), "gpu/tst3d_v_o2_thermal_plasma_medium.py")


@pytest.mark.parametrize("order,ref_file", [(2, "tst3d_v_o2_thermal_plasma.py"), (4, "tst3d_v_o4_thermal_plasma.py"),
                                            (2, "tst3d_01_thermal_plasma.py"), (2, "gpu/tst3d_gpu_o2_thermal_plasma.py")])
def test_tst3d_thermal_restatement(order, ref_file):
    from test_reference_streams import tst3d_thermal
    _check(tst3d_thermal(order), ref_file)


def test_thermal_short_restatement():
    """BASELINE.json configs[0]: gpu/tst3d_gpu_o2_thermal_plasma_short.py (32^3, 4^3 patches, random positions)."""
    from test_reference_streams import thermal_short
    _check(thermal_short(32, (4, 4, 4), "random", 2001), "gpu/tst3d_gpu_o2_thermal_plasma_short.py")
