"""ctypes access to the CPU oracle (oracle/libsmilei_oracle.so) and, when built, to the
reference's own operators (oracle/_ref/libsmilei_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libsmilei_oracle.so")
REF_SO = os.environ.get("SB200_REF_SO", os.path.join(ORACLE_DIR, "_ref", "libsmilei_ref.so"))
REF_FAST_SO = os.path.join(ORACLE_DIR, "_ref", "libsmilei_ref_fast.so")


class Grid(C.Structure):
    """orc_grid of oracle/smilei_oracle.h."""
    _fields_ = [("n", C.c_int * 3), ("o", C.c_int * 3), ("cell", C.c_double * 3), ("dt", C.c_double),
                ("pcoord", C.c_int * 3), ("npatch", C.c_int * 3), ("n_moved", C.c_int)]


def make_grid(n, order, cell, dt, pcoord=(0, 0, 0), npatch=(1, 1, 1), n_moved=0):
    return Grid((C.c_int * 3)(*n), (C.c_int * 3)(order, order, order), (C.c_double * 3)(*cell), float(dt),
                (C.c_int * 3)(*pcoord), (C.c_int * 3)(*npatch), int(n_moved))


def make_grid_moved(g, n_moved):
    """The same patch after the moving window advanced n_moved cells in total."""
    return make_grid(tuple(g.n), g.o[0], tuple(g.cell), g.dt, tuple(g.pcoord), tuple(g.npatch), n_moved)


# (dual_x, dual_y, dual_z) per field, ElectroMagn3D.cpp:115-123
DUAL = {"Ex": (1, 0, 0), "Ey": (0, 1, 0), "Ez": (0, 0, 1), "Bx": (0, 1, 1), "By": (1, 0, 1), "Bz": (1, 1, 0),
        "Jx": (1, 0, 0), "Jy": (0, 1, 0), "Jz": (0, 0, 1), "Bxm": (0, 1, 1), "Bym": (1, 0, 1), "Bzm": (1, 1, 0),
        "rho": (0, 0, 0)}


def field_dims(g, name):
    return tuple(g.n[i] + 2 * g.o[i] + 1 + DUAL[name][i] for i in range(3))


def patch_bounds(g):
    """Patch::initStep3 (Patch.cpp:146-147), same expression order."""
    mn = [(g.pcoord[i]) * (g.n[i] * g.cell[i]) for i in range(3)]
    mx = [(g.pcoord[i] + 1) * (g.n[i] * g.cell[i]) for i in range(3)]
    mn[0] += g.n_moved * g.cell[0]          # Patch.cpp:159-163
    mx[0] += g.n_moved * g.cell[0]
    return mn, mx


def ensure_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(ORACLE_DIR, "smilei_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libsmilei_oracle.so"])
    return ORACLE_SO


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class _Ops:
    """Common call surface of the oracle (`orc_*`) and the reference build (`ref_*`)."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        for name in ("field_norm2", "ukin", "uelm"):
            f = getattr(self.lib, prefix + name, None)
            if f is not None:
                f.restype = C.c_double
        for name in ("time_dynamics", "time_maxwell"):
            f = getattr(self.lib, prefix + name, None)
            if f is not None:
                f.restype = C.c_double

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def interp(self, g, order, F, x, y, z, istart=0, iend=None):
        n = len(x)
        iend = n if iend is None else iend
        E = np.zeros(3 * n)
        B = np.zeros(3 * n)
        iold = np.zeros(3 * n, dtype=np.int32)
        delta = np.zeros(3 * n)
        self._f("interp")(C.byref(g), order, _p(F["Ex"]), _p(F["Ey"]), _p(F["Ez"]), _p(F["Bxm"]), _p(F["Bym"]),
                          _p(F["Bzm"]), _p(x), _p(y), _p(z), n, istart, iend, _p(E), _p(B), _p(iold), _p(delta))
        return E, B, iold, delta

    def push(self, g, pusher, mass, x, y, z, px, py, pz, q, E, B, istart=0, iend=None):
        """In place on x..pz; returns invgf."""
        n = len(x)
        iend = n if iend is None else iend
        invgf = np.zeros(n)
        self._f("push")(C.byref(g), pusher, C.c_double(mass), _p(x), _p(y), _p(z), _p(px), _p(py), _p(pz), _p(q), n,
                        istart, iend, _p(E), _p(B), _p(invgf))
        return invgf

    def bc_tag(self, g, x, y, z):
        n = len(x)
        keys = np.zeros(n, dtype=np.int32)
        self._f("bc_tag")(C.byref(g), _p(x), _p(y), _p(z), _p(keys), 0, n)
        return keys

    def bc_apply(self, g, bc_remove, part):
        """PartBoundCond::apply with `remove` on the flagged global sides.  Returns (keys, q_after, energy_lost)."""
        n = len(part["x"])
        keys = np.zeros(n, dtype=np.int32)
        q = part["q"].copy()
        lost = C.c_double(0.)
        flags = (C.c_int * 6)(*[int(v) for v in bc_remove])
        self._f("bc_apply")(C.byref(g), flags, _p(part["x"]), _p(part["y"]), _p(part["z"]), _p(part["px"]),
                            _p(part["py"]), _p(part["pz"]), _p(part["w"]), _p(q), _p(keys), 0, n, C.byref(lost))
        return keys, q, lost.value

    def apply_SM(self, g, i_boundary, k, is_boundary, F, db1=None, db2=None):
        """ElectroMagnBC3D_SM::apply on Bx, By, Bz of F (in place)."""
        kk = (C.c_double * 3)(*[float(v) for v in k])
        isb = (C.c_int * 4)(*[int(v) for v in is_boundary])
        self._f("apply_SM")(C.byref(g), int(i_boundary), kk, isb, _p(F["Ex"]), _p(F["Ey"]), _p(F["Ez"]), _p(F["Bx"]),
                            _p(F["By"]), _p(F["Bz"]), _p(db1), _p(db2))

    def project(self, g, order, J, x, y, z, q, w, iold, delta, istart=0, iend=None):
        n = len(x)
        iend = n if iend is None else iend
        self._f("project")(C.byref(g), order, _p(J["Jx"]), _p(J["Jy"]), _p(J["Jz"]), _p(x), _p(y), _p(z), _p(q),
                           _p(w), n, istart, iend, _p(iold), _p(delta))

    def project_rho(self, g, order, J, x, y, z, q, w, iold, delta):
        """currentsAndDensityWrapper with diag_flag, either order, in place on J = {Jx, Jy, Jz, rho} (oracle only)."""
        n = len(x)
        self._f("project_rho")(C.byref(g), order, _p(J["Jx"]), _p(J["Jy"]), _p(J["Jz"]), _p(J["rho"]), _p(x), _p(y),
                               _p(z), _p(q), _p(w), n, 0, n, _p(iold), _p(delta))

    def project_rho_o2(self, g, J, x, y, z, q, w, iold, delta):
        n = len(x)
        self._f("project_rho_o2")(C.byref(g), _p(J["Jx"]), _p(J["Jy"]), _p(J["Jz"]), _p(J["rho"]), _p(x), _p(y),
                                  _p(z), _p(q), _p(w), n, 0, n, _p(iold), _p(delta))

    def save_B(self, g, F):
        self._f("save_B")(C.byref(g), _p(F["Bx"]), _p(F["By"]), _p(F["Bz"]), _p(F["Bxm"]), _p(F["Bym"]), _p(F["Bzm"]))

    def maxwell_ampere(self, g, F):
        self._f("maxwell_ampere")(C.byref(g), _p(F["Ex"]), _p(F["Ey"]), _p(F["Ez"]), _p(F["Bx"]), _p(F["By"]),
                                  _p(F["Bz"]), _p(F["Jx"]), _p(F["Jy"]), _p(F["Jz"]))

    def maxwell_faraday(self, g, F):
        self._f("maxwell_faraday")(C.byref(g), _p(F["Ex"]), _p(F["Ey"]), _p(F["Ez"]), _p(F["Bx"]), _p(F["By"]),
                                   _p(F["Bz"]))

    def center_B(self, g, F):
        self._f("center_B")(C.byref(g), _p(F["Bx"]), _p(F["By"]), _p(F["Bz"]), _p(F["Bxm"]), _p(F["Bym"]),
                            _p(F["Bzm"]))

    def cell_keys(self, g, x, y, z, keys=None, count=None):
        n = len(x)
        if keys is None:
            keys = np.zeros(n, dtype=np.int32)
        self._f("cell_keys")(C.byref(g), _p(x), _p(y), _p(z), _p(keys), _p(count), 0, n)
        return keys

    def field_norm2(self, g, f, name):
        d = DUAL[name]
        return self._f("field_norm2")(C.byref(g), _p(f), d[0], d[1], d[2])


class Oracle(_Ops):
    def __init__(self):
        super().__init__(ensure_oracle(), "orc_")
        self.lib.orc_counting_sort_perm.restype = C.c_int

    def counting_sort_perm(self, keys, ncells):
        n = len(keys)
        first = np.zeros(ncells + 1, dtype=np.int32)
        perm = np.zeros(n, dtype=np.int32)
        kept = self.lib.orc_counting_sort_perm(_p(keys), n, ncells, _p(first), _p(perm))
        return first, perm[:kept]

    def ukin(self, mass, px, py, pz, w):
        return self.lib.orc_ukin(C.c_double(mass), _p(px), _p(py), _p(pz), _p(w), len(px))

    def uelm(self, g, F):
        return self.lib.orc_uelm(C.byref(g), _p(F["Ex"]), _p(F["Ey"]), _p(F["Ez"]), _p(F["Bxm"]), _p(F["Bym"]),
                                 _p(F["Bzm"]))

    def compute_total_rhoJ(self, g, J, Js):
        self.lib.orc_compute_total_rhoJ(C.byref(g), _p(J["Jx"]), _p(J["Jy"]), _p(J["Jz"]), _p(J["rho"]),
                                        _p(Js["Jx"]), _p(Js["Jy"]), _p(Js["Jz"]), _p(Js["rho"]))

    def sum_pair(self, g, dim, name, L, R):
        d = DUAL[name]
        self.lib.orc_sum_pair(C.byref(g), dim, d[0], d[1], d[2], _p(L), _p(R))

    def exchange_pair(self, g, dim, name, L, R):
        d = DUAL[name]
        self.lib.orc_exchange_pair(C.byref(g), dim, d[0], d[1], d[2], _p(L), _p(R))


def have_ref():
    return os.path.exists(REF_SO)


class Reference(_Ops):
    """The reference's own operator classes (oracle/_ref), see oracle/ref_build/."""

    def __init__(self, fast=False):
        path = REF_FAST_SO if fast and os.path.exists(REF_FAST_SO) else REF_SO
        super().__init__(path, "ref_")
        self.path = path

    def time_dynamics(self, g, order, pusher, mass, F, part, npatches, nsteps, nthreads):
        fields6 = np.concatenate([F[k].ravel() for k in ("Ex", "Ey", "Ez", "Bxm", "Bym", "Bzm")])
        chk = C.c_double(0.)
        t = self.lib.ref_time_dynamics(C.byref(g), order, pusher, C.c_double(mass), _p(fields6), _p(part["x"]),
                                       _p(part["y"]), _p(part["z"]), _p(part["px"]), _p(part["py"]), _p(part["pz"]),
                                       _p(part["w"]), _p(part["q"]), len(part["x"]), npatches, nsteps, nthreads,
                                       C.byref(chk))
        return t, chk.value

    def time_dynamics_V(self, g, pusher, mass, F, part, first, npatches, nsteps, nthreads, with_sort=True, J_out=None):
        """Seconds for nsteps steps of the reference's VECTORISED species path (Interpolator3D2OrderV, pusher,
        computeParticleCellKeys, Projector3D2OrderV, sortParticles) on npatches patches; `part` cell-sorted with
        first_index `first`.  Returns (seconds, checksum of Jx of patch 0)."""
        fields6 = np.concatenate([F[k].ravel() for k in ("Ex", "Ey", "Ez", "Bxm", "Bym", "Bzm")])
        chk = C.c_double(0.)
        self.lib.ref_time_dynamics_V.restype = C.c_double
        first = np.ascontiguousarray(first, dtype=np.int32)
        t = self.lib.ref_time_dynamics_V(C.byref(g), pusher, C.c_double(mass), _p(fields6), _p(part["x"]), _p(part["y"]),
                                         _p(part["z"]), _p(part["px"]), _p(part["py"]), _p(part["pz"]), _p(part["w"]),
                                         _p(part["q"]), _p(first), len(part["x"]), npatches, nsteps, nthreads,
                                         int(with_sort), C.byref(chk), _p(J_out))
        return t, chk.value

    def time_maxwell(self, g, npatches, nsteps, nthreads):
        return self.lib.ref_time_maxwell(C.byref(g), npatches, nsteps, nthreads)

    def project_rho_species(self, g, order, J, Js, x, y, z, q, w, iold, delta, compute_total=True):
        """The reference's currentsAndDensityWrapper with diag_flag = true; Js = the species' own {Jx, Jy, Jz, rho}
        (or None: deposit into the totals J), then ElectroMagn3D::computeTotalRhoJ.  In place on J and Js."""
        n = len(x)
        sp = Js is not None
        z0 = [None] * 4 if not sp else [_p(Js[k]) for k in ("Jx", "Jy", "Jz", "rho")]
        self.lib.ref_project_rho_species(C.byref(g), order, int(sp), _p(J["Jx"]), _p(J["Jy"]), _p(J["Jz"]), _p(J["rho"]),
                                         *z0, _p(x), _p(y), _p(z), _p(q), _p(w), n, 0, n, _p(iold), _p(delta),
                                         int(compute_total))

    def sort(self, g, part, tags, arrivals=None):
        """The reference's own SpeciesV::computeParticleCellKeys + SpeciesV::sortParticles (SpeciesV.cpp:599-762) on
        the resident particles `part` (tags < 0: leavers, erased by the sort) with `arrivals` = six particle dicts
        (xmin xmax ymin ymax zmin zmax neighbour, or None) in MPI_buffer_.partRecv.  Returns (particles after the
        sort incl. 'key', first_index of ncell+1 entries)."""
        cols = ("x", "y", "z", "px", "py", "pz", "w")
        n = len(part["x"])
        arrivals = arrivals or [None] * 6
        narr = (C.c_int * 6)(*[0 if a is None else len(a["x"]) for a in arrivals])
        na = sum(narr)
        A = {k: np.ascontiguousarray(np.concatenate([a[k] for a in arrivals if a is not None] or [np.zeros(0)]))
             for k in cols}
        Aq = np.ascontiguousarray(np.concatenate([a["q"] for a in arrivals if a is not None] or
                                                 [np.zeros(0, dtype=np.int16)]).astype(np.int16))
        cap = n + na
        out = {k: np.zeros(cap) for k in cols}
        out["q"] = np.zeros(cap, dtype=np.int16)
        out["key"] = np.zeros(cap, dtype=np.int32)
        ncell = (g.n[0] + 1) * (g.n[1] + 1) * (g.n[2] + 1)
        first = np.zeros(ncell + 1, dtype=np.int32)
        tags = np.ascontiguousarray(tags, dtype=np.int32)
        self.lib.ref_sort.restype = C.c_int
        nout = self.lib.ref_sort(C.byref(g), n, *[_p(part[k]) for k in cols], _p(part["q"]), _p(tags), narr,
                                 *[_p(A[k]) for k in cols], _p(Aq), cap, *[_p(out[k]) for k in cols], _p(out["q"]),
                                 _p(out["key"]), _p(first))
        assert 0 <= nout <= cap
        return {k: v[:nout] for k, v in out.items()}, first


# ---------------------------------------------------------------------------------------
# seeded synthetic inputs shared by the oracle / reference / GPU parity tests
# ---------------------------------------------------------------------------------------

def random_fields(g, rng, names=("Ex", "Ey", "Ez", "Bx", "By", "Bz", "Bxm", "Bym", "Bzm", "Jx", "Jy", "Jz"),
                  scale=1.0):
    return {k: np.ascontiguousarray(scale * rng.standard_normal(field_dims(g, k))) for k in names}


def random_particles(g, rng, n, p_scale=0.3, charge=-1):
    mn, mx = patch_bounds(g)
    part = {}
    for i, c in enumerate("xyz"):
        # strictly inside [min, max): what Species::dynamics sees at the start of a step
        part[c] = np.ascontiguousarray(mn[i] + rng.random(n) * (mx[i] - mn[i]) * (1 - 1e-12))
        part["p" + c] = np.ascontiguousarray(p_scale * rng.standard_normal(n))
    part["w"] = np.ascontiguousarray(0.5 + rng.random(n))
    part["q"] = np.full(n, charge, dtype=np.int16)
    return part
