"""ctypes binding of the C ABI declared in include/smilei_b200.h.

The shared library is built in-tree (smilei_b200/csrc/libsmilei_b200.so, see
__graft_entry__.build()).  There is no fallback of any kind: if the library is missing, or a
call fails, an exception is raised.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsmilei_b200.so")

FIELDS = ("Ex", "Ey", "Ez", "Bx", "By", "Bz", "Bxm", "Bym", "Bzm", "Jx", "Jy", "Jz", "rho")
FIELD_ID = {name: i for i, name in enumerate(FIELDS)}
# namelist aliases (src/ElectroMagn/ElectroMagn.h:163-190)
FIELD_ID.update({"Bx_m": FIELD_ID["Bxm"], "By_m": FIELD_ID["Bym"], "Bz_m": FIELD_ID["Bzm"], "Rho": FIELD_ID["rho"]})
SPECIES_FIELDS = ("Jx", "Jy", "Jz", "rho")


def field_id(name):
    """Field id of a name: one of FIELDS (or a namelist alias), or a species' own array written ("Jx", ispec) /
    "Jx_s<ispec>" (SB200_SPECIES_FIELD of include/smilei_b200.h; ElectroMagn::Jx_s .. rho_s)."""
    if isinstance(name, tuple):
        return len(FIELDS) + 4 * int(name[1]) + SPECIES_FIELDS.index(name[0])
    if "_s" in name and name.split("_s")[0] in SPECIES_FIELDS and name.split("_s")[1].isdigit():
        return len(FIELDS) + 4 * int(name.split("_s")[1]) + SPECIES_FIELDS.index(name.split("_s")[0])
    return FIELD_ID[name]


PUSHERS = {"boris": 0, "vay": 1, "higueracary": 2}
PBC = {"periodic": 0, "remove": 1}
DYN_KEEP_SCRATCH = 1
DYN_DIAG_RHO = 2
UNPACK_COPY, UNPACK_ADD = 0, 1
RECORD_DOUBLES = 8

# every symbol include/smilei_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = (
    "sb200_last_error", "sb200_abi_version", "sb200_device_count", "sb200_patch_create", "sb200_patch_destroy",
    "sb200_patch_set_stream", "sb200_patch_synchronize", "sb200_species_config", "sb200_species_set",
    "sb200_species_get", "sb200_species_count", "sb200_species_device_ptr", "sb200_species_first_index",
    "sb200_field_size", "sb200_field_set", "sb200_field_get", "sb200_field_device_ptr", "sb200_restart_rhoJ",
    "sb200_dynamics", "sb200_scratch_get", "sb200_maxwell", "sb200_center_B", "sb200_sort", "sb200_energy",
    "sb200_halo_plane_elems", "sb200_halo_pack", "sb200_halo_unpack", "sb200_halo_sum_self",
    "sb200_halo_exchange_self", "sb200_leaving_count", "sb200_leaving_pack", "sb200_leaving_pack_known", "sb200_arriving_unpack",
    "sb200_debug_flags", "sb200_species_init_thermal", "sb200_launch_count",
    "sb200_species_set_bc", "sb200_species_lost_energy", "sb200_apply_SM", "sb200_window_shift", "sb200_species_append",
    "sb200_hilbert_index3d", "sb200_create_particles_ref", "sb200_species_diag_fields", "sb200_compute_total_rhoJ",
    "sb200_species_append_regular",
)


class SmileiB200Error(RuntimeError):
    """Raised for every non-zero return of the C ABI (the reference's ERROR(), Tools.h:127-133)."""


class Grid(C.Structure):
    """sb200_grid."""
    _fields_ = [("n", C.c_int * 3), ("oversize", C.c_int * 3), ("cell_length", C.c_double * 3), ("dt", C.c_double),
                ("pcoord", C.c_int * 3), ("npatch", C.c_int * 3), ("interp_order", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SmileiB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C smilei_b200/csrc). There is no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.sb200_last_error.restype = C.c_char_p
        for s in SYMBOLS:
            getattr(_lib, s)  # AttributeError if the library and the header disagree
    return _lib


def _check(rc, what):
    if rc != 0:
        raise SmileiB200Error(f"{what}: {lib().sb200_last_error().decode()}")


def _p(a, dtype):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags["C_CONTIGUOUS"], (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


def device_count():
    n = C.c_int(0)
    _check(lib().sb200_device_count(C.byref(n)), "sb200_device_count")
    return n.value


def launch_count():
    n = C.c_ulonglong(0)
    _check(lib().sb200_launch_count(C.byref(n)), "sb200_launch_count")
    return n.value


# ---- initial particles on the reference's random streams (host side, init time only)
POSITION_INIT = {"regular": 0, "random": 1, "centered": 2, None: 3}
MOMENTUM_INIT = {"cold": 0, "maxwell-juettner": 1, "mj": 1}
_mj_tables = None


def mj_tables():
    """The two tables of ParticleCreator::maxwellJuttner (smilei_b200/data/mj_tables.npz)."""
    global _mj_tables
    if _mj_tables is None:
        d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "mj_tables.npz"))
        _mj_tables = (np.ascontiguousarray(d["lnInvF"]), np.ascontiguousarray(d["lnInvH"]))
    return _mj_tables


def hilbert_index3d(m, coords):
    """generalhilbertindex( m0, m1, m2, x, y, z ) of the reference (Hilbert_functions.cpp:246)."""
    h = C.c_uint(0)
    _check(lib().sb200_hilbert_index3d(C.c_uint(m[0]), C.c_uint(m[1]), C.c_uint(m[2]), int(coords[0]), int(coords[1]),
                                       int(coords[2]), C.byref(h)), "sb200_hilbert_index3d")
    return h.value


def create_particles_ref(rng_state, position_init, momentum_init, box, box_min, cell_length, nppc, n_real, charge,
                         temperature, mass, regular_number=None, positions=None):
    """sb200_create_particles_ref: (arrays, new rng_state) for one species in one reference patch."""
    nppc = np.ascontiguousarray(nppc, dtype=np.int32).ravel()
    n_real = np.ascontiguousarray(n_real, dtype=np.float64).ravel()
    charge = np.ascontiguousarray(charge, dtype=np.float64).ravel()
    temperature = np.ascontiguousarray(temperature, dtype=np.float64).ravel()
    cap = int(nppc[(n_real > 0) & (nppc > 0)].sum())
    out = {k: np.zeros(cap) for k in ("x", "y", "z", "px", "py", "pz", "w")}
    out["q"] = np.zeros(cap, dtype=np.int16)
    pinit = POSITION_INIT[position_init]
    if pinit == 3:
        for k, v in zip("xyz", positions):
            if len(v) != cap:
                raise SmileiB200Error("Copying particles: the two species should have the same number of particles")
            out[k][:] = v
    state = C.c_uint(int(rng_state) & 0xffffffff)
    n = C.c_size_t(0)
    F, H = mj_tables()
    reg = (C.c_int * 3)(*([int(v) for v in regular_number] if regular_number else [0, 0, 0]))
    _check(lib().sb200_create_particles_ref(
        C.byref(state), pinit, MOMENTUM_INIT[momentum_init], (C.c_int * 3)(*[int(v) for v in box]),
        (C.c_double * 3)(*box_min), (C.c_double * 3)(*cell_length), _p(nppc, np.int32), _p(n_real, np.float64),
        _p(charge, np.float64), _p(temperature, np.float64), C.c_double(mass), reg, _p(F, np.float64), _p(H, np.float64),
        *[_p(out[k], np.float64) for k in ("x", "y", "z", "px", "py", "pz", "w")], _p(out["q"], np.int16),
        C.c_size_t(cap), C.byref(n)), "sb200_create_particles_ref")
    assert n.value == cap
    return out, state.value


class Patch:
    """One patch on one GPU: owns the device fields and the per-species SoA (sb200_patch)."""

    def __init__(self, n, cell_length, dt, interp_order=2, n_species=0, pcoord=(0, 0, 0), npatch=(1, 1, 1),
                 oversize=None, device=0):
        oversize = (interp_order,) * 3 if oversize is None else tuple(oversize)
        self.grid = Grid((C.c_int * 3)(*n), (C.c_int * 3)(*oversize), (C.c_double * 3)(*cell_length), float(dt),
                         (C.c_int * 3)(*pcoord), (C.c_int * 3)(*npatch), int(interp_order))
        self.n = tuple(n)
        self.oversize = oversize
        self.order = interp_order
        self.n_species = n_species
        self.ncells = (n[0] + 1) * (n[1] + 1) * (n[2] + 1)
        self._h = C.c_void_p()
        _check(lib().sb200_patch_create(C.byref(self._h), C.byref(self.grid), n_species, device), "sb200_patch_create")

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().sb200_patch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        self.own_stream = bool(cuda_stream)
        _check(lib().sb200_patch_set_stream(self._h, C.c_void_p(cuda_stream)), "sb200_patch_set_stream")

    def bind_stream(self, cuda_stream):
        """Run the patch on the caller's stream (torch's current stream) and remember which one it was: the exchange
        layer checks that torch's current stream is still this one before every collective."""
        _check(lib().sb200_patch_set_stream(self._h, C.c_void_p(cuda_stream)), "sb200_patch_set_stream")
        self.bound_stream = int(cuda_stream)
        self.own_stream = False

    def synchronize(self):
        _check(lib().sb200_patch_synchronize(self._h), "sb200_patch_synchronize")

    # -- species
    def species_config(self, ispec, mass, pusher, capacity):
        pusher = PUSHERS[pusher] if isinstance(pusher, str) else int(pusher)
        _check(lib().sb200_species_config(self._h, ispec, C.c_double(mass), pusher, C.c_size_t(capacity)),
               "sb200_species_config")

    def species_set_bc(self, ispec, bc):
        """bc: six entries (xmin xmax ymin ymax zmin zmax), each 'periodic' / 'remove' or SB200_PBC_*."""
        codes = [PBC[b] if isinstance(b, str) else int(b) for b in bc]
        arr = (C.c_int * 6)(*codes)
        _check(lib().sb200_species_set_bc(self._h, ispec, arr), "sb200_species_set_bc")

    def species_lost_energy(self, ispec, reset=False):
        v = C.c_double(0.)
        _check(lib().sb200_species_lost_energy(self._h, ispec, C.byref(v), int(reset)), "sb200_species_lost_energy")
        return v.value

    def apply_SM(self, i_boundary, k, is_boundary=(0, 0, 0, 0), db1=None, db2=None):
        """ElectroMagnBC3D_SM::apply on one global box side; db1/db2 = laser amplitudes on the face or None."""
        kk = (C.c_double * 3)(*[float(v) for v in k])
        isb = (C.c_int * 4)(*[int(v) for v in is_boundary])
        a1 = None if db1 is None else np.ascontiguousarray(db1, dtype=np.float64)
        a2 = None if db2 is None else np.ascontiguousarray(db2, dtype=np.float64)
        _check(lib().sb200_apply_SM(self._h, int(i_boundary), kk, isb, None if a1 is None else _p(a1, np.float64),
                                    None if a2 is None else _p(a2, np.float64)), "sb200_apply_SM")

    def species_set(self, ispec, x, y, z, px, py, pz, w, q):
        n = len(x)
        cols = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z, px, py, pz, w)]
        q = np.ascontiguousarray(q, dtype=np.int16)
        _check(lib().sb200_species_set(self._h, ispec, *[_p(a, np.float64) for a in cols], _p(q, np.int16),
                                       C.c_size_t(n)), "sb200_species_set")

    def species_append(self, ispec, x, y, z, px, py, pz, w, q):
        n = len(x)
        if n == 0:
            return
        cols = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z, px, py, pz, w)]
        q = np.ascontiguousarray(q, dtype=np.int16)
        _check(lib().sb200_species_append(self._h, ispec, *[_p(a, np.float64) for a in cols], _p(q, np.int16),
                                          C.c_size_t(n)), "sb200_species_append")

    def species_append_regular(self, ispec, origin, box, regular_number, regular_inv, cells, weight, charge):
        """Device-side ParticleCreator (regular positions, cold): see sb200_species_append_regular."""
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        weight = np.ascontiguousarray(weight, dtype=np.float64)
        charge = np.ascontiguousarray(charge, dtype=np.int16)
        if len(cells) == 0:
            return
        _check(lib().sb200_species_append_regular(
            self._h, ispec, (C.c_double * 3)(*[float(v) for v in origin]), (C.c_int * 3)(*[int(v) for v in box]),
            (C.c_int * 3)(*[int(v) for v in regular_number]), (C.c_double * 3)(*[float(v) for v in regular_inv]),
            _p(cells, np.int32), _p(weight, np.float64), _p(charge, np.int16), C.c_size_t(len(cells))),
            "sb200_species_append_regular")

    def window_shift(self, ncells):
        _check(lib().sb200_window_shift(self._h, int(ncells)), "sb200_window_shift")

    def species_init_thermal(self, ispec, ppc, density, charge, temperature, seed=0):
        _check(lib().sb200_species_init_thermal(self._h, ispec, (C.c_int * 3)(*ppc), C.c_double(density), int(charge),
                                                C.c_double(temperature), C.c_ulonglong(seed)),
               "sb200_species_init_thermal")

    def species_count(self, ispec):
        n = C.c_size_t(0)
        _check(lib().sb200_species_count(self._h, ispec, C.byref(n)), "sb200_species_count")
        return n.value

    def species_get(self, ispec):
        n = self.species_count(ispec)
        out = {k: np.empty(n) for k in ("x", "y", "z", "px", "py", "pz", "w")}
        out["q"] = np.empty(n, dtype=np.int16)
        out["key"] = np.empty(n, dtype=np.int32)
        _check(lib().sb200_species_get(self._h, ispec, *[_p(out[k], np.float64) for k in
                                                         ("x", "y", "z", "px", "py", "pz", "w")],
                                       _p(out["q"], np.int16), _p(out["key"], np.int32), C.c_size_t(n)),
               "sb200_species_get")
        return out

    def species_device_ptr(self, ispec, column):
        ptr = C.c_void_p()
        _check(lib().sb200_species_device_ptr(self._h, ispec, column, C.byref(ptr)), "sb200_species_device_ptr")
        return ptr.value

    def first_index(self, ispec):
        first = np.empty(self.ncells + 1, dtype=np.int32)
        _check(lib().sb200_species_first_index(self._h, ispec, _p(first, np.int32), C.c_size_t(len(first))),
               "sb200_species_first_index")
        return first

    # -- fields
    def field_dims(self, name):
        dims = (C.c_int * 3)()
        n = C.c_size_t(0)
        _check(lib().sb200_field_size(self._h, field_id(name), C.byref(n), dims), "sb200_field_size")
        return tuple(dims)

    def field_set(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        _check(lib().sb200_field_set(self._h, field_id(name), _p(a, np.float64), C.c_size_t(a.size)),
               "sb200_field_set")

    def field_get(self, name):
        a = np.empty(self.field_dims(name))
        _check(lib().sb200_field_get(self._h, field_id(name), _p(a, np.float64), C.c_size_t(a.size)),
               "sb200_field_get")
        return a

    def field_device_ptr(self, name):
        ptr = C.c_void_p()
        alloc = (C.c_int * 3)()
        _check(lib().sb200_field_device_ptr(self._h, field_id(name), C.byref(ptr), alloc), "sb200_field_device_ptr")
        return ptr.value, tuple(alloc)

    # -- time step
    def species_diag_fields(self, ispec, Jx=True, Jy=True, Jz=True, rho=True):
        """The species' own Jx_s / Jy_s / Jz_s / rho_s (what a DiagFields of that species asks for): diag-step deposits
        of this species go there; field_get(("Jx", ispec)) reads them."""
        mask = int(bool(Jx)) | int(bool(Jy)) << 1 | int(bool(Jz)) << 2 | int(bool(rho)) << 3
        _check(lib().sb200_species_diag_fields(self._h, ispec, mask), "sb200_species_diag_fields")

    def compute_total_rhoJ(self):
        _check(lib().sb200_compute_total_rhoJ(self._h), "sb200_compute_total_rhoJ")

    def restart_rhoJ(self):
        _check(lib().sb200_restart_rhoJ(self._h), "sb200_restart_rhoJ")

    def dynamics(self, ispec, flags=0):
        _check(lib().sb200_dynamics(self._h, ispec, flags), "sb200_dynamics")

    def scratch_get(self, n):
        E = np.empty(3 * n)
        B = np.empty(3 * n)
        invgf = np.empty(n)
        iold = np.empty(3 * n, dtype=np.int32)
        delta = np.empty(3 * n)
        _check(lib().sb200_scratch_get(self._h, _p(E, np.float64), _p(B, np.float64), _p(invgf, np.float64),
                                       _p(iold, np.int32), _p(delta, np.float64), C.c_size_t(n)), "sb200_scratch_get")
        return E, B, invgf, iold, delta

    def maxwell(self):
        _check(lib().sb200_maxwell(self._h), "sb200_maxwell")

    def center_B(self):
        _check(lib().sb200_center_B(self._h), "sb200_center_B")

    def sort(self, ispec):
        _check(lib().sb200_sort(self._h, ispec), "sb200_sort")

    def energy(self):
        ukin = np.zeros(max(self.n_species, 1))
        uelm = C.c_double(0.)
        _check(lib().sb200_energy(self._h, _p(ukin, np.float64), C.byref(uelm)), "sb200_energy")
        return ukin[:self.n_species], uelm.value

    def debug_flags(self):
        f = np.zeros(8, dtype=np.int32)
        _check(lib().sb200_debug_flags(self._h, _p(f, np.int32)), "sb200_debug_flags")
        return f

    # -- halos (device buffers are raw pointers: torch tensors' data_ptr())
    def halo_plane_elems(self, name, dim):
        n = C.c_size_t(0)
        _check(lib().sb200_halo_plane_elems(self._h, field_id(name), dim, C.byref(n)), "sb200_halo_plane_elems")
        return n.value

    def halo_pack(self, name, dim, first_plane, nplanes, dev_ptr):
        _check(lib().sb200_halo_pack(self._h, field_id(name), dim, first_plane, nplanes, C.c_void_p(dev_ptr)),
               "sb200_halo_pack")

    def halo_unpack(self, name, dim, first_plane, nplanes, dev_ptr, mode):
        _check(lib().sb200_halo_unpack(self._h, field_id(name), dim, first_plane, nplanes, C.c_void_p(dev_ptr), mode),
               "sb200_halo_unpack")

    def halo_sum_self(self, name, dim):
        _check(lib().sb200_halo_sum_self(self._h, field_id(name), dim), "sb200_halo_sum_self")

    def halo_exchange_self(self, name, dim):
        _check(lib().sb200_halo_exchange_self(self._h, field_id(name), dim), "sb200_halo_exchange_self")

    # -- particle migration
    def leaving_count(self, ispec):
        c = (C.c_int * 6)()
        _check(lib().sb200_leaving_count(self._h, ispec, c), "sb200_leaving_count")
        return list(c)

    def leaving_pack(self, ispec, dim, side, wrap, dev_ptr, max_records):
        n = C.c_size_t(0)
        _check(lib().sb200_leaving_pack(self._h, ispec, dim, side, C.c_double(wrap), C.c_void_p(dev_ptr),
                                        C.c_size_t(max_records), C.byref(n)), "sb200_leaving_pack")
        return n.value

    def leaving_pack_known(self, ispec, dim, side, wrap, dev_ptr, max_records, n_known):
        _check(lib().sb200_leaving_pack_known(self._h, ispec, dim, side, C.c_double(wrap), C.c_void_p(dev_ptr),
                                              C.c_size_t(max_records), C.c_size_t(n_known)), "sb200_leaving_pack_known")
        return n_known

    def arriving_unpack(self, ispec, dev_ptr, n):
        _check(lib().sb200_arriving_unpack(self._h, ispec, C.c_void_p(dev_ptr), C.c_size_t(n)),
               "sb200_arriving_unpack")
