"""OraclePatch: the smilei_b200.capi.Patch interface implemented on the CPU with the oracle.

TEST INFRASTRUCTURE ONLY.  It lets the product's host logic (smilei_b200.simulation /
smilei_b200.exchange: call order, neighbour plan, message layout) run without a GPU — in
particular under torch.distributed/gloo with several processes — and provides the multi-step
oracle trajectory the GPU runs are compared with.  Buffers are torch CPU tensors addressed
by data_ptr(), exactly like the CUDA library addresses device memory.
"""
import ctypes as C

import numpy as np

import oracle_lib as ol

RECORD = 8


def _view(ptr, n):
    if n == 0:
        return np.zeros(0)
    return np.ctypeslib.as_array((C.c_double * n).from_address(ptr))


class OraclePatch:
    def __init__(self, n, cell_length, dt, interp_order=2, n_species=0, pcoord=(0, 0, 0), npatch=(1, 1, 1),
                 oversize=None, device=0):
        self.g = ol.make_grid(n, interp_order, cell_length, dt, pcoord, npatch)
        self.n = tuple(n)
        self.order = interp_order
        self.n_species = n_species
        self.orc = ol.Oracle()
        self.F = {k: np.zeros(ol.field_dims(self.g, k)) for k in
                  ("Ex", "Ey", "Ez", "Bx", "By", "Bz", "Bxm", "Bym", "Bzm", "Jx", "Jy", "Jz", "rho")}
        self.sp = [dict(mass=1., pusher=0, P=None, first=None, bc=[0] * 6, lost=0.) for _ in range(n_species)]
        self.ncells = (n[0] + 1) * (n[1] + 1) * (n[2] + 1)
        self.mn, self.mx = ol.patch_bounds(self.g)

    def close(self):
        pass

    def synchronize(self):
        pass

    # -- species
    def species_config(self, ispec, mass, pusher, capacity):
        from smilei_b200.capi import PUSHERS
        self.sp[ispec]["mass"] = mass
        self.sp[ispec]["pusher"] = PUSHERS[pusher] if isinstance(pusher, str) else int(pusher)

    def species_set_bc(self, ispec, bc):
        from smilei_b200.capi import PBC
        self.sp[ispec]["bc"] = [PBC[b] if isinstance(b, str) else int(b) for b in bc]

    def species_lost_energy(self, ispec, reset=False):
        v = self.sp[ispec]["mass"] * self.sp[ispec]["lost"]
        if reset:
            self.sp[ispec]["lost"] = 0.
        return v

    def apply_SM(self, i_boundary, k, is_boundary=(0, 0, 0, 0), db1=None, db2=None):
        a1 = None if db1 is None else np.ascontiguousarray(db1, dtype=np.float64)
        a2 = None if db2 is None else np.ascontiguousarray(db2, dtype=np.float64)
        self.orc.apply_SM(self.g, i_boundary, k, is_boundary, self.F, a1, a2)

    def species_append(self, ispec, x, y, z, px, py, pz, w, q):
        P = self.sp[ispec]["P"]
        new = dict(x=x, y=y, z=z, px=px, py=py, pz=pz, w=w)
        for k, v in new.items():
            P[k] = np.concatenate([P[k], np.asarray(v, dtype=np.float64)])
        P["q"] = np.concatenate([P["q"], np.asarray(q, dtype=np.int16)])
        P["key"] = np.concatenate([P["key"], np.zeros(len(x), dtype=np.int32)])
        self.sp[ispec]["sorted"] = False

    def window_shift(self, ncells):
        """Cell-granular SimWindow::shift of a patch spanning the box along x (see sb200_window_shift)."""
        for k in ("Ex", "Ey", "Ez", "Bx", "By", "Bz", "Bxm", "Bym", "Bzm"):
            a = self.F[k]
            a[:-ncells] = a[ncells:].copy()
            a[-ncells:] = 0.
        self.n_moved = getattr(self, "n_moved", 0) + ncells
        self.g = ol.make_grid_moved(self.g, self.n_moved) if hasattr(ol, "make_grid_moved") else self.g
        self.mn, self.mx = ol.patch_bounds(self.g)
        for s in self.sp:
            P = s["P"]
            if P is None:
                continue
            behind = P["x"] < self.mn[0]
            if self.g.pcoord[0] > 0:
                P["key"] = np.where(behind, -2, np.where(P["key"] < 0, -1, 0)).astype(np.int32)    # the -x neighbour takes them
            else:
                for k in P:
                    P[k] = np.ascontiguousarray(P[k][~behind])
                P["key"] = np.where(P["key"] < 0, -1, 0).astype(np.int32)
            s["sorted"] = False

    def species_set(self, ispec, x, y, z, px, py, pz, w, q):
        P = dict(x=x, y=y, z=z, px=px, py=py, pz=pz, w=w)
        P = {k: np.ascontiguousarray(v, dtype=np.float64).copy() for k, v in P.items()}
        P["q"] = np.ascontiguousarray(q, dtype=np.int16).copy()
        P["key"] = np.zeros(len(x), dtype=np.int32)
        self.sp[ispec]["P"] = P
        self.sp[ispec]["sorted"] = False

    def species_count(self, ispec):
        P = self.sp[ispec]["P"]
        return 0 if P is None else len(P["x"])

    def species_get(self, ispec):
        return {k: v.copy() for k, v in self.sp[ispec]["P"].items()}

    def first_index(self, ispec):
        return self.sp[ispec]["first"].copy()

    # -- fields
    def _key(self, name):
        """Patch field name, or a species' own array: ("Jx", ispec) / "Jx_s<ispec>"."""
        if isinstance(name, tuple):
            name = "%s_s%d" % (name[0], int(name[1]))
        if name not in self.F:
            from smilei_b200.capi import SmileiB200Error
            raise SmileiB200Error("bad field id (or a species array that was not requested): %s" % name)
        return name

    def _base(self, name):
        name = self._key(name)
        return name.split("_s")[0] if "_s" in name else name

    def field_dims(self, name):
        return self.F[self._key(name)].shape

    def field_set(self, name, a):
        self.F[self._key(name)][...] = a

    def field_get(self, name):
        return self.F[self._key(name)].copy()

    def species_diag_fields(self, ispec, Jx=True, Jy=True, Jz=True, rho=True):
        for n, on in zip(("Jx", "Jy", "Jz", "rho"), (Jx, Jy, Jz, rho)):
            k = "%s_s%d" % (n, ispec)
            if on and k not in self.F:
                self.F[k] = np.zeros(ol.field_dims(self.g, n))
            elif not on:
                self.F.pop(k, None)

    def compute_total_rhoJ(self):
        for k in list(self.F):
            if "_s" in k:
                self.F[k.split("_s")[0]] += self.F[k]

    # -- time step
    def restart_rhoJ(self):
        for k in self.F:
            if k.split("_s")[0] in ("Jx", "Jy", "Jz", "rho"):
                self.F[k][...] = 0.

    def dynamics(self, ispec, flags=0):
        s = self.sp[ispec]
        P = s["P"]
        if P is None or len(P["x"]) == 0:
            return
        g, o = self.g, self.orc
        E, B, iold, delta = o.interp(g, self.order, self.F, P["x"], P["y"], P["z"])
        o.push(g, s["pusher"], s["mass"], P["x"], P["y"], P["z"], P["px"], P["py"], P["pz"], P["q"], E, B)
        if any(s["bc"]):
            # `remove` only acts where the patch touches the global box side (PartBoundCond.cpp:99-245)
            eff = [s["bc"][2 * d + sd] == 1 and g.pcoord[d] == (0 if sd == 0 else g.npatch[d] - 1)
                   for d in range(3) for sd in range(2)]
            tags, q_after, lost = o.bc_apply(g, eff, P)
            P["q"] = q_after
            s["lost"] += lost
        else:
            tags = o.bc_tag(g, P["x"], P["y"], P["z"])
        if flags & 2:     # diag step: currentsAndDensityWrapper with diag_flag, into the species' own arrays where it has them
            J = {n: self.F.get("%s_s%d" % (n, ispec), self.F[n]) for n in ("Jx", "Jy", "Jz", "rho")}
            o.project_rho(g, self.order, J, P["x"], P["y"], P["z"], P["q"], P["w"], iold, delta)
        else:
            o.project(g, self.order, self.F, P["x"], P["y"], P["z"], P["q"], P["w"], iold, delta)
        keys = tags.copy()
        o.cell_keys(g, P["x"], P["y"], P["z"], keys=keys)
        P["key"] = keys

    def maxwell(self):
        self.orc.save_B(self.g, self.F)
        self.orc.maxwell_ampere(self.g, self.F)
        self.orc.maxwell_faraday(self.g, self.F)

    def center_B(self):
        self.orc.center_B(self.g, self.F)

    def sort(self, ispec):
        s = self.sp[ispec]
        P = s["P"]
        if P is None:
            s["first"] = np.zeros(self.ncells + 1, dtype=np.int32)
            return
        if not s.get("sorted", False):
            keys = P["key"].copy()
            keys[keys >= 0] = 0
            self.orc.cell_keys(self.g, P["x"], P["y"], P["z"], keys=keys)
            P["key"] = keys
        first, perm = self.orc.counting_sort_perm(P["key"], self.ncells)
        s["P"] = {k: np.ascontiguousarray(v[perm]) for k, v in P.items()}
        s["first"] = first
        s["sorted"] = True

    def energy(self):
        uk = np.array([0. if s["P"] is None else self.orc.ukin(s["mass"], s["P"]["px"], s["P"]["py"], s["P"]["pz"],
                                                                 s["P"]["w"]) for s in self.sp])
        return uk, self.orc.uelm(self.g, self.F)

    # -- halos
    def halo_plane_elems(self, name, dim):
        name = self._key(name)
        d = self.F[name].shape
        return int(np.prod([d[i] for i in range(3) if i != dim]))

    def _slab(self, name, dim, first, npl):
        sl = [slice(None)] * 3
        sl[dim] = slice(first, first + npl)
        return tuple(sl)

    def halo_pack(self, name, dim, first_plane, nplanes, ptr):
        name = self._key(name)
        a = np.moveaxis(self.F[name][self._slab(name, dim, first_plane, nplanes)], dim, 0)
        _view(ptr, a.size)[:] = a.reshape(-1)

    def halo_unpack(self, name, dim, first_plane, nplanes, ptr, mode):
        name = self._key(name)
        sl = self._slab(name, dim, first_plane, nplanes)
        shape = np.moveaxis(self.F[name][sl], dim, 0).shape
        a = np.moveaxis(_view(ptr, int(np.prod(shape))).reshape(shape), 0, dim)
        if mode == 1:
            self.F[name][sl] += a
        else:
            self.F[name][sl] = a

    def halo_sum_self(self, name, dim):
        self.orc.sum_pair(self.g, dim, self._base(name), self.F[self._key(name)], self.F[self._key(name)])

    def halo_exchange_self(self, name, dim):
        self.orc.exchange_pair(self.g, dim, self._base(name), self.F[self._key(name)], self.F[self._key(name)])

    # -- particles
    def leaving_count(self, ispec):
        P = self.sp[ispec]["P"]
        if P is None:
            return [0] * 6
        return [int((P["key"] == -2 - t).sum()) for t in range(6)]

    def leaving_pack(self, ispec, dim, side, wrap, ptr, max_records):
        P = self.sp[ispec]["P"]
        if P is None:
            return 0
        idx = np.flatnonzero(P["key"] == -2 - 2 * dim - side)
        k = len(idx)
        assert k <= max_records
        rec = np.empty((k, RECORD))
        for c, name in enumerate(("x", "y", "z", "px", "py", "pz", "w")):
            rec[:, c] = P[name][idx]
        rec[:, 7] = P["q"][idx]
        hi = self.g.cell[dim] * float(self.g.n[dim] * self.g.npatch[dim])
        if wrap > 0:
            m = rec[:, dim] < 0.
            rec[m, dim] += wrap
        elif wrap < 0:
            m = rec[:, dim] >= hi
            rec[m, dim] += wrap
        _view(ptr, k * RECORD)[:] = rec.reshape(-1)
        return k

    def leaving_pack_known(self, ispec, dim, side, wrap, ptr, max_records, n_known):
        k = self.leaving_pack(ispec, dim, side, wrap, ptr, max_records)
        assert k == n_known
        return k

    def arriving_unpack(self, ispec, ptr, n):
        if n == 0:
            return
        rec = _view(ptr, n * RECORD).reshape(n, RECORD).copy()
        P = self.sp[ispec]["P"]
        x, y, z = (np.ascontiguousarray(rec[:, c]) for c in range(3))
        keys = self.orc.bc_tag(self.g, x, y, z)
        self.orc.cell_keys(self.g, x, y, z, keys=keys)
        for c, name in enumerate(("x", "y", "z", "px", "py", "pz", "w")):
            P[name] = np.concatenate([P[name], rec[:, c]])
        P["q"] = np.concatenate([P["q"], rec[:, 7].astype(np.int16)])
        P["key"] = np.concatenate([P["key"], keys])
