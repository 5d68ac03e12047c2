"""Initial particles of a namelist species (host side, init time only — not on the hot path).

Follows the reference's ParticleCreator (src/Particles/ParticleCreator.cpp) for what the 3D
thermal-plasma namelists use:
  * position_initialization "regular" (:627-667: x = cell_origin + dx*0.975*(0.5+i%c)/c, with
    regular_number or the cube root of particles_per_cell), "random", or the name of another
    species (same positions);
  * weight = density*cell_volume/nppc (:259, :933-939), charge (:964-974);
  * momentum_initialization "cold", "maxwell-juettner"/"mj" (:840-851,1002-1071).
The random stream is numpy's, not the reference's per-patch xorshift32 (src/Tools/Random.h), so
states are statistically, not bitwise, those of the reference; bitwise parity tests import
explicit arrays instead (the reference supports that too, ParticleCreator.cpp:344-529).
"""
import zlib

import numpy as np


def _profile(value, X, Y, Z):
    """Evaluate a namelist profile (constant or callable f(x,y,z)) on cell positions."""
    if callable(value):
        try:
            out = value(X, Y, Z)
            out = np.broadcast_to(np.asarray(out, dtype=float), X.shape).copy()
            return out
        except Exception:
            f = np.vectorize(lambda a, b, c: float(value(a, b, c)))
            return f(X, Y, Z)
    return np.full(X.shape, float(value))


def _regular_counts(nppc, regular_number):
    if regular_number:
        c = [int(v) for v in regular_number]
        if len(c) != 3 or c[0] * c[1] * c[2] != nppc:
            raise ValueError("regular_number is not coherent with particles_per_cell")          # :627-640
        return c
    coeff = round(nppc ** (1. / 3.))
    if coeff ** 3 != nppc:
        raise ValueError(f"Impossible to put {nppc} particles regularly spaced in one cell")   # :647-649
    return [coeff] * 3


def _maxwell_juttner(T, n, rng):
    """Momentum modulus p (units of m c) from a Maxwell-Juttner distribution of temperature T (units of
    m c^2).  T >= 0.1: Sobol's rejection method; colder: the relativistic correction to the modulus
    distribution is O(T), far below the sampling noise, and the rejection rate of Sobol's method diverges,
    so the Maxwellian modulus (chi distribution with 3 degrees of freedom) is drawn instead."""
    if T < 0.1:
        g = rng.standard_normal((3, n))
        return np.sqrt(T) * np.sqrt((g * g).sum(axis=0))
    out = np.empty(n)
    filled = 0
    while filled < n:
        m = int((n - filled) * 1.6) + 64
        u = rng.random((4, m))
        u = np.clip(u, 1e-300, 1.0)
        eta = -T * np.log(u[0] * u[1] * u[2])
        zeta = eta - T * np.log(u[3])
        ok = zeta * zeta - eta * eta > 1.0
        k = min(int(ok.sum()), n - filled)
        out[filled:filled + k] = eta[ok][:k]
        filled += k
    return out


def create(params, sp, n, pcoord, seed, rank, positions=None, origin_cells=None):
    """Arrays (x,y,z,px,py,pz,w,q) of species `sp` inside the patch at `pcoord` with `n` cells, or — with
    `origin_cells` — inside the box of `n` cells that starts at that global cell (the cells a moving window
    uncovers, SimWindow.cpp:372-392)."""
    cell = params.cell_length
    nppc_prof = sp.particles_per_cell
    if origin_cells is None:
        origin = [pcoord[d] * n[d] * cell[d] for d in range(3)]
    else:
        origin = [origin_cells[d] * cell[d] for d in range(3)]
    ic, jc, kc = np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing="ij")
    # profiles are evaluated at the cell centre (ParticleCreator.cpp:170-210: x_cell + 0.5 dx)
    X = origin[0] + (ic + 0.5) * cell[0]
    Y = origin[1] + (jc + 0.5) * cell[1]
    Z = origin[2] + (kc + 0.5) * cell[2]
    nppc = _profile(nppc_prof, X, Y, Z).astype(int)
    charge = _profile(sp.charge, X, Y, Z)
    if sp.charge_density is not None:
        dens = _profile(sp.charge_density, X, Y, Z)
        with np.errstate(divide="ignore", invalid="ignore"):
            dens = np.where(charge != 0, np.abs(dens / charge), 0.)                              # :250-256
    else:
        dens = np.abs(_profile(sp.number_density, X, Y, Z))
    dens = dens * params.cell_volume                                                             # :259
    keep = (dens > 0) & (nppc > 0)
    cells = np.flatnonzero(keep.ravel())
    npc = nppc.ravel()[cells]
    total = int(npc.sum())
    rng = np.random.default_rng([int(seed), int(rank), zlib.crc32(sp.name.encode()) & 0xffff])
    cell_of = np.repeat(cells, npc)
    first = np.concatenate([[0], np.cumsum(npc)[:-1]])
    local = np.arange(total) - np.repeat(first, npc)                    # index of the particle inside its cell
    ci = np.stack(np.unravel_index(cell_of, n), axis=0)
    pos = np.empty((3, total))
    posinit = sp.position_initialization
    if positions is not None:
        # position_initialization = "<other species>": same positions (ParticleCreator.cpp:300-340)
        if positions[0].shape[0] != total:
            raise ValueError("position_initialization from another species needs the same number of particles")
        pos[0], pos[1], pos[2] = positions
    elif posinit == "regular":
        uniq = np.unique(npc)
        for u in uniq:
            c = _regular_counts(int(u), sp.regular_number)
            sel = np.repeat(npc == u, npc)
            i = local[sel].copy()
            for d in range(3):
                pos[d, sel] = origin[d] + ci[d, sel] * cell[d] + cell[d] * 0.975 * (1. / c[d]) * (0.5 + i % c[d])
                i //= c[d]
    elif posinit == "random":
        for d in range(3):
            pos[d] = origin[d] + (ci[d] + rng.random(total)) * cell[d]
    else:
        raise ValueError(f"position_initialization `{posinit}` handled by the caller (species name) or unsupported")
    w = np.repeat(dens.ravel()[cells] / npc, npc)                                                # :933-939
    q = np.repeat(charge.ravel()[cells], npc).astype(np.int16)                                   # :964-974 (integer charges)
    mom = np.zeros((3, total))
    minit = sp.momentum_initialization
    if minit in ("maxwell-juettner", "mj"):
        T = float(sp.temperature[0]) / sp.mass                                                   # T in units of m c^2
        pmod = _maxwell_juttner(T, total, rng)
        # isotropic direction (ParticleCreator.cpp:1002-1071)
        phi = np.arccos(1. - 2. * rng.random(total))
        theta = 2. * np.pi * rng.random(total)
        mom[0] = pmod * np.sin(phi) * np.cos(theta)
        mom[1] = pmod * np.sin(phi) * np.sin(theta)
        mom[2] = pmod * np.cos(phi)
    elif minit != "cold":
        raise ValueError(f"momentum_initialization `{minit}` is not supported by this initialiser")
    if any(abs(v) > 0 for v in sp.mean_velocity):
        raise ValueError("mean_velocity != 0 is not supported by this initialiser")
    return dict(x=pos[0], y=pos[1], z=pos[2], px=mom[0], py=mom[1], pz=mom[2], w=w, q=q)
