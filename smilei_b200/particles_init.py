"""Initial particles of a namelist species (host side, init time only — not on the hot path).

Follows the reference's ParticleCreator (src/Particles/ParticleCreator.cpp) for what the 3D
thermal-plasma namelists use:
  * position_initialization "regular" (:627-667: x = cell_origin + dx*0.975*(0.5+i%c)/c, with
    regular_number or the cube root of particles_per_cell), "random", or the name of another
    species (same positions);
  * weight = density*cell_volume/nppc (:259, :933-939), charge (:964-974);
  * momentum_initialization "cold", "maxwell-juettner"/"mj" (:840-851,1002-1071).
Two random streams:
  * `create`: numpy's generator — states are statistically those of the reference (synthetic workloads);
  * `create_reference_streams`: the reference's own per-patch xorshift32 streams
    (src/Tools/Random.h, seed random_seed + Hilbert index of the patch, src/Patch/Patch.cpp:129) walked in
    the reference's order by the C ABI (sb200_create_particles_ref), so that a namelist run starts from
    the reference's particles bit for bit whatever the rank layout.
"""
import zlib

import numpy as np


_UNIFORM_CACHE = {}          # regular_cold_cells: the cell list of the last uniform (species, box) asked for


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
        return c, [1. / v for v in c]
    # :645-653 as written: the count per dimension is the double pow( nppc, 1/3 ) stored in an int (64 -> 3, because
    # pow( 64., 1./3. ) = 3.9999999999999996) while the spacing uses the double
    coeff = float(nppc) ** (1. / 3.)
    if round(coeff) ** 3 != nppc:
        raise ValueError(f"Impossible to put {nppc} particles regularly spaced in one cell")   # :647-649
    return [int(coeff)] * 3, [1. / coeff] * 3


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


def _log2_exact(v, what):
    m = int(v).bit_length() - 1
    if v <= 0 or (1 << m) != v:
        raise ValueError(f"{what} = {v} must be a power of 2 (Params.cpp:716-720)")
    return m


def _cell_profiles(params, sp, box, origin):
    """nppc, n_real (= |density| x cell volume), charge and temperature of the `box` cells starting at `origin`
    (ParticleCreator.cpp:150-263; profiles are evaluated at the cell centre)."""
    cell = params.cell_length
    ic, jc, kc = np.meshgrid(np.arange(box[0]), np.arange(box[1]), np.arange(box[2]), indexing="ij")
    X = origin[0] + (ic + 0.5) * cell[0]
    Y = origin[1] + (jc + 0.5) * cell[1]
    Z = origin[2] + (kc + 0.5) * cell[2]
    nppc = np.floor(_profile(sp.particles_per_cell, X, Y, Z)).astype(np.int32)
    charge = _profile(sp.charge, X, Y, Z)
    if sp.charge_density is not None:
        dens = _profile(sp.charge_density, X, Y, Z)
        dens = np.where(np.abs(dens) < 1e-200, 0., dens)                                         # :236-238
        with np.errstate(divide="ignore", invalid="ignore"):
            dens = np.where(dens != 0, np.abs(dens / charge), 0.)                                # :250-256
    else:
        dens = _profile(sp.number_density, X, Y, Z)
        dens = np.abs(np.where(np.abs(dens) < 1e-200, 0., dens))
    n_real = dens * params.cell_volume                                                           # :259
    n_real = np.where(nppc > 0, n_real, 0.)
    T = sp.temperature[0] if sp.temperature and sp.temperature[0] is not None else 1e-10       # :164 default
    temperature = _profile(T, X, Y, Z)
    return nppc, n_real, charge, temperature


def create_reference_streams(params, species, n, pcoord):
    """Initial particles of every species of the namelist inside the rank box (`n` cells at rank coordinates
    `pcoord`), drawn from the reference's per-patch streams: {species name: arrays}.

    The box is walked by reference patches (Main.number_of_patches of the namelist, Hilbert-numbered,
    HilbertDomainDecomposition.cpp:83-87); each patch owns one xorshift32 stream seeded with random_seed +
    hindex (Patch.cpp:129) that the species consume one after the other (SpeciesFactory.h:1153-1157), each
    cell by cell (ParticleCreator.cpp:300-338)."""
    from . import capi
    if getattr(params, "patch_arrangement", "hilbertian") != "hilbertian":
        raise ValueError("reference streams need patch_arrangement = 'hilbertian'")
    npatch = params.number_of_patches
    m = [_log2_exact(npatch[d], f"number_of_patches[{d}]") for d in range(3)]
    psize = [params.global_size[d] // npatch[d] for d in range(3)]
    for d in range(3):
        if psize[d] * npatch[d] != params.global_size[d] or n[d] % psize[d]:
            raise ValueError("the rank box must be made of whole reference patches")
    first = [pcoord[d] * n[d] // psize[d] for d in range(3)]
    count = [n[d] // psize[d] for d in range(3)]
    patches = []
    for a in range(count[0]):
        for b in range(count[1]):
            for c in range(count[2]):
                P = (first[0] + a, first[1] + b, first[2] + c)
                h = capi.hilbert_index3d(m, P)
                patches.append((h, P))
    patches.sort()                                                    # the reference's patch order on a rank
    state = {h: (params.random_seed + h) & 0xffffffff or 0xffffffff for h, _ in patches}       # Random.h:91-99
    columns = ("x", "y", "z", "px", "py", "pz", "w", "q")
    created = {}
    for sp in species:
        if any(abs(v) > 0 for v in sp.mean_velocity):
            raise ValueError("mean_velocity != 0 is not supported by this initialiser")
        posinit = sp.position_initialization
        src = created.get(posinit)
        if src is None and posinit not in ("regular", "random", "centered"):
            raise ValueError(f"position_initialization `{posinit}` is neither a method nor an earlier species")
        parts = []
        for ip, (h, P) in enumerate(patches):
            box_min = [P[d] * (psize[d] * params.cell_length[d]) for d in range(3)]              # Patch.cpp:149
            nppc, n_real, charge, temperature = _cell_profiles(params, sp, psize, box_min)
            arrays, state[h] = capi.create_particles_ref(
                state[h], None if src is not None else posinit, sp.momentum_initialization, psize, box_min,
                params.cell_length, nppc, n_real, charge, temperature, sp.mass, sp.regular_number,
                positions=None if src is None else [src["_parts"][ip][k] for k in "xyz"])
            parts.append(arrays)
        created[sp.name] = {k: np.concatenate([p[k] for p in parts]) for k in columns}
        created[sp.name]["_parts"] = parts
    for v in created.values():
        del v["_parts"]
    return created


def regular_cold_cells(params, sp, n, pcoord, origin_cells=None):
    """What the device-side creator (sb200_species_append_regular) needs for species `sp` in a box of `n` cells, or None
    when the species is not of that kind: position_initialization "regular", momentum_initialization "cold", zero mean
    velocity and the same particles_per_cell in every kept cell.  Returns (origin, cells, weight, charge,
    regular_number, regular_inv): the profiles are evaluated here exactly as `create` does."""
    if sp.position_initialization != "regular" or sp.momentum_initialization != "cold":
        return None
    if any(abs(v) > 0 for v in sp.mean_velocity):
        return None
    cell = params.cell_length
    if origin_cells is None:
        origin = [pcoord[d] * n[d] * cell[d] for d in range(3)]
    else:
        origin = [origin_cells[d] * cell[d] for d in range(3)]
    dens_prof = sp.charge_density if sp.charge_density is not None else sp.number_density
    if not (callable(sp.particles_per_cell) or callable(sp.charge) or callable(dens_prof)):
        # uniform plasma (every profile a number): the same values as below without the arrays over the cells; the
        # slab a moving window uncovers asks for this every few steps (22 ms per call through the general branch
        # for 8 x 256 x 256 cells, a third of a laser-wake step)
        key = (sp.name, tuple(int(v) for v in n))
        hit = _UNIFORM_CACHE.get(key)
        sig = (float(sp.particles_per_cell), float(sp.charge), float(dens_prof), sp.charge_density is not None,
               float(params.cell_volume), tuple(sp.regular_number or ()))
        if hit is None or hit[0] != sig:
            nppc1 = int(np.float64(sp.particles_per_cell))
            charge1 = np.float64(sp.charge)
            if sp.charge_density is not None:
                dens1 = np.abs(np.float64(dens_prof) / charge1) if charge1 != 0 else np.float64(0.)
            else:
                dens1 = np.abs(np.float64(dens_prof))
            dens1 = dens1 * params.cell_volume
            ncell = int(n[0]) * int(n[1]) * int(n[2])
            if not (dens1 > 0 and nppc1 > 0):
                hit = (sig, (np.zeros(0, dtype=np.int32), np.zeros(0), np.zeros(0, dtype=np.int16), [1, 1, 1], [1., 1., 1.]))
            else:
                c, inv = _regular_counts(nppc1, sp.regular_number)
                hit = (sig, (np.arange(ncell, dtype=np.int32), np.full(ncell, dens1 / nppc1),
                             np.full(ncell, charge1).astype(np.int16), c, inv))
            for arr in hit[1][:3]:
                arr.setflags(write=False)              # handed out again on the next call: nobody may write into them
            _UNIFORM_CACHE.clear()
            _UNIFORM_CACHE[key] = hit
        cells, weight, charge_i, c, inv = hit[1]
        return origin, cells, weight, charge_i, c, inv
    ic, jc, kc = np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing="ij")
    X = origin[0] + (ic + 0.5) * cell[0]
    Y = origin[1] + (jc + 0.5) * cell[1]
    Z = origin[2] + (kc + 0.5) * cell[2]
    nppc = _profile(sp.particles_per_cell, X, Y, Z).astype(int)
    charge = _profile(sp.charge, X, Y, Z)
    if sp.charge_density is not None:
        dens = _profile(sp.charge_density, X, Y, Z)
        with np.errstate(divide="ignore", invalid="ignore"):
            dens = np.where(charge != 0, np.abs(dens / charge), 0.)
    else:
        dens = np.abs(_profile(sp.number_density, X, Y, Z))
    dens = dens * params.cell_volume
    keep = (dens > 0) & (nppc > 0)
    cells = np.flatnonzero(keep.ravel())
    npc = nppc.ravel()[cells]
    if len(cells) == 0:
        return origin, cells.astype(np.int32), np.zeros(0), np.zeros(0, dtype=np.int16), [1, 1, 1], [1., 1., 1.]
    if npc.min() != npc.max():
        return None
    c, inv = _regular_counts(int(npc[0]), sp.regular_number)
    return origin, cells.astype(np.int32), dens.ravel()[cells] / npc, charge.ravel()[cells].astype(np.int16), c, inv


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
            c, inv = _regular_counts(int(u), sp.regular_number)
            sel = np.repeat(npc == u, npc)
            i = local[sel].copy()
            for d in range(3):
                pos[d, sel] = origin[d] + ci[d, sel] * cell[d] + cell[d] * 0.975 * inv[d] * (0.5 + i % c[d])
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
