"""The time loop of the hot path: one patch per GPU, neighbours exchanged through
smilei_b200.exchange.  Mirrors the call order of the reference driver,

    main loop                     src/Smilei.cpp:483-749
    VectorPatch::dynamics         src/Patch/VectorPatch.cpp:325-374, 4763-4834
    VectorPatch::sumDensities     :819
    VectorPatch::solveMaxwell     :965-1085
    finalizeExchParticlesAndSort  :435,  SyncVectorPatch.cpp:44-60
    finalizeSyncAndBCFields       :1144-1180

with the reference's vocabulary (Species, ElectroMagn, patch, smpi).  Everything numerical runs
in the CUDA library behind the C ABI; this file only sequences calls.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import particles_init
from .exchange import Exchanger, J_FIELDS, rank_to_pcoord
from .operators import InterpolatorFactory, PusherFactory, ProjectorFactory, SolverFactory


class Particles:
    """Handle on the device SoA of one species (the reference's Particles, src/Particles/Particles.h)."""

    def __init__(self, patch, ispec):
        self.patch = patch
        self.ispec = ispec
        self._stage = None

    def numberOfParticles(self):
        return self.patch.species_count(self.ispec)


class Species:
    """src/Species/Species.h — owns Particles and the three operators (Species::initOperators, Species.cpp:260-307)."""

    def __init__(self, params, sparams, patch, ispec):
        self.params = params
        self.name = sparams.name
        self.mass_ = sparams.mass
        self.pusher = sparams.pusher
        self.sparams = sparams
        self.patch = patch
        self.ispec = ispec
        self.particles = Particles(patch, ispec)
        self.Interp = InterpolatorFactory.create(params, patch)
        self.Push = PusherFactory.create(params, self)
        self.Proj = ProjectorFactory.create(params, patch)

    def dynamics(self, EMfields, smpi, diag_flag=False):
        """Species::dynamics (Species.cpp:524-875): interpolate, push, boundary conditions, project."""
        n = self.particles.numberOfParticles()
        istart, iend = 0, n
        self.Interp.fieldsWrapper(EMfields, self.particles, smpi, istart, iend, 0)          # Species.cpp:591
        self.Push(self.particles, smpi, 0, n, 0)                                             # Species.cpp:727
        # partBoundCond->apply (Species.cpp:757) is fused in the kernel
        self.Proj.currentsAndDensityWrapper(EMfields, self.particles, smpi, istart, iend, 0, diag_flag, False,
                                            self.ispec)                                      # Species.cpp:782


class ElectroMagn:
    """src/ElectroMagn/ElectroMagn.h — field container of the patch + the two Maxwell solvers."""

    def __init__(self, params, patch):
        self.patch = patch
        self.MaxwellAmpereSolver_ = SolverFactory.createMA(params, patch)        # ElectroMagn.cpp:54
        self.MaxwellFaradaySolver_ = SolverFactory.createMF(params, patch)       # ElectroMagn.cpp:55

    def restartRhoJ(self):
        self.patch.restart_rhoJ()

    def allocateSpeciesFields(self, ispec, Jx=True, Jy=True, Jz=True, rho=True):
        """ElectroMagn::Jx_s[ispec] .. rho_s[ispec] (ElectroMagn.h:129-141): the arrays a DiagFields naming that
        species needs; diag-step deposits of the species then go there (Projector3D2Order.cpp:756-759)."""
        self.patch.species_diag_fields(ispec, Jx, Jy, Jz, rho)
        self.species_fields = getattr(self, "species_fields", {})
        self.species_fields[ispec] = tuple(n for n, on in zip(("Jx", "Jy", "Jz", "rho"), (Jx, Jy, Jz, rho)) if on)

    def computeTotalRhoJ(self):
        """ElectroMagn3D::computeTotalRhoJ (ElectroMagn3D.cpp:1753-1799)."""
        self.patch.compute_total_rhoJ()

    def centerMagneticFields(self):
        self.patch.center_B()


class _Smpi:
    keep_scratch = False


class SimWindow:
    """src/MovWindow/SimWindow.{h,cpp}: the moving window along x.  The reference slides by one patch
    (params.patch_size_[0] cells) whenever (time_dual - time_start)*velocity_x exceeds the distance already
    moved; here the GPU holds one patch spanning the box along x and the same stride — the patch size of the
    NAMELIST's number_of_patches — is applied as a shift of that patch by that many cells."""

    def __init__(self, params):
        w = params.window
        self.active = w is not None
        self.time_start = float(w.time_start) if self.active else float("inf")
        self.velocity_x = float(w.velocity_x) if self.active else 1.
        self.cell_length_x_ = params.cell_length[0]
        self.n_space_x_ = params.global_size[0] // params.number_of_patches[0]      # params.patch_size_[0]
        self.x_moved = 0.
        self.n_moved = 0

    def isMoving(self, time_dual):                                                    # SimWindow.cpp:93-96
        return self.active and (time_dual - self.time_start) * self.velocity_x > self.x_moved


class Simulation:
    """One rank's share of the run.  `rank_grid` = ranks per dimension (the GPU grid); the global box
    of the namelist is split evenly, one patch per rank, periodic."""

    def __init__(self, params, rank_grid=(1, 1, 1), rank=0, patch_factory=None, device=None, group=None,
                 capacity_factor=1.25, overlap_exchange=True):
        params.check_hot_path()
        self.params = params
        self.rank_grid = tuple(int(v) for v in rank_grid)
        self.world = int(np.prod(self.rank_grid))
        self.rank = rank
        self.group = group
        self.pcoord = rank_to_pcoord(rank, self.rank_grid)
        for d in range(3):
            if params.global_size[d] % self.rank_grid[d]:
                raise ValueError(f"global size {params.global_size[d]} not divisible by {self.rank_grid[d]} ranks along dim {d}")
        self.n = tuple(params.global_size[d] // self.rank_grid[d] for d in range(3))
        self.capacity_factor = capacity_factor
        if patch_factory is None:
            from .capi import Patch
            dev_index = torch.cuda.current_device() if device is None else torch.device(device).index or 0
            self.device = torch.device("cuda", dev_index)

            def patch_factory(**kw):
                return Patch(device=dev_index, **kw)
        else:
            self.device = torch.device("cpu") if device is None else torch.device(device)
        self.patch = patch_factory(n=self.n, cell_length=tuple(params.cell_length), dt=params.timestep,
                                   interp_order=params.interpolation_order, n_species=len(params.species),
                                   pcoord=self.pcoord, npatch=self.rank_grid, oversize=tuple(params.oversize))
        self.overlap_exchange = overlap_exchange
        if self.device.type == "cuda" and hasattr(self.patch, "bind_stream"):
            # the library's kernels and torch.distributed's collectives share ONE stream: torch's current one
            self.patch.bind_stream(torch.cuda.current_stream(self.device).cuda_stream)
            self._side_stream = torch.cuda.Stream(self.device)      # the particle exchange overlaps the field part on it
        self.EMfields = ElectroMagn(params, self.patch)
        self.vecSpecies = [Species(params, sp, self.patch, i) for i, sp in enumerate(params.species)]
        self.smpi = _Smpi()
        # box sides: periodic or open (Silver-Mueller fields, `remove` particles)
        self.periodic = tuple(params.EM_BCs[d][0] == "periodic" for d in range(3))
        self.exchanger = Exchanger(self.patch, self.n, params.oversize, params.cell_length, self.rank_grid,
                                   self.pcoord, self.device, group, periodic=self.periodic)
        # Patch::isBoundary( axis, side ): no neighbour there
        self.is_boundary = [[(not self.periodic[d]) and self.pcoord[d] == (0 if s == 0 else self.rank_grid[d] - 1)
                             for s in range(2)] for d in range(3)]
        # lasers enter through the Silver-Mueller faces this patch holds (ElectroMagnBC::vecLaser, LaserFactory)
        from .laser import Laser
        self.lasers = []
        min_local = [self.pcoord[d] * (self.n[d] * params.cell_length[d]) for d in range(3)]
        for block in getattr(params, "laser_blocks", []):
            L = Laser(block, params)
            if self.is_boundary[L.i_boundary_ // 2][L.i_boundary_ % 2]:
                L.init_fields(self.n, params.oversize, params.cell_length, min_local)
                self.lasers.append(L)
        self.simWindow = SimWindow(params)
        self.itime = 0

    # ------------------------------------------------------------------ initial state
    def create_particles(self, seed=None, reference_streams=False):
        """ParticleCreator for the species of the namelist (host-side, init time only).  With
        `reference_streams` the particles are those the reference creates for this namelist and random_seed
        (per-patch xorshift32 streams, particles_init.create_reference_streams)."""
        if reference_streams:
            created = particles_init.create_reference_streams(self.params, [sp.sparams for sp in self.vecSpecies],
                                                              self.n, self.pcoord)
            for sp in self.vecSpecies:
                self.set_particles(sp.ispec, **created[sp.name])
            return
        seed = self.params.random_seed if seed is None else seed
        created = {}
        for sp in self.vecSpecies:
            dev = particles_init.regular_cold_cells(self.params, sp.sparams, self.n, self.pcoord) \
                if hasattr(self.patch, "species_append_regular") else None
            if dev is not None:
                # regular positions, cold: created on the device (the same doubles as the host creator)
                origin, cells, weight, charge, c, inv = dev
                n = len(cells) * c[0] * c[1] * c[2]
                self.patch.species_config(sp.ispec, sp.mass_, sp.pusher, int(max(n * self.capacity_factor, n + 4096)))
                self._set_species_bc(sp.ispec)
                self.patch.species_append_regular(sp.ispec, origin, self.n, c, inv, cells, weight, charge)
                self.patch.sort(sp.ispec)
                continue
            src = created.get(sp.sparams.position_initialization)
            arrays = particles_init.create(self.params, sp.sparams, self.n, self.pcoord, seed, self.rank,
                                           positions=None if src is None else (src["x"], src["y"], src["z"]))
            created[sp.name] = arrays
            self.set_particles(sp.ispec, **arrays)

    def set_particles(self, ispec, x, y, z, px, py, pz, w, q, capacity=None):
        n = len(x)
        cap = int(max(n * self.capacity_factor, n + 4096)) if capacity is None else capacity
        self.patch.species_config(ispec, self.vecSpecies[ispec].mass_, self.vecSpecies[ispec].pusher, cap)
        self._set_species_bc(ispec)
        self.patch.species_set(ispec, x, y, z, px, py, pz, w, q)
        self.patch.sort(ispec)                       # VectorPatch::initialParticleSorting (VectorPatch.cpp:300-320)

    def init_thermal(self, ppc, density=1.0, temperature=None, seed=0):
        """Synthetic uniform thermal plasma created on the device (bench / smoke)."""
        ncell = self.n[0] * self.n[1] * self.n[2]
        nppc = ppc[0] * ppc[1] * ppc[2]
        for sp in self.vecSpecies:
            T = sp.sparams.temperature[0] if temperature is None else temperature
            self.patch.species_config(sp.ispec, sp.mass_, sp.pusher, int(ncell * nppc * self.capacity_factor) + 4096)
            self._set_species_bc(sp.ispec)
            q = int(sp.sparams.charge)
            # same seed for every species: identical positions => rho = 0 at t = 0, no Poisson solve needed
            self.patch.species_init_thermal(sp.ispec, ppc, density, q, T, seed)
            self.patch.sort(sp.ispec)

    def _set_species_bc(self, ispec):
        bc = self.vecSpecies[ispec].sparams.boundary_conditions
        flat = [bc[d][s] for d in range(3) for s in range(2)]
        if any(b != "periodic" for b in flat):
            self.patch.species_set_bc(ispec, flat)

    def boundaryConditions(self, time_dual, sides=range(6)):
        """ElectroMagn::boundaryConditions (ElectroMagn.cpp:371-394): the six sides in order, Silver-Mueller where
        the box is open; the laser amplitudes of a side are summed on the host (ElectroMagnBC3D_SM.cpp:189-199).
        The x sides are skipped while the window moves (:374): SimWindow::shift applies xmin itself."""
        for ib in sides:
            axis0, side = ib // 2, ib % 2
            if self.periodic[axis0] or not self.is_boundary[axis0][side]:
                continue
            axis1, axis2 = (1 if axis0 == 0 else 0), (1 if axis0 == 2 else 2)
            isb = (self.is_boundary[axis1][0], self.is_boundary[axis1][1], self.is_boundary[axis2][0], self.is_boundary[axis2][1])
            db1 = db2 = None
            for L in self.lasers:
                if L.i_boundary_ == ib:
                    a0, a1 = L.amplitude(0, time_dual), L.amplitude(1, time_dual)
                    db1 = a0 if db1 is None else db1 + a0
                    db2 = a1 if db2 is None else db2 + a1
            self.patch.apply_SM(ib, self.params.EM_BCs_k[ib], isb, db1, db2)

    # ------------------------------------------------------------------ one time step
    def step(self, diag_flag=False):
        p = self.patch
        # ---- VectorPatch::dynamics
        self.EMfields.restartRhoJ()                                   # VectorPatch.cpp:4779
        for sp in self.vecSpecies:
            sp.dynamics(self.EMfields, self.smpi, diag_flag)          # VectorPatch.cpp:4821
        self.exchange_and_solve(diag_flag)
        # ---- importAndSortParticles (Smilei.cpp:637)
        for sp in self.vecSpecies:
            p.sort(sp.ispec)
        # ---- finalizeSyncAndBCFields (Smilei.cpp:649): boundary conditions on the open sides, then centre B
        time_dual = (self.itime + 1.5) * self.params.timestep                      # Smilei.cpp:172,488
        moving = self.simWindow.isMoving(time_dual)
        if not all(self.periodic):
            self.boundaryConditions(time_dual, sides=range(2, 6) if moving else range(6))
        self.EMfields.centerMagneticFields()
        # ---- moveWindow (Smilei.cpp:667)
        if moving:
            self.moveWindow(time_dual)
        self.itime += 1

    def exchange_and_solve(self, diag_flag=False, maxwell_events=None):
        """The middle of a time step: particle exchange, density sum, Maxwell solve, B exchange (Smilei.cpp:528-547,637).
        `maxwell_events`: a pair of CUDA events recorded around the two solver calls (bench.py)."""
        p = self.patch
        # ---- initExchParticles (Smilei.cpp:528) .. sumDensities (:531) .. solveMaxwell (:547) .. finalizeExchParticles
        #      (:637).  The reference starts the particle exchange, sums the densities and solves Maxwell while the
        #      particles travel, and completes the exchange before the sort.  Across ranks on the GPU the same overlap:
        #      the field part is enqueued on the main stream, the particle exchange (pack kernels, NCCL messages,
        #      unpack kernels: they touch no field array and none of the field scratch) runs on a side stream that
        #      only waits for the dynamics kernels, and the sort waits for both.
        overlap = self.overlap_exchange and self.world > 1 and self.device.type == "cuda" and hasattr(p, "bind_stream")
        if overlap:
            main = torch.cuda.current_stream(self.device)
            after_dynamics = torch.cuda.Event()
            after_dynamics.record(main)
        else:
            self.exchanger.exchange_particles(len(self.vecSpecies))
        # ---- sumDensities (Smilei.cpp:531; VectorPatch.cpp:905-963).  On a diag step the totals first receive the
        #      species' own arrays (computeTotalRhoJ), rho is summed with J (sumRhoJ), and so are the species' arrays
        #      (sumRhoJs) so that a field diagnostic reads complete values on the shared planes
        if diag_flag:
            self.EMfields.computeTotalRhoJ()
            self.exchanger.sum_J(fields=J_FIELDS + ("rho",))
            for ispec, names in getattr(self.EMfields, "species_fields", {}).items():
                self.exchanger.sum_J(fields=tuple((n, ispec) for n in names))
        else:
            self.exchanger.sum_J()
        # ---- solveMaxwell (Smilei.cpp:547): saveMagneticFields, Ampere, Faraday, exchangeB
        if maxwell_events is not None:
            maxwell_events[0].record()
        self.EMfields.MaxwellAmpereSolver_(self.EMfields)
        self.EMfields.MaxwellFaradaySolver_(self.EMfields)
        if maxwell_events is not None:
            maxwell_events[1].record()
        self.exchanger.exchange_B()
        if overlap:
            with torch.cuda.stream(self._side_stream):
                self._side_stream.wait_event(after_dynamics)
                p.bind_stream(self._side_stream.cuda_stream)
                try:
                    self.exchanger.exchange_particles(len(self.vecSpecies))
                finally:
                    p.bind_stream(main.cuda_stream)
                exchanged = torch.cuda.Event()
                exchanged.record(self._side_stream)
            main.wait_event(exchanged)

    def moveWindow(self, time_dual):
        """SimWindow::shift (SimWindow.cpp:98-550) for one patch spanning the box along x."""
        w = self.simWindow
        S = w.n_space_x_
        w.n_moved += S                                                               # :136
        self.lasers = []                                                             # laserDisabled, :144
        # with several ranks along x every patch hands its leftmost interior planes and the particles it leaves
        # behind to its -x neighbour (the reference sends whole patches to the left, :166-222)
        incoming = self.exchanger.window_fields_begin(S)
        self.patch.window_shift(S)
        self.exchanger.window_fields_end(S, incoming)
        self.exchanger.exchange_window_particles(len(self.vecSpecies))
        # particles of the cells uncovered at the right end of the box (ParticleCreator over the new patch, :372-392)
        created = {}
        for sp in self.vecSpecies:
            if self.pcoord[0] == self.rank_grid[0] - 1:
                first_new = self.params.global_size[0] + w.n_moved - S
                box = (S, self.n[1], self.n[2])
                origin = (first_new, self.pcoord[1] * self.n[1], self.pcoord[2] * self.n[2])
                dev = particles_init.regular_cold_cells(self.params, sp.sparams, box, self.pcoord, origin_cells=origin) \
                    if hasattr(self.patch, "species_append_regular") else None
                if dev is not None:
                    o_, cells, weight, charge, c, inv = dev
                    self.patch.species_append_regular(sp.ispec, o_, box, c, inv, cells, weight, charge)
                    self.patch.sort(sp.ispec)
                    continue
                src = created.get(sp.sparams.position_initialization)
                arrays = particles_init.create(self.params, sp.sparams, box, self.pcoord, self.params.random_seed + w.n_moved,
                                               self.rank, positions=None if src is None else (src["x"], src["y"], src["z"]),
                                               origin_cells=origin)
                created[sp.name] = arrays
                self.patch.species_append(sp.ispec, **arrays)
            self.patch.sort(sp.ispec)
        # xmin boundary condition now that the patch has moved (:420-430), lasers off
        if not self.periodic[0]:
            self.boundaryConditions(time_dual, sides=[0])
        w.x_moved += w.cell_length_x_ * S                                            # :549

    def run(self, n_steps, scalars_every=None):
        out = []
        for _ in range(n_steps):
            self.step()
            if scalars_every and self.itime % scalars_every == 0:
                out.append((self.itime,) + self.scalars())
        return out

    # ------------------------------------------------------------------ diagnostics (parity observable)
    def scalars(self):
        """DiagnosticScalar Ukin (per species), Uelm, summed over ranks (SmileiMPI::computeGlobalDiags)."""
        ukin, uelm = self.patch.energy()
        v = torch.tensor(list(ukin) + [uelm], dtype=torch.float64, device=self.device)
        if self.world > 1:
            dist.all_reduce(v, group=self.group)
        v = v.cpu().numpy()
        return v[:-1].copy(), float(v[-1])

    def n_particles(self):
        c = torch.tensor([float(self.patch.species_count(s.ispec)) for s in self.vecSpecies], dtype=torch.float64,
                         device=self.device)
        if self.world > 1:
            dist.all_reduce(c, group=self.group)
        return [int(x) for x in c.cpu().numpy()]

    def close(self):
        self.patch.close()
