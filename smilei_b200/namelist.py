"""Smilei's Python namelist, for the blocks the hot path reads.

A namelist is an ordinary Python file that instantiates `Main(...)`, `Species(...)`, ...
(reference: src/Python/pyinit.py:156-654, executed by Params::Params, src/Params/Params.cpp:126-187).
This module provides those block classes with the reference's keyword names and defaults for
the 3D Cartesian / Yee / periodic path, executes an unmodified namelist and derives the
quantities Params::compute derives (src/Params/Params.cpp:1148-1260).  Blocks that concern
subsystems outside the hot path (diagnostics, checkpoints, load balancing, ...) are accepted
and recorded but drive nothing here.
"""
import math
import os

_OUT_OF_SCOPE_BLOCKS = (
    "LoadBalancing", "Checkpoints", "DiagFields", "DiagProbe", "DiagScreen", "DiagParticleBinning",
    "DiagTrackParticles", "DiagPerformances", "DiagRadiationSpectrum", "DiagNewParticles", "CurrentFilter",
    "FieldFilter", "MultipleDecomposition", "Collisions", "RadiationReaction", "MultiphotonBreitWheeler",
    "ParticleInjector", "ExternalField", "PrescribedField", "Antenna", "PartWall", "LaserEnvelope",
    "LaserPlanar1D", "LaserGaussian2D", "LaserOffset", "LaserGaussianAM", "LaserEnvelopePlanar1D",
    "LaserEnvelopeGaussian2D", "LaserEnvelopeGaussian3D", "LaserEnvelopeGaussianAM",
)


# spatial / temporal profile helpers of src/Python/pyprofiles.py: only `constant` is evaluated on the hot
# path; the others are accepted so that namelists parse, and raise if something tries to evaluate them
_PROFILE_HELPERS = ("trapezoidal", "gaussian", "polygonal", "cosine", "polynomial", "tpolygonal", "tcosine",
                    "tpolynomial")
# Laser block (src/Python/pyinit.py:455-467)
_LASER_DEFAULTS = dict(box_side="xmin", omega=1., chirp_profile=1., time_envelope=1., space_envelope=[1., 0.],
                       phase=[0., 0.], delay_phase=[0., 0.], space_time_profile=None, space_time_profile_AM=None,
                       file=None, _offset=None)


def _opaque_profile(name):
    def make(*args, **kwargs):
        def profile(*a):
            raise NamelistError(f"profile helper `{name}` is not evaluated by the B200 hot path")
        profile.helper = (name, args, kwargs)
        return profile
    return make


class NamelistError(Exception):
    """The reference's ERROR_NAMELIST (src/Tools/Tools.h)."""


class _Block:
    _defaults = {}

    def __init__(self, **kwargs):
        for k, v in self._defaults.items():
            setattr(self, k, v)
        for k, v in kwargs.items():
            if self._defaults and k not in self._defaults:
                raise NamelistError(f"{type(self).__name__}: unknown keyword `{k}`")
            setattr(self, k, v)
        if self._singleton:
            # singletons expose their values as class attributes: namelists read e.g. Main.grid_length
            # (SmileiSingleton, src/Python/pyinit.py:118-150)
            for k in list(self._defaults) + list(kwargs):
                setattr(type(self), k, getattr(self, k))
        type(self)._instances.append(self)


def _make_block(name, defaults, singleton=False):
    cls = type(name, (_Block,), {"_defaults": dict(defaults), "_instances": [], "_singleton": singleton})
    return cls


# defaults: src/Python/pyinit.py:156-230 (Main), :330-390 (Species), Vectorization, MovingWindow, DiagScalar
_MAIN_DEFAULTS = dict(
    geometry=None, cell_length=[], grid_length=[], number_of_cells=[], timestep=None, simulation_time=None,
    number_of_timesteps=None, interpolation_order=2, interpolator="momentum-conserving", custom_oversize=2,
    number_of_patches=None, patch_arrangement="hilbertian", cluster_width=-1, every_clean_particles_overhead=100,
    timestep_over_CFL=None, cell_sorting=None, gpu_computing=False, solve_poisson=True,
    poisson_max_iteration=50000, poisson_max_error=1.e-14, maxwell_solver="Yee",
    EM_boundary_conditions=[["periodic"]], EM_boundary_conditions_k=[], time_fields_frozen=0.,
    reference_angular_frequency_SI=0., print_every=None, random_seed=None, print_expected_disk_usage=True,
    terminal_mode=True, number_of_AM=2, use_BTIS3_interpolation=False, save_magnectic_fields_for_SM=True,
    number_of_pml_cells=[[10]], Laser_Envelope_model=False, solve_relativistic_poisson=False,
    spectral_solver_order=[], initial_rotational_cleaning=False,
)
_SPECIES_DEFAULTS = dict(
    name=None, position_initialization=None, regular_number=[], momentum_initialization="cold",
    particles_per_cell=None, c_part_max=1.0, mass=None, charge=None, charge_density=None, number_density=None,
    mean_velocity=[0., 0., 0.], mean_velocity_AM=None, temperature=[1e-10], thermal_boundary_temperature=[],
    thermal_boundary_velocity=[0., 0., 0.], pusher="boris", boundary_conditions=[["periodic"]],
    time_frozen=0.0, radiation_model="none", ionization_model="none", is_test=False,
    relativistic_field_initialization=False, keep_interpolated_fields=[], atomic_number=None,
)
_VECTO_DEFAULTS = dict(mode="off", reconfigure_every=20, initial_mode="on")
_MW_DEFAULTS = dict(time_start=0., velocity_x=1., number_of_additional_shifts=0, additional_shifts_time=0.)
_SCALAR_DEFAULTS = dict(every=None, precision=10, vars=[])


def _fresh_namespace():
    import numpy
    ns = {"math": math, "np": numpy, "numpy": numpy, "os": os}
    ns["Main"] = _make_block("Main", _MAIN_DEFAULTS, singleton=True)
    ns["Species"] = _make_block("Species", _SPECIES_DEFAULTS)
    ns["Vectorization"] = _make_block("Vectorization", _VECTO_DEFAULTS, singleton=True)
    ns["MovingWindow"] = _make_block("MovingWindow", _MW_DEFAULTS, singleton=True)
    ns["DiagScalar"] = _make_block("DiagScalar", _SCALAR_DEFAULTS)
    ns["Laser"] = _make_block("Laser", _LASER_DEFAULTS)
    for b in _OUT_OF_SCOPE_BLOCKS:
        ns[b] = _make_block(b, {})
    # time profiles and the Gaussian-beam helper of src/Python/pyprofiles.py, numpy-aware (smilei_b200/laser.py)
    from . import laser as _laser

    def _sim_time():
        M = ns["Main"]
        if not M._instances:
            raise NamelistError("time profile defined before `Main()`")
        m = M._instances[0]
        if m.simulation_time is not None:
            return float(m.simulation_time)
        return float(m.number_of_timesteps) * float(m.timestep) if m.timestep is not None else 0.

    ns["tconstant"] = _laser.tconstant
    ns["ttrapezoidal"] = lambda *a, **k: _laser.ttrapezoidal(_sim_time(), *a, **k)
    ns["tgaussian"] = lambda *a, **k: _laser.tgaussian(_sim_time(), *a, **k)
    ns["tsin2plateau"] = lambda *a, **k: _laser.tsin2plateau(_sim_time(), *a, **k)
    ns["transformPolarization"] = lambda phi, e: list(_laser.polarization(phi, e))

    def _gaussian3d(**kw):
        M = ns["Main"]
        if not M._instances:
            raise NamelistError("LaserGaussian3D profile has been defined before `Main()`")
        m = M._instances[0]
        gl = list(m.grid_length) if m.grid_length else [n * c for n, c in zip(m.number_of_cells, m.cell_length)]
        if kw.get("time_envelope") is None:
            kw["time_envelope"] = _laser.tconstant()
        return _laser.gaussian3d_block(lambda **b: ns["Laser"](**b), gl, **kw)
    ns["LaserGaussian3D"] = _gaussian3d
    # helpers namelists commonly use (src/Python/pyprofiles.py); only the constant profile matters here
    ns["constant"] = lambda v, **kw: (lambda *a: v)
    for helper in _PROFILE_HELPERS:
        ns[helper] = _opaque_profile(helper)
    return ns


class SpeciesParams:
    def __init__(self, block, ispec):
        self.name = block.name if block.name is not None else f"species{ispec}"
        self.mass = float(block.mass)
        self.charge = block.charge
        self.pusher = str(block.pusher)
        self.particles_per_cell = block.particles_per_cell
        self.position_initialization = block.position_initialization
        self.momentum_initialization = block.momentum_initialization
        self.regular_number = list(block.regular_number)
        self.temperature = list(block.temperature) if isinstance(block.temperature, (list, tuple)) else [block.temperature]
        self.mean_velocity = list(block.mean_velocity)
        self.charge_density = block.charge_density
        self.number_density = block.number_density
        # Species::boundary_conditions_: [[xmin,xmax],[ymin,ymax],[zmin,zmax]], short forms expanded (pyinit / Params)
        bcs = block.boundary_conditions
        full = [list(bcs[min(i, len(bcs) - 1)]) for i in range(3)]
        for bc in full:
            if len(bc) == 1:
                bc.append(bc[0])
        self.boundary_conditions = full
        self.block = block


class Params:
    """The subset of the reference's Params the hot path reads (src/Params/Params.h)."""

    def __init__(self, ns):
        mains = ns["Main"]._instances
        if len(mains) != 1:
            raise NamelistError("the namelist must contain exactly one Main() block")
        m = mains[0]
        self.main = m
        self.geometry = m.geometry
        if self.geometry != "3Dcartesian":
            raise NamelistError(f"geometry `{self.geometry}` is outside the B200 hot path (3Dcartesian only)")
        if str(m.maxwell_solver) != "Yee":
            raise NamelistError(f"maxwell_solver `{m.maxwell_solver}` is outside the B200 hot path (Yee only)")
        self.interpolation_order = int(m.interpolation_order)
        if self.interpolation_order not in (2, 4):
            raise NamelistError("interpolation_order must be 2 or 4")
        self.cell_length = [float(v) for v in m.cell_length]
        if len(self.cell_length) != 3:
            raise NamelistError("cell_length must have 3 entries")
        if m.grid_length:
            # Params.cpp:1179: patch_size_ = round(grid_length/cell_length) before the division by patches
            self.global_size = [int(round(float(g) / c)) for g, c in zip(m.grid_length, self.cell_length)]
        elif m.number_of_cells:
            self.global_size = [int(v) for v in m.number_of_cells]
        else:
            raise NamelistError("grid_length or number_of_cells must be defined")
        self.grid_length = [n * c for n, c in zip(self.global_size, self.cell_length)]
        # timestep (pyinit.py:233-262: timestep_over_CFL / sqrt(sum 1/dx^2) for Yee)
        if m.timestep is not None:
            self.timestep = float(m.timestep)
        elif m.timestep_over_CFL is not None:
            self.timestep = float(m.timestep_over_CFL) / math.sqrt(sum(1. / c ** 2 for c in self.cell_length))
        else:
            raise NamelistError("timestep and timestep_over_CFL not defined")
        if m.number_of_timesteps is not None:
            self.n_time = int(m.number_of_timesteps)
        elif m.simulation_time is not None:
            self.n_time = int(float(m.simulation_time) / self.timestep)        # Params.cpp:1154
        else:
            raise NamelistError("simulation_time or number_of_timesteps must be defined")
        self.simulation_time = self.n_time * self.timestep
        self.number_of_patches = [int(v) for v in (m.number_of_patches or [1, 1, 1])]
        self.patch_arrangement = str(m.patch_arrangement)
        bcs = m.EM_boundary_conditions
        self.EM_BCs = [list(bcs[min(i, len(bcs) - 1)]) for i in range(3)]
        for bc in self.EM_BCs:
            if len(bc) == 1:
                bc.append(bc[0])
        # Params.cpp:1202: oversize = max(interpolation_order, spectral_solver_order/2+1, custom_oversize)
        self.oversize = [max(self.interpolation_order, 1, int(m.custom_oversize))] * 3
        self.cell_volume = 1.0
        for c in self.cell_length:
            self.cell_volume *= c
        self.random_seed = 0 if m.random_seed is None else int(m.random_seed)       # Params.cpp:217
        self.gpu_computing = bool(m.gpu_computing)
        v = ns["Vectorization"]._instances
        self.vectorization_mode = v[0].mode if v else "off"
        self.has_window = len(ns["MovingWindow"]._instances) > 0
        self.window = ns["MovingWindow"]._instances[0] if self.has_window else None
        self.species = [SpeciesParams(b, i) for i, b in enumerate(ns["Species"]._instances)]
        sc = ns["DiagScalar"]._instances
        self.scalar_every = sc[0].every if sc else None
        self.ignored_blocks = {b: len(ns[b]._instances) for b in _OUT_OF_SCOPE_BLOCKS if ns[b]._instances}
        self.laser_blocks = list(ns["Laser"]._instances)
        # EM_boundary_conditions_k: incidence vectors of the Silver-Mueller sides, default = the inward normals, one
        # vector = the same on every side (Params.cpp:475-515)
        ks = list(m.EM_boundary_conditions_k)
        if len(ks) == 1:
            ks = ks * 6
        elif len(ks) not in (0, 6):
            raise NamelistError("EM_boundary_conditions_k must be the same size as the number of faces.")
        default_k = [[1., 0., 0.], [-1., 0., 0.], [0., 1., 0.], [0., -1., 0.], [0., 0., 1.], [0., 0., -1.]]
        self.EM_BCs_k = [[float(v) for v in ks[i]] if i < len(ks) else default_k[i] for i in range(6)]

    def check_hot_path(self):
        """Raise for anything the B200 path does not cover (no silent fallback)."""
        for d in range(3):
            bc = self.EM_BCs[d]
            if bc != ["periodic", "periodic"] and any(b != "silver-muller" for b in bc):
                raise NamelistError(f"EM_boundary_conditions {bc} along dim {d}: periodic and silver-muller are on the B200 hot path")
        if self.has_window and self.EM_BCs[0][0] == "periodic":
            raise NamelistError("MovingWindow with a periodic x direction is not supported")
        if self.has_window and int(getattr(self.window, "number_of_additional_shifts", 0) or 0) != 0:
            # SimWindow::isMoving / shift take extra shifts at additional_shifts_time (SimWindow.cpp:95,100): not built
            raise NamelistError("MovingWindow.number_of_additional_shifts != 0 is outside the B200 hot path")
        if self.has_window and self.number_of_patches[0] < 2:
            # the window slides by one patch of the NAMELIST (SimWindow.cpp:136): with a single patch along x the
            # stride would be the whole box
            raise NamelistError("MovingWindow needs number_of_patches[0] >= 2 (the window slides by one patch along x)")
        for s in self.species:
            if s.pusher not in ("boris", "vay", "higueracary"):
                raise NamelistError(f"pusher `{s.pusher}` is outside the B200 hot path (boris, vay, higueracary)")
            if s.mass <= 0:
                raise NamelistError("photon species are outside the B200 hot path")
            for d in range(3):
                em_periodic = self.EM_BCs[d][0] == "periodic"
                for side in range(2):
                    pbc = s.boundary_conditions[d][side]
                    if pbc not in ("periodic", "remove"):
                        raise NamelistError(f"particle boundary condition `{pbc}` is outside the B200 hot path (periodic, remove)")
                    if em_periodic != (pbc == "periodic"):         # PartBoundCond.cpp:76-80
                        raise NamelistError(f"species {s.name}: periodic EM boundaries along dim {d} go with periodic particle boundaries, and only with them")
        for L in self.laser_blocks:
            d = {"x": 0, "y": 1, "z": 2}[str(L.box_side)[0]]
            if self.EM_BCs[d][0 if str(L.box_side).endswith("min") else 1] != "silver-muller":
                raise NamelistError(f"Laser on {L.box_side} needs a silver-muller boundary there")
        for key in ("Collisions", "RadiationReaction", "MultiphotonBreitWheeler", "ParticleInjector", "LaserOffset",
                    "LaserEnvelope", "LaserEnvelopePlanar1D", "LaserEnvelopeGaussian2D", "LaserEnvelopeGaussian3D",
                    "LaserEnvelopeGaussianAM", "LaserPlanar1D", "LaserGaussian2D", "LaserGaussianAM", "ExternalField",
                    "PrescribedField", "Antenna", "PartWall"):
            if key in self.ignored_blocks:
                raise NamelistError(f"{key} blocks are outside the B200 hot path")


def load_namelist(path_or_source, is_source=False):
    """Execute a Smilei namelist and return Params."""
    ns = _fresh_namespace()
    if is_source:
        src, fname = path_or_source, "<namelist>"
    else:
        with open(path_or_source) as f:
            src = f.read()
        fname = os.path.abspath(path_or_source)
    exec(compile(src, fname, "exec"), ns)
    return Params(ns)
