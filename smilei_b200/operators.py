"""Host-side mirror of the reference's operator surface for the hot path.

Same names, same factories, same call order as the reference:

    InterpolatorFactory::create   src/Interpolator/InterpolatorFactory.h:44
    PusherFactory::create         src/Pusher/PusherFactory.h:41-129
    ProjectorFactory::create      src/Projector/ProjectorFactory.h:72-89
    SolverFactory::createMA / MF  src/ElectroMagnSolver/SolverFactory.h:55,128
    Species::dynamics             src/Species/Species.cpp:524-875  (fieldsWrapper -> Push -> BC -> currentsAndDensityWrapper)

The three particle operators are ONE fused kernel on the device (sb200_dynamics).  As in the
C++ adapter (include/smilei_b200_operators.hpp), `fieldsWrapper` and `Pusher.__call__` only
record that they were requested; `currentsAndDensityWrapper` launches the fused kernel and
checks that the two earlier stages were requested for the same species in the same step —
legal because nothing reads the inter-operator scratch in between when ionization /
radiation are off (Species.cpp:596-676), and those are rejected by Params.check_hot_path().
Unsupported variants raise: there is no fallback path.
"""
from .capi import SmileiB200Error, DYN_KEEP_SCRATCH, DYN_DIAG_RHO


class _Operator:
    def __init__(self, params, patch):
        self.params = params
        self.patch = patch


class Interpolator3D(_Operator):
    order = None

    def fieldsWrapper(self, EMfields, particles, smpi, istart, iend, ithread, scell=0, ipart_ref=0):
        """Interpolator::fieldsWrapper (src/Interpolator/Interpolator.h:23).  Deferred: see module doc."""
        particles._stage = {"interp": self.order}


class Interpolator3D2Order(Interpolator3D):
    """src/Interpolator/Interpolator3D2Order.{h,cpp}"""
    order = 2


class Interpolator3D4Order(Interpolator3D):
    """src/Interpolator/Interpolator3D4Order.{h,cpp}"""
    order = 4


class InterpolatorFactory:
    @staticmethod
    def create(params, patch, vectorization=False):
        if params.geometry != "3Dcartesian":
            raise SmileiB200Error(f"InterpolatorFactory: geometry {params.geometry} is not on the B200 hot path")
        if params.interpolation_order == 2:
            return Interpolator3D2Order(params, patch)
        if params.interpolation_order == 4:
            return Interpolator3D4Order(params, patch)
        raise SmileiB200Error(f"InterpolatorFactory: unknown interpolation_order {params.interpolation_order}")


class Pusher(_Operator):
    name = None

    def __init__(self, params, species):
        super().__init__(params, species.patch)
        self.species = species

    def __call__(self, particles, smpi, istart, iend, ithread, ipart_buffer_offset=0):
        """Pusher::operator() (src/Pusher/Pusher.h:22).  Deferred: see module doc."""
        if getattr(particles, "_stage", None) is None or "interp" not in particles._stage:
            raise SmileiB200Error("Pusher called before Interpolator::fieldsWrapper for this species")
        particles._stage["push"] = self.name


class PusherBoris(Pusher):
    """src/Pusher/PusherBoris.cpp:23-127"""
    name = "boris"


class PusherVay(Pusher):
    """src/Pusher/PusherVay.cpp:32-172"""
    name = "vay"


class PusherHigueraCary(Pusher):
    """src/Pusher/PusherHigueraCary.cpp:32-166"""
    name = "higueracary"


class PusherFactory:
    _table = {"boris": PusherBoris, "vay": PusherVay, "higueracary": PusherHigueraCary}

    @staticmethod
    def create(params, species):
        # PusherFactory.h:49-70; borisnr, ponderomotive_boris, borisBTIS3, norm (photons) are not on this path
        try:
            return PusherFactory._table[species.pusher](params, species)
        except KeyError:
            raise SmileiB200Error(f"PusherFactory: pusher `{species.pusher}` is not on the B200 hot path "
                                  "(boris, vay, higueracary)") from None


class Projector3D(_Operator):
    order = None

    def currentsAndDensityWrapper(self, EMfields, particles, smpi, istart, iend, ithread, diag_flag, is_spectral,
                                  ispec, icell=0, ipart_ref=0):
        """Projector::currentsAndDensityWrapper (src/Projector/Projector.h:44): launches the fused
        gather+push+BC+deposit kernel for species `ispec`."""
        st = getattr(particles, "_stage", None) or {}
        if st.get("interp") != self.order or "push" not in st:
            raise SmileiB200Error("Projector called without the interpolator and pusher stages of the same step")
        if is_spectral:
            raise SmileiB200Error("spectral solvers are not on the B200 hot path")
        flags = (DYN_DIAG_RHO if diag_flag else 0) | (DYN_KEEP_SCRATCH if getattr(smpi, "keep_scratch", False) else 0)
        self.patch.dynamics(ispec, flags)
        particles._stage = None


class Projector3D2Order(Projector3D):
    """src/Projector/Projector3D2Order.cpp"""
    order = 2


class Projector3D4Order(Projector3D):
    """src/Projector/Projector3D4Order.cpp"""
    order = 4


class ProjectorFactory:
    @staticmethod
    def create(params, patch, vectorization=False):
        if params.interpolation_order == 2:
            return Projector3D2Order(params, patch)
        if params.interpolation_order == 4:
            return Projector3D4Order(params, patch)
        raise SmileiB200Error(f"ProjectorFactory: unknown interpolation_order {params.interpolation_order}")


class MA_Solver3D_norm(_Operator):
    """src/ElectroMagnSolver/MA_Solver3D_norm.cpp.  The Ampère and Faraday sweeps are launched
    together by the Faraday operator (sb200_maxwell runs E then B); calling MA first and MF
    second, as VectorPatch::solveMaxwell does (VectorPatch.cpp:1017,1023), is required."""

    def __call__(self, EMfields):
        EMfields._ampere_requested = True


class MF_Solver3D_Yee(_Operator):
    """src/ElectroMagnSolver/MF_Solver3D_Yee.cpp"""

    def __call__(self, EMfields):
        if not getattr(EMfields, "_ampere_requested", False):
            raise SmileiB200Error("MF_Solver3D_Yee called before MA_Solver3D_norm in this step")
        EMfields._ampere_requested = False
        self.patch.maxwell()


class SolverFactory:
    @staticmethod
    def createMA(params, patch):
        if params.geometry == "3Dcartesian":
            return MA_Solver3D_norm(params, patch)                   # SolverFactory.h:72-76
        raise SmileiB200Error("SolverFactory.createMA: only 3Dcartesian is on the B200 hot path")

    @staticmethod
    def createMF(params, patch):
        if params.geometry == "3Dcartesian" and str(params.main.maxwell_solver) == "Yee":
            return MF_Solver3D_Yee(params, patch)                    # SolverFactory.h:160-164
        raise SmileiB200Error("SolverFactory.createMF: only the 3D Yee solver is on the B200 hot path")
