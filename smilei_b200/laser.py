"""Lasers injected through a Silver-Mueller face: the host side of `Laser` (src/ElectroMagn/Laser.{h,cpp}).

The reference keeps, per laser, two `LaserProfileSeparable` objects (first and second tangential B
component of the face).  Each pre-computes its space envelope and phase on the face grid of the patch
(LaserProfileSeparable::createFields / initFields, Laser.cpp:340-450) and returns, every step,
    time_envelope( t - (phase + delay_phase)/omega ) * space_envelope * sin( omega*t - phase ),
omega = omega0 * chirp(t)                                  (LaserProfileSeparable::getAmplitude, :453-460)
which ElectroMagnBC3D_SM::apply sums into its b1 / b2 arrays (ElectroMagnBC3D_SM.cpp:189-199, 265-275).
Here the two arrays are evaluated with numpy on the host each step (a face holds ~n^2 points) and handed
to sb200_apply_SM, which applies the boundary condition on the device.

Only what the 3D Cartesian path needs is mirrored: separable profiles and the `LaserGaussian3D` helper
(src/Python/pyprofiles.py:600-666).  Space-time and file profiles raise.
"""
import math

import numpy as np

BOX_SIDES = {"xmin": 0, "xmax": 1, "ymin": 2, "ymax": 3, "zmin": 4, "zmax": 5}


def _call(f, *a):
    """Evaluate a namelist profile on arrays: numpy-aware profiles directly, scalar Python callables point-wise."""
    if not callable(f):
        return np.full(np.broadcast(*a).shape, float(f))
    if getattr(f, "vectorized", False):
        return np.asarray(f(*a), dtype=np.float64)
    return np.vectorize(f, otypes=[np.float64])(*a)


class Laser:
    def __init__(self, block, params):
        side = getattr(block, "box_side", "xmin")
        if side not in BOX_SIDES:
            raise ValueError("Laser: box_side must be xmin, xmax, ymin, ymax, zmin or zmax")           # Laser.cpp:24-46
        self.i_boundary_ = BOX_SIDES[side]
        if getattr(block, "file", None) is not None:
            raise ValueError("Laser: profiles read from a file (LaserOffset) are not on the B200 path")
        # space_time_profile = [By(y,z,t), Bz(y,z,t)]: LaserProfileNonSeparable (Laser.h:130-155), None = zero component
        self.space_time = getattr(block, "space_time_profile", None)
        self.pos = [None, None]
        self.omega = float(getattr(block, "omega", 1.))
        self.chirp = getattr(block, "chirp_profile", 1.)
        self.time = getattr(block, "time_envelope", 1.)
        self.space = list(getattr(block, "space_envelope", [1., 0.]))
        self.phase = list(getattr(block, "phase", [0., 0.]))
        self.delay = [float(v) for v in getattr(block, "delay_phase", [0., 0.])]
        self.env = [None, None]
        self.phi = [None, None]

    def init_fields(self, n, oversize, cell_length, min_local):
        """LaserProfileSeparable::initFields, 3Dcartesian branch (Laser.cpp:424-450): profile 0 sits on the
        (primal, dual) points of the face, profile 1 on the (dual, primal) ones."""
        axis = self.i_boundary_ // 2
        ax1 = 1 if axis == 0 else 0
        ax2 = 1 if axis == 2 else 2
        n1p, n2p = n[ax1] + 1 + 2 * oversize[ax1], n[ax2] + 1 + 2 * oversize[ax2]
        d1, d2 = cell_length[ax1], cell_length[ax2]
        for comp, primal in ((0, True), (1, False)):
            dim1 = n1p if primal else n1p + 1
            dim2 = n2p + 1 if primal else n2p
            p1 = np.empty(dim1)
            p2 = np.empty(dim2)
            v = min_local[ax1] - ((0. if primal else 0.5) + oversize[ax1]) * d1
            for j in range(dim1):                    # the reference accumulates pos += d
                p1[j] = v
                v += d1
            v = min_local[ax2] - ((0.5 if primal else 0.) + oversize[ax2]) * d2
            for k in range(dim2):
                p2[k] = v
                v += d2
            Y, Z = np.meshgrid(p1, p2, indexing="ij")
            if self.space_time is not None:
                # the non-separable profile is evaluated at the positions ElectroMagnBC3D_SM::apply computes
                # (ElectroMagnBC3D_SM.cpp:191-195, 267-271): min + (j - oversize)*d, minus half a cell on the dual axis
                q1 = min_local[ax1] + (np.arange(dim1) - (0. if primal else 0.5) - oversize[ax1]) * d1
                q2 = min_local[ax2] + (np.arange(dim2) - (0.5 if primal else 0.) - oversize[ax2]) * d2
                self.pos[comp] = np.meshgrid(q1, q2, indexing="ij")
                continue
            self.env[comp] = np.ascontiguousarray(_call(self.space[comp], Y, Z))
            self.phi[comp] = np.ascontiguousarray(_call(self.phase[comp], Y, Z))

    def amplitude(self, comp, t):
        """Laser::getAmplitude0 / getAmplitude1 on the whole face at time t."""
        if self.space_time is not None:
            f = self.space_time[comp]
            Y, Z = self.pos[comp]
            if f is None:
                return np.zeros(Y.shape)
            return np.ascontiguousarray(_call(f, Y, Z, np.float64(t)))
        omega = self.omega * float(_call(self.chirp, np.float64(t)))
        phi = self.phi[comp]
        envt = _call(self.time, t - (phi + self.delay[comp]) / omega)
        return envt * self.env[comp] * np.sin(omega * t - phi)


# ---------------------------------------------------------------------------------------------------------
# time profiles of src/Python/pyprofiles.py that laser namelists use, numpy-aware
# ---------------------------------------------------------------------------------------------------------

def _vec(f):
    f.vectorized = True
    return f


def tconstant(start=0.):
    return _vec(lambda t: np.where(np.asarray(t) >= start, 1., 0.))


def ttrapezoidal(simulation_time, start=0., plateau=None, slope1=0., slope2=0.):
    if plateau is None:
        plateau = simulation_time - start

    def f(t):
        t = np.asarray(t, dtype=np.float64)
        r = np.zeros_like(t)
        up = (t >= start) & (t < start + slope1)
        if slope1 > 0:
            r = np.where(up, (t - start) / slope1, r)
        r = np.where((t >= start + slope1) & (t < start + slope1 + plateau), 1., r)
        dn = (t >= start + slope1 + plateau) & (t < start + slope1 + plateau + slope2)
        if slope2 > 0:
            r = np.where(dn, 1. - (t - (start + slope1 + plateau)) / slope2, r)
        return r
    return _vec(f)


def tgaussian(simulation_time, start=0., duration=None, fwhm=None, center=None, order=2):
    if duration is None:
        duration = simulation_time - start
    if fwhm is None:
        fwhm = duration / 3.
    if center is None:
        center = start + duration / 2.
    sigma = (0.5 * fwhm) ** order / math.log(2.0)

    def f(t):
        t = np.asarray(t, dtype=np.float64)
        inside = (t >= start) & (t < start + duration)
        return np.where(inside, np.exp(-(np.where(inside, t, center) - center) ** order / sigma), 0.)
    return _vec(f)


def tsin2plateau(simulation_time, start=0., fwhm=0., plateau=None, slope1=None, slope2=None):
    if plateau is None:
        plateau = 0.
    if slope1 is None:
        slope1 = fwhm
    if slope2 is None:
        slope2 = slope1

    def f(t):
        t = np.asarray(t, dtype=np.float64)
        r = np.zeros_like(t)
        if slope1 > 0:
            r = np.where((t >= start) & (t < start + slope1), np.sin(0.5 * math.pi * (t - start) / slope1) ** 2, r)
        r = np.where((t >= start + slope1) & (t < start + slope1 + plateau), 1., r)
        if slope2 > 0:
            r = np.where((t >= start + slope1 + plateau) & (t < start + slope1 + plateau + slope2),
                         np.cos(0.5 * math.pi * (t - start - slope1 - plateau) / slope2) ** 2, r)
        return r
    return _vec(f)


def polarization(polarization_phi, ellipticity):
    """transformPolarization (pyprofiles.py:458-468): [dephasing, amplitude on the first axis, on the second]."""
    e2 = ellipticity ** 2
    p = (1. - e2) * math.sin(2. * polarization_phi) / 2.
    dephasing = math.atan2(ellipticity, p)
    amplitude = math.sqrt(1. / (1. + e2))
    c2 = math.cos(polarization_phi) ** 2
    s2 = 1. - c2
    return dephasing, amplitude * math.sqrt(c2 + e2 * s2), amplitude * math.sqrt(s2 + e2 * c2)


def gaussian3d_block(make_laser, grid_length, box_side="xmin", a0=1., omega=1., focus=None, waist=3.,
                     incidence_angle=(0., 0.), polarization_phi=0., ellipticity=0., time_envelope=None,
                     phase_offset=0.):
    """LaserGaussian3D (pyprofiles.py:600-666): a Gaussian beam focused at `focus`, entering through `box_side`."""
    assert focus is not None and len(focus) == 3, "LaserGaussian3D: focus must be a list of length 3."
    dephasing, ampZ, ampY = polarization(polarization_phi, ellipticity)
    ampY *= a0 * omega
    ampZ *= a0 * omega
    focus = list(focus)
    gl = list(grid_length)
    if box_side[0] == "y":
        focus = [focus[1], focus[0], focus[2]]
        gl = [gl[1], gl[0], gl[2]]
        ampY = -ampY
    elif box_side[0] == "z":
        focus = [focus[2], focus[0], focus[1]]
        gl = [gl[2], gl[0], gl[1]]
    if box_side.endswith("max"):
        focus[0] = gl[0] - focus[0]
    Zr = omega * waist ** 2 / 2.
    if list(incidence_angle) == [0., 0.]:
        w = math.sqrt(1. / (1. + (focus[0] / Zr) ** 2))
        invWaist2 = (w / waist) ** 2
        coeff = -omega * focus[0] * w ** 2 / (2. * Zr ** 2)

        def spatial(y, z):
            return w * np.exp(-invWaist2 * ((y - focus[1]) ** 2 + (z - focus[2]) ** 2))

        def phase(y, z):
            return coeff * ((y - focus[1]) ** 2 + (z - focus[2]) ** 2)
    else:
        invZr, invW, alpha = 1. / Zr, 1. / waist, omega * Zr
        cy, sy = math.cos(incidence_angle[0]), math.sin(incidence_angle[0])
        cz, sz = math.cos(incidence_angle[1]), math.sin(incidence_angle[1])
        cycz, cysz, sycz, sysz = cy * cz, cy * sz, sy * cz, sy * sz
        ampZ = sysz * ampY + cy * ampZ
        ampY *= cz

        def _xyz(y, z, s):
            X = invZr * (-focus[0] * cycz + (y - focus[1]) * cysz - (z - focus[2]) * sy)
            Y = s * (focus[0] * sz + (y - focus[1]) * cz)
            Z = s * (-focus[0] * sycz + (y - focus[1]) * sysz + (z - focus[2]) * cy)
            return X, Y, Z

        def spatial(y, z):
            X, Y, Z = _xyz(y, z, invW)
            invW2 = 1. / (1. + X ** 2)
            return np.sqrt(invW2) * np.exp(-(Y ** 2 + Z ** 2) * invW2)

        def phase(y, z):
            X, Y, Z = _xyz(y, z, invZr)
            return alpha * X * (1. + 0.5 * (Y ** 2 + Z ** 2) / (1. + X ** 2)) - np.arctan(X)
        faces = (focus[0], focus[1], focus[2], focus[1] - gl[1], focus[2] - gl[2])
        denominators = (cycz, cysz, -sy, cysz, -sy)
        dist = min(N / D for N, D in zip(faces, denominators) if D != 0 and N / D > 0)
        phase_offset -= omega * dist - math.atan(dist / Zr)
    return make_laser(
        box_side=box_side, omega=omega, chirp_profile=tconstant(), time_envelope=time_envelope,
        space_envelope=[_vec(lambda y, z: ampY * spatial(y, z)), _vec(lambda y, z: ampZ * spatial(y, z))],
        phase=[_vec(lambda y, z: phase(y, z) - phase_offset + dephasing), _vec(lambda y, z: phase(y, z) - phase_offset)],
        delay_phase=[0., dephasing])
