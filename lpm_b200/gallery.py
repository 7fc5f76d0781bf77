"""Initial conditions used by the four benchmark configs (host-side, numpy).

Reference: src/lpm_vorticity_gallery.hpp (SolidBodyRotation :31-57, GaussianVortexSphere :59-102,
RossbyHaurwitz54 :104-147, SphereTestCase2Vorticity :264-278), src/lpm_surface_gallery.hpp:120-134
(SphereTestCase2InitialSurface), src/util/lpm_math.hpp (atan4), examples/sphere_swe_tc2.cpp:231-251.
"""
import numpy as np

PI = 3.1415926535897932384626433832795027975  # lpm_constants.hpp:11
ZERO_TOL = 2.220446049250313e-16


def atan4(y, x):
    """Longitude in [0, 2pi): atan4 (src/util/lpm_math.hpp:66-104), quadrant logic on atan2(|y|, |x|)."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    xz = np.abs(x) < ZERO_TOL
    yz = np.abs(y) < ZERO_TOL
    theta = np.arctan2(np.abs(y), np.abs(x))
    res = np.zeros_like(theta)
    gen = ~xz & ~yz
    res = np.where(gen & (x > 0) & (y > 0), theta, res)
    res = np.where(gen & (x < 0) & (y > 0), PI - theta, res)
    res = np.where(gen & (x < 0) & (y < 0), PI + theta, res)
    res = np.where(gen & (x > 0) & (y < 0), 2 * PI - theta, res)
    res = np.where(~xz & yz & (x < 0), PI, res)
    res = np.where(xz & (y > 0), 0.5 * PI, res)
    res = np.where(xz & (y < 0), 1.5 * PI, res)
    return res


class SolidBodyRotation:
    """zeta = 2*OMEGA*z, exact velocity OMEGA*(-y, x, 0) (lpm_vorticity_gallery.hpp:31-57)."""
    OMEGA = 2 * PI

    def __call__(self, xyz):
        return 2 * self.OMEGA * xyz[:, 2]

    def velocity(self, xyz):
        return np.stack([-self.OMEGA * xyz[:, 1], self.OMEGA * xyz[:, 0], np.zeros(len(xyz))], axis=1)

    def stream_fn(self, xyz):
        """psi = OMEGA*z (examples/bve_rotation.cpp:152-153: stream_fn - 2*pi*z)."""
        return self.OMEGA * xyz[:, 2]


class GaussianVortexSphere:
    """lpm_vorticity_gallery.hpp:59-102."""

    def __init__(self, strength=4 * PI, shape=4.0, init_lon=0.0, init_lat=PI / 20):
        self.gauss_const = 0.0
        self.vortex_strength = strength
        self.shape_parameter = shape
        self.xyz_ctr = np.array([np.cos(init_lon) * np.cos(init_lat), np.sin(init_lon) * np.cos(init_lat),
                                 np.sin(init_lat)])

    def set_gauss_const(self, vorticity_sum):
        self.gauss_const = vorticity_sum / (4 * PI)

    def __call__(self, xyz):
        distsq = 1.0 - xyz[:, 0] * self.xyz_ctr[0] - xyz[:, 1] * self.xyz_ctr[1] - xyz[:, 2] * self.xyz_ctr[2]
        return self.vortex_strength * np.exp(-self.shape_parameter ** 2 * distsq) - self.gauss_const


class RossbyHaurwitz54:
    """lpm_vorticity_gallery.hpp:104-147."""

    def __init__(self, u0=0.0, amp=1.0):
        self.u0 = u0
        self.amp = amp

    def set_stationary_wave_speed(self, Omega=2 * PI):
        self.u0 = Omega / 14

    def __call__(self, xyz):
        z = xyz[:, 2]
        lon = atan4(xyz[:, 1], xyz[:, 0])
        return 2 * self.u0 * z + 30 * self.amp * np.cos(4 * lon) * (z * (z * z - 1) ** 2)


class SphereTestCase2:
    """Williamson test case 2 (examples/sphere_swe_tc2.cpp:231-251, lpm_vorticity_gallery.hpp:264-278,
    lpm_surface_gallery.hpp:120-134): u = u0(-y, x, 0), zeta = 2 u0 z, sigma = 0,
    initial surface AS CODED in SphereTestCase2InitialSurface: s = h0 + Omega u0 cos^2(lat) / g (no u0^2/2 term,
    lpm_surface_gallery.hpp:126-131); the example's exact surface is h0 + (u0^2/2 + Omega u0) cos^2(lat) / g
    (examples/sphere_swe_tc2.cpp:239)."""

    def __init__(self, u0=2 * PI / 12, h0=10.0, g=1.0, Omega=2 * PI):
        self.u0, self.h0, self.g, self.Omega = u0, h0, g, Omega

    def vorticity(self, xyz):
        return 2 * self.u0 * xyz[:, 2]

    def divergence(self, xyz):
        return np.zeros(len(xyz))

    def velocity(self, xyz):
        return np.stack([-self.u0 * xyz[:, 1], self.u0 * xyz[:, 0], np.zeros(len(xyz))], axis=1)

    def double_dot(self, xyz):
        """grad u : grad u^T = -2 u0^2 z^2 (examples/sphere_swe_tc2.cpp:243-251)."""
        return -2 * self.u0 ** 2 * xyz[:, 2] ** 2

    def surface(self, xyz):
        return self.h0 + self.Omega * self.u0 * (1 - xyz[:, 2] ** 2) / self.g

    def surface_exact(self, xyz):
        return self.h0 + (0.5 * self.u0 ** 2 + self.Omega * self.u0) * (1 - xyz[:, 2] ** 2) / self.g

    def surface_laplacian_exact(self, xyz):
        """slap_exact (examples/sphere_swe_tc2.cpp:243-244): (u0^2 + 2 Omega u0)(2 sin^2 - cos^2)/g"""
        return (self.u0 ** 2 + 2 * self.Omega * self.u0) * (3 * xyz[:, 2] ** 2 - 1) / self.g


def synthetic_sphere_points(n, seed=20261017):
    """Collocated synthetic particle set of the sweep: normalised N(0,1)^3 points from a
    counter-based generator, RH54-like smooth vorticity, equal areas 4pi/N."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    x = rng.standard_normal((n, 3))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    rh = RossbyHaurwitz54(u0=2 * PI / 14)
    zeta = rh(x)
    area = np.full(n, 4 * PI / n)
    return x, zeta, area


def fibonacci_sphere_points(n):
    """Quasi-uniform synthetic particle set for the N-sweep: Fibonacci (golden-angle) lattice, smooth
    RH54 vorticity, equal areas 4pi/N.  Deterministic, no RNG; nearest-neighbour spacing ~ sqrt(4pi/N),
    so the 1 - x.y kernel stays as well conditioned as on LPM's own meshes."""
    i = np.arange(n, dtype=np.float64)
    z = 1.0 - (2.0 * i + 1.0) / n
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    phi = i * (PI * (3.0 - np.sqrt(5.0)))
    x = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    rh = RossbyHaurwitz54(u0=2 * PI / 14)
    return x, rh(x), np.full(n, 4 * PI / n)
