"""ctypes binding of the planar CPU oracle (oracle/lpm_oracle_plane.c).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.  Every function takes L= to run the same call
against oracle/_ref/liblpm_ref.so (the reference's own functors compiled in place) where that library exports it."""
import ctypes

import numpy as np

from . import oracle as _o

_dp = ctypes.POINTER(ctypes.c_double)
_up = ctypes.POINTER(ctypes.c_ubyte)
_d, _p, _m = _o._d, _o._p, _o._m

TOPO_ZERO = 0
TOPO_PLANAR_GAUSSIAN_MOUNTAIN = 1

SUM_FIELDS = ("ddot", "du1dx1", "du1dx2", "du2dx1", "du2dx2", "laps", "psi", "phi")


def lib():
    return _o.lib()


def ic2d_plane_sums(tgt_xy, src_xy, vort, area, mask, eps=0.0, targets_are_sources=False, with_psi=True, L=None):
    L = L or lib()
    src_xy, vort, area = _d(src_xy), _d(vort), _d(area)
    mask, mp = _m(mask)
    tx = src_xy if targets_are_sources else _d(tgt_xy)
    vel = np.zeros((tx.shape[0], 2))
    psi = np.zeros(tx.shape[0]) if with_psi else None
    L.oracle_ic2d_plane_sums(ctypes.c_int(tx.shape[0]), _p(tx), ctypes.c_int(src_xy.shape[0]), _p(src_xy), _p(vort),
                             _p(area), mp, ctypes.c_double(eps), ctypes.c_int(int(targets_are_sources)), _p(vel),
                             _p(psi))
    return vel, psi


def ic2d_plane_rk2_step(dt, f0, beta, eps, px, pz, pu, ppsi, ax, az, au, apsi, aa, am, n_steps=1, L=None):
    """In place on float64 C-contiguous arrays."""
    L = L or lib()
    am, mp = _m(am)
    for a in (px, pz, pu, ppsi, ax, az, au, apsi):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    aa = _d(aa)
    L.oracle_ic2d_plane_rk2_step(ctypes.c_double(dt), ctypes.c_double(f0), ctypes.c_double(beta), ctypes.c_double(eps),
                                 ctypes.c_int(px.shape[0]), _p(px), _p(pz), _p(pu), _p(ppsi),
                                 ctypes.c_int(ax.shape[0]), _p(ax), _p(az), _p(au), _p(apsi), _p(aa), mp,
                                 ctypes.c_int(n_steps))


def planar_swe_pair(x, y, zeta, sigma, area, src_s, tgt_s, eps, pse_eps, L=None):
    """planar_swe_sums_rhs_pse for one pair: the 9-tuple."""
    L = L or lib()
    x, y = _d(x), _d(y)
    r = np.zeros(9)
    L.oracle_planar_swe_sums_rhs_pse(_p(r), _p(x), _p(y), *(ctypes.c_double(float(v)) for v in
                                                             (zeta, sigma, area, src_s, tgt_s, eps, pse_eps)))
    return r


def swe_plane_sums(tgt_xy, tgt_surf, src_xy, vort, div, area, mask, src_surf, eps, pse_eps, targets_are_sources=False,
                   do_velocity=True, L=None):
    """PlanarSWEVertexSums / PlanarSWEFaceSums: returns dict(vel, ddot, du1dx1, ..., laps, psi, phi)."""
    L = L or lib()
    src_xy, vort, div, area, src_surf = map(_d, (src_xy, vort, div, area, src_surf))
    mask, mp = _m(mask)
    if targets_are_sources:
        tx, ts = src_xy, src_surf
    else:
        tx, ts = _d(tgt_xy), _d(tgt_surf)
    n = tx.shape[0]
    out = {"vel": np.zeros((n, 2))}
    for k in SUM_FIELDS:
        out[k] = np.zeros(n)
    L.oracle_swe_plane_sums(ctypes.c_int(n), _p(tx), _p(ts), ctypes.c_int(src_xy.shape[0]), _p(src_xy), _p(vort),
                            _p(div), _p(area), mp, _p(src_surf), ctypes.c_double(eps), ctypes.c_double(pse_eps),
                            ctypes.c_int(int(targets_are_sources)), ctypes.c_int(int(do_velocity)), _p(out["vel"]),
                            *(_p(out[k]) for k in SUM_FIELDS))
    return out


def swe_plane_tendencies(is_area, x, u, zeta, sigma, third, ddot, laps, f0, beta, g, dt, L=None):
    L = L or lib()
    x, u, zeta, sigma, third, ddot, laps = map(_d, (x, u, zeta, sigma, third, ddot, laps))
    n = x.shape[0]
    dz, ds, d3 = np.zeros(n), np.zeros(n), np.zeros(n)
    L.oracle_swe_plane_tendencies(ctypes.c_int(n), ctypes.c_int(int(is_area)), _p(dz), _p(ds), _p(d3), _p(x), _p(u),
                                  _p(zeta), _p(sigma), _p(third), _p(ddot), _p(laps), ctypes.c_double(f0),
                                  ctypes.c_double(beta), ctypes.c_double(g), ctypes.c_double(dt))
    return dz, ds, d3


def plane_topography(topo, xy, L=None):
    L = L or lib()
    L.oracle_plane_topography.restype = ctypes.c_double
    xy = _d(xy)
    return np.array([L.oracle_plane_topography(ctypes.c_int(topo), _p(np.ascontiguousarray(p))) for p in xy])


def swe_plane_surfaces(topo, px, ph, ax, amass, aarea, amask, asurf0, adepth0, abot0, L=None):
    """SetSurfaceFromDepth (passive) and SetDepthAndSurfaceFromMassAndArea (active; masked entries keep their input
    values): returns (psurf, pbot, adepth, asurf, abot)."""
    L = L or lib()
    px, ph, ax, amass, aarea = map(_d, (px, ph, ax, amass, aarea))
    amask, mp = _m(amask)
    n_p, n_a = px.shape[0], ax.shape[0]
    ps, pb = np.zeros(n_p), np.zeros(n_p)
    ah, asf, ab = _d(adepth0).copy(), _d(asurf0).copy(), _d(abot0).copy()
    L.oracle_swe_plane_set_surface_from_depth(ctypes.c_int(n_p), _p(ps), _p(pb), _p(px), _p(ph), ctypes.c_int(topo))
    L.oracle_swe_plane_set_depth_surface_from_mass_area(ctypes.c_int(n_a), _p(ah), _p(asf), _p(ab), _p(ax), _p(amass),
                                                        _p(aarea), mp, ctypes.c_int(topo))
    return ps, pb, ah, asf, ab


class _Side(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int)] + [(k, _dp) for k in (
        "xy", "vort", "div", "third", "mass", "depth", "surf", "bottom", "vel", "ddot", "du1dx1", "du1dx2", "du2dx1",
        "du2dx2", "laps", "psi", "phi")] + [("mask", _up)]


class PlaneSWEState:
    """The planar SWE<Seed> fields SWERK4 touches, float64 C-contiguous (src/lpm_swe.hpp:29-88).
    passive: depth is the prognostic third variable; active: area is."""
    PASSIVE = ("xy", "vort", "div", "depth", "surf", "bottom", "vel", "ddot", "du1dx1", "du1dx2", "du2dx1", "du2dx2",
               "laps", "psi", "phi")
    ACTIVE = ("xy", "vort", "div", "area", "mass", "depth", "surf", "bottom", "vel", "ddot", "du1dx1", "du1dx2",
              "du2dx1", "du2dx2", "laps", "psi", "phi")

    def __init__(self, passive, active, mask):
        n_p = np.asarray(passive["xy"]).shape[0]
        n_a = np.asarray(active["xy"]).shape[0]

        def take(src, k, n):
            if k in src and src[k] is not None:
                return np.ascontiguousarray(src[k], dtype=np.float64).copy()
            return np.zeros((n, 2)) if k == "vel" else np.zeros(n)
        self.p = {k: take(passive, k, n_p) for k in self.PASSIVE}
        self.a = {k: take(active, k, n_a) for k in self.ACTIVE}
        self.mask = np.ascontiguousarray(mask, dtype=np.uint8).copy()

    def copy(self):
        return PlaneSWEState(self.p, self.a, self.mask)

    def _sides(self):
        p, a = self.p, self.a
        P = _Side(n=p["xy"].shape[0], mask=None)
        A = _Side(n=a["xy"].shape[0], mask=self.mask.ctypes.data_as(_up))
        for k in ("xy", "vort", "div", "surf", "bottom", "vel", "ddot", "du1dx1", "du1dx2", "du2dx1", "du2dx2", "laps",
                  "psi", "phi"):
            setattr(P, k, _p(p[k]))
            setattr(A, k, _p(a[k]))
        P.third, P.depth, P.mass = _p(p["depth"]), _p(p["depth"]), None
        A.third, A.depth, A.mass = _p(a["area"]), _p(a["depth"]), _p(a["mass"])
        return P, A


def swe_plane_init_direct_sums(st, eps, pse_eps, do_velocity=True, L=None):
    """SWE<Seed>::init_direct_sums for PlaneGeometry (src/lpm_swe_impl.hpp:401-422), in place on `st`."""
    p, a = st.p, st.a
    rp = swe_plane_sums(p["xy"], p["surf"], a["xy"], a["vort"], a["div"], a["area"], st.mask, a["surf"], eps, pse_eps,
                        False, do_velocity, L=L)
    ra = swe_plane_sums(None, None, a["xy"], a["vort"], a["div"], a["area"], st.mask, a["surf"], eps, pse_eps, True,
                        do_velocity, L=L)
    for side, r in ((p, rp), (a, ra)):
        for k in SUM_FIELDS:
            side[k][:] = r[k]
        if do_velocity:
            side["vel"][:] = r["vel"]
    return st


def swe_plane_rk4_step(dt, f0, beta, g, eps, pse_eps, topo, st, n_steps=1):
    """SWERK4::advance_timestep for PlaneGeometry, in place on `st` (PlaneSWEState).  Restatement only: the
    reference's stepper class cannot be compiled here (mesh/VTK/Compadre headers), DESIGN.md section 3."""
    P, A = st._sides()
    lib().oracle_swe_plane_rk4_step(ctypes.c_double(dt), ctypes.c_double(f0), ctypes.c_double(beta),
                                    ctypes.c_double(g), ctypes.c_double(eps), ctypes.c_double(pse_eps),
                                    ctypes.c_int(topo), ctypes.byref(P), ctypes.byref(A), ctypes.c_int(n_steps))
    return st
