"""ctypes binding of the CPU oracle (oracle/lpm_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liblpm_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "liblpm_ref.so")

_dp = ctypes.POINTER(ctypes.c_double)
_up = ctypes.POINTER(ctypes.c_ubyte)


def build(ref=False):
    targets = ["all"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", _HERE] + targets, check=True, capture_output=True)


def _load(path):
    if not os.path.exists(path):
        if path == LIB:
            build()
        else:
            raise FileNotFoundError(path)
    return ctypes.CDLL(path)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load(LIB)
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _m(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(_up)


def num_threads():
    return lib().oracle_num_threads()


def bve_velocity(tgt_xyz, src_xyz, vort, area, mask, collocated=False, long_double=False, L=None):
    L = L or lib()
    src_xyz, vort, area = _d(src_xyz), _d(vort), _d(area)
    mask, mp = _m(mask)
    tx = src_xyz if collocated else _d(tgt_xyz)
    out = np.zeros((tx.shape[0], 3))
    f = L.oracle_bve_velocity_ld if long_double else L.oracle_bve_velocity
    f(ctypes.c_int(tx.shape[0]), _p(tx), ctypes.c_int(src_xyz.shape[0]), _p(src_xyz), _p(vort), _p(area), mp,
      ctypes.c_int(int(collocated)), _p(out))
    return out


def bve_velocity_subset(idx, xyz, vort, area, mask, long_double=False, L=None):
    """Collocated velocity (BVEFaceVelocity) at the particles idx only: all unmasked sources j != idx[k]."""
    L = L or lib()
    xyz, vort, area = _d(xyz), _d(vort), _d(area)
    mask, mp = _m(mask)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.zeros((idx.shape[0], 3))
    f = L.oracle_bve_velocity_subset_ld if long_double else L.oracle_bve_velocity_subset
    f(ctypes.c_int(idx.shape[0]), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), ctypes.c_int(xyz.shape[0]), _p(xyz),
      _p(vort), _p(area), mp, _p(out))
    return out


def bve_streamfn(tgt_xyz, src_xyz, vort, area, mask, collocated=False, L=None):
    L = L or lib()
    src_xyz, vort, area = _d(src_xyz), _d(vort), _d(area)
    mask, mp = _m(mask)
    tx = src_xyz if collocated else _d(tgt_xyz)
    out = np.zeros(tx.shape[0])
    L.oracle_bve_streamfn(ctypes.c_int(tx.shape[0]), _p(tx), ctypes.c_int(src_xyz.shape[0]), _p(src_xyz), _p(vort),
                          _p(area), mp, ctypes.c_int(int(collocated)), _p(out))
    return out


def bve_rk4_step(dt, Omega, vx, vz, vu, fx, fz, fu, fa, fm, n_steps=1, L=None):
    """In place on float64 C-contiguous arrays."""
    L = L or lib()
    fm, mp = _m(fm)
    for a in (vx, vz, vu, fx, fz, fu):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    fa = _d(fa)
    L.oracle_bve_rk4_step(ctypes.c_double(dt), ctypes.c_double(Omega), ctypes.c_int(vx.shape[0]), _p(vx), _p(vz),
                          _p(vu), ctypes.c_int(fx.shape[0]), _p(fx), _p(fz), _p(fu), _p(fa), mp,
                          ctypes.c_int(n_steps))


def ic2d_sums(tgt_xyz, src_xyz, vort, area, mask, eps=0.0, targets_are_sources=False, with_psi=True, L=None):
    L = L or lib()
    src_xyz, vort, area = _d(src_xyz), _d(vort), _d(area)
    mask, mp = _m(mask)
    tx = src_xyz if targets_are_sources else _d(tgt_xyz)
    vel = np.zeros((tx.shape[0], 3))
    psi = np.zeros(tx.shape[0]) if with_psi else None
    L.oracle_ic2d_sums(ctypes.c_int(tx.shape[0]), _p(tx), ctypes.c_int(src_xyz.shape[0]), _p(src_xyz), _p(vort),
                       _p(area), mp, ctypes.c_double(eps), ctypes.c_int(int(targets_are_sources)), _p(vel), _p(psi))
    return vel, psi


def ic2d_rk2_step(dt, Omega, eps, px, pz, pu, ppsi, ax, az, au, apsi, aa, am, n_steps=1, L=None):
    L = L or lib()
    am, mp = _m(am)
    for a in (px, pz, pu, ppsi, ax, az, au, apsi):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    aa = _d(aa)
    L.oracle_ic2d_rk2_step(ctypes.c_double(dt), ctypes.c_double(Omega), ctypes.c_double(eps),
                           ctypes.c_int(px.shape[0]), _p(px), _p(pz), _p(pu), _p(ppsi), ctypes.c_int(ax.shape[0]),
                           _p(ax), _p(az), _p(au), _p(apsi), _p(aa), mp, ctypes.c_int(n_steps))


def swe_pair(x, y, eps=0.0, L=None):
    """kzeta, ksigma (unit strength: vort*area = 1), grad_kzeta, grad_ksigma for one pair."""
    L = L or lib()
    x, y = _d(x), _d(y)
    kz, ks, gkz, gks = np.zeros(3), np.zeros(3), np.zeros(9), np.zeros(9)
    L.oracle_kzeta_sphere(_p(kz), _p(x), _p(y), ctypes.c_double(1.0), ctypes.c_double(1.0), ctypes.c_double(eps))
    L.oracle_ksigma_sphere(_p(ks), _p(x), _p(y), ctypes.c_double(1.0), ctypes.c_double(1.0), ctypes.c_double(eps))
    L.oracle_grad_kzeta(_p(gkz), _p(x), _p(y), ctypes.c_double(eps))
    L.oracle_grad_ksigma(_p(gks), _p(x), _p(y), ctypes.c_double(eps))
    return kz, ks, gkz, gks


def swe_sphere_sums(tgt_xyz, src_xyz, vort, div, area, mask, eps=0.0, targets_are_sources=False, do_velocity=True,
                    L=None):
    L = L or lib()
    src_xyz, vort, div, area = _d(src_xyz), _d(vort), _d(div), _d(area)
    mask, mp = _m(mask)
    tx = src_xyz if targets_are_sources else _d(tgt_xyz)
    n = tx.shape[0]
    vel, ddot, grad = np.zeros((n, 3)), np.zeros(n), np.zeros((n, 9))
    L.oracle_swe_sphere_sums(ctypes.c_int(n), _p(tx), ctypes.c_int(src_xyz.shape[0]), _p(src_xyz), _p(vort), _p(div),
                             _p(area), mp, ctypes.c_double(eps), ctypes.c_int(int(targets_are_sources)),
                             ctypes.c_int(int(do_velocity)), _p(vel), _p(ddot), _p(grad))
    return vel, ddot, grad


def swe_sphere_sums_ld(tgt_xyz, src_xyz, vort, div, area, mask, eps=0.0, targets_are_sources=False, L=None):
    """Long-double adjudicator of swe_sphere_sums (oracle_swe_sphere_sums_ld): (vel, ddot, grad9) rounded to double."""
    L = L or lib()
    src_xyz, vort, div, area = _d(src_xyz), _d(vort), _d(div), _d(area)
    mask, mp = _m(mask)
    tx = src_xyz if targets_are_sources else _d(tgt_xyz)
    n = tx.shape[0]
    vel, ddot, grad = np.zeros((n, 3)), np.zeros(n), np.zeros((n, 9))
    L.oracle_swe_sphere_sums_ld(ctypes.c_int(n), _p(tx), ctypes.c_int(src_xyz.shape[0]), _p(src_xyz), _p(vort), _p(div),
                                _p(area), mp, ctypes.c_double(eps), ctypes.c_int(int(targets_are_sources)), _p(vel), _p(ddot),
                                _p(grad))
    return vel, ddot, grad


# ---- SWE RK2 (family C stepper) ---------------------------------------------------------------------
LAPS_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, ctypes.c_int, _dp, _dp,
                           _up, _dp)


def swe_tendencies(is_area, x, u, zeta, sigma, third, ddot, laps, Omega, g, dt, L=None):
    """SWEVorticityDivergence{Height,Area}Tendencies<SphereGeometry>: returns (dzeta, dsigma, dh | darea)."""
    L = L or lib()
    x, u, zeta, sigma, third, ddot, laps = map(_d, (x, u, zeta, sigma, third, ddot, laps))
    n = x.shape[0]
    dz, ds, d3 = np.zeros(n), np.zeros(n), np.zeros(n)
    L.oracle_swe_tendencies(ctypes.c_int(n), ctypes.c_int(int(is_area)), _p(dz), _p(ds), _p(d3), _p(x), _p(u),
                            _p(zeta), _p(sigma), _p(third), _p(ddot), _p(laps), ctypes.c_double(Omega),
                            ctypes.c_double(g), ctypes.c_double(dt))
    return dz, ds, d3


class SWEState:
    """The SWE<Seed> fields the stepper touches, as float64 C-contiguous numpy arrays (src/lpm_swe.hpp:29-88)."""
    PASSIVE = ("xyz", "vort", "div", "depth", "surf", "bottom", "vel", "ddot", "laps")
    ACTIVE = ("xyz", "vort", "div", "area", "mass", "depth", "surf", "bottom", "vel", "ddot", "laps")

    def __init__(self, passive, active, mask):
        self.p = {k: np.ascontiguousarray(passive[k], dtype=np.float64).copy() for k in self.PASSIVE}
        self.a = {k: np.ascontiguousarray(active[k], dtype=np.float64).copy() for k in self.ACTIVE}
        self.mask = np.ascontiguousarray(mask, dtype=np.uint8).copy()

    def copy(self):
        return SWEState(self.p, self.a, self.mask)


def swe_rk2_step(dt, Omega, g, eps, st, laps_fn=None, n_steps=1, L=None):
    """SWERK2::advance_timestep_impl on the sphere, in place on `st` (SWEState).  laps_fn(stage, px, psurf, ax,
    asurf, amask) -> (plaps, alaps) stands in for the GMLS Laplacian; None leaves st.*['laps'] unchanged."""
    L = L or lib()
    p, a = st.p, st.a
    np_, na = p["xyz"].shape[0], a["xyz"].shape[0]

    def cb(user, stage, n_p, px, psurf, plaps, n_a, ax, asurf, amask, alaps):
        pxv = np.ctypeslib.as_array(px, shape=(n_p, 3))
        axv = np.ctypeslib.as_array(ax, shape=(n_a, 3))
        pl, al = laps_fn(stage, pxv, np.ctypeslib.as_array(psurf, shape=(n_p,)), axv,
                         np.ctypeslib.as_array(asurf, shape=(n_a,)), np.ctypeslib.as_array(amask, shape=(n_a,)))
        np.ctypeslib.as_array(plaps, shape=(n_p,))[:] = pl
        np.ctypeslib.as_array(alaps, shape=(n_a,))[:] = al

    fn = LAPS_FN(cb) if laps_fn is not None else ctypes.cast(None, LAPS_FN)
    L.oracle_swe_rk2_step(ctypes.c_double(dt), ctypes.c_double(Omega), ctypes.c_double(g), ctypes.c_double(eps),
                          ctypes.c_int(np_), _p(p["xyz"]), _p(p["vort"]), _p(p["div"]), _p(p["depth"]), _p(p["surf"]),
                          _p(p["bottom"]), _p(p["vel"]), _p(p["ddot"]), _p(p["laps"]), ctypes.c_int(na),
                          _p(a["xyz"]), _p(a["vort"]), _p(a["div"]), _p(a["area"]), _p(a["mass"]), _p(a["depth"]),
                          _p(a["surf"]), _p(a["bottom"]), _p(a["vel"]), _p(a["ddot"]), _p(a["laps"]),
                          st.mask.ctypes.data_as(_up), fn, None, ctypes.c_int(n_steps))
    return st


def ftle(geom, vert_phys, vert_ref, face_phys, face_ref, face_verts, mask, L=None):
    """ComputeFTLE for quadrilateral faces (mesh/lpm_ftle.hpp); geom 0 = sphere, 1 = plane.  Returns (ftle with
    zeros at masked faces, face_phys after the call -- the sphere branch normalises it in place, max_ftle)."""
    L = L or lib()
    vert_phys, vert_ref, face_ref = _d(vert_phys), _d(vert_ref), _d(face_ref)
    face_phys = _d(face_phys).copy()
    fv = np.ascontiguousarray(face_verts, dtype=np.int32)
    mask, mp = _m(mask)
    nf, nv = face_ref.shape[0], vert_ref.shape[0]
    out = np.zeros(nf)
    L.oracle_ftle(ctypes.c_int(geom), ctypes.c_int(nv), _p(vert_phys), _p(vert_ref), ctypes.c_int(nf), _p(face_phys),
                  _p(face_ref), fv.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), mp, _p(out))
    L.oracle_max_ftle.restype = ctypes.c_double
    mx = L.oracle_max_ftle(ctypes.c_int(nf), _p(out), mp)
    return out, face_phys, float(mx)
