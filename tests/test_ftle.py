"""FTLE diagnostic: ComputeFTLE<Seed> / get_max_ftle, src/mesh/lpm_ftle.hpp.
CPU: the oracle restatement against golden outputs of the reference's functor compiled in place
(tests/golden/ref_ftle.npz, tests/golden/make_ftle_golden.py) and, where oracle/_ref exists, against the live build.
GPU: lpmx_ftle through the C ABI against the oracle and the goldens.  Tolerance: 1e-12 field-relative on log(lambda_1)
(different FMA contraction), face coordinates after the in-place normalisation to 2 ulp."""
import ctypes
import os

import numpy as np
import pytest

import ftle_cases
from conftest import field_rel_err
from lpm_b200.api import LAYOUT_LEFT, PolyMesh2d

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ftle.npz")
TOL = 1e-12


def _cases():
    return {"cubed3": ftle_cases.sphere_case(PolyMesh2d("cubed", 3)), "plane12": ftle_cases.plane_case()}


def test_oracle_ftle_matches_golden_reference_outputs(oracle):
    g = np.load(GOLDEN)
    for name, case in _cases().items():
        f, fp, mx = oracle.ftle(**case)
        assert np.isfinite(f).all()
        assert field_rel_err(f, g[name + "_ftle"]) < TOL, name
        assert np.abs(fp - g[name + "_face_phys"]).max() < 5e-16, name
        assert abs(mx - float(g[name + "_max"])) < TOL * abs(mx), name
        assert (f[case["mask"] != 0] == 0).all()  # divided faces are not written


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref",
                                                    "liblpm_ref.so")), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_ftle_matches_live_reference_build(oracle):
    R = ctypes.CDLL(oracle.REF_LIB)
    for name, case in _cases().items():
        a, b = oracle.ftle(**case), oracle.ftle(**case, L=R)
        assert field_rel_err(a[0], b[0]) < TOL and np.abs(a[1] - b[1]).max() < 5e-16 and abs(a[2] - b[2]) < TOL


def test_oracle_ftle_quirks(oracle):
    """As coded: (1) the sphere branch normalises the face's physical coordinates in place, vertices are left alone;
    (2) the in-place "shift to vertex 1" loop leaves vertices 2 and 3 unshifted, so the planar value depends on where
    the panel sits (translating the whole configuration changes it); (3) the identity map gives lambda_1 = 1 up to the
    square root of round-off -- or NaN where half_trace^2 - det rounds to a tiny negative number, in the reference too
    (the elementwise F_ij F_ji is not F^T F, so not even a rigid rotation gives 0; not asserted)."""
    c = ftle_cases.sphere_case(PolyMesh2d("cubed", 2))
    f, fp, _ = oracle.ftle(**c)
    leaf = c["mask"] == 0
    assert np.abs(np.linalg.norm(fp[leaf], axis=1) - 1).max() < 4e-16
    assert np.abs(np.linalg.norm(c["face_phys"][leaf], axis=1) - 1).max() > 1e-8
    assert np.array_equal(fp[~leaf], c["face_phys"][~leaf])
    p = ftle_cases.plane_case()
    f0 = oracle.ftle(**p)[0]
    shift = np.array([0.5, -0.25])
    q = dict(p, vert_ref=p["vert_ref"] + shift, face_ref=p["face_ref"] + shift, vert_phys=p["vert_phys"] + shift,
             face_phys=p["face_phys"] + shift)
    assert np.abs(oracle.ftle(**q)[0] - f0).max() > 1e-3
    r = dict(c, vert_phys=c["vert_ref"], face_phys=c["face_ref"])
    fr = oracle.ftle(**r)[0]
    ok = np.isfinite(fr)
    assert ok.sum() > 0.5 * leaf.sum() and np.abs(fr[ok]).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cubed3", "plane12", "cubed5"])
def test_gpu_ftle_matches_oracle_and_golden(engine, oracle, name):
    case = _cases()[name] if name != "cubed5" else ftle_cases.sphere_case(PolyMesh2d("cubed", 5))
    ref_f, ref_fp, ref_mx = oracle.ftle(**case)
    fp = np.ascontiguousarray(case["face_phys"]).copy()
    pre = np.full(case["mask"].shape[0], -7.0)
    f, mx = engine.ftle(case["geom"], case["vert_phys"], case["vert_ref"], fp, case["face_ref"], case["face_verts"],
                        case["mask"], ftle=pre.copy())
    leaf = case["mask"] == 0
    assert (f[~leaf] == -7.0).all()  # entries of divided faces are left alone
    assert field_rel_err(f[leaf], ref_f[leaf]) < TOL
    assert abs(mx - ref_mx) < TOL * abs(ref_mx)
    if case["geom"] == 0:
        assert np.abs(fp - ref_fp).max() < 5e-16
    else:
        assert np.array_equal(fp, case["face_phys"])
    if name != "cubed5":
        g = np.load(GOLDEN)
        assert field_rel_err(f[leaf], g[name + "_ftle"][leaf]) < TOL


@pytest.mark.gpu
def test_gpu_ftle_layout_left_device_pointers(engine, oracle):
    """Kokkos-CUDA layouts (LayoutLeft coordinates and connectivity) with device-resident arrays."""
    import torch
    case = ftle_cases.sphere_case(PolyMesh2d("cubed", 4))
    ref_f, ref_fp, ref_mx = oracle.ftle(**case)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)  # noqa: E731
    fp = t(case["face_phys"])
    out = torch.zeros(case["mask"].shape[0], dtype=torch.float64, device=dev)
    f, mx = engine.ftle(0, t(case["vert_phys"]), t(case["vert_ref"]), fp, t(case["face_ref"]),
                        torch.from_numpy(np.ascontiguousarray(case["face_verts"].T)).to(dev),
                        torch.from_numpy(case["mask"]).to(dev), ftle=out, layout=LAYOUT_LEFT, verts_layout=LAYOUT_LEFT)
    engine.sync()
    leaf = case["mask"] == 0
    assert field_rel_err(f.cpu().numpy()[leaf], ref_f[leaf]) < TOL
    assert np.abs(fp.cpu().numpy().T - ref_fp).max() < 5e-16
    assert abs(mx - ref_mx) < TOL * abs(ref_mx)
