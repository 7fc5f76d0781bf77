"""Velocity pair sums through the constant bank (lpm_b200/csrc/lpmx_const_stream.cu; automatic for >= 1e6 targets per rank,
DESIGN.md section 4.1b).  CPU: the launch-shape planner and the kernel body on the host.  GPU: parity with the oracle and with
the default stream-K kernel, forced onto small meshes (LPMX_CONST_MIN_TARGETS) and at icos-7 on sampled targets."""
import ctypes
import os

import numpy as np
import pytest

from lpm_b200 import _lib

SMS = 148


def _shape(n_tgt, sms=SMS):
    T, nw, grid = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = _lib.lib().lpmx_const_stream_shape(sms, n_tgt, ctypes.byref(T), ctypes.byref(nw), ctypes.byref(grid))
    assert rc == 0
    return T.value, nw.value, grid.value


@pytest.mark.parametrize("n_tgt", [229376, 9382, 600742, 2402982, 9611942, 300000, 1201491, 189440, 1, 12345])
def test_shape_covers_the_targets_and_fits_a_cta(n_tgt):
    T, nw, grid = _shape(n_tgt)
    assert T in (5, 6, 7) and nw == 8  # 8 warps = 2 per scheduler; the measured shapes (profiles/r2e_icos8_const_shapes.txt)
    tb = T * nw * 32
    assert grid * tb >= n_tgt > (grid - 1) * tb


@pytest.mark.parametrize("n_tgt,least", [(1000000, 0.94), (2402982, 0.97), (9611942, 0.98), (1201491, 0.90)])
def test_shape_wastes_little_of_the_chip(n_tgt, least):
    """The sizes the automatic mode takes (>= 1e6 targets per rank): the threshold itself, icos-8 and icos-9 on one GPU, icos-8 on
    two = icos-9 on eight: fraction of (waves x SMs x targets per CTA) that is work."""
    T, nw, grid = _shape(n_tgt)
    waves = -(-grid // SMS)
    assert n_tgt / (waves * SMS * T * nw * 32) >= least


def test_shape_rejects_bad_arguments():
    T = ctypes.c_int()
    assert _lib.lib().lpmx_const_stream_shape(0, 10, ctypes.byref(T), ctypes.byref(T), ctypes.byref(T)) != 0
    assert _lib.lib().lpmx_const_stream_shape(148, 0, ctypes.byref(T), ctypes.byref(T), ctypes.byref(T)) != 0


def test_kernel_body_host_model(tmp_path):
    """The body of pair_sum_const_kernel (lpmx_const_stream_body.h) run on the host, one loop iteration per CUDA thread, around
    a restatement of the launch sequence (640-record batches, alternating halves, zero padding, `first`): T = 4..8, ragged
    target counts, both target layouts, collocated self-pair exclusion -- equal to a direct double loop."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "const_stream_model")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(root, "lpm_b200", "csrc"),
                    os.path.join(root, "tests", "cpp", "const_stream_model.cpp"), "-o", exe], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.count(" ok") == 6 and "FAILED" not in p.stdout, p.stdout + p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("seed,depth", [("icos", 4), ("cubed", 5)])
def test_gpu_const_stream_velocity_matches_oracle_and_default_kernel(oracle, mode, seed, depth, monkeypatch):
    from lpm_b200 import gallery
    from lpm_b200.api import Engine, PolyMesh2d
    from conftest import field_rel_err
    monkeypatch.setenv("LPMX_CONST_MIN_TARGETS", "1")
    m = PolyMesh2d(seed, depth)
    f = gallery.RossbyHaurwitz54()
    f.set_stationary_wave_speed()
    fz = f(m.face_xyz)
    ref_v = oracle.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
    ref_f = oracle.bve_velocity(None, m.face_xyz, fz, m.face_area, m.face_mask, collocated=True)
    e0, e1 = Engine(0), Engine(0)
    try:
        e0.pair_sum_const_stream(0)
        e1.pair_sum_const_stream(mode)
        got0_v = e0.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
        got1_v = e1.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
        got1_f = e1.bve_velocity(None, m.face_xyz, fz, m.face_area, m.face_mask, collocated=True)
        leaf = m.face_mask == 0  # divided icosahedral faces sit on their centre child: 0/0 in the reference too
        assert field_rel_err(got1_v, ref_v) <= 1e-12
        assert field_rel_err(got1_f, ref_f, leaf) <= 1e-12   # collocated: the self pair is removed by index
        # two summation orders of the same terms, each within 1e-12 of the oracle (4e-15 .. 1.1e-13 measured, r2b)
        assert field_rel_err(got1_v, got0_v) <= 5e-13
        # and through a stepper: two RK4 steps
        got1_f[~leaf] = 0.0
        st0 = [m.vert_xyz.copy(), f(m.vert_xyz), got0_v.copy(), m.face_xyz.copy(), fz.copy(), got1_f.copy()]
        st1 = [a.copy() for a in st0]
        e0.bve_rk4_step(0.01, 2 * np.pi, *st0, m.face_area, m.face_mask, n_steps=2)
        e1.bve_rk4_step(0.01, 2 * np.pi, *st1, m.face_area, m.face_mask, n_steps=2)
        assert max(field_rel_err(st1[0], st0[0]), field_rel_err(st1[1], st0[1]), field_rel_err(st1[3], st0[3], leaf)) <= 1e-12
    finally:
        e0.close()
        e1.close()


@pytest.mark.gpu
def test_gpu_const_stream_at_icos7_sampled_targets(oracle, monkeypatch):
    """600 742 targets x 327 680 leaf sources through the constant bank (forced: the automatic mode starts at 1e6 targets),
    512 launches of 640 sources with the bank halves refilled behind them: 2 048 sampled vertex targets against the oracle, all
    vertex targets against the default kernel, and a second handle on the same device keeps the default kernel (one bank)."""
    from lpm_b200 import gallery
    from lpm_b200.api import Engine, PolyMesh2d
    from conftest import field_rel_err
    monkeypatch.setenv("LPMX_CONST_MIN_TARGETS", "1")
    m = PolyMesh2d("icos", 7)
    fz = gallery.GaussianVortexSphere()(m.face_xyz)
    e0, e1, e2 = Engine(0), Engine(0), Engine(0)
    try:
        e0.pair_sum_const_stream(0)
        e1.pair_sum_const_stream(1)
        e2.pair_sum_const_stream(1)
        l1 = e1.launch_count()
        got1 = e1.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
        n1 = e1.launch_count() - l1
        l2 = e2.launch_count()
        got2 = e2.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)  # bank owned by e1: ring kernel
        n2 = e2.launch_count() - l2
        got0 = e0.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
        assert n1 > 500 and n2 < 20
        assert np.array_equal(got2, got0)
        idx = np.sort(np.random.default_rng(5).choice(m.n_verts, 2048, replace=False))
        ref = oracle.bve_velocity(m.vert_xyz[idx], m.face_xyz, fz, m.face_area, m.face_mask)
        assert field_rel_err(got1[idx], ref) <= 1e-12
        assert field_rel_err(got1, got0) <= 1e-12
    finally:
        e0.close()
        e1.close()
        e2.close()
