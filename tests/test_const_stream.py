"""Velocity pair sums through the constant banks (lpm_b200/csrc/lpmx_const_stream.cu; automatic from one full wave of the chip
per rank).  CPU: the planner that splits the targets between the bank path (whole waves) and the ring
kernel (the remainder), and the kernel body on the host.  GPU: parity with the oracle and with the default stream-K kernel,
forced onto small meshes (LPMX_CONST_MIN_TARGETS), at icos-7 on sampled targets, and the automatic split at cubed-7."""
import ctypes
import os

import numpy as np
import pytest

from lpm_b200 import _lib

SMS = 148


def _split(n_tgt, n_src, sms=SMS):
    T, nw, ctas, n_const = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    t, tr = ctypes.c_double(), ctypes.c_double()
    rc = _lib.lib().lpmx_const_stream_split(sms, n_tgt, n_src, ctypes.byref(T), ctypes.byref(nw), ctypes.byref(ctas),
                                            ctypes.byref(n_const), ctypes.byref(t), ctypes.byref(tr))
    assert rc == 0
    return T.value, nw.value, ctas.value, n_const.value, t.value, tr.value


@pytest.mark.parametrize("n_tgt,n_src", [(229376, 98304), (9382, 5120), (2402982, 1310720), (9611942, 5242880), (1, 5120)])
def test_pipelined_plan_puts_every_target_through_the_banks(n_tgt, n_src):
    """The default: bank launches pipelined with programmatic dependent launch -- no waves to fill, no remainder; CTAs of
    T = 6 x 4 compute warps = 768 targets, three per SM."""
    T, nw, ctas, n_const, t, tr = _split(n_tgt, n_src)
    assert (T, nw) == (6, 4) and n_const == n_tgt and ctas == -(-n_tgt // 768) and t > 0 and tr > 0
    if n_tgt >= 189440:
        assert t < 0.98 * tr  # the automatic mode takes it


def test_pipelined_model_against_the_measured_rank_sizes():
    """The planner's model against what was measured on one GPU for a rank's target counts x the 98 304 sources of cubed-7
    (profiles/r3a_auto_rank_sizes.txt, profiles/r2z_rank_size_sweep.jsonl): the bank path's modelled time within 6 % of the
    measured one from 28 672 targets up, and the decision (model below 0.98 x the ring kernel's model) the measured winner."""
    measured = {16384: (1.051, 0.987), 28672: (1.789, 1.589), 57344: (3.524, 2.997), 114688: (6.946, 5.869), 229376: (13.809, 11.521)}
    for n, (ring_ms, bank_ms) in measured.items():
        T, nw, ctas, n_const, t, tr = _split(n, 98304)
        assert abs(tr * 1e3 - ring_ms) <= 0.03 * ring_ms, (n, tr * 1e3, ring_ms)
        if n >= 28672:
            assert abs(t * 1e3 - bank_ms) <= 0.06 * bank_ms, (n, t * 1e3, bank_ms)
        assert (t < 0.98 * tr) == (bank_ms < ring_ms), (n, t, tr)
    # 12 288 targets (list A of a rank of eight): measured 0.884 ms through the banks against 0.767 ms -- below the floor of
    # the automatic mode (16 384), so the question is never put to the model
    assert 12288 < 16384


@pytest.mark.parametrize("n_tgt,n_src", [(229376, 98304), (9382, 5120), (600742, 327680), (2402982, 1310720), (9611942, 5242880),
                                         (300000, 300000), (1201491, 1310720), (189440, 98304), (1, 5120), (12345, 6000)])
def test_split_covers_the_targets_with_whole_waves_plus_a_remainder(n_tgt, n_src, monkeypatch):
    """Without the pipelining (LPMX_CONST_PDL=0) a launch has to fill whole waves of the chip."""
    monkeypatch.setenv("LPMX_CONST_PDL", "0")
    T, nw, ctas, n_const, t, tr = _split(n_tgt, n_src)
    assert T in (5, 6, 7) and nw == 8  # 8 warps = 2 per scheduler; the measured shapes (profiles/r2e_icos8_const_shapes.txt)
    tb = T * nw * 32
    assert 0 < n_const <= n_tgt and t > 0 and tr > 0
    if n_const < n_tgt:  # a remainder follows: the bank path's CTAs are all full and come in whole waves
        assert n_const == ctas * tb and ctas % SMS == 0
    else:  # no remainder: the last CTA may be partly filled
        assert ctas * tb >= n_tgt > (ctas - 1) * tb


def test_split_at_the_headline_sizes(monkeypatch):
    """(LPMX_CONST_PDL=0) cubed-7 (BASELINE configs[1]): one wave of T = 6 = 227 328 targets through the banks, 2 048 through the ring kernel, and
    the model prefers that to the ring kernel alone; icos-8 on one GPU: whole waves + a remainder, also preferred; a rank's
    share of cubed-7 on two GPUs (114 688 targets) cannot fill one wave and stays with the ring kernel in the automatic mode."""
    monkeypatch.setenv("LPMX_CONST_PDL", "0")
    T, nw, ctas, n_const, t, tr = _split(229376, 98304)
    assert (T, nw, ctas, n_const) == (6, 8, 148, 227328) and t < 0.98 * tr
    T, nw, ctas, n_const, t, tr = _split(2402982, 1310720)
    assert n_const >= 0.9 * 2402982 and t < 0.98 * tr
    T, nw, ctas, n_const, t, tr = _split(114688, 98304)
    assert 114688 < SMS * 5 * 256  # below the automatic mode's floor (make_const_plan)


def test_split_rejects_bad_arguments():
    T = ctypes.c_int()
    args = [ctypes.byref(T)] * 4 + [None, None]
    assert _lib.lib().lpmx_const_stream_split(0, 10, 10, *args) != 0
    assert _lib.lib().lpmx_const_stream_split(148, 0, 10, *args) != 0
    assert _lib.lib().lpmx_const_stream_split(148, 10, 0, *args) != 0


def test_kernel_body_host_model(tmp_path):
    """The body of pair_sum_const_kernel (lpmx_const_stream_body.h) run on the host, one loop iteration per CUDA thread, around
    a restatement of the launch sequence (1 280-record batches, alternating banks, zero padding, `first`): T = 4..8, ragged
    target counts, both target layouts, index lists of targets, collocated self-pair exclusion -- equal to a direct double loop."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "const_stream_model")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(root, "lpm_b200", "csrc"),
                    os.path.join(root, "tests", "cpp", "const_stream_model.cpp"), "-o", exe], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.count(" ok") == 8 and "FAILED" not in p.stdout, p.stdout + p.stderr


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
@pytest.mark.parametrize("pdl", ["1", "0"])
def test_gpu_const_stream_automatic_split_at_cubed7(pdl, monkeypatch):
    """BASELINE configs[1] as bench.py runs it, nothing forced.  pdl = 1 (the default): all 229 376 targets of the resident
    solver's evaluation through the banks, launches pipelined; pdl = 0: one wave through the banks (227 328) and 2 048 through
    the ring kernel.  One BVERK4 step against the same step with the path switched off (two summation orders of the same
    terms; the tail of the face list is the ring kernel's share when pdl = 0), and the launch count shows the path was taken
    (77 bank launches per evaluation).  The same step against the ORACLE -- every face target,
    so the remainder too, and 4 096 sampled vertices -- is test_rk4_step_at_cubed7_sampled_targets_against_the_oracle
    (tests/test_gpu_parity_bve.py), which runs on the automatic mode as well."""
    from lpm_b200 import gallery
    from lpm_b200.api import BVESolver, Engine, PolyMesh2d
    from conftest import field_rel_err
    monkeypatch.setenv("LPMX_CONST_PDL", pdl)
    m = PolyMesh2d("cubed", 7)
    f = gallery.RossbyHaurwitz54()
    f.set_stationary_wave_speed()
    vz, fz = f(m.vert_xyz), f(m.face_xyz)
    area, mask = np.ascontiguousarray(m.face_area), np.ascontiguousarray(m.face_mask)
    states, counts = [], []
    for mode in (0, -1):
        e = Engine(0)
        try:
            e.pair_sum_const_stream(mode)
            s = BVESolver(e, m.n_verts, m.n_faces)
            s.set_state(m.vert_xyz, vz, None, m.face_xyz, fz, None, area, mask)
            s.init_velocity()
            l0 = e.launch_count()
            s.advance(0.003, 2 * np.pi, 1)
            counts.append(e.launch_count() - l0)
            out = [np.empty((m.n_verts, 3)), np.empty(m.n_verts), np.empty((m.n_verts, 3)), np.empty((m.n_faces, 3)),
                   np.empty(m.n_faces), np.empty((m.n_faces, 3))]
            s.get_state(*out)
            s.close()
            states.append(out)
        finally:
            e.close()
    assert counts[0] < 20 and counts[1] >= 4 * 77, counts
    a, b = states
    for k, tol in ((0, 1e-13), (1, 1e-12), (2, 1e-12), (3, 1e-13), (4, 1e-12), (5, 1e-12)):
        assert field_rel_err(b[k], a[k]) <= tol, (k, field_rel_err(b[k], a[k]))
    tail = slice(m.n_faces - 2048, m.n_faces)  # the ring kernel's share
    assert field_rel_err(b[5][tail], a[5][tail]) <= 1e-12 and np.abs(b[5][tail]).max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("merge", ["0", "1"])
def test_gpu_const_stream_with_split_target_lists(oracle, merge, monkeypatch):
    """What a rank of a multi-GPU run does -- list A (its leaf faces) and list B (vertices and divided faces) as index lists --
    forced onto one GPU and a small mesh (LPMX_FORCE_SPLIT=1, LPMX_CONST_MIN_TARGETS=1).  merge = 0: each list through its own
    bank sequence (large meshes: the exchange of list A's records runs beside list B's sum); merge = 1: both lists through ONE
    sequence into one accumulator array that the two stage kernels read at their offsets (small meshes: LPMX_MERGE_LISTS).
    Two BVERK4 steps and two Incompressible2DRK2 steps against the oracle, and the bank launches are counted."""
    from lpm_b200 import gallery
    from lpm_b200.api import Engine, PolyMesh2d
    from conftest import field_rel_err
    monkeypatch.setenv("LPMX_FORCE_SPLIT", "1")
    monkeypatch.setenv("LPMX_CONST_MIN_TARGETS", "1")
    monkeypatch.setenv("LPMX_MERGE_LISTS", merge)
    lists = 1 if merge == "1" else 2
    for seed, depth in (("cubed", 5), ("icos", 5)):
        m = PolyMesh2d(seed, depth)
        f = gallery.RossbyHaurwitz54()
        f.set_stationary_wave_speed()
        leaf = m.face_mask == 0
        fz = f(m.face_xyz)
        vu = oracle.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
        fu = oracle.bve_velocity(None, m.face_xyz, fz, m.face_area, m.face_mask, collocated=True)
        fu[~leaf] = 0.0
        ref = [m.vert_xyz.copy(), f(m.vert_xyz), vu, m.face_xyz.copy(), fz.copy(), fu]
        got = [a.copy() for a in ref]
        oracle.bve_rk4_step(0.01, 2 * np.pi, *ref, m.face_area, m.face_mask, n_steps=2)
        e = Engine(0)
        try:
            e.pair_sum_const_stream(1)
            c0 = e.const_stream_launch_count()
            e.bve_rk4_step(0.01, 2 * np.pi, *got, m.face_area, m.face_mask, n_steps=2)
            n = e.const_stream_launch_count() - c0
            n_banks = -(-int(leaf.sum()) // 640)  # target sets below 100 000: half a bank (640 records) per launch
            assert n == 2 * 4 * lists * n_banks, (n, lists, n_banks)  # 2 steps x 4 evaluations x sequences x bank launches
            # Incompressible2DRK2 with eps = 0 and a lazy psi: the velocity-only evaluations take the same path
            ic_ref = [m.vert_xyz.copy(), f(m.vert_xyz), vu.copy(), np.zeros(m.n_verts), m.face_xyz.copy(), fz.copy(), fu.copy(),
                      np.zeros(m.n_faces)]
            ic_got = [a.copy() for a in ic_ref]
            oracle.ic2d_rk2_step(0.01, 2 * np.pi, 0.0, *ic_ref, m.face_area, m.face_mask, n_steps=2)
            e.ic2d_rk2_step(0.01, 2 * np.pi, 0.0, *ic_got, m.face_area, m.face_mask, n_steps=2)
            for k, sel, tol in ((0, None, 1e-12), (1, None, 1e-10), (2, None, 1e-12), (4, leaf, 1e-12), (5, leaf, 1e-10), (6, leaf, 1e-12)):
                assert field_rel_err(ic_got[k], ic_ref[k], sel) <= tol, ("ic2d", seed, k, field_rel_err(ic_got[k], ic_ref[k], sel))
        finally:
            e.close()
        for k, sel, tol in ((0, None, 1e-12), (1, None, 1e-10), (2, None, 1e-12), (3, leaf, 1e-12), (4, leaf, 1e-10), (5, leaf, 1e-12)):
            assert field_rel_err(got[k], ref[k], sel) <= tol, (seed, k, field_rel_err(got[k], ref[k], sel))


@pytest.mark.gpu
def test_gpu_const_stream_captured_sequence_is_bit_identical(monkeypatch):
    """The launch sequence of an evaluation replayed from its CUDA graph (captured the second time a sequence comes by) against
    the same sequence enqueued launch by launch (LPMX_CONST_GRAPH=0), with and without the prefetch warp, with and without
    pipelined launches: the same kernels on the same data in the same order -- bit-identical states after three BVERK4 steps,
    and the same launch counts."""
    from lpm_b200 import gallery
    from lpm_b200.api import BVESolver, Engine, PolyMesh2d
    monkeypatch.setenv("LPMX_CONST_MIN_TARGETS", "1")
    m = PolyMesh2d("cubed", 5)
    f = gallery.RossbyHaurwitz54()
    f.set_stationary_wave_speed()
    vz, fz = f(m.vert_xyz), f(m.face_xyz)
    area, mask = np.ascontiguousarray(m.face_area), np.ascontiguousarray(m.face_mask)
    runs = []
    for graph, prefetch, pdl in (("0", "64", "0"), ("1", "64", "0"), ("1", "0", "0"), ("0", "128", "1"), ("1", "128", "1")):
        monkeypatch.setenv("LPMX_CONST_GRAPH", graph)
        monkeypatch.setenv("LPMX_CONST_PREFETCH", prefetch)
        monkeypatch.setenv("LPMX_CONST_PDL", pdl)
        e = Engine(0)
        try:
            e.pair_sum_const_stream(1)
            s = BVESolver(e, m.n_verts, m.n_faces)
            s.set_state(m.vert_xyz, vz, None, m.face_xyz, fz, None, area, mask)
            s.init_velocity()
            l0, c0 = e.launch_count(), e.const_stream_launch_count()
            for _ in range(3):
                s.advance(0.01, 2 * np.pi, 1)
            counts = (e.launch_count() - l0, e.const_stream_launch_count() - c0)
            out = [np.empty((m.n_verts, 3)), np.empty(m.n_verts), np.empty((m.n_verts, 3)), np.empty((m.n_faces, 3)),
                   np.empty(m.n_faces), np.empty((m.n_faces, 3))]
            s.get_state(*out)
            s.close()
            runs.append((counts, out))
        finally:
            e.close()
    # runs 0-2: whole waves (+ remainder); runs 3-4: pipelined launches, every target through the banks -- two different
    # splits of the targets (the same per-target sums either way: a target's terms are added in the same order by whichever
    # kernel has it only within a split, so bit-identity is asserted within each group)
    assert runs[0][0] == runs[1][0] == runs[2][0] and runs[0][0][1] >= 12 * 5 and runs[3][0] == runs[4][0]
    for k in range(6):
        assert np.array_equal(runs[0][1][k], runs[1][1][k]) and np.array_equal(runs[0][1][k], runs[2][1][k]), k
        assert np.array_equal(runs[3][1][k], runs[4][1][k]), k


@pytest.mark.gpu
def test_gpu_const_stream_at_icos7_sampled_targets(oracle, monkeypatch):
    """600 742 targets x 327 680 leaf sources through the constant banks, 256 launches of 1 280 sources with the other bank
    refilled behind them: 2 048 sampled vertex targets against the oracle, all vertex targets against the default kernel, and a
    second handle on the same device keeps the default kernel (one pair of banks per device)."""
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
        assert n1 > 250 and n2 < 20
        assert np.array_equal(got2, got0)
        idx = np.sort(np.random.default_rng(5).choice(m.n_verts, 2048, replace=False))
        ref = oracle.bve_velocity(m.vert_xyz[idx], m.face_xyz, fz, m.face_area, m.face_mask)
        assert field_rel_err(got1[idx], ref) <= 1e-12
        assert field_rel_err(got1, got0) <= 1e-12
    finally:
        e0.close()
        e1.close()
        e2.close()
