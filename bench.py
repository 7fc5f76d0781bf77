#!/usr/bin/env python
"""bench.py -- the contract benchmark of the direct-sum hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--stepper bve_rk4|ic2d_rk2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU path on the host cores

A "step" is one time step of the stepper over the whole particle set: BVERK4::advance_timestep = 4 velocity
evaluations (default; BASELINE.json's metric names the RK4 step), or Incompressible2DRK2 = 2 evaluations.
metric = FP64 particle-pair interactions per second (one interaction = one (target,
unmasked source, j != i) kernel evaluation; I_eval = (n_v + n_f) n_leaf - n_leaf).

One JSON line on stdout (rank 0).  value: device-resident stepping (state in HBM, CUDA events on the
engine's stream, L2 flushed between steps).  e2e: the same step through the in-place C-ABI entry point
lpmx_bve_rk4_step with PINNED HOST buffers, i.e. H2D of the state + step + D2H of the result in the timed
region.  roofline: the pair-sum kernel alone (CUDA events around each launch) in algorithmic FP64 flops
(24 per BVE interaction) against the FP64 FMA peak measured live by a DFMA probe (MEASURED_PEAKS.json has no
FP64 entry).  cpu_baseline: the reference's CPU path timed on the host cores on a bounded target sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic flops per interaction; ic2d: one 24-flop velocity eval + one 31-flop (u, psi) eval per step
FLOPS_PER_INTERACTION = {"bve_rk4": 24.0, "ic2d_rk2": 27.5, "swe_rk2": 124.0}
EVALS_PER_STEP = {"bve_rk4": 4, "ic2d_rk2": 2, "swe_rk2": 2}

WORKLOADS = {
    # name: (seed, depth, vorticity, description)
    "rh54_cubed7": ("cubed", 7, "rh54", "examples/sphere_rh54: Rossby-Haurwitz 54 on cubedSphereSeed depth 7 "
                                        "(98306 vertices + 131070 faces, 98304 leaf sources)"),
    "rotation_icos4": ("icos", 4, "rotation", "examples/bve_rotation: solid-body rotation on icosTriSphereSeed depth 4"),
    "gauss_icos8": ("icos", 8, "gauss", "examples/sphere_gaussian_vortex on icosTriSphereSeed depth 8 (1.97M particles)"),
    "gauss_icos9": ("icos", 9, "gauss", "examples/sphere_gaussian_vortex on icosTriSphereSeed depth 9 (7.86M particles)"),
    "rh54_cubed6": ("cubed", 6, "rh54", "Rossby-Haurwitz 54 on cubedSphereSeed depth 6"),
    "rh54_cubed5": ("cubed", 5, "rh54", "Rossby-Haurwitz 54 on cubedSphereSeed depth 5 (CPU-sized)"),
    "tc2_icos8": ("icos", 8, "tc2", "examples/sphere_swe_tc2 (Williamson test case 2) on icosTriSphereSeed depth 8"),
    "tc2_cubed7": ("cubed", 7, "tc2", "examples/sphere_swe_tc2 (Williamson test case 2) on cubedSphereSeed depth 7"),
    "tc2_cubed5": ("cubed", 5, "tc2", "Williamson test case 2 on cubedSphereSeed depth 5"),
}


def build_case(workload):
    from lpm_b200 import gallery
    from lpm_b200.api import PolyMesh2d
    seed, depth, vort, desc = WORKLOADS[workload]
    m = PolyMesh2d(seed, depth)
    if vort == "rh54":
        f = gallery.RossbyHaurwitz54()
        f.set_stationary_wave_speed()  # u0 = Omega/14, Omega = 2 pi (examples/sphere_rh54.cpp:108-111)
    elif vort == "tc2":
        f = gallery.SphereTestCase2().vorticity
    elif vort == "rotation":
        f = gallery.SolidBodyRotation()
    else:
        f = gallery.GaussianVortexSphere()
    return m, f(m.vert_xyz), f(m.face_xyz), desc


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.reasons = set()
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append((float(parts[0]), float(parts[1]), float(parts[2])))
                    for n, v in zip(names, parts[3:7]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(self.reasons), "samples": 0}
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1],
                "power_w_max": max(s[2] for s in self.samples), "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def host_threads():
    """Host threads this process may use (cgroup/affinity aware)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_lib():
    """The reference's own functors compiled in place (oracle/_ref) when present, else the restatement.  The OpenMP team size is
    set explicitly: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which round 1's reference arm inherited."""
    import ctypes
    from oracle import oracle
    kind, L = "port", None
    if os.path.exists(oracle.REF_LIB):
        try:
            L = ctypes.CDLL(oracle.REF_LIB)
            kind = "reference"
        except OSError:
            L = None
    if L is None:
        L = oracle.lib()
    L.oracle_num_threads.restype = ctypes.c_int
    try:
        L.oracle_set_num_threads(ctypes.c_int(host_threads()))
    except AttributeError:
        pass
    return oracle, L, kind


def cpu_sample_indices(m, n_sample):
    """A bounded sample of the workload's targets in the mesh's own proportion of vertices (BVEVertexVelocity: distinct targets)
    and faces (BVEFaceVelocity: collocated targets with the i != j branch), evenly spaced over each list."""
    nt = m.n_verts + m.n_faces
    n_sample = max(2, min(n_sample, nt))
    nvs = max(1, min(m.n_verts, int(round(n_sample * m.n_verts / nt))))
    nfs = max(1, min(m.n_faces, n_sample - nvs))
    vi = np.unique(np.linspace(0, m.n_verts - 1, nvs).astype(np.int64))
    fi = np.unique(np.linspace(0, m.n_faces - 1, nfs).astype(np.int32))
    return vi, fi


def time_cpu_sample(m, fz, n_sample, reps=1):
    """One velocity evaluation of a bounded target sample against all sources with the reference CPU path (OpenMP over
    targets, sequential j, per-pair divide).  Returns (interactions/s, seconds, kind, threads, n_vert_targets, n_face_targets)."""
    oracle, L, kind = cpu_reference_lib()
    vi, fi = cpu_sample_indices(m, n_sample)
    tx = np.ascontiguousarray(m.vert_xyz[vi])
    if not getattr(time_cpu_sample, "_warm", False):
        # the first few parallel regions of a process run several times slower (thread-pool start-up)
        w = np.ascontiguousarray(m.vert_xyz[:max(64, min(len(vi) // 16, 2048))])
        for _ in range(4):
            oracle.bve_velocity(w, m.face_xyz, fz, m.face_area, m.face_mask, collocated=False, L=L)
        time_cpu_sample._warm = True
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.bve_velocity(tx, m.face_xyz, fz, m.face_area, m.face_mask, collocated=False, L=L)
        oracle.bve_velocity_subset(fi, m.face_xyz, fz, m.face_area, m.face_mask, L=L)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    inter = float(len(vi) + len(fi)) * m.n_face_leaves - float((m.face_mask[fi] == 0).sum())
    return inter / best, best, kind, L.oracle_num_threads(), len(vi), len(fi)


def size_cpu_sample(m, fz, seconds):
    """Targets that take about `seconds` per evaluation on this host, from a short calibration run."""
    rate, _, _, _, _, _ = time_cpu_sample(m, fz, 2048)
    rate2, _, _, _, _, _ = time_cpu_sample(m, fz, 2048)
    n = int(max(rate, rate2) * seconds / max(1, m.n_face_leaves))
    return max(256, min(n, m.n_verts + m.n_faces))


def time_reference_icos4_example():
    """BASELINE.md section 4's CPU-runnable case: examples/bve_rotation-style run at icosTriSphereSeed depth 4, 3 BVERK4 steps of
    solid-body rotation, with the REFERENCE's own BVESphere + BVERK4::advance_timestep compiled in place (oracle/_ref/
    liblpm_ref_mesh.so) when present, else the oracle's restatement of the step.  Returns a dict or None."""
    from lpm_b200 import gallery
    from lpm_b200.api import PolyMesh2d
    try:
        m = PolyMesh2d("icos", 4)
        f = gallery.SolidBodyRotation()
        vz, fz = f(m.vert_xyz), f(m.face_xyz)
        inter = 3 * 4 * (float(m.n_verts + m.n_faces) * m.n_face_leaves - m.n_face_leaves)
        from oracle import ref_mesh
        if ref_mesh.available():
            import ctypes
            ref_mesh.lib().ref_mesh_set_num_threads(ctypes.c_int(host_threads()))
            ref_mesh.bve_rk4_run("icos", 4, 0.0025, 0.0, 0, vz, fz)  # warm-up + what is not the stepper (mesh, init_velocity)
            t0 = time.perf_counter()
            ref_mesh.bve_rk4_run("icos", 4, 0.0025, 0.0, 0, vz, fz)
            t_init = time.perf_counter() - t0
            t0 = time.perf_counter()
            ref_mesh.bve_rk4_run("icos", 4, 0.0025, 0.0, 3, vz, fz)
            secs = max(time.perf_counter() - t0 - t_init, 1e-9)
            kind, threads = "reference", ref_mesh.lib().ref_mesh_num_threads()
            what = "BVESphere<IcosTriSphereSeed> + 3 x BVERK4::advance_timestep compiled in place (tree_init and init_velocity subtracted)"
        else:
            oracle, L, kind = cpu_reference_lib()
            L = oracle.lib()
            L.oracle_set_num_threads(host_threads())
            vu = oracle.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
            fu = oracle.bve_velocity(None, m.face_xyz, fz, m.face_area, m.face_mask, collocated=True)
            st = [m.vert_xyz.copy(), vz.copy(), vu, m.face_xyz.copy(), fz.copy(), fu]
            t0 = time.perf_counter()
            oracle.bve_rk4_step(0.0025, 0.0, *st, m.face_area, m.face_mask, n_steps=3)
            secs = time.perf_counter() - t0
            kind, threads = "port", L.oracle_num_threads()
            what = "oracle_bve_rk4_step x 3 (the restatement pinned against the compiled BVERK4)"
        return {"workload": "rotation_icos4, 3 BVERK4 steps (BASELINE configs[0])", "seconds": secs, "ms_per_step": secs / 3 * 1e3,
                "interactions_per_s": inter / secs, "kind": kind, "cores": threads, "what": what}
    except Exception as e:  # a baseline extra must never take the bench line down
        return {"error": repr(e)}


def workload_config(args, m, desc, world):
    """The `config` object, identical for both arms (the driver compares them)."""
    evals = EVALS_PER_STEP[args.stepper]
    i_eval = float(m.n_verts + m.n_faces) * m.n_face_leaves - m.n_face_leaves
    cfg = {"workload": args.workload, "description": desc, "stepper": args.stepper, "evals_per_step": evals,
           "n_verts": m.n_verts, "n_faces": m.n_faces, "n_leaf_sources": m.n_face_leaves, "interactions_per_eval": i_eval,
           "dt": args.dt, "Omega": 2 * np.pi,
           "parallelism": f"targets sharded over {world} GPU(s), per-stage allgather of leaf source records",
           "l2": "flushed between timed steps (256 MiB write)"}
    if args.stepper == "swe_rk2":
        cfg["surface_laplacian"] = (f"device GMLS order {args.gmls_order}, both stages" if args.laplacian == "gmls"
                                    else "frozen at the TC2 closed form")
    return cfg


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores (oracle/_ref: its own functors compiled
    in place), all host threads, rank 0 only.  Each step is ONE velocity evaluation of a bounded sample of the workload's targets
    (vertices through BVEVertexVelocity, faces through BVEFaceVelocity) against all sources, sized by a calibration run to
    --ref-seconds per step, so that the whole --steps K --warmup W run stays within a few minutes whatever the host."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if world > 1:
        args.gpus = world
    m, vz, fz, desc = build_case(args.workload)
    if args.dt is None:
        args.dt = 0.025 * m.appx_mesh_size() / 0.09045016
    evals = EVALS_PER_STEP[args.stepper]
    n_sample = args.cpu_sample if args.cpu_sample > 0 else size_cpu_sample(m, fz, args.ref_seconds)
    for _ in range(max(args.warmup, 0)):
        time_cpu_sample(m, fz, max(n_sample // 8, 64))
    rates, secs = [], []
    kind, threads, nvs, nfs = "port", 1, 0, 0
    for _ in range(args.steps):
        r, sec, kind, threads, nvs, nfs = time_cpu_sample(m, fz, n_sample)
        rates.append(r)
        secs.append(sec)
    value = float(np.sum(rates) / len(rates))
    cfg = workload_config(args, m, desc, args.gpus)
    i_eval = cfg["interactions_per_eval"]
    sample = (f"per step ONE velocity evaluation of {nvs} vertex targets (BVEVertexVelocity) + {nfs} face targets "
              f"(BVEFaceVelocity, i != j) x all {m.n_faces} faces ({m.n_face_leaves} leaf sources); ms_per_step is the "
              f"timed sample, not a stepper step")
    line = {
        "impl": "reference", "metric": "fp64_pair_interactions_per_s", "value": value, "unit": "interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(secs)) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "full_step_ms_extrapolated": evals * i_eval / value * 1e3,
        "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "native_code_note": "lpm_b200/liblpmx.so is loaded by this arm for the HOST mesh generator only (input generation; "
                            "no kernel runs, no device is opened); the timed code is oracle/_ref",
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rh54_cubed7", choices=sorted(WORKLOADS))
    ap.add_argument("--stepper", default="bve_rk4", choices=["bve_rk4", "ic2d_rk2", "swe_rk2"])
    ap.add_argument("--dt", type=float, default=None,
                    help="time step; default 0.025 * h / h(cubed-4): the reference's sphere_rh54 default (tfinal 0.025, "
                         "1 step, depth 4; examples/sphere_rh54.cpp:442-452) at constant Courant number")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="targets (vertices + faces, in the mesh's proportion) in the CPU sample; 0 = sized by a calibration "
                         "run to --ref-seconds (reference arm) / --cpu-seconds (cpu_baseline leg) per evaluation")
    ap.add_argument("--ref-seconds", type=float, default=6.0, help="reference arm: host seconds per timed step")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="cpu_baseline leg: host seconds per evaluation (timed twice)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra records of the N = 1 line (n1m synthetic set, ic2d_rk2 stepper, icos-4 CPU example)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run parity block (sampled targets vs the oracle)")
    ap.add_argument("--laplacian", default="frozen", choices=["frozen", "gmls"],
                    help="swe_rk2 only: surface Laplacian frozen at the TC2 closed form (the pair sums alone), or the "
                         "device-side GMLS provider of order --gmls-order evaluated at both stages of every step (the whole "
                         "reference step, src/lpm_swe_rk2_impl.hpp:134-154,233-252)")
    ap.add_argument("--gmls-order", type=int, default=4, help="examples/sphere_swe_tc2.cpp:136 uses 4")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE JSON line: anything a library prints there while we run (NCCL's version banner
    # under NCCL_DEBUG=VERSION, for one) is sent to stderr; the line itself goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    from lpm_b200.api import BVESolver, Engine, IC2DSolver, SWESolver
    from lpm_b200.dist import env_rank_world, init_engine_comm

    rank, world, local_rank = env_rank_world()
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # keep stdout to the one JSON line: NCCL's version / debug lines go to a file unless the user chose one
        import tempfile
        os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(tempfile.gettempdir(), "lpmx_nccl.%h.%p.log"))
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = Engine(local_rank)
    init_engine_comm(eng, rank, world)
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local_rank))

    m, vz, fz, desc = build_case(args.workload)
    nv, nf, nleaf = m.n_verts, m.n_faces, m.n_face_leaves
    if args.dt is None:
        args.dt = 0.025 * m.appx_mesh_size() / 0.09045016  # Courant number of the reference default (~0.4)
    evals = EVALS_PER_STEP[args.stepper]
    i_eval = float(nv + nf) * nleaf - nleaf
    Omega = 2 * np.pi
    area = np.ascontiguousarray(m.face_area)
    mask = np.ascontiguousarray(m.face_mask)

    if args.stepper == "bve_rk4":
        solver = BVESolver(eng, nv, nf)
        solver.set_state(m.vert_xyz, vz, None, m.face_xyz, fz, None, area, mask)
        solver.init_velocity()
    elif args.stepper == "ic2d_rk2":
        solver = IC2DSolver(eng, nv, nf, eps=0.0)
        solver.set_state(m.vert_xyz, vz, None, m.face_xyz, fz, None, area, mask)
        solver.init_direct_sums()
    else:
        # SWE fields of examples/sphere_swe_tc2.cpp; the (out-of-path) GMLS Laplacian is replaced by the closed-form
        # Laplacian of the TC2 surface, frozen at the initial particle positions
        from lpm_b200 import gallery as _g
        tc = _g.SphereTestCase2()
        swe_p = {"xyz": m.vert_xyz, "vort": vz, "div": np.zeros(nv), "depth": tc.surface(m.vert_xyz),
                 "surf": tc.surface(m.vert_xyz), "bottom": np.zeros(nv), "laps": tc.surface_laplacian_exact(m.vert_xyz)}
        swe_a = {"xyz": m.face_xyz, "vort": fz, "div": np.zeros(nf), "area": area, "mass": tc.surface(m.face_xyz) * area,
                 "depth": tc.surface(m.face_xyz), "surf": tc.surface(m.face_xyz), "bottom": np.zeros(nf),
                 "laps": tc.surface_laplacian_exact(m.face_xyz)}
        swe_p = {k: np.ascontiguousarray(v) for k, v in swe_p.items()}
        swe_a = {k: np.ascontiguousarray(v) for k, v in swe_a.items()}
        swe_solver = SWESolver(eng, nv, nf, eps=0.0)
        swe_solver.set_state(swe_p, swe_a, mask)
        swe_solver.init_direct_sums(True)
        swe_g = tc.g
        swe_lap = eng.gmls_provider(args.gmls_order) if args.laplacian == "gmls" else None

        class _Adv:  # same advance(dt, Omega, n) surface as the other two solvers
            def advance(self, dt, Omega, n):
                swe_solver.advance(dt, Omega, swe_g, swe_lap, n)
        solver = _Adv()
    eng.sync()
    fp64_peak = eng.fp64_peak_tflops()

    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(args.warmup):
        solver.advance(args.dt, Omega, 1)
    eng.sync()

    # ---- timed: K device-resident steps, L2 flushed between steps, events on the engine's stream ----
    sampler = ClockSampler(local_rank)
    launches0 = eng.launch_count()
    cs_launches0 = eng.const_stream_launch_count()
    eng.profile_enable(True)
    eng.profile_read()
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    evs = []
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush_buf.fill_(1)  # untimed L2 flush
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            solver.advance(args.dt, Omega, 1)
            e1.record(stream)
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))
    n_k, k_ms, k_pairs = eng.profile_read()
    eng.profile_enable(False)
    launches = eng.launch_count() - launches0
    cs_launches = eng.const_stream_launch_count() - cs_launches0
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = evals * i_eval / (ms_per_step * 1e-3)

    # ---- sanity of the advanced state (a blown-up run would still time the same): finite, still on the sphere ----
    chk = [np.zeros((nv, 3)), np.zeros(nv), np.zeros((nv, 3)), np.zeros((nf, 3)), np.zeros(nf), np.zeros((nf, 3))]
    if args.stepper == "bve_rk4":
        solver.get_state(*chk)
    elif args.stepper == "ic2d_rk2":
        solver.get_state(chk[0], chk[1], chk[2], None, chk[3], chk[4], chk[5], None)
    else:
        swe_solver.get_state({"xyz": chk[0], "vort": chk[1], "vel": chk[2]}, {"xyz": chk[3], "vort": chk[4], "vel": chk[5]})
    leafsel = mask == 0
    state_check = {
        "finite": bool(np.isfinite(chk[0]).all() and np.isfinite(chk[3][leafsel]).all() and np.isfinite(chk[5][leafsel]).all()),
        "max_abs_radius_minus_1": float(np.abs(np.linalg.norm(chk[3][leafsel], axis=1) - 1).max()),
        "max_speed": float(np.linalg.norm(chk[5][leafsel], axis=1).max()),
        "steps_advanced": args.warmup + args.steps,
    }

    # ---- parity of what the timed steps left behind (rank 0; the oracle is the CHECKER here, outside every timed region):
    # the velocity the last evaluation stored, against the reference arithmetic (oracle/_ref when present) evaluated on the
    # same advanced state, on sampled vertex targets (distinct) and sampled leaf-face targets (collocated, i != j).  With
    # N > 1 the state was gathered from all ranks' shards, so this is the multi-GPU parity record of the SCALE runs.
    parity = None
    if rank == 0 and not args.no_parity:
        try:
            oracle, Lref, pkind = cpu_reference_lib()
            rng = np.random.default_rng(20261018)
            vi = np.sort(rng.choice(nv, min(1024, nv), replace=False))
            leaf_ids = np.nonzero(leafsel)[0]
            fi = np.sort(rng.choice(leaf_ids, min(1024, len(leaf_ids)), replace=False)).astype(np.int32)
            if args.stepper == "swe_rk2":
                sdiv_p, sdiv_a, sarea = np.zeros(nv), np.zeros(nf), np.zeros(nf)
                swe_solver.get_state({"div": sdiv_p}, {"div": sdiv_a, "area": sarea})
                ov, _, _ = oracle.swe_sphere_sums(chk[0][vi], chk[3], chk[4], sdiv_a, sarea, mask, eps=0.0, L=Lref)
                of, gf = None, None
            else:
                ov = oracle.bve_velocity(chk[0][vi], chk[3], chk[4], area, mask, L=Lref)
                of = oracle.bve_velocity_subset(fi, chk[3], chk[4], area, mask, L=Lref)
            scale = float(np.linalg.norm(ov, axis=1).max())
            ev = float(np.linalg.norm(chk[2][vi] - ov, axis=1).max() / scale)
            adjud = None
            if args.stepper != "swe_rk2":
                # whose round-off is it: both against a long-double sum of the same formula on the same state (vertex sample)
                old = oracle.bve_velocity(chk[0][vi], chk[3], chk[4], area, mask, long_double=True)
                adjud = {"engine_vs_long_double": float(np.linalg.norm(chk[2][vi] - old, axis=1).max() / scale),
                         "reference_fp64_vs_long_double": float(np.linalg.norm(ov - old, axis=1).max() / scale)}
            ef = float(np.linalg.norm(chk[5][fi] - of, axis=1).max() / scale) if of is not None else None
            parity = {"quantity": "velocity stored by the last evaluation of the timed steps vs the reference arithmetic on the "
                                  "same advanced state (field-relative max-norm)",
                      "max_rel_err": max(ev, ef) if ef is not None else ev, "vertex_targets": {"n": int(len(vi)), "rel_err": ev},
                      "leaf_face_targets": ({"n": int(len(fi)), "rel_err": ef} if ef is not None else None),
                      "adjudication_vertex_sample": adjud, "tolerance": 1e-12, "checker": "oracle/_ref (reference functors compiled in place)" if pkind == "reference"
                      else "oracle/ (C restatement)", "n_gpus_that_produced_the_state": world}
        except Exception as e:
            parity = {"error": repr(e)}

    # ---- roofline of the dominant kernel (rank-local): algorithmic flops / CUDA-event launch time ----
    if args.stepper == "swe_rk2":  # the SWE solver shards the concatenated list; BVE / IC2D shard leaves and non-sources separately
        n_local_targets = solver_local_targets(nv + nf, rank, world)
    else:
        n_local_targets = solver_local_targets(nleaf, rank, world) + solver_local_targets(nv + nf - nleaf, rank, world)
    local_inter = evals * args.steps * (float(n_local_targets) * nleaf)
    flops_per = FLOPS_PER_INTERACTION[args.stepper]
    achieved_tf = local_inter * flops_per / (k_ms * 1e-3) * 1e-12 if k_ms > 0 else None
    roofline = {
        "bound": "fp64", "kernel": "lpmx::pair_sum_kernel", "achieved": achieved_tf, "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": (achieved_tf / fp64_peak) if achieved_tf else None,
        "peak_source": "measured live: lpmx_fp64_peak_tflops, DFMA R, R, c[0][..], R probe, loop unrolled x16 (MEASURED_PEAKS.json "
                       "has no FP64 entry; nominal 148 SM x 64 FMA x 2 x 1.965 GHz = 37.2)",
        "flops_per_interaction": flops_per, "launches": n_k, "avg_launch_ms": (k_ms / n_k) if n_k else None,
        "kernel_share_of_step": (k_ms / (sum(step_ms))) if step_ms else None,
        "fp64_pipe_instr_per_interaction": {"bve_rk4": 9, "ic2d_rk2": 13.5, "swe_rk2": 53}[args.stepper],  # ic2d: (9 + 18) / 2
        "traffic": None,
    }
    if cs_launches > 0:
        # the velocity sums of this run went through the constant banks: `launches` above counts evaluations
        # (each = one sequence of pipelined bank launches, timed as a whole)
        roofline["kernel"] = "lpmx::pair_sum_const_kernel (sources through the constant banks as uniform-register operands; bank launches pipelined, one CUDA graph launch per evaluation)"
        roofline["evaluations"] = n_k
        roofline["bank_launches"] = int(cs_launches)
        roofline["avg_bank_launch_ms"] = (k_ms / cs_launches) if cs_launches else None
    # what the FP64 pipe actually issued (2 flop per DFMA) against the same measured peak: the number that corresponds to
    # ncu's sm__pipe_fp64_cycles_active, and the reason `frac` can exceed 1 (9 DFMAs do the work of the 24-flop reference pair)
    if achieved_tf:
        instr = roofline["fp64_pipe_instr_per_interaction"]
        roofline["issued_tflops"] = achieved_tf * (2.0 * instr) / flops_per
        roofline["issued_frac"] = roofline["issued_tflops"] / fp64_peak if fp64_peak else None
    prof = os.path.join(ROOT, "profiles", "r2_pair_sum_const_dram.json" if cs_launches > 0 else "r2_pair_sum_dram.json")
    if os.path.exists(prof) and args.workload == "rh54_cubed7" and args.stepper == "bve_rk4" and world == 1:  # captured on that launch shape
        try:
            rec = json.load(open(prof))
            roofline["traffic"] = rec.get("dram_bytes_per_launch")
            roofline["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one launch of this shape from the ncu --set "
                                          f"full capture committed as profiles/{os.path.basename(prof)} ({rec.get('tag', 'r2')}); NOT measured in this run")
        except Exception:
            pass

    # ---- e2e: in-place C-ABI call on pinned host buffers (H2D + step + D2H inside the timed region) ----
    e2e = None
    if True:
        # N = 1: the full state goes up and comes back every step.  N > 1: sharded host I/O (lpmx_set_io_sharded) -- each rank's
        # host arrays carry its own target rows, as in any distributed-memory program: it uploads those rows (+ area and mask
        # of all faces) and downloads those rows; nothing is replicated through the host.
        sharded_io = world > 1 and args.stepper in ("bve_rk4", "ic2d_rk2")
        own_a, own_b = eng.local_targets(nv, nf, mask)
        n_own = len(own_a) + len(own_b)

        def pin(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t
        if args.stepper == "bve_rk4":
            st_np = [m.vert_xyz.copy(), vz.copy(), np.zeros((nv, 3)), m.face_xyz.copy(), fz.copy(), np.zeros((nf, 3))]
            solver.get_state(*st_np)
            host = [pin(a) for a in st_np]
            h_area, h_mask = pin(area), pin(mask)
            args_np = [t.numpy() for t in host]

            def one():
                eng.bve_rk4_step(args.dt, Omega, *args_np, h_area.numpy(), h_mask.numpy(), n_steps=1)
            h2d = sum(t.numel() * 8 for t in host) + area.nbytes + mask.nbytes
            d2h = sum(t.numel() * 8 for t in host)
            if sharded_io:
                h2d, d2h = n_own * 7 * 8 + area.nbytes + mask.nbytes, n_own * 7 * 8
        elif args.stepper == "swe_rk2":
            from lpm_b200.api import ACTIVE_FIELDS, PASSIVE_FIELDS, swe_rk2_step
            hp = {k: pin(np.zeros((nv, 3)) if k in ("xyz", "vel") else np.zeros(nv)) for k in PASSIVE_FIELDS}
            ha = {k: pin(np.zeros((nf, 3)) if k in ("xyz", "vel") else np.zeros(nf)) for k in ACTIVE_FIELDS}
            h_mask = pin(mask)
            hp_np, ha_np = {k: t.numpy() for k, t in hp.items()}, {k: t.numpy() for k, t in ha.items()}
            swe_solver.get_state(hp_np, ha_np)

            def one():
                swe_rk2_step(eng, args.dt, Omega, swe_g, 0.0, hp_np, ha_np, h_mask.numpy(), swe_lap, n_steps=1)
            h2d = sum(t.numel() * 8 for t in hp.values()) + sum(t.numel() * 8 for t in ha.values()) + mask.nbytes
            d2h = h2d - mask.nbytes
        else:
            st_np = [m.vert_xyz.copy(), vz.copy(), np.zeros((nv, 3)), np.zeros(nv), m.face_xyz.copy(), fz.copy(),
                     np.zeros((nf, 3)), np.zeros(nf)]
            solver.get_state(*st_np)
            host = [pin(a) for a in st_np]
            h_area, h_mask = pin(area), pin(mask)
            args_np = [t.numpy() for t in host]

            def one():
                eng.ic2d_rk2_step(args.dt, Omega, 0.0, *args_np, h_area.numpy(), h_mask.numpy(), n_steps=1)
            h2d = sum(t.numel() * 8 for t in host) - 8 * (nv + nf) + area.nbytes + mask.nbytes
            d2h = sum(t.numel() * 8 for t in host)
            if sharded_io:
                h2d, d2h = n_own * 7 * 8 + area.nbytes + mask.nbytes, n_own * 8 * 8
        if sharded_io:
            eng.set_io_sharded(True)
        one()  # warm-up: allocates the cached solver and the staging buffers
        torch.cuda.synchronize()
        t_calls = 0.0
        for _ in range(args.steps):
            flush_buf.fill_(1)  # L2 flush between calls, outside the timed call
            barrier()
            c0 = time.perf_counter()
            one()  # returns when the results are back in the host buffers
            t_calls += time.perf_counter() - c0
        if dist is not None:
            t = torch.tensor([t_calls], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_calls = float(t.item())
        e2e = {"value": evals * i_eval * args.steps / t_calls, "unit": "interactions/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": t_calls / args.steps * 1e3,
               "api": {"bve_rk4": "lpmx_bve_rk4_step", "ic2d_rk2": "lpmx_ic2d_rk2_step", "swe_rk2": "lpmx_swe_rk2_step"}[args.stepper],
               "host_buffers": "pinned",
               "host_io": ("sharded: each rank moves its own target rows (lpmx_set_io_sharded); bytes are per rank" if sharded_io
                           else "full state per rank")}
        if sharded_io:
            eng.set_io_sharded(False)

    # ---- extras of the N = 1 line: the synthetic N = 1e6 set (north_star's ">= 1M particles"), and the stepper the reference's
    # sphere_rh54 / sphere_gaussian_vortex drivers actually use (Incompressible2DRK2) on the same mesh ----
    extras = {}
    if world == 1 and not args.no_extras and args.stepper == "bve_rk4":
        extras["n1m"] = bench_synthetic_n1m(eng, stream, torch, fp64_peak)
    if not args.no_extras and args.stepper == "bve_rk4":
        extras["ic2d_rk2"] = bench_ic2d_extra(eng, stream, torch, dist, m, vz, fz, area, mask, args.dt, Omega, i_eval, flush_buf)

    # ---- CPU baseline on the host cores (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_sample = args.cpu_sample if args.cpu_sample > 0 else size_cpu_sample(m, fz, args.cpu_seconds)
        rate, secs, kind, threads, nvs, nfs = time_cpu_sample(m, fz, n_sample, reps=2)
        cpu = {"value": rate, "unit": "interactions/s", "cores": threads, "kind": kind,
               "sample": f"{nvs} vertex targets (BVEVertexVelocity) + {nfs} face targets (BVEFaceVelocity, i != j) x all {nf} faces "
                         f"({nleaf} leaf sources), one velocity evaluation, best of 2, {secs:.1f} s each"}
        if not args.no_extras:
            cpu["icos4_3steps"] = time_reference_icos4_example()

    peer_on, peer_slabs = eng.comm_peer_exchange_enabled()
    exchange = ("none (one GPU)" if world == 1 else
                "one kernel storing each rank's records into the peers' slabs over NVLink, on the copy stream beside the pair sum of the "
                "rank's non-source targets (default for world <= 8)"
                if peer_on and peer_slabs else "grouped ncclBroadcast in line (LPMX_PEER_EXCHANGE=0, or peer mapping unavailable)")
    if rank == 0:
        line = {
            "metric": "fp64_pair_interactions_per_s", "value": value, "unit": "interactions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, m, desc, world), "exchange": exchange,
            "rk_step_ms": ms_per_step, "step_ms_each": step_ms, "wall_s_timed_region": t_wall,
            "state_check": state_check,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "parity": parity, **extras,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def bench_synthetic_n1m(eng, stream, torch, fp64_peak, n=1_000_000, steps=3):
    """BASELINE configs[4] at N = 1e6 (tools/synthetic_sweep.py's particle set): `steps` BVERK4 steps of N collocated i.i.d.
    particles, device-resident, CUDA events on the engine's stream.  north_star's ">= 70 % of FP64 peak at N >= 1M"."""
    from lpm_b200.api import BVESolver
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        from synthetic_sweep import particles
        x, zeta, area = particles(n)
        mask = np.zeros(n, dtype=np.uint8)
        s = BVESolver(eng, 0, n)
        s.set_state(None, None, None, x, zeta, None, area, mask)
        s.init_velocity()
        s.advance(1e-4, 2 * np.pi, 1)  # warm-up step
        eng.sync()
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            s.advance(1e-4, 2 * np.pi, steps)
            e1.record(stream)
        eng.sync()
        ms = e0.elapsed_time(e1) / steps
        s.close()
        rate = 4.0 * n * (n - 1.0) / (ms * 1e-3)
        return {"workload": "synthetic_collocated (Philox key 20261017, RH54 vorticity formula)", "n_particles": n, "steps": steps,
                "rk4_step_ms": ms, "interactions_per_s": rate, "alg_tflops": rate * 24e-12,
                "frac_of_measured_fp64_peak": rate * 24e-12 / fp64_peak if fp64_peak else None,
                "issued_frac": rate * 18e-12 / fp64_peak if fp64_peak else None, "inputs": "larger than L2 (64 MB of records per pass)"}
    except Exception as e:
        return {"error": repr(e)}


def bench_ic2d_extra(eng, stream, torch, dist, m, vz, fz, area, mask, dt, Omega, i_eval, flush_buf, steps=3):
    """Incompressible2DRK2 (src/lpm_incompressible2d_rk2_impl.hpp:75-172; 2 evaluations per step, psi with the second) on the
    bench's mesh, one step per call as the reference's drivers call it, device-resident; max over ranks."""
    from lpm_b200.api import IC2DSolver
    try:
        s = IC2DSolver(eng, m.n_verts, m.n_faces, eps=0.0)
        s.set_state(m.vert_xyz, vz, None, m.face_xyz, fz, None, area, mask)
        s.init_direct_sums()
        for _ in range(3):  # warm-up as in the main line (W >= 3): the second sighting of a launch sequence captures its graph
            s.advance(dt, Omega, 1)
        eng.sync()
        if dist is not None:
            dist.barrier()
        total = 0.0
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush_buf.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                s.advance(dt, Omega, 1)
                e1.record(stream)
            eng.sync()
            total += e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        s.close()
        ms = total / steps
        return {"stepper": "Incompressible2DRK2, one step per call", "steps": steps, "ms_per_step": ms,
                "interactions_per_s": 2 * i_eval / (ms * 1e-3), "evals_per_step": 2}
    except Exception as e:
        return {"error": repr(e)}


def solver_local_targets(nt, rank, world):
    return ((rank + 1) * nt) // world - (rank * nt) // world


if __name__ == "__main__":
    sys.exit(main())
