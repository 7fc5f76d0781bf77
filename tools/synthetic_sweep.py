"""BASELINE.json configs[4]: synthetic direct-sum Biot-Savart sweep, N collocated particles (targets = sources, no mask),
through the resident BVERK4 solver (4 velocity evaluations + fused stage updates per step).

    python tools/synthetic_sweep.py [--sizes 1e4,3e4,1e5,3e5,1e6] [--steps 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/synthetic_sweep.py ...

Inputs: points = normalised i.i.d. N(0,1)^3 from the counter-based Philox generator with key 20261017,
zeta_j = 2 (2 pi / 14) z_j + 30 cos(4 lambda_j) z_j (z_j^2 - 1)^2 (the RH54 formula), A_j = 4 pi / N, I = N (N - 1) per
evaluation.  One JSON line per size on stdout (rank 0): interactions/s (max over ranks of the CUDA-event time), RK4 step
time, algorithmic TFLOP/s at 24 flop per interaction and its ratio to the FP64 DFMA peak measured in the same process.
A 64-target subset of the initial velocity is checked against a long-double host sum of the same formula (no oracle/
import: this is a tool, not a test)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def particles(n):
    g = np.random.Generator(np.random.Philox(key=20261017))
    x = g.standard_normal((n, 3))
    x /= np.linalg.norm(x, axis=1)[:, None]
    lam = np.arctan2(x[:, 1], x[:, 0])
    z = x[:, 2]
    zeta = 2 * (2 * np.pi / 14) * z + 30 * np.cos(4 * lam) * z * (z * z - 1) ** 2
    return np.ascontiguousarray(x), zeta, np.full(n, 4 * np.pi / n)


def subset_check(x, zeta, area, vel, n_check=64):
    """A few targets against a long-double host sum of u_i = sum_j (x_i x y_j)(-zeta_j A_j)/(4 pi (1 - x_i.y_j)).
    Returns (max relative error, conditioning bound).  For i.i.d. points the nearest pairs have d = 1 - x.y ~ 1/N, and
    rounding x.y to double perturbs d by 2^-53 whatever the evaluation order, i.e. u by |Gamma||x x y| 2^-53 / d^2: the
    bound sums that over j.  Errors at the bound are the formula's conditioning (the reference's double-precision
    evaluation carries the same), not the kernel's; on the quasi-uniform meshes the bound is ~1e-15."""
    n = x.shape[0]
    idx = np.linspace(0, n - 1, n_check).astype(np.int64)
    xl = x.astype(np.longdouble)
    g = (-(zeta * area) / (4 * np.pi)).astype(np.longdouble)
    worst, bound = 0.0, 0.0
    scale = float(np.abs(vel).max())
    for i in idx:
        d = 1 - xl @ xl[i]
        d[i] = 1  # self pair skipped by index
        w = g / d
        w[i] = 0
        m = (w[:, None] * xl).sum(axis=0)
        u = np.cross(xl[i], m)
        worst = max(worst, float(np.abs(u - vel[i]).max()) / scale)
        c = np.linalg.norm(np.cross(x[i], x), axis=1)
        bound = max(bound, float((np.abs(w) * c / d).sum()) * 2.0 ** -53 / scale)
    return worst, bound


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1e4,3e4,1e5,3e5,1e6")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--dt", type=float, default=1e-4)
    ap.add_argument("--eval-only-above", type=float, default=4e6,
                    help="sizes above this time ONE velocity evaluation (BVESphere::init_velocity on the resident state) instead of "
                         "RK4 steps: N = 1e7 is 1e14 interactions, a minute per evaluation on one GPU")
    ap.add_argument("--n-check", type=int, default=64, help="targets of the long-double host check (rank 0; 64 x 1e7 takes a minute)")
    args = ap.parse_args()
    import torch
    from lpm_b200.api import BVESolver, Engine
    from lpm_b200.dist import env_rank_world, init_engine_comm

    rank, world, local = env_rank_world()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    init_engine_comm(eng, rank, world)
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local))
    peak = eng.fp64_peak_tflops()
    for tok in args.sizes.split(","):
        n = int(float(tok))
        x, zeta, area = particles(n)
        mask = np.zeros(n, dtype=np.uint8)
        s = BVESolver(eng, 0, n)
        s.set_state(None, None, None, x, zeta, None, area, mask)
        eval_only = n > args.eval_only_above
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record(stream)
            s.init_velocity()
            v1.record(stream)
        eng.sync()
        eval_ms = v0.elapsed_time(v1)
        vel = np.zeros((n, 3))
        s.get_state(None, None, None, None, None, vel)
        err, bound = subset_check(x, zeta, area, vel, args.n_check) if rank == 0 else (0.0, 0.0)
        if dist is not None:
            dist.barrier()  # rank 0 alone ran the host-side check: do not let the others wait for it inside a kernel
        if eval_only:
            if dist is not None:
                t = torch.tensor([eval_ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                eval_ms = float(t.item())
            if rank == 0:
                rate = n * (n - 1.0) / (eval_ms * 1e-3)
                print(json.dumps({"workload": "synthetic_collocated", "n_particles": n, "n_gpus": world, "timed": "one velocity evaluation "
                                  "(init_velocity on the resident state, first call)", "eval_ms": eval_ms,
                                  "interactions_per_s": rate, "alg_tflops": rate * 24e-12,
                                  "frac_of_measured_fp64_peak_per_gpu": rate * 24e-12 / (peak * world),
                                  "fp64_peak_tflops_measured": peak, "bank_launches": eng.const_stream_launch_count(),
                                  "velocity_rel_err_vs_float128_subset": err,
                                  "conditioning_bound_of_1_minus_xdoty": bound}), flush=True)
            s.close()
            continue
        s.advance(args.dt, 2 * np.pi, 1)  # warm-up step
        eng.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            s.advance(args.dt, 2 * np.pi, args.steps)
            e1.record(stream)
        eng.sync()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        ms /= args.steps
        inter = 4.0 * n * (n - 1.0)
        if rank == 0:
            rate = inter / (ms * 1e-3)
            print(json.dumps({"workload": "synthetic_collocated", "n_particles": n, "n_gpus": world, "rk4_step_ms": ms,
                              "interactions_per_s": rate, "alg_tflops": rate * 24e-12,
                              "frac_of_measured_fp64_peak_per_gpu": rate * 24e-12 / (peak * world),
                              "fp64_peak_tflops_measured": peak, "velocity_rel_err_vs_float128_subset": err,
                              "conditioning_bound_of_1_minus_xdoty": bound}), flush=True)
        s.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
