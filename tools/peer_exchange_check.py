"""Peer exchange (lpm_b200/csrc/lpmx_peer.cu) against the NCCL exchange, under torchrun with one rank per GPU:

    LPMX_PEER_TIMEOUT_S=10 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/peer_exchange_check.py [--time cubed-7,icos-4] [--steps 8]

Every process holds two engines on its GPU, one per exchange; both run the same sharded steppers.
  1. parity: the states after 2 steps must be BIT-identical between the exchanges (the same records land in the
     same places; the pair sums do not depend on how they got there) and within 1e-10 of the CPU oracle;
  2. timing: ms per BVERK4 step of a device-resident solver, max over ranks, for each exchange.
Exits non-zero on any mismatch, on a timeout inside the exchange kernel, or when the peer path was not taken.
(Development/measurement tool: the oracle is only the checker here.)"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", default="icos-4,cubed-6,cubed-7")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--skip-oracle", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from conftest import field_rel_err
    from lpm_b200 import gallery
    from lpm_b200.api import BVESolver, Engine, LpmxError, PolyMesh2d
    from lpm_b200.dist import env_rank_world, init_engine_comm

    rank, world, local = env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["LPMX_PEER_EXCHANGE"] = "0"  # the two engines are configured explicitly (the default would turn it on for both)
    e_nccl, e_peer = Engine(local), Engine(local)
    init_engine_comm(e_nccl, rank, world)
    init_engine_comm(e_peer, rank, world)
    try:
        e_peer.comm_enable_peer_exchange(True)
    except LpmxError as ex:
        print(f"[rank {rank}] peer exchange unavailable: {ex}", flush=True)
        return 2
    fails = 0

    def state(m):
        f = gallery.RossbyHaurwitz54()
        f.set_stationary_wave_speed()
        return [m.vert_xyz.copy(), f(m.vert_xyz), np.zeros_like(m.vert_xyz), m.face_xyz.copy(), f(m.face_xyz),
                np.zeros_like(m.face_xyz)]

    # ---- 1. parity -------------------------------------------------------------------------------------------
    for seed, depth in (("icos", 3), ("cubed", 4)):
        m = PolyMesh2d(seed, depth)
        leaf = m.face_mask == 0
        outs = {}
        for name, e in (("nccl", e_nccl), ("peer", e_peer)):
            st = state(m)
            e.bve_rk4_step(0.01, 2 * np.pi, *st, m.face_area, m.face_mask, n_steps=2)
            ic = state(m)
            ic = [ic[0], ic[1], ic[2], np.zeros(ic[0].shape[0]), ic[3], ic[4], ic[5], np.zeros(ic[3].shape[0])]
            e.ic2d_rk2_step(0.01, 2 * np.pi, 0.0, *ic, m.face_area, m.face_mask, n_steps=2)
            e.sync()
            outs[name] = st + ic
        same = all(np.array_equal(a[i] if i not in (3, 4, 5, 10, 11, 12, 13) else a[i][leaf],
                                  b[i] if i not in (3, 4, 5, 10, 11, 12, 13) else b[i][leaf])
                   for a, b in [(outs["nccl"], outs["peer"])] for i in range(len(a)))
        print(f"[rank {rank}/{world}] {seed}-{depth}: peer == nccl bitwise: {same}", flush=True)
        fails += 0 if same else 1
        if not args.skip_oracle:
            from oracle import oracle
            ref = state(m)
            oracle.bve_rk4_step(0.01, 2 * np.pi, *ref, m.face_area, m.face_mask, n_steps=2)
            got = outs["peer"]
            err = max(field_rel_err(got[0], ref[0]), field_rel_err(got[1], ref[1]), field_rel_err(got[3], ref[3], leaf),
                      field_rel_err(got[4], ref[4], leaf))
            print(f"[rank {rank}/{world}] {seed}-{depth}: peer vs oracle {err:.3e}", flush=True)
            fails += 0 if err <= 1e-10 else 1
    en, nreg = e_peer.comm_peer_exchange_enabled()
    print(f"[rank {rank}/{world}] peer path enabled={en} mapped slabs={nreg}", flush=True)
    fails += 0 if (en and nreg >= 1) else 1
    en0, nreg0 = e_nccl.comm_peer_exchange_enabled()
    fails += 0 if (not en0 and nreg0 == 0) else 1

    # ---- 2. timing -------------------------------------------------------------------------------------------
    for case in [c for c in args.time.split(",") if c]:
        seed, depth = case.split("-")
        m = PolyMesh2d(seed, int(depth))
        row = {"case": case, "world": world, "steps": args.steps}
        for name, e in (("nccl", e_nccl), ("peer", e_peer)):
            s = BVESolver(e, m.vert_xyz.shape[0], m.face_xyz.shape[0])
            st = state(m)
            s.set_state(*st, m.face_area, m.face_mask)
            s.init_velocity()
            s.advance(0.001, 2 * np.pi, 2)
            e.sync()
            best = None
            for _ in range(3):
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                s.advance(0.001, 2 * np.pi, args.steps)
                e.sync()
                dt = torch.tensor([time.perf_counter() - t0], device="cuda")
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                ms = float(dt.item()) * 1e3 / args.steps
                best = ms if best is None else min(best, ms)
            row[name + "_ms_per_step"] = best
            s.close()
        if rank == 0:
            print(json.dumps(row), flush=True)
    dist.barrier()
    e_peer.close()
    e_nccl.close()
    dist.destroy_process_group()
    if fails:
        print(f"[rank {rank}] FAILED ({fails})", flush=True)
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
