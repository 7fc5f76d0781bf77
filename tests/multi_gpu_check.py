"""Run under torchrun (one rank per GPU): sharded BVERK4 / IC2D RK2 / SWERK2 steps must equal the CPU oracle on every rank.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Used by tests/test_gpu_multi.py; exits non-zero on any mismatch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import field_rel_err
    from lpm_b200 import gallery
    from lpm_b200.api import Engine, PolyMesh2d, swe_rk2_step
    from lpm_b200.dist import env_rank_world, init_engine_comm
    from oracle import oracle
    import test_gpu_parity_swe_rk2 as T

    rank, world, local = env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    e = Engine(local)
    init_engine_comm(e, rank, world)
    worst = 0.0
    for seed, depth in (("icos", 3), ("cubed", 4)):
        m = PolyMesh2d(seed, depth)
        leaf = m.face_mask == 0
        f = gallery.RossbyHaurwitz54()
        f.set_stationary_wave_speed()
        vz, fz = f(m.vert_xyz), f(m.face_xyz)
        a = (m.face_xyz, fz, m.face_area, m.face_mask)
        vu = oracle.bve_velocity(m.vert_xyz, *a)
        fu = oracle.bve_velocity(None, *a, collocated=True)
        fu[~leaf] = 0.0
        got = [m.vert_xyz.copy(), vz.copy(), vu.copy(), m.face_xyz.copy(), fz.copy(), fu.copy()]
        ref = [x.copy() for x in got]
        e.bve_rk4_step(0.01, 2 * np.pi, *got, m.face_area, m.face_mask, n_steps=2)
        oracle.bve_rk4_step(0.01, 2 * np.pi, *ref, m.face_area, m.face_mask, n_steps=2)
        errs = [field_rel_err(got[0], ref[0]), field_rel_err(got[1], ref[1]), field_rel_err(got[2], ref[2]),
                field_rel_err(got[3], ref[3], leaf), field_rel_err(got[4], ref[4], leaf), field_rel_err(got[5], ref[5], leaf)]
        worst = max(worst, max(errs))
        print(f"[rank {rank}/{world}] bve_rk4 {seed}-{depth}: max field-rel err {max(errs):.3e}", flush=True)
        # IC2D RK2
        pu, ppsi = oracle.ic2d_sums(m.vert_xyz, *a)
        au, apsi = oracle.ic2d_sums(None, *a, targets_are_sources=True)
        au[~leaf], apsi[~leaf] = 0.0, 0.0
        got = [m.vert_xyz.copy(), vz.copy(), pu.copy(), ppsi.copy(), m.face_xyz.copy(), fz.copy(), au.copy(), apsi.copy()]
        ref = [x.copy() for x in got]
        e.ic2d_rk2_step(0.01, 2 * np.pi, 0.0, *got, m.face_area, m.face_mask, n_steps=2)
        oracle.ic2d_rk2_step(0.01, 2 * np.pi, 0.0, *ref, m.face_area, m.face_mask, n_steps=2)
        errs = [field_rel_err(got[i], ref[i]) for i in (0, 1, 2, 3)] + [field_rel_err(got[i], ref[i], leaf) for i in (4, 5, 6, 7)]
        worst = max(worst, max(errs))
        print(f"[rank {rank}/{world}] ic2d_rk2 {seed}-{depth}: max field-rel err {max(errs):.3e}", flush=True)
    # sharded host I/O (lpmx_set_io_sharded): every rank passes arrays whose rows of OTHER ranks are poisoned; after two steps its
    # own rows equal the oracle's and the others are still untouched
    m = PolyMesh2d("cubed", 4)
    leaf = m.face_mask == 0
    f = gallery.RossbyHaurwitz54()
    f.set_stationary_wave_speed()
    vz, fz = f(m.vert_xyz), f(m.face_xyz)
    a = (m.face_xyz, fz, m.face_area, m.face_mask)
    ref = [m.vert_xyz.copy(), vz.copy(), oracle.bve_velocity(m.vert_xyz, *a), m.face_xyz.copy(), fz.copy(),
           oracle.bve_velocity(None, *a, collocated=True)]
    got = [x.copy() for x in ref]
    la, lb = e.local_targets(m.n_verts, m.n_faces, m.face_mask)
    own_all = np.zeros(m.n_verts + m.n_faces, bool)
    own_all[la], own_all[lb] = True, True
    own = [own_all[:m.n_verts], own_all[m.n_verts:]]
    for k, arr in enumerate(got):
        arr[~own[k // 3]] = np.nan
    e.set_io_sharded(True)
    e.bve_rk4_step(0.01, 2 * np.pi, *got, m.face_area, m.face_mask, n_steps=2)
    e.set_io_sharded(False)
    oracle.bve_rk4_step(0.01, 2 * np.pi, *ref, m.face_area, m.face_mask, n_steps=2)
    ok = True
    for k, (g, r) in enumerate(zip(got, ref)):
        o = own[k // 3]
        ok &= bool(np.isnan(g[~o]).all())
        if o.any():
            ok &= field_rel_err(g[o], r[o]) <= 1e-10 * max(1.0, np.abs(r).max() / max(np.abs(r[o]).max(), 1e-300))
    print(f"[rank {rank}/{world}] sharded host I/O ({len(la)} leaf faces + {len(lb)} other targets of {m.n_verts + m.n_faces}): "
          f"{'ok' if ok else 'FAILED'}", flush=True)
    if not ok:
        worst = np.inf
    # SWE RK2 with the Laplacian provider (the provider sees the gathered arrays on every rank)
    m = PolyMesh2d("cubed", 3)
    st0 = T.tc2_state(oracle, m, eps=0.0, div_amp=0.02)
    ref = oracle.swe_rk2_step(0.0125, T.OMEGA, T.G, 0.0, st0.copy(), T.host_laplacian, n_steps=2)
    got = st0.copy()
    swe_rk2_step(e, 0.0125, T.OMEGA, T.G, 0.0, got.p, got.a, got.mask, T.device_laplacian, n_steps=2)
    T.compare(got, ref, m.face_mask)
    print(f"[rank {rank}/{world}] swe_rk2 cubed-3: ok", flush=True)
    dist.barrier()
    e.close()
    dist.destroy_process_group()
    if not worst <= 1e-10:
        print(f"[rank {rank}] FAILED: worst {worst:.3e}", flush=True)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
