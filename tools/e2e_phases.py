"""Where the time of one in-place step goes at N > 1 (development probe): set_state / advance / get_state of the resident solver,
each followed by lpmx_sync, sharded host I/O on, pinned host arrays; run under torchrun, with LPMX_PEER_EXCHANGE=0 and =1."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from lpm_b200 import gallery
    from lpm_b200.api import BVESolver, Engine, PolyMesh2d
    from lpm_b200.dist import env_rank_world, init_engine_comm
    rank, world, local = env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    e = Engine(local)
    init_engine_comm(e, rank, world)
    m = PolyMesh2d("cubed", 7)
    f = gallery.RossbyHaurwitz54()
    f.set_stationary_wave_speed()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    st = [pin(m.vert_xyz), pin(f(m.vert_xyz)), pin(np.zeros((m.n_verts, 3))), pin(m.face_xyz), pin(f(m.face_xyz)),
          pin(np.zeros((m.n_faces, 3)))]
    area, mask = pin(m.face_area), pin(m.face_mask)
    s = BVESolver(e, m.n_verts, m.n_faces)
    s.set_state(*st, area, mask)
    s.init_velocity()
    s.get_state(*st)
    e.set_io_sharded(True)
    t = np.zeros(4)
    reps = 10
    for it in range(reps + 2):
        dist.barrier()
        torch.cuda.synchronize()
        c0 = time.perf_counter()
        s.set_state(*st, area, mask)
        e.sync()
        c1 = time.perf_counter()
        s.advance(0.003, 2 * np.pi, 1)
        e.sync()
        c2 = time.perf_counter()
        s.get_state(*st)
        e.sync()
        c3 = time.perf_counter()
        if it >= 2:
            t += [c1 - c0, c2 - c1, c3 - c2, c3 - c0]
    if os.environ.get("LPMX_PROFILE_DUMP"):
        e.profile_enable(True)
        e.profile_read()
        for it in range(2):
            dist.barrier()
            torch.cuda.synchronize()
            s.set_state(*st, area, mask)
            e.sync()
            s.advance(0.003, 2 * np.pi, 1)
            e.sync()
        e.profile_read()
        e.profile_enable(False)
    t *= 1e3 / reps
    print(f"[rank {rank}/{world}] peer={os.environ.get('LPMX_PEER_EXCHANGE', '0')} set_state {t[0]:.3f} ms  advance {t[1]:.3f} ms  "
          f"get_state {t[2]:.3f} ms  total {t[3]:.3f} ms", flush=True)
    s.close()
    e.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
