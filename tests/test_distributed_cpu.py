"""CPU, world_size 2 over gloo: the host-side logic of the target-sharded evaluation.

Each rank evaluates its two target lists (A: own leaf faces, B: own vertices and divided faces -- lpm_b200/partition.py, the
mirror of solver_set_state) against ALL leaf sources (with the oracle standing in for the GPU kernel), the vertex and the face
row ranges are all-gathered, and the result must equal the unsharded evaluation bit for bit -- the property the multi-GPU
stepper relies on.  Also checks the partition helpers
and that the NCCL-unique-id hand-off reaches every rank identically."""
import os
import socket

import numpy as np
import pytest

from lpm_b200 import partition


def test_offsets_cover_without_overlap():
    for n, w in [(9382, 1), (9382, 2), (229376, 8), (7, 8), (0, 4)]:
        off = partition.target_offsets(n, w)
        assert off[0] == 0 and off[-1] == n and all(a <= b for a, b in zip(off, off[1:]))
        assert max(b - a for a, b in zip(off, off[1:])) - min(b - a for a, b in zip(off, off[1:])) <= 1


def test_target_lists_are_class_balanced_partitions(meshes):
    from lpm_b200.api import PolyMesh2d
    amr = PolyMesh2d("cubed", 2, amr_buffer=2, amr_limit=2)
    rng = np.random.default_rng(3)
    for _ in range(2):  # an adaptively refined mesh: leaves and divided faces interleave in face order
        amr.divide_flagged_faces(((rng.random(amr.n_faces) < 0.4) & (amr.face_mask == 0)).astype(np.uint8))
    for m in (meshes("icos", 3), meshes("cubed", 3), amr):
        for w in (1, 2, 3, 8):
            l = partition.leaf_offsets(m.face_mask, w)
            assert l[0] == 0 and l[-1] == m.n_face_leaves
            seen, na, nb = [], [], []
            for r in range(w):
                a, b = partition.target_lists(m.n_verts, m.face_mask, r, w)
                assert len(a) == l[r + 1] - l[r]
                assert (m.face_mask[a - m.n_verts] == 0).all()                       # A: leaf faces only (the sources)
                assert ((b < m.n_verts) | (m.face_mask[np.maximum(b - m.n_verts, 0)] == 1)).all()  # B: never sources
                seen += [a, b]
                na.append(len(a)), nb.append(len(b))
            assert max(na) - min(na) <= 1 and max(nb) - min(nb) <= 1                 # both classes balanced over the ranks
            assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(m.n_verts + m.n_faces))  # a partition of all targets
    assert partition.interactions_per_eval(2562, 6820, 5120) == (2562 + 6820) * 5120 - 5120


@pytest.mark.gpu
def test_engine_target_lists_equal_the_mirror(meshes):
    """lpmx_local_targets (host code of the engine, no device needed) against lpm_b200/partition.py."""
    m = meshes("icos", 3)
    mask = np.ascontiguousarray(m.face_mask, dtype=np.uint8)
    from lpm_b200.api import Engine
    e = Engine(0)
    try:
        for w in (1, 3, 8):
            for r in range(w):
                e.set_partition(r, w)
                a, b = e.local_targets(m.n_verts, m.n_faces, mask)
                pa, pb = partition.target_lists(m.n_verts, mask, r, w)
                assert np.array_equal(a, pa) and np.array_equal(b, pb)
        e.set_partition(0, 1)
    finally:
        e.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lpm_b200 import gallery
        from lpm_b200.api import PolyMesh2d
        from lpm_b200.dist import broadcast_unique_id
        from oracle import oracle
        m = PolyMesh2d("cubed", 3)
        fz = gallery.SolidBodyRotation()(m.face_xyz)
        nv, nf = m.n_verts, m.n_faces
        la, lb = partition.target_lists(nv, m.face_mask, rank, world)
        # this rank's targets: list A then list B; faces need the collocated (skip-self) rule
        full_v = oracle.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
        full_f = oracle.bve_velocity(None, m.face_xyz, fz, m.face_area, m.face_mask, collocated=True)
        full = np.concatenate([full_v, full_f])
        xyz = np.concatenate([m.vert_xyz, m.face_xyz])
        gathered = torch.zeros((nv + nf, 3), dtype=torch.float64)
        for g in np.concatenate([la, lb]):
            mask = m.face_mask.copy()
            if g >= nv:
                mask[g - nv] = 1  # skip the self pair by index
            gathered[g] = torch.from_numpy(oracle.bve_velocity(xyz[g:g + 1], m.face_xyz, fz, m.face_area, mask)[0])
        # the row gather in list order: rows -> perm order, every rank broadcasts its segment, scatter back
        lists = [np.concatenate(partition.target_lists(nv, m.face_mask, r, world)) for r in range(world)]
        perm = np.concatenate(lists)
        off = np.concatenate([[0], np.cumsum([len(x) for x in lists])])
        buf = gathered[torch.from_numpy(perm)].contiguous()
        for r in range(world):
            seg = buf[off[r]:off[r + 1]].contiguous()
            dist.broadcast(seg, src=r)
            buf[off[r]:off[r + 1]] = seg
        gathered[torch.from_numpy(perm)] = buf
        ok_sum = bool(np.array_equal(gathered.numpy(), full))
        # leaf ranges: the packed records this rank would own
        lo = partition.leaf_offsets(m.face_mask, world)
        uid = broadcast_unique_id(lambda: bytes(range(128)), rank)
        q.put((rank, ok_sum, lo, uid == bytes(range(128))))
    finally:
        dist.destroy_process_group()


def test_sharded_evaluation_equals_unsharded_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] for r in res), "gathered shards differ from the unsharded evaluation"
    assert res[0][2] == res[1][2]
    assert all(r[3] for r in res)
