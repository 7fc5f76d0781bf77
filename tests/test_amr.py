"""Adaptive refinement (the AMR half of the remesh hand-off).

CPU (-m "not gpu"):
  * PolyMesh2d::divide_flagged_faces in the host generator against tests/golden/mesh_amr_*.npz (the independent Python replay
    of src/mesh/lpm_polymesh2d_impl.hpp:124-173 driven by the reference's seed files): every array bit-exact, the
    (refine_count, outcome) of every pass equal, incl. "level limit reached" and "not enough memory";
  * the live Python replay where /root/reference is mounted;
  * invariants of an adaptively refined mesh (area 4 pi, hanging nodes, leaf scan);
  * the numpy restatement of the flag functors against tests/golden/ref_flags.npz (the reference header compiled in place)
    and against the live _ref build when present.
GPU (-m gpu): the flag kernels behind lpmx_refine_flag / lpmx_refine_flag_max against the same goldens (bit-exact flags and
tolerances), host and device pointers; the direct sums on an adaptively refined mesh against the oracle."""
import os

import numpy as np
import pytest

from lpm_b200.api import PolyMesh2d

from conftest import field_rel_err

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AMR_CASES = [("icos", 2, "circ"), ("cubed", 2, "circ"), ("icos", 1, "random"), ("cubed", 1, "random"),
             ("quad_rect", 1, "random"), ("tri_hex", 1, "random")]
INT_ARRAYS = ["edge_origs", "edge_dests", "edge_lefts", "edge_rights", "edge_parents", "edge_kids", "face_verts",
              "face_edges", "face_parent", "face_kids", "face_level", "face_leaf_idx", "face_mask"]
REAL_ARRAYS = ["vert_xyz", "vert_lag_xyz", "face_xyz", "face_lag_xyz", "face_area"]
FLAG_ARGS = {"scalar_max": ("face_vals",), "scalar_integral": ("face_vals", "area"),
             "scalar_variation": ("face_vals", "vert_vals", "face_verts"), "flow_map_variation": ("face_verts", "vert_lag")}
KINDS = tuple(FLAG_ARGS)


def replay(g):
    m = PolyMesh2d(str(g["seed"]), int(g["depth"]), amr_buffer=int(g["amr_buffer"]), amr_limit=int(g["amr_limit"]))
    assert m.nmaxfaces == int(g["nmaxfaces"])
    results = []
    for it in range(g["results"].shape[0]):
        results.append(m.divide_flagged_faces(g[f"flags_{it}"]))
    return m, np.array(results, dtype=np.int32)


@pytest.mark.parametrize("seed,depth,kind", AMR_CASES)
def test_divide_flagged_faces_matches_golden_bit_exact(seed, depth, kind):
    g = np.load(os.path.join(GOLDEN, f"mesh_amr_{seed}_{depth}_{kind}.npz"))
    m, results = replay(g)
    assert np.array_equal(results, g["results"])
    for k in INT_ARRAYS:
        assert np.array_equal(g[k], getattr(m, k)), k
    for k in REAL_ARRAYS:
        assert np.array_equal(g[k].view(np.int64), getattr(m, k).view(np.int64)), k


def test_outcomes_cover_limit_and_no_space():
    g = np.load(os.path.join(GOLDEN, "mesh_amr_icos_1_random.npz"))
    outcomes = set(g["results"][:, 1].tolist())
    assert outcomes == {PolyMesh2d.AMR_DIVIDED_ALL, PolyMesh2d.AMR_LIMIT_REACHED, PolyMesh2d.AMR_NO_SPACE}
    # "not enough memory" divides nothing (the reference warns and returns)
    assert g["results"][-1].tolist() == [0, PolyMesh2d.AMR_NO_SPACE]


@pytest.mark.skipif(not os.path.isdir("/root/reference/mesh_seeds"), reason="reference not mounted")
def test_live_python_replay_of_divide_flagged_faces():
    from oracle import mesh_oracle
    rng = np.random.default_rng(7)
    for seed in ("icos", "cubed", "quad_rect", "tri_hex"):
        radius = 2.5 if seed in ("quad_rect", "tri_hex") else 1.0
        ref = mesh_oracle.TreeMesh(seed, 1, radius=radius)
        m = PolyMesh2d(seed, 1, radius=radius, amr_buffer=3, amr_limit=3)
        for _ in range(3):
            flags = ((rng.random(m.n_faces) < 0.3) & (m.face_mask == 0)).astype(np.uint8)
            assert ref.divide_flagged_faces(flags, m.nmaxfaces, 3) == m.divide_flagged_faces(flags)
        for k, v in ref.arrays().items():
            assert np.array_equal(v, getattr(m, k)), k


def test_live_reference_divide_flagged_faces_every_array_bit_exact():
    """PolyMesh2d<Seed>::divide_flagged_faces of the reference compiled in place (src/mesh/lpm_polymesh2d_impl.hpp:124-173,
    oracle/_ref/liblpm_ref_mesh.so) against the product's generator on fresh pseudo-random flags: three passes per seed, deep
    enough to meet neighbours two levels apart, the level limit and the "not enough memory" return; every outcome, index array,
    coordinate and area."""
    from oracle import ref_mesh
    if not ref_mesh.available():
        pytest.skip("oracle/_ref/liblpm_ref_mesh.so or /root/reference not present (build container only)")
    rng = np.random.default_rng(11)
    seen = set()
    for seed in ("icos", "cubed", "quad_rect", "tri_hex"):
        radius = 2.5 if seed in ("quad_rect", "tri_hex") else 1.0
        for depth, buf, lim, p in ((1, 3, 3, 0.3), (2, 1, 1, 0.5)):
            ref = ref_mesh.RefMesh(seed, depth, radius, buf, lim)
            m = PolyMesh2d(seed, depth, radius=radius, amr_buffer=buf, amr_limit=lim)
            assert ref.counts()["nmaxfaces"] == m.nmaxfaces
            for _ in range(4):
                flags = ((rng.random(m.n_faces) < p) & (m.face_mask == 0)).astype(np.uint8)
                got, want = m.divide_flagged_faces(flags), ref.divide_flagged_faces(flags)
                assert got == want, (seed, depth, got, want)
                seen.add(got[1])
            a = ref.arrays()
            ref.close()
            for k in INT_ARRAYS:
                assert np.array_equal(a[k], getattr(m, k)), (seed, k)
            for k in REAL_ARRAYS:
                assert np.array_equal(a[k].view(np.int64), getattr(m, k).view(np.int64)), (seed, k)
    assert seen == {PolyMesh2d.AMR_DIVIDED_ALL, PolyMesh2d.AMR_LIMIT_REACHED, PolyMesh2d.AMR_NO_SPACE}


def test_flagging_a_divided_face_is_rejected_and_short_flag_arrays_too():
    from lpm_b200.api import LpmxError
    m = PolyMesh2d("icos", 1, amr_buffer=1, amr_limit=1)
    flags = np.zeros(m.n_faces, dtype=np.uint8)
    flags[0] = 1  # a root face, divided by tree_init
    with pytest.raises(LpmxError):
        m.divide_flagged_faces(flags)
    with pytest.raises(LpmxError):
        m.divide_flagged_faces(np.zeros(m.n_faces - 1, dtype=np.uint8))
    assert m.divide_flagged_faces(np.zeros(m.n_faces, dtype=np.uint8)) == (0, PolyMesh2d.AMR_DIVIDED_ALL)


def _fnv(arrays):
    h = 1469598103934665603
    for a in arrays:
        for b in np.ascontiguousarray(a).tobytes():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.parametrize("seed,depth,amr,passes", [("icos", 2, 2, 3), ("cubed", 2, 1, 3), ("quad_rect", 2, 1, 3), ("tri_hex", 1, 2, 3)])
def test_cpp_shim_divide_flagged_faces_equals_the_binding(seed, depth, amr, passes):
    """include/lpm/lpm_polymesh2d.hpp (PolyMesh2d<Seed>::divide_flagged_faces: nmax-sized views refilled in place, coordinates
    pushed first, Logger warnings) against the ctypes binding on the same flag rule; host-only program, no engine."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from lpm_b200 import build
    build.build()
    exe = os.path.join(root, "tests", "cpp", "_build", "amr_mesh_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(root, "include"), "-o", exe,
                    os.path.join(root, "tests", "cpp", "amr_mesh_check.cpp"), "-L" + os.path.join(root, "lpm_b200"), "-llpmx",
                    "-Wl,-rpath," + os.path.join(root, "lpm_b200")], check=True)
    p = subprocess.run([exe, seed, str(depth), str(amr), str(passes)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    counts, digest = p.stdout.split("\n")[:2]
    m = PolyMesh2d(seed, depth, radius=3.0 if seed in ("quad_rect", "tri_hex") else 1.0, amr_buffer=amr, amr_limit=amr)
    warnings = 0
    for _ in range(passes):
        flags = np.zeros(m.nmaxfaces, dtype=np.uint8)
        leaves = np.nonzero(m.face_mask == 0)[0]
        flags[leaves[::3]] = 1
        _, oc = m.divide_flagged_faces(flags)
        warnings += oc != PolyMesh2d.AMR_DIVIDED_ALL
    assert [int(x) for x in counts.split()] == [m.n_verts, m.n_edges, m.n_faces, m.n_face_leaves, m.n_edge_leaves, warnings]
    assert warnings >= 1  # the last pass runs into the level limit / the memory check
    arrays = [m.vert_xyz, m.vert_lag_xyz, m.face_xyz, m.face_lag_xyz, m.face_area, m.face_mask, m.face_verts, m.face_edges,
              m.face_kids, m.face_parent, m.face_level, m.face_leaf_idx, m.edge_origs, m.edge_dests, m.edge_lefts,
              m.edge_rights, m.edge_parents, m.edge_kids]
    assert int(digest, 16) == _fnv(arrays)


@pytest.mark.parametrize("seed", ["icos", "cubed"])
def test_adaptive_mesh_invariants(seed):
    g = np.load(os.path.join(GOLDEN, f"mesh_amr_{seed}_1_random.npz"))
    m, _ = replay(g)
    leaf = m.face_mask == 0
    assert np.array_equal(~leaf, m.face_kids[:, 0] > 0)
    assert (m.face_area[~leaf] == 0).all() and (m.face_area[leaf] > 0).all()
    assert np.array_equal(m.face_leaf_idx, np.concatenate([[0], np.cumsum(leaf)[:-1]]).astype(np.int32))
    assert m.n_face_leaves == leaf.sum()
    assert abs(m.face_area.sum() - 4 * np.pi) < 1e-13  # children tile their parent exactly (great-circle edges)
    assert m.face_level.max() == m.depth + m.amr_limit + 1 and m.face_level[leaf].min() == m.depth + 1
    # kids point back to their parent and are one level deeper
    for k in range(4):
        kids = m.face_kids[~leaf, k]
        assert np.array_equal(m.face_parent[kids], np.nonzero(~leaf)[0])
        assert np.array_equal(m.face_level[kids], m.face_level[~leaf] + 1)
    # hanging nodes: a divided edge whose recorded side is still an undivided face -- that coarser face lists the edge
    # itself or one of its ancestors (it was never told about the subdivision)
    eleaf = m.edge_kids[:, 0] <= 0
    hanging = 0
    for e in np.nonzero(~eleaf)[0]:
        for f in (m.edge_lefts[e], m.edge_rights[e]):
            if leaf[f]:
                a = e
                while a >= 0 and a not in m.face_edges[f]:
                    a = m.edge_parents[a]
                assert a >= 0
                hanging += 1
    assert hanging > 0
    # all particles on the unit sphere; vertex count = seed vertices + one per divided edge (+ one per divided quad)
    assert np.abs(np.linalg.norm(m.vert_xyz, axis=1) - 1).max() < 4e-16
    n_seed = 12 if seed == "icos" else 8
    assert m.n_verts == n_seed + (~eleaf).sum() + (0 if seed == "icos" else (~leaf).sum())


@pytest.mark.parametrize("seed", ["icos", "cubed"])
@pytest.mark.parametrize("kind", KINDS)
def test_flag_restatement_matches_reference_golden(seed, kind):
    from oracle import refinement_oracle as ro
    g = np.load(os.path.join(GOLDEN, "ref_flags.npz"))
    arr = {k: g[f"{seed}_{k}"] for k in FLAG_ARGS[kind]}
    mask = g[f"{seed}_face_mask"]
    for tag in ("rel", "abs"):
        rtol, tol_ref, count, start, end, relative = g[f"{seed}_{kind}_{tag}_meta"]
        tol = rtol * ro.flag_max(kind, mask, **arr) if relative else rtol
        assert tol == tol_ref
        flags, ct = ro.iterate(kind, mask, tol, int(start), int(end), **arr)
        assert ct == int(count)
        assert np.array_equal(flags, g[f"{seed}_{kind}_{tag}_flags"])


def test_flag_restatement_matches_live_reference_build():
    from oracle import refinement_oracle as ro
    L = ro.ref_lib()
    if L is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(11)
    m = PolyMesh2d("cubed", 2, amr_buffer=2, amr_limit=2)
    m.divide_flagged_faces(((rng.random(m.n_faces) < 0.4) & (m.face_mask == 0)).astype(np.uint8))
    fz, vz = rng.standard_normal(m.n_faces), rng.standard_normal(m.n_verts)
    lag = m.vert_xyz + 0.1 * rng.standard_normal(m.vert_xyz.shape)
    full = dict(face_vals=fz, area=m.face_area, vert_vals=vz, face_verts=m.face_verts, vert_lag=lag)
    for kind in KINDS:
        arr = {k: full[k] for k in FLAG_ARGS[kind]}
        fr, cr, tolr = ro.ref_iterate(L, kind, m.face_mask, 0.5, 1, 3, m.n_faces - 2, **arr)
        tol = 0.5 * ro.flag_max(kind, m.face_mask, **arr)
        fo, co = ro.iterate(kind, m.face_mask, tol, 3, m.n_faces - 2, **arr)
        assert tol == tolr and co == cr and np.array_equal(fo, fr), kind


def test_flag_max_of_nothing_is_kokkos_identity():
    from oracle import refinement_oracle as ro
    assert ro.flag_max("scalar_variation", np.ones(4, dtype=np.uint8), face_vals=np.zeros(4), vert_vals=np.zeros(3),
                       face_verts=np.zeros((4, 3), dtype=np.int32)) == -np.finfo(np.float64).max


# ---------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("seed", ["icos", "cubed"])
@pytest.mark.parametrize("kind", KINDS)
def test_gpu_flags_match_reference_golden_bit_exact(engine, seed, kind):
    g = np.load(os.path.join(GOLDEN, "ref_flags.npz"))
    arr = {k: g[f"{seed}_{k}"] for k in FLAG_ARGS[kind]}
    mask = g[f"{seed}_face_mask"]
    for tag in ("rel", "abs"):
        rtol, tol_ref, count, start, end, relative = g[f"{seed}_{kind}_{tag}_meta"]
        tol = rtol * engine.refine_flag_max(kind, mask, **arr) if relative else rtol
        assert tol == tol_ref  # a maximum and one multiplication: bit-exact
        flags, ct = engine.refine_flag(kind, mask, tol, int(start), int(end), **arr)
        assert ct == int(count)
        assert np.array_equal(flags, g[f"{seed}_{kind}_{tag}_flags"])


@pytest.mark.gpu
def test_gpu_flags_accumulate_and_accept_device_pointers(engine):
    import torch
    from oracle import refinement_oracle as ro
    g = np.load(os.path.join(GOLDEN, "ref_flags.npz"))
    mask, fz, area = g["icos_face_mask"], g["icos_face_vals"], g["icos_area"]
    n = mask.shape[0]
    # two functors into the same flag array (Refinement::iterate with two flags, lpm_refinement.hpp:43-62)
    f1, c1 = engine.refine_flag("scalar_integral", mask, 0.03, 0, n, face_vals=fz, area=area)
    f2, c2 = engine.refine_flag("scalar_variation", mask, 0.5, 0, n, flags=f1, face_vals=fz, vert_vals=g["icos_vert_vals"],
                                face_verts=g["icos_face_verts"])
    o1, _ = ro.iterate("scalar_integral", mask, 0.03, 0, n, face_vals=fz, area=area)
    o2, oc2 = ro.iterate("scalar_variation", mask, 0.5, 0, n, flags=o1.copy(), face_vals=fz, vert_vals=g["icos_vert_vals"],
                         face_verts=g["icos_face_verts"])
    assert f2 is f1 and np.array_equal(f2, o2) and c2 == oc2 and c2 >= c1
    # device-resident arrays
    dev = torch.device("cuda", 0)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in (("face_vals", fz), ("area", area))}
    tmask = torch.from_numpy(mask).to(dev)
    tflags = torch.zeros(n, dtype=torch.uint8, device=dev)
    mx = engine.refine_flag_max("scalar_integral", tmask, **t)
    assert mx == ro.flag_max("scalar_integral", mask, face_vals=fz, area=area)
    _, ct = engine.refine_flag("scalar_integral", tmask, 0.03, 0, n, flags=tflags, **t)
    assert ct == c1 and np.array_equal(tflags.cpu().numpy(), o1)
    # empty range and empty mesh
    f0, c0 = engine.refine_flag("scalar_max", mask, 0.0, 5, 5, face_vals=fz)
    assert c0 == 0 and not f0.any()
    assert engine.refine_flag_max("scalar_max", np.zeros(0, dtype=np.uint8), face_vals=np.zeros(0)) == -np.finfo(np.float64).max


@pytest.mark.gpu
@pytest.mark.parametrize("seed", ["icos", "cubed"])
def test_gpu_direct_sums_on_an_adaptive_mesh(engine, oracle, seed):
    """The pair sums only see (coordinates, strength, area, mask): an AMR mesh (mixed levels, scattered divided faces) goes
    through the same path.  Tolerance 1e-12 field-relative (north_star)."""
    from lpm_b200 import gallery
    m = PolyMesh2d(seed, 3, amr_buffer=2, amr_limit=2)
    gv = gallery.GaussianVortexSphere()
    start, tol = 0, None
    for _ in range(2):
        z = gv(m.face_xyz)
        if tol is None:
            tol = 0.2 * engine.refine_flag_max("scalar_integral", m.face_mask, face_vals=z, area=m.face_area)
        n = m.n_faces
        flags, ct = engine.refine_flag("scalar_integral", m.face_mask, tol, start, n, face_vals=z, area=m.face_area)
        nd, oc = m.divide_flagged_faces(flags)
        assert nd == ct and oc == PolyMesh2d.AMR_DIVIDED_ALL
        start = n
    assert m.face_level.max() == 3 + 2 + 1
    fz = gv(m.face_xyz)
    leaf = m.face_mask == 0
    vu = engine.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)
    fu = engine.bve_velocity(None, m.face_xyz, fz, m.face_area, m.face_mask, collocated=True)
    assert field_rel_err(vu, oracle.bve_velocity(m.vert_xyz, m.face_xyz, fz, m.face_area, m.face_mask)) <= 1e-12
    ou = oracle.bve_velocity(m.face_xyz, m.face_xyz, fz, m.face_area, m.face_mask, collocated=True)
    assert field_rel_err(fu, ou, leaf) <= 1e-12
