"""Independent restatement (pure Python, small depths only) of the reference's tree refinement, uniform and adaptive.

TEST INFRASTRUCTURE ONLY.  It replays, in the reference's own insertion order,
  MeshSeed::read_file            /root/reference/src/mesh/lpm_mesh_seed.cpp:20-206  (the .dat parser)
  PolyMesh2d::tree_init          src/mesh/lpm_polymesh2d_impl.hpp:25-42
  PolyMesh2d::divide_flagged_faces  src/mesh/lpm_polymesh2d_impl.hpp:124-173 (+ Faces::scan_leaves)
  Edges::divide                  src/mesh/lpm_edges.cpp:58-96
  FaceDivider<.., TriFace>       src/mesh/lpm_faces_impl.hpp:284-431
  FaceDivider<.., QuadFace>      src/mesh/lpm_faces_impl.hpp:433-574
  SphereGeometry helpers         src/lpm_geometry.hpp:449-642
so that the product's C++ generator (lpm_b200/csrc/lpmx_mesh.cpp) can be checked bit-for-bit on every
integer array and on coordinates/areas.  tests/golden/make_mesh_golden.py runs it against the reference's
seed files (in the build container, where /root/reference is mounted) and commits the results as fixtures.

Floating point: Python floats are IEEE doubles and each operation rounds once, i.e. the same as the C++
generator compiled with -ffp-contract=off; math.sqrt/atan2/tan/atan are the platform libm's.
"""
import math

NULL = -1
ZERO_TOL = 2.220446049250313e-16


def read_seed(path, nverts, nfaces, nedges, nfv, ndim=3):
    """Parse a mesh seed .dat file the way MeshSeed::read_file does (line-number driven)."""
    crds, edges, fverts, fedges = [], [], [], []
    edge_hdr = fv_hdr = fe_hdr = None
    ncrds = nverts + nfaces
    with open(path) as f:
        for lineno, line in enumerate(f, start=1):
            if "edgeO" in line:
                edge_hdr = lineno
            if "faceverts" in line:
                fv_hdr = lineno
            if "faceedges" in line:
                fe_hdr = lineno
            tok = line.split()
            if 1 < lineno < ncrds + 2:
                crds.append([float(t) for t in tok[:ndim]])
            elif edge_hdr and edge_hdr < lineno < edge_hdr + nedges + 1:
                edges.append([int(t) for t in tok[:4]])
            elif fv_hdr and fv_hdr < lineno < fv_hdr + nfaces + 1:
                fverts.append([int(t) for t in tok[:nfv]])
            elif fe_hdr and fe_hdr < lineno < fe_hdr + nfaces + 1:
                fedges.append([int(t) for t in tok[:nfv]])
    assert len(crds) == ncrds and len(edges) == nedges and len(fverts) == nfaces and len(fedges) == nfaces
    return crds, edges, fverts, fedges


SEEDS = {
    "icos": dict(file="icosTriSphereSeed.dat", nverts=12, nfaces=20, nedges=30, nfv=3),
    "cubed": dict(file="cubedSphereSeed.dat", nverts=8, nfaces=6, nedges=12, nfv=4),
    "quad_rect": dict(file="quadRectSeed.dat", nverts=9, nfaces=4, nedges=12, nfv=4, ndim=2),
    "tri_hex": dict(file="triHexSeed.dat", nverts=7, nfaces=6, nedges=12, nfv=3, ndim=2),
}


# ---- PlaneGeometry (src/lpm_geometry.hpp:69-160) ----
def _plane_tri_area(a, b, c):
    ar = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
    return 0.5 * abs(ar)


def _plane_poly_area(ctr, vs):
    n = len(vs)
    ar = 0.0
    for i in range(n):
        ar += _plane_tri_area(ctr, vs[i], vs[(i + 1) % n])
    return ar


def _plane_barycenter(vs):
    n = len(vs)
    v = [0.0, 0.0]
    for p in vs:
        v[0] += p[0]
        v[1] += p[1]
    s = 1.0 / n
    return [v[0] * s, v[1] * s]


def _plane_midpoint(a, b):
    return [0.5 * (a[0] + b[0]), 0.5 * (a[1] + b[1])]


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _normalized(v):
    s = 1.0 / math.sqrt(_dot(v, v))
    return [v[0] * s, v[1] * s, v[2] * s]


def _dist(a, b):
    cp = _cross(a, b)
    return math.atan2(math.sqrt(_dot(cp, cp)), _dot(a, b))


def _tri_area(a, b, c):
    s1, s2, s3 = _dist(a, b), _dist(b, c), _dist(c, a)
    hp = 0.5 * (s1 + s2 + s3)
    zz = math.tan(0.5 * hp) * math.tan(0.5 * (hp - s1)) * math.tan(0.5 * (hp - s2)) * math.tan(0.5 * (hp - s3))
    if abs(zz) < ZERO_TOL:
        zz = 0
    return 4 * math.atan(math.sqrt(zz))


def _poly_area(ctr, vs):
    n = len(vs)
    ar = 0
    for i in range(n):
        ar += _tri_area(ctr, vs[i], vs[(i + 1) % n])
    return ar


def _barycenter(vs):
    n = len(vs)
    v = [0.0, 0.0, 0.0]
    for p in vs:
        v[0] += p[0]
        v[1] += p[1]
        v[2] += p[2]
    s = 1.0 / n
    return _normalized([v[0] * s, v[1] * s, v[2] * s])


def _midpoint(a, b):
    return _normalized([0.5 * (a[0] + b[0]), 0.5 * (a[1] + b[1]), 0.5 * (a[2] + b[2])])


class TreeMesh:
    def __init__(self, seed, depth, seed_dir="/root/reference/mesh_seeds", radius=1.0):
        import os
        d = SEEDS[seed]
        self.nfv = nfv = d["nfv"]
        self.ndim = d.get("ndim", 3)
        if self.ndim == 2:
            self._midpoint, self._barycenter, self._poly_area = _plane_midpoint, _plane_barycenter, _plane_poly_area
        else:
            self._midpoint, self._barycenter, self._poly_area = _midpoint, _barycenter, _poly_area
        crds, edges, fverts, fedges = read_seed(os.path.join(seed_dir, d["file"]), d["nverts"], d["nfaces"],
                                                d["nedges"], nfv, self.ndim)
        crds = [[c * radius for c in row] for row in crds]  # MeshSeed(maxr) (lpm_mesh_seed.cpp:10-18)
        nv = d["nverts"]
        self.vx = [list(c) for c in crds[:nv]]
        self.vlag = [list(c) for c in crds[:nv]]
        self.eo = [e[0] for e in edges]
        self.ed = [e[1] for e in edges]
        self.el = [e[2] for e in edges]
        self.er = [e[3] for e in edges]
        self.ep = [NULL] * len(edges)
        self.ek = [[NULL, NULL] for _ in edges]
        self.fx, self.flag, self.fverts, self.fedges = [], [], [], []
        self.fparent, self.fkids, self.flevel, self.fmask, self.farea = [], [], [], [], []
        for i in range(d["nfaces"]):
            ctr = list(crds[nv + i])
            vs = [self.vx[v] for v in fverts[i]]
            self._add_face(ctr, list(ctr), list(fverts[i]), list(fedges[i]), NULL, self._poly_area(ctr, vs))
        start = 0
        for _ in range(depth):
            stop = len(self.fx)
            for j in range(start, stop):
                if not self.fkids[j][0] > 0:
                    self._divide(j)
            start = stop - 1
        self.init_depth = depth
        self.scan_leaves()

    def scan_leaves(self):
        self.leaf_idx = []
        acc = 0
        for k in self.fkids:
            self.leaf_idx.append(acc)
            acc += 0 if k[0] > 0 else 1

    def divide_flagged_faces(self, flags, nmaxfaces, amr_limit):
        """Returns (refine_count, outcome): outcome 0 all divided, 1 not enough memory (nothing divided), 2 level limit
        reached for some flagged faces."""
        n_in = len(self.fx)
        flag_count = sum(1 for i in range(n_in) if flags[i])
        space_left = nmaxfaces - n_in
        if flag_count > space_left // 4:
            return 0, 1
        refine_count, limit_reached = 0, False
        for i in range(n_in):
            if flags[i]:
                if self.flevel[i] <= self.init_depth + amr_limit:
                    self._divide(i)
                    refine_count += 1
                else:
                    limit_reached = True
        self.scan_leaves()
        return refine_count, (2 if limit_reached else 0)

    def _add_face(self, ctr, lctr, verts, edges, parent, area):
        self.fx.append(ctr)
        self.flag.append(lctr)
        self.fverts.append(verts)
        self.fedges.append(edges)
        self.fparent.append(parent)
        self.fkids.append([NULL] * 4)
        self.flevel.append(1 if parent == NULL else self.flevel[parent] + 1)
        self.fmask.append(0)
        self.farea.append(area)

    def _add_edge(self, o, d, left, right, parent=NULL):
        self.eo.append(o)
        self.ed.append(d)
        self.el.append(left)
        self.er.append(right)
        self.ep.append(parent)
        self.ek.append([NULL, NULL])
        return len(self.eo) - 1

    def _split_edge(self, e):
        mid_v = len(self.vx)
        self.vx.append(self._midpoint(self.vx[self.eo[e]], self.vx[self.ed[e]]))
        # the reference takes the Lagrangian destination from the PHYSICAL array (lpm_edges.cpp:81)
        self.vlag.append(self._midpoint(self.vlag[self.eo[e]], self.vx[self.ed[e]]))
        k0 = self._add_edge(self.eo[e], mid_v, self.el[e], self.er[e], e)
        k1 = self._add_edge(mid_v, self.ed[e], self.el[e], self.er[e], e)
        self.ek[e] = [k0, k1]
        return k0, k1

    def _divide(self, f):
        n = self.nfv
        kid0 = len(self.fx)
        kv = [[NULL] * n for _ in range(4)]
        ke = [[NULL] * n for _ in range(4)]
        for i in range(n):
            kv[i][i] = self.fverts[f][i]
        for i in range(n):
            pe = self.fedges[f][i]
            if self.ek[pe][0] > 0:
                k0, k1 = self.ek[pe]
            else:
                k0, k1 = self._split_edge(pe)
            a, b = i, (i + 1) % n
            if self.el[pe] == f:
                ke[a][i] = k0
                self.el[k0] = kid0 + a
                ke[b][i] = k1
                self.el[k1] = kid0 + b
            else:
                ke[a][i] = k1
                self.er[k1] = kid0 + a
                ke[b][i] = k0
                self.er[k0] = kid0 + b
            m = self.ed[k0]
            if n == 3:
                slots = {0: ((0, 1), (1, 0), (3, 2)), 1: ((1, 2), (2, 1), (3, 0)), 2: ((2, 0), (0, 2), (3, 1))}[i]
                for kk, ss in slots:
                    kv[kk][ss] = m
            else:
                kv[a][b] = m
                kv[b][a] = m
        if n == 3:
            e0 = len(self.eo)
            for i in range(3):
                ke[3][i] = e0 + i
            ke[0][1] = e0 + 1
            ke[1][2] = e0 + 2
            ke[2][0] = e0
            self._add_edge(kv[2][1], kv[2][0], kid0 + 3, kid0 + 2)
            self._add_edge(kv[0][2], kv[0][1], kid0 + 3, kid0 + 0)
            self._add_edge(kv[1][0], kv[1][2], kid0 + 3, kid0 + 1)
        else:
            cv = len(self.vx)
            self.vx.append(list(self.fx[f]))
            self.vlag.append(list(self.flag[f]))
            for i in range(4):
                kv[i][(i + 2) % 4] = cv
            e0 = len(self.eo)
            self._add_edge(kv[0][1], kv[0][2], kid0 + 0, kid0 + 1)
            ke[0][1] = e0
            ke[1][3] = e0
            self._add_edge(kv[2][0], kv[2][3], kid0 + 3, kid0 + 2)
            ke[2][3] = e0 + 1
            ke[3][1] = e0 + 1
            self._add_edge(kv[2][1], kv[2][0], kid0 + 1, kid0 + 2)
            ke[1][2] = e0 + 2
            ke[2][0] = e0 + 2
            self._add_edge(kv[3][1], kv[3][0], kid0 + 0, kid0 + 3)
            ke[0][2] = e0 + 3
            ke[3][0] = e0 + 3
        kids = []
        for i in range(4):
            vs = [self.vx[v] for v in kv[i]]
            ls = [self.vlag[v] for v in kv[i]]
            ctr = self._barycenter(vs)
            kids.append((ctr, self._barycenter(ls), self._poly_area(ctr, vs)))
        for i in range(4):
            self._add_face(kids[i][0], kids[i][1], kv[i], ke[i], f, kids[i][2])
        self.fkids[f] = [kid0 + i for i in range(4)]
        self.farea[f] = 0.0
        self.fmask[f] = 1

    def arrays(self):
        import numpy as np
        return dict(
            vert_xyz=np.array(self.vx, dtype=np.float64), vert_lag_xyz=np.array(self.vlag, dtype=np.float64),
            edge_origs=np.array(self.eo, dtype=np.int32), edge_dests=np.array(self.ed, dtype=np.int32),
            edge_lefts=np.array(self.el, dtype=np.int32), edge_rights=np.array(self.er, dtype=np.int32),
            edge_parents=np.array(self.ep, dtype=np.int32), edge_kids=np.array(self.ek, dtype=np.int32),
            face_xyz=np.array(self.fx, dtype=np.float64), face_lag_xyz=np.array(self.flag, dtype=np.float64),
            face_area=np.array(self.farea, dtype=np.float64), face_mask=np.array(self.fmask, dtype=np.uint8),
            face_verts=np.array(self.fverts, dtype=np.int32), face_edges=np.array(self.fedges, dtype=np.int32),
            face_parent=np.array(self.fparent, dtype=np.int32), face_kids=np.array(self.fkids, dtype=np.int32),
            face_level=np.array(self.flevel, dtype=np.int32), face_leaf_idx=np.array(self.leaf_idx, dtype=np.int32),
        )
