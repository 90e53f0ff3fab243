"""Setup-time geometry: the host-side mirror of the reference's ``Geometry`` module.

Only what the contact-wrench path needs as *inputs* is restated here (SURVEY.md section 2, rows
"OUT OF SCOPE (setup-time)"): the ``eMesh`` container, the mesh generators used by the
reference's tests/configs and a deterministic bounding-volume-tree builder whose flattened
output is what both the CUDA library and the CPU oracle consume.

Reference (paths relative to /root/reference):
  * eMesh, as_tri_eMesh / as_tet_eMesh ............ src/geometry/mesh.jl:10-78
  * eMesh_half_plane / eMesh_sphere / eMesh_box ... src/geometry/mesh.jl:430-575
  * extrude_mesh ................................... src/geometry/mesh.jl:600-659
  * eMesh_to_tree (bottom-up blob merge) ........... src/geometry/blob_types.jl:136-190
  * recursive_top_down ............................. src/geometry/top_down.jl:10-32
  * leaf OBB fitting ............................... src/obb/obb_construction.jl:13-41

The reference's tree topology depends on Julia's Dict/Set/PriorityQueue iteration order
 and therefore cannot be reproduced bit-for-bit; this builder follows the same
cost function and merge rule but breaks ties by (cost, id_a, id_b).  Indices are 0-based here.
"""
from __future__ import annotations

import heapq
import math
from dataclasses import dataclass
from typing import Optional

import numpy as np

__all__ = [
    "eMesh", "as_tri_eMesh", "as_tet_eMesh", "transform", "eMesh_box", "eMesh_half_plane", "eMesh_sphere",
    "eMesh_grid_square", "extrude_mesh", "create_swept_mesh", "f_swept_triv", "f_swept_circle", "FlatTree", "eMesh_to_tree", "tet_volume", "fit_tri_obb", "fit_tet_obb", "fit_tri_obb_batch", "fit_tet_obb_batch",
]


# ------------------------------------------------------------------------------------------------
# eMesh
# ------------------------------------------------------------------------------------------------
def tet_volume(p: np.ndarray) -> float:
    """src/math_kernel/geometry_kernel.jl:27-39 (sign convention only; plain numpy arithmetic)."""
    return _vol(*np.asarray(p, dtype=np.float64))


def _vol(a, b, c, d):
    a1, a2, a3 = a
    b1, b2, b3 = b
    c1, c2, c3 = c
    d1, d2, d3 = d
    V = (b1 - a1) * (c2 * d3 - c3 * d2)
    V += (b2 - a2) * (c3 * d1 - c1 * d3)
    V += (b3 - a3) * (c1 * d2 - c2 * d1)
    V += (c1 - d1) * (a3 * b2 - a2 * b3)
    V += (c2 - d2) * (a1 * b3 - a3 * b1)
    V += (c3 - d3) * (a2 * b1 - a1 * b2)
    return V / 6.0


class eMesh:
    """Geometry container: points, optional triangles, optional tets with per-vertex normalized
    penetration extent eps (src/geometry/mesh.jl:10-46)."""

    def __init__(self, point, tri=None, tet=None, eps=None, check: bool = True):
        self.point = np.array(point, dtype=np.float64).reshape(-1, 3)
        self.tri = None if tri is None else np.array(tri, dtype=np.int64).reshape(-1, 3)
        self.tet = None if tet is None else np.array(tet, dtype=np.int64).reshape(-1, 4)
        self.eps = None if eps is None else np.array(eps, dtype=np.float64).reshape(-1)
        if self.tri is None and self.tet is None:
            raise ValueError("a whole lot of nothing")
        if self.tet is not None:
            if self.eps is None or len(self.eps) != len(self.point):
                raise ValueError("length(eps) must equal length(point)")
            if check and len(self.eps):
                if not (0.0 < self.eps.max()):
                    raise ValueError("normalized penetration extent must be non-negative")
                if self.eps.min() != 0.0:
                    raise ValueError("normalized penetration extent must be zero on the surface of the volume mesh")
                p = self.point[self.tet]
                if len(p) and not (0.0 < _vol(p[:, 0].T, p[:, 1].T, p[:, 2].T, p[:, 3].T)).all():
                    raise ValueError("inverted tetrahedron")
        elif self.eps is not None:
            raise ValueError("eps given for a mesh without tets")

    @property
    def is_tri(self) -> bool:
        return self.tri is not None

    @property
    def is_tet(self) -> bool:
        return self.tet is not None

    def n_point(self) -> int:
        return len(self.point)

    def n_tri(self) -> int:
        return 0 if self.tri is None else len(self.tri)

    def n_tet(self) -> int:
        return 0 if self.tet is None else len(self.tet)

    def copy(self) -> "eMesh":
        return eMesh(self.point.copy(), None if self.tri is None else self.tri.copy(), None if self.tet is None else self.tet.copy(),
                     None if self.eps is None else self.eps.copy(), check=False)


def as_tet_eMesh(m: eMesh) -> eMesh:
    """src/geometry/mesh.jl:54-55"""
    if m.tet is None:
        raise ValueError("mesh has no tets")
    return eMesh(m.point, None, m.tet, m.eps, check=False)


def as_tri_eMesh(m: eMesh) -> eMesh:
    """src/geometry/mesh.jl:63-64 (the Tri,Tet and Tri,Nothing methods)."""
    if m.tri is None:
        raise ValueError("mesh has no triangles (the Nothing,Tet method of as_tri_eMesh is not needed on this path)")
    return eMesh(m.point, m.tri, None, None, check=False)


def transform(m: eMesh, R=None, t=None) -> eMesh:
    """transform!(e_mesh, basic_dh(...)) -- src/geometry/mesh.jl:171-178.  Returns m (mutated)."""
    P = m.point
    if R is not None:
        R = np.asarray(R, dtype=np.float64)
        if R.ndim == 0:
            R = np.eye(3) * float(R)
        elif R.ndim == 1:
            R = np.diag(R)
        P = P @ R.T
    if t is not None:
        P = P + np.asarray(t, dtype=np.float64)
    m.point = np.ascontiguousarray(P)
    return m


# ------------------------------------------------------------------------------------------------
# Basic shapes
# ------------------------------------------------------------------------------------------------
def eMesh_half_plane(plane_w: float = 1.0, is_include_vis_sides: bool = False) -> eMesh:
    """src/geometry/mesh.jl:430-442"""
    th = (0.0, 2 * math.pi / 3, 4 * math.pi / 3)
    pts = [[math.cos(a), math.sin(a), 0.0] for a in th] + [[0.0, 0.0, -1.0 * plane_w]]
    if is_include_vis_sides:
        tri = [[0, 1, 2], [0, 2, 3], [0, 3, 1], [1, 3, 2]]
    else:
        tri = [[0, 1, 2]]
    return eMesh(pts, tri, [[3, 0, 1, 2]], [0.0, 0.0, 0.0, plane_w])


def _output_box_ind():
    """src/geometry/mesh.jl:527-550 (0-based)."""
    faces = np.array([[1, 3, 5, 7], [2, 6, 4, 8], [1, 5, 2, 6], [3, 4, 7, 8], [1, 2, 3, 4], [5, 7, 6, 8]]) - 1
    tri = []
    for bf in faces:
        tri.append(bf[[0, 2, 3]])
        tri.append(bf[[0, 3, 1]])
    tri = np.array(tri)
    tet = np.concatenate([np.full((len(tri), 1), 8), tri], axis=1)
    eps = np.zeros(9)
    eps[8] = 1.0
    return tri, tet, eps


def eMesh_box(r=1.0, c=(0.0, 0.0, 0.0)) -> eMesh:
    """src/geometry/mesh.jl:557-575"""
    pts = np.array([[-1, -1, -1], [+1, -1, -1], [-1, +1, -1], [+1, +1, -1], [-1, -1, +1], [+1, -1, +1], [-1, +1, +1], [+1, +1, +1], [0, 0, 0]],
                   dtype=np.float64)
    tri, tet, eps = _output_box_ind()
    r = np.ones(3) * np.asarray(r, dtype=np.float64)
    m = eMesh(pts, tri, tet, eps)
    transform(m, R=np.diag(r))
    transform(m, t=np.asarray(c, dtype=np.float64))
    return m


def _dedupe_points(m: eMesh) -> None:
    """mesh_repair! restricted to what the generators here need: merge coincident points (first
    occurrence wins) and drop unused ones -- src/geometry/mesh.jl:235-319."""
    from scipy.spatial import cKDTree

    prims = [a for a in (m.tri, m.tet) if a is not None]
    side = np.inf
    for a in prims:
        p = m.point[a]
        for i in range(a.shape[1]):
            for j in range(i):
                side = min(side, float(np.linalg.norm(p[:, i] - p[:, j], axis=1).min()))
    tree = cKDTree(m.point)
    groups = tree.query_ball_point(m.point, side * 0.499)
    new_key = np.array([min(g) for g in groups], dtype=np.int64)
    for a in prims:
        a[:] = new_key[a]
    used = np.zeros(len(m.point), dtype=bool)
    for a in prims:
        used[a.reshape(-1)] = True
    remap = np.cumsum(used) - 1
    for a in prims:
        a[:] = remap[a]
    m.point = np.ascontiguousarray(m.point[used])
    if m.eps is not None:
        m.eps = m.eps[used]


def _delete_opposing_triangles(m: eMesh) -> None:
    """delete_triangles! -- src/geometry/mesh.jl:322-361"""
    if m.tri is None:
        return
    seen = {}
    for k, t in enumerate(m.tri):
        seen.setdefault(tuple(sorted(int(x) for x in t)), []).append(k)
    kill = [k for v in seen.values() if len(v) == 2 for k in v]
    if kill:
        m.tri = np.delete(m.tri, sorted(kill), axis=0)


def _sub_div_triangle(p: np.ndarray, n_div: int):
    """sub_div_triangle -- src/geometry/mesh.jl:367-413 (0-based output)."""
    n_end = lambda n: (n + 1) * n // 2
    n_start = lambda n: 1 + n_end(n - 1)
    tri = []
    for k in range(1, n_div + 1):
        for kk in range(k):
            i1 = n_start(k) + kk
            i2 = i1 + k
            tri.append((i1, i2, i2 + 1))
        for kk in range(k - 1):
            i1 = n_start(k) + kk
            i2 = i1 + k + 1
            tri.append((i1, i2, i2 - k))
    pts = []
    for n_vert in range(1, n_end(n_div + 1) + 1):
        i_end_layer, i_layer = 1, 1
        while i_end_layer < n_vert:
            i_layer += 1
            i_end_layer += i_layer
        ext = 0.0 if n_vert == 1 else (i_end_layer - n_vert) / (i_layer - 1)
        phi1 = (n_div - i_layer + 1) / n_div
        phi2 = (1 - phi1) * ext
        phi3 = 1 - phi1 - phi2
        pts.append(p[0] * phi1 + p[1] * phi2 + p[2] * phi3)
    return np.array(pts), np.array(tri, dtype=np.int64) - 1


def eMesh_sphere(rad=1.0, n_div: int = 4) -> eMesh:
    """src/geometry/mesh.jl:449-525: subdivided icosahedron projected to the sphere, volumised
    about the centre (20 n_div^2 triangles and tets)."""
    phi = (1 + math.sqrt(5.0)) / 2
    v = []
    for s1 in (-1.0, 1.0):
        for s2 in (-1.0, 1.0):
            v += [[0.0, s1, phi * s2], [s1, phi * s2, 0.0], [phi * s2, 0.0, s1]]
    v = np.array(v)
    d = np.linalg.norm(v[:, None, :] - v[None, :, :], axis=2)
    b = np.isclose(d, 2.0)
    faces = []
    for i1 in range(12):
        for i2 in range(i1 + 1, 12):
            for i3 in range(i2 + 1, 12):
                if b[i1, i2] and b[i2, i3] and b[i1, i3]:
                    n = np.cross(v[i2] - v[i1], v[i3] - v[i2])
                    c = v[i1] + v[i2] + v[i3]
                    faces.append((i1, i2, i3) if np.dot(n, c) > 0 else (i1, i3, i2))
    pts_all, tri_all = [], []
    off = 0
    for f in faces:
        p, t = _sub_div_triangle(v[list(f)], n_div)
        pts_all.append(p)
        tri_all.append(t + off)
        off += len(p)
    surf = eMesh(np.concatenate(pts_all), np.concatenate(tri_all))
    _dedupe_points(surf)
    surf.point = surf.point / np.linalg.norm(surf.point, axis=1, keepdims=True)
    surf.point = surf.point * (np.ones(3) * np.asarray(rad, dtype=np.float64))
    n_vert = len(surf.point)
    tet = np.concatenate([np.full((len(surf.tri), 1), n_vert), surf.tri], axis=1)
    eps = np.concatenate([np.zeros(n_vert), [1.0]])
    point = np.concatenate([surf.point, np.zeros((1, 3))])
    return eMesh(point, surf.tri, tet, eps)


def f_swept_triv(theta: float):
    """Straight path along +y (src/geometry/mesh_create_swept.jl:19-23): position, radial direction, path direction."""
    n1 = np.array([0.0, 0.0, -1.0])
    n2 = np.array([0.0, 1.0, 0.0])
    return n2 * theta, n1, n2


def f_swept_circle(r: float, theta: float):
    """Circular path of radius r in the xy plane (src/geometry/mesh_create_swept.jl:8-12)."""
    n1 = np.array([math.cos(theta), math.sin(theta), 0.0])
    n2 = np.array([-math.sin(theta), math.cos(theta), 0.0])
    return r * n1, n1, n2


def _rodrigues(angle: float, axis: np.ndarray, v: np.ndarray) -> np.ndarray:
    a = axis / np.linalg.norm(axis)
    return v * math.cos(angle) + np.cross(a, v) * math.sin(angle) + a * float(a @ v) * (1.0 - math.cos(angle))


def _remove_degenerate(m: eMesh, tol: float = 1.0e-6) -> None:
    """remove_degenerate! (src/geometry/mesh.jl:242-255): drop primitives whose area / volume is below tol x the largest."""
    if m.tet is not None and len(m.tet):
        p = m.point[m.tet]
        vol = _vol(p[:, 0].T, p[:, 1].T, p[:, 2].T, p[:, 3].T)
        m.tet = m.tet[~(vol < vol.max() * tol)]
    if m.tri is not None and len(m.tri):
        p = m.point[m.tri]
        area = 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 1]), axis=1)
        m.tri = m.tri[~(area < area.max() * tol)]


def create_swept_mesh(fun_gen, lr, rad, n_side: int = 4, is_open: bool = True, rot_half: bool = True) -> eMesh:
    """create_swept_mesh (src/geometry/mesh_create_swept.jl:76-108): an n_side-gon swept along the path fun_gen
    through the arc-length nodes lr with (circumscribed-corrected) radii rad.  Each (path segment, side) adds 7
    points, 2 surface triangles (+1 cap triangle at an open end) and 4 tets (add_rot_sym_segment!, :25-62);
    eps = 1 on the path, 0 on the surface and at the open ends."""
    lr = np.asarray(lr, dtype=np.float64)
    rad = np.full(len(lr), float(rad)) if np.isscalar(rad) else np.asarray(rad, dtype=np.float64)
    if len(rad) != len(lr):
        raise ValueError("the length of lr and length of rad must be the same")
    d_phi = 2.0 * math.pi / n_side
    rad = rad / math.cos(d_phi / 2.0)
    pts, tris, tets, eps = [], [], [], []
    n_theta = len(lr) - 1
    for k_theta in range(n_theta):
        for k_phi in range(1, n_side + 1):
            phi0 = d_phi * (k_phi - (0.5 if rot_half else 0.0))
            phi1 = phi0 + d_phi
            open_lo = is_open and k_theta == 0
            open_hi = is_open and k_theta == n_theta - 1
            pa, xa, ya = fun_gen(float(lr[k_theta]))
            pb, xb, yb = fun_gen(float(lr[k_theta + 1]))
            seg = [pa, pb, (pa + pb) * 0.5,
                   pa + _rodrigues(phi0, ya, xa) * rad[k_theta], pb + _rodrigues(phi0, yb, xb) * rad[k_theta + 1],
                   pa + _rodrigues(phi1, ya, xa) * rad[k_theta], pb + _rodrigues(phi1, yb, xb) * rad[k_theta + 1]]
            o = len(pts)
            pts += seg
            tris += [(o + 3, o + 5, o + 6), (o + 3, o + 6, o + 4)]
            tets += [(o + 0, o + 2, o + 3, o + 5), (o + 2, o + 1, o + 4, o + 6), (o + 2, o + 3, o + 5, o + 6), (o + 3, o + 2, o + 4, o + 6)]
            e = [1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0]
            if open_lo:
                e[0] = 0.0
                tris.append((o + 0, o + 5, o + 3))
            if open_hi:
                e[1] = 0.0
                tris.append((o + 1, o + 4, o + 6))
            eps += e
    m = eMesh(np.array(pts), np.array(tris, dtype=np.int64), np.array(tets, dtype=np.int64), np.array(eps), check=False)
    _remove_degenerate(m)
    _dedupe_points(m)
    _delete_opposing_triangles(m)
    return m


def eMesh_grid_square(side: float, n_cell: int) -> eMesh:
    """A planar, +z-facing triangulated square of n_cell x n_cell cells (2 triangles each) -- the
    planar input extrude_mesh needs for the slab of config C4."""
    xs = np.linspace(-side / 2, side / 2, n_cell + 1)
    X, Y = np.meshgrid(xs, xs, indexing="xy")
    pts = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], axis=1)
    idx = lambda i, j: j * (n_cell + 1) + i
    tri = []
    for j in range(n_cell):
        for i in range(n_cell):
            a, b2, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            tri.append((a, b2, c))
            tri.append((a, c, d))
    return eMesh(pts, np.array(tri))


def extrude_mesh(surf: eMesh, thick: float) -> eMesh:
    """src/geometry/mesh.jl:600-659: each planar triangle becomes a prism of 8 tets around its
    centroid (eps = 1 at the centroid, 0 on both faces)."""
    P, T = surf.point, surf.tri
    n_hat_all = np.cross(P[T[:, 1]] - P[T[:, 0]], P[T[:, 2]] - P[T[:, 1]])
    n_hat_all /= np.linalg.norm(n_hat_all, axis=1, keepdims=True)
    if not np.allclose(n_hat_all, n_hat_all[0]):
        raise ValueError("All triangles must have the same normal.")
    n_hat = n_hat_all[0]
    n_pt, n_tri = len(P), len(T)
    point = np.concatenate([P - n_hat * thick / 2, P + n_hat * thick / 2, P[T].sum(axis=1) * (1.0 / 3.0)])
    eps = np.concatenate([np.zeros(2 * n_pt), np.ones(n_tri)])
    tri_out, tet_out = [], []
    for k, (b1, b2, b3) in enumerate(T):
        t4, t5, t6 = b1 + n_pt, b2 + n_pt, b3 + n_pt
        i_center = k + 2 * n_pt
        tri_add = [(b1, b3, b2), (t4, t5, t6)]
        for f in ((b1, b2, t5, t4), (b2, b3, t6, t5), (b1, t4, t6, b3)):
            i = int(np.argmin(f))
            f = (f[i], f[(i + 1) % 4], f[(i + 2) % 4], f[(i + 3) % 4])
            tri_add.append((f[0], f[1], f[2]))
            tri_add.append((f[0], f[2], f[3]))
        for t in tri_add:
            tri_out.append(t)
            tet_out.append((i_center,) + tuple(t))
    m = eMesh(point, np.array(tri_out), np.array(tet_out), eps)
    _delete_opposing_triangles(m)
    return m


# ------------------------------------------------------------------------------------------------
# Leaf OBB fitting (src/obb/obb_construction.jl:13-41, src/obb/util.jl:54-66)
# ------------------------------------------------------------------------------------------------
def _normalize(a):
    return a / math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def _make_obb(p: np.ndarray, i_start: int):
    i_next = (i_start % 3) + 1
    e1 = _normalize(p[i_next - 1] - p[i_start - 1])
    e3 = _normalize(np.cross(p[1] - p[0], p[2] - p[1]) * 0.5)
    e2 = np.cross(e3, e1)
    R = np.stack([e1, e2, e3], axis=1)
    proj = p @ R
    lo, hi = proj.min(axis=0), proj.max(axis=0)
    c = (hi + lo) * 0.5
    e = (hi - lo) * 0.5
    return R @ c, e, R


def fit_tri_obb(p: np.ndarray):
    return _make_obb(np.asarray(p, dtype=np.float64), 1)


_TET_PERM = {1: (2, 4, 3, 1), 2: (4, 1, 3, 2), 3: (1, 4, 2, 3), 4: (1, 2, 3, 4)}


def fit_tet_obb(p: np.ndarray, eps: np.ndarray):
    p = np.asarray(p, dtype=np.float64)
    if not (0.0 < _vol(*p)):
        raise ValueError("inverted tet")
    i = int(np.argmax(np.abs(eps))) + 1
    p = p[[k - 1 for k in _TET_PERM[i]]]
    boxes = [_make_obb(p, s) for s in (1, 2, 3)]
    area = [8 * (b[1][0] * b[1][1] + b[1][1] * b[1][2] + b[1][2] * b[1][0]) for b in boxes]
    if max(area[1], area[2]) <= area[0]:
        return boxes[0]
    if max(area[0], area[2]) <= area[1]:
        return boxes[1]
    return boxes[2]


def _make_obb_batch(p: np.ndarray, i_start: int):
    """_make_obb for a batch p[n, N, 3] (same arithmetic, vectorised over the primitives)."""
    i_next = (i_start % 3) + 1
    d = p[:, i_next - 1] - p[:, i_start - 1]
    e1 = d / np.sqrt((d * d).sum(axis=1, keepdims=True))
    nrm = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 1]) * 0.5
    e3 = nrm / np.sqrt((nrm * nrm).sum(axis=1, keepdims=True))
    e2 = np.cross(e3, e1)
    R = np.stack([e1, e2, e3], axis=2)                    # [n, 3, 3], columns are the axes
    proj = np.einsum("nki,nij->nkj", p, R)
    lo, hi = proj.min(axis=1), proj.max(axis=1)
    c = (hi + lo) * 0.5
    e = (hi - lo) * 0.5
    return np.einsum("nij,nj->ni", R, c), e, R


def fit_tri_obb_batch(p: np.ndarray):
    return _make_obb_batch(np.asarray(p, dtype=np.float64), 1)


def fit_tet_obb_batch(p: np.ndarray, eps: np.ndarray):
    p = np.asarray(p, dtype=np.float64)
    if not (0.0 < _vol(p[:, 0].T, p[:, 1].T, p[:, 2].T, p[:, 3].T)).all():
        raise ValueError("inverted tet")
    i = np.argmax(np.abs(eps), axis=1) + 1
    perm = np.array([[k - 1 for k in _TET_PERM[j]] for j in (1, 2, 3, 4)])[i - 1]
    p = np.take_along_axis(p, perm[:, :, None], axis=1)
    boxes = [_make_obb_batch(p, s) for s in (1, 2, 3)]
    area = np.stack([8 * (b[1][:, 0] * b[1][:, 1] + b[1][:, 1] * b[1][:, 2] + b[1][:, 2] * b[1][:, 0]) for b in boxes], axis=1)
    pick = np.where(np.maximum(area[:, 1], area[:, 2]) <= area[:, 0], 0, np.where(np.maximum(area[:, 0], area[:, 2]) <= area[:, 1], 1, 2))
    sel = lambda j: np.choose(pick.reshape((-1,) + (1,) * (boxes[0][j].ndim - 1)), [b[j] for b in boxes])
    return sel(0), sel(1), sel(2)


# ------------------------------------------------------------------------------------------------
# Bounding-volume tree
# ------------------------------------------------------------------------------------------------
@dataclass
class FlatTree:
    """Pre-order flattening of the reference's pointer-linked bin_BB_Tree
    (src/obb/tree_types.jl:1-16).  Node 0 is the root; leaf_id is the 0-based primitive index
    for leaves and -1 for internal nodes (the reference's id == -9999)."""
    c: np.ndarray        # (n_node, 3)
    e: np.ndarray        # (n_node, 3)
    R: np.ndarray        # (n_node, 9) column-major 3x3
    left: np.ndarray     # (n_node,) int32, -1 for leaves
    right: np.ndarray    # (n_node,) int32
    leaf_id: np.ndarray  # (n_node,) int32

    @property
    def n_node(self) -> int:
        return len(self.left)

    def depth(self) -> int:
        d = np.zeros(self.n_node, dtype=np.int64)
        for k in range(self.n_node):  # pre-order => parents come first
            if self.left[k] >= 0:
                d[self.left[k]] = d[k] + 1
                d[self.right[k]] = d[k] + 1
        return int(d.max()) + 1


class _Node:
    __slots__ = ("lo", "hi", "a", "b", "leaf")

    def __init__(self, lo, hi, a=None, b=None, leaf=-1):
        self.lo, self.hi, self.a, self.b, self.leaf = lo, hi, a, b, leaf


def _merge_box(lo1, hi1, lo2, hi2):
    """OBB(a, b) for two axis-aligned boxes (src/obb/box_types.jl:11-15 via calc_min_max):
    the boxes are stored as centre/extent, so the round trip through (c, e) is kept."""
    c1, e1 = (hi1 + lo1) * 0.5, (hi1 - lo1) * 0.5
    c2, e2 = (hi2 + lo2) * 0.5, (hi2 - lo2) * 0.5
    lo = np.minimum(np.minimum(c1 - e1, c1 + e1), np.minimum(c2 - e2, c2 + e2))
    hi = np.maximum(np.maximum(c1 - e1, c1 + e1), np.maximum(c2 - e2, c2 + e2))
    return lo, hi


def _blob_cost(lo, hi, n_below: int, scale: float) -> float:
    """blobCost -- src/geometry/blob_types.jl:73-81"""
    e = (hi - lo) * 0.5
    area = 8 * (e[0] * e[1] + e[1] * e[2] + e[2] * e[0])
    vol = 8 * e[0] * e[1] * e[2]
    return n_below * math.log2(2 * n_below) + area / scale ** 2 + vol / scale ** 3


def _neighbors(prims: np.ndarray):
    """extractTriTetNeighborInformation -- src/geometry/blob_types.jl:29-71"""
    n_v = prims.shape[1]
    shared = {}
    for k, iv in enumerate(prims):
        for j in range(n_v):
            key = tuple(sorted(int(iv[(j + d) % n_v]) for d in range(1, n_v)))
            ent = shared.get(key)
            if ent is None:
                shared[key] = [k, -1]
            else:
                if ent[1] != -1:
                    raise ValueError("three primitives share the same edge/face something is wrong")
                ent[1] = k
    nb = [set() for _ in range(len(prims))]
    for a, b in shared.values():
        if b == -1:
            if n_v == 3:
                raise ValueError("not implemented error: disconnected mesh (triangle edge without a partner)")
            continue
        nb[a].add(b)
        nb[b].add(a)
    return nb


def _top_down(nodes):
    """recursive_top_down -- src/geometry/top_down.jl:10-32"""
    n = len(nodes)
    if n == 1:
        return nodes[0]
    if n == 2:
        lo, hi = _merge_box(nodes[0].lo, nodes[0].hi, nodes[1].lo, nodes[1].hi)
        return _Node(lo, hi, nodes[0], nodes[1])
    lo, hi = nodes[0].lo, nodes[0].hi
    for nd in nodes:
        lo, hi = _merge_box(lo, hi, nd.lo, nd.hi)
    mi = int(np.argmax((hi - lo) * 0.5))
    all_c = np.array([(nd.hi[mi] + nd.lo[mi]) * 0.5 for nd in nodes])
    perm = np.argsort(all_c, kind="stable")
    n_mid = int(math.ceil(n / 2))
    t1 = _top_down([nodes[i] for i in perm[: n_mid - 1]])
    t2 = _top_down([nodes[i] for i in perm[n_mid - 1:]])
    lo, hi = _merge_box(t1.lo, t1.hi, t2.lo, t2.hi)
    return _Node(lo, hi, t1, t2)


def eMesh_to_tree(m: eMesh, method: str = "auto") -> FlatTree:
    """Deterministic restatement of eMesh_to_tree (src/geometry/blob_types.jl:136-173).

    method: "bottom_up" = the reference's cost-driven neighbour merging followed by
    recursive_top_down over whatever blobs remain; "top_down" = recursive_top_down over the
    leaves only (used for very large meshes where the pure-Python merge is slow);
    "auto" picks bottom_up up to 20 000 primitives.
    """
    if m.is_tri and m.is_tet:
        raise ValueError("Cannot create tree for eMesh{Tri,Tet} use as_tri_eMesh or as_tet_eMesh on input first.")
    prims = m.tri if m.is_tri else m.tet
    n_leaf = len(prims)
    P = m.point
    pp = P[prims]
    leaf_lo, leaf_hi = pp.min(axis=1), pp.max(axis=1)
    if n_leaf == 1:  # lone leaf keeps its axis-aligned box (blob_types.jl:139-146)
        lo, hi = leaf_lo[0], leaf_hi[0]
        return FlatTree(((hi + lo) * 0.5)[None], ((hi - lo) * 0.5)[None], np.eye(3).reshape(1, 9), np.array([-1], np.int32),
                        np.array([-1], np.int32), np.array([0], np.int32))
    if method == "auto":
        method = "bottom_up" if n_leaf <= 20000 else "top_down"
    leaves = [_Node(leaf_lo[k], leaf_hi[k], leaf=k) for k in range(n_leaf)]
    if method == "top_down":
        root = _top_down_fast(leaves, leaf_lo, leaf_hi)
    else:
        g_lo, g_hi = P.min(axis=0), P.max(axis=0)
        scale = float(((g_hi - g_lo) * 0.5).sum() / 3)
        nb = _neighbors(prims)
        blob = {}  # id -> [node, n_below, cost, neighbor-set]
        for k in range(n_leaf):
            lo, hi = _merge_box(leaf_lo[k], leaf_hi[k], leaf_lo[k], leaf_hi[k])
            blob[k] = [leaves[k], 1, _blob_cost(lo, hi, 1, scale), nb[k]]

        def marginal(a, b):
            lo, hi = _merge_box(a[0].lo, a[0].hi, b[0].lo, b[0].hi)
            return _blob_cost(lo, hi, a[1] + b[1], scale) - a[2] - b[2]

        heap = []
        for ka in range(n_leaf):
            for kb in blob[ka][3]:
                if ka < kb:
                    heap.append((marginal(blob[ka], blob[kb]), ka, kb))
        heapq.heapify(heap)
        k_next = n_leaf
        while heap:
            _, ka, kb = heapq.heappop(heap)
            if ka not in blob or kb not in blob:
                continue  # stale entry (lazy deletion)
            a, b = blob[ka], blob[kb]
            a[3].discard(kb)
            b[3].discard(ka)
            lo, hi = _merge_box(a[0].lo, a[0].hi, b[0].lo, b[0].hi)
            node_c = _Node(lo, hi, a[0], b[0])
            n_c = a[1] + b[1]
            lo2, hi2 = _merge_box(lo, hi, lo, hi)
            c = [node_c, n_c, _blob_cost(lo2, hi2, n_c, scale), a[3] | b[3]]
            kc = k_next
            k_next += 1
            del blob[ka], blob[kb]
            blob[kc] = c
            for kn in c[3]:
                other = blob[kn]
                other[3].discard(ka)
                other[3].discard(kb)
                other[3].add(kc)
                heapq.heappush(heap, (marginal(other, c), kn, kc))
        root = _top_down([blob[k][0] for k in sorted(blob)])
    return _flatten(root, m, prims)


def _top_down_fast(leaves, leaf_lo, leaf_hi):
    """recursive_top_down over leaves with numpy index arrays (same split rule)."""
    centers = (leaf_hi + leaf_lo) * 0.5

    def rec(ids):
        n = len(ids)
        if n == 1:
            return leaves[ids[0]]
        if n == 2:
            a, b = leaves[ids[0]], leaves[ids[1]]
            lo, hi = _merge_box(a.lo, a.hi, b.lo, b.hi)
            return _Node(lo, hi, a, b)
        lo, hi = leaf_lo[ids].min(axis=0), leaf_hi[ids].max(axis=0)
        mi = int(np.argmax((hi - lo) * 0.5))
        perm = np.argsort(centers[ids, mi], kind="stable")
        n_mid = int(math.ceil(n / 2))
        t1 = rec(ids[perm[: n_mid - 1]])
        t2 = rec(ids[perm[n_mid - 1:]])
        lo, hi = _merge_box(t1.lo, t1.hi, t2.lo, t2.hi)
        return _Node(lo, hi, t1, t2)

    import sys
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    return rec(np.arange(len(leaves)))


def _flatten(root: _Node, m: eMesh, prims: np.ndarray) -> FlatTree:
    """Pre-order flattening + tight_fit_leaves! (src/geometry/blob_types.jl:170-190)."""
    order = []
    stack = [root]
    while stack:
        nd = stack.pop()
        order.append(nd)
        if nd.leaf < 0:
            stack.append(nd.b)
            stack.append(nd.a)
    index = {id(nd): k for k, nd in enumerate(order)}
    n = len(order)
    c = np.zeros((n, 3))
    e = np.zeros((n, 3))
    R = np.zeros((n, 9))
    left = np.full(n, -1, np.int32)
    right = np.full(n, -1, np.int32)
    leaf_id = np.full(n, -1, np.int32)
    leaf_nodes = np.array([k for k, nd in enumerate(order) if nd.leaf >= 0])
    leaf_prims = np.array([order[k].leaf for k in leaf_nodes])
    pts = m.point[prims[leaf_prims]]
    if m.is_tri:
        cb, eb, Rb = fit_tri_obb_batch(pts)
    else:
        cb, eb, Rb = fit_tet_obb_batch(pts, m.eps[prims[leaf_prims]])
    c[leaf_nodes], e[leaf_nodes] = cb, eb
    R[leaf_nodes] = np.transpose(Rb, (0, 2, 1)).reshape(-1, 9)  # column-major
    leaf_id[leaf_nodes] = leaf_prims
    for k, nd in enumerate(order):
        if nd.leaf < 0:
            c[k] = (nd.hi + nd.lo) * 0.5
            e[k] = (nd.hi - nd.lo) * 0.5
            R[k] = np.eye(3).reshape(9)
            left[k] = index[id(nd.a)]
            right[k] = index[id(nd.b)]
    return FlatTree(c, e, R, left, right, leaf_id)


def refit_tree(tree: FlatTree, m: eMesh) -> FlatTree:
    """The boxes of an existing tree for moved vertices (same connectivity, same topology), with the construction rules of eMesh_to_tree:
    leaves get fit_tri_obb / fit_tet_obb (tight_fit_leaves!), internal nodes OBB(a, b) of their children's axis-aligned boxes, bottom-up,
    the leaves entering with calc_obb of their vertices (recursive_top_down).  Host mirror of pfc_refit_mesh."""
    prims = m.tri if m.is_tri else m.tet
    n = tree.n_node
    c, e, R = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 9))
    lo, hi = np.zeros((n, 3)), np.zeros((n, 3))
    leaf_nodes = np.nonzero(tree.leaf_id >= 0)[0]
    pts = m.point[prims[tree.leaf_id[leaf_nodes]]]
    lo[leaf_nodes], hi[leaf_nodes] = pts.min(axis=1), pts.max(axis=1)
    cb, eb, Rb = fit_tri_obb_batch(pts) if m.is_tri else fit_tet_obb_batch(pts, m.eps[prims[tree.leaf_id[leaf_nodes]]])
    c[leaf_nodes], e[leaf_nodes] = cb, eb
    R[leaf_nodes] = np.transpose(Rb, (0, 2, 1)).reshape(-1, 9)
    order = np.arange(n)
    # children always come after their parent in the pre-order flattening _flatten produces; in general, order by depth
    depth = np.zeros(n, dtype=np.int64)
    for k in range(n):
        if tree.left[k] >= 0:
            if tree.left[k] < k or tree.right[k] < k:
                raise ValueError("refit_tree expects a pre-order tree (parents before children)")
            depth[tree.left[k]] = depth[tree.right[k]] = depth[k] + 1
    for k in order[::-1]:
        if tree.left[k] >= 0:
            a, b = tree.left[k], tree.right[k]
            lo[k], hi[k] = _merge_box(lo[a], hi[a], lo[b], hi[b])
            c[k], e[k] = (hi[k] + lo[k]) * 0.5, (hi[k] - lo[k]) * 0.5
            R[k] = np.eye(3).reshape(9)
    return FlatTree(c, e, R, tree.left.copy(), tree.right.copy(), tree.leaf_id.copy())
