"""Pins the CPU oracle to the reference's own unit tests (restated; the reference is Julia and
cannot run here).  Each test names the reference test it restates."""
import math

import numpy as np
import pytest

from helpers import angle_axis, rot_x, rot_y, rot_z, rotation_between
from oracle import orc

I3 = np.eye(3)


# ---- test/test_obb/test_intersection.jl ---------------------------------------------------------
FACES = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]
CORNERS = [(-1, -1, -1), (1, -1, -1), (-1, 1, -1), (1, 1, -1), (-1, -1, 1), (1, -1, 1), (-1, 1, 1), (1, 1, 1)]
EDGES = [(0, -1, -1), (0, 1, -1), (0, -1, 1), (0, 1, 1), (-1, 0, -1), (1, 0, -1), (-1, 0, 1), (1, 0, 1), (-1, -1, 0), (1, -1, 0), (-1, 1, 0),
         (1, 1, 0)]
TOL = 1.0e-6


def _face_corner(dir_1, dir_2):
    e1, e2 = np.array([1.0, 2.0, 3.0]), np.array([2.1, 2.2, 2.3])
    a, b = (np.zeros(3), e1, I3), (np.zeros(3), e2, I3)
    vec_1, vec_2 = np.array(dir_1) * e1, np.array(dir_2) * e2
    R = rotation_between(dir_2, dir_1)
    sep = vec_1 + R @ vec_2
    return orc.sat(a, b, R, sep * (1 - TOL)), orc.sat(a, b, R, sep * (1 + TOL))


def test_sat_face_corner():
    """test_intersection.jl:77-88"""
    for f in FACES:
        for c in CORNERS:
            pos, neg = _face_corner(f, c)
            assert pos and not neg
            pos, neg = _face_corner(c, f)
            assert pos and not neg


def test_sat_edge_edge():
    """test_intersection.jl:90-104"""
    rng = np.random.default_rng(7)
    one = np.ones(3)
    a = b = (np.zeros(3), one, I3)
    for k_edge in range(6):
        for theta_axis in rng.random(15) * 2 * math.pi:
            for k in range(5):
                theta_extra = k * math.pi / 2
                vec_1 = np.array(EDGES[k_edge], float)
                for Rx in (rot_x, rot_y, rot_z):
                    R = angle_axis(theta_axis, vec_1) @ Rx(theta_extra)
                    sep = vec_1 * 2
                    assert orc.sat(a, b, R, sep * (1 - TOL))
                    assert not orc.sat(a, b, R, sep * (1 + TOL))


# ---- test/test_clip/*.jl --------------------------------------------------------------------------
def _roll_tet(rng):
    while True:
        v = rng.standard_normal((4, 3))
        if orc.lib().orc_tet_volume(np.ascontiguousarray(v.reshape(12))) >= 0.25:
            return v


def _A(tet):
    A = np.ones((4, 4))
    A[:3, :] = tet.T
    return A


def _tri_area(a, b, c, n):
    return float(n @ (np.cross(b - a, c - b) * 0.5))


def _tri_normal(a, b, c):
    v = np.cross(b - a, c - b) * 0.5
    return v / np.linalg.norm(v)


def test_clip_plane_tet_properties():
    """test_plane_tet_intersection.jl:15-62"""
    rng = np.random.default_rng(11)
    n0 = n3 = n4 = 0
    for _ in range(60):
        tet = _roll_tet(rng)
        inv_A = np.linalg.inv(_A(tet))
        for _ in range(60):
            n = rng.standard_normal(3)
            n /= np.linalg.norm(n)
            plane = np.array([n[0], n[1], n[2], rng.standard_normal()])
            c = orc.clip_plane_tet(plane, tet)
            d = tet @ n + plane[3]
            zeta = lambda p: inv_A @ np.append(p, 1.0)
            if len(c) == 3:
                assert (d < 0).sum() == 1 or (d > 0).sum() == 1
                assert np.allclose(_tri_normal(*c), n)
                n3 += 1
            elif len(c) == 4:
                assert (d < 0).sum() == 2 and (d > 0).sum() == 2
                for k in range(4):
                    assert np.allclose(_tri_normal(c[k], c[(k + 1) % 4], c[(k + 2) % 4]), n)
                n4 += 1
            else:
                assert (d <= 0).all() or (d >= 0).all()
                n0 += 1
            for p in c:
                assert (np.abs(zeta(p)) < 1e-13).sum() >= 2  # on an edge of the tet
                assert abs(p @ n + plane[3]) <= 1e-13
    assert n0 and n3 and n4


def _make_4_sided(rng):
    while True:
        tet = _roll_tet(rng)
        n = rng.standard_normal(3)
        n /= np.linalg.norm(n)
        plane = np.array([n[0], n[1], n[2], rng.standard_normal()])
        c = orc.clip_plane_tet(plane, tet)
        if len(c) == 4:
            return c


def _min_area(poly, n, r):
    if len(poly) == 0:
        return -math.inf
    return min(_tri_area(poly[k], poly[(k + 1) % len(poly)], r, n) for k in range(len(poly)))


def test_static_clip_monte_carlo():
    """test_static_clip.jl:13-64 (bounded sample; the reference loops until 3 octagons appear)."""
    rng = np.random.default_rng(5)
    tol = 1.0e-13
    n_empty, n_hits = 0, np.zeros(9, int)
    for _ in range(6000):
        r_orig = _make_4_sided(rng)
        n = _tri_normal(r_orig[0], r_orig[1], r_orig[2])
        plane = np.append(n, -n @ r_orig[0])
        tet = _roll_tet(rng)
        A = _A(tet)
        inv_A = np.linalg.inv(A)
        zeta_orig = (inv_A @ np.concatenate([r_orig, np.ones((4, 1))], axis=1).T).T
        zeta_clip = orc.clip_in_tet_coordinates(zeta_orig)
        r_clip = (A @ zeta_clip.T).T[:, :3] if len(zeta_clip) else np.zeros((0, 3))
        for p in r_clip:
            assert abs(p @ n + plane[3]) < 2000 * tol
        for _ in range(6):
            q = rng.standard_normal(3)
            q = q - (q @ n + plane[3]) * n
            min_zeta = (inv_A @ np.append(q, 1.0)).min()
            a_orig, a_clip = _min_area(r_orig, n, q), _min_area(r_clip, n, q)
            if tol < a_clip:
                assert -tol < min_zeta and -tol < a_orig
            else:
                assert min_zeta < tol or a_orig < tol
        if len(r_clip) == 0:
            n_empty += 1
        else:
            n_hits[len(r_clip)] += 1
    assert n_empty > 1000
    assert n_hits[3:8].all()  # every vertex count 3..7 reached in this sample


def test_clip_reaches_octagon_and_asymmetry():
    """An octagon needs a quad cut by all four faces; built deterministically: a large square
    through a regular tet's mid-section is a quad, shrunk so each tet face trims one corner."""
    # regular tet, plane z = const through it gives triangle/quad sections; search a seeded sample
    rng = np.random.default_rng(123)
    best = 0
    for _ in range(200000):
        r_orig = _make_4_sided(rng)
        tet = _roll_tet(rng)
        inv_A = np.linalg.inv(_A(tet))
        zeta = (inv_A @ np.concatenate([r_orig, np.ones((4, 1))], axis=1).T).T
        n = len(orc.clip_in_tet_coordinates(zeta))
        best = max(best, n)
        if best == 8:
            break
    assert best == 8


def test_centroid_exact():
    """test_poly_eight.jl:2-28"""
    p1, p2, p3, p4 = np.array([0.0, 0, 0]), np.array([1.0, 0, 0]), np.array([1.0, 1, 0]), np.array([0.0, 1, 0])
    n = _tri_normal(p1, p2, p3)
    a, c = orc.poly_centroid([p1, p2, p3, p4], n)
    assert a == 1.0 and (c == [0.5, 0.5, 0.0]).all()
    a, c = orc.poly_centroid([p1, p2, p3, p4, p1, p1, p1, p1], n)
    assert a == 1.0 and (c == [0.5, 0.5, 0.0]).all()
    a, c = orc.poly_centroid([p1, p2, p2, p3, p4], n)
    assert a == 1.0 and (c == [0.5, 0.5, 0.0]).all()
    a, c = orc.poly_centroid([p1, p2, p4], n)
    assert a == 0.5 and (c == [1 / 3, 1 / 3, 0.0]).all()
    a, c = orc.poly_centroid([p1, p2, p2], n)
    assert a == 0.0 and not np.isnan(c).any()


def test_zero_small_coordinates_exact():
    """test_poly_eight.jl:30-51"""
    rng = np.random.default_rng(3)
    for i_size in range(1, 9):
        for i_vert in range(i_size):
            for i_ind in range(4):
                A = rng.random((8, 4)) + 0.5
                A[i_vert, i_ind] = (rng.random() - 0.5) * 3.0e-15
                out = orc.zero_small_coordinates(A[:i_size])
                for v in range(i_size):
                    for i in range(4):
                        if v == i_vert and i == i_ind:
                            assert out[v, i] == 0.0
                        else:
                            assert out[v, i] == A[v, i]


# ---- test/test_friction.jl -------------------------------------------------------------------------
def test_calc_clamped_piecewise():
    """test_friction.jl:17-31"""
    x1, x2, y1, y2 = 0.3, 0.5, 1.1, 0.1
    f = orc.calc_clamped_piecewise
    eps = np.finfo(float).eps
    assert np.isclose(y1, f(x1 - 0.1, x1, x2, y1, y2))
    assert np.isclose(y1, f(x1, x1, x2, y1, y2))
    assert np.isclose(y1, f(x1 + 10 * eps, x1, x2, y1, y2))
    assert np.isclose((y1 + y2) / 2, f((x1 + x2) / 2, x1, x2, y1, y2))
    assert np.isclose(y2, f(x2 - 10 * eps, x1, x2, y1, y2))
    assert np.isclose(y2, f(x2, x1, x2, y1, y2))
    assert np.isclose(y2, f(x2 + 0.1, x1, x2, y1, y2))


MU_S, MU_D, V_C, P_DA = 1.1, 0.3, 1.0e-4, 0.133
DIR = np.array([1.0, 2.0, 3.0]) / math.sqrt(14.0)


def test_traction_bristle_law():
    """test_friction.jl:33-79"""
    Ts_mu_s, Ts_mu_d = 2 * MU_S, 3 * MU_S
    for mag in np.linspace(0.0, 4 * MU_S, 100):
        Ts = mag * DIR
        Tc = orc.traction_bristle(0.01, 1000.0, MU_S, MU_D, 1.0e-2, Ts, P_DA)
        if mag <= MU_S:
            ref = Ts * P_DA
        elif mag <= Ts_mu_s:
            ref = MU_S * Ts / mag * P_DA
        elif Ts_mu_d <= mag:
            ref = MU_D * Ts / mag * P_DA
        else:
            w_d = (mag - Ts_mu_s) / (Ts_mu_d - Ts_mu_s)
            ref = (MU_S * (1 - w_d) + MU_D * w_d) * Ts / mag * P_DA
        assert np.allclose(Tc, ref, rtol=1e-12, atol=1e-15)


def test_traction_regularized_law():
    """test_friction.jl:81-90"""
    v_mu_s, v_mu_d = 2 * V_C, 3 * V_C
    for mag in np.linspace(0.0, 4 * V_C, 100):
        vel = mag * DIR
        Tc = orc.traction_regularized(V_C, MU_S, MU_D, vel, P_DA)
        if mag <= V_C:
            ref = -MU_S * vel / V_C * P_DA
        elif mag <= v_mu_s:
            ref = -MU_S * vel / mag * P_DA
        elif v_mu_d <= mag:
            ref = -MU_D * vel / mag * P_DA
        else:
            w_d = (mag - v_mu_s) / (v_mu_d - v_mu_s)
            ref = -(MU_S * (1 - w_d) + MU_D * w_d) * vel / mag * P_DA
        assert np.allclose(Tc, ref, rtol=1e-11, atol=1e-18)


def _rand_pd(rng, n=6):
    U, s, _ = np.linalg.svd(rng.standard_normal((n, n)))
    return U @ np.diag(s) @ U.T


def test_decompose_K_identities_f64_and_dual():
    """test_friction.jl:163-176 (the reference runs it for Float64 and a 9-partial Dual; the
    chunk size on the path is 6)."""
    rng = np.random.default_rng(2)
    M = np.diag([1.0, 1.0, 1.0, 1000, 1000, 1000])
    K = M @ _rand_pd(rng) @ M
    magic = 1.0e-2
    Sinv, Kh = orc.decompose_K(K, magic)
    Kbar = np.linalg.matrix_power(np.linalg.inv(Kh), 2)
    t1, t2 = np.trace(Kbar[:3, :3]), np.trace(Kbar[3:, 3:])
    assert np.isclose(t1, t2 * magic ** 2)
    Sm = np.diag(1 / Sinv)
    assert np.allclose(K, Sm @ Kbar @ Sm, rtol=1e-9)
    # Dual: value part obeys the same identities, partials match central finite differences
    dK = np.stack([(lambda B: B + B.T)(rng.standard_normal((6, 6))) for _ in range(6)], axis=-1)
    dK = np.einsum("ij,jkd,kl->ild", M, dK, M)
    K7 = np.concatenate([K[..., None], dK], axis=-1)
    Sinv7, Kh7 = orc.decompose_K_dual6(K7, magic)
    assert np.allclose(Sinv7[:, 0], Sinv, rtol=1e-13) and np.allclose(Kh7[..., 0], Kh, rtol=1e-10, atol=1e-14)
    for d in range(6):
        h = 1e-6
        Sp, Khp = orc.decompose_K(K + h * dK[..., d], magic)
        Sm_, Khm = orc.decompose_K(K - h * dK[..., d], magic)
        assert np.allclose(Sinv7[:, 1 + d], (Sp - Sm_) / (2 * h), rtol=1e-5, atol=1e-9 * np.abs(Sinv).max())
        fd = (Khp - Khm) / (2 * h)
        assert np.allclose(Kh7[..., 1 + d], fd, rtol=1e-4, atol=1e-6 * np.abs(fd).max())


def test_inv44_and_obb_fit_against_numpy():
    """StaticArrays' 4x4 inv and the leaf OBB fit are third-party / setup-time; the oracle's
    restatement is cross-checked against numpy and the independent numpy fit in pfc_b200.geometry."""
    from pfc_b200 import geometry as G
    rng = np.random.default_rng(9)
    for _ in range(50):
        tet = _roll_tet(rng)
        A = _A(tet)
        assert np.allclose(orc.inv44(A), np.linalg.inv(A), rtol=1e-11, atol=1e-13)
        eps = np.zeros(4)
        eps[rng.integers(4)] = 1.0
        c, e, R = orc.fit_tet_obb(tet, eps)
        c2, e2, R2 = G.fit_tet_obb(tet, eps)
        assert np.allclose(c, c2) and np.allclose(e, e2) and np.allclose(R, R2)
        assert np.allclose(R.T @ R, np.eye(3), atol=1e-13)
        local = (tet - c) @ R
        assert (np.abs(local) <= e + 1e-12).all()  # the box contains the tet
        c, e, R = orc.fit_tri_obb(tet[:3])
        c2, e2, R2 = G.fit_tri_obb(tet[:3])
        assert np.allclose(c, c2) and np.allclose(e, e2) and np.allclose(R, R2)
        assert abs(e[2]) < 1e-14  # a triangle's box is flat


# ---- the reference's remaining unit tests of kernels ON the path, restated on the oracle (through oracle_capi.cpp::orc_kat) ----------
def test_weight_poly_known_answers():
    """test/test_math_kernel/test_utility.jl:9-15 (weightPoly, src/math_kernel/utility.jl:21-26): the three exact cases."""
    p1, p2 = np.array([1.0, 2.0, 3.0]), np.array([2.0, 3.0, 4.0])
    assert np.array_equal(orc.kat(0, [*p1, *p2, 1.0, 0.0], 3), p2)
    assert np.array_equal(orc.kat(0, [*p1, *p2, 0.0, 1.0], 3), p1)
    assert np.array_equal(orc.kat(0, [*p1, *p2, -0.7, 0.7], 3), (p1 + p2) * 0.5)


def test_vector_projections_known_answers():
    """test/test_math_kernel/test_vector_projections.jl:1-25: vec_sub_vec_proj and a_dot_one_pad_b, exact cases."""
    n = [0.0, 0.0, 1.0]
    assert np.array_equal(orc.kat(1, [1.0, 0.0, 0.0, *n], 3), [1.0, 0.0, 0.0])
    assert np.array_equal(orc.kat(1, [*n, *n], 3), [0.0, 0.0, 0.0])
    assert np.array_equal(orc.kat(1, [0.0, 1.0, 1.0, *n], 3), [0.0, 1.0, 0.0])
    a = [1.0, 2.0, 3.0, 4.0]
    assert orc.kat(2, [*a, 2.0, 0.0, 0.0], 1)[0] == 6.0
    assert orc.kat(2, [*a, 1.0, 2.0, 0.0], 1)[0] == 9.0
    assert orc.kat(2, [*a, 1.0, 2.0, 3.0], 1)[0] == 18.0


def test_triangle_kernels_known_answers():
    """test/test_math_kernel/test_geometry_kernel.jl:6-16: area 1/2, centroid (1/3, 1/3, 0), normal (0, 0, 1) of the unit right triangle;
    the tetrahedron cases (:18-25) through orc_tet_volume: volume 1/6."""
    out = orc.kat(3, [0, 0, 0, 1, 0, 0, 0, 1, 0], 7)
    assert abs(out[0] - 0.5) <= 1e-15
    assert np.allclose(out[1:4], [1 / 3, 1 / 3, 0.0], rtol=0, atol=1e-16)
    assert np.allclose(out[4:7], [0.0, 0.0, 1.0], rtol=0, atol=1e-16)
    tet = np.ascontiguousarray([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
    assert abs(abs(orc.lib().orc_tet_volume(tet)) - 1 / 6) <= 1e-16


def test_quadrature_rules_properties():
    """test/test_clip/test_quadrature.jl:1-21 for the two triangle rules the path uses (n_quad_rule 1 and 2, src/clip/quadrature.jl:21-41):
    weights sum to 1, every point's barycentric coordinates sum to 1, the weighted points average to the centroid, and the three points
    of rule 2 are permutations of one another."""
    for rule, n_point in ((1, 1), (2, 3)):
        out = orc.kat(4, [rule], 13)
        n = int(out[0])
        assert n == n_point
        w, zeta = out[1:1 + n], out[4:13].reshape(3, 3)[:n]
        assert abs(w.sum() - 1.0) <= 1e-15
        assert np.abs(zeta.sum(axis=1) - 1.0).max() <= 1e-15
        assert np.abs((w[:, None] * zeta).sum(axis=0) - 1 / 3).max() <= 1e-15
    z2 = np.sort(orc.kat(4, [2], 13)[4:13].reshape(3, 3), axis=1)
    assert np.array_equal(z2[0], z2[1]) and np.array_equal(z2[1], z2[2])


def test_basic_dh_algebra():
    """test/test_math_kernel/test_basic_dh.jl:56-75 (inverse, multiply, dh_vector_mul) on the oracle's homogeneous-transform helpers, the
    ones BB_BB_intersect composes (src/obb/bb_intersection.jl:2-12)."""
    rng = np.random.default_rng(5)
    for _ in range(20):
        q, _r = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        t, p = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        out = orc.kat(5, [*q.reshape(9), *t, *p], 9)
        assert np.abs(out[0:3] - (q @ p + t)).max() <= 1e-15 * 4          # dh_vector_mul(dh, p) == R p + t
        assert np.abs(out[3:6] - p).max() <= 1e-14                         # inv(dh) * dh == I
        assert np.abs(out[6:9] - (q @ (q @ p + t) + t)).max() <= 1e-14     # (dh * dh).mat == dh.mat * dh.mat
