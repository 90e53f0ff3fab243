"""Scene-level pins of the CPU oracle: the reference's analytic / known-answer tests for the
whole path, re-expressed through the host mirror (pfc_b200.scenario) with the oracle backend."""
import math

import numpy as np
import pytest

import pfc_b200  # noqa: F401
from helpers import rot_z, scene_boxes
from oracle import orc
from pfc_b200 import geometry as G
from pfc_b200 import scenario as S


@pytest.mark.parametrize("k_quad_rule", [1, 2])
def test_normal_wrench_kat(k_quad_rule):
    """test/test_normal.jl:2-49: rigid box pressed 0.1 r into a compliant half-space."""
    p_pos = (0.1, 0.2)
    r = 0.05
    m = S.MechanismScenario()
    id_plane = S.add_contact(m, "plane", G.as_tet_eMesh(G.eMesh_half_plane()), c_prop=S.ContactProperties(1.0e9))
    eM_box = G.transform(G.as_tri_eMesh(G.eMesh_box(r)), t=(0.0, 0.0, r))
    body, joint, id_box = S.add_body_contact(m, "box", eM_box, i_prop=S.InertiaProperties(400.0, d=0.09))
    ci = S.add_friction_bristle(m, id_box, id_plane, mu_d=0.3, chi=0.6, k_bar=1.0e6, tau=0.03, n_quad_rule=k_quad_rule)
    assert (ci.id_1, ci.id_2) == (id_box, id_plane)  # (Tri, Tet) ordering rule
    S.finalize(m, orc.OracleContext())
    pene = 0.1 * r
    S.set_state_spq(m, body, trans=(p_pos[0], p_pos[1], -pene))
    out = S.force_all_elastic_intersections(m)
    # zero velocity and zero bristle state => friction wrench is zero => total == normal_wrench(b)
    check = 1.0e9 * pene / 1.0 * r ** 2 * 4
    f3 = np.array([0.0, 0.0, check])
    a3 = np.cross([p_pos[0], p_pos[1], 0.0], f3)
    assert np.allclose(out["wrench"][0], -np.concatenate([a3, f3]), rtol=1e-10)
    assert out["flags"][0] == 1 and out["n_pairs"][0] > 0
    # generalized force on the box: third law => +check along z in world
    assert np.isclose(out["f_generalized"][5], check, rtol=1e-10)


def _box_and_plane(n_quad_rule, v_tol=None):
    """create_box_and_plane of test/test_friction.jl:92-128 (without the dynamics)."""
    r, E = 0.05, 1.0e9
    m = S.MechanismScenario()
    id_plane = S.add_contact(m, "plane", G.as_tet_eMesh(G.eMesh_half_plane(1.0)), c_prop=S.ContactProperties(E))
    eM = G.transform(G.eMesh_box(r * np.ones(3)), t=(0, 0, r))
    body, _, id_box = S.add_body_contact(m, "box_1", G.as_tri_eMesh(eM), i_prop=S.InertiaProperties(400.0, d=0.09))
    mu_d = 0.3
    if v_tol is not None:
        S.add_friction_regularize(m, id_plane, id_box, mu_d=mu_d, v_tol=v_tol, n_quad_rule=n_quad_rule)
        vel = (0.0, v_tol, 0.0)
    else:
        S.add_friction_bristle(m, id_plane, id_box, mu_d=mu_d, k_bar=1.0e4, tau=0.03, n_quad_rule=n_quad_rule)
        vel = (0.0, 0.0, 0.0)
    S.finalize(m, orc.OracleContext())
    mass_g = 9.8054 * 400.0 * (2 * r) ** 2 * 0.09 * 6  # any positive load works for the KAT below
    pene = mass_g / (E * 4 * r ** 2)
    S.set_state_spq(m, body, vel=vel, trans=(0.0, 0.0, -pene))
    return m, mu_d, E * pene * 4 * r ** 2


@pytest.mark.parametrize("n_quad_rule", [1, 2])
def test_regularized_friction_strength(n_quad_rule):
    """test/test_friction.jl:133-143: the box decelerates when pushed with 0.999 mu m g and
    accelerates with 1.001 mu m g, i.e. the friction force at |v_t| = v_tol lies within 0.1 % of
    mu N (the 2 um deep side faces add ~2e-5 of extra friction)."""
    m, mu_d, N = _box_and_plane(n_quad_rule, v_tol=1.0e-4)
    out = S.force_all_elastic_intersections(m)
    f = out["f_generalized"]
    # chi = default 0.5 and the box only slides tangentially => damping term is exactly 1
    assert np.isclose(f[5], N, rtol=1e-10)
    assert 0.999 * mu_d * N < -f[4] < 1.001 * mu_d * N
    assert abs(f[3]) < 1e-9 * N


def test_bristle_stiffness_analytic():
    """test/test_friction.jl:178-237: K_44 = K_55 ~ 4 hol_rad^2 k_bar E pene / hol_rad (within 1 %)."""
    box_rad, E = 0.05, 1.0e9
    hol_rad = 0.2 * box_rad
    m = S.MechanismScenario()
    _, _, id_part = S.add_body_contact(m, "part", G.as_tri_eMesh(G.eMesh_half_plane(1.0)), i_prop=S.InertiaProperties(400.0, d=0.09))
    eM = G.transform(G.eMesh_box(hol_rad * np.ones(3)), t=(0.0, 0.0, hol_rad))
    b_hol, _, id_hol = S.add_body_contact(m, "hol_1", G.as_tet_eMesh(eM), c_prop=S.ContactProperties(E), i_prop=S.InertiaProperties(400.0),
                                          joint=S.Prismatic((0.0, 0.0, 1.0)))
    k_bar = 1.0e6
    S.add_friction_bristle(m, id_part, id_hol, mu_d=0.3, chi=0.6, k_bar=k_bar, tau=0.03, n_quad_rule=2)
    ctx = orc.OracleContext()
    S.finalize(m, ctx)
    pene = hol_rad * 0.001
    S.set_configuration(m, b_hol, [-pene])
    X, tw, s = S.boundary_arrays(m, S.get_state(m))
    out = ctx.eval_f64(X, tw, s.reshape(1, 1, 6), keep=True)
    assert out["flags"][0, 0] & 1
    cop, w, K = orc.patch_stiffness(ctx.get_traction(0, 0), k_bar)
    Sinv, Kh = orc.decompose_K(K, 1.0e-3)
    Sm = np.diag(1 / Sinv)
    K2 = Sm @ np.linalg.inv(Kh @ Kh) @ Sm
    K_ana = hol_rad ** 2 * 4 * k_bar * (E * (pene / hol_rad))
    assert np.isclose(K2[3, 3], K2[4, 4], rtol=1e-9)
    assert 0.99 * K_ana < K2[4, 4] < 1.01 * K_ana


def _calc_it(t):
    """calc_it of test/test_friction.jl:239-257"""
    r = 0.05
    m = S.MechanismScenario()
    id_plane = S.add_contact(m, "plane", G.as_tet_eMesh(G.eMesh_half_plane()), c_prop=S.ContactProperties(1.0e6))
    body, _, id_box = S.add_body_contact(m, "box_1", G.as_tri_eMesh(G.eMesh_box(r)), i_prop=S.InertiaProperties(400.0, d=r))
    S.add_friction_bristle(m, id_plane, id_box, mu_d=1.0, chi=2.2, n_quad_rule=2)
    ctx = orc.OracleContext()
    S.finalize(m, ctx)
    S.set_state_spq(m, body, trans=np.asarray(t) + [0.0, 0.0, 0.99 * r], w=(0.4, 0.3, 1.0))
    X, tw, s = S.boundary_arrays(m, S.get_state(m))
    ctx.eval_f64(X, tw, s.reshape(1, 1, 6), keep=True)
    cop, _, K = orc.patch_stiffness(ctx.get_traction(0, 0), 1.0e4)
    return K, cop


def test_spatial_stiffness_translation_invariance():
    """test/test_friction.jl:259-266.  The wrench lives in frame r2 (the plane, fixed to the world)
    so translating the box translates cop and leaves K unchanged."""
    t = np.array([0.35, 0.10, 0.0])
    K_t, cop_t = _calc_it(t)
    K_0, cop_0 = _calc_it(np.zeros(3))
    assert np.allclose(K_0, K_t, rtol=1e-9, atol=1e-9 * np.abs(K_0).max())
    assert np.allclose(cop_t, cop_0 + t, rtol=1e-10)


def test_tet_tet_frictionless_spin_has_no_z_torque():
    """test/test_vol_vol.jl:2-31 integrates 5 s and checks the spin is conserved; the property that
    makes it hold is that a frictionless (mu = 0, chi = 0) tet-tet patch exerts no torque about z
    on a box spinning about z.  Also checks the normal force against E * pene * A / 2: with both
    bodies compliant (same E) the equal-pressure surface sits half-way."""
    r = 0.05
    c_prop = S.ContactProperties(1.0e6)
    m = S.MechanismScenario()
    id_plane = S.add_contact(m, "plane", G.as_tet_eMesh(G.eMesh_half_plane()), c_prop=c_prop)
    body, _, id_box = S.add_body_contact(m, "box_1", G.as_tet_eMesh(G.eMesh_box(r)), i_prop=S.InertiaProperties(400.0), c_prop=c_prop)
    ci = S.add_friction_regularize(m, id_plane, id_box, mu_d=0.0, chi=0.0, n_quad_rule=2)
    assert (ci.id_1, ci.id_2) == (id_plane, id_box)  # (Tet, Tet) keeps the given order
    S.finalize(m, orc.OracleContext())
    pene = 0.002
    S.set_state_spq(m, body, rot=rot_z(0.3), trans=(0.0, 0.0, r - pene), w=(0.0, 0.0, 1.14))
    out = S.force_all_elastic_intersections(m)
    f = out["f_generalized"]
    assert out["flags"][0] & 1
    assert abs(f[2]) < 1e-12 * abs(f[5])       # no torque about z
    assert abs(f[0]) < 1e-9 and abs(f[1]) < 1e-9
    # pressure fields: plane eps = depth / 1.0, box eps = depth_from_face / r.  Equal pressure
    # surface at depth d_p below the plane top where E d_p / 1 = E (pene - d_p) / r
    # (an O(h / r) rim of the patch lies in the box's side pyramids where the pressure is lower)
    d_p = pene / (1 + r)
    F_ana = 1.0e6 * d_p * 4 * r * r
    h = pene - d_p
    assert F_ana * (1 - 4 * h / (2 * r)) < f[5] < F_ana


def test_boxes_scene_runs_and_third_law():
    """test/boxes.jl (config C1) at a settled-stack state: each instruction finds candidate pairs,
    contact flags are set, the generalized force balances per the third law."""
    m, bodies = scene_boxes(orc.OracleContext())
    r = 0.05
    for k, b in enumerate(bodies):
        S.set_state_spq(m, b[0], rot=rot_z(0.1 * k), trans=(0.001 * k, 0.0, (2 * k + 1) * r - 1e-4 * (k + 1)), w=(0, 0, k + 1.0))
    out = S.force_all_elastic_intersections(m)
    assert (out["n_pairs"] > 0).all() and (out["flags"] & 1).all()
    assert (out["n_pairs"] <= np.array([12, 144, 144, 144])).all()
    assert (out["wrench"][:, 5] != 0).all()


def test_ordering_rule_and_errors():
    """src/mechanism_scenario.jl:298-306, 399-416, 45"""
    m = S.MechanismScenario()
    tri = G.as_tri_eMesh(G.eMesh_box(0.05))
    tet = G.as_tet_eMesh(G.eMesh_box(0.05))
    with pytest.raises(ValueError):
        S.add_contact(m, "both", G.eMesh_box(0.05), c_prop=S.ContactProperties(1e6))
    with pytest.raises(ValueError):
        S.add_contact(m, "tri", tri, c_prop=S.ContactProperties(1e6))
    with pytest.raises(ValueError):
        S.add_contact(m, "tet", tet)
    with pytest.raises(ValueError):
        S.ContactProperties(1.0)
    a = S.add_contact(m, "a", tri)
    b = S.add_contact(m, "b", tri)
    c = S.add_contact(m, "c", tet, c_prop=S.ContactProperties(1e6))
    with pytest.raises(TypeError):
        S.add_friction_regularize(m, a, b)
    ci = S.add_friction_regularize(m, c, a)
    assert (ci.id_1, ci.id_2) == (a, c)
    with pytest.raises(ValueError):
        S.add_friction_regularize(m, a, c, n_quad_rule=3)
    with pytest.raises(ValueError):
        S.add_friction_bristle(m, a, c, mu_d=0.0)
    ci = S.add_friction_regularize(m, a, c)
    assert ci.friction_model.mu_s == 0.5 and ci.friction_model.mu_d == 0.5  # default_chi quirk (:350)


def test_tree_build_properties():
    """test/test_geometry/test_blob.jl:2-18"""
    eM = G.eMesh_sphere()
    with pytest.raises(ValueError):
        G.eMesh_to_tree(eM)
    for sub, n in ((G.as_tri_eMesh(eM), eM.n_tri()), (G.as_tet_eMesh(eM), eM.n_tet())):
        tree = G.eMesh_to_tree(sub)
        leaves = tree.leaf_id[tree.leaf_id >= 0]
        assert len(leaves) == n and set(leaves.tolist()) == set(range(n))
        assert tree.depth() - 1 < math.log2(n) * 1.3
        assert tree.n_node == 2 * n - 1
    for k in range(1, 5):  # test/test_geometry/test_mesh.jl:112-132
        s = G.eMesh_sphere(2.0, k)
        assert s.n_tri() == s.n_tet() == 20 * k * k
        assert s.n_point() == 1 + 12 + (k - 1) * 30 + (k - 1) * (k - 2) // 2 * 20
        nrm = np.linalg.norm(s.point, axis=1)
        assert (nrm == 0).sum() == 1 and np.allclose(nrm[nrm > 0], 2.0)


def test_sdot_reproducibility_vs_lapack():
    """Documents the inherent sensitivity of the bristle state derivative (not a parity bar): the
    reference computes Kb^-1/2 with LAPACK's symmetric eigensolver (friction.jl:88).  Re-stating
    bristle_wrench_in_world in numpy (LAPACK eigh) on the oracle's own TractionCache reproduces the
    oracle's wrench to ~1e-7 and its s-dot only to ~1e-5 on small curved patches, because
    1 / sqrt(max(lambda, 1e-16 lambda_max)) amplifies eps-level differences by up to 1e8.  GPU-vs-
    oracle s-dot tolerances in tests/test_gpu_parity.py are derived from this."""
    m = S.MechanismScenario()
    sph = G.eMesh_sphere(0.05, 1)
    b1 = S.add_body_contact(m, "s_tri", G.as_tri_eMesh(sph), i_prop=S.InertiaProperties(400.0, d=0.01))
    b2 = S.add_body_contact(m, "s_tet", G.as_tet_eMesh(sph), i_prop=S.InertiaProperties(400.0), c_prop=S.ContactProperties(1.0e6))
    ci = S.add_friction_bristle(m, b1[2], b2[2], mu_d=0.4, k_bar=2.0e4, tau=0.05, n_quad_rule=2)
    fm = ci.friction_model
    ctx = orc.OracleContext()
    S.finalize(m, ctx)
    rng = np.random.default_rng(5)
    n_env = 48
    x = np.zeros((n_env, S.num_x(m)))
    for e in range(n_env):
        d = rng.standard_normal(3); d /= np.linalg.norm(d)
        x[e, 0:3] = rng.uniform(-0.3, 0.3, 3)
        x[e, 6:9] = rng.uniform(-0.3, 0.3, 3)
        x[e, 9:12] = d * rng.uniform(0.07, 0.098)
        x[e, 12:24] = rng.uniform(-1, 1, 12) * 0.3
        x[e, 24:] = rng.uniform(-1, 1, 6) * 1e-4
    X, tw, s = S.boundary_arrays(m, x)
    c = ctx.eval_f64(X, tw, s.reshape(n_env, 1, 6), keep=True)
    skew = lambda r: np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    worst_w = worst_s = 0.0
    n_contact = 0
    for e in range(n_env):
        if not c["flags"][e, 0] & 1:
            continue
        n_contact += 1
        tr = ctx.get_traction(e, 0)
        n, r, pdA = tr[:, 0:3], tr[:, 3:6], tr[:, 6] * tr[:, 7]
        cop = (pdA[:, None] * r).sum(0) / pdA.sum()
        K11, K12, K22 = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3))
        for i in range(len(tr)):
            q = r[i] - cop
            cx = np.cross(q, n[i])
            K22 += pdA[i] * (np.eye(3) - np.outer(n[i], n[i]))
            K12 += pdA[i] * (skew(q) - np.outer(cx, n[i]))
            K11 -= pdA[i] * (skew(q) @ skew(q) + np.outer(cx, cx))
        K = np.block([[K11, K12], [K12.T, K22]]) * fm.k_bar
        Sinv = np.array([fm.magic / np.sqrt(np.trace(K[:3, :3]))] * 3 + [1 / np.sqrt(np.trace(K[3:, 3:]))] * 3)
        lam, V = np.linalg.eigh(np.diag(Sinv) @ K @ np.diag(Sinv))
        Kh = V @ np.diag(1 / np.sqrt(np.maximum(lam, lam.max() * 1e-16))) @ V.T
        Delta = Sinv * (Kh @ s[e])
        lin, ang = np.zeros(3), np.zeros(3)
        for i in range(len(tr)):
            x2 = r[i] - cop
            Ts = -fm.k_bar * (Delta[3:] + np.cross(Delta[:3], x2) + fm.tau * (tw[e, 0, 3:] + np.cross(tw[e, 0, :3], r[i])))
            Ts = Ts - np.dot(Ts, n[i]) * n[i]
            mag = np.linalg.norm(Ts)
            if mag * mag >= fm.mu_s ** 2:
                Ts = np.clip(fm.mu_s + (mag - 2 * fm.mu_s) * (fm.mu_d - fm.mu_s) / fm.mu_s, fm.mu_d, fm.mu_s) * Ts / mag
            lin += Ts * pdA[i]
            ang += np.cross(x2, Ts * pdA[i])
        w_n = np.concatenate([(np.cross(r, pdA[:, None] * n)).sum(0), (pdA[:, None] * n).sum(0)])
        w = w_n + np.concatenate([ang + np.cross(cop, lin), lin])
        sd = -(1 / fm.tau) * (Kh @ (Sinv * np.concatenate([ang, lin])) + s[e])
        worst_w = max(worst_w, np.abs(w - c["wrench"][e, 0]).max() / np.abs(w).max())
        worst_s = max(worst_s, np.abs(sd - c["sdot"][e, 0]).max() / np.abs(sd).max())
    assert n_contact > 10
    assert worst_w < 1e-5          # the friction part of the wrench goes through the same Kb^-1/2 (via Delta)
    assert 1e-9 < worst_s < 1e-3   # s-dot is not: LAPACK vs Jacobi differ far above 1e-9


def test_c2_scenes_build_and_chain_kinematics():
    """Config C2 host mirror: the swept-mesh pencil has the reference's 48 triangles (12 tip + 24 barrel + 12 cap after
    remove_degenerate!), the spoon stand-in 5004; generalized forces through the arm chain (prismatic -> revolute ->
    prismatic pads) equal a finite-difference of the contact power  f . v = sum_k w_k . twist_k."""
    from pfc_b200 import scenes
    octx = orc.OracleContext()
    m, bodies = scenes.scene_c2_pencil(True, octx)
    assert m.MeshCache[S.find_mesh_id(m, "name")].mesh.n_tri() == 48
    assert m.MeshCache[S.find_mesh_id(m, "pad_n")].mesh.n_tet() == 320
    assert [ci.friction_model.model for ci in m.ContactInstructions] == [1, 1, 0, 0]
    xs = scenes.pencil_sample_states(m, bodies, 8)
    n_checked = 0
    for x in xs:
        X, tw, s = S.boundary_arrays(m, x)
        out = octx.eval_f64(X, tw, s.reshape(1, m.n_bristle, 6))
        w = out["wrench"][0]
        f = S.generalized_forces(m, x, w)
        # power balance: the wrench on body 2 (in r2) times the relative twist of r2 w.r.t. r1 (in r2) summed over
        # instructions equals f_generalized . v  (third law + J' w, non_friction.jl:267-286)
        p_contact = float(sum(w[k] @ tw[0, k] for k in range(len(w))))
        p_joint = float(f @ x[m.nq:m.nq + m.nv])
        if np.abs(w).max() > 0:
            assert abs(p_contact - p_joint) <= 1e-9 * max(abs(p_contact), np.abs(w).max() * np.abs(tw).max()), (p_contact, p_joint)
            n_checked += 1
    assert n_checked >= 6
    m2, _ = scenes.scene_c2_spoon(octx.__class__())
    assert m2.MeshCache[S.find_mesh_id(m2, "spoon")].mesh.n_tri() == 5004
