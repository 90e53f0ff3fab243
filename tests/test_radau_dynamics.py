"""CPU tests of the caller-side mirrors that drive the contact path the way the reference does:
the adaptive Radau IIA integrator (src/radau; KATs of test/test_radau) and calcXd! for floating-body
scenes (src/contact_algorithms_non_friction.jl:18-52), with the CPU oracle as the contact backend."""
import numpy as np
import pytest

import pfc_b200  # noqa: F401
from helpers import boxes_env_states, scene_boxes
from oracle import orc
from pfc_b200 import dynamics as D
from pfc_b200 import radau as R
from pfc_b200 import scenario as S


# ---- Butcher data generated from the definitions (the reference reads them from src/radau/table/*_rule) -------------
def test_radau_tables_satisfy_their_definitions():
    for n_rule in (1, 2, 3):
        t = R.radau_table(n_rule)
        s = t.n_stage
        assert s == 2 * n_rule - 1 and t.c[-1] == 1.0
        for k in range(1, s + 1):   # collocation conditions C(s) and B(2s - 1) for the weights
            assert np.allclose(t.A @ t.c ** (k - 1), t.c ** k / k, atol=1e-14)
        for k in range(1, 2 * s):
            assert abs(t.b @ t.c ** (k - 1) - 1.0 / k) < 1e-14
        assert np.allclose(np.linalg.inv(t.A) @ t.T, t.T * t.lam[None, :], atol=1e-12)     # inv(A) T = T diag(lambda)
        assert np.allclose(t.T @ t.Tinv, np.eye(s), atol=1e-13)
        assert abs(t.lam[0].imag) == 0.0 and abs(t.b_hat_0 - 1.0 / t.lam[0].real) < 1e-15
        r = np.array([1.0 / k for k in range(1, s + 1)])
        r[0] -= t.b_hat_0
        assert np.allclose(np.vander(t.c, s, increasing=True).T @ t.b_hat, r, atol=1e-14)
    # published constants of RADAU5 (Hairer & Wanner IV.8): nodes (4 -+ sqrt 6) / 10, real eigenvalue of inv(A); and the values in the
    # reference's table files (src/radau/table/2_rule/{c,lambda,b_hat}.txt, compared with the generator's output to 1e-14 when written)
    t = R.radau_table(2)
    assert np.allclose(t.c, [(4 - np.sqrt(6)) / 10, (4 + np.sqrt(6)) / 10, 1.0], atol=1e-15)
    assert abs(t.lam[0].real - 3.637834252744496) < 1e-13 and abs(t.lam[1] - (2.6810828736277523 + 3.0504301992474105j)) < 1e-13
    assert np.allclose(np.r_[t.b_hat_0, t.b_hat], [0.27488882959567734, -0.05189523141490083, 0.7575249005733381, 0.01948150124588532], atol=1e-14)


class _Linear:   # test/test_radau/basic_test.jl
    def de(self, xx, x, t=0.0):
        xx[:] = -1.0 * x


def test_radau_basic_exponential_decay():
    x0 = np.ones(4)
    for NC in (1, 3, 6, 10):
        rr = R.makeRadauIntegrator(_Linear(), x0, 1.0e-16, 3, NC)
        for k_rule in (1, 2, 3):
            rr.rule.s = 3
            R.update_h(rr, 0.2)
            h, x_final, _ = R.solveRadau(rr, x0)
            assert abs(np.exp(-0.2) - x_final[0]) < 1e-8 * np.exp(-0.2)


class _Robertson:   # test/test_radau/test_robertson.jl (Hairer & Wanner, eq. IV.1.4)
    def de(self, xx, x, t=0.0):
        y1, y2, y3 = x
        xx[0] = -0.04 * y1 + 1.0e4 * y2 * y3
        xx[1] = 0.04 * y1 - 1.0e4 * y2 * y3 - 3.0e7 * y2 * y2
        xx[2] = 3.0e7 * y2 * y2


def test_radau_robertson():
    x = np.array([1.0, 0.0, 0.0])
    rr = R.makeRadauIntegrator(_Robertson(), x, 1.0e-16, 2, 1)
    R.update_h(rr, 1.0e-4)
    rr.rule.s = 2
    rr.step.h_max = np.inf
    t = 0.0
    for _ in range(10):
        h, x, _ = R.solveRadau(rr, x)
        t += h
    assert 3.45e-5 < x[1] < 3.7e-5
    assert t == pytest.approx(1.0e-4 * (2 ** 10 - 1), rel=1e-12)   # h can at most double every step


class _TimeDep:   # test/test_radau/test_time_dep.jl
    def de(self, xx, x, t=0.0):
        xx[:] = t


def test_radau_time_dependent():
    rr = R.makeRadauIntegrator(_TimeDep(), np.zeros(1), 1.0e-16, 2)
    rr.rule.s = 2
    R.update_h(rr, 1.0)
    h, x_final, t_f = R.solveRadau(rr, np.zeros(1), 0.0)
    assert x_final[0] == pytest.approx(0.5, abs=1e-12) and t_f == pytest.approx(1.0)


# ---- rigid-body side of calcXd! ------------------------------------------------------------------------------------
def test_inertia_of_the_boxes():
    m, _ = scene_boxes(None)
    r, rho = 0.05, 400.0
    I_c, com, mass, vol = D.make_inertia_info(m.MeshCache[2].mesh, S.InertiaProperties(rho))          # solid cube (tet mesh)
    assert mass == pytest.approx(rho * (2 * r) ** 3, rel=1e-12) and np.allclose(com, 0, atol=1e-15)
    assert np.allclose(I_c, np.eye(3) * mass * (2 * r) ** 2 / 6, rtol=1e-12, atol=1e-18)
    I_s, com, mass, vol = D.make_inertia_info(m.MeshCache[1].mesh, S.InertiaProperties(rho, d=r))      # shell of thickness d (tri mesh)
    face = (2 * r) ** 2 * r * rho                                                                   # mass of one face
    assert mass == pytest.approx(6 * face, rel=1e-12)
    # per axis: 2 faces normal to it (a^2/6 each about their centre) + 4 faces containing it (a^2/12 + r^2 each)
    assert np.allclose(I_s, np.eye(3) * (2 * face * (2 * r) ** 2 / 6 + 4 * face * ((2 * r) ** 2 / 12 + r ** 2)), rtol=1e-12, atol=1e-18)


def test_mrp_rate_matches_rotation_derivative():
    rng = np.random.default_rng(1)
    for _ in range(5):
        p, w = rng.uniform(-0.4, 0.4, 3), rng.uniform(-2, 2, 3)
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        pc = p.astype(np.complex128) + 1e-30j * (D.mrp_rate_matrix(p) @ w)
        dR = S.mrp_to_rotation(pc).imag / 1e-30                      # d/dt R(p(t)) by complex step
        assert np.allclose(dR, S.mrp_to_rotation(p) @ K, atol=1e-12)  # = R [w]x for a body-frame angular velocity


def test_free_flight_conserves_momentum_and_energy():
    """Without gravity and contact, energy and the world-frame linear / angular momentum of every body are invariants of the
    flow: their derivative along calcXd's output vanishes (checked by a complex step along x_dot), and a short Radau run keeps
    them to the integrator's tolerance."""
    m, bodies = scene_boxes(orc.OracleContext())
    m.gravity = np.zeros(3)
    dyn = D.FloatingBodyDynamics(m)
    x0 = S.get_state(m)                      # the boxes.jl drop: nothing touches at t = 0
    rng = np.random.default_rng(2)
    x0[:m.nq] += rng.uniform(-0.2, 0.2, m.nq) * np.tile([1, 1, 1, 0, 0, 0], 4)
    x0[m.nq:m.nq + m.nv] = rng.uniform(-1, 1, m.nv)

    def invariants(x):
        out = []
        for k, b in enumerate(m.bodies):
            if b.joint is None:
                continue
            v = x[m.nq + b.v0:m.nq + b.v0 + 6]
            Rm = S.mrp_to_rotation(x[b.q0:b.q0 + 3])
            h = dyn.H[k] @ v
            out += [0.5 * v @ h, *(Rm @ h[3:]), *(Rm @ h[:3] + np.cross(x[b.q0 + 3:b.q0 + 6], Rm @ h[3:]))]   # energy, linear and angular momentum (world)
        return np.array(out)

    xd = dyn.calcXd(x0)
    rate = invariants(x0.astype(np.complex128) + 1e-30j * xd).imag / 1e-30
    assert np.abs(rate).max() <= 1e-12 * max(1.0, np.abs(invariants(x0)).max())
    rr = R.makeRadauIntegrator(dyn, S.num_x(m), 1.0e-16, 2, 6)
    ts, xs = R.integrate_radau(rr, x0, t_final=0.02, max_steps=40, after_step=lambda x: D.principal_value(m, x))
    assert ts[-1] > 0.01
    assert np.allclose(invariants(xs[0]), invariants(xs[-1]), rtol=1e-3, atol=1e-4)


def test_gravity_is_a_uniform_world_acceleration():
    m, _ = scene_boxes(orc.OracleContext())
    dyn = D.FloatingBodyDynamics(m)
    x0 = S.get_state(m)
    rng = np.random.default_rng(3)
    x0[:m.nq] += rng.uniform(-0.2, 0.2, m.nq) * np.tile([1, 1, 1, 0, 0, 0], 4)
    x0[m.nq:m.nq + m.nv] = rng.uniform(-1, 1, m.nv)
    xd = dyn.calcXd(x0)
    for b in m.bodies[1:]:
        Rm = S.mrp_to_rotation(x0[b.q0:b.q0 + 3])
        w, u = x0[m.nq + b.v0:m.nq + b.v0 + 3], x0[m.nq + b.v0 + 3:m.nq + b.v0 + 6]
        # world acceleration of the body origin = d/dt (R u) = R (u_dot + w x u): the centre of mass is the origin for the boxes
        acc = Rm @ (xd[m.nq + b.v0 + 3:m.nq + b.v0 + 6] + np.cross(w, u))
        assert np.allclose(acc, m.gravity, atol=1e-12)


def test_jacobian_chunks_match_finite_differences():
    m, _ = scene_boxes(orc.OracleContext())
    dyn = D.FloatingBodyDynamics(m)
    x = boxes_env_states(m, 1)[0]
    nx = S.num_x(m)
    J = np.zeros((nx, nx))
    for i0 in range(0, nx, 6):
        xx0, cols = dyn.de_jacobian_chunk(x, i0, i0 + 6)
        assert np.array_equal(xx0, dyn.calcXd(x))
        J[:, i0:i0 + 6] = cols
    Jfd = np.zeros((nx, nx))
    for j in range(nx):
        h = 1e-7
        xp, xm = x.copy(), x.copy()
        xp[j] += h
        xm[j] -= h
        Jfd[:, j] = (dyn.calcXd(xp) - dyn.calcXd(xm)) / (2 * h)
    assert np.abs(J - Jfd).max() <= 1e-6 * np.abs(J).max()


def test_boxes_settle_under_radau_with_the_oracle_backend():
    """A short run of test/boxes.jl's scene from a settled stack: the integrator advances, contact forces hold the stack up
    (no box falls through), and the evaluation counters show both kinds of calls the reference makes."""
    m, _ = scene_boxes(orc.OracleContext())
    dyn = D.FloatingBodyDynamics(m)
    x0 = boxes_env_states(m, 1)[0]
    x0[m.nq:m.nq + m.nv] = 0.0
    rr = R.makeRadauIntegrator(dyn, S.num_x(m), 1.0e-16, 2, 6)
    rr.step.h_max = 0.05
    ts, xs = R.integrate_radau(rr, x0, t_final=1.0, max_steps=25, after_step=lambda x: D.principal_value(m, x))
    assert len(ts) == 26 and ts[-1] > 1e-3
    assert rr.n_de_chunk == 25 * 8 and rr.n_de_float >= 25
    z = xs[-1][[5, 11, 17, 23]]
    assert np.all(np.abs(z - xs[0][[5, 11, 17, 23]]) < 0.01)
