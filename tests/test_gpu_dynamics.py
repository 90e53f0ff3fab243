"""GPU tests of the rows that sit AFTER the contact-wrench path in the reference's call stack:
pfc_calcxd_f64 (calcXd! on the device for floating-body scenes) against the host mirror fed by the CPU oracle, and
"state parity over N Radau steps" (BASELINE.json configs[0]): the same adaptive Radau IIA integration of test/boxes.jl
driven once by the CUDA library and once by the CPU oracle, through the same de / Dual-6 Jacobian-chunk interface."""
import numpy as np
import pytest

import pfc_b200  # noqa: F401
from helpers import boxes_env_states, scene_boxes
from oracle import orc
from pfc_b200 import dynamics as D
from pfc_b200 import radau as R
from pfc_b200 import scenario as S

pytestmark = pytest.mark.gpu
TOL = 1.0e-9


def _ctx():
    from pfc_b200 import capi
    return capi.Context(0)


def test_calcxd_device_matches_host_mirror_with_oracle_contact():
    """x_dot = [q_dot; v_dot; s_dot] of 256 boxes.jl environments: device (prologue, contact kernels, J' w, 6x6 solves, MRP rates)
    against calcXd! restated on the host with the oracle's wrenches; each 3-vector of x_dot within 1e-9 of the reference's."""
    n_env = 256
    m_gpu = scene_boxes(_ctx(), max_env=n_env)[0]
    m_cpu = scene_boxes(orc.OracleContext())[0]
    assert m_gpu.device_dynamics
    dyn = D.FloatingBodyDynamics(m_cpu)
    x = boxes_env_states(m_gpu, n_env)
    out = m_gpu.backend.calcxd_f64(x)
    ref = np.array([dyn.calcXd(x[e]) for e in range(n_env)])
    assert (out["flags"] & 1).sum() > n_env
    for e in range(n_env):
        for i in range(0, ref.shape[1], 3):
            den = max(np.abs(ref[e, i:i + 3]).max(), 1e-9 * np.abs(ref[e]).max())
            assert np.abs(out["xdot"][e, i:i + 3] - ref[e, i:i + 3]).max() <= TOL * den, (e, i)
    # external generalized forces enter the right-hand side like the controller's tau_ext (sum_all_forces!)
    rng = np.random.default_rng(5)
    tau = rng.uniform(-1, 1, (n_env, m_gpu.nv))
    out_t = m_gpu.backend.calcxd_f64(x, tau)
    dv = out_t["xdot"][:, m_gpu.nq:] - out["xdot"][:, m_gpu.nq:]
    for k, b in enumerate(m_cpu.bodies):
        if b.joint is not None:
            want = tau[:, b.v0:b.v0 + 6] @ dyn.Hinv[k].T
            assert np.abs(dv[:, b.v0:b.v0 + 6] - want).max() <= 1e-9 * np.abs(want).max()
    assert np.array_equal(out_t["xdot"][:, :m_gpu.nq], out["xdot"][:, :m_gpu.nq])


def test_calcxd_dual6_device_matches_host_mirror_with_oracle_contact():
    """Jacobian chunks d x_dot / d x[6k : 6k + 6] of 16 boxes.jl environments, everything on the device (Dual-6 prologue, contact,
    J' w, rigid-body terms), against the host mirror (complex-step rigid-body terms + the oracle's Dual-6 wrenches)."""
    n_env = 16
    m_gpu = scene_boxes(_ctx(), max_env=n_env)[0]
    m_cpu = scene_boxes(orc.OracleContext())[0]
    dyn = D.FloatingBodyDynamics(m_cpu)
    x = boxes_env_states(m_gpu, n_env)
    nx = S.num_x(m_gpu)
    for i0 in range(0, nx, 6):
        out = m_gpu.backend.calcxd_dual6(x, i0)
        for e in range(n_env):
            xx0, cols = dyn.de_jacobian_chunk(x[e], i0, i0 + 6)
            g = out["xdot7"][e]
            for i in range(0, nx, 3):
                den = max(np.abs(xx0[i:i + 3]).max(), 1e-9 * np.abs(xx0).max())
                assert np.abs(g[i:i + 3, 0] - xx0[i:i + 3]).max() <= TOL * den
            for d in range(6):   # column by column: ||delta||_inf <= 1e-9 ||column||_inf (+ a floor for columns that vanish)
                col = cols[:, d]
                assert np.abs(g[:, 1 + d] - col).max() <= TOL * max(np.abs(col).max(), 1e-6 * np.abs(cols).max()), (i0, e, d)


@pytest.mark.parametrize("bristle", [False, True])
def test_whole_jacobian_equals_its_chunks_bit_for_bit(bristle):
    """pfc_calcxd_jacobian (one Float64 broad phase, the seed chunks as a grid axis of the Dual kernels; bristle scenes chunk by chunk over
    the shared pair lists) against ceil(n_x / 6) calls of pfc_calcxd_dual6: the same bits, value and partials; then against the host
    mirror + oracle like the chunk test.  n_x = 48 (8 chunks) and, with two bristle instructions, 60 (10 chunks)."""
    n_env = 12
    m_gpu = scene_boxes(_ctx(), max_env=n_env, bristle=bristle)[0]
    x = boxes_env_states(m_gpu, n_env)
    nx = S.num_x(m_gpu)
    assert nx == (60 if bristle else 48)
    rng = np.random.default_rng(11)
    tau = rng.uniform(-1, 1, (n_env, m_gpu.nv))
    whole = m_gpu.backend.calcxd_jacobian(x, tau)
    assert (whole["flags"] & 1).any()
    for i0 in range(0, nx, 6):
        i1 = min(i0 + 6, nx)
        chunk = m_gpu.backend.calcxd_dual6(x, i0, tau)
        assert np.array_equal(whole["jac"][:, :, i0:i1], chunk["xdot7"][:, :, 1:1 + (i1 - i0)]), i0
        if i0 == 0:   # x_dot is the value part of the first pass (an instruction none of whose bodies is seeded is summed in Float64 order:
            assert np.array_equal(whole["xdot"], chunk["xdot7"][:, :, 0])   # the value parts of different passes differ by rounding)
        else:
            assert np.abs(whole["xdot"] - chunk["xdot7"][:, :, 0]).max() <= 1e-12 * np.abs(whole["xdot"]).max()
        assert np.array_equal(whole["n_pairs"], chunk["n_pairs"]) and np.array_equal(whole["flags"], chunk["flags"])
    # and the Float64 evaluation agrees with the value part (same pair lists; Dual value parts are Float64 operations)
    plain = m_gpu.backend.calcxd_f64(x, tau)
    assert np.abs(plain["xdot"] - whole["xdot"]).max() <= 1e-9 * np.abs(plain["xdot"]).max()
    # against the host mirror (complex-step rigid-body terms + the oracle's Dual-6 wrenches), two environments (tau_ext is additive: the
    # partials do not depend on it)
    m_cpu = scene_boxes(orc.OracleContext(), bristle=bristle)[0]
    dyn = D.FloatingBodyDynamics(m_cpu)
    for e in (0, n_env - 1):
        for i0 in range(0, nx, 6):
            i1 = min(i0 + 6, nx)
            _, cols = dyn.de_jacobian_chunk(x[e], i0, i1)
            for d in range(i1 - i0):
                col = cols[:, d]
                assert np.abs(whole["jac"][e, :, i0 + d] - col).max() <= TOL * max(np.abs(col).max(), 1e-6 * np.abs(cols).max()), (e, i0, d)


def test_whole_jacobian_large_path_scene():
    """The same on a scene whose one instruction takes the LARGE path (tet-tet sphere on slab, ~9 k candidate pairs): the Dual kernel reads
    the sorted pair list of the Float64 traversal, n_x = 12 (2 chunks).  Whole Jacobian == chunk calls bit for bit; columns vs the host
    mirror + oracle within 1e-9.  The sphere sits at a generic pose: in the scene's own pose (centred, unrotated over a regular grid) mesh
    vertices lie EXACTLY on faces of the other mesh, the wrench has kinks there, and which one-sided derivative forward-mode AD returns is
    decided by rounding noise in a sign test -- two correct implementations differ by 1e-6 of a column that vanishes by symmetry (measured),
    although the values agree to 1e-14."""
    from pfc_b200 import scenes
    m_gpu, x = scenes.scene_c4_sphere_on_slab(24, 27, backend=_ctx())
    assert m_gpu.device_dynamics
    nx = S.num_x(m_gpu)
    x = x.copy()
    x[0:3] = [0.011, -0.017, 0.023]        # MRP
    x[3:6] = [0.0123, -0.0071, 0.0451]     # translation
    whole = m_gpu.backend.calcxd_jacobian(x[None, :])
    assert whole["n_pairs"].sum() > 5000 and (whole["flags"] & 1).all()
    for i0 in range(0, nx, 6):
        chunk = m_gpu.backend.calcxd_dual6(x[None, :], i0)
        assert np.array_equal(whole["jac"][:, :, i0:i0 + 6], chunk["xdot7"][:, :, 1:7]), i0
    m_cpu, _ = scenes.scene_c4_sphere_on_slab(24, 27, backend=orc.OracleContext(n_threads=orc.lib().orc_max_threads()))
    dyn = D.FloatingBodyDynamics(m_cpu)
    for i0 in range(0, nx, 6):
        xx0, cols = dyn.de_jacobian_chunk(x, i0, i0 + 6)
        assert np.abs(whole["xdot"][0] - xx0).max() <= TOL * np.abs(xx0).max()
        for d in range(6):
            col = cols[:, d]
            assert np.abs(whole["jac"][0, :, i0 + d] - col).max() <= TOL * max(np.abs(col).max(), 1e-6 * np.abs(cols).max()), (i0, d)


def _integrate(backend, x0, n_steps, h_max=0.05, device=False):
    m = scene_boxes(backend)[0]
    dyn = D.FloatingBodyDynamics(m, device=device)
    rr = R.makeRadauIntegrator(dyn, S.num_x(m), 1.0e-16, 2, 6)
    rr.step.h_max = h_max
    ts, xs = R.integrate_radau(rr, x0, t_final=1e9, max_steps=n_steps, after_step=lambda x: D.principal_value(m, x))
    return ts, xs, rr, m


@pytest.mark.parametrize("start", ["drop", "settled"])
def test_state_parity_over_radau_steps(start):
    """test/boxes.jl under the reference's integrator, CUDA backend vs oracle backend: the step-size sequences are identical
    and the integrated states agree within 1e-9 relative (per 3-vector of the state) after every one of N steps.  'drop' is the
    reference's own initial condition (boxes 3 r apart, spinning; the first contacts happen during the run); 'settled' starts
    in contact."""
    m0 = scene_boxes(None)[0]
    if start == "drop":
        x0, n_steps = S.get_state(m0), 120
    else:
        x0, n_steps = boxes_env_states(m0, 1)[0], 60
    ts_g, xs_g, rr_g, m_g = _integrate(_ctx(), x0, n_steps, device=(start == "drop"))   # 'drop': calcXd! and its Jacobian entirely on the device
    ts_c, xs_c, rr_c, _ = _integrate(orc.OracleContext(), x0, n_steps)
    assert len(ts_g) == len(ts_c) == n_steps + 1
    assert (rr_g.n_de_float, rr_g.n_de_chunk) == (rr_c.n_de_float, rr_c.n_de_chunk)
    assert np.abs(ts_g - ts_c).max() <= 1e-9 * ts_c[-1]
    worst = 0.0
    for k in range(1, n_steps + 1):
        scale = np.abs(xs_c[k]).max()
        for i in range(0, xs_c.shape[1], 3):
            den = max(np.abs(xs_c[k, i:i + 3]).max(), 1e-6 * scale)
            worst = max(worst, np.abs(xs_g[k, i:i + 3] - xs_c[k, i:i + 3]).max() / den)
    assert worst <= TOL, worst
    if start == "settled":
        out = S.force_all_elastic_intersections(m_g, xs_g[-1])
        assert (out["flags"] & 1).sum() >= 3     # still a stack in contact at the end


def test_batched_radau_follows_the_single_scene_integrator():
    """BatchedRadau (all environments advanced together on the GPU: batched calcXd!, batched Dual-6 Jacobian chunks, batched complex
    inverses, per-environment Newton / step-size / order control) against the single-scene mirror of the reference's integrator run
    environment by environment on the same device entry points: same accepted step sizes, states within 1e-7 relative after every
    step (the two differ only in the linear-algebra library and in rounding)."""
    from pfc_b200.radau_batched import BatchedRadau
    n_env, n_steps = 6, 25
    m = scene_boxes(_ctx(), max_env=4 * n_env)[0]
    x0 = boxes_env_states(m, n_env)
    x0[0] = S.get_state(m)                      # environment 0: the reference's own drop
    x0[1, m.nq:] = 0.0                          # environment 1: a settled stack at rest
    br = BatchedRadau(m, n_env, h_max=0.05)
    ts_b, xs_b = br.integrate(x0, n_steps)
    assert br.n_chunk_states == n_steps * 8 * n_env
    for e in range(n_env):
        dyn = D.FloatingBodyDynamics(m, device=True)
        rr = R.makeRadauIntegrator(dyn, S.num_x(m), 1.0e-16, 2, 6)
        rr.step.h_max = 0.05
        ts, xs = R.integrate_radau(rr, x0[e], t_final=1e9, max_steps=n_steps, after_step=lambda x: D.principal_value(m, x))
        assert np.abs(ts_b[:, e] - ts).max() <= 1e-7 * ts[-1], e
        for k in range(1, n_steps + 1):
            scale = np.abs(xs[k]).max()
            for i in range(0, xs.shape[1], 3):
                den = max(np.abs(xs[k, i:i + 3]).max(), 1e-6 * scale)
                assert np.abs(xs_b[k, e, i:i + 3] - xs[k, i:i + 3]).max() <= 1e-7 * den, (e, k, i)
