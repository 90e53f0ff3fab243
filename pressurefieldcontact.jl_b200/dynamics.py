"""dynamics.py -- host-side mirror of calcXd! for scenes whose bodies float on the world
(reference: src/contact_algorithms_non_friction.jl:18-52, with the
RigidBodyDynamics 1.4.0 pieces it calls -- mass_matrix!, dynamics_bias!, configuration_derivative!,
cholesky/ldiv -- restated for SPQuatFloating joints attached to the world, which is what test/boxes.jl
and the MPC batch use).  This is CALLER-side code: the contact wrenches come from the backend
(CUDA library or CPU oracle) through the same boundary as everywhere else; nothing here is a fallback
for the device path.  The device version of the same function is pfc_calcxd_f64 (csrc/pfc_state.cu).

  make_inertia_info ........ src/body_inertia.jl:22-66 (makeInertiaTensor / centroidVolumeCombo); the
                             reference integrates r r' with a degree >= 3 quadrature rule, which is exact for
                             this quadratic, so the closed-form simplex moments used here agree to rounding
  spatial_inertia .......... newBodyFromInertia, src/body_inertia.jl:2-9
  FloatingBodyDynamics ..... calcXd! :18-38, sum_all_forces! :40-52, copyto! src/extensions.jl:21-50;
                             .de / .de_jacobian_chunk are what RadauIntegrator calls (radau_functions.jl:9,67)
  principal_value .......... principal_value!, src/extensions.jl:2-7 (MRP -> shadow set when |p| > 1)

For one floating body with spatial inertia H (6x6, body frame, [angular; linear]), twist v = [w; u]
in the body frame, pose (R, t):
    H (vdot - [0; R' g]) + v x* (H v) = f        q_dot = [ B(p) w ; R u ],   B(p) = 1/4 ((1 - p'p) I + 2 [p]x + 2 p p')
"""
from __future__ import annotations

import numpy as np

from . import scenario as S

__all__ = ["make_inertia_info", "spatial_inertia", "FloatingBodyDynamics", "principal_value", "mrp_rate_matrix"]


def make_inertia_info(e_mesh, i_prop):
    """-> (tensor_I about the centre of mass, com, mass, volume)."""
    pts = np.asarray(e_mesh.point, dtype=np.float64)
    if e_mesh.is_tet:
        idx = np.asarray(e_mesh.tet)
        v = pts[idx]                                                     # [n][4][3]
        vol = np.abs(np.einsum("ni,ni->n", np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]), v[:, 3] - v[:, 0])) / 6.0
        denom = 20.0
    else:
        if i_prop.d is None:
            raise ValueError("a triangle mesh needs InertiaProperties.d (shell thickness)")
        idx = np.asarray(e_mesh.tri)
        v = pts[idx]                                                     # [n][3][3]
        vol = 0.5 * np.linalg.norm(np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]), axis=1) * i_prop.d
        denom = 12.0
    cen = v.mean(axis=1)
    volume = float(vol.sum())
    com = (vol[:, None] * cen).sum(axis=0) / volume
    r = v - com                                                          # vertices relative to the centre of mass
    ssum = r.sum(axis=1)
    # int r r' over a simplex = vol / denom * (sum_i r_i r_i' + (sum_i r_i)(sum_i r_i)')
    second = np.einsum("n,nij->ij", vol / denom, np.einsum("nki,nkj->nij", r, r) + np.einsum("ni,nj->nij", ssum, ssum))
    tensor_I = i_prop.rho * (np.eye(3) * np.trace(second) - second)
    return tensor_I, com, volume * i_prop.rho, volume


def _skew(c):
    return np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])


def spatial_inertia(tensor_I, com, mass):
    """6x6 spatial inertia about the body origin, [angular; linear] ordering."""
    cx = _skew(com)
    J = tensor_I + mass * cx @ cx.T
    H = np.zeros((6, 6))
    H[:3, :3], H[:3, 3:], H[3:, :3], H[3:, 3:] = J, mass * cx, mass * cx.T, mass * np.eye(3)
    return H


def mrp_rate_matrix(p):
    """p_dot = B(p) w for a body-frame angular velocity w (complex-safe)."""
    p = np.asarray(p)
    px = np.array([[0 * p[0], -p[2], p[1]], [p[2], 0 * p[0], -p[0]], [-p[1], p[0], 0 * p[0]]])
    return 0.25 * ((1 - p @ p) * np.eye(3) + 2 * px + 2 * np.outer(p, p))


def principal_value(m, x) -> None:
    """In place: every SPQuatFloating joint's MRP is moved to the shadow set when |p|^2 > 1."""
    for b in m.bodies:
        if isinstance(b.joint, S.SPQuatFloating):
            p = x[b.q0:b.q0 + 3]
            n2 = float(p @ p)
            if n2 > 1.0:
                x[b.q0:b.q0 + 3] = -p / n2


class FloatingBodyDynamics:
    """The ODE right-hand side x -> [q_dot; v_dot; s_dot] of a MechanismScenario whose bodies all float on the
    world (the reference's calcXd!), with the contact wrenches evaluated by m.backend."""

    def __init__(self, m, device: bool = False):
        """device=True: x_dot and its Jacobian chunks come from pfc_calcxd_f64 / pfc_calcxd_dual6 (everything on the GPU);
        device=False: rigid-body terms on the host (this class), contact wrenches from m.backend at the boundary level."""
        if m.backend is None:
            raise RuntimeError("finalize(m, backend=...) first")
        if device and not getattr(m, "device_dynamics", False):
            raise RuntimeError("device=True needs a CUDA backend with pfc_set_dynamics (floating bodies with InertiaProperties)")
        self.m = m
        self.device = device
        self.bodies = [b for b in m.bodies if b.joint is not None]
        for b in self.bodies:
            if not (isinstance(b.joint, S.SPQuatFloating) and b.parent == 0 and np.array_equal(b.pose_R, np.eye(3)) and not b.pose_t.any()):
                raise NotImplementedError("FloatingBodyDynamics handles SPQuatFloating joints attached to the world at the identity pose")
        self.H, self.Hinv = {}, {}
        for k, b in enumerate(m.bodies):
            if b.joint is None:
                continue
            meshes = [mc for mc in m.MeshCache if mc.body_id == k]
            props = getattr(b, "i_prop", None)
            if props is None or len(meshes) != 1:
                raise ValueError(f"body {b.name} needs exactly one mesh with InertiaProperties (add_body_contact(..., i_prop=...))")
            H = spatial_inertia(*make_inertia_info(meshes[0].mesh, props)[:3])
            self.H[k], self.Hinv[k] = H, np.linalg.inv(H)
        self.nq, self.nv, self.nx = m.nq, m.nv, S.num_x(m)
        self.n_float = 0
        self.n_chunk = 0

    # ---- pieces ---------------------------------------------------------------------------------------
    def _assemble(self, x, wrench_r2, sdot):
        """[q_dot; v_dot; s_dot] from a state and the per-instruction wrenches (real or complex arrays)."""
        m, nq, nv = self.m, self.nq, self.nv
        dt = np.result_type(x.dtype, wrench_r2.dtype)
        xx = np.zeros(self.nx, dtype=dt)
        pose = {}
        for k, b in enumerate(m.bodies):
            if b.joint is None:
                pose[k] = (np.eye(3), np.zeros(3))
            else:
                pose[k] = (S.mrp_to_rotation(x[b.q0:b.q0 + 3]), x[b.q0 + 3:b.q0 + 6])
        f = {k: np.zeros(6, dtype=dt) for k in self.H}
        for i, ci in enumerate(m.ContactInstructions):     # addGeneralizedForcesThirdLaw!: +J2' w on body 2, -J1' w on body 1
            b1, b2 = m.MeshCache[ci.id_1].body_id, m.MeshCache[ci.id_2].body_id
            R2, t2 = pose[b2]
            lin_w = R2 @ wrench_r2[i, 3:]
            ang_w = R2 @ wrench_r2[i, :3] + np.cross(t2, lin_w)
            for bid, sign in ((b2, 1.0), (b1, -1.0)):
                if bid in f:
                    R, t = pose[bid]
                    f[bid][:3] += sign * (R.T @ (ang_w - np.cross(t, lin_w)))
                    f[bid][3:] += sign * (R.T @ lin_w)
        for k, b in enumerate(m.bodies):
            if b.joint is None:
                continue
            R, _ = pose[k]
            v = x[nq + b.v0:nq + b.v0 + 6]
            w, u = v[:3], v[3:]
            h = self.H[k] @ v
            bias = np.concatenate([np.cross(w, h[:3]) + np.cross(u, h[3:]), np.cross(w, h[3:])])     # v x* (H v)
            vdot = self.Hinv[k] @ (f[k] - bias)
            vdot[3:] = vdot[3:] + R.T @ m.gravity
            xx[b.q0:b.q0 + 3] = mrp_rate_matrix(x[b.q0:b.q0 + 3]) @ w
            xx[b.q0 + 3:b.q0 + 6] = R @ u
            xx[nq + b.v0:nq + b.v0 + 6] = vdot
        if m.n_bristle:
            xx[nq + nv:] = sdot
        return xx

    def calcXd(self, x, t: float = 0.0):
        x = np.asarray(x, dtype=np.float64)
        if self.device:
            self.n_float += 1
            return self.m.backend.calcxd_f64(x)["xdot"][0]
        out = S.force_all_elastic_intersections(self.m, x)
        self.n_float += 1
        return self._assemble(x, out["wrench"], out["sdot"])

    # ---- the two entry points the integrator uses ---------------------------------------------------------
    def de(self, xx, x, t: float = 0.0) -> None:
        xx[:] = self.calcXd(x, t)

    def de_jacobian_chunk(self, x, i0: int, i1: int, t: float = 0.0):
        """One Dual-6 evaluation seeded on x[i0:i1] (i1 - i0 <= 6): returns (xx(x), d xx / d x[i0:i1]).
        Kinematics and rigid-body terms are differentiated by complex step (analytic functions, exact to rounding);
        the contact wrenches and their 6 partials come from the backend's Dual-6 evaluation (pfc_eval_dual6)."""
        if i1 - i0 > 6:
            raise ValueError("the device evaluates Jacobian chunks of 6 seeds (Dual{Nothing,Float64,6})")
        m = self.m
        x = np.asarray(x, dtype=np.float64)
        if self.device:
            self.n_chunk += 1
            xd7 = m.backend.calcxd_dual6(x, i0)["xdot7"][0]
            return xd7[:, 0].copy(), xd7[:, 1:1 + (i1 - i0)].copy()
        X0, X7, tw7, s7 = S.boundary_arrays_dual6(m, x, i0)
        out = m.backend.eval_dual6(X0, X7, tw7, s7 if m.n_bristle else None)
        self.n_chunk += 1
        w7 = out["wrench"][0]                                       # [ins][6][7]
        sd7 = out["sdot"][0].reshape(-1, 7) if m.n_bristle else np.zeros((0, 7))
        xx0 = self._assemble(x, w7[..., 0], sd7[:, 0])
        cols = np.zeros((self.nx, i1 - i0))
        h = 1e-30
        for d, j in enumerate(range(i0, i1)):
            xc = x.astype(np.complex128)
            xc[j] += 1j * h
            cols[:, d] = self._assemble(xc, w7[..., 0] + 1j * h * w7[..., 1 + d], sd7[:, 0] + 1j * h * sd7[:, 1 + d]).imag / h
        return xx0, cols
