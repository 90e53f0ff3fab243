"""Host-side mirror of the reference's scenario API for the contact-wrench path.

Mirrors (names, argument meaning, defaults and error behaviour) -- paths relative to
/root/reference:
  * Regularized / Bristle friction models ............ src/mechanism_scenario.jl:5-34
  * ContactInstructions + the Tri/Tet ordering rule ... src/mechanism_scenario.jl:36-49, 399-416
  * ContactProperties / InertiaProperties ............. src/structs.jl:9-31
  * MeshCache ......................................... src/structs.jl:33-54
  * MechanismScenario, add_contact!, add_body_contact!,
    add_friction_regularize!, add_friction_bristle!,
    finalize!, set_state_spq!, get_state, num_x ........ src/mechanism_scenario.jl:166-397
  * state layout x = [q; v; s], xdot = [qdot; vdot; sdot] src/extensions.jl:21-50
  * refreshBodyBodyTransform! / refreshBodyBodyCache! .. src/contact_algorithms_non_friction.jl:103-134
  * addGeneralizedForcesThirdLaw! ....................... src/contact_algorithms_non_friction.jl:267-286
  * forceAllElasticIntersections! ....................... src/contact_algorithms_non_friction.jl:60-68

The reference delegates rigid-body kinematics to RigidBodyDynamics 1.4.0 (not vendored).  The
subset needed at the boundary -- world-attached bodies, SPQuatFloating joints (q = [MRP; trans],
v = [omega; vel] in the body frame), Prismatic and Revolute joints, in chains -- is restated here in
numpy; it runs on the host exactly where the reference runs it.
The device work is done by a *backend* (pfc_b200.capi.Context in production).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from .geometry import FlatTree, eMesh, eMesh_to_tree

__all__ = [
    "Regularized", "Bristle", "ContactProperties", "InertiaProperties", "ContactInstructions", "MeshCache", "Prismatic", "Revolute", "add_body",
    "SPQuatFloating", "MechanismScenario", "add_contact", "add_body_contact", "add_friction_regularize", "add_friction_bristle",
    "finalize", "set_state_spq", "set_configuration", "get_state", "num_x", "boundary_arrays", "boundary_arrays_dual6",
    "force_all_elastic_intersections", "force_all_elastic_intersections_batch", "mrp_to_rotation", "rotation_to_mrp",
]


def default_chi() -> float:
    return 0.5


def default_mu() -> float:
    return 0.3


def determine_mu_s_mu_d(mu_s, mu_d):
    """src/mechanism_scenario.jl:350-356 (including the quirk that both-None yields default_chi)."""
    if mu_s is None and mu_d is None:
        return default_chi(), default_chi()
    if mu_d is None:
        raise ValueError("need to specify μd")
    if mu_s is None:
        return mu_d, mu_d
    if not (mu_d <= mu_s):
        raise ValueError("something is wrong")
    return mu_s, mu_d


@dataclass
class Regularized:
    mu_s: float
    mu_d: float
    v_c: float

    def __post_init__(self):
        self.mu_s, self.mu_d = determine_mu_s_mu_d(self.mu_s, self.mu_d)
        self.v_mu_s = 2 * self.v_c
        self.v_mu_d = 3 * self.v_c

    model = 0

    def params(self):
        return [self.mu_s, self.mu_d, self.v_c]


@dataclass
class Bristle:
    bristle_id: int
    tau: float
    k_bar: float
    mu_s: float
    mu_d: float
    magic: float

    def __post_init__(self):
        self.mu_s, self.mu_d = determine_mu_s_mu_d(self.mu_s, self.mu_d)
        self.Ts_mu_s = 2 * self.mu_s
        self.Ts_mu_d = 3 * self.mu_s

    model = 1

    def params(self):
        return [self.tau, self.k_bar, self.mu_s, self.mu_d, self.magic]


@dataclass
class ContactProperties:
    E_bar: float

    def __post_init__(self):
        if not (1.0e4 <= self.E_bar <= 3.0e11):
            raise ValueError("E_effective in unexpected range.")


@dataclass
class InertiaProperties:
    rho: float
    d: Optional[float] = None

    def __post_init__(self):
        if self.d is not None and not (0.001 <= self.d <= 0.1):
            raise ValueError("thickness in unexpected range.")
        if not (50.0 <= self.rho):
            raise ValueError("rho in unexpected range.")


@dataclass
class ContactInstructions:
    id_1: int
    id_2: int
    chi: float
    friction_model: object
    n_quad_rule: int

    def __post_init__(self):
        if not (1 <= self.n_quad_rule <= 2):
            raise ValueError("only quadrature rules 1 (first order) and 2 (second? order) are currently implemented")


@dataclass
class SPQuatFloating:
    nq: int = 6
    nv: int = 6


@dataclass
class Prismatic:
    axis: Sequence[float] = (0.0, 0.0, 1.0)
    nq: int = 1
    nv: int = 1


@dataclass
class Revolute:
    axis: Sequence[float] = (0.0, 0.0, 1.0)
    nq: int = 1
    nv: int = 1


@dataclass
class Body:
    name: str
    joint: object            # None => the world (root) body
    q0: int = 0              # offset of this joint's q in the q vector
    v0: int = 0
    pose_R: np.ndarray = field(default_factory=lambda: np.eye(3))   # joint_pose on the parent body
    pose_t: np.ndarray = field(default_factory=lambda: np.zeros(3))
    parent: int = 0          # index of the parent body (0 = world); bodies are stored parents first
    i_prop: Optional["InertiaProperties"] = None


@dataclass
class MeshCache:
    name: str
    body_id: int
    mesh: eMesh
    tree: FlatTree
    c_prop: Optional[ContactProperties]

    @property
    def is_tet(self) -> bool:
        return self.mesh.is_tet


class MechanismScenario:
    """Contact information for an entire mechanism (src/mechanism_scenario.jl:166-199).  ``de`` is
    the reference's plug-in point; here it defaults to None and the drop-in function is
    force_all_elastic_intersections (the narrower seam of SURVEY.md section 8b)."""

    def __init__(self, de=None, N_chunk: int = 6, gravity=(0.0, 0.0, -9.8054)):
        self.de = de
        self.N_chunk = N_chunk
        self.gravity = np.asarray(gravity, dtype=np.float64)
        self.bodies: List[Body] = [Body("world", None)]
        self.MeshCache: List[MeshCache] = []
        self.ContactInstructions: List[ContactInstructions] = []
        self.n_bristle = 0
        self.q = np.zeros(0)
        self.v = np.zeros(0)
        self.s = np.zeros(0)
        self.backend = None
        self.finalized = False
        self.TT_Cache_n_pairs = None
        self.device_kinematics = False
        self.device_dynamics = False

    # state sizes
    @property
    def nq(self) -> int:
        return sum(b.joint.nq for b in self.bodies if b.joint is not None)

    @property
    def nv(self) -> int:
        return sum(b.joint.nv for b in self.bodies if b.joint is not None)


# --------------------------------------------------------------------------------------------------
# builders
# --------------------------------------------------------------------------------------------------
def find_mesh_id(m: MechanismScenario, name: str) -> int:
    ids = [k for k, mc in enumerate(m.MeshCache) if mc.name == name]
    if len(ids) > 1:
        raise ValueError("multiple")
    if not ids:
        raise ValueError(f"no mesh found by name: {name}")
    return ids[0]


def add_contact(m: MechanismScenario, name: str, e_mesh: eMesh, c_prop: Optional[ContactProperties] = None, body: Optional[int] = None,
                tree: Optional[FlatTree] = None) -> int:
    """add_contact! (src/mechanism_scenario.jl:298-314).  Returns the mesh id."""
    if e_mesh.is_tri and e_mesh.is_tet:
        raise ValueError("eMesh has triangles and tets. Use as_tri_eMesh or as_tet_eMesh to convert eMesh.")
    if e_mesh.is_tri and c_prop is not None:
        raise ValueError("Using ContactProperties for triangular eMesh")
    if e_mesh.is_tet and c_prop is None:
        raise ValueError("Using nothing as ContactProperties for tet eMesh")
    body = 0 if body is None else body
    if tree is None:
        tree = eMesh_to_tree(e_mesh)
    m.MeshCache.append(MeshCache(name, body, e_mesh, tree, c_prop))
    return len(m.MeshCache) - 1


def add_body(m: MechanismScenario, name: str, joint=None, pose_R=None, pose_t=None, body: Optional[int] = None) -> int:
    """add_body! restricted to what the contact path needs: the joint type, its parent body
    (``body``, default the world) and its pose on the parent (src/mechanism_scenario.jl:324-345;
    inertia stays with the caller's dynamics)."""
    joint = SPQuatFloating() if joint is None else joint
    b = Body(name, joint, q0=m.nq, v0=m.nv, parent=0 if body is None else body)
    if pose_R is not None:
        b.pose_R = np.asarray(pose_R, dtype=np.float64)
    if pose_t is not None:
        b.pose_t = np.asarray(pose_t, dtype=np.float64)
    m.bodies.append(b)
    return len(m.bodies) - 1


def add_body_contact(m: MechanismScenario, name: str, e_mesh: eMesh, i_prop: Optional[InertiaProperties] = None,
                     c_prop: Optional[ContactProperties] = None, joint=None, tree: Optional[FlatTree] = None, body: Optional[int] = None):
    """add_body_contact! (src/mechanism_scenario.jl:279-289); ``body`` is the parent.  Returns (body, joint, id)."""
    new_body = add_body(m, name, joint, body=body)
    m.bodies[new_body].i_prop = i_prop      # used by the caller-side dynamics (dynamics.py), not by the contact path
    mesh_id = add_contact(m, name, e_mesh, c_prop=c_prop, body=new_body, tree=tree)
    return new_body, m.bodies[new_body].joint, mesh_id


def _add_friction(m: MechanismScenario, id_1: int, id_2: int, fric_model, chi: float, n_quad_rule: int) -> ContactInstructions:
    """add_friction! ordering rule (src/mechanism_scenario.jl:399-416): id_2 is always a Tet mesh;
    (Tet, Tri) is swapped; (Tri, Tri) has no method."""
    m_1, m_2 = m.MeshCache[id_1], m.MeshCache[id_2]
    if m_1.is_tet and not m_2.is_tet:
        return _add_friction(m, id_2, id_1, fric_model, chi, n_quad_rule)
    if not m_2.is_tet:
        raise TypeError("MethodError: no method matching add_friction!(::MeshCache{Tri,Nothing}, ::MeshCache{Tri,Nothing})")
    c_ins = ContactInstructions(id_1, id_2, chi, fric_model, n_quad_rule)
    m.ContactInstructions.append(c_ins)
    return c_ins


def add_friction_regularize(m: MechanismScenario, mesh_id_1: int, mesh_id_2: int, mu_s=None, mu_d=None, chi: float = None,
                            v_tol: float = 0.01, n_quad_rule: int = 2) -> ContactInstructions:
    """add_friction_regularize! (src/mechanism_scenario.jl:365-375)"""
    chi = default_chi() if chi is None else chi
    mu_s, mu_d = determine_mu_s_mu_d(mu_s, mu_d)
    return _add_friction(m, mesh_id_1, mesh_id_2, Regularized(mu_s, mu_d, v_tol), chi, n_quad_rule)


def add_friction_bristle(m: MechanismScenario, mesh_id_1: int, mesh_id_c: int, tau: float = 0.05, k_bar: float = 1.0e4, mu_s=None, mu_d=None,
                         chi: float = None, n_quad_rule: int = 2, magic: float = 1.0e-3) -> ContactInstructions:
    """add_friction_bristle! (src/mechanism_scenario.jl:384-397)"""
    chi = default_chi() if chi is None else chi
    mu_s, mu_d = determine_mu_s_mu_d(mu_s, mu_d)
    if not (0 < mu_d):
        raise ValueError("μd cannot be 0 for bristle friction")
    bf = Bristle(m.n_bristle, tau, k_bar, mu_s, mu_d, magic)
    m.n_bristle += 1
    return _add_friction(m, mesh_id_1, mesh_id_c, bf, chi, n_quad_rule)


def finalize(m: MechanismScenario, backend=None, max_env: int = 1) -> None:
    """finalize! (src/mechanism_scenario.jl:206-231) + the one-time upload of the static scene to
    the backend (the finalize_gpu! of SURVEY.md section 8b)."""
    m.q = np.zeros(m.nq)
    m.v = np.zeros(m.nv)
    m.s = np.zeros(6 * m.n_bristle)
    m.finalized = True
    if backend is not None:
        attach_backend(m, backend, max_env)


def attach_backend(m: MechanismScenario, backend, max_env: int = 1) -> None:
    for mc in m.MeshCache:
        em = mc.mesh
        if em.is_tet:
            backend.add_mesh(1, em.point, em.tet, em.eps, mc.c_prop.E_bar, mc.tree)
        else:
            backend.add_mesh(0, em.point, em.tri, None, 0.0, mc.tree)
    for ci in m.ContactInstructions:
        fm = ci.friction_model
        backend.add_instruction(ci.id_1, ci.id_2, ci.chi, fm.model, fm.params(), ci.n_quad_rule)
    backend.finalize(max_env)
    m.backend = backend
    # device-side kinematics (pfc_set_bodies) when every body is world-attached or floats on the world
    if hasattr(backend, "set_bodies") and all(b.joint is None or (isinstance(b.joint, SPQuatFloating) and b.parent == 0) for b in m.bodies):
        pose = np.array([np.concatenate([b.pose_R.reshape(9), b.pose_t]) for b in m.bodies])
        backend.set_bodies([0 if b.joint is None else 1 for b in m.bodies], [b.q0 for b in m.bodies], [b.v0 for b in m.bodies],
                           [mc.body_id for mc in m.MeshCache], m.nq, m.nv, pose=pose)
        m.device_kinematics = True
        # device-side calcXd! (pfc_set_dynamics) when every floating body carries one mesh with InertiaProperties
        floating = [k for k, b in enumerate(m.bodies) if b.joint is not None]
        if hasattr(backend, "set_dynamics") and floating and all(
                m.bodies[k].i_prop is not None and sum(mc.body_id == k for mc in m.MeshCache) == 1 for k in floating):
            from . import dynamics as _dyn
            H = np.zeros((len(m.bodies), 6, 6))
            for k in floating:
                mesh = next(mc.mesh for mc in m.MeshCache if mc.body_id == k)
                H[k] = _dyn.spatial_inertia(*_dyn.make_inertia_info(mesh, m.bodies[k].i_prop)[:3])
            backend.set_dynamics(H, m.gravity)
            m.device_dynamics = True


def refit_mesh(m: MechanismScenario, mesh_id: int, new_points) -> None:
    """Moves the vertices of one mesh (same connectivity): the host description gets the refitted tree (geometry.refit_tree) and, with a
    CUDA backend attached, the device refits its primitive records and boxes itself (pfc_refit_mesh) -- no re-upload of the scene."""
    from .geometry import refit_tree
    mc = m.MeshCache[mesh_id]
    new_points = np.ascontiguousarray(new_points, dtype=np.float64).reshape(mc.mesh.point.shape)
    mc.mesh = eMesh(point=new_points, tri=mc.mesh.tri, tet=mc.mesh.tet, eps=mc.mesh.eps)
    mc.tree = refit_tree(mc.tree, mc.mesh)
    if m.backend is not None:
        if not hasattr(m.backend, "refit_mesh"):
            raise RuntimeError("this backend has no refit: build a new scene from the moved mesh and its refitted tree")
        m.backend.refit_mesh(mesh_id, new_points)


# --------------------------------------------------------------------------------------------------
# state
# --------------------------------------------------------------------------------------------------
def num_x(m: MechanismScenario) -> int:
    return m.nq + m.nv + 6 * m.n_bristle


def get_state(m: MechanismScenario) -> np.ndarray:
    return np.concatenate([m.q, m.v, m.s])


def rotation_to_mrp(R: np.ndarray) -> np.ndarray:
    """MRP(rot): stereographic projection of the unit quaternion with w >= 0."""
    R = np.asarray(R, dtype=np.float64)
    w = 0.5 * np.sqrt(max(0.0, 1.0 + R[0, 0] + R[1, 1] + R[2, 2]))
    if w > 1e-6:
        xyz = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (4 * w)
    else:  # rotation by ~pi
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        xi = 0.5 * np.sqrt(max(0.0, 1 + R[i, i] - R[j, j] - R[k, k]))
        xyz = np.zeros(3)
        xyz[i] = xi
        xyz[j] = (R[j, i] + R[i, j]) / (4 * xi)
        xyz[k] = (R[k, i] + R[i, k]) / (4 * xi)
        w = (R[k, j] - R[j, k]) / (4 * xi)
    return xyz / (1.0 + w)


def mrp_to_rotation(p):
    """SPQuat/MRP -> rotation matrix (Rotations.jl, not vendored): q = ((1-a2)/(1+a2), 2p/(1+a2)).
    Works for real and complex input (complex-step differentiation is used for Dual seeds)."""
    x, y, z = p[0], p[1], p[2]
    a2 = x * x + y * y + z * z
    w = (1 - a2) / (a2 + 1)
    x, y, z = 2 * x / (a2 + 1), 2 * y / (a2 + 1), 2 * z / (a2 + 1)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def set_state_spq(m: MechanismScenario, body: int, rot=None, trans=(0.0, 0.0, 0.0), w=(0.0, 0.0, 0.0), vel=(0.0, 0.0, 0.0)) -> None:
    """set_state_spq! (src/mechanism_scenario.jl:247-256).  ``body`` is the index returned by
    add_body_contact (the reference passes the joint object)."""
    b = m.bodies[body]
    if not isinstance(b.joint, SPQuatFloating):
        raise TypeError("set_state_spq needs an SPQuatFloating joint")
    mrp = np.zeros(3) if rot is None else rotation_to_mrp(rot)
    m.q[b.q0:b.q0 + 6] = np.concatenate([mrp, np.asarray(trans, dtype=np.float64)])
    m.v[b.v0:b.v0 + 6] = np.concatenate([np.asarray(w, dtype=np.float64), np.asarray(vel, dtype=np.float64)])


def set_configuration(m: MechanismScenario, body: int, config) -> None:
    b = m.bodies[body]
    m.q[b.q0:b.q0 + b.joint.nq] = np.asarray(config, dtype=np.float64)


# --------------------------------------------------------------------------------------------------
# kinematics at the boundary (RigidBodyDynamics restated for the supported joints)
# --------------------------------------------------------------------------------------------------
def _joint_axis(b: Body):
    return np.asarray(b.joint.axis, dtype=np.float64)


def _body_kinematics(m: MechanismScenario, q, v):
    """Per body: (R, t) = transform_to_root, (ang, lin) = twist_wrt_world expressed in world (about
    the world origin).  Bodies are stored parents first, so one forward sweep covers every chain."""
    dt = np.result_type(q.dtype, v.dtype)
    out = []
    for b in m.bodies:
        if b.joint is None:
            out.append((np.eye(3, dtype=dt), np.zeros(3, dtype=dt), np.zeros(3, dtype=dt), np.zeros(3, dtype=dt)))
            continue
        if isinstance(b.joint, SPQuatFloating):
            Rj = mrp_to_rotation(q[b.q0:b.q0 + 3])
            tj = q[b.q0 + 3:b.q0 + 6]
            om_b, vel_b = v[b.v0:b.v0 + 3], v[b.v0 + 3:b.v0 + 6]
        elif isinstance(b.joint, Revolute):
            ax = _joint_axis(b)
            ax = ax / np.linalg.norm(ax)
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            th = q[b.q0]
            Rj = np.eye(3, dtype=dt) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
            tj = np.zeros(3, dtype=dt)
            om_b, vel_b = ax * v[b.v0], np.zeros(3, dtype=dt)
        else:  # Prismatic
            ax = _joint_axis(b)
            Rj = np.eye(3, dtype=dt)
            tj = ax * q[b.q0]
            om_b, vel_b = np.zeros(3, dtype=dt), ax * v[b.v0]
        Rp, tp, ap, lp = out[b.parent]
        Rpose = Rp @ b.pose_R
        R = Rpose @ Rj
        t = Rpose @ tj + Rp @ b.pose_t + tp
        ang_j = R @ om_b                                   # joint twist, body frame -> world, about the world origin
        lin_j = R @ vel_b + np.cross(t, ang_j)
        out.append((R, t, ap + ang_j, lp + lin_j))
    return out


def _motion_subspace_world(m: MechanismScenario, kin, bid: int):
    """Columns of the geometric Jacobian contributed by body bid's own joint: a list of
    (velocity index, angular part, linear part about the world origin), in world."""
    b = m.bodies[bid]
    R, t, _, _ = kin[bid]
    cols = []
    if isinstance(b.joint, SPQuatFloating):
        for k in range(3):
            cols.append((b.v0 + k, R[:, k], np.cross(t, R[:, k])))
        for k in range(3):
            cols.append((b.v0 + 3 + k, np.zeros(3), R[:, k]))
    elif isinstance(b.joint, Revolute):
        ax = _joint_axis(b)
        a_w = R @ (ax / np.linalg.norm(ax))
        cols.append((b.v0, a_w, np.cross(t, a_w)))
    else:
        cols.append((b.v0, np.zeros(3), R @ _joint_axis(b)))
    return cols


def _ins_boundary(kin, id_body_1, id_body_2):
    R1, t1, a1, l1 = kin[id_body_1]
    R2, t2, a2, l2 = kin[id_body_2]
    # x_r2_rw = inv(x_rw_r2);  x_r2_r1 = x_r2_rw * x_rw_r1   (non_friction.jl:109-113)
    R2t = R2.T
    t_inv = -(R2t @ t2)
    R21 = R2t @ R1
    t21 = R2t @ t1 + t_inv
    # twist_r2_r1 = -twist_w_r1 + twist_w_r2 in world, then transform(., x_r2_rw)  (:125-128)
    ang_w = -a1 + a2
    lin_w = -l1 + l2
    ang = R2t @ ang_w
    lin = R2t @ lin_w + np.cross(t_inv, ang)
    X = np.zeros((4, 4), dtype=R21.dtype)
    X[:3, :3] = R21
    X[:3, 3] = t21
    X[3, 3] = 1.0
    return X, np.concatenate([ang, lin])


def boundary_arrays(m: MechanismScenario, x: np.ndarray):
    """State vector(s) -> the C-ABI boundary arrays: X_r2_r1 [env][ins][16] (column-major 4x4) and
    twist_r2 [env][ins][6] (angular, linear); s [env][6 n_bristle]."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    n_env, n_ins = x.shape[0], len(m.ContactInstructions)
    X = np.zeros((n_env, n_ins, 16))
    tw = np.zeros((n_env, n_ins, 6))
    nq, nv = m.nq, m.nv
    for e in range(n_env):
        kin = _body_kinematics(m, x[e, :nq], x[e, nq:nq + nv])
        for k, ci in enumerate(m.ContactInstructions):
            Xk, twk = _ins_boundary(kin, m.MeshCache[ci.id_1].body_id, m.MeshCache[ci.id_2].body_id)
            X[e, k] = Xk.T.reshape(16)
            tw[e, k] = twk
    return X, tw, np.ascontiguousarray(x[:, nq + nv:])


def boundary_arrays_dual6(m: MechanismScenario, x: np.ndarray, seed_start: int):
    """Boundary arrays in Dual-6 form for the Jacobian chunk that seeds x[seed_start : seed_start+6]
    (src/radau/radau_functions.jl:16-26).  Partials are obtained by complex-step differentiation
    of the (analytic) kinematics, exact to rounding."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    X0, tw0, s0 = boundary_arrays(m, x)
    n_ins = len(m.ContactInstructions)
    X7 = np.zeros((1, n_ins, 16, 7))
    tw7 = np.zeros((1, n_ins, 6, 7))
    s7 = np.zeros((1, 6 * m.n_bristle, 7))
    X7[..., 0], tw7[..., 0], s7[..., 0] = X0, tw0, s0
    nq, nv = m.nq, m.nv
    h = 1e-30
    for d in range(6):
        j = seed_start + d
        if j >= len(x):
            break
        if j >= nq + nv:
            s7[0, j - nq - nv, 1 + d] = 1.0
            continue
        xc = x.astype(np.complex128)
        xc[j] += 1j * h
        kin = _body_kinematics(m, xc[:nq], xc[nq:nq + nv])
        for k, ci in enumerate(m.ContactInstructions):
            Xk, twk = _ins_boundary(kin, m.MeshCache[ci.id_1].body_id, m.MeshCache[ci.id_2].body_id)
            X7[0, k, :, 1 + d] = Xk.T.reshape(16).imag / h
            tw7[0, k, :, 1 + d] = twk.imag / h
    return X0, X7, tw7, s7.reshape(1, m.n_bristle, 6, 7) if m.n_bristle else s7


def generalized_forces(m: MechanismScenario, x: np.ndarray, wrench_r2: np.ndarray) -> np.ndarray:
    """addGeneralizedForcesThirdLaw! (src/contact_algorithms_non_friction.jl:267-286): wrench (about
    the r2 origin, in r2, applied to body 2) -> world, then +J2' w - J1' w."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    nq, nv = m.nq, m.nv
    kin = _body_kinematics(m, x[:nq], x[nq:nq + nv])
    f = np.zeros(nv)
    for k, ci in enumerate(m.ContactInstructions):
        b1, b2 = m.MeshCache[ci.id_1].body_id, m.MeshCache[ci.id_2].body_id
        R2, t2, _, _ = kin[b2]
        ang, lin = wrench_r2[k, :3], wrench_r2[k, 3:]
        lin_w = R2 @ lin
        ang_w = R2 @ ang + np.cross(t2, lin_w)
        for bid, sign in ((b2, +1.0), (b1, -1.0)):
            while bid != 0:  # world-attached meshes have jac == nothing; every joint on the path to the root gets J' w
                for iv, s_ang, s_lin in _motion_subspace_world(m, kin, bid):
                    f[iv] += sign * float(s_ang @ ang_w + s_lin @ lin_w)
                bid = m.bodies[bid].parent
    return f


def force_all_elastic_intersections_batch(m: MechanismScenario, x: np.ndarray):
    """forceAllElasticIntersections! for a batch of states x[env][num_x] with kinematics and J' w on the device
    (floating-joint scenes): dict(f_generalized[env][nv], sdot[env][n_bristle][6], n_pairs, flags)."""
    if m.backend is None or not m.device_kinematics:
        raise RuntimeError("the scene has joints the device prologue does not handle (or no CUDA backend): use force_all_elastic_intersections")
    return m.backend.eval_state_f64(np.atleast_2d(np.asarray(x, dtype=np.float64)))


def force_all_elastic_intersections(m: MechanismScenario, x: Optional[np.ndarray] = None):
    """forceAllElasticIntersections!(m, tm) for the Float64 mode: returns a dict with
    f_generalized (nv), sdot (6 n_bristle), wrench [ins][6], n_pairs [ins], flags [ins]."""
    if m.backend is None:
        raise RuntimeError("finalize(m, backend=...) first")
    x = get_state(m) if x is None else np.asarray(x, dtype=np.float64)
    X, tw, s = boundary_arrays(m, x)
    out = m.backend.eval_f64(X, tw, s.reshape(1, m.n_bristle, 6) if m.n_bristle else None)
    w = out["wrench"][0]
    f = generalized_forces(m, x, w)
    sdot = out["sdot"][0].reshape(-1) if m.n_bristle else np.zeros(0)
    m.TT_Cache_n_pairs = out["n_pairs"][0]
    return dict(f_generalized=f, sdot=sdot, wrench=w, n_pairs=out["n_pairs"][0], flags=out["flags"][0])
