"""Synthetic scenes of the named benchmark configurations (BASELINE.json `configs`),
built with the host mirror of the reference's scenario API."""
from __future__ import annotations

import os

import numpy as np

from . import geometry as G
from . import scenario as S


def _splitmix(seed):
    state = seed & 0xFFFFFFFFFFFFFFFF

    def nxt():
        nonlocal state
        state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        z = z ^ (z >> 31)
        return (z >> 11) * (1.0 / 9007199254740992.0)

    return nxt


def scene_c4_sphere_on_slab(n_div: int = 71, n_cell: int = 79, backend=None):
    """C4: test_vol_vol-style tet-tet contact at scale.  eMesh_sphere(0.05, n_div) (20 n_div^2 tets) pressed
    5 mm into a compliant slab = extrude_mesh of an n_cell x n_cell triangulated square (16 n_cell^2 tets),
    side 0.2, thickness 0.05, E = 1e6 both; relative twist (0, 0, 1; 0.05, 0, -0.1); regularized friction
    mu 0.3, chi 0.5, quadrature rule 2.  One tet-tet instruction.  Returns (scenario, state x)."""
    m = S.MechanismScenario()
    slab_mesh = G.as_tet_eMesh(G.extrude_mesh(G.eMesh_grid_square(0.2, n_cell), 0.05))
    G.transform(slab_mesh, t=(0.0, 0.0, -0.025))  # top face at z = 0
    sph_mesh = G.as_tet_eMesh(G.eMesh_sphere(0.05, n_div))
    c_prop = S.ContactProperties(1.0e6)
    slab = S.add_contact(m, "slab", slab_mesh, c_prop=c_prop, tree=G.eMesh_to_tree(slab_mesh, method="top_down"))
    body, _, sph = S.add_body_contact(m, "sphere", sph_mesh, i_prop=S.InertiaProperties(400.0), c_prop=c_prop,
                                      tree=G.eMesh_to_tree(sph_mesh, method="top_down"))
    S.add_friction_regularize(m, slab, sph, mu_d=0.3, chi=0.5, n_quad_rule=2)
    S.finalize(m, backend)
    S.set_state_spq(m, body, trans=(0.0, 0.0, 0.045), w=(0.0, 0.0, 1.0), vel=(0.05, 0.0, -0.1))
    return m, S.get_state(m)


def scene_c5_pile(n_side: int = 4, n_div: int = 8, backend=None, bristle_every: int = 0):
    """C5: n_side^3 compliant spheres (eMesh_sphere(0.05, n_div), 20 n_div^2 primitives), each registered as a
    triangle mesh AND a tet mesh; instructions tri_i - tet_j for all i < j; centres on a lattice with
    spacing 0.09 (10 % overlap) jittered U(-0.005, 0.005)^3, random rotations and twists, splitmix64 seed
    0x5EED0064.  Regularized friction (every `bristle_every`-th instruction bristle if > 0)."""
    u = _splitmix(0x5EED0064)
    U = lambda lo, hi: lo + (hi - lo) * u()
    sph = G.eMesh_sphere(0.05, n_div)
    tri_mesh, tet_mesh = G.as_tri_eMesh(sph), G.as_tet_eMesh(sph)
    tri_tree, tet_tree = G.eMesh_to_tree(tri_mesh), G.eMesh_to_tree(tet_mesh)
    m = S.MechanismScenario()
    c_prop = S.ContactProperties(1.0e6)
    n_body = n_side ** 3
    ids = []
    for b in range(n_body):
        body = S.add_body(m, f"ball_{b}")
        ids.append((body, S.add_contact(m, f"ball_{b}_tri", tri_mesh, body=body, tree=tri_tree),
                    S.add_contact(m, f"ball_{b}_tet", tet_mesh, c_prop=c_prop, body=body, tree=tet_tree)))
    k = 0
    for i in range(n_body):
        for j in range(i + 1, n_body):
            k += 1
            if bristle_every and k % bristle_every == 0:
                S.add_friction_bristle(m, ids[i][1], ids[j][2], mu_d=0.3, chi=0.5, k_bar=2.0e4, tau=0.05, n_quad_rule=2)
            else:
                S.add_friction_regularize(m, ids[i][1], ids[j][2], mu_d=0.3, chi=0.5, n_quad_rule=2)
    S.finalize(m, backend)
    x = np.zeros(S.num_x(m))
    nq = m.nq
    for b in range(n_body):
        body = m.bodies[ids[b][0]]
        ix, iy, iz = b % n_side, (b // n_side) % n_side, b // (n_side * n_side)
        pos = [0.09 * ix + U(-0.005, 0.005), 0.09 * iy + U(-0.005, 0.005), 0.09 * iz + U(-0.005, 0.005)]
        mrp = [U(-0.3, 0.3) for _ in range(3)]
        x[body.q0:body.q0 + 6] = mrp + pos
        x[nq + body.v0:nq + body.v0 + 6] = [U(-1, 1) for _ in range(3)] + [U(-0.1, 0.1) for _ in range(3)]
    x[nq + m.nv:] = [U(-1e-4, 1e-4) for _ in range(6 * m.n_bristle)]
    return m, x


def scene_c2_pencil(is_bristle: bool = True, backend=None):
    """C2: the gripper-and-pencil scene of test/pencil.jl:23-33,185-236 (is_bristle = true variant): a 1-tet
    half-space, two compliant ellipsoidal pads (eMesh_sphere(pad_rad, 4) scaled (2, 1, 2): 320 tets each) on
    prismatic joints of an arm (prismatic z -> revolute y), a rigid box on the arm, and a 12-sided swept-mesh
    pencil (48 triangles) on a floating joint.  Instructions, in the reference's order: pencil-pad_n and
    pencil-pad_p (bristle: k_bar 8e4, tau 0.01, magic 1e-2; mu_d 0.5, chi 0.6), pencil-plane and pad_n-pad_p
    (tet-tet, mu_d 0) regularized with v_tol 1e-5.  Returns (scenario, dict of body indices)."""
    pad_rad, penci_length, penci_rad = 0.0035, 0.16, 0.0035
    m = S.MechanismScenario()
    c_prop = S.ContactProperties(1.0e6)
    plane = G.transform(G.eMesh_half_plane(), R=0.6)
    pads = []
    for sgn in (-1.0, +1.0):
        pad = G.as_tet_eMesh(G.eMesh_sphere(pad_rad, 4))
        G.transform(pad, t=(0.0, sgn * (pad_rad + penci_rad), 0.0))
        G.transform(pad, R=np.diag([2.0, 1.0, 2.0]))
        pads.append(pad)
    arm_box = G.as_tri_eMesh(G.eMesh_box(pad_rad * np.array([1.0, 7.0, 1.0]), pad_rad * np.array([0.0, -4.0, 8.0])))
    pencil = G.as_tri_eMesh(G.create_swept_mesh(G.f_swept_triv, [0.0, 0.013, penci_length], [0.0, penci_rad, penci_rad], 12, True, rot_half=True))
    G.transform(pencil, t=(0.0, 0.0, penci_rad))

    id_plane = S.add_contact(m, "plane", G.as_tet_eMesh(plane), c_prop=c_prop)
    b_tra_z = S.add_body(m, "tra_z", joint=S.Prismatic((0.0, 0.0, 1.0)))
    b_rev_y = S.add_body(m, "rev_y", joint=S.Revolute((0.0, 1.0, 0.0)), body=b_tra_z)
    S.add_contact(m, "rev_y", arm_box, body=b_rev_y)
    b_pad_n, _, id_pad_n = S.add_body_contact(m, "pad_n", pads[0], c_prop=c_prop, joint=S.Prismatic((0.0, +1.0, 0.0)), body=b_rev_y)
    b_pad_p, _, id_pad_p = S.add_body_contact(m, "pad_p", pads[1], c_prop=c_prop, joint=S.Prismatic((0.0, -1.0, 0.0)), body=b_rev_y)
    b_penci, _, id_penci = S.add_body_contact(m, "name", pencil)
    if is_bristle:
        S.add_friction_bristle(m, id_penci, id_pad_n, mu_d=0.5, chi=0.6, k_bar=8.0e4, magic=1.0e-2, tau=0.01)
        S.add_friction_bristle(m, id_penci, id_pad_p, mu_d=0.5, chi=0.6, k_bar=8.0e4, magic=1.0e-2, tau=0.01)
    else:
        S.add_friction_regularize(m, id_penci, id_pad_n, mu_d=0.5, chi=0.6, v_tol=1.0e-5)
        S.add_friction_regularize(m, id_penci, id_pad_p, mu_d=0.5, chi=0.6, v_tol=1.0e-5)
    S.add_friction_regularize(m, id_penci, id_plane, mu_d=0.5, chi=0.6, v_tol=1.0e-5)
    S.add_friction_regularize(m, id_pad_n, id_pad_p, mu_d=0.0, chi=0.6, v_tol=1.0e-5)
    S.finalize(m, backend)
    bodies = dict(tra_z=b_tra_z, rev_y=b_rev_y, pad_n=b_pad_n, pad_p=b_pad_p, pencil=b_penci)
    return m, bodies


def pencil_sample_states(m, bodies, n: int = 8, seed: int = 0x5EED0C2):
    """Sampled states along the pencil task (there is no integrator here): the
    initial state of test/pencil.jl:225-228, the pencil gripped on the table, lifted, swung by the arm, and
    the pads closed on each other; randomised velocities and bristle states.  Returns x[n][num_x]."""
    u = _splitmix(seed)
    U = lambda lo, hi: lo + (hi - lo) * u()
    L, r = 0.16, 0.0035
    rz = lambda a: np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
    ry = lambda a: np.array([[np.cos(a), 0.0, np.sin(a)], [0.0, 1.0, 0.0], [-np.sin(a), 0.0, np.cos(a)]])
    xs = []
    for k in range(n):
        mode = k % 4
        m.q[:] = 0.0
        m.v[:] = 0.0
        if mode == 0:      # test/pencil.jl initial configuration: arm raised and turned, pencil lying on the table
            z, th, grip = 0.10, np.pi / 4, 0.0
            S.set_state_spq(m, bodies["pencil"], rot=rz(-np.pi / 2), trans=(-0.8 * L, 0.0, -U(2e-5, 2e-4)))
        elif mode == 1:    # arm lowered onto the pencil, pads squeezing it against the table
            z, th, grip = r, 0.0, U(1e-4, 5e-4)
            S.set_state_spq(m, bodies["pencil"], rot=rz(-np.pi / 2 + U(-0.002, 0.002)), trans=(-0.5 * L, U(-1e-4, 1e-4), -U(2e-5, 2e-4)))
        elif mode == 2:    # lifted and swung: the pencil moves with the arm frame
            z, th, grip = 0.08, U(0.2, 1.2), U(1e-4, 5e-4)
            Rw = ry(th)    # the pencil's frame expressed in the arm frame is RotZ(-pi/2) with its axis through the pads
            S.set_state_spq(m, bodies["pencil"], rot=Rw @ rz(-np.pi / 2), trans=tuple(np.array([0.0, 0.0, z]) + Rw @ np.array([-0.4 * L, 0.0, -r])))
        else:              # pencil away on the table, the pads pressed against each other (tet-tet contact)
            z, th, grip = 0.05, U(-0.5, 0.5), r + U(1e-4, 4e-4)
            S.set_state_spq(m, bodies["pencil"], rot=rz(U(-1, 1)), trans=(0.3, 0.1, -U(2e-5, 2e-4)))
        S.set_configuration(m, bodies["tra_z"], [z])
        S.set_configuration(m, bodies["rev_y"], [th])
        S.set_configuration(m, bodies["pad_n"], [grip])
        S.set_configuration(m, bodies["pad_p"], [grip])
        v_amp = 0.05 if (k // 4) % 2 else 0.002   # fast relative motion switches contacts off through the damping term
        m.v[:] = [U(-v_amp, v_amp) for _ in range(m.nv)]
        m.s[:] = [U(-2e-4, 2e-4) for _ in range(6 * m.n_bristle)]
        xs.append(S.get_state(m).copy())
    return np.array(xs)


def eMesh_spoon_like(n_ring: int = 139, n_side: int = 18, length: float = 0.15) -> "G.eMesh":
    """A closed spoon-shaped triangle surface with 2 * n_side * n_ring triangles (5004 for the defaults: the size of
    the reference's test/data/spoon.obj, 2502 quads, which is reference data and is not copied here).  Elliptical
    cross-sections swept along x: a narrow handle widening into a bowl; flat facets on the bottom (z = 0) and top."""
    xs = np.linspace(-0.25 * length, 0.75 * length, n_ring)
    half_w = 0.004 + 0.014 * np.exp(-((xs / (0.22 * length)) ** 2))        # bowl centred at x = 0
    half_t = np.full(n_ring, 0.0015)
    phi = (np.arange(n_side) + 0.5) * (2.0 * np.pi / n_side) - 0.5 * np.pi  # a facet, not a vertex, faces -z and +z
    pts = [np.stack([np.full(n_side, x), a * np.cos(phi), b + b * np.sin(phi) / np.sin(phi).max()], axis=1) for x, a, b in zip(xs, half_w, half_t)]
    pts = np.concatenate(pts + [np.array([[xs[0], 0.0, half_t[0]], [xs[-1], 0.0, half_t[-1]]])])
    tri = []
    ring = lambda i, j: i * n_side + (j % n_side)
    for i in range(n_ring - 1):
        for j in range(n_side):
            a, b, c, d = ring(i, j), ring(i + 1, j), ring(i + 1, j + 1), ring(i, j + 1)
            tri += [(a, b, c), (a, c, d)]
    c0, c1 = n_ring * n_side, n_ring * n_side + 1
    for j in range(n_side):
        tri.append((c0, ring(0, j), ring(0, j + 1)))
        tri.append((c1, ring(n_ring - 1, j + 1), ring(n_ring - 1, j)))
    m = G.eMesh(pts, np.array(tri))
    # outward orientation: flip everything if the signed volume is negative
    p = m.point[m.tri]
    if np.einsum("ij,ij->i", p[:, 0], np.cross(p[:, 1], p[:, 2])).sum() < 0.0:
        m.tri = np.ascontiguousarray(m.tri[:, ::-1])
    return m


SPOON_FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "spoon_mesh.npz")


def eMesh_spoon_obj(path: str = SPOON_FIXTURE) -> "G.eMesh":
    """The reference's own spoon (test/data/spoon.obj: 2504 vertices, 2502 quadrilaterals -> 5004 triangles), from the committed
    vertex / face fixture (scripts/make_spoon_fixture.py), scaled by 0.01 as test/spoon.jl:37-38 does.  Every quadrilateral
    (a, b, c, d) is split into (a, b, c), (a, c, d) -- the fan GeometryTypes' decompose uses when MeshIO hands FileIO.load's faces over
    as triangles."""
    d = np.load(path)
    q = d["quads"].astype(np.int64)
    tri = np.concatenate([q[:, [0, 1, 2]], q[:, [0, 2, 3]]], axis=1).reshape(-1, 3)   # the two triangles of a quad stay adjacent
    return G.eMesh(d["vertices"] * 0.01, tri)


def scene_c2_spoon(backend=None, real_mesh=None):
    """C2 (spoon): test/spoon.jl:23-54 re-expressed in the current API: a rigid 5004-triangle spoon
    surface between a world-fixed compliant box and a compliant box on a z-prismatic joint (eMesh_box(0.02), 12
    tets, E 1e6); two bristle instructions mu 0.2, chi 0.2, quadrature rule 1.  Returns (scenario, bodies).
    real_mesh: True = the reference's spoon.obj (fixture), False = a swept stand-in of the same size, None = the fixture when present."""
    if real_mesh is None:
        real_mesh = os.path.exists(SPOON_FIXTURE)
    rad_box = 0.02
    m = S.MechanismScenario()
    c_prop = S.ContactProperties(1.0e6)
    box = G.as_tet_eMesh(G.eMesh_box(rad_box, (0.0, 0.0, -rad_box)))
    id_lo = S.add_contact(m, "box_lo", box, c_prop=c_prop)
    b_up, _, id_up = S.add_body_contact(m, "box_up", box.copy(), c_prop=c_prop, joint=S.Prismatic((0.0, 0.0, 1.0)))
    spoon = eMesh_spoon_obj() if real_mesh else eMesh_spoon_like()
    b_spoon, _, id_spoon = S.add_body_contact(m, "spoon", spoon)
    S.add_friction_bristle(m, id_spoon, id_lo, mu_d=0.2, chi=0.2, n_quad_rule=1)
    S.add_friction_bristle(m, id_spoon, id_up, mu_d=0.2, chi=0.2, n_quad_rule=1)
    S.finalize(m, backend)
    m.spoon_is_real = bool(real_mesh)
    m.spoon_points = spoon.point
    return m, dict(box_up=b_up, spoon=b_spoon)


def spoon_sample_states(m, bodies, n: int = 6, seed: int = 0x5EED5B00):
    """Sampled states: the spoon resting on / pressed into the lower box, the upper box above it (initial state of
    test/spoon.jl: q_up = 0.10) or clamping it; small random twists and bristle states."""
    u = _splitmix(seed)
    U = lambda lo, hi: lo + (hi - lo) * u()
    rz = lambda a: np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
    rx90 = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]])   # RotX(pi / 2), test/spoon.jl:51
    xs = []
    for k in range(n):
        m.q[:] = 0.0
        m.v[:] = 0.0
        pen_lo = U(5e-5, 4e-4)
        if getattr(m, "spoon_is_real", False):
            # the reference's pose (RotX(pi/2), a little yaw and xy jitter); heights from the part of the spoon above the boxes' footprint
            R = rz(U(-0.3, 0.3)) @ rx90
            t_xy = np.array([U(-2e-3, 2e-3), U(-2e-3, 2e-3)])
            w = m.spoon_points @ R.T
            inside = (np.abs(w[:, 0] + t_xy[0]) < 0.02) & (np.abs(w[:, 1] + t_xy[1]) < 0.02)
            z_lo, z_hi = w[inside, 2].min(), w[inside, 2].max()
            S.set_state_spq(m, bodies["spoon"], rot=R, trans=(t_xy[0], t_xy[1], -z_lo - pen_lo))
            q_up = 0.10 if k % 3 == 0 else 0.04 + (z_hi - z_lo) - pen_lo - U(5e-5, 4e-4)   # box_up spans [q - 0.04, q]
            S.set_configuration(m, bodies["box_up"], [q_up])
            m.v[:] = [U(-0.003, 0.003) for _ in range(m.nv)]
            m.s[:] = [U(-2e-4, 2e-4) for _ in range(6 * m.n_bristle)]
            xs.append(S.get_state(m).copy())
            continue
        S.set_state_spq(m, bodies["spoon"], rot=rz(U(-0.3, 0.3)), trans=(U(-2e-3, 2e-3), U(-2e-3, 2e-3), -pen_lo))
        q_up = 0.10 if k % 3 == 0 else 0.04 + 0.003 - pen_lo - U(5e-5, 4e-4)  # box_up spans [q - 0.04, q]; the spoon is 3 mm thick
        S.set_configuration(m, bodies["box_up"], [q_up])
        m.v[:] = [U(-0.003, 0.003) for _ in range(m.nv)]
        m.s[:] = [U(-2e-4, 2e-4) for _ in range(6 * m.n_bristle)]
        xs.append(S.get_state(m).copy())
    return np.array(xs)
