"""Synthetic scenes of the named benchmark configurations (BASELINE.json `configs`),
built with the host mirror of the reference's scenario API."""
from __future__ import annotations

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
