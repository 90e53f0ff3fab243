"""Refit of a mesh whose vertices moved (refit half): geometry.refit_tree on the host (CPU tests) and
pfc_refit_mesh on the device against a scene rebuilt from the moved mesh and evaluated by the oracle (GPU tests)."""
import numpy as np
import pytest

import pfc_b200  # noqa: F401
from helpers import boxes_env_states, scene_boxes, wrench_rel_err
from oracle import orc
from pfc_b200 import geometry as G
from pfc_b200 import scenario as S


def _deform(points, seed):
    """A smooth deformation: anisotropic scaling, a small shear and a bend; keeps every tetrahedron of the test meshes positively oriented."""
    rng = np.random.default_rng(seed)
    A = np.diag([1.15, 0.9, 1.05]) + 0.05 * rng.uniform(-1, 1, (3, 3))
    q = points @ A.T
    q[:, 2] += 2.0 * q[:, 0] ** 2
    return q


def test_refit_tree_of_unmoved_mesh_is_the_tree():
    for mesh in (G.as_tet_eMesh(G.eMesh_box(0.05)), G.as_tri_eMesh(G.eMesh_sphere(0.05, 3)), G.as_tet_eMesh(G.eMesh_sphere(0.05, 3))):
        tree = G.eMesh_to_tree(mesh)
        again = G.refit_tree(tree, mesh)
        for a, b in ((tree.c, again.c), (tree.e, again.e), (tree.R, again.R)):
            assert np.array_equal(a, b)


def test_refit_boxes_contain_their_primitives():
    mesh = G.as_tet_eMesh(G.eMesh_sphere(0.05, 3))
    tree = G.eMesh_to_tree(mesh)
    moved = G.eMesh(point=_deform(mesh.point, 1), tet=mesh.tet, eps=mesh.eps)
    new = G.refit_tree(tree, moved)
    prims_below = [None] * new.n_node
    for k in range(new.n_node - 1, -1, -1):     # pre-order: children have larger indices
        prims_below[k] = [int(new.leaf_id[k])] if new.leaf_id[k] >= 0 else prims_below[new.left[k]] + prims_below[new.right[k]]
    for k in range(new.n_node):
        R = new.R[k].reshape(3, 3).T            # column-major storage
        pts = moved.point[moved.tet[prims_below[k]]].reshape(-1, 3)
        local = (pts - new.c[k]) @ R
        assert (np.abs(local) <= new.e[k] * (1 + 1e-12) + 1e-15).all(), k
    assert not np.allclose(new.c, tree.c)


@pytest.mark.gpu
def test_device_refit_matches_a_rebuilt_scene():
    """test/boxes.jl with box 2 (tets) and box 3 (triangles) deformed: the device refit (primitive records, leaf and internal boxes) gives
    the same candidate-pair lists (bit-exact, incl. order) and the same wrenches (1e-9) as a fresh scene built from the moved meshes and
    their host-refitted trees, evaluated by the oracle; a second refit back to the original geometry restores the original results."""
    from pfc_b200 import capi
    n_env = 64
    m_gpu = scene_boxes(capi.Context(0), max_env=n_env)[0]
    x = boxes_env_states(m_gpu, n_env)
    X, tw, _ = S.boundary_arrays(m_gpu, x)
    before = m_gpu.backend.eval_f64(X, tw, keep=True)
    orig = {k: m_gpu.MeshCache[k].mesh.point.copy() for k in (2, 3)}
    for k in (2, 3):
        S.refit_mesh(m_gpu, k, _deform(orig[k], k) * 1.02)
    after = m_gpu.backend.eval_f64(X, tw, keep=True)
    # the same scene built from scratch for the oracle: moved meshes + host-refitted trees
    m_cpu = scene_boxes(None)[0]
    for k in (2, 3):
        S.refit_mesh(m_cpu, k, _deform(orig[k], k) * 1.02)
    S.attach_backend(m_cpu, orc.OracleContext())
    ref = m_cpu.backend.eval_f64(X, tw, keep=True)
    assert not np.array_equal(before["n_pairs"], after["n_pairs"])
    assert np.array_equal(after["n_pairs"], ref["n_pairs"]) and np.array_equal(after["flags"], ref["flags"])
    for e in range(0, n_env, 7):
        for k in range(4):
            assert np.array_equal(m_gpu.backend.get_pairs(e, k), m_cpu.backend.get_pairs(e, k))
    assert wrench_rel_err(after["wrench"], ref["wrench"], floor=1e-9 * np.abs(ref["wrench"]).max()) <= 1e-9
    for k in (2, 3):
        S.refit_mesh(m_gpu, k, orig[k])
    again = m_gpu.backend.eval_f64(X, tw)
    assert np.array_equal(again["n_pairs"], before["n_pairs"])
    assert wrench_rel_err(again["wrench"], before["wrench"], floor=1e-9 * np.abs(before["wrench"]).max()) <= 1e-12


@pytest.mark.gpu
def test_device_refit_large_path_and_errors():
    """A 1280-primitive sphere pair on the large path (tri-tet), deformed; and the error path: a refit that inverts a tetrahedron is refused."""
    from pfc_b200 import capi

    def build(backend):
        m = S.MechanismScenario()
        sph = G.eMesh_sphere(0.05, 8)
        a = S.add_body_contact(m, "a", G.as_tri_eMesh(sph), i_prop=S.InertiaProperties(400.0, d=0.01))
        b = S.add_body_contact(m, "b", G.as_tet_eMesh(sph), i_prop=S.InertiaProperties(400.0), c_prop=S.ContactProperties(1.0e6))
        S.add_friction_regularize(m, a[2], b[2], mu_d=0.3, chi=0.5, n_quad_rule=2)
        S.finalize(m, backend)
        S.set_state_spq(m, a[0], trans=(0.0, 0.0, 0.0))
        S.set_state_spq(m, b[0], trans=(0.004, -0.003, 0.08), w=(0.3, 0.2, 1.0), vel=(0.01, 0.0, -0.1))
        return m

    m_gpu = build(capi.Context(0))
    x = S.get_state(m_gpu)
    X, tw, _ = S.boundary_arrays(m_gpu, x)
    pts = {k: m_gpu.MeshCache[k].mesh.point.copy() for k in (0, 1)}
    squash = lambda p: p * np.array([1.1, 0.95, 0.9])
    for k in (0, 1):
        S.refit_mesh(m_gpu, k, squash(pts[k]))
    g = m_gpu.backend.eval_f64(X, tw, keep=True)
    m_cpu = build(None)
    for k in (0, 1):
        S.refit_mesh(m_cpu, k, squash(pts[k]))
    S.attach_backend(m_cpu, orc.OracleContext())
    c = m_cpu.backend.eval_f64(X, tw, keep=True)
    assert c["n_pairs"][0, 0] > 100 and (c["flags"] & 1).all()
    assert np.array_equal(g["n_pairs"], c["n_pairs"]) and np.array_equal(m_gpu.backend.get_pairs(0, 0), m_cpu.backend.get_pairs(0, 0))
    assert wrench_rel_err(g["wrench"], c["wrench"], floor=1e-9 * np.abs(c["wrench"]).max()) <= 1e-9
    with pytest.raises(capi.PfcError) as ei:
        m_gpu.backend.refit_mesh(1, squash(pts[1]) * np.array([1.0, 1.0, -1.0]))      # mirrored: every tetrahedron inverted
    assert ei.value.code == -5
    with pytest.raises(capi.PfcError):
        m_gpu.backend.refit_mesh(1, pts[1][:-1])                                       # a different number of points
