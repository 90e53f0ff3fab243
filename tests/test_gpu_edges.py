"""GPU edge cases of the C ABI: empty batches, argument / call-order errors, the reference's "Non-finite vertex likely" error path
(src/clip/static_clip.jl:52), growth of the candidate-pair buffers (the reference's VectorCache doubling, src/obb/vector_cache.jl:11-15),
ragged scenes (instructions of very different sizes, small and large paths in one evaluation)."""
import numpy as np
import pytest

import pfc_b200  # noqa: F401
from helpers import boxes_env_states, scene_boxes, wrench_rel_err
from oracle import orc
from pfc_b200 import capi, scenes
from pfc_b200 import scenario as S

pytestmark = pytest.mark.gpu


def test_empty_batches_are_no_ops():
    m = scene_boxes(capi.Context(0), max_env=4)[0]
    L, h = capi.lib(), m.backend._h
    assert L.pfc_eval_f64(h, 0, None, None, None, None, None, None, None) == 0
    out = m.backend.eval_f64(np.zeros((0, 4, 16)), np.zeros((0, 4, 6)))
    assert out["wrench"].shape == (0, 4, 6)
    assert m.backend.eval_state_f64(np.zeros((0, S.num_x(m))))["f_generalized"].shape == (0, m.nv)
    assert m.backend.calcxd_f64(np.zeros((0, S.num_x(m))))["xdot"].shape == (0, S.num_x(m))


def test_argument_and_call_order_errors():
    ctx = capi.Context(0)
    L = capi.lib()
    x = np.zeros(48)
    # nothing works before pfc_finalize; the message names the function
    assert L.pfc_eval_f64(ctx._h, 1, x.ctypes.data, x.ctypes.data, None, x.ctypes.data, None, None, None) == -1
    assert b"pfc_eval_f64" in L.pfc_last_error()
    m = scene_boxes(ctx, max_env=2)[0]
    X, tw, _ = S.boundary_arrays(m, boxes_env_states(m, 2))
    w = np.zeros((2, 4, 6))
    assert L.pfc_eval_f64(ctx._h, 2, None, tw.ctypes.data, None, w.ctypes.data, None, None, None) == -1      # NULL input
    assert L.pfc_eval_f64(ctx._h, -1, X.ctypes.data, tw.ctypes.data, None, w.ctypes.data, None, None, None) == -1
    with pytest.raises(capi.PfcError):
        m.backend.calcxd_dual6(boxes_env_states(m, 1), 48)                                                    # seed_start out of range
    with pytest.raises(capi.PfcError):
        m.backend.set_dynamics(-np.tile(np.eye(6), (5, 1, 1)), [0.0, 0.0, -9.8])                              # inertia not positive definite
    with pytest.raises(capi.PfcError):
        m.backend.set_dynamics(np.tile(np.triu(np.ones((6, 6))) + 5 * np.eye(6), (5, 1, 1)), [0.0, 0.0, -9.8])  # not symmetric
    # a scene with revolute / prismatic chains has no device kinematics: the state-level calls are refused, the boundary level works
    m_chain, _ = scenes.scene_c2_pencil(True, capi.Context(0))
    assert L.pfc_eval_state_f64(m_chain.backend._h, 1, x.ctypes.data, x.ctypes.data, None, None, None) == -1
    out = S.force_all_elastic_intersections(m_chain)
    assert np.isfinite(out["f_generalized"]).all()


def test_non_finite_transform_reports_the_reference_error():
    """A NaN in x_r2_r1 makes every comparison of the clipper false: the reference throws "Non-finite vertex likely"; here the call
    returns PFC_E_NONFINITE, the offending (environment, instruction) carries PFC_FLAG_NONFINITE, the other environments are unaffected,
    and the oracle flags the same entries."""
    n_env = 3
    m_gpu = scene_boxes(capi.Context(0), max_env=n_env)[0]
    m_cpu = scene_boxes(orc.OracleContext())[0]
    X, tw, _ = S.boundary_arrays(m_gpu, boxes_env_states(m_gpu, n_env))
    good = m_gpu.backend.eval_f64(X, tw)
    X[1, 2, 12] = np.nan
    out = dict(wrench=np.zeros((n_env, 4, 6)), sdot=None, n_pairs=np.zeros((n_env, 4), np.int64), flags=np.zeros((n_env, 4), np.int32))
    with pytest.raises(capi.PfcError) as ei:
        m_gpu.backend.eval_f64(X, tw, out=out)
    assert ei.value.code == -4 and "Non-finite vertex likely" in str(ei.value)
    assert out["flags"][1, 2] & 2
    c = m_cpu.backend.eval_f64(X, tw)
    assert np.array_equal(out["flags"] & 2, c["flags"] & 2)
    assert np.array_equal(out["wrench"][[0, 2]], good["wrench"][[0, 2]])


def test_pair_buffers_grow_like_vector_cache():
    """Three environments of the 64-body pile at n_div = 16: 1.2 M candidate pairs, more than the initial 2^20-entry pair buffer, so
    the traversal overflows, the buffers double and the evaluation re-runs; results still match the oracle."""
    m_gpu, x = scenes.scene_c5_pile(4, 16, backend=capi.Context(0))
    m_cpu, _ = scenes.scene_c5_pile(4, 16, backend=orc.OracleContext(n_threads=orc.lib().orc_max_threads()))
    X1, tw1, _ = S.boundary_arrays(m_gpu, x)
    X, tw = np.repeat(X1, 3, axis=0), np.repeat(tw1, 3, axis=0)
    X[1, :, 12:15] += 1e-4          # three slightly different states
    X[2, :, 12:15] -= 1e-4
    g = m_gpu.backend.eval_f64(X, tw)
    assert g["n_pairs"].sum() > (1 << 20)
    c = m_cpu.backend.eval_f64(X, tw)
    assert np.array_equal(g["n_pairs"], c["n_pairs"]) and np.array_equal(g["flags"], c["flags"])
    assert wrench_rel_err(g["wrench"], c["wrench"], floor=1e-9 * np.abs(c["wrench"]).max()) <= 1e-9
    # and a second, smaller evaluation on the grown buffers is still right
    g1 = m_gpu.backend.eval_f64(X[:1], tw[:1])
    assert np.array_equal(g1["n_pairs"], c["n_pairs"][:1]) and np.array_equal(g1["wrench"], g["wrench"][:1])


def test_ragged_scene_mixes_small_and_large_instructions():
    """The pencil scene (config C2): a 1-tet half-space, 320-tet pads and a swept triangle mesh give instructions from 48 x 1 to
    320 x 320 leaf pairs -- the small and the large path in one evaluation, for a batch of different states."""
    m_gpu, bodies = scenes.scene_c2_pencil(True, capi.Context(0))
    m_cpu, _ = scenes.scene_c2_pencil(True, orc.OracleContext())
    xs = scenes.pencil_sample_states(m_gpu, bodies, n=5)
    arrs = [S.boundary_arrays(m_gpu, x) for x in xs]
    X, tw, s = (np.concatenate([a[i] for a in arrs], axis=0) for i in range(3))
    nb = m_gpu.n_bristle
    g = m_gpu.backend.eval_f64(X, tw, s.reshape(len(xs), nb, 6))
    c = m_cpu.backend.eval_f64(X, tw, s.reshape(len(xs), nb, 6))
    assert np.array_equal(g["n_pairs"], c["n_pairs"]) and np.array_equal(g["flags"], c["flags"])
    assert g["n_pairs"].max() > 200 and (g["n_pairs"] == 0).any() and (g["flags"] & 1).sum() >= 4   # ragged: empty lists next to long ones
    # regularized and bristle instructions alike at 1e-9 (the bristle ones run in the reference's operation order: csrc/pfc_exact.cuh).
    # Halves of a wrench that nearly vanish by symmetry carry eps |F| L rounding noise: floor of 1e-6 of the largest component, as in
    # test_gpu_parity.C2_FLOOR
    floor = 1e-6 * np.abs(c["wrench"]).max()
    assert wrench_rel_err(g["wrench"], c["wrench"], floor=floor) <= 1e-9
    assert wrench_rel_err(g["sdot"], c["sdot"], floor=1e-9 * np.abs(c["sdot"]).max()) <= 1e-9
