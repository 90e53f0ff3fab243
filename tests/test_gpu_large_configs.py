"""GPU parity on the two single-large-scene configurations at FULL size, and the multi-GPU split of a large scene at world 2 and 4.

  C4  sphere (100 820 tets) on slab (99 856 tets), tet-tet, one instruction, ~86 k candidate pairs (BASELINE.json configs[3])
  C5  64-body pile at n_div 24: 2016 tri-tet instructions, ~0.9 M candidate pairs per evaluation (configs[4])

Bars: pair counts and flags of every instruction equal; EVERY candidate-pair list equal to the oracle's, in order; wrenches <= 1e-9.
The split is emulated with `world` contexts on one GPU (each context owns rank r of world: hash-partitioned sub-trees, disjoint pair
lists), the partial buffers summed on the host -- what the NCCL exchange of bench.py / parallel.eval_sharded does across GPUs."""
import ctypes

import numpy as np
import pytest

import pfc_b200  # noqa: F401
from helpers import wrench_rel_err
from oracle import orc
from pfc_b200 import capi, scenes
from pfc_b200 import scenario as S

pytestmark = pytest.mark.gpu
TOL = 1.0e-9


def _full_parity(m, x):
    ctx = capi.Context(0)
    S.attach_backend(m, ctx)
    X, tw, _ = S.boundary_arrays(m, x)
    g = ctx.eval_f64(X, tw, None, keep=True)
    octx = orc.OracleContext(n_threads=orc.lib().orc_max_threads())
    S.attach_backend(m, octx)
    c = octx.eval_f64(X, tw, None, keep=True)
    assert np.array_equal(g["n_pairs"], c["n_pairs"]) and np.array_equal(g["flags"], c["flags"])
    for k in range(ctx.n_ins):                     # every list, in the reference's traversal order
        if c["n_pairs"][0, k]:
            assert np.array_equal(ctx.get_pairs(0, k), octx.get_pairs(0, k)), k
    assert wrench_rel_err(g["wrench"], c["wrench"], floor=1e-9 * np.abs(c["wrench"]).max()) <= TOL
    n_tests, n_pairs = ctx.counters()
    assert n_pairs == int(c["n_pairs"].sum())
    return c


def test_c4_sphere_on_slab_full_size():
    m, x = scenes.scene_c4_sphere_on_slab(71, 79)
    c = _full_parity(m, x)
    assert c["n_pairs"].sum() > 80000 and (c["flags"] & 1).all()


def test_c5_pile_full_size():
    m, x = scenes.scene_c5_pile(4, 24)
    c = _full_parity(m, x)
    assert c["n_pairs"].sum() > 850000 and (c["flags"] & 1).sum() > 100


def test_replayed_graph_of_a_many_kernel_scene_follows_the_inputs():
    """Scenes with large or bristle instructions replay a captured CUDA graph from their third identical evaluation on (same batch size,
    same device buffers).  The pile with bristle instructions, evaluated six times with OTHER states in the same buffers -- other pair
    counts, other contacts: every result equals what a fresh context (first evaluation: never a graph) computes from the same state."""
    m, x0 = scenes.scene_c5_pile(3, 8, bristle_every=7)
    ctx = capi.Context(0)
    S.attach_backend(m, ctx)
    nb = ctx.n_bristle
    rng = np.random.default_rng(42)
    counts = set()
    for it in range(6):
        x = x0.copy()
        x[:m.nq] += rng.uniform(-0.004, 0.004, m.nq)            # nudged poses: other pair lists
        X, tw, s = S.boundary_arrays(m, x)
        s_arr = s.reshape(1, nb, 6)
        got = ctx.eval_f64(X, tw, s_arr)
        m2, _ = scenes.scene_c5_pile(3, 8, bristle_every=7)
        fresh = capi.Context(0)
        S.attach_backend(m2, fresh)
        want = fresh.eval_f64(X, tw, s_arr)
        assert np.array_equal(got["n_pairs"], want["n_pairs"]) and np.array_equal(got["flags"], want["flags"]), it
        assert got["wrench"].tobytes() == want["wrench"].tobytes() and got["sdot"].tobytes() == want["sdot"].tobytes(), it
        counts.add(int(got["n_pairs"].sum()))
    assert len(counts) > 1


def _sharded_on_one_gpu(build, world):
    """Returns (unsharded result, per-rank results after the exchange, per-rank pair lists of every instruction)."""
    import torch
    m_ref, x = build()
    ref = capi.Context(0)
    S.attach_backend(m_ref, ref)
    X, tw, _ = S.boundary_arrays(m_ref, x)
    full = ref.eval_f64(X, tw, None, keep=True)
    n_ins = ref.n_ins
    dev = torch.device("cuda", 0)
    cudart = ctypes.CDLL("libcudart.so")
    ranks, bufs = [], []
    for r in range(world):
        m_r, _ = build()
        ctx = capi.Context(0)
        S.attach_backend(m_r, ctx)
        ctx.set_shard(r, world)
        ctx.set_debug(True)
        b = dict(X=torch.from_numpy(X).to(dev), tw=torch.from_numpy(tw).to(dev), w=torch.zeros((1, n_ins, 6), dtype=torch.float64, device=dev),
                 np_=torch.zeros((1, n_ins), dtype=torch.int64, device=dev), fl=torch.zeros((1, n_ins), dtype=torch.int32, device=dev))
        ctx.eval_sharded_begin(1, b["X"].data_ptr(), b["tw"].data_ptr(), None, b["w"].data_ptr(), None, b["np_"].data_ptr(), b["fl"].data_ptr())
        ranks.append(ctx)
        bufs.append(b)
    parts = []
    for ctx in ranks:
        ctx.sync()
        parts.append(ctx.eval_sharded_partials())
    host = [np.zeros(cnt) for _, cnt in parts]
    for h, (ptr, cnt) in zip(host, parts):
        assert cudart.cudaMemcpy(h.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(ptr), ctypes.c_size_t(8 * cnt), 2) == 0
    total = host[0].copy()
    for h in host[1:]:                             # fixed rank order
        total += h
    for ptr, cnt in parts:
        assert cudart.cudaMemcpy(ctypes.c_void_p(ptr), total.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(8 * cnt), 1) == 0
    for ctx in ranks:
        assert ctx.eval_sharded_step() == 0        # one exchange per evaluation
        ctx.sync()
    return ref, full, ranks, bufs


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("scene", ["C4", "C5"])
def test_sharded_split_matches_the_unsharded_evaluation(scene, world):
    build = (lambda: scenes.scene_c4_sphere_on_slab(24, 27)) if scene == "C4" else (lambda: scenes.scene_c5_pile(3, 8))
    ref, full, ranks, bufs = _sharded_on_one_gpu(build, world)
    scale = np.abs(full["wrench"]).max()
    n_ins = ref.n_ins
    for ctx, b in zip(ranks, bufs):
        assert wrench_rel_err(b["w"].cpu().numpy(), full["wrench"], floor=1e-9 * scale) <= 1e-11   # same traction points, another association
        assert np.array_equal(b["np_"].cpu().numpy(), full["n_pairs"])
        assert np.array_equal(b["fl"].cpu().numpy() & 1, full["flags"] & 1)
    # the ranks' pair lists are disjoint and their union is the full list; each part keeps the reference's order
    busiest = np.argsort(-full["n_pairs"][0])[:8]
    for k in busiest:
        whole = [tuple(p) for p in ref.get_pairs(0, int(k))]
        pos = {p: i for i, p in enumerate(whole)}
        seen = set()
        for ctx in ranks:
            part = [tuple(p) for p in ctx.get_pairs(0, int(k))]
            idx = [pos[p] for p in part]
            assert idx == sorted(idx)
            assert not (seen & set(part))
            seen |= set(part)
        assert seen == set(whole)
    bits = [b["w"].cpu().numpy().tobytes() for b in bufs]
    assert all(x == bits[0] for x in bits)         # every rank ends with the same bits


def test_sort_key_capacity_is_an_error_not_a_wrong_answer(monkeypatch):
    """Advisor finding: the sort key packs (problem, DFS key) in 64 bits.  Two caterpillar trees of depth 31 need 62 key bits; with 8
    environments (3 problem bits) the key does not fit: the call must fail with PFC_E_CAPACITY instead of aliasing problems."""
    from pfc_b200 import geometry as G
    monkeypatch.setenv("PFC_MID_LEAVES", "0")      # 32-leaf trees would take the on-chip path: this test is about the large path's sort key

    def caterpillar_tet_mesh(n):
        # n thin tetrahedra in a row; the tree is a chain: node = (leaf k, rest)
        pts, tets = [], []
        for k in range(n):
            x0 = 0.01 * k
            pts += [[x0, 0.0, 0.0], [x0 + 0.009, 0.0, 0.0], [x0, 0.009, 0.0], [x0, 0.0, 0.009]]
            tets.append([4 * k, 4 * k + 1, 4 * k + 2, 4 * k + 3])
        eps = np.zeros(len(pts)); eps[3::4] = 1.0
        mesh = G.eMesh(np.array(pts), None, np.array(tets), eps)
        lo = np.array(pts).reshape(n, 4, 3).min(axis=1); hi = np.array(pts).reshape(n, 4, 3).max(axis=1)
        c, e, R, left, right, leaf = [], [], [], [], [], []

        def add(node_lo, node_hi, l, r, lf):
            c.append((node_hi + node_lo) / 2); e.append((node_hi - node_lo) / 2); R.append(np.eye(3).reshape(9)); left.append(l); right.append(r); leaf.append(lf)
            return len(c) - 1
        # pre-order chain: internal node i has children (leaf i, internal i + 1); the last internal node holds the last two leaves
        for k in range(n - 1):
            add(lo[k:].min(axis=0), hi[k:].max(axis=0), 2 * k + 1, 2 * k + 2, -1)
            add(lo[k], hi[k], -1, -1, k)
        add(lo[n - 1], hi[n - 1], -1, -1, n - 1)
        tree = G.FlatTree(np.array(c), np.array(e), np.array(R), np.array(left, np.int32), np.array(right, np.int32), np.array(leaf, np.int32))
        return mesh, tree

    n = 32                                          # chain depth 31 per tree, 32 x 32 = 1024 leaf pairs > 512: the large path
    mesh, tree = caterpillar_tet_mesh(n)
    ctx = capi.Context(0)
    a = ctx.add_mesh(1, mesh.point, mesh.tet, mesh.eps, 1.0e6, tree)
    b = ctx.add_mesh(1, mesh.point, mesh.tet, mesh.eps, 1.0e6, tree)
    ctx.add_instruction(a, b, 0.5, 0, np.array([0.3, 0.3, 0.01]), 2)
    ctx.finalize(8)
    X = np.tile(np.eye(4).reshape(1, 1, 16), (8, 1, 1)); X[:, 0, 12] = 0.001
    tw = np.zeros((8, 1, 6))
    ok = ctx.eval_f64(X[:1], tw[:1])                # one environment: 62 key bits + 0 problem bits fit
    assert ok["n_pairs"][0, 0] > 0
    with pytest.raises(capi.PfcError) as ei:
        ctx.eval_f64(X, tw)                         # 8 environments: 62 + 3 bits do not
    assert ei.value.code == -3
