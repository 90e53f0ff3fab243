"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in pfc_b200.parallel: environment
sharding without a collective on the path, and the partial-sum allreduce of a sharded large scene.
The compute is done by the CPU oracle here; on GPUs the same driver code runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pfc_b200  # noqa: F401
from helpers import boxes_env_states, scene_boxes


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import orc
        from pfc_b200 import geometry as G
        from pfc_b200 import parallel
        from pfc_b200 import scenario as S

        # (1) batched environments: contiguous env ranges, results gathered only for checking
        n_env = 37
        m, _ = scene_boxes(orc.OracleContext())
        X, tw, _ = S.boundary_arrays(m, boxes_env_states(m, n_env))
        lo, hi = parallel.env_range(n_env, rank, world)
        local = m.backend.eval_f64(X[lo:hi], tw[lo:hi])["wrench"]
        gathered = parallel.gather_env_results(local, n_env)
        full = m.backend.eval_f64(X, tw)["wrench"]
        ok_env = np.array_equal(gathered, full)

        # (2) one large scene: slice partials + allreduce(sum) == full wrench
        sc = S.MechanismScenario()
        ground = S.add_contact(sc, "ground", G.as_tet_eMesh(G.eMesh_sphere(0.06, 6)), c_prop=S.ContactProperties(2.0e6))
        b = S.add_body_contact(sc, "ball", G.as_tri_eMesh(G.eMesh_sphere(0.05, 6)), i_prop=S.InertiaProperties(400.0, d=0.01))
        S.add_friction_regularize(sc, b[2], ground, mu_s=0.4, mu_d=0.3, chi=0.5, n_quad_rule=2)
        ctx = orc.OracleContext()
        S.finalize(sc, ctx)
        S.set_state_spq(sc, b[0], trans=(0.0, 0.02, 0.095), w=(0.3, -0.2, 0.5), vel=(0.1, 0.0, -0.05))
        Xs, tws, _ = S.boundary_arrays(sc, S.get_state(sc))
        part, n_pairs = ctx.eval_slice_regularized(Xs, tws, 0, rank, world)
        t = torch.from_numpy(part.copy())
        parallel.allreduce_sum_(t)
        ref = ctx.eval_f64(Xs, tws)
        ok_large = n_pairs == int(ref["n_pairs"][0, 0]) and n_pairs > 256 and np.allclose(t.numpy(), ref["wrench"][0, 0], rtol=1e-12, atol=1e-15)
        q.put((rank, ok_env, ok_large, n_pairs))
    except Exception as exc:  # report instead of letting the parent wait for its timeout
        q.put((rank, False, False, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_env, ok_large, n_pairs in res:
        assert ok_env, f"rank {rank}: env-sharded results differ from the full batch"
        assert ok_large, f"rank {rank}: allreduced slice partials differ from the full wrench ({n_pairs} pairs)"


def test_env_range_partition():
    from pfc_b200 import parallel
    for n in (1, 7, 4096):
        for world in (1, 2, 3, 8):
            edges = [parallel.env_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
