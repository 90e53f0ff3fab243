#!/usr/bin/env python
"""One very large scene split across the GPUs of a box (config C4): every rank traverses and sorts (identical pair
lists), evaluates its slice of the 256-pair chunks, and the per-instruction partial sums are all-reduced over NCCL.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/run_sharded.py

Rank 0 checks the result against the CPU oracle (wrench <= 1e-9) and prints one JSON line (time = max over ranks,
CUDA events on the library's stream)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class _RawCuda:
    """A device pointer as a __cuda_array_interface__ object (float64 vector)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-div", type=int, default=71)
    ap.add_argument("--n-cell", type=int, default=79)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--bristle", action="store_true")
    ap.add_argument("--scene", default="C4", choices=["C4", "C5"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import pfc_b200  # noqa: F401
    from helpers import wrench_rel_err
    from pfc_b200 import capi, parallel, scenes
    from pfc_b200 import scenario as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m, x = scenes.scene_c4_sphere_on_slab(args.n_div, args.n_cell) if args.scene == "C4" else scenes.scene_c5_pile(4, 24)
    ctx = capi.Context(local_rank)
    S.attach_backend(m, ctx)
    ctx.set_shard(rank, world)
    X, tw, s = S.boundary_arrays(m, x)
    n_ins = ctx.n_ins
    Xd, twd = torch.from_numpy(X).to(dev), torch.from_numpy(tw).to(dev)
    w = torch.zeros((1, n_ins, 6), dtype=torch.float64, device=dev)
    npairs = torch.zeros((1, n_ins), dtype=torch.int64, device=dev)
    fl = torch.zeros((1, n_ins), dtype=torch.int32, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def reduce_partials(ptr, count):
        t = torch.as_tensor(_RawCuda(ptr, count), device=dev)
        with torch.cuda.stream(stream):
            parallel.allreduce_sum_(t)

    def step():
        return parallel.eval_sharded(ctx, 1, Xd, twd, None, w, None, npairs, fl, reduce_partials)

    for _ in range(3):
        n_exchange = step()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    ctx.sync()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    n_tests, n_pairs = ctx.counters()
    if rank == 0:
        from oracle import orc
        octx = orc.OracleContext(n_threads=orc.lib().orc_max_threads())
        S.attach_backend(m, octx)
        ref = octx.eval_f64(X, tw, None)
        wg = w.cpu().numpy()
        err = wrench_rel_err(wg, ref["wrench"], floor=1e-9 * np.abs(ref["wrench"]).max())
        assert np.array_equal(npairs.cpu().numpy(), ref["n_pairs"]), "pair counts differ"
        n_total = int(ref["n_pairs"].sum())
        assert err <= 1e-9, err
        print(json.dumps({"scene": args.scene + " sharded", "n_gpus": world, "candidate_pairs": n_total, "candidate_pairs_listed_by_rank_0": int(n_pairs),
                          "node_pairs_tested_by_rank_0": int(n_tests),
                          "ms_per_eval": float(ms[0]), "candidate_pairs_per_sec": n_total / (float(ms[0]) * 1e-3), "nccl_exchanges_per_eval": n_exchange,
                          "wrench_rel_err_vs_oracle": err, "collective": "NCCL all_reduce(sum) of 23 doubles per large instruction" if world > 1 else "none"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
