#!/usr/bin/env python
"""Where the time of a split large scene goes, rank by rank: rank r of `world` is emulated on ONE GPU (a context with
pfc_set_shard(r, world) does exactly the work that rank does before the exchange: shared top levels, its own sub-trees, sort, narrow
phase, partial sums) and timed with CUDA events on the library's stream; the node-pair tests and candidate pairs of the rank come from the
library's counters.  Prints one JSON line; under `ncu --metrics gpu__time_duration.sum` the launch list gives the per-kernel split of a rank.

  python scripts/shard_breakdown.py C5 4 [rank]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import pfc_b200  # noqa: F401
    from pfc_b200 import capi, scenes
    from pfc_b200 import scenario as S
    scene = sys.argv[1] if len(sys.argv) > 1 else "C5"
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    only = int(sys.argv[3]) if len(sys.argv) > 3 else -1
    build = (lambda: scenes.scene_c4_sphere_on_slab(71, 79)) if scene == "C4" else (lambda: scenes.scene_c5_pile(4, 24))
    dev = torch.device("cuda", 0)
    out = {"scene": scene, "world": world, "ranks": []}
    for r in ([only] if only >= 0 else list(range(world)) + [-1]):   # -1: the unsplit evaluation
        m, x = build()
        ctx = capi.Context(0)
        S.attach_backend(m, ctx)
        if r >= 0:
            ctx.set_shard(r, world)
        X, tw, _ = S.boundary_arrays(m, x)
        n_ins = ctx.n_ins
        Xd, twd = torch.from_numpy(X).to(dev), torch.from_numpy(tw).to(dev)
        w = torch.zeros((1, n_ins, 6), dtype=torch.float64, device=dev)
        npairs = torch.zeros((1, n_ins), dtype=torch.int64, device=dev)
        fl = torch.zeros((1, n_ins), dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        if r >= 0:
            call = lambda: ctx.eval_sharded_begin(1, Xd.data_ptr(), twd.data_ptr(), None, w.data_ptr(), None, npairs.data_ptr(), fl.data_ptr())
        else:
            call = lambda: ctx.eval_f64_device(1, Xd.data_ptr(), twd.data_ptr(), None, w.data_ptr(), None, npairs.data_ptr(), fl.data_ptr())
        for _ in range(4):
            call(); ctx.sync()
        reps = 10
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        with torch.cuda.stream(stream):
            ev[0].record()
            for k in range(reps):
                call()
                ev[k + 1].record()
        ctx.sync()
        ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(reps))
        n_tests, n_pairs = ctx.counters()
        out["ranks"].append({"rank": r if r >= 0 else "unsplit", "ms_median": ms[len(ms) // 2], "node_pairs_tested": int(n_tests), "candidate_pairs": int(n_pairs)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
