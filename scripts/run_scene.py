#!/usr/bin/env python
"""Runs one of the large named configurations (C4 / C5) on the GPU, checks it against the CPU oracle
(pair lists bit-exact incl. order, wrench 1e-9) and prints one JSON line with device timings."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="C5", choices=["C4", "C5"])
    ap.add_argument("--n-div", type=int, default=None)
    ap.add_argument("--n-cell", type=int, default=79)
    ap.add_argument("--n-side", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    import torch
    import pfc_b200  # noqa: F401
    from helpers import wrench_rel_err
    from oracle import orc
    from pfc_b200 import capi, scenes
    from pfc_b200 import scenario as S

    t0 = time.time()
    if args.scene == "C4":
        m, x = scenes.scene_c4_sphere_on_slab(args.n_div or 71, args.n_cell)
    else:
        m, x = scenes.scene_c5_pile(args.n_side, args.n_div or 24)   # ~0.9 M candidate pairs (n_div 8: 0.11 M, 16: 0.41 M)
    t_build = time.time() - t0
    ctx = capi.Context(0)
    S.attach_backend(m, ctx)
    X, tw, s = S.boundary_arrays(m, x)
    n_ins = ctx.n_ins
    dev = torch.device("cuda", 0)
    Xd, twd = torch.from_numpy(X).to(dev), torch.from_numpy(tw).to(dev)
    w = torch.zeros((1, n_ins, 6), dtype=torch.float64, device=dev)
    npairs = torch.zeros((1, n_ins), dtype=torch.int64, device=dev)
    fl = torch.zeros((1, n_ins), dtype=torch.int32, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    step = lambda: ctx.eval_f64_device(1, Xd.data_ptr(), twd.data_ptr(), None, w.data_ptr(), None, npairs.data_ptr(), fl.data_ptr())
    for _ in range(3):
        step()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    ctx.sync()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (ctx.launch_count() - l0) // args.steps
    n_tests, n_pairs = ctx.counters()
    out = {"scene": args.scene, "n_instructions": n_ins, "meshes": len(m.MeshCache), "candidate_pairs": n_pairs, "node_pairs_tested": n_tests,
           "ms_per_eval": ms, "candidate_pairs_per_sec": n_pairs / (ms * 1e-3), "node_pairs_per_sec": n_tests / (ms * 1e-3),
           "kernel_launches_per_eval": launches, "contacts": int((fl.cpu().numpy() & 1).sum()), "build_s": t_build,
           "broad_phase_algorithmic_GBps": (272 * n_tests + 12 * n_pairs) / (ms * 1e-3) * 1e-9}
    if not args.no_check:
        octx = orc.OracleContext(n_threads=orc.lib().orc_max_threads())
        S.attach_backend(m, octx)
        t0 = time.time()
        ref = octx.eval_f64(X, tw, None, keep=True)
        out["oracle_s_all_cores_1_env"] = time.time() - t0
        o1 = orc.OracleContext(n_threads=1)
        S.attach_backend(m, o1)
        t0 = time.time()
        o1.eval_f64(X, tw, None)
        out["oracle_s_1_thread"] = time.time() - t0
        g = ctx.eval_f64(X, tw, None, keep=True)
        assert (g["n_pairs"] == ref["n_pairs"]).all(), "pair counts differ"
        assert (g["flags"] == ref["flags"]).all(), "flags differ"
        busiest = np.argsort(-ref["n_pairs"][0])[:4]
        for k in busiest:
            assert np.array_equal(ctx.get_pairs(0, int(k)), octx.get_pairs(0, int(k))), f"pair list of instruction {k} differs"
        scale = np.abs(ref["wrench"]).max()
        out["wrench_rel_err"] = wrench_rel_err(g["wrench"], ref["wrench"], floor=1e-9 * scale)
        assert out["wrench_rel_err"] <= 1e-9, out["wrench_rel_err"]
        out["parity"] = "pairs bit-exact (count all instructions, lists of the 4 busiest incl. order), wrench <= 1e-9"
        assert n_pairs == int(ref["n_pairs"].sum())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
