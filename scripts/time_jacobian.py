#!/usr/bin/env python
"""Whole-Jacobian timing on the C3 workload (n_env x test/boxes.jl): pfc_calcxd_jacobian_device, CUDA events on the library's stream,
plus a checksum of the Jacobian (to see at a glance that an experiment changed no result).  Under `ncu --metrics gpu__time_duration.sum`
its launch list gives the per-kernel split."""
import hashlib
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
    from helpers import boxes_env_states, scene_boxes
    from pfc_b200 import capi
    from pfc_b200 import scenario as S
    n_env = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    ctx = capi.Context(0)
    m, _ = scene_boxes(ctx, max_env=n_env)
    x = boxes_env_states(m, n_env)
    nx = x.shape[1]
    dev = torch.device("cuda", 0)
    xd = torch.from_numpy(x).to(dev)
    jac = torch.empty((n_env, nx, nx), dtype=torch.float64, device=dev)
    xdot = torch.empty((n_env, nx), dtype=torch.float64, device=dev)
    npairs = torch.zeros((n_env, ctx.n_ins), dtype=torch.int64, device=dev)
    fl = torch.zeros((n_env, ctx.n_ins), dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    call = lambda: ctx.calcxd_jacobian_device(n_env, xd.data_ptr(), None, jac.data_ptr(), xdot.data_ptr(), npairs.data_ptr(), fl.data_ptr())
    call(); ctx.sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    with torch.cuda.stream(stream):
        ev[0].record()
        for k in range(reps):
            call()
            ev[k + 1].record()
    ctx.sync()
    ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(reps))
    j = jac.cpu().numpy()
    print(json.dumps({"n_env": n_env, "n_x": nx, "ms_per_jacobian_batch_median": ms[len(ms) // 2], "ms_min": ms[0],
                      "jac_sha": hashlib.sha256(j.tobytes()).hexdigest()[:16], "jac_abs_sum": float(np.abs(j).sum())}))


if __name__ == "__main__":
    main()
