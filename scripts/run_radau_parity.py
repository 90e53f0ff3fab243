#!/usr/bin/env python
"""State parity over N Radau steps on test/boxes.jl (config C1): integrates the scene with the reference's adaptive Radau IIA
scheme (pfc_b200.radau, a mirror of src/radau) twice -- contact wrenches and Dual-6 Jacobian chunks from the CUDA library, then
from the CPU oracle -- and prints one JSON line with the worst relative deviation of the integrated state over all steps."""
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
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--h-max", type=float, default=0.05)
    args = ap.parse_args()
    import pfc_b200  # noqa: F401
    from helpers import scene_boxes
    from oracle import orc
    from pfc_b200 import capi
    from pfc_b200 import dynamics as D
    from pfc_b200 import radau as R
    from pfc_b200 import scenario as S

    def integrate(backend):
        m = scene_boxes(backend)[0]
        dyn = D.FloatingBodyDynamics(m)
        rr = R.makeRadauIntegrator(dyn, S.num_x(m), 1.0e-16, 2, 6)
        rr.step.h_max = args.h_max
        t0 = time.time()
        ts, xs = R.integrate_radau(rr, S.get_state(m), t_final=1e9, max_steps=args.steps, after_step=lambda x: D.principal_value(m, x))
        return ts, xs, rr, time.time() - t0

    ts_g, xs_g, rr_g, wall_g = integrate(capi.Context(0))
    ts_c, xs_c, rr_c, wall_c = integrate(orc.OracleContext())
    worst, worst_step = 0.0, 0
    for k in range(1, len(ts_c)):
        scale = np.abs(xs_c[k]).max()
        for i in range(0, xs_c.shape[1], 3):
            den = max(np.abs(xs_c[k, i:i + 3]).max(), 1e-6 * scale)
            e = np.abs(xs_g[k, i:i + 3] - xs_c[k, i:i + 3]).max() / den
            if e > worst:
                worst, worst_step = float(e), k
    print(json.dumps({"scene": "C1 test/boxes.jl (drop)", "radau_steps": len(ts_c) - 1, "t_end": float(ts_c[-1]),
                      "worst_state_rel_err": worst, "at_step": worst_step, "time_grid_max_abs_diff": float(np.abs(ts_g - ts_c).max()),
                      "float_evals": rr_c.n_de_float, "dual6_chunk_evals": rr_c.n_de_chunk, "same_eval_counts": (rr_g.n_de_float, rr_g.n_de_chunk) == (rr_c.n_de_float, rr_c.n_de_chunk),
                      "final_z": [float(v) for v in xs_g[-1][[5, 11, 17, 23]]], "wall_s_gpu_backend": wall_g, "wall_s_oracle_backend": wall_c,
                      "note": "wall times are dominated by the Python host mirror of the integrator and RigidBodyDynamics, one scene at a time"}))


if __name__ == "__main__":
    main()
