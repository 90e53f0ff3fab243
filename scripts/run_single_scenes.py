#!/usr/bin/env python
"""Single-scene latency of the contact-wrench evaluation on the reference's own test scenes (configs C1 / C2): one environment per
call through pfc_eval_f64 (host pointers, synchronous), beside the CPU oracle on one thread.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import pfc_b200  # noqa: F401
    from helpers import boxes_env_states, scene_boxes
    from oracle import orc
    from pfc_b200 import capi, scenes
    from pfc_b200 import scenario as S

    def rate(backend, X, tw, s, reps):
        backend.eval_f64(X, tw, s)
        t0 = time.perf_counter()
        for _ in range(reps):
            out = backend.eval_f64(X, tw, s)
        return reps / (time.perf_counter() - t0), out

    res = {}
    builders = {
        "C1 boxes (settled stack)": lambda b: (scene_boxes(b)[0], None),
        "C2 pencil (bristle)": lambda b: scenes.scene_c2_pencil(True, b),
        "C2 spoon (test/data/spoon.obj, bristle)": lambda b: scenes.scene_c2_spoon(b),
    }
    only = sys.argv[1] if len(sys.argv) > 1 else ""      # e.g. "pencil": that scene only (for profiling)
    reps_gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    for name, build in builders.items():
        if only and only not in name:
            continue
        m_g, bodies = build(capi.Context(0))
        m_c, _ = build(orc.OracleContext())
        if name.startswith("C1"):
            x = boxes_env_states(m_g, 1)[0]
        elif "pencil" in name:
            x = scenes.pencil_sample_states(m_g, bodies, n=2)[1]
        else:
            x = scenes.spoon_sample_states(m_g, bodies, n=2)[1]
        X, tw, s = S.boundary_arrays(m_g, x)
        s = s.reshape(1, m_g.n_bristle, 6) if m_g.n_bristle else None
        g_rate, g = rate(m_g.backend, X, tw, s, reps_gpu)
        c_rate, c = rate(m_c.backend, X, tw, s, 20)
        assert np.array_equal(g["n_pairs"], c["n_pairs"])
        res[name] = {"gpu_evals_per_sec": g_rate, "gpu_us_per_eval": 1e6 / g_rate, "oracle_1_thread_evals_per_sec": c_rate,
                     "candidate_pairs": int(c["n_pairs"].sum()), "instructions": int(c["n_pairs"].size), "contacts": int((c["flags"] & 1).sum())}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
