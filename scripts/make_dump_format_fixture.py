#!/usr/bin/env python
"""Writes tests/golden/dump_format_selfcheck.json: a dump in the format of baseline/julia/dump_reference.jl ("pfc-reference-dump-1"),
produced by THIS repository's host mirror (pressurefieldcontact.jl_b200/scenario.py) and CPU oracle -- NOT by the Julia reference.
Its only purpose is to keep tests/test_reference_dump.py's loader and checks exercised where no Julia dump exists.  Scene: the pad /
box / plane mix of tests (tri-tet regularized, tet-tet regularized, tri-tet bristle), three sampled states."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pfc_b200  # noqa: E402,F401
from oracle import orc  # noqa: E402
from pfc_b200 import geometry as G  # noqa: E402
from pfc_b200 import scenario as S  # noqa: E402


def build(backend):
    r = 0.05
    c_prop = S.ContactProperties(1.0e6)
    m = S.MechanismScenario()
    plane = S.add_contact(m, "plane", G.as_tet_eMesh(G.eMesh_half_plane()), c_prop=c_prop)
    b1 = S.add_body_contact(m, "box_1", G.as_tri_eMesh(G.eMesh_box(r)), i_prop=S.InertiaProperties(400.0, d=r))
    b2 = S.add_body_contact(m, "box_2", G.as_tet_eMesh(G.eMesh_box(r)), i_prop=S.InertiaProperties(400.0), c_prop=c_prop)
    b3 = S.add_body_contact(m, "box_3", G.as_tet_eMesh(G.eMesh_box(r)), i_prop=S.InertiaProperties(400.0), c_prop=S.ContactProperties(3.0e6))
    S.add_friction_regularize(m, plane, b1[2], mu_d=0.0, chi=2.2, n_quad_rule=2)
    S.add_friction_bristle(m, b1[2], b2[2], mu_d=0.3, chi=0.3, k_bar=1.0e5, tau=0.02, n_quad_rule=1)
    S.add_friction_regularize(m, b2[2], b3[2], mu_s=0.4, mu_d=0.3, chi=0.7, v_tol=1e-3, n_quad_rule=2)
    S.finalize(m, backend, 1)
    return m


def main():
    m = build(orc.OracleContext())
    rng = np.random.default_rng(12)
    r = 0.05
    doc = {"format": "pfc-reference-dump-1", "scene": "selfcheck",
           "provenance": "NOT Julia output: written by scripts/make_dump_format_fixture.py from the host mirror + CPU oracle of this repository",
           "nq": m.nq, "nv": m.nv, "n_bristle": m.n_bristle, "gravity": list(map(float, m.gravity)), "meshes": [], "instructions": [], "samples": []}
    for mc in m.MeshCache:
        em, t = mc.mesh, mc.tree
        prim = em.tet if em.is_tet else em.tri
        doc["meshes"].append({"name": mc.name, "kind": "tet" if em.is_tet else "tri", "body_id": int(mc.body_id) + 1,
                              "Ebar": None if mc.c_prop is None else mc.c_prop.E_bar, "point": em.point.tolist(), "prim": (np.asarray(prim) + 1).tolist(),
                              "eps": None if em.eps is None else np.asarray(em.eps).tolist(),
                              "tree": {"c": t.c.tolist(), "e": t.e.tolist(), "R": t.R.tolist(), "left": t.left.tolist(), "right": t.right.tolist(), "leaf_id": t.leaf_id.tolist()}})
    for ci in m.ContactInstructions:
        fm = ci.friction_model
        doc["instructions"].append({"id_1": int(ci.id_1), "id_2": int(ci.id_2), "chi": ci.chi, "model": fm.model, "params": list(map(float, fm.params())),
                                    "n_quad_rule": ci.n_quad_rule, "bristle_id": int(getattr(fm, "bristle_id", -1)) if fm.model == 1 else -1})
    for k in range(3):
        x = np.zeros(S.num_x(m))
        nq = m.nq
        x[0:3] = rng.uniform(-0.03, 0.03, 3); x[3:6] = [0.004, -0.003, r - 0.002]
        x[6:9] = rng.uniform(-0.03, 0.03, 3); x[9:12] = [0.01, 0.0, 3 * r - 0.005]
        x[12:15] = rng.uniform(-0.03, 0.03, 3); x[15:18] = [0.0, 0.01, 5 * r - 0.009]
        x[nq:nq + m.nv] = rng.uniform(-1, 1, m.nv) * 0.2
        x[nq + m.nv:] = rng.uniform(-1, 1, 6 * m.n_bristle) * 1e-4
        X, tw, s = S.boundary_arrays(m, x)
        out = m.backend.eval_f64(X, tw, s.reshape(1, m.n_bristle, 6), keep=True)
        per = []
        for i in range(len(m.ContactInstructions)):
            per.append({"X_r2_r1": X[0, i].tolist(), "twist_r2": tw[0, i].tolist(), "pairs": m.backend.get_pairs(0, i).tolist(),
                        "traction": m.backend.get_traction(0, i).tolist(), "contact": bool(out["flags"][0, i] & 1), "wrench_r2": out["wrench"][0, i].tolist()})
        doc["samples"].append({"x": x.tolist(), "per_instruction": per, "sdot": out["sdot"][0].reshape(-1).tolist(),
                               "f_generalized": S.generalized_forces(m, x, out["wrench"][0]).tolist()})
    path = os.path.join(ROOT, "tests", "golden", "dump_format_selfcheck.json")
    json.dump(doc, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes;", sum(p["contact"] for s_ in doc["samples"] for p in s_["per_instruction"]), "contacts")


if __name__ == "__main__":
    main()
