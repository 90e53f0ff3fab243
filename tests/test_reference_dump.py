"""Parity against OUTPUTS OF THE JULIA REFERENCE ITSELF, when they are available.

baseline/julia/dump_reference.jl (run on a machine with Julia and the reference checkout) writes tests/golden/julia/<scene>.json:
meshes, flattened bin_BB_Trees, contact instructions, and for sampled states of a Radau run the boundary arrays of every
instruction with what the reference computed from them -- candidate-pair lists (m.TT_Cache), traction points
(tm.bodyBodyCache.TractionCache), the wrench yes_contact! returned, s-dot.  This test rebuilds each scene through the C ABI FROM THE
DUMP ALONE and holds the CPU oracle (always) and the CUDA library (-m gpu) to

    pair lists       bit-exact, in order                      src/obb/tree_types.jl:88-111
    traction points  same count; n, r, dA, p to 1e-11         src/contact_algorithms_non_friction.jl:217-265
    wrench           1e-9 per force / torque 3-vector         src/contact_algorithms_friction.jl:50-72, 119-143
    s-dot            1e-9 (regularized and full-rank bristle patches; see note)

Note on s-dot / bristle wrenches: the reference calls LAPACK's symmetric eigensolver inside K^(-1/2); on rank-deficient patches the
result is reproducible to 1e-9 only by the same eigensolver on the same bits (csrc/pfc_exact.cuh).  Entries where the dump's own K
is rank deficient (an eigenvalue at the 1e-16 clamp) are therefore compared on the well-conditioned part only: the wrench's normal
component and the pair / traction lists.

Without dumps the Julia part skips.  tests/golden/dump_format_selfcheck.json keeps the loader exercised: it has the same format but
was written by scripts/make_dump_format_fixture.py from this repository's own host mirror + oracle (NOT Julia output)."""
import glob
import json
import os

import numpy as np
import pytest

import pfc_b200  # noqa: F401
from helpers import wrench_rel_err
from oracle import orc
from pfc_b200 import geometry as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JULIA_DUMPS = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "julia", "*.json")))
SELFCHECK = os.path.join(ROOT, "tests", "golden", "dump_format_selfcheck.json")
TOL = 1.0e-9


def build_from_dump(doc, backend):
    """Scene -> backend through add_mesh / add_instruction, exactly the calls the Julia shim (julia/PFCGpu.jl) makes."""
    assert doc["format"] == "pfc-reference-dump-1"
    for md in doc["meshes"]:
        t = md["tree"]
        tree = G.FlatTree(np.array(t["c"], float), np.array(t["e"], float), np.array(t["R"], float), np.array(t["left"], np.int32),
                          np.array(t["right"], np.int32), np.array(t["leaf_id"], np.int32))
        prim = np.array(md["prim"], np.int64) - 1          # the reference's indices are 1-based
        kind = 1 if md["kind"] == "tet" else 0
        backend.add_mesh(kind, np.array(md["point"], float), prim.astype(np.int32), None if md["eps"] is None else np.array(md["eps"], float),
                         md["Ebar"] or 0.0, tree)
    for ci in doc["instructions"]:
        backend.add_instruction(ci["id_1"], ci["id_2"], ci["chi"], ci["model"], np.array(ci["params"], float), ci["n_quad_rule"])
    backend.finalize(1)
    return backend


def check_against_dump(doc, backend, what):
    n_ins = len(doc["instructions"])
    nb = doc["n_bristle"]
    nq, nv = doc["nq"], doc["nv"]
    n_contact = 0
    for smp in doc["samples"]:
        per = smp["per_instruction"]
        X = np.array([p["X_r2_r1"] for p in per], float).reshape(1, n_ins, 16)
        tw = np.array([p["twist_r2"] for p in per], float).reshape(1, n_ins, 6)
        s = np.array(smp["x"], float)[nq + nv:].reshape(1, nb, 6) if nb else None
        out = backend.eval_f64(X, tw, s, keep=True)
        for k, p in enumerate(per):
            ref_pairs = np.array(p["pairs"], np.int32).reshape(-1, 2)
            assert out["n_pairs"][0, k] == len(ref_pairs), (what, k)
            assert np.array_equal(backend.get_pairs(0, k), ref_pairs), (what, "pair list", k)
            ref_tr = np.array(p["traction"], float).reshape(-1, 8)
            tr = backend.get_traction(0, k)
            assert tr.shape == ref_tr.shape, (what, "traction count", k)
            if len(ref_tr):
                assert np.allclose(tr, ref_tr, rtol=1e-11, atol=1e-13 * np.abs(ref_tr).max()), (what, "traction", k)
            assert bool(out["flags"][0, k] & 1) == bool(p["contact"]), (what, "contact flag", k)
            n_contact += int(p["contact"])
            w_ref = np.array(p["wrench_r2"], float)
            ci = doc["instructions"][k]
            scale = max(np.abs(w_ref).max(), 1e-300)
            if ci["model"] == 0 or not p.get("rank_deficient_K", False):
                assert wrench_rel_err(out["wrench"][0, k], w_ref, floor=1e-9 * scale) <= TOL, (what, "wrench", k)
        if nb:
            sd_ref = np.array(smp["sdot"], float).reshape(nb, 6)
            for ci, p in zip(doc["instructions"], per):
                if ci["model"] == 1 and not p.get("rank_deficient_K", False):
                    b = ci["bristle_id"]
                    assert wrench_rel_err(out["sdot"][0, b], sd_ref[b], floor=1e-9 * max(np.abs(sd_ref[b]).max(), 1e-300)) <= TOL, (what, "sdot", b)
    return n_contact


def _mark_rank_deficient(doc):
    """Flags bristle entries whose scaled stiffness has an eigenvalue at the clamp (from the dump's own traction points)."""
    for smp in doc["samples"]:
        for ci, p in zip(doc["instructions"], smp["per_instruction"]):
            if ci["model"] != 1 or not p["contact"]:
                continue
            _, _, K = orc.patch_stiffness(np.array(p["traction"], float).reshape(-1, 8), ci["params"][1])
            Sinv, _ = orc.decompose_K(K, ci["params"][4])
            Kf = np.triu(K) + np.triu(K, 1).T
            lam = np.linalg.eigvalsh(np.diag(Sinv) @ Kf @ np.diag(Sinv))
            p["rank_deficient_K"] = bool(lam.min() < 1e-12 * lam.max())


def test_dump_format_selfcheck_oracle():
    """The loader and the checks themselves, on a dump of the same format written by this repository's own mirror (not Julia)."""
    doc = json.load(open(SELFCHECK))
    assert doc["format"] == "pfc-reference-dump-1" and "NOT Julia" in doc["provenance"]
    assert check_against_dump(doc, build_from_dump(doc, orc.OracleContext()), "selfcheck/oracle") > 0


@pytest.mark.gpu
def test_dump_format_selfcheck_cuda():
    from pfc_b200 import capi
    doc = json.load(open(SELFCHECK))
    assert check_against_dump(doc, build_from_dump(doc, capi.Context(0)), "selfcheck/cuda") > 0


@pytest.mark.skipif(not JULIA_DUMPS, reason="no tests/golden/julia/*.json: run baseline/julia/dump_reference.jl on a machine with Julia")
@pytest.mark.parametrize("path", JULIA_DUMPS or ["-"])
def test_oracle_against_julia_dump(path):
    doc = json.load(open(path))
    _mark_rank_deficient(doc)
    assert check_against_dump(doc, build_from_dump(doc, orc.OracleContext()), os.path.basename(path) + "/oracle") > 0


@pytest.mark.gpu
@pytest.mark.skipif(not JULIA_DUMPS, reason="no tests/golden/julia/*.json: run baseline/julia/dump_reference.jl on a machine with Julia")
@pytest.mark.parametrize("path", JULIA_DUMPS or ["-"])
def test_cuda_against_julia_dump(path):
    from pfc_b200 import capi
    doc = json.load(open(path))
    _mark_rank_deficient(doc)
    assert check_against_dump(doc, build_from_dump(doc, capi.Context(0)), os.path.basename(path) + "/cuda") > 0
