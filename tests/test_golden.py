"""Committed fixtures (tests/golden/, written by scripts/make_golden.py).

radau_tables.json IS pinned by reference data: when it was generated, the Butcher data were compared with the reference's own table
files (src/radau/table/{1,2,3}_rule) and the largest deviations are recorded in the file.  boxes_c1_oracle.json holds outputs of the
CPU oracle (the Julia reference cannot run in the build image): a regression fixture that ties the oracle and the CUDA path to each
other across commits, not a reference output."""
import json
import os

import numpy as np
import pytest

import pfc_b200  # noqa: F401
from helpers import scene_boxes, wrench_rel_err
from oracle import orc
from pfc_b200 import radau as R
from pfc_b200 import scenario as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_radau_tables_match_the_fixture_and_the_reference_files():
    tabs = json.load(open(os.path.join(GOLD, "radau_tables.json")))
    for n in (1, 2, 3):
        rec, t = tabs[str(n)], R.radau_table(n)
        assert np.allclose(rec["A"], t.A, atol=1e-15) and np.allclose(rec["c"], t.c, atol=1e-15) and np.allclose(rec["b_hat"], t.b_hat, atol=1e-14)
        assert np.allclose(np.array(rec["lambda_re"]) + 1j * np.array(rec["lambda_im"]), t.lam, atol=1e-13)
        dev = rec["max_abs_deviation_from_reference_table_files"]     # recorded against /root/reference at generation time
        assert max(dev.values()) <= 2e-13


def _check(backend_out, gold, tol):
    assert np.array_equal(backend_out["n_pairs"], np.array(gold["n_pairs"]))
    assert np.array_equal(backend_out["flags"], np.array(gold["flags"]))
    ref = np.array(gold["wrench"])
    assert wrench_rel_err(backend_out["wrench"], ref, floor=1e-9 * np.abs(ref).max()) <= tol


def test_oracle_reproduces_the_boxes_fixture():
    gold = json.load(open(os.path.join(GOLD, "boxes_c1_oracle.json")))
    m, _ = scene_boxes(orc.OracleContext())
    x = np.array(gold["x"])
    X, tw, _ = S.boundary_arrays(m, x)
    out = m.backend.eval_f64(X, tw, None, keep=True)
    _check(out, gold, 1e-12)
    for e in range(x.shape[0]):
        for k in range(4):
            assert m.backend.get_pairs(e, k)[:6].tolist() == gold["first_pairs"][e][k]
    assert (np.array(gold["flags"])[0] & 1).sum() == 0 and (np.array(gold["flags"])[1:] & 1).sum() >= 24   # the drop touches nothing; the stacks do


@pytest.mark.gpu
def test_cuda_reproduces_the_boxes_fixture():
    from pfc_b200 import capi
    gold = json.load(open(os.path.join(GOLD, "boxes_c1_oracle.json")))
    x = np.array(gold["x"])
    m, _ = scene_boxes(capi.Context(0), max_env=x.shape[0])
    m.backend.set_debug(True)
    X, tw, _ = S.boundary_arrays(m, x)
    out = m.backend.eval_f64(X, tw, None, keep=True)
    _check(out, gold, 1e-9)
    for e in range(x.shape[0]):
        for k in range(4):
            assert m.backend.get_pairs(e, k)[:6].tolist() == gold["first_pairs"][e][k]
