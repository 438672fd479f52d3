"""Shared helpers of the test-suite (golden fixtures, tolerance checks)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def golden_chunk(golden, case):
    """Rebuild the (gene_info_str, DataFrame) chunk of a golden case."""
    import pandas as pd
    from scape_b200 import synth
    spec = golden["cases"][case]
    if spec["source"] == "synth":
        utrs = [synth.make_utr(u, spec["synth_reads"]) for u in spec["synth_utrs"]]
        return [(u.gene_info_str, synth.to_dataframe(u)) for u in utrs]
    chunk = []
    for i, rec in enumerate(spec["utrs"]):
        g = lambda k: golden["inputs"][f"{case}/{i}/{k}"]
        n = len(g("x"))
        nan = np.full(n, np.nan)
        df = pd.DataFrame({"x": g("x").astype(np.int64), "l": g("l").astype(np.int64), "r": g("r").astype(np.float64),
                           "pa": g("pa").astype(np.float64), "cb_id": g("cb_id").astype(np.int64),
                           "read_id": np.arange(n, dtype=np.int64), "junction": np.zeros(n, np.int64),
                           "seg1_en": nan, "seg2_en": nan}, columns=synth.COLUMNS)
        chunk.append((rec["gene_info_str"], df))
    return chunk


class PrePara:
    def __init__(self, d):
        self.alpha_arr = np.array(d["alpha_arr"])
        self.beta_arr = np.array(d["beta_arr"], dtype=float)
        self.L = d["L"]
        self.K = d["K"]


def check_against_golden(res, rec, labels, tight=True, lb_rtol=1e-9, ws_atol=1e-9):
    """BASELINE.json tolerances: K identical, |d alpha| <= 1 bp, |d beta|, |d ws| <= 1e-3,
    lb within 1e-6 relative.  `tight` additionally requires what FP64 parity delivers in practice."""
    assert int(res.K) == rec["K"]
    assert int(res.L) == rec["L"]
    assert res.title == rec["title"]
    assert np.max(np.abs(np.asarray(res.alpha_arr) - np.array(rec["alpha_arr"]))) <= 1
    assert np.max(np.abs(np.asarray(res.beta_arr) - np.array(rec["beta_arr"]))) <= 1e-3
    ws = np.array([float(v) for v in rec["ws"]])
    assert np.max(np.abs(np.asarray(res.ws) - ws)) <= 1e-3
    lb = np.array([float(v) for v in rec["lb_arr"]])
    assert abs(res.lb_arr[-1] - lb[-1]) <= 1e-6 * abs(lb[-1])
    assert abs(float(res.bic) - float(rec["bic"])) <= 1e-6 * abs(float(rec["bic"]))
    agree = np.mean(np.asarray(res.label_arr) == labels)
    assert agree >= 0.999
    if tight:
        assert len(res.lb_arr) == len(lb)
        assert np.array_equal(np.asarray(res.alpha_arr), np.array(rec["alpha_arr"]))
        assert np.allclose(np.asarray(res.lb_arr, dtype=float), lb, rtol=lb_rtol, atol=0)
        assert np.allclose(np.asarray(res.ws), ws, rtol=0, atol=ws_atol)
        assert agree == 1.0
