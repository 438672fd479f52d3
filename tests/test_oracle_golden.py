"""The travelling oracle against the reference's own outputs (tests/golden/, produced by
oracle/gen_golden.py from the unmodified reference).  These pin the oracle on every machine."""
import numpy as np
import pytest

from oracle import scape_oracle as so
from _helpers import PrePara, check_against_golden, golden_chunk

FAST_CASES = ["chr19", "chr17", "synth_fixed", "synth_rerun"]


def _run_oracle(golden, case, limit=None):
    spec = golden["cases"][case]
    chunk = golden_chunk(golden, case)[:limit]
    rng = np.random.RandomState(1)
    out = []
    for gi, df in chunk:
        cols = (np.array(df["x"]), np.array(df["l"]), np.array(df["r"]), np.array(df["pa"]))
        if "pre_para" in spec:
            pp = PrePara(spec["pre_para"])
            res = so.fit_utr_fixed(*cols, rng, pp.alpha_arr, pp.beta_arr, pp.L, **spec["params"])
        else:
            res = so.fit_utr(*cols, rng, **spec["params"])
        out.append(res)
    return out


@pytest.mark.parametrize("case", FAST_CASES)
def test_oracle_reproduces_reference(golden, case):
    res = _run_oracle(golden, case)
    for i, r in enumerate(res):
        check_against_golden(r, golden["cases"][case]["utrs"][i], golden["labels"][f"{case}/{i}"], tight=True)


def test_oracle_serial_stream_with_prunes(golden):
    """First 4 UTRs of synth8: one RNG stream, results depend on the draws consumed by the prunes
    of the UTRs before (apa_core.py:843)."""
    res = _run_oracle(golden, "synth8", limit=4)
    for i, r in enumerate(res):
        check_against_golden(r, golden["cases"]["synth8"]["utrs"][i], golden["labels"][f"synth8/{i}"], tight=True)


def test_shipped_result_pickles_are_loose_goldens(golden):
    """The pickles shipped with the reference came from an older RNG stream:
    they pin K and alpha (within one theta step) and the hard labels (>= 99 %)."""
    for case in ("chr17", "chr19"):
        for i, rec in enumerate(golden["cases"][case]["utrs"]):
            old_a = golden["inputs"][f"{case}/{i}/shipped_alpha"]
            assert len(old_a) == rec["K"]
            assert np.max(np.abs(old_a - np.array(rec["alpha_arr"]))) <= 9
            old_l = golden["inputs"][f"{case}/{i}/shipped_label"]
            assert np.mean(old_l == golden["labels"][f"{case}/{i}"]) >= 0.99
