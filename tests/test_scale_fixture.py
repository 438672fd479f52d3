"""The at-scale fixtures (tests/golden/scale/, written by oracle/gen_scale_golden.py) are what the GPU
acceptance gate compares with.  Here, without a GPU: the oracle reproduces leading records of every
suite from the recorded RNG offsets (so fixture, generator and oracle stay in step), and the suites
have the composition VERDICT r01 asked for."""
import json
import os

import numpy as np
import pytest

from oracle import scape_oracle as so
from scape_b200 import synth
from _helpers import GOLD

SCALE = os.path.join(GOLD, "scale")


def _load(name):
    with open(os.path.join(SCALE, name + ".json")) as fh:
        meta = json.load(fh)
    return dict(np.load(os.path.join(SCALE, name + ".npz"))), meta


def _rng_at(offset):
    g = np.random.RandomState(1)
    if offset > 0:
        g.bytes(4 * int(offset))
    return g


def test_suite_composition():
    fx, _ = _load("cfg2")
    assert len(fx["index"]) == 1000 and len(set(fx["file_id"])) == 10 and np.all(fx["reads"] == 500)
    fx, _ = _load("cfg3")
    assert len(fx["index"]) >= 300 and int(np.sum(fx["reads"] >= 100000)) >= 5 and int(np.sum(fx["long_utr"])) >= 10
    fx, meta = _load("cfg4")
    assert len(fx["index"]) == 200 and len(meta["pre"]) == 4
    for name, kmax in (("kmax8", 8), ("kmax10", 10)):
        fx, meta = _load(name)
        assert len(fx["index"]) == 100 and meta["params"]["n_max_apa"] == kmax
    fx, meta = _load("cfg5")                       # giant UTRs (BASELINE.json configs[4], the stress sweep's far corner): 1M and 300k reads at n_max_apa = 10
    assert sorted(int(r) for r in fx["reads"]) == [300000, 1000000] and meta["params"]["n_max_apa"] == 10
    assert np.all(fx["chains_run"] >= 100)
    for name in ("cfg2", "cfg3", "cfg4", "kmax8", "kmax10", "cfg5"):
        fx, meta = _load(name)
        assert not meta["errors"]
        # rng_off is cumulative within a file and restarts with every file
        for f in set(fx["file_id"]):
            off = fx["rng_off"][fx["file_id"] == f]
            assert off[0] == 0 and np.all(np.diff(off) > 0)


@pytest.mark.parametrize("suite,rows", [("cfg2", [0, 1, 100]), ("kmax8", [0]), ("cfg4", [0, 1, 150])])
def test_oracle_reproduces_fixture_records(suite, rows):
    fx, meta = _load(suite)
    for j in rows:
        u = synth.make_utr(int(fx["index"][j]), int(fx["reads"][j]), long_utr=bool(fx["long_utr"][j]))
        rng = _rng_at(fx["rng_off"][j])                      # the file's stream where this UTR starts
        pre = meta["pre"][int(fx["file_id"][j])] if meta.get("pre") else None
        if pre:
            res = so.fit_utr_fixed(u.x, u.l, u.r, u.pa, rng, pre["alpha_arr"], pre["beta_arr"], pre["L"], **meta["params"])
        else:
            res = so.fit_utr(u.x, u.l, u.r, u.pa, rng, **meta["params"])
        K = int(fx["K"][j])
        assert res.K == K and res.L == fx["L"][j] and len(res.lb_arr) == fx["n_iter"][j]
        assert np.array_equal(res.alpha_arr, fx["alpha"][j, :K]) and np.array_equal(res.beta_arr, fx["beta"][j, :K])
        assert np.array_equal(res.ws, fx["ws"][j, :K + 1]) and res.bic == fx["bic"][j] and res.lb_arr[-1] == fx["lb_last"][j]
        lab = fx["labels"][fx["label_off"][j]:fx["label_off"][j + 1]]
        assert np.array_equal(np.asarray(res.label_arr), lab)
        if j + 1 < len(fx["index"]) and fx["file_id"][j + 1] == fx["file_id"][j]:
            # the stream ends where the fixture says the next UTR of the file starts
            st, want = rng.get_state(), _rng_at(fx["rng_off"][j + 1]).get_state()
            assert np.array_equal(st[1], want[1]) and st[2] == want[2]
