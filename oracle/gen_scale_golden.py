"""Generate tests/golden/scale/* : the at-scale parity fixtures (north-star acceptance gate).

TEST INFRASTRUCTURE.  Run in the build container (`python -m oracle.gen_scale_golden [suite ...]`);
it fits whole synthetic chunk files with `oracle/scape_oracle.py` (pinned bit-for-bit against the
unmodified reference by tests/test_oracle_vs_reference.py) under the reference's seed policy --
`np.random.seed(1)` once per chunk file, UTRs fitted serially on that one stream
(/root/reference/src/scape/apa_core.py:125, 1104-1137) -- one process per file, and freezes compact
per-UTR records.  The inputs are not stored: `scape_b200.synth.make_utr(index, reads, long_utr)` is
deterministic, the fixture keeps (index, reads, long_utr) per UTR.

Suites (BASELINE.json configs):
  cfg2     10 chunk files x 100 UTRs x 500 reads, UTR indices 0..999 (the first 10 files of bench.py's
           cfg-2 set), n_max_apa 5
  cfg3     stratified sample of the 20k heavy-tailed set: 30 UTRs per reads-per-UTR decile + 6 UTRs
           with >= 100k reads + 6 more of the long-UTR class (top 1 % by reads), dealt to 12 files
  cfg4     `--pre_para_pkl_file` mode (fixed_run, apa_core.py:883-928): the first 200 UTRs of the
           heavy-tailed set in 4 files; files 0-1 use a fixed 3-site pre_para, files 2-3 the
           normal-mode result of their own first UTR
  kmax8    2 files x 50 UTRs x 500 reads (indices 2000..2099), n_max_apa 8
  kmax10   2 files x 50 UTRs x 500 reads (indices 2100..2199), n_max_apa 10
  cfg5     the stress corner: one UTR with 1,000,000 reads and one with 300,000, n_max_apa 10

Per UTR: K, alpha, beta, ws, bic, lb_arr[-1], n_iter, n_frag, n_theta, chains_run, the (k_max,
k_selected, K) path of every sweep, the per-read labels (int8), and `rng_off` = number of 32-bit
MT19937 outputs the file's stream had produced when this UTR started, so a test can also start
every UTR from the reference's exact RNG state (RandomState(1).bytes(4 * rng_off)) and separate a
UTR's own divergence from the cascade a divergence causes in the rest of its file (apa_core.py:843,
1023-1030 consume result-dependent draws).
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import pickle
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import scape_oracle as so  # noqa: E402
from scape_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "scale")
CACHE = os.environ.get("SCALE_CACHE", "/tmp/work/scale_cache")
KCAP = 15
FIXED_PRE = dict(alpha_arr=[500, 900, 1400], beta_arr=[20.0, 35.0, 30.0], L=21000)   # bench.py's cfg-4 pre_para


def suites():
    """suite -> dict(params, files=[[(index, reads, long_utr), ...], ...], pre=[None | dict | "first"])."""
    out = {}
    out["cfg2"] = dict(params={}, files=[[(f * 100 + i, 500, False) for i in range(100)] for f in range(10)])
    counts = synth.heavy_tail_read_counts(20000)
    cut = np.quantile(counts, 0.99)
    order = np.argsort(counts, kind="stable")
    pick = []
    for d in range(10):
        dec = np.sort(order[d * 2000:(d + 1) * 2000])
        pick += [int(i) for i in dec[:30]]
    giants = [int(i) for i in np.sort(np.nonzero(counts >= 100000)[0])[:6]]
    longs = [int(i) for i in np.sort(np.nonzero((counts >= cut) & (counts < 100000))[0])[:6]]
    pick = sorted(set(pick) - set(giants) - set(longs))
    files = [[] for _ in range(12)]
    for j, i in enumerate(pick):
        files[j % 12].append(i)
    for j, i in enumerate(giants):           # one giant per file, a few UTRs into the stream
        files[j].insert(5, i)
    for j, i in enumerate(longs):
        files[6 + j].insert(5, i)
    out["cfg3"] = dict(params={}, files=[[(i, int(counts[i]), bool(counts[i] >= cut)) for i in f] for f in files])
    out["cfg4"] = dict(params={}, files=[[(i, int(counts[i]), bool(counts[i] >= cut)) for i in range(f * 50, f * 50 + 50)]
                                         for f in range(4)], pre=[FIXED_PRE, FIXED_PRE, "first", "first"])
    out["kmax8"] = dict(params={"n_max_apa": 8}, files=[[(2000 + f * 50 + i, 500, False) for i in range(50)] for f in range(2)])
    # cfg-5 corner: one UTR with a million reads (N ~ 11,500 fragments) and one with 300k, n_max_apa 10
    out["cfg5"] = dict(params={"n_max_apa": 10}, files=[[(900000, 1000000, True)], [(900001, 300000, True)]])
    out["kmax10"] = dict(params={"n_max_apa": 10}, files=[[(2100 + f * 50 + i, 500, False) for i in range(50)] for f in range(2)])
    return out


def _advance_count(s1, s2):
    """Number of 32-bit MT19937 outputs between two legacy RandomState states (s2 after s1)."""
    k1, p1, k2, p2 = s1[1], s1[2], s2[1], s2[2]
    if np.array_equal(k1, k2):
        assert p2 >= p1
        return p2 - p1
    g = np.random.RandomState()
    g.set_state(s1)
    n = 624 - p1
    if n > 0:                                # (bytes(0) draws one output: C division in mtrand)
        g.bytes(4 * n)                       # pos -> 624, key unchanged
    while True:
        g.bytes(4 * 624)                     # one twist, pos -> 624 again
        st = g.get_state()
        if np.array_equal(st[1], k2):
            return n + p2
        n += 624
        if n > 1 << 34:
            raise RuntimeError("RNG states are not on one stream")


def run_file(job):
    suite, f, utrs, params, pre = job
    cache = os.path.join(CACHE, f"{suite}_{f}.pkl")
    if os.path.exists(cache):
        with open(cache, "rb") as fh:
            return pickle.load(fh)
    warnings.simplefilter("ignore")
    rng = np.random.RandomState(1)                     # np.random.seed(1) per file (apa_core.py:125)
    t0 = time.perf_counter()
    recs, off, pre_used = [], 0, None
    if pre == "first":
        # --pre_para_pkl_file: the FIRST object of the given pickle (apa_core.py:1002-1003); here the
        # normal-mode result of this file's first UTR on its own fresh stream
        i, reads, long_utr = utrs[0]
        u = synth.make_utr(i, reads, long_utr=long_utr)
        r0 = so.fit_utr(u.x, u.l, u.r, u.pa, np.random.RandomState(1), **params)
        pre_used = dict(alpha_arr=[int(a) for a in r0.alpha_arr], beta_arr=[float(b) for b in r0.beta_arr], L=int(r0.L))
    elif pre is not None:
        pre_used = pre
    for i, reads, long_utr in utrs:
        u = synth.make_utr(i, reads, long_utr=long_utr)
        s1 = rng.get_state()
        t1 = time.perf_counter()
        try:
            if pre_used is not None:
                res = so.fit_utr_fixed(u.x, u.l, u.r, u.pa, rng, pre_used["alpha_arr"], pre_used["beta_arr"], pre_used["L"], **params)
            else:
                res = so.fit_utr(u.x, u.l, u.r, u.pa, rng, **params)
            err = ""
        except Exception as e:                         # the reference would abort the file here
            res, err = None, f"{type(e).__name__}: {e}"
        dt = time.perf_counter() - t1
        rec = dict(index=i, reads=reads, long_utr=long_utr, rng_off=off, err=err, secs=dt)
        if res is not None:
            rec.update(K=int(res.K), L=int(res.L), alpha=[int(a) for a in res.alpha_arr],
                       beta=[float(b) for b in res.beta_arr], ws=[float(w) for w in res.ws], bic=float(res.bic),
                       lb_last=float(res.lb_arr[-1]), n_iter=len(res.lb_arr), n_frag=int(res.n_frag),
                       n_theta=int(res.n_theta), chains_run=int(res.chains_run),
                       path=[[int(v) for v in p] for p in res.path], labels=np.asarray(res.label_arr, dtype=np.int8))
        recs.append(rec)
        off += _advance_count(s1, rng.get_state())
        if err:
            break
    out = dict(suite=suite, file=f, recs=recs, pre=pre_used, secs=time.perf_counter() - t0)
    os.makedirs(CACHE, exist_ok=True)
    with open(cache + ".tmp", "wb") as fh:
        pickle.dump(out, fh)
    os.replace(cache + ".tmp", cache)
    return out


def write_suite(name, spec, results):
    results = sorted(results, key=lambda r: r["file"])
    n = sum(len(r["recs"]) for r in results)
    a = dict(file_id=np.zeros(n, np.int32), index=np.zeros(n, np.int32), reads=np.zeros(n, np.int64),
             long_utr=np.zeros(n, np.int8), rng_off=np.zeros(n, np.int64), K=np.full(n, -1, np.int32),
             L=np.zeros(n, np.int64), alpha=np.zeros((n, KCAP), np.int32), beta=np.zeros((n, KCAP)),
             ws=np.zeros((n, KCAP + 1)), bic=np.zeros(n), lb_last=np.zeros(n), n_iter=np.zeros(n, np.int32),
             n_frag=np.zeros(n, np.int32), n_theta=np.zeros(n, np.int32), chains_run=np.zeros(n, np.int32),
             path=np.zeros((n, 8, 3), np.int32), n_path=np.zeros(n, np.int32), label_off=np.zeros(n + 1, np.int64),
             oracle_secs=np.zeros(n, np.float32))
    labels, errs, j = [], {}, 0
    for r in results:
        for rec in r["recs"]:
            a["file_id"][j], a["index"][j], a["reads"][j], a["long_utr"][j] = r["file"], rec["index"], rec["reads"], rec["long_utr"]
            a["rng_off"][j], a["oracle_secs"][j] = rec["rng_off"], rec["secs"]
            if rec["err"]:
                errs[str(j)] = rec["err"]
                a["label_off"][j + 1] = a["label_off"][j]
            else:
                K = rec["K"]
                a["K"][j], a["L"][j] = K, rec["L"]
                a["alpha"][j, :K], a["beta"][j, :K], a["ws"][j, :K + 1] = rec["alpha"], rec["beta"], rec["ws"]
                a["bic"][j], a["lb_last"][j], a["n_iter"][j] = rec["bic"], rec["lb_last"], rec["n_iter"]
                a["n_frag"][j], a["n_theta"][j], a["chains_run"][j] = rec["n_frag"], rec["n_theta"], rec["chains_run"]
                p = rec["path"][:8]
                a["n_path"][j] = len(p)
                if p:
                    a["path"][j, :len(p)] = p
                labels.append(rec["labels"])
                a["label_off"][j + 1] = a["label_off"][j] + len(rec["labels"])
            j += 1
    a["labels"] = np.concatenate(labels) if labels else np.zeros(0, np.int8)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **a)
    meta = dict(params=spec["params"], n_files=len(spec["files"]), n_utr=n, errors=errs,
                pre=[r["pre"] for r in results], oracle_cpu_seconds=float(sum(r["secs"] for r in results)),
                generator="python -m oracle.gen_scale_golden " + name)
    with open(os.path.join(OUT, name + ".json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print(f"{name}: {n} UTRs, {len(errs)} errors, oracle {meta['oracle_cpu_seconds']:.0f} core-s -> {OUT}/{name}.npz", flush=True)


def main(argv):
    specs = suites()
    names = argv or list(specs)
    procs = int(os.environ.get("SCALE_PROCS", os.cpu_count() or 1))
    jobs = []
    for nm in names:
        s = specs[nm]
        pre = s.get("pre") or [None] * len(s["files"])
        for f, utrs in enumerate(s["files"]):
            jobs.append((nm, f, utrs, s["params"], pre[f]))
    # most expensive files first (cost ~ total reads, giants dominate)
    jobs.sort(key=lambda j: -sum(min(r, 20000) + (3e5 if r >= 100000 else 0) for _, r, _ in j[2]))
    done = {nm: [] for nm in names}
    with mp.get_context("fork").Pool(procs) as pool:
        for res in pool.imap_unordered(run_file, jobs, chunksize=1):
            done[res["suite"]].append(res)
            print(f"  {res['suite']} file {res['file']}: {len(res['recs'])} UTRs in {res['secs']:.0f} s", flush=True)
            if len(done[res["suite"]]) == len(specs[res["suite"]]["files"]):
                write_suite(res["suite"], specs[res["suite"]], done[res["suite"]])


if __name__ == "__main__":
    main(sys.argv[1:])
