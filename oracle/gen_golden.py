"""Generate tests/golden/* by running the UNMODIFIED reference (oracle/ref_harness.py).

TEST INFRASTRUCTURE.  Run in the build container only (`python -m oracle.gen_golden`); the outputs
are committed so that the travelling tests can pin `oracle/scape_oracle.py` and the CUDA path
without `/root/reference`.

Fixtures written:
  tests/golden/example_inputs.npz    read columns of the 4 shipped example UTRs (x, l, r, pa, cb_id)
                                     + K / alpha / beta / labels of the shipped result pickles
                                     (loose goldens from an older RNG stream)
  tests/golden/reference_results.json  per case: parameters, inputs (by name / synthetic index) and
                                     the reference's Parameters fields (floats as repr -> exact)
  tests/golden/reference_labels.npz  label_arr per (case, utr)

Cases (every one is `np.random.seed(1)` once per chunk, like `_infer_pa`, apa_core.py:125):
  toy, chr17, chr19    shipped chunks, default TOML            (cfg-1 of BASELINE.json)
  synth8               synthetic UTRs 0..7 x 300 reads          (serial RNG stream with prunes)
  synth_rerun          UTRs 0, 22, 31 with n_max_apa=3          (re-run loop, apa_core.py:1023-1030)
  synth_fixed          UTRs 3, 6, 15 with pre_para = result of UTR 3   (fixed_run, apa_core.py:883-928)
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import pickle
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from scape_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
EXAMPLES = {
    "toy": "/root/reference/examples/toy-example/pkl_{}/example.100.1.1.{}.pkl",
    "chr17": "/root/reference/examples/SCZ-nowa-scape/pkl_{}/chr17_merge.100.1.1.{}.pkl",
    "chr19": "/root/reference/examples/SCZ-nowa-scape/pkl_{}/chr19_merge.100.1.1.{}.pkl",
}
SYNTH_CASES = {
    "synth8": dict(utrs=list(range(8)), reads=300, params={}),
    "synth_rerun": dict(utrs=[0, 22, 31], reads=300, params={"n_max_apa": 3}),
    "synth_fixed": dict(utrs=[3, 6, 15], reads=300, params={}, fixed_from=3),
}


def _load_stream(path):
    out = []
    with open(path, "rb") as fh:
        while True:
            try:
                out.append(pickle.load(fh))
            except EOFError:
                return out


def _record(res):
    return dict(title=res.title, K=int(res.K), L=int(res.L), alpha_arr=[int(a) for a in res.alpha_arr],
                beta_arr=[float(b) for b in res.beta_arr], ws=[repr(float(w)) for w in res.ws],
                bic=repr(float(res.bic)), lb_arr=[repr(float(v)) for v in res.lb_arr],
                gene_info_str=res.gene_info_str, n_reads=int(len(res.label_arr)))


def _run_reference(ref, chunk, params, pre_para_file=None):
    """The body of reference `infer` (apa_core.py:1104-1137) on an in-memory chunk."""
    kw = {"n_max_apa": 5}
    kw.update(params)
    if pre_para_file:
        kw.update(fixed_run_mode=True, pre_para_pkl_file=pre_para_file)
    np.random.seed(1)
    out = []
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for gi, df in chunk:
            out.append(ref.subsample_run(data=df, gene_info_str=gi, **kw))
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_harness.load_reference_apa_core()
    sys.modules.setdefault("scape", sys.modules["scape_ref"])          # shipped pickles name scape.apa_core
    sys.modules.setdefault("scape.apa_core", ref)
    results, labels, inputs = {}, {}, {}
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only:   # refresh selected cases, keep the rest of the committed fixtures
        with open(os.path.join(GOLD, "reference_results.json")) as fh:
            results = json.load(fh)["cases"]
        labels = dict(np.load(os.path.join(GOLD, "reference_labels.npz")))
        inputs = dict(np.load(os.path.join(GOLD, "example_inputs.npz")))
        for name in only:
            for k in [k for k in labels if k.startswith(name + "/")]:
                del labels[k]

    for name, pat in EXAMPLES.items():
        if only and name not in only:
            continue
        chunk = _load_stream(pat.format("input", "input"))
        shipped = _load_stream(pat.format("output", "res"))
        t0 = time.time()
        res = _run_reference(ref, chunk, {})
        print(f"{name}: {len(chunk)} UTRs, reference took {time.time() - t0:.1f}s", flush=True)
        results[name] = dict(params={}, source="example", utrs=[_record(r) for r in res])
        for i, ((gi, df), r, old) in enumerate(zip(chunk, res, shipped)):
            assert np.array_equal(df["read_id"], np.arange(len(df)))
            key = f"{name}/{i}"
            inputs[key + "/x"] = np.asarray(df["x"], dtype=np.int32)
            inputs[key + "/l"] = np.asarray(df["l"], dtype=np.int32)
            inputs[key + "/r"] = np.asarray(df["r"], dtype=np.float32)
            inputs[key + "/pa"] = np.asarray(df["pa"], dtype=np.float32)
            inputs[key + "/cb_id"] = np.asarray(df["cb_id"], dtype=np.int32)
            assert np.array_equal(inputs[key + "/pa"].astype(float), np.asarray(df["pa"]), equal_nan=True)
            inputs[key + "/shipped_alpha"] = np.asarray(old.alpha_arr, dtype=np.int64)
            inputs[key + "/shipped_beta"] = np.asarray(old.beta_arr, dtype=np.float64)
            inputs[key + "/shipped_label"] = np.asarray(old.label_arr, dtype=np.int8)
            labels[key] = np.asarray(r.label_arr, dtype=np.int8)

    for name, spec in SYNTH_CASES.items():
        if only and name not in only:
            continue
        utrs = [synth.make_utr(u, spec["reads"]) for u in spec["utrs"]]
        chunk = [(u.gene_info_str, synth.to_dataframe(u)) for u in utrs]
        pre_file = None
        extra = {}
        if "fixed_from" in spec:
            src = synth.make_utr(spec["fixed_from"], spec["reads"])
            pre = _run_reference(ref, [(src.gene_info_str, synth.to_dataframe(src))], spec["params"])[0]
            pre_file = "/tmp/_gen_golden_pre_para.pkl"
            with open(pre_file, "wb") as fh:
                pickle.dump(pre, fh)
            extra["pre_para"] = dict(alpha_arr=[int(a) for a in pre.alpha_arr],
                                     beta_arr=[float(b) for b in pre.beta_arr], L=int(pre.L), K=int(pre.K))
        t0 = time.time()
        res = _run_reference(ref, chunk, spec["params"], pre_file)
        print(f"{name}: {len(chunk)} UTRs, reference took {time.time() - t0:.1f}s", flush=True)
        results[name] = dict(params=spec["params"], source="synth", synth_utrs=spec["utrs"],
                             synth_reads=spec["reads"], utrs=[_record(r) for r in res], **extra)
        for i, r in enumerate(res):
            labels[f"{name}/{i}"] = np.asarray(r.label_arr, dtype=np.int8)

    with open(os.path.join(GOLD, "reference_results.json"), "w") as fh:
        json.dump(dict(generator="oracle/gen_golden.py", reference="chengl7-lab/scape src/scape/apa_core.py (unmodified), "
                       "taichi_core replaced by oracle.scape_oracle FP64 stand-in",
                       numpy=np.__version__, cases=results), fh, indent=1)
    np.savez_compressed(os.path.join(GOLD, "reference_labels.npz"), **labels)
    np.savez_compressed(os.path.join(GOLD, "example_inputs.npz"), **inputs)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
