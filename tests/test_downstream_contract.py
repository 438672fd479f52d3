"""Consumer contract (SURVEY.md section 8c gate iv): the reference's own `merge_pa`, `cal_exp_pa_len`
and `ex_pa_cnt_mat` must run UNTOUCHED on the result pickles this repository writes.

What is under test is the product's result assembly (`scape_b200.apa_core.results_to_parameters`:
attribute names, dtypes, class path, per-read arrays in input order) and the on-disk layout, not the
kernels: the per-UTR numbers come from the CPU oracle here so the test runs without a GPU.  The
downstream stages are the unmodified reference files, loaded in place from /root/reference (this
build container only -- the test skips where the tree is not mounted)."""
import importlib
import os
import pickle
import sys
import types

import numpy as np
import pytest

from oracle import ref_harness
from oracle import scape_oracle as so
from scape_b200 import _lib, synth
from scape_b200.apa_core import ChunkBatch, results_to_parameters

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted")


def _reference_modules():
    ref_harness.load_reference_apa_core()                       # registers the scape_ref package
    for name in ("pybedtools", "gffutils"):                     # imported by utils.py for the annotation stage only
        sys.modules.setdefault(name, types.ModuleType(name))
    jh = importlib.import_module("scape_ref.junction_handler")
    ut = importlib.import_module("scape_ref.utils")
    return jh, ut


def _oracle_fit_output(utrs):
    """FitOutput (the C ABI's result arrays) filled from the oracle, one RNG stream for the file."""
    out = _lib.FitOutput(len(utrs), sum(u.n_reads for u in utrs))
    rng = np.random.RandomState(1)
    pos = 0
    for i, u in enumerate(utrs):
        w = so.fit_utr(u.x, u.l, u.r, u.pa, rng)
        out.K[i] = w.K
        out.L[i] = w.L
        out.alpha[i, :w.K] = w.alpha_arr
        out.beta[i, :w.K] = w.beta_arr
        out.ws[i, :w.K + 1] = w.ws
        out.bic[i] = w.bic
        out.n_lb[i] = len(w.lb_arr)
        out.lb_arr[i, :len(w.lb_arr)] = w.lb_arr
        out.label[pos:pos + u.n_reads] = w.label_arr
        pos += u.n_reads
    return out


def test_reference_downstream_stages_run_on_our_pickles(tmp_path):
    jh, ut = _reference_modules()
    out_dir = str(tmp_path)
    utrs = [synth.make_utr(40 + i, 180 + 40 * i) for i in range(6)]
    paths = synth.write_chunk_files(utrs, out_dir, per_file=3, stem="contract")
    os.makedirs(os.path.join(out_dir, "pkl_output"))
    assigned = {}                                                # gene -> reads our result gives to a pA site
    for f, path in enumerate(paths):
        mine = utrs[3 * f:3 * f + 3]
        batch = ChunkBatch()
        for u in mine:
            batch.add(u.gene_info_str, synth.to_dataframe(u), f)
        paras = results_to_parameters(batch, _oracle_fit_output(mine), fixed_run_mode=False)
        name = os.path.basename(path)[:-10]                      # apa_core.py:127
        with open(os.path.join(out_dir, "pkl_output", name + ".res.pkl"), "wb") as fh:
            for p in paras:
                assert type(p).__module__ == "scape.apa_core" and type(p).__name__ == "Parameters"
                assert p.alpha_arr.dtype.kind == "i" and p.label_arr.dtype == np.int64
                pickle.dump(p, fh)
                assigned[p.gene_info_str.split(":")[1]] = int(np.sum(p.label_arr < p.K))
    with open(os.path.join(out_dir, "barcode_index.csv"), "w") as fh:
        fh.write("CB,index\n")
        for i in range(10000):
            fh.write(f"CB{i:05d}-1,{i}\n")

    jh._merge_pa(out_dir, utr_merge=True)                        # junction_handler.py:44
    merged = os.path.join(out_dir, "res.gene.pkl")
    assert os.path.exists(merged)
    genes = []
    with open(merged, "rb") as fh:
        while True:
            try:
                genes.append(pickle.load(fh))
            except EOFError:
                break
    assert len(genes) == len(utrs)
    by_gene = {g.gene_info_str.split(":")[1]: g for g in genes}
    for u in utrs:
        gene = u.gene_info_str.split(":")[1]
        g = by_gene[gene]
        # merge_pa keeps the reads assigned to a pA site and drops the uniform component
        assert g.K >= 1 and len(g.label_arr) == assigned[gene] == len(g.cb_id_arr)

    ut.cal_exp_pa_len.callback(out_dir, "None", "res.gene.pkl")  # utils.py:339
    csv_path = os.path.join(out_dir, "all_cell.gene.pa.len.csv")
    assert os.path.exists(csv_path)
    import pandas as pd
    df = pd.read_csv(csv_path)
    assert len(df) == len(utrs) and set(df.columns) == {"gene_id", "exp_length", "num_pa"}

    ut.ex_pa_cnt_mat.callback(out_dir, "res.gene.pkl")           # utils.py:451
    cnt_path = os.path.join(out_dir, "res.gene.cnt.tsv.gz")
    assert os.path.exists(cnt_path)
    cnt = pd.read_csv(cnt_path, index_col=0)                     # comma-separated, every field quoted (utils.py:542)
    assert cnt.shape[0] == sum(g.K for g in genes)               # one row per pA site
    assert int(cnt.to_numpy().sum()) == sum(int(np.sum(g.label_arr < g.K)) for g in genes)
