"""The C-ABI boundary without a GPU: the library loads, exports exactly what include/scape_b200.h
declares, ctypes structs match the C layout, and the product path fails loudly without CUDA."""
import ctypes
import os
import pickle
import re
import subprocess

import numpy as np
import pytest

from scape_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scape_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scape_b200_[a-z0-9_]+)\s*\(", src)) - {"scape_b200_argsort_fn"})


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for n in names:
        assert re.search(rf"\bT {n}\b", nm), n


def test_struct_layouts_match_the_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof / offsetof with ctypes."""
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "scape_b200.h"\n'
        "int main(){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(scape_b200_params), sizeof(scape_b200_batch),"
        " sizeof(scape_b200_results), sizeof(scape_b200_timing), sizeof(scape_b200_chain_io),"
        " offsetof(scape_b200_params, betas), offsetof(scape_b200_params, smooth_w),"
        " offsetof(scape_b200_chain_io, bic), offsetof(scape_b200_batch, stream_state));return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(_lib.Params), ctypes.sizeof(_lib.Batch), ctypes.sizeof(_lib.Results),
            ctypes.sizeof(_lib.Timing), ctypes.sizeof(_lib.ChainIO), _lib.Params.betas.offset,
            _lib.Params.smooth_w.offset, _lib.ChainIO.bic.offset, _lib.Batch.stream_state.offset]
    assert got == want


def test_params_tables_are_the_reference_expressions():
    p = _lib.make_params()
    assert list(p.betas[:p.n_beta]) == list(np.arange(5, 70, 5) + 0.0)
    s = np.arange(20, 150, 10)
    pmf = np.repeat(1 / 13, 13)
    pmf = pmf / sum(pmf)
    assert list(p.s_dis[:p.n_s]) == list(s) and list(p.pmf_s[:p.n_s]) == list(pmf)
    w = np.exp(-np.arange(-45, 46) ** 2 / (2 * 15 * 15))
    assert p.n_smooth == 91 and list(p.smooth_w[:91]) == list(w)


def test_parameter_errors_match_the_reference():
    with pytest.raises(Exception, match="n_max_apa has to be greater than n_min_apa"):
        _lib.make_params(n_max_apa=2, n_min_apa=3)          # apa_core.py:931-933
    with pytest.raises(Exception, match="max_beta has to be greater than beta_step_size"):
        _lib.make_params(max_beta=3, beta_step=5)           # apa_core.py:935-937


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to exist (and nothing routes to the oracle)."""
    lib = _lib.load()
    if lib.scape_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.ScapeB200Error, match="no CUDA device"):
        _lib.Engine(_lib.make_params())
    import scape_b200.apa_core as host
    src = open(host.__file__).read() + open(_lib.__file__).read()
    assert "oracle" not in src.replace("no CPU / Taichi fallback", "")


def test_result_pickles_keep_the_reference_class_path():
    from scape.apa_core import Parameters
    p = Parameters(title="Final Result", alpha_arr=np.array([100, 200]), beta_arr=np.array([5., 10.]),
                   ws=np.array([.5, .4, .1]), L=2000, cb_id_arr=np.arange(3), readID_arr=np.arange(3))
    blob = pickle.dumps(p)
    assert b"scape.apa_core" in blob and b"Parameters" in blob and b"scape_b200" not in blob
    q = pickle.loads(blob)
    assert q.K == 2 and "K=2 L=2000" in str(q)


def test_cli_contract(tmp_path):
    """Flag names and failure modes of `scape infer_pa` (apa_core.py:40-104, 120-132)."""
    from click.testing import CliRunner
    from scape.cli import cli
    out = tmp_path / "out"
    out.mkdir()
    r = CliRunner().invoke(cli, ["infer_pa", "--pkl_input_file", "x.input.pkl", "--output_dir", str(out)])
    assert isinstance(r.exception, AssertionError)            # parameters.toml must exist (:85)
    (out / "parameters.toml").write_text("n_max_apa = 5\n")
    r = CliRunner().invoke(cli, ["infer_pa", "--pkl_input_file", str(tmp_path / "nope.input.pkl"), "--output_dir", str(out)])
    assert "Given input file does not exists" in str(r.exception)
    tmp_in = tmp_path / "a.100.tmp.1.input.pkl"
    tmp_in.write_bytes(b"")
    r = CliRunner().invoke(cli, ["infer_pa", "--pkl_input_file", str(tmp_in), "--output_dir", str(out)])
    assert "is incomplete" in str(r.exception)
    r = CliRunner().invoke(cli, ["infer_pa", "--help"])
    for flag in ("--pkl_input_file", "--output_dir", "--toml_para_file", "--pre_para_pkl_file"):
        assert flag in r.output


def test_chunk_batch_packs_prepare_input_frames_exactly():
    """ChunkBatch.add takes the four read columns and the two id columns out of a prepare_input frame
    through one ndarray conversion; values and dtypes must equal the column-by-column access, and frames
    it cannot treat that way (ids that are not integers) must take the slow path and stay exact."""
    import numpy as np
    from scape_b200 import apa_core, synth
    batch = apa_core.ChunkBatch()
    frames = [synth.to_dataframe(synth.make_utr(300 + i, 50 + 40 * i)) for i in range(6)]
    frames[2] = frames[2][["read_id", "cb_id", "pa", "r", "l", "x", "junction", "seg1_en", "seg2_en"]]   # other column order
    odd = frames[4].copy()
    odd["cb_id"] = odd["cb_id"].astype(float)
    odd.loc[3, "cb_id"] = np.nan
    frames[4] = odd
    for i, df in enumerate(frames):
        batch.add(f"g{i}", df, i % 2)
    off, x, l, r, pa, sid = batch.packed()
    assert list(off) == list(np.cumsum([0] + [len(df) for df in frames]))
    assert list(sid) == [0, 1, 0, 1, 0, 1]
    for i, df in enumerate(frames):
        sl = slice(off[i], off[i + 1])
        assert np.array_equal(x[sl], np.asarray(df["x"], dtype=np.float64))
        assert np.array_equal(l[sl], np.asarray(df["l"], dtype=np.float64))
        assert np.array_equal(r[sl], np.asarray(df["r"], dtype=np.float64), equal_nan=True)
        assert np.array_equal(pa[sl], np.asarray(df["pa"], dtype=np.float64), equal_nan=True)
        cb, rid = batch.frames[i]
        assert cb.dtype == df["cb_id"].dtype and np.array_equal(cb, np.array(df["cb_id"]), equal_nan=True)
        assert rid.dtype == df["read_id"].dtype and np.array_equal(rid, np.array(df["read_id"]))
