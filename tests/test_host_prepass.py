"""libscape_b200's host pre-pass (no GPU): binning, coverage profile, peaks and the numpy-legacy
RNG replay must be bit-identical to numpy / scipy / the oracle, because every later result depends
on the initial draws (hard part #1)."""
import os

import numpy as np
import pytest

from oracle import scape_oracle as so
from scape_b200 import _lib, synth


@pytest.mark.parametrize("seed", [1, 7, 2**31 + 5])
def test_mt19937_streams(seed):
    rs = np.random.RandomState(seed)
    assert np.array_equal(rs.random_sample(1500), _lib.rng_draw(seed, 0, 0, 1500))
    rs = np.random.RandomState(seed)
    assert np.array_equal(rs.randint(0, 13, size=700), _lib.rng_draw(seed, 1, 13, 700))
    # the shuffle is the hot loop of the chain initialisation and is written without the rejection
    # branch (np_rng.hpp): every mask boundary, the 624-word refill boundary, and the stream after it
    for n in (1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 31, 32, 33, 623, 624, 625, 1000, 2000, 4370, 70000):
        rs = np.random.RandomState(seed)
        assert np.array_equal(rs.permutation(n), _lib.rng_draw(seed, 2, n, n))
        assert np.array_equal(rs.random_sample(40), _lib.rng_draw(seed, 3, n, 40))


def _oracle_model(u, **kw):
    prm = dict(so.DEFAULTS)
    prm.update(kw)
    prm["utr_length"] = so.resolve_utr_length(u.x, u.l, prm)
    m = so.build_model(u.x, u.l, u.r, u.pa, prm)
    m.prof_x, m.prof_y = so.coverage_profile(m)
    m.peak_idx, m.peak_w = so.find_profile_peaks(m)
    return m


@pytest.mark.parametrize("ui,reads", [(0, 300), (1, 3000), (5, 40), (39, 300), (17, 12), (22, 20000)])
def test_binning_profile_peaks(ui, reads):
    u = synth.make_utr(ui, reads)
    want = so.bin_reads(u.x, u.l, u.r, u.pa)
    got = _lib.bin_reads(u.x, u.l, u.r, u.pa)
    for i in range(4):
        assert np.array_equal(want[i], got[i], equal_nan=True)
    assert np.array_equal(want[4], got[4]) and np.array_equal(want[5], got[5])
    m = _oracle_model(u)
    pr = _lib.profile(_lib.make_params(), u.x, u.l, u.r, u.pa)
    assert pr["L"] == m.L
    assert np.array_equal(pr["theta"], m.theta)
    assert np.array_equal(pr["prof_y"], m.prof_y)          # numpy pairwise-sum order reproduced
    assert np.array_equal(pr["peak_idx"], m.peak_idx)      # incl. np.argsort tie order (UTR 39)
    assert np.array_equal(pr["peak_w"], m.peak_w)


@pytest.mark.parametrize("beta_step", [2, 3, 10])
def test_smoothing_window_sizes(beta_step):
    """ker_smooth's window is 6 * 3 * beta_step + 1 taps: 37 and 55 go through the four-positions-at-a-time
    path (numpy pairwise order for 8 <= n <= 128), 181 through the recursive pairwise sum; profile, peaks
    and peak weights must stay bit-identical to numpy / scipy either way."""
    for ui, reads in ((0, 300), (3, 2000)):
        u = synth.make_utr(ui, reads)
        m = _oracle_model(u, beta_step=beta_step, max_beta=70)
        pr = _lib.profile(_lib.make_params(beta_step=beta_step, max_beta=70), u.x, u.l, u.r, u.pa)
        assert np.array_equal(pr["prof_y"], m.prof_y)
        assert np.array_equal(pr["peak_idx"], m.peak_idx)
        assert np.array_equal(pr["peak_w"], m.peak_w)


def test_smoothing_without_avx2_is_the_same():
    """The smoothing kernel is compiled for AVX2 and for baseline x86-64 and picked at run time; the
    baseline build must give the same bits (own process: the choice is made once)."""
    import hashlib
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, hashlib, numpy as np; sys.path.insert(0, %r)\n"
            "from scape_b200 import _lib, synth\n"
            "h = hashlib.sha256()\n"
            "for ui, reads in ((0, 300), (1, 3000), (22, 20000)):\n"
            "    u = synth.make_utr(ui, reads)\n"
            "    pr = _lib.profile(_lib.make_params(), u.x, u.l, u.r, u.pa)\n"
            "    h.update(pr['prof_y'].tobytes()); h.update(pr['peak_w'].tobytes())\n"
            "print('DIGEST', h.hexdigest())\n") % root
    outs = []
    for extra in ({}, {"SCAPE_B200_NO_AVX2": "1"}):
        env = dict(os.environ)
        env.update(extra)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-1500:]
        outs.append([l for l in r.stdout.splitlines() if l.startswith("DIGEST")][-1])
    assert outs[0] == outs[1]


def test_binning_with_polya_lengths_and_nan_columns():
    rng = np.random.default_rng(3)
    n = 400
    x = rng.integers(0, 1500, n).astype(float)
    l = rng.integers(20, 133, n).astype(float)
    r = np.where(rng.random(n) < 0.5, rng.integers(5, 120, n), np.nan).astype(float)
    pa = np.where(rng.random(n) < 0.1, x + l - 1, np.nan)
    want = so.bin_reads(x, l, r, pa)
    got = _lib.bin_reads(x, l, r, pa)
    for i in range(4):
        assert np.array_equal(want[i], got[i], equal_nan=True)
    assert np.array_equal(want[5], got[5])


@pytest.mark.parametrize("ui,reads", [(0, 300), (3, 300), (12, 60), (22, 2000)])
def test_chain_initialisation_draws(ui, reads):
    """init_para + gen_k_arr for a whole K sweep, a prune refit, and a re-run block, back to back
    on one stream (apa_core.py:817-829, 720, 708-711)."""
    u = synth.make_utr(ui, reads)
    m = _oracle_model(u)
    ks = [5] * 10 + [4] * 10 + [3] * 10 + [2] * 10 + [1] * 10 + [-3] + [7] * 3 + [6] * 2 + [12]
    got = _lib.draw_chains(_lib.make_params(), u.x, u.l, u.r, u.pa, 1, ks)
    rng = np.random.RandomState(1)
    for i, k in enumerate(ks):
        if k < 0:
            K, w = -k, so.draw_weights(m, -k, rng)
        else:
            K = k
            ch = so.draw_chain(m, k, rng)
            w = ch.ws
            assert list(got[i].a_idx[:K]) == list(ch.a_idx)
            assert list(got[i].b_idx[:K]) == list(ch.b_idx)
        order = so.draw_component_order(K, 50, rng)
        assert list(got[i].ws[:K + 1]) == list(w)
        assert list(got[i].k_order[:50]) == list(order)


def test_chain_initialisation_draws_many_utrs():
    """A wider net for the arithmetic grid snap (start positions below the first / above the last theta,
    exact ties), the branch-free shuffles and the weighted choice: 40 UTRs of mixed size, a K sweep each,
    every alpha / beta index, weight and component order equal to the oracle's draw for draw."""
    P = _lib.make_params()
    ks = [5, 5, 5, 4, 4, 3, 3, 2, 1, 6, 8]
    for ui in range(200, 240):
        u = synth.make_utr(ui, [40, 150, 500, 3000][ui % 4], long_utr=(ui % 8 == 7))
        m = _oracle_model(u)
        got = _lib.draw_chains(P, u.x, u.l, u.r, u.pa, 1 + ui, ks)
        rng = np.random.RandomState(1 + ui)
        for i, k in enumerate(ks):
            ch = so.draw_chain(m, k, rng)
            order = so.draw_component_order(k, 50, rng)
            assert list(got[i].a_idx[:k]) == list(ch.a_idx), (ui, k)
            assert list(got[i].b_idx[:k]) == list(ch.b_idx)
            assert list(got[i].ws[:k + 1]) == list(ch.ws)
            assert list(got[i].k_order[:50]) == list(order)


def test_reference_assertion_is_reported():
    """assert 0 <= x < utr_length (apa_core.py:388) -> negative status instead of a silent fit."""
    x = np.array([-5.0, 10.0, 30.0]); l = np.array([50.0, 50.0, 50.0]); nan = np.full(3, np.nan)
    with pytest.raises(_lib.ScapeB200Error):
        _lib.profile(_lib.make_params(), x, l, nan, nan)


def test_chunk_batch_layout_cache_is_per_dtype():
    """Frames with the same columns but different id dtypes (int ids, then float ids carrying NaN) must
    each come back with THEIR dtype, like the reference's np.array(data[col]) (apa_core.py:1013-1014)."""
    import pandas as pd
    from scape_b200 import synth
    from scape_b200.apa_core import ChunkBatch
    u = synth.make_utr(5, 40)
    a = synth.to_dataframe(u)
    b = synth.to_dataframe(u)
    b["cb_id"] = b["cb_id"].astype(float)
    b.loc[3, "cb_id"] = np.nan
    batch = ChunkBatch()
    batch.add("g1", a, 0)
    batch.add("g2", b, 0)
    batch.add("g3", a, 0)
    assert batch.frames[0][0].dtype == np.int64 and np.array_equal(batch.frames[0][0], np.array(a["cb_id"]))
    assert batch.frames[1][0].dtype == np.float64 and np.isnan(batch.frames[1][0][3])
    assert np.array_equal(batch.frames[1][0][:3], np.array(b["cb_id"])[:3])
    assert batch.frames[2][0].dtype == np.int64


def test_chunk_batch_packed_into_scratch_equals_fresh_arrays():
    """`packed(scratch)` (pooled infer_files: columns concatenated into arrays kept on the engine) returns
    the same columns as `packed()`, also when a smaller batch reuses the scratch of a bigger one."""
    from scape_b200 import synth
    from scape_b200.apa_core import ChunkBatch
    scratch = {}
    for ids in ([5, 6, 7, 8], [9], []):
        batch = ChunkBatch()
        for k, i in enumerate(ids):
            u = synth.make_utr(i, 60 + 10 * k)
            batch.add(u.gene_info_str, synth.to_dataframe(u), k % 2)
        fresh = batch.packed()
        kept = batch.packed(scratch)
        assert len(fresh) == len(kept)
        for a, b in zip(fresh, kept):
            assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)
    assert len(scratch["cols"]) == 4


def test_chunk_batch_block_path_equals_column_by_column():
    """`ChunkBatch._columns_blocks` (columns straight from the block manager) returns what the reference's
    per-column reads return (np.array(data[col]), apa_core.py:1007-1014): plain frames, float ids with
    NaN, reordered columns, an extra non-numeric column, a float32 column."""
    import pandas as pd
    from scape_b200 import synth
    from scape_b200.apa_core import ChunkBatch
    u = synth.make_utr(7, 80)
    base = synth.to_dataframe(u)
    frames = [base]
    f = base.copy(); f["cb_id"] = f["cb_id"].astype(float); f.loc[2, "cb_id"] = np.nan; frames.append(f)
    frames.append(base[list(reversed(base.columns))])
    f = base.copy(); f["note"] = "a"; frames.append(f)
    f = base.copy(); f["r"] = f["r"].astype(np.float32); frames.append(f)
    frames.append(base.iloc[:0])
    import pickle
    frames = [pickle.loads(pickle.dumps(f)) for f in frames]      # like frames read from a chunk file
    for df in frames:
        got = ChunkBatch._columns_blocks(df)
        assert got is not None
        for name, col in zip(("x", "l", "r", "pa"), got[:4]):
            want = np.asarray(df[name], dtype=np.float64)
            assert col.dtype == np.float64 and col.flags.c_contiguous and np.array_equal(col, want, equal_nan=True)
        for name, col in zip(("cb_id", "read_id"), got[4:]):
            want = np.array(df[name])
            assert col.dtype == want.dtype and np.array_equal(col, want, equal_nan=True)
        batch = ChunkBatch()
        batch.add("g", df, 0)
        assert batch.n_reads == [len(df)]
    # a frame the block path cannot serve (object column among the wanted ones) falls through
    f = base.copy(); f["pa"] = f["pa"].astype(object)
    assert ChunkBatch._columns_blocks(f) is None
    batch = ChunkBatch()
    batch.add("g", f, 0)
    assert np.array_equal(batch.cols[3][0], np.asarray(base["pa"], dtype=np.float64), equal_nan=True)


def test_chunk_files_read_without_dataframes_equal_the_pandas_path(tmp_path):
    """`_read_chunk_light` (stand-ins for the pandas classes while unpickling) yields the same columns as
    pickle.load + pandas, frame by frame; layouts it does not know make `_load_chunk_packed` fall back."""
    import pickle
    import pandas as pd
    from scape_b200 import apa_core, synth
    us = [synth.make_utr(i, 40 + 7 * i) for i in range(6)]
    frames = [synth.to_dataframe(u) for u in us]
    frames[1]["cb_id"] = frames[1]["cb_id"].astype(float)
    frames[1].loc[3, "cb_id"] = np.nan
    frames[2] = frames[2][list(reversed(frames[2].columns))]
    frames[3]["note"] = "a"
    frames[4]["r"] = frames[4]["r"].astype(np.float32)
    frames[5] = frames[5].iloc[:0]
    path = tmp_path / "demo.100.1.1.input.pkl"
    with open(path, "wb") as fh:
        for u, df in zip(us, frames):
            pickle.dump((u.gene_info_str, df), fh)
    light = apa_core._read_chunk_light(str(path))
    heavy = apa_core.read_chunk_file(str(path))
    assert [g for g, _ in light] == [g for g, _ in heavy]
    for (_, cols), (_, df) in zip(light, heavy):
        for name, col in zip(("x", "l", "r", "pa"), cols[:4]):
            want = np.asarray(df[name], dtype=np.float64)
            assert col.dtype == np.float64 and np.array_equal(col, want, equal_nan=True)
        for name, col in zip(("cb_id", "read_id"), cols[4:]):
            want = np.array(df[name])
            assert col.dtype == want.dtype and np.array_equal(col, want, equal_nan=True)
    # the packed form (what a worker process returns) is the same through both readers
    a = apa_core._load_chunk_packed(str(path))
    b = apa_core.ChunkBatch()
    for g, df in heavy:
        b.add(g, df, 0)
    assert a[0] == b.gene_info and a[1] == b.n_reads
    for got, parts in zip(a[2:6], b.cols):
        assert np.array_equal(got, np.concatenate(parts), equal_nan=True)
    assert np.array_equal(a[6], np.concatenate([f[0] for f in b.frames]), equal_nan=True)
    assert np.array_equal(a[7], np.concatenate([f[1] for f in b.frames]))
    # a file whose wanted column is not a numeric block row: the light reader refuses, the packed loader falls back
    bad = frames[0].copy()
    bad["pa"] = bad["pa"].astype(object)
    path2 = tmp_path / "odd.100.1.1.input.pkl"
    with open(path2, "wb") as fh:
        pickle.dump((us[0].gene_info_str, bad), fh)
    with pytest.raises(Exception):
        apa_core._read_chunk_light(str(path2))
    got = apa_core._load_chunk_packed(str(path2))
    assert np.array_equal(got[5], np.asarray(frames[0]["pa"], dtype=np.float64), equal_nan=True)
    # a file that holds something else than (str, DataFrame) tuples
    path3 = tmp_path / "other.100.1.1.input.pkl"
    with open(path3, "wb") as fh:
        pickle.dump((us[0].gene_info_str, {"x": 1}), fh)
    with pytest.raises(Exception):
        apa_core._read_chunk_light(str(path3))


def test_tmpfs_exchange_of_packed_chunks(tmp_path, monkeypatch):
    """Worker processes hand the packed columns of a chunk file to the parent through a tmpfs file
    (`_load_chunk_shm` / `_open_chunk`); without a usable directory the record carries the arrays itself.
    Both forms equal `_load_chunk_packed`; the file layout survives odd sizes and dtypes."""
    import pickle
    from scape_b200 import apa_core, synth
    us = [synth.make_utr(i, 33 + i) for i in range(3)]
    path = synth.write_chunk_files(us, str(tmp_path), per_file=100)[0]
    want = apa_core._load_chunk_packed(path)
    monkeypatch.setattr(apa_core, "_SHM_DIR", str(tmp_path))
    rec = apa_core._load_chunk_shm(path)
    assert rec[0] == "shm" and os.path.exists(rec[3])
    got = apa_core._open_chunk(rec)
    assert got[8] == (rec[3], len(want[2]), want[6].dtype.str, want[7].dtype.str)
    os.unlink(rec[3])                                   # the mapping outlives the name
    assert got[0] == want[0] and got[1] == want[1]
    for a, b in zip(got[2:8], want[2:8]):
        assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)
    monkeypatch.setattr(apa_core, "_SHM_DIR", str(tmp_path / "missing"))
    rec = apa_core._load_chunk_shm(path)
    assert rec[0] == "inline"
    got = apa_core._open_chunk(rec)
    assert got[8] is None
    for a, b in zip(got[2:8], want[2:8]):
        assert np.array_equal(a, b, equal_nan=True)
    # raw layout: every array starts 8-byte aligned whatever came before it
    monkeypatch.setattr(apa_core, "_SHM_DIR", str(tmp_path))
    arrays = [np.arange(5, dtype=np.int32), np.array([1.5, np.nan, -2.0]), np.arange(3, dtype=np.int16), np.zeros(0), np.arange(2, dtype=np.int64)]
    name = apa_core._shm_write(arrays)
    back = apa_core._shm_read(name, [(a.dtype.str, len(a)) for a in arrays])
    os.unlink(name)
    for a, b in zip(arrays, back):
        assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)
