"""Host-side mirror of the reference's `scape.apa_core` for the infer_pa path.

Same names, arguments and error behaviour as the reference module (apa_core.py:40-184, 236-258,
984-1137) so that the `scape infer_pa` CLI, the TOML parameters, `--pre_para_pkl_file` mode and the
result pickles are unchanged -- but the per-UTR work (binning, likelihood table, marginal tensor,
EM chains, BIC selection, pruning, re-run, labels) happens in libscape_b200.so on the GPU.  There
is no CPU / Taichi fallback: if the library or a CUDA device is missing, these functions raise.

The module also keeps `exp_pa_len` / `cal_exp_pa_len_by_cluster` importable (the reference's
`utils.py:3` imports them from `scape.apa_core`; they are small numpy helpers outside the hot path).
"""
from __future__ import annotations

import datetime
import os
import pickle
import threading
import time
import tomllib
from multiprocessing import Event, Process
from pathlib import Path
from timeit import default_timer as timer
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _lib

try:  # the CLI decorators need click (a reference dependency); the Python API works without it
    import click
except Exception:  # pragma: no cover
    click = None


class Parameters:
    """Result record, attribute-for-attribute the reference's (apa_core.py:236-258).  Pickled under
    the class path `scape.apa_core.Parameters` so `merge_pa` / `cal_exp_pa_len` / `ex_pa_cnt_mat`
    load the results untouched (junction_handler.py:7,76)."""

    def __init__(self, title='', alpha_arr=None, beta_arr=None, ws=None, L=None, cb_id_arr=None, readID_arr=None,
                 K=None):
        self.title = title
        self.alpha_arr = alpha_arr
        self.beta_arr = beta_arr
        self.ws = ws
        self.K = len(self.alpha_arr)
        self.L = L
        self.cb_id_arr = cb_id_arr
        self.readID_arr = readID_arr

    def __str__(self):
        outstr = '-' * 10 + f'{self.title} K={self.K}' + '-' * 10 + '\n'
        if hasattr(self, 'gene_info_str'):
            outstr += f'gene info: {self.gene_info_str}\n'
        outstr += f'K={self.K} L={self.L} Last component is uniform component.\n'
        outstr += f'alpha_arr={self.alpha_arr}\n'
        outstr += f'beta_arr={self.beta_arr}\n'
        outstr += f'ws={np.around(self.ws, decimals=2)}\n'
        if hasattr(self, 'bic'):
            outstr += f'bic={np.around(self.bic, decimals=2)}\n'
        outstr += '-' * 30 + '\n'
        return outstr


Parameters.__module__ = "scape.apa_core"

STATUS_MESSAGES = {
    -1: "AssertionError: read start outside [0, utr_length) (apa_core.py:388)",
    -2: "read coordinates exceed the supported bin range",
    -3: "UTR without reads",
    -4: f"re-run would need more than {_lib.KCAP} pA components",
    -5: "invalid parameters",
    -6: "ValueError: Fewer non-zero entries in p than size (np.random.choice in sample_alpha, apa_core.py:797)",
}


def _result_class():
    """The class object pickles must reference: whatever `scape.apa_core.Parameters` resolves to
    (this repo's shim, or the reference package when this path is grafted into it)."""
    try:
        import importlib
        mod = importlib.import_module("scape.apa_core")
        return getattr(mod, "Parameters", Parameters)
    except Exception:
        return Parameters


# ------------------------------------------------------------------------------------------------
# batch driver (new API: many chunks at once)
# ------------------------------------------------------------------------------------------------
class ChunkBatch:
    """UTRs of one or more chunk files packed as CSR read columns."""

    def __init__(self):
        self.gene_info: List[str] = []
        self.frames = []          # (cb_id, read_id) per UTR
        self.cols = [[], [], [], []]
        self.n_reads: List[int] = []
        self.stream: List[int] = []

    def add_columns(self, gene_info_str, got, stream: int):
        """One UTR from its six columns: x, l, r, pa as float64, cb_id and read_id in their own dtype."""
        self.gene_info.append(gene_info_str)
        for c, col in zip(self.cols, got[:4]):
            c.append(col)
        self.frames.append((got[4], got[5]))
        self.n_reads.append(len(got[0]))
        self.stream.append(stream)

    def add(self, gene_info_str, df, stream: int):
        self.gene_info.append(gene_info_str)
        got = self._columns_blocks(df)
        if got is None:
            got = self._columns_fast(df)
        if got is None:                       # non-numeric extra columns etc.: column by column
            got = [np.asarray(df[name], dtype=np.float64) for name in ("x", "l", "r", "pa")] + \
                  [np.array(df["cb_id"]), np.array(df["read_id"])]
        for c, col in zip(self.cols, got[:4]):
            c.append(col)
        self.frames.append((got[4], got[5]))
        self.n_reads.append(len(got[0]))
        self.stream.append(stream)

    def add_packed(self, stream: int, gene_infos, n_reads, x, l, r, pa, cb, rid):
        """All UTRs of one chunk file at once, as produced by `_load_chunk_packed` (possibly in a worker
        process): concatenated read columns plus the reads-per-UTR list."""
        for c, col in zip(self.cols, (x, l, r, pa)):      # one piece per file: packed() only concatenates
            c.append(col)
        pos = 0
        for gi, n in zip(gene_infos, n_reads):
            sl = slice(pos, pos + n)
            self.gene_info.append(gi)
            self.frames.append((cb[sl], rid[sl]))
            self.n_reads.append(int(n))
            self.stream.append(stream)
            pos += n

    _block_layouts: dict = {}  # block signature -> (columns Index, [(block, row) of x, l, r, pa, cb_id, read_id])

    @classmethod
    def _columns_blocks(cls, df):
        """x, l, r, pa (float64) and cb_id, read_id (their own dtype, like the reference's
        np.array(data[col]), apa_core.py:1013-1014) straight from the frame's block manager: a
        prepare_input frame is one int64 block (x, l, cb_id, read_id, junction) and one float64 block (r, pa,
        seg1_en, seg2_en) whose rows ARE the columns, so nothing is interleaved or converted except x and
        l (int64 -> float64, exact below 2**53 like any float64 view of them).  ~15 us per frame against
        ~150 us through DataFrame.to_numpy.  Uses pandas internals (`_mgr.blocks`, `mgr_locs`): anything
        unexpected returns None and the caller takes the documented-API paths."""
        try:
            blocks = df._mgr.blocks
            sig = tuple((str(b.dtype), b.mgr_locs.as_array.tobytes()) for b in blocks)
            cols = df.columns
            ent = cls._block_layouts.get(sig)
            if ent is None or not (cols is ent[0] or cols.equals(ent[0])):
                loc = [cols.get_loc(n) for n in ("x", "l", "r", "pa", "cb_id", "read_id")]
                if not all(isinstance(j, (int, np.integer)) for j in loc):
                    return None
                where = {}
                for bi, b in enumerate(blocks):
                    for ri, j in enumerate(b.mgr_locs.as_array):
                        where[int(j)] = (bi, ri)
                plan = [where[int(j)] for j in loc]
                for bi, _ in plan:
                    v = blocks[bi].values
                    if type(v) is not np.ndarray or v.ndim != 2 or v.dtype.kind not in "iuf":
                        return None
                ent = (cols, plan)
                cls._block_layouts[sig] = ent
            out = []
            for k, (bi, ri) in enumerate(ent[1]):
                row = blocks[bi].values[ri]
                if k < 4:
                    out.append(np.ascontiguousarray(row, dtype=np.float64))
                else:
                    out.append(np.array(row))
            n = len(out[0])
            if any(len(c) != n or c.ndim != 1 for c in out):
                return None
            return out
        except Exception:                                   # internals moved: documented-API paths
            return None

    _layouts: dict = {}       # (column names, dtypes) -> (positions of x, l, r, pa, cb_id, read_id; dtypes of the two ids)

    @classmethod
    def _columns_fast(cls, df):
        """x, l, r, pa (float64) and cb_id, read_id (their own dtype) of a prepare_input frame through
        ONE DataFrame -> ndarray conversion.  Per-column Series access, `get_loc` and `dtypes` cost
        50-150 us each in pandas, which at ~10k UTR/s is more than the GPU spends on the UTR; the frames
        of a run all share one layout, so positions and id dtypes are looked up once per layout.
        Integer ids come back exactly (checked: integral and below 2**53), otherwise the caller falls
        back to the column-by-column path."""
        try:
            # the layout is keyed by column names AND dtypes: a later frame whose ids have another dtype
            # (e.g. float ids with NaN after int ids) gets its own entry instead of the first frame's
            try:
                dkey = tuple(df._mgr.get_dtypes())      # 5 us; df.dtypes costs 45 us per frame
            except AttributeError:                      # pandas without the block-manager accessor
                dkey = tuple(df.dtypes)
            key = (tuple(df.columns.values), dkey)
            lay = cls._layouts.get(key)
            if lay is None:
                loc = [df.columns.get_loc(n) for n in ("x", "l", "r", "pa", "cb_id", "read_id")]
                if not all(isinstance(j, (int, np.integer)) for j in loc):
                    return None
                dts = df.dtypes
                lay = (loc, [dts.iloc[loc[4]], dts.iloc[loc[5]]])
                if any(d.kind not in "iuf" for d in lay[1]):
                    return None
                cls._layouts[key] = lay
            loc, id_dtypes = lay
            arr = df.to_numpy(dtype=np.float64)
            out = [np.ascontiguousarray(arr[:, j]) for j in loc[:4]]
            for j, dt in zip(loc[4:], id_dtypes):
                col = arr[:, j]
                if dt.kind in "iu":
                    if len(col) and not (np.abs(col).max() < 2.0 ** 53 and np.array_equal(col, np.floor(col))):
                        return None        # NaN, fractional or huge: this frame's ids are not that integer type
                    out.append(col.astype(dt))
                else:
                    out.append(np.ascontiguousarray(col).astype(dt, copy=False))
            return out
        except (TypeError, ValueError, KeyError):
            return None

    def __len__(self):
        return len(self.gene_info)

    def packed(self, scratch: Optional[dict] = None):
        """CSR offsets, the four float64 read columns and the stream ids of the batch.

        `scratch` (a dict the caller keeps, e.g. on its Engine): the columns are concatenated into arrays
        kept there and overwritten by the next call with the same dict (valid until then: enough for
        one `Engine.fit` call).  A fresh 160 MB of columns for 10k UTRs costs ~100 ms in first-touch
        page faults alone."""
        off = np.zeros(len(self) + 1, np.int64)
        np.cumsum(self.n_reads, out=off[1:])
        if scratch is None:
            cat = [np.concatenate(c) if c else np.zeros(0) for c in self.cols]
            return off, cat[0], cat[1], cat[2], cat[3], np.asarray(self.stream, np.int32)
        n = int(off[-1])
        bufs = scratch.get("cols")
        if bufs is None or len(bufs[0]) < n:
            bufs = scratch["cols"] = [np.empty(max(n, 1) + max(n, 1) // 8, np.float64) for _ in range(4)]
        cat = []
        for c, buf in zip(self.cols, bufs):
            view = buf[:n]
            if c:
                np.concatenate(c, out=view)
            cat.append(view)
        return off, cat[0], cat[1], cat[2], cat[3], np.asarray(self.stream, np.int32)


def read_chunk_file(path) -> list:
    """All (gene_info_str, DataFrame) tuples of a prepare_input chunk (apa_core.py:1117-1132)."""
    out = []
    with open(path, 'rb') as fh:
        while True:
            try:
                out.append(pickle.load(fh))
            except EOFError:
                return out


# ---- chunk files without building DataFrames ------------------------------------------------------
# A prepare_input chunk file is a stream of pickled (gene_info_str, DataFrame) tuples
# (input_processor.py:224-259, read back at apa_core.py:1117-1132).  Half of what a worker process spends
# on a file is pandas rebuilding 100 DataFrames (block manager, arrow-backed column Index) that are
# taken apart again two lines later.  `_LightUnpickler` reads the same pickles with stand-ins for the
# pandas classes that only keep what was pickled: the blocks' ndarrays, their placements and the column
# names.  It knows the layout pandas has written since 1.3 (`_unpickle_block`, `BlockManager(blocks,
# axes)`); anything else raises and the file is read with pickle.load + pandas as before.
class _LightFrame:
    def __setstate__(self, state):
        self.state = state


class _LightMgr:
    def __init__(self, blocks, axes):
        self.blocks, self.axes = blocks, axes


class _LightStrings:                 # pandas.arrays.ArrowStringArray stand-in: keeps the pyarrow array
    def __setstate__(self, state):
        self.state = state

    def tolist(self):
        return self.state["_pa_array"].to_pylist()


def _light_block(values, placement, ndim=2):
    return values, placement


def _light_index(cls, d):
    return d                         # {'data': names, 'name': ...} or a RangeIndex's {'start', 'stop', 'step'}


class _LightUnpickler(pickle.Unpickler):
    _swap = {
        ("pandas", "DataFrame"): _LightFrame, ("pandas.core.frame", "DataFrame"): _LightFrame,
        ("pandas.core.internals.managers", "BlockManager"): _LightMgr,
        ("pandas._libs.internals", "_unpickle_block"): _light_block,
        ("pandas.core.indexes.base", "_new_Index"): _light_index,
        ("pandas.arrays", "ArrowStringArray"): _LightStrings,
        ("pandas.core.arrays.string_arrow", "ArrowStringArray"): _LightStrings,
    }

    def find_class(self, module, name):
        got = self._swap.get((module, name))
        return got if got is not None else super().find_class(module, name)


def _light_columns(frame):
    """x, l, r, pa (float64) and cb_id, read_id (own dtype) of one light frame, like
    ChunkBatch._columns_blocks; raises on anything that is not the expected layout."""
    mgr = frame.state["_mgr"]
    if type(mgr) is not _LightMgr or len(mgr.axes) != 2:
        raise ValueError("unexpected manager")
    names = mgr.axes[0]["data"]
    names = names.tolist() if hasattr(names, "tolist") else list(names)
    rows = mgr.axes[1]
    if "stop" in rows:                                   # RangeIndex
        n = len(range(int(rows.get("start") or 0), int(rows["stop"]), int(rows.get("step") or 1)))
    else:
        n = len(rows["data"])
    where = {}
    for values, placement in mgr.blocks:
        locs = range(*placement.indices(len(names))) if isinstance(placement, slice) else np.asarray(placement).tolist()
        for ri, j in enumerate(locs):
            where[int(j)] = (values, ri)
    if sorted(where) != list(range(len(names))):
        raise ValueError("block placements do not cover the columns")
    out = []
    for k, want in enumerate(("x", "l", "r", "pa", "cb_id", "read_id")):
        if names.count(want) != 1:
            raise ValueError("column " + want)
        values, ri = where[names.index(want)]
        if type(values) is not np.ndarray or values.ndim != 2 or values.dtype.kind not in "iuf" or values.shape[1] != n:
            raise ValueError("column " + want + " is not a plain numeric block row")
        row = values[ri]
        out.append(np.ascontiguousarray(row, dtype=np.float64) if k < 4 else np.array(row))
    return out


def _read_chunk_light(path):
    """[(gene_info_str, [x, l, r, pa, cb_id, read_id]), ...] of a chunk file without pandas objects."""
    out = []
    with open(path, 'rb') as fh:
        while True:
            try:
                gene_info_str, frame = _LightUnpickler(fh).load()
            except EOFError:
                return out
            if type(frame) is not _LightFrame:
                raise ValueError("not a DataFrame pickle")
            out.append((gene_info_str, _light_columns(frame)))


def _load_chunk_packed(path):
    """Unpickle one chunk file and pack it (runs in a worker process for many-file calls): the
    DataFrames stay in the worker (or are never built, `_read_chunk_light`), a handful of flat arrays
    come back."""
    b = ChunkBatch()
    try:
        for gene_info_str, cols6 in _read_chunk_light(path):
            b.add_columns(gene_info_str, cols6, 0)
    except Exception:                                    # another pickle layout: pandas reads it
        b = ChunkBatch()
        for gene_info_str, df in read_chunk_file(path):
            b.add(gene_info_str, df, 0)
    cat = lambda parts, dt=None: (np.concatenate(parts) if parts else np.zeros(0, dt or np.float64))
    cols = [cat(c) for c in b.cols]
    cb = cat([f[0] for f in b.frames], np.int64)
    rid = cat([f[1] for f in b.frames], np.int64)
    return b.gene_info, b.n_reads, cols[0], cols[1], cols[2], cols[3], cb, rid


def _write_result_file(path, fixed_run_mode, gene_infos, n_reads, K, L, alpha, beta, ws, bic, n_lb, lb_arr, label, cb, rid):
    """Build the Parameters objects of one chunk file from flat result arrays and pickle them
    (apa_core.py:1134-1137); the mirror image of `_load_chunk_packed`, also worker-side."""
    cls = _result_class()
    pos = 0
    with open(path, 'wb') as fh:
        for u, gi in enumerate(gene_infos):
            k, n = int(K[u]), int(n_reads[u])
            para = cls(title='Final Result (subsample run)' if fixed_run_mode else 'Final Result',
                       alpha_arr=alpha[u, :k].astype('int'), beta_arr=beta[u, :k].copy(), ws=ws[u, :k + 1].copy(),
                       L=int(L[u]), cb_id_arr=cb[pos:pos + n], readID_arr=rid[pos:pos + n])
            para.bic = np.float64(bic[u])
            para.lb_arr = [np.float64(v) for v in lb_arr[u, :int(n_lb[u])]]
            para.label_arr = label[pos:pos + n].copy()
            para.gene_info_str = gi
            pickle.dump(para, fh)
            pos += n
    return path


# ---- big arrays between the worker processes and this one: tmpfs files instead of pipes -----------------
# A packed chunk file is ~2.4 MB of columns and a result file needs ~1.2 MB of labels and ids; through the
# executor's pipes this process receives / sends them at ~1 GB/s, one file after the other (0.25 s + 0.1 s
# for 100 cfg-2 files, a quarter of the whole call).  The workers therefore exchange them through files in
# a tmpfs directory, which both sides map; only names and small per-UTR arrays travel through the pipes.
_SHM_DIR = os.environ.get("SCAPE_B200_SHM_DIR", "/dev/shm")


def _shm_ok() -> bool:
    return bool(_SHM_DIR) and os.path.isdir(_SHM_DIR) and os.access(_SHM_DIR, os.W_OK)


def _shm_write(arrays) -> str:
    """The arrays back to back in a new tmpfs file (each one 8-byte aligned); returns its name."""
    import tempfile
    fd, name = tempfile.mkstemp(prefix="scape_b200_", dir=_SHM_DIR)
    try:
        with os.fdopen(fd, "wb") as fh:
            for a in arrays:
                a = np.ascontiguousarray(a)
                fh.write(memoryview(a).cast("B"))
                pad = (-a.nbytes) % 8
                if pad:
                    fh.write(b"\0" * pad)
    except BaseException:
        _shm_unlink(name)
        raise
    return name


def _shm_read(name: str, specs):
    """The arrays `_shm_write` put into `name`, as read-only views of one mapping; specs = [(dtype str, n)]."""
    total = sum(np.dtype(d).itemsize * n + (-(np.dtype(d).itemsize * n)) % 8 for d, n in specs)
    if total == 0:
        return [np.zeros(0, np.dtype(d)) for d, _ in specs]
    mm = np.memmap(name, dtype=np.uint8, mode="r", shape=(total,))
    out, pos = [], 0
    for d, n in specs:
        nb = np.dtype(d).itemsize * n
        out.append(np.frombuffer(mm, dtype=np.dtype(d), count=n, offset=pos))
        pos += nb + (-nb) % 8
    return out


def _shm_unlink(name):
    try:
        os.unlink(name)
    except OSError:
        pass


def _load_chunk_shm(path):
    """`_load_chunk_packed` in a worker process, the six columns left in a tmpfs file for the parent."""
    gene_infos, n_reads, x, l, r, pa, cb, rid = _load_chunk_packed(path)
    n = len(x)
    if n == 0 or not _shm_ok():
        return ("inline", gene_infos, n_reads, x, l, r, pa, cb, rid)
    try:
        name = _shm_write([x, l, r, pa, cb, rid])
    except Exception:                                     # tmpfs full or unusable: through the pipe after all
        return ("inline", gene_infos, n_reads, x, l, r, pa, cb, rid)
    return ("shm", gene_infos, n_reads, name, n, cb.dtype.str, rid.dtype.str)


def _open_chunk(rec):
    """What `_load_chunk_packed` returns, from a `_load_chunk_shm` record (+ the tmpfs file's name and the id
    dtypes, for `_write_result_file_shm`; None for inline records)."""
    if rec[0] == "inline":
        return rec[1:] + (None,)
    _, gene_infos, n_reads, name, n, cb_dt, rid_dt = rec
    x, l, r, pa, cb, rid = _shm_read(name, [("<f8", n)] * 4 + [(cb_dt, n), (rid_dt, n)])
    return gene_infos, n_reads, x, l, r, pa, cb, rid, (name, n, cb_dt, rid_dt)


def _write_result_file_shm(path, fixed_run_mode, gene_infos, n_reads, K, L, alpha, beta, ws, bic, n_lb, lb_arr,
                           label_ref, ids_ref):
    """`_write_result_file` with the per-read arrays taken from tmpfs files: label_ref = (file, first read,
    reads) into the fit's int64 label array, ids_ref = the chunk's `_load_chunk_shm` file."""
    lname, first, n = label_ref
    label = np.memmap(lname, dtype=np.int64, mode="r")[first:first + n] if n else np.zeros(0, np.int64)
    name, n_ids, cb_dt, rid_dt = ids_ref
    cb, rid = _shm_read(name, [("<f8", n_ids)] * 4 + [(cb_dt, n_ids), (rid_dt, n_ids)])[4:]
    # (np.array: the Parameters own their arrays, like the reference's; the mappings go away with this call)
    return _write_result_file(path, fixed_run_mode, gene_infos, n_reads, K, L, alpha, beta, ws, bic, n_lb, lb_arr,
                              np.array(label), np.array(cb), np.array(rid))


_io_pool = None
_engines = {}
_engines_lock = __import__("threading").Lock()


def _cached_engine(device: int, params, host_threads: int = 0, tensor_dtype: Optional[str] = None):
    """One Engine per (device, parameter set), kept across infer_files calls: its device arenas, pinned
    staging and host pools are expensive to build.  A handle is not re-entrant; callers use a device
    from one thread at a time."""
    import threading
    key = (int(device), bytes(params), tensor_dtype)
    with _engines_lock:
        eng = _engines.get(key)
        if eng is None:
            eng = _engines[key] = _lib.Engine(params, device=device, tensor_dtype=tensor_dtype)
            eng.fit_lock = threading.Lock()          # a handle is not re-entrant (two threads given the same device)
    eng.set_host_threads(host_threads)
    return eng


def _get_io_pool(workers: int):
    """Persistent worker processes for unpickling / pickling chunk files.  forkserver, not fork: the
    caller's process holds a CUDA context."""
    global _io_pool
    if _io_pool is None or _io_pool[0] != workers:
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        if _io_pool is not None:
            _io_pool[1].shutdown(wait=False)
        ctx = mp.get_context("forkserver")
        try:
            ctx.set_forkserver_preload(["numpy", "pandas", "scape_b200.apa_core"])
        except Exception:  # pragma: no cover
            pass
        _io_pool = (workers, ProcessPoolExecutor(max_workers=workers, mp_context=ctx))
    return _io_pool[1]


def results_to_parameters(batch: ChunkBatch, out, fixed_run_mode: bool) -> list:
    """FitOutput -> list of Parameters (one per UTR, input order)."""
    cls = _result_class()
    off = np.zeros(len(batch) + 1, np.int64)
    np.cumsum(batch.n_reads, out=off[1:])
    res = []
    for u in range(len(batch)):
        if out.status[u] != 0:
            msg = STATUS_MESSAGES.get(int(out.status[u]), f"status {int(out.status[u])}")
            if out.status[u] == -1:
                raise AssertionError(f"{batch.gene_info[u]}: {msg}")
            raise Exception(f"{batch.gene_info[u]}: {msg}")
        K = int(out.K[u])
        para = cls(title='Final Result (subsample run)' if fixed_run_mode else 'Final Result',
                   alpha_arr=out.alpha[u, :K].astype('int'), beta_arr=out.beta[u, :K].copy(),
                   ws=out.ws[u, :K + 1].copy(), L=int(out.L[u]),
                   cb_id_arr=batch.frames[u][0], readID_arr=batch.frames[u][1])
        para.bic = np.float64(out.bic[u])
        para.lb_arr = [np.float64(v) for v in out.lb_arr[u, :int(out.n_lb[u])]]
        para.label_arr = out.label[off[u]:off[u + 1]].copy()
        para.gene_info_str = batch.gene_info[u]
        res.append(para)
    return res


def _load_pre_para(kwargs):
    assert kwargs["pre_para_pkl_file"]
    assert os.path.exists(kwargs["pre_para_pkl_file"])
    with open(kwargs["pre_para_pkl_file"], 'rb') as fh:
        return pickle.load(fh)          # first object only (apa_core.py:1002-1003)


def fit_chunks(chunks: Sequence[Sequence[tuple]], seeds: Optional[Sequence[int]] = None, device: int = 0,
               engine: Optional[_lib.Engine] = None, return_raw: bool = False, stream_state=None,
               tensor_dtype: Optional[str] = None, **kwargs):
    """Fit every UTR of several in-memory chunks in ONE library call.  Each chunk is one RNG stream
    (seed 1 by default, like `_infer_pa`), so results equal running `infer` per file.  Returns a
    list (per chunk) of lists of Parameters."""
    fixed = bool(kwargs.get("fixed_run_mode", False))
    pre_para = _load_pre_para(kwargs) if fixed else None
    batch = ChunkBatch()
    for s, chunk in enumerate(chunks):
        for gene_info_str, df in chunk:
            batch.add(gene_info_str, df, s)
    if seeds is None:
        seeds = [1] * len(chunks)
    own = engine is None
    if own:
        engine = _lib.Engine(_lib.make_params(pre_para=pre_para, **kwargs), device=device, tensor_dtype=tensor_dtype)
    try:
        off, x, l, r, pa, sid = batch.packed()
        out = engine.fit(off, x, l, r, pa, sid, np.asarray(seeds, np.uint32), stream_state=stream_state)
    finally:
        if own:
            engine.close()
    paras = results_to_parameters(batch, out, fixed)
    per_chunk, pos = [], 0
    for chunk in chunks:
        per_chunk.append(paras[pos:pos + len(chunk)])
        pos += len(chunk)
    return (per_chunk, out) if return_raw else per_chunk


# ------------------------------------------------------------------------------------------------
# reference API (same signatures)
# ------------------------------------------------------------------------------------------------
def subsample_run(return_model=False, re_run_mode=True, gene_info_str="None", fixed_run_mode=False, **kwargs):
    """apa_core.py:984-1035 for one UTR (`data=DataFrame`).  Consumes and advances the global numpy
    legacy RNG exactly like the reference."""
    if return_model:
        raise NotImplementedError("return_model=True exposes the reference's ApaModel object; not part of the hot path")
    tbl = kwargs.pop('data')
    state = np.random.get_state()
    ss = np.empty((1, 625), np.uint32)
    ss[0, :624] = state[1]
    ss[0, 624] = state[2]
    res = fit_chunks([[(gene_info_str, tbl)]], re_run_mode=re_run_mode, fixed_run_mode=fixed_run_mode,
                     stream_state=ss, **kwargs)[0][0]
    np.random.set_state((state[0], ss[0, :624].copy(), int(ss[0, 624]), 0, 0.0))
    return res


def infer(pickle_input_file, pickle_output_file, **kwargs):
    """apa_core.py:1104-1137: fit every UTR of a chunk pickle, then dump the Parameters objects in
    input order.  The whole file is one library call; the global numpy RNG is consumed and advanced
    as the reference would."""
    print(f"start inferring APA events from input pickle file = {pickle_input_file}. Output file = {pickle_input_file}")
    start_t = timer()
    chunk = read_chunk_file(pickle_input_file)
    state = np.random.get_state()
    ss = np.empty((1, 625), np.uint32)
    ss[0, :624] = state[1]
    ss[0, 624] = state[2]
    res_lst = fit_chunks([chunk], stream_state=ss, **kwargs)[0]
    np.random.set_state((state[0], ss[0, :624].copy(), int(ss[0, 624]), 0, 0.0))
    end_t = timer()
    print(f"Done {len(res_lst)} UTR regions in {(end_t - start_t) / 60} min.")
    with open(pickle_output_file, 'wb') as fh:
        for res in res_lst:
            print(f"save result of {res.gene_info_str}")
            pickle.dump(res, fh)


def infer_files(pkl_input_files: Sequence[str], output_dir: str, device: int = 0,
                devices: Optional[Sequence[int]] = None, io_workers: Optional[int] = None, **kwargs) -> List[str]:
    """Multi-file entry point: same per-file seeding and output naming as
    `_infer_pa`, but all files share one library call so their UTRs run concurrently.

    `devices=[0, 1, ...]`: the files (RNG streams) are bin-packed over the GPUs by estimated cost
    (`scape_b200.shard`) and fitted by one host thread per GPU (the C call releases the
    GIL); there is nothing to exchange between GPUs.  Results do not depend on the partition: every
    file is its own stream, seeded 1.

    `io_workers` (default: one per core for calls with 64 or more files on one GPU, else 0): worker
    processes that unpickle / pack the inputs and build / pickle the outputs; 0 keeps everything in
    this process."""
    global _io_pool
    os.makedirs(os.path.join(output_dir, "pkl_output"), exist_ok=True)
    names, outs = [], []
    for f in pkl_input_files:
        if not os.path.exists(f):
            raise Exception("Given input file does not exists")
        name = os.path.basename(f)[:-10]
        if ".tmp." in name:
            raise Exception("The input file " + name + " is incomplete. Please re-run prepare_input() on " +
                            name.split(".")[0] + ".bam")
        names.append(name)
        outs.append(os.path.join(output_dir, "pkl_output", name + ".res.pkl"))
    for o in outs:
        if os.path.exists(o):
            os.remove(o)
    devs = list(devices) if devices else [device]
    n_files = len(pkl_input_files)
    if io_workers is None:
        # starting the worker processes costs a few seconds once per process (forkserver + pandas import)
        io_workers = min(os.cpu_count() or 1, n_files) if (n_files >= 64 or _io_pool is not None) else 0
    io_workers = int(io_workers)
    if io_workers > 0:
        from concurrent.futures.process import BrokenProcessPool
        try:
            return _infer_files_pooled(pkl_input_files, outs, devs, io_workers, **kwargs)
        except BrokenProcessPool as e:      # e.g. a __main__ that worker processes cannot re-import
            _io_pool = None
            print(f"infer_files: worker processes unavailable ({e}); continuing in this process")
    chunks = [read_chunk_file(f) for f in pkl_input_files]
    if len(devs) <= 1:
        results = fit_chunks(chunks, seeds=[1] * len(chunks), device=devs[0], **kwargs)
    else:
        results = _fit_chunks_multi_gpu(chunks, devs, **kwargs)
    for o, res_lst in zip(outs, results):
        with open(o, 'wb') as fh:
            for res in res_lst:
                pickle.dump(res, fh)
    return outs


def _infer_files_pooled(paths, outs, devices, io_workers, **kwargs):
    """`infer_files` for many files: the DataFrames are unpickled and packed, and the Parameters built
    and pickled, in worker processes (one task per file); this process only sees flat arrays, packs
    them into one `fit_batch` call per GPU (one host thread each; files dealt by estimated cost) and
    slices the results per file.  Same streams, same seeds, same result files as the in-process path."""
    import threading
    from . import shard
    pool = _get_io_pool(io_workers)
    fixed = bool(kwargs.get("fixed_run_mode", False))
    pre_para = _load_pre_para(kwargs) if fixed else None
    shm_files: List[str] = []
    try:
        return _infer_files_pooled_run(pool, paths, outs, devices, fixed, pre_para, shm_files, **kwargs)
    finally:
        for name in shm_files:
            _shm_unlink(name)


def _infer_files_pooled_run(pool, paths, outs, devices, fixed, pre_para, shm_files, **kwargs):
    import threading
    from . import shard
    loads = [pool.submit(_load_chunk_shm, p) for p in paths]
    records, first_error = [], None
    for fut in loads:                   # every record is collected, so that every tmpfs file gets unlinked
        try:
            rec = fut.result()
            if rec[0] == "shm":
                shm_files.append(rec[3])
            records.append(rec)
        except BaseException as e:
            first_error = first_error or e
    if first_error is not None:
        raise first_error
    packed = [_open_chunk(rec) for rec in records]
    if len(devices) > 1:
        hints = [int(np.max(p[2]) + np.max(p[3]) + 50) if len(p[2]) else 2000 for p in packed]   # max x + max l of the file
        costs = shard.stream_costs([p[1] for p in packed], [[h] * len(p[1]) for h, p in zip(hints, packed)])
        parts = shard.lpt_partition(costs, len(devices))
    else:
        parts = [list(range(len(paths)))]
    errors: List[BaseException] = []
    futures, flock = [], threading.Lock()

    def work(dev, mine):
        try:
            if not mine:
                return
            batch = ChunkBatch()
            for s, f in enumerate(mine):
                batch.add_packed(s, *packed[f][:8])
            engine = _cached_engine(dev, _lib.make_params(pre_para=pre_para, **kwargs),
                                    host_threads=max(1, (os.cpu_count() or 1) // len(devices)) if len(devices) > 1 else 0,
                                    tensor_dtype=kwargs.get("tensor_dtype"))
            with getattr(engine, "fit_lock", _engines_lock):
                # (the engine's scratch columns: overwritten by the next fit of this engine, under the same lock)
                scratch = engine.__dict__.setdefault("_packed_scratch", {}) if hasattr(engine, "__dict__") else None
                off, x, l, r, pa, sid = batch.packed(scratch)
                out = engine.fit(off, x, l, r, pa, sid, np.ones(len(mine), np.uint32))
            if np.any(out.status != 0):
                results_to_parameters(batch, out, fixed)      # raises the reference's error for the first bad UTR
            # the fit's per-read labels go to the workers through one tmpfs file (when the chunks came that way)
            label_file = None
            if _shm_ok() and any(packed[f][8] is not None for f in mine) and len(out.label):
                try:
                    label_file = _shm_write([np.asarray(out.label, dtype=np.int64)])
                    with flock:
                        shm_files.append(label_file)
                except Exception:
                    label_file = None
            u0 = 0
            for f in mine:
                gene_infos, n_reads, cb, rid, ids_ref = packed[f][0], packed[f][1], packed[f][6], packed[f][7], packed[f][8]
                n_utr = len(gene_infos)
                us = slice(u0, u0 + n_utr)
                rs = slice(int(off[u0]), int(off[u0 + n_utr]))
                small = (outs[f], fixed, gene_infos, n_reads, out.K[us], out.L[us], out.alpha[us], out.beta[us], out.ws[us],
                         out.bic[us], out.n_lb[us], out.lb_arr[us])
                if label_file is not None and ids_ref is not None:
                    fut = pool.submit(_write_result_file_shm, *small, (label_file, rs.start, rs.stop - rs.start), ids_ref)
                else:
                    fut = pool.submit(_write_result_file, *small, out.label[rs], np.array(cb), np.array(rid))
                with flock:
                    futures.append(fut)
                u0 += n_utr
        except BaseException as e:          # re-raised in the caller's thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(d, p)) for d, p in zip(devices, parts)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    for f in futures:
        f.result()
    return outs


def plan_multi_gpu(chunks: Sequence[Sequence[tuple]], n_gpus: int) -> List[List[int]]:
    """Which chunk (stream) goes to which GPU: LPT packing of the a-priori stream costs."""
    from . import shard
    reads = [[len(df) for _, df in chunk] for chunk in chunks]
    hints = [[int(np.max(np.asarray(df["x"])) + np.max(np.asarray(df["l"])) + 50) if len(df) else 2000
              for _, df in chunk] for chunk in chunks]
    return shard.lpt_partition(shard.stream_costs(reads, hints), n_gpus)


def _fit_chunks_multi_gpu(chunks, devices, **kwargs):
    import threading
    parts = plan_multi_gpu(chunks, len(devices))
    results: List[Optional[list]] = [None] * len(chunks)
    errors: List[BaseException] = []

    def work(dev, mine):
        try:
            if mine:
                got = fit_chunks([chunks[i] for i in mine], seeds=[1] * len(mine), device=dev, **kwargs)
                for i, r in zip(mine, got):
                    results[i] = r
        except BaseException as e:      # re-raised in the caller's thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(d, p)) for d, p in zip(devices, parts)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def _infer_pa(pkl_input_file: str, output_dir: str, **kwargs):
    """apa_core.py:107-147."""
    if not (os.path.exists(pkl_input_file)):
        raise Exception("Given input file does not exists")
    if not (os.path.exists(os.path.join(output_dir, "pkl_output"))):
        os.makedirs(os.path.join(output_dir, "pkl_output"))
    np.random.seed(1)
    filename = os.path.basename(pkl_input_file)[:-10]
    if ".tmp." in filename:
        raise Exception("The input file " + filename + " is incomplete. Please re-run prepare_input() on " +
                        filename.split(".")[0] + ".bam")
    out_pkl_file = os.path.join(output_dir, "pkl_output", filename + ".res.pkl")
    if os.path.exists(out_pkl_file):
        os.remove(out_pkl_file)
    watch_dog_flag = kwargs.get("watch_dog_flag", False)
    if watch_dog_flag:
        exit_event = Event()
        log_file = os.path.join(output_dir, "pkl_output", filename + "log.txt")
        infer_with_watchdog = watch_dog(log_file, exit_event)(infer)
        infer_with_watchdog(pkl_input_file, out_pkl_file, **kwargs)
    else:
        infer(pkl_input_file, out_pkl_file, **kwargs)


def run_infer_pa(pkl_input_file: str, output_dir: str, toml_para_file: str = None, pre_para_pkl_file=None):
    """Body of the `infer_pa` click command (apa_core.py:67-104)."""
    assert Path(output_dir).exists()
    para_dict = {"n_max_apa": 5, "pre_para_pkl_file": pre_para_pkl_file}

    if toml_para_file is None:
        toml_para_file = Path(output_dir) / "parameters.toml"

    if toml_para_file:
        assert os.path.exists(toml_para_file)
        with open(toml_para_file, "rb") as fh:
            user_para_dict = tomllib.load(fh)
            para_dict.update(user_para_dict)
        print(f"Parameter file {toml_para_file} loaded.")
        for k, v in user_para_dict.items():
            print(f"{k} = {v}")
        print()

    if pre_para_pkl_file:
        assert os.path.exists(pre_para_pkl_file)
        para_dict["fixed_run_mode"] = True
        para_dict["pre_para_pkl_file"] = pre_para_pkl_file
        import tomli_w
        with open(toml_para_file, 'wb') as fh:
            tomli_w.dump(para_dict, fh)

    if "output_dir" in para_dict:
        del para_dict["output_dir"]

    _infer_pa(pkl_input_file, output_dir, **para_dict)


if click is not None:
    @click.command(name="infer_pa")
    @click.option('--pkl_input_file', type=str, help='input pickle file (result of prepare_input)', required=True)
    @click.option('--output_dir', type=str, help='output directory', required=True)
    @click.option('--toml_para_file', type=str, help='a TOML file specifies user-defined parameters', default=None,
                  required=False)
    @click.option('--pre_para_pkl_file', type=str,
                  help='a pickle file with pre-specified pA sites and utr length, result file of scape analysis',
                  default=None, required=False)
    def infer_pa(pkl_input_file: str, output_dir: str, toml_para_file: str = None, pre_para_pkl_file=None):
        """
        INPUT:
        - pkl_input_file: file path (pickle) including information for each UTR region
        - output_dir: path to output_dir folder
        - toml_para_file: a TOML file specifies user-defined parameters

        OUTPUT:
        - Pickle file including Parameters for each UTR region
        """
        run_infer_pa(pkl_input_file, output_dir, toml_para_file, pre_para_pkl_file)
else:  # pragma: no cover
    infer_pa = run_infer_pa


# ---- helpers the reference's downstream stages import from scape.apa_core (utils.py:3) ------------
def exp_pa_len(apamix_res, label_arr):
    """Expected pA length statistic (apa_core.py:1038-1052)."""
    if apamix_res.K == 1:
        return 1.0
    if len(label_arr) == 0:
        return np.nan
    is_pa = label_arr < apamix_res.K
    if not np.any(is_pa):
        return np.nan
    labs, cnt = np.unique(label_arr[is_pa], return_counts=True)
    ws = np.zeros(apamix_res.K)
    ws[labs] = cnt
    ws = ws / np.sum(ws)
    a = apamix_res.alpha_arr
    return np.sum(ws * (1.0 + 9.0 * (a - a[0]) / (a[-1] - a[0])))


def cal_exp_pa_len_by_cluster(apamix_res, partition):
    """apa_core.py:1055-1063."""
    partition = np.array(partition)
    clusters = np.unique(partition)
    out = np.zeros(len(clusters))
    for i, c in enumerate(clusters):
        out[i] = exp_pa_len(apamix_res, apamix_res.label_arr[partition == c])
    return clusters, out


# ---- watchdog (apa_core.py:1066-1101) -------------------------------------------------------------
def _watch_dog(log_file: str, exit_event):
    import psutil
    gib = 1024.0 ** 3
    with open(log_file, "w") as fh:
        while not exit_event.is_set():
            mem = psutil.virtual_memory()
            used, avail, total = round(mem.used / gib, 2), round(mem.available / gib, 2), round(mem.total / gib, 2)
            fh.write(datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S") + "\n")
            fh.write(f'The CPU usage is: {psutil.cpu_percent(4)}%\n')
            fh.write(f"Memory usage: used = {used} GB ({round(used / total * 100, 2)}%);  "
                     f"available={avail} GB ({round(avail / total * 100, 2)}%); total={total} GB\n")
            fh.write(str(mem))
            fh.write("\n\n")
            fh.flush()
            exit_event.wait(60)


def watch_dog(log_file: str, exit_event):
    def decorate(task_func):
        def wrapper(*args, **kwargs):
            print(f"Launching watch dog. log_file = {log_file}")
            proc = Process(target=_watch_dog, args=(log_file, exit_event))
            proc.start()
            start_t = timer()
            try:
                return task_func(*args, **kwargs)
            finally:
                print(f"Task takes {(timer() - start_t) / 60} minutes.")
                print("Task finished, terminating watch dog process.")
                exit_event.set()
                proc.join()
        return wrapper
    return decorate
