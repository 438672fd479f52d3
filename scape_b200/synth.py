"""Synthetic `prepare_input`-format UTR chunks (SURVEY.md section 8d, cfg-2 / cfg-3 generator spec).

The reference ships 4 example UTRs only; BASELINE.json's configs are quoted on synthetic batches,
so the generator is part of the product's bench/test tooling.  It draws reads from the model the
reference fits (apa_core.py:576-640): pA site theta ~ N(alpha_k, beta_k), polyA length s ~ U{20..149},
read start x ~ N(theta + s - mu_f, sigma_f), read length l <= theta - x, 5 % uniform noise, a 1.5 % subset of
junction reads carrying an observed pA site.  Chunk files use the exact on-disk layout
`prepare_input` writes (input_processor.py:224-259, 610-636): concatenated
`pickle.dump((gene_info_str, DataFrame[x,l,r,pa,cb_id,read_id,junction,seg1_en,seg2_en]))`.
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

COLUMNS = ["x", "l", "r", "pa", "cb_id", "read_id", "junction", "seg1_en", "seg2_en"]
SEED_BASE = 20260000


@dataclass
class SynthUtr:
    gene_info_str: str
    x: np.ndarray       # int64[R]
    l: np.ndarray       # int64[R]
    r: np.ndarray       # float64[R] (NaN: polyA length unknown, always for 10x data)
    pa: np.ndarray      # float64[R] (NaN unless junction read)
    cb_id: np.ndarray   # int64[R]
    read_id: np.ndarray  # int64[R]
    true_alpha: np.ndarray
    true_beta: np.ndarray
    true_ws: np.ndarray
    L_true: int

    @property
    def n_reads(self):
        return len(self.x)


def _read_lengths(rng, n):
    kind = rng.random(n)
    out = np.where(kind < 0.6, 98, np.where(kind < 0.9, 132, 0))
    short = out == 0
    out[short] = rng.integers(31, 98, size=int(short.sum()))
    return out.astype(np.int64)


def make_utr(u: int, n_reads: int = 500, long_utr: bool = False) -> SynthUtr:
    """One UTR, fully determined by its index `u` (np.random.default_rng(20260000 + u))."""
    rng = np.random.default_rng(SEED_BASE + u)
    L_true = int(rng.integers(800, 20001 if long_utr else 4001))
    K = int(rng.choice([1, 2, 3, 4], p=[0.4, 0.3, 0.2, 0.1]))
    span = L_true - 100 - 300
    while K > 1 and (K - 1) * 150 > span:
        K -= 1
    free = span - (K - 1) * 150
    alpha = np.sort(rng.integers(0, free + 1, size=K)) + 300 + 150 * np.arange(K)
    beta = rng.choice(np.arange(10, 55, 5), size=K).astype(float)
    w = rng.dirichlet(2.0 * np.ones(K)) * 0.95

    n_noise = int(rng.binomial(n_reads, 0.05))
    n_sig = n_reads - n_noise
    xs, ls, pas = [], [], []
    need = n_sig
    while need > 0:
        m = int(need * 1.5) + 16
        comp = rng.choice(K, size=m, p=w / w.sum())
        theta = np.rint(rng.normal(alpha[comp], beta[comp]))
        s = rng.integers(20, 150, size=m)
        x = np.rint(theta + s - rng.normal(300.0, 50.0, size=m))
        l = _read_lengths(rng, m).astype(float)
        # 1.5 % junction reads: the read ends on the pA site, so l = theta - x + 1 (<= 132) and pa = theta
        junction = rng.random(m) < 0.015
        lj = rng.integers(31, 133, size=m).astype(float)
        x = np.where(junction, theta - lj + 1, x)
        room = theta - x
        l = np.where(junction, room + 1, np.minimum(l, room))
        pa = np.where(junction, x + l - 1, np.nan)
        ok = (x >= 0) & (l >= 20)
        xs.append(x[ok][:need]); ls.append(l[ok][:need]); pas.append(pa[ok][:need])
        need -= len(xs[-1])
    if n_noise:
        xs.append(np.floor(rng.random(n_noise) * max(1, L_true - 150)))
        ls.append(_read_lengths(rng, n_noise).astype(float))
        pas.append(np.full(n_noise, np.nan))
    x = np.concatenate(xs); l = np.concatenate(ls); pa = np.concatenate(pas)
    order = rng.permutation(len(x))
    x, l, pa = x[order], l[order], pa[order]
    R = len(x)
    return SynthUtr(
        gene_info_str=f"1:SYN{u:06d}:1:{1000}-{1000 + L_true}:+",
        x=x.astype(np.int64), l=l.astype(np.int64), r=np.full(R, np.nan), pa=pa,
        cb_id=rng.integers(0, 10000, size=R).astype(np.int64), read_id=np.arange(R, dtype=np.int64),
        true_alpha=alpha, true_beta=beta, true_ws=w, L_true=L_true)


def heavy_tail_read_counts(n_utr: int, seed: int = SEED_BASE) -> np.ndarray:
    """cfg-3: R_u = clip(round(exp(N(ln 600, 1.6^2))), 10, 200000)."""
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(np.exp(rng.normal(np.log(600.0), 1.6, size=n_utr))), 10, 200000).astype(np.int64)


def make_batch(n_utr: int, reads_per_utr: Optional[int] = 500, heavy_tail: bool = False,
               first: int = 0) -> List[SynthUtr]:
    """cfg-2 (`reads_per_utr=500`) or cfg-3 (`heavy_tail=True`) batches."""
    if heavy_tail:
        counts = heavy_tail_read_counts(first + n_utr)[first:]
        cut = np.quantile(heavy_tail_read_counts(max(first + n_utr, 1000)), 0.99)
        return [make_utr(first + i, int(c), long_utr=bool(c >= cut)) for i, c in enumerate(counts)]
    return [make_utr(first + i, int(reads_per_utr)) for i in range(n_utr)]


def to_dataframe(u: SynthUtr):
    """The 9-column frame `prepare_input` pickles (input_processor.py:636)."""
    import pandas as pd
    R = u.n_reads
    nan = np.full(R, np.nan)
    return pd.DataFrame({"x": u.x, "l": u.l, "r": u.r, "pa": u.pa, "cb_id": u.cb_id, "read_id": u.read_id,
                         "junction": np.zeros(R, dtype=np.int64), "seg1_en": nan, "seg2_en": nan},
                        columns=COLUMNS)


def write_chunk_files(utrs: Sequence[SynthUtr], out_dir: str, per_file: int = 100,
                      stem: str = "synth") -> List[str]:
    """Write `<out_dir>/pkl_input/<stem>.<per_file>.<nfiles>.<i>.input.pkl` like prepare_input
    (input_processor.py:258-259) and return the paths."""
    d = os.path.join(out_dir, "pkl_input")
    os.makedirs(d, exist_ok=True)
    n_files = (len(utrs) + per_file - 1) // per_file
    paths = []
    for f in range(n_files):
        p = os.path.join(d, f"{stem}.{per_file}.{n_files}.{f + 1}.input.pkl")
        with open(p, "wb") as fh:
            for u in utrs[f * per_file:(f + 1) * per_file]:
                pickle.dump((u.gene_info_str, to_dataframe(u)), fh)
        paths.append(p)
    return paths


DEFAULT_TOML = """\
watch_dog_flag = false
re_run_mode = true
debug = false
n_max_apa = 5
n_min_apa = 1
utr_length = 2000
min_LA = 20
max_LA = 150
mu_f = 300
sigma_f = 50
min_pa_gap = 100
max_beta = 70
theta_step = 9
beta_step = 5
min_ws = 0.05
max_unif_ws = 0.15
"""


def write_default_toml(out_dir: str) -> str:
    """`<out_dir>/parameters.toml` with the reference defaults (tutorial/default_config.toml)."""
    p = os.path.join(out_dir, "parameters.toml")
    with open(p, "w") as fh:
        fh.write(DEFAULT_TOML)
    return p
