"""CPU ORACLE for the `scape infer_pa` hot path -- TEST INFRASTRUCTURE, NOT THE PRODUCT.

A from-scratch FP64 numpy restatement of the reference algorithm (SCAPE-APA 1.0.4,
`/root/reference/src/scape/apa_core.py` + `taichi_core.py`), written so that it can travel to the
GPU box (which has no `/root/reference`).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product package
`scape_b200` never does.

PARITY STATUS: PINNED.  `tests/test_oracle_vs_reference.py` (runs in the build container, where
the reference is mounted) drives the UNMODIFIED reference `apa_core.py` through
`oracle/ref_harness.py` on the 4 shipped example UTRs and on synthetic chunks and requires this
restatement to reproduce K / alpha / beta / labels / iteration counts exactly and ws / bic / lb_arr
to 1e-12 relative; the reference's outputs are committed as fixtures in `tests/golden/` by
`oracle/gen_golden.py`, and the travelling tests check this file against those fixtures.  The
kernel-level known-answer cases of `src/scape/taichi_code_test.py:514-593` are restated in
`tests/test_oracle_kat.py`.

Third-party arithmetic used exactly as the reference uses it (same library calls, so same
rounding): numpy legacy `RandomState` (MT19937) streams, `scipy.signal.find_peaks`,
`scipy.stats.entropy`, `np.matmul`, numpy pairwise `np.sum`.  Taichi 1.7.2 (un-vendored, absent)
only supplied exp/log/sqrt + a parallel-for; its four interface functions are restated below from
the reference's own pure-Python twins (`src/scape/taichi_code_test.py:249-430`).

Everything is organised around grid *indices* (theta index, beta index) instead of the reference's
value + searchsorted lookups; that is equivalent because every alpha / beta the EM ever holds is a
grid point (apa_core.py:805, 819, 523).
"""
from __future__ import annotations

import math
import pickle
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
from scipy import stats as _sp_stats
from scipy.signal import find_peaks as _find_peaks

SENTINEL = float(np.finfo("f").min)      # apa_core.py:428, taichi_core.py:8  (-3.4028235e38)
POS_SENTINEL = float(np.finfo("f").max)  # apa_core.py:427
_PI = 3.141592653589793                  # taichi_core.py:9
N_ROUND = 50                             # apa_core.py:422
N_TRIAL = 10                             # apa_core.py:847

DEFAULTS = dict(  # apa_core.py:333-363 == tutorial/default_config.toml
    n_max_apa=5, n_min_apa=1, utr_length=2000, min_LA=20, max_LA=150, mu_f=300, sigma_f=50,
    min_pa_gap=100, max_beta=70, theta_step=9, beta_step=5, min_ws=0.05, max_unif_ws=0.15,
)


# ----------------------------------------------------------------------------------------------
# Taichi layer (taichi_core.py:24-246), restated from taichi_code_test.py twins
# ----------------------------------------------------------------------------------------------
def _log_normal(x, mu, sigma):
    """taichi_core.py:32-33 / taichi_code_test.py:46-51."""
    return -0.5 * ((x - mu) / sigma) ** 2 - np.log(sigma) - 0.5 * math.log(2 * _PI)


def _pdf_normal(x, mu, sigma):
    """taichi_core.py:36-37 / taichi_code_test.py:64-69."""
    return np.exp(-0.5 * ((x - mu) / sigma) ** 2) / math.sqrt(2 * _PI) / sigma


def _loglik_l_given_xt(x, l, theta):
    """taichi_core.py:57-62: -log(theta-x) where the read fits (l <= theta-x), else the sentinel."""
    span = theta - x
    ok = l <= span
    out = np.full(np.shape(span), SENTINEL)
    np.negative(np.log(span, where=ok, out=np.zeros_like(span)), where=ok, out=out)
    return out


def _lik_l_given_xt(x, l, theta):
    """taichi_core.py:65-70."""
    span = theta - x
    ok = l <= span
    out = np.zeros(np.shape(span))
    np.divide(1.0, span, where=ok, out=out)
    return out


def _seq_sum_last(a):
    """Strictly sequential (left-to-right) sum over the last axis, like the serialised Taichi
    loops (taichi_core.py:46-53, 147-152)."""
    a = np.asarray(a)
    if a.shape[-1] == 0:
        return np.zeros(a.shape[:-1])
    return np.cumsum(a, axis=-1)[..., -1]


def _lse_rows(mat):
    """taichi_core.py:41-54 / taichi_code_test.py:82-93: two-pass max / sum-exp / log per row."""
    m = np.max(mat, axis=-1)
    return np.log(_seq_sum_last(np.exp(mat - m[..., None]))) + m


def loglik_xlr_t_pa(x_arr, l_arr, pa_arr, theta, sigma_f):
    """Junction-pA reads (taichi_core.py:101-107,183-197; twin taichi_code_test.py:249-254)."""
    x_arr = np.asarray(x_arr, float)
    return _loglik_l_given_xt(x_arr, np.asarray(l_arr, float), theta) + \
        _log_normal(np.asarray(pa_arr, float) - theta, 0, sigma_f)


def loglik_xlr_t_r_known(x_arr, l_arr, r_arr, s_dis_arr, pmf_s_arr, theta, mu_f, sigma_f):
    """Reads with an observed polyA length r (taichi_core.py:111-132,200-207; twin :285-309)."""
    x = np.asarray(x_arr, float)[:, None]
    l = np.asarray(l_arr, float)[:, None]
    r = np.asarray(r_arr, float)[:, None]
    s = np.asarray(s_dis_arr, float)[None, :]
    pmf = np.asarray(pmf_s_arr, float)
    logpmf = np.log(pmf)
    live = ~(s < r)                                      # `if s < r: sentinel; continue`
    mass = _seq_sum_last(np.where(live, pmf[None, :], 0.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        log_r = np.where(r <= s, -np.log(s), SENTINEL)   # taichi_core.py:86-90
        term = log_r + _log_normal(x, theta + s - mu_f, sigma_f) + _loglik_l_given_xt(x, l, theta) + logpmf[None, :]
        term = np.where(live, term, SENTINEL)
        return _lse_rows(term) - np.log(mass)


def loglik_xlr_t_r_unknown(x_arr, l_arr, r_arr, s_dis_arr, pmf_s_arr, theta, mu_f, sigma_f):
    """Reads without polyA length (taichi_core.py:141-157,210-215; twin :353-365)."""
    x = np.asarray(x_arr, float)[:, None]
    l = np.asarray(l_arr, float)[:, None]
    s = np.asarray(s_dis_arr, float)[None, :]
    pmf = np.asarray(pmf_s_arr, float)[None, :]
    term = 1 / s * _pdf_normal(x, theta + s - mu_f, sigma_f) * _lik_l_given_xt(x, l, theta) * pmf
    acc = _seq_sum_last(term)
    acc = np.where(acc < 1e-300, 0.0, acc)               # taichi_core.py:154-155
    out = np.full(acc.shape, SENTINEL)
    np.log(acc, where=acc > 0.0, out=out)                # my_log, taichi_core.py:25-29
    return out


def marginal_window(all_theta, alpha, beta):
    """taichi_core.py:221-222 (identical to the mask form at apa_core.py:643)."""
    lo = int(np.searchsorted(all_theta, alpha - 3 * beta, side="left"))
    hi = int(np.searchsorted(all_theta, alpha + 3 * beta, side="right") - 1)
    return lo, hi


def loglik_marginal_lxr(alpha, beta, all_theta, table):
    """One (alpha, beta) slab (taichi_core.py:160-179,218-234; twin taichi_code_test.py:398-430)."""
    lo, hi = marginal_window(all_theta, alpha, beta)
    logp = _log_normal(all_theta[lo:hi + 1], alpha, beta)
    logp_sum = math.log(float(_seq_sum_last(np.exp(logp))))
    return _lse_rows(table[:, lo:hi + 1] + logp[None, :] - logp_sum)


def loglik_marginal_lxr_scipy(alpha, beta, all_theta, table):
    """The numpy/scipy variant fixed_run uses instead of the Taichi kernel (apa_core.py:642-651):
    same window and the same mathematics, rounded by scipy's logpdf / logsumexp."""
    from scipy.special import logsumexp
    sel = np.where(np.logical_and(all_theta >= alpha - 3 * beta, all_theta <= alpha + 3 * beta))[0]
    logp = _sp_stats.norm(loc=alpha, scale=beta).logpdf(all_theta[sel])
    res = np.zeros((table.shape[0], len(sel))) + SENTINEL
    for i, t in enumerate(sel):
        res[:, i] = table[:, t] + logp[i]
    return logsumexp(res, axis=1) - logsumexp(logp)


def get_loglik_marginal_tensor(all_theta, predef_beta_arr, table):
    """tensor[i, j, n] (taichi_core.py:237-246)."""
    out = np.empty((len(all_theta), len(predef_beta_arr), table.shape[0]))
    for i, a in enumerate(all_theta):
        for j, b in enumerate(predef_beta_arr):
            out[i, j] = loglik_marginal_lxr(a, b, all_theta, table)
    return out


# ----------------------------------------------------------------------------------------------
# Binning and per-UTR model set-up (apa_core.py:285-327, 332-462, 576-584, 681-700)
# ----------------------------------------------------------------------------------------------
def bin_reads(x, l, r, pa, steps=(5, 10, 10, 5)):
    """apa_core.py:285-327.  Returns (x, l, r, pa) bin means, cnt, and read->bin index."""
    cols = [np.array(x), np.array(l), np.array(r), np.array(pa)]

    def edges(col, step):
        with np.errstate(all="ignore"):
            top = np.nanmax(col) if len(col) else np.nan
        if np.isnan(top):
            return np.array([0, step])           # apa_core.py:305-312 (all-NaN column)
        return np.arange(0, step + top, step)

    labels = []
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        for col, step in zip(cols, steps):
            e = edges(col, step)
            tmp = col.copy()
            tmp[np.isnan(tmp)] = -1
            labels.append(np.digitize(tmp, e, right=False))
    key = np.column_stack(labels)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    inv = np.asarray(inv).reshape(-1)
    means = [np.bincount(inv, c) / cnt for c in cols]
    return means[0], means[1], means[2], means[3], cnt, inv


def smooth_profile(y, weights):
    """Truncated, edge-renormalised Gaussian smoothing (apa_core.py:681-700)."""
    half = (len(weights) - 1) // 2
    n = len(y)
    wsum = np.sum(weights)
    out = np.full_like(y, 0)
    for i in range(half):
        keep = np.arange(i - half, i + half + 1) >= 0
        out[i] = np.sum(weights[keep] * y[0:i + half + 1]) / np.sum(weights[keep])
    for i in range(half, n - half):
        out[i] = np.sum(weights * y[i - half:i + half + 1]) / wsum
    for i in range(n - half, n):
        keep = np.arange(i - half, i + half + 1) < n
        out[i] = np.sum(weights[keep] * y[i - half:n]) / np.sum(weights[keep])
    return out


def smoothing_weights(bw):
    """apa_core.py:684-685."""
    return np.exp(-np.arange(-3 * bw, 3 * bw + 1) ** 2 / (2 * bw * bw))


@dataclass
class UtrModel:
    """Everything the EM needs for one UTR (the reference keeps this on `ApaModel`)."""
    x: np.ndarray
    l: np.ndarray
    r: np.ndarray
    pa: np.ndarray
    cnt: np.ndarray
    read_to_bin: np.ndarray
    L: int
    min_theta: float
    theta: np.ndarray
    betas: np.ndarray
    s_dis: np.ndarray
    pmf_s: np.ndarray
    prm: dict
    unif_loglik: float = 0.0
    prof_x: Optional[np.ndarray] = None
    prof_y: Optional[np.ndarray] = None
    peak_idx: Optional[np.ndarray] = None
    peak_w: Optional[np.ndarray] = None
    table: Optional[np.ndarray] = None
    tensor: Optional[np.ndarray] = None
    trace: Optional[list] = None
    path: list = field(default_factory=list)   # (k_max, K selected by BIC, K after pruning) per sweep

    @property
    def n(self):
        return len(self.cnt)


def build_model(x, l, r, pa, prm) -> UtrModel:
    """ApaModel.__init__ (apa_core.py:365-437) for already-resolved `utr_length`."""
    bx, bl, br, bpa, cnt, inv = bin_reads(x, l, r, pa)
    utr_length = prm["utr_length"]
    L = utr_length if utr_length > 2000 else 2000                       # :387
    if not all(0 <= v < utr_length for v in bx):                        # :388
        raise AssertionError("read start outside [0, utr_length)")
    s_dis = np.arange(prm["min_LA"], prm["max_LA"], 10)                 # :394
    pmf = np.repeat(1 / len(s_dis), len(s_dis))
    pmf = pmf / sum(pmf)                                                # :395-396
    min_theta = int(min(bl)) + 0.0                                      # :407
    theta = np.arange(int(min_theta), int(L), int(prm["theta_step"])) + 0.0   # :409 / :940
    betas = np.arange(prm["beta_step"], prm["max_beta"], prm["beta_step"]) + 0.0  # :942
    return UtrModel(bx, bl, br, bpa, cnt, inv, L, min_theta, theta, betas, s_dis, pmf, dict(prm))


def uniform_loglik(m: UtrModel) -> float:
    """apa_core.py:576-584."""
    return math.log(1 / m.L * (1 / m.L) * (1 / m.prm["max_LA"]))


def coverage_profile(m: UtrModel):
    """apa_core.py:454-462."""
    cov = np.zeros(m.L)
    for i in range(m.n):
        a = int(m.x[i])
        cov[a:a + int(m.l[i])] += m.cnt[i]
    xs = np.hstack([np.arange(-100, 0), np.arange(m.L), m.L + np.arange(100)])
    ys = np.hstack([np.zeros(100), cov, np.zeros(100)])
    ys = smooth_profile(ys, smoothing_weights(m.prm["beta_step"] * 3))
    return xs, ys


def find_profile_peaks(m: UtrModel):
    """First half of sample_alpha (apa_core.py:784-794): RNG-free, so hoisted out of the chains."""
    idx, _ = _find_peaks(m.prof_y, distance=m.prm["min_pa_gap"])
    bw = m.prm["beta_step"] * 3
    w = np.zeros(len(idx))
    for i, p in enumerate(idx):
        w[i] = sum(m.prof_y[p - bw:p + bw + 1])
    w = w / sum(w)
    return idx, w


def theta_table(m: UtrModel, theta: Sequence[float]) -> np.ndarray:
    """loglik_xlr_t over a theta list (apa_core.py:620-640, 954-957): table[n, t]."""
    has_pa = ~np.isnan(m.pa)
    has_r = ~np.isnan(m.r)
    sel_pa = has_pa
    sel_known = has_r & ~has_pa
    sel_unknown = ~has_r & ~has_pa                                      # :439-452
    out = np.zeros((m.n, len(theta)))
    p = m.prm
    for t, th in enumerate(theta):
        if sel_pa.any():
            out[sel_pa, t] = loglik_xlr_t_pa(m.x[sel_pa], m.l[sel_pa], m.pa[sel_pa], th, p["sigma_f"])
        if sel_known.any():
            out[sel_known, t] = loglik_xlr_t_r_known(m.x[sel_known], m.l[sel_known], m.r[sel_known],
                                                     m.s_dis, m.pmf_s, th, p["mu_f"], p["sigma_f"])
        if sel_unknown.any():
            out[sel_unknown, t] = loglik_xlr_t_r_unknown(m.x[sel_unknown], m.l[sel_unknown], m.r[sel_unknown],
                                                         m.s_dis, m.pmf_s, th, p["mu_f"], p["sigma_f"])
    return out


# ----------------------------------------------------------------------------------------------
# Random initialisation -- literal numpy legacy-RNG call sequence (apa_core.py:655-677, 781-829)
# ----------------------------------------------------------------------------------------------
def snap_to_grid(grid, vals):
    """find_nearest (apa_core.py:537-549): nearest grid point, ties go up."""
    pos = np.searchsorted(grid, vals, side="left")
    out = pos.copy()
    for i, p in enumerate(pos):
        if p == 0:
            continue
        if p == len(grid):
            out[i] = len(grid) - 1
        elif vals[i] - grid[p - 1] >= grid[p] - vals[i]:
            out[i] = p
        else:
            out[i] = p - 1
    return out


def draw_alpha_idx(m: UtrModel, k: int, rng) -> np.ndarray:
    """sample_alpha (apa_core.py:781-807) -> theta-grid indices."""
    peaks = m.prof_x[m.peak_idx]
    if k <= len(peaks):
        picked = rng.choice(peaks, size=k, replace=False, p=m.peak_w)
    else:
        extra = rng.choice(m.L, size=k - len(peaks), replace=False)
        picked = np.concatenate((peaks, extra))
    jitter = np.rint(5 * m.prm["beta_step"] * (2 * rng.uniform(low=0.0, high=1.0, size=k) - 1))
    return snap_to_grid(m.theta, np.sort(picked + jitter))


def draw_weights(m: UtrModel, k: int, rng) -> np.ndarray:
    """init_ws (apa_core.py:809-815).  NB the capped branch does not renormalise."""
    w = rng.uniform(size=(k + 1))
    w = w / sum(w)
    cap = m.prm["max_unif_ws"]
    if w[-1] > cap:
        w[:-1] = w[:-1] * (1 - cap)
        w[-1] = cap
    return w


def draw_component_order(k: int, n: int, rng) -> np.ndarray:
    """gen_k_arr (apa_core.py:655-677).  The 'no repeat' swap at :663-665 can never fire because
    :667 is a comparison, so last_ind stays -1."""
    if k == 0 or k == 1:
        return np.zeros(n, dtype="int")
    arr = rng.permutation(k)
    out = []
    pos = 0
    for _ in range(n):
        if pos % k == 0:
            rng.shuffle(arr)
            pos = 0
        out.append(arr[pos])
        pos += 1
    return np.array(out, dtype="int")


@dataclass
class Chain:
    """One EM chain's parameters, by grid index."""
    a_idx: np.ndarray
    b_idx: np.ndarray
    ws: np.ndarray
    bic: float = float("nan")
    lb_arr: List[float] = field(default_factory=list)
    k_order: Optional[np.ndarray] = None

    @property
    def K(self):
        return len(self.a_idx)


def draw_chain(m: UtrModel, k: int, rng) -> Chain:
    """init_para (apa_core.py:817-829)."""
    a = draw_alpha_idx(m, k, rng)
    b_val = rng.choice(m.betas, size=k, replace=True)
    b = np.searchsorted(m.betas, b_val, side="left")
    w = draw_weights(m, k, rng)
    return Chain(a_idx=np.asarray(a, dtype=np.int64), b_idx=np.asarray(b, dtype=np.int64), ws=w)


# ----------------------------------------------------------------------------------------------
# EM (apa_core.py:473-573, 702-779)
# ----------------------------------------------------------------------------------------------
def _log_w(w):
    return SENTINEL if w <= 0.0 else np.log(w)                          # :478, :515


def _refresh_column(m: UtrModel, ch: Chain, k: int, log_z: np.ndarray):
    """cal_z_k (apa_core.py:473-488)."""
    if k < ch.K:
        log_z[:, k] = _log_w(ch.ws[k]) + m.tensor[ch.a_idx[k]][ch.b_idx[k]]
    else:
        log_z[:, k] = _log_w(ch.ws[k]) + m.unif_loglik


def _responsibilities(m: UtrModel, log_z: np.ndarray) -> np.ndarray:
    """norm_z (apa_core.py:490-495): count-tempered softmax."""
    z = log_z - np.max(log_z, axis=1, keepdims=True)
    z = np.multiply(z, m.cnt[:, np.newaxis])
    z = np.exp(z)
    return z / np.sum(z, axis=1, keepdims=True)


def _update_weights(m: UtrModel, z: np.ndarray) -> np.ndarray:
    """maximize_ws (apa_core.py:498-505)."""
    w = np.matmul(m.cnt, z)
    w = w / np.sum(w)
    cap = m.prm["max_unif_ws"]
    if w[-1] > cap:
        w[:-1] = (1 - cap) * w[:-1] / np.sum(w[:-1])
        w[-1] = cap
    return w


_GRID_BLOCK_BYTES = 1 << 21


def _grid_argmax(m: UtrModel, ch: Chain, z: np.ndarray, k: int):
    """max_alpha_beta (apa_core.py:507-523), evaluated slab-wise instead of one np.sum per
    candidate.  Per candidate the arithmetic is the reference's: ((log w_k + tensor) * Z_k) * cnt,
    pairwise-summed over the N contiguous reads; first maximum in (alpha asc, beta asc) order."""
    lo = 0 if k == 0 else int(ch.a_idx[k - 1])
    hi = len(m.theta) - 1 if k == ch.K - 1 else int(ch.a_idx[k + 1])
    lw = _log_w(ch.ws[k])
    zk = np.ascontiguousarray(z[:, k])
    nb = len(m.betas)
    step = max(1, _GRID_BLOCK_BYTES // (8 * nb * m.n))
    best, best_at = None, (lo, 0)
    for a0 in range(lo, hi + 1, step):
        a1 = min(hi + 1, a0 + step)
        slab = lw + m.tensor[a0:a1]
        slab *= zk
        slab *= m.cnt
        score = np.sum(slab, axis=-1).reshape(-1)
        j = int(np.argmax(score))
        if best is None or score[j] > best:
            best, best_at = score[j], (a0 + j // nb, j % nb)
    return best_at


def _elbo_terms(m: UtrModel, log_z: np.ndarray, z: np.ndarray) -> float:
    """exp_log_lik (apa_core.py:570-573)."""
    zz = np.multiply(z, m.cnt[:, np.newaxis])
    nz = z != 0
    return np.sum(zz[nz] * log_z[nz])


def run_chain(m: UtrModel, ch: Chain, rng, weights_only: bool = False) -> Chain:
    """em_algo (apa_core.py:714-779).  Coordinate-wise: one column of log_z is refreshed per
    iteration, every other column keeps stale parameters; the uniform column is never refreshed."""
    K = ch.K
    lb = SENTINEL
    ch.lb_arr = []
    ch.k_order = draw_component_order(K, N_ROUND, rng)
    log_z = np.zeros((m.n, K + 1))
    for k in range(K + 1):
        _refresh_column(m, ch, k, log_z)
    z = None
    for it in range(N_ROUND):
        k = int(ch.k_order[it])
        _refresh_column(m, ch, k, log_z)
        z = _responsibilities(m, log_z)
        if np.sum(z[:, k]) < 1e-8:                                      # mstep :526-529 / :554-555
            z[:, k] += 1e-8
        ch.ws = _update_weights(m, z)
        if not weights_only:
            ai, bi = _grid_argmax(m, ch, z, k)
            ch.a_idx[k], ch.b_idx[k] = ai, bi
        lb_new = _elbo_terms(m, log_z, z) + np.sum(m.cnt * _sp_stats.entropy(z, axis=1))   # :559-561
        ch.lb_arr.append(lb_new)
        if m.trace is not None:
            m.trace.append(dict(K=K, it=it, k=k, a_idx=ch.a_idx.copy(), b_idx=ch.b_idx.copy(),
                                ws=ch.ws.copy(), lb=float(lb_new)))
        if np.abs(lb_new - lb) < np.abs(1e-6 * lb):                      # :743
            break
        lb = lb_new
    ch.bic = -2 * _elbo_terms(m, log_z, z) + (3 * K + 1) * np.log(m.n)   # cal_bic :702-706
    order = np.argsort(m.theta[ch.a_idx])                               # :768-772 (always identity)
    ch.a_idx, ch.b_idx = ch.a_idx[order], ch.b_idx[order]
    ch.ws[0:K] = ch.ws[order]
    return ch


def best_of_restarts(m: UtrModel, k: int, rng) -> Chain:
    """em_optim0 (apa_core.py:846-871): 10 restarts, first arg-min BIC."""
    bic = np.full(N_TRIAL, np.finfo("f").max)   # NB float32 array (:427, :849): BICs are compared after
    runs = []                                    # rounding to float32, first of the tied minima wins
    for i in range(N_TRIAL):
        runs.append(run_chain(m, draw_chain(m, k, rng), rng))
        bic[i] = runs[i].bic
    return runs[int(np.argmin(bic))]


def prune_and_refit(m: UtrModel, ch: Chain, rng) -> Chain:
    """rm_component + fixed_inference (apa_core.py:832-844, 708-711): single pass, weights-only refit."""
    min_ws = m.prm["min_ws"]
    keep = np.array([i for i in range(ch.K) if not ch.ws[i] < min_ws], dtype=np.int64)
    if len(keep) == ch.K:
        return ch
    slim = Chain(a_idx=ch.a_idx[keep], b_idx=ch.b_idx[keep], ws=None)
    slim.ws = draw_weights(m, slim.K, rng)
    return run_chain(m, slim, rng, weights_only=True)


def hard_labels(m: UtrModel, ch: Chain) -> np.ndarray:
    """get_label (apa_core.py:873-881): full E-step with the final parameters; label K = noise."""
    log_z = np.zeros((m.n, ch.K + 1), dtype="float")
    for k in range(ch.K + 1):
        _refresh_column(m, ch, k, log_z)
    return np.argmax(_responsibilities(m, log_z), axis=1)


# ----------------------------------------------------------------------------------------------
# Orchestration (apa_core.py:883-1035, 1104-1137)
# ----------------------------------------------------------------------------------------------
@dataclass
class FitResult:
    """Field-for-field twin of scape.apa_core.Parameters (apa_core.py:236-258) as left by `infer`."""
    title: str
    alpha_arr: np.ndarray
    beta_arr: np.ndarray
    ws: np.ndarray
    K: int
    L: int
    bic: float
    lb_arr: List[float]
    label_arr: np.ndarray
    gene_info_str: str = "None"
    cb_id_arr: Optional[np.ndarray] = None
    readID_arr: Optional[np.ndarray] = None
    n_frag: int = 0
    n_theta: int = 0
    chains_run: int = 0
    path: list = field(default_factory=list)


def _finish(m: UtrModel, ch: Chain, title: str) -> FitResult:
    return FitResult(title=title,
                     alpha_arr=np.rint(m.theta[ch.a_idx]).astype("int"),
                     beta_arr=m.betas[ch.b_idx].copy(), ws=ch.ws.copy(), K=ch.K, L=m.L,
                     bic=ch.bic, lb_arr=list(ch.lb_arr),
                     label_arr=hard_labels(m, ch)[m.read_to_bin], n_frag=m.n, n_theta=len(m.theta),
                     path=list(m.path))


def _sweep(m: UtrModel, k_max: int, k_min: int, rng) -> Chain:
    """The K loop + selection + pruning of run() (apa_core.py:965-975)."""
    if k_min > k_max:
        raise Exception("n_min_apa=%s n_max_apa=%s, n_max_apa has to be greater than n_min_apa!" % (k_min, k_max))
    if m.prm["max_beta"] < m.prm["beta_step"]:
        raise Exception("max_beta has to be greater than beta_step_size!")
    ks = list(range(k_max, k_min - 1, -1))
    best = [best_of_restarts(m, k, rng) for k in ks]
    bic_k = np.full(len(ks), np.finfo("f").max)  # float32 again (:945)
    for i, c in enumerate(best):
        bic_k[i] = c.bic
    pick = best[int(np.argmin(bic_k))]
    k_sel = pick.K
    out = prune_and_refit(m, pick, rng)
    m.path.append((k_max, k_sel, out.K))
    return out


def resolve_utr_length(x, l, prm) -> int:
    """subsample_run (apa_core.py:995-997)."""
    return max(prm.get("utr_length", -1), max(x) + max(l) + 50)


def fit_utr(x, l, r, pa, rng, re_run_mode=True, trace=None, **params) -> FitResult:
    """subsample_run, normal mode (apa_core.py:984-997, 1019-1035) + ApaModel.run (:930-981)."""
    prm = dict(DEFAULTS)
    prm.update({k: v for k, v in params.items() if k in DEFAULTS})
    prm["utr_length"] = resolve_utr_length(x, l, prm)
    m = build_model(x, l, r, pa, prm)
    m.trace = trace
    m.unif_loglik = uniform_loglik(m)
    m.prof_x, m.prof_y = coverage_profile(m)
    m.peak_idx, m.peak_w = find_profile_peaks(m)
    m.table = theta_table(m, m.theta)
    m.tensor = get_loglik_marginal_tensor(m.theta, m.betas, m.table)
    k_max, k_min = prm["n_max_apa"], prm["n_min_apa"]
    ch = _sweep(m, k_max, k_min, rng)
    n_chain = N_TRIAL * (k_max - k_min + 1)
    while re_run_mode and ch.K == k_max:                                # :1023-1030
        k_min, k_max = k_max, k_max + 2
        ch = _sweep(m, k_max, k_min, rng)
        n_chain += N_TRIAL * 3
    res = _finish(m, ch, "Final Result")
    res.chains_run = n_chain
    return res


def fit_utr_fixed(x, l, r, pa, rng, pre_alpha, pre_beta, pre_L, trace=None, **params) -> FitResult:
    """subsample_run fixed mode (apa_core.py:999-1017) + ApaModel.fixed_run (:883-928)."""
    prm = dict(DEFAULTS)
    prm.update({k: v for k, v in params.items() if k in DEFAULTS})
    prm["utr_length"] = max(resolve_utr_length(x, l, prm), pre_L)       # :1004
    m = build_model(x, l, r, pa, prm)
    m.trace = trace
    pre_alpha = np.asarray(pre_alpha)
    pre_beta = np.asarray(pre_beta)
    b_hi, b_lo = np.max(pre_beta), np.min(pre_beta)
    full = m.theta
    parts = []
    for a in pre_alpha:                                                 # :891-894
        ends = snap_to_grid(full, np.array([a - 3 * b_hi, a + 3 * b_hi]))
        parts.append(full[ends[0]:ends[1]])
    m.theta = np.unique(np.concatenate(parts))                          # :895
    m.betas = np.arange(b_lo, b_hi + prm["beta_step"], prm["beta_step"]) + 0.0   # :896
    m.unif_loglik = uniform_loglik(m)
    m.table = theta_table(m, m.theta)
    m.prof_x, m.prof_y = coverage_profile(m)
    m.peak_idx, m.peak_w = find_profile_peaks(m)
    m.tensor = np.zeros((len(m.theta), len(m.betas), m.n))              # :910-914
    for i, a in enumerate(m.theta):
        for j, b in enumerate(m.betas):
            m.tensor[i][j] = loglik_marginal_lxr_scipy(a, b, m.theta, m.table)
    ch = best_of_restarts(m, len(pre_alpha), rng)                       # :920
    res = _finish(m, ch, "Final Result (subsample run)")
    res.chains_run = N_TRIAL
    return res


def read_chunk(path):
    """Stream the (gene_info_str, DataFrame) tuples of a prepare_input chunk (apa_core.py:1117-1132)."""
    out = []
    with open(path, "rb") as fh:
        while True:
            try:
                out.append(pickle.load(fh))
            except EOFError:
                return out


def infer_chunk(path, seed=1, pre_para=None, limit=None, **params) -> List[FitResult]:
    """_infer_pa + infer (apa_core.py:107-147, 1104-1137): one legacy RNG stream per file."""
    rng = np.random.RandomState(seed)                                   # np.random.seed(1), :125
    results = []
    for gi, df in read_chunk(path)[:limit]:
        cols = (df["x"], df["l"], df["r"], df["pa"])
        if pre_para is not None:
            res = fit_utr_fixed(*cols, rng, pre_para["alpha_arr"], pre_para["beta_arr"], pre_para["L"], **params)
        else:
            res = fit_utr(*cols, rng, **params)
        res.gene_info_str = gi
        res.cb_id_arr = np.array(df["cb_id"])
        res.readID_arr = np.array(df["read_id"])
        results.append(res)
    return results
