"""ctypes binding of libscape_b200.so (include/scape_b200.h).  Thin: structs, argument marshalling,
error translation.  No compute happens in this file and there is no fallback: if the shared
library is missing the import of the product path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

KCAP = 15
NROUND = 50
NTRIAL = 10
MAX_BETA = 64
MAX_S = 32
MAX_SMOOTH = 1024

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscape_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint32_p = C.POINTER(C.c_uint32)


class Params(C.Structure):
    _fields_ = [
        ("n_max_apa", C.c_int32), ("n_min_apa", C.c_int32),
        ("utr_length", C.c_int64),
        ("min_LA", C.c_double), ("max_LA", C.c_double), ("mu_f", C.c_double), ("sigma_f", C.c_double),
        ("min_pa_gap", C.c_double), ("max_beta", C.c_double),
        ("theta_step", C.c_int32), ("beta_step", C.c_int32),
        ("min_ws", C.c_double), ("max_unif_ws", C.c_double),
        ("re_run_mode", C.c_int32), ("fixed_run_mode", C.c_int32),
        ("pre_K", C.c_int32), ("_pad0", C.c_int32),
        ("pre_L", C.c_int64),
        ("pre_alpha", C.c_double * KCAP), ("pre_beta", C.c_double * KCAP),
        ("n_beta", C.c_int32), ("n_s", C.c_int32), ("n_smooth", C.c_int32), ("_pad1", C.c_int32),
        ("betas", C.c_double * MAX_BETA), ("s_dis", C.c_double * MAX_S), ("pmf_s", C.c_double * MAX_S),
        ("smooth_w", C.c_double * MAX_SMOOTH),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_utr", C.c_int64), ("read_off", c_int64_p),
        ("x", c_double_p), ("l", c_double_p), ("r", c_double_p), ("pa", c_double_p),
        ("n_streams", C.c_int32), ("_pad0", C.c_int32),
        ("stream_id", c_int32_p), ("stream_seed", c_uint32_p), ("stream_state", c_uint32_p),
    ]


class Results(C.Structure):
    _fields_ = [
        ("status", c_int32_p), ("K", c_int32_p), ("L", c_int64_p),
        ("alpha", c_double_p), ("beta", c_double_p), ("ws", c_double_p), ("bic", c_double_p),
        ("n_lb", c_int32_p), ("lb_arr", c_double_p), ("label", c_int64_p),
        ("n_frag", c_int32_p), ("n_theta", c_int32_p), ("path", c_int32_p), ("em_work", c_double_p),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("table_ms", C.c_double), ("tensor_ms", C.c_double), ("em_ms", C.c_double), ("label_ms", C.c_double),
        ("host_prep_ms", C.c_double), ("host_rng_ms", C.c_double), ("device_busy_ms", C.c_double), ("d2h_ms", C.c_double),
        ("total_ms", C.c_double),
        ("launches", C.c_int64), ("waves", C.c_int64),
        ("em_grid_bytes", C.c_double), ("em_grid_flops", C.c_double), ("tensor_exp", C.c_double),
        ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double), ("em_scan_bytes", C.c_double),
        ("estep_ms", C.c_double), ("scan_ms", C.c_double), ("scan_launches", C.c_int64),
        ("table_exp", C.c_double), ("resident_ms", C.c_double), ("resident_grid_flops", C.c_double), ("resident_launches", C.c_int64),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class ChainIO(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("weights_only", C.c_int32),
        ("a_idx", C.c_int32 * KCAP), ("b_idx", C.c_int32 * KCAP),
        ("ws", C.c_double * (KCAP + 1)),
        ("k_order", C.c_uint8 * (NROUND + 6)),
        ("n_iter", C.c_int32), ("_pad", C.c_int32),
        ("bic", C.c_double),
        ("lb_arr", C.c_double * NROUND),
    ]


EXPORTS = [
    "scape_b200_last_error", "scape_b200_version", "scape_b200_device_count", "scape_b200_create",
    "scape_b200_destroy", "scape_b200_fit_batch", "scape_b200_get_timing", "scape_b200_loglik_table",
    "scape_b200_marginal_tensor", "scape_b200_em_chains", "scape_b200_bin_reads", "scape_b200_profile",
    "scape_b200_draw_chains", "scape_b200_rng_draw", "scape_b200_set_argsort_callback",
    "scape_b200_set_tensor_dtype", "scape_b200_set_overlap", "scape_b200_set_host_threads", "scape_b200_fp64_peaks", "scape_b200_sfu_peaks",
]

ARGSORT_FN = C.CFUNCTYPE(None, c_double_p, C.c_int64, c_int64_p)


@ARGSORT_FN
def _numpy_argsort(values, n, out):
    # scipy.signal.find_peaks ranks peaks with np.argsort (unstable, build-specific tie order)
    order = np.argsort(np.ctypeslib.as_array(values, shape=(n,)))
    np.ctypeslib.as_array(out, shape=(n,))[:] = order


_lib = None


class ScapeB200Error(RuntimeError):
    pass


def load():
    """dlopen libscape_b200.so (built in-tree by scape_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScapeB200Error(
            f"{LIB_PATH} is missing: build it with `python -m scape_b200.build` "
            "(nvcc, sm_100a).  There is no CPU / Taichi fallback for infer_pa.")
    lib = C.CDLL(LIB_PATH)
    lib.scape_b200_last_error.restype = C.c_char_p
    lib.scape_b200_create.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(C.c_void_p)]
    lib.scape_b200_destroy.argtypes = [C.c_void_p]
    lib.scape_b200_fit_batch.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(Results)]
    lib.scape_b200_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
    lib.scape_b200_set_tensor_dtype.argtypes = [C.c_void_p, C.c_int]
    lib.scape_b200_set_overlap.argtypes = [C.c_void_p, C.c_int]
    lib.scape_b200_set_host_threads.argtypes = [C.c_void_p, C.c_int]
    lib.scape_b200_fp64_peaks.argtypes = [C.c_void_p, c_double_p, c_double_p]
    lib.scape_b200_sfu_peaks.argtypes = [C.c_void_p, c_double_p]
    lib.scape_b200_loglik_table.argtypes = [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p,
                                            C.c_int64, c_double_p, c_double_p]
    lib.scape_b200_marginal_tensor.argtypes = [C.c_void_p, C.c_int64, C.c_int64, c_double_p, C.c_int64, c_double_p,
                                               c_double_p, c_double_p]
    lib.scape_b200_em_chains.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, c_double_p, c_double_p,
                                         C.c_double, C.c_int64, C.POINTER(ChainIO), c_int32_p, c_int32_p, c_double_p]
    lib.scape_b200_bin_reads.argtypes = [C.c_int64] + [c_double_p] * 9 + [c_int32_p, c_int64_p]
    lib.scape_b200_profile.argtypes = [C.POINTER(Params), C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p,
                                       c_int64_p, c_int64_p, c_double_p, c_double_p, c_int64_p, c_int64_p,
                                       c_double_p, C.c_int64]
    lib.scape_b200_draw_chains.argtypes = [C.POINTER(Params), C.c_int64, c_double_p, c_double_p, c_double_p,
                                           c_double_p, C.c_uint32, C.c_int64, c_int32_p, C.POINTER(ChainIO)]
    lib.scape_b200_rng_draw.argtypes = [C.c_uint32, C.c_int, C.c_int64, C.c_int64, c_double_p]
    lib.scape_b200_set_argsort_callback.argtypes = [ARGSORT_FN]
    lib.scape_b200_set_argsort_callback(_numpy_argsort)
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise ScapeB200Error(f"libscape_b200 error {rc}: {load().scape_b200_last_error().decode()}")


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


DEFAULTS = dict(  # tutorial/default_config.toml == ApaModel.__init__ defaults (apa_core.py:333-363)
    n_max_apa=5, n_min_apa=1, utr_length=2000, min_LA=20, max_LA=150, mu_f=300, sigma_f=50,
    min_pa_gap=100, max_beta=70, theta_step=9, beta_step=5, min_ws=0.05, max_unif_ws=0.15,
    re_run_mode=True,
)


def make_params(pre_para=None, **kwargs) -> Params:
    """TOML / kwargs -> scape_b200_params.  The small derived tables are computed here with the
    same numpy expressions the reference uses so they are bit-identical (apa_core.py:394-396,
    684-685, 942, 896)."""
    kw = dict(DEFAULTS)
    kw.update({k: v for k, v in kwargs.items() if k in DEFAULTS})
    p = Params()
    p.n_max_apa, p.n_min_apa = int(kw["n_max_apa"]), int(kw["n_min_apa"])
    p.utr_length = int(kw["utr_length"])
    p.min_LA, p.max_LA, p.mu_f, p.sigma_f = float(kw["min_LA"]), float(kw["max_LA"]), float(kw["mu_f"]), float(kw["sigma_f"])
    p.min_pa_gap, p.max_beta = float(kw["min_pa_gap"]), float(kw["max_beta"])
    if int(kw["beta_step"]) != kw["beta_step"]:
        raise ValueError("beta_step must be an integer (the reference slices arrays with it, apa_core.py:788-793)")
    p.theta_step, p.beta_step = int(kw["theta_step"]), int(kw["beta_step"])
    p.min_ws, p.max_unif_ws = float(kw["min_ws"]), float(kw["max_unif_ws"])
    p.re_run_mode = 1 if kw["re_run_mode"] else 0
    if p.n_min_apa > p.n_max_apa:                       # apa_core.py:931-933
        raise Exception("n_min_apa=" + str(p.n_min_apa) + " n_max_apa=" + str(p.n_max_apa) +
                        ", n_max_apa has to be greater than n_min_apa!")
    if p.max_beta < p.beta_step:                        # apa_core.py:935-937
        raise Exception("max_beta=" + str(kw["max_beta"]) + " beta_step_size=" + str(kw["beta_step"]) +
                        ", max_beta has to be greater than beta_step_size!")
    s_dis = np.arange(kw["min_LA"], kw["max_LA"], 10)
    pmf = np.repeat(1 / len(s_dis), len(s_dis))
    pmf = pmf / sum(pmf)
    if pre_para is None:
        betas = np.arange(kw["beta_step"], kw["max_beta"], kw["beta_step"]) + 0.0
        p.fixed_run_mode = 0
    else:
        b = np.asarray(pre_para.beta_arr, dtype=float)
        betas = np.arange(np.min(b), np.max(b) + kw["beta_step"], kw["beta_step"]) + 0.0
        p.fixed_run_mode = 1
        p.pre_K = len(pre_para.alpha_arr)
        if p.pre_K > KCAP:
            raise ValueError(f"pre_para has {p.pre_K} pA sites; libscape_b200 supports at most {KCAP}")
        p.pre_L = int(pre_para.L)
        for i in range(p.pre_K):
            p.pre_alpha[i] = float(pre_para.alpha_arr[i])
            p.pre_beta[i] = float(b[i])
    bw = kw["beta_step"] * 3
    w = np.exp(-np.arange(-3 * bw, 3 * bw + 1) ** 2 / (2 * bw * bw))
    for name, arr, cap in (("betas", betas, MAX_BETA), ("s_dis", s_dis, MAX_S), ("pmf_s", pmf, MAX_S),
                           ("smooth_w", w, MAX_SMOOTH)):
        if len(arr) > cap or len(arr) == 0:
            raise ValueError(f"{name}: {len(arr)} entries (supported: 1..{cap})")
        dst = getattr(p, name)
        for i, v in enumerate(arr):
            dst[i] = float(v)
    p.n_beta, p.n_s, p.n_smooth = len(betas), len(s_dis), len(w)
    return p


class FitOutput:
    """Per-UTR result arrays of one fit_batch call (host memory)."""

    def __init__(self, n_utr, n_reads):
        self.status = np.zeros(n_utr, np.int32)
        self.K = np.zeros(n_utr, np.int32)
        self.L = np.zeros(n_utr, np.int64)
        self.alpha = np.zeros((n_utr, KCAP), np.float64)
        self.beta = np.zeros((n_utr, KCAP), np.float64)
        self.ws = np.zeros((n_utr, KCAP + 1), np.float64)
        self.bic = np.zeros(n_utr, np.float64)
        self.n_lb = np.zeros(n_utr, np.int32)
        self.lb_arr = np.zeros((n_utr, NROUND), np.float64)
        self.label = np.zeros(n_reads, np.int64)
        self.n_frag = np.zeros(n_utr, np.int32)
        self.n_theta = np.zeros(n_utr, np.int32)
        self.path = np.zeros((n_utr, 4), np.int32)
        self.em_work = np.zeros((n_utr, 2), np.float64)
        self.timing = {}

    def _struct(self):
        r = Results()
        for f, t in Results._fields_:
            setattr(r, f, getattr(self, f).ctypes.data_as(t))
        return r


class Engine:
    """One handle = one GPU + one parameter set (ApaModel(**kwargs) lifetime)."""

    def __init__(self, params: Params, device: int = 0, tensor_dtype: Optional[str] = None):
        """tensor_dtype: "f32" (default) or "f64" storage of the marginal tensor; arithmetic is FP64."""
        self._lib = load()
        self._h = C.c_void_p()
        self.params = params
        _check(self._lib.scape_b200_create(device, C.byref(params), C.byref(self._h)))
        if tensor_dtype is not None:
            _check(self._lib.scape_b200_set_tensor_dtype(self._h, {"f32": 4, "f64": 8}[tensor_dtype]))

    def set_host_threads(self, n: int):
        """Host threads of this handle's pre-pass / RNG-replay pools (0 = process default)."""
        _check(self._lib.scape_b200_set_host_threads(self._h, int(n)))

    def set_overlap(self, on: bool):
        """Pipeline the likelihood phase of the next wave under the EM of the current one (default on)."""
        _check(self._lib.scape_b200_set_overlap(self._h, int(bool(on))))

    def close(self):
        if self._h:
            self._lib.scape_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def fit(self, read_off, x, l, r, pa, stream_id, stream_seed=None, stream_state=None) -> FitOutput:
        """stream_state: optional uint32[n_streams, 625] (MT19937 key + pos), updated in place."""
        read_off = np.ascontiguousarray(read_off, np.int64)
        x, l, r, pa = _f64(x), _f64(l), _f64(r), _f64(pa)
        stream_id = np.ascontiguousarray(stream_id, np.int32)
        if stream_state is not None:
            assert stream_state.dtype == np.uint32 and stream_state.flags.c_contiguous and stream_state.shape[1] == 625
            stream_seed = np.zeros(stream_state.shape[0], np.uint32)
        stream_seed = np.ascontiguousarray(stream_seed, np.uint32)
        n_utr = len(read_off) - 1
        assert len(stream_id) == n_utr and len(x) == read_off[-1]
        b = Batch()
        b.n_utr = n_utr
        b.read_off = read_off.ctypes.data_as(c_int64_p)
        b.x, b.l, b.r, b.pa = _dp(x), _dp(l), _dp(r), _dp(pa)
        b.n_streams = len(stream_seed)
        b.stream_id = stream_id.ctypes.data_as(c_int32_p)
        b.stream_seed = stream_seed.ctypes.data_as(c_uint32_p)
        b.stream_state = stream_state.ctypes.data_as(c_uint32_p) if stream_state is not None else None
        out = FitOutput(n_utr, int(read_off[-1]))
        res = out._struct()
        _check(self._lib.scape_b200_fit_batch(self._h, C.byref(b), C.byref(res)))
        out.timing = self.timing()
        return out

    def fp64_peaks(self) -> dict:
        """Measured FP64 TFLOP/s of this GPU: CUDA-core DFMA stream and tensor-core DMMA stream."""
        a, b = C.c_double(), C.c_double()
        _check(self._lib.scape_b200_fp64_peaks(self._h, C.byref(a), C.byref(b)))
        return {"dfma_tflops": a.value, "dmma_tflops": b.value}

    def sfu_peaks(self) -> dict:
        """Measured FP32 FMA TFLOP/s, MUFU ex2 Gop/s and FP64 exp() / log() Gop/s of this GPU."""
        out = np.zeros(4)
        _check(self._lib.scape_b200_sfu_peaks(self._h, _dp(out)))
        return {"ffma_tflops": out[0], "mufu_ex2_gops": out[1], "exp_f64_gops": out[2], "log_f64_gops": out[3]}

    def timing(self) -> dict:
        t = Timing()
        _check(self._lib.scape_b200_get_timing(self._h, C.byref(t)))
        return t.as_dict()

    # ---- kernel-seam entry points (parity tests) ------------------------------------------------
    def loglik_table(self, x, l, r, pa, theta) -> np.ndarray:
        x, l, r, pa, theta = _f64(x), _f64(l), _f64(r), _f64(pa), _f64(theta)
        out = np.empty((len(x), len(theta)))
        _check(self._lib.scape_b200_loglik_table(self._h, len(x), _dp(x), _dp(l), _dp(r), _dp(pa), len(theta),
                                                 _dp(theta), _dp(out)))
        return out

    def marginal_tensor(self, theta, betas, table) -> np.ndarray:
        theta, betas, table = _f64(theta), _f64(betas), _f64(table)
        n = table.shape[0]
        out = np.empty((len(theta), len(betas), n))
        _check(self._lib.scape_b200_marginal_tensor(self._h, n, len(theta), _dp(theta), len(betas), _dp(betas),
                                                    _dp(table), _dp(out)))
        return out

    def em_chains(self, tensor, cnt, unif_loglik, chains: Sequence[dict], trace=False):
        """chains: dicts with K, a_idx, b_idx, ws, k_order[, weights_only].  Returns the filled
        ChainIO array (+ per-iteration traces)."""
        tensor, cnt = _f64(tensor), _f64(cnt)
        T, B, N = tensor.shape
        arr = (ChainIO * len(chains))()
        for io, c in zip(arr, chains):
            K = int(c["K"])
            io.K = K
            io.weights_only = int(c.get("weights_only", 0))
            for i in range(K):
                io.a_idx[i] = int(c["a_idx"][i])
                io.b_idx[i] = int(c["b_idx"][i])
            for i in range(K + 1):
                io.ws[i] = float(c["ws"][i])
            for i in range(NROUND):
                io.k_order[i] = int(c["k_order"][i])
        ta = tb = tw = None
        if trace:
            shape = (len(chains), NROUND, KCAP + 1)
            ta, tb, tw = np.zeros(shape, np.int32), np.zeros(shape, np.int32), np.zeros(shape, np.float64)
        _check(self._lib.scape_b200_em_chains(
            self._h, N, T, B, _dp(tensor), _dp(cnt), float(unif_loglik), len(chains), arr,
            ta.ctypes.data_as(c_int32_p) if trace else None, tb.ctypes.data_as(c_int32_p) if trace else None,
            _dp(tw) if trace else None))
        return arr, (ta, tb, tw)


# ---- host pre-pass entry points (CPU only) --------------------------------------------------------
def bin_reads(x, l, r, pa):
    lib = load()
    x, l, r, pa = _f64(x), _f64(l), _f64(r), _f64(pa)
    n = len(x)
    outs = [np.empty(n) for _ in range(5)]
    inv = np.empty(n, np.int32)
    nb = C.c_int64()
    _check(lib.scape_b200_bin_reads(n, _dp(x), _dp(l), _dp(r), _dp(pa), *[_dp(o) for o in outs],
                                    inv.ctypes.data_as(c_int32_p), C.byref(nb)))
    k = nb.value
    return tuple(o[:k] for o in outs) + (inv,)


def profile(params: Params, x, l, r, pa):
    lib = load()
    x, l, r, pa = _f64(x), _f64(l), _f64(r), _f64(pa)
    L, T, npk = C.c_int64(), C.c_int64(), C.c_int64()
    _check(lib.scape_b200_profile(C.byref(params), len(x), _dp(x), _dp(l), _dp(r), _dp(pa), C.byref(L), C.byref(T),
                                  None, None, C.byref(npk), None, None, 0))
    cap = max(L.value + 200, T.value)
    theta, prof = np.empty(cap), np.empty(cap)
    pk, pw = np.empty(max(npk.value, 1), np.int64), np.empty(max(npk.value, 1))
    _check(lib.scape_b200_profile(C.byref(params), len(x), _dp(x), _dp(l), _dp(r), _dp(pa), C.byref(L), C.byref(T),
                                  _dp(theta), _dp(prof), C.byref(npk), pk.ctypes.data_as(c_int64_p), _dp(pw), cap))
    return dict(L=L.value, theta=theta[:T.value], prof_y=prof[:L.value + 200], peak_idx=pk[:npk.value],
                peak_w=pw[:npk.value])


def draw_chains(params: Params, x, l, r, pa, seed: int, ks: Sequence[int]):
    lib = load()
    x, l, r, pa = _f64(x), _f64(l), _f64(r), _f64(pa)
    ks = np.ascontiguousarray(ks, np.int32)
    arr = (ChainIO * len(ks))()
    _check(lib.scape_b200_draw_chains(C.byref(params), len(x), _dp(x), _dp(l), _dp(r), _dp(pa), seed, len(ks),
                                      ks.ctypes.data_as(c_int32_p), arr))
    return arr


def rng_draw(seed: int, kind: int, arg: int, n: int) -> np.ndarray:
    out = np.empty(n)
    _check(load().scape_b200_rng_draw(seed, kind, arg, n, _dp(out)))
    return out
