/* libscape_b200.so -- C ABI of the B200-native `scape infer_pa` hot path.
 *
 * The reference (chengl7-lab/scape, SCAPE-APA 1.0.4) has no FFI of its own: its hot path sits behind
 * Python seams.  The entry points below are what a binding of that path
 * needs; each one names the reference interface it replaces.  The reference-side ctypes stub a
 * maintainer would add is shown in INTEGRATION.md; the in-repo binding is scape_b200/_lib.py.
 *
 * Conventions: every function returns 0 on success and a negative code on failure
 * (scape_b200_last_error() describes the last failure on the calling thread); all pointers are
 * caller-owned, contiguous host memory, read-only unless documented as output; the library never
 * keeps a caller pointer past the call; one handle per device, a handle is not re-entrant.
 * There is no CPU fallback: without a CUDA device scape_b200_create fails.
 */
#ifndef SCAPE_B200_H
#define SCAPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCAPE_B200_KCAP 15     /* max pA components of one chain (columns = KCAP + 1 with the uniform one) */
#define SCAPE_B200_NROUND 50   /* apa_core.py:422  nround  */
#define SCAPE_B200_NTRIAL 10   /* apa_core.py:847  n_trial */
#define SCAPE_B200_MAX_BETA 64
#define SCAPE_B200_MAX_S 32
#define SCAPE_B200_MAX_SMOOTH 1024

/* Model parameters: the TOML keys of `scape infer_pa` (tutorial/default_config.toml,
 * ApaModel.__init__ apa_core.py:333-363) plus the small tables the Python host derives from them
 * with numpy so that they are bit-identical to the reference's (apa_core.py:394-396, 684-686, 942). */
typedef struct scape_b200_params {
  int32_t n_max_apa, n_min_apa;
  int64_t utr_length;
  double min_LA, max_LA, mu_f, sigma_f;
  double min_pa_gap, max_beta;
  int32_t theta_step, beta_step;
  double min_ws, max_unif_ws;
  int32_t re_run_mode;          /* subsample_run(re_run_mode=...), apa_core.py:984,1023 */
  int32_t fixed_run_mode;       /* --pre_para_pkl_file, apa_core.py:94-99, 999-1017      */
  int32_t pre_K;                /* fixed mode: first Parameters object of the pre_para file */
  int32_t _pad0;
  int64_t pre_L;
  double pre_alpha[SCAPE_B200_KCAP];
  double pre_beta[SCAPE_B200_KCAP];
  int32_t n_beta, n_s, n_smooth, _pad1;
  double betas[SCAPE_B200_MAX_BETA];      /* predef_beta_arr (apa_core.py:942 / :896)              */
  double s_dis[SCAPE_B200_MAX_S];         /* polyA length grid (apa_core.py:394)                   */
  double pmf_s[SCAPE_B200_MAX_S];         /* its pmf (apa_core.py:395-396)                         */
  double smooth_w[SCAPE_B200_MAX_SMOOTH]; /* exp(-arange(-3bw,3bw+1)^2 / (2 bw^2)) (apa_core.py:684-685) */
} scape_b200_params;

/* One batch of UTRs = what `infer` streams out of chunk pickles (apa_core.py:1117-1132), with the
 * DataFrame columns x, l, r, pa concatenated over UTRs (CSR by read_off).  Each UTR belongs to one
 * RNG stream; a stream is seeded once (np.random.seed(1) per chunk file, apa_core.py:125) and its
 * UTRs are consumed in batch order. */
typedef struct scape_b200_batch {
  int64_t n_utr;
  const int64_t* read_off;     /* [n_utr + 1] */
  const double* x;             /* [read_off[n_utr]]  read start (integer valued)            */
  const double* l;             /*                     UTR part length of the read            */
  const double* r;             /*                     polyA length, NaN = unknown            */
  const double* pa;            /*                     junction pA site, NaN = not a junction */
  int32_t n_streams;
  int32_t _pad0;
  const int32_t* stream_id;    /* [n_utr], values in [0, n_streams) */
  const uint32_t* stream_seed; /* [n_streams] np.random.seed(seed) per stream; ignored if stream_state != NULL */
  uint32_t* stream_state;      /* optional in/out [n_streams * 625]: MT19937 key[624] + pos, i.e.
                                  np.random.get_state()[1:3]; on return holds the state the reference
                                  would have left behind, so callers can np.random.set_state() it */
} scape_b200_batch;

/* Per-UTR results = the fields of scape.apa_core.Parameters (apa_core.py:236-258) that `infer`
 * pickles.  All arrays are caller-allocated. */
typedef struct scape_b200_results {
  int32_t* status;   /* [n_utr] 0 = ok, <0 = the reference would have raised (see host_prep.hpp)  */
  int32_t* K;        /* [n_utr]                                                                   */
  int64_t* L;        /* [n_utr]                                                                   */
  double* alpha;     /* [n_utr * KCAP]       sorted pA sites (bp from the UTR 5' end)             */
  double* beta;      /* [n_utr * KCAP]                                                            */
  double* ws;        /* [n_utr * (KCAP + 1)] last used entry = uniform component                  */
  double* bic;       /* [n_utr]                                                                   */
  int32_t* n_lb;     /* [n_utr]              length of lb_arr                                     */
  double* lb_arr;    /* [n_utr * NROUND]                                                          */
  int64_t* label;    /* [read_off[n_utr]]    per-read hard label, K = noise (apa_core.py:873-881,976) */
  int32_t* n_frag;   /* [n_utr]              number of bins N (apa_core.py:379)                   */
  int32_t* n_theta;  /* [n_utr]              theta grid size T                                    */
  int32_t* path;     /* [n_utr * 4]          sweeps run, K selected by BIC, K after pruning, chains run */
  double* em_work;   /* [n_utr * 2]          sum over chains/iterations of N*(K+1); total EM iterations */
} scape_b200_results;

/* Timing of the last fit_batch.  Kernel durations come from CUDA events on the library's own
 * streams (one per lane); device_busy_ms is the length of the UNION of all kernel intervals, i.e.
 * the time the GPU was executing this library's kernels (lanes overlap, so the per-phase sums can
 * exceed it). */
typedef struct scape_b200_timing {
  double table_ms, tensor_ms, em_ms, label_ms;    /* summed kernel durations (all lanes)      */
  double host_prep_ms, host_rng_ms, device_busy_ms, d2h_ms, total_ms; /* wall clock, except device_busy_ms */
  int64_t launches;                               /* kernels launched by this library          */
  int64_t waves;
  double em_grid_bytes;                           /* algorithmic tensor bytes read by the EM grid search */
  double em_grid_flops;
  double tensor_exp;                              /* exp() evaluations of the marginal kernel  */
  double h2d_bytes, d2h_bytes;
  double em_scan_bytes;                           /* tensor bytes the grid search actually loads (fragment hull only) */
  double estep_ms, scan_ms;                       /* em_ms split: E-step kernels / arg-max scan kernels */
  int64_t scan_launches;
  double table_exp;                               /* exp() evaluations of the theta-table kernel (N T S per UTR) */
  double resident_ms;                             /* em_ms spent in the resident EM kernels (chain-resident tail / cluster-resident): E passes + grid search, many iterations per launch */
  double resident_grid_flops;                     /* algorithmic grid-search flops (2 W_k B N per chain iteration) done inside those kernels */
  int64_t resident_launches;
} scape_b200_timing;

typedef struct scape_b200_handle scape_b200_handle;

const char* scape_b200_last_error(void);
int scape_b200_version(void);
int scape_b200_device_count(void);

/* ApaModel(**kwargs) lifetime: parameters are fixed per handle. */
int scape_b200_create(int device, const scape_b200_params* params, scape_b200_handle** out);
int scape_b200_destroy(scape_b200_handle* h);

/* Replaces the body of `infer` (apa_core.py:1104-1137): subsample_run (+ run / fixed_run,
 * rm_component, re-run loop, get_label) for every UTR of the batch. */
int scape_b200_fit_batch(scape_b200_handle* h, const scape_b200_batch* batch, scape_b200_results* out);
int scape_b200_get_timing(scape_b200_handle* h, scape_b200_timing* out);

/* FP64 peak microbenchmarks on the handle's device (TFLOP/s, 2 flop per FMA): CUDA-core DFMA stream
 * and tensor-core DMMA (mma.m8n8k4.f64) stream.  They are the roofline denominators of the EM
 * kernels; MEASURED_PEAKS.json carries no FP64 figure. */
int scape_b200_fp64_peaks(scape_b200_handle* h, double* dfma_tflops, double* dmma_tflops);
/* out4 = {FP32 FMA TFLOP/s, MUFU ex2.approx Gop/s, FP64 exp() Gop/s, FP64 log() Gop/s} of this GPU, measured
 * with register-resident instruction streams: the roofline denominators of the likelihood phases
 * (loglik_xlr_t: 13 exp per table entry, taichi_core.py:141-157; get_loglik_marginal_tensor: 307 exp
 * per (fragment, alpha), taichi_core.py:160-179). */
int scape_b200_sfu_peaks(scape_b200_handle* h, double* out4);

/* Storage type of the marginal tensor in HBM: 4 = float (default; values are computed in FP64 and
 * rounded once, every sum / product stays FP64), 8 = double (strict mode).  Environment override at
 * create time: SCAPE_B200_TENSOR=f64. */
int scape_b200_set_tensor_dtype(scape_b200_handle* h, int bytes);

/* Wave pipelining: 1 (default) = the likelihood phase (uploads, table, marginal tensor) of wave w+1
 * runs on a second stream under the EM of wave w; 0 = strictly one phase at a time (per-kernel
 * timings are then taken with the kernel alone on the GPU).  Results are identical either way.
 * Environment override at create time: SCAPE_B200_OVERLAP=0. */
int scape_b200_set_overlap(scape_b200_handle* h, int on);

/* Host threads this handle may use for its pre-pass / RNG-replay pools (0 = default: the CPUs of the
 * process divided by LOCAL_WORLD_SIZE).  A process that drives several GPUs gives every handle its
 * share (scape_b200.apa_core.infer_files(devices=[...])); the reference itself is single-threaded
 * (apa_core.py:1104-1137).  Environment override at create time: SCAPE_B200_THREADS. */
int scape_b200_set_host_threads(scape_b200_handle* h, int n);

/* ---- kernel-seam entry points (parity tests; reference seam B3, apa_core.py:23) ------------- */

/* loglik_xlr_t over a theta list (apa_core.py:620-640, taichi_core.py:183-215):
 * fragments in, table[n_frag * n_theta] (row-major [n][t]) out. */
int scape_b200_loglik_table(scape_b200_handle* h, int64_t n_frag, const double* x, const double* l,
                            const double* r, const double* pa, int64_t n_theta, const double* theta,
                            double* table_out);

/* get_loglik_marginal_tensor (taichi_core.py:237-246): table[n][t] in, tensor[t][b][n] out. */
int scape_b200_marginal_tensor(scape_b200_handle* h, int64_t n_frag, int64_t n_theta, const double* theta,
                               int64_t n_beta, const double* betas, const double* table,
                               double* tensor_out);

/* em_algo for explicit initial chains on one UTR's tensor (apa_core.py:714-779): inputs are the
 * init_para blobs; outputs per chain: alpha/beta grid indices, ws, bic, lb_arr, n_iter and the
 * per-iteration trace (alpha idx, beta idx, ws after each iteration). */
typedef struct scape_b200_chain_io {
  int32_t K;
  int32_t weights_only;                 /* fixed_inference_flag (apa_core.py:735-736) */
  int32_t a_idx[SCAPE_B200_KCAP];
  int32_t b_idx[SCAPE_B200_KCAP];
  double ws[SCAPE_B200_KCAP + 1];
  uint8_t k_order[SCAPE_B200_NROUND + 6];
  /* outputs */
  int32_t n_iter;
  int32_t _pad;
  double bic;
  double lb_arr[SCAPE_B200_NROUND];
} scape_b200_chain_io;

int scape_b200_em_chains(scape_b200_handle* h, int64_t n_frag, int64_t n_theta, int64_t n_beta,
                         const double* tensor, const double* cnt, double unif_loglik,
                         int64_t n_chains, scape_b200_chain_io* chains,
                         int32_t* trace_a, int32_t* trace_b, double* trace_ws /* optional, may be NULL */);

/* ---- host pre-pass entry points (no GPU needed; CPU tests) --------------------------------- */

/* bin_data (apa_core.py:285-327).  Outputs sized n_reads; returns the number of bins in *n_bins. */
int scape_b200_bin_reads(int64_t n_reads, const double* x, const double* l, const double* r, const double* pa,
                         double* bx, double* bl, double* br, double* bpa, double* cnt, int32_t* read_to_bin,
                         int64_t* n_bins);

/* Model set-up + coverage profile + peaks for one UTR (apa_core.py:365-462, 681-700, 784-794).
 * prof_y must hold L + 200 doubles (query with prof_y == NULL first: *L_out is still written). */
int scape_b200_profile(const scape_b200_params* params, int64_t n_reads, const double* x, const double* l,
                       const double* r, const double* pa, int64_t* L_out, int64_t* n_theta, double* theta,
                       double* prof_y, int64_t* n_peaks, int64_t* peak_idx, double* peak_w, int64_t cap);

/* Replay of the reference's initialisation draws on one UTR (apa_core.py:817-829, 720):
 * n_chains chains with K = ks[i] drawn back to back from a stream seeded with `seed`,
 * after skipping `skip_u32` 32-bit draws. */
int scape_b200_draw_chains(const scape_b200_params* params, int64_t n_reads, const double* x, const double* l,
                           const double* r, const double* pa, uint32_t seed, int64_t n_chains,
                           const int32_t* ks, scape_b200_chain_io* out);

/* Process-wide hook: order of np.argsort(values) as the installed numpy computes it.  Called (from
 * worker threads) only when two candidate coverage peaks have exactly equal height, because
 * scipy.signal.find_peaks ranks peaks with an unstable np.argsort whose tie order is build / CPU
 * specific (apa_core.py:784).  NULL (default) = stable order. */
typedef void (*scape_b200_argsort_fn)(const double* values, int64_t n, int64_t* order_out);
int scape_b200_set_argsort_callback(scape_b200_argsort_fn fn);

/* Raw generator access for the RNG tests: fills out[n] with RandomState(seed).random_sample(n)
 * (kind 0), .randint(0, arg, n) (kind 1), .permutation(arg) (kind 2, n == arg), or the n
 * .random_sample() values that FOLLOW a .permutation(arg) (kind 3: state continuity). */
int scape_b200_rng_draw(uint32_t seed, int kind, int64_t arg, int64_t n, double* out);

#ifdef __cplusplus
}
#endif
#endif /* SCAPE_B200_H */
