// libscape_b200.so: C ABI (include/scape_b200.h) + the wave scheduler that drives the kernels.
//
// Replaces the body of the reference's `infer` loop (apa_core.py:1104-1137): for every UTR of a
// batch it does what subsample_run -> ApaModel.run / fixed_run do (apa_core.py:883-1035), with the
// likelihood phases and every EM chain on the GPU and the (tiny, serial) model-selection logic and
// the numpy-legacy RNG replay on the host.
//
// Scheduling: the reference seeds the global RNG once per chunk file and the amount of randomness
// a UTR consumes depends on its own result (rm_component :843, re-run loop :1023-1030), so the
// UTRs of one stream form a serial chain at the RNG level.  A *wave* therefore takes the next
// unfinished UTR of every stream; all chains of all UTRs of a wave run concurrently.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <functional>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <memory>
#include <thread>
#include <vector>

#include "host_prep.hpp"
#include "kernels.cuh"
#include "work_pool.hpp"

using namespace scape;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(-100, std::string(#call) + ": " + cudaGetErrorString(e__));                      \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = n + n / 4 + 64;
    cudaError_t e = cudaMalloc((void**)&p, cap * sizeof(T));
    if (e != cudaSuccess) cap = 0;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// Page-locked host staging that keeps its capacity across waves (no page faults, fast DMA).
template <class T>
struct PinnedBuf {
  T* p = nullptr;
  size_t cap = 0, n = 0;
  cudaError_t resize(size_t m) {
    if (m > cap) {
      T* q = nullptr;
      const size_t c = m + m / 2 + 64;
      cudaError_t e = cudaMallocHost((void**)&q, c * sizeof(T));
      if (e != cudaSuccess) return e;
      if (p) {
        memcpy(q, p, n * sizeof(T));
        cudaFreeHost(p);
      }
      p = q;
      cap = c;
    }
    n = m;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = n = 0;
  }
};

// Device-side state of one lane.  A fit splits its RNG streams over the lanes; every lane runs its
// own wave loop in its own host thread on its own CUDA stream, so the host work of one lane (RNG
// replay, model selection, result assembly) overlaps the kernels of the others.
// Likelihood-phase arrays of the wave AFTER the one whose EM is running: the table and the marginal
// tensor depend on the data only (no RNG), so they are produced one wave ahead on a second,
// low-priority stream and fill the SMs the latency-bound EM steps leave idle.
struct StagedBufs {
  DevBuf<double> d_fx, d_fl, d_fr, d_fpa, d_cnt, d_theta, d_table, d_tensor;
  DevBuf<UtrDev> d_utrs;
  DevBuf<RowRef> d_rows, d_trows;
  DevBuf<TileRef> d_tiles;
  DevBuf<int32_t> d_r2b;                             // read -> fragment map of the wave's UTRs (label expansion)
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};   // before table, before tensor, after tensor
  void release() {
    d_fx.release(); d_fl.release(); d_fr.release(); d_fpa.release(); d_cnt.release(); d_theta.release();
    d_table.release(); d_tensor.release(); d_utrs.release(); d_rows.release(); d_trows.release(); d_tiles.release();
    d_r2b.release();
  }
};

struct Lane {
  cudaStream_t st = nullptr;       // EM, labels (high priority)
  cudaStream_t st_lik = nullptr;   // uploads, table, tensor of the staged wave (low priority)
  cudaStream_t st_aux[3] = {nullptr, nullptr, nullptr};   // parts 2..4 of a split bulk-synchronous wave (high priority)
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  cudaStream_t st_big[4] = {nullptr, nullptr, nullptr, nullptr};   // per part: CTA-per-chain E-step kernel beside the warp-per-chain one (EstepPlan::st_big)
  cudaEvent_t ev_big[4][2] = {};
  StagedBufs staged;
  std::unique_ptr<WorkPool> pool;  // host workers of this lane's wave loop (sleep between regions)
  PinnedBuf<char> h_stage[2];      // pinned upload staging of the staged wave, alternating per wave
  cudaEvent_t ev_mid = nullptr;    // recorded on `st` where the next wave's likelihood phase may start
  int stage_step = 40;
  DevBuf<double> d_fx, d_fl, d_fr, d_fpa, d_cnt, d_theta, d_table, d_tensor, d_lz, d_trace_ws;   // d_lz = [log_zmat scratch | V]
  DevBuf<UtrDev> d_utrs;
  DevBuf<RowRef> d_rows, d_trows;
  DevBuf<TileRef> d_tiles;
  DevBuf<ChainDev> d_chains;
  DevBuf<int32_t> d_labels, d_trace_a, d_trace_b, d_r2b;
  DevBuf<int64_t> d_labels64;           // per-read labels of the wave (label_expand_kernel)
  PinnedBuf<int64_t> h_labels64;
  DevBuf<ScanRef> d_refs;
  DevBuf<ClusterJob> d_cjobs;
  DevBuf<long long> d_clstats;
  cudaEvent_t ev_cl[2] = {nullptr, nullptr};   // around the cluster-resident EM launch
  DevBuf<ScanDesc> d_descs;
  DevBuf<int32_t> d_chain_off, d_chain_idx, d_lists, d_counts;
  DevBuf<double> d_partials, d_counter;
  DevBuf<LabelDev> d_jobs;
  PinnedBuf<ChainDev> h_chains, h_refits;
  PinnedBuf<LabelDev> h_jobs;
  cudaEvent_t ev[8];
  EmStepEvents em_events, em_events2;   // bulk-synchronous runs: all steps / the head of the tail route
  EmStepEvents em_events_part[3];       // parts 2..4 of a split wave
  PinnedBuf<char> h_runmeta;            // per-run index lists, scan items, cluster jobs (pinned upload staging)
  // a run whose results are collected later (the prune refits + labels of a wave finish on the GPU while
  // the host prepares the next wave): its own pinned staging and events, so the next run can be
  // enqueued behind it on the same stream without a host synchronisation in between
  PinnedBuf<char> h_runmeta2;
  PinnedBuf<double> h_scalars;          // [0] scan element counter of a normal run, [1] of a deferred one
  cudaEvent_t ev_def[3] = {nullptr, nullptr, nullptr};   // deferred run: begin, end of the EM, everything downloaded
  struct Deferred {
    bool active = false;
    ChainDev* chains = nullptr;
    size_t n = 0;
  } deferred;
  scape_b200_timing tm;
  std::vector<std::pair<float, float>> busy;   // kernel intervals (ms since the fit's base event)
  std::string err;
  int rc = 0;
  void release() {
    d_fx.release(); d_fl.release(); d_fr.release(); d_fpa.release(); d_cnt.release(); d_theta.release();
    d_table.release(); d_tensor.release(); d_lz.release(); d_trace_ws.release(); d_utrs.release();
    d_rows.release(); d_trows.release(); d_tiles.release(); d_chains.release(); d_labels.release(); d_r2b.release(); d_labels64.release(); h_labels64.release(); d_trace_a.release(); d_trace_b.release();
    d_refs.release(); d_cjobs.release(); d_clstats.release(); d_descs.release(); d_chain_off.release(); d_chain_idx.release(); d_partials.release();
    d_lists.release(); d_counts.release();
    d_counter.release(); d_jobs.release();
    h_chains.release(); h_refits.release(); h_runmeta.release(); h_runmeta2.release(); h_scalars.release(); h_jobs.release();
    h_stage[0].release(); h_stage[1].release();
    staged.release();
  }
  // the staged wave becomes the current one (pointer swaps only)
  void adopt_staged() {
    std::swap(d_fx, staged.d_fx); std::swap(d_fl, staged.d_fl); std::swap(d_fr, staged.d_fr);
    std::swap(d_fpa, staged.d_fpa); std::swap(d_cnt, staged.d_cnt); std::swap(d_theta, staged.d_theta);
    std::swap(d_table, staged.d_table); std::swap(d_tensor, staged.d_tensor); std::swap(d_utrs, staged.d_utrs);
    std::swap(d_rows, staged.d_rows); std::swap(d_trows, staged.d_trows); std::swap(d_tiles, staged.d_tiles);
    std::swap(d_r2b, staged.d_r2b);
    for (int i = 0; i < 3; i++) std::swap(ev[i], staged.ev[i]);
  }
};

constexpr int kMaxLanes = 4;
constexpr int kMaxSplit = 4;   // parts a bulk-synchronous wave can be split into (one stream each)

struct scape_b200_handle {
  int device = 0;
  int n_sm = 148;
  scape_b200_params P;
  ModelConst mc;
  Lane lanes[kMaxLanes];
  // Lanes: independent wave loops (host thread, streams, buffers) over disjoint sets of RNG streams.
  // Between two waves a lane's host does selection, the next wave's draws and launch lists (~1.2 ms for a
  // cfg-2 wave) while its streams are empty; with two lanes the other lane's EM fills the GPU meanwhile
  // (measured: e2e +8 % at 100 streams, +15 % at 12).  Un-pipelined passes (overlap off) use one lane.
  // Fewer streams per GPU want more lanes (12 streams: 4,567 / 4,694 / 4,820 UTR/s e2e with 2 / 3 / 4 lanes;
  // 100 streams: 11,619 with 2, 11,363 with 3; a 25-stream slice of the heavy-tailed cfg-3: 4,197 with 2, 3,982
  // with 4, 3,718 with 1): n_lanes = 0 picks 4 lanes up to 16 streams, 3 up to 40, else 2.
  int n_lanes = 0;            // SCAPE_B200_LANES; 0 = by the batch's stream count
  int lanes_in_use = 1;       // of the running fit
  bool tensor_fast = false;   // default grid shape: alpha rows use the constant-weight kernel
  bool tensor_fast_edges = true;   // ... including the rows whose windows are clipped by the grid ends (SCAPE_B200_TENSOR_EDGES=0: generic kernel)
  double tf_g[kTfB * kTfW], tf_lp[kTfB * kTfW], tf_lps[kTfB];
  int tf_hw[kTfB];
  cudaEvent_t base_ev = nullptr;
  scape_b200_timing tm;
  double wave_budget_bytes = 24e9;
  bool tensor_f32 = true;   // tensor storage: FP32 (default) or FP64; all arithmetic is FP64 either way
  int host_threads = 0;     // 0 = CPUs of this process / ranks sharing the host
  std::unique_ptr<WorkPool> prep_pool;
  bool overlap = true;      // likelihood phase of wave w+1 runs under the EM of wave w
  size_t l2_persist_bytes = 0, l2_window_max = 0;   // persisting-L2 set-aside for the EM state (SCAPE_B200_L2_PERSIST_MB, 0 = off)
  bool poison = false;      // SCAPE_B200_POISON=1 (tests): fill the tensor arena with NaN bits before every wave
};

static double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

static int pad4(int64_t n) { return int((n + 3) / 4 * 4); }

static void fill_model_const(const scape_b200_params& P, ModelConst& mc) {
  memset(&mc, 0, sizeof(mc));
  mc.mu_f = P.mu_f;
  mc.sigma_f = P.sigma_f;
  mc.max_unif_ws = P.max_unif_ws;
  mc.n_s = P.n_s;
  mc.n_beta = P.n_beta;
  for (int i = 0; i < P.n_s; i++) {
    mc.s_dis[i] = P.s_dis[i];
    mc.pmf_s[i] = P.pmf_s[i];
    mc.logpmf_s[i] = std::log(P.pmf_s[i]);
  }
  for (int i = 0; i < P.n_beta; i++) mc.betas[i] = P.betas[i];
  // exp_nonpos (em_device.cuh): log2(e), 1.5 * 2^52, -ln2_hi, -ln2_lo, 1/13! ... 1/2!, 1, clamp
  static const double expc[18] = {
      1.4426950408889634074, 6755399441055744.0, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
      1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06,
      2.48015873015873e-05, 0.0001984126984126984, 0.001388888888888889, 0.008333333333333333, 0.041666666666666664,
      0.16666666666666666, 0.5, 1.0, -746.0};
  memcpy(mc.expc, expc, sizeof(expc));
}

static int check_params(const scape_b200_params& P) {
  if (P.n_beta <= 0 || P.n_beta > SCAPE_B200_MAX_BETA) return fail(-5, "n_beta out of range");
  if (P.n_s <= 0 || P.n_s > SCAPE_B200_MAX_S) return fail(-5, "n_s out of range");
  if (P.n_smooth <= 0 || P.n_smooth > SCAPE_B200_MAX_SMOOTH || P.n_smooth % 2 == 0)
    return fail(-5, "n_smooth out of range");
  if (P.theta_step <= 0 || P.beta_step <= 0) return fail(-5, "theta_step / beta_step must be positive");
  if (P.n_max_apa > SCAPE_B200_KCAP || P.n_min_apa < 1)
    return fail(-5, "n_max_apa above SCAPE_B200_KCAP or n_min_apa < 1");
  if (P.fixed_run_mode && (P.pre_K < 1 || P.pre_K > SCAPE_B200_KCAP)) return fail(-5, "pre_K out of range");
  return 0;
}

// widest marginal window (in grid points) any (alpha, beta) of this parameter set can have
static int max_window(const scape_b200_params& P) {
  double bmax = 0;
  for (int i = 0; i < P.n_beta; i++) bmax = std::max(bmax, P.betas[i]);
  return 2 * int(std::floor(3 * bmax / P.theta_step)) + 1;
}

// ------------------------------------------------------------------------------------------------
extern "C" {

const char* scape_b200_last_error(void) { return g_err.c_str(); }
int scape_b200_version(void) { return 100; }

int scape_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int scape_b200_create(int device, const scape_b200_params* params, scape_b200_handle** out) {
  if (!params || !out) return fail(-5, "null argument");
  if (int rc = check_params(*params)) return rc;
  int n = scape_b200_device_count();
  if (n <= 0) return fail(-101, "no CUDA device: libscape_b200 has no CPU fallback");
  if (device < 0 || device >= n) return fail(-101, "device index out of range");
  CU(cudaSetDevice(device));
  scape_b200_handle* h = new scape_b200_handle();
  h->device = device;
  {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    h->n_sm = prop.multiProcessorCount;
  }
  h->P = *params;
  fill_model_const(h->P, h->mc);
  {
    // Fast marginal kernel applies to the default grid shape (13 betas, widest window 43 points):
    // weights of the (alpha-independent) interior windows, taichi_core.py:160-169.
    const scape_b200_params& P = h->P;
    bool ok = !P.fixed_run_mode && P.n_beta == kTfB && max_window(P) == kTfW;
    if (const char* s = getenv("SCAPE_B200_TENSOR_FAST")) ok = ok && atoi(s) != 0;
    if (ok) {
      for (int j = 0; j < kTfB; j++) {
        const double beta = P.betas[j];
        const int hw = int(std::floor(3 * beta / P.theta_step));
        h->tf_hw[j] = hw;
        double sum = 0.0;
        for (int d = 0; d < kTfW; d++) {
          const double x = double(P.theta_step) * (d - kTfHalf) / beta;
          const double lp = -0.5 * (x * x) - std::log(beta) - 0.5 * std::log(2 * 3.141592653589793);
          h->tf_lp[j * kTfW + d] = lp;
          if (std::abs(d - kTfHalf) <= hw) sum += std::exp(lp);      // theta order, like the reference
        }
        h->tf_lps[j] = std::log(sum);
        for (int d = 0; d < kTfW; d++)
          h->tf_g[j * kTfW + d] = std::abs(d - kTfHalf) <= hw ? std::exp(h->tf_lp[j * kTfW + d] - h->tf_lps[j]) : 0.0;
      }
      h->tensor_fast = true;
      for (int j = 0; j < kTfB; j++) h->tensor_fast = h->tensor_fast && h->tf_hw[j] == tf_default_hw(j);   // the kernel's compile-time windows
      if (const char* s2 = getenv("SCAPE_B200_TENSOR_EDGES")) h->tensor_fast_edges = atoi(s2) != 0;
    }
  }
  int prio_lo = 0, prio_hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  for (Lane& L : h->lanes) {
    CU(cudaStreamCreateWithPriority(&L.st, cudaStreamNonBlocking, prio_hi));
    CU(cudaStreamCreateWithPriority(&L.st_lik, cudaStreamNonBlocking, prio_lo));
    for (auto& sa : L.st_aux) CU(cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, prio_hi));
    CU(cudaEventCreateWithFlags(&L.ev_fork, cudaEventDisableTiming));
    for (auto& e : L.ev_join) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& sb : L.st_big) CU(cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, prio_hi));
    for (auto& eb : L.ev_big)
      for (auto& e : eb) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : L.ev) CU(cudaEventCreate(&e));
    for (auto& e : L.ev_def) CU(cudaEventCreate(&e));
    for (auto& e : L.ev_cl) CU(cudaEventCreate(&e));
    for (auto& e : L.staged.ev) CU(cudaEventCreate(&e));
    CU(cudaEventCreateWithFlags(&L.ev_mid, cudaEventDisableTiming));
    if (const char* s = getenv("SCAPE_B200_STAGE_STEP")) L.stage_step = atoi(s);
    memset(&L.tm, 0, sizeof(L.tm));
  }
  {
    // persisting L2 set-aside for the EM working state (see run_chains)
    int max_persist = 0, max_window = 0;
    CU(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device));
    CU(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device));
    double want_mb = 0;      // measured: no effect on the E step; opt-in knob
    if (const char* s = getenv("SCAPE_B200_L2_PERSIST_MB")) want_mb = atof(s);
    const size_t want = size_t(std::max(0.0, want_mb) * 1024 * 1024);
    h->l2_persist_bytes = std::min(want, size_t(std::max(0, max_persist)));
    h->l2_window_max = size_t(std::max(0, max_window));
    if (h->l2_persist_bytes > 0 && h->l2_window_max > 0)
      CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, h->l2_persist_bytes));
    else
      h->l2_persist_bytes = 0;
  }
  CU(cudaEventCreate(&h->base_ev));
  CU(cudaEventRecord(h->base_ev, 0));
  CU(cudaEventSynchronize(h->base_ev));
  memset(&h->tm, 0, sizeof(h->tm));
  if (const char* s = getenv("SCAPE_B200_WAVE_GB")) h->wave_budget_bytes = atof(s) * 1e9;
  if (const char* s = getenv("SCAPE_B200_THREADS")) h->host_threads = atoi(s);
  if (const char* s = getenv("SCAPE_B200_TENSOR")) h->tensor_f32 = (strcmp(s, "f64") != 0);
  if (const char* s = getenv("SCAPE_B200_OVERLAP")) h->overlap = atoi(s) != 0;
  if (const char* s = getenv("SCAPE_B200_POISON")) h->poison = atoi(s) != 0;
  if (const char* s = getenv("SCAPE_B200_LANES")) h->n_lanes = std::max(0, std::min(kMaxLanes, atoi(s)));
  *out = h;
  return 0;
}

int scape_b200_destroy(scape_b200_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  for (Lane& L : h->lanes) {
    cudaStreamSynchronize(L.st);
    cudaStreamSynchronize(L.st_lik);
    for (auto& sa : L.st_aux) cudaStreamSynchronize(sa);
    for (auto& sb : L.st_big) cudaStreamSynchronize(sb);
    L.release();
    for (auto& e : L.ev) cudaEventDestroy(e);
    for (auto& e : L.ev_def) cudaEventDestroy(e);
    for (auto& e : L.ev_cl) cudaEventDestroy(e);
    for (auto& e : L.staged.ev) cudaEventDestroy(e);
    cudaEventDestroy(L.ev_mid);
    cudaStreamDestroy(L.st);
    cudaStreamDestroy(L.st_lik);
    for (auto& sa : L.st_aux) cudaStreamDestroy(sa);
    cudaEventDestroy(L.ev_fork);
    for (auto& e : L.ev_join) cudaEventDestroy(e);
    for (auto& sb : L.st_big) cudaStreamDestroy(sb);
    for (auto& eb : L.ev_big)
      for (auto& e : eb) cudaEventDestroy(e);
  }
  cudaEventDestroy(h->base_ev);
  delete h;
  return 0;
}

int scape_b200_set_tensor_dtype(scape_b200_handle* h, int bytes) {
  if (!h) return fail(-5, "null handle");
  if (bytes != 4 && bytes != 8) return fail(-5, "tensor dtype must be 4 (float) or 8 (double) bytes");
  h->tensor_f32 = (bytes == 4);
  return 0;
}

int scape_b200_set_host_threads(scape_b200_handle* h, int n) {
  if (!h) return fail(-5, "null handle");
  if (n < 0) return fail(-5, "host thread count must be >= 0 (0 = the process default)");
  h->host_threads = n;      // the pools are rebuilt by the next fit if the count changed
  return 0;
}

int scape_b200_set_overlap(scape_b200_handle* h, int on) {
  if (!h) return fail(-5, "null handle");
  h->overlap = on != 0;
  return 0;
}

int scape_b200_fp64_peaks(scape_b200_handle* h, double* dfma_tflops, double* dmma_tflops) {
  if (!h || !dfma_tflops || !dmma_tflops) return fail(-5, "null argument");
  CU(cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, h->device));
  if (measure_fp64_peaks(prop.multiProcessorCount, dfma_tflops, dmma_tflops, h->lanes[0].st)) return fail(-100, "peak kernels failed");
  return 0;
}

int scape_b200_sfu_peaks(scape_b200_handle* h, double* out4) {
  if (!h || !out4) return fail(-5, "null argument");
  CU(cudaSetDevice(h->device));
  if (measure_sfu_peaks(h->n_sm, out4, h->lanes[0].st)) return fail(-100, "peak kernels failed");
  return 0;
}

int scape_b200_get_timing(scape_b200_handle* h, scape_b200_timing* out) {
  if (!h || !out) return fail(-5, "null argument");
  *out = h->tm;
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// wave machinery
// ------------------------------------------------------------------------------------------------
namespace {

struct WaveUtr {
  int64_t u;          // index into the batch
  int k_max, k_min;   // current sweep range
  int sweeps = 0, chains_run = 0, k_selected = 0;
  bool done = false;
  ChainDev best;      // current result chain (a_idx/b_idx/ws/bic/lb_arr/n_iter)
  double work = 0, iters = 0;
};

// np.argmin over the reference's bic_arr.  That array is created with np.full(n, np.finfo('f').max)
// (apa_core.py:427, 849, 945) and is therefore FLOAT32: BICs are compared after rounding to float32
// and the first of the tied minima wins.
int np_argmin(const std::vector<double>& v) {
  int best = 0;
  for (int i = 0; i < (int)v.size(); i++) {
    if (std::isnan(v[size_t(i)])) return i;
    if (float(v[size_t(i)]) < float(v[size_t(best)])) best = i;
  }
  return best;
}

// How the EM chains of a run are executed (SCAPE_B200_EM; measurements in DESIGN.md section 5):
//   bsp      all 51 steps as bulk-synchronous {E step, scan} launch pairs over all chains of the wave
//            (batched FP64 MMA scan).  Fastest whenever a wave holds enough chains to fill the GPU
//            (cfg-2: 100 UTRs = 5,000 chains per wave), and what giant UTRs always take.
//   tail     the first SCAPE_B200_TAIL_STEP (8) iterations bulk-synchronous, then ONE launch of the
//            chain-resident kernel (em_tail.cu): every chain still running gets a CTA and iterates to
//            convergence on its own.  2 launches + 16 instead of 101 per wave: wins on small waves
//            (few RNG streams per GPU), where the step kernels are latency floors.
//   cluster  one thread-block cluster per UTR, all iterations in one launch (em_cluster.cu).  Measured
//            slower than both everywhere (idle warps at the cluster barriers); kept as an experiment.
//   auto (default)  tail for runs of fewer than SCAPE_B200_TAIL_CHAINS chains, else bsp.  The default
//            threshold is 0 (always bsp): the tail route wins a few per cent on small waves of SMALL
//            UTRs (cfg-2 shape, 12-25 UTRs per wave) but loses badly on small heavy-tailed waves (cfg-3
//            over 8 GPUs: 24.2k -> 16.0k UTR/s), where a chain's own window no longer sits in L2.
enum EmRoute : char { kRouteBsp = 0, kRouteCluster = 1, kRouteTail = 2, kRouteBspPart = 3 };   // kRouteBspPart + p: part p + 1 of a split wave

// The chains of a set of UTRs prepared for the bulk-synchronous step kernels: the E-step launch
// order (few-fragment chains first, sorted by K and N) and the scan's work items.
struct StepSet {
  std::vector<int32_t> index;
  int64_t n_small = 0, n_big = 0;
  std::vector<ScanRef> refs;        // chunked-path items first (n_refs_chunk of them), then tile-path items
  int64_t n_refs_chunk = 0;
  bool big_k = false;
};

template <class Chains>
int build_step_set(scape_b200_handle* h, const Chains& chains, const std::vector<UtrDev>& utrs_host,
                   const std::vector<int32_t>& chain_off, const std::vector<char>& scans, const std::vector<char>& route,
                   char want, StepSet& out) {
  const size_t W = utrs_host.size();
  std::vector<int32_t>& index = out.index;
  index.reserve(chains.size());
  static const int warp_max_n = getenv("SCAPE_B200_WARP_MAXN") ? atoi(getenv("SCAPE_B200_WARP_MAXN")) : kWarpEstepMaxN;
  auto mine = [&](size_t i) { return route[size_t(chains[i].utr)] == want; };
  for (size_t i = 0; i < chains.size(); i++)
    if (mine(i) && utrs_host[size_t(chains[i].utr)].N <= warp_max_n) index.push_back(int32_t(i));
  out.n_small = int64_t(index.size());
  // same-K chains next to each other: the warps resident on an SM then run the same template
  // instantiation of the E step (instruction-cache locality; 24 % 'no instruction' stalls otherwise);
  // within one K the chains with the longest fragment loop go first (a warp's time is ~ N / 32
  // fragment passes, the launch ends with the slowest warp)
  // (stable counting sort on (K descending, N descending): both keys are small integers, and this
  // runs on the host's critical path between two waves)
  if (warp_max_n <= 4096) {
    const size_t stride = size_t(warp_max_n) + 1;
    std::vector<int32_t> bucket_of(index.size());
    std::vector<int32_t> start((SCAPE_B200_KCAP + 1) * stride + 1, 0);
    for (size_t j = 0; j < index.size(); j++) {
      const ChainDev& c = chains[size_t(index[j])];
      const int N = utrs_host[size_t(c.utr)].N;
      if (c.K < 0 || c.K > SCAPE_B200_KCAP || N < 0 || N > warp_max_n) return fail(-5, "internal: chain outside the E-step sort range");
      const size_t b = size_t(SCAPE_B200_KCAP - c.K) * stride + size_t(warp_max_n - N);
      bucket_of[j] = int32_t(b);
      start[b + 1]++;
    }
    for (size_t b = 1; b < start.size(); b++) start[b] += start[b - 1];
    std::vector<int32_t> sorted(index.size());
    for (size_t j = 0; j < index.size(); j++) sorted[size_t(start[size_t(bucket_of[j])]++)] = index[j];
    index.swap(sorted);
  } else {
    std::stable_sort(index.begin(), index.end(), [&](int32_t a, int32_t b) {
      const ChainDev &ca = chains[size_t(a)], &cb = chains[size_t(b)];
      if (ca.K != cb.K) return ca.K > cb.K;
      return utrs_host[size_t(ca.utr)].N > utrs_host[size_t(cb.utr)].N;
    });
  }
  for (size_t i = 0; i < chains.size(); i++)
    if (mine(i) && utrs_host[size_t(chains[i].utr)].N > warp_max_n) index.push_back(int32_t(i));
  out.n_big = int64_t(index.size()) - out.n_small;
  for (int32_t ci : index) out.big_k = out.big_k || chains[size_t(ci)].K > 7;
  // Scan work items.  A CTA's cost is (fragments of the UTR) x (chains it multiplies); the step ends
  // with the slowest CTA.  UTRs whose single-CTA cost is above half of an even share of the wave's
  // work over the resident CTA slots get their chain sub-batches narrowed (32 -> 16 -> 8 chains) and
  // dealt to separate CTAs (the tensor block is then re-read from L2, which a big UTR can afford).
  // Items are issued most expensive first.
  std::vector<ScanRef>& refs = out.refs;
  static const int split_mode = getenv("SCAPE_B200_SCAN_SPLIT") ? atoi(getenv("SCAPE_B200_SCAN_SPLIT")) : 1;
  const double slots = 2.0 * h->n_sm;
  double total = 0;
  for (size_t i = 0; i < W; i++)
    if (scans[i] && route[i] == want) {
      const UtrDev& u = utrs_host[i];
      const double n_blk = double((int64_t(u.T) * u.B + kScanRows - 1) / kScanRows);
      total += n_blk * u.N * double(chain_off[i + 1] - chain_off[i]);
    }
  const double limit = std::max(0.5 * total / slots, 2048.0);
  std::vector<double> cost;
  for (size_t i = 0; i < W; i++)
    if (scans[i] && route[i] == want) {
      const UtrDev& u = utrs_host[i];
      const int32_t n_blk = int32_t((int64_t(u.T) * u.B + kScanRows - 1) / kScanRows);
      const int C = chain_off[i + 1] - chain_off[i];
      int gb = 32, nsb = 1;
      // tile path when whole V rows of >= 8 chains fit the scan CTA's shared memory (N <= ~1280)
      static const bool tiles_on = getenv("SCAPE_B200_SCAN_TILES") ? atoi(getenv("SCAPE_B200_SCAN_TILES")) != 0 : false;   // opt-in: measured 16 % slower
      const int tile_gb = tiles_on ? scan_tile_chains(u.N) : 0;
      if (tile_gb) gb = tile_gb;
      if (split_mode && double(u.N) * C > limit) {
        while (gb > 8 && double(u.N) * std::min(gb, C) > limit) gb /= 2;
        nsb = (C + gb - 1) / gb;
      }
      for (int32_t b = 0; b < n_blk; b++)
        for (int sb = 0; sb < nsb; sb++) {
          refs.push_back(ScanRef{int32_t(i), b, int16_t(sb), int16_t(nsb), int16_t(gb), int16_t(tile_gb ? 1 : 0)});
          cost.push_back(double(u.N) * std::min(nsb == 1 ? C : gb, C));
        }
    }
  // work items of the chunked path first, then those of the tile path (two kernels), each most expensive first
  {
    std::vector<size_t> ord(refs.size());
    for (size_t i = 0; i < ord.size(); i++) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
      if (refs[a].pad != refs[b].pad) return refs[a].pad < refs[b].pad;
      return split_mode ? cost[a] > cost[b] : false;
    });
    std::vector<ScanRef> sorted(refs.size());
    for (size_t i = 0; i < ord.size(); i++) sorted[i] = refs[ord[i]];
    refs.swap(sorted);
    out.n_refs_chunk = 0;
    for (const ScanRef& r : refs) out.n_refs_chunk += r.pad == 0;
  }
  return 0;
}

// Upload chains, run them to convergence, bring them back.  `utrs_host` is the wave's UtrDev array;
// chains must be ordered by UTR (they are generated that way).
// `enqueued` (optional) runs on the host after every launch and the download have been enqueued and
// before the stream is synchronised: host work that hides under the run, or more work for the stream.
// `defer`: return once everything is enqueued (no synchronisation, no accounting); the caller collects the
// run with finish_deferred_run() before it touches the chains or starts another deferred run.
int run_chains(scape_b200_handle* h, Lane& L, ChainDev* chains_p, size_t n_chains, const std::vector<UtrDev>& utrs_host,
               bool want_trace = false, const std::function<int()>& enqueued = nullptr, bool defer = false) {
  static const bool host_dbg_rc = scape_env_on("SCAPE_B200_DBG_HOST");
  const double t_rc0 = now_ms();
  if (defer) {
    if (L.deferred.active) return fail(-5, "internal: a deferred run is still open");
    L.deferred.active = true;
    L.deferred.chains = chains_p;
    L.deferred.n = n_chains;
  }
  CU(L.h_scalars.resize(2));
  if (n_chains == 0) {                 // nothing to run: the caller's follow-up work still gets its synchronisation
    if (enqueued) {
      if (int rc = enqueued()) return rc;
      if (!defer) CU(cudaStreamSynchronize(L.st));
    }
    if (defer) CU(cudaEventRecord(L.ev_def[2], L.st));
    return 0;
  }
  struct Span {
    ChainDev* p; size_t n;
    size_t size() const { return n; }
    ChainDev* data() const { return p; }
    ChainDev& operator[](size_t i) const { return p[i]; }
    ChainDev* begin() const { return p; }
    ChainDev* end() const { return p + n; }
  } chains{chains_p, n_chains};
  const size_t W = utrs_host.size();
  int64_t lz = 0, vsz = 0, tr = 0, pb = 0;
  std::vector<int32_t> chain_off(W + 1, 0);
  std::vector<char> scans(W, 0);
  bool any_scan = false;
  static const char* em_env = getenv("SCAPE_B200_EM");
  static const size_t tail_chains = getenv("SCAPE_B200_TAIL_CHAINS") ? size_t(atol(getenv("SCAPE_B200_TAIL_CHAINS"))) : 0;
  const char mode = (!em_env || !strcmp(em_env, "auto")) ? (chains.size() < tail_chains ? kRouteTail : kRouteBsp)
                    : !strcmp(em_env, "bsp") ? kRouteBsp : !strcmp(em_env, "cluster") ? kRouteCluster : kRouteTail;
  static const int tail_step = std::max(0, std::min(SCAPE_B200_NROUND, getenv("SCAPE_B200_TAIL_STEP") ? atoi(getenv("SCAPE_B200_TAIL_STEP")) : 8));
  // per-iteration cost (fragment x candidate row x chain products) above which a UTR stays on the
  // bulk-synchronous kernels for the whole run: the resident kernels give a UTR / a chain a few SMs,
  // a giant UTR's grid search needs all of them
  static const double resident_max_cost = getenv("SCAPE_B200_CLUSTER_COST") ? atof(getenv("SCAPE_B200_CLUSTER_COST")) : 4e8;
  std::vector<char> route(W, kRouteBsp);
  {
    std::vector<int32_t> n_scan(W, 0);
    for (size_t i = 0; i < chains.size(); i++)
      if (!chains[i].weights_only) n_scan[size_t(chains[i].utr)]++;
    for (size_t i = 0; i < W && mode != kRouteBsp; i++) {
      const UtrDev& u = utrs_host[i];
      const double cost = double(u.N) * double(u.T) * u.B * n_scan[i];
      if (n_scan[i] == 0 || cost > resident_max_cost) continue;
      if (mode == kRouteCluster && cluster_chains_per_pass(u.N) >= 8) route[i] = kRouteCluster;
      if (mode == kRouteTail && u.N <= 4096) route[i] = kRouteTail;
    }
  }
  // Split a bulk-synchronous wave into two halves that step independently on two streams: one half's
  // E step (latency-bound, FP64 pipe ~20 % busy) then runs under the other half's scan (tensor-pipe
  // bound).  UTRs are dealt alternately in order of decreasing cost.  SCAPE_B200_SPLIT=0 switches it
  // off; runs with traces, few UTRs or weights-only chains are never split.
  // (default: 2 parts with a single lane; with several lanes the lanes already interleave, and splitting their
  // half-size waves again measured slower: 10,951 vs 11,619 UTR/s e2e on cfg-2)
  static const int split_knob = getenv("SCAPE_B200_SPLIT") ? atoi(getenv("SCAPE_B200_SPLIT")) : 0;
  const int split_env = std::max(1, std::min(kMaxSplit, split_knob > 0 ? split_knob : (h->lanes_in_use > 1 ? 1 : 2)));
  int n_parts = 1;
  if (split_env > 1 && h->overlap && !want_trace && mode == kRouteBsp) {
    std::vector<std::pair<double, size_t>> order;
    for (size_t i = 0; i < W; i++)
      if (route[i] == kRouteBsp) order.emplace_back(-double(utrs_host[i].N) * utrs_host[i].T, i);
    bool any_weights_only = false;
    for (size_t i = 0; i < chains.size(); i++) any_weights_only = any_weights_only || chains[i].weights_only;
    if (order.size() >= size_t(8 * split_env) && !any_weights_only) {
      std::sort(order.begin(), order.end());
      n_parts = split_env;
      for (size_t j = 0; j < order.size(); j++)
        if (j % size_t(n_parts)) route[order[j].second] = char(kRouteBspPart + int(j % size_t(n_parts)) - 1);
    }
  }
  for (size_t i = 0; i < chains.size(); i++) {
    ChainDev& c = chains[i];
    if (i > 0 && c.utr < chains[i - 1].utr) return fail(-5, "internal: chains not ordered by UTR");
    const UtrDev& u = utrs_host[size_t(c.utr)];
    const int rows_per_partial = route[size_t(c.utr)] == kRouteCluster ? kClusterTileRows : kScanRows;
    const int64_t n_blk = (int64_t(u.T) * u.B + rows_per_partial - 1) / rows_per_partial;
    c.lz_off = lz;
    lz += int64_t(c.K + 1) * u.Npad;
    c.v_off = vsz;
    vsz += (int64_t(u.N) + 7) / 8 * 8;
    c.pb_off = pb;
    pb += n_blk;
    c.trace_off = want_trace ? tr : -1;
    tr += int64_t(SCAPE_B200_NROUND) * (SCAPE_B200_KCAP + 1);
    c.n_iter = 0;
    c.grid_rows = 0;
    c.lb_prev = kSentinel;
    c.last_a = 0;
    c.state = 1;
    c.pending = 0;
    c.trace_pending = 0;
    chain_off[size_t(c.utr) + 1]++;
    if (!c.weights_only) { scans[size_t(c.utr)] = 1; any_scan = true; }
  }
  for (size_t i = 0; i < W; i++) {
    if (scans[i] && chain_off[i + 1] > kScanMaxChains)
      return fail(-5, "internal: " + std::to_string(chain_off[i + 1]) + " chains of one UTR in one run; the scan lists at most " +
                          std::to_string(kScanMaxChains));
    chain_off[i + 1] += chain_off[i];
  }
  // The chain records are complete (offsets assigned): their upload (4.8 MB for a cfg-2 wave) starts now and
  // runs while the host builds the launch lists below.
  CU(L.d_chains.ensure(chains.size()));
  CU(cudaMemcpyAsync(L.d_chains.p, chains.data(), sizeof(ChainDev) * chains.size(), cudaMemcpyHostToDevice, L.st));
  StepSet full, head, parts[kMaxSplit - 1];   // all 51 steps / the first tail_step steps before the chain-resident kernel / parts 2.. of a split wave
  if (int rc = build_step_set(h, chains, utrs_host, chain_off, scans, route, kRouteBsp, full)) return rc;
  for (int p = 1; p < n_parts; p++)
    if (int rc = build_step_set(h, chains, utrs_host, chain_off, scans, route, char(kRouteBspPart + p - 1), parts[p - 1])) return rc;
  size_t parts_index = 0, parts_refs = 0;
  for (int p = 1; p < n_parts; p++) { parts_index += parts[p - 1].index.size(); parts_refs += parts[p - 1].refs.size(); }
  if (int rc = build_step_set(h, chains, utrs_host, chain_off, scans, route, kRouteTail, head)) return rc;
  // chains of the tail route in UTR order: neighbours in the launch share a tensor (L2)
  std::vector<int32_t> tail_list;
  int tail_max_n = 0;
  for (size_t i = 0; i < chains.size(); i++)
    if (route[size_t(chains[i].utr)] == kRouteTail) {
      tail_list.push_back(int32_t(i));
      tail_max_n = std::max(tail_max_n, utrs_host[size_t(chains[i].utr)].N);
    }
  // Cluster jobs, most expensive first (the hardware dispatches clusters in launch order as SMs free up)
  std::vector<ClusterJob> cjobs;
  {
    std::vector<double> ccost;
    for (size_t i = 0; i < W; i++)
      if (route[i] == kRouteCluster) {
        const UtrDev& u = utrs_host[i];
        // all chains of the UTR (weights-only ones included: the kernel iterates them to convergence in place)
        cjobs.push_back(ClusterJob{int32_t(i), chain_off[i], chain_off[i + 1] - chain_off[i], cluster_chains_per_pass(u.N)});
        ccost.push_back(double(u.N) * u.T * double(chain_off[i + 1] - chain_off[i]));
      }
    std::vector<size_t> ord(cjobs.size());
    for (size_t i = 0; i < ord.size(); i++) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return ccost[a] > ccost[b]; });
    std::vector<ClusterJob> sorted(cjobs.size());
    for (size_t i = 0; i < ord.size(); i++) sorted[i] = cjobs[ord[i]];
    cjobs.swap(sorted);
  }
  // ---- device buffers + uploads ------------------------------------------------------------------
  const size_t n_index = full.index.size() + head.index.size() + tail_list.size() + parts_index;
  const size_t n_refs = full.refs.size() + head.refs.size() + parts_refs;
  CU(L.d_cjobs.ensure(cjobs.size() + 1));
  // log_zmat scratch and V live in ONE allocation: [lz | V].  They are what every EM iteration reads
  // and rewrites (a wave's 5,000 chains: ~55 MB + 10 MB), while each scan streams the wave's marginal
  // tensors (250-400 MB) through the L2 in between and evicts them.  An access-policy window on the EM
  // stream marks [lz | V] as persisting L2 lines (set-aside sized at create), so the E passes' stale-column
  // reads and the scan's V staging hit L2 instead of HBM; the tensor keeps streaming through the rest.
  const size_t lzv_elems = size_t(lz) + size_t(vsz) + 8;
  CU(L.d_lz.ensure(lzv_elems));
  double* const d_v = L.d_lz.p + size_t(lz);
  if (h->l2_persist_bytes > 0) {
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    const size_t win = std::min(lzv_elems * sizeof(double), h->l2_window_max);
    attr.accessPolicyWindow.base_ptr = L.d_lz.p;
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = float(std::min(1.0, double(h->l2_persist_bytes) / double(std::max<size_t>(win, 1))));
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CU(cudaStreamSetAttribute(L.st, cudaStreamAttributeAccessPolicyWindow, &attr));
  }
  CU(L.d_chain_off.ensure(W + 1));
  CU(L.d_descs.ensure(chains.size()));
  CU(cudaMemsetAsync(L.d_descs.p, 0, sizeof(ScanDesc) * chains.size(), L.st));
  CU(L.d_chain_idx.ensure(n_index + 1));
  CU(L.d_refs.ensure(n_refs + 1));
  CU(L.d_partials.ensure(size_t(pb) * 2 + 2));
  CU(L.d_counter.ensure(1));
  if (want_trace) {
    CU(L.d_trace_a.ensure(size_t(tr)));
    CU(L.d_trace_b.ensure(size_t(tr)));
    CU(L.d_trace_ws.ensure(size_t(tr)));
  }
  // the small host arrays go through one pinned blob (pageable sources would make the copies synchronous)
  int32_t *d_idx_full = L.d_chain_idx.p, *d_idx_head = d_idx_full + full.index.size(),
          *d_tail = d_idx_head + head.index.size(), *d_idx_parts = d_tail + tail_list.size();
  ScanRef *d_refs_full = L.d_refs.p, *d_refs_head = d_refs_full + full.refs.size(), *d_refs_parts = d_refs_head + head.refs.size();
  {
    auto up16 = [](size_t v) { return (v + 15) / 16 * 16; };
    const size_t o_idx = 0, o_refs = o_idx + up16(4 * n_index), o_off = o_refs + up16(sizeof(ScanRef) * n_refs),
                 o_jobs = o_off + up16(4 * (W + 1)), o_end = o_jobs + up16(sizeof(ClusterJob) * cjobs.size());
    PinnedBuf<char>& meta = defer ? L.h_runmeta2 : L.h_runmeta;
    CU(meta.resize(o_end + 16));
    char* b = meta.p;
    int32_t* hi = (int32_t*)(b + o_idx);
    std::copy(full.index.begin(), full.index.end(), hi);
    std::copy(head.index.begin(), head.index.end(), hi + full.index.size());
    std::copy(tail_list.begin(), tail_list.end(), hi + full.index.size() + head.index.size());
    {
      int32_t* o = hi + full.index.size() + head.index.size() + tail_list.size();
      for (int p = 1; p < n_parts; p++) o = std::copy(parts[p - 1].index.begin(), parts[p - 1].index.end(), o);
    }
    ScanRef* hr = (ScanRef*)(b + o_refs);
    std::copy(full.refs.begin(), full.refs.end(), hr);
    std::copy(head.refs.begin(), head.refs.end(), hr + full.refs.size());
    {
      ScanRef* o = hr + full.refs.size() + head.refs.size();
      for (int p = 1; p < n_parts; p++) o = std::copy(parts[p - 1].refs.begin(), parts[p - 1].refs.end(), o);
    }
    std::copy(chain_off.begin(), chain_off.end(), (int32_t*)(b + o_off));
    std::copy(cjobs.begin(), cjobs.end(), (ClusterJob*)(b + o_jobs));
    CU(cudaMemcpyAsync(L.d_chain_off.p, b + o_off, sizeof(int32_t) * (W + 1), cudaMemcpyHostToDevice, L.st));
    if (n_index) CU(cudaMemcpyAsync(L.d_chain_idx.p, hi, sizeof(int32_t) * n_index, cudaMemcpyHostToDevice, L.st));
    if (n_refs) CU(cudaMemcpyAsync(L.d_refs.p, hr, sizeof(ScanRef) * n_refs, cudaMemcpyHostToDevice, L.st));
    if (!cjobs.empty())
      CU(cudaMemcpyAsync(L.d_cjobs.p, b + o_jobs, sizeof(ClusterJob) * cjobs.size(), cudaMemcpyHostToDevice, L.st));
  }
  CU(cudaMemsetAsync(L.d_counter.p, 0, sizeof(double), L.st));
  L.tm.h2d_bytes += double(sizeof(ChainDev) * chains.size() + sizeof(ScanRef) * n_refs + 4 * (W + 1) + 4 * n_index);
  CU(cudaEventRecord(defer ? L.ev_def[0] : L.ev[4], L.st));
  const double t_rc1 = now_ms();
  {
    // per-step events only where somebody reads them (un-pipelined passes, SCAPE_B200_STEP_EVENTS=1)
    static const int ev_env = getenv("SCAPE_B200_STEP_EVENTS") ? atoi(getenv("SCAPE_B200_STEP_EVENTS")) : -1;
    const bool timing = ev_env >= 0 ? ev_env != 0 : !h->overlap;
    L.em_events.timing = L.em_events2.timing = timing;
    for (auto& ee : L.em_events_part) ee.timing = timing;
  }
  EstepPlan plan;
  {
    static const bool group_steps = getenv("SCAPE_B200_ESTEP") && atoi(getenv("SCAPE_B200_ESTEP")) != 0;
    static const bool warp_pf = getenv("SCAPE_B200_WARP_PF") ? atoi(getenv("SCAPE_B200_WARP_PF")) != 0 : true;
    static const int g_env = getenv("SCAPE_B200_ESTEP_G") ? atoi(getenv("SCAPE_B200_ESTEP_G")) : 0;
    // list / count scratch of the group kernel: one region per concurrently stepping set (split parts, head)
    CU(L.d_lists.ensure((kMaxSplit + 1) * (2 * chains.size() + 2)));
    CU(L.d_counts.ensure((kMaxSplit + 1) * 2 * (SCAPE_B200_NROUND + 2)));
    plan.lists = L.d_lists.p;
    plan.counts = L.d_counts.p;
    plan.n_sm = h->n_sm;
    plan.group_steps = group_steps;
    plan.warp_prefetch = warp_pf;
    static const int stage_chain = getenv("SCAPE_B200_STAGE_CHAIN") ? atoi(getenv("SCAPE_B200_STAGE_CHAIN")) : 1;
    plan.stage_chain = stage_chain;
    static const int wpc_env = getenv("SCAPE_B200_WARP_WPC") ? atoi(getenv("SCAPE_B200_WARP_WPC")) : 2;   // measured: 2 warps per chain, cfg-2 E step -12 %, value +3 %; 4: no gain
    plan.warp_wpc = (wpc_env == 1 || wpc_env == 2 || wpc_env == 4) ? wpc_env : 2;
    // 4 warps per chain for the small class whatever else is in the wave: a chain's sums must not
    // depend on the composition of its wave
    plan.g_small = (g_env == 1 || g_env == 2 || g_env == 4 || g_env == 8) ? g_env : 4;
  }
  // SCAPE_B200_ESTEP_FORK (default 1): the two E-step kernels of a wide step on two streams (EstepPlan::st_big)
  static const bool estep_fork = getenv("SCAPE_B200_ESTEP_FORK") ? atoi(getenv("SCAPE_B200_ESTEP_FORK")) != 0 : true;
  auto fork_plan = [&](EstepPlan& pl, int part) {
    if (!estep_fork) return;
    pl.st_big = L.st_big[part];
    pl.ev_big[0] = L.ev_big[part][0];
    pl.ev_big[1] = L.ev_big[part][1];
  };
  // The wave scheduler's hook (staging the next wave's likelihood phase on the low-priority stream)
  // fires at a step of the bulk-synchronous loop of `full`; a run without one releases it before its
  // resident kernel, so the staged wave's table / tensor kernels fill the SMs as this wave's CTAs retire.
  const bool have_full = !full.index.empty();
  bool hook_pending = bool(L.em_events.hook) && !have_full;
  auto fire_mark = [&]() { if (hook_pending && L.em_events.mark) L.em_events.mark(); };
  auto fire_hook = [&]() { if (hook_pending) { hook_pending = false; L.em_events.hook(); } };
  int nl = 0;
  bool resident_timed = false;
  static const bool cl_dbg = scape_env_on("SCAPE_B200_DBG");
  if (!cjobs.empty()) {
    static const int c_env = getenv("SCAPE_B200_CLUSTER") ? atoi(getenv("SCAPE_B200_CLUSTER")) : 0;
    int csize = 8;
    if (c_env == 1 || c_env == 2 || c_env == 4 || c_env == 8) csize = c_env;
    static const int solo_max = getenv("SCAPE_B200_SOLO") ? atoi(getenv("SCAPE_B200_SOLO")) : 8;   // chains left when a cluster splits up
    fire_mark();
    if (cl_dbg) CU(L.d_clstats.ensure(cjobs.size() * 10));
    CU(cudaEventRecord(L.ev_cl[0], L.st));
    CU(launch_em_cluster(L.d_cjobs.p, int(cjobs.size()), csize, L.d_chains.p, L.d_descs.p, L.d_utrs.p, L.d_tensor.p,
                         h->tensor_f32, L.d_cnt.p, L.d_lz.p, d_v, L.d_partials.p, L.d_counter.p, L.d_trace_a.p,
                         L.d_trace_b.p, L.d_trace_ws.p, cl_dbg ? L.d_clstats.p : nullptr, solo_max, L.st));
    CU(cudaEventRecord(L.ev_cl[1], L.st));
    resident_timed = true;
    if (cl_dbg) {
      constexpr int NS = 10;
      std::vector<long long> st(cjobs.size() * NS);
      CU(cudaMemcpyAsync(st.data(), L.d_clstats.p, sizeof(long long) * st.size(), cudaMemcpyDeviceToHost, L.st));
      CU(cudaStreamSynchronize(L.st));
      float cms = 0;
      CU(cudaEventElapsedTime(&cms, L.ev_cl[0], L.ev_cl[1]));
      double sum[NS] = {0}, mx[NS] = {0};
      for (size_t j = 0; j < cjobs.size(); j++)
        for (int k = 0; k < NS; k++) { sum[k] += double(st[j * NS + k]); mx[k] = std::max(mx[k], double(st[j * NS + k])); }
      const double n = double(cjobs.size()), us = 1.0 / 1965.0;     // cycles -> us at the B200's 1965 MHz
      fprintf(stderr, "cluster EM: %zu jobs x %d CTAs, kernel %.0f us | per job mean (max): rounds %.1f (%.0f) total %.0f (%.0f) us; "
                      "rounds < 12: E %.0f scan %.0f; later: E %.0f (%.0f) scan %.0f (%.0f); wait1 %.0f (%.0f) wait2 %.0f (%.0f) | "
                      "solo from round %.1f, %.0f us\n",
              cjobs.size(), csize, cms * 1e3, sum[0] / n, mx[0], sum[1] / n * us, mx[1] * us, sum[8] / n * us, sum[9] / n * us,
              sum[2] / n * us, mx[2] * us, sum[4] / n * us, mx[4] * us, sum[3] / n * us, mx[3] * us, sum[5] / n * us, mx[5] * us,
              sum[6] / n, sum[7] / n * us);
    }
    nl += 1;
    fire_hook();
  }
  L.em_events2.kinds.clear();
  if (!tail_list.empty()) {
    // first tail_step iterations bulk-synchronous (E step + batched scan per iteration, no closing E step) ...
    if (tail_step > 0 && !head.index.empty()) {
      L.em_events2.hook = nullptr;
      EstepPlan hplan = plan;
      hplan.lists = plan.lists + size_t(kMaxSplit) * (2 * chains.size() + 2);
      hplan.counts = plan.counts + size_t(kMaxSplit) * 2 * (SCAPE_B200_NROUND + 2);
      nl += launch_em_steps(L.d_chains.p, L.d_descs.p, d_idx_head, head.n_small, head.n_big, any_scan, head.big_k, d_refs_head,
                            head.n_refs_chunk, int64_t(head.refs.size()) - head.n_refs_chunk, L.d_utrs.p, L.d_chain_off.p, L.d_tensor.p, h->tensor_f32, L.d_cnt.p,
                            L.d_lz.p, d_v, L.d_partials.p, L.d_counter.p, L.d_trace_a.p, L.d_trace_b.p, L.d_trace_ws.p,
                            L.st, L.em_events2, hplan, tail_step);
    }
    // ... then every chain that still runs iterates to convergence in a CTA of its own
    fire_mark();
    if (cl_dbg) {
      CU(L.d_clstats.ensure(16));
      CU(cudaMemsetAsync(L.d_clstats.p, 0, 8 * 16, L.st));
    }
    CU(cudaEventRecord(L.ev_cl[0], L.st));
    CU(launch_em_tail(L.d_chains.p, L.d_descs.p, d_tail, 0, int(tail_list.size()), tail_max_n, kScanRows, L.d_utrs.p,
                      L.d_tensor.p, h->tensor_f32, L.d_cnt.p, L.d_lz.p, L.d_partials.p, L.d_trace_a.p, L.d_trace_b.p,
                      L.d_trace_ws.p, cl_dbg ? (unsigned long long*)L.d_clstats.p : nullptr, L.st));
    CU(cudaEventRecord(L.ev_cl[1], L.st));
    if (cl_dbg) {
      unsigned long long st[6];
      CU(cudaMemcpyAsync(st, L.d_clstats.p, sizeof(st), cudaMemcpyDeviceToHost, L.st));
      CU(cudaStreamSynchronize(L.st));
      float cms = 0;
      CU(cudaEventElapsedTime(&cms, L.ev_cl[0], L.ev_cl[1]));
      const double it = std::max(1.0, double(st[0])), us = 1.0 / 1965.0;
      fprintf(stderr, "tail EM: %zu chains listed, %llu ran %llu iterations, kernel %.0f us | per iteration: total %.1f us = E %.1f + scan %.1f "
                      "+ apply %.1f (CTA time)\n", tail_list.size(), st[5], st[0], cms * 1e3, double(st[1]) / it * us,
              double(st[2]) / it * us, double(st[3]) / it * us, double(st[4]) / it * us);
    }
    resident_timed = true;
    nl += 1;
    fire_hook();
  }
  if (n_parts > 1) {
    // parts 2.. on the auxiliary streams: fork after the uploads, join before the download.  The
    // group-kernel list / count scratch is not used by runs with a grid search, so the parts share nothing.
    CU(cudaEventRecord(L.ev_fork, L.st));
    int32_t* di = d_idx_parts;
    ScanRef* dr = d_refs_parts;
    for (int p = 1; p < n_parts; p++) {
      StepSet& S = parts[p - 1];
      cudaStream_t sp = L.st_aux[p - 1];
      EmStepEvents& ee = L.em_events_part[p - 1];
      ee.hook = nullptr;
      ee.mark = nullptr;
      ee.kinds.clear();
      if (!S.index.empty()) {
        EstepPlan pplan = plan;
        pplan.lists = plan.lists + size_t(p) * (2 * chains.size() + 2);
        pplan.counts = plan.counts + size_t(p) * 2 * (SCAPE_B200_NROUND + 2);
        fork_plan(pplan, p);
        CU(cudaStreamWaitEvent(sp, L.ev_fork, 0));
        nl += launch_em_steps(L.d_chains.p, L.d_descs.p, di, S.n_small, S.n_big, any_scan, S.big_k, dr, S.n_refs_chunk,
                              int64_t(S.refs.size()) - S.n_refs_chunk, L.d_utrs.p, L.d_chain_off.p, L.d_tensor.p, h->tensor_f32,
                              L.d_cnt.p, L.d_lz.p, d_v, L.d_partials.p, L.d_counter.p, L.d_trace_a.p, L.d_trace_b.p,
                              L.d_trace_ws.p, sp, ee, pplan, SCAPE_B200_NROUND + 1);
        CU(cudaEventRecord(L.ev_join[p - 1], sp));
      }
      di += S.index.size();
      dr += S.refs.size();
    }
  }
  if (have_full) fork_plan(plan, 0);
  if (have_full)
    nl += launch_em_steps(L.d_chains.p, L.d_descs.p, d_idx_full, full.n_small, full.n_big, any_scan, full.big_k, d_refs_full,
                          full.n_refs_chunk, int64_t(full.refs.size()) - full.n_refs_chunk, L.d_utrs.p, L.d_chain_off.p, L.d_tensor.p, h->tensor_f32, L.d_cnt.p,
                          L.d_lz.p, d_v, L.d_partials.p, L.d_counter.p, L.d_trace_a.p, L.d_trace_b.p, L.d_trace_ws.p, L.st,
                          L.em_events, plan, SCAPE_B200_NROUND + 1);
  else
    L.em_events.kinds.clear();
  for (int p = 1; p < n_parts; p++)
    if (!parts[p - 1].index.empty()) CU(cudaStreamWaitEvent(L.st, L.ev_join[p - 1], 0));
  CU(cudaGetLastError());
  CU(cudaEventRecord(defer ? L.ev_def[1] : L.ev[5], L.st));
  // (pinned destinations: a pageable one would make the copy call wait for the whole run)
  double& scan_elems = L.h_scalars.p[defer ? 1 : 0];
  CU(cudaMemcpyAsync(chains.data(), L.d_chains.p, sizeof(ChainDev) * chains.size(), cudaMemcpyDeviceToHost, L.st));
  CU(cudaMemcpyAsync(&scan_elems, L.d_counter.p, sizeof(double), cudaMemcpyDeviceToHost, L.st));
  const double t_rc2 = now_ms();
  if (enqueued)
    if (int rc = enqueued()) { cudaStreamSynchronize(L.st); return rc; }
  const double t_rc3 = now_ms();
  L.tm.launches += nl;
  if (defer) {
    CU(cudaEventRecord(L.ev_def[2], L.st));
    return 0;
  }
  CU(cudaStreamSynchronize(L.st));
  if (host_dbg_rc && chains.size() >= 1000) {
    static double acc[5] = {0, 0, 0, 0, 0};
    static int n_acc = 0;
    acc[0] += t_rc1 - t_rc0; acc[1] += t_rc2 - t_rc1; acc[2] += t_rc3 - t_rc2; acc[3] += now_ms() - t_rc3;
    if (++n_acc % 100 == 0) {
      fprintf(stderr, "run_chains host wall per call (us): prepare+upload %.0f | enqueue launches %.0f | callback %.0f | wait %.0f\n",
              10 * acc[0], 10 * acc[1], 10 * acc[2], 10 * acc[3]);
      acc[0] = acc[1] = acc[2] = acc[3] = 0;
    }
  }
  L.tm.d2h_bytes += double(sizeof(ChainDev) * chains.size());
  float ms = 0, t0 = 0;
  CU(cudaEventElapsedTime(&ms, L.ev[4], L.ev[5]));
  CU(cudaEventElapsedTime(&t0, h->base_ev, L.ev[4]));
  L.tm.em_ms += ms;
  L.busy.emplace_back(t0, t0 + ms);
  for (int p = 1; p < kMaxSplit; p++)
    if (p >= n_parts) L.em_events_part[p - 1].kinds.clear();
  for (const EmStepEvents* ee : {&L.em_events, &L.em_events2, &L.em_events_part[0], &L.em_events_part[1], &L.em_events_part[2]}) {
    double e_ms = 0, s_ms = 0;
    em_steps_elapsed(*ee, &e_ms, &s_ms);
    L.tm.estep_ms += e_ms;
    L.tm.scan_ms += s_ms;
    L.tm.scan_launches += ee->kinds.empty() ? 0 : ee->scan_launches;
  }
  if (resident_timed) {
    float cms = 0;
    CU(cudaEventElapsedTime(&cms, L.ev_cl[0], L.ev_cl[1]));
    L.tm.resident_ms += cms;
    L.tm.resident_launches += 1;
  }
  for (auto& c : chains)
    if (c.error) return fail(-7, "non-finite grid-search scores (a NaN reached max_alpha_beta)");
  for (auto& c : chains) {
    const UtrDev& u = utrs_host[size_t(c.utr)];
    L.tm.em_grid_bytes += c.grid_rows * double(u.N) * 8.0;     // FP64 tensor, one chain at a time
    L.tm.em_grid_flops += c.grid_rows * double(u.N) * 2.0;
    if (route[size_t(c.utr)] != kRouteBsp) L.tm.resident_grid_flops += (c.grid_rows - c.grid_rows_head) * double(u.N) * 2.0;
  }
  L.tm.em_scan_bytes += scan_elems * (h->tensor_f32 ? 4.0 : 8.0);   // what the batched scans really load
  return 0;
}

// Collect a deferred run: wait for its downloads, then the accounting run_chains does after its own
// synchronisation (deferred runs are never timed per step and never take the resident routes).
int finish_deferred_run(scape_b200_handle* h, Lane& L, const std::vector<UtrDev>& utrs_host) {
  if (!L.deferred.active) return 0;
  L.deferred.active = false;
  CU(cudaEventSynchronize(L.ev_def[2]));
  if (L.deferred.n == 0) return 0;
  L.tm.d2h_bytes += double(sizeof(ChainDev) * L.deferred.n);
  float ms = 0, t0 = 0;
  CU(cudaEventElapsedTime(&ms, L.ev_def[0], L.ev_def[1]));
  CU(cudaEventElapsedTime(&t0, h->base_ev, L.ev_def[0]));
  L.tm.em_ms += ms;
  L.busy.emplace_back(t0, t0 + ms);
  for (size_t i = 0; i < L.deferred.n; i++) {
    const ChainDev& c = L.deferred.chains[i];
    if (c.error) return fail(-7, "non-finite grid-search scores (a NaN reached max_alpha_beta)");
    const UtrDev& u = utrs_host[size_t(c.utr)];
    L.tm.em_grid_bytes += c.grid_rows * double(u.N) * 8.0;
    L.tm.em_grid_flops += c.grid_rows * double(u.N) * 2.0;
  }
  L.tm.em_scan_bytes += L.h_scalars.p[1] * (h->tensor_f32 ? 4.0 : 8.0);
  return 0;
}

}  // namespace

namespace {

struct FitShared {
  scape_b200_handle* h;
  const scape_b200_batch* bt;
  scape_b200_results* out;
  std::vector<UtrPrep>* prep;
  std::vector<std::atomic<int>>* ready;
  std::vector<std::vector<int64_t>>* stream_utrs;
  std::vector<NpRandomState>* rng;
  std::vector<size_t>* cursor;
  std::vector<int32_t>* stream_of;
  int maxwin;
};

// The wave loop of one lane over its own streams.
int run_lane(FitShared& F, Lane& L, const std::vector<int>& my_streams) {
  scape_b200_handle* h = F.h;
  const scape_b200_batch* bt = F.bt;
  scape_b200_results* out = F.out;
  const scape_b200_params& P = h->P;
  std::vector<UtrPrep>& prep = *F.prep;
  std::vector<std::vector<int64_t>>& stream_utrs = *F.stream_utrs;
  std::vector<NpRandomState>& rng = *F.rng;
  std::vector<size_t>& cursor = *F.cursor;
  std::vector<int32_t>& stream_of = *F.stream_of;
  const int maxwin = F.maxwin;
  CU(cudaSetDevice(h->device));
  const bool overlap = h->overlap;
  const size_t esz = h->tensor_f32 ? 4 : 8;

  // Pick the next wave and put its likelihood phase (uploads, table, tensor) on the lane's second
  // stream, into the staged buffer set.  Nothing here depends on the RNG, so it runs one wave ahead.
  std::vector<WaveUtr> st_wave;
  std::vector<UtrDev> st_ud;
  int st_max_n = 0;
  unsigned stage_no = 0;
  auto stage = [&]() -> int {
    StagedBufs& S = L.staged;
    cudaStream_t sq = L.st_lik;
    // ---- pick the wave: next UTR of every stream, within the memory budget ---------------------
    st_wave.clear();
    st_ud.clear();
    st_max_n = 0;
    double bytes = 0;
    for (int s : my_streams) {
      while (cursor[size_t(s)] < stream_utrs[size_t(s)].size()) {
        int64_t u = stream_utrs[size_t(s)][cursor[size_t(s)]];
        while (!(*F.ready)[size_t(u)].load(std::memory_order_acquire))      // pre-pass still running
          std::this_thread::sleep_for(std::chrono::microseconds(50));
        if (prep[size_t(u)].status != kOk) { cursor[size_t(s)]++; continue; }   // the reference would have raised
        const UtrPrep& p = prep[size_t(u)];
        double need = double(p.T() * p.B() + 4) * pad4(p.n()) * double(esz) + double(p.T()) * pad4(p.n()) * 8.0;
        if (!st_wave.empty() && bytes + need > h->wave_budget_bytes) break;
        bytes += need;
        WaveUtr w;
        w.u = u;
        w.k_max = P.fixed_run_mode ? P.pre_K : P.n_max_apa;
        w.k_min = P.fixed_run_mode ? P.pre_K : P.n_min_apa;
        memset(&w.best, 0, sizeof(ChainDev));
        st_wave.push_back(w);
        cursor[size_t(s)]++;
        break;
      }
    }
    if (st_wave.empty()) return 0;
    L.tm.waves++;

    // ---- device layout of the wave --------------------------------------------------------------
    const size_t W = st_wave.size();
    std::vector<UtrDev>& ud = st_ud;
    ud.resize(W);
    int64_t nf = 0, nt = 0, ntab = 0, nten = 0, n_rows = 0, n_trows = 0, n_tiles = 0, nr = 0;
    int max_n = 0;
    const int i_lo = kTfHalf;
    for (size_t i = 0; i < W; i++) {
      const UtrPrep& p = prep[size_t(st_wave[i].u)];
      UtrDev& d = ud[i];
      d.N = int32_t(p.n()); d.Npad = pad4(p.n()); d.T = int32_t(p.T()); d.B = int32_t(p.B());
      d.ldR = pad4(int64_t(d.T) * d.B);
      d.frag_off = nf; d.theta_off = nt; d.table_off = ntab; d.tensor_off = nten;
      d.n_reads = int32_t(p.n_reads); d.pad_ = 0; d.read_off = nr;
      nr += p.n_reads;
      d.unif_loglik = p.unif_loglik;
      nf += d.Npad; nt += d.T; ntab += int64_t(d.T) * d.Npad; nten += d.ldR * d.N;
      max_n = std::max(max_n, d.Npad);
      // marginal tensor: interior alpha rows go to the constant-weight kernel in tiles, the rest
      // (window clipped by the grid ends) to the generic kernel
      const int n_int = !h->tensor_fast ? 0 : h->tensor_fast_edges ? d.T : std::max(0, d.T - 2 * kTfHalf);
      n_rows += d.T;
      n_trows += d.T - n_int;
      n_tiles += (n_int + kTfTile - 1) / kTfTile;
    }
    st_max_n = max_n;
    // one pinned blob per wave parity: the uploads are truly asynchronous (a pageable source would
    // make cudaMemcpyAsync wait for the stream, which is parked behind the EM steps of this wave)
    PinnedBuf<char>& blob = L.h_stage[stage_no++ & 1];
    auto up16 = [](size_t v) { return (v + 15) / 16 * 16; };
    const size_t fb = sizeof(double) * size_t(nf);
    size_t o_x = 0, o_l = o_x + up16(fb), o_r = o_l + up16(fb), o_pa = o_r + up16(fb), o_c = o_pa + up16(fb),
           o_th = o_c + up16(fb), o_ud = o_th + up16(8 * size_t(nt)), o_rows = o_ud + up16(sizeof(UtrDev) * W),
           o_trows = o_rows + up16(sizeof(RowRef) * size_t(n_rows)),
           o_tiles = o_trows + up16(sizeof(RowRef) * size_t(n_trows)),
           o_r2b = o_tiles + up16(sizeof(TileRef) * size_t(n_tiles)), o_end = o_r2b + up16(sizeof(int32_t) * size_t(nr));
    CU(blob.resize(o_end));
    double *hx = (double*)(blob.p + o_x), *hl = (double*)(blob.p + o_l), *hr = (double*)(blob.p + o_r),
           *hpa = (double*)(blob.p + o_pa), *hc = (double*)(blob.p + o_c), *hth = (double*)(blob.p + o_th);
    UtrDev* hud = (UtrDev*)(blob.p + o_ud);
    RowRef *rows = (RowRef*)(blob.p + o_rows), *trows = (RowRef*)(blob.p + o_trows);
    TileRef* tiles = (TileRef*)(blob.p + o_tiles);
    memset(blob.p, 0, o_th);     // fragment columns are padded with zeros
    {
      int64_t ir = 0, it = 0, il = 0;
      for (size_t i = 0; i < W; i++) {
        const UtrPrep& p = prep[size_t(st_wave[i].u)];
        const UtrDev& d = ud[i];
        std::copy(p.x.begin(), p.x.end(), hx + d.frag_off);
        std::copy(p.l.begin(), p.l.end(), hl + d.frag_off);
        std::copy(p.r.begin(), p.r.end(), hr + d.frag_off);
        std::copy(p.pa.begin(), p.pa.end(), hpa + d.frag_off);
        std::copy(p.cnt.begin(), p.cnt.end(), hc + d.frag_off);
        std::copy(p.theta.begin(), p.theta.end(), hth + d.theta_off);
        std::copy(p.read_to_bin.begin(), p.read_to_bin.end(), (int32_t*)(blob.p + o_r2b) + d.read_off);
        hud[i] = d;
        const bool all_fast = h->tensor_fast && h->tensor_fast_edges;
        const int t_lo = all_fast ? 0 : i_lo, t_hi = all_fast ? d.T - 1 : d.T - 1 - kTfHalf;
        for (int t = 0; t < d.T; t++) rows[ir++] = {int32_t(i), t};
        for (int t = 0; t < d.T; t++)
          if (!h->tensor_fast || t < t_lo || t > t_hi) trows[it++] = {int32_t(i), t};
        if (h->tensor_fast)
          for (int t = t_lo; t <= t_hi; t += kTfTile) tiles[il++] = {int32_t(i), t, std::min(kTfTile, t_hi - t + 1)};
      }
      if (ir != n_rows || it != n_trows || il != n_tiles) return fail(-5, "internal: wave row lists out of step");
    }
    CU(S.d_fx.ensure(size_t(nf))); CU(S.d_fl.ensure(size_t(nf))); CU(S.d_fr.ensure(size_t(nf)));
    CU(S.d_fpa.ensure(size_t(nf))); CU(S.d_cnt.ensure(size_t(nf))); CU(S.d_theta.ensure(size_t(nt)));
    int64_t max_ldr = 0;
    for (size_t i = 0; i < W; i++) max_ldr = std::max<int64_t>(max_ldr, ud[i].ldR);
    const int64_t slack = int64_t(kTensorSlackRows) * max_ldr;     // elements; see scan_subbatch's register ring
    CU(S.d_table.ensure(size_t(ntab))); CU(S.d_tensor.ensure(size_t(nten + slack)));
    if (h->poison) CU(cudaMemsetAsync(S.d_tensor.p, 0xFF, S.d_tensor.cap * sizeof(double), sq));   // every byte the kernels do not write is NaN
    CU(cudaMemsetAsync((char*)S.d_tensor.p + size_t(nten) * esz, 0, size_t(slack) * esz, sq));
    CU(S.d_utrs.ensure(W)); CU(S.d_rows.ensure(size_t(n_rows)));
    CU(S.d_trows.ensure(size_t(n_trows) + 1)); CU(S.d_tiles.ensure(size_t(n_tiles) + 1));
    CU(S.d_r2b.ensure(size_t(nr) + 1));
    if (nr > 0) CU(cudaMemcpyAsync(S.d_r2b.p, blob.p + o_r2b, sizeof(int32_t) * size_t(nr), cudaMemcpyHostToDevice, sq));
    L.tm.h2d_bytes += double(sizeof(int32_t) * size_t(nr));
    CU(cudaMemcpyAsync(S.d_fx.p, hx, fb, cudaMemcpyHostToDevice, sq));
    CU(cudaMemcpyAsync(S.d_fl.p, hl, fb, cudaMemcpyHostToDevice, sq));
    CU(cudaMemcpyAsync(S.d_fr.p, hr, fb, cudaMemcpyHostToDevice, sq));
    CU(cudaMemcpyAsync(S.d_fpa.p, hpa, fb, cudaMemcpyHostToDevice, sq));
    CU(cudaMemcpyAsync(S.d_cnt.p, hc, fb, cudaMemcpyHostToDevice, sq));
    CU(cudaMemcpyAsync(S.d_theta.p, hth, sizeof(double) * size_t(nt), cudaMemcpyHostToDevice, sq));
    CU(cudaMemcpyAsync(S.d_utrs.p, hud, sizeof(UtrDev) * W, cudaMemcpyHostToDevice, sq));
    CU(cudaMemcpyAsync(S.d_rows.p, rows, sizeof(RowRef) * size_t(n_rows), cudaMemcpyHostToDevice, sq));
    if (n_trows > 0)
      CU(cudaMemcpyAsync(S.d_trows.p, trows, sizeof(RowRef) * size_t(n_trows), cudaMemcpyHostToDevice, sq));
    if (n_tiles > 0)
      CU(cudaMemcpyAsync(S.d_tiles.p, tiles, sizeof(TileRef) * size_t(n_tiles), cudaMemcpyHostToDevice, sq));
    L.tm.h2d_bytes += double(sizeof(RowRef) * size_t(n_trows) + sizeof(TileRef) * size_t(n_tiles));
    L.tm.h2d_bytes += double(5 * fb + sizeof(double) * size_t(nt) + sizeof(UtrDev) * W + sizeof(RowRef) * size_t(n_rows));

    // ---- likelihood phases ----------------------------------------------------------------------
    CU(cudaEventRecord(S.ev[0], sq));
    launch_table(S.d_utrs.p, S.d_rows.p, n_rows, max_n, S.d_fx.p, S.d_fl.p, S.d_fr.p,
                 S.d_fpa.p, S.d_theta.p, S.d_table.p, sq);
    CU(cudaEventRecord(S.ev[1], sq));
    launch_tensor(S.d_utrs.p, S.d_trows.p, n_trows, max_n, P.n_beta, maxwin, S.d_theta.p,
                  S.d_table.p, S.d_tensor.p, h->tensor_f32, sq);
    launch_tensor_interior(S.d_utrs.p, S.d_tiles.p, n_tiles, max_n, S.d_table.p, S.d_tensor.p,
                           h->tensor_f32, sq);
    CU(cudaEventRecord(S.ev[2], sq));
    CU(cudaGetLastError());
    L.tm.launches += 1 + (n_trows == 0 ? 0 : 1) + (n_tiles == 0 ? 0 : 1);
    for (size_t i = 0; i < W; i++) {
      // exp() evaluations = N * sum over (t, beta) of the clipped window sizes (regular grid)
      const UtrDev& d = ud[i];
      double win = 0;
      for (int j = 0; j < P.n_beta; j++) {
        const int half = int(std::floor(3 * P.betas[j] / P.theta_step));
        for (int t = 0; t < d.T; t++) win += std::min(d.T - 1, t + half) - std::max(0, t - half) + 1;
      }
      L.tm.tensor_exp += double(d.N) * win;
      L.tm.table_exp += double(d.N) * d.T * P.n_s;
    }
    return 0;
  };

  // RNG replay of the initial chains of a wave (apa_core.py:781-829, 655-677): serial per stream,
  // streams are independent -> one task per UTR of the wave, each writing its chains in place into
  // the lane's pinned staging buffer.  Returns the number of chains (UTRs whose initialisation
  // raised inside the reference are marked done and their chains dropped).
  auto draw_wave = [&](std::vector<WaveUtr>& wv, size_t& n_chains_out) -> int {
    const size_t W = wv.size();
    std::vector<size_t> first_chain(W + 1, 0);
    for (size_t i = 0; i < W; i++)
      first_chain[i + 1] = first_chain[i] + (wv[i].done ? 0 : size_t(wv[i].k_max - wv[i].k_min + 1) * SCAPE_B200_NTRIAL);
    CU(L.h_chains.resize(first_chain[W]));
    ChainDev* chains = L.h_chains.p;
    std::atomic<int> n_failed(0);
    L.pool->run(int64_t(W), [&](int64_t ii) {
      const size_t i = size_t(ii);
      WaveUtr& w = wv[i];
      if (w.done) return;
      const UtrPrep& p = prep[size_t(w.u)];
      NpRandomState& g = rng[size_t(stream_of[size_t(w.u)])];
      ChainDev* mine = chains + first_chain[i];
      size_t j = 0;
      for (int K = w.k_max; K >= w.k_min && !w.done; K--)
        for (int trial = 0; trial < SCAPE_B200_NTRIAL; trial++, j++) {
          ChainInit ci;
          int32_t rc = draw_chain(g, P, p, K, ci);
          if (rc != kOk) {       // numpy's choice() would have raised inside the reference
            out->status[w.u] = rc;
            w.done = true;
            n_failed++;
            break;
          }
          ChainDev& c = mine[j];
          memset(&c, 0, sizeof(c));
          c.utr = int32_t(i); c.K = K; c.weights_only = 0;
          memcpy(c.a_idx, ci.a_idx, sizeof(ci.a_idx));
          memcpy(c.b_idx, ci.b_idx, sizeof(ci.b_idx));
          memcpy(c.ws, ci.ws, sizeof(ci.ws));
          memcpy(c.k_order, ci.k_order, SCAPE_B200_NROUND);
        }
    });
    size_t n_chains = first_chain[W];
    if (n_failed.load() > 0) {   // rare: drop the chains of UTRs whose initialisation raised
      size_t o = 0;
      for (size_t i = 0; i < W; i++) {
        const size_t cnt = first_chain[i + 1] - first_chain[i];
        if (cnt == 0) continue;
        if (!wv[i].done) {
          if (o != first_chain[i]) memmove(chains + o, chains + first_chain[i], cnt * sizeof(ChainDev));
          o += cnt;
        }
      }
      n_chains = o;
    }
    n_chains_out = n_chains;
    return 0;
  };
  // The initial chains of the NEXT wave are drawn by a helper thread while the GPU runs this wave's
  // prune refits and labels: once a wave's selection is done and no UTR of it re-runs, the RNG
  // streams are where the next UTRs start.
  static const bool predraw_on = getenv("SCAPE_B200_PREDRAW") ? atoi(getenv("SCAPE_B200_PREDRAW")) != 0 : true;
  std::thread predraw_thread;
  struct Joiner {            // no early return may leave the helper running
    std::thread& t;
    ~Joiner() { if (t.joinable()) t.join(); }
  } predraw_joiner{predraw_thread};
  bool predrawn = false;
  size_t predrawn_chains = 0;
  int predraw_rc = 0;
  double predraw_ms = 0;

  // development aid (SCAPE_B200_DBG_HOST=1): wall clock of the wave loop's sections
  static const bool host_dbg = scape_env_on("SCAPE_B200_DBG_HOST");
  double hp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double hp_mark = now_ms();
  auto lap = [&](int k) { const double t = now_ms(); hp[k] += t - hp_mark; hp_mark = t; };
  struct HostProfile {
    const bool on; double* hp; int64_t* waves;
    ~HostProfile() {
      if (on)
        fprintf(stderr, "host wall per wave (us): draw %.0f | EM run_chains (incl. GPU wait) %.0f | select+refit draws %.0f | refit run %.0f | "
                        "labels launch+wait %.0f | assemble+expand %.0f | stage/predraw join %.0f  (%lld waves)\n",
                1e3 * hp[0] / double(std::max<int64_t>(*waves, 1)), 1e3 * hp[1] / double(std::max<int64_t>(*waves, 1)),
                1e3 * hp[2] / double(std::max<int64_t>(*waves, 1)), 1e3 * hp[3] / double(std::max<int64_t>(*waves, 1)),
                1e3 * hp[4] / double(std::max<int64_t>(*waves, 1)), 1e3 * hp[5] / double(std::max<int64_t>(*waves, 1)),
                1e3 * hp[6] / double(std::max<int64_t>(*waves, 1)), (long long)*waves);
    }
  } host_profile{host_dbg, hp, &L.tm.waves};
  std::function<int()> pending_assemble;      // deferred collection (refits, labels) + result assembly of the previous wave
  struct FlushPending {
    std::function<int()>& f;
    ~FlushPending() { if (f) f(); }            // (error paths: the stream is drained, the results are dropped with the error)
  } flush_pending{pending_assemble};
  // The prune refits + labels of a wave's last sweep are not waited for: the host goes on to the next
  // wave (draws joined, chain records, launch lists, uploads) while they run, and collects them from the
  // next run's `enqueued` callback.  Off in un-pipelined / per-step-timed passes.
  static const bool defer_env = getenv("SCAPE_B200_DEFER") ? atoi(getenv("SCAPE_B200_DEFER")) != 0 : true;
  static const int ev_env_lane = getenv("SCAPE_B200_STEP_EVENTS") ? atoi(getenv("SCAPE_B200_STEP_EVENTS")) : -1;
  const bool defer_ok = defer_env && overlap && !(ev_env_lane >= 0 ? ev_env_lane != 0 : !overlap);
  if (int rc = stage()) return rc;
  for (;;) {
    if (st_wave.empty()) break;
    // the staged wave becomes the current one; the EM stream waits for its tensor
    std::vector<WaveUtr> wave = std::move(st_wave);
    const std::vector<UtrDev> ud = std::move(st_ud);
    const int max_n = st_max_n;
    const size_t W = wave.size();
    L.adopt_staged();
    CU(cudaStreamWaitEvent(L.st, L.ev[2], 0));
    // The next wave's likelihood phase is slotted under this wave's EM: from step `stage_step` on
    // (most chains have converged by then and the step kernels leave SMs idle), or right away.
    bool next_staged = false;
    int stage_rc = 0;
    if (overlap && L.stage_step < 0) {
      if (int rc = stage()) return rc;
      next_staged = true;
    }
    bool timed_lik = false;

    std::vector<LabelDev> jobs;
    bool labels_enqueued = false, final_sweep = false;
    size_t deferred_refits = 0;
    std::vector<size_t> deferred_owner;
    bool wave_deferred = false;
    // ---- sweeps: main K range, then re-run ranges while K == n_max (apa_core.py:1023-1030) -------
    for (;;) {
      double tr0 = now_ms();
      size_t n_chains = 0;
      if (predrawn) {            // drawn while the previous wave finished on the GPU
        predrawn = false;
        n_chains = predrawn_chains;
      } else if (int rc = draw_wave(wave, n_chains)) {
        return rc;
      }
      ChainDev* chains = L.h_chains.p;
      L.tm.host_rng_ms += now_ms() - tr0;
      lap(0);
      if (n_chains == 0) break;
      if (overlap && !next_staged) {
        L.em_events.hook_step = std::min(L.stage_step, SCAPE_B200_NROUND - 1);
        L.em_events.mark = [&]() {
          cudaEventRecord(L.ev_mid, L.st);
          cudaStreamWaitEvent(L.st_lik, L.ev_mid, 0);
        };
        L.em_events.hook = [&]() {
          stage_rc = stage();
          next_staged = true;
        };
      }
      {
        // the previous wave's results are written out while the GPU runs this wave's EM
        std::function<int()> cb;
        if (pending_assemble) cb = [&]() { const int rc = pending_assemble(); pending_assemble = nullptr; return rc; };
        const int rc = run_chains(h, L, chains, n_chains, ud, false, cb);
        L.em_events.hook = nullptr;
        L.em_events.mark = nullptr;
        if (rc) return rc;
        if (stage_rc) return stage_rc;
      }
      lap(1);
      if (!timed_lik) {
        float a = 0, b = 0;
        CU(cudaEventElapsedTime(&a, L.ev[0], L.ev[1]));
        CU(cudaEventElapsedTime(&b, L.ev[1], L.ev[2]));
        L.tm.table_ms += a;
        L.tm.tensor_ms += b;
        float tb = 0;
        CU(cudaEventElapsedTime(&tb, h->base_ev, L.ev[0]));
        L.busy.emplace_back(tb, tb + a + b);
        timed_lik = true;
      }
      // ---- selection (em_optim0 :865, run :972) + pruning (rm_component :832-844) -----------------
      CU(L.h_refits.resize(W));
      ChainDev* refits = L.h_refits.p;
      size_t n_refits = 0;
      std::vector<size_t> refit_owner;
      tr0 = now_ms();
      std::vector<size_t> pos_of(W, 0);
      {
        size_t pos = 0;
        for (size_t i = 0; i < W; i++)
          if (!wave[i].done) {
            pos_of[i] = pos;
            pos += size_t(wave[i].k_max - wave[i].k_min + 1) * SCAPE_B200_NTRIAL;
          }
      }
      std::vector<char> need_refit(W, 0);
      L.pool->run(int64_t(W), [&](int64_t ii) {     // one task per UTR: its own chains, its own RNG stream
        const size_t i = size_t(ii);
        WaveUtr& w = wave[i];
        if (w.done) return;
        const size_t pos = pos_of[i];
        const int nK = w.k_max - w.k_min + 1;
        std::vector<double> bic_k(static_cast<size_t>(nK));
        std::vector<size_t> best_k(static_cast<size_t>(nK));
        std::vector<double> b(SCAPE_B200_NTRIAL);
        for (int ik = 0; ik < nK; ik++) {
          for (int t = 0; t < SCAPE_B200_NTRIAL; t++) {
            const ChainDev& c = chains[pos + size_t(ik) * SCAPE_B200_NTRIAL + size_t(t)];
            b[size_t(t)] = c.bic;
            w.work += double(c.n_iter) * ud[i].N * (c.K + 1);
            w.iters += c.n_iter;
          }
          int m = np_argmin(b);
          best_k[size_t(ik)] = pos + size_t(ik) * SCAPE_B200_NTRIAL + size_t(m);
          bic_k[size_t(ik)] = b[size_t(m)];
        }
        w.chains_run += nK * SCAPE_B200_NTRIAL;
        w.sweeps++;
        w.best = chains[best_k[size_t(np_argmin(bic_k))]];
        w.k_selected = w.best.K;
        if (!P.fixed_run_mode) {
          int keep[SCAPE_B200_KCAP], nk = 0;
          for (int k = 0; k < w.best.K; k++)
            if (!(w.best.ws[k] < P.min_ws)) keep[nk++] = k;
          if (nk < w.best.K) {
            ChainDev& c = refits[i];          // slot i; compacted below
            memset(&c, 0, sizeof(c));
            c.utr = int32_t(i); c.K = nk; c.weights_only = 1;
            for (int k = 0; k < nk; k++) { c.a_idx[k] = w.best.a_idx[keep[k]]; c.b_idx[k] = w.best.b_idx[keep[k]]; }
            ChainInit ci;
            ci.K = nk;
            draw_refit(rng[size_t(stream_of[size_t(w.u)])], P, ci);
            memcpy(c.ws, ci.ws, sizeof(ci.ws));
            memcpy(c.k_order, ci.k_order, SCAPE_B200_NROUND);
            need_refit[i] = 1;
          }
        }
      });
      for (size_t i = 0; i < W; i++)
        if (need_refit[i]) {
          if (n_refits != i) refits[n_refits] = refits[i];
          refit_owner.push_back(i);
          n_refits++;
        }
      L.tm.host_rng_ms += now_ms() - tr0;
      {
        bool any_rerun = false;
        for (size_t i = 0; i < W; i++)
          if (!wave[i].done && !need_refit[i] && !P.fixed_run_mode && P.re_run_mode && wave[i].best.K == wave[i].k_max)
            any_rerun = true;
        final_sweep = !any_rerun;
        if (predraw_on && !any_rerun && next_staged && !st_wave.empty() && !predraw_thread.joinable()) {
          predraw_thread = std::thread([&]() {
            const double t0 = now_ms();
            predraw_rc = draw_wave(st_wave, predrawn_chains);
            predraw_ms = now_ms() - t0;
          });
        }
      }
      lap(2);
      // ---- labels (get_label :873-881) + per-read expansion (:976), enqueued right behind the prune refits
      // when this is the last sweep of the wave: the label kernel takes a refitted UTR's parameters
      // from the refit's device record, so refits + labels cost ONE stream synchronisation
      auto enqueue_labels = [&]() -> int {
        std::vector<int32_t> refit_of(W, -1);
        for (size_t j = 0; j < n_refits; j++) refit_of[refit_owner[j]] = int32_t(j);
        jobs.assign(W, LabelDev{});
        int64_t nl = 0;
        int max_reads = 0;
        for (size_t i = 0; i < W; i++) {
          LabelDev& j = jobs[i];
          memset(&j, 0, sizeof(j));
          j.utr = int32_t(i); j.K = wave[i].best.K; j.out_off = nl; j.chain = refit_of[i];
          memcpy(j.a_idx, wave[i].best.a_idx, sizeof(j.a_idx));
          memcpy(j.b_idx, wave[i].best.b_idx, sizeof(j.b_idx));
          memcpy(j.ws, wave[i].best.ws, sizeof(j.ws));
          nl += ud[i].N;
          max_reads = std::max(max_reads, int(ud[i].n_reads));
        }
        const int64_t n_reads_wave = W ? ud[W - 1].read_off + ud[W - 1].n_reads : 0;
        CU(L.d_jobs.ensure(W));
        CU(L.d_labels.ensure(size_t(nl) + 1));
        CU(L.d_labels64.ensure(size_t(n_reads_wave) + 1));
        CU(L.h_labels64.resize(size_t(n_reads_wave) + 1));
        CU(L.h_jobs.resize(W));
        std::copy(jobs.begin(), jobs.end(), L.h_jobs.p);
        CU(cudaMemcpyAsync(L.d_jobs.p, L.h_jobs.p, sizeof(LabelDev) * W, cudaMemcpyHostToDevice, L.st));
        CU(cudaEventRecord(L.ev[6], L.st));
        launch_labels(L.d_jobs.p, int64_t(W), max_n, L.d_utrs.p, L.d_tensor.p, h->tensor_f32, L.d_cnt.p, L.d_chains.p,
                      L.d_labels.p, L.st);
        launch_label_expand(L.d_jobs.p, int64_t(W), max_reads, L.d_utrs.p, L.d_labels.p, L.d_r2b.p, L.d_labels64.p, L.st);
        CU(cudaEventRecord(L.ev[7], L.st));
        CU(cudaGetLastError());
        if (n_reads_wave > 0)
          CU(cudaMemcpyAsync(L.h_labels64.p, L.d_labels64.p, sizeof(int64_t) * size_t(n_reads_wave), cudaMemcpyDeviceToHost, L.st));
        L.tm.launches += 2;
        L.tm.h2d_bytes += double(sizeof(LabelDev) * W);
        L.tm.d2h_bytes += double(sizeof(int64_t) * size_t(n_reads_wave));
        labels_enqueued = true;
        return 0;
      };
      const bool deferred = final_sweep && defer_ok;
      {
        const std::function<int()> cb = final_sweep ? std::function<int()>(enqueue_labels) : std::function<int()>();
        if (int rc = run_chains(h, L, refits, n_refits, ud, false, cb, deferred)) {
          if (predraw_thread.joinable()) predraw_thread.join();
          return rc;
        }
      }
      if (deferred) {
        // (a last sweep never asks for a re-run: refitted UTRs have K < k_max, the others were checked above)
        for (size_t i = 0; i < W; i++) wave[i].done = true;
        deferred_refits = n_refits;
        deferred_owner = refit_owner;
        wave_deferred = true;
        break;
      }
      for (size_t j = 0; j < n_refits; j++) {
        WaveUtr& w = wave[refit_owner[j]];
        w.best = refits[j];
        w.chains_run += 1;
        w.work += double(refits[j].n_iter) * ud[refit_owner[j]].N * (refits[j].K + 1);
        w.iters += refits[j].n_iter;
      }
      bool again = false;
      for (size_t i = 0; i < W; i++) {
        WaveUtr& w = wave[i];
        if (w.done) continue;
        if (!P.fixed_run_mode && P.re_run_mode && w.best.K == w.k_max) {
          if (w.k_max + 2 > SCAPE_B200_KCAP) {
            out->status[w.u] = kErrKcap;
            w.done = true;
          } else {
            w.k_min = w.k_max;
            w.k_max += 2;
            again = true;
          }
        } else {
          w.done = true;
        }
      }
      if (!again) {
        if (!labels_enqueued) {         // (a sweep that expected a re-run and got none, e.g. K hit SCAPE_B200_KCAP)
          n_refits = 0;                 // every refit result is already in wave[i].best
          if (int rc = enqueue_labels()) return rc;
          CU(cudaStreamSynchronize(L.st));
        }
        break;
      }
    }
    lap(3);
    // label timing + the per-read labels out of the pinned download buffer (the next wave's download reuses it)
    auto collect_labels = [h, &L, out, bt](const std::vector<WaveUtr>& wv, const std::vector<UtrDev>& udv) -> int {
      float a = 0;
      CU(cudaEventElapsedTime(&a, L.ev[6], L.ev[7]));
      L.tm.label_ms += a;
      float tb = 0;
      CU(cudaEventElapsedTime(&tb, h->base_ev, L.ev[6]));
      L.busy.emplace_back(tb, tb + a);
      for (size_t i = 0; i < wv.size(); i++) {
        const int64_t u = wv[i].u;
        memcpy(out->label + bt->read_off[u], L.h_labels64.p + udv[i].read_off, sizeof(int64_t) * size_t(udv[i].n_reads));
      }
      return 0;
    };
    // ---- results: assembled on the host while the GPU runs the next wave's EM (or at the end) ---------
    {
      std::shared_ptr<std::vector<WaveUtr>> wv = std::make_shared<std::vector<WaveUtr>>(std::move(wave));
      std::shared_ptr<std::vector<UtrDev>> udv = std::make_shared<std::vector<UtrDev>>(ud);
      if (!wave_deferred)
        if (int rc = collect_labels(*wv, *udv)) return rc;
      lap(4);
      const size_t n_def = deferred_refits;
      const std::vector<size_t> owner = deferred_owner;
      pending_assemble = [&prep, out, wv, udv, h, &L, wave_deferred, n_def, owner, collect_labels]() -> int {
        if (wave_deferred) {
          // the refits + labels enqueued a wave ago: wait for their downloads (long there by now)
          if (int rc = finish_deferred_run(h, L, *udv)) return rc;
          const ChainDev* refits = L.h_refits.p;
          for (size_t j = 0; j < n_def; j++) {
            WaveUtr& w = (*wv)[owner[j]];
            w.best = refits[j];
            w.chains_run += 1;
            w.work += double(refits[j].n_iter) * (*udv)[owner[j]].N * (refits[j].K + 1);
            w.iters += refits[j].n_iter;
          }
          if (int rc = collect_labels(*wv, *udv)) return rc;
        }
        for (size_t i = 0; i < wv->size(); i++) {
          const WaveUtr& w = (*wv)[i];
          const UtrPrep& p = prep[size_t(w.u)];
          const int64_t u = w.u;
          const ChainDev& c = w.best;
          out->K[u] = c.K;
          for (int k = 0; k < c.K; k++) {
            out->alpha[u * SCAPE_B200_KCAP + k] = std::nearbyint(p.theta[size_t(c.a_idx[k])]);   // np.rint (:770)
            out->beta[u * SCAPE_B200_KCAP + k] = p.betas[size_t(c.b_idx[k])];
          }
          for (int k = 0; k <= c.K; k++) out->ws[u * (SCAPE_B200_KCAP + 1) + k] = c.ws[k];
          out->bic[u] = c.bic;
          out->n_lb[u] = c.n_iter;
          for (int k = 0; k < c.n_iter; k++) out->lb_arr[u * SCAPE_B200_NROUND + k] = c.lb_arr[k];
          out->path[u * 4 + 0] = w.sweeps; out->path[u * 4 + 1] = w.k_selected;
          out->path[u * 4 + 2] = c.K; out->path[u * 4 + 3] = w.chains_run;
          out->em_work[u * 2] = w.work; out->em_work[u * 2 + 1] = w.iters;
        }
        return 0;
      };
    }
    lap(5);
    if (predraw_thread.joinable()) {
      predraw_thread.join();
      if (predraw_rc) return predraw_rc;
      predrawn = true;
      L.tm.host_rng_ms += predraw_ms;
    }
    if (!next_staged)
      if (int rc = stage()) return rc;
    lap(6);
  }
  if (pending_assemble) {
    const int rc = pending_assemble();
    pending_assemble = nullptr;
    if (rc) return rc;
  }
  return 0;
}

}  // namespace

extern "C" int scape_b200_fit_batch(scape_b200_handle* h, const scape_b200_batch* bt, scape_b200_results* out) {
  if (!h || !bt || !out) return fail(-5, "null argument");
  CU(cudaSetDevice(h->device));
  const scape_b200_params& P = h->P;
  const int64_t U = bt->n_utr;
  memset(&h->tm, 0, sizeof(h->tm));
  const double t_begin = now_ms();
  CU(upload_model_const(h->mc));
  if (h->tensor_fast) CU(upload_tensor_fast_tables(h->tf_g, h->tf_lp, h->tf_lps, h->tf_hw));
  CU(cudaEventRecord(h->base_ev, 0));
  CU(cudaEventSynchronize(h->base_ev));

  // ---- host pre-pass (RNG free): background thread, parallel over UTRs in wave order -------------
  // UTRs are prepared in the order the waves need them (position within the stream first), each
  // one publishes a ready flag; the lanes start their first wave as soon as its UTRs are ready, so
  // the pre-pass of later waves overlaps the kernels of earlier ones.
  std::vector<UtrPrep> prep(static_cast<size_t>(U));
  std::vector<std::atomic<int>> ready(static_cast<size_t>(U));
  for (auto& r : ready) r.store(0, std::memory_order_relaxed);
  std::vector<int64_t> prep_order(static_cast<size_t>(U));
  {
    std::vector<int64_t> seen(size_t(std::max(1, bt->n_streams)), 0), rank(static_cast<size_t>(U));
    for (int64_t u = 0; u < U; u++) {
      const int s = bt->stream_id[u];
      if (s < 0 || s >= bt->n_streams) return fail(-5, "stream_id out of range");
      rank[size_t(u)] = seen[size_t(s)]++;
    }
    for (int64_t u = 0; u < U; u++) prep_order[size_t(u)] = u;
    std::stable_sort(prep_order.begin(), prep_order.end(),
                     [&](int64_t a, int64_t b) { return rank[size_t(a)] < rank[size_t(b)]; });
  }
  double prep_ms = 0;
  {
    const int nt = h->host_threads > 0 ? h->host_threads : default_host_threads();
    if (!h->prep_pool || h->prep_pool->threads() != nt) h->prep_pool.reset(new WorkPool(nt, true));
  }
  std::thread prep_thread([&]() {
    WorkPool::lower_priority();
    const double t0 = now_ms();
    h->prep_pool->run(U, [&](int64_t i) {
      const int64_t u = prep_order[size_t(i)];
      const int64_t a = bt->read_off[u], n = bt->read_off[u + 1] - a;
      UtrPrep& p = prep[size_t(u)];
      if (bin_reads(bt->x + a, bt->l + a, bt->r + a, bt->pa + a, n, p) == kOk &&
          setup_model(P, bt->x + a, bt->l + a, n, p) == kOk)
        coverage_and_peaks(P, p);
      out->status[u] = p.status;
      out->K[u] = 0;
      out->L[u] = p.L;
      out->n_frag[u] = int32_t(p.n());
      out->n_theta[u] = int32_t(p.T());
      out->n_lb[u] = 0;
      out->bic[u] = NAN;
      for (int k = 0; k < 4; k++) out->path[u * 4 + k] = 0;
      out->em_work[u * 2] = out->em_work[u * 2 + 1] = 0;
      ready[size_t(u)].store(1, std::memory_order_release);
    });
    prep_ms = now_ms() - t0;
  });
  struct Joiner {
    std::thread& t;
    ~Joiner() { if (t.joinable()) t.join(); }
  } prep_joiner{prep_thread};

  // ---- streams --------------------------------------------------------------------------------
  const int S = bt->n_streams;
  std::vector<std::vector<int64_t>> stream_utrs(static_cast<size_t>(S));
  for (int64_t u = 0; u < U; u++) {
    int s = bt->stream_id[u];
    if (s < 0 || s >= S) return fail(-5, "stream_id out of range");
    stream_utrs[size_t(s)].push_back(u);
  }
  std::vector<NpRandomState> rng;
  rng.reserve(size_t(S));
  for (int s = 0; s < S; s++) {
    if (bt->stream_state) {
      rng.emplace_back(0u);
      memcpy(rng.back().key, bt->stream_state + size_t(s) * 625, 624 * sizeof(uint32_t));
      rng.back().pos = int(bt->stream_state[size_t(s) * 625 + 624]);
      if (rng.back().pos < 0 || rng.back().pos > 624) return fail(-5, "stream_state: bad MT19937 position");
    } else {
      rng.emplace_back(bt->stream_seed[s]);
    }
  }
  std::vector<size_t> cursor(size_t(S), 0);

  const int maxwin = max_window(P);
  std::vector<int32_t> stream_of(static_cast<size_t>(U));
  for (int64_t u = 0; u < U; u++) stream_of[size_t(u)] = bt->stream_id[u];

  // ---- lanes: streams are dealt round-robin, every lane runs its waves in its own thread ---------
  const int lanes_wanted = h->n_lanes > 0 ? h->n_lanes : (S <= 16 ? 4 : S <= 40 ? 3 : 2);
  const int n_lanes = std::max(1, std::min(h->overlap ? lanes_wanted : 1, S));
  h->lanes_in_use = n_lanes;
  std::vector<std::vector<int>> lane_streams(static_cast<size_t>(n_lanes));
  for (int s = 0; s < S; s++) lane_streams[size_t(s % n_lanes)].push_back(s);
  const int total_threads = h->host_threads > 0 ? h->host_threads : default_host_threads();
  const int per_lane = std::max(1, total_threads / n_lanes);
  for (int l = 0; l < n_lanes; l++)
    if (!h->lanes[l].pool || h->lanes[l].pool->threads() != per_lane) h->lanes[l].pool.reset(new WorkPool(per_lane));
  FitShared F{h, bt, out, &prep, &ready, &stream_utrs, &rng, &cursor, &stream_of, maxwin};
  for (int l = 0; l < n_lanes; l++) {
    Lane& L = h->lanes[l];
    memset(&L.tm, 0, sizeof(L.tm));
    L.busy.clear();
    L.rc = 0;
    L.err.clear();
    L.deferred = Lane::Deferred{};     // (a fit that ended with an error may have left one open)
  }
  {
    std::vector<std::thread> workers;
    for (int l = 0; l < n_lanes; l++)
      workers.emplace_back([&, l]() {
        Lane& L = h->lanes[l];
        L.rc = run_lane(F, L, lane_streams[size_t(l)]);
        if (L.rc) L.err = g_err;
      });
    for (auto& w : workers) w.join();
  }
  prep_thread.join();
  h->tm.host_prep_ms = prep_ms;
  for (int l = 0; l < n_lanes; l++)
    if (h->lanes[l].rc) return fail(h->lanes[l].rc, h->lanes[l].err);
  // merge the lanes' accounting; device time = length of the union of all kernel intervals
  std::vector<std::pair<float, float>> iv;
  for (int l = 0; l < n_lanes; l++) {
    const scape_b200_timing& t = h->lanes[l].tm;
    h->tm.table_ms += t.table_ms; h->tm.tensor_ms += t.tensor_ms; h->tm.em_ms += t.em_ms; h->tm.label_ms += t.label_ms;
    h->tm.host_rng_ms += t.host_rng_ms; h->tm.launches += t.launches; h->tm.waves += t.waves;
    h->tm.em_grid_bytes += t.em_grid_bytes; h->tm.em_grid_flops += t.em_grid_flops; h->tm.tensor_exp += t.tensor_exp; h->tm.table_exp += t.table_exp;
    h->tm.h2d_bytes += t.h2d_bytes; h->tm.d2h_bytes += t.d2h_bytes; h->tm.em_scan_bytes += t.em_scan_bytes;
    h->tm.estep_ms += t.estep_ms; h->tm.scan_ms += t.scan_ms; h->tm.scan_launches += t.scan_launches;
    h->tm.resident_ms += t.resident_ms; h->tm.resident_launches += t.resident_launches; h->tm.resident_grid_flops += t.resident_grid_flops;
    iv.insert(iv.end(), h->lanes[l].busy.begin(), h->lanes[l].busy.end());
  }
  std::sort(iv.begin(), iv.end());
  double busy = 0, cur_a = 0, cur_b = -1;
  for (auto& p : iv) {
    if (p.first > cur_b) { if (cur_b > cur_a) busy += cur_b - cur_a; cur_a = p.first; cur_b = p.second; }
    else cur_b = std::max<double>(cur_b, p.second);
  }
  if (cur_b > cur_a) busy += cur_b - cur_a;
  h->tm.device_busy_ms = busy;     // device busy time: union of the kernel intervals of all lanes
  if (bt->stream_state)
    for (int s = 0; s < S; s++) {
      memcpy(bt->stream_state + size_t(s) * 625, rng[size_t(s)].key, 624 * sizeof(uint32_t));
      bt->stream_state[size_t(s) * 625 + 624] = uint32_t(rng[size_t(s)].pos);
    }
  h->tm.total_ms = now_ms() - t_begin;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// kernel-seam entry points for the parity tests
// ------------------------------------------------------------------------------------------------
extern "C" int scape_b200_loglik_table(scape_b200_handle* h, int64_t n_frag, const double* x, const double* l,
                                       const double* r, const double* pa, int64_t n_theta, const double* theta,
                                       double* table_out) {
  if (!h) return fail(-5, "null handle");
  Lane& L = h->lanes[0];
  CU(cudaSetDevice(h->device));
  CU(upload_model_const(h->mc));
  UtrDev d;
  memset(&d, 0, sizeof(d));
  d.N = int32_t(n_frag); d.Npad = pad4(n_frag); d.T = int32_t(n_theta); d.B = h->P.n_beta;
  d.ldR = pad4(int64_t(d.T) * d.B);
  std::vector<RowRef> rows;
  for (int t = 0; t < d.T; t++) rows.push_back({0, t});
  const size_t np_ = size_t(d.Npad);
  std::vector<double> px(np_, 0.0), pl(np_, 0.0), pr(np_, 0.0), ppa(np_, 0.0);
  std::copy(x, x + n_frag, px.begin()); std::copy(l, l + n_frag, pl.begin());
  std::copy(r, r + n_frag, pr.begin()); std::copy(pa, pa + n_frag, ppa.begin());
  CU(L.d_fx.ensure(np_)); CU(L.d_fl.ensure(np_)); CU(L.d_fr.ensure(np_)); CU(L.d_fpa.ensure(np_));
  CU(L.d_theta.ensure(size_t(n_theta))); CU(L.d_table.ensure(size_t(n_theta) * np_));
  CU(L.d_utrs.ensure(1)); CU(L.d_rows.ensure(rows.size()));
  CU(cudaMemcpyAsync(L.d_fx.p, px.data(), 8 * np_, cudaMemcpyHostToDevice, L.st));
  CU(cudaMemcpyAsync(L.d_fl.p, pl.data(), 8 * np_, cudaMemcpyHostToDevice, L.st));
  CU(cudaMemcpyAsync(L.d_fr.p, pr.data(), 8 * np_, cudaMemcpyHostToDevice, L.st));
  CU(cudaMemcpyAsync(L.d_fpa.p, ppa.data(), 8 * np_, cudaMemcpyHostToDevice, L.st));
  CU(cudaMemcpyAsync(L.d_theta.p, theta, 8 * size_t(n_theta), cudaMemcpyHostToDevice, L.st));
  CU(cudaMemcpyAsync(L.d_utrs.p, &d, sizeof(d), cudaMemcpyHostToDevice, L.st));
  CU(cudaMemcpyAsync(L.d_rows.p, rows.data(), sizeof(RowRef) * rows.size(), cudaMemcpyHostToDevice, L.st));
  launch_table(L.d_utrs.p, L.d_rows.p, int64_t(rows.size()), d.Npad, L.d_fx.p, L.d_fl.p, L.d_fr.p, L.d_fpa.p,
               L.d_theta.p, L.d_table.p, L.st);
  CU(cudaGetLastError());
  std::vector<double> tt(size_t(n_theta) * np_);
  CU(cudaMemcpyAsync(tt.data(), L.d_table.p, 8 * tt.size(), cudaMemcpyDeviceToHost, L.st));
  CU(cudaStreamSynchronize(L.st));
  for (int64_t n = 0; n < n_frag; n++)
    for (int64_t t = 0; t < n_theta; t++) table_out[n * n_theta + t] = tt[size_t(t) * np_ + size_t(n)];
  return 0;
}

extern "C" int scape_b200_marginal_tensor(scape_b200_handle* h, int64_t n_frag, int64_t n_theta, const double* theta,
                                          int64_t n_beta, const double* betas, const double* table,
                                          double* tensor_out) {
  if (!h) return fail(-5, "null handle");
  if (n_beta > SCAPE_B200_MAX_BETA) return fail(-5, "n_beta too large");
  Lane& L = h->lanes[0];
  CU(cudaSetDevice(h->device));
  ModelConst mc = h->mc;
  mc.n_beta = int32_t(n_beta);
  double bmax = 0;
  for (int i = 0; i < n_beta; i++) { mc.betas[i] = betas[i]; bmax = std::max(bmax, betas[i]); }
  CU(upload_model_const(mc));
  // widest window for an arbitrary sorted theta list
  int maxwin = 1;
  for (int64_t i = 0; i < n_theta; i++) {
    const double* lo = std::lower_bound(theta, theta + n_theta, theta[i] - 3 * bmax);
    const double* hi = std::upper_bound(theta, theta + n_theta, theta[i] + 3 * bmax);
    maxwin = std::max(maxwin, int(hi - lo));
  }
  UtrDev d;
  memset(&d, 0, sizeof(d));
  d.N = int32_t(n_frag); d.Npad = pad4(n_frag); d.T = int32_t(n_theta); d.B = int32_t(n_beta);
  d.ldR = pad4(int64_t(d.T) * d.B);
  const size_t np_ = size_t(d.Npad);
  std::vector<RowRef> rows;
  for (int t = 0; t < d.T; t++) rows.push_back({0, t});
  std::vector<double> tt(size_t(n_theta) * np_, 0.0);
  for (int64_t n = 0; n < n_frag; n++)
    for (int64_t t = 0; t < n_theta; t++) tt[size_t(t) * np_ + size_t(n)] = table[n * n_theta + t];
  CU(L.d_theta.ensure(size_t(n_theta))); CU(L.d_table.ensure(tt.size()));
  CU(L.d_tensor.ensure(size_t(d.ldR) * size_t(n_frag)));
  CU(L.d_utrs.ensure(1)); CU(L.d_rows.ensure(rows.size()));
  CU(cudaMemcpyAsync(L.d_theta.p, theta, 8 * size_t(n_theta), cudaMemcpyHostToDevice, L.st));
  CU(cudaMemcpyAsync(L.d_table.p, tt.data(), 8 * tt.size(), cudaMemcpyHostToDevice, L.st));
  CU(cudaMemcpyAsync(L.d_utrs.p, &d, sizeof(d), cudaMemcpyHostToDevice, L.st));
  CU(cudaMemcpyAsync(L.d_rows.p, rows.data(), sizeof(RowRef) * rows.size(), cudaMemcpyHostToDevice, L.st));
  // same dispatch as fit_batch: interior rows of a default-shaped regular grid take the constant-weight kernel
  bool fast = h->tensor_fast && n_beta == kTfB && (h->tensor_fast_edges || (maxwin == kTfW && n_theta > 2 * kTfHalf));
  for (int j = 0; fast && j < kTfB; j++) fast = betas[j] == h->P.betas[j];
  for (int64_t t = 1; fast && t < n_theta; t++) fast = theta[t] - theta[t - 1] == double(h->P.theta_step);
  if (fast) {
    CU(upload_tensor_fast_tables(h->tf_g, h->tf_lp, h->tf_lps, h->tf_hw));
    std::vector<RowRef> edge;
    std::vector<TileRef> tiles;
    const int i_lo = h->tensor_fast_edges ? 0 : kTfHalf, i_hi = h->tensor_fast_edges ? d.T - 1 : d.T - 1 - kTfHalf;
    for (int t = 0; t < d.T; t++)
      if (t < i_lo || t > i_hi) edge.push_back({0, t});
    for (int t = i_lo; t <= i_hi; t += kTfTile) tiles.push_back({0, t, std::min(kTfTile, i_hi - t + 1)});
    CU(L.d_trows.ensure(edge.size() + 1)); CU(L.d_tiles.ensure(tiles.size() + 1));
    CU(cudaMemcpyAsync(L.d_trows.p, edge.data(), sizeof(RowRef) * edge.size(), cudaMemcpyHostToDevice, L.st));
    CU(cudaMemcpyAsync(L.d_tiles.p, tiles.data(), sizeof(TileRef) * tiles.size(), cudaMemcpyHostToDevice, L.st));
    launch_tensor(L.d_utrs.p, L.d_trows.p, int64_t(edge.size()), d.Npad, int(n_beta), maxwin, L.d_theta.p, L.d_table.p,
                  L.d_tensor.p, h->tensor_f32, L.st);
    launch_tensor_interior(L.d_utrs.p, L.d_tiles.p, int64_t(tiles.size()), d.Npad, L.d_table.p, L.d_tensor.p,
                           h->tensor_f32, L.st);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(L.st));   // the host lists above must outlive the copies
  } else {
    launch_tensor(L.d_utrs.p, L.d_rows.p, int64_t(rows.size()), d.Npad, int(n_beta), maxwin, L.d_theta.p,
                  L.d_table.p, L.d_tensor.p, h->tensor_f32, L.st);
    CU(cudaGetLastError());
  }
  const size_t Rr = size_t(n_theta) * size_t(n_beta), ldr = size_t(d.ldR);
  std::vector<double> ten(ldr * size_t(n_frag));
  if (h->tensor_f32) {
    std::vector<float> tf(ten.size());
    CU(cudaMemcpyAsync(tf.data(), L.d_tensor.p, 4 * tf.size(), cudaMemcpyDeviceToHost, L.st));
    CU(cudaStreamSynchronize(L.st));
    for (size_t i = 0; i < tf.size(); i++) ten[i] = double(tf[i]);
  } else {
    CU(cudaMemcpyAsync(ten.data(), L.d_tensor.p, 8 * ten.size(), cudaMemcpyDeviceToHost, L.st));
    CU(cudaStreamSynchronize(L.st));
  }
  for (size_t n = 0; n < size_t(n_frag); n++)          // device [n][t][b] -> reference [t][b][n]
    for (size_t r = 0; r < Rr; r++) tensor_out[r * size_t(n_frag) + n] = ten[n * ldr + r];
  CU(upload_model_const(h->mc));
  return 0;
}

extern "C" int scape_b200_em_chains(scape_b200_handle* h, int64_t n_frag, int64_t n_theta, int64_t n_beta,
                                    const double* tensor, const double* cnt, double unif_loglik, int64_t n_chains,
                                    scape_b200_chain_io* io, int32_t* trace_a, int32_t* trace_b, double* trace_ws) {
  if (!h) return fail(-5, "null handle");
  Lane& L = h->lanes[0];
  CU(cudaSetDevice(h->device));
  CU(upload_model_const(h->mc));
  UtrDev d;
  memset(&d, 0, sizeof(d));
  d.N = int32_t(n_frag); d.Npad = pad4(n_frag); d.T = int32_t(n_theta); d.B = int32_t(n_beta);
  d.ldR = pad4(int64_t(d.T) * d.B);
  d.unif_loglik = unif_loglik;
  const size_t np_ = size_t(d.Npad);
  const size_t Rr = size_t(n_theta) * size_t(n_beta), ldr = size_t(d.ldR);
  std::vector<double> ten(ldr * size_t(n_frag), 0.0), pc(np_, 0.0);
  for (size_t n = 0; n < size_t(n_frag); n++)          // reference [t][b][n] -> device [n][t][b]
    for (size_t r = 0; r < Rr; r++) ten[n * ldr + r] = tensor[r * size_t(n_frag) + n];
  std::copy(cnt, cnt + n_frag, pc.begin());
  CU(L.d_tensor.ensure(ten.size() + size_t(kTensorSlackRows) * ldr)); CU(L.d_cnt.ensure(np_)); CU(L.d_utrs.ensure(1));
  CU(cudaMemsetAsync(L.d_tensor.p, 0, (ten.size() + size_t(kTensorSlackRows) * ldr) * (h->tensor_f32 ? 4 : 8), L.st));
  std::vector<float> tf;
  if (h->tensor_f32) {
    tf.resize(ten.size());
    for (size_t i = 0; i < ten.size(); i++) tf[i] = float(ten[i]);
    CU(cudaMemcpyAsync(L.d_tensor.p, tf.data(), 4 * tf.size(), cudaMemcpyHostToDevice, L.st));
  } else {
    CU(cudaMemcpyAsync(L.d_tensor.p, ten.data(), 8 * ten.size(), cudaMemcpyHostToDevice, L.st));
  }
  CU(cudaMemcpyAsync(L.d_cnt.p, pc.data(), 8 * np_, cudaMemcpyHostToDevice, L.st));
  CU(cudaMemcpyAsync(L.d_utrs.p, &d, sizeof(d), cudaMemcpyHostToDevice, L.st));
  std::vector<UtrDev> ud(1, d);
  std::vector<ChainDev> chains(static_cast<size_t>(n_chains));
  for (int64_t i = 0; i < n_chains; i++) {
    ChainDev& c = chains[size_t(i)];
    memset(&c, 0, sizeof(c));
    if (io[i].K < 1 || io[i].K > SCAPE_B200_KCAP) return fail(-5, "chain K out of range");
    c.utr = 0; c.K = io[i].K; c.weights_only = io[i].weights_only;
    memcpy(c.a_idx, io[i].a_idx, sizeof(c.a_idx));
    memcpy(c.b_idx, io[i].b_idx, sizeof(c.b_idx));
    memcpy(c.ws, io[i].ws, sizeof(c.ws));
    memcpy(c.k_order, io[i].k_order, SCAPE_B200_NROUND);
  }
  const bool want_trace = trace_a && trace_b && trace_ws;
  if (int rc = run_chains(h, L, chains.data(), chains.size(), ud, want_trace)) return rc;
  for (int64_t i = 0; i < n_chains; i++) {
    const ChainDev& c = chains[size_t(i)];
    memcpy(io[i].a_idx, c.a_idx, sizeof(c.a_idx));
    memcpy(io[i].b_idx, c.b_idx, sizeof(c.b_idx));
    memcpy(io[i].ws, c.ws, sizeof(c.ws));
    io[i].n_iter = c.n_iter;
    io[i].bic = c.bic;
    memcpy(io[i].lb_arr, c.lb_arr, sizeof(c.lb_arr));
  }
  if (want_trace) {
    const size_t tr = size_t(n_chains) * SCAPE_B200_NROUND * (SCAPE_B200_KCAP + 1);
    CU(cudaMemcpy(trace_a, L.d_trace_a.p, 4 * tr, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(trace_b, L.d_trace_b.p, 4 * tr, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(trace_ws, L.d_trace_ws.p, 8 * tr, cudaMemcpyDeviceToHost));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// host pre-pass entry points (CPU only)
// ------------------------------------------------------------------------------------------------
extern "C" int scape_b200_bin_reads(int64_t n_reads, const double* x, const double* l, const double* r,
                                    const double* pa, double* bx, double* bl, double* br, double* bpa, double* cnt,
                                    int32_t* read_to_bin, int64_t* n_bins) {
  UtrPrep p;
  int32_t rc = bin_reads(x, l, r, pa, n_reads, p);
  if (rc != kOk) return fail(rc, "bin_reads failed");
  *n_bins = p.n();
  std::copy(p.x.begin(), p.x.end(), bx); std::copy(p.l.begin(), p.l.end(), bl);
  std::copy(p.r.begin(), p.r.end(), br); std::copy(p.pa.begin(), p.pa.end(), bpa);
  std::copy(p.cnt.begin(), p.cnt.end(), cnt);
  std::copy(p.read_to_bin.begin(), p.read_to_bin.end(), read_to_bin);
  return 0;
}

static int prep_one(const scape_b200_params* P, int64_t n, const double* x, const double* l, const double* r,
                    const double* pa, UtrPrep& p) {
  if (int rc = check_params(*P)) return rc;
  if (bin_reads(x, l, r, pa, n, p) != kOk) return fail(p.status, "bin_reads failed");
  if (setup_model(*P, x, l, n, p) != kOk) return fail(p.status, "model set-up failed (reference would raise)");
  coverage_and_peaks(*P, p);
  return 0;
}

extern "C" int scape_b200_profile(const scape_b200_params* P, int64_t n_reads, const double* x, const double* l,
                                  const double* r, const double* pa, int64_t* L_out, int64_t* n_theta,
                                  double* theta, double* prof_y, int64_t* n_peaks, int64_t* peak_idx,
                                  double* peak_w, int64_t cap) {
  UtrPrep p;
  if (int rc = prep_one(P, n_reads, x, l, r, pa, p)) return rc;
  *L_out = p.L;
  *n_theta = p.T();
  *n_peaks = int64_t(p.peak_idx.size());
  if (!prof_y) return 0;
  if (cap < p.L + 200 || cap < p.T()) return fail(-5, "output capacity too small");
  std::copy(p.theta.begin(), p.theta.end(), theta);
  std::copy(p.prof_y.begin(), p.prof_y.end(), prof_y);
  std::copy(p.peak_idx.begin(), p.peak_idx.end(), peak_idx);
  std::copy(p.peak_w.begin(), p.peak_w.end(), peak_w);
  return 0;
}

extern "C" int scape_b200_draw_chains(const scape_b200_params* P, int64_t n_reads, const double* x, const double* l,
                                      const double* r, const double* pa, uint32_t seed, int64_t n_chains,
                                      const int32_t* ks, scape_b200_chain_io* out) {
  UtrPrep p;
  if (int rc = prep_one(P, n_reads, x, l, r, pa, p)) return rc;
  NpRandomState g(seed);
  for (int64_t i = 0; i < n_chains; i++) {
    ChainInit ci;
    if (ks[i] < 0) {           // negative K: a prune refit draw (init_ws + gen_k_arr only)
      ci.K = -ks[i];
      draw_refit(g, *P, ci);
    } else {
      int32_t rc = draw_chain(g, *P, p, ks[i], ci);
      if (rc != kOk) return fail(rc, "draw_chain failed");
    }
    memset(&out[i], 0, sizeof(out[i]));
    out[i].K = ci.K;
    memcpy(out[i].a_idx, ci.a_idx, sizeof(ci.a_idx));
    memcpy(out[i].b_idx, ci.b_idx, sizeof(ci.b_idx));
    memcpy(out[i].ws, ci.ws, sizeof(ci.ws));
    memcpy(out[i].k_order, ci.k_order, SCAPE_B200_NROUND);
  }
  return 0;
}

extern "C" int scape_b200_set_argsort_callback(scape_b200_argsort_fn fn) {
  argsort_callback() = fn;
  return 0;
}

extern "C" int scape_b200_rng_draw(uint32_t seed, int kind, int64_t arg, int64_t n, double* out) {
  NpRandomState g(seed);
  if (kind == 0) {
    for (int64_t i = 0; i < n; i++) out[i] = g.next_double();
  } else if (kind == 1) {
    for (int64_t i = 0; i < n; i++) out[i] = double(g.randint_below(arg));
  } else if (kind == 2) {
    std::vector<int64_t> p;
    g.permutation(arg, p);
    for (int64_t i = 0; i < n && i < arg; i++) out[i] = double(p[size_t(i)]);
  } else if (kind == 3) {      // the stream AFTER a permutation(arg): state continuity across the shuffle
    std::vector<int64_t> p;
    g.permutation(arg, p);
    for (int64_t i = 0; i < n; i++) out[i] = g.next_double();
  } else {
    return fail(-5, "unknown kind");
  }
  return 0;
}
