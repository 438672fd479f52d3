// Device code shared by the EM kernels (kernels.cu: bulk-synchronous step kernels; em_cluster.cu:
// the cluster-resident EM kernel): the E pass of one chain (cal_z_k, norm_z, maximize_ws, elbo --
// apa_core.py:473-573), the application of a grid arg-max found by a scan (max_alpha_beta :507-523)
// and the FP64 tensor-core MMA wrapper.  Every translation unit that includes this header owns a
// copy of the model constants (`c_mc`) and must upload it (upload_model_const_tu).
#pragma once
#include "kernels.cuh"

namespace scape {

static __constant__ ModelConst c_mc;
static inline cudaError_t upload_model_const_tu(const ModelConst& mc) { return cudaMemcpyToSymbol(c_mc, &mc, sizeof(ModelConst)); }

constexpr int GT = 256;                    // threads per CTA
constexpr int GW = GT / 32;                // warps
constexpr int SCAN_ROWS = GT;              // candidate rows per block (1 per thread)
constexpr int SCAN_GB = 32;                // chains per register sub-batch (32 FP64 accumulators per thread)
constexpr int SCAN_MAXCH = 160;            // running chains of one UTR a scan CTA can list
static_assert(SCAN_MAXCH == kScanMaxChains, "api.cu checks the chain count of a UTR against kScanMaxChains");
constexpr int SCAN_VCHUNK = 256;           // fragments of V staged per chunk
constexpr int SCAN_VPITCH = SCAN_VCHUNK + 4;  // pitch = 4 mod 16 doubles: the 8x4 B-fragment loads are bank-conflict free

__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
template <typename TT> __device__ __forceinline__ double lds_elem(uint32_t addr);
template <> __device__ __forceinline__ double lds_elem<float>(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return (double)v;
}
template <> __device__ __forceinline__ double lds_elem<double>(uint32_t addr) { return lds_f64(addr); }

// ------------------------------------------------------------------------------------------------
// E step
// ------------------------------------------------------------------------------------------------
struct ScanPartial {
  double score;
  int row;
  int pad;
};
struct EShared {
  double red[GW][SCAPE_B200_KCAP + 4];
  double tot[SCAPE_B200_KCAP + 4];
  double lwk;
  long long rk;
  int k, go, hull[2];
};

template <int NV>
__device__ __forceinline__ void block_reduce_sum(double (&val)[NV], EShared& sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    double x = val[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sh.red[warp][i] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < GW; w++) acc += sh.red[w][threadIdx.x];
    sh.tot[threadIdx.x] = acc;
  }
  __syncthreads();
}

__device__ __forceinline__ void finalize_chain(ChainDev& ch, int N) {
  const int K = ch.K;
  ch.bic = -2.0 * ch.last_a + (3 * K + 1) * log((double)N);   // cal_bic (:702-706)
  ch.state = 0;
}

// Executed by ONE thread: take over the arg-max (`row`, when the chain was waiting for one), write the
// trace, finalise a converged chain.  Returns 1 if the chain still has an E pass to do.
__device__ __forceinline__ int apply_row(ChainDev& ch, ScanDesc& sd, const UtrDev& u, bool have_row, int row,
                                         int32_t* trace_a, int32_t* trace_b, double* trace_ws) {
  int go = 1;
  if (have_row) {
    if (row < ch.row0 || row >= ch.row1) {
      // no candidate of the window won a `>` comparison: every score was NaN.  Fail the chain (the
      // host turns this into an error) instead of indexing the tensor with a row that is not one.
      ch.error = 1;
      ch.state = 0;
      ch.bic = CUDART_NAN;
      go = 0;
    } else {
      ch.a_idx[ch.cur_k] = row / u.B;
      ch.b_idx[ch.cur_k] = row % u.B;
    }
    ch.pending = 0;
    sd.pending = 0;
    if (ch.trace_off >= 0) ch.trace_pending = ch.n_iter;
  }
  if (ch.trace_off >= 0 && ch.trace_pending > 0) {
    const int64_t o = ch.trace_off + (int64_t)(ch.trace_pending - 1) * (SCAPE_B200_KCAP + 1);
    for (int j = 0; j < ch.K; j++) { trace_a[o + j] = ch.a_idx[j]; trace_b[o + j] = ch.b_idx[j]; }
    for (int j = 0; j <= ch.K; j++) trace_ws[o + j] = ch.ws[j];
    ch.trace_pending = 0;
  }
  if (ch.state == 2) { finalize_chain(ch, u.N); go = 0; }
  if (ch.n_iter >= SCAPE_B200_NROUND || ch.error) go = 0;
  return go;
}

// (0) of a step, executed by one full warp: apply the arg-max the previous scan found for this chain
// (first maximum in row order over the per-block partials), write the trace, finalise converged
// chains.  Returns 1 (in every lane) if the chain still has an E step to do.
template <int PROWS = SCAN_ROWS>
__device__ __forceinline__ int apply_pending(ChainDev& ch, ScanDesc& sd, const UtrDev& u,
                                             const ScanPartial* __restrict__ partials, int32_t* trace_a,
                                             int32_t* trace_b, double* trace_ws) {
  const int lane = threadIdx.x & 31;
  int row = 0x7fffffff;
  const bool pending = ch.pending != 0;                  // read before lane 0 clears it
  if (pending) {
    const int b0 = ch.row0 / PROWS, b1 = (ch.row1 - 1) / PROWS;     // PROWS = candidate rows per partial
    double best = -CUDART_INF;
    for (int b = b0 + lane; b <= b1; b += 32) {
      // L2 load: in the cluster kernel the partials were written by other SMs a cluster barrier ago
      const double2 raw = __ldcg(reinterpret_cast<const double2*>(partials + ch.pb_off + b));
      ScanPartial p;
      p.score = raw.x;
      p.row = (int)(__double_as_longlong(raw.y) & 0xffffffffll);
      if (p.score > best || (p.score == best && p.row < row)) { best = p.score; row = p.row; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int orow = __shfl_xor_sync(0xffffffffu, row, o);
      if (ob > best || (ob == best && orow < row)) { best = ob; row = orow; }
    }
  }
  __syncwarp();
  int go = 1;
  if (lane == 0) go = apply_row(ch, sd, u, pending, row, trace_a, trace_b, trace_ws);
  __syncwarp();
  return __shfl_sync(0xffffffffu, go, 0);
}

// After the E pass: weights (maximize_ws :498-505), ELBO (:559-561), convergence (:743), and the
// candidate window / fragment hull the scan needs.  `tot` = [cnt@Z (NK), sum Z[:,k], A term, H term].
template <int NK>
__device__ __forceinline__ void estep_epilogue(ChainDev& ch, ScanDesc& sd, const UtrDev& u, const double* tot, int k,
                                               int it, int hull_lo, int hull_hi) {
  constexpr int K = NK - 1;
  const double cap = c_mc.max_unif_ws;
  const int N = u.N, B = u.B;
  double w[NK];
  double sum = 0.0;
#pragma unroll
  for (int j = 0; j < NK; j++) sum += tot[j];
#pragma unroll
  for (int j = 0; j < NK; j++) w[j] = tot[j] / sum;
  if (w[K] > cap) {
    double rest = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) rest += w[j];
#pragma unroll
    for (int j = 0; j < K; j++) w[j] = (1 - cap) * w[j] / rest;
    w[K] = cap;
  }
#pragma unroll
  for (int j = 0; j < NK; j++) {
    ch.ws[j] = w[j];
    ch.lw[j] = (w[j] <= 0.0) ? SCAPE_SENTINEL : log(w[j]);
  }
  const double lb_new = tot[NK + 1] + tot[NK + 2];
  ch.last_a = tot[NK + 1];
  ch.lb_arr[it] = lb_new;
  ch.n_iter = it + 1;
  const double lb = ch.lb_prev;
  const bool conv = fabs(lb_new - lb) < fabs(1e-6 * lb);
  if (!conv) ch.lb_prev = lb_new;
  const bool last = conv || it == SCAPE_B200_NROUND - 1;
  ch.cur_k = k;
  if (ch.weights_only) {                                        // mstep_fixed (:552-557): no grid search
    if (last) finalize_chain(ch, N); else ch.state = 1;
  } else {
    // max_alpha_beta (:507-523): candidate window of component k
    const int lo = (k == 0) ? 0 : ch.a_idx[k - 1];
    const int hi = (k == K - 1) ? u.T - 1 : ch.a_idx[k + 1];
    ch.row0 = lo * B;
    ch.row1 = (hi + 1) * B;
    ch.hlo = hull_lo;
    ch.hhi = hull_hi;
    ch.grid_rows += (double)(ch.row1 - ch.row0);
    ch.pending = 1;
    ch.state = last ? 2 : 1;
    sd.row0 = ch.row0; sd.row1 = ch.row1; sd.hlo = hull_lo; sd.hhi = hull_hi;
    sd.v_off = ch.v_off; sd.pb_off = ch.pb_off;
    sd.pending = 1;
  }
}

// exp(x) for x <= 0, the only case the count-tempered softmax has (norm_z :491-493 subtracts the row maximum):
// branch-free -- magic-number rounding of x log2(e), two-constant Cody-Waite reduction, degree-13 Taylor
// polynomial on |r| <= ln(2)/2, two exponent scalings so that results down to the subnormals (and exactly 0
// below -745.13, the sentinel's -3.4e38 included) come out right.  Measured against long-double expl on 2e7
// arguments in [-750, 0]: max error 0.88 ulp (CUDA's exp(): 1 ulp), exp_nonpos(0) == 1.  CUDA's general exp()
// costs about twice the instructions and carries branches for its special cases (BSSY / BSYNC pairs).
// The constants sit in the uploaded model constants (c_mc.expc, filled by fill_model_const in api.cu), so
// they are constant-bank operands of the DFMAs: as literals -- or as an initialised __constant__ array,
// which the compiler folds back into literals -- every one costs two UMOVs in front of its DFMA (2,177
// UMOVs for 2,813 DFMAs in the warp E step).
#define c_expc c_mc.expc
__device__ __forceinline__ double exp_nonpos(double x) {
  x = fmax(x, c_expc[17]);
  const double t = fma(x, c_expc[0], c_expc[1]);   // c[1] = 1.5 * 2^52: rounds x log2(e) to an integer in the low word
  const int k = __double2loint(t);
  const double kd = t - c_expc[1];
  double r = fma(kd, c_expc[2], x);                // Cody-Waite: -ln2_hi, -ln2_lo
  r = fma(kd, c_expc[3], r);
  double p = c_expc[4];                             // 1/13! ... 1/2!, 1, 1
#pragma unroll
  for (int i = 5; i <= 15; i++) p = fma(p, r, c_expc[i]);
  p = fma(p, r, c_expc[16]);
  p = fma(p, r, c_expc[16]);
  const int k1 = k >> 1, k2 = k - k1;               // k in [-1077, 0]: both factors stay normal
  p *= __hiloint2double((k1 + 1023) << 20, 0);
  p *= __hiloint2double((k2 + 1023) << 20, 0);
  return p;
}

// One fragment of the E pass (shared by the warp- and the block-per-chain kernels).
// What one fragment of the E pass reads from HBM / L2: its count, its tensor entry for the refreshed
// component (a strided gather: DRAM latency) and the stale log_zmat columns.  Kept apart from the
// arithmetic so that a caller can fetch the next fragment while it computes the current one.
template <int NK, typename TT>
struct FragIn {
  double c;
  TT a;
  double lz[NK];
};
template <int NK, typename TT>
__device__ __forceinline__ void estep_load(FragIn<NK, TT>& f, int n, int k, int64_t rk, int npad, int64_t R,
                                           const TT* __restrict__ A, const double* __restrict__ cnt,
                                           const double* __restrict__ lz) {
  f.c = cnt[n];
  f.a = A[(int64_t)n * R + rk];
#pragma unroll
  for (int j = 0; j < NK; j++) f.lz[j] = (j == k) ? 0.0 : lz[(int64_t)j * npad + n];
}

template <int NK, typename TT>
__device__ __forceinline__ void estep_compute(const FragIn<NK, TT>& f, int n, int k, double lwk, bool guard, int npad,
                                              double* __restrict__ lz, double* __restrict__ V, double (&red)[NK + 3],
                                              int& h_lo, int& h_hi);

template <int NK, typename TT>
__device__ __forceinline__ void estep_fragment(int n, int k, double lwk, int64_t rk, bool guard, int npad, int64_t R,
                                               const TT* __restrict__ A, const double* __restrict__ cnt,
                                               double* __restrict__ lz, double* __restrict__ V, double (&red)[NK + 3],
                                               int& h_lo, int& h_hi) {
  FragIn<NK, TT> f;
  estep_load<NK, TT>(f, n, k, rk, npad, R, A, cnt, lz);
  estep_compute<NK, TT>(f, n, k, lwk, guard, npad, lz, V, red, h_lo, h_hi);
}

template <int NK, typename TT>
__device__ __forceinline__ void estep_compute(const FragIn<NK, TT>& f, int n, int k, double lwk, bool guard, int npad,
                                              double* __restrict__ lz, double* __restrict__ V, double (&red)[NK + 3],
                                              int& h_lo, int& h_hi) {
  const double c = f.c;
  const double fresh = lwk + (double)f.a;
  double z[NK], lzv[NK];
  double m = -CUDART_INF;
#pragma unroll
  for (int j = 0; j < NK; j++) {
    lzv[j] = (j == k) ? fresh : f.lz[j];
    m = fmax(m, lzv[j]);
  }
  lz[(int64_t)k * npad + n] = fresh;
  double s = 0.0, ex[NK];
#pragma unroll
  for (int j = 0; j < NK; j++) {
    ex[j] = (lzv[j] - m) * c;                // exponent of the count-tempered softmax (norm_z :491-493), <= 0
    z[j] = exp_nonpos(ex[j]);
    s += z[j];
  }
  const double inv_s = 1.0 / s;              // one reciprocal instead of NK divisions (<= 1 ulp per entry)
  double zk = 0.0;
#pragma unroll
  for (int j = 0; j < NK; j++) {
    z[j] = z[j] * inv_s;
    if (j == k) zk = z[j];
  }
  red[NK] += zk;                             // np.sum(Z[:, k]) before the guard
  if (guard) {
    zk += 1e-8;
#pragma unroll
    for (int j = 0; j < NK; j++)
      if (j == k) z[j] = zk;
  }
  double ps = 0.0, Aterm = 0.0;
#pragma unroll
  for (int j = 0; j < NK; j++) {
    red[j] = fma(c, z[j], red[j]);           // cnt @ Z
    if (z[j] != 0.0) Aterm += (z[j] * c) * lzv[j];
    ps += z[j];
  }
  // scipy.stats.entropy(Z[n, :]) = -sum p log p with p = Z / sum(Z).  Without the guard,
  // log p_j = ex_j - log(s) - log(ps) exactly in real arithmetic: two logs instead of NK.
  double h = 0.0;
  const double inv_ps = 1.0 / ps;
  if (!guard) {
    // ps = sum of the normalised z = 1 + d with |d| ~ 1e-16: log(ps) = d - d^2 / 2 + ..., and d^2 ~ 1e-32 is
    // far below an ulp of log(s) + d, so the second logarithm is the subtraction
    const double lnorm = log(s) + (ps - 1.0);
#pragma unroll
    for (int j = 0; j < NK; j++) {
      const double p = z[j] * inv_ps;
      if (p > 0.0) h -= p * (ex[j] - lnorm);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NK; j++) {
      const double p = z[j] * inv_ps;
      if (p > 0.0) h -= p * log(p);
    }
  }
  red[NK + 1] += Aterm;
  red[NK + 2] = fma(c, h, red[NK + 2]);
  const double vn = zk * c;
  V[n] = vn;
  if (vn != 0.0) { h_lo = min(h_lo, n); h_hi = n; }
}

// Warp-per-chain E step (small fragment counts): no block barriers, reductions by shuffles.
// WPC > 1: WPC warps share a chain (fragments dealt round-robin over 32 WPC lanes); the warps' sums
// meet once per pass in shared memory behind a named barrier and are added in warp order.  A chain's
// pass is a serial loop of N / (32 WPC) fragment passes, and a launch lasts as long as its longest chain.
struct EPairShared {
  double red[2][4][SCAPE_B200_KCAP + 4];     // [pass parity (guard retry)][warp of the chain][sum]
  int hull[2][4][2];
  int go;
};
__device__ __forceinline__ void pair_sync(int bar_id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
}

template <int NK, typename TT, bool PF, int WPC = 1>
__device__ void estep_warp_run(ChainDev& ch, ScanDesc& sd, const UtrDev& u, const TT* __restrict__ A,
                               const double* __restrict__ cnt, double* __restrict__ lz, double* __restrict__ V,
                               EPairShared* ps = nullptr, int bar_id = 0) {
  constexpr int K = NK - 1;
  constexpr int STRIDE = 32 * WPC;
  const int lane = threadIdx.x & 31;
  const int sub = WPC > 1 ? ((threadIdx.x >> 5) & (WPC - 1)) : 0;
  const int tig = sub * 32 + lane;
  const int N = u.N, npad = u.Npad, B = u.B;
  const int64_t R = u.ldR;
  const int it = ch.n_iter;
  if (it == 0) {
    for (int j = 0; j < NK; j++) {
      const double w = ch.ws[j];
      const double lw = (w <= 0.0) ? SCAPE_SENTINEL : log(w);
      if (tig == 0) ch.lw[j] = lw;
      if (j < K) {
        const int64_t rj = (int64_t)ch.a_idx[j] * B + ch.b_idx[j];
        for (int n = tig; n < N; n += STRIDE) lz[(int64_t)j * npad + n] = lw + (double)A[(int64_t)n * R + rj];
      } else {
        const double val = lw + u.unif_loglik;
        for (int n = tig; n < N; n += STRIDE) lz[(int64_t)j * npad + n] = val;
      }
    }
    if (WPC > 1) pair_sync(bar_id, STRIDE); else __syncwarp();
  }
  const int k = ch.k_order[it];
  const double lwk = ch.lw[k];
  const int64_t rk = (int64_t)ch.a_idx[k] * B + ch.b_idx[k];
  bool guard = false;
  double red[NK + 3];
  int h_lo, h_hi;
  while (true) {
#pragma unroll
    for (int j = 0; j < NK + 3; j++) red[j] = 0.0;
    h_lo = N;
    h_hi = -1;
    if (PF) {
      // software pipeline: the next fragment's loads (a strided tensor gather and the stale columns,
      // DRAM / L2 latency) are in flight while this one is computed -- ncu: 54 % of the stall
      // samples of the unpipelined loop sit on the first use of those loads
      FragIn<NK, TT> cur, nxt;
      int n = tig;
      if (n < N) estep_load<NK, TT>(cur, n, k, rk, npad, R, A, cnt, lz);
      while (n < N) {
        const int nn = n + STRIDE;
        if (nn < N) estep_load<NK, TT>(nxt, nn, k, rk, npad, R, A, cnt, lz);
        estep_compute<NK, TT>(cur, n, k, lwk, guard, npad, lz, V, red, h_lo, h_hi);
        cur = nxt;
        n = nn;
      }
    } else {
      for (int n = tig; n < N; n += STRIDE) estep_fragment<NK, TT>(n, k, lwk, rk, guard, npad, R, A, cnt, lz, V, red, h_lo, h_hi);
    }
#pragma unroll
    for (int j = 0; j < NK + 3; j++) {
      double x = red[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      red[j] = x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      h_lo = min(h_lo, __shfl_xor_sync(0xffffffffu, h_lo, o));
      h_hi = max(h_hi, __shfl_xor_sync(0xffffffffu, h_hi, o));
    }
    if (WPC > 1) {
      const int par = guard ? 1 : 0;         // the retry pass uses the other half: no barrier between the two
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NK + 3; j++) ps->red[par][sub][j] = red[j];
        ps->hull[par][sub][0] = h_lo;
        ps->hull[par][sub][1] = h_hi;
      }
      pair_sync(bar_id, STRIDE);
#pragma unroll
      for (int j = 0; j < NK + 3; j++) {
        double x = ps->red[par][0][j];
#pragma unroll
        for (int w = 1; w < WPC; w++) x += ps->red[par][w][j];
        red[j] = x;
      }
#pragma unroll
      for (int w = 0; w < WPC; w++) {
        h_lo = min(h_lo, ps->hull[par][w][0]);
        h_hi = max(h_hi, ps->hull[par][w][1]);
      }
    }
    if (!guard && red[NK] < 1e-8) {          // mstep guard (:526-529), uniform across the chain's warps
      guard = true;
      continue;
    }
    break;
  }
  __syncwarp();
  if (tig == 0) estep_epilogue<NK>(ch, sd, u, red, k, it, h_lo, h_hi);
}

// ---- G warps per chain: named-barrier groups (used by em_estep_group_kernel and the cluster kernel) ----
struct EGroupShared {
  double red[GW][SCAPE_B200_KCAP + 4];
  double tot[SCAPE_B200_KCAP + 4];
  double lwk;
  long long rk;
  int k, go, hull[2];
};

__device__ __forceinline__ void group_sync(int gid, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(nthreads) : "memory");
}

template <int NV>
__device__ __forceinline__ void group_reduce_sum(double (&val)[NV], EGroupShared& sh, int G, int gid, int tig) {
  const int lane = tig & 31, wig = tig >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    double x = val[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sh.red[wig][i] = x;
  }
  group_sync(gid, 32 * G);
  if (tig < NV) {
    double acc = 0.0;
    for (int w = 0; w < G; w++) acc += sh.red[w][tig];
    sh.tot[tig] = acc;
  }
  group_sync(gid, 32 * G);
}

template <int NK, typename TT>
__device__ __noinline__ void estep_group_run(EGroupShared& sh, int G, int gid, int tig, ChainDev& ch, ScanDesc& sd, const UtrDev& u,
                                const TT* __restrict__ A, const double* __restrict__ cnt, double* __restrict__ lz,
                                double* __restrict__ V) {
  constexpr int K = NK - 1;
  const int gthreads = 32 * G;
  const int N = u.N, npad = u.Npad, B = u.B;
  const int64_t R = u.ldR;
  const int it = ch.n_iter;
  if (it == 0) {
    // initial log_zmat: all K+1 columns (em_algo :722-724)
    for (int j = 0; j < NK; j++) {
      const double w = ch.ws[j];
      const double lw = (w <= 0.0) ? SCAPE_SENTINEL : log(w);
      if (tig == 0) ch.lw[j] = lw;
      if (j < K) {
        const int64_t rj = (int64_t)ch.a_idx[j] * B + ch.b_idx[j];
        for (int n = tig; n < N; n += gthreads) lz[(int64_t)j * npad + n] = lw + (double)A[(int64_t)n * R + rj];
      } else {
        const double val = lw + u.unif_loglik;
        for (int n = tig; n < N; n += gthreads) lz[(int64_t)j * npad + n] = val;
      }
    }
  }
  if (tig == 0) {
    const int k = ch.k_order[it];
    sh.k = k;
    sh.lwk = ch.lw[k];
    sh.rk = (long long)ch.a_idx[k] * B + ch.b_idx[k];
  }
  group_sync(gid, gthreads);
  const int k = sh.k;
  const double lwk = sh.lwk;
  const int64_t rk = sh.rk;
  bool guard = false;
  double red[NK + 3];
  while (true) {
#pragma unroll
    for (int j = 0; j < NK + 3; j++) red[j] = 0.0;
    int h_lo = N, h_hi = -1;
    if (tig == 0) { sh.hull[0] = N; sh.hull[1] = -1; }
    {
      // software pipeline: the next fragment's loads are in flight while this one is computed
      FragIn<NK, TT> cur, nxt;
      int n = tig;
      if (n < N) estep_load<NK, TT>(cur, n, k, rk, npad, R, A, cnt, lz);
      while (n < N) {
        const int nn = n + gthreads;
        if (nn < N) estep_load<NK, TT>(nxt, nn, k, rk, npad, R, A, cnt, lz);
        estep_compute<NK, TT>(cur, n, k, lwk, guard, npad, lz, V, red, h_lo, h_hi);
        cur = nxt;
        n = nn;
      }
    }
    group_sync(gid, gthreads);
    if (h_hi >= 0) { atomicMin(&sh.hull[0], h_lo); atomicMax(&sh.hull[1], h_hi); }
    group_reduce_sum<NK + 3>(red, sh, G, gid, tig);
    if (!guard && sh.tot[NK] < 1e-8) {       // mstep guard (:526-529); uniform across the group
      guard = true;
      group_sync(gid, gthreads);
      continue;
    }
    break;
  }
  if (tig == 0) estep_epilogue<NK>(ch, sd, u, sh.tot, k, it, sh.hull[0], sh.hull[1]);
}

// A chain's record (944 bytes over 8 cache lines) is read field by field through several dependent
// steps of an E pass (state -> window -> k_order -> log w, alpha, beta ...): staged once into shared
// memory by a coalesced copy, worked on there and copied back, the pass pays one global round trip
// for it instead of one per step.
static_assert(sizeof(ChainDev) % 8 == 0, "ChainDev is copied in 8-byte words");
__device__ __forceinline__ void copy_chain(ChainDev* dst, const ChainDev* src, int t, int nt) {
  const double* s = reinterpret_cast<const double*>(src);
  double* d = reinterpret_cast<double*>(dst);
  for (int i = t; i < (int)(sizeof(ChainDev) / 8); i += nt) d[i] = s[i];
}

// D(8x8) += A(8x4) * B(4x8), FP64 tensor-core MMA.  Fragment layout (PTX ISA, mma.m8n8k4 .f64):
//   A: a0 = A[lane>>2][lane&3]     B: b0 = B[lane&3][lane>>2]     D: d{0,1} = D[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}


// ---- grid arg-max as 32-row FP64 MMA tiles (used by em_scan_kernel's tile path and the cluster kernel) ----
// The chains of a pass: <= 32 chains of one UTR whose V rows (fragments [NA, NB), pitch P doubles, plus
// one row of zeros at index `cnt`) sit in shared memory.
struct PassCtx {
  int row0[kClusterPassMax], row1[kClusterPassMax], hlo[kClusterPassMax], hhi[kClusterPassMax];
  long long voff[kClusterPassMax], pboff[kClusterPassMax];
  int NA, NB, P;
  unsigned char wlist[GW][kClusterPassMax];   // per warp: the slots of the chains that cover the warp's current tile, ascending
};

// scores of one 32-row tile against the nt <= 8 NG chains of the pass that cover it (bit `s` of `mask`:
// chain slot s of the pass), fragments [h0, h1), h0 % 4 == 0.
// 4 consecutive candidate rows of one fragment as one 128-bit load (float storage) or two (double):
// row groups start at multiples of 4 and the tensor pitch is a multiple of 4 elements.
template <typename TT> struct ARow4;
template <> struct ARow4<float> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ double get(int i) const { return (double)(i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w); }
};
template <> struct ARow4<double> {
  double2 a, b;
  __device__ __forceinline__ void load(const double* p) {
    a = __ldg(reinterpret_cast<const double2*>(p));
    b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  }
  __device__ __forceinline__ double get(int i) const { return i == 0 ? a.x : i == 1 ? a.y : i == 2 ? b.x : b.y; }
};

// scores of one 32-row tile against the nt <= 8 NG chains of the pass that cover it (bit `s` of `mask`:
// chain slot s of the pass), fragments [h0, h1), h0 % 4 == 0.
//
// MMA rows are dealt to candidate rows as row(mi, g) = base + 4 g + mi: the four A operands a lane needs
// for one k-step (mi = 0..3, fragment k0 + q) are then 16 contiguous bytes of the [fragment][row]
// tensor -- ONE 128-bit load per lane and k-step instead of four 32-bit ones, and the 8 lanes of a
// quad column read one whole 128-byte line.  The loads are the scan's bottleneck (the tensor streams
// from HBM / L2 with no reuse inside a warp): fewer, wider loads leave room for a deeper ring.
template <int NG, typename TT, bool TO_SMEM>
__device__ __forceinline__ void scan_tile_mma(const PassCtx& sh, const UtrDev& u, const TT* __restrict__ A,
                                              const double* Vs, ScanPartial* partials, int t, int cntp,
                                              unsigned mask, int nt, int h0, int h1, double* wbest, int* wrow) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;          // MMA group id / thread-in-group
  const int64_t R = u.ldR;
  const int Rv = u.T * u.B;
  const int base = t * kClusterTileRows;
  const int tile_end = min(base + kClusterTileRows, Rv);
  const int NA = sh.NA, P = sh.P;
  const unsigned char* wl = sh.wlist[(threadIdx.x >> 5) & (GW - 1)];
  // B operand of this lane: V[chain slot of column 8 ni + g][fragment k0 + q]; columns past nt read the zero row
  const uint32_t vs_base = (uint32_t)__cvta_generic_to_shared(Vs);
  uint32_t vb[NG];
#pragma unroll
  for (int ni = 0; ni < NG; ni++) {
    const int idx = 8 * ni + g;
    const int slot = idx < nt ? (int)wl[idx] : cntp;
    vb[ni] = vs_base + (uint32_t)(slot * P + (h0 - NA) + q) * 8u;
  }
  double acc[4][NG][2];
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < NG; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  const int len4 = (h1 - h0 + 3) & ~3;
  // Register ring, refilled unconditionally with pointer increments (see scan_subbatch in kernels.cu):
  // reads run up to 4 PFD + 3 fragments past h1, into the next UTR's tensor or the zeroed slack rows
  // (all finite), against V = 0.  Row groups past the grid end are clamped to the last group of the
  // pitch (masked below).
  constexpr int PFD = (NG == 1 ? 14 : NG == 2 ? 12 : NG == 3 ? 8 : 6) / (sizeof(TT) == 8 ? 2 : 1);
  static_assert(4 * PFD + 3 < kTensorSlackRows, "the ring may read at most kTensorSlackRows fragments past a UTR");
  const int64_t kstep = 4 * R;
  const TT* pp = A + min((int64_t)base + 4 * g, R - 4) + (int64_t)(h0 + q) * R;
  ARow4<TT> pre[PFD];
#pragma unroll
  for (int p = 0; p < PFD; p++) {
    pre[p].load(pp);
    pp += kstep;
  }
  int kk = 0;
  for (; kk + 4 * PFD <= len4; kk += 4 * PFD) {
#pragma unroll
    for (int p = 0; p < PFD; p++) {
      double a[4];
#pragma unroll
      for (int mi = 0; mi < 4; mi++) a[mi] = pre[p].get(mi);
      pre[p].load(pp);
      pp += kstep;
      double b[NG];
#pragma unroll
      for (int ni = 0; ni < NG; ni++) b[ni] = lds_f64(vb[ni] + (uint32_t)(kk + 4 * p) * 8u);
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < NG; ni++) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
    }
  }
#pragma unroll
  for (int p = 0; p < PFD; p++) {                          // remainder: the ring already holds it
    if (kk + 4 * p < len4) {
      double b[NG];
#pragma unroll
      for (int ni = 0; ni < NG; ni++) b[ni] = lds_f64(vb[ni] + (uint32_t)(kk + 4 * p) * 8u);
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < NG; ni++) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], pre[p].get(mi), b[ni]);
    }
  }
  // first maximum of the tile per chain: larger score wins, ties go to the smaller row.
  // acc[mi][ni][i] = score[row = base + 4 g + mi][column 8 ni + 2 q + i]
#pragma unroll
  for (int ni = 0; ni < NG; ni++) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int c = 8 * ni + 2 * q + i;
      const bool live = c < nt;
      const int slot = live ? (int)wl[c] : 0;
      const int w0 = live ? sh.row0[slot] : 0, w1 = live ? min(sh.row1[slot], tile_end) : 0;
      double b = -CUDART_INF;
      int r = 0x7fffffff;
#pragma unroll
      for (int mi = 0; mi < 4; mi++) {
        const int row = base + 4 * g + mi;
        if (row >= w0 && row < w1 && acc[mi][ni][i] > b) { b = acc[mi][ni][i]; r = row; }   // rows ascend with mi
      }
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {         // lanes with the same q hold the same column
        const double ob = __shfl_xor_sync(0xffffffffu, b, o);
        const int orow = __shfl_xor_sync(0xffffffffu, r, o);
        if (ob > b || (ob == b && orow < r)) { b = ob; r = orow; }
      }
      if (live && g == 0) {
        if (TO_SMEM) {                                   // the caller combines the tiles of its row block
          wbest[slot] = b;
          wrow[slot] = r;
        } else {
          ScanPartial p;
          p.score = b; p.row = r; p.pad = 0;
          partials[sh.pboff[slot] + t] = p;
        }
      }
    }
  }
}

template <typename TT, bool TO_SMEM>
__device__ __forceinline__ void scan_tile(PassCtx& sh, const UtrDev& u, const TT* __restrict__ A,
                                          const double* Vs, ScanPartial* partials, int t, int cntp,
                                          double* scan_elems, double* wbest, int* wrow) {
  const int lane = threadIdx.x & 31;
  const int base = t * kClusterTileRows;
  const int tile_end = min(base + kClusterTileRows, u.T * u.B);
  const bool cover = lane < cntp && sh.row0[lane] < tile_end && sh.row1[lane] > base;
  const unsigned mask = __ballot_sync(0xffffffffu, cover);
  const int nt = __popc(mask);
  if (nt == 0) return;
  // compact list of the covering chains' slots (MMA column -> slot), per warp
  __syncwarp();                                   // the previous tile's readers are done
  if (cover) sh.wlist[(threadIdx.x >> 5) & (GW - 1)][__popc(mask & ((1u << lane) - 1u))] = (unsigned char)lane;
  __syncwarp();
  int h0 = 1 << 30, h1 = 0;
  if (cover && sh.hhi[lane] >= 0) { h0 = sh.hlo[lane]; h1 = sh.hhi[lane] + 1; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    h0 = min(h0, __shfl_xor_sync(0xffffffffu, h0, o));
    h1 = max(h1, __shfl_xor_sync(0xffffffffu, h1, o));
  }
  if (h1 <= h0) { h0 = 0; h1 = 0; }              // every v is zero: all scores 0, the first row of each window wins
  h0 &= ~3;
  if (scan_elems && lane == 0) atomicAdd(scan_elems, (double)(tile_end - base) * (double)(h1 - h0));
  const int NG = (nt + 7) >> 3;
  if (NG == 1) scan_tile_mma<1, TT, TO_SMEM>(sh, u, A, Vs, partials, t, cntp, mask, nt, h0, h1, wbest, wrow);
  else if (NG == 2) scan_tile_mma<2, TT, TO_SMEM>(sh, u, A, Vs, partials, t, cntp, mask, nt, h0, h1, wbest, wrow);
  else if (NG == 3) scan_tile_mma<3, TT, TO_SMEM>(sh, u, A, Vs, partials, t, cntp, mask, nt, h0, h1, wbest, wrow);
  else scan_tile_mma<4, TT, TO_SMEM>(sh, u, A, Vs, partials, t, cntp, mask, nt, h0, h1, wbest, wrow);
}


}  // namespace scape
