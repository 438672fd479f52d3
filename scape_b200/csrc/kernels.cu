// See kernels.cuh for the kernel list, the reference lines each kernel follows and the HBM layout.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "em_device.cuh"

namespace scape {

static bool dbg_env() {
  static const bool on = scape_env_on("SCAPE_B200_DBG");
  return on;
}

cudaError_t upload_model_const_cluster(const ModelConst& mc);   // em_cluster.cu's copy
cudaError_t upload_model_const_tail(const ModelConst& mc);      // em_tail.cu's copy
cudaError_t upload_model_const(const ModelConst& mc) {
  cudaError_t e = upload_model_const_tu(mc);
  if (e == cudaSuccess) e = upload_model_const_cluster(mc);
  return e == cudaSuccess ? upload_model_const_tail(mc) : e;
}

// ------------------------------------------------------------------------------------------------
// scalar helpers (taichi_core.py:24-97)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double logpdf_normal(double x, double mu, double sigma) {
  double d = (x - mu) / sigma;
  return -0.5 * (d * d) - log(sigma) - 0.5 * log(2 * SCAPE_PI);
}

__device__ __forceinline__ double pdf_normal(double x, double mu, double sigma) {
  double d = (x - mu) / sigma;
  return exp(-0.5 * (d * d)) / sqrt(2 * SCAPE_PI) / sigma;
}

// ------------------------------------------------------------------------------------------------
// K2: theta table.  One thread per (theta row, fragment); n contiguous -> coalesced stores.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) table_kernel(const UtrDev* __restrict__ utrs, const RowRef* __restrict__ rows,
                                                    const double* __restrict__ fx, const double* __restrict__ fl,
                                                    const double* __restrict__ fr, const double* __restrict__ fpa,
                                                    const double* __restrict__ theta, double* __restrict__ table) {
  const RowRef rr = rows[blockIdx.x];
  const UtrDev u = utrs[rr.utr];
  const int n = blockIdx.y * blockDim.x + threadIdx.x;
  if (n >= u.Npad) return;
  double out = 0.0;  // padding columns stay finite; their responsibilities are never read
  if (n < u.N) {
    const double th = theta[u.theta_off + rr.t];
    const double x = fx[u.frag_off + n], l = fl[u.frag_off + n];
    const double r = fr[u.frag_off + n], pa = fpa[u.frag_off + n];
    const double span = th - x;
    const bool fits = (l <= span);
    const double mu_f = c_mc.mu_f, sigma_f = c_mc.sigma_f;
    if (!isnan(pa)) {
      // loglik_xlr_t_pa_kernel (taichi_core.py:101-107)
      const double ll = fits ? -log(span) : SCAPE_SENTINEL;
      out = ll + logpdf_normal(pa - th, 0.0, sigma_f);
    } else if (!isnan(r)) {
      // loglik_xlr_t_r_known_kernel (taichi_core.py:111-132): LSE over s >= r, minus log of the kept pmf mass
      const double ll = fits ? -log(span) : SCAPE_SENTINEL;
      double mass = 0.0, mx = 0.0;
      bool first = true;
      for (int j = 0; j < c_mc.n_s; j++) {
        const double s = c_mc.s_dis[j];
        double v = SCAPE_SENTINEL;
        if (!(s < r)) {
          mass += c_mc.pmf_s[j];
          const double lr = (r <= s) ? -log(s) : SCAPE_SENTINEL;
          v = lr + logpdf_normal(x, th + s - mu_f, sigma_f) + ll + c_mc.logpmf_s[j];
        }
        if (first || v > mx) mx = v;
        first = false;
      }
      double acc = 0.0;
      for (int j = 0; j < c_mc.n_s; j++) {
        const double s = c_mc.s_dis[j];
        double v = SCAPE_SENTINEL;
        if (!(s < r)) {
          const double lr = (r <= s) ? -log(s) : SCAPE_SENTINEL;
          v = lr + logpdf_normal(x, th + s - mu_f, sigma_f) + ll + c_mc.logpmf_s[j];
        }
        acc += exp(v - mx);
      }
      out = (log(acc) + mx) - log(mass);
    } else {
      // loglik_xlr_t_r_unknown_kernel (taichi_core.py:141-157).  A read that does not fit (l > theta - x)
      // has lik_l_xt = 0: every term of the sum is exactly +0, the sum is 0 < 1e-300 and the entry is
      // the sentinel -- 71-74 % of all entries, decided here without the 13 exp (fragments are sorted
      // by x, so whole warps take the same side).
      if (!fits) {
        out = SCAPE_SENTINEL;
      } else {
        const double inv_span = 1.0 / span;
        double acc = 0.0;
        for (int j = 0; j < c_mc.n_s; j++) {
          const double s = c_mc.s_dis[j];
          acc += 1.0 / s * pdf_normal(x, th + s - mu_f, sigma_f) * inv_span * c_mc.pmf_s[j];
        }
        if (acc < 1e-300) acc = 0.0;
        out = (acc <= 0.0) ? SCAPE_SENTINEL : log(acc);
      }
    }
  }
  table[u.table_off + (int64_t)rr.t * u.Npad + n] = out;
}

void launch_table(const UtrDev* utrs, const RowRef* rows, int64_t n_rows, int max_n, const double* fx,
                  const double* fl, const double* fr, const double* fpa, const double* theta, double* table,
                  cudaStream_t st) {
  if (n_rows <= 0) return;
  dim3 grid((unsigned)n_rows, (unsigned)((max_n + 255) / 256));
  table_kernel<<<grid, 256, 0, st>>>(utrs, rows, fx, fl, fr, fpa, theta, table);
}

// ------------------------------------------------------------------------------------------------
// K3: marginal tensor.  One CTA per (alpha row, 256-fragment tile); the 13 beta windows and their
// normalised weights are built once per CTA in shared memory, then every thread owns one fragment.
// The reference does a two-pass log-sum-exp per (alpha, beta) (taichi_core.py:41-54, 172-179): 307
// exp per (fragment, alpha).  Here the exp of each table entry is taken ONCE relative to the maximum
// over the widest window and reused by all betas (43 exp + 13*43 FMA); a beta whose own window would
// underflow falls back to the reference's exact form.
// ------------------------------------------------------------------------------------------------
template <typename TT>
__global__ void __launch_bounds__(256) tensor_kernel(const UtrDev* __restrict__ utrs, const RowRef* __restrict__ rows,
                                                     int max_win, const double* __restrict__ theta,
                                                     const double* __restrict__ table, TT* __restrict__ tensor) {
  extern __shared__ double sm[];
  const RowRef rr = rows[blockIdx.x];
  const UtrDev u = utrs[rr.utr];
  const int B = u.B;
  if ((int)(blockIdx.y * blockDim.x) >= u.Npad) return;
  double* s_logp = sm;                       // [B][max_win]
  double* s_p = sm + (size_t)B * max_win;    // [B][max_win]
  double* s_lps = s_p + (size_t)B * max_win; // [B]
  int* s_lo = (int*)(s_lps + B);             // [B]
  int* s_w = s_lo + B;                       // [B]
  const double* th = theta + u.theta_off;
  const double alpha = th[rr.t];
  const int tid = threadIdx.x;
  if (tid < B) {
    // np.searchsorted(all_theta, alpha - 3 beta, 'left') / (alpha + 3 beta, 'right') - 1  (taichi_core.py:221-222)
    const double beta = c_mc.betas[tid];
    const double lo_v = alpha - 3 * beta, hi_v = alpha + 3 * beta;
    int a = 0, b = u.T;
    while (a < b) { int m = (a + b) >> 1; if (th[m] < lo_v) a = m + 1; else b = m; }
    const int lo = a;
    a = 0; b = u.T;
    while (a < b) { int m = (a + b) >> 1; if (th[m] <= hi_v) a = m + 1; else b = m; }
    s_lo[tid] = lo;
    s_w[tid] = a - lo;  // hi - lo + 1
  }
  __syncthreads();
  for (int e = tid; e < B * max_win; e += blockDim.x) {
    const int j = e / max_win, d = e % max_win;
    if (d < s_w[j]) {
      const double lp = logpdf_normal(th[s_lo[j] + d], alpha, c_mc.betas[j]);
      s_logp[e] = lp;
      s_p[e] = exp(lp);
    }
  }
  __syncthreads();
  if (tid < B) {
    // call_logp_theta_sum_kernel (taichi_core.py:160-169), summed in theta order like the CPU twin
    double acc = 0.0;
    for (int d = 0; d < s_w[tid]; d++) acc += s_p[tid * max_win + d];
    s_lps[tid] = log(acc);
  }
  __syncthreads();
  // normalised window weights g[j][d] = N(theta_d; alpha, beta_j) / sum, laid out on the UNION window
  // [lo_all, lo_all + w_all) and zero outside beta_j's own window (s_p is reused for them)
  int lo_all = s_lo[0], hi_all = s_lo[0] + s_w[0];
  for (int j = 1; j < B; j++) { lo_all = min(lo_all, s_lo[j]); hi_all = max(hi_all, s_lo[j] + s_w[j]); }
  const int w_all = hi_all - lo_all;
  __syncthreads();
  for (int e = tid; e < B * max_win; e += blockDim.x) {
    const int j = e / max_win, d = e % max_win;
    const int dj = d - (s_lo[j] - lo_all);                 // index inside beta_j's own window
    s_p[e] = (d < w_all && dj >= 0 && dj < s_w[j]) ? exp(s_logp[j * max_win + dj] - s_lps[j]) : 0.0;
  }
  __syncthreads();
  const int n = blockIdx.y * blockDim.x + tid;
  if (n >= u.N) return;   // the tensor has no padding fragments
  const double* tab = table + u.table_off + n;
  // tensor layout [n][R], R = T*B candidate rows (alpha-major, beta-minor) contiguous per fragment
  TT* out = tensor + u.tensor_off + (int64_t)n * u.ldR + (int64_t)rr.t * B;
  // The pitch ldR rounds T*B up to 4: the pad columns must hold finite values, because the scan's
  // k-loop runs a few fragments past a UTR's last one (multiplied by V = 0) and those reads can land
  // on the pad columns of the next UTR's region.  The last alpha row always takes this kernel.
  if (rr.t == u.T - 1)
    for (int64_t c = (int64_t)u.T * B; c < u.ldR; c++) tensor[u.tensor_off + (int64_t)n * u.ldR + c] = TT(0);
  const int64_t ld = u.Npad;
  const double* col_all = tab + (int64_t)lo_all * ld;
  // One exp per theta of the union window, shared by all betas:
  //   sum_d exp(table_d + logp_jd - lps_j) = exp(M) * sum_d exp(table_d - M) * g_jd,   M = max_d table_d.
  // Identical to the reference's per-beta two-pass log-sum-exp up to rounding; when a beta's own
  // window lies so far below M that the shared sum underflows, that beta falls back to the exact
  // two-pass form (this also reproduces the all-sentinel case exactly).
  double M = col_all[0];
  for (int d = 1; d < w_all; d++) M = fmax(M, col_all[(int64_t)d * ld]);
  if (M < -1e30) {
    // the read is incompatible with every theta near this alpha (71-74 % of all entries): each beta's
    // log-sum-exp is log(w) + sentinel, which IS the sentinel in FP64 (ulp(3.4e38) = 3.8e22)
    for (int j = 0; j < B; j++) out[j] = TT(SCAPE_SENTINEL);
    return;
  }
  constexpr int JB = 16;                                   // betas handled per pass over the window
  double acc[JB];
  for (int j0 = 0; j0 < B; j0 += JB) {
    const int nj = min(JB, B - j0);
#pragma unroll
    for (int j = 0; j < JB; j++) acc[j] = 0.0;
    for (int d = 0; d < w_all; d++) {
      const double a = col_all[(int64_t)d * ld] - M;
      if (a > -746.0) {                                    // below that exp() is exactly 0 (sentinel: a ~ -3.4e38)
        const double e = exp(a);
#pragma unroll
        for (int j = 0; j < JB; j++)
          if (j < nj) acc[j] = fma(e, s_p[(j0 + j) * max_win + d], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < JB; j++) {
      if (j < nj) {
        double res;
        if (acc[j] > 1e-290 && M > -1e30) {
          res = log(acc[j]) + M;
        } else {
          // exact two-pass log-sum-exp over beta_j's own window (taichi_core.py:41-54, 172-179)
          const int jj = j0 + j, lo = s_lo[jj], w = s_w[jj];
          const double lps = s_lps[jj];
          const double* lp = s_logp + jj * max_win;
          const double* col = tab + (int64_t)lo * ld;
          double m = (col[0] + lp[0]) - lps;
          for (int d = 1; d < w; d++) m = fmax(m, (col[(int64_t)d * ld] + lp[d]) - lps);
          if (m < -1e30) {
            res = log((double)w) + m;   // every term is the sentinel: exp(0) each, log(w) + sentinel == sentinel
          } else {
            double sum = 0.0;
            for (int d = 0; d < w; d++) {
              const double a = ((col[(int64_t)d * ld] + lp[d]) - lps) - m;
              if (a > -746.0) sum += exp(a);
            }
            res = log(sum) + m;
          }
        }
        out[j0 + j] = TT(res);   // float storage keeps the sentinel exactly (it IS float's lowest)
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K3 fast path: interior alpha rows of a regular theta grid (every normal `run()`: theta_step-spaced
// grid, the 13 default betas, windows 3..43 wide).  There the normalised window weights
// g[j][d] = N(theta_d; alpha, beta_j) / sum depend on (j, d) only, not on alpha, so they live in
// constant memory and enter the FMAs as constant operands (no shared-memory broadcast), and a thread
// (= one fragment) takes exp() of each table entry ONCE for a tile of TI consecutive alpha rows:
// (TI + 42) / TI = 6.25 exp per (fragment, alpha) instead of 307 in the reference.
// Rows whose widest window is clipped by the grid ends, irregular (fixed-mode) grids and any other
// parameter set go through tensor_kernel above.
// ------------------------------------------------------------------------------------------------
constexpr int TF_B = 13, TF_W = 43, TF_HALF = 21, TF_TI = 8, TF_COLS = TF_TI + TF_W - 1;
constexpr int TF_THREADS = 64;   // small CTAs: fragments are sorted by read start, so whole warps exit on the
                                 // all-sentinel fast path; small CTAs free their shared memory sooner
__constant__ double c_tf_g[TF_B * TF_W];     // weights on the widest window, 0 outside beta_j's own window
__constant__ double c_tf_lp[TF_B * TF_W];    // log pdf (for the exact fallback)
__constant__ double c_tf_lps[TF_B];
__constant__ int c_tf_hw[TF_B];              // half width of beta_j's window in grid points
__constant__ double c_tf_pre[TF_B * (TF_W + 1)];   // prefix sums of exp(lp[j][d]) over d: clipped-window normalisers of the edge rows

cudaError_t upload_tensor_fast_tables(const double* g, const double* lp, const double* lps, const int* hw) {
  cudaError_t e = cudaMemcpyToSymbol(c_tf_g, g, sizeof(double) * TF_B * TF_W);
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_tf_lp, lp, sizeof(double) * TF_B * TF_W);
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_tf_lps, lps, sizeof(double) * TF_B);
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_tf_hw, hw, sizeof(int) * TF_B);
  if (e == cudaSuccess) {
    // prefix sums of exp(lp) in theta order (0 outside beta_j's own window)
    double pre[TF_B * (TF_W + 1)];
    for (int j = 0; j < TF_B; j++) {
      pre[j * (TF_W + 1)] = 0.0;
      for (int d = 0; d < TF_W; d++) {
        const bool in = d >= TF_HALF - hw[j] && d <= TF_HALF + hw[j];
        pre[j * (TF_W + 1) + d + 1] = pre[j * (TF_W + 1) + d] + (in ? std::exp(lp[j * TF_W + d]) : 0.0);
      }
    }
    e = cudaMemcpyToSymbol(c_tf_pre, pre, sizeof(pre));
  }
  return e;
}


// Alpha rows whose widest window is clipped by the grid ends (the first and last 21 rows of a UTR) take
// the same kernel: grid points outside [0, T) contribute nothing (e = 0, dead), and the weights of a
// clipped window are the full-window weights times exp(lps_full - lps_clip), lps_clip = log of the sum
// of exp(lp) over the points that exist (call_logp_theta_sum_kernel sums exactly those,
// taichi_core.py:160-169, 221-222), so
//     res = log(sum_d e_d g_full[j][d]) + M + (lps_full[j] - lps_clip[j][row]).
// lps_clip comes from a prefix-sum table (one log per (row, beta) of the tile, not per fragment).
template <typename TT>
__global__ void __launch_bounds__(TF_THREADS) tensor_interior_kernel(const UtrDev* __restrict__ utrs,
                                                              const TileRef* __restrict__ tiles,
                                                              const double* __restrict__ table,
                                                              TT* __restrict__ tensor) {
  extern __shared__ double sm_e[];             // [TF_COLS][TF_THREADS]: this thread's exp(table - M) column
  __shared__ double s_dl[TF_TI][TF_B];         // edge tiles: lps_full - lps_clip per (row of the tile, beta)
  __shared__ int s_lo[TF_TI][TF_B], s_hi[TF_TI][TF_B];   // edge tiles: first / last existing point of beta_j's window (index d)
  const TileRef tr = tiles[blockIdx.x];
  const UtrDev u = utrs[tr.utr];
  const int tid = threadIdx.x;
  const int n = blockIdx.y * blockDim.x + tid;
  if ((int)(blockIdx.y * blockDim.x) >= u.N) return;   // CTA-uniform
  const bool edge = tr.i0 < TF_HALF || tr.i0 + tr.cnt - 1 > u.T - 1 - TF_HALF;
  if (edge) {
    for (int e = tid; e < tr.cnt * TF_B; e += TF_THREADS) {
      const int ii = e / TF_B, j = e - ii * TF_B;
      const int hw = c_tf_hw[j], row = tr.i0 + ii;
      const int lo = TF_HALF - min(hw, row), hi = TF_HALF + min(hw, u.T - 1 - row);   // existing points of the window, as d
      s_lo[ii][j] = lo;
      s_hi[ii][j] = hi;
      s_dl[ii][j] = c_tf_lps[j] - log(c_tf_pre[j * (TF_W + 1) + hi + 1] - c_tf_pre[j * (TF_W + 1) + lo]);
    }
    __syncthreads();
  }
  if (n >= u.N) return;                        // no barrier below: every thread only touches its own column
  const int64_t ld = u.Npad;
  const int c_first = tr.i0 - TF_HALF;         // grid index of column 0 (negative at the left edge)
  const double* tab = table + u.table_off + n;
  const int cols = tr.cnt + TF_W - 1;
  double M = -CUDART_INF;
  uint64_t live = 0ull;                        // bit c: theta column c exists and is compatible with this fragment
  for (int c = 0; c < cols; c++) {
    const int gc = c_first + c;
    if (gc < 0 || gc >= u.T) continue;
    const double v = tab[(int64_t)gc * ld];
    M = fmax(M, v);
    live |= (v > -1e30 ? 1ull : 0ull) << c;
  }
  TT* out = tensor + u.tensor_off + (int64_t)n * u.ldR + (int64_t)tr.i0 * TF_B;
  if (tr.i0 + tr.cnt == u.T)                   // the tile with the last alpha row also writes the pitch padding (finite)
    for (int64_t c = (int64_t)u.T * TF_B; c < u.ldR; c++) tensor[u.tensor_off + (int64_t)n * u.ldR + c] = TT(0);
  if (M < -1e30) {                             // incompatible with every theta of the tile: all sentinel
    for (int e = 0; e < tr.cnt * TF_B; e++) out[e] = TT(SCAPE_SENTINEL);
    return;
  }
  for (int c = 0; c < cols; c++) {
    const int gc = c_first + c;
    double ev = 0.0;
    if (gc >= 0 && gc < u.T) {
      const double a = tab[(int64_t)gc * ld] - M;
      ev = (a > -746.0) ? exp(a) : 0.0;
    }
    sm_e[c * TF_THREADS + tid] = ev;
  }
  for (int ii = 0; ii < tr.cnt; ii++) {
    double acc[TF_B];
#pragma unroll
    for (int j = 0; j < TF_B; j++) acc[j] = 0.0;
#pragma unroll
    for (int d = 0; d < TF_W; d++) {
      const double e = sm_e[(ii + d) * TF_THREADS + tid];
#pragma unroll
      for (int j = 0; j < TF_B; j++)           // weights outside beta_j's own window are 0: skipped at compile time (307 of 559 FMAs remain)
        if (d >= TF_HALF - tf_default_hw(j) && d <= TF_HALF + tf_default_hw(j)) acc[j] = fma(e, c_tf_g[j * TF_W + d], acc[j]);   // constant operand
    }
    unsigned need = 0;                         // betas that take the exact path below
#pragma unroll
    for (int j = 0; j < TF_B; j++) {
      double res = SCAPE_SENTINEL;
      const int hw = tf_default_hw(j);
      const double dl = edge ? s_dl[ii][j] : 0.0;
      if (acc[j] > 1e-290) {
        res = (log(acc[j]) + M) + dl;
      } else if ((live & (((2ull << (2 * hw)) - 1ull) << (ii + TF_HALF - hw))) != 0ull) {
        need |= 1u << j;
      }
      // (else: beta_j's own window holds incompatible thetas only -- the common case at the edge of the
      // fragment's compatible range -- and the exact path would return the sentinel)
      out[ii * TF_B + j] = TT(res);
    }
    // exact two-pass log-sum-exp over beta_j's own (existing) window (taichi_core.py:41-54, 172-179): rare,
    // and ONE copy of the code behind the 13 unrolled common paths (inlined into each of them it spread
    // the hot path over 190 KB of instructions: 27 % of the stall samples were instruction fetch)
    while (need) {
      const int j = __ffs(need) - 1;
      need &= need - 1;
      const int hw = c_tf_hw[j];
      const double dl = edge ? s_dl[ii][j] : 0.0;
      const int lo = edge ? s_lo[ii][j] : TF_HALF - hw, hi = edge ? s_hi[ii][j] : TF_HALF + hw;
      const double* col = tab + (int64_t)(c_first + ii) * ld;     // column of d = 0 for this row
      const double* lp = c_tf_lp + j * TF_W;
      const double lps = c_tf_lps[j] - dl;
      double m = -CUDART_INF;
      for (int d = lo; d <= hi; d++) m = fmax(m, (col[(int64_t)d * ld] + lp[d]) - lps);
      double res = SCAPE_SENTINEL;             // log(w) + sentinel == sentinel in FP64
      if (m >= -1e30) {
        double sum = 0.0;
        for (int d = lo; d <= hi; d++) {
          const double a = ((col[(int64_t)d * ld] + lp[d]) - lps) - m;
          if (a > -746.0) sum += exp(a);
        }
        res = log(sum) + m;
      }
      out[ii * TF_B + j] = TT(res);
    }
  }
}

void launch_tensor_interior(const UtrDev* utrs, const TileRef* tiles, int64_t n_tiles, int max_n, const double* table,
                            void* tensor, bool f32, cudaStream_t st) {
  if (n_tiles <= 0) return;
  dim3 grid((unsigned)n_tiles, (unsigned)((max_n + TF_THREADS - 1) / TF_THREADS));
  const size_t smem = (size_t)TF_COLS * TF_THREADS * sizeof(double);
  if (f32) {
    tensor_interior_kernel<float><<<grid, TF_THREADS, smem, st>>>(utrs, tiles, table, (float*)tensor);
  } else {
    tensor_interior_kernel<double><<<grid, TF_THREADS, smem, st>>>(utrs, tiles, table, (double*)tensor);
  }
}

void launch_tensor(const UtrDev* utrs, const RowRef* rows, int64_t n_rows, int max_n, int n_beta, int max_win,
                   const double* theta, const double* table, void* tensor, bool f32, cudaStream_t st) {
  if (n_rows <= 0) return;
  dim3 grid((unsigned)n_rows, (unsigned)((max_n + 255) / 256));
  size_t smem = (size_t)n_beta * max_win * 2 * sizeof(double) + n_beta * sizeof(double) + 2 * n_beta * sizeof(int);
  if (f32) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(tensor_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tensor_kernel<float><<<grid, 256, smem, st>>>(utrs, rows, max_win, theta, table, (float*)tensor);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(tensor_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tensor_kernel<double><<<grid, 256, smem, st>>>(utrs, rows, max_win, theta, table, (double*)tensor);
  }
}

// ------------------------------------------------------------------------------------------------
// K4: EM, bulk-synchronous.  All chains of a wave advance one coordinate-EM iteration per step; a
// step is two launches:
//
//   E step            em_estep_warp_kernel (one warp per chain, few fragments, early steps),
//                     em_estep_kernel (one CTA per chain) or em_estep_group_kernel (G warps per chain,
//                     device-side work lists; prune refits).  (0) applies the arg-max the previous scan
//                     found for this chain (alpha_k, beta_k update, trace, finalisation of converged
//                     chains), then (1) column refresh (cal_z_k :473-488), count-tempered softmax
//                     (norm_z :490-495), weight update (maximize_ws :498-505, mstep guard :526-529),
//                     ELBO (:559-573) and the convergence test (:743).  Z is never materialised; the
//                     pass leaves v[n] = Z[n,k] cnt[n] in the chain's row of V.
//   em_scan_kernel    max_alpha_beta (:507-523) as a blocked product on the FP64 tensor cores (DMMA
//                     m8n8k4): one CTA per (UTR, block of 256 candidate rows, share of the chain
//                     sub-batches) computes scores[row][chain] = sum_n tensor[n][row] * V[chain][n]
//                     for the running chains of that UTR whose window touches the block (all K, all
//                     restarts), so a tensor block is fetched once per step however many chains need
//                     it, and the step's work is spread over all SMs.  Tensor rows come straight from
//                     HBM / L2 through a branch-free register ring; V is staged in shared memory per
//                     sub-batch of <= 32 chains.  Each (chain, block) leaves its first-maximum
//                     (score, row) in a partials array.
// Only the hull of fragments with v != 0 is visited (other terms are exactly +-0 in the reference's
// sum).  BIC at the end (:702-706).
// ------------------------------------------------------------------------------------------------
// (the E pass, apply_pending and the MMA wrapper live in em_device.cuh)
template <int NK, typename TT>
__device__ void estep_run(EShared& sh, ChainDev& ch, ScanDesc& sd, const UtrDev& u, const TT* __restrict__ A,
                          const double* __restrict__ cnt, double* __restrict__ lz, double* __restrict__ V) {
  constexpr int K = NK - 1;
  const int tid = threadIdx.x;
  const int N = u.N, npad = u.Npad, B = u.B;
  const int64_t R = u.ldR;
  const int it = ch.n_iter;
  if (it == 0) {
    // initial log_zmat: all K+1 columns (em_algo :722-724)
    for (int j = 0; j < NK; j++) {
      const double w = ch.ws[j];
      const double lw = (w <= 0.0) ? SCAPE_SENTINEL : log(w);
      if (tid == 0) ch.lw[j] = lw;
      if (j < K) {
        const int64_t rj = (int64_t)ch.a_idx[j] * B + ch.b_idx[j];
        for (int n = tid; n < N; n += GT) lz[(int64_t)j * npad + n] = lw + (double)A[(int64_t)n * R + rj];
      } else {
        const double val = lw + u.unif_loglik;
        for (int n = tid; n < N; n += GT) lz[(int64_t)j * npad + n] = val;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const int k = ch.k_order[it];
    sh.k = k;
    sh.lwk = ch.lw[k];
    sh.rk = (long long)ch.a_idx[k] * B + ch.b_idx[k];
  }
  __syncthreads();
  const int k = sh.k;
  const double lwk = sh.lwk;
  const int64_t rk = sh.rk;
  bool guard = false;
  double red[NK + 3];
  while (true) {
#pragma unroll
    for (int j = 0; j < NK + 3; j++) red[j] = 0.0;
    int h_lo = N, h_hi = -1;
    if (tid == 0) { sh.hull[0] = N; sh.hull[1] = -1; }
    for (int n = tid; n < N; n += GT) estep_fragment<NK, TT>(n, k, lwk, rk, guard, npad, R, A, cnt, lz, V, red, h_lo, h_hi);
    __syncthreads();
    if (h_hi >= 0) { atomicMin(&sh.hull[0], h_lo); atomicMax(&sh.hull[1], h_hi); }
    block_reduce_sum<NK + 3>(red, sh);
    if (!guard && sh.tot[NK] < 1e-8) {       // mstep guard (:526-529); uniform across the CTA
      guard = true;
      __syncthreads();
      continue;
    }
    break;
  }
  if (tid == 0) estep_epilogue<NK>(ch, sd, u, sh.tot, k, it, sh.hull[0], sh.hull[1]);
}

// one CTA per chain; `chains` are the wave's chains in launch order
template <typename TT>
__global__ void __launch_bounds__(GT, 2)
em_estep_kernel(ChainDev* chains, ScanDesc* descs, const int32_t* __restrict__ index, const UtrDev* __restrict__ utrs,
                const void* __restrict__ tensor, const double* __restrict__ cnt_all, double* lz_all, double* v_all,
                const ScanPartial* __restrict__ partials, int32_t* trace_a, int32_t* trace_b, double* trace_ws,
                int stage) {
  __shared__ EShared sh;
  __shared__ ChainDev s_ch;
  ChainDev& gch = chains[index[blockIdx.x]];
  ScanDesc& sd = descs[index[blockIdx.x]];
  if (gch.state == 0) return;
  const int tid = threadIdx.x;
  if (stage) {
    copy_chain(&s_ch, &gch, tid, GT);
    __syncthreads();
  }
  ChainDev& ch = stage ? s_ch : gch;
  const UtrDev u = utrs[ch.utr];
  if (tid < 32) {
    const int go = apply_pending(ch, sd, u, partials, trace_a, trace_b, trace_ws);
    if (tid == 0) sh.go = go;
  }
  __syncthreads();
  if (!sh.go) {
    if (stage) copy_chain(&gch, &s_ch, tid, GT);
    return;
  }
  const TT* A = (const TT*)tensor + u.tensor_off;
  const double* cnt = cnt_all + u.frag_off;
  double* lz = lz_all + ch.lz_off;
  double* V = v_all + ch.v_off;
  switch (ch.K) {
    case 1: estep_run<2, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 2: estep_run<3, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 3: estep_run<4, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 4: estep_run<5, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 5: estep_run<6, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 6: estep_run<7, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 7: estep_run<8, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 8: estep_run<9, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 9: estep_run<10, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 10: estep_run<11, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 11: estep_run<12, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 12: estep_run<13, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 13: estep_run<14, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 14: estep_run<15, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    case 15: estep_run<16, TT>(sh, ch, sd, u, A, cnt, lz, V); break;
    default: break;
  }
  if (tid == 0 && ch.weights_only && ch.trace_off >= 0) {     // weights-only chains never wait for a scan
    const int64_t o = ch.trace_off + (int64_t)(ch.n_iter - 1) * (SCAPE_B200_KCAP + 1);
    for (int j = 0; j < ch.K; j++) { trace_a[o + j] = ch.a_idx[j]; trace_b[o + j] = ch.b_idx[j]; }
    for (int j = 0; j <= ch.K; j++) trace_ws[o + j] = ch.ws[j];
  }
  if (stage) {
    __syncthreads();
    copy_chain(&gch, &s_ch, tid, GT);
  }
}

// one WARP per chain (8 chains per CTA) for UTRs with few fragments: no block barriers at all.
// BIGK = false handles K = 1..7 (every normal run) with 4 CTAs per SM; BIGK = true handles the
// K = 8..15 chains that only re-runs can create and may use twice the registers.
// WPC = 2 / 4: that many warps per chain (8 / WPC chains per CTA), see estep_warp_run; the chain record
// is always staged in shared memory then, and the first warp of a chain does the serial parts.
template <typename TT, bool BIGK, bool PF, int WPC = 1>
__global__ void __launch_bounds__(GT, BIGK ? 2 : (PF ? 3 : 4))
em_estep_warp_kernel(ChainDev* chains, ScanDesc* descs, const int32_t* __restrict__ index, int n_index,
                     const UtrDev* __restrict__ utrs, const void* __restrict__ tensor,
                     const double* __restrict__ cnt_all, double* lz_all, double* v_all,
                     const ScanPartial* __restrict__ partials, int32_t* trace_a, int32_t* trace_b,
                     double* trace_ws, int stage) {
  constexpr int CPB = GW / WPC;                          // chains per CTA
  __shared__ ChainDev s_ch[CPB];
  __shared__ EPairShared s_pair[WPC > 1 ? CPB : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cs = warp / WPC, sub = warp % WPC, tig = sub * 32 + lane;
  const int slot = blockIdx.x * CPB + cs;
  if (slot >= n_index) return;                           // (every exit up to the E pass is uniform across a chain's warps)
  ChainDev& gch = chains[index[slot]];
  ScanDesc& sd = descs[index[slot]];
  if (gch.state == 0) return;
  if ((gch.K > 7) != BIGK) return;
  if (WPC > 1) stage = 1;
  ChainDev& ch = stage ? s_ch[cs] : gch;
  if (stage) {
    copy_chain(&ch, &gch, tig, 32 * WPC);
    if (WPC > 1) pair_sync(cs + 1, 32 * WPC); else __syncwarp();
  }
  const UtrDev u = utrs[ch.utr];
  int go;
  if (WPC > 1) {
    if (sub == 0) {
      go = apply_pending(ch, sd, u, partials, trace_a, trace_b, trace_ws);
      if (lane == 0) s_pair[cs].go = go;
    }
    pair_sync(cs + 1, 32 * WPC);
    go = s_pair[cs].go;
  } else {
    go = apply_pending(ch, sd, u, partials, trace_a, trace_b, trace_ws);
  }
  if (!go) {
    if (stage && sub == 0) { __syncwarp(); copy_chain(&gch, &ch, lane, 32); }
    return;
  }
  const TT* A = (const TT*)tensor + u.tensor_off;
  const double* cnt = cnt_all + u.frag_off;
  double* lz = lz_all + ch.lz_off;
  double* V = v_all + ch.v_off;
  EPairShared* ps = &s_pair[WPC > 1 ? cs : 0];
  {
    if (!BIGK) {
      switch (ch.K) {
        case 1: estep_warp_run<2, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 2: estep_warp_run<3, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 3: estep_warp_run<4, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 4: estep_warp_run<5, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 5: estep_warp_run<6, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 6: estep_warp_run<7, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 7: estep_warp_run<8, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        default: break;
      }
    } else {
      switch (ch.K) {
        case 8: estep_warp_run<9, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 9: estep_warp_run<10, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 10: estep_warp_run<11, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 11: estep_warp_run<12, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 12: estep_warp_run<13, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 13: estep_warp_run<14, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 14: estep_warp_run<15, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        case 15: estep_warp_run<16, TT, PF, WPC>(ch, sd, u, A, cnt, lz, V, ps, cs + 1); break;
        default: break;
      }
    }
    if (sub != 0) return;                                // the chain's first warp ran the epilogue
    __syncwarp();
    if (lane == 0 && ch.weights_only && ch.trace_off >= 0) {
      const int64_t o = ch.trace_off + (int64_t)(ch.n_iter - 1) * (SCAPE_B200_KCAP + 1);
      for (int j = 0; j < ch.K; j++) { trace_a[o + j] = ch.a_idx[j]; trace_b[o + j] = ch.b_idx[j]; }
      for (int j = 0; j <= ch.K; j++) trace_ws[o + j] = ch.ws[j];
    }
    __syncwarp();
  }
  if (stage) copy_chain(&gch, &ch, lane, 32);
}

// ------------------------------------------------------------------------------------------------
// E step, group kernel: G warps (1, 2, 4 or 8) per chain, 8 / G chains per CTA, persistent CTAs over a
// device-side list of the chains that are still running.
//   * a chain's E pass is a serial loop of ~1 us fragment passes per lane; G warps cut the loop to
//     N / (32 G) passes, so the step is no longer bound by one warp walking a whole UTR, and the
//     jobs are short enough to pack the SMs evenly;
//   * every launch appends the chains that continue to the next step's list, so the late steps
//     launch work for the few running chains only (an all-chains grid costs ~18 us of empty CTAs).
// Groups synchronise with named barriers (bar.sync id, 32 G); __syncthreads() is never used here.
// ------------------------------------------------------------------------------------------------
// (EGroupShared, group_sync, group_reduce_sum and estep_group_run live in em_device.cuh)
// list_in / n_in: the chains to step (n_host >= 0: the host knows the count, else *n_in);
// list_out / n_out: the chains that still run after this step, appended here.
template <typename TT>
__global__ void __launch_bounds__(GT, 2)
em_estep_group_kernel(ChainDev* chains, ScanDesc* descs, const int32_t* __restrict__ list_in, const int32_t* n_in,
                      int n_host, int32_t* list_out, int32_t* n_out, int G, int loop, const UtrDev* __restrict__ utrs,
                      const void* __restrict__ tensor, const double* __restrict__ cnt_all, double* lz_all,
                      double* v_all, const ScanPartial* __restrict__ partials, int32_t* trace_a, int32_t* trace_b,
                      double* trace_ws) {
  __shared__ EGroupShared shg[GW];
  const int gthreads = 32 * G;
  const int gid = threadIdx.x / gthreads, tig = threadIdx.x - gid * gthreads;
  const int gpc = GW / G;
  EGroupShared& sh = shg[gid];
  const int n_jobs = n_host >= 0 ? n_host : *n_in;
  for (int job = blockIdx.x * gpc + gid; job < n_jobs; job += gridDim.x * gpc) {
    const int ci = list_in[job];
    ChainDev& ch = chains[ci];
    ScanDesc& sd = descs[ci];
    if (ch.state == 0) continue;                       // group-uniform (nobody writes this chain meanwhile)
    const UtrDev u = utrs[ch.utr];
    if (tig < 32) {
      const int go = apply_pending(ch, sd, u, partials, trace_a, trace_b, trace_ws);
      if (tig == 0) sh.go = go;
    }
    group_sync(gid, gthreads);
    const bool go = sh.go != 0;
    if (go) {
      const TT* A = (const TT*)tensor + u.tensor_off;
      const double* cnt = cnt_all + u.frag_off;
      double* lz = lz_all + ch.lz_off;
      double* V = v_all + ch.v_off;
      bool again;
      do {
        switch (ch.K) {
          case 1: estep_group_run<2, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 2: estep_group_run<3, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 3: estep_group_run<4, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 4: estep_group_run<5, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 5: estep_group_run<6, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 6: estep_group_run<7, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 7: estep_group_run<8, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 8: estep_group_run<9, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 9: estep_group_run<10, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 10: estep_group_run<11, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 11: estep_group_run<12, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 12: estep_group_run<13, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 13: estep_group_run<14, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 14: estep_group_run<15, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          case 15: estep_group_run<16, TT>(sh, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
          default: break;
        }
        if (tig == 0 && ch.weights_only && ch.trace_off >= 0) {   // weights-only chains never wait for a scan
          const int64_t o = ch.trace_off + (int64_t)(ch.n_iter - 1) * (SCAPE_B200_KCAP + 1);
          for (int j = 0; j < ch.K; j++) { trace_a[o + j] = ch.a_idx[j]; trace_b[o + j] = ch.b_idx[j]; }
          for (int j = 0; j <= ch.K; j++) trace_ws[o + j] = ch.ws[j];
        }
        group_sync(gid, gthreads);                     // the epilogue's writes to the chain are visible
        // weights-only chains (prune refits) never wait for a scan: run them to convergence here
        again = loop && ch.weights_only && ch.state == 1 && ch.n_iter < SCAPE_B200_NROUND;
      } while (again);
    }
    if (tig == 0 && ch.state != 0) list_out[atomicAdd(n_out, 1)] = ci;
    group_sync(gid, gthreads);                         // sh is reused by the next job
  }
}

// ------------------------------------------------------------------------------------------------
// scan
// ------------------------------------------------------------------------------------------------
struct ScanShared {
  int list[SCAN_MAXCH];          // chain indices (into the wave's chain array) that need this block
  int n_list, N0, N1;
  double wbest[GW][SCAN_GB];
  int wrow[GW][SCAN_GB];
  int w0[SCAN_GB], w1[SCAN_GB];  // candidate windows of the sub-batch's chains
  long long voff[SCAN_GB], pboff[SCAN_GB];
  PassCtx px;                    // tile path: the sub-batch's chains (windows, hulls, staged V rows)
};

// One sub-batch of up to 8*NG chains against this CTA's 256 candidate rows:
//   scores[row][chain] = sum_n tensor[n][row] * V[chain][n]        (max_alpha_beta's np.sum, :522)
// as a blocked FP64 matrix product on the tensor cores.  M = candidate rows (8 per MMA, 32 per
// warp), N = chains (8 per MMA), K = fragments (4 per MMA).  A fragments come straight from the
// [n][row] tensor (32-byte segments, software-prefetched one k-step ahead, converted to FP64 once),
// B fragments from V staged in shared memory with a conflict-free pitch.  The operand reuse that a
// CUDA-core version has to buy with shared-memory broadcasts (128 B/clk/SM, i.e. <= 1 FMA pair per
// 4 SM cycles) happens inside the MMA datapath here.
// f32 -> f64 of a tensor element.  F2F.F64.F32 runs on the XU pipe (16 per clock and SM); for the finite,
// normal values a tensor holds (log-likelihoods and the float sentinel) the widening is also an exponent
// re-bias and a mantissa shift in integer instructions.  CVT: 0 = all on the XU pipe, 1 = rows with odd mi
// in integer instructions, 2 = all in integer instructions.  (+-0, which only the pitch padding holds,
// would come out as 2^-127; those columns are never candidate rows, and slack fragments meet V = 0.)
template <int CVT>
__device__ __forceinline__ double scan_widen(float f, int mi) {
  if (CVT == 0 || (CVT == 1 && (mi & 1) == 0)) return (double)f;
  const uint32_t x = __float_as_uint(f);
  const uint32_t hi = (((x & 0x7fffffffu) >> 3) + 0x38000000u) | (x & 0x80000000u);
  return __hiloint2double((int)hi, (int)(x << 29));
}
template <int CVT>
__device__ __forceinline__ double scan_widen(double d, int) { return d; }

template <int NG, typename TT, int CVT = 0>
__device__ __forceinline__ void scan_subbatch(ScanShared& sh, const ScanDesc* __restrict__ descs, const UtrDev& u,
                                              const TT* __restrict__ A, const double* __restrict__ v_all,
                                              ScanPartial* partials, int first, int cnt, int blk, double* Vs,
                                              double* scan_elems) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;         // MMA group id / thread-in-group
  const int64_t R = u.ldR;
  const int N = u.N;
  const int Rv = u.T * u.B;                     // valid candidate rows
  const int base = blk * SCAN_ROWS;
  const int blk_end = min(base + SCAN_ROWS, Rv);
  // fragment hull of this sub-batch (union over its chains) and the chains' windows
  if (tid < 32) {                                // one full warp (the shuffles below need all 32 lanes)
    int lo = 0, hi = 0, h0 = 1 << 30, h1 = 0;
    if (tid < cnt) {
      const ScanDesc d = descs[sh.list[first + tid]];
      lo = d.row0; hi = d.row1;
      if (d.hhi >= 0) { h0 = d.hlo; h1 = d.hhi + 1; }
      sh.voff[tid] = d.v_off;
      sh.pboff[tid] = d.pb_off;
    }
    if (tid < SCAN_GB) {
      sh.w0[tid] = lo;
      sh.w1[tid] = hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      h0 = min(h0, __shfl_xor_sync(0xffffffffu, h0, o));
      h1 = max(h1, __shfl_xor_sync(0xffffffffu, h1, o));
    }
    if (tid == 0) {
      if (h1 <= h0) { h0 = 0; h1 = 0; }         // every v is zero: all scores 0, first row of each window wins
      sh.N0 = h0 & ~7;                          // aligned start (V is exactly 0 outside a chain's hull)
      sh.N1 = h1;
      if (scan_elems) atomicAdd(scan_elems, (double)(blk_end - base) * (double)(h1 - (h0 & ~7)));
    }
  }
  __syncthreads();
  const int N0 = sh.N0, N1 = sh.N1;
  const uint32_t vs_base = (uint32_t)__cvta_generic_to_shared(Vs);
  // this lane's A rows: row(mi) = base + 32*warp + 8*mi + g  (clamped; out-of-range rows are masked later)
  const TT* arow[4];
#pragma unroll
  for (int mi = 0; mi < 4; mi++) arow[mi] = A + min(base + 32 * warp + 8 * mi + g, Rv - 1);

  double acc[4][NG][2];
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < NG; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  for (int c0 = N0; c0 < N1; c0 += SCAN_VCHUNK) {
    const int c1 = min(c0 + SCAN_VCHUNK, N1);
    const int len = c1 - c0, len4 = (len + 3) & ~3;
    __syncthreads();                            // previous chunk / sub-batch fully consumed
    for (int e = tid; e < 8 * NG * len4; e += GT) {
      const int j = e / len4, o = e - j * len4;
      Vs[j * SCAN_VPITCH + o] = (j < cnt && o < len) ? v_all[sh.voff[j] + c0 + o] : 0.0;
    }
    __syncthreads();
    // B fragment of this lane: V[chain = 8*ni + g][n = k0 + q]
    const uint32_t vb = vs_base + (uint32_t)(g * SCAN_VPITCH + q) * 8u;
    // A fragments are fetched PFD k-steps (4*PFD fragments) ahead through a register ring.  The ring
    // is refilled unconditionally with plain pointer increments (no clamp, no branch: a conditional
    // refill makes the compiler insert register moves that wait on the load just issued).  Reads may
    // run up to 4*PFD+3 fragments past the chunk: the tensor arena carries zeroed slack for that, and
    // V is zero-padded, so nothing past the hull contributes.
    constexpr int PFD = NG == 1 ? 12 : NG == 2 ? 8 : 4;   // fewer MMAs per k-step -> prefetch further ahead
    const int64_t kstep = 4 * R;
    const TT* pp[4];
    TT pre[PFD][4];
#pragma unroll
    for (int mi = 0; mi < 4; mi++) pp[mi] = arow[mi] + (int64_t)(c0 + q) * R;
#pragma unroll
    for (int p = 0; p < PFD; p++)
#pragma unroll
      for (int mi = 0; mi < 4; mi++) {
        pre[p][mi] = __ldg(pp[mi]);
        pp[mi] += kstep;
      }
    int kk = 0;
    for (; kk + 4 * PFD <= len4; kk += 4 * PFD) {
#pragma unroll
      for (int p = 0; p < PFD; p++) {
        double a[4];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
          a[mi] = scan_widen<CVT>(pre[p][mi], mi);
          pre[p][mi] = __ldg(pp[mi]);
          pp[mi] += kstep;
        }
        double b[NG];
#pragma unroll
        for (int ni = 0; ni < NG; ni++) b[ni] = lds_f64(vb + (uint32_t)(ni * 8 * SCAN_VPITCH + kk + 4 * p) * 8u);
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
          for (int ni = 0; ni < NG; ni++) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
      }
    }
#pragma unroll
    for (int p = 0; p < PFD; p++) {                        // remainder: fewer than PFD k-steps, ring already holds them
      if (kk + 4 * p < len4) {
        double b[NG];
#pragma unroll
        for (int ni = 0; ni < NG; ni++) b[ni] = lds_f64(vb + (uint32_t)(ni * 8 * SCAN_VPITCH + kk + 4 * p) * 8u);
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
          for (int ni = 0; ni < NG; ni++) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], scan_widen<CVT>(pre[p][mi], mi), b[ni]);
      }
    }
  }
  // first maximum of this block per chain: larger score wins, ties go to the smaller row.
  // acc[mi][ni][i] = score[row = base + 32*warp + 8*mi + g][chain = 8*ni + 2*q + i]
#pragma unroll
  for (int ni = 0; ni < NG; ni++) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int c = 8 * ni + 2 * q + i;
      const int w0 = sh.w0[c], w1 = min(sh.w1[c], blk_end);
      double b = -CUDART_INF;
      int r = 0x7fffffff;
#pragma unroll
      for (int mi = 0; mi < 4; mi++) {
        const int row = base + 32 * warp + 8 * mi + g;
        if (row >= w0 && row < w1 && acc[mi][ni][i] > b) { b = acc[mi][ni][i]; r = row; }   // rows ascend with mi
      }
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {        // lanes with the same q hold the same chain
        const double ob = __shfl_xor_sync(0xffffffffu, b, o);
        const int orow = __shfl_xor_sync(0xffffffffu, r, o);
        if (ob > b || (ob == b && orow < r)) { b = ob; r = orow; }
      }
      if (g == 0) { sh.wbest[warp][c] = b; sh.wrow[warp][c] = r; }
    }
  }
  __syncthreads();
  if (tid < cnt) {
    double b = sh.wbest[0][tid];
    int r = sh.wrow[0][tid];
    for (int w = 1; w < GW; w++) {
      const double ob = sh.wbest[w][tid];
      const int orow = sh.wrow[w][tid];
      if (ob > b || (ob == b && orow < r)) { b = ob; r = orow; }
    }
    ScanPartial p;
    p.score = b; p.row = r; p.pad = 0;
    partials[sh.pboff[tid] + blk] = p;
  }
}

// Tile path of a sub-batch (UTRs whose whole V rows fit the CTA's shared memory): the 8 warps take the
// 8 32-row tiles of the block, and every warp multiplies only the chains of the sub-batch whose window
// covers ITS tile, over the hull of THOSE chains (scan_tile, em_device.cuh).  scan_subbatch above
// multiplies all 256 rows by all chains of the sub-batch over the union of all their hulls: measured
// 1.9x the algorithmic MMAs (3x what windows and hulls strictly need).  Per-chain sums are
// bit-identical in both paths (4-fragment MMA steps aligned to multiples of 4, ascending).
template <typename TT>
__device__ __forceinline__ void scan_subbatch_tiles(ScanShared& sh, const ScanDesc* __restrict__ descs, const UtrDev& u,
                                                    const TT* __restrict__ A, const double* __restrict__ v_all,
                                                    ScanPartial* partials, int first, int cnt, int blk, double* Vs,
                                                    double* scan_elems) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  PassCtx& px = sh.px;
  if (warp == 0) {
    int h0 = 1 << 30, h1 = 0;
    if (lane < cnt) {
      const ScanDesc d = descs[sh.list[first + lane]];
      px.row0[lane] = d.row0; px.row1[lane] = d.row1; px.hlo[lane] = d.hlo; px.hhi[lane] = d.hhi;
      px.voff[lane] = d.v_off;
      px.pboff[lane] = d.pb_off;
      if (d.hhi >= 0) { h0 = d.hlo; h1 = d.hhi + 1; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      h0 = min(h0, __shfl_xor_sync(0xffffffffu, h0, o));
      h1 = max(h1, __shfl_xor_sync(0xffffffffu, h1, o));
    }
    if (lane == 0) {
      if (h1 <= h0) { h0 = 0; h1 = 0; }
      px.NA = h0 & ~7;
      px.NB = h1;
      px.P = cluster_v_pitch(h1 - (h0 & ~7));
    }
  }
  sh.wbest[warp][lane] = -CUDART_INF;                    // chains that do not cover this warp's tile
  sh.wrow[warp][lane] = 0x7fffffff;
  __syncthreads();
  const int NA = px.NA, NB = px.NB, P = px.P;
  for (int j = warp; j <= cnt; j += GW) {                // row cnt = zeros (MMA columns without a chain)
    double* dst = Vs + j * P;
    const double* src = v_all + (j < cnt ? px.voff[j] : 0) + NA;
    for (int o = lane; o < P; o += 32) dst[o] = (j < cnt && NA + o < NB) ? src[o] : 0.0;
  }
  __syncthreads();
  scan_tile<TT, true>(px, u, A, Vs, partials, blk * (SCAN_ROWS / kClusterTileRows) + warp, cnt, scan_elems, sh.wbest[warp],
                      sh.wrow[warp]);
  __syncthreads();
  if (tid < cnt) {
    double b = sh.wbest[0][tid];
    int r = sh.wrow[0][tid];
    for (int w = 1; w < GW; w++) {                       // tiles ascend with the warp index: ties keep the smaller row
      const double ob = sh.wbest[w][tid];
      const int orow = sh.wrow[w][tid];
      if (ob > b || (ob == b && orow < r)) { b = ob; r = orow; }
    }
    ScanPartial p;
    p.score = b; p.row = r; p.pad = 0;
    partials[px.pboff[tid] + blk] = p;
  }
}

// one CTA per (UTR, block of SCAN_ROWS candidate rows); TILES: the tile path (its own kernel: the two
// paths' register allocations do not disturb each other)
template <typename TT, bool TILES, int CVT = 0>
__global__ void __launch_bounds__(GT, 2)
em_scan_kernel(const ScanRef* __restrict__ refs, const ScanDesc* __restrict__ descs, const UtrDev* __restrict__ utrs,
               const int32_t* __restrict__ utr_chain_off, const void* __restrict__ tensor,
               const double* __restrict__ v_all, ScanPartial* partials, double* scan_elems) {
  extern __shared__ double sm_dyn[];
  __shared__ ScanShared sh;
  const ScanRef ref = refs[blockIdx.x];
  const UtrDev u = utrs[ref.utr];
  const int tid = threadIdx.x;
  const int Rv = u.T * u.B;
  const int lo = ref.blk * SCAN_ROWS, hi = min(lo + SCAN_ROWS, Rv);
  if (tid == 0) sh.n_list = 0;
  __syncthreads();
  const int c_begin = utr_chain_off[ref.utr], c_end = utr_chain_off[ref.utr + 1];
  for (int c = c_begin + tid; c < c_end; c += GT) {
    const ScanDesc d = descs[c];
    if (d.pending && d.row0 < hi && d.row1 > lo) {
      const int slot = atomicAdd(&sh.n_list, 1);
      if (slot < SCAN_MAXCH) sh.list[slot] = c;
    }
  }
  __syncthreads();
  const int n_list = min(sh.n_list, SCAN_MAXCH);
  if (ref.sb * ref.gb >= n_list) return;                          // nothing (left) for this CTA's share
  // deterministic sub-batches: order the list by chain index (rank sort, one thread per entry;
  // the indices are distinct, so the ranks are a permutation)
  {
    int mine = 0, rank = 0;
    if (tid < n_list) {
      mine = sh.list[tid];
      for (int i = 0; i < n_list; i++) rank += sh.list[i] < mine;
    }
    __syncthreads();
    if (tid < n_list) sh.list[rank] = mine;
    __syncthreads();
  }
  double* Vs = sm_dyn;                                           // [SCAN_GB][SCAN_VPITCH]
  const TT* A = (const TT*)tensor + u.tensor_off;
#define SCAN_CALL(G) scan_subbatch<G, TT, CVT>(sh, descs, u, A, v_all, partials, first, cnt, ref.blk, Vs, scan_elems)
  const int gb = ref.gb;
  for (int first = ref.sb * gb; first < n_list; first += ref.nsb * gb) {
    const int cnt = min(gb, n_list - first);
    if (TILES) scan_subbatch_tiles<TT>(sh, descs, u, A, v_all, partials, first, cnt, ref.blk, Vs, scan_elems);
    else if (cnt <= 8) SCAN_CALL(1);
    else if (cnt <= 16) SCAN_CALL(2);
    else if (cnt <= 24) SCAN_CALL(3);
    else SCAN_CALL(4);
    __syncthreads();
  }
#undef SCAN_CALL
}

// One EM run = NROUND steps of {estep, scan} plus a closing estep that applies the last arg-max.
// `index_dev` lists the chains with few fragments first (warp-per-chain kernel, n_small of them),
// then the others (block-per-chain kernel).
template <typename TT>
static int launch_em_steps_t(ChainDev* chains_dev, ScanDesc* descs_dev, const int32_t* index_dev, int64_t n_small, int64_t n_big,
                             bool any_scan, bool big_k, const ScanRef* refs_dev, int64_t n_refs, int64_t n_refs_tile, const UtrDev* utrs_dev,
                             const int32_t* utr_chain_off_dev, const void* tensor, const double* cnt, double* lz,
                             double* vbuf, void* partials, double* scan_elems, int32_t* trace_a, int32_t* trace_b,
                             double* trace_ws, cudaStream_t st, std::vector<cudaEvent_t>& evs,
                             std::vector<int>& kinds, int& scan_launches, const std::function<void()>& hook,
                             const std::function<void()>& hook_mark, int hook_step, const EstepPlan& plan, int n_steps, bool timing) {
  const size_t smem = (size_t)SCAN_GB * SCAN_VPITCH * sizeof(double);
  // f32 -> f64 widening of the tensor elements in the scan: SCAPE_B200_SCAN_CVT = 0 XU pipe, 1 half in integer
  // instructions, 2 all (only with f32 storage; identical values)
  static const int scan_cvt_env = getenv("SCAPE_B200_SCAN_CVT") ? atoi(getenv("SCAPE_B200_SCAN_CVT")) : 0;
  const int scan_cvt = sizeof(TT) == 4 ? scan_cvt_env : 0;
  cudaFuncSetAttribute(em_scan_kernel<TT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (scan_cvt == 1) cudaFuncSetAttribute(em_scan_kernel<TT, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (scan_cvt == 2) cudaFuncSetAttribute(em_scan_kernel<TT, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(em_scan_kernel<TT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanTileVBytes);
  // the scan of one step: work items of the chunked path first (n_refs of them), then those of the tile path
  auto scan_step = [&]() {
    if (n_refs > 0) {
      if (scan_cvt == 1)
        em_scan_kernel<TT, false, 1><<<(unsigned)n_refs, GT, smem, st>>>(refs_dev, descs_dev, utrs_dev, utr_chain_off_dev, tensor,
                                                                           vbuf, (ScanPartial*)partials, scan_elems);
      else if (scan_cvt == 2)
        em_scan_kernel<TT, false, 2><<<(unsigned)n_refs, GT, smem, st>>>(refs_dev, descs_dev, utrs_dev, utr_chain_off_dev, tensor,
                                                                           vbuf, (ScanPartial*)partials, scan_elems);
      else
        em_scan_kernel<TT, false><<<(unsigned)n_refs, GT, smem, st>>>(refs_dev, descs_dev, utrs_dev, utr_chain_off_dev, tensor,
                                                                        vbuf, (ScanPartial*)partials, scan_elems);
      scan_launches++;
    }
    if (n_refs_tile > 0) {
      em_scan_kernel<TT, true><<<(unsigned)n_refs_tile, GT, kScanTileVBytes, st>>>(refs_dev + n_refs, descs_dev, utrs_dev,
                                                                                    utr_chain_off_dev, tensor, vbuf,
                                                                                    (ScanPartial*)partials, scan_elems);
      scan_launches++;
    }
    return (n_refs > 0) + (n_refs_tile > 0);
  };
  int launches = 0;
  static const int warp_steps = getenv("SCAPE_B200_WARP_STEPS") ? atoi(getenv("SCAPE_B200_WARP_STEPS")) : 32;   // (24 with one warp per chain; 32 measured best with two: E step -4 %)
  static const bool dbg = scape_env_on("SCAPE_B200_DBG");   // print per-launch timings (development aid)
  // events around every launch group: [E step | scan] per step; read back by em_steps_elapsed()
  evs.resize(size_t(2 * (SCAPE_B200_NROUND + 1) + 1));
  for (auto& e : evs)
    if (!e) cudaEventCreateWithFlags(&e, cudaEventDefault);
  kinds.clear();
  size_t ne = 0;
  auto mark = [&](int kind) {
    if (!timing) return;
    cudaEventRecord(evs[ne++], st);
    kinds.push_back(kind);
  };
  mark(-1);
  const bool group = plan.group_steps || !any_scan;
  if (group) {
    // ---- group E step: persistent CTAs over device-side lists of the running chains ---------------
    int32_t* Ls[2] = {plan.lists, plan.lists + n_small};
    int32_t* Lb[2] = {plan.lists + 2 * n_small, plan.lists + 2 * n_small + n_big};
    int32_t* cs = plan.counts;
    int32_t* cb = plan.counts + (SCAPE_B200_NROUND + 2);
    cudaMemsetAsync(plan.counts, 0, sizeof(int32_t) * 2 * (SCAPE_B200_NROUND + 2), st);
    const int Gs = plan.g_small, Gb = GW;
    const unsigned max_grid = (unsigned)(plan.n_sm * 2);
    const unsigned grid_s = (unsigned)std::min<int64_t>((n_small + GW / Gs - 1) / (GW / Gs), max_grid);
    const unsigned grid_b = (unsigned)std::min<int64_t>(n_big, max_grid);
    auto estep = [&](int step, int loop) {
      if (n_small > 0) {
        em_estep_group_kernel<TT><<<grid_s, GT, 0, st>>>(
            chains_dev, descs_dev, step == 0 ? index_dev : Ls[step & 1], cs + step, step == 0 ? (int)n_small : -1,
            Ls[(step + 1) & 1], cs + step + 1, Gs, loop, utrs_dev, tensor, cnt, lz, vbuf, (const ScanPartial*)partials,
            trace_a, trace_b, trace_ws);
        launches++;
      }
      if (n_big > 0) {
        em_estep_group_kernel<TT><<<grid_b, GT, 0, st>>>(
            chains_dev, descs_dev, step == 0 ? index_dev + n_small : Lb[step & 1], cb + step, step == 0 ? (int)n_big : -1,
            Lb[(step + 1) & 1], cb + step + 1, Gb, loop, utrs_dev, tensor, cnt, lz, vbuf, (const ScanPartial*)partials,
            trace_a, trace_b, trace_ws);
        launches++;
      }
      mark(0);
    };
    if (!any_scan) {                                   // prune refits: whole chains inside one launch
      estep(0, 1);
    } else {
      for (int step = 0; step < n_steps; step++) {
        estep(step, 0);
        if (step == SCAPE_B200_NROUND || n_refs + n_refs_tile == 0) continue;
        launches += scan_step();
        mark(1);
        if (hook && step == hook_step) { if (hook_mark) hook_mark(); hook(); }
      }
    }
  }
  // Late steps (>= warp_steps) of a run with a grid search: few chains still run.  late_group: the group
  // kernel over device-side lists of the running chains (persistent CTAs, 4 / 8 warps per chain) --
  // a launch of <= 2 CTAs per SM instead of one CTA per chain of the wave, most of which exit at once.
  // Measured: no gain (scheduling 5,000 empty CTAs costs ~4 us, scripts/micro/launch_rate.cu).
  static const bool late_group = getenv("SCAPE_B200_LATE_GROUP") ? atoi(getenv("SCAPE_B200_LATE_GROUP")) != 0 : false;   // opt-in: measured no gain
  int32_t* hLs[2] = {plan.lists, plan.lists + n_small};
  int32_t* hLb[2] = {plan.lists + 2 * n_small, plan.lists + 2 * n_small + n_big};
  int32_t* hcs = plan.counts;
  int32_t* hcb = plan.counts + (SCAPE_B200_NROUND + 2);
  if (!group && late_group && n_steps > warp_steps)
    cudaMemsetAsync(plan.counts, 0, sizeof(int32_t) * 2 * (SCAPE_B200_NROUND + 2), st);
  for (int step = 0; !group && step < n_steps; step++) {
    // Early steps: most chains run -> one warp per chain (throughput).  Late steps: few chains run
    // and the step time is the latency of ONE chain's E pass -> several warps per chain.
    const bool wide = step < warp_steps;
    if (!wide && late_group) {
      const bool first = step == warp_steps;
      const unsigned max_grid = (unsigned)(plan.n_sm * 2);
      if (n_small > 0) {
        const int Gs = plan.g_small;
        const unsigned grid_s = (unsigned)std::min<int64_t>((n_small + GW / Gs - 1) / (GW / Gs), max_grid);
        em_estep_group_kernel<TT><<<grid_s, GT, 0, st>>>(
            chains_dev, descs_dev, first ? index_dev : hLs[step & 1], hcs + step, first ? (int)n_small : -1, hLs[(step + 1) & 1],
            hcs + step + 1, Gs, 0, utrs_dev, tensor, cnt, lz, vbuf, (const ScanPartial*)partials, trace_a, trace_b, trace_ws);
        launches++;
      }
      if (n_big > 0) {
        const unsigned grid_b = (unsigned)std::min<int64_t>(n_big, max_grid);
        em_estep_group_kernel<TT><<<grid_b, GT, 0, st>>>(
            chains_dev, descs_dev, first ? index_dev + n_small : hLb[step & 1], hcb + step, first ? (int)n_big : -1,
            hLb[(step + 1) & 1], hcb + step + 1, GW, 0, utrs_dev, tensor, cnt, lz, vbuf, (const ScanPartial*)partials, trace_a,
            trace_b, trace_ws);
        launches++;
      }
      mark(0);
      if (step == SCAPE_B200_NROUND || !any_scan || n_refs + n_refs_tile == 0) continue;
      launches += scan_step();
      mark(1);
      if (hook && step == hook_step) { if (hook_mark) hook_mark(); hook(); }
      continue;
    }
    // wide steps: the CTA-per-chain kernel (many-fragment chains, the longer latency floor) first, on its
    // own stream when the plan has one; the warp-per-chain kernel runs beside it
    const int64_t n_blk = wide ? n_big : n_small + n_big;
    const bool fork = wide && n_small > 0 && n_blk > 0 && plan.st_big != nullptr;
    if (n_blk > 0) {
      cudaStream_t sb = fork ? plan.st_big : st;
      if (fork) {
        cudaEventRecord(plan.ev_big[0], st);               // behind the previous step's scan
        cudaStreamWaitEvent(sb, plan.ev_big[0], 0);
      }
      em_estep_kernel<TT><<<(unsigned)n_blk, GT, 0, sb>>>(chains_dev, descs_dev, wide ? index_dev + n_small : index_dev,
                                                           utrs_dev, tensor, cnt, lz, vbuf, (const ScanPartial*)partials,
                                                           trace_a, trace_b, trace_ws, plan.stage_chain);
      launches++;
      if (fork) cudaEventRecord(plan.ev_big[1], sb);
    }
    if (n_small > 0 && wide && plan.warp_wpc > 1) {
      const int cpb = GW / plan.warp_wpc;
      const unsigned g = (unsigned)((n_small + cpb - 1) / cpb);
#define SCAPE_WARP_ARGS chains_dev, descs_dev, index_dev, (int)n_small, utrs_dev, tensor, cnt, lz, vbuf, (const ScanPartial*)partials, trace_a, trace_b, trace_ws, plan.stage_chain
      if (plan.warp_wpc == 2 && !plan.warp_prefetch) em_estep_warp_kernel<TT, false, false, 2><<<g, GT, 0, st>>>(SCAPE_WARP_ARGS);   // 64 registers, 4 CTAs per SM
      else if (plan.warp_wpc == 2) em_estep_warp_kernel<TT, false, true, 2><<<g, GT, 0, st>>>(SCAPE_WARP_ARGS);
      else em_estep_warp_kernel<TT, false, true, 4><<<g, GT, 0, st>>>(SCAPE_WARP_ARGS);
      launches++;
      if (big_k) {
        if (plan.warp_wpc == 2) em_estep_warp_kernel<TT, true, true, 2><<<g, GT, 0, st>>>(SCAPE_WARP_ARGS);
        else em_estep_warp_kernel<TT, true, true, 4><<<g, GT, 0, st>>>(SCAPE_WARP_ARGS);
        launches++;
      }
#undef SCAPE_WARP_ARGS
    } else if (n_small > 0 && wide) {
      if (plan.warp_prefetch)
        em_estep_warp_kernel<TT, false, true><<<(unsigned)((n_small + GW - 1) / GW), GT, 0, st>>>(
            chains_dev, descs_dev, index_dev, (int)n_small, utrs_dev, tensor, cnt, lz, vbuf, (const ScanPartial*)partials, trace_a,
            trace_b, trace_ws, plan.stage_chain);
      else
        em_estep_warp_kernel<TT, false, false><<<(unsigned)((n_small + GW - 1) / GW), GT, 0, st>>>(
            chains_dev, descs_dev, index_dev, (int)n_small, utrs_dev, tensor, cnt, lz, vbuf, (const ScanPartial*)partials, trace_a,
            trace_b, trace_ws, plan.stage_chain);
      launches++;
      if (big_k) {
        em_estep_warp_kernel<TT, true, true><<<(unsigned)((n_small + GW - 1) / GW), GT, 0, st>>>(
            chains_dev, descs_dev, index_dev, (int)n_small, utrs_dev, tensor, cnt, lz, vbuf, (const ScanPartial*)partials,
            trace_a, trace_b, trace_ws, plan.stage_chain);
        launches++;
      }
    }
    if (fork) cudaStreamWaitEvent(st, plan.ev_big[1], 0);
    mark(0);
    if (step == SCAPE_B200_NROUND || !any_scan || n_refs + n_refs_tile == 0) continue;
    launches += scan_step();
    mark(1);
    if (hook && step == hook_step) { if (hook_mark) hook_mark(); hook(); }
  }
  if (dbg) {
    cudaStreamSynchronize(st);
    std::string le, ls;
    for (size_t i = 1; i < kinds.size(); i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, evs[i - 1], evs[i]);
      (kinds[i] == 1 ? ls : le) += std::to_string((int)(ms * 1000)) + " ";
    }
    fprintf(stderr, "em run: small=%lld big=%lld refs=%lld\n  estep us/step: %s\n  scan us/step: %s\n", (long long)n_small, (long long)n_big, (long long)n_refs, le.c_str(), ls.c_str());
  }
  return launches;
}

int launch_em_steps(ChainDev* chains_dev, ScanDesc* descs_dev, const int32_t* index_dev, int64_t n_small, int64_t n_big, bool any_scan,
                    bool big_k,
                    const ScanRef* refs_dev, int64_t n_refs, int64_t n_refs_tile, const UtrDev* utrs_dev,
                    const int32_t* utr_chain_off_dev, const void* tensor, bool f32, const double* cnt, double* lz,
                    double* vbuf, void* partials, double* scan_elems, int32_t* trace_a, int32_t* trace_b,
                    double* trace_ws, cudaStream_t st, EmStepEvents& ee, const EstepPlan& plan, int n_steps) {
  ee.scan_launches = 0;
  if (f32)
    return launch_em_steps_t<float>(chains_dev, descs_dev, index_dev, n_small, n_big, any_scan, big_k, refs_dev, n_refs, n_refs_tile, utrs_dev,
                                    utr_chain_off_dev, tensor, cnt, lz, vbuf, partials, scan_elems, trace_a, trace_b,
                                    trace_ws, st, ee.evs, ee.kinds, ee.scan_launches, ee.hook, ee.mark, ee.hook_step, plan, n_steps, ee.timing || dbg_env());
  return launch_em_steps_t<double>(chains_dev, descs_dev, index_dev, n_small, n_big, any_scan, big_k, refs_dev, n_refs, n_refs_tile, utrs_dev,
                                   utr_chain_off_dev, tensor, cnt, lz, vbuf, partials, scan_elems, trace_a, trace_b,
                                   trace_ws, st, ee.evs, ee.kinds, ee.scan_launches, ee.hook, ee.mark, ee.hook_step, plan, n_steps, ee.timing || dbg_env());
}

// After the stream has been synchronised: total E-step and scan kernel time of the last run.
void em_steps_elapsed(const EmStepEvents& ee, double* estep_ms, double* scan_ms) {
  *estep_ms = *scan_ms = 0;
  for (size_t i = 1; i < ee.kinds.size(); i++) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ee.evs[i - 1], ee.evs[i]);
    (ee.kinds[i] == 1 ? *scan_ms : *estep_ms) += ms;
  }
}

// ------------------------------------------------------------------------------------------------
// FP64 peak microbenchmarks (the roofline denominators of the EM kernels; MEASURED_PEAKS.json has
// no FP64 entry).  Dependent-free FMA / MMA streams in registers, one CTA of 256 threads x 8 per SM.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, int iters, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-9 + i;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters, double seed) {
  double d[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) d[i][0] = d[i][1] = seed + i;
  const double a = 1e-3 * (threadIdx.x & 3), b = 1e-3 * (threadIdx.x >> 2);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma_8x8x4(d[i][0], d[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += d[i][0] + d[i][1];
  if (s == 12345.678) out[0] = s;
}

// FP32 FMA, MUFU (ex2.approx) and FP64 exp() / log() streams: the denominators of the likelihood
// phases (theta table: 13 exp per entry; marginal tensor: 307 exp per (fragment, alpha) in the
// reference) -- BASELINE.md section 3 asks for them, MEASURED_PEAKS.json does not have them.
__global__ void __launch_bounds__(256) peak_ffma_kernel(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-6f + i;
  const float m = 1.0000001f, c = 1e-7f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], m, c);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_mufu_kernel(float* out, int iters, float seed) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = seed * 0.01f + threadIdx.x * 1e-6f + 0.1f * i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));   // stays in [1, 2) after one step: 2^x - ... bounded
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] -= 1.0f;                                                 // back into [0, 1): one FADD per MUFU
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_exp64_kernel(double* out, int iters, double seed) {
  double a[4];
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = seed + threadIdx.x * 1e-9 + 0.1 * i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = exp(a[i] - 1.0);      // fixed point 1: arguments stay in exp()'s main path
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) s += a[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_log64_kernel(double* out, int iters, double seed) {
  double a[4];
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = seed + threadIdx.x * 1e-9 + 0.1 * i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = log(a[i] + 2.0);      // fixed point ~1.146
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) s += a[i];
  if (s == 12345.678) out[0] = s;
}

// out[0..3] = FP32 FMA TFLOP/s, MUFU ex2 Gop/s, FP64 exp() Gop/s, FP64 log() Gop/s
int measure_sfu_peaks(int n_sm, double* out, cudaStream_t st) {
  void* buf = nullptr;
  if (cudaMalloc(&buf, 64) != cudaSuccess) return -1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = n_sm * 8;
  auto timed = [&](int which, int iters) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0, st);
      if (which == 0) peak_ffma_kernel<<<grid, 256, 0, st>>>((float*)buf, iters, 1.0f);
      else if (which == 1) peak_mufu_kernel<<<grid, 256, 0, st>>>((float*)buf, iters, 1.0f);
      else if (which == 2) peak_exp64_kernel<<<grid, 256, 0, st>>>((double*)buf, iters, 1.0);
      else peak_log64_kernel<<<grid, 256, 0, st>>>((double*)buf, iters, 1.0);
      cudaEventRecord(e1, st);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep) best = std::min(best, ms);
    }
    return double(best) * 1e-3;
  };
  const double threads = double(grid) * 256.0;
  out[0] = 2.0 * threads * 8.0 * (1 << 14) / timed(0, 1 << 14) / 1e12;
  out[1] = threads * 8.0 * (1 << 13) / timed(1, 1 << 13) / 1e9;
  out[2] = threads * 4.0 * (1 << 10) / timed(2, 1 << 10) / 1e9;
  out[3] = threads * 4.0 * (1 << 10) / timed(3, 1 << 10) / 1e9;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// returns FP64 TFLOP/s (2 flop per FMA) of CUDA-core DFMA and tensor-core DMMA streams
int measure_fp64_peaks(int n_sm, double* dfma_tflops, double* dmma_tflops, cudaStream_t st) {
  double* buf = nullptr;
  if (cudaMalloc((void**)&buf, 64) != cudaSuccess) return -1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = n_sm * 8, iters = 1 << 14;
  float best_f = 1e30f, best_m = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0, st);
    peak_dfma_kernel<<<grid, 256, 0, st>>>(buf, iters, 1.0);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) best_f = std::min(best_f, ms);
    cudaEventRecord(e0, st);
    peak_dmma_kernel<<<grid, 256, 0, st>>>(buf, iters, 1.0);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) best_m = std::min(best_m, ms);
  }
  *dfma_tflops = 2.0 * grid * 256.0 * 8.0 * iters / (best_f * 1e-3) / 1e12;
  *dmma_tflops = 2.0 * grid * 8.0 /*warps*/ * 8.0 /*mma*/ * 256.0 /*fma per mma*/ * iters / (best_m * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------
// K5: labels.  get_label (:873-881): refresh all columns with the final parameters, tempered
// softmax, first arg-max per fragment.
// ------------------------------------------------------------------------------------------------
template <typename TT>
__global__ void __launch_bounds__(256) label_kernel(const LabelDev* __restrict__ jobs,
                                                    const UtrDev* __restrict__ utrs,
                                                    const TT* __restrict__ tensor,
                                                    const double* __restrict__ cnt, const ChainDev* __restrict__ chains,
                                                    int32_t* __restrict__ labels) {
  const LabelDev& jb = jobs[blockIdx.x];
  const UtrDev u = utrs[jb.utr];
  const int n = blockIdx.y * blockDim.x + threadIdx.x;
  if (n >= u.N) return;
  // final parameters: inline, or the record of a prune refit that ran just before on this stream
  const ChainDev* src = jb.chain >= 0 ? chains + jb.chain : nullptr;
  const int K = src ? src->K : jb.K;
  const int32_t* a_idx = src ? src->a_idx : jb.a_idx;
  const int32_t* b_idx = src ? src->b_idx : jb.b_idx;
  const double* ws = src ? src->ws : jb.ws;
  const double c = cnt[u.frag_off + n];
  double lzv[SCAPE_B200_KCAP + 1];
  double m = -CUDART_INF;
  for (int j = 0; j <= K; j++) {
    const double w = ws[j];
    const double lw = (w <= 0.0) ? SCAPE_SENTINEL : log(w);
    double val;
    if (j < K)
      val = lw + (double)tensor[u.tensor_off + (int64_t)n * u.ldR + (int64_t)a_idx[j] * u.B + b_idx[j]];
    else
      val = lw + u.unif_loglik;
    lzv[j] = val;
    m = fmax(m, val);
  }
  double s = 0.0;
  for (int j = 0; j <= K; j++) {
    lzv[j] = exp((lzv[j] - m) * c);
    s += lzv[j];
  }
  int best = 0;
  double bz = lzv[0] / s;
  for (int j = 1; j <= K; j++) {
    const double z = lzv[j] / s;
    if (z > bz) { bz = z; best = j; }
  }
  labels[jb.out_off + n] = best;
}

void launch_labels(const LabelDev* jobs, int64_t n_jobs, int max_n, const UtrDev* utrs, const void* tensor, bool f32,
                   const double* cnt, const ChainDev* chains, int32_t* labels, cudaStream_t st) {
  if (n_jobs <= 0) return;
  dim3 grid((unsigned)n_jobs, (unsigned)((max_n + 255) / 256));
  if (f32)
    label_kernel<float><<<grid, 256, 0, st>>>(jobs, utrs, (const float*)tensor, cnt, chains, labels);
  else
    label_kernel<double><<<grid, 256, 0, st>>>(jobs, utrs, (const double*)tensor, cnt, chains, labels);
}

// label_arr = label_arr_of_the_bins[idx_arr] (apa_core.py:976): per read, int64 like the reference's array
__global__ void __launch_bounds__(256) label_expand_kernel(const LabelDev* __restrict__ jobs, const UtrDev* __restrict__ utrs,
                                                           const int32_t* __restrict__ labels,
                                                           const int32_t* __restrict__ read_to_bin,
                                                           int64_t* __restrict__ out) {
  const LabelDev& jb = jobs[blockIdx.x];
  const UtrDev u = utrs[jb.utr];
  for (int r = blockIdx.y * blockDim.x + threadIdx.x; r < u.n_reads; r += gridDim.y * blockDim.x)
    out[u.read_off + r] = (int64_t)labels[jb.out_off + read_to_bin[u.read_off + r]];
}

void launch_label_expand(const LabelDev* jobs, int64_t n_jobs, int max_reads, const UtrDev* utrs, const int32_t* labels,
                         const int32_t* read_to_bin, int64_t* labels_per_read, cudaStream_t st) {
  if (n_jobs <= 0) return;
  dim3 grid((unsigned)n_jobs, (unsigned)std::min(64, (max_reads + 255) / 256));
  label_expand_kernel<<<grid, 256, 0, st>>>(jobs, utrs, labels, read_to_bin, labels_per_read);
}

}  // namespace scape
