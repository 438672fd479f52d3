// See kernels.cuh for the kernel list, the reference lines each kernel follows and the HBM layout.
#include <algorithm>
#include <vector>

#include "kernels.cuh"

namespace scape {

__constant__ ModelConst c_mc;

cudaError_t upload_model_const(const ModelConst& mc) { return cudaMemcpyToSymbol(c_mc, &mc, sizeof(ModelConst)); }

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
      // loglik_xlr_t_r_unknown_kernel (taichi_core.py:141-157)
      const double inv_span = fits ? 1.0 / span : 0.0;
      double acc = 0.0;
      for (int j = 0; j < c_mc.n_s; j++) {
        const double s = c_mc.s_dis[j];
        acc += 1.0 / s * pdf_normal(x, th + s - mu_f, sigma_f) * inv_span * c_mc.pmf_s[j];
      }
      if (acc < 1e-300) acc = 0.0;
      out = (acc <= 0.0) ? SCAPE_SENTINEL : log(acc);
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
// normalised log weights are built once per CTA in shared memory, then every thread owns one
// fragment and does the reference's two-pass log-sum-exp per beta (taichi_core.py:41-54, 172-179).
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
  const int n = blockIdx.y * blockDim.x + tid;
  if (n >= u.Npad) return;
  const double* tab = table + u.table_off + n;
  TT* out = tensor + u.tensor_off + (int64_t)rr.t * B * u.Npad + n;
  const int64_t ld = u.Npad;
  if (n >= u.N) {
    for (int j = 0; j < B; j++) out[(int64_t)j * ld] = TT(0);
    return;
  }
  for (int j = 0; j < B; j++) {
    const int lo = s_lo[j], w = s_w[j];
    const double lps = s_lps[j];
    const double* lp = s_logp + j * max_win;
    const double* col = tab + (int64_t)lo * ld;
    double m = (col[0] + lp[0]) - lps;
    for (int d = 1; d < w; d++) m = fmax(m, (col[(int64_t)d * ld] + lp[d]) - lps);
    double res;
    if (m < -1e30) {
      // every term is the sentinel: exp(0) each, log(w) + sentinel == sentinel in FP64
      res = log((double)w) + m;
    } else {
      double acc = 0.0;
      for (int d = 0; d < w; d++) {
        const double a = ((col[(int64_t)d * ld] + lp[d]) - lps) - m;
        if (a > -746.0) acc += exp(a);   // below that exp() is exactly 0 (sentinel terms: a ~ -3.4e38)
      }
      res = log(acc) + m;
    }
    out[(int64_t)j * ld] = TT(res);   // float storage keeps the sentinel exactly (it IS float's lowest)
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
// K4: EM chain kernel.  One CTA (256 threads) runs one chain from its init blob to convergence:
//   E column refresh (cal_z_k :473-488) -> count-tempered softmax (norm_z :490-495) -> weight
//   update (maximize_ws :498-505, mstep guard :526-529) -> grid arg-max (max_alpha_beta :507-523)
//   -> ELBO (:559-573) -> convergence test (:743); BIC at the end (:702-706).
// Z is never materialised: one pass over the fragments produces every reduction the iteration
// needs plus v[n] = Z[n,k] cnt[n], which the grid search then contracts against the candidate
// rows of the tensor (rows are contiguous: [alpha][beta][n]).
// ------------------------------------------------------------------------------------------------
constexpr int EM_THREADS = 256;
constexpr int EM_WARPS = EM_THREADS / 32;

template <int NV>
__device__ __forceinline__ void block_reduce_sum(double (&val)[NV], double (*s_red)[SCAPE_B200_KCAP + 4],
                                                 double* s_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    double x = val[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[warp][i] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < EM_WARPS; w++) acc += s_red[w][threadIdx.x];
    s_tot[threadIdx.x] = acc;
  }
  __syncthreads();
}

// Candidate scan with LW lanes per tensor row, 16-byte loads (2 doubles or 4 floats).  Only the hull
// [i_lo, i_hi) (in 16-byte units) of the fragments with v != 0 is read: terms with v == 0 contribute
// exactly +-0 to the reference's sum.  Products and sums are FP64 whatever the storage type.
template <typename TT> struct Vec16;
template <> struct Vec16<double> {
  static constexpr int E = 2;
  static __device__ __forceinline__ void dot(const double* row, const double* v, int i, double& a0, double& a1) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(row) + i);
    const double2 w = reinterpret_cast<const double2*>(v)[i];
    a0 = fma(t.x, w.x, a0);
    a1 = fma(t.y, w.y, a1);
  }
};
template <> struct Vec16<float> {
  static constexpr int E = 4;
  static __device__ __forceinline__ void dot(const float* row, const double* v, int i, double& a0, double& a1) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(row) + i);
    const double2 w0 = reinterpret_cast<const double2*>(v)[2 * i];
    const double2 w1 = reinterpret_cast<const double2*>(v)[2 * i + 1];
    a0 = fma((double)t.x, w0.x, a0);
    a1 = fma((double)t.y, w0.y, a1);
    a0 = fma((double)t.z, w1.x, a0);
    a1 = fma((double)t.w, w1.y, a1);
  }
};

template <int LW, typename TT>
__device__ __forceinline__ void grid_scan(const TT* __restrict__ T, const double* __restrict__ v, int npad,
                                          int i_lo, int i_hi, int row0, int row1, double& best_score,
                                          int& best_row) {
  constexpr int G = EM_THREADS / LW;
  const int g = threadIdx.x / LW, lg = threadIdx.x % LW;
  for (int rb = row0; rb < row1; rb += G) {
    const int r = rb + g;
    double acc0 = 0.0, acc1 = 0.0;
    if (r < row1) {
      const TT* row = T + (int64_t)r * npad;
#pragma unroll 4
      for (int i = i_lo + lg; i < i_hi; i += LW) Vec16<TT>::dot(row, v, i, acc0, acc1);
    }
    double acc = acc0 + acc1;
#pragma unroll
    for (int o = LW / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lg == 0 && r < row1 && acc > best_score) {
      best_score = acc;
      best_row = r;
    }
  }
}

struct EmShared {
  double w[SCAPE_B200_KCAP + 1], lw[SCAPE_B200_KCAP + 1];
  int a[SCAPE_B200_KCAP], b[SCAPE_B200_KCAP];
  double red[EM_WARPS][SCAPE_B200_KCAP + 4];
  double tot[SCAPE_B200_KCAP + 4];
  double bscore[EM_THREADS / 4];
  int brow[EM_THREADS / 4];
  int ctl[2];    // [1] converged
  int hull[2];   // first / last fragment with v != 0
};

template <int NK, typename TT>  // NK = K + 1 columns, TT = tensor storage type
__device__ void em_chain_run(EmShared& sh, ChainDev& ch, const UtrDev& u, const TT* __restrict__ T,
                             const double* __restrict__ cnt, double* __restrict__ lz, double* __restrict__ v,
                             int32_t* trace_a, int32_t* trace_b, double* trace_ws) {
  constexpr int K = NK - 1;
  double (&s_w)[SCAPE_B200_KCAP + 1] = sh.w;
  double (&s_lw)[SCAPE_B200_KCAP + 1] = sh.lw;
  int (&s_a)[SCAPE_B200_KCAP] = sh.a;
  int (&s_b)[SCAPE_B200_KCAP] = sh.b;
  double (*s_red)[SCAPE_B200_KCAP + 4] = sh.red;
  double* s_tot = sh.tot;
  double* s_bscore = sh.bscore;
  int* s_brow = sh.brow;
  int* s_ctl = sh.ctl;
  int* s_hull = sh.hull;

  const int tid = threadIdx.x;
  const int N = u.N, npad = u.Npad, B = u.B;
  const double cap = c_mc.max_unif_ws;

  if (tid < NK) {
    const double w = ch.ws[tid];
    s_w[tid] = w;
    s_lw[tid] = (w <= 0.0) ? SCAPE_SENTINEL : log(w);
    if (tid < K) {
      s_a[tid] = ch.a_idx[tid];
      s_b[tid] = ch.b_idx[tid];
    }
  }
  for (int n = tid + N; n < npad; n += EM_THREADS) v[n] = 0.0;  // padding never contributes
  __syncthreads();
  // initial log_zmat: all K+1 columns (em_algo :722-724)
  for (int j = 0; j < NK; j++) {
    const double lw = s_lw[j];
    if (j < K) {
      const TT* row = T + ((int64_t)s_a[j] * B + s_b[j]) * npad;
      for (int n = tid; n < N; n += EM_THREADS) lz[(int64_t)j * npad + n] = lw + (double)row[n];
    } else {
      const double val = lw + u.unif_loglik;
      for (int n = tid; n < N; n += EM_THREADS) lz[(int64_t)j * npad + n] = val;
    }
  }
  __syncthreads();

  double lb = SCAPE_SENTINEL;  // meaningful in thread 0 only
  double last_A = 0.0;
  double grid_rows = 0.0, grid_elems = 0.0;
  int n_iter = 0;

  for (int it = 0; it < SCAPE_B200_NROUND; it++) {
    const int k = ch.k_order[it];
    bool guard = false;
    double red[NK + 3];
    while (true) {
      const double lwk = s_lw[k];
      const TT* trow = T + ((int64_t)s_a[k] * B + s_b[k]) * npad;
#pragma unroll
      for (int j = 0; j < NK + 3; j++) red[j] = 0.0;
      int h_lo = N, h_hi = -1;
      if (tid == 0) { s_hull[0] = N; s_hull[1] = -1; }
      for (int n = tid; n < N; n += EM_THREADS) {
        const double c = cnt[n];
        const double fresh = lwk + (double)trow[n];
        double z[NK], lzv[NK];
        double m = -CUDART_INF;
#pragma unroll
        for (int j = 0; j < NK; j++) {
          lzv[j] = (j == k) ? fresh : lz[(int64_t)j * npad + n];
          m = fmax(m, lzv[j]);
        }
        lz[(int64_t)k * npad + n] = fresh;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NK; j++) {
          z[j] = exp((lzv[j] - m) * c);
          s += z[j];
        }
        double zk = 0.0;
#pragma unroll
        for (int j = 0; j < NK; j++) {
          z[j] = z[j] / s;
          if (j == k) zk = z[j];
        }
        red[NK] += zk;                         // np.sum(Z[:, k]) before the guard
        if (guard) {
          zk += 1e-8;
#pragma unroll
          for (int j = 0; j < NK; j++)
            if (j == k) z[j] = zk;
        }
        double ps = 0.0, A = 0.0;
#pragma unroll
        for (int j = 0; j < NK; j++) {
          red[j] = fma(c, z[j], red[j]);       // cnt @ Z
          if (z[j] != 0.0) A += (z[j] * c) * lzv[j];
          ps += z[j];
        }
        double h = 0.0;                        // scipy.stats.entropy(Z[n, :])
#pragma unroll
        for (int j = 0; j < NK; j++) {
          const double p = z[j] / ps;
          if (p > 0.0) h -= p * log(p);
        }
        red[NK + 1] += A;
        red[NK + 2] = fma(c, h, red[NK + 2]);
        const double vn = zk * c;
        v[n] = vn;
        if (vn != 0.0) { h_lo = min(h_lo, n); h_hi = n; }
      }
      __syncthreads();
      if (h_hi >= 0) { atomicMin(&s_hull[0], h_lo); atomicMax(&s_hull[1], h_hi); }
      block_reduce_sum<NK + 3>(red, s_red, s_tot);
      if (!guard && s_tot[NK] < 1e-8) {        // mstep guard (:526-529); uniform across the CTA
        guard = true;
        __syncthreads();
        continue;
      }
      break;
    }
    if (tid == 0) {
      // maximize_ws (:498-505)
      double w[NK];
      double tot = 0.0;
#pragma unroll
      for (int j = 0; j < NK; j++) tot += s_tot[j];
#pragma unroll
      for (int j = 0; j < NK; j++) w[j] = s_tot[j] / tot;
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
        s_w[j] = w[j];
        s_lw[j] = (w[j] <= 0.0) ? SCAPE_SENTINEL : log(w[j]);
      }
    }
    __syncthreads();
    if (!ch.weights_only) {
      // max_alpha_beta (:507-523): rows [lo*B, (hi+1)*B) of the tensor are one contiguous slab
      const int lo = (k == 0) ? 0 : s_a[k - 1];
      const int hi = (k == K - 1) ? u.T - 1 : s_a[k + 1];
      const int row0 = lo * B, row1 = (hi + 1) * B;
      double bscore = -CUDART_INF;
      int brow = row0;
      int groups;
      constexpr int VE = Vec16<TT>::E;   // elements per 16-byte load
      const int i_lo = s_hull[0] / VE, i_hi = s_hull[1] < 0 ? 0 : s_hull[1] / VE + 1;   // empty hull -> no loads, all scores 0
      const int span = i_hi - i_lo;
      if (span <= 32) {
        grid_scan<4, TT>(T, v, npad, i_lo, i_hi, row0, row1, bscore, brow);
        groups = EM_THREADS / 4;
        if ((tid & 3) == 0) { s_bscore[tid >> 2] = bscore; s_brow[tid >> 2] = brow; }
      } else if (span <= 256) {
        grid_scan<8, TT>(T, v, npad, i_lo, i_hi, row0, row1, bscore, brow);
        groups = EM_THREADS / 8;
        if ((tid & 7) == 0) { s_bscore[tid >> 3] = bscore; s_brow[tid >> 3] = brow; }
      } else {
        grid_scan<32, TT>(T, v, npad, i_lo, i_hi, row0, row1, bscore, brow);
        groups = EM_THREADS / 32;
        if ((tid & 31) == 0) { s_bscore[tid >> 5] = bscore; s_brow[tid >> 5] = brow; }
      }
      __syncthreads();
      if (tid == 0) {
        double best = s_bscore[0];
        int row = s_brow[0];
        for (int g = 1; g < groups; g++) {
          if (s_bscore[g] > best || (s_bscore[g] == best && s_brow[g] < row)) {
            best = s_bscore[g];
            row = s_brow[g];
          }
        }
        s_a[k] = row / B;
        s_b[k] = row % B;
        grid_rows += (double)(row1 - row0);
        grid_elems += (double)(row1 - row0) * (double)(VE * max(span, 0));
      }
    }
    if (tid == 0) {
      const double lb_new = s_tot[NK + 1] + s_tot[NK + 2];   // elbo (:559-561)
      last_A = s_tot[NK + 1];
      ch.lb_arr[it] = lb_new;
      n_iter = it + 1;
      if (ch.trace_off >= 0) {
        const int64_t o = ch.trace_off + (int64_t)it * (SCAPE_B200_KCAP + 1);
        for (int j = 0; j < K; j++) { trace_a[o + j] = s_a[j]; trace_b[o + j] = s_b[j]; }
        for (int j = 0; j < NK; j++) trace_ws[o + j] = s_w[j];
      }
      const bool conv = fabs(lb_new - lb) < fabs(1e-6 * lb);   // (:743)
      s_ctl[1] = conv ? 1 : 0;
      if (!conv) lb = lb_new;
    }
    __syncthreads();
    if (s_ctl[1]) break;
  }
  if (tid == 0) {
    ch.n_iter = n_iter;
    ch.bic = -2.0 * last_A + (3 * K + 1) * log((double)N);      // cal_bic (:702-706)
    ch.grid_rows = grid_rows;
    ch.grid_elems = grid_elems;
    for (int j = 0; j < K; j++) { ch.a_idx[j] = s_a[j]; ch.b_idx[j] = s_b[j]; }
    for (int j = 0; j < NK; j++) ch.ws[j] = s_w[j];
  }
}

constexpr int EM_MIN_BLOCKS = 3;
constexpr int EM_MULTI_KMAX = 7;

// K >= 8 (only reachable through re-runs): one instantiation per K.
template <int NK, typename TT>
__global__ void __launch_bounds__(EM_THREADS, 2) em_chain_kernel(ChainDev* chains, const int32_t* __restrict__ order,
                                                                 const UtrDev* __restrict__ utrs,
                                                                 const void* __restrict__ tensor,
                                                                 const double* __restrict__ cnt, double* lz_all,
                                                                 double* v_all, int smem_doubles, int32_t* trace_a,
                                                                 int32_t* trace_b, double* trace_ws) {
  extern __shared__ double sm_v[];
  __shared__ EmShared sh;
  ChainDev& ch = chains[order[blockIdx.x]];
  const UtrDev u = utrs[ch.utr];
  double* v = (u.Npad <= smem_doubles) ? sm_v : (v_all + ch.v_off);
  em_chain_run<NK, TT>(sh, ch, u, (const TT*)tensor + u.tensor_off, cnt + u.frag_off, lz_all + ch.lz_off, v, trace_a,
                       trace_b, trace_ws);
}

// K = 1..7 in ONE launch: CTAs are ordered (UTR, K, restart), so the ~50 chains that share a UTR's
// tensor are resident together and the tensor stays in L2 while they scan it.
template <typename TT>
__global__ void __launch_bounds__(EM_THREADS, EM_MIN_BLOCKS)
em_chain_kernel_multi(ChainDev* chains, const int32_t* __restrict__ order, const UtrDev* __restrict__ utrs,
                      const void* __restrict__ tensor, const double* __restrict__ cnt, double* lz_all,
                      double* v_all, int smem_doubles, int32_t* trace_a, int32_t* trace_b, double* trace_ws) {
  extern __shared__ double sm_v[];
  __shared__ EmShared sh;
  ChainDev& ch = chains[order[blockIdx.x]];
  const UtrDev u = utrs[ch.utr];
  double* v = (u.Npad <= smem_doubles) ? sm_v : (v_all + ch.v_off);
  const TT* T = (const TT*)tensor + u.tensor_off;
  const double* c = cnt + u.frag_off;
  double* lz = lz_all + ch.lz_off;
  switch (ch.K) {
    case 1: em_chain_run<2, TT>(sh, ch, u, T, c, lz, v, trace_a, trace_b, trace_ws); break;
    case 2: em_chain_run<3, TT>(sh, ch, u, T, c, lz, v, trace_a, trace_b, trace_ws); break;
    case 3: em_chain_run<4, TT>(sh, ch, u, T, c, lz, v, trace_a, trace_b, trace_ws); break;
    case 4: em_chain_run<5, TT>(sh, ch, u, T, c, lz, v, trace_a, trace_b, trace_ws); break;
    case 5: em_chain_run<6, TT>(sh, ch, u, T, c, lz, v, trace_a, trace_b, trace_ws); break;
    case 6: em_chain_run<7, TT>(sh, ch, u, T, c, lz, v, trace_a, trace_b, trace_ws); break;
    case 7: em_chain_run<8, TT>(sh, ch, u, T, c, lz, v, trace_a, trace_b, trace_ws); break;
    default: break;
  }
}

typedef void (*em_kernel_t)(ChainDev*, const int32_t*, const UtrDev*, const void*, const double*, double*, double*,
                            int, int32_t*, int32_t*, double*);

template <typename TT>
static em_kernel_t em_kernel_for(int K) {
  switch (K) {
    case 0: return em_chain_kernel_multi<TT>;
    case 8: return em_chain_kernel<9, TT>;
    case 9: return em_chain_kernel<10, TT>;
    case 10: return em_chain_kernel<11, TT>;
    case 11: return em_chain_kernel<12, TT>;
    case 12: return em_chain_kernel<13, TT>;
    case 13: return em_chain_kernel<14, TT>;
    case 14: return em_chain_kernel<15, TT>;
    case 15: return em_chain_kernel<16, TT>;
  }
  return nullptr;
}

// Host-side launch plan: chains are grouped by (K, small/large fragment count); `order_dev` must
// hold n_chains int32 and is filled here through `order_host` (pinned or pageable).
int launch_em_groups(ChainDev* chains_dev, const ChainDev* chains_host, int64_t n_chains, const UtrDev* utrs_host,
                     const UtrDev* utrs_dev, const void* tensor, bool f32, const double* cnt, double* lz,
                     double* vbuf, int32_t* order_dev, int32_t* order_host, int32_t* trace_a, int32_t* trace_b,
                     double* trace_ws, cudaStream_t st) {
  constexpr int SMALL = 1024, LARGE_CAP = 24576;
  std::vector<int32_t> buckets[SCAPE_B200_KCAP + 1][2];
  int big_max[SCAPE_B200_KCAP + 1] = {0};
  for (int64_t i = 0; i < n_chains; i++) {
    const ChainDev& c = chains_host[i];
    const int npad = utrs_host[c.utr].Npad;
    const int cls = npad <= SMALL ? 0 : 1;
    const int kb = c.K <= EM_MULTI_KMAX ? 0 : c.K;       // bucket 0 = the multi-K kernel
    buckets[kb][cls].push_back((int32_t)i);
    if (cls) big_max[kb] = std::max(big_max[kb], npad);
  }
  int64_t pos = 0;
  int launches = 0;
  struct Plan { int K, cls; int64_t off, n; };
  std::vector<Plan> plans;
  for (int K = 0; K <= SCAPE_B200_KCAP; K++)
    for (int cls = 0; cls < 2; cls++) {
      auto& b = buckets[K][cls];
      if (b.empty()) continue;
      std::copy(b.begin(), b.end(), order_host + pos);
      plans.push_back({K, cls, pos, (int64_t)b.size()});
      pos += (int64_t)b.size();
    }
  cudaMemcpyAsync(order_dev, order_host, sizeof(int32_t) * (size_t)n_chains, cudaMemcpyHostToDevice, st);
  for (const Plan& p : plans) {
    em_kernel_t kern = f32 ? em_kernel_for<float>(p.K) : em_kernel_for<double>(p.K);
    int smem_doubles = p.cls == 0 ? SMALL : std::min(big_max[p.K], LARGE_CAP);
    size_t smem = (size_t)smem_doubles * sizeof(double);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)p.n, EM_THREADS, smem, st>>>(chains_dev, order_dev + p.off, utrs_dev, tensor, cnt, lz, vbuf,
                                                  smem_doubles, trace_a, trace_b, trace_ws);
    launches++;
  }
  return launches;
}

// ------------------------------------------------------------------------------------------------
// K5: labels.  get_label (:873-881): refresh all columns with the final parameters, tempered
// softmax, first arg-max per fragment.
// ------------------------------------------------------------------------------------------------
template <typename TT>
__global__ void __launch_bounds__(256) label_kernel(const LabelDev* __restrict__ jobs,
                                                    const UtrDev* __restrict__ utrs,
                                                    const TT* __restrict__ tensor,
                                                    const double* __restrict__ cnt, int32_t* __restrict__ labels) {
  const LabelDev& jb = jobs[blockIdx.x];
  const UtrDev u = utrs[jb.utr];
  const int n = blockIdx.y * blockDim.x + threadIdx.x;
  if (n >= u.N) return;
  const int K = jb.K;
  const double c = cnt[u.frag_off + n];
  double lzv[SCAPE_B200_KCAP + 1];
  double m = -CUDART_INF;
  for (int j = 0; j <= K; j++) {
    const double w = jb.ws[j];
    const double lw = (w <= 0.0) ? SCAPE_SENTINEL : log(w);
    double val;
    if (j < K)
      val = lw + (double)tensor[u.tensor_off + ((int64_t)jb.a_idx[j] * u.B + jb.b_idx[j]) * u.Npad + n];
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
                   const double* cnt, int32_t* labels, cudaStream_t st) {
  if (n_jobs <= 0) return;
  dim3 grid((unsigned)n_jobs, (unsigned)((max_n + 255) / 256));
  if (f32)
    label_kernel<float><<<grid, 256, 0, st>>>(jobs, utrs, (const float*)tensor, cnt, labels);
  else
    label_kernel<double><<<grid, 256, 0, st>>>(jobs, utrs, (const double*)tensor, cnt, labels);
}

}  // namespace scape
