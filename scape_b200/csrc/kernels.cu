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
  // tensor layout [n][R], R = T*B candidate rows (alpha-major, beta-minor) contiguous per fragment
  TT* out = tensor + u.tensor_off + (int64_t)n * u.ldR + (int64_t)rr.t * B;
  const int64_t ld = u.Npad;
  if (n >= u.N) return;   // the tensor has no padding fragments
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
    out[j] = TT(res);   // float storage keeps the sentinel exactly (it IS float's lowest)
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
//   em_estep_kernel   one CTA per chain.  (0) applies the arg-max the previous scan found for this
//                     chain (alpha_k, beta_k update, trace, finalisation of converged chains), then
//                     (1) column refresh (cal_z_k :473-488), count-tempered softmax (norm_z :490-495),
//                     weight update (maximize_ws :498-505, mstep guard :526-529), ELBO (:559-573) and
//                     the convergence test (:743).  Z is never materialised; the pass leaves
//                     v[n] = Z[n,k] cnt[n] in the chain's row of V.
//   em_scan_kernel    max_alpha_beta (:507-523) as a blocked product: one CTA per (UTR, block of 512
//                     candidate rows) computes scores[row][chain] = sum_n tensor[n][row] * V[chain][n]
//                     for EVERY running chain of that UTR whose window touches the block (all K, all
//                     restarts), so a tensor block is fetched once per step however many chains
//                     need it, and the step's work is spread over all SMs (no long-tailed CTAs).
//                     Tensor rows stream through a TMA (1-D bulk copy) + mbarrier ring; thread <->
//                     2 rows; V is staged in shared memory per sub-batch of <= 8 chains; FP64 FMAs.
//                     Each (chain, block) leaves its first-maximum (score, row) in a partials array.
// Only the hull of fragments with v != 0 is visited (other terms are exactly +-0 in the reference's
// sum).  BIC at the end (:702-706).
// ------------------------------------------------------------------------------------------------
constexpr int GT = 256;                    // threads per CTA
constexpr int GW = GT / 32;                // warps
constexpr int SCAN_ROWS = 2 * GT;          // candidate rows per block (2 per thread)
constexpr int SCAN_GB = 8;                 // chains per register sub-batch
constexpr int SCAN_MAXCH = 96;             // running chains of one UTR a scan CTA can list
constexpr int SCAN_VCHUNK = 512;           // fragments of V staged per chunk (8 x 512 doubles = 32 KB)
constexpr int RING_STAGES = 3;             // TMA ring: stages in flight
constexpr int RING_CH = 8;                 // fragments (tensor n-rows) per stage
constexpr int RING_PITCH = SCAN_ROWS + 8;  // elements per staged n-row (up to 3 + 3 elements of 16-byte alignment slack)

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

// ---- TMA (1-D bulk async copy) + mbarrier, raw PTX ------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// E step
// ------------------------------------------------------------------------------------------------
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

template <int NK, typename TT>
__device__ void estep_run(EShared& sh, ChainDev& ch, const UtrDev& u, const TT* __restrict__ A,
                          const double* __restrict__ cnt, double* __restrict__ lz, double* __restrict__ V) {
  constexpr int K = NK - 1;
  const int tid = threadIdx.x;
  const int N = u.N, npad = u.Npad, B = u.B;
  const int64_t R = u.ldR;
  const double cap = c_mc.max_unif_ws;
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
    for (int n = tid; n < N; n += GT) {
      const double c = cnt[n];
      const double fresh = lwk + (double)A[(int64_t)n * R + rk];
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
      double ps = 0.0, Aterm = 0.0;
#pragma unroll
      for (int j = 0; j < NK; j++) {
        red[j] = fma(c, z[j], red[j]);       // cnt @ Z
        if (z[j] != 0.0) Aterm += (z[j] * c) * lzv[j];
        ps += z[j];
      }
      double h = 0.0;                        // scipy.stats.entropy(Z[n, :])
#pragma unroll
      for (int j = 0; j < NK; j++) {
        const double p = z[j] / ps;
        if (p > 0.0) h -= p * log(p);
      }
      red[NK + 1] += Aterm;
      red[NK + 2] = fma(c, h, red[NK + 2]);
      const double vn = zk * c;
      V[n] = vn;
      if (vn != 0.0) { h_lo = min(h_lo, n); h_hi = n; }
    }
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
  if (tid == 0) {
    // maximize_ws (:498-505)
    double w[NK];
    double tot = 0.0;
#pragma unroll
    for (int j = 0; j < NK; j++) tot += sh.tot[j];
#pragma unroll
    for (int j = 0; j < NK; j++) w[j] = sh.tot[j] / tot;
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
    const double lb_new = sh.tot[NK + 1] + sh.tot[NK + 2];      // elbo (:559-561)
    ch.last_a = sh.tot[NK + 1];
    ch.lb_arr[it] = lb_new;
    ch.n_iter = it + 1;
    const double lb = ch.lb_prev;
    const bool conv = fabs(lb_new - lb) < fabs(1e-6 * lb);      // (:743)
    if (!conv) ch.lb_prev = lb_new;
    const bool last = conv || it == SCAPE_B200_NROUND - 1;
    ch.cur_k = k;
    if (ch.weights_only) {                                      // mstep_fixed (:552-557): no grid search
      if (last) finalize_chain(ch, N); else ch.state = 1;
    } else {
      // max_alpha_beta (:507-523): candidate window of component k
      const int lo = (k == 0) ? 0 : ch.a_idx[k - 1];
      const int hi = (k == K - 1) ? u.T - 1 : ch.a_idx[k + 1];
      ch.row0 = lo * B;
      ch.row1 = (hi + 1) * B;
      ch.hlo = sh.hull[0];
      ch.hhi = sh.hull[1];
      ch.grid_rows += (double)(ch.row1 - ch.row0);
      ch.pending = 1;
      ch.state = last ? 2 : 1;
    }
  }
}

struct ScanPartial {
  double score;
  int row;
  int pad;
};

// one CTA per chain; `chains` are the wave's chains in launch order
template <typename TT>
__global__ void __launch_bounds__(GT, 2)
em_estep_kernel(ChainDev* chains, const UtrDev* __restrict__ utrs, const void* __restrict__ tensor,
                const double* __restrict__ cnt_all, double* lz_all, double* v_all,
                const ScanPartial* __restrict__ partials, int32_t* trace_a, int32_t* trace_b, double* trace_ws) {
  __shared__ EShared sh;
  ChainDev& ch = chains[blockIdx.x];
  if (ch.state == 0) return;
  const UtrDev u = utrs[ch.utr];
  const int tid = threadIdx.x;
  // ---- (0) apply the arg-max of the previous step's scan: first maximum in row order ----
  if (tid < 32) {
    int go = 1;
    if (ch.pending) {
      const int b0 = ch.row0 / SCAN_ROWS, b1 = (ch.row1 - 1) / SCAN_ROWS;
      double best = -CUDART_INF;
      int row = 0x7fffffff;
      for (int b = b0 + tid; b <= b1; b += 32) {
        const ScanPartial p = partials[ch.pb_off + b];
        if (p.score > best || (p.score == best && p.row < row)) { best = p.score; row = p.row; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int orow = __shfl_xor_sync(0xffffffffu, row, o);
        if (ob > best || (ob == best && orow < row)) { best = ob; row = orow; }
      }
      if (tid == 0) {
        ch.a_idx[ch.cur_k] = row / u.B;
        ch.b_idx[ch.cur_k] = row % u.B;
        ch.pending = 0;
        if (ch.trace_off >= 0) ch.trace_pending = ch.n_iter;
      }
    }
    __syncwarp();
    if (tid == 0) {
      if (ch.trace_off >= 0 && ch.trace_pending > 0) {
        const int64_t o = ch.trace_off + (int64_t)(ch.trace_pending - 1) * (SCAPE_B200_KCAP + 1);
        for (int j = 0; j < ch.K; j++) { trace_a[o + j] = ch.a_idx[j]; trace_b[o + j] = ch.b_idx[j]; }
        for (int j = 0; j <= ch.K; j++) trace_ws[o + j] = ch.ws[j];
        ch.trace_pending = 0;
      }
      if (ch.state == 2) { finalize_chain(ch, u.N); go = 0; }
      if (ch.n_iter >= SCAPE_B200_NROUND) go = 0;
      sh.go = go;
    }
  }
  __syncthreads();
  if (!sh.go) return;
  const TT* A = (const TT*)tensor + u.tensor_off;
  const double* cnt = cnt_all + u.frag_off;
  double* lz = lz_all + ch.lz_off;
  double* V = v_all + ch.v_off;
  switch (ch.K) {
    case 1: estep_run<2, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 2: estep_run<3, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 3: estep_run<4, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 4: estep_run<5, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 5: estep_run<6, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 6: estep_run<7, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 7: estep_run<8, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 8: estep_run<9, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 9: estep_run<10, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 10: estep_run<11, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 11: estep_run<12, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 12: estep_run<13, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 13: estep_run<14, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 14: estep_run<15, TT>(sh, ch, u, A, cnt, lz, V); break;
    case 15: estep_run<16, TT>(sh, ch, u, A, cnt, lz, V); break;
    default: break;
  }
  if (tid == 0 && ch.weights_only && ch.trace_off >= 0) {     // weights-only chains never wait for a scan
    const int64_t o = ch.trace_off + (int64_t)(ch.n_iter - 1) * (SCAPE_B200_KCAP + 1);
    for (int j = 0; j < ch.K; j++) { trace_a[o + j] = ch.a_idx[j]; trace_b[o + j] = ch.b_idx[j]; }
    for (int j = 0; j <= ch.K; j++) trace_ws[o + j] = ch.ws[j];
  }
}

// ------------------------------------------------------------------------------------------------
// scan
// ------------------------------------------------------------------------------------------------
struct ScanShared {
  int list[SCAN_MAXCH];          // chain indices (into the wave's chain array) that need this block
  int n_list, N0, N1;
  unsigned long long full_bar[RING_STAGES];
  double wbest[GW][SCAN_GB];
  int wrow[GW][SCAN_GB];
};

template <int GB, typename TT>
__device__ __forceinline__ void scan_subbatch(ScanShared& sh, ChainDev* chains, const UtrDev& u,
                                              const TT* __restrict__ A, const double* __restrict__ v_all,
                                              ScanPartial* partials, int first, int cnt, int blk, double* Vs,
                                              TT* ring, uint32_t& ring_it, double* scan_elems) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t R = u.ldR;
  const int Rv = u.T * u.B;                     // valid candidate rows
  const int base = blk * SCAN_ROWS;
  const int blk_end = min(base + SCAN_ROWS, Rv);
  // hull of this sub-batch
  int N0 = 1 << 30, N1 = 0;
  int w0[GB], w1[GB];
  const double* vrow[GB];
#pragma unroll
  for (int j = 0; j < GB; j++) {
    const ChainDev& c = chains[sh.list[first + (j < cnt ? j : 0)]];
    w0[j] = c.row0;
    w1[j] = c.row1;
    vrow[j] = v_all + c.v_off;
    if (j < cnt && c.hhi >= 0) { N0 = min(N0, c.hlo); N1 = max(N1, c.hhi + 1); }
  }
  if (N1 <= N0) { N0 = 0; N1 = 0; }             // every v is zero: all scores 0, first row of each window wins
  N0 &= ~7;
  if (threadIdx.x == 0 && scan_elems) atomicAdd(scan_elems, (double)(min(base + SCAN_ROWS, Rv) - base) * (double)(N1 - N0));                                     // aligned start (V is exactly 0 outside a chain's hull)
  const int r0 = base + tid, r1 = r0 + GT;
  const int base_al = base & ~3;
  const int end_al = min((blk_end + 3) & ~3, (int)R);
  const uint32_t row_bytes = (uint32_t)(end_al - base_al) * (uint32_t)sizeof(TT);
  const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t bar_base = (uint32_t)__cvta_generic_to_shared(sh.full_bar);
  constexpr uint32_t STAGE_BYTES = RING_CH * RING_PITCH * sizeof(TT);
  const uint32_t off0 = (uint32_t)(r0 - base_al) * (uint32_t)sizeof(TT);
  const uint32_t off1 = off0 + GT * (uint32_t)sizeof(TT);
  const uint32_t vs_base = (uint32_t)__cvta_generic_to_shared(Vs);

  double acc0[GB], acc1[GB];
#pragma unroll
  for (int j = 0; j < GB; j++) acc0[j] = acc1[j] = 0.0;
  for (int c0 = N0; c0 < N1; c0 += SCAN_VCHUNK) {
    const int c1 = min(c0 + SCAN_VCHUNK, N1);
    __syncthreads();                            // previous chunk / sub-batch fully consumed
    for (int e = tid; e < GB * (c1 - c0); e += GT) {
      const int j = e / (c1 - c0), o = e % (c1 - c0);
      Vs[j * SCAN_VCHUNK + o] = (j < cnt) ? vrow[j < cnt ? j : 0][c0 + o] : 0.0;
    }
    const int n_it = (c1 - c0 + RING_CH - 1) / RING_CH;
    if (tid == 0) {                             // producer prologue: fill the ring
      for (int p = 0; p < min(n_it, RING_STAGES); p++) {
        const uint32_t st = (ring_it + p) % RING_STAGES;
        const int nb = c0 + p * RING_CH, ne = min(nb + RING_CH, c1);
        const uint32_t bar = bar_base + st * 8u;
        mbar_expect_tx(bar, row_bytes * (uint32_t)(ne - nb));
        for (int n = nb; n < ne; n++)
          tma_load_1d(ring_base + st * STAGE_BYTES + (uint32_t)(n - nb) * RING_PITCH * (uint32_t)sizeof(TT),
                      A + (int64_t)n * R + base_al, row_bytes, bar);
      }
    }
    __syncthreads();
    uint32_t vj = vs_base;
    for (int it = 0; it < n_it; it++) {
      const uint32_t st = ring_it % RING_STAGES, parity = (ring_it / RING_STAGES) & 1u;
      mbar_wait(bar_base + st * 8u, parity);
      const uint32_t sa = ring_base + st * STAGE_BYTES;
      const int nb = c0 + it * RING_CH, ne = min(nb + RING_CH, c1);
      if (ne - nb == RING_CH) {
#pragma unroll
        for (int i = 0; i < RING_CH; i += 2) {
          const double a0 = lds_elem<TT>(sa + (uint32_t)i * RING_PITCH * (uint32_t)sizeof(TT) + off0);
          const double b0 = lds_elem<TT>(sa + (uint32_t)i * RING_PITCH * (uint32_t)sizeof(TT) + off1);
          const double a1 = lds_elem<TT>(sa + (uint32_t)(i + 1) * RING_PITCH * (uint32_t)sizeof(TT) + off0);
          const double b1 = lds_elem<TT>(sa + (uint32_t)(i + 1) * RING_PITCH * (uint32_t)sizeof(TT) + off1);
#pragma unroll
          for (int j = 0; j < GB; j++) {
            const double2 v = lds_f64x2(vj + (uint32_t)(j * SCAN_VCHUNK + i) * 8u);
            acc0[j] = fma(a0, v.x, acc0[j]);
            acc1[j] = fma(b0, v.x, acc1[j]);
            acc0[j] = fma(a1, v.y, acc0[j]);
            acc1[j] = fma(b1, v.y, acc1[j]);
          }
        }
      } else {
        for (int i = 0; i < ne - nb; i++) {
          const double a0 = lds_elem<TT>(sa + (uint32_t)i * RING_PITCH * (uint32_t)sizeof(TT) + off0);
          const double b0 = lds_elem<TT>(sa + (uint32_t)i * RING_PITCH * (uint32_t)sizeof(TT) + off1);
#pragma unroll
          for (int j = 0; j < GB; j++) {
            const double v = lds_f64(vj + (uint32_t)(j * SCAN_VCHUNK + i) * 8u);
            acc0[j] = fma(a0, v, acc0[j]);
            acc1[j] = fma(b0, v, acc1[j]);
          }
        }
      }
      vj += RING_CH * 8u;
      __syncthreads();                          // every thread is done with this stage
      if (tid == 0 && it + RING_STAGES < n_it) {
        const int fb = c0 + (it + RING_STAGES) * RING_CH, fe = min(fb + RING_CH, c1);
        const uint32_t bar = bar_base + st * 8u;
        mbar_expect_tx(bar, row_bytes * (uint32_t)(fe - fb));
        for (int n = fb; n < fe; n++)
          tma_load_1d(sa + (uint32_t)(n - fb) * RING_PITCH * (uint32_t)sizeof(TT), A + (int64_t)n * R + base_al,
                      row_bytes, bar);
      }
      ring_it++;
    }
  }
  // first maximum of this block per chain: larger score wins, ties go to the smaller row
#pragma unroll
  for (int j = 0; j < GB; j++) {
    double b = -CUDART_INF;
    int r = 0x7fffffff;
    if (r0 >= w0[j] && r0 < w1[j] && r0 < blk_end) { b = acc0[j]; r = r0; }
    if (r1 >= w0[j] && r1 < w1[j] && r1 < blk_end && acc1[j] > b) { b = acc1[j]; r = r1; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, b, o);
      const int orow = __shfl_xor_sync(0xffffffffu, r, o);
      if (ob > b || (ob == b && orow < r)) { b = ob; r = orow; }
    }
    if (lane == 0) { sh.wbest[warp][j] = b; sh.wrow[warp][j] = r; }
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
    partials[chains[sh.list[first + tid]].pb_off + blk] = p;
  }
}

// one CTA per (UTR, block of SCAN_ROWS candidate rows)
template <typename TT>
__global__ void __launch_bounds__(GT, 2)
em_scan_kernel(const ScanRef* __restrict__ refs, ChainDev* chains, const UtrDev* __restrict__ utrs,
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
    const ChainDev& ch = chains[c];
    if (ch.pending && ch.row0 < hi && ch.row1 > lo) {
      const int slot = atomicAdd(&sh.n_list, 1);
      if (slot < SCAN_MAXCH) sh.list[slot] = c;
    }
  }
  __syncthreads();
  const int n_list = min(sh.n_list, SCAN_MAXCH);
  if (n_list == 0) return;
  // deterministic sub-batches: order the list by chain index (tiny insertion sort by one thread)
  if (tid == 0) {
    for (int i = 1; i < n_list; i++) {
      const int v = sh.list[i];
      int j = i - 1;
      while (j >= 0 && sh.list[j] > v) { sh.list[j + 1] = sh.list[j]; j--; }
      sh.list[j + 1] = v;
    }
    for (int st = 0; st < RING_STAGES; st++) mbar_init((uint32_t)__cvta_generic_to_shared(&sh.full_bar[st]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  double* Vs = sm_dyn;                                           // [SCAN_GB][SCAN_VCHUNK]
  TT* ring = reinterpret_cast<TT*>(sm_dyn + SCAN_GB * SCAN_VCHUNK);
  const TT* A = (const TT*)tensor + u.tensor_off;
  uint32_t ring_it = 0;
  for (int first = 0; first < n_list; first += SCAN_GB) {
    const int cnt = min(SCAN_GB, n_list - first);
    if (cnt <= 1) scan_subbatch<1, TT>(sh, chains, u, A, v_all, partials, first, cnt, ref.blk, Vs, ring, ring_it, scan_elems);
    else if (cnt <= 2) scan_subbatch<2, TT>(sh, chains, u, A, v_all, partials, first, cnt, ref.blk, Vs, ring, ring_it, scan_elems);
    else if (cnt <= 3) scan_subbatch<3, TT>(sh, chains, u, A, v_all, partials, first, cnt, ref.blk, Vs, ring, ring_it, scan_elems);
    else if (cnt <= 4) scan_subbatch<4, TT>(sh, chains, u, A, v_all, partials, first, cnt, ref.blk, Vs, ring, ring_it, scan_elems);
    else if (cnt <= 6) scan_subbatch<6, TT>(sh, chains, u, A, v_all, partials, first, cnt, ref.blk, Vs, ring, ring_it, scan_elems);
    else scan_subbatch<8, TT>(sh, chains, u, A, v_all, partials, first, cnt, ref.blk, Vs, ring, ring_it, scan_elems);
    __syncthreads();
  }
}

// One EM run = NROUND steps of {estep, scan} plus a closing estep that applies the last arg-max.
int launch_em_steps(ChainDev* chains_dev, int64_t n_chains, bool any_scan, const ScanRef* refs_dev, int64_t n_refs,
                    const UtrDev* utrs_dev, const int32_t* utr_chain_off_dev, const void* tensor, bool f32,
                    const double* cnt, double* lz, double* vbuf, void* partials, double* scan_elems,
                    int32_t* trace_a, int32_t* trace_b, double* trace_ws, cudaStream_t st) {
  const size_t smem = (size_t)SCAN_GB * SCAN_VCHUNK * sizeof(double) +
                      (size_t)RING_STAGES * RING_CH * RING_PITCH * (f32 ? sizeof(float) : sizeof(double));
  if (f32) cudaFuncSetAttribute(em_scan_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  else cudaFuncSetAttribute(em_scan_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int launches = 0;
  for (int step = 0; step <= SCAPE_B200_NROUND; step++) {
    if (f32)
      em_estep_kernel<float><<<(unsigned)n_chains, GT, 0, st>>>(chains_dev, utrs_dev, tensor, cnt, lz, vbuf,
                                                                (const ScanPartial*)partials, trace_a, trace_b, trace_ws);
    else
      em_estep_kernel<double><<<(unsigned)n_chains, GT, 0, st>>>(chains_dev, utrs_dev, tensor, cnt, lz, vbuf,
                                                                 (const ScanPartial*)partials, trace_a, trace_b, trace_ws);
    launches++;
    if (step == SCAPE_B200_NROUND || !any_scan || n_refs == 0) continue;
    if (f32)
      em_scan_kernel<float><<<(unsigned)n_refs, GT, smem, st>>>(refs_dev, chains_dev, utrs_dev, utr_chain_off_dev,
                                                                tensor, vbuf, (ScanPartial*)partials, scan_elems);
    else
      em_scan_kernel<double><<<(unsigned)n_refs, GT, smem, st>>>(refs_dev, chains_dev, utrs_dev, utr_chain_off_dev,
                                                                 tensor, vbuf, (ScanPartial*)partials, scan_elems);
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
      val = lw + (double)tensor[u.tensor_off + (int64_t)n * u.ldR + (int64_t)jb.a_idx[j] * u.B + jb.b_idx[j]];
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
