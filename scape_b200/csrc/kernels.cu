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
// K4: EM group kernel.  One CTA runs the (up to 10) restarts of one (UTR, K) pair in lockstep:
//
//   E phase   one warp per chain: column refresh (cal_z_k :473-488), count-tempered softmax (norm_z
//             :490-495), weight update (maximize_ws :498-505, mstep guard :526-529), ELBO (:559-573)
//             and the convergence test (:743).  Z is never materialised; the pass leaves
//             v_g[n] = Z[n,k] cnt[n] in shared memory as V[n][slot].
//   scan      max_alpha_beta (:507-523) for ALL running chains at once: thread <-> candidate row,
//             scores[row][slot] = sum_n tensor[n][row] * V[n][slot].  With the tensor stored
//             [n][row] the loads are perfectly coalesced, V is a shared-memory broadcast, and every
//             tensor element fetched from L2 feeds up to 10 FP64 FMAs (one per restart) instead of
//             one.  Only the hull of fragments with v != 0 is visited (other terms are exactly +-0).
//   arg-max   per chain over its own window [alpha_{k-1}, alpha_{k+1}] x all beta, first maximum
//             in (alpha asc, beta asc) order; then BIC at the end (:702-706).
// ------------------------------------------------------------------------------------------------
constexpr int GT = 256;                    // threads per CTA
constexpr int GW = GT / 32;                // warps
constexpr int GMAX = SCAPE_B200_NTRIAL;    // chains per group
constexpr int EM_MULTI_KMAX = 7;
constexpr int SCAN_ROWS = 2 * GT;          // candidate rows per block (2 per thread)
constexpr int SCAN_MAXBLK = 96;
constexpr int RING_STAGES = 3;             // TMA ring: stages in flight
constexpr int RING_CH = 8;                 // fragments (tensor n-rows) per stage
constexpr int RING_PITCH = SCAN_ROWS + 8;  // elements per staged n-row (up to 3 + 3 elements of 16-byte alignment slack)

struct GroupShared {
  double w[GMAX][SCAPE_B200_KCAP + 1], lw[GMAX][SCAPE_B200_KCAP + 1];
  int a[GMAX][SCAPE_B200_KCAP], b[GMAX][SCAPE_B200_KCAP];
  double lb[GMAX], last_a[GMAX], grid_rows[GMAX];
  int k[GMAX], row0[GMAX], row1[GMAX], hlo[GMAX], hhi[GMAX], n_iter[GMAX];
  int state[GMAX];        // 0 finished, 1 running, 2 converged in this step (finishes after the scan)
  int slot_of[GMAX];      // column of V, or -1 (weights-only chains do not scan)
  int chain_of[GMAX];
  int n_run, n_scan, ga, R0, R1, N0, N1;
  double bscore[GW][GMAX];
  int brow[GW][GMAX];
  double grid_elems;
  int n_blk;
  int blk_cnt[SCAN_MAXBLK];
  unsigned char blk_idx[SCAN_MAXBLK][GMAX];   // slots whose window intersects the row block
  unsigned long long full_bar[RING_STAGES];   // mbarriers: stage filled by TMA
};

__device__ __forceinline__ int ga_bucket(int n) { return n <= 1 ? 1 : n <= 2 ? 2 : n <= 4 ? 4 : n <= 6 ? 6 : n <= 8 ? 8 : 10; }

template <int NK, typename TT>
__device__ __forceinline__ void e_step_warp(GroupShared& sh, int g, int it, ChainDev& ch, const UtrDev& u,
                                            const TT* __restrict__ A, int64_t R, const double* __restrict__ cnt,
                                            double* __restrict__ lz, double* V, int nv, int slot) {
  constexpr int K = NK - 1;
  const int lane = threadIdx.x & 31;
  const int N = u.N, npad = u.Npad, B = u.B;
  const int k = ch.k_order[it];
  const double lwk = sh.lw[g][k];
  const int64_t rk = (int64_t)sh.a[g][k] * B + sh.b[g][k];
  const double cap = c_mc.max_unif_ws;
  bool guard = false;
  double red[NK + 3];
  int h_lo, h_hi;
  while (true) {
#pragma unroll
    for (int j = 0; j < NK + 3; j++) red[j] = 0.0;
    h_lo = N;
    h_hi = -1;
    for (int n = lane; n < N; n += 32) {
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
      if (slot >= 0) V[(size_t)slot * nv + n] = vn;
      if (vn != 0.0) { h_lo = min(h_lo, n); h_hi = n; }
    }
#pragma unroll
    for (int j = 0; j < NK + 3; j++) {
      double x = red[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      red[j] = x;
    }
    if (!guard && red[NK] < 1e-8) {          // mstep guard (:526-529), uniform across the warp
      guard = true;
      continue;
    }
    break;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    h_lo = min(h_lo, __shfl_xor_sync(0xffffffffu, h_lo, o));
    h_hi = max(h_hi, __shfl_xor_sync(0xffffffffu, h_hi, o));
  }
  if (lane == 0) {
    // maximize_ws (:498-505)
    double w[NK];
    double tot = 0.0;
#pragma unroll
    for (int j = 0; j < NK; j++) tot += red[j];
#pragma unroll
    for (int j = 0; j < NK; j++) w[j] = red[j] / tot;
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
      sh.w[g][j] = w[j];
      sh.lw[g][j] = (w[j] <= 0.0) ? SCAPE_SENTINEL : log(w[j]);
    }
    sh.k[g] = k;
    const int lo = (k == 0) ? 0 : sh.a[g][k - 1];
    const int hi = (k == K - 1) ? u.T - 1 : sh.a[g][k + 1];
    sh.row0[g] = lo * B;
    sh.row1[g] = (hi + 1) * B;
    sh.hlo[g] = h_lo;
    sh.hhi[g] = h_hi;
    const double lb_new = red[NK + 1] + red[NK + 2];            // elbo (:559-561)
    sh.last_a[g] = red[NK + 1];
    ch.lb_arr[it] = lb_new;
    sh.n_iter[g] = it + 1;
    const double lb = sh.lb[g];
    const bool conv = fabs(lb_new - lb) < fabs(1e-6 * lb);      // (:743)
    if (!conv) sh.lb[g] = lb_new;
    sh.state[g] = (conv || it == SCAPE_B200_NROUND - 1) ? 2 : 1;
  }
  __syncwarp();
}

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

// One block of SCAN_ROWS candidate rows against the GB chains whose window intersects it:
// acc[row][j] = sum_n tensor[n][row] * V[slot_j][n].
//   * the tensor slab [n0..n1) x [block rows] streams through a RING_STAGES-deep shared-memory
//     ring filled by 1-D TMA bulk copies (one per fragment row: the block's rows are contiguous in
//     the [n][row] layout), completion tracked by mbarriers: deep prefetch without registers;
//   * thread <-> 2 rows (conflict-free LDS from the stage), V is slot-major in shared memory so two
//     consecutive fragments of one chain come in one 16-byte broadcast LDS;
//   * all sums / products are FP64; accumulators stay in registers for the whole block;
//   * when V does not fit shared memory (`staged`) it lives in global scratch and is staged chunk by
//     chunk (compacted to the block's chains);
//   * per-thread running maxima live in shared memory (s_best[slot][tid]) because the chain subset
//     changes from block to block.
template <int GB, typename TT>
__device__ __forceinline__ void block_scan(GroupShared& sh, const TT* __restrict__ A, int64_t R, double* Vs,
                                           const double* Vg, int nv, bool staged, int vcap, int blk, int base,
                                           double* s_best, int* s_brow, TT* ring, uint32_t& ring_it) {
  const int tid = threadIdx.x;
  const int R1 = sh.R1, N1 = sh.N1;
  const int N0 = sh.N0 & ~3;                    // aligned start (V is exactly 0 outside the hull)
  const int cnt = sh.blk_cnt[blk];
  int idx[GB];
#pragma unroll
  for (int j = 0; j < GB; j++) idx[j] = sh.blk_idx[blk][j < cnt ? j : 0];
  const int r0 = base + tid, r1 = r0 + GT;
  // 16-byte aligned row range of this block that the TMA copies fetch per fragment
  const int base_al = base & ~3;
  const int end_al = min((min(base + SCAN_ROWS, R1) + 3) & ~3, (int)R);
  const uint32_t row_bytes = (uint32_t)(end_al - base_al) * (uint32_t)sizeof(TT);
  const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t bar_base = (uint32_t)__cvta_generic_to_shared(sh.full_bar);
  constexpr uint32_t STAGE_BYTES = RING_CH * RING_PITCH * sizeof(TT);
  const uint32_t off0 = (uint32_t)(r0 - base_al) * (uint32_t)sizeof(TT);
  const uint32_t off1 = off0 + GT * (uint32_t)sizeof(TT);

  double acc0[GB], acc1[GB];
#pragma unroll
  for (int j = 0; j < GB; j++) acc0[j] = acc1[j] = 0.0;
  const uint32_t vs_base = (uint32_t)__cvta_generic_to_shared(Vs);
  const int chunk = staged ? ((vcap / GB) & ~7) : (((N1 - N0) + 8) & ~7);
  for (int c0 = N0; c0 < N1; c0 += chunk) {
    const int c1 = min(c0 + chunk, N1);
    uint32_t vj[GB];
    if (staged) {
      __syncthreads();                          // previous chunk fully consumed
      const int len = (c1 - c0 + 3) & ~3;
      for (int e = tid; e < GB * len; e += GT) {
        const int j = e / len, o = e % len;
        Vs[j * chunk + o] = (j < cnt) ? Vg[(size_t)idx[j < cnt ? j : 0] * nv + c0 + o] : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < GB; j++) vj[j] = vs_base + (uint32_t)(j * chunk) * 8u;
    } else {
#pragma unroll
      for (int j = 0; j < GB; j++) vj[j] = vs_base + (uint32_t)(idx[j] * nv + c0) * 8u;
    }
    const int n_it = (c1 - c0 + RING_CH - 1) / RING_CH;
    // producer prologue: fill the ring
    if (tid == 0) {
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
            const double2 v = lds_f64x2(vj[j] + (uint32_t)i * 8u);
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
            const double v = lds_f64(vj[j] + (uint32_t)i * 8u);
            acc0[j] = fma(a0, v, acc0[j]);
            acc1[j] = fma(b0, v, acc1[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < GB; j++) vj[j] += RING_CH * 8u;
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
#pragma unroll
  for (int j = 0; j < GB; j++) {
    if (j < cnt) {
      const int s = idx[j], c = sh.chain_of[s];
      const int w0 = sh.row0[c], w1 = sh.row1[c];
      double b = s_best[s * GT + tid];
      int r = s_brow[s * GT + tid];
      bool upd = false;
      if (r0 >= w0 && r0 < w1 && acc0[j] > b) { b = acc0[j]; r = r0; upd = true; }   // rows ascend: first max wins
      if (r1 >= w0 && r1 < w1 && acc1[j] > b) { b = acc1[j]; r = r1; upd = true; }
      if (upd) { s_best[s * GT + tid] = b; s_brow[s * GT + tid] = r; }
    }
  }
}

// Row blocks [blk0, blk0 + n_blk) of the current super-range, lists already in shared memory.
template <typename TT>
__device__ __forceinline__ void group_scan_blocks(GroupShared& sh, const TT* __restrict__ A, int64_t R,
                                                  double* Vs, const double* Vg, int nv, bool staged, int vcap,
                                                  int row_base, double* s_best, int* s_brow, TT* ring,
                                                  uint32_t& ring_it) {
  const int n_blk = sh.n_blk;
  for (int blk = 0; blk < n_blk; blk++) {
    const int base = row_base + blk * SCAN_ROWS;
    const int cnt = sh.blk_cnt[blk];            // uniform across the CTA
    if (cnt == 0) continue;
    if (cnt <= 1) block_scan<1, TT>(sh, A, R, Vs, Vg, nv, staged, vcap, blk, base, s_best, s_brow, ring, ring_it);
    else if (cnt <= 2) block_scan<2, TT>(sh, A, R, Vs, Vg, nv, staged, vcap, blk, base, s_best, s_brow, ring, ring_it);
    else if (cnt <= 3) block_scan<3, TT>(sh, A, R, Vs, Vg, nv, staged, vcap, blk, base, s_best, s_brow, ring, ring_it);
    else if (cnt <= 4) block_scan<4, TT>(sh, A, R, Vs, Vg, nv, staged, vcap, blk, base, s_best, s_brow, ring, ring_it);
    else if (cnt <= 6) block_scan<6, TT>(sh, A, R, Vs, Vg, nv, staged, vcap, blk, base, s_best, s_brow, ring, ring_it);
    else if (cnt <= 8) block_scan<8, TT>(sh, A, R, Vs, Vg, nv, staged, vcap, blk, base, s_best, s_brow, ring, ring_it);
    else block_scan<10, TT>(sh, A, R, Vs, Vg, nv, staged, vcap, blk, base, s_best, s_brow, ring, ring_it);
  }
}

// first maximum in row order: larger score wins, ties go to the smaller row
__device__ __forceinline__ void group_scan_reduce(GroupShared& sh, const double* s_best, const int* s_brow) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int s = 0; s < sh.n_scan; s++) {
    double b = s_best[s * GT + tid];
    int r = s_brow[s * GT + tid];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, b, o);
      const int orow = __shfl_xor_sync(0xffffffffu, r, o);
      if (ob > b || (ob == b && orow < r)) { b = ob; r = orow; }
    }
    if (lane == 0) { sh.bscore[warp][s] = b; sh.brow[warp][s] = r; }
  }
}

template <int NK, typename TT>
__device__ void em_group_run(GroupShared& sh, const GroupDev& grp, ChainDev* chains, const UtrDev& u,
                             const TT* __restrict__ A, const double* __restrict__ cnt, double* lz_all,
                             double* Vs, double* Vg, int vcap, double* s_best, int* s_brow, TT* ring,
                             int32_t* trace_a, int32_t* trace_b, double* trace_ws) {
  constexpr int K = NK - 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = grp.n_chains;
  const int N = u.N, npad = u.Npad, B = u.B;
  const int64_t R = u.ldR;                      // pitch of one fragment's candidate rows
  ChainDev* my = chains + grp.first_chain;
  const int nv = (N + 3) & ~3;                  // row pitch of V (slot-major), 32-byte aligned rows
  const bool staged = (int64_t)nv * GMAX > vcap; // V in global scratch, staged through shared memory by the scan
  double* V = staged ? Vg : Vs;

  uint32_t ring_it = 0;                          // stage / phase bookkeeping of the TMA ring (uniform across the CTA)
  if (tid == 0) {
    for (int st = 0; st < RING_STAGES; st++) mbar_init((uint32_t)__cvta_generic_to_shared(&sh.full_bar[st]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // ---- load the init blobs, initial log_zmat (em_algo :722-724) ----
  for (int g = warp; g < G; g += GW) {
    ChainDev& ch = my[g];
    if (lane < NK) {
      const double w = ch.ws[lane];
      sh.w[g][lane] = w;
      sh.lw[g][lane] = (w <= 0.0) ? SCAPE_SENTINEL : log(w);
      if (lane < K) { sh.a[g][lane] = ch.a_idx[lane]; sh.b[g][lane] = ch.b_idx[lane]; }
    }
    if (lane == 0) {
      sh.lb[g] = SCAPE_SENTINEL;
      sh.last_a[g] = 0.0;
      sh.grid_rows[g] = 0.0;
      sh.n_iter[g] = 0;
      sh.state[g] = 1;
    }
    __syncwarp();
    double* lz = lz_all + ch.lz_off;
    for (int j = 0; j < NK; j++) {
      const double lw = sh.lw[g][j];
      if (j < K) {
        const int64_t rj = (int64_t)sh.a[g][j] * B + sh.b[g][j];
        for (int n = lane; n < N; n += 32) lz[(int64_t)j * npad + n] = lw + (double)A[(int64_t)n * R + rj];
      } else {
        const double val = lw + u.unif_loglik;
        for (int n = lane; n < N; n += 32) lz[(int64_t)j * npad + n] = val;
      }
    }
  }
  if (tid == 0) {
    int ns = 0;
    for (int g = 0; g < G; g++) {
      const bool scans = !my[g].weights_only;
      sh.slot_of[g] = scans ? ns : -1;
      if (scans) sh.chain_of[ns++] = g;
    }
    sh.n_scan = ns;
    sh.n_run = G;
    sh.ga = ga_bucket(ns);
    sh.grid_elems = 0.0;
  }
  __syncthreads();

  long long t_e = 0, t_s = 0, t_b = 0, steps = 0;
  for (int it = 0; it < SCAPE_B200_NROUND; it++) {
    const int n_scan = sh.n_scan;
    const long long c_a = clock64();
    // ---- E phase: one warp per running chain ----
    for (int g = warp; g < G; g += GW)
      if (sh.state[g] == 1)
        e_step_warp<NK, TT>(sh, g, it, my[g], u, A, R, cnt, lz_all + my[g].lz_off, V, nv, sh.slot_of[g]);
    __syncthreads();
    const long long c_b = clock64();
    if (n_scan > 0) {
      if (tid == 0) {
        int R0 = 1 << 30, R1 = 0, N0 = 1 << 30, N1 = 0;
        for (int s = 0; s < n_scan; s++) {
          const int c = sh.chain_of[s];
          R0 = min(R0, sh.row0[c]);
          R1 = max(R1, sh.row1[c]);
          if (sh.hhi[c] >= 0) { N0 = min(N0, sh.hlo[c]); N1 = max(N1, sh.hhi[c] + 1); }
          sh.grid_rows[c] += (double)(sh.row1[c] - sh.row0[c]);
        }
        if (N1 <= N0) { N0 = 0; N1 = 0; }   // every v is zero: all scores 0, first row of each window wins
        sh.R0 = R0; sh.R1 = R1; sh.N0 = N0; sh.N1 = N1;
      }
      for (int s2 = 0; s2 < n_scan; s2++) { s_best[s2 * GT + tid] = -CUDART_INF; s_brow[s2 * GT + tid] = 0x7fffffff; }
      __syncthreads();
      // super-ranges of at most SCAN_MAXBLK row blocks (one is enough unless R > 49152)
      for (int sb = sh.R0; sb < sh.R1; sb += SCAN_MAXBLK * SCAN_ROWS) {
        const int nb = min(SCAN_MAXBLK, (sh.R1 - sb + SCAN_ROWS - 1) / SCAN_ROWS);
        // which chains need which row block (window intersection)
        if (tid < nb) {
          const int lo = sb + tid * SCAN_ROWS, hi = min(lo + SCAN_ROWS, sh.R1);
          int cnt = 0;
          for (int s2 = 0; s2 < n_scan; s2++) {
            const int c = sh.chain_of[s2];
            if (sh.row0[c] < hi && sh.row1[c] > lo) sh.blk_idx[tid][cnt++] = (unsigned char)s2;
          }
          sh.blk_cnt[tid] = cnt;
          if (cnt) atomicAdd(&sh.grid_elems, (double)(hi - lo) * (double)(sh.N1 - sh.N0));
        }
        if (tid == 0) sh.n_blk = nb;
        __syncthreads();
        group_scan_blocks<TT>(sh, A, R, Vs, Vg, nv, staged, vcap, sb, s_best, s_brow, ring, ring_it);
        __syncthreads();
      }
      group_scan_reduce(sh, s_best, s_brow);
      __syncthreads();
      if (tid < n_scan) {
        double b = sh.bscore[0][tid];
        int r = sh.brow[0][tid];
        for (int w = 1; w < GW; w++) {
          const double ob = sh.bscore[w][tid];
          const int orow = sh.brow[w][tid];
          if (ob > b || (ob == b && orow < r)) { b = ob; r = orow; }
        }
        const int c = sh.chain_of[tid];
        const int k = sh.k[c];
        sh.a[c][k] = r / B;
        sh.b[c][k] = r % B;
      }
      __syncthreads();
    }
    const long long c_c = clock64();
    // ---- bookkeeping: traces, finished chains, slots for the next step ----
    if (tid < G && sh.state[tid] >= 1) {
      const int g = tid;
      ChainDev& ch = my[g];
      if (ch.trace_off >= 0) {
        const int64_t o = ch.trace_off + (int64_t)it * (SCAPE_B200_KCAP + 1);
        for (int j = 0; j < K; j++) { trace_a[o + j] = sh.a[g][j]; trace_b[o + j] = sh.b[g][j]; }
        for (int j = 0; j < NK; j++) trace_ws[o + j] = sh.w[g][j];
      }
      if (sh.state[g] == 2) {
        ch.n_iter = sh.n_iter[g];
        ch.bic = -2.0 * sh.last_a[g] + (3 * K + 1) * log((double)N);   // cal_bic (:702-706)
        ch.grid_rows = sh.grid_rows[g];
        for (int j = 0; j < K; j++) { ch.a_idx[j] = sh.a[g][j]; ch.b_idx[j] = sh.b[g][j]; }
        for (int j = 0; j < NK; j++) ch.ws[j] = sh.w[g][j];
      }
    }
    __syncthreads();
    if (tid == 0) {
      int ns = 0, nr = 0;
      for (int g = 0; g < G; g++) {
        if (sh.state[g] == 2) sh.state[g] = 0;
        if (sh.state[g] == 1) {
          nr++;
          const bool scans = !my[g].weights_only;
          sh.slot_of[g] = scans ? ns : -1;
          if (scans) sh.chain_of[ns++] = g;
        } else {
          sh.slot_of[g] = -1;
        }
      }
      sh.n_scan = ns;
      sh.n_run = nr;
      sh.ga = ga_bucket(ns);
    }
    __syncthreads();
    t_e += c_b - c_a; t_s += c_c - c_b; t_b += clock64() - c_c; steps++;
    if (sh.n_run == 0) break;
  }
  if (tid == 0) {
    my[0].grid_elems = sh.grid_elems;   // tensor elements the whole group loaded
    my[0].dbg[0] = (double)t_e; my[0].dbg[1] = (double)t_s; my[0].dbg[2] = (double)t_b; my[0].dbg[3] = (double)steps;
  }
}

// dynamic shared memory: [ V: smem_doubles ][ s_best: GMAX*GT doubles ][ s_brow: GMAX*GT ints ][ TMA ring ]
#define EM_GROUP_PROLOGUE                                                                       \
  extern __shared__ double sm_dyn[];                                                            \
  __shared__ GroupShared sh;                                                                    \
  const GroupDev grp = groups[blockIdx.x];                                                      \
  const UtrDev u = utrs[grp.utr];                                                               \
  double* Vs = sm_dyn;                                                                          \
  double* Vg = v_all + grp.v_off;                                                               \
  double* s_best = sm_dyn + smem_doubles;                                                       \
  int* s_brow = reinterpret_cast<int*>(s_best + GMAX * GT);                                     \
  TT* ring = reinterpret_cast<TT*>(s_brow + GMAX * GT);                                         \
  const TT* A = (const TT*)tensor + u.tensor_off;                                               \
  const double* c = cnt + u.frag_off;

template <int NK, typename TT>
__global__ void __launch_bounds__(GT, 2)
em_group_kernel(const GroupDev* __restrict__ groups, ChainDev* chains, const UtrDev* __restrict__ utrs,
                const void* __restrict__ tensor, const double* __restrict__ cnt, double* lz_all, double* v_all,
                int smem_doubles, int32_t* trace_a, int32_t* trace_b, double* trace_ws) {
  EM_GROUP_PROLOGUE
  em_group_run<NK, TT>(sh, grp, chains, u, A, c, lz_all, Vs, Vg, smem_doubles, s_best, s_brow, ring, trace_a, trace_b, trace_ws);
}

// K = 1..7 in one launch; groups are ordered by UTR so the CTAs that share a tensor run together
template <typename TT>
__global__ void __launch_bounds__(GT, 2)
em_group_kernel_multi(const GroupDev* __restrict__ groups, ChainDev* chains, const UtrDev* __restrict__ utrs,
                      const void* __restrict__ tensor, const double* __restrict__ cnt, double* lz_all,
                      double* v_all, int smem_doubles, int32_t* trace_a, int32_t* trace_b, double* trace_ws) {
  EM_GROUP_PROLOGUE
  switch (grp.K) {
    case 1: em_group_run<2, TT>(sh, grp, chains, u, A, c, lz_all, Vs, Vg, smem_doubles, s_best, s_brow, ring, trace_a, trace_b, trace_ws); break;
    case 2: em_group_run<3, TT>(sh, grp, chains, u, A, c, lz_all, Vs, Vg, smem_doubles, s_best, s_brow, ring, trace_a, trace_b, trace_ws); break;
    case 3: em_group_run<4, TT>(sh, grp, chains, u, A, c, lz_all, Vs, Vg, smem_doubles, s_best, s_brow, ring, trace_a, trace_b, trace_ws); break;
    case 4: em_group_run<5, TT>(sh, grp, chains, u, A, c, lz_all, Vs, Vg, smem_doubles, s_best, s_brow, ring, trace_a, trace_b, trace_ws); break;
    case 5: em_group_run<6, TT>(sh, grp, chains, u, A, c, lz_all, Vs, Vg, smem_doubles, s_best, s_brow, ring, trace_a, trace_b, trace_ws); break;
    case 6: em_group_run<7, TT>(sh, grp, chains, u, A, c, lz_all, Vs, Vg, smem_doubles, s_best, s_brow, ring, trace_a, trace_b, trace_ws); break;
    case 7: em_group_run<8, TT>(sh, grp, chains, u, A, c, lz_all, Vs, Vg, smem_doubles, s_best, s_brow, ring, trace_a, trace_b, trace_ws); break;
    default: break;
  }
}

typedef void (*em_kernel_t)(const GroupDev*, ChainDev*, const UtrDev*, const void*, const double*, double*, double*,
                            int, int32_t*, int32_t*, double*);

template <typename TT>
static em_kernel_t em_kernel_for(int K) {
  switch (K) {
    case 0: return em_group_kernel_multi<TT>;
    case 8: return em_group_kernel<9, TT>;
    case 9: return em_group_kernel<10, TT>;
    case 10: return em_group_kernel<11, TT>;
    case 11: return em_group_kernel<12, TT>;
    case 12: return em_group_kernel<13, TT>;
    case 13: return em_group_kernel<14, TT>;
    case 14: return em_group_kernel<15, TT>;
    case 15: return em_group_kernel<16, TT>;
  }
  return nullptr;
}

// Host-side launch plan.  `groups_host` lists (UTR, K) groups whose chains are contiguous in the
// chain array.  Groups are bucketed by kernel (multi-K / one per K >= 8) and by the shared-memory
// class of their fragment count; inside a bucket they keep the caller's order (UTR-major, K
// descending), so the CTAs that scan one UTR's tensor are resident together (L2 locality).
int launch_em_groups(const std::vector<GroupDev>& groups_host, GroupDev* groups_dev, GroupDev* staging,
                     ChainDev* chains_dev, const UtrDev* utrs_host, const UtrDev* utrs_dev, const void* tensor,
                     bool f32, const double* cnt, double* lz, double* vbuf, int32_t* trace_a, int32_t* trace_b,
                     double* trace_ws, cudaStream_t st) {
  constexpr int SMALL_N = 512, MID_N = 1024;   // V = N x 10 doubles (40 KB / 95 KB) + 30 KB running maxima + TMA ring (50-100 KB)
  struct Plan { int kb, cls; std::vector<GroupDev> g; int max_n = 0; };
  std::vector<Plan> plans;
  auto plan_for = [&](int kb, int cls) -> Plan& {
    for (auto& p : plans)
      if (p.kb == kb && p.cls == cls) return p;
    plans.push_back(Plan{kb, cls, {}, 0});
    return plans.back();
  };
  for (const GroupDev& g : groups_host) {
    const int n = utrs_host[g.utr].N;
    const int cls = n <= SMALL_N ? 0 : n <= MID_N ? 1 : 2;
    Plan& p = plan_for(g.K <= EM_MULTI_KMAX ? 0 : g.K, cls);
    p.g.push_back(g);
    p.max_n = std::max(p.max_n, n);
  }
  size_t pos = 0;
  for (auto& p : plans) {
    std::copy(p.g.begin(), p.g.end(), staging + pos);
    pos += p.g.size();
  }
  cudaMemcpyAsync(groups_dev, staging, sizeof(GroupDev) * pos, cudaMemcpyHostToDevice, st);
  int launches = 0;
  pos = 0;
  for (auto& p : plans) {
    em_kernel_t kern = f32 ? em_kernel_for<float>(p.kb) : em_kernel_for<double>(p.kb);
    const int smem_doubles = p.cls == 2 ? ((MID_N + 3) & ~3) * GMAX : ((p.max_n + 3) & ~3) * GMAX;
    const size_t smem = (size_t)smem_doubles * sizeof(double) + (size_t)GMAX * GT * (sizeof(double) + sizeof(int)) +
                        (size_t)RING_STAGES * RING_CH * RING_PITCH * (f32 ? sizeof(float) : sizeof(double));
    if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)p.g.size(), GT, smem, st>>>(groups_dev + pos, chains_dev, utrs_dev, tensor, cnt, lz, vbuf,
                                                 smem_doubles, trace_a, trace_b, trace_ws);
    pos += p.g.size();
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
