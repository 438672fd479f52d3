// Chain-resident EM: one CTA runs ONE chain's remaining EM iterations to convergence, with no kernel
// boundary, no inter-CTA barrier and no batching (em_algo's loop, apa_core.py:726-746, per chain).
//
// Why unbatched.  The bulk-synchronous kernels batch the chains of a UTR so that the grid arg-max
// becomes a dense FP64 MMA product; that pays while most chains run.  After a few steps most chains
// have converged and what is left is a few stragglers per UTR: their batched steps are latency floors
// (two launches per iteration for the whole wave, MMA tiles that are 7/8 padding).  Chains are
// independent of one another, so here every straggler gets a CTA of its own and iterates as fast as
// its own dependent work allows (~10 us per iteration):
//
//   E pass     the chain's fragments spread over the CTA's 8 warps (estep_group_run, G = 8):
//              cal_z_k :473-488, norm_z :490-495, maximize_ws :498-505, elbo :559-573, convergence :743
//   grid scan  max_alpha_beta (:507-523): scores[row] = sum_n tensor[n][row] * V[n] over the chain's own
//              window and fragment hull, one thread per 4 candidate rows, V broadcast from shared
//              memory, tensor rows streamed from L2 (chains of one UTR are neighbours in the launch
//              order, so the ~300 resident CTAs work on a dozen UTRs whose tensors stay in L2).  FP64
//              FMAs in ascending fragment order; on B200 the CUDA-core FP64 rate equals the DMMA rate,
//              and a single chain cannot fill an 8-wide MMA anyway.  The float -> double conversion of
//              the tensor elements is integer arithmetic (the F2F instruction runs on the 16-lane XU
//              pipe, which is exactly as wide as the L2 port: it would halve the rate).
//   arg-max    first maximum in row order, applied at once (apply_row)
#include "em_device.cuh"

namespace scape {

cudaError_t upload_model_const_tail(const ModelConst& mc) { return upload_model_const_tu(mc); }

// float -> double.  xu = false: integer instructions only, exact for finite NORMAL inputs -- tensor
// entries are log-likelihoods (|v| far above float's denormal range, never exactly 0) or the sentinel
// (float's lowest finite value), so the exponent field is never 0 or 255, and the scan reads genuine
// entries only (rows < T*B, fragments < N).  Half of the elements take each path: the XU pipe
// converts 16 per clock and SM (exactly the L2 port's rate), the integer version costs 5 issue slots.
__device__ __forceinline__ double tensor_elem(float f, bool xu) {   // `xu` is a compile-time constant after unrolling
  if (xu) return (double)f;                              // F2F.F64.F32 on the XU pipe
  const uint32_t x = __float_as_uint(f);
  const uint32_t hi = (((x & 0x7fffffffu) >> 3) + 0x38000000u) | (x & 0x80000000u);
  return __hiloint2double((int)hi, (int)(x << 29));
}
__device__ __forceinline__ double tensor_elem(double d, bool) { return d; }

constexpr int TAIL_UNROLL = 8;                  // fragments (128-bit loads of 4 rows) in flight per thread

struct TailShared {
  ChainDev ch;
  EGroupShared gs;
  double wbest[GW];
  int wrow[GW];
  int go;
};

template <typename TT>
__device__ __forceinline__ void tail_estep(EGroupShared& gs, int tid, ChainDev& ch, ScanDesc& sd, const UtrDev& u,
                                           const TT* __restrict__ A, const double* __restrict__ cnt,
                                           double* __restrict__ lz, double* __restrict__ V) {
  switch (ch.K) {
    case 1: estep_group_run<2, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 2: estep_group_run<3, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 3: estep_group_run<4, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 4: estep_group_run<5, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 5: estep_group_run<6, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 6: estep_group_run<7, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 7: estep_group_run<8, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 8: estep_group_run<9, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 9: estep_group_run<10, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 10: estep_group_run<11, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 11: estep_group_run<12, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 12: estep_group_run<13, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 13: estep_group_run<14, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 14: estep_group_run<15, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    case 15: estep_group_run<16, TT>(gs, GW, 0, tid, ch, sd, u, A, cnt, lz, V); break;
    default: break;
  }
}

// 4 consecutive candidate rows of fragment n: one 128-bit load (float storage) or two (double storage).
// Row groups start at multiples of 4 and the tensor pitch is a multiple of 4 elements (16-byte aligned).
template <typename TT> struct Row4;
template <> struct Row4<float> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = __ldg(reinterpret_cast<const float4*>(p)); }
  // three of four conversions on the XU pipe (F2F), one with integer instructions: the XU pipe
  // converts 16 elements per clock and SM, exactly what the L2 port delivers
  __device__ __forceinline__ double e0() const { return (double)v.x; }
  __device__ __forceinline__ double e1() const { return (double)v.y; }
  __device__ __forceinline__ double e2() const { return (double)v.z; }
  __device__ __forceinline__ double e3() const { return tensor_elem(v.w, false); }
};
template <> struct Row4<double> {
  double2 a, b;
  __device__ __forceinline__ void load(const double* p) {
    a = __ldg(reinterpret_cast<const double2*>(p));
    b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  }
  __device__ __forceinline__ double e0() const { return a.x; }
  __device__ __forceinline__ double e1() const { return a.y; }
  __device__ __forceinline__ double e2() const { return b.x; }
  __device__ __forceinline__ double e3() const { return b.y; }
};

// first maximum (larger score, then smaller row) of scores[row] = sum_{n in [h0, h1)} A[n][row] * V[n]
// over rows [w0, w1); every thread returns the CTA-wide result.  A thread owns 4 consecutive rows
// per sweep of 1024 rows; TAIL_UNROLL fragments (128-bit loads) are in flight per thread.
template <typename TT>
__device__ __forceinline__ int tail_scan(TailShared& sh, const UtrDev& u, const TT* __restrict__ A, const double* Vs,
                                         int w0, int w1, int h0, int h1) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t R = u.ldR;
  double best = -CUDART_INF;
  int brow = 0x7fffffff;
  for (int rbase = w0 & ~3; rbase < w1; rbase += GT * 4) {
    const int r0 = rbase + 4 * tid;                       // this thread's rows r0 .. r0 + 3
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (r0 < w1) {                                        // (r0 < w1 <= T*B <= ldR: the 4 rows stay inside the pitch)
      const TT* p = A + r0 + (int64_t)h0 * R;
      int n = h0;
      for (; n + TAIL_UNROLL <= h1; n += TAIL_UNROLL) {
        Row4<TT> a[TAIL_UNROLL];
#pragma unroll
        for (int i = 0; i < TAIL_UNROLL; i++) a[i].load(p + (int64_t)i * R);
        p += (int64_t)TAIL_UNROLL * R;
#pragma unroll
        for (int i = 0; i < TAIL_UNROLL; i++) {
          const double v = Vs[n + i];
          acc[0] = fma(a[i].e0(), v, acc[0]);
          acc[1] = fma(a[i].e1(), v, acc[1]);
          acc[2] = fma(a[i].e2(), v, acc[2]);
          acc[3] = fma(a[i].e3(), v, acc[3]);
        }
      }
      for (; n < h1; n++) {
        Row4<TT> a;
        a.load(p);
        p += R;
        const double v = Vs[n];
        acc[0] = fma(a.e0(), v, acc[0]);
        acc[1] = fma(a.e1(), v, acc[1]);
        acc[2] = fma(a.e2(), v, acc[2]);
        acc[3] = fma(a.e3(), v, acc[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int row = r0 + j;                           // rows ascend with j, and sweeps ascend: `>` keeps the first
        if (row >= w0 && row < w1 && acc[j] > best) { best = acc[j]; brow = row; }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
    if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
  }
  if (lane == 0) { sh.wbest[warp] = best; sh.wrow[warp] = brow; }
  __syncthreads();
  best = sh.wbest[0];
  brow = sh.wrow[0];
#pragma unroll
  for (int w = 1; w < GW; w++) {
    const double ob = sh.wbest[w];
    const int orow = sh.wrow[w];
    if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
  }
  return brow;
}

// PROWS: candidate rows per partial of the scan that ran before this kernel (a pending arg-max of it
// is applied first).  `list` = chains to run (or nullptr: chains first .. first + gridDim.x - 1).
template <typename TT, int PROWS>
__global__ void __launch_bounds__(GT, 2)
em_tail_kernel(ChainDev* chains, ScanDesc* descs, const int32_t* __restrict__ list, int first,
               const UtrDev* __restrict__ utrs, const void* __restrict__ tensor, const double* __restrict__ cnt_all,
               double* lz_all, const ScanPartial* __restrict__ partials, int32_t* trace_a, int32_t* trace_b,
               double* trace_ws, unsigned long long* stats) {
  extern __shared__ double Vs[];                         // V[n] = Z[n,k] cnt[n] of the current iteration
  __shared__ TailShared sh;
  const int tid = threadIdx.x;
  const int ci = list ? list[blockIdx.x] : first + (int)blockIdx.x;
  ChainDev& gch = chains[ci];
  if (gch.state == 0) return;                            // uniform
  ChainDev& ch = sh.ch;
  copy_chain(&ch, &gch, tid, GT);
  __syncthreads();
  if (tid == 0) ch.grid_rows_head = ch.grid_rows;        // accounting: what the bulk-synchronous head already scanned
  ScanDesc& sd = descs[ci];
  const UtrDev u = utrs[ch.utr];
  const TT* A = (const TT*)tensor + u.tensor_off;
  const double* cnt = cnt_all + u.frag_off;
  double* lz = lz_all + ch.lz_off;
  if (tid < 32) {
    const int go = apply_pending<PROWS>(ch, sd, u, partials, trace_a, trace_b, trace_ws);
    if (tid == 0) sh.go = go;
  }
  __syncthreads();
  // development aid (SCAPE_B200_DBG): cycles thread 0 sees per phase, summed over all CTAs
  long long t_e = 0, t_scan = 0, t_apply = 0, t_mark = 0, t_begin = 0;
  int iters = 0;
  const bool prof = stats != nullptr && tid == 0;
  if (prof) t_begin = t_mark = clock64();
#define TL_LAP(acc) do { if (prof) { const long long now__ = clock64(); acc += now__ - t_mark; t_mark = now__; } } while (0)
  while (sh.go) {
    tail_estep<TT>(sh.gs, tid, ch, sd, u, A, cnt, lz, Vs);
    __syncthreads();                                     // the epilogue's writes to the chain, V complete
    TL_LAP(t_e);
    iters++;
    if (ch.weights_only) {                               // mstep_fixed (:552-557): no grid search
      if (tid == 0) {
        if (ch.trace_off >= 0) {
          const int64_t o = ch.trace_off + (int64_t)(ch.n_iter - 1) * (SCAPE_B200_KCAP + 1);
          for (int k = 0; k < ch.K; k++) { trace_a[o + k] = ch.a_idx[k]; trace_b[o + k] = ch.b_idx[k]; }
          for (int k = 0; k <= ch.K; k++) trace_ws[o + k] = ch.ws[k];
        }
        sh.go = ch.state == 1 && ch.n_iter < SCAPE_B200_NROUND;
      }
      __syncthreads();
      continue;
    }
    int h0 = 0, h1 = 0;
    if (ch.hhi >= 0) { h0 = ch.hlo; h1 = ch.hhi + 1; }   // no fragment with v != 0: all scores 0, the first row wins
    const int row = tail_scan<TT>(sh, u, A, Vs, ch.row0, ch.row1, h0, h1);
    __syncthreads();                                     // everybody has read the chain's window
    TL_LAP(t_scan);
    if (tid == 0) sh.go = apply_row(ch, sd, u, true, row, trace_a, trace_b, trace_ws);
    __syncthreads();
    TL_LAP(t_apply);
  }
  copy_chain(&gch, &ch, tid, GT);
  if (prof) {
    atomicAdd(stats + 0, (unsigned long long)iters);
    atomicAdd(stats + 1, (unsigned long long)(clock64() - t_begin));
    atomicAdd(stats + 2, (unsigned long long)t_e);
    atomicAdd(stats + 3, (unsigned long long)t_scan);
    atomicAdd(stats + 4, (unsigned long long)t_apply);
    atomicAdd(stats + 5, 1ull);
  }
#undef TL_LAP
}

template <typename TT>
static cudaError_t launch_tail_t(ChainDev* chains_dev, ScanDesc* descs_dev, const int32_t* list_dev, int first, int n,
                                 int max_n, int prows, const UtrDev* utrs_dev, const void* tensor, const double* cnt,
                                 double* lz, const void* partials, int32_t* trace_a, int32_t* trace_b,
                                 double* trace_ws, unsigned long long* stats, cudaStream_t st) {
  const size_t smem = sizeof(double) * size_t((max_n + 7) / 8 * 8 + 8);
  const ScanPartial* pb = (const ScanPartial*)partials;
  if (prows == kClusterTileRows) {
    if (smem > 40 * 1024) cudaFuncSetAttribute(em_tail_kernel<TT, kClusterTileRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    em_tail_kernel<TT, kClusterTileRows><<<(unsigned)n, GT, smem, st>>>(chains_dev, descs_dev, list_dev, first, utrs_dev, tensor,
                                                                          cnt, lz, pb, trace_a, trace_b, trace_ws, stats);
  } else {
    if (smem > 40 * 1024) cudaFuncSetAttribute(em_tail_kernel<TT, SCAN_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    em_tail_kernel<TT, SCAN_ROWS><<<(unsigned)n, GT, smem, st>>>(chains_dev, descs_dev, list_dev, first, utrs_dev, tensor, cnt,
                                                                  lz, pb, trace_a, trace_b, trace_ws, stats);
  }
  return cudaGetLastError();
}

cudaError_t launch_em_tail(ChainDev* chains_dev, ScanDesc* descs_dev, const int32_t* list_dev, int first, int n,
                           int max_n, int prows, const UtrDev* utrs_dev, const void* tensor, bool f32,
                           const double* cnt, double* lz, const void* partials, int32_t* trace_a, int32_t* trace_b,
                           double* trace_ws, unsigned long long* stats, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (f32)
    return launch_tail_t<float>(chains_dev, descs_dev, list_dev, first, n, max_n, prows, utrs_dev, tensor, cnt, lz, partials,
                                trace_a, trace_b, trace_ws, stats, st);
  return launch_tail_t<double>(chains_dev, descs_dev, list_dev, first, n, max_n, prows, utrs_dev, tensor, cnt, lz, partials,
                               trace_a, trace_b, trace_ws, stats, st);
}

}  // namespace scape
