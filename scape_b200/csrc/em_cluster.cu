// Cluster-resident EM: em_algo's iteration loop (apa_core.py:726-746) for ALL chains of a UTR inside
// one kernel launch, one thread-block cluster per UTR.
//
// The bulk-synchronous kernels (kernels.cu) advance every chain of a wave by one iteration per
// launch pair: a wave costs 51 x {E step, scan} launches whose duration is the slowest CTA of the
// whole wave, however few chains still run.  Here a UTR's chains iterate at their own pace:
//
//   repeat
//     E phase    every warp of the cluster owns chains j = gw, gw + W, ... of the UTR: apply the
//                arg-max the last scan found (apply_pending), then one E pass (estep_warp_run:
//                cal_z_k :473-488, norm_z :490-495, maximize_ws :498-505, elbo :559-573, convergence
//                :743); leaves V[n] = Z[n,k] cnt[n] and the candidate window / fragment hull.
//                Weights-only chains (prune refits, :708-711) iterate to convergence right here.
//     barrier.cluster  (V, windows and hulls of all chains visible to every CTA of the cluster)
//     scan       max_alpha_beta (:507-523) on the FP64 tensor cores.  The pending chains are staged
//                into shared memory `cpp` at a time (whole V rows), and the candidate rows are cut into
//                32-row tiles dealt round-robin to the warps of the cluster.  A task = (tile, the
//                chains of the pass whose window covers the tile): scores[32 x 8 NG] = tensor[32 rows]
//                [fragments] * V[fragments][chains] over the union of those chains' hulls, DMMA
//                m8n8k4, tensor rows straight from L2 through a register ring.  Padding is at most
//                31 rows x 7 chains per task (the bulk-synchronous scan pads to 256 rows x the hull
//                union of 32 chains).  Each (chain, tile) leaves its first maximum in `partials`.
//     barrier.cluster
//   until no chain of the UTR has a pending arg-max
//
// Per-chain sums are bit-identical to the bulk-synchronous scan: both accumulate 4-fragment MMA
// steps aligned to multiples of 4 in ascending fragment order, and a fragment outside a chain's hull
// contributes exactly 0.
#include <cooperative_groups.h>

#include "em_device.cuh"

namespace cg = cooperative_groups;

namespace scape {

cudaError_t upload_model_const_cluster(const ModelConst& mc) { return upload_model_const_tu(mc); }

constexpr int CL_TILE = kClusterTileRows;
constexpr int CL_PASS = kClusterPassMax;
constexpr int CL_MAXCH = SCAN_MAXCH;

struct ClShared {
  ChainDev ch[GW];                      // chain record of the chain a warp is working on (E phase)
  int list[CL_MAXCH];                   // pending chains of the UTR (index within the UTR), ascending
  int warp_cnt[GW];
  // chains of the current pass
  int row0[CL_PASS], row1[CL_PASS], hlo[CL_PASS], hhi[CL_PASS];
  long long voff[CL_PASS], pboff[CL_PASS];
  int NA, NB, P, tile_lo, tile_hi;
};

// scores of one 32-row tile against the nt <= 8 NG chains of the pass that cover it (bit `s` of `mask`:
// chain slot s of the pass), fragments [h0, h1), h0 % 4 == 0.
template <int NG, typename TT>
__device__ __forceinline__ void scan_tile_mma(const ClShared& sh, const UtrDev& u, const TT* __restrict__ A,
                                              const double* Vs, ScanPartial* partials, int t, int cntp,
                                              unsigned mask, int nt, int h0, int h1) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;          // MMA group id / thread-in-group
  const int64_t R = u.ldR;
  const int Rv = u.T * u.B;
  const int base = t * CL_TILE;
  const int tile_end = min(base + CL_TILE, Rv);
  const int NA = sh.NA, P = sh.P;
  // B operand of this lane: V[chain slot of column 8 ni + g][fragment k0 + q]; columns past nt read the zero row
  const uint32_t vs_base = (uint32_t)__cvta_generic_to_shared(Vs);
  uint32_t vb[NG];
#pragma unroll
  for (int ni = 0; ni < NG; ni++) {
    const int idx = 8 * ni + g;
    const int slot = idx < nt ? (int)__fns(mask, 0, idx + 1) : cntp;
    vb[ni] = vs_base + (uint32_t)(slot * P + (h0 - NA) + q) * 8u;
  }
  // A rows of this lane: row(mi) = base + 8 mi + g (clamped; rows past the grid are masked below)
  const TT* arow[4];
#pragma unroll
  for (int mi = 0; mi < 4; mi++) arow[mi] = A + min(base + 8 * mi + g, Rv - 1);
  double acc[4][NG][2];
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < NG; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  const int len4 = (h1 - h0 + 3) & ~3;
  // Register ring, refilled unconditionally with pointer increments (see scan_subbatch in kernels.cu):
  // reads run up to 4 PFD + 3 fragments past h1, into the next UTR's tensor or the zeroed slack rows
  // (all finite), against V = 0.
  constexpr int PFD = (NG == 1 ? 12 : NG == 2 ? 8 : 4) / (sizeof(TT) == 8 ? 2 : 1);   // FP64 storage: half the depth, same registers
  const int64_t kstep = 4 * R;
  const TT* pp[4];
  TT pre[PFD][4];
#pragma unroll
  for (int mi = 0; mi < 4; mi++) pp[mi] = arow[mi] + (int64_t)(h0 + q) * R;
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
        a[mi] = (double)pre[p][mi];
        pre[p][mi] = __ldg(pp[mi]);
        pp[mi] += kstep;
      }
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
        for (int ni = 0; ni < NG; ni++) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], (double)pre[p][mi], b[ni]);
    }
  }
  // first maximum of the tile per chain: larger score wins, ties go to the smaller row.
  // acc[mi][ni][i] = score[row = base + 8 mi + g][column 8 ni + 2 q + i]
#pragma unroll
  for (int ni = 0; ni < NG; ni++) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int c = 8 * ni + 2 * q + i;
      const bool live = c < nt;
      const int slot = live ? (int)__fns(mask, 0, c + 1) : 0;
      const int w0 = live ? sh.row0[slot] : 0, w1 = live ? min(sh.row1[slot], tile_end) : 0;
      double b = -CUDART_INF;
      int r = 0x7fffffff;
#pragma unroll
      for (int mi = 0; mi < 4; mi++) {
        const int row = base + 8 * mi + g;
        if (row >= w0 && row < w1 && acc[mi][ni][i] > b) { b = acc[mi][ni][i]; r = row; }   // rows ascend with mi
      }
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {         // lanes with the same q hold the same column
        const double ob = __shfl_xor_sync(0xffffffffu, b, o);
        const int orow = __shfl_xor_sync(0xffffffffu, r, o);
        if (ob > b || (ob == b && orow < r)) { b = ob; r = orow; }
      }
      if (live && g == 0) {
        ScanPartial p;
        p.score = b; p.row = r; p.pad = 0;
        partials[sh.pboff[slot] + t] = p;
      }
    }
  }
}

template <typename TT>
__device__ __forceinline__ void scan_tile(const ClShared& sh, const UtrDev& u, const TT* __restrict__ A,
                                          const double* Vs, ScanPartial* partials, int t, int cntp,
                                          double* scan_elems) {
  const int lane = threadIdx.x & 31;
  const int base = t * CL_TILE;
  const int tile_end = min(base + CL_TILE, u.T * u.B);
  const bool cover = lane < cntp && sh.row0[lane] < tile_end && sh.row1[lane] > base;
  const unsigned mask = __ballot_sync(0xffffffffu, cover);
  const int nt = __popc(mask);
  if (nt == 0) return;
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
  if (NG == 1) scan_tile_mma<1, TT>(sh, u, A, Vs, partials, t, cntp, mask, nt, h0, h1);
  else if (NG == 2) scan_tile_mma<2, TT>(sh, u, A, Vs, partials, t, cntp, mask, nt, h0, h1);
  else if (NG == 3) scan_tile_mma<3, TT>(sh, u, A, Vs, partials, t, cntp, mask, nt, h0, h1);
  else scan_tile_mma<4, TT>(sh, u, A, Vs, partials, t, cntp, mask, nt, h0, h1);
}

template <typename TT, bool PF>
__device__ __noinline__ void cluster_estep(ChainDev& ch, ScanDesc& sd, const UtrDev& u, const TT* __restrict__ A,
                                           const double* __restrict__ cnt, double* __restrict__ lz,
                                           double* __restrict__ V) {
  switch (ch.K) {
    case 1: estep_warp_run<2, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 2: estep_warp_run<3, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 3: estep_warp_run<4, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 4: estep_warp_run<5, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 5: estep_warp_run<6, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 6: estep_warp_run<7, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 7: estep_warp_run<8, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 8: estep_warp_run<9, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 9: estep_warp_run<10, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 10: estep_warp_run<11, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 11: estep_warp_run<12, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 12: estep_warp_run<13, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 13: estep_warp_run<14, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 14: estep_warp_run<15, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    case 15: estep_warp_run<16, TT, PF>(ch, sd, u, A, cnt, lz, V); break;
    default: break;
  }
}

template <typename TT>
__global__ void __launch_bounds__(GT, 2)
em_cluster_kernel(const ClusterJob* __restrict__ jobs, ChainDev* chains, ScanDesc* descs,
                  const UtrDev* __restrict__ utrs, const void* __restrict__ tensor,
                  const double* __restrict__ cnt_all, double* lz_all, double* v_all, ScanPartial* partials,
                  int32_t* trace_a, int32_t* trace_b, double* trace_ws, double* scan_elems) {
  extern __shared__ double Vs[];
  __shared__ ClShared sh;
  cg::cluster_group cl = cg::this_cluster();
  const int C = (int)cl.num_blocks();
  const int crank = (int)cl.block_rank();
  const ClusterJob job = jobs[blockIdx.x / C];
  const UtrDev u = utrs[job.utr];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = GW * C, gw = crank * GW + warp;          // warps of the cluster, this warp's rank among them
  const TT* A = (const TT*)tensor + u.tensor_off;
  const double* cnt = cnt_all + u.frag_off;

  for (int round = 0; round <= SCAPE_B200_NROUND + 1; round++) {
    // ---- E phase: chain j of the UTR belongs to warp j mod W for the whole run ---------------------
    for (int j = gw; j < job.n_chains; j += W) {
      const int ci = job.chain_begin + j;
      ChainDev& gch = chains[ci];
      if (gch.state == 0) continue;                       // warp-uniform; only this warp writes the chain
      ChainDev& ch = sh.ch[warp];
      copy_chain(&ch, &gch, lane, 32);
      __syncwarp();
      ScanDesc& sd = descs[ci];
      if (apply_pending<CL_TILE>(ch, sd, u, partials, trace_a, trace_b, trace_ws)) {
        double* lz = lz_all + ch.lz_off;
        double* V = v_all + ch.v_off;
        bool again;
        do {
          cluster_estep<TT, true>(ch, sd, u, A, cnt, lz, V);
          __syncwarp();
          if (lane == 0 && ch.weights_only && ch.trace_off >= 0) {   // weights-only chains never wait for a scan
            const int64_t o = ch.trace_off + (int64_t)(ch.n_iter - 1) * (SCAPE_B200_KCAP + 1);
            for (int k = 0; k < ch.K; k++) { trace_a[o + k] = ch.a_idx[k]; trace_b[o + k] = ch.b_idx[k]; }
            for (int k = 0; k <= ch.K; k++) trace_ws[o + k] = ch.ws[k];
          }
          __syncwarp();
          again = ch.weights_only && ch.state == 1 && ch.n_iter < SCAPE_B200_NROUND;
        } while (again);
      }
      __syncwarp();
      copy_chain(&gch, &ch, lane, 32);
      __syncwarp();
    }
    __threadfence();
    cl.sync();                                            // V, windows, hulls of every chain of the UTR are visible

    // ---- the chains with a pending arg-max, ascending (identical in every CTA of the cluster) -------
    int pend = 0;
    if (tid < job.n_chains) pend = __ldcg(&descs[job.chain_begin + tid].pending);
    const unsigned pm = __ballot_sync(0xffffffffu, pend != 0);
    if (lane == 0) sh.warp_cnt[warp] = __popc(pm);
    __syncthreads();
    int before = 0, n_list = 0;
#pragma unroll
    for (int w = 0; w < GW; w++) {
      const int c = sh.warp_cnt[w];
      if (w < warp) before += c;
      n_list += c;
    }
    if (pend) sh.list[before + __popc(pm & ((1u << lane) - 1u))] = tid;
    if (n_list == 0) break;                               // uniform over the cluster: every CTA read the same flags

    // ---- scan: passes of <= cpp chains whose whole V rows sit in shared memory -----------------------
    for (int p0 = 0; p0 < n_list; p0 += job.cpp) {
      const int cntp = min(job.cpp, n_list - p0);
      __syncthreads();                                    // list complete / previous pass fully consumed
      if (warp == 0) {
        int h0 = 1 << 30, h1 = 0, t0 = 1 << 30, t1 = -1;
        if (lane < cntp) {
          const ScanDesc* dp = descs + job.chain_begin + sh.list[p0 + lane];
          const int r0 = __ldcg(&dp->row0), r1 = __ldcg(&dp->row1);
          const int a = __ldcg(&dp->hlo), b = __ldcg(&dp->hhi);
          sh.row0[lane] = r0; sh.row1[lane] = r1; sh.hlo[lane] = a; sh.hhi[lane] = b;
          sh.voff[lane] = __ldcg(&dp->v_off);
          sh.pboff[lane] = __ldcg(&dp->pb_off);
          if (b >= 0) { h0 = a; h1 = b + 1; }
          t0 = r0 / CL_TILE;
          t1 = (r1 - 1) / CL_TILE;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          h0 = min(h0, __shfl_xor_sync(0xffffffffu, h0, o));
          h1 = max(h1, __shfl_xor_sync(0xffffffffu, h1, o));
          t0 = min(t0, __shfl_xor_sync(0xffffffffu, t0, o));
          t1 = max(t1, __shfl_xor_sync(0xffffffffu, t1, o));
        }
        if (lane == 0) {
          if (h1 <= h0) { h0 = 0; h1 = 0; }
          sh.NA = h0 & ~7;
          sh.NB = h1;
          sh.P = cluster_v_pitch(h1 - (h0 & ~7));
          sh.tile_lo = t0;
          sh.tile_hi = t1;
        }
      }
      __syncthreads();
      const int NA = sh.NA, NB = sh.NB, P = sh.P;
      for (int j = warp; j <= cntp; j += GW) {            // row cntp = zeros (MMA columns without a chain)
        double* dst = Vs + j * P;
        const double* src = v_all + (j < cntp ? sh.voff[j] : 0) + NA;
        for (int o = lane; o < P; o += 32) dst[o] = (j < cntp && NA + o < NB) ? __ldcg(src + o) : 0.0;
      }
      __syncthreads();
      for (int t = sh.tile_lo + gw; t <= sh.tile_hi; t += W) scan_tile<TT>(sh, u, A, Vs, partials, t, cntp, scan_elems);
    }
    __threadfence();
    cl.sync();                                            // partials visible to the chains' owner warps
  }
}

cudaError_t launch_em_cluster(const ClusterJob* jobs_dev, int n_jobs, int cluster_size, ChainDev* chains_dev,
                              ScanDesc* descs_dev, const UtrDev* utrs_dev, const void* tensor, bool f32,
                              const double* cnt, double* lz, double* vbuf, void* partials, double* scan_elems,
                              int32_t* trace_a, int32_t* trace_b, double* trace_ws, cudaStream_t st) {
  if (n_jobs <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(n_jobs * cluster_size));
  cfg.blockDim = dim3(GT);
  cfg.dynamicSmemBytes = kClusterVBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ScanPartial* pb = (ScanPartial*)partials;
  cudaError_t e;
  if (f32) {
    e = cudaFuncSetAttribute(em_cluster_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kClusterVBytes);
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, em_cluster_kernel<float>, jobs_dev, chains_dev, descs_dev, utrs_dev, tensor, cnt, lz,
                              vbuf, pb, trace_a, trace_b, trace_ws, scan_elems);
  }
  e = cudaFuncSetAttribute(em_cluster_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, kClusterVBytes);
  if (e != cudaSuccess) return e;
  return cudaLaunchKernelEx(&cfg, em_cluster_kernel<double>, jobs_dev, chains_dev, descs_dev, utrs_dev, tensor, cnt, lz,
                            vbuf, pb, trace_a, trace_b, trace_ws, scan_elems);
}

}  // namespace scape
