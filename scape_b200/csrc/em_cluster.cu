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
constexpr int CL_MAXCH = SCAN_MAXCH;

struct ClShared {
  ChainDev ch[GW];                      // chain record of the chain a warp group is working on (E phase)
  EGroupShared grp[GW];                 // reduction scratch of the E-phase warp groups
  int list[CL_MAXCH];                   // pending chains of the UTR (index within the UTR), ascending
  int warp_cnt[GW];
  PassCtx px;                           // chains of the current pass
  int tile_lo, tile_hi;
  // what the last E pass of every pending chain left for the scan (indexed by chain within the UTR)
  int d_row0[CL_MAXCH], d_row1[CL_MAXCH], d_hlo[CL_MAXCH], d_hhi[CL_MAXCH];
  long long d_voff[CL_MAXCH], d_pboff[CL_MAXCH];
  int owner[CL_MAXCH];                  // after the split: the CTA a remaining chain belongs to
};

template <typename TT>
__device__ __forceinline__ void cluster_estep(EGroupShared& gs, int G, int gid, int tig, ChainDev& ch, ScanDesc& sd,
                                              const UtrDev& u, const TT* __restrict__ A, const double* __restrict__ cnt,
                                              double* __restrict__ lz, double* __restrict__ V) {
  switch (ch.K) {
    case 1: estep_group_run<2, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 2: estep_group_run<3, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 3: estep_group_run<4, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 4: estep_group_run<5, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 5: estep_group_run<6, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 6: estep_group_run<7, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 7: estep_group_run<8, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 8: estep_group_run<9, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 9: estep_group_run<10, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 10: estep_group_run<11, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 11: estep_group_run<12, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 12: estep_group_run<13, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 13: estep_group_run<14, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 14: estep_group_run<15, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    case 15: estep_group_run<16, TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V); break;
    default: break;
  }
}

template <typename TT>
__global__ void __launch_bounds__(GT, 2)
em_cluster_kernel(const ClusterJob* __restrict__ jobs, ChainDev* chains, ScanDesc* descs,
                  const UtrDev* __restrict__ utrs, const void* __restrict__ tensor,
                  const double* __restrict__ cnt_all, double* lz_all, double* v_all, ScanPartial* partials,
                  int32_t* trace_a, int32_t* trace_b, double* trace_ws, double* scan_elems, long long* stats,
                  int solo_max) {
  extern __shared__ double Vs[];
  __shared__ ClShared sh;
  cg::cluster_group cl = cg::this_cluster();
  const int C = (int)cl.num_blocks();
  const int crank = (int)cl.block_rank();
  const ClusterJob job = jobs[blockIdx.x / C];
  const UtrDev u = utrs[job.utr];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // warps that share the UTR's work, this warp's rank among them.  The cluster shrinks to its first
  // CTA ("solo") once few chains are left: the others exit and free their SM slots for the next UTR.
  int W = GW * C, gw = crank * GW + warp;
  bool solo = C == 1, split = false;     // split: the chains left were dealt to the CTAs (sh.owner), every CTA on its own
  const TT* A = (const TT*)tensor + u.tensor_off;
  const double* cnt = cnt_all + u.frag_off;

  // development aid (SCAPE_B200_DBG): cycles CTA 0 / thread 0 of the cluster spends per phase
  long long t_e = 0, t_w1 = 0, t_scan = 0, t_w2 = 0, t_mark = 0, t_begin = 0, t_solo = 0, t_e_early = 0, t_scan_early = 0;
  int n_rounds = 0, solo_round = -1;
  const bool prof = stats != nullptr && crank == 0 && tid == 0;
  if (prof) t_begin = t_mark = clock64();
#define CL_LAP(acc) do { if (prof) { const long long now__ = clock64(); acc += now__ - t_mark; t_mark = now__; } } while (0)
#define CL_SYNC() do { if (solo) __syncthreads(); else cl.sync(); } while (0)

  // round 0: every chain of the UTR takes an E pass
  int n_list = job.n_chains;
  if (tid < n_list) sh.list[tid] = tid;
  __syncthreads();

  for (int round = 0; round <= SCAPE_B200_NROUND + 1; round++) {
    // ---- E phase: G warps per chain, the i-th chain of the list goes to group i mod (groups of the cluster).
    // A fragment pass of one warp is ~800 dependent instructions (~2.5 us): with few chains left the
    // round's latency is the E pass, so the fragments of a chain are spread over up to 8 warps.  G
    // depends on the number of listed chains only (not on the cluster size): a chain's sums do not
    // depend on how the UTR was scheduled.
    {
      const int G = n_list <= 8 ? 8 : n_list <= 16 ? 4 : n_list <= 32 ? 2 : 1;
      const int gthreads = 32 * G, gpc = GW / G;
      const int gid = warp / G, tig = tid - gid * gthreads;
      const int n_groups = gpc * (solo ? 1 : C), ggid = (solo ? 0 : crank) * gpc + gid;
      EGroupShared& gs = sh.grp[gid];
      ChainDev& ch = sh.ch[gid];
      for (int i = ggid; i < n_list; i += n_groups) {
        const int ci = job.chain_begin + sh.list[i];
        ChainDev& gch = chains[ci];
        copy_chain(&ch, &gch, tig, gthreads);
        group_sync(gid, gthreads);
        if (ch.state != 0) {                               // group-uniform
          ScanDesc& sd = descs[ci];
          if (tig < 32) {
            const int go = apply_pending<CL_TILE>(ch, sd, u, partials, trace_a, trace_b, trace_ws);
            if (tig == 0) gs.go = go;
          }
          group_sync(gid, gthreads);
          if (gs.go) {
            double* lz = lz_all + ch.lz_off;
            double* V = v_all + ch.v_off;
            bool again;
            do {
              cluster_estep<TT>(gs, G, gid, tig, ch, sd, u, A, cnt, lz, V);
              if (tig == 0 && ch.weights_only && ch.trace_off >= 0) {   // weights-only chains never wait for a scan
                const int64_t o = ch.trace_off + (int64_t)(ch.n_iter - 1) * (SCAPE_B200_KCAP + 1);
                for (int k = 0; k < ch.K; k++) { trace_a[o + k] = ch.a_idx[k]; trace_b[o + k] = ch.b_idx[k]; }
                for (int k = 0; k <= ch.K; k++) trace_ws[o + k] = ch.ws[k];
              }
              group_sync(gid, gthreads);                   // the epilogue's writes to the chain are visible to the group
              again = ch.weights_only && ch.state == 1 && ch.n_iter < SCAPE_B200_NROUND;
            } while (again);
          }
          copy_chain(&gch, &ch, tig, gthreads);
        }
        group_sync(gid, gthreads);                         // ch / gs are reused by the group's next chain
      }
    }
    __threadfence();
    if (round < 12) CL_LAP(t_e_early); else CL_LAP(t_e);
    CL_SYNC();                                            // V, windows, hulls of every chain of the UTR are visible
    CL_LAP(t_w1);
    n_rounds++;

    // ---- the chains with a pending arg-max, ascending (identical in every CTA of the cluster) -------
    int pend = 0;
    if (tid < job.n_chains && (!split || sh.owner[tid] == crank)) {
      const ScanDesc* dp = descs + job.chain_begin + tid;
      pend = __ldcg(&dp->pending);
      if (pend) {
        sh.d_row0[tid] = __ldcg(&dp->row0); sh.d_row1[tid] = __ldcg(&dp->row1);
        sh.d_hlo[tid] = __ldcg(&dp->hlo); sh.d_hhi[tid] = __ldcg(&dp->hhi);
        sh.d_voff[tid] = __ldcg(&dp->v_off); sh.d_pboff[tid] = __ldcg(&dp->pb_off);
      }
    }
    const unsigned pm = __ballot_sync(0xffffffffu, pend != 0);
    if (lane == 0) sh.warp_cnt[warp] = __popc(pm);
    __syncthreads();
    int before = 0;
    n_list = 0;
#pragma unroll
    for (int w = 0; w < GW; w++) {
      const int c = sh.warp_cnt[w];
      if (w < warp) before += c;
      n_list += c;
    }
    if (pend) sh.list[before + __popc(pm & ((1u << lane) - 1u))] = tid;
    if (n_list == 0) break;                               // uniform over the cluster: every CTA read the same flags
    if (!solo && n_list <= solo_max) {
      // Few chains left.  Chains are independent of one another (they were batched for the MMA's sake
      // only), and a round of one or two chains is pure latency: the i-th remaining chain moves to CTA
      // i mod C for good and every CTA finishes its own chains alone -- no cluster barriers, a CTA
      // without a chain exits and frees its SM slot for the next UTR.
      __syncthreads();                                    // sh.list complete
      for (int i = tid; i < n_list; i += GT) sh.owner[sh.list[i]] = i % C;
      __syncthreads();
      int mine = 0;
      for (int i = crank; i < n_list; i += C) mine++;
      if (mine == 0) return;                              // uniform over the CTA
      if (tid < mine) pend = sh.list[crank + tid * C]; // (reuse `pend` as a scratch register)
      __syncthreads();
      if (tid < mine) sh.list[tid] = pend;
      n_list = mine;
      solo = true;
      split = true;
      W = GW;
      gw = warp;
      if (prof) { solo_round = n_rounds; t_solo = clock64(); }
    }

    // ---- scan: passes of <= cpp chains whose whole V rows sit in shared memory -----------------------
    for (int p0 = 0; p0 < n_list; p0 += job.cpp) {
      const int cntp = min(job.cpp, n_list - p0);
      __syncthreads();                                    // list complete / previous pass fully consumed
      if (warp == 0) {
        int h0 = 1 << 30, h1 = 0, t0 = 1 << 30, t1 = -1;
        if (lane < cntp) {
          const int j = sh.list[p0 + lane];
          const int r0 = sh.d_row0[j], r1 = sh.d_row1[j], a = sh.d_hlo[j], b = sh.d_hhi[j];
          sh.px.row0[lane] = r0; sh.px.row1[lane] = r1; sh.px.hlo[lane] = a; sh.px.hhi[lane] = b;
          sh.px.voff[lane] = sh.d_voff[j];
          sh.px.pboff[lane] = sh.d_pboff[j];
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
          sh.px.NA = h0 & ~7;
          sh.px.NB = h1;
          sh.px.P = cluster_v_pitch(h1 - (h0 & ~7));
          sh.tile_lo = t0;
          sh.tile_hi = t1;
        }
      }
      __syncthreads();
      const int NA = sh.px.NA, NB = sh.px.NB, P = sh.px.P;
      for (int j = warp; j <= cntp; j += GW) {            // row cntp = zeros (MMA columns without a chain)
        double* dst = Vs + j * P;
        const double* src = v_all + (j < cntp ? sh.px.voff[j] : 0) + NA;
        for (int o = lane; o < P; o += 32) dst[o] = (j < cntp && NA + o < NB) ? __ldcg(src + o) : 0.0;
      }
      __syncthreads();
      for (int t = sh.tile_lo + gw; t <= sh.tile_hi; t += W) scan_tile<TT, false>(sh.px, u, A, Vs, partials, t, cntp, scan_elems, nullptr, nullptr);
    }
    __threadfence();
    if (round < 12) CL_LAP(t_scan_early); else CL_LAP(t_scan);
    CL_SYNC();                                            // partials visible to the chains' owner warps
    CL_LAP(t_w2);
  }
  if (prof) {
    long long* o = stats + (blockIdx.x / C) * 10;
    const long long t_end = clock64();
    o[0] = n_rounds; o[1] = t_end - t_begin; o[2] = t_e; o[3] = t_w1; o[4] = t_scan; o[5] = t_w2;
    o[6] = solo_round; o[7] = solo_round >= 0 ? t_end - t_solo : 0; o[8] = t_e_early; o[9] = t_scan_early;
  }
#undef CL_LAP
#undef CL_SYNC
}

cudaError_t launch_em_cluster(const ClusterJob* jobs_dev, int n_jobs, int cluster_size, ChainDev* chains_dev,
                              ScanDesc* descs_dev, const UtrDev* utrs_dev, const void* tensor, bool f32,
                              const double* cnt, double* lz, double* vbuf, void* partials, double* scan_elems,
                              int32_t* trace_a, int32_t* trace_b, double* trace_ws, long long* stats, int solo_max,
                              cudaStream_t st) {
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
                              vbuf, pb, trace_a, trace_b, trace_ws, scan_elems, stats, solo_max);
  }
  e = cudaFuncSetAttribute(em_cluster_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, kClusterVBytes);
  if (e != cudaSuccess) return e;
  return cudaLaunchKernelEx(&cfg, em_cluster_kernel<double>, jobs_dev, chains_dev, descs_dev, utrs_dev, tensor, cnt, lz,
                            vbuf, pb, trace_a, trace_b, trace_ws, scan_elems, stats, solo_max);
}

}  // namespace scape
