// sm_100a CUDA kernels of the infer_pa path.  All arithmetic is FP64 like the reference
// (taichi_core.py:11 `default_fp=ti.f64`).  The grid arg-max, batched over the chains of a UTR, is a
// dense FP64 product and runs on the FP64 tensor cores (mma.sync m8n8k4 -- FP64 has no tcgen05 path);
// everything else is CUDA-core FP64.
//
//   table_kernel    K2  loglik_xlr_t for every (fragment, theta)         apa_core.py:620-640, taichi_core.py:101-157
//   tensor_kernel   K3  marginal log-likelihood tensor[t][b][n]          taichi_core.py:160-179, 218-246
//   em_estep_{warp,,group}_kernel + em_scan_kernel  K4  bulk-synchronous EM iterations    apa_core.py:473-573, 702-779
//   label_kernel    K5  full E-step + row arg-max                        apa_core.py:873-881
//
// HBM layout (per wave of UTRs, one arena):
//   frag columns x,l,r,pa,cnt  double[sum N]            CSR by UtrDev.frag_off
//   theta grids                double[sum T]            CSR by UtrDev.theta_off
//   table                      double[sum T*Npad]       [t][n]  (transposed w.r.t. the reference so that n is contiguous)
//   tensor                     float|double[sum N*T*B]     [n][t][b]: for one fragment the R = T*B candidate rows
//                              (alpha-major, beta-minor) are contiguous, so the grid search (thread <-> row)
//                              reads it perfectly coalesced;
//                              FP32 storage by default (values are computed in FP64 and rounded once; the
//                              sentinel is exactly float's lowest), FP64 storage on request
//   log_zmat scratch           double[sum_chains (K+1)*Npad]  [k][n]
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <functional>
#include <vector>

#include "../../include/scape_b200.h"

#include <cstdlib>
#include <cstring>
// development-aid switches: on when the variable is set to anything but "" or "0"
inline bool scape_env_on(const char* name) {
  const char* s = getenv(name);
  return s && *s && strcmp(s, "0") != 0;
}

namespace scape {

#define SCAPE_SENTINEL (-3.4028234663852886e38)
#define SCAPE_PI 3.141592653589793

struct UtrDev {
  int32_t N, Npad, T, B;
  int32_t n_reads, pad_;          // raw reads of the UTR (label expansion)
  int64_t read_off;               // into the wave's read -> fragment map and per-read label buffer
  int64_t ldR;                    // pitch of the tensor: T*B candidate rows rounded up to 4 (16-byte aligned TMA rows)
  int64_t frag_off;
  int64_t theta_off;
  int64_t table_off;
  int64_t tensor_off;
  double unif_loglik;
};

struct ModelConst {
  double mu_f, sigma_f, max_unif_ws;
  int32_t n_s, n_beta;
  double s_dis[SCAPE_B200_MAX_S];
  double pmf_s[SCAPE_B200_MAX_S];
  double logpmf_s[SCAPE_B200_MAX_S];
  double betas[SCAPE_B200_MAX_BETA];
  double expc[18];                // constants of exp_nonpos (em_device.cuh): uploaded, so that they are c[bank][offset] operands
};

struct ChainDev {
  int32_t utr;
  int32_t K;
  int32_t weights_only;
  int32_t n_iter;                 // in/out: iterations done
  int64_t lz_off;                 // into the log_zmat scratch
  int64_t v_off;                  // this chain's row of V (fragment pitch = UTR's N rounded up to 8)
  int64_t trace_off;              // into the trace buffers, or -1
  int64_t pb_off;                 // into the scan partials: one ScanPartial per row block of the UTR
  int32_t a_idx[SCAPE_B200_KCAP];
  int32_t b_idx[SCAPE_B200_KCAP];
  double ws[SCAPE_B200_KCAP + 1];
  uint8_t k_order[SCAPE_B200_NROUND + 6];
  double bic;                     // out
  double lb_arr[SCAPE_B200_NROUND];  // out
  double grid_rows;               // out: sum over iterations of candidate rows scanned (W*B)
  double grid_rows_head;          // out: grid_rows when a resident kernel took the chain over (0 if it ran the whole chain)
  // device working state
  double lw[SCAPE_B200_KCAP + 1];
  double lb_prev, last_a;
  int32_t state;                  // 0 finished, 1 running, 2 converged (finishes once its last arg-max is applied)
  int32_t pending;                // the scan of this step must cover this chain
  int32_t cur_k, row0, row1, hlo, hhi, trace_pending;
  int32_t error, pad_;            // out: 1 = the grid search saw only non-finite scores (the chain is abandoned)
};

// What the scan needs to know about a chain, compact (the E step writes it, scan CTAs read 50 of them)
struct ScanDesc {
  int32_t pending, row0, row1, hlo, hhi, pad;
  int64_t v_off, pb_off;
};

// (UTR, block of candidate rows, share of the chain sub-batches) work item of the scan kernel.
// The CTA lists the chains that need the block, cuts the list into sub-batches of `gb` chains and
// takes sub-batches sb, sb + nsb, sb + 2 nsb ...  nsb = 1 means one CTA does them all.  Big UTRs
// (long fragment loop, few row blocks) are cut finer by the host so that one step's work spreads
// over all SMs instead of a handful (the step time is the slowest CTA).
struct ScanRef {
  int32_t utr;
  int32_t blk;
  int16_t sb, nsb, gb;
  int16_t pad;                    // 1: tile path (whole V rows of a sub-batch staged; per-warp chain masks and hulls)
};

struct LabelDev {
  int32_t utr;
  int32_t K;
  int64_t out_off;                // into the label buffer (per fragment)
  int32_t chain;                  // >= 0: take K, alpha, beta, ws from this record of the device chain array (a refit that has just run)
  int32_t pad_;
  int32_t a_idx[SCAPE_B200_KCAP];
  int32_t b_idx[SCAPE_B200_KCAP];
  double ws[SCAPE_B200_KCAP + 1];
};

// (utr, theta index) rows of a wave, flattened for the table / tensor launches
struct RowRef {
  int32_t utr;
  int32_t t;
};

// tile of consecutive interior alpha rows for the fast marginal kernel
struct TileRef {
  int32_t utr;
  int32_t i0;
  int32_t cnt;
};
constexpr int kTfB = 13, kTfW = 43, kTfHalf = 21, kTfTile = 8;
// half widths (grid points) of the 13 beta windows of the default grid (betas 5..65 step 5, theta step 9:
// floor(3 beta / 9), taichi_core.py:221-222): the constant-weight kernel skips the zero weights at compile time
__host__ __device__ constexpr int tf_default_hw(int j) {
  return j == 0 ? 1 : j == 1 ? 3 : j == 2 ? 5 : j == 3 ? 6 : j == 4 ? 8 : j == 5 ? 10 : j == 6 ? 11 : j == 7 ? 13 :
         j == 8 ? 15 : j == 9 ? 16 : j == 10 ? 18 : j == 11 ? 20 : 21;
}
cudaError_t upload_tensor_fast_tables(const double* g, const double* lp, const double* lps, const int* hw);
void launch_tensor_interior(const UtrDev* utrs, const TileRef* tiles, int64_t n_tiles, int max_n, const double* table,
                            void* tensor, bool f32, cudaStream_t st);
void launch_table(const UtrDev* utrs, const RowRef* rows, int64_t n_rows, int max_n, const double* fx,
                  const double* fl, const double* fr, const double* fpa, const double* theta, double* table,
                  cudaStream_t st);
void launch_tensor(const UtrDev* utrs, const RowRef* rows, int64_t n_rows, int max_n, int n_beta, int max_win,
                   const double* theta, const double* table, void* tensor, bool f32, cudaStream_t st);
void launch_labels(const LabelDev* jobs, int64_t n_jobs, int max_n, const UtrDev* utrs, const void* tensor, bool f32,
                   const double* cnt, const ChainDev* chains, int32_t* labels, cudaStream_t st);
// label_arr = labels of the fragments expanded by idx_arr (apa_core.py:976): one thread per read
void launch_label_expand(const LabelDev* jobs, int64_t n_jobs, int max_reads, const UtrDev* utrs, const int32_t* labels,
                         const int32_t* read_to_bin, int64_t* labels_per_read, cudaStream_t st);
cudaError_t upload_model_const(const ModelConst& mc);
constexpr int kTensorSlackRows = 64;   // zeroed fragment rows after the last UTR's tensor: the scan's register ring prefetches past the hull
constexpr int kScanRows = 256;    // candidate rows per scan CTA (must equal SCAN_ROWS in kernels.cu)
constexpr int kScanMaxChains = 160;  // running chains of one UTR a scan CTA can list (must equal SCAN_MAXCH)
constexpr int kPartialBytes = 16; // sizeof(ScanPartial)
// returns the number of kernel launches made
// CUDA events bracketing every launch group of one EM run (kept per lane, reused)
struct EmStepEvents {
  std::vector<cudaEvent_t> evs;
  std::vector<int> kinds;
  int scan_launches = 0;
  // called once, on the host, after the launches of step `hook_step` have been issued (the wave
  // scheduler uses it to slot the next wave's likelihood phase under the thinly filled late steps)
  std::function<void()> hook;
  // recorded on the EM stream where the staged wave's work may start (cheap; `hook` does the host work)
  std::function<void()> mark;
  int hook_step = 0;
  // record an event around every launch group (E-step / scan split of em_ms).  Off in the pipelined
  // production passes: on this platform an API call costs ~15-20 us, and two extra calls per step make
  // the 101-step loop launch-bound; on for the un-pipelined pass bench.py takes its per-kernel times from.
  bool timing = true;
};
void em_steps_elapsed(const EmStepEvents& ee, double* estep_ms, double* scan_ms);
int measure_fp64_peaks(int n_sm, double* dfma_tflops, double* dmma_tflops, cudaStream_t st);
int measure_sfu_peaks(int n_sm, double* out4, cudaStream_t st);   // FP32 FMA TFLOP/s, MUFU ex2 Gop/s, FP64 exp Gop/s, FP64 log Gop/s
constexpr int kWarpEstepMaxN = 384;    // UTRs with at most this many fragments use the warp-per-chain E step (a warp needs ~4 us per 32 fragments: above this the launch is bound by its longest warp)
// How the E step is launched.  Runs of weights-only chains (prune refits) always use the group
// kernel (g_small warps per chain for the chains with few fragments, a whole CTA for the others;
// persistent CTAs over device-side lists of running chains; every chain iterates to convergence
// inside one launch).  Runs with a grid search use the warp-per-chain / CTA-per-chain kernels over
// the full chain index (group_steps = true switches them to the group kernel too; measured slower).
// lists = 2 * n_chains ints, counts = 2 * (NROUND + 2) ints.
struct EstepPlan {
  int32_t* lists = nullptr;
  int32_t* counts = nullptr;
  int g_small = 4;
  int n_sm = 148;
  bool group_steps = false;
  bool warp_prefetch = true;
  int stage_chain = 1;          // E-step kernels work on a shared-memory copy of the chain record
  int warp_wpc = 1;             // warps per chain of the warp E-step kernel (1, 2 or 4)
  // wide steps launch two independent E-step kernels (warp per chain / CTA per chain): with st_big set,
  // the CTA-per-chain kernel goes to that stream (forked after the previous scan, joined before the
  // next), so that a step's E time is the longer of the two latency floors instead of their sum
  cudaStream_t st_big = nullptr;
  cudaEvent_t ev_big[2] = {nullptr, nullptr};
};
// ---- cluster-resident EM (em_cluster.cu) -----------------------------------------------------------
// One thread-block cluster per UTR runs every chain of the UTR through all its EM iterations inside
// ONE launch: E pass (warp per chain) -> cluster barrier -> grid arg-max on the FP64 tensor cores
// (32-row tiles x the chains whose window covers the tile) -> cluster barrier -> next iteration.
// A UTR never waits for another UTR; its tensor, log_zmat and V stay in L2 for the whole run.
struct ClusterJob {
  int32_t utr;           // index into the wave's UtrDev array
  int32_t chain_begin;   // first chain of the UTR in the run's chain array (chains are ordered by UTR)
  int32_t n_chains;      // <= kScanMaxChains
  int32_t cpp;           // chains whose V rows fit the kernel's shared-memory budget per pass (1..32)
};
constexpr int kClusterTileRows = 32;            // candidate rows per scan task = rows per partial
constexpr int kClusterVBytes = 80 * 1024;       // dynamic shared memory per CTA: staged V rows (2 CTAs per SM)
constexpr int kClusterPassMax = 32;             // chains per pass (one ballot)
// pitch (in doubles) of a staged V row that holds `len` fragments: multiple of 4, = 4 mod 16
inline __host__ __device__ int cluster_v_pitch(int len) {
  const int len4 = (len + 3) & ~3;
  return len4 + ((20 - (len4 & 15)) & 15);
}
constexpr int kScanTileVBytes = 92 * 1024;      // em_scan_kernel's dynamic shared memory (tile path: staged V rows; 2 CTAs per SM)
// chains per sub-batch of em_scan_kernel's tile path for a UTR with N fragments (multiple of 8), 0 = does not fit
inline int scan_tile_chains(int N) {
  const int rows = int(kScanTileVBytes / sizeof(double)) / cluster_v_pitch(N + 8) - 1;   // one row of zeros
  return rows < 8 ? 0 : (rows >= 32 ? 32 : rows / 8 * 8);
}
// chains per pass for a UTR with N fragments, 0 = does not fit (use the bulk-synchronous kernels)
inline int cluster_chains_per_pass(int N) {
  const int rows = int(kClusterVBytes / sizeof(double)) / cluster_v_pitch(N + 8) - 1;   // one row of zeros
  return rows < 1 ? 0 : (rows > kClusterPassMax ? kClusterPassMax : rows);
}
cudaError_t launch_em_cluster(const ClusterJob* jobs_dev, int n_jobs, int cluster_size, ChainDev* chains_dev,
                              ScanDesc* descs_dev, const UtrDev* utrs_dev, const void* tensor, bool f32,
                              const double* cnt, double* lz, double* vbuf, void* partials, double* scan_elems,
                              int32_t* trace_a, int32_t* trace_b, double* trace_ws, long long* stats, int solo_max,
                              cudaStream_t st);

int launch_em_steps(ChainDev* chains_dev, ScanDesc* descs_dev, const int32_t* index_dev, int64_t n_small, int64_t n_big, bool any_scan,
                    bool big_k,
                    const ScanRef* refs_dev, int64_t n_refs, int64_t n_refs_tile, const UtrDev* utrs_dev,
                    const int32_t* utr_chain_off_dev, const void* tensor, bool f32, const double* cnt, double* lz,
                    double* vbuf, void* partials, double* scan_elems, int32_t* trace_a, int32_t* trace_b,
                    double* trace_ws, cudaStream_t st, EmStepEvents& ee, const EstepPlan& plan,
                    int n_steps = SCAPE_B200_NROUND + 1);
// n_steps = SCAPE_B200_NROUND + 1: the whole run (50 x {E step, scan} + the closing E step that applies
// the last arg-max); n_steps = S <= NROUND: the first S x {E step, scan} only -- the chains then wait
// for their S-th arg-max to be applied by whoever continues them (em_tail.cu).

// ---- chain-resident EM (em_tail.cu): one CTA per chain, all remaining iterations in one launch ---------
cudaError_t launch_em_tail(ChainDev* chains_dev, ScanDesc* descs_dev, const int32_t* list_dev, int first, int n,
                           int max_n, int prows, const UtrDev* utrs_dev, const void* tensor, bool f32,
                           const double* cnt, double* lz, const void* partials, int32_t* trace_a, int32_t* trace_b,
                           double* trace_ws, unsigned long long* stats, cudaStream_t st);

}  // namespace scape
