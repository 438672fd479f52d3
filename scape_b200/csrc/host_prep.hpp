// Host pre-pass of the infer_pa path, in native code:
//   read binning                         apa_core.py:285-327   (bin_data)
//   grids, read-type split, uniform lik  apa_core.py:365-452, 576-584
//   coverage profile + smoothing         apa_core.py:454-462, 681-700
//   profile peaks                        apa_core.py:784-794 + scipy.signal.find_peaks(distance=)
//   random chain initialisation          apa_core.py:655-677, 781-829 (numpy legacy RNG replay)
// Everything here is RNG-free except `draw_chain` / `draw_refit`, which consume the per-file
// stream in the reference's exact call order.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <string>
#include <vector>

#include "../../include/scape_b200.h"
#include "np_rng.hpp"

namespace scape {

constexpr double kSentinel = -3.4028234663852886e38;  // float(np.finfo('f').min), apa_core.py:428

enum : int32_t {
  kOk = 0,
  kErrReadStart = -1,   // assert 0 <= x < utr_length failed (apa_core.py:388)
  kErrBinRange = -2,    // bin label does not fit the packed sort key
  kErrEmpty = -3,       // UTR without reads
  kErrKcap = -4,        // re-run would need more than SCAPE_B200_KCAP components
  kErrParams = -5,      // n_min_apa > n_max_apa, max_beta < beta_step ... (apa_core.py:931-937)
  kErrNoPeakMass = -6,  // fewer non-zero peak weights than K (numpy choice raises ValueError)
};

// numpy's pairwise summation for n <= 128 contiguous doubles (numpy/_core/src/umath/loops_utils.h.src,
// DOUBLE_pairwise_sum); np.sum(a) on a fresh 1-d array is 0 + pairwise(a).
inline double np_pairwise_sum(const double* a, int64_t n) {
  if (n < 8) {
    double res = 0.;
    for (int64_t i = 0; i < n; i++) res += a[i];
    return res;
  }
  if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int64_t i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
  }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// np.argsort(priority) as the *installed numpy* does it.  scipy's _select_by_peak_distance ranks
// peaks with an unstable np.argsort, whose order among exactly equal heights depends on the
// numpy build / CPU dispatch (AVX-512 sorting networks on this image).  Equal heights are common
// (isolated noise reads of equal length give identical bumps), so when -- and only when -- two
// candidate peaks tie, the host binding is asked for the order.  Without a callback: stable sort.
typedef void (*argsort_fn)(const double* values, int64_t n, int64_t* order_out);
inline argsort_fn& argsort_callback() {
  static argsort_fn fn = nullptr;
  return fn;
}

struct UtrPrep {
  int32_t status = kOk;
  int64_t n_reads = 0;
  // fragments (bins), lexicographic in (x, l, r, pa) bin label like np.unique(axis=0)
  std::vector<double> x, l, r, pa, cnt;
  std::vector<int32_t> read_to_bin;
  int64_t L = 0;
  double min_theta = 0;
  std::vector<double> theta_full;  // arange(int(min_theta), int(L), theta_step)
  std::vector<double> theta;       // grid the EM runs on (== theta_full except in fixed mode)
  std::vector<double> betas;
  double unif_loglik = 0;
  std::vector<double> prof_y;      // smoothed, padded coverage (length L + 200)
  std::vector<int64_t> peak_idx;   // indices into prof_y; position = idx - 100
  std::vector<double> peak_w;
  int64_t n() const { return int64_t(cnt.size()); }
  int64_t T() const { return int64_t(theta.size()); }
  int64_t B() const { return int64_t(betas.size()); }
};

// ---- binning ---------------------------------------------------------------------------------
inline int64_t bin_label(double v, double step) {
  // np.digitize(v, arange(0, step + max, step)) with NaN -> -1 -> label 0 (apa_core.py:296-299)
  if (!(v >= 0.0)) return 0;
  return int64_t(std::floor(v / step)) + 1;
}

inline int32_t bin_reads(const double* x, const double* l, const double* r, const double* pa, int64_t n,
                         UtrPrep& u) {
  u.n_reads = n;
  if (n <= 0) return u.status = kErrEmpty;
  struct Rec { uint64_t key; int32_t idx; };
  std::vector<Rec> recs(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; i++) {
    int64_t bx = bin_label(x[i], 5), bl = bin_label(l[i], 10), br = bin_label(r[i], 10), bp = bin_label(pa[i], 5);
    if (bx >= (1 << 20) || bl >= (1 << 12) || br >= (1 << 12) || bp >= (1 << 20)) return u.status = kErrBinRange;
    recs[size_t(i)] = {(uint64_t(bx) << 44) | (uint64_t(bl) << 32) | (uint64_t(br) << 20) | uint64_t(bp), int32_t(i)};
  }
  // Order by key.  The leading field is the x bin (a few hundred distinct values, one or two reads
  // each), so a counting sort on it followed by insertion sorts inside the x bins does the job in
  // O(n + range); a comparison sort is the fallback for absurd coordinate ranges.
  {
    uint64_t max_bx = 0;
    for (const Rec& rc : recs) max_bx = std::max(max_bx, rc.key >> 44);
    if (max_bx < uint64_t(4 * n + 4096)) {
      static thread_local std::vector<int32_t> start;
      static thread_local std::vector<Rec> sorted;
      start.assign(size_t(max_bx) + 2, 0);
      for (const Rec& rc : recs) start[size_t(rc.key >> 44) + 1]++;
      for (size_t b = 1; b < start.size(); b++) start[b] += start[b - 1];
      sorted.resize(recs.size());
      for (const Rec& rc : recs) sorted[size_t(start[size_t(rc.key >> 44)]++)] = rc;   // start[b] becomes the END of bin b
      size_t lo = 0;
      for (size_t b = 0; b + 1 < start.size(); b++) {
        const size_t hi = size_t(start[b]);
        if (hi - lo > 24) {                             // crowded x bin (deep UTRs): comparison sort
          std::sort(sorted.begin() + long(lo), sorted.begin() + long(hi), [](const Rec& a, const Rec& b) { return a.key < b.key; });
          lo = hi;
          continue;
        }
        for (size_t a = lo + 1; a < hi; a++) {          // insertion sort of one x bin by the full key
          const Rec v = sorted[a];
          size_t q = a;
          while (q > lo && sorted[q - 1].key > v.key) { sorted[q] = sorted[q - 1]; q--; }
          sorted[q] = v;
        }
        lo = hi;
      }
      recs.swap(sorted);
    } else {
      std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) { return a.key < b.key; });
    }
  }
  u.read_to_bin.assign(size_t(n), 0);
  u.x.clear(); u.l.clear(); u.r.clear(); u.pa.clear(); u.cnt.clear();
  size_t i = 0;
  while (i < recs.size()) {
    size_t j = i;
    double sx = 0, sl = 0, sr = 0, sp = 0;
    int32_t bin = int32_t(u.cnt.size());
    while (j < recs.size() && recs[j].key == recs[i].key) {
      int32_t k = recs[j].idx;
      sx += x[k]; sl += l[k]; sr += r[k]; sp += pa[k];   // NaN bins stay NaN (np.bincount weights)
      u.read_to_bin[size_t(k)] = bin;
      j++;
    }
    double c = double(j - i);
    u.x.push_back(sx / c); u.l.push_back(sl / c); u.r.push_back(sr / c); u.pa.push_back(sp / c);
    u.cnt.push_back(c);
    i = j;
  }
  return kOk;
}

// ---- grids -----------------------------------------------------------------------------------
inline std::vector<double> int_arange(int64_t lo, int64_t hi, int64_t step) {
  std::vector<double> v;
  for (int64_t t = lo; t < hi; t += step) v.push_back(double(t));
  return v;
}

// find_nearest (apa_core.py:537-549) for one value: nearest grid point, ties go up.
// `step` > 0: the grid is known to be g[0] + t * step (the full theta grid), so searchsorted's
// position is guessed arithmetically and only corrected by comparisons against the grid itself.
inline int64_t snap_to_grid(const std::vector<double>& g, double v, double step = 0.0) {
  int64_t p;
  if (step > 0.0 && !g.empty()) {
    const int64_t n = int64_t(g.size());
    const double q = std::ceil((v - g[0]) / step);
    p = q <= 0.0 ? 0 : (q >= double(n) ? n : int64_t(q));
    while (p > 0 && g[size_t(p - 1)] >= v) p--;          // lower bound: first index with g[p] >= v
    while (p < n && g[size_t(p)] < v) p++;
  } else {
    p = int64_t(std::lower_bound(g.begin(), g.end(), v) - g.begin());
  }
  if (p == 0) return 0;
  if (p == int64_t(g.size())) return int64_t(g.size()) - 1;
  return (v - g[size_t(p - 1)] >= g[size_t(p)] - v) ? p : p - 1;
}

inline int32_t setup_model(const scape_b200_params& P, const double* x_raw, const double* l_raw, int64_t n_raw,
                           UtrPrep& u) {
  if (u.status != kOk) return u.status;
  if (P.n_min_apa > P.n_max_apa || P.max_beta < P.beta_step) return u.status = kErrParams;
  // subsample_run (apa_core.py:995-997, 1004)
  double mx = x_raw[0], ml = l_raw[0];
  for (int64_t i = 1; i < n_raw; i++) { mx = std::max(mx, x_raw[i]); ml = std::max(ml, l_raw[i]); }
  int64_t utr_len = std::max<int64_t>(P.utr_length, int64_t(mx) + int64_t(ml) + 50);
  if (P.fixed_run_mode) utr_len = std::max<int64_t>(utr_len, P.pre_L);
  u.L = utr_len > 2000 ? utr_len : 2000;                                   // apa_core.py:387
  for (double v : u.x)
    if (!(v >= 0 && v < double(utr_len))) return u.status = kErrReadStart;  // apa_core.py:388
  double lmin = u.l[0];
  for (double v : u.l) lmin = std::min(lmin, v);
  u.min_theta = double(int64_t(lmin));                                     // apa_core.py:407
  u.theta_full = int_arange(int64_t(u.min_theta), u.L, P.theta_step);      // apa_core.py:409 / 940
  u.unif_loglik = std::log(1.0 / double(u.L) * (1.0 / double(u.L)) * (1.0 / P.max_LA));  // apa_core.py:576-584
  u.betas.assign(P.betas, P.betas + P.n_beta);
  if (!P.fixed_run_mode) {
    u.theta = u.theta_full;
  } else {
    // fixed_run (apa_core.py:888-896): union of [alpha - 3 b_max, alpha + 3 b_max) grid slices
    double b_hi = P.pre_beta[0];
    for (int k = 1; k < P.pre_K; k++) b_hi = std::max(b_hi, P.pre_beta[k]);
    std::vector<char> take(u.theta_full.size(), 0);
    for (int k = 0; k < P.pre_K; k++) {
      int64_t a = snap_to_grid(u.theta_full, P.pre_alpha[k] - 3 * b_hi);
      int64_t b = snap_to_grid(u.theta_full, P.pre_alpha[k] + 3 * b_hi);
      for (int64_t t = a; t < b; t++) take[size_t(t)] = 1;
    }
    u.theta.clear();
    for (size_t t = 0; t < take.size(); t++)
      if (take[t]) u.theta.push_back(u.theta_full[t]);
  }
  return kOk;
}

// ---- coverage profile, smoothing, peaks --------------------------------------------------------
// ker_smooth for the positions whose window is not clipped by the ends of the profile, four
// positions at a time.  Every position runs numpy's pairwise summation for 8 <= n <= 128 exactly as
// np_pairwise_sum does (8 running lanes over blocks of 8, the fixed combine tree, the tail added in
// order); the SIMD lanes are four neighbouring POSITIONS, so the order of operations of each sum is
// untouched.  Products are rounded before they are added (no FMA: the targets below have none).
typedef double v4d_t __attribute__((vector_size(32)));
__attribute__((always_inline)) inline void smooth_interior_impl(const double* w, int64_t nw, double wsum, const double* y,
                                                                int64_t half, int64_t lo, int64_t hi, double* out) {
  const int64_t nb = nw - (nw % 8);
  for (int64_t i = lo; i + 4 <= hi; i += 4) {
    const double* yy = y + (i - half);
    v4d_t r[8];
    for (int j = 0; j < 8; j++) {
      v4d_t v;
      __builtin_memcpy(&v, yy + j, sizeof(v));
      const v4d_t wv = {w[j], w[j], w[j], w[j]};
      r[j] = wv * v;
    }
    for (int64_t b = 8; b < nb; b += 8)
      for (int j = 0; j < 8; j++) {
        v4d_t v;
        __builtin_memcpy(&v, yy + b + j, sizeof(v));
        const v4d_t wv = {w[b + j], w[b + j], w[b + j], w[b + j]};
        const v4d_t prod = wv * v;
        r[j] = r[j] + prod;
      }
    v4d_t res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (int64_t t = nb; t < nw; t++) {
      v4d_t v;
      __builtin_memcpy(&v, yy + t, sizeof(v));
      const v4d_t wv = {w[t], w[t], w[t], w[t]};
      const v4d_t prod = wv * v;
      res = res + prod;
    }
    const v4d_t den = {wsum, wsum, wsum, wsum};
    res = res / den;
    __builtin_memcpy(out + i, &res, sizeof(res));
  }
}
__attribute__((target("avx2"))) inline void smooth_interior_avx2(const double* w, int64_t nw, double wsum, const double* y,
                                                                 int64_t half, int64_t lo, int64_t hi, double* out) {
  smooth_interior_impl(w, nw, wsum, y, half, lo, hi, out);
}
inline void smooth_interior_base(const double* w, int64_t nw, double wsum, const double* y, int64_t half, int64_t lo,
                                 int64_t hi, double* out) {
  smooth_interior_impl(w, nw, wsum, y, half, lo, hi, out);
}

inline void coverage_and_peaks(const scape_b200_params& P, UtrPrep& u) {
  if (u.status != kOk) return;
  const int64_t L = u.L, ny = L + 200;
  // coverage_cnt[int(x) : int(x)+int(l)] += cnt  (integer-valued, so a difference array is exact)
  std::vector<double> y(size_t(ny) + 1, 0.0);
  for (int64_t i = 0; i < u.n(); i++) {
    int64_t a = int64_t(u.x[size_t(i)]), len = int64_t(u.l[size_t(i)]);
    if (len <= 0) continue;
    int64_t b = std::min<int64_t>(a + len, L);
    y[size_t(100 + a)] += u.cnt[size_t(i)];
    y[size_t(100 + b)] -= u.cnt[size_t(i)];
  }
  double run = 0;
  for (int64_t i = 0; i < ny; i++) { run += y[size_t(i)]; y[size_t(i)] = run; }
  y.resize(size_t(ny));
  // ker_smooth (apa_core.py:681-700) with numpy's pairwise summation order
  const int64_t nw = P.n_smooth, half = (nw - 1) / 2;
  const double* w = P.smooth_w;
  const double wsum = np_pairwise_sum(w, nw);
  u.prof_y.assign(size_t(ny), 0.0);
  std::vector<double> tmp(static_cast<size_t>(nw));
  // prefix count of non-zero coverage so all-zero windows are skipped (0/wsum == 0 exactly)
  std::vector<int32_t> nzp(size_t(ny) + 1, 0);
  for (int64_t i = 0; i < ny; i++) nzp[size_t(i + 1)] = nzp[size_t(i)] + (y[size_t(i)] != 0.0);
  // interior positions (full window) in blocks of four; the ends and the remainder below
  int64_t blk_lo = half, blk_hi = half;
  if (nw >= 8 && nw <= 128 && ny - half > half + 4) {
    blk_hi = half + (ny - half - half) / 4 * 4;
    static const bool has_avx2 = __builtin_cpu_supports("avx2") && !getenv("SCAPE_B200_NO_AVX2");
    if (has_avx2) smooth_interior_avx2(w, nw, wsum, y.data(), half, blk_lo, blk_hi, u.prof_y.data());
    else smooth_interior_base(w, nw, wsum, y.data(), half, blk_lo, blk_hi, u.prof_y.data());
  }
  for (int64_t i = 0; i < ny; i++) {
    if (i == blk_lo && blk_hi > blk_lo) { i = blk_hi - 1; continue; }
    int64_t st = std::max<int64_t>(0, i - half), en = std::min<int64_t>(ny - 1, i + half);
    if (nzp[size_t(en + 1)] == nzp[size_t(st)]) continue;
    int64_t w0 = st - (i - half), m = en - st + 1;
    for (int64_t j = 0; j < m; j++) tmp[size_t(j)] = w[w0 + j] * y[size_t(st + j)];
    double num = np_pairwise_sum(tmp.data(), m);
    double den = (m == nw) ? wsum : np_pairwise_sum(w + w0, m);
    u.prof_y[size_t(i)] = num / den;
  }
  // scipy.signal._peak_finding_utils._local_maxima_1d (plateau midpoints)
  const std::vector<double>& s = u.prof_y;
  std::vector<int64_t> cand;
  {
    int64_t i = 1, imax = ny - 1;
    while (i < imax) {
      if (s[size_t(i - 1)] < s[size_t(i)]) {
        int64_t ahead = i + 1;
        while (ahead < imax && s[size_t(ahead)] == s[size_t(i)]) ahead++;
        if (s[size_t(ahead)] < s[size_t(i)]) {
          cand.push_back((i + ahead - 1) / 2);
          i = ahead;
        }
      }
      i++;
    }
  }
  // _select_by_peak_distance: highest first, drop neighbours closer than ceil(distance)
  const int64_t dist = int64_t(std::ceil(P.min_pa_gap));
  std::vector<int64_t> order(cand.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = int64_t(i);
  std::stable_sort(order.begin(), order.end(),
                   [&](int64_t a, int64_t b) { return s[size_t(cand[size_t(a)])] < s[size_t(cand[size_t(b)])]; });
  bool tie = false;
  for (size_t i = 1; i < order.size() && !tie; i++)
    tie = s[size_t(cand[size_t(order[i])])] == s[size_t(cand[size_t(order[i - 1])])];
  if (tie && argsort_callback()) {
    std::vector<double> pri(cand.size());
    for (size_t i = 0; i < cand.size(); i++) pri[i] = s[size_t(cand[i])];
    argsort_callback()(pri.data(), int64_t(pri.size()), order.data());
  }
  std::vector<char> keep(cand.size(), 1);
  for (int64_t i = int64_t(order.size()) - 1; i >= 0; i--) {
    int64_t j = order[size_t(i)];
    if (!keep[size_t(j)]) continue;
    for (int64_t k = j - 1; k >= 0 && cand[size_t(j)] - cand[size_t(k)] < dist; k--) keep[size_t(k)] = 0;
    for (int64_t k = j + 1; k < int64_t(cand.size()) && cand[size_t(k)] - cand[size_t(j)] < dist; k++) keep[size_t(k)] = 0;
  }
  u.peak_idx.clear();
  for (size_t i = 0; i < cand.size(); i++)
    if (keep[i]) u.peak_idx.push_back(cand[i]);
  // peak weights (apa_core.py:788-794): sequential Python sum over +-bw
  const int64_t bw = int64_t(P.beta_step) * 3;
  u.peak_w.assign(u.peak_idx.size(), 0.0);
  double tot = 0;
  for (size_t i = 0; i < u.peak_idx.size(); i++) {
    int64_t a = std::max<int64_t>(0, u.peak_idx[i] - bw), b = std::min<int64_t>(ny, u.peak_idx[i] + bw + 1);
    double acc = 0;
    for (int64_t j = a; j < b; j++) acc += s[size_t(j)];
    u.peak_w[i] = acc;
    tot += acc;
  }
  for (double& v : u.peak_w) v /= tot;
}

// ---- chain initialisation (RNG) ----------------------------------------------------------------
struct ChainInit {
  int32_t K = 0;
  int32_t a_idx[SCAPE_B200_KCAP] = {0};
  int32_t b_idx[SCAPE_B200_KCAP] = {0};
  double ws[SCAPE_B200_KCAP + 1] = {0};
  uint8_t k_order[SCAPE_B200_NROUND] = {0};
};

// init_ws (apa_core.py:809-815)
inline void draw_weights(NpRandomState& rng, int K, double cap, double* w) {
  for (int i = 0; i <= K; i++) w[i] = rng.next_double();
  double tot = 0;
  for (int i = 0; i <= K; i++) tot += w[i];         // Python sum(): sequential from 0
  for (int i = 0; i <= K; i++) w[i] = w[i] / tot;
  if (w[K] > cap) {
    for (int i = 0; i < K; i++) w[i] = w[i] * (1 - cap);   // NB: not renormalised (apa_core.py:812-814)
    w[K] = cap;
  }
}

// gen_k_arr (apa_core.py:655-677)
inline void draw_component_order(NpRandomState& rng, int K, uint8_t* out) {
  if (K <= 1) {
    for (int i = 0; i < SCAPE_B200_NROUND; i++) out[i] = 0;
    return;
  }
  static thread_local std::vector<int64_t> arr;
  rng.permutation(K, arr);
  int pos = 0;
  for (int i = 0; i < SCAPE_B200_NROUND; i++) {
    if (pos % K == 0) {
      rng.shuffle(arr.data(), K);
      pos = 0;
    }
    out[i] = uint8_t(arr[size_t(pos)]);
    pos++;
  }
}

// init_para + the gen_k_arr call at the top of em_algo (apa_core.py:817-829, 720)
inline int32_t draw_chain(NpRandomState& rng, const scape_b200_params& P, const UtrPrep& u, int K, ChainInit& c) {
  c.K = K;
  const int64_t n_peak = int64_t(u.peak_idx.size());
  double picked[SCAPE_B200_KCAP];
  if (K <= n_peak) {
    int64_t nz = 0;
    for (double v : u.peak_w) nz += (v > 0);
    if (nz < K) return kErrNoPeakMass;
    static thread_local std::vector<int64_t> found;
    rng.choice_weighted_noreplace(u.peak_w.data(), n_peak, K, found);
    for (int i = 0; i < K; i++) picked[size_t(i)] = double(u.peak_idx[size_t(found[size_t(i)])] - 100);
  } else {
    static thread_local std::vector<int64_t> perm;    // reused: L entries, every chain with K > n_peak
    rng.permutation(u.L, perm);                       // choice(L, size, replace=False) == permutation(L)[:size]
    for (int64_t i = 0; i < n_peak; i++) picked[size_t(i)] = double(u.peak_idx[size_t(i)] - 100);
    for (int64_t i = n_peak; i < K; i++) picked[size_t(i)] = double(perm[size_t(i - n_peak)]);
  }
  const double amp = double(5 * int64_t(P.beta_step));
  for (int i = 0; i < K; i++) {
    double uni = 0.0 + (1.0 - 0.0) * rng.next_double();
    picked[size_t(i)] += std::nearbyint(amp * (2 * uni - 1));
  }
  std::sort(picked, picked + K);
  const double grid_step = P.fixed_run_mode ? 0.0 : double(P.theta_step);   // fixed mode: irregular union of slices
  for (int i = 0; i < K; i++) c.a_idx[i] = int32_t(snap_to_grid(u.theta, picked[size_t(i)], grid_step));
  for (int i = 0; i < K; i++) c.b_idx[i] = int32_t(rng.randint_below(u.B()));
  draw_weights(rng, K, P.max_unif_ws, c.ws);
  draw_component_order(rng, K, c.k_order);
  return kOk;
}

// fixed_inference (apa_core.py:708-711): new weights, then em_algo's gen_k_arr
inline void draw_refit(NpRandomState& rng, const scape_b200_params& P, ChainInit& c) {
  draw_weights(rng, c.K, P.max_unif_ws, c.ws);
  draw_component_order(rng, c.K, c.k_order);
}

}  // namespace scape
