// Bit-exact replay of the numpy *legacy* global RNG (np.random.seed / RandomState, MT19937) for
// exactly the calls the reference's initialisation makes (apa_core.py:125, 655-677, 781-829):
//   seed(int), uniform(size=n) / rand(n), choice(a, size, replace=False, p), choice(n, size,
//   replace=False), choice(arr, size, replace=True), permutation(n), shuffle(arr).
// numpy is a third-party dependency of the reference; its legacy stream is frozen by NEP 19, and
// the algorithms restated here are the published ones (numpy/random/mtrand.pyx `choice`,
// `shuffle`, `permutation`; src/legacy/legacy-distributions.c `legacy_random_interval`;
// src/mt19937/mt19937.c).  tests/test_host_rng.py checks every entry point draw-for-draw against
// the numpy installed in the image.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

namespace scape {

struct NpRandomState {
  static constexpr int N = 624, M = 397;
  uint32_t key[N];
  int pos;

  explicit NpRandomState(uint32_t seed = 1) { this->seed(seed); }

  // mt19937_seed(): Knuth's LCG initialisation used for integer seeds.
  void seed(uint32_t s) {
    for (int i = 0; i < N; i++) {
      key[i] = s;
      s = 1812433253u * (s ^ (s >> 30)) + uint32_t(i) + 1u;
    }
    pos = N;
  }

  void refill() {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
    int i;
    uint32_t y;
    for (i = 0; i < N - M; i++) {
      y = (key[i] & UPPER) | (key[i + 1] & LOWER);
      key[i] = key[i + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);
    }
    for (; i < N - 1; i++) {
      y = (key[i] & UPPER) | (key[i + 1] & LOWER);
      key[i] = key[i + (M - N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);
    }
    y = (key[N - 1] & UPPER) | (key[0] & LOWER);
    key[N - 1] = key[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);
    pos = 0;
  }

  inline uint32_t next_u32() {
    if (pos == N) refill();
    return temper(key[pos++]);
  }

  // mt19937_next_double(): 53-bit double from two 32-bit draws.
  inline double next_double() {
    int32_t a = int32_t(next_u32() >> 5), b = int32_t(next_u32() >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
  }

  // legacy random_interval(max): masked rejection, 32-bit draws while max fits.
  inline uint64_t interval(uint64_t max) {
    if (max == 0) return 0;
    uint64_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
    mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    uint64_t v;
    if (max <= 0xffffffffull) {
      while ((v = (next_u32() & mask)) > max) {}
    } else {
      while ((v = (((uint64_t(next_u32()) << 32) | next_u32()) & mask)) > max) {}
    }
    return v;
  }

  static inline uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }

  // RandomState.shuffle on a 1-d array: Fisher-Yates from the top, j = random_interval(i) per step.
  // This is the hot loop of the chain initialisation (choice(L, size, replace=False) shuffles all L
  // positions of the UTR), so the masked rejection is written without an unpredictable branch: a
  // rejected draw swaps a[i] with itself and leaves i where it is; the mask (smallest 2^k - 1 >= i)
  // is carried along instead of being rebuilt per step.  Same draws, same swaps, same final state.
  template <class T>
  void shuffle(T* a, int64_t n) {
    int64_t i = n - 1;
    if (i < 1) return;
    if (uint64_t(i) > 0xffffffffull) {              // 64-bit draws: not on any path of the reference, kept exact
      for (; i >= 1; i--) std::swap(a[i], a[int64_t(interval(uint64_t(i)))]);
      return;
    }
    uint32_t mask = uint32_t(i);
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    // phase 1: the accepted draws j_i for i = n-1 .. 1 (register-only loop-carried state); tiny arrays
    // (gen_k_arr shuffles K entries ~11 times per chain) keep them on the stack
    uint32_t small[64];
    uint32_t* jp = small;
    if (n > 64) {
      static thread_local std::vector<uint32_t> js;
      js.resize(size_t(n));
      jp = js.data();
    }
    while (i >= 1) {
      if (pos == N) refill();
      int p = pos;
      while (p < N && i >= 1) {
        const uint32_t lo = mask >> 1;              // the mask holds while i is in (lo, mask]
        while (p < N && uint32_t(i) > lo) {
          const uint32_t v = temper(key[p++]) & mask;
          jp[i] = v;                                // overwritten until a draw is accepted for this i
          i -= (v <= uint32_t(i));
        }
        if (uint32_t(i) <= lo) mask = lo;
      }
      pos = p;
    }
    // phase 2: the swaps, in the same order
    for (i = n - 1; i >= 1; i--) {
      const uint32_t j = jp[i];
      const T t = a[i];
      a[i] = a[j];
      a[j] = t;
    }
  }

  // RandomState.permutation(int n)
  void permutation(int64_t n, std::vector<int64_t>& out) {
    out.resize(size_t(n));
    std::iota(out.begin(), out.end(), int64_t(0));
    shuffle(out.data(), n);
  }

  // RandomState.randint(0, n) for one value (default int64 dtype: masked rejection on 32 bits).
  inline int64_t randint_below(int64_t n) {
    uint64_t rng = uint64_t(n - 1);
    if (rng == 0) return 0;
    return int64_t(interval(rng));  // same masked-rejection loop as _bounded_uint64 (use_masked)
  }

  // RandomState.choice(pop, size, replace=False, p=p) -> indices into the population.
  void choice_weighted_noreplace(const double* p, int64_t pop, int64_t size, std::vector<int64_t>& found) {
    static thread_local std::vector<double> pw, cdf, x;      // scratch reused across calls (hot path)
    static thread_local std::vector<int64_t> fresh;
    pw.assign(p, p + pop);
    cdf.assign(size_t(pop), 0.0);
    found.assign(size_t(size), 0);
    int64_t n_uniq = 0;
    while (n_uniq < size) {
      int64_t m = size - n_uniq;
      x.resize(size_t(m));
      for (int64_t i = 0; i < m; i++) x[size_t(i)] = next_double();
      for (int64_t i = 0; i < n_uniq; i++) pw[size_t(found[size_t(i)])] = 0.0;
      double run = 0.0;
      for (int64_t i = 0; i < pop; i++) { run += pw[size_t(i)]; cdf[size_t(i)] = run; }  // np.cumsum
      double last = cdf[size_t(pop - 1)];
      for (int64_t i = 0; i < pop; i++) cdf[size_t(i)] /= last;
      fresh.clear();
      for (int64_t i = 0; i < m; i++) {
        // searchsorted(side='right'): first index with cdf > x
        int64_t idx;
        if (pop <= 32) {                 // few peaks: count the entries <= x (same answer, no branches)
          idx = 0;
          const double xi = x[size_t(i)];
          const double* cp = cdf.data();
          for (int64_t q = 0; q < pop; q++) idx += (cp[q] <= xi);
        } else {
          idx = int64_t(std::upper_bound(cdf.begin(), cdf.end(), x[size_t(i)]) - cdf.begin());
        }
        // keep first occurrences, in draw order (np.unique(return_index) + sort + take)
        if (std::find(fresh.begin(), fresh.end(), idx) == fresh.end()) fresh.push_back(idx);
      }
      for (size_t i = 0; i < fresh.size(); i++) found[size_t(n_uniq) + i] = fresh[i];
      n_uniq += int64_t(fresh.size());
    }
  }
};

}  // namespace scape
