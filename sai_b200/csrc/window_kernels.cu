// K4: per-window U, N(Variants), exact Q and the candidate position lists (one launch).
//
// One warp owns one (job, window).  A window [start, end] (inclusive,
// sai/generators/window_generator.py:173-174) maps to the site range [lo, hi)
// by a cooperative 32-ary search on the sorted positions; its U count is the
// popcount of the tile masks over that range (sai/stats/u_statistic.py:94-96)
// and Q is the numpy 'linear' quantile of the flagged sites' target frequencies
// (sai/stats/q_statistic.py:96-101; numpy/lib/_function_base_impl.py _lerp).
// Order statistics are selected exactly on the float64 bit patterns (all
// values are non-negative, so unsigned order == numeric order).
//
// The kernel is issue-bound (see profiles/), so all site / tile indices are
// 32-bit (n_sites < 2^31 is checked on the host) and warp sums use REDUX.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"

namespace sai {

constexpr int kWinWarps = 4;
constexpr int kBufCap = 256;    // doubles buffered in shared memory per warp (more: global walk)
constexpr int kMaskCache = 64;  // tiles whose masks stay in shared memory between the two walks

struct WinParams {
  const int32_t* pos;
  int32_t n_sites;
  int32_t n_tiles;
  const int64_t* ws;
  const int64_t* we;
  const int32_t* seg_lo;  // optional: window i only sees the sites [seg_lo[i], seg_hi[i]) --
  const int32_t* seg_hi;  // several chromosomes (pieces) in one launch
  int32_t W;
  int32_t n_jobs;
  int32_t u_enabled[SAI_MAX_JOBS];
  int32_t q_enabled[SAI_MAX_JOBS];
  double quantile[SAI_MAX_JOBS];
  const uint32_t* mask_u;
  const uint32_t* mask_q;
  const double* qval;
  int64_t qval_stride;
  int32_t* nsnps;
  int64_t* u;
  double* q;
  int32_t* q_cnt;
  int64_t* u_start;  // [J][W] first candidate of the window inside the job's u_cand
  int64_t* q_start;
  unsigned long long* totals;  // [J][2] candidates reserved so far (U, Q)
  int32_t* u_cand;
  int64_t cap_u;
  int32_t* q_cand;
  int64_t cap_q;
};

__device__ __forceinline__ int clamp_key(int64_t k) {
  return k > 2147483647ll ? 2147483647 : (k < -2147483647ll ? -2147483647 : (int)k);
}

// Cooperative 32-ary lower bounds of two keys at once: first index whose
// position is >= key.  Each round probes 32 evenly spaced positions per key, so
// a 6 M-site chromosome needs 5 dependent rounds instead of 2 x 23 binary steps.
__device__ __forceinline__ void warp_lower_bound2(const int32_t* __restrict__ pos, int n,
                                                  int64_t key_a, int64_t key_b, int lane,
                                                  int& out_a, int& out_b) {
  // positions are int32: clamp the keys so that the comparisons can be 32-bit
  const int ka = clamp_key(key_a), kb = clamp_key(key_b);
  // a key beyond INT32_MAX is above every position
  int lo_a = key_a > 2147483647ll ? n : 0, hi_a = n;  // answer in [lo, hi]
  int lo_b = key_b > 2147483647ll ? n : 0, hi_b = n;
  while (hi_a > lo_a || hi_b > lo_b) {
    const int na = hi_a - lo_a, nb = hi_b - lo_b;
    const int sa = na > 32 ? (na + 31) >> 5 : 1, sb = nb > 32 ? (nb + 31) >> 5 : 1;
    const unsigned ia = (unsigned)lo_a + (unsigned)(lane + 1) * (unsigned)sa - 1u;
    const unsigned ib = (unsigned)lo_b + (unsigned)(lane + 1) * (unsigned)sb - 1u;
    const bool pa = ia < (unsigned)hi_a && __ldg(pos + ia) < ka;
    const bool pb = ib < (unsigned)hi_b && __ldg(pos + ib) < kb;
    const int ca = __popc(__ballot_sync(0xffffffffu, pa));  // sorted => the true lanes are a prefix
    const int cb = __popc(__ballot_sync(0xffffffffu, pb));
    if (na > 0) {
      // probes 0..ca-1 are < key, probe ca (if it exists) is >= key
      const unsigned nhi = (unsigned)lo_a + (unsigned)(ca + 1) * (unsigned)sa - 1u;
      lo_a += ca * sa;
      if (nhi < (unsigned)hi_a) hi_a = (int)nhi;
      if (lo_a > hi_a) lo_a = hi_a;
    }
    if (nb > 0) {
      const unsigned nhi = (unsigned)lo_b + (unsigned)(cb + 1) * (unsigned)sb - 1u;
      lo_b += cb * sb;
      if (nhi < (unsigned)hi_b) hi_b = (int)nhi;
      if (lo_b > hi_b) lo_b = hi_b;
    }
  }
  out_a = lo_a;
  out_b = lo_b;
}

__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_excl_scan(int v, int lane) {
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  return x - v;
}
__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
  unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
  unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long warp_min64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = shfl64(v, (threadIdx.x & 31) ^ o);
    v = w < v ? w : v;
  }
  return v;
}

// Site range of a window restricted to tile T: mask of the bits inside [lo, hi).
__device__ __forceinline__ uint32_t range_mask(int T, int lo, int hi) {
  const int base = T * kTile;
  uint32_t m = 0xffffffffu;
  if (lo > base) m <<= (lo - base);
  if (hi < base + kTile) m &= 0xffffffffu >> (base + kTile - hi);
  return m;
}

// Element source walking the Q mask of the window (values gathered from the
// dense qval array).  for_each calls f(key) once per flagged site, on some lane.
struct MaskSource {
  const uint32_t* mask;  // job's mask_q
  const double* qval;    // job's qval
  int lo, hi, T0, T1;
  int lane;
  template <typename F>
  __device__ __forceinline__ void for_each(F f) const {
    for (int T = T0 + lane; T <= T1; T += 32) {
      uint32_t m = __ldg(mask + T) & range_mask(T, lo, hi);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        f((unsigned long long)__double_as_longlong(__ldg(qval + (size_t)T * kTile + b)));
      }
    }
  }
};

struct BufSource {
  const unsigned long long* buf;
  int n;
  int lane;
  template <typename F>
  __device__ __forceinline__ void for_each(F f) const {
    for (int i = lane; i < n; i += 32) f(buf[i]);
  }
};

// Exact k-th smallest (0-based) of the source's keys by MSB-first radix select
// with a per-warp 256-bin shared histogram.  Also returns how many keys are
// <= the selected key.
template <typename Src>
__device__ unsigned long long warp_radix_select(const Src& src, int k, int* hist, int lane,
                                                int& count_le) {
  unsigned long long prefix = 0;
  int kk = k;
  int eq = 0;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
#pragma unroll
    for (int i = 0; i < 8; ++i) hist[lane * 8 + i] = 0;
    __syncwarp();
    src.for_each([&](unsigned long long key) {
      if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8)))
        atomicAdd(&hist[(int)((key >> shift) & 255ull)], 1);
    });
    __syncwarp();
    int c[8], local = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      c[i] = hist[lane * 8 + i];
      local += c[i];
    }
    const int excl = warp_excl_scan(local, lane);
    const bool mine = kk >= excl && kk < excl + local;
    int bin = 0, below = 0, cnt = 0;
    if (mine) {
      int run = excl;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (cnt == 0 && kk < run + c[i]) {
          bin = lane * 8 + i;
          below = run;
          cnt = c[i];
        }
        run += c[i];
      }
    }
    const unsigned who = __ballot_sync(0xffffffffu, mine);
    const int srcl = __ffs(who) - 1;
    bin = __shfl_sync(0xffffffffu, bin, srcl);
    below = __shfl_sync(0xffffffffu, below, srcl);
    cnt = __shfl_sync(0xffffffffu, cnt, srcl);
    prefix |= (unsigned long long)bin << shift;
    kk -= below;
    eq = cnt;
    __syncwarp();
    if (cnt == 1 && pass < 7) {
      // a single key carries this prefix: fetch it and stop
      unsigned long long found = ~0ull;
      src.for_each([&](unsigned long long key) {
        if ((key >> shift) == (prefix >> shift)) found = key;
      });
      prefix = warp_min64(found);
      break;
    }
  }
  count_le = (k - kk) + eq;
  return prefix;
}

template <typename Src>
__device__ unsigned long long warp_min_above(const Src& src, unsigned long long a) {
  unsigned long long best = ~0ull;
  src.for_each([&](unsigned long long key) {
    if (key > a && key < best) best = key;
  });
  return warp_min64(best);
}

template <typename Src>
__device__ int warp_count_ge(const Src& src, double thr) {
  int c = 0;
  src.for_each([&](unsigned long long key) { c += (__longlong_as_double((long long)key) >= thr); });
  return warp_sum(c);
}

// numpy 'linear' quantile from the two neighbouring order statistics
// (_get_indexes / _get_gamma / _lerp): separately rounded, no FMA.
__device__ __forceinline__ double lerp_numpy(double a, double b, double g) {
  const double d = __dsub_rn(b, a);
  if (g >= 0.5) return __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, g)));
  return __dadd_rn(a, __dmul_rn(d, g));
}

template <typename Src>
__device__ double warp_quantile(const Src& src, int n, double q, int* hist, int lane) {
  const double vi = __dmul_rn((double)(n - 1), q);
  int cle;
  if (vi >= (double)(n - 1)) {
    return __longlong_as_double((long long)warp_radix_select(src, n - 1, hist, lane, cle));
  }
  const double fl = floor(vi);
  const int k = (int)fl;
  const double g = __dsub_rn(vi, fl);
  const unsigned long long ka = warp_radix_select(src, k, hist, lane, cle);
  const unsigned long long kb = (k + 1 < cle) ? ka : warp_min_above(src, ka);
  return lerp_numpy(__longlong_as_double((long long)ka), __longlong_as_double((long long)kb), g);
}

// n <= 32 values, one per lane: ranks by all-pairs comparison over shuffles.
__device__ double warp_quantile_small(unsigned long long key, int n, double q, int lane) {
  int rank = 0;
  for (int j = 0; j < n; ++j) {
    const unsigned long long o = shfl64(key, j);
    rank += (o < key) || (o == key && j < lane);
  }
  const double vi = __dmul_rn((double)(n - 1), q);
  int k, k2;
  double g = 0.0;
  if (vi >= (double)(n - 1)) {
    k = k2 = n - 1;
  } else {
    const double fl = floor(vi);
    k = (int)fl;
    k2 = k + 1;
    g = __dsub_rn(vi, fl);
  }
  const unsigned ma = __ballot_sync(0xffffffffu, lane < n && rank == k);
  const unsigned mb = __ballot_sync(0xffffffffu, lane < n && rank == k2);
  const double a = __longlong_as_double((long long)shfl64(key, __ffs(ma) - 1));
  const double b = __longlong_as_double((long long)shfl64(key, __ffs(mb) - 1));
  if (k == k2) return a;
  return lerp_numpy(a, b, g);
}

// ---------------------------------------------------------------------------
// Per-warp scratch.  The batched fast path and the generic path never use it at
// the same time, so they share the bytes.
// ---------------------------------------------------------------------------
constexpr int kFastN = 32;   // flagged sites per window (U and Q each) on the fast path
constexpr int kFastTiles = 64;

template <int kBatch>
struct FastScratch {
  unsigned long long qv[kBatch][kFastN];  // flagged target frequencies (bit patterns)
  int32_t qpos[kBatch][kFastN];           // their positions
  int32_t upos[kBatch][kFastN];           // positions of the U-flagged sites
};
struct SlowScratch {
  unsigned long long buf[kBufCap];
  int hist[256];
  uint32_t mu[kMaskCache], mq[kMaskCache];
};
template <int kBatch>
union WarpScratch {
  FastScratch<kBatch> f;
  SlowScratch s;
};

struct JobView {
  const uint32_t* __restrict__ mu;
  const uint32_t* __restrict__ mq;
  const double* __restrict__ qv;
  int32_t* __restrict__ uc;
  int32_t* __restrict__ qc;
  bool want_u, want_q;
  double qq;
  int j;
};

// Generic path for one window whose site range [lo, hi) is known: any number of
// tiles and of flagged sites.
__device__ void window_generic(const WinParams& P, const JobView& J, SlowScratch& S, int i, int lo,
                               int hi, int lane) {
  const size_t item = (size_t)J.j * P.W + i;
  int u_cnt = 0, q_n = 0;
  int T0 = 0, T1 = -1;
  if (hi > lo) {
    T0 = lo >> 5;
    T1 = (hi - 1) >> 5;
    for (int Tb = T0; Tb <= T1; Tb += 32) {
      const int T = Tb + lane;
      uint32_t a = 0, b = 0;
      if (T <= T1) {
        const uint32_t rm = range_mask(T, lo, hi);
        if (J.want_u) a = __ldg(J.mu + T) & rm;
        if (J.want_q) b = __ldg(J.mq + T) & rm;
        if (T - T0 < kMaskCache) {
          S.mu[T - T0] = a;
          S.mq[T - T0] = b;
        }
      }
      u_cnt += __popc(a);
      if (__any_sync(0xffffffffu, b != 0)) {
        const int c = __popc(b);
        int w = q_n + warp_excl_scan(c, lane);
        q_n += warp_sum(c);
        while (b) {
          const int bit = __ffs(b) - 1;
          b &= b - 1;
          if (w < kBufCap)
            S.buf[w] = (unsigned long long)__double_as_longlong(__ldg(J.qv + (size_t)T * kTile + bit));
          ++w;
        }
      }
    }
    u_cnt = warp_sum(u_cnt);
    __syncwarp();
  }
  double qres = CUDART_NAN;
  int q_cand = 0;
  if (q_n > 0) {
    if (q_n <= 32) {
      const unsigned long long key = lane < q_n ? S.buf[lane] : ~0ull;
      qres = warp_quantile_small(key, q_n, J.qq, lane);
      const bool ge = lane < q_n && __longlong_as_double((long long)key) >= qres;
      q_cand = __popc(__ballot_sync(0xffffffffu, ge));
    } else if (q_n <= kBufCap) {
      BufSource src{S.buf, q_n, lane};
      qres = warp_quantile(src, q_n, J.qq, S.hist, lane);
      q_cand = warp_count_ge(src, qres);
    } else {
      MaskSource src{J.mq, J.qv, lo, hi, T0, T1, lane};
      qres = warp_quantile(src, q_n, J.qq, S.hist, lane);
      q_cand = warp_count_ge(src, qres);
    }
  }
  // reserve this window's slices of the candidate buffers (one atomic each)
  unsigned long long ub = 0, qb = 0;
  if (lane == 0) {
    if (u_cnt > 0) ub = atomicAdd(P.totals + 2 * J.j, (unsigned long long)u_cnt);
    if (q_cand > 0) qb = atomicAdd(P.totals + 2 * J.j + 1, (unsigned long long)q_cand);
    P.nsnps[item] = hi - lo;
    P.u[item] = u_cnt;
    P.q[item] = qres;
    P.q_cnt[item] = q_cand;
    P.u_start[item] = (int64_t)ub;
    P.q_start[item] = (int64_t)qb;
  }
  if (u_cnt > 0 || q_cand > 0) {
    ub = shfl64(ub, 0);
    qb = shfl64(qb, 0);
    // second walk: candidate positions in site order (u_statistic.py:95, q_statistic.py:101)
    long long uw = (long long)ub, qw = (long long)qb;
    for (int Tb = T0; Tb <= T1; Tb += 32) {
      const int T = Tb + lane;
      uint32_t a = 0, b = 0;
      if (T <= T1) {
        uint32_t m;
        if (T - T0 < kMaskCache) {
          a = S.mu[T - T0];
          m = S.mq[T - T0];
        } else {
          const uint32_t rm = range_mask(T, lo, hi);
          a = J.want_u ? (__ldg(J.mu + T) & rm) : 0u;
          m = J.want_q ? (__ldg(J.mq + T) & rm) : 0u;
        }
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          if (__ldg(J.qv + (size_t)T * kTile + bit) >= qres) b |= 1u << bit;
        }
      }
      if (__any_sync(0xffffffffu, a != 0)) {
        const int ca = __popc(a);
        long long wa = uw + warp_excl_scan(ca, lane);
        uw += warp_sum(ca);
        while (a) {
          const int bit = __ffs(a) - 1;
          a &= a - 1;
          if (wa < P.cap_u) J.uc[wa] = __ldg(P.pos + T * kTile + bit);
          ++wa;
        }
      }
      if (__any_sync(0xffffffffu, b != 0)) {
        const int cb = __popc(b);
        long long wb = qw + warp_excl_scan(cb, lane);
        qw += warp_sum(cb);
        while (b) {
          const int bit = __ffs(b) - 1;
          b &= b - 1;
          if (wb < P.cap_q) J.qc[wb] = __ldg(P.pos + T * kTile + bit);
          ++wb;
        }
      }
    }
  }
  __syncwarp();
}

// 32-ary lower bounds of 2*kBatch keys at once (same invariant as a single
// search: answer in [lo, hi]); every round has 2*kBatch loads in flight.
template <int kBatch>
__device__ __forceinline__ void warp_lower_bound_batch(const int32_t* __restrict__ pos,
                                                       const int (&first)[kBatch], const int (&last)[kBatch],
                                                       const int64_t (&key)[2 * kBatch], int lane,
                                                       int (&out)[2 * kBatch]) {
  int lo[2 * kBatch], hi[2 * kBatch], k32[2 * kBatch];
#pragma unroll
  for (int t = 0; t < 2 * kBatch; ++t) {
    k32[t] = clamp_key(key[t]);
    lo[t] = key[t] > 2147483647ll ? last[t >> 1] : first[t >> 1];  // search inside [first, last)
    hi[t] = last[t >> 1];
  }
  while (true) {
    bool any = false;
#pragma unroll
    for (int t = 0; t < 2 * kBatch; ++t) any = any || (hi[t] > lo[t]);
    if (!any) break;
    bool pred[2 * kBatch];
    int step[2 * kBatch];
#pragma unroll
    for (int t = 0; t < 2 * kBatch; ++t) {
      const int nn = hi[t] - lo[t];
      step[t] = nn > 32 ? (nn + 31) >> 5 : 1;
      const unsigned idx = (unsigned)lo[t] + (unsigned)(lane + 1) * (unsigned)step[t] - 1u;
      pred[t] = idx < (unsigned)hi[t] && __ldg(pos + idx) < k32[t];
    }
#pragma unroll
    for (int t = 0; t < 2 * kBatch; ++t) {
      const int c = __popc(__ballot_sync(0xffffffffu, pred[t]));
      if (hi[t] > lo[t]) {
        const unsigned nhi = (unsigned)lo[t] + (unsigned)(c + 1) * (unsigned)step[t] - 1u;
        lo[t] += c * step[t];
        if (nhi < (unsigned)hi[t]) hi[t] = (int)nhi;
        if (lo[t] > hi[t]) lo[t] = hi[t];
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 2 * kBatch; ++t) out[t] = lo[t];
}

// grid: x = window batches (kWinWarps * kBatch windows per block, grid-stride), y = job.
// Each warp works on kBatch consecutive windows at once so that every dependent
// memory round trip (window bounds, searches, masks, flagged values, slice
// reservation) is shared by kBatch windows; windows that do not fit the fast
// path (> 64 tiles or > 32 flagged sites) fall through to window_generic.
template <int kBatch, int kMinBlocks>
__global__ void __launch_bounds__(kWinWarps * 32, kMinBlocks)
    k_window_stats(const __grid_constant__ WinParams P) {
  __shared__ WarpScratch<kBatch> s_scratch[kWinWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  JobView J;
  J.j = blockIdx.y;
  J.mu = P.mask_u + (size_t)J.j * P.n_tiles;
  J.mq = P.mask_q + (size_t)J.j * P.n_tiles;
  J.qv = P.qval + (size_t)J.j * P.qval_stride;
  J.want_u = P.u_enabled[J.j] != 0;
  J.want_q = P.q_enabled[J.j] != 0;
  J.qq = P.quantile[J.j];
  J.uc = P.u_cand + (size_t)J.j * P.cap_u;
  J.qc = P.q_cand + (size_t)J.j * P.cap_q;
  WarpScratch<kBatch>& SC = s_scratch[warp];
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int i0 = (blockIdx.x * kWinWarps + warp) * kBatch; i0 < P.W; i0 += gridDim.x * kWinWarps * kBatch) {
    // ---- A: window bounds as keys (one coalesced load for the whole batch) ----
    int64_t key[2 * kBatch];
    {
      const int k = lane >> 1;  // lanes 0..2*kBatch-1: lane 2k -> start of window k, 2k+1 -> end+1
      int64_t mine = 0;
      if (lane < 2 * kBatch && i0 + k < P.W) mine = (lane & 1) ? P.we[i0 + k] + 1 : P.ws[i0 + k];
#pragma unroll
      for (int t = 0; t < 2 * kBatch; ++t) key[t] = (int64_t)shfl64((unsigned long long)mine, t);
    }
    // ---- B: all searches together (inside the window's own piece when pieces are given) ----
    int bnd[2 * kBatch];
    {
      int first[kBatch], last[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const bool seg = P.seg_lo != nullptr && i0 + k < P.W;
        first[k] = seg ? __ldg(P.seg_lo + i0 + k) : 0;
        last[k] = seg ? __ldg(P.seg_hi + i0 + k) : P.n_sites;
      }
      warp_lower_bound_batch<kBatch>(P.pos, first, last, key, lane, bnd);
    }
#pragma unroll
    for (int k = 0; k < kBatch; ++k)  // end < start: the reference's masks select nothing (window_generator.py:173-174)
      if (bnd[2 * k + 1] < bnd[2 * k]) bnd[2 * k + 1] = bnd[2 * k];

    // ---- C: masks of all windows ----
    uint32_t a[kBatch][2], b[kBatch][2];
    int T0[kBatch];
    bool valid[kBatch], fast[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int lo = bnd[2 * k], hi = bnd[2 * k + 1];
      valid[k] = i0 + k < P.W;
      T0[k] = lo >> 5;
      const int T1 = hi > lo ? (hi - 1) >> 5 : T0[k] - 1;
      fast[k] = valid[k] && (T1 - T0[k] + 1 <= kFastTiles);
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int T = T0[k] + it * 32 + lane;
        a[k][it] = b[k][it] = 0;
        if (fast[k] && T <= T1) {
          const uint32_t rm = range_mask(T, lo, hi);
          if (J.want_u) a[k][it] = __ldg(J.mu + T) & rm;
          if (J.want_q) b[k][it] = __ldg(J.mq + T) & rm;
        }
      }
    }
    int u_cnt[kBatch], q_n[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      u_cnt[k] = warp_sum(__popc(a[k][0]) + __popc(a[k][1]));
      q_n[k] = warp_sum(__popc(b[k][0]) + __popc(b[k][1]));
      fast[k] = fast[k] && u_cnt[k] <= kFastN && q_n[k] <= kFastN;
    }

    // ---- D: flagged values / positions into shared memory, loads issued before any store ----
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      if (!fast[k] || (u_cnt[k] == 0 && q_n[k] == 0)) continue;
      int wq[2], wu[2];
      {
        const int c0 = __popc(b[k][0]), c1 = __popc(b[k][1]);
        wq[0] = q_n[k] ? warp_excl_scan(c0, lane) : 0;
        wq[1] = q_n[k] ? warp_sum(c0) + warp_excl_scan(c1, lane) : 0;
        const int d0 = __popc(a[k][0]), d1 = __popc(a[k][1]);
        wu[0] = u_cnt[k] ? warp_excl_scan(d0, lane) : 0;
        wu[1] = u_cnt[k] ? warp_sum(d0) + warp_excl_scan(d1, lane) : 0;
      }
      unsigned long long vq[2] = {0ull, 0ull};
      int pq[2] = {0, 0}, pu[2] = {0, 0};
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int site0 = (T0[k] + it * 32 + lane) * kTile;
        if (b[k][it]) {
          const int bit = __ffs(b[k][it]) - 1;
          vq[it] = (unsigned long long)__double_as_longlong(__ldg(J.qv + (size_t)site0 + bit));
          pq[it] = __ldg(P.pos + site0 + bit);
        }
        if (a[k][it]) pu[it] = __ldg(P.pos + site0 + __ffs(a[k][it]) - 1);
      }
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int site0 = (T0[k] + it * 32 + lane) * kTile;
        if (b[k][it]) {
          SC.f.qv[k][wq[it]] = vq[it];
          SC.f.qpos[k][wq[it]] = pq[it];
          uint32_t m = b[k][it] & (b[k][it] - 1);  // further flagged sites of the same tile (rare)
          int w = wq[it] + 1;
          while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            SC.f.qv[k][w] = (unsigned long long)__double_as_longlong(__ldg(J.qv + (size_t)site0 + bit));
            SC.f.qpos[k][w] = __ldg(P.pos + site0 + bit);
            ++w;
          }
        }
        if (a[k][it]) {
          SC.f.upos[k][wu[it]] = pu[it];
          uint32_t m = a[k][it] & (a[k][it] - 1);
          int w = wu[it] + 1;
          while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            SC.f.upos[k][w] = __ldg(P.pos + site0 + bit);
            ++w;
          }
        }
      }
    }
    __syncwarp();

    // ---- E: quantiles (n <= 32: one value per lane) ----
    double qres[kBatch];
    unsigned ge[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      qres[k] = CUDART_NAN;
      ge[k] = 0;
      if (fast[k] && q_n[k] > 0) {
        const unsigned long long kv = lane < q_n[k] ? SC.f.qv[k][lane] : ~0ull;
        qres[k] = warp_quantile_small(kv, q_n[k], J.qq, lane);
        ge[k] = __ballot_sync(0xffffffffu, lane < q_n[k] && __longlong_as_double((long long)kv) >= qres[k]);
      }
    }
    // ---- F: reserve the candidate slices of all fast windows with one atomic instruction ----
    unsigned long long base = 0;
    {
      int want = 0;
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        if (lane == 2 * k && fast[k]) want = u_cnt[k];
        if (lane == 2 * k + 1 && fast[k]) want = __popc(ge[k]);
      }
      if (want > 0) base = atomicAdd(P.totals + 2 * J.j + (lane & 1), (unsigned long long)want);
    }
    // ---- G: results and candidate positions ----
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      if (!fast[k]) continue;
      const long long ub = (long long)shfl64(base, 2 * k), qb = (long long)shfl64(base, 2 * k + 1);
      const int qc_n = __popc(ge[k]);
      if (lane == 0) {
        const size_t item = (size_t)J.j * P.W + i0 + k;
        P.nsnps[item] = bnd[2 * k + 1] - bnd[2 * k];
        P.u[item] = u_cnt[k];
        P.q[item] = qres[k];
        P.q_cnt[item] = qc_n;
        P.u_start[item] = ub;
        P.q_start[item] = qb;
      }
      if (lane < u_cnt[k] && ub + lane < P.cap_u) J.uc[ub + lane] = SC.f.upos[k][lane];
      if ((ge[k] >> lane) & 1u) {
        const long long w = qb + __popc(ge[k] & lt_mask);
        if (w < P.cap_q) J.qc[w] = SC.f.qpos[k][lane];
      }
    }
    __syncwarp();
    // ---- windows that need the generic path ----
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      if (valid[k] && !fast[k]) window_generic(P, J, SC.s, i0 + k, bnd[2 * k], bnd[2 * k + 1], lane);
    }
  }
}

}  // namespace sai

using namespace sai;

static int window_stats_impl(const int32_t* d_pos, int64_t n_sites, const int64_t* d_win_start,
                             const int64_t* d_win_end, const int32_t* d_seg_lo, const int32_t* d_seg_hi,
                             int64_t n_windows, const sai_job* jobs,
                             int32_t n_jobs, const uint32_t* d_mask_u, const uint32_t* d_mask_q,
                             const double* d_qval, int64_t qval_stride, int32_t* d_nsnps,
                             int64_t* d_u, double* d_q, int32_t* d_q_cnt, int64_t* d_u_start,
                             int64_t* d_q_start, int64_t* d_totals, int32_t* d_u_cand,
                             int64_t cap_u, int32_t* d_q_cand, int64_t cap_q, void* stream) {
  SAI_REQUIRE(d_pos && d_mask_u && d_mask_q && d_qval && d_totals, "NULL device pointer");
  SAI_REQUIRE(n_sites >= 0 && n_sites < (1ll << 31) - 64, "n_sites out of range");
  SAI_REQUIRE(n_windows >= 0 && n_windows < (1ll << 31) - 1024, "window count out of range");
  SAI_REQUIRE(jobs && n_jobs >= 1 && n_jobs <= SAI_MAX_JOBS, "n_jobs %d outside [1,%d]", n_jobs,
              SAI_MAX_JOBS);
  SAI_REQUIRE(qval_stride >= n_sites, "qval_stride too small");
  SAI_REQUIRE(n_windows == 0 || (d_win_start && d_win_end && d_nsnps && d_u && d_q && d_q_cnt &&
                                 d_u_start && d_q_start),
              "NULL device pointer");
  SAI_REQUIRE(cap_u >= 0 && cap_q >= 0 && (cap_u == 0 || d_u_cand) && (cap_q == 0 || d_q_cand),
              "bad candidate buffers");
  WinParams P{};
  P.pos = d_pos;
  P.n_sites = (int32_t)n_sites;
  P.n_tiles = (int32_t)sai_num_tiles(n_sites);
  P.ws = d_win_start;
  P.we = d_win_end;
  P.seg_lo = d_seg_lo;
  P.seg_hi = d_seg_hi;
  P.W = (int32_t)n_windows;
  P.n_jobs = n_jobs;
  for (int j = 0; j < n_jobs; ++j) {
    P.u_enabled[j] = jobs[j].u.enabled;
    P.q_enabled[j] = jobs[j].q.enabled;
    P.quantile[j] = jobs[j].quantile;
    if (jobs[j].q.enabled)
      SAI_REQUIRE(jobs[j].quantile >= 0.0 && jobs[j].quantile <= 1.0,
                  "Quantiles must be in the range [0, 1]");
  }
  P.mask_u = d_mask_u;
  P.mask_q = d_mask_q;
  P.qval = d_qval;
  P.qval_stride = qval_stride;
  P.nsnps = d_nsnps;
  P.u = d_u;
  P.q = d_q;
  P.q_cnt = d_q_cnt;
  P.u_start = d_u_start;
  P.q_start = d_q_start;
  P.totals = reinterpret_cast<unsigned long long*>(d_totals);
  P.u_cand = d_u_cand;
  P.cap_u = cap_u;
  P.q_cand = d_q_cand;
  P.cap_q = cap_q;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SAI_CUDA_CHECK(cudaMemsetAsync(d_totals, 0, sizeof(int64_t) * 2 * n_jobs, st));
  if (n_windows == 0) return SAI_OK;
  const int64_t cap = (int64_t)sm_count() * 32;
#ifdef SAI_EXPERIMENTS
  // tuning knobs of tools/ builds: windows per warp and the occupancy bound
  static const int batch = [] {
    const char* e = getenv("SAI_WIN_BATCH");
    const int b = e ? atoi(e) : 1;
    return (b == 2 || b == 4) ? b : 1;
  }();
  static const int minb = [] {
    const char* e = getenv("SAI_WIN_MINB");
    return e ? atoi(e) : 12;
  }();
  const int64_t per_block = kWinWarps * batch;
  const int64_t want = (n_windows + per_block - 1) / per_block;
  const dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)n_jobs);
  if (batch == 1 && minb == 10)
    k_window_stats<1, 10><<<grid, kWinWarps * 32, 0, st>>>(P);
  else if (batch == 1 && minb == 12)
    k_window_stats<1, 12><<<grid, kWinWarps * 32, 0, st>>>(P);
  else if (batch == 1)
    k_window_stats<1, 8><<<grid, kWinWarps * 32, 0, st>>>(P);
  else if (batch == 2)
    k_window_stats<2, 6><<<grid, kWinWarps * 32, 0, st>>>(P);
  else
    k_window_stats<4, 4><<<grid, kWinWarps * 32, 0, st>>>(P);
#else
  // one window per warp, 12 blocks per SM: the configuration that measured fastest
  // (profiles/round1_notes.md); the batched variants live in -DSAI_EXPERIMENTS builds
  const int64_t want = (n_windows + kWinWarps - 1) / kWinWarps;
  const dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)n_jobs);
  k_window_stats<1, 12><<<grid, kWinWarps * 32, 0, st>>>(P);
#endif
  SAI_CUDA_CHECK(cudaGetLastError());
  return SAI_OK;
}

extern "C" int sai_window_stats(const int32_t* d_pos, int64_t n_sites, const int64_t* d_win_start,
                                const int64_t* d_win_end, int64_t n_windows, const sai_job* jobs,
                                int32_t n_jobs, const uint32_t* d_mask_u, const uint32_t* d_mask_q,
                                const double* d_qval, int64_t qval_stride, int32_t* d_nsnps,
                                int64_t* d_u, double* d_q, int32_t* d_q_cnt, int64_t* d_u_start,
                                int64_t* d_q_start, int64_t* d_totals, int32_t* d_u_cand,
                                int64_t cap_u, int32_t* d_q_cand, int64_t cap_q, void* stream) {
  return window_stats_impl(d_pos, n_sites, d_win_start, d_win_end, nullptr, nullptr, n_windows, jobs, n_jobs,
                           d_mask_u, d_mask_q, d_qval, qval_stride, d_nsnps, d_u, d_q, d_q_cnt, d_u_start,
                           d_q_start, d_totals, d_u_cand, cap_u, d_q_cand, cap_q, stream);
}

extern "C" int sai_window_stats_pieces(const int32_t* d_pos, int64_t n_sites, const int64_t* d_win_start,
                                       const int64_t* d_win_end, const int32_t* d_win_first_site,
                                       const int32_t* d_win_last_site, int64_t n_windows, const sai_job* jobs,
                                       int32_t n_jobs, const uint32_t* d_mask_u, const uint32_t* d_mask_q,
                                       const double* d_qval, int64_t qval_stride, int32_t* d_nsnps,
                                       int64_t* d_u, double* d_q, int32_t* d_q_cnt, int64_t* d_u_start,
                                       int64_t* d_q_start, int64_t* d_totals, int32_t* d_u_cand,
                                       int64_t cap_u, int32_t* d_q_cand, int64_t cap_q, void* stream) {
  SAI_REQUIRE(n_windows == 0 || (d_win_first_site && d_win_last_site), "NULL piece bounds");
  return window_stats_impl(d_pos, n_sites, d_win_start, d_win_end, d_win_first_site, d_win_last_site, n_windows,
                           jobs, n_jobs, d_mask_u, d_mask_q, d_qval, qval_stride, d_nsnps, d_u, d_q, d_q_cnt,
                           d_u_start, d_q_start, d_totals, d_u_cand, cap_u, d_q_cand, cap_q, stream);
}
