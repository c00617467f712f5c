// K4..K6: per-window U, N(Variants), exact Q and the candidate position lists.
//
// One warp owns one (job, window).  A window [start, end] (inclusive,
// sai/generators/window_generator.py:173-174) maps to the site range [lo, hi)
// by binary search on the sorted positions; its U count is the popcount of the
// tile masks over that range (sai/stats/u_statistic.py:94-96) and Q is the
// numpy 'linear' quantile of the flagged sites' target frequencies
// (sai/stats/q_statistic.py:96-101; numpy/lib/_function_base_impl.py _lerp).
// Order statistics are selected exactly on the float64 bit patterns (all
// values are non-negative, so unsigned order == numeric order).
#include <math_constants.h>

#include "common.cuh"

namespace sai {

constexpr int kWinWarps = 4;
constexpr int kBufCap = 1024;  // doubles buffered in shared memory per warp

struct WinParams {
  const int32_t* pos;
  int64_t n_sites;
  int64_t n_tiles;
  const int64_t* ws;
  const int64_t* we;
  int64_t W;
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
  int64_t* u_off;  // [J][W+1]; K4 stores counts, K5 scans in place
  int64_t* q_off;
  int32_t* u_cand;
  int64_t cap_u;
  int32_t* q_cand;
  int64_t cap_q;
};

__device__ __forceinline__ int64_t lower_bound_pos(const int32_t* pos, int64_t n, int64_t key) {
  int64_t lo = 0, hi = n;  // first index with pos >= key
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)__ldg(pos + mid) < key)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
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
__device__ __forceinline__ uint32_t range_mask(int64_t T, int64_t lo, int64_t hi) {
  const int64_t base = T * kTile;
  uint32_t m = 0xffffffffu;
  if (lo > base) m &= 0xffffffffu << (int)(lo - base);
  if (hi < base + kTile) m &= 0xffffffffu >> (int)(base + kTile - hi);
  return m;
}

// Element source walking the Q mask of the window (values gathered from the
// dense qval array).  for_each calls f(key) once per flagged site, on some lane.
struct MaskSource {
  const uint32_t* mask;  // job's mask_q
  const double* qval;    // job's qval
  int64_t lo, hi, T0, T1;
  int lane;
  template <typename F>
  __device__ __forceinline__ void for_each(F f) const {
    for (int64_t T = T0 + lane; T <= T1; T += 32) {
      uint32_t m = __ldg(mask + T) & range_mask(T, lo, hi);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        f((unsigned long long)__double_as_longlong(__ldg(qval + T * kTile + b)));
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

__global__ void __launch_bounds__(kWinWarps * 32) k_window_stats(const __grid_constant__ WinParams P) {
  __shared__ unsigned long long s_buf[kWinWarps][kBufCap];
  __shared__ int s_hist[kWinWarps][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total = (int64_t)P.n_jobs * P.W;
  for (int64_t item = (int64_t)blockIdx.x * kWinWarps + warp; item < total;
       item += (int64_t)gridDim.x * kWinWarps) {
    const int j = (int)(item / P.W);
    const int64_t i = item - (int64_t)j * P.W;
    const int64_t lo = lower_bound_pos(P.pos, P.n_sites, P.ws[i]);
    const int64_t hi = lower_bound_pos(P.pos, P.n_sites, P.we[i] + 1);
    const uint32_t* mu = P.mask_u + (size_t)j * P.n_tiles;
    const uint32_t* mq = P.mask_q + (size_t)j * P.n_tiles;
    const double* qv = P.qval + (size_t)j * P.qval_stride;
    int u_cnt = 0, q_n = 0;
    int64_t T0 = 0, T1 = -1;
    if (hi > lo) {
      T0 = lo / kTile;
      T1 = (hi - 1) / kTile;
      unsigned long long* buf = s_buf[warp];
      for (int64_t Tb = T0; Tb <= T1; Tb += 32) {
        const int64_t T = Tb + lane;
        uint32_t a = 0, b = 0;
        if (T <= T1) {
          const uint32_t rm = range_mask(T, lo, hi);
          a = __ldg(mu + T) & rm;
          b = __ldg(mq + T) & rm;
        }
        u_cnt += __popc(a);
        const int c = __popc(b);
        const int off = q_n + warp_excl_scan(c, lane);
        q_n += warp_sum(c);
        // gather this lane's flagged values into the warp buffer (site order)
        int w = off;
        while (b) {
          const int bit = __ffs(b) - 1;
          b &= b - 1;
          if (w < kBufCap)
            buf[w] = (unsigned long long)__double_as_longlong(__ldg(qv + T * kTile + bit));
          ++w;
        }
      }
      u_cnt = warp_sum(u_cnt);
      __syncwarp();
    }
    double qres = CUDART_NAN;
    int q_cand = 0;
    if (P.q_enabled[j] && q_n > 0) {
      const double qq = P.quantile[j];
      if (q_n <= 32) {
        const unsigned long long key = lane < q_n ? s_buf[warp][lane] : ~0ull;
        qres = warp_quantile_small(key, q_n, qq, lane);
        const bool ge = lane < q_n && __longlong_as_double((long long)key) >= qres;
        q_cand = __popc(__ballot_sync(0xffffffffu, ge));
      } else if (q_n <= kBufCap) {
        BufSource src{s_buf[warp], q_n, lane};
        qres = warp_quantile(src, q_n, qq, s_hist[warp], lane);
        q_cand = warp_count_ge(src, qres);
      } else {
        MaskSource src{mq, qv, lo, hi, T0, T1, lane};
        qres = warp_quantile(src, q_n, qq, s_hist[warp], lane);
        q_cand = warp_count_ge(src, qres);
      }
    }
    __syncwarp();
    if (lane == 0) {
      P.nsnps[item] = (int32_t)(hi - lo);
      P.u[item] = P.u_enabled[j] ? u_cnt : 0;
      P.q[item] = qres;
      P.u_off[(size_t)j * (P.W + 1) + i] = P.u_enabled[j] ? u_cnt : 0;
      P.q_off[(size_t)j * (P.W + 1) + i] = q_cand;
      if (i == P.W - 1) {
        P.u_off[(size_t)j * (P.W + 1) + P.W] = 0;
        P.q_off[(size_t)j * (P.W + 1) + P.W] = 0;
      }
    }
  }
}

// K5: in-place exclusive scan of the per-window candidate counts -> CSR offsets.
// One block per (job, statistic); chunks of blockDim with a running carry.
__global__ void __launch_bounds__(1024) k_scan_offsets(int64_t* u_off, int64_t* q_off, int64_t W) {
  __shared__ long long s_warp[32];
  __shared__ long long s_total;
  __shared__ long long s_carry;
  int64_t* a = (blockIdx.y == 0 ? u_off : q_off) + (size_t)blockIdx.x * (W + 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base <= W; base += blockDim.x) {
    const int64_t idx = base + threadIdx.x;
    const long long v = idx <= W ? a[idx] : 0;
    long long x = v;  // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      const long long w = s_warp[lane];
      long long xx = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, xx, o);
        if (lane >= o) xx += y;
      }
      s_warp[lane] = xx - w;  // exclusive offset of each warp
      if (lane == 31) s_total = xx;
    }
    __syncthreads();
    if (idx <= W) a[idx] = s_carry + s_warp[warp] + (x - v);
    __syncthreads();
    if (threadIdx.x == 0) s_carry += s_total;
    __syncthreads();
  }
}

// K6: candidate positions (u_statistic.py:95, q_statistic.py:101) in site order.
__global__ void __launch_bounds__(kWinWarps * 32) k_fill_candidates(const __grid_constant__ WinParams P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total = (int64_t)P.n_jobs * P.W;
  for (int64_t item = (int64_t)blockIdx.x * kWinWarps + warp; item < total;
       item += (int64_t)gridDim.x * kWinWarps) {
    const int j = (int)(item / P.W);
    const int64_t i = item - (int64_t)j * P.W;
    const int64_t u0 = P.u_off[(size_t)j * (P.W + 1) + i], u1 = P.u_off[(size_t)j * (P.W + 1) + i + 1];
    const int64_t q0 = P.q_off[(size_t)j * (P.W + 1) + i], q1 = P.q_off[(size_t)j * (P.W + 1) + i + 1];
    if (u1 == u0 && q1 == q0) continue;
    const int64_t lo = lower_bound_pos(P.pos, P.n_sites, P.ws[i]);
    const int64_t hi = lower_bound_pos(P.pos, P.n_sites, P.we[i] + 1);
    if (hi <= lo) continue;
    const uint32_t* mu = P.mask_u + (size_t)j * P.n_tiles;
    const uint32_t* mq = P.mask_q + (size_t)j * P.n_tiles;
    const double* qv = P.qval + (size_t)j * P.qval_stride;
    const double thr = P.q[item];
    int32_t* uc = P.u_cand + (size_t)j * P.cap_u;
    int32_t* qc = P.q_cand + (size_t)j * P.cap_q;
    const int64_t T0 = lo / kTile, T1 = (hi - 1) / kTile;
    int64_t uw = u0, qw = q0;
    for (int64_t Tb = T0; Tb <= T1; Tb += 32) {
      const int64_t T = Tb + lane;
      uint32_t a = 0, b = 0;
      if (T <= T1) {
        const uint32_t rm = range_mask(T, lo, hi);
        if (u1 > u0) a = __ldg(mu + T) & rm;
        if (q1 > q0) {
          uint32_t m = __ldg(mq + T) & rm;
          while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            if (__ldg(qv + T * kTile + bit) >= thr) b |= 1u << bit;
          }
        }
      }
      const int ca = __popc(a), cb = __popc(b);
      int64_t wa = uw + warp_excl_scan(ca, lane);
      int64_t wb = qw + warp_excl_scan(cb, lane);
      uw += warp_sum(ca);
      qw += warp_sum(cb);
      while (a) {
        const int bit = __ffs(a) - 1;
        a &= a - 1;
        if (wa < P.cap_u) uc[wa] = __ldg(P.pos + T * kTile + bit);
        ++wa;
      }
      while (b) {
        const int bit = __ffs(b) - 1;
        b &= b - 1;
        if (wb < P.cap_q) qc[wb] = __ldg(P.pos + T * kTile + bit);
        ++wb;
      }
    }
  }
}

static int fill_params(WinParams& P, const int32_t* d_pos, int64_t n_sites, const int64_t* ws,
                       const int64_t* we, int64_t W, int32_t n_jobs, const uint32_t* mask_u,
                       const uint32_t* mask_q, const double* qval, int64_t qval_stride) {
  SAI_REQUIRE(d_pos && ws && we && mask_u && mask_q && qval, "NULL device pointer");
  SAI_REQUIRE(n_sites >= 0 && n_sites < (1ll << 31), "n_sites out of range");
  SAI_REQUIRE(W >= 0, "negative window count");
  SAI_REQUIRE(n_jobs >= 1 && n_jobs <= SAI_MAX_JOBS, "n_jobs %d outside [1,%d]", n_jobs, SAI_MAX_JOBS);
  SAI_REQUIRE(qval_stride >= n_sites, "qval_stride too small");
  P.pos = d_pos;
  P.n_sites = n_sites;
  P.n_tiles = sai_num_tiles(n_sites);
  P.ws = ws;
  P.we = we;
  P.W = W;
  P.n_jobs = n_jobs;
  P.mask_u = mask_u;
  P.mask_q = mask_q;
  P.qval = qval;
  P.qval_stride = qval_stride;
  return SAI_OK;
}

static int win_grid(int64_t items) {
  const int64_t want = (items + kWinWarps - 1) / kWinWarps;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace sai

using namespace sai;

extern "C" {

int sai_window_stats(const int32_t* d_pos, int64_t n_sites, const int64_t* d_win_start,
                     const int64_t* d_win_end, int64_t n_windows, const sai_job* jobs,
                     int32_t n_jobs, const uint32_t* d_mask_u, const uint32_t* d_mask_q,
                     const double* d_qval, int64_t qval_stride, int32_t* d_nsnps, int64_t* d_u,
                     double* d_q, int64_t* d_u_off, int64_t* d_q_off, int32_t* d_u_cand,
                     int64_t cap_u, int32_t* d_q_cand, int64_t cap_q, void* stream) {
  WinParams P{};
  if (int rc = fill_params(P, d_pos, n_sites, d_win_start, d_win_end, n_windows, n_jobs, d_mask_u,
                           d_mask_q, d_qval, qval_stride))
    return rc;
  SAI_REQUIRE(jobs && d_nsnps && d_u && d_q && d_u_off && d_q_off, "NULL device pointer");
  SAI_REQUIRE(cap_u >= 0 && cap_q >= 0 && (cap_u == 0 || d_u_cand) && (cap_q == 0 || d_q_cand),
              "bad candidate buffers");
  for (int j = 0; j < n_jobs; ++j) {
    P.u_enabled[j] = jobs[j].u.enabled;
    P.q_enabled[j] = jobs[j].q.enabled;
    P.quantile[j] = jobs[j].quantile;
    if (jobs[j].q.enabled)
      SAI_REQUIRE(jobs[j].quantile >= 0.0 && jobs[j].quantile <= 1.0,
                  "Quantiles must be in the range [0, 1]");
  }
  P.nsnps = d_nsnps;
  P.u = d_u;
  P.q = d_q;
  P.u_off = d_u_off;
  P.q_off = d_q_off;
  P.u_cand = d_u_cand;
  P.cap_u = cap_u;
  P.q_cand = d_q_cand;
  P.cap_q = cap_q;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_windows == 0) {
    SAI_CUDA_CHECK(cudaMemsetAsync(d_u_off, 0, sizeof(int64_t) * n_jobs, st));
    SAI_CUDA_CHECK(cudaMemsetAsync(d_q_off, 0, sizeof(int64_t) * n_jobs, st));
    return SAI_OK;
  }
  const int grid = win_grid((int64_t)n_jobs * n_windows);
  k_window_stats<<<grid, kWinWarps * 32, 0, st>>>(P);
  SAI_CUDA_CHECK(cudaGetLastError());
  k_scan_offsets<<<dim3(n_jobs, 2), 1024, 0, st>>>(d_u_off, d_q_off, n_windows);
  SAI_CUDA_CHECK(cudaGetLastError());
  if (cap_u > 0 || cap_q > 0) {
    k_fill_candidates<<<grid, kWinWarps * 32, 0, st>>>(P);
    SAI_CUDA_CHECK(cudaGetLastError());
  }
  return SAI_OK;
}

int sai_fill_candidates(const int32_t* d_pos, int64_t n_sites, const int64_t* d_win_start,
                        const int64_t* d_win_end, int64_t n_windows, int32_t n_jobs,
                        const uint32_t* d_mask_u, const uint32_t* d_mask_q, const double* d_qval,
                        int64_t qval_stride, const double* d_q, const int64_t* d_u_off,
                        const int64_t* d_q_off, int32_t* d_u_cand, int64_t cap_u,
                        int32_t* d_q_cand, int64_t cap_q, void* stream) {
  WinParams P{};
  if (int rc = fill_params(P, d_pos, n_sites, d_win_start, d_win_end, n_windows, n_jobs, d_mask_u,
                           d_mask_q, d_qval, qval_stride))
    return rc;
  SAI_REQUIRE(d_q && d_u_off && d_q_off, "NULL device pointer");
  SAI_REQUIRE(cap_u >= 0 && cap_q >= 0 && (cap_u == 0 || d_u_cand) && (cap_q == 0 || d_q_cand),
              "bad candidate buffers");
  if (n_windows == 0) return SAI_OK;
  P.q = const_cast<double*>(d_q);
  P.u_off = const_cast<int64_t*>(d_u_off);
  P.q_off = const_cast<int64_t*>(d_q_off);
  P.u_cand = d_u_cand;
  P.cap_u = cap_u;
  P.q_cand = d_q_cand;
  P.cap_q = cap_q;
  const int grid = win_grid((int64_t)n_jobs * n_windows);
  k_fill_candidates<<<grid, kWinWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(P);
  SAI_CUDA_CHECK(cudaGetLastError());
  return SAI_OK;
}

}  // extern "C"
