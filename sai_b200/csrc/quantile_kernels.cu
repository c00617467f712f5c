// N1: genome-wide outlier threshold of a score column (`sai outlier`, sai/sai.py:192-214) on the
// device: the linear quantile of the column's non-NaN values (pandas `Series.quantile(q)` ->
// numpy 'linear': virtual index (n-1) q, the two neighbouring order statistics, separately rounded
// lerp -- numpy/lib/_function_base_impl.py _get_indexes / _lerp).
//
// The column arrives as `n_chunks` equally long pieces (one per rank after ONE
// all_gather_into_tensor over NVLink; NaN = padding or a window without a value), so nothing is
// copied or sorted on the host.  One 1024-thread block per column selects the order statistic
// exactly by an MSB-first radix select over the order-preserving 64-bit image of the doubles
// (sign-flipped, so negative statistics such as Danc / fd work too): 8 passes with a 256-bin
// shared histogram, stopping as soon as one key is left.  Everything is exact: the result equals
// what sorting the column gives.
#include <math_constants.h>

#include "common.cuh"

namespace sai {

constexpr int kQThreads = 1024;

struct ColParams {
  const double* vals;
  int32_t n_chunks;
  int64_t chunk_stride;  // doubles between the pieces of one column
  int64_t col_stride;    // doubles between columns inside a piece
  int64_t len;           // doubles per piece
  double q;
  double* out;  // [n_cols][4]: threshold (NaN if undefined), n_valid, min, max
};

__device__ __forceinline__ unsigned long long ordered_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(__dadd_rn(v, 0.0));  // -0.0 -> +0.0
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// f(key) for every non-NaN value of the column, spread over the block's threads
template <typename F>
__device__ __forceinline__ void for_each_key(const ColParams& P, int col, F f) {
  for (int c = 0; c < P.n_chunks; ++c) {
    const double* p = P.vals + (size_t)c * P.chunk_stride + (size_t)col * P.col_stride;
    int64_t i = threadIdx.x;
    for (; i + 3 * kQThreads < P.len; i += 4 * kQThreads) {  // four loads in flight per thread
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(p + i + u * kQThreads);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (v[u] == v[u]) f(ordered_key(v[u]));
    }
    for (; i < P.len; i += kQThreads) {
      const double v = __ldg(p + i);
      if (v == v) f(ordered_key(v));
    }
  }
}

__device__ __forceinline__ unsigned long long block_min64(unsigned long long v, unsigned long long* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, o);
    const unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), o);
    const unsigned long long w = ((unsigned long long)hi << 32) | lo;
    v = w < v ? w : v;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long r = s_red[0];
  for (int w = 1; w < kQThreads / 32; ++w) r = s_red[w] < r ? s_red[w] : r;
  return r;
}

// exact k-th smallest key (0-based); count_le = number of keys <= it
__device__ unsigned long long block_radix_select(const ColParams& P, int col, long long k, int* s_hist,
                                                 long long* s_bcast, unsigned long long* s_red, long long& count_le) {
  unsigned long long prefix = 0;
  long long kk = k, eq = 0;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    __syncthreads();
    if (threadIdx.x < 256) s_hist[threadIdx.x] = 0;
    __syncthreads();
    for_each_key(P, col, [&](unsigned long long key) {
      if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&s_hist[(int)((key >> shift) & 255ull)], 1);
    });
    __syncthreads();
    if (threadIdx.x == 0) {  // 256 bins: a serial scan is cheaper than a block scan here
      long long run = 0;
      int bin = 255;
      for (int b = 0; b < 256; ++b) {
        if (kk < run + s_hist[b]) {
          bin = b;
          break;
        }
        run += s_hist[b];
      }
      s_bcast[0] = bin;
      s_bcast[1] = run;
      s_bcast[2] = s_hist[bin];
    }
    __syncthreads();
    const int bin = (int)s_bcast[0];
    const long long below = s_bcast[1], cnt = s_bcast[2];
    prefix |= (unsigned long long)bin << shift;
    kk -= below;
    eq = cnt;
    if (cnt == 1 && pass < 7) {  // a single key carries this prefix: fetch it and stop
      unsigned long long found = ~0ull;
      for_each_key(P, col, [&](unsigned long long key) {
        if ((key >> shift) == (prefix >> shift)) found = key;
      });
      prefix = block_min64(found, s_red);
      break;
    }
  }
  count_le = (k - kk) + eq;
  return prefix;
}

__global__ void __launch_bounds__(kQThreads) k_column_quantile(const __grid_constant__ ColParams P) {
  __shared__ int s_hist[256];
  __shared__ long long s_bcast[4];
  __shared__ unsigned long long s_red[kQThreads / 32];
  __shared__ unsigned long long s_stat[3];
  const int col = blockIdx.x;
  if (threadIdx.x == 0) {
    s_stat[0] = 0;       // n_valid
    s_stat[1] = ~0ull;   // min key
    s_stat[2] = 0;       // max key
  }
  __syncthreads();
  unsigned long long n_loc = 0, mn = ~0ull, mx = 0;
  for_each_key(P, col, [&](unsigned long long key) {
    ++n_loc;
    mn = key < mn ? key : mn;
    mx = key > mx ? key : mx;
  });
  atomicAdd(&s_stat[0], n_loc);
  atomicMin(&s_stat[1], mn);
  atomicMax(&s_stat[2], mx);
  __syncthreads();
  const long long n = (long long)s_stat[0];
  const unsigned long long kmin = s_stat[1], kmax = s_stat[2];
  double thr = CUDART_NAN;
  if (n > 0 && kmin != kmax) {  // empty or single-valued column: no threshold (sai.py:195-207)
    const double vi = __dmul_rn((double)(n - 1), P.q);
    long long cle;
    if (vi >= (double)(n - 1)) {
      thr = key_value(kmax);
    } else {
      const double fl = floor(vi);
      const long long k = (long long)fl;
      const double g = __dsub_rn(vi, fl);
      const unsigned long long ka = block_radix_select(P, col, k, s_hist, s_bcast, s_red, cle);
      unsigned long long kb = ka;
      if (k + 1 >= cle) {  // the next order statistic is the smallest key above ka
        unsigned long long best = ~0ull;
        for_each_key(P, col, [&](unsigned long long key) {
          if (key > ka && key < best) best = key;
        });
        kb = block_min64(best, s_red);
      }
      const double a = key_value(ka), b = key_value(kb);
      const double d = __dsub_rn(b, a);
      thr = g >= 0.5 ? __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, g))) : __dadd_rn(a, __dmul_rn(d, g));
    }
  }
  if (threadIdx.x == 0) {
    double* o = P.out + 4 * (size_t)col;
    o[0] = thr;
    o[1] = (double)n;
    o[2] = n > 0 ? key_value(kmin) : CUDART_NAN;
    o[3] = n > 0 ? key_value(kmax) : CUDART_NAN;
  }
}

}  // namespace sai

using namespace sai;

extern "C" int sai_column_quantiles(const double* d_vals, int32_t n_cols, int32_t n_chunks, int64_t chunk_stride,
                                    int64_t col_stride, int64_t len, double q, double* d_out, void* stream) {
  SAI_REQUIRE(n_cols >= 0 && n_chunks >= 1 && len >= 0 && chunk_stride >= 0 && col_stride >= 0, "bad sizes");
  SAI_REQUIRE(q >= 0.0 && q <= 1.0, "Quantiles must be in the range [0, 1]");
  if (n_cols == 0) return SAI_OK;
  SAI_REQUIRE(d_out && (len == 0 || d_vals), "NULL device pointer");
  ColParams P{d_vals, n_chunks, chunk_stride, col_stride, len, q, d_out};
  k_column_quantile<<<n_cols, kQThreads, 0, static_cast<cudaStream_t>(stream)>>>(P);
  SAI_CUDA_CHECK(cudaGetLastError());
  return SAI_OK;
}
