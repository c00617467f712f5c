// N3: per-window ABBA/BABA-style site-pattern sums for Danc, Dplus, df and fd.
//
// Follows calc_four_pops_freq / calc_pattern_sum (sai/stats/stat_utils.py:171-272):
// per site the frequencies of (ref, tgt, src, outgroup) -- outgroup 0 when absent
// (:212-213) -- and for a pattern such as "abba" the product
// ((((1) * (1-ref)) * tgt) * src) * (1-out), multiplied in that order (:261-270),
// then summed over the sites of the window.  A site nobody is called at has
// frequency NaN (0/0) and makes the window's sums NaN, as np.sum does.
// The per-site products are bit-exact, and so are the window sums: they are accumulated in
// the order of numpy's pairwise summation (`np.sum` of a contiguous float64 vector,
// numpy/core/src/umath/loops_utils.h.src `@TYPE@_pairwise_sum`, unchanged since 1.9):
//   n < 8          sequential;
//   n <= 128       eight accumulators r[j] += a[8 i + j], combined as
//                  ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the n % 8 tail sequentially;
//   n > 128        pairwise(a, n2) + pairwise(a + n2, n - n2) with n2 = n/2 rounded down to a
//                  multiple of 8.
// The model was checked against numpy itself for n = 1 .. 100 000 (tests/test_oracle_golden.py).
//
// Sums per (source population, window), in this order:
//   0 abba  1 baba  2 baaa  3 abaa  4 bbaa  5 abba_d  6 baba_d
// where the _d sums use dnr = max(tgt, src) for both middle populations
// (sai/stats/fd_statistic.py:78-81).
#include <algorithm>

#include "common.cuh"
#include "scratch.cuh"

namespace sai {

constexpr int kPatWarps = 4;
constexpr int kPatSums = 7;

struct PatParams {
  const int32_t* pos;
  int32_t n_sites;
  const int64_t* ws;
  const int64_t* we;
  int32_t W;
  const int32_t* num;
  const int32_t* called;
  int64_t count_stride;
  int32_t ref_pop, tgt_pop, out_pop;  // out_pop < 0: no outgroup
  int32_t n_src;
  int32_t src_pop[SAI_MAX_SRC];
  int32_t ploidy[SAI_MAX_POPS];
  double* sums;  // [n_src][W][7]
};

__device__ __forceinline__ int clampk(int64_t k) {
  return k > 2147483647ll ? 2147483647 : (k < -2147483647ll ? -2147483647 : (int)k);
}

// same cooperative 32-ary search as the window kernel, one key
__device__ __forceinline__ int warp_lower_bound1(const int32_t* __restrict__ pos, int n, int64_t key,
                                                 int lane) {
  const int k = clampk(key);
  int lo = key > 2147483647ll ? n : 0, hi = n;
  while (hi > lo) {
    const int nn = hi - lo;
    const int s = nn > 32 ? (nn + 31) >> 5 : 1;
    const unsigned idx = (unsigned)lo + (unsigned)(lane + 1) * (unsigned)s - 1u;
    const bool p = idx < (unsigned)hi && __ldg(pos + idx) < k;
    const int c = __popc(__ballot_sync(0xffffffffu, p));
    const unsigned nhi = (unsigned)lo + (unsigned)(c + 1) * (unsigned)s - 1u;
    lo += c * s;
    if (nhi < (unsigned)hi) hi = (int)nhi;
    if (lo > hi) lo = hi;
  }
  return lo;
}

__device__ __forceinline__ double freq_of(const PatParams& P, int pop, int site) {
  const int n = __ldg(P.num + (size_t)pop * P.count_stride + site);
  const int d = __ldg(P.called + (size_t)pop * P.count_stride + site) * P.ploidy[pop];
  return __ddiv_rn((double)n, (double)d);  // 0/0 -> NaN like calc_freq (stat_utils.py:51-52)
}

constexpr int kLeaf = 128;   // numpy's PW_BLOCKSIZE
constexpr int kMaxDepth = 32;  // recursion depth of pairwise_sum for n < 2^31 is <= 25

// ---- pass 1: the seven products of every site, once (a site lies in win_len / win_step windows:
// computing them per window would repeat the four float64 divisions that many times) ----------
// prod[(k * 7 + t) * stride + site], t in the order above; one thread per site, all sources.
struct ProdParams {
  PatParams P;
  double* prod;
  int64_t stride;
};

__global__ void __launch_bounds__(256) k_site_products(const __grid_constant__ ProdParams Q) {
  const PatParams& P = Q.P;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < P.n_sites; s += (int64_t)gridDim.x * blockDim.x) {
    const int site = (int)s;
    const double fr = freq_of(P, P.ref_pop, site);
    const double ft = freq_of(P, P.tgt_pop, site);
    const double fo = P.out_pop >= 0 ? freq_of(P, P.out_pop, site) : 0.0;
    const double ar = __dsub_rn(1.0, fr), at = __dsub_rn(1.0, ft), ao = __dsub_rn(1.0, fo);
    for (int k = 0; k < P.n_src; ++k) {
      const double fs = freq_of(P, P.src_pop[k], site);
      const double as = __dsub_rn(1.0, fs);
      // np.maximum propagates NaN
      const double dn = (ft != ft || fs != fs) ? (ft + fs) : (ft > fs ? ft : fs);
      const double adn = __dsub_rn(1.0, dn);
      double* o = Q.prod + (size_t)k * kPatSums * Q.stride + s;
      // product = 1; product *= f(ref); *= f(tgt); *= f(src); *= f(out)   (stat_utils.py:261-270)
      o[0 * Q.stride] = __dmul_rn(__dmul_rn(__dmul_rn(ar, ft), fs), ao);   // abba
      o[1 * Q.stride] = __dmul_rn(__dmul_rn(__dmul_rn(fr, at), fs), ao);   // baba
      o[2 * Q.stride] = __dmul_rn(__dmul_rn(__dmul_rn(fr, at), as), ao);   // baaa
      o[3 * Q.stride] = __dmul_rn(__dmul_rn(__dmul_rn(ar, ft), as), ao);   // abaa
      o[4 * Q.stride] = __dmul_rn(__dmul_rn(__dmul_rn(fr, ft), as), ao);   // bbaa
      o[5 * Q.stride] = __dmul_rn(__dmul_rn(__dmul_rn(ar, dn), dn), ao);   // abba_d
      o[6 * Q.stride] = __dmul_rn(__dmul_rn(__dmul_rn(fr, adn), dn), ao);  // baba_d
    }
  }
}

// ---- pass 2: window sums in numpy's pairwise order ---------------------------------------------
// One warp per (source population, window).  Lane (g, j) = (lane >> 3, lane & 7) runs accumulator
// j of the sums g and g + 4 (sum 7 does not exist), straight from the product rows in L2; the
// recursion above 128 elements is walked depth first with a per-lane stack (every lane follows
// the same path, so no shared memory and no warp barrier is needed).
__device__ __forceinline__ void leaf_sums(const double* __restrict__ a0, const double* __restrict__ a1, bool has1,
                                          int L, int j, double& s0, double& s1) {
  if (L < 8) {
    double r0 = 0.0, r1 = 0.0;
    for (int i = 0; i < L; ++i) {
      r0 = __dadd_rn(r0, __ldg(a0 + i));
      if (has1) r1 = __dadd_rn(r1, __ldg(a1 + i));
    }
    s0 = r0, s1 = r1;
    return;
  }
  const int body = L - (L & 7);
  double r0 = __ldg(a0 + j), r1 = has1 ? __ldg(a1 + j) : 0.0;
  int i = 8 + j;
  for (; i + 24 < body; i += 32) {  // four independent loads per row in flight, added in order
    double v0[4], v1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v0[u] = __ldg(a0 + i + 8 * u);
      v1[u] = has1 ? __ldg(a1 + i + 8 * u) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      r0 = __dadd_rn(r0, v0[u]);
      r1 = __dadd_rn(r1, v1[u]);
    }
  }
  for (; i < body; i += 8) {
    r0 = __dadd_rn(r0, __ldg(a0 + i));
    if (has1) r1 = __dadd_rn(r1, __ldg(a1 + i));
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {  // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) inside every group of 8 lanes
    r0 = __dadd_rn(r0, __shfl_xor_sync(0xffffffffu, r0, o));
    r1 = __dadd_rn(r1, __shfl_xor_sync(0xffffffffu, r1, o));
  }
  for (int t = body; t < L; ++t) {
    r0 = __dadd_rn(r0, __ldg(a0 + t));
    if (has1) r1 = __dadd_rn(r1, __ldg(a1 + t));
  }
  s0 = r0, s1 = r1;
}

// grid: x = windows (one warp each, grid-stride), y = source population
__global__ void __launch_bounds__(kPatWarps * 32) k_window_patterns(const __grid_constant__ ProdParams Q) {
  const PatParams& P = Q.P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.y;
  const int g = lane >> 3, j = lane & 7;
  const bool has1 = g + 4 < kPatSums;
  const double* row0 = Q.prod + ((size_t)k * kPatSums + g) * Q.stride;
  const double* row1 = Q.prod + ((size_t)k * kPatSums + (has1 ? g + 4 : g)) * Q.stride;
  for (int i = blockIdx.x * kPatWarps + warp; i < P.W; i += gridDim.x * kPatWarps) {
    const int lo = warp_lower_bound1(P.pos, P.n_sites, P.ws[i], lane);
    int hi = warp_lower_bound1(P.pos, P.n_sites, P.we[i] + 1, lane);
    if (hi < lo) hi = lo;
    // numpy's recursion, depth first; the stack is lane-uniform
    int st_start[kMaxDepth], st_len[kMaxDepth], st_state[kMaxDepth];
    double st_left0[kMaxDepth], st_left1[kMaxDepth];
    double ret0 = 0.0, ret1 = 0.0;  // this lane's sums g and g + 4 of the node just finished
    int sp_ = 1;
    st_start[0] = lo, st_len[0] = hi - lo, st_state[0] = 0;
    while (sp_ > 0) {
      const int top = sp_ - 1;
      const int start = st_start[top], len = st_len[top];
      if (len <= kLeaf) {
        leaf_sums(row0 + start, row1 + start, has1, len, j, ret0, ret1);
        --sp_;
        while (sp_ > 0) {  // hand the result to the ancestors
          const int par = sp_ - 1;
          if (st_state[par] == 1) {  // the left half is done: keep it, descend into the right half
            const int n2 = (st_len[par] / 2) - ((st_len[par] / 2) % 8);
            st_left0[par] = ret0, st_left1[par] = ret1;
            st_state[par] = 2;
            st_start[sp_] = st_start[par] + n2, st_len[sp_] = st_len[par] - n2, st_state[sp_] = 0;
            ++sp_;
            break;
          }
          // both halves are done: pairwise(left) + pairwise(right)
          ret0 = __dadd_rn(st_left0[par], ret0);
          ret1 = __dadd_rn(st_left1[par], ret1);
          --sp_;
        }
      } else {  // state 0: descend into the left half
        const int n2 = (len / 2) - ((len / 2) % 8);
        st_state[top] = 1;
        st_start[sp_] = start, st_len[sp_] = n2, st_state[sp_] = 0;
        ++sp_;
      }
    }
    if (j == 0) {
      double* out = P.sums + ((size_t)k * P.W + i) * kPatSums;
      out[g] = hi > lo ? ret0 : 0.0;
      if (has1) out[g + 4] = hi > lo ? ret1 : 0.0;
    }
  }
}

}  // namespace sai

using namespace sai;

extern "C" int sai_window_patterns(const sai_layout* lay, const int32_t* d_pos, int64_t n_sites,
                                   const int64_t* d_win_start, const int64_t* d_win_end,
                                   int64_t n_windows, const int32_t* d_num, const int32_t* d_called,
                                   int64_t count_stride, int32_t ref_pop, int32_t tgt_pop,
                                   int32_t out_pop, const int32_t* src_pops, int32_t n_src,
                                   double* d_sums, void* stream) {
  if (int rc = validate_layout(lay)) return rc;
  SAI_REQUIRE(d_pos && d_num && d_called && d_sums && src_pops, "NULL pointer");
  SAI_REQUIRE(n_sites >= 0 && n_sites < (1ll << 31) - 64 && count_stride >= n_sites, "bad n_sites / stride");
  SAI_REQUIRE(n_windows >= 0 && n_windows < (1ll << 31) - 1024, "window count out of range");
  SAI_REQUIRE(n_windows == 0 || (d_win_start && d_win_end), "NULL windows");
  SAI_REQUIRE(ref_pop >= 0 && ref_pop < lay->n_pops && tgt_pop >= 0 && tgt_pop < lay->n_pops &&
                  out_pop < lay->n_pops,
              "bad population index");
  SAI_REQUIRE(n_src >= 1 && n_src <= SAI_MAX_SRC, "n_src %d outside [1,%d]", n_src, SAI_MAX_SRC);
  PatParams P{};
  P.pos = d_pos;
  P.n_sites = (int32_t)n_sites;
  P.ws = d_win_start;
  P.we = d_win_end;
  P.W = (int32_t)n_windows;
  P.num = d_num;
  P.called = d_called;
  P.count_stride = count_stride;
  P.ref_pop = ref_pop;
  P.tgt_pop = tgt_pop;
  P.out_pop = out_pop;
  P.n_src = n_src;
  for (int k = 0; k < n_src; ++k) {
    SAI_REQUIRE(src_pops[k] >= 0 && src_pops[k] < lay->n_pops, "bad source population index");
    P.src_pop[k] = src_pops[k];
  }
  for (int p = 0; p < lay->n_pops; ++p) P.ploidy[p] = lay->pop[p].ploidy;
  P.sums = d_sums;
  if (n_windows == 0) return SAI_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // scratch for the per-site products: stream-ordered allocation, freed after the window pass
  const int64_t stride = (n_sites + 31) / 32 * 32;
  ProdParams Q{};
  Q.P = P;
  Q.stride = stride;
  void* scratch = nullptr;
  if (int rc = scratch_alloc(&scratch, sizeof(double) * kPatSums * (size_t)n_src * (size_t)std::max<int64_t>(stride, 32), st)) return rc;
  Q.prod = static_cast<double*>(scratch);
  if (n_sites > 0) {
    const int64_t blocks = (n_sites + 255) / 256;
    const int64_t capb = (int64_t)sm_count() * 16;
    k_site_products<<<(unsigned)(blocks < capb ? blocks : capb), 256, 0, st>>>(Q);
    SAI_CUDA_CHECK(cudaGetLastError());
  }
  const int64_t want = (n_windows + kPatWarps - 1) / kPatWarps;
  const int64_t cap = (int64_t)sm_count() * 16;
  const dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)n_src);
  k_window_patterns<<<grid, kPatWarps * 32, 0, st>>>(Q);
  SAI_CUDA_CHECK(cudaGetLastError());
  return scratch_free(scratch, st);
}
