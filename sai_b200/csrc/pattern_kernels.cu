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
#include "common.cuh"

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
constexpr int kMaxDepth = 40;

// per-warp scratch: the seven products of one leaf's sites, and the recursion stack
struct PatScratch {
  double a[kPatSums][kLeaf];
  double left[kMaxDepth][kPatSums];
  int start[kMaxDepth], len[kMaxDepth], state[kMaxDepth];
};

// numpy's pairwise_sum of the <= 128 values of every product held in S.a: lane (g, j) = (lane >> 3,
// lane & 7) runs accumulator j of the sums g and g + 4; returns this lane's two sums
__device__ __forceinline__ void leaf_sums(const PatScratch& S, int L, int lane, double& s0, double& s1) {
  const int g = lane >> 3, j = lane & 7;
  const int t0 = g, t1 = g + 4;  // t1 == 7 does not exist
  if (L < 8) {
    double r0 = 0.0, r1 = 0.0;
    for (int i = 0; i < L; ++i) {
      r0 = __dadd_rn(r0, S.a[t0][i]);
      if (t1 < kPatSums) r1 = __dadd_rn(r1, S.a[t1][i]);
    }
    s0 = r0, s1 = r1;
    return;
  }
  const int body = L - (L & 7);
  double r0 = S.a[t0][j], r1 = t1 < kPatSums ? S.a[t1][j] : 0.0;
  for (int i = 8 + j; i < body; i += 8) {
    r0 = __dadd_rn(r0, S.a[t0][i]);
    if (t1 < kPatSums) r1 = __dadd_rn(r1, S.a[t1][i]);
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {  // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) inside every group of 8 lanes
    r0 = __dadd_rn(r0, __shfl_xor_sync(0xffffffffu, r0, o));
    r1 = __dadd_rn(r1, __shfl_xor_sync(0xffffffffu, r1, o));
  }
  for (int i = body; i < L; ++i) {
    r0 = __dadd_rn(r0, S.a[t0][i]);
    if (t1 < kPatSums) r1 = __dadd_rn(r1, S.a[t1][i]);
  }
  s0 = r0, s1 = r1;
}

// grid: x = windows (one warp each, grid-stride), y = source population
__global__ void __launch_bounds__(kPatWarps * 32) k_window_patterns(const __grid_constant__ PatParams P) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PatScratch& S = reinterpret_cast<PatScratch*>(s_raw)[warp];
  const int k = blockIdx.y;
  const int sp = P.src_pop[k];
  const int g = lane >> 3;
  for (int i = blockIdx.x * kPatWarps + warp; i < P.W; i += gridDim.x * kPatWarps) {
    const int lo = warp_lower_bound1(P.pos, P.n_sites, P.ws[i], lane);
    int hi = warp_lower_bound1(P.pos, P.n_sites, P.we[i] + 1, lane);
    if (hi < lo) hi = lo;
    // the products of sites [start, start + L) into S.a
    auto fill = [&](int start, int L) {
      for (int e = lane; e < L; e += 32) {
        const int s = start + e;
        const double fr = freq_of(P, P.ref_pop, s);
        const double ft = freq_of(P, P.tgt_pop, s);
        const double fs = freq_of(P, sp, s);
        const double fo = P.out_pop >= 0 ? freq_of(P, P.out_pop, s) : 0.0;
        const double ar = __dsub_rn(1.0, fr), at = __dsub_rn(1.0, ft), as = __dsub_rn(1.0, fs),
                     ao = __dsub_rn(1.0, fo);
        // np.maximum propagates NaN
        const double dn = (ft != ft || fs != fs) ? (ft + fs) : (ft > fs ? ft : fs);
        const double adn = __dsub_rn(1.0, dn);
        // product = 1; product *= f(ref); *= f(tgt); *= f(src); *= f(out)
        S.a[0][e] = __dmul_rn(__dmul_rn(__dmul_rn(ar, ft), fs), ao);   // abba
        S.a[1][e] = __dmul_rn(__dmul_rn(__dmul_rn(fr, at), fs), ao);   // baba
        S.a[2][e] = __dmul_rn(__dmul_rn(__dmul_rn(fr, at), as), ao);   // baaa
        S.a[3][e] = __dmul_rn(__dmul_rn(__dmul_rn(ar, ft), as), ao);   // abaa
        S.a[4][e] = __dmul_rn(__dmul_rn(__dmul_rn(fr, ft), as), ao);   // bbaa
        S.a[5][e] = __dmul_rn(__dmul_rn(__dmul_rn(ar, dn), dn), ao);   // abba_d
        S.a[6][e] = __dmul_rn(__dmul_rn(__dmul_rn(fr, adn), dn), ao);  // baba_d
      }
      __syncwarp();
    };
    // numpy's recursion, depth first with an explicit stack; every lane follows the same path
    double ret0 = 0.0, ret1 = 0.0;  // this lane's sums g and g + 4 of the node just finished
    int sp_ = 0;
    if (lane == 0) S.start[0] = lo, S.len[0] = hi - lo, S.state[0] = 0;
    sp_ = 1;
    __syncwarp();
    while (sp_ > 0) {
      const int top = sp_ - 1;
      const int start = S.start[top], len = S.len[top], state = S.state[top];
      if (len <= kLeaf) {
        fill(start, len);
        leaf_sums(S, len, lane, ret0, ret1);
        __syncwarp();
        --sp_;
        // hand the result to the ancestors
        while (sp_ > 0) {
          const int par = sp_ - 1;
          if (S.state[par] == 1) {  // the left half is done: keep it, descend into the right half
            const int n2 = (S.len[par] / 2) - ((S.len[par] / 2) % 8);
            __syncwarp();
            if ((lane & 7) == 0) {
              S.left[par][g] = ret0;
              if (g + 4 < kPatSums) S.left[par][g + 4] = ret1;
            }
            if (lane == 0) {
              S.state[par] = 2;
              S.start[sp_] = S.start[par] + n2, S.len[sp_] = S.len[par] - n2, S.state[sp_] = 0;
            }
            ++sp_;
            __syncwarp();
            break;
          }
          // both halves are done: pairwise(left) + pairwise(right)
          ret0 = __dadd_rn(S.left[par][g], ret0);
          if (g + 4 < kPatSums) ret1 = __dadd_rn(S.left[par][g + 4], ret1);
          --sp_;
        }
      } else if (state == 0) {  // descend into the left half
        const int n2 = (len / 2) - ((len / 2) % 8);
        __syncwarp();
        if (lane == 0) {
          S.state[top] = 1;
          S.start[sp_] = start, S.len[sp_] = n2, S.state[sp_] = 0;
        }
        ++sp_;
        __syncwarp();
      }
    }
    if ((lane & 7) == 0) {
      double* out = P.sums + ((size_t)k * P.W + i) * kPatSums;
      out[g] = hi > lo ? ret0 : 0.0;
      if (g + 4 < kPatSums) out[g + 4] = hi > lo ? ret1 : 0.0;
    }
    __syncwarp();
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
  const int64_t want = (n_windows + kPatWarps - 1) / kPatWarps;
  const int64_t cap = (int64_t)sm_count() * 8;
  const dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)n_src);
  const size_t smem = sizeof(PatScratch) * kPatWarps;
  SAI_CUDA_CHECK(cudaFuncSetAttribute(k_window_patterns, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_window_patterns<<<grid, kPatWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(P);
  SAI_CUDA_CHECK(cudaGetLastError());
  return SAI_OK;
}
