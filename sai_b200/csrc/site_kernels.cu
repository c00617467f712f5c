// K1: the genotype pass.  One warp owns one 32-site tile, lane == site.
//
// Replaces, per population, the four numpy passes of calc_freq
// (sai/stats/stat_utils.py:45-49) by a carry-save (Harley-Seal) popcount over
// the tile's bit-planes, and -- fused in the epilogue -- the site conditions of
// compute_matching_loci (stat_utils.py:114-166) and UStatistic / QStatistic
// (u_statistic.py:92, q_statistic.py:92).
//
// HBM-bound: every packed byte is read exactly once with coalesced 256-byte
// warp requests (8 bytes per lane); outputs are two 32-bit masks per tile and
// job plus one double per Q-flagged site.
#include "common.cuh"
#include "site_cond.cuh"

namespace sai {

// streaming 8-byte load: read-only path, do not allocate in L1.  (An L2
// evict-first cache hint was measured: it made the window kernel 14 us faster
// but this kernel 80 us slower -- profiles/round1_notes.md.)
__device__ __forceinline__ uint2 ld_stream(const uint2* p) {
  uint2 r;
  asm("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// Bit-sliced counter of one word stream: ones/twos/fours hold weights 1/2/4,
// `high` counts completed weight-8 carries.
struct SliceCounter {
  uint32_t ones = 0, twos = 0, fours = 0;
  int high = 0;
  __device__ __forceinline__ void add8(const uint32_t (&w)[8]) {
    uint32_t tA = maj3(ones, w[0], w[1]);
    ones = xor3(ones, w[0], w[1]);
    uint32_t tB = maj3(ones, w[2], w[3]);
    ones = xor3(ones, w[2], w[3]);
    uint32_t fA = maj3(twos, tA, tB);
    twos = xor3(twos, tA, tB);
    tA = maj3(ones, w[4], w[5]);
    ones = xor3(ones, w[4], w[5]);
    tB = maj3(ones, w[6], w[7]);
    ones = xor3(ones, w[6], w[7]);
    uint32_t fB = maj3(twos, tA, tB);
    twos = xor3(twos, tA, tB);
    uint32_t e = maj3(fours, fA, fB);
    fours = xor3(fours, fA, fB);
    high += __popc(e);
  }
  __device__ __forceinline__ int total() const {
    return 8 * high + 4 * __popc(fours) + 2 * __popc(twos) + __popc(ones);
  }
};

// B == 2 population: planes (a = bit0, b = bit1) of one group sit in one pair.
//   num  = popc(a) + 2 popc(b) - 3 popc(a&b),   missing = popc(a&b)
template <int MODE>
__device__ __forceinline__ void count_b2(const uint2* __restrict__ col, int n_pairs, int& num,
                                         int& miss) {
  int p = 0;
  int acc_a = 0, acc_b = 0, acc_m = 0;
  if (MODE != 2) {
    SliceCounter ca, cb, cm;
    if (MODE == 1) {
      // 16 loads in flight per lane
      for (; p + 16 <= n_pairs; p += 16) {
        uint2 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = ld_stream(col + (size_t)(p + i) * kTile);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t a[8], b[8], m[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            a[i] = v[8 * h + i].x;
            b[i] = v[8 * h + i].y;
            m[i] = v[8 * h + i].x & v[8 * h + i].y;
          }
          ca.add8(a);
          cb.add8(b);
          cm.add8(m);
        }
      }
    }
    for (; p + 8 <= n_pairs; p += 8) {
      uint2 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ld_stream(col + (size_t)(p + i) * kTile);
      uint32_t a[8], b[8], m[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] = v[i].x;
        b[i] = v[i].y;
        m[i] = v[i].x & v[i].y;
      }
      ca.add8(a);
      cb.add8(b);
      cm.add8(m);
    }
    acc_a = ca.total();
    acc_b = cb.total();
    acc_m = cm.total();
  }
  // remainder (and MODE 2): direct popcounts
#pragma unroll 4
  for (; p < n_pairs; ++p) {
    uint2 v = ld_stream(col + (size_t)p * kTile);
    acc_a += __popc(v.x);
    acc_b += __popc(v.y);
    acc_m += __popc(v.x & v.y);
  }
  num = acc_a + 2 * acc_b - 3 * acc_m;
  miss = acc_m;
}

// General B (3 or 4): groups of B words packed back to back into pairs.
__device__ __forceinline__ void count_generic(const uint2* __restrict__ col, int n_groups, int B,
                                              int& num, int& miss) {
  const uint32_t* base = reinterpret_cast<const uint32_t*>(col);
  const int all = (1 << B) - 1;
  int acc = 0, accm = 0;
  for (int g = 0; g < n_groups; ++g) {
    uint32_t andw = 0xffffffffu;
    int s = 0;
    for (int b = 0; b < B; ++b) {
      const int word = g * B + b;
      // word w of this lane: pair (w>>1) is 32 lanes * 2 words further on
      uint32_t x = __ldg(base + (size_t)(word >> 1) * (kTile * 2) + (word & 1));
      s += __popc(x) << b;
      andw &= x;
    }
    const int m = __popc(andw);
    acc += s - all * m;
    accm += m;
  }
  num = acc;
  miss = accm;
}

struct SiteParams {
  sai_layout lay;
  const uint2* packed;  // tile 0
  int64_t tile0, n_tiles, n_tiles_total;
  uint32_t* mask_u;
  uint32_t* mask_q;
  double* qval;
  int64_t qval_stride;
  int32_t* num;
  int32_t* called;
  int64_t count_stride;
};

constexpr int kSiteWarps = 8;

template <int MODE, bool FUSED>
__global__ void __launch_bounds__(kSiteWarps * 32, MODE == 1 ? 3 : 1)
    k_site(const __grid_constant__ SiteParams P, const __grid_constant__ JobBlock JB) {
  extern __shared__ int s_counts[];  // [warp][2][n_pops][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pops = P.lay.n_pops;
  int* s_num = s_counts + warp * (2 * n_pops * kTile);
  int* s_cal = s_num + n_pops * kTile;
  const int64_t pps = P.lay.pairs_per_site;

  for (int64_t t = (int64_t)blockIdx.x * kSiteWarps + warp; t < P.n_tiles;
       t += (int64_t)gridDim.x * kSiteWarps) {
    const int64_t T = P.tile0 + t;
    const uint2* tile = P.packed + (size_t)T * pps * kTile + lane;
    for (int pi = 0; pi < n_pops; ++pi) {
      const sai_pop_layout& L = P.lay.pop[pi];
      const uint2* col = tile + (size_t)L.pair_off * kTile;
      int num, miss;
      if (L.bits == 2)
        count_b2<MODE>(col, L.n_pairs, num, miss);
      else
        count_generic(col, L.n_groups, L.bits, num, miss);
      const int called = L.n_groups * 32 - miss;
      s_num[pi * kTile + lane] = num;
      s_cal[pi * kTile + lane] = called;
      if (P.num) {
        const int64_t site = T * kTile + lane;
        P.num[(size_t)pi * P.count_stride + site] = num;
        P.called[(size_t)pi * P.count_stride + site] = called;
      }
    }
    if (FUSED) {
      const int64_t site = T * kTile + lane;
      for (int j = 0; j < JB.n_jobs; ++j) {
        const sai_job& J = JB.job[j];
        SiteFlags f = eval_site(
            J, P.lay, [&](int pop) { return s_num[pop * kTile + lane]; },
            [&](int pop) { return s_cal[pop * kTile + lane]; });
        const uint32_t mu = __ballot_sync(0xffffffffu, f.u);
        const uint32_t mq = __ballot_sync(0xffffffffu, f.q);
        if (lane == 0) {
          P.mask_u[(size_t)j * P.n_tiles_total + T] = mu;
          P.mask_q[(size_t)j * P.n_tiles_total + T] = mq;
        }
        if (f.q) P.qval[(size_t)j * P.qval_stride + site] = f.q_tgt_freq;
      }
    }
  }
}

struct CountFlagParams {
  sai_layout lay;
  const int32_t* num;
  const int32_t* called;
  int64_t count_stride;
  int64_t n_sites;
  uint32_t* mask_u;
  uint32_t* mask_q;
  double* qval;
  int64_t qval_stride;
  int64_t n_tiles_total;
};

// Site conditions from cached counts: one lane per site, one warp per tile.
__global__ void __launch_bounds__(256)
    k_flags_from_counts(const __grid_constant__ CountFlagParams P,
                        const __grid_constant__ JobBlock JB) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t T = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; T < P.n_tiles_total;
       T += warps) {
    const int64_t site = T * kTile + lane;
    const bool live = site < P.n_sites;
    for (int j = 0; j < JB.n_jobs; ++j) {
      const sai_job& J = JB.job[j];
      SiteFlags f = eval_site(
          J, P.lay,
          [&](int pop) { return live ? P.num[(size_t)pop * P.count_stride + site] : 0; },
          [&](int pop) { return live ? P.called[(size_t)pop * P.count_stride + site] : 0; });
      const uint32_t mu = __ballot_sync(0xffffffffu, f.u);
      const uint32_t mq = __ballot_sync(0xffffffffu, f.q);
      if (lane == 0) {
        P.mask_u[(size_t)j * P.n_tiles_total + T] = mu;
        P.mask_q[(size_t)j * P.n_tiles_total + T] = mq;
      }
      if (f.q) P.qval[(size_t)j * P.qval_stride + site] = f.q_tgt_freq;
    }
  }
}

template <int MODE, bool FUSED>
static int launch_site(const SiteParams& P, const JobBlock& JB, cudaStream_t st) {
  if (P.n_tiles == 0) return SAI_OK;
  const size_t smem = (size_t)kSiteWarps * 2 * P.lay.n_pops * kTile * sizeof(int);
  int occ = 0;
  SAI_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_site<MODE, FUSED>,
                                                               kSiteWarps * 32, smem));
  if (occ < 1) occ = 1;
  int64_t want = (P.n_tiles + kSiteWarps - 1) / kSiteWarps;
  int64_t cap = (int64_t)sm_count() * occ;
  int grid = (int)(want < cap ? want : cap);
  k_site<MODE, FUSED><<<grid, kSiteWarps * 32, smem, st>>>(P, JB);
  SAI_CUDA_CHECK(cudaGetLastError());
  return SAI_OK;
}

}  // namespace sai

using namespace sai;

extern "C" {

int sai_site_counts(const sai_layout* lay, const void* d_packed, int64_t tile0, int64_t n_tiles,
                    int32_t* d_num, int32_t* d_called, int64_t stride, int32_t variant,
                    void* stream) {
  if (int rc = validate_layout(lay)) return rc;
  SAI_REQUIRE(d_packed && d_num && d_called, "NULL device pointer");
  SAI_REQUIRE(tile0 >= 0 && n_tiles >= 0 && stride >= (tile0 + n_tiles) * kTile, "bad tile range / stride");
  SiteParams P{};
  P.lay = *lay;
  P.packed = static_cast<const uint2*>(d_packed);
  P.tile0 = tile0;
  P.n_tiles = n_tiles;
  P.n_tiles_total = tile0 + n_tiles;
  P.num = d_num;
  P.called = d_called;
  P.count_stride = stride;
  JobBlock JB{};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (variant == 2) return launch_site<2, false>(P, JB, st);
  if (variant == 1) return launch_site<1, false>(P, JB, st);
  return launch_site<0, false>(P, JB, st);
}

int sai_site_flags(const sai_layout* lay, const void* d_packed, int64_t tile0, int64_t n_tiles,
                   int64_t n_tiles_total, const sai_job* jobs, int32_t n_jobs, uint32_t* d_mask_u,
                   uint32_t* d_mask_q, double* d_qval, int64_t qval_stride, int32_t* d_num,
                   int32_t* d_called, int64_t count_stride, int32_t variant, void* stream) {
  if (int rc = validate_layout(lay)) return rc;
  if (int rc = validate_jobs(lay, jobs, n_jobs)) return rc;
  SAI_REQUIRE(d_packed && d_mask_u && d_mask_q && d_qval, "NULL device pointer");
  SAI_REQUIRE(tile0 >= 0 && n_tiles >= 0 && tile0 + n_tiles <= n_tiles_total, "bad tile range");
  SAI_REQUIRE(qval_stride >= n_tiles_total * kTile, "qval_stride too small");
  SAI_REQUIRE((d_num == nullptr) == (d_called == nullptr), "d_num/d_called must both be set or NULL");
  SAI_REQUIRE(!d_num || count_stride >= n_tiles_total * kTile, "count_stride too small");
  SiteParams P{};
  P.lay = *lay;
  P.packed = static_cast<const uint2*>(d_packed);
  P.tile0 = tile0;
  P.n_tiles = n_tiles;
  P.n_tiles_total = n_tiles_total;
  P.mask_u = d_mask_u;
  P.mask_q = d_mask_q;
  P.qval = d_qval;
  P.qval_stride = qval_stride;
  P.num = d_num;
  P.called = d_called;
  P.count_stride = count_stride;
  JobBlock JB{};
  JB.n_jobs = n_jobs;
  for (int j = 0; j < n_jobs; ++j) JB.job[j] = jobs[j];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (variant == 2) return launch_site<2, true>(P, JB, st);
  if (variant == 1) return launch_site<1, true>(P, JB, st);
  return launch_site<0, true>(P, JB, st);
}

int sai_flags_from_counts(const sai_layout* lay, const int32_t* d_num, const int32_t* d_called,
                          int64_t count_stride, int64_t n_sites, const sai_job* jobs,
                          int32_t n_jobs, uint32_t* d_mask_u, uint32_t* d_mask_q, double* d_qval,
                          int64_t qval_stride, void* stream) {
  if (int rc = validate_layout(lay)) return rc;
  if (int rc = validate_jobs(lay, jobs, n_jobs)) return rc;
  SAI_REQUIRE(d_num && d_called && d_mask_u && d_mask_q && d_qval, "NULL device pointer");
  SAI_REQUIRE(n_sites >= 0 && count_stride >= n_sites && qval_stride >= n_sites, "bad strides");
  if (n_sites == 0) return SAI_OK;
  CountFlagParams P{};
  P.lay = *lay;
  P.num = d_num;
  P.called = d_called;
  P.count_stride = count_stride;
  P.n_sites = n_sites;
  P.mask_u = d_mask_u;
  P.mask_q = d_mask_q;
  P.qval = d_qval;
  P.qval_stride = qval_stride;
  P.n_tiles_total = sai_num_tiles(n_sites);
  JobBlock JB{};
  JB.n_jobs = n_jobs;
  for (int j = 0; j < n_jobs; ++j) JB.job[j] = jobs[j];
  const int64_t want = (P.n_tiles_total + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(want < cap ? want : cap);
  k_flags_from_counts<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(P, JB);
  SAI_CUDA_CHECK(cudaGetLastError());
  return SAI_OK;
}

}  // extern "C"
