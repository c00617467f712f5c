// Library housekeeping, packed-layout arithmetic and the host-side encoder.
//
// The encoder replaces the reference's in-memory representation (int64
// per-individual allele sums, sai/utils/utils.py:405-410) with the tiled
// bit-plane layout described in include/sai_b200.h.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"
#include "host_pack.h"
#include "scratch.cuh"

namespace sai {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int scratch_alloc(void** out, size_t bytes, cudaStream_t st) {
  static cudaMemPool_t pools[64] = {nullptr};
  static std::mutex mu;
  int dev = 0;
  SAI_CUDA_CHECK(cudaGetDevice(&dev));
  SAI_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!pools[dev]) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      SAI_CUDA_CHECK(cudaMemPoolCreate(&pools[dev], &props));
      uint64_t never = ~0ull;
      SAI_CUDA_CHECK(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &never));
    }
  }
  cudaError_t e = cudaMallocFromPoolAsync(out, bytes ? bytes : 256, pools[dev], st);
  if (e != cudaSuccess) {
    set_error("scratch allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? SAI_E_NOMEM : SAI_E_CUDA;
  }
  return SAI_OK;
}

int scratch_free(void* p, cudaStream_t st) {
  if (p) SAI_CUDA_CHECK(cudaFreeAsync(p, st));
  return SAI_OK;
}

int validate_layout(const sai_layout* lay) {
  SAI_REQUIRE(lay != nullptr, "layout is NULL");
  SAI_REQUIRE(lay->n_pops >= 1 && lay->n_pops <= SAI_MAX_POPS, "n_pops %d outside [1,%d]",
              lay->n_pops, SAI_MAX_POPS);
  int32_t at = 0;
  for (int p = 0; p < lay->n_pops; ++p) {
    const sai_pop_layout& L = lay->pop[p];
    SAI_REQUIRE(L.n_samples >= 1, "population %d has no samples", p);
    SAI_REQUIRE(L.ploidy >= 1, "ploidy must be a positive integer.");
    SAI_REQUIRE(L.bits >= 2 && L.bits <= SAI_MAX_BITS, "population %d: bits %d outside [2,%d]", p, L.bits, SAI_MAX_BITS);
    SAI_REQUIRE(L.n_groups == (L.n_samples + 31) / 32, "population %d: bad n_groups", p);
    SAI_REQUIRE(L.n_pairs == (L.n_groups * L.bits + 1) / 2, "population %d: bad n_pairs", p);
    SAI_REQUIRE(L.pair_off == at, "population %d: bad pair_off", p);
    at += L.n_pairs;
  }
  SAI_REQUIRE(lay->pairs_per_site == at, "bad pairs_per_site");
  return SAI_OK;
}

int validate_jobs(const sai_layout* lay, const sai_job* jobs, int32_t n_jobs) {
  SAI_REQUIRE(jobs != nullptr && n_jobs >= 1 && n_jobs <= SAI_MAX_JOBS, "n_jobs %d outside [1,%d]",
              n_jobs, SAI_MAX_JOBS);
  for (int j = 0; j < n_jobs; ++j) {
    const sai_job& J = jobs[j];
    SAI_REQUIRE(J.ref_pop >= 0 && J.ref_pop < lay->n_pops, "job %d: bad ref_pop", j);
    SAI_REQUIRE(J.tgt_pop >= 0 && J.tgt_pop < lay->n_pops, "job %d: bad tgt_pop", j);
    SAI_REQUIRE(J.n_src >= 1 && J.n_src <= SAI_MAX_SRC, "job %d: n_src %d outside [1,%d]", j,
                J.n_src, SAI_MAX_SRC);
    for (int k = 0; k < J.n_src; ++k)
      SAI_REQUIRE(J.src_pop[k] >= 0 && J.src_pop[k] < lay->n_pops, "job %d: bad src_pop[%d]", j, k);
    const sai_cond* cs[2] = {&J.u, &J.q};
    for (const sai_cond* c : cs) {
      if (!c->enabled) continue;
      // same ranges the reference enforces (sai/stats/stat_utils.py:99-108)
      SAI_REQUIRE(c->w >= 0.0 && c->w <= 1.0, "Parameters w must be within the range [0, 1].");
      for (int k = 0; k < J.n_src; ++k) {
        SAI_REQUIRE(c->y[k] >= 0.0 && c->y[k] <= 1.0, "Invalid value in y_list: %g. within the range [0, 1].",
                    c->y[k]);
        SAI_REQUIRE(c->op[k] >= SAI_OP_EQ && c->op[k] <= SAI_OP_GE, "Invalid operator in y_list");
      }
    }
    if (J.q.enabled)
      SAI_REQUIRE(J.quantile >= 0.0 && J.quantile <= 1.0, "Quantiles must be in the range [0, 1]");
  }
  return SAI_OK;
}

}  // namespace sai

using namespace sai;

extern "C" {

const char* sai_version(void) { return "sai_b200 0.1 (sm_100a)"; }
const char* sai_last_error(void) { return g_err; }

int32_t sai_bits_for_max_value(int32_t max_value) {
  // codes 0..max_value plus the all-ones missing code
  int32_t b = 2;
  while (((1 << b) - 1) <= max_value) ++b;
  return b;
}

int sai_layout_init(sai_layout* lay, int32_t n_pops, const int32_t* n_samples,
                    const int32_t* ploidy, const int32_t* bits) {
  SAI_REQUIRE(lay && n_samples && ploidy, "NULL argument");
  SAI_REQUIRE(n_pops >= 1 && n_pops <= SAI_MAX_POPS, "n_pops %d outside [1,%d]", n_pops,
              SAI_MAX_POPS);
  memset(lay, 0, sizeof(*lay));
  lay->n_pops = n_pops;
  int32_t at = 0;
  for (int p = 0; p < n_pops; ++p) {
    sai_pop_layout& L = lay->pop[p];
    SAI_REQUIRE(ploidy[p] >= 1, "ploidy must be a positive integer.");
    SAI_REQUIRE(n_samples[p] >= 1, "population %d has no samples", p);
    L.n_samples = n_samples[p];
    L.ploidy = ploidy[p];
    L.bits = (bits && bits[p] > 0) ? bits[p] : sai_bits_for_max_value(ploidy[p]);
    SAI_REQUIRE(L.bits >= 2 && L.bits <= SAI_MAX_BITS, "population %d needs %d bit-planes (max %d)", p, L.bits, SAI_MAX_BITS);
    L.n_groups = (L.n_samples + 31) / 32;
    L.n_pairs = (L.n_groups * L.bits + 1) / 2;
    L.pair_off = at;
    at += L.n_pairs;
  }
  lay->pairs_per_site = at;
  return SAI_OK;
}

int64_t sai_num_tiles(int64_t n_sites) { return (n_sites + kTile - 1) / kTile; }

uint64_t sai_packed_bytes(const sai_layout* lay, int64_t n_sites) {
  if (!lay || n_sites < 0) return 0;
  return (uint64_t)sai_num_tiles(n_sites) * (uint64_t)lay->pairs_per_site * kTile * 8ull;
}

// word w (0-based within the population's word sequence) of site s of tile T
static inline uint32_t* word_ptr(const sai_layout* lay, const sai_pop_layout& L, uint8_t* packed,
                                 int64_t tile, int site_in_tile, int word) {
  int64_t pair = (int64_t)L.pair_off + (word >> 1);
  int64_t off = ((tile * lay->pairs_per_site + pair) * kTile + site_in_tile) * 8 + (word & 1) * 4;
  return reinterpret_cast<uint32_t*>(packed + off);
}

int sai_pack_i8(const sai_layout* lay, int32_t pop, const int8_t* gt, int64_t n_sites,
                int64_t row_stride, uint8_t* packed, int32_t n_threads) {
  return sai_pack_i8_isa(lay, pop, gt, n_sites, row_stride, packed, n_threads, 0);
}

int sai_pack_i8_isa(const sai_layout* lay, int32_t pop, const int8_t* gt, int64_t n_sites,
                    int64_t row_stride, uint8_t* packed, int32_t n_threads, int32_t isa) {
  if (int rc = validate_layout(lay)) return rc;
  SAI_REQUIRE(pop >= 0 && pop < lay->n_pops, "bad population index %d", pop);
  SAI_REQUIRE(gt && packed && n_sites >= 0, "NULL argument");
  const sai_pop_layout& L = lay->pop[pop];
  SAI_REQUIRE(row_stride >= L.n_samples, "row_stride smaller than n_samples");
  const int64_t n_tiles = sai_num_tiles(n_sites);
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, n_tiles / 8));
  std::atomic<int> domain_err{0};
  // blocks of tiles handed out dynamically (the vector row packers live in pack_simd.cpp)
  const int64_t block = 64;
  std::atomic<int64_t> next{0};
  auto work = [&]() {
    bool bad = false;
    for (int64_t t0 = next.fetch_add(block); t0 < n_tiles; t0 = next.fetch_add(block))
      bad |= pack_tiles_i8(*lay, pop, gt, n_sites, row_stride, t0, std::min(n_tiles, t0 + block), 0, packed, isa);
    if (bad) domain_err.store(1, std::memory_order_relaxed);
  };
  if (n_threads <= 1) {
    work();
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  if (domain_err.load()) {
    set_error("population %d: a genotype value does not fit %d bit-planes", pop, L.bits);
    return SAI_E_DOMAIN;
  }
  return SAI_OK;
}

int sai_pack_i8_all(const sai_layout* lay, const int8_t* const* gt, const int64_t* row_stride, int64_t n_sites,
                    uint8_t* packed, int32_t n_threads) {
  if (int rc = validate_layout(lay)) return rc;
  SAI_REQUIRE(gt && row_stride && packed && n_sites >= 0, "NULL argument");
  for (int p = 0; p < lay->n_pops; ++p) {
    SAI_REQUIRE(n_sites == 0 || gt[p], "NULL genotype matrix of population %d", p);
    SAI_REQUIRE(row_stride[p] >= lay->pop[p].n_samples, "row_stride of population %d smaller than n_samples", p);
  }
  const int64_t n_tiles = sai_num_tiles(n_sites);
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, n_tiles / 8));
  std::atomic<int> domain_err{0};
  const int64_t block = 32;
  std::atomic<int64_t> next{0};
  auto work = [&]() {
    bool bad = false;
    for (int64_t t0 = next.fetch_add(block); t0 < n_tiles; t0 = next.fetch_add(block))
      bad |= pack_tiles_i8_all(*lay, gt, row_stride, n_sites, t0, std::min(n_tiles, t0 + block), 0, packed, 0);
    if (bad) domain_err.store(1, std::memory_order_relaxed);
  };
  if (n_threads <= 1) {
    work();
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  if (domain_err.load()) {
    set_error("a genotype value does not fit the bit-planes of its population");
    return SAI_E_DOMAIN;
  }
  return SAI_OK;
}

const char* sai_pack_isa(void) { return pack_isa(); }

// Negative-value table of one int8 population matrix (DD, include/sai_b200.h "N4"): every entry v < 0 in
// row-major order.  Two parallel passes over site blocks: count, then fill at the prefix offsets.
int64_t sai_neg_table_i8(const int8_t* gt, int64_t n_sites, int32_t n_samples, int64_t row_stride,
                         int32_t* site, int32_t* ind, int32_t* val, int64_t cap, int32_t n_threads) {
  if (!gt || n_sites < 0 || n_samples < 1 || row_stride < n_samples || cap < 0 || (cap > 0 && (!site || !ind || !val))) {
    set_error("sai_neg_table_i8: bad argument");
    return SAI_E_ARG;
  }
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  const int64_t n_blocks = std::max<int64_t>(1, std::min<int64_t>(n_threads * 4, n_sites / 256));
  const int64_t per = (n_sites + n_blocks - 1) / n_blocks;
  std::vector<int64_t> count(n_blocks + 1, 0);
  auto run = [&](auto&& fn) {
    std::atomic<int64_t> next{0};
    auto work = [&]() {
      for (int64_t b = next.fetch_add(1); b < n_blocks; b = next.fetch_add(1)) fn(b);
    };
    const int nt = (int)std::min<int64_t>(n_threads, n_blocks);
    if (nt <= 1) {
      work();
      return;
    }
    std::vector<std::thread> th;
    for (int i = 0; i < nt; ++i) th.emplace_back(work);
    for (auto& t : th) t.join();
  };
  const uint64_t signs = 0x8080808080808080ull;
  run([&](int64_t b) {
    int64_t c = 0;
    for (int64_t s = b * per; s < std::min(n_sites, (b + 1) * per); ++s) {
      const int8_t* row = gt + s * row_stride;
      int i = 0;
      for (; i + 8 <= n_samples; i += 8) {
        uint64_t x;
        memcpy(&x, row + i, 8);
        c += __builtin_popcountll(x & signs);
      }
      for (; i < n_samples; ++i) c += row[i] < 0;
    }
    count[b + 1] = c;
  });
  for (int64_t b = 0; b < n_blocks; ++b) count[b + 1] += count[b];
  const int64_t total = count[n_blocks];
  if (cap == 0 || total > cap) return total;  // counting call / buffers too small: nothing written
  run([&](int64_t b) {
    int64_t at = count[b];
    for (int64_t s = b * per; s < std::min(n_sites, (b + 1) * per); ++s) {
      const int8_t* row = gt + s * row_stride;
      int i = 0;
      for (; i + 8 <= n_samples; i += 8) {
        uint64_t x;
        memcpy(&x, row + i, 8);
        if (!(x & signs)) continue;
        for (int k = 0; k < 8; ++k)
          if (row[i + k] < 0) site[at] = (int32_t)s, ind[at] = i + k, val[at] = row[i + k], ++at;
      }
      for (; i < n_samples; ++i)
        if (row[i] < 0) site[at] = (int32_t)s, ind[at] = i, val[at] = row[i], ++at;
    }
  });
  return total;
}

int sai_unpack_i8(const sai_layout* lay, int32_t pop, const uint8_t* packed, int64_t n_sites_total,
                  int64_t site0, int64_t n, int8_t* gt, int64_t row_stride) {
  if (int rc = validate_layout(lay)) return rc;
  SAI_REQUIRE(pop >= 0 && pop < lay->n_pops, "bad population index %d", pop);
  SAI_REQUIRE(packed && gt && site0 >= 0 && n >= 0 && site0 + n <= sai_num_tiles(n_sites_total) * kTile,
              "bad site range");
  const sai_pop_layout& L = lay->pop[pop];
  SAI_REQUIRE(row_stride >= L.n_samples, "row_stride smaller than n_samples");
  const int B = L.bits;
  const int miss_code = (1 << B) - 1;
  for (int64_t site = site0; site < site0 + n; ++site) {
    const int64_t T = site / kTile;
    const int s = (int)(site % kTile);
    int8_t* row = gt + (site - site0) * row_stride;
    for (int g = 0; g < L.n_groups; ++g) {
      uint32_t plane[SAI_MAX_BITS] = {0};
      for (int b = 0; b < B; ++b)
        plane[b] = *word_ptr(lay, L, const_cast<uint8_t*>(packed), T, s, g * B + b);
      const int i0 = g * 32;
      const int cnt = std::min(32, L.n_samples - i0);
      for (int i = 0; i < cnt; ++i) {
        int code = 0;
        for (int b = 0; b < B; ++b) code |= (int)((plane[b] >> i) & 1u) << b;
        row[i0 + i] = (int8_t)(code == miss_code ? -1 : code);
      }
    }
  }
  return SAI_OK;
}

}  // extern "C"
