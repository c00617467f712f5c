// "zt": zero-suppressed packed tiles -- the wire format between host memory and HBM.
//
// End to end the scoring path is bound by the host->device copy of the packed
// tiles (PCIe Gen5: ~53 GB/s against 6.4 TB/s for the genotype pass), and most
// of a genotype matrix is the homozygous-reference code 0: with a realistic
// site-frequency spectrum most sites are rare variants.  The host therefore
// sends a two-level zero-suppressed form of every tile and a kernel rebuilds
// the dense tiles in HBM (include/sai_b200.h "Packed genotype layout"); the
// genotype pass and everything behind it run on the dense tiles, unchanged.
//
// Record of tile T (P = pairs_per_site rows of 32 pairs, one pair per site) at
// byte offset tile_off[T] & ~SAI_ZT_RAW of the stream (8-byte aligned):
//     u32 n1                 number of non-zero pairs of the tile
//     u32 nz[P]              bit s of nz[r]: pair (row r, site s) is non-zero
//     u8  mask[n1]           per non-zero pair, in (r, s) order: bit k = byte k of the pair is non-zero
//     u8  data[n2]           the non-zero bytes, in the same order, ascending k
// "Pair" here is the stored pair XOR the row's padding constant (the unused
// individuals of a population's last group are coded missing, i.e. all-ones in
// every plane; XOR-ing the constant out makes those bytes zero too).
// A tile whose record would not be smaller than the dense tile is stored as its
// P*256 raw bytes and flagged with SAI_ZT_RAW (bit 63) in tile_off[T].
// tile_off[n_tiles] = total stream length.
#include <string.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include "common.cuh"
#include "zt_format.cuh"
#include "zt_simd.h"

namespace sai {

constexpr uint64_t kZtRaw = 1ull << 63;
constexpr int kZtWarps = 8;

static void run_parallel(int64_t n, int n_threads, const std::function<void(int64_t, int64_t)>& fn) {
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, n / 64));
  if (n_threads <= 1) {
    fn(0, n);
    return;
  }
  std::vector<std::thread> th;
  const int64_t per = (n + n_threads - 1) / n_threads;
  for (int i = 0; i < n_threads; ++i) {
    const int64_t a = i * per, b = std::min<int64_t>(n, a + per);
    if (a < b) th.emplace_back(fn, a, b);
  }
  for (auto& t : th) t.join();
}

struct ZtDecodeParams {
  sai_layout lay;
  const uint8_t* stream;
  unsigned long long stream_bytes;  // loads never reach past it
  const unsigned long long* tile_off;
  int64_t tile0, n_tiles;
  unsigned long long* packed;  // dense tiles, tile 0
};

// Shared tables of the decoder.  For a byte mask m (which of a pair's 8 bytes are present) the
// packed bytes c0, c1, ... go to the set positions of m: two PRMT selectors (byte k of the
// output <- packed byte rank_k) and two byte masks that clear the absent positions.
struct ZtTables {
  uint2 sel[256];   // PRMT selectors for output bytes 0-3 / 4-7
  uint2 keep[256];  // 0xff in every present byte
};

// one warp per tile, lane == site
__global__ void __launch_bounds__(kZtWarps * 32) k_zt_decode(const __grid_constant__ ZtDecodeParams P) {
  extern __shared__ __align__(16) unsigned char s_zt[];
  ZtTables& TB = *reinterpret_cast<ZtTables*>(s_zt);
  unsigned long long* s_pad = reinterpret_cast<unsigned long long*>(s_zt + sizeof(ZtTables));  // [pairs_per_site]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pps = P.lay.pairs_per_site;
  for (int m = threadIdx.x; m < 256; m += blockDim.x) {
    uint32_t sel[2] = {0u, 0u}, keep[2] = {0u, 0u};
    int rank = 0;
    for (int k = 0; k < 8; ++k) {
      if ((m >> k) & 1) {
        sel[k >> 2] |= (uint32_t)rank << (4 * (k & 3));
        keep[k >> 2] |= 0xffu << (8 * (k & 3));
        ++rank;
      }
    }
    TB.sel[m] = make_uint2(sel[0], sel[1]);
    TB.keep[m] = make_uint2(keep[0], keep[1]);
  }
  for (int r = threadIdx.x; r < pps; r += blockDim.x) s_pad[r] = pad_constant(P.lay, r);
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
  const uint8_t* const stream_end = P.stream + P.stream_bytes;
  for (int64_t t = (int64_t)blockIdx.x * kZtWarps + warp; t < P.n_tiles; t += (int64_t)gridDim.x * kZtWarps) {
    const int64_t T = P.tile0 + t;
    const unsigned long long off = __ldg(P.tile_off + T);
    unsigned long long* out = P.packed + (size_t)T * pps * kTile + lane;
    const uint8_t* rec = P.stream + (off & ~kZtRaw);
    if (off & kZtRaw) {
      const unsigned long long* in = reinterpret_cast<const unsigned long long*>(rec) + lane;
      for (int r = 0; r < pps; ++r) out[(size_t)r * kTile] = __ldg(in + (size_t)r * kTile);
      continue;
    }
    const uint32_t n1 = __ldg(reinterpret_cast<const uint32_t*>(rec));
    const uint32_t* nz = reinterpret_cast<const uint32_t*>(rec + 4);
    const uint8_t* mask = rec + 4 + 4 * (size_t)pps;
    const uint8_t* data = mask + n1;
    uint32_t base1 = 0, base2 = 0;
    for (int r0 = 0; r0 < pps; r0 += 32) {
      const uint32_t mine = r0 + lane < pps ? __ldg(nz + r0 + lane) : 0u;
      const int rows = pps - r0 < 32 ? pps - r0 : 32;
      for (int rr = 0; rr < rows; ++rr) {
        const uint32_t w = __shfl_sync(0xffffffffu, mine, rr);
        const int r = r0 + rr;
        const unsigned long long padc = s_pad[r];
        if (w == 0u) {
          out[(size_t)r * kTile] = padc;
          continue;
        }
        const bool has = (w >> lane) & 1u;
        const uint32_t m = has ? (uint32_t)__ldg(mask + base1 + __popc(w & lt)) : 0u;
        const int nb = __popc(m);
        int incl = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += up;
        }
        unsigned long long v = 0;
        if (nb) {
          // the lane's nb <= 8 bytes start at p (any alignment): two aligned 8-byte loads + funnel shift
          const uint8_t* p = data + base2 + (incl - nb);
          const uintptr_t a = reinterpret_cast<uintptr_t>(p);
          const unsigned long long* q = reinterpret_cast<const unsigned long long*>(a & ~(uintptr_t)7);
          const int sh = (int)(a & 7) * 8;
          const unsigned long long lo = __ldg(q);
          const unsigned long long hi =
              (sh && reinterpret_cast<const uint8_t*>(q + 2) <= stream_end) ? __ldg(q + 1) : 0ull;
          const unsigned long long c = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
          const uint32_t c0 = (uint32_t)c, c1 = (uint32_t)(c >> 32);
          const uint2 sel = TB.sel[m], keep = TB.keep[m];
          const uint32_t v0 = __byte_perm(c0, c1, sel.x) & keep.x;
          const uint32_t v1 = __byte_perm(c0, c1, sel.y) & keep.y;
          v = (unsigned long long)v0 | ((unsigned long long)v1 << 32);
        }
        out[(size_t)r * kTile] = v ^ padc;
        base1 += __popc(w);
        base2 += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
  }
}

}  // namespace sai

using namespace sai;

extern "C" {

uint64_t sai_zt_bound(const sai_layout* lay, int64_t n_sites) {
  if (!lay || n_sites < 0) return 0;
  // no record is larger than its dense tile
  return sai_packed_bytes(lay, n_sites) + 8;
}

int64_t sai_zt_encode_isa(const sai_layout* lay, const uint8_t* packed, int64_t n_sites, uint8_t* out,
                          uint64_t out_cap, uint64_t* tile_off, int32_t n_threads, int32_t isa) {
  if (int rc = validate_layout(lay)) return rc;
  SAI_REQUIRE(n_sites >= 0 && tile_off, "bad argument");
  const int64_t n_tiles = sai_num_tiles(n_sites);
  const int P = lay->pairs_per_site;
  const size_t dense = (size_t)P * kTile * 8;
  tile_off[0] = 0;
  if (n_tiles == 0) return 0;
  SAI_REQUIRE(packed && out, "NULL argument");
  std::vector<uint64_t> padc(P);
  for (int r = 0; r < P; ++r) padc[r] = pad_constant(*lay, r);
  // pass 1: record sizes (tile_off[T + 1] = size of tile T, raw flag kept aside)
  std::vector<uint8_t> raw(n_tiles);
  run_parallel(n_tiles, n_threads, [&](int64_t t0, int64_t t1) {
    for (int64_t T = t0; T < t1; ++T) {
      const uint64_t* d = reinterpret_cast<const uint64_t*>(packed + (size_t)T * dense);
      const size_t rec = (zt_tile_size(d, P, padc.data(), isa) + 7) & ~size_t(7);
      raw[T] = rec >= dense;
      tile_off[T + 1] = raw[T] ? dense : rec;
    }
  });
  uint64_t at = 0;
  for (int64_t T = 0; T < n_tiles; ++T) {
    const uint64_t sz = tile_off[T + 1];
    tile_off[T] = at | (raw[T] ? kZtRaw : 0ull);
    at += sz;
  }
  tile_off[n_tiles] = at;
  if (at > out_cap) {
    set_error("zt stream needs %llu bytes, buffer has %llu", (unsigned long long)at, (unsigned long long)out_cap);
    return SAI_E_CAPACITY;
  }
  // pass 2: write the records (built in a private buffer: the vector encoder stores whole vectors)
  run_parallel(n_tiles, n_threads, [&](int64_t t0, int64_t t1) {
    std::vector<uint8_t> tmp(zt_tmp_cap(P)), rec(zt_record_cap(P));
    for (int64_t T = t0; T < t1; ++T) {
      const uint8_t* src = packed + (size_t)T * dense;
      uint8_t* dst = out + (tile_off[T] & ~kZtRaw);
      if (tile_off[T] & kZtRaw) {
        memcpy(dst, src, dense);
        continue;
      }
      const size_t used = zt_encode_tile(reinterpret_cast<const uint64_t*>(src), P, padc.data(), rec.data(), tmp.data(), isa);
      const size_t end = (tile_off[T + 1] & ~kZtRaw) - (tile_off[T] & ~kZtRaw);
      memset(rec.data() + used, 0, end - used);
      memcpy(dst, rec.data(), end);
    }
  });
  return (int64_t)at;
}

int64_t sai_zt_encode(const sai_layout* lay, const uint8_t* packed, int64_t n_sites, uint8_t* out,
                      uint64_t out_cap, uint64_t* tile_off, int32_t n_threads) {
  return sai_zt_encode_isa(lay, packed, n_sites, out, out_cap, tile_off, n_threads, 0);
}

const char* sai_zt_isa(void) { return zt_isa(); }

int64_t sai_zt_pack_i8(const sai_layout* lay, const int8_t* const* gt, const int64_t* row_stride, int64_t n_sites,
                       uint8_t* out, uint64_t out_cap, uint64_t* tile_off, int32_t n_threads) {
  if (int rc = validate_layout(lay)) return rc;
  SAI_REQUIRE(gt && row_stride && tile_off && n_sites >= 0, "NULL argument");
  for (int p = 0; p < lay->n_pops; ++p) {
    SAI_REQUIRE(n_sites == 0 || gt[p], "NULL genotype matrix of population %d", p);
    SAI_REQUIRE(row_stride[p] >= lay->pop[p].n_samples, "row_stride of population %d smaller than n_samples", p);
  }
  const int64_t n_tiles = sai_num_tiles(n_sites);
  tile_off[0] = 0;
  if (n_tiles == 0) return 0;
  SAI_REQUIRE(out, "NULL argument");
  const int P = lay->pairs_per_site;
  const size_t tile_bytes = (size_t)P * kTile * 8;
  std::vector<uint64_t> padc(P);
  for (int r = 0; r < P; ++r) padc[r] = pad_constant(*lay, r);
  // blocks of tiles are encoded into private arenas (their lengths are not known in advance),
  // then laid end to end in `out`.  Arenas grow in 16 MB chunks: one allocation per ~70 blocks
  // instead of one per block (16 threads allocating 200 KB buffers serialise on the address space)
  const int64_t block = 32, n_blocks = (n_tiles + block - 1) / block;
  const size_t block_cap = (size_t)block * tile_bytes + 64;
  const size_t chunk_bytes = std::max<size_t>(16u << 20, block_cap);
  struct Piece {
    const uint8_t* p;
    size_t n;
  };
  std::vector<Piece> recs(n_blocks);
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  n_threads = (int)std::min<int64_t>(n_threads, n_blocks);
  std::vector<std::vector<std::unique_ptr<uint8_t[]>>> arenas(n_threads);
  std::atomic<int64_t> next{0};
  std::atomic<int> domain_err{0};
  auto encode = [&](int me) {
    ZtBlockScratch sc(P);
    uint8_t* at = nullptr;  // 64-byte aligned write position in the current chunk
    size_t room = 0;
    bool bad = false;
    for (int64_t b = next.fetch_add(1); b < n_blocks; b = next.fetch_add(1)) {
      if (room < block_cap) {
        arenas[me].emplace_back(new uint8_t[chunk_bytes + 64]);
        at = arenas[me].back().get();
        at += (64 - reinterpret_cast<uintptr_t>(at) % 64) % 64;
        room = chunk_bytes;
      }
      const int64_t t0 = b * block, t1 = std::min(n_tiles, t0 + block);
      const size_t used = zt_pack_block_i8(*lay, gt, row_stride, n_sites, t0, t1, padc.data(), at, 0, tile_off, sc, false, &bad);
      recs[b] = Piece{at, used};
      const size_t step = (used + 63) & ~size_t(63);
      at += step;
      room -= step;
    }
    if (bad) domain_err.store(1, std::memory_order_relaxed);
  };
  {
    std::vector<std::thread> th;
    for (int i = 1; i < n_threads; ++i) th.emplace_back(encode, i);
    encode(0);
    for (auto& t : th) t.join();
  }
  if (domain_err.load()) {
    set_error("a genotype value does not fit the bit-planes of its population");
    return SAI_E_DOMAIN;
  }
  std::vector<uint64_t> base(n_blocks + 1, 0);
  for (int64_t b = 0; b < n_blocks; ++b) base[b + 1] = base[b] + recs[b].n;
  const uint64_t total = base[n_blocks];
  if (total > out_cap) {
    set_error("zt stream needs %llu bytes, buffer has %llu", (unsigned long long)total, (unsigned long long)out_cap);
    return SAI_E_CAPACITY;
  }
  run_parallel(n_blocks, n_threads, [&](int64_t b0, int64_t b1) {
    for (int64_t b = b0; b < b1; ++b) {
      memcpy(out + base[b], recs[b].p, recs[b].n);
      for (int64_t T = b * block; T < std::min(n_tiles, (b + 1) * block); ++T) tile_off[T] += base[b];  // raw flag: bit 63 untouched
    }
  });
  tile_off[n_tiles] = total;
  return (int64_t)total;
}

int sai_zt_decode_host(const sai_layout* lay, const uint8_t* stream, const uint64_t* tile_off, int64_t n_sites,
                       uint8_t* packed) {
  if (int rc = validate_layout(lay)) return rc;
  SAI_REQUIRE(n_sites >= 0 && tile_off, "bad argument");
  const int64_t n_tiles = sai_num_tiles(n_sites);
  if (n_tiles == 0) return SAI_OK;
  SAI_REQUIRE(stream && packed, "NULL argument");
  const int P = lay->pairs_per_site;
  const size_t dense = (size_t)P * kTile * 8;
  for (int64_t T = 0; T < n_tiles; ++T) {
    const uint8_t* rec = stream + (tile_off[T] & ~kZtRaw);
    uint8_t* dst = packed + (size_t)T * dense;
    if (tile_off[T] & kZtRaw) {
      memcpy(dst, rec, dense);
      continue;
    }
    uint32_t n1;
    memcpy(&n1, rec, 4);
    const uint8_t* mp = rec + 4 + 4 * (size_t)P;
    const uint8_t* dp = mp + n1;
    for (int r = 0; r < P; ++r) {
      uint32_t w;
      memcpy(&w, rec + 4 + 4 * (size_t)r, 4);
      const uint64_t c = pad_constant(*lay, r);
      for (int s = 0; s < kTile; ++s) {
        uint64_t v = 0;
        if ((w >> s) & 1u) {
          const uint32_t m = *mp++;
          for (int k = 0; k < 8; ++k)
            if ((m >> k) & 1u) v |= (uint64_t)(*dp++) << (8 * k);
        }
        v ^= c;
        memcpy(dst + ((size_t)r * kTile + s) * 8, &v, 8);
      }
    }
  }
  return SAI_OK;
}

int sai_zt_decode(const sai_layout* lay, const void* d_stream, uint64_t stream_bytes, const uint64_t* d_tile_off,
                  int64_t tile0, int64_t n_tiles, void* d_packed, void* stream) {
  if (int rc = validate_layout(lay)) return rc;
  SAI_REQUIRE(tile0 >= 0 && n_tiles >= 0, "bad tile range");
  if (n_tiles == 0) return SAI_OK;
  SAI_REQUIRE(d_stream && d_tile_off && d_packed, "NULL device pointer");
  ZtDecodeParams P{};
  P.lay = *lay;
  P.stream = static_cast<const uint8_t*>(d_stream);
  P.stream_bytes = stream_bytes;
  P.tile_off = reinterpret_cast<const unsigned long long*>(d_tile_off);
  P.tile0 = tile0;
  P.n_tiles = n_tiles;
  P.packed = static_cast<unsigned long long*>(d_packed);
  const int64_t want = (n_tiles + kZtWarps - 1) / kZtWarps;
  const int64_t cap = (int64_t)sm_count() * 8;
  SAI_REQUIRE((reinterpret_cast<uintptr_t>(d_stream) & 7) == 0 && (stream_bytes & 7) == 0,
              "zt stream must be 8-byte aligned and a multiple of 8 bytes long");
  const size_t smem = sizeof(ZtTables) + sizeof(unsigned long long) * (size_t)lay->pairs_per_site;
  if (smem > 48 * 1024)
    SAI_CUDA_CHECK(cudaFuncSetAttribute(k_zt_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_zt_decode<<<(unsigned)(want < cap ? want : cap), kZtWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(P);
  SAI_CUDA_CHECK(cudaGetLastError());
  return SAI_OK;
}

}  // extern "C"
