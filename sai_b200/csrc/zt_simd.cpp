// Host encoder of one "zt" record (zero-suppressed tile, format in zt_codec.cu / include/sai_b200.h):
// plain C++ (g++) with a run-time selected AVX-512 path.  A dense tile is P rows of 32 pairs
// (8 bytes each); per row the encoder XORs the row's padding constant out and emits
//     nz[r]   which of the 32 pairs are non-zero            (vptestmq: 8 pairs per instruction)
//     mask[]  per non-zero pair, which of its bytes are     (vptestmb: the 64 byte flags of 8 pairs,
//             non-zero                                        compacted by nz with vpcompressb)
//     data[]  the non-zero bytes                             (vpcompressb of the 64 bytes)
// The vector path needs AVX-512 BW + VL + VBMI2 (Ice Lake and later); the portable path is the
// byte-at-a-time restatement.  Both write identical bytes (tests/test_cabi_host.py).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "zt_simd.h"

#include <algorithm>

#include "host_pack.h"

#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#define SAI_X86 1
#endif

namespace sai {

namespace {

constexpr int kTileSites = SAI_TILE_SITES;

inline uint32_t byte_mask(uint64_t v) {  // bit k set iff byte k of v is non-zero
  const uint64_t lo7 = 0x7f7f7f7f7f7f7f7full;
  const uint64_t m = (((v & lo7) + lo7) | v) & ~lo7;  // 0x80 in every non-zero byte
  return (uint32_t)(((m >> 7) * 0x0102040810204080ull) >> 56);
}

// Returns the unpadded record length; `rec` has room for zt_record_cap(P) bytes, `tmp` for the
// data bytes of a whole tile + 64.
size_t encode_portable(const uint64_t* tile, int P, const uint64_t* padc, uint8_t* rec, uint8_t* tmp) {
  uint8_t* mp = rec + 4 + 4 * (size_t)P;
  uint8_t* dp = tmp;
  uint32_t n1 = 0;
  for (int r = 0; r < P; ++r) {
    const uint64_t c = padc[r];
    uint32_t w = 0;
    for (int s = 0; s < kTileSites; ++s) {
      uint64_t v = tile[r * kTileSites + s] ^ c;
      if (!v) continue;
      w |= 1u << s;
      mp[n1++] = (uint8_t)byte_mask(v);
      for (int k = 0; k < 8; ++k, v >>= 8)
        if (v & 0xff) *dp++ = (uint8_t)v;
    }
    memcpy(rec + 4 + 4 * (size_t)r, &w, 4);
  }
  memcpy(rec, &n1, 4);
  const size_t n2 = dp - tmp;
  memcpy(mp + n1, tmp, n2);
  return 4 + 4 * (size_t)P + n1 + n2;
}

size_t size_portable(const uint64_t* tile, int P, const uint64_t* padc) {
  size_t n1 = 0, n2 = 0;
  for (int r = 0; r < P; ++r) {
    const uint64_t c = padc[r];
    for (int s = 0; s < kTileSites; ++s) {
      const uint64_t v = tile[r * kTileSites + s] ^ c;
      if (v) {
        ++n1;
        n2 += __builtin_popcount(byte_mask(v));
      }
    }
  }
  return 4 + 4 * (size_t)P + n1 + n2;
}

#ifdef SAI_X86
#pragma GCC push_options
#pragma GCC target("avx512f,avx512bw,avx512vl,avx512vbmi2,popcnt")
size_t encode_avx512(const uint64_t* tile, int P, const uint64_t* padc, uint8_t* rec, uint8_t* tmp) {
  uint8_t* mp = rec + 4 + 4 * (size_t)P;
  uint8_t* dp = tmp;
  uint32_t* nz = reinterpret_cast<uint32_t*>(rec + 4);
  for (int r = 0; r < P; ++r) {
    const __m512i c = _mm512_set1_epi64((long long)padc[r]);
    const uint64_t* row = tile + (size_t)r * kTileSites;
    const __m512i v0 = _mm512_xor_si512(_mm512_loadu_si512(row), c);
    const __m512i v1 = _mm512_xor_si512(_mm512_loadu_si512(row + 8), c);
    const __m512i v2 = _mm512_xor_si512(_mm512_loadu_si512(row + 16), c);
    const __m512i v3 = _mm512_xor_si512(_mm512_loadu_si512(row + 24), c);
    const uint32_t w = (uint32_t)_mm512_test_epi64_mask(v0, v0) | ((uint32_t)_mm512_test_epi64_mask(v1, v1) << 8) |
                       ((uint32_t)_mm512_test_epi64_mask(v2, v2) << 16) | ((uint32_t)_mm512_test_epi64_mask(v3, v3) << 24);
    nz[r] = w;
    if (w == 0u) continue;
    const uint64_t b0 = _mm512_test_epi8_mask(v0, v0), b1 = _mm512_test_epi8_mask(v1, v1);
    const uint64_t b2 = _mm512_test_epi8_mask(v2, v2), b3 = _mm512_test_epi8_mask(v3, v3);
    // the byte masks of the 32 pairs are the bytes of b0..b3; keep those of the non-zero pairs
    const __m256i masks = _mm256_set_epi64x((long long)b3, (long long)b2, (long long)b1, (long long)b0);
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(mp), _mm256_maskz_compress_epi8((__mmask32)w, masks));
    mp += __builtin_popcount(w);
    _mm512_storeu_si512(dp, _mm512_maskz_compress_epi8(b0, v0));
    dp += __builtin_popcountll(b0);
    _mm512_storeu_si512(dp, _mm512_maskz_compress_epi8(b1, v1));
    dp += __builtin_popcountll(b1);
    _mm512_storeu_si512(dp, _mm512_maskz_compress_epi8(b2, v2));
    dp += __builtin_popcountll(b2);
    _mm512_storeu_si512(dp, _mm512_maskz_compress_epi8(b3, v3));
    dp += __builtin_popcountll(b3);
  }
  const uint32_t n1 = (uint32_t)(mp - (rec + 4 + 4 * (size_t)P));
  memcpy(rec, &n1, 4);
  const size_t n2 = dp - tmp;
  memcpy(mp, tmp, n2);
  return 4 + 4 * (size_t)P + n1 + n2;
}

size_t size_avx512(const uint64_t* tile, int P, const uint64_t* padc) {
  size_t n1 = 0, n2 = 0;
  for (int r = 0; r < P; ++r) {
    const __m512i c = _mm512_set1_epi64((long long)padc[r]);
    const uint64_t* row = tile + (size_t)r * kTileSites;
    for (int i = 0; i < 4; ++i) {
      const __m512i v = _mm512_xor_si512(_mm512_loadu_si512(row + 8 * i), c);
      n1 += __builtin_popcount((uint32_t)_mm512_test_epi64_mask(v, v));
      n2 += __builtin_popcountll(_mm512_test_epi8_mask(v, v));
    }
  }
  return 4 + 4 * (size_t)P + n1 + n2;
}
#pragma GCC pop_options

bool cpu_has_vbmi2() {
  __builtin_cpu_init();
  return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
         __builtin_cpu_supports("avx512vbmi2") && __builtin_cpu_supports("popcnt");
}
#endif

bool use_vector(int isa) {
#ifdef SAI_X86
  static const bool has = cpu_has_vbmi2();
  return has && isa != 1;
#else
  (void)isa;
  return false;
#endif
}

}  // namespace

const char* zt_isa() { return use_vector(0) ? "avx512vbmi2" : "portable"; }

size_t zt_tile_size(const uint64_t* tile, int P, const uint64_t* padc, int isa) {
#ifdef SAI_X86
  if (use_vector(isa)) return size_avx512(tile, P, padc);
#endif
  return size_portable(tile, P, padc);
}

size_t zt_encode_tile(const uint64_t* tile, int P, const uint64_t* padc, uint8_t* rec, uint8_t* tmp, int isa) {
#ifdef SAI_X86
  if (use_vector(isa)) return encode_avx512(tile, P, padc, rec, tmp);
#endif
  return encode_portable(tile, P, padc, rec, tmp);
}

ZtBlockScratch::ZtBlockScratch(int P) {
  const size_t tile_bytes = (size_t)P * kTileSites * 8;
  const size_t stage_bytes = (64 + std::max(zt_record_cap(P), tile_bytes) + 63) & ~size_t(63);
  mem = new uint8_t[64 + tile_bytes + stage_bytes + zt_tmp_cap(P)];
  tilebuf = mem + ((64 - reinterpret_cast<uintptr_t>(mem) % 64) % 64);
  stage = tilebuf + tile_bytes;
  tmp = stage + stage_bytes;
}

ZtBlockScratch::~ZtBlockScratch() { delete[] mem; }

size_t zt_pack_block_i8(const sai_layout& lay, const int8_t* const* gt, const int64_t* row_stride, int64_t n_sites,
                        int64_t t0, int64_t t1, const uint64_t* padc, uint8_t* region, uint64_t base_off,
                        uint64_t* tile_off, ZtBlockScratch& sc, bool nontemporal, bool* bad) {
  const int P = lay.pairs_per_site;
  const size_t tile_bytes = (size_t)P * kTileSites * 8;
  size_t emitted = 0, carry = 0;  // bytes written to the region / waiting at the head of the stage
  for (int64_t T = t0; T < t1; ++T) {
    if (pack_tiles_i8_all(lay, gt, row_stride, n_sites, T, T + 1, T, sc.tilebuf, 0, false)) *bad = true;
    uint8_t* rec = sc.stage + carry;
    const size_t unpadded = zt_encode_tile(reinterpret_cast<const uint64_t*>(sc.tilebuf), P, padc, rec, sc.tmp, 0);
    size_t len = (unpadded + 7) & ~size_t(7);
    uint64_t flag = 0;
    if (len >= tile_bytes) {  // not smaller than the tile itself: raw
      memcpy(rec, sc.tilebuf, tile_bytes);
      len = tile_bytes;
      flag = 1ull << 63;
    } else {
      memset(rec + unpadded, 0, len - unpadded);
    }
    tile_off[T] = (base_off + emitted + carry) | flag;
    if (!nontemporal) {
      memcpy(region + emitted, rec, len);
      emitted += len;
      continue;
    }
    const size_t total = carry + len, lines = total / 64;
    stream_lines(region + emitted, sc.stage, lines);
    emitted += lines * 64;
    carry = total - lines * 64;
    if (carry) memmove(sc.stage, sc.stage + lines * 64, carry);
  }
  if (carry) {
    memset(sc.stage + carry, 0, 64 - carry);
    stream_lines(region + emitted, sc.stage, 1);
    emitted += 64;
  }
  if (nontemporal) stream_fence();
  return emitted;
}

}  // namespace sai
