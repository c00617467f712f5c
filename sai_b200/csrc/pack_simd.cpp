// Host packer core: int8 per-individual allele sums -> tiled bit-planes (include/sai_b200.h),
// the encoding that replaces the reference's int64 matrices (reshape_genotypes,
// sai/utils/utils.py:405-410).  Plain C++ (compiled by g++, not nvcc) so that the x86 vector
// paths can be selected at run time: AVX-512 with GFNI + VBMI (64 individuals per 8x8 bit-matrix
// transpose + one byte permute), AVX-512BW (64 per mask-test group), AVX2 (32), SSE2 (16) or a
// portable 64-bit multiply "movemask" (8).
//
// One call packs one population of a run of tiles.  A tile is 32 sites; per site the 32-individual
// groups of the row are turned into B plane words (bit i of plane b = bit b of individual i's
// code; a negative value = missing = all-ones code) and stored at the site's slot of each pair
// row, 256 bytes apart -- the whole tile (pairs_per_site * 256 B, 20 KB for 2504 diploids) stays
// in L1/L2 while its 32 input rows stream through.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "host_pack.h"

#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#define SAI_X86 1
#endif

namespace sai {

namespace {

constexpr int kTileSites = SAI_TILE_SITES;

// Row packers: words[g * B + b] for g < n_groups; returns true when a called value exceeds the
// planes (>= 2^B - 1).  `n` individuals; the last group is padded with the missing code.
typedef bool (*row_fn)(const int8_t* row, int n, int n_groups, int B, uint32_t* words);

bool row_portable(const int8_t* row, int n, int n_groups, int B, uint32_t* words) {
  const uint64_t ones = 0x0101010101010101ull;
  const int miss_code = (1 << B) - 1;
  bool bad = false;
  int8_t tail[32];
  for (int g = 0; g < n_groups; ++g) {
    const int i0 = g * 32, cnt = std::min(32, n - i0);
    const int8_t* v = row + i0;
    if (cnt < 32) {
      memset(tail, 0xff, sizeof(tail));
      memcpy(tail, row + i0, cnt);
      v = tail;
    }
    uint32_t* plane = words + (size_t)g * B;
    for (int b = 0; b < B; ++b) plane[b] = 0;
    for (int q = 0; q < 4; ++q) {
      uint64_t x;
      memcpy(&x, v + 8 * q, 8);
      const uint64_t negm = ((x >> 7) & ones) * 0xffull;  // 0xff in every negative byte
      const uint64_t val = x & ~negm;
      // a called value must be < miss_code: byte + (128 - miss_code) sets bit 7 otherwise
      // (8 planes hold every non-negative int8)
      if (miss_code < 128) bad |= (((val + ones * (uint64_t)(128 - miss_code)) | val) & (ones << 7)) != 0;
      const uint64_t code = val | (negm & (ones * (uint64_t)miss_code));
      for (int b = 0; b < B; ++b)
        plane[b] |= (uint32_t)((((code >> b) & ones) * 0x0102040810204080ull) >> 56) << (8 * q);
    }
  }
  return bad;
}

#ifdef SAI_X86
// SSE2 is part of x86-64: 16 individuals per instruction group, two per 32-individual group.
// (The B200 boxes' virtual CPUs expose AVX but neither AVX2 nor AVX-512: this is their path.)
bool row_sse2(const int8_t* row, int n, int n_groups, int B, uint32_t* words) {
  const __m128i lim = _mm_set1_epi8((char)std::min((1 << B) - 2, 127));  // largest called value
  __m128i bad = _mm_setzero_si128();
  auto half = [&](__m128i v, uint32_t (&p)[SAI_MAX_BITS]) {
    const uint32_t neg = (uint32_t)_mm_movemask_epi8(v);  // sign bits: missing calls
    bad = _mm_or_si128(bad, _mm_cmpgt_epi8(v, lim));
    p[0] = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v, 7)) | neg;
    p[1] = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v, 6)) | neg;
    if (B > 2) p[2] = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v, 5)) | neg;
    if (B > 3) p[3] = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v, 4)) | neg;
    if (B > 4) p[4] = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v, 3)) | neg;
    if (B > 5) p[5] = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v, 2)) | neg;
    if (B > 6) p[6] = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v, 1)) | neg;
    if (B > 7) p[7] = neg;  // bit 7 of a called value is never set: the plane is the missing mask
  };
  const int full = n / 32;
  for (int g = 0; g < full; ++g) {
    uint32_t lo[SAI_MAX_BITS], hi[SAI_MAX_BITS];
    half(_mm_loadu_si128((const __m128i*)(row + 32 * g)), lo);
    half(_mm_loadu_si128((const __m128i*)(row + 32 * g + 16)), hi);
    for (int b = 0; b < B; ++b) words[(size_t)g * B + b] = lo[b] | (hi[b] << 16);
  }
  if (full < n_groups) {
    alignas(16) int8_t tail[32];
    memset(tail, 0xff, sizeof(tail));
    memcpy(tail, row + 32 * full, n - 32 * full);
    uint32_t lo[SAI_MAX_BITS], hi[SAI_MAX_BITS];
    half(_mm_load_si128((const __m128i*)tail), lo);
    half(_mm_load_si128((const __m128i*)(tail + 16)), hi);
    for (int b = 0; b < B; ++b) words[(size_t)full * B + b] = lo[b] | (hi[b] << 16);
  }
  return _mm_movemask_epi8(bad) != 0;
}

// (pragma rather than the function attribute: the lambdas inside must carry the target too)
#pragma GCC push_options
#pragma GCC target("avx2")
bool row_avx2(const int8_t* row, int n, int n_groups, int B, uint32_t* words) {
  const __m256i lim = _mm256_set1_epi8((char)std::min((1 << B) - 2, 127));  // largest called value
  __m256i bad = _mm256_setzero_si256();
  const int full = n / 32;
  auto group = [&](__m256i v, uint32_t* plane) {
    const uint32_t neg = (uint32_t)_mm256_movemask_epi8(v);  // sign bits: missing calls
    bad = _mm256_or_si256(bad, _mm256_cmpgt_epi8(v, lim));
    // bit b of every byte moved to the byte's sign position, then gathered
    plane[0] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 7)) | neg;
    plane[1] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6)) | neg;
    if (B > 2) plane[2] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5)) | neg;
    if (B > 3) plane[3] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 4)) | neg;
    if (B > 4) plane[4] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 3)) | neg;
    if (B > 5) plane[5] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 2)) | neg;
    if (B > 6) plane[6] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 1)) | neg;
    if (B > 7) plane[7] = neg;  // bit 7 of a called value is never set: the plane is the missing mask
  };
  for (int g = 0; g < full; ++g) group(_mm256_loadu_si256((const __m256i*)(row + 32 * g)), words + (size_t)g * B);
  if (full < n_groups) {
    alignas(32) int8_t tail[32];
    memset(tail, 0xff, sizeof(tail));
    memcpy(tail, row + 32 * full, n - 32 * full);
    group(_mm256_load_si256((const __m256i*)tail), words + (size_t)full * B);
  }
  return !_mm256_testz_si256(bad, bad);
}

#pragma GCC pop_options

#pragma GCC push_options
#pragma GCC target("avx512f,avx512bw")
bool row_avx512(const int8_t* row, int n, int n_groups, int B, uint32_t* words) {
  const __m512i lim = _mm512_set1_epi8((char)std::min((1 << B) - 2, 127));
  __mmask64 bad = 0;
  const __m512i bit0 = _mm512_set1_epi8(1), bit1 = _mm512_set1_epi8(2), bit2 = _mm512_set1_epi8(4), bit3 = _mm512_set1_epi8(8);
  auto pair = [&](__m512i v, int g, int groups_here) {  // 64 individuals = groups g and g + 1
    const uint64_t neg = _mm512_movepi8_mask(v);
    bad |= _mm512_cmpgt_epi8_mask(v, lim);
    uint64_t p[SAI_MAX_BITS];
    p[0] = _mm512_test_epi8_mask(v, bit0) | neg;
    p[1] = _mm512_test_epi8_mask(v, bit1) | neg;
    if (B > 2) p[2] = _mm512_test_epi8_mask(v, bit2) | neg;
    if (B > 3) p[3] = _mm512_test_epi8_mask(v, bit3) | neg;
    for (int b = 4; b < B; ++b) p[b] = _mm512_test_epi8_mask(v, _mm512_set1_epi8((char)(1 << b))) | neg;
    for (int h = 0; h < groups_here; ++h)
      for (int b = 0; b < B; ++b) words[(size_t)(g + h) * B + b] = (uint32_t)(p[b] >> (32 * h));
  };
  const int full = n / 64;
  for (int k = 0; k < full; ++k) pair(_mm512_loadu_si512(row + 64 * k), 2 * k, 2);
  const int rest = n - 64 * full;
  if (rest > 0) {  // masked load; lanes beyond the row read as -1 (missing)
    const __mmask64 m = rest >= 64 ? ~0ull : ((1ull << rest) - 1ull);
    const __m512i v = _mm512_mask_loadu_epi8(_mm512_set1_epi8(-1), m, row + 64 * full);
    pair(v, 2 * full, n_groups - 2 * full);
  }
  return bad != 0;
}
#pragma GCC pop_options

// GFNI + VBMI (Ice Lake and later): no mask registers on the data path.  Per 64 individuals:
// the code bytes (missing -> 0xff) are byte-reversed inside every 8-byte group, vgf2p8affineqb with
// the DATA as the bit matrix and a one-hot selector per output byte transposes each group (output
// byte b = bit b of its 8 individuals, individual i at bit i), and one vpermb gathers byte b of four
// neighbouring groups into the plane word -- for the two 32-individual groups at once, already in
// the stored order [group][plane], so the row costs 8 vector instructions per 64 input bytes
// whatever the plane count.
#pragma GCC push_options
#pragma GCC target("avx512f,avx512bw,avx512vbmi,gfni")
struct GatherTables {
  alignas(64) uint8_t idx[SAI_MAX_BITS + 1][64];
  GatherTables() {
    memset(idx, 0, sizeof(idx));
    for (int B = 1; B <= SAI_MAX_BITS; ++B)
      for (int h = 0; h < 2; ++h)
        for (int b = 0; b < B; ++b)
          for (int k = 0; k < 4; ++k) idx[B][(h * B + b) * 4 + k] = (uint8_t)((4 * h + k) * 8 + b);
  }
};

bool row_avx512gfni(const int8_t* row, int n, int n_groups, int B, uint32_t* words) {
  static const GatherTables tables;
  const __m512i ones = _mm512_set1_epi8(-1);
  const __m512i one_hot = _mm512_set1_epi64((long long)0x8040201008040201ull);  // output byte b selects bit b
  const __m512i rev8 = _mm512_broadcast_i32x4(_mm_set_epi8(8, 9, 10, 11, 12, 13, 14, 15, 0, 1, 2, 3, 4, 5, 6, 7));
  const __m512i gather = _mm512_load_si512(tables.idx[B]);
  __m512i vmax = _mm512_set1_epi8(-128);
  auto planes = [&](__m512i v) {
    vmax = _mm512_max_epi8(vmax, v);
    const __m512i code = _mm512_mask_mov_epi8(v, _mm512_movepi8_mask(v), ones);
    const __m512i t = _mm512_gf2p8affine_epi64_epi8(one_hot, _mm512_shuffle_epi8(code, rev8), 0);
    return _mm512_permutexvar_epi8(gather, t);
  };
  const int full = n / 64;
  if (B == 2) {
    for (int k = 0; k < full; ++k)
      _mm_storeu_si128((__m128i*)(words + 4 * k), _mm512_castsi512_si128(planes(_mm512_loadu_si512(row + 64 * k))));
  } else if (B == 4) {
    for (int k = 0; k < full; ++k)
      _mm256_storeu_si256((__m256i*)(words + 8 * k), _mm512_castsi512_si256(planes(_mm512_loadu_si512(row + 64 * k))));
  } else {
    const __mmask64 m = B == 8 ? ~0ull : ((1ull << (8 * B)) - 1ull);
    for (int k = 0; k < full; ++k)
      _mm512_mask_storeu_epi8(words + (size_t)2 * B * k, m, planes(_mm512_loadu_si512(row + 64 * k)));
  }
  const int rest = n - 64 * full;
  if (rest > 0) {  // masked load; lanes beyond the row read as -1 (missing)
    const __m512i v = _mm512_mask_loadu_epi8(ones, (1ull << rest) - 1ull, row + 64 * full);
    const int out_bytes = (n_groups - 2 * full) * B * 4;
    _mm512_mask_storeu_epi8(words + (size_t)2 * B * full, out_bytes >= 64 ? ~0ull : ((1ull << out_bytes) - 1ull), planes(v));
  }
  const __m512i lim = _mm512_set1_epi8((char)std::min((1 << B) - 2, 127));
  return _mm512_cmpgt_epi8_mask(vmax, lim) != 0;
}
#pragma GCC pop_options

// The words of 8 consecutive sites (rows of `wbuf`, `wstride` words apart) -> one 64-byte line per
// pair row at `dst0 + p * 256`.  AVX-512: 8 x 8 transpose of 64-bit pairs in registers, one
// 64-byte (non-temporal) store per line.
typedef void (*flush_fn)(const uint32_t* wbuf, int wstride, int pps, uint8_t* dst0, bool nt);

#pragma GCC push_options
#pragma GCC target("avx512f,avx512bw")
void flush8_avx512(const uint32_t* wbuf, int wstride, int pps, uint8_t* dst0, bool nt) {
  int p = 0;
  for (; p + 8 <= pps; p += 8) {
    __m512i r[8], t[8], u[8];
    for (int j = 0; j < 8; ++j) r[j] = _mm512_loadu_si512(wbuf + (size_t)j * wstride + 2 * p);
    for (int j = 0; j < 8; j += 2) {
      t[j] = _mm512_unpacklo_epi64(r[j], r[j + 1]);      // pairs 0,2,4,6 of sites j, j+1
      t[j + 1] = _mm512_unpackhi_epi64(r[j], r[j + 1]);  // pairs 1,3,5,7
    }
    for (int odd = 0; odd < 2; ++odd) {
      u[4 * odd + 0] = _mm512_shuffle_i64x2(t[odd], t[2 + odd], 0x88);      // lanes 0,2 of sites 0-1 | 2-3
      u[4 * odd + 1] = _mm512_shuffle_i64x2(t[odd], t[2 + odd], 0xdd);      // lanes 1,3
      u[4 * odd + 2] = _mm512_shuffle_i64x2(t[4 + odd], t[6 + odd], 0x88);  // sites 4-5 | 6-7
      u[4 * odd + 3] = _mm512_shuffle_i64x2(t[4 + odd], t[6 + odd], 0xdd);
    }
    __m512i out[8];
    for (int odd = 0; odd < 2; ++odd) {
      out[0 + odd] = _mm512_shuffle_i64x2(u[4 * odd + 0], u[4 * odd + 2], 0x88);  // pair 0 / 1 of all 8 sites
      out[4 + odd] = _mm512_shuffle_i64x2(u[4 * odd + 0], u[4 * odd + 2], 0xdd);  // pair 4 / 5
      out[2 + odd] = _mm512_shuffle_i64x2(u[4 * odd + 1], u[4 * odd + 3], 0x88);  // pair 2 / 3
      out[6 + odd] = _mm512_shuffle_i64x2(u[4 * odd + 1], u[4 * odd + 3], 0xdd);  // pair 6 / 7
    }
    uint8_t* dst = dst0 + (size_t)p * kTileSites * 8;
    if (nt) {
      for (int c = 0; c < 8; ++c) _mm512_stream_si512((__m512i*)(dst + (size_t)c * kTileSites * 8), out[c]);
    } else {
      for (int c = 0; c < 8; ++c) _mm512_storeu_si512(dst + (size_t)c * kTileSites * 8, out[c]);
    }
  }
  for (; p < pps; ++p) {
    uint64_t line[8];
    for (int j = 0; j < 8; ++j) memcpy(&line[j], wbuf + (size_t)j * wstride + 2 * p, 8);
    uint8_t* dst = dst0 + (size_t)p * kTileSites * 8;
    if (nt) {
      for (int j = 0; j < 8; ++j) _mm_stream_si64(reinterpret_cast<long long*>(dst) + j, (long long)line[j]);
    } else {
      memcpy(dst, line, sizeof(line));
    }
  }
}
#pragma GCC pop_options
#endif

namespace {
void flush8_scalar(const uint32_t* wbuf, int wstride, int pps, uint8_t* dst0, bool nt) {
  for (int p = 0; p < pps; ++p) {
    uint64_t line[8];
    for (int j = 0; j < 8; ++j) memcpy(&line[j], wbuf + (size_t)j * wstride + 2 * p, 8);
    uint8_t* dst = dst0 + (size_t)p * kTileSites * 8;
#ifdef SAI_X86
    if (nt) {
      for (int j = 0; j < 8; ++j) _mm_stream_si64(reinterpret_cast<long long*>(dst) + j, (long long)line[j]);
      continue;
    }
#endif
    memcpy(dst, line, sizeof(line));
  }
}
}  // namespace

#ifdef SAI_X86
bool cpu_has_avx512bw() { return __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f"); }
bool cpu_has_gfni_vbmi() { return cpu_has_avx512bw() && __builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("gfni"); }
#endif

row_fn pick_row_fn() {
#ifdef SAI_X86
  __builtin_cpu_init();
  if (cpu_has_gfni_vbmi()) return row_avx512gfni;
  if (cpu_has_avx512bw()) return row_avx512;
  if (__builtin_cpu_supports("avx2")) return row_avx2;
  return row_sse2;
#endif
  return row_portable;
}

}  // namespace

const char* pack_isa() {
#ifdef SAI_X86
  __builtin_cpu_init();
  if (cpu_has_gfni_vbmi()) return "avx512gfni";
  if (cpu_has_avx512bw()) return "avx512bw";
  if (__builtin_cpu_supports("avx2")) return "avx2";
  return "sse2";
#endif
  return "portable";
}

namespace {
row_fn row_fn_for(int isa) {
  static const row_fn best = pick_row_fn();
  row_fn fn = best;
#ifdef SAI_EXPERIMENTS
  static const int forced = [] { const char* e = getenv("SAI_PACK_ISA"); return e ? atoi(e) : 0; }();  // A/B knob
  if (isa == 0) isa = forced;
#endif
  if (isa == 1) fn = row_portable;
#ifdef SAI_X86
  if (isa == 2) fn = row_sse2;
  if (isa == 3 && __builtin_cpu_supports("avx2")) fn = row_avx2;
  if (isa == 4 && cpu_has_avx512bw()) fn = row_avx512;
  if (isa == 5 && cpu_has_gfni_vbmi()) fn = row_avx512gfni;
#endif
  return fn;
}

// isa 1 (portable) keeps the scalar line writer so that the tests compare both
flush_fn flush_fn_for(int isa) {
#ifdef SAI_X86
  static const bool wide = (__builtin_cpu_init(), cpu_has_avx512bw());
#ifdef SAI_EXPERIMENTS
  static const bool off = [] { const char* e = getenv("SAI_PACK_FLUSH"); return e && e[0] == '0'; }();  // A/B knob
  if (off) return flush8_scalar;
#endif
  if (wide && (isa == 0 || isa >= 4)) return flush8_avx512;
#endif
  return flush8_scalar;
}

inline void prefetch_row(const int8_t* row, int n) {
#ifdef SAI_X86
#ifdef SAI_EXPERIMENTS
  static const int hint = [] { const char* e = getenv("SAI_PACK_HINT"); return e ? atoi(e) : 0; }();  // A/B knob
  if (hint == 1) {
    for (int i = 0; i < n; i += 64) _mm_prefetch(reinterpret_cast<const char*>(row + i), _MM_HINT_T1);
    return;
  }
  if (hint == 2) {
    for (int i = 0; i < n; i += 64) _mm_prefetch(reinterpret_cast<const char*>(row + i), _MM_HINT_T2);
    return;
  }
  if (hint == 3) {
    for (int i = 0; i < n; i += 64) _mm_prefetch(reinterpret_cast<const char*>(row + i), _MM_HINT_NTA);
    return;
  }
  if (hint == 4) {  // first line of every 4 KB page only: leave the rest to the hardware streamer
    const uintptr_t a = reinterpret_cast<uintptr_t>(row), b = a + n;
    _mm_prefetch(reinterpret_cast<const char*>(row), _MM_HINT_T0);
    for (uintptr_t pg = (a | 4095) + 1; pg < b; pg += 4096) _mm_prefetch(reinterpret_cast<const char*>(pg), _MM_HINT_T0);
    return;
  }
#endif
  for (int i = 0; i < n; i += 64) _mm_prefetch(reinterpret_cast<const char*>(row + i), _MM_HINT_T0);
#else
  (void)row, (void)n;
#endif
}

// the words of one site's row -> its slot (site s of the tile) of the population's pair rows
inline void scatter_site(const uint32_t* words, int n_pairs, uint8_t* tile_pop, int s) {
  uint64_t* dst = reinterpret_cast<uint64_t*>(tile_pop) + s;
  for (int p = 0; p < n_pairs; ++p) {
    uint64_t pair;
    memcpy(&pair, words + 2 * p, 8);
    dst[(size_t)p * kTileSites] = pair;
  }
}
}  // namespace

// All populations of tiles [t0, t1), site by site: when the populations are column blocks of one
// row-major matrix (what a VCF parse leaves behind) the whole matrix is read as ONE sequential
// stream, which is what the hardware prefetcher wants; the rows a few sites ahead are prefetched
// explicitly as well.  Output: the words of 8 consecutive sites are collected in a small buffer
// and every pair row then receives its 8 x 8 bytes as ONE full cache line of non-temporal stores
// (no read-for-ownership of the destination, no cache pollution: the tiles are consumed by the
// DMA engine or a later pass, never by this core).  Returns true when a value does not fit its
// population's bit-planes.
bool pack_tiles_i8_all(const sai_layout& lay, const int8_t* const* gt, const int64_t* row_stride, int64_t n_sites,
                       int64_t t0, int64_t t1, int64_t tile_base, uint8_t* packed_base, int isa, bool allow_nt) {
  const row_fn fn = row_fn_for(isa);
  const flush_fn flush = flush_fn_for(isa);
  const int pps = lay.pairs_per_site;
  const size_t tile_bytes = (size_t)pps * kTileSites * 8;
  const int col_words = 2 * pps;  // words of one site's whole column (zero pad words included)
  constexpr int kBatch = 8;       // sites per output cache line
  // +16: the vector row packers and the line writer move whole 64-byte vectors
  uint32_t stack_words[kBatch * 512 + 16];
  const int wstride = col_words + 2;
  uint32_t* wbuf = wstride <= 512 ? stack_words : new uint32_t[(size_t)kBatch * wstride + 16];
#ifdef SAI_EXPERIMENTS
  const char* ahead_env = getenv("SAI_PACK_AHEAD");  // tools/ A/B knob (read per call: one process can sweep it)
  const int kAhead = ahead_env ? atoi(ahead_env) : 4;
#else
  constexpr int kAhead = 4;  // sites
#endif
  // SAI_PACK_NT=0 (read once) switches the non-temporal stores off: an A/B knob for tools/pack_bench.py
  static const bool nt_enabled = [] {
    const char* e = getenv("SAI_PACK_NT");
    return !(e && e[0] == '0');
  }();
  const bool aligned = allow_nt && nt_enabled && (reinterpret_cast<uintptr_t>(packed_base) & 63) == 0;
  bool bad = false;
  for (int64_t T = t0; T < t1; ++T) {
    uint8_t* tile = packed_base + (size_t)(T - tile_base) * tile_bytes;
    for (int s8 = 0; s8 < kTileSites; s8 += kBatch) {
      for (int j = 0; j < kBatch; ++j) {
        const int64_t site = T * kTileSites + s8 + j;
        uint32_t* col = wbuf + (size_t)j * wstride;
        if (site + kAhead < n_sites)
          for (int p = 0; p < lay.n_pops; ++p) prefetch_row(gt[p] + (site + kAhead) * row_stride[p], lay.pop[p].n_samples);
        for (int p = 0; p < lay.n_pops; ++p) {
          const sai_pop_layout& L = lay.pop[p];
          const int n_words = L.n_groups * L.bits;
          uint32_t* words = col + 2 * L.pair_off;
          if (site < n_sites) {
            bad |= fn(gt[p] + site * row_stride[p], L.n_samples, L.n_groups, L.bits, words);
          } else {
            for (int w = 0; w < n_words; ++w) words[w] = 0xffffffffu;  // padding site: all missing
          }
          if (n_words & 1) words[n_words] = 0u;  // the population's pad word
        }
      }
      flush(wbuf, wstride, pps, tile + (size_t)s8 * 8, aligned);
    }
  }
#ifdef SAI_X86
  _mm_sfence();  // the non-temporal stores are visible before the caller publishes the tiles
#endif
  if (wbuf != stack_words) delete[] wbuf;
  return bad;
}

// n_lines 64-byte lines, src -> dst (both 64-byte aligned), written around the caches.
void stream_lines(uint8_t* dst, const uint8_t* src, size_t n_lines) {
#ifdef SAI_X86
  for (size_t i = 0; i < n_lines * 4; ++i)
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + i, _mm_load_si128(reinterpret_cast<const __m128i*>(src) + i));
#else
  memcpy(dst, src, n_lines * 64);
#endif
}

void stream_fence() {
#ifdef SAI_X86
  _mm_sfence();
#endif
}

// One population of tiles [t0, t1) (sai_pack_i8).
bool pack_tiles_i8(const sai_layout& lay, int pop, const int8_t* gt, int64_t n_sites, int64_t row_stride,
                   int64_t t0, int64_t t1, int64_t tile_base, uint8_t* packed_base, int isa) {
  const row_fn fn = row_fn_for(isa);
  const sai_pop_layout& L = lay.pop[pop];
  const int B = L.bits;
  const int n_words = L.n_groups * B;
  const size_t tile_bytes = (size_t)lay.pairs_per_site * kTileSites * 8;
  // words of one row, padded to whole pairs (+1 zero word when odd)
  uint32_t stack_words[1024];
  uint32_t* words = n_words + 1 <= 1024 ? stack_words : new uint32_t[n_words + 1];
  constexpr int kAhead = 3;
  bool bad = false;
  for (int64_t T = t0; T < t1; ++T) {
    uint8_t* tile = packed_base + (size_t)(T - tile_base) * tile_bytes + (size_t)L.pair_off * kTileSites * 8;
    for (int s = 0; s < kTileSites; ++s) {
      const int64_t site = T * kTileSites + s;
      if (site + kAhead < n_sites) prefetch_row(gt + (site + kAhead) * row_stride, L.n_samples);
      if (site < n_sites) {
        bad |= fn(gt + site * row_stride, L.n_samples, L.n_groups, B, words);
      } else {
        for (int w = 0; w < n_words; ++w) words[w] = 0xffffffffu;  // padding site: all missing
      }
      words[n_words] = 0u;
      scatter_site(words, L.n_pairs, tile, s);
    }
  }
  if (words != stack_words) delete[] words;
  return bad;
}

}  // namespace sai
