// Fast path of the host VCF genotype parser (N2: ingest) for "regular" records: every sample field
// is exactly three characters `x|y` or `x/y` (x, y one digit or `.`), fields separated by single
// tabs -- what phased / unphased diploid GT-only panels (1000 Genomes style) look like.  Such a
// sample region is a sequence of 4-byte words [x, sep, y, TAB]; 16 of them are classified and
// converted per AVX-512 step (one word at a time in the portable path).  Anything else -- a
// multi-digit allele, another ploidy, extra FORMAT sub-fields, CRLF surprises -- makes the function
// return false and the caller falls back to the general field walker of vcf_parse.cu, so the
// fast path can only ever produce what the general parser produces (cross-checked in the tests).
//
// Plain C++ (compiled by g++) so that the vector path is selected at run time.
#include <stdint.h>
#include <string.h>

#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#define SAI_X86 1
#endif

namespace sai {

namespace {

// one field, scalar: w = little-endian [x, sep, y, end]; `end` is checked by the caller for the last field
inline bool field_alleles(uint32_t w, bool check_tab, int8_t& a0, int8_t& a1) {
  const unsigned c0 = w & 0xffu, sep = (w >> 8) & 0xffu, c1 = (w >> 16) & 0xffu, tab = w >> 24;
  const unsigned d0 = c0 - '0', d1 = c1 - '0';
  const bool ok0 = d0 < 10u || c0 == '.', ok1 = d1 < 10u || c1 == '.';
  if (!(ok0 && ok1 && (sep == '|' || sep == '/') && (!check_tab || tab == '\t'))) return false;
  a0 = d0 < 10u ? (int8_t)d0 : (int8_t)-1;
  a1 = d1 < 10u ? (int8_t)d1 : (int8_t)-1;
  return true;
}

bool regular_portable(const char* s, int64_t n, int8_t* a0, int8_t* a1) {
  for (int64_t i = 0; i + 1 < n; ++i) {
    uint32_t w;
    memcpy(&w, s + 4 * i, 4);
    if (!field_alleles(w, true, a0[i], a1[i])) return false;
  }
  return true;
}

inline int8_t flipped(int8_t a) { return (int8_t)(a > 0 ? a - 1 : 1 - a); }  // |a - 1| (utils.py:555)

bool regular_sum_portable(const char* s, int64_t n, bool flip, int8_t* sum2) {
  for (int64_t i = 0; i + 1 < n; ++i) {
    uint32_t w;
    memcpy(&w, s + 4 * i, 4);
    int8_t a0, a1;
    if (!field_alleles(w, true, a0, a1)) return false;
    sum2[i] = flip ? (int8_t)(flipped(a0) + flipped(a1)) : (int8_t)(a0 + a1);
  }
  return true;
}

#ifdef SAI_X86
#pragma GCC push_options
#pragma GCC target("avx512f,avx512bw")
bool regular_sum_avx512(const char* s, int64_t n, bool flip, int8_t* sum2) {
  const __m512i zero = _mm512_set1_epi8('0'), dot = _mm512_set1_epi8('.'), bar = _mm512_set1_epi8('|'),
                slash = _mm512_set1_epi8('/'), tab = _mm512_set1_epi8('\t'), ten = _mm512_set1_epi8(10),
                minus1 = _mm512_set1_epi8(-1), one = _mm512_set1_epi8(1);
  int64_t i = 0;
  for (; i + 16 < n; i += 16) {
    const __m512i v = _mm512_loadu_si512(s + 4 * i);
    const __m512i d = _mm512_sub_epi8(v, zero);
    const __mmask64 is_digit = _mm512_cmplt_epu8_mask(d, ten);
    const __mmask64 is_dot = _mm512_cmpeq_epi8_mask(v, dot);
    const __mmask64 is_sep = _mm512_cmpeq_epi8_mask(v, bar) | _mm512_cmpeq_epi8_mask(v, slash);
    const __mmask64 is_tab = _mm512_cmpeq_epi8_mask(v, tab);
    const __mmask64 allele_ok = is_digit | is_dot;
    if ((allele_ok & 0x5555555555555555ull) != 0x5555555555555555ull || (is_sep & 0x2222222222222222ull) != 0x2222222222222222ull ||
        (is_tab & 0x8888888888888888ull) != 0x8888888888888888ull)
      return false;
    __m512i val = _mm512_mask_mov_epi8(d, is_dot, minus1);  // bytes 0 and 2 of every field: allele, or -1 for "."
    if (flip) val = _mm512_abs_epi8(_mm512_sub_epi8(val, one));
    const __m512i sum = _mm512_add_epi8(val, _mm512_srli_epi32(val, 16));  // byte 0 of every field: a0 + a1
    _mm_storeu_si128(reinterpret_cast<__m128i*>(sum2 + i), _mm512_cvtepi32_epi8(sum));
  }
  return regular_sum_portable(s + 4 * i, n - i, flip, sum2 + i);
}
#pragma GCC pop_options
#endif

#ifdef SAI_X86
#pragma GCC push_options
#pragma GCC target("avx512f,avx512bw")
bool regular_avx512(const char* s, int64_t n, int8_t* a0, int8_t* a1) {
  const __m512i zero = _mm512_set1_epi8('0'), dot = _mm512_set1_epi8('.'), bar = _mm512_set1_epi8('|'),
                slash = _mm512_set1_epi8('/'), tab = _mm512_set1_epi8('\t'), ten = _mm512_set1_epi8(10),
                minus1 = _mm512_set1_epi8(-1);
  int64_t i = 0;
  // 16 complete fields (with their trailing tabs) per step; the last field of the line has no tab
  for (; i + 16 < n; i += 16) {
    const __m512i v = _mm512_loadu_si512(s + 4 * i);
    const __m512i d = _mm512_sub_epi8(v, zero);
    const __mmask64 is_digit = _mm512_cmplt_epu8_mask(d, ten);
    const __mmask64 is_dot = _mm512_cmpeq_epi8_mask(v, dot);
    const __mmask64 is_sep = _mm512_cmpeq_epi8_mask(v, bar) | _mm512_cmpeq_epi8_mask(v, slash);
    const __mmask64 is_tab = _mm512_cmpeq_epi8_mask(v, tab);
    const __mmask64 allele_ok = is_digit | is_dot;
    if ((allele_ok & 0x5555555555555555ull) != 0x5555555555555555ull || (is_sep & 0x2222222222222222ull) != 0x2222222222222222ull ||
        (is_tab & 0x8888888888888888ull) != 0x8888888888888888ull)
      return false;
    const __m512i val = _mm512_mask_mov_epi8(d, is_dot, minus1);  // digit value, or -1 for "."
    _mm_storeu_si128(reinterpret_cast<__m128i*>(a0 + i), _mm512_cvtepi32_epi8(val));
    _mm_storeu_si128(reinterpret_cast<__m128i*>(a1 + i), _mm512_cvtepi32_epi8(_mm512_srli_epi32(val, 16)));
  }
  return regular_portable(s + 4 * i, n - i, a0 + i, a1 + i);
}
#pragma GCC pop_options
#endif

}  // namespace

// Sample region [s, lend) of one record (first character after the FORMAT column up to the line
// end, CR already stripped).  If it is regular, writes the two alleles of every field (digit value
// or -1 for ".") to a0[i], a1[i], sets *n_fields and returns true; `cap` = capacity of a0 / a1.
bool vcf_regular_diploid(const char* s, const char* lend, int8_t* a0, int8_t* a1, int64_t cap, int64_t* n_fields) {
  const int64_t len = lend - s;
  if (len < 3 || ((len + 1) & 3) != 0) return false;
  const int64_t n = (len + 1) >> 2;
  if (n > cap) return false;
  bool ok;
#ifdef SAI_X86
  static const bool has512 = [] {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f");
  }();
  ok = has512 ? regular_avx512(s, n, a0, a1) : regular_portable(s, n, a0, a1);
#else
  ok = regular_portable(s, n, a0, a1);
#endif
  if (!ok) return false;
  // the last field: three characters, then the line end
  uint32_t w = 0;
  memcpy(&w, s + 4 * (n - 1), 3);
  if (!field_alleles(w, false, a0[n - 1], a1[n - 1])) return false;
  *n_fields = n;
  return true;
}

bool vcf_regular_diploid_sum(const char* s, const char* lend, bool flip, int8_t* sum2, int64_t cap, int64_t* n_fields) {
  const int64_t len = lend - s;
  if (len < 3 || ((len + 1) & 3) != 0) return false;
  const int64_t n = (len + 1) >> 2;
  if (n > cap) return false;
  bool ok;
#ifdef SAI_X86
  static const bool has512 = [] {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f");
  }();
  ok = has512 ? regular_sum_avx512(s, n, flip, sum2) : regular_sum_portable(s, n, flip, sum2);
#else
  ok = regular_sum_portable(s, n, flip, sum2);
#endif
  if (!ok) return false;
  uint32_t w = 0;
  memcpy(&w, s + 4 * (n - 1), 3);  // the last field: three characters, then the line end
  int8_t a0, a1;
  if (!field_alleles(w, false, a0, a1)) return false;
  sum2[n - 1] = flip ? (int8_t)(flipped(a0) + flipped(a1)) : (int8_t)(a0 + a1);
  *n_fields = n;
  return true;
}

}  // namespace sai
