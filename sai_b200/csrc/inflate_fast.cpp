// Raw-deflate decoder and CRC-32 for the BGZF reader (bgzf.cu; N2: ingest).  Plain C++ (g++).
//
// A bgzip block is at most 64 KB of text behind a raw deflate stream (RFC 1951) whose inflated
// size and CRC-32 are known in advance (the gzip trailer).  That allows a decoder that is much
// simpler and faster than a general streaming inflate: whole input and whole output in memory,
// a 64-bit bit buffer refilled eight bytes at a time, two-level decode tables (11 bits for
// literal/length codes, 8 for distances; second level indexed by the remaining bits up to the
// 15-bit maximum) and word-wise match copies (VCF genotype columns are mostly matches of
// distance 4 = one "0|0<TAB>" field, written as a repeated 8-byte pattern).  It never reads
// outside [in, in + in_len) nor writes outside [out, out + out_len), returns false on anything
// unexpected -- the caller (bgzf.cu) then repeats the block with zlib, so a decoder bug could cost
// time but not correctness; every block is also checked against its CRC-32.
//
// CRC-32 (the gzip / zlib polynomial) uses carry-less multiplication (PCLMULQDQ: four 128-bit
// lanes folded per 64 bytes, then Barrett reduction; constants of Gopal et al., "Fast CRC
// computation for generic polynomials using PCLMULQDQ") with a slicing-by-8 table fallback.
// Both are verified against zlib in the CPU tests.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#define SAI_X86 1
#endif

namespace sai {

namespace {

constexpr int kLitBits = 11, kDistBits = 8, kMaxCodeLen = 15;
constexpr int kLitSub = kMaxCodeLen - kLitBits, kDistSub = kMaxCodeLen - kDistBits;
constexpr int kNumLit = 288, kNumDist = 32;

// decode-table entry: bits 0-3 code length, 4-6 kind, 8-12 number of extra bits, 16-31 value
enum Kind : uint32_t { kInvalid = 0, kLiteral = 1, kLength = 2, kEndOfBlock = 3, kSubtable = 4, kDistance = 5 };
inline uint32_t entry(uint32_t len, Kind kind, uint32_t extra, uint32_t value) {
  return len | ((uint32_t)kind << 4) | (extra << 8) | (value << 16);
}
inline uint32_t e_len(uint32_t e) { return e & 15u; }
inline uint32_t e_kind(uint32_t e) { return (e >> 4) & 7u; }
inline uint32_t e_extra(uint32_t e) { return (e >> 8) & 31u; }
inline uint32_t e_value(uint32_t e) { return e >> 16; }

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kCodeLenOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct ByteReverse {
  uint8_t r[256];
  ByteReverse() {
    for (int i = 0; i < 256; ++i) {
      int v = 0;
      for (int b = 0; b < 8; ++b) v |= ((i >> b) & 1) << (7 - b);
      r[i] = (uint8_t)v;
    }
  }
};
inline uint32_t reverse_bits(uint32_t code, int len) {  // len <= 15
  static const ByteReverse R;
  return (((uint32_t)R.r[code & 0xffu] << 8) | R.r[(code >> 8) & 0xffu]) >> (16 - len);
}

struct Tables {
  uint32_t lit[(1 << kLitBits) + kNumLit * (1 << kLitSub)];
  uint32_t dist[(1 << kDistBits) + kNumDist * (1 << kDistSub)];
};

// Canonical Huffman code (lens[0..n), 0 = unused) -> two-level decode table.  `symbol_entry`
// gives the entry of a symbol with its code length.  Returns false for an over-subscribed code;
// incomplete codes are allowed (their holes decode as kInvalid), as zlib allows the single-code case.
template <typename F>
bool build_table(const uint8_t* lens, int n, int primary_bits, int sub_bits, uint32_t* table, size_t table_cap, F symbol_entry) {
  int count[kMaxCodeLen + 1] = {0};
  for (int i = 0; i < n; ++i) ++count[lens[i]];
  count[0] = 0;
  uint32_t next_code[kMaxCodeLen + 2];
  uint32_t code = 0;
  int64_t left = 1;
  for (int len = 1; len <= kMaxCodeLen; ++len) {
    left = (left << 1) - count[len];
    if (left < 0) return false;  // over-subscribed
    code = (code + (uint32_t)count[len - 1]) << 1;
    next_code[len] = code;
  }
  const size_t primary = (size_t)1 << primary_bits;
  // A complete code writes every primary slot, so the table (reused from block to block) is not
  // cleared as a whole -- only the slots that will hold subtable links, because a link is
  // recognised by reading its slot: left-aligned, the codes of up to primary_bits bits occupy the
  // prefixes [0, short_slots) and the longer codes share the prefixes above.
  if (left != 0) {
    for (size_t i = 0; i < primary; ++i) table[i] = 0;
  } else {
    size_t short_slots = 0;
    for (int len = 1; len <= primary_bits; ++len) short_slots += (size_t)count[len] << (primary_bits - len);
    for (size_t v = short_slots; v < primary; ++v) table[reverse_bits((uint32_t)v, primary_bits)] = 0;
  }
  size_t used = primary;
  for (int sym = 0; sym < n; ++sym) {
    const int len = lens[sym];
    if (len == 0) continue;
    const uint32_t rev = reverse_bits(next_code[len]++, len);
    const uint32_t e = symbol_entry(sym, (uint32_t)len);
    if (len <= primary_bits) {
      for (size_t i = rev; i < primary; i += (size_t)1 << len) table[i] = e;
      continue;
    }
    const uint32_t lo = rev & (uint32_t)(primary - 1);
    if (e_kind(table[lo]) != kSubtable) {
      if (table[lo] != 0 || used + ((size_t)1 << sub_bits) > table_cap) return false;
      for (size_t i = 0; i < ((size_t)1 << sub_bits); ++i) table[used + i] = 0;
      table[lo] = entry(0, kSubtable, 0, (uint32_t)used);
      used += (size_t)1 << sub_bits;
    }
    uint32_t* sub = table + e_value(table[lo]);
    for (size_t i = rev >> primary_bits; i < ((size_t)1 << sub_bits); i += (size_t)1 << (len - primary_bits)) sub[i] = e;
  }
  return true;
}

inline uint32_t lit_entry(int sym, uint32_t len) {
  if (sym < 256) return entry(len, kLiteral, 0, (uint32_t)sym);
  if (sym == 256) return entry(len, kEndOfBlock, 0, 0);
  if (sym <= 285) return entry(len, kLength, kLenExtra[sym - 257], kLenBase[sym - 257]);
  return entry(len, kInvalid, 0, 0);  // 286, 287: never valid in a stream
}
inline uint32_t dist_entry(int sym, uint32_t len) {
  if (sym < 30) return entry(len, kDistance, kDistExtra[sym], kDistBase[sym]);
  return entry(len, kInvalid, 0, 0);
}

struct BitReader {
  const uint8_t* in;
  const uint8_t* end;
  uint64_t buf = 0;
  int bits = 0;       // valid bits in buf
  int overrun = 0;    // zero bytes shifted in past the end of the input
  BitReader(const uint8_t* p, size_t n) : in(p), end(p + n) {}
  inline void refill() {  // at least 56 valid bits afterwards
    if (end - in >= 8) {
      uint64_t w;
      memcpy(&w, in, 8);
      buf |= w << bits;
      const int take = (63 - bits) >> 3;
      in += take;
      bits += take * 8;
    } else {
      while (bits <= 55) {
        if (in < end) {
          buf |= (uint64_t)*in++ << bits;
        } else {
          ++overrun;
        }
        bits += 8;
      }
    }
  }
  inline uint32_t peek(int n) const { return (uint32_t)(buf & (((uint64_t)1 << n) - 1)); }
  inline void drop(int n) {
    buf >>= n;
    bits -= n;
  }
  inline uint32_t take(int n) {
    const uint32_t v = peek(n);
    drop(n);
    return v;
  }
  // true when more bits were consumed than the input holds
  inline bool past_end() const { return overrun * 8 > bits; }
};

bool read_dynamic_tables(BitReader& br, Tables& T) {
  br.refill();
  const int hlit = (int)br.take(5) + 257, hdist = (int)br.take(5) + 1, hclen = (int)br.take(4) + 4;
  if (hlit > 286 || hdist > 30) return false;
  uint8_t cl_lens[19] = {0};
  for (int i = 0; i < hclen; ++i) {
    if (br.bits < 3) br.refill();
    cl_lens[kCodeLenOrder[i]] = (uint8_t)br.take(3);
  }
  uint32_t cl_table[(1 << 7) + 1];
  if (!build_table(cl_lens, 19, 7, 0, cl_table, 1 << 7, [](int sym, uint32_t len) { return entry(len, kLiteral, 0, (uint32_t)sym); }))
    return false;
  uint8_t lens[kNumLit + kNumDist] = {0};
  int i = 0;
  while (i < hlit + hdist) {
    if (br.bits < 14) br.refill();  // a code of <= 7 bits and <= 7 extra bits
    const uint32_t e = cl_table[br.peek(7)];
    if (e_kind(e) != kLiteral) return false;
    br.drop((int)e_len(e));
    const int sym = (int)e_value(e);
    if (sym < 16) {
      lens[i++] = (uint8_t)sym;
      continue;
    }
    int rep;
    uint8_t val = 0;
    if (sym == 16) {
      if (i == 0) return false;
      val = lens[i - 1];
      rep = 3 + (int)br.take(2);
    } else if (sym == 17) {
      rep = 3 + (int)br.take(3);
    } else {
      rep = 11 + (int)br.take(7);
    }
    if (i + rep > hlit + hdist) return false;
    while (rep--) lens[i++] = val;
  }
  if (br.past_end() || lens[256] == 0) return false;
  uint8_t dlens[kNumDist] = {0};
  memcpy(dlens, lens + hlit, (size_t)hdist);
  uint8_t llens[kNumLit] = {0};
  memcpy(llens, lens, (size_t)hlit);
  return build_table(llens, kNumLit, kLitBits, kLitSub, T.lit, sizeof(T.lit) / 4, lit_entry) &&
         build_table(dlens, kNumDist, kDistBits, kDistSub, T.dist, sizeof(T.dist) / 4, dist_entry);
}

bool build_fixed_tables(Tables& T) {
  uint8_t llens[kNumLit], dlens[kNumDist];
  for (int i = 0; i < 144; ++i) llens[i] = 8;
  for (int i = 144; i < 256; ++i) llens[i] = 9;
  for (int i = 256; i < 280; ++i) llens[i] = 7;
  for (int i = 280; i < 288; ++i) llens[i] = 8;
  for (int i = 0; i < kNumDist; ++i) dlens[i] = 5;
  return build_table(llens, kNumLit, kLitBits, kLitSub, T.lit, sizeof(T.lit) / 4, lit_entry) &&
         build_table(dlens, kNumDist, kDistBits, kDistSub, T.dist, sizeof(T.dist) / 4, dist_entry);
}

// One Huffman-coded block.  false = invalid data or output overflow.  The table entry of the NEXT
// symbol is loaded before a match is copied, so that the load overlaps the copy.
bool inflate_block(BitReader& br, const Tables& T, uint8_t* const out_begin, uint8_t*& out_ref, uint8_t* const out_end) {
  uint8_t* out = out_ref;
  br.refill();
  uint32_t e = T.lit[br.peek(kLitBits)];
  for (;;) {
    // here: e = primary entry at the current position, >= 15 valid bits in the buffer
    if (e_kind(e) == kSubtable) e = T.lit[e_value(e) + ((br.buf >> kLitBits) & ((1u << kLitSub) - 1))];
    br.drop((int)e_len(e));
    const uint32_t kind = e_kind(e);
    if (kind == kLiteral) {
      if (out >= out_end) return false;
      *out++ = (uint8_t)e_value(e);
      if (br.bits < 32) br.refill();
      e = T.lit[br.peek(kLitBits)];
      continue;
    }
    if (kind == kLength) {
      if (br.bits < 5 + kMaxCodeLen + 13) br.refill();  // length extra bits, distance code, distance extra bits
      const uint32_t length = e_value(e) + br.take((int)e_extra(e));
      uint32_t d = T.dist[br.peek(kDistBits)];
      if (e_kind(d) == kSubtable) d = T.dist[e_value(d) + ((br.buf >> kDistBits) & ((1u << kDistSub) - 1))];
      if (e_kind(d) != kDistance) return false;
      br.drop((int)e_len(d));
      const uint32_t distance = e_value(d) + br.take((int)e_extra(d));
      if (br.past_end()) return false;
      if (distance > (size_t)(out - out_begin) || length > (size_t)(out_end - out)) return false;
      br.refill();
      e = T.lit[br.peek(kLitBits)];  // the next symbol's entry, in flight during the copy
      const uint8_t* src = out - distance;
      uint8_t* dst = out;
      out += length;
      const size_t room = (size_t)(out_end - dst);
      if (distance >= 16 && room >= (size_t)length + 64) {  // the usual match: the same columns one line up
        // 64 bytes unconditionally (most matches are shorter: no loop-exit misprediction), 16 at a
        // time so that an overlap at distance >= 16 still reads what was just written
        for (int k = 0; k < 64; k += 16) {
          uint64_t w[2];
          memcpy(w, src + k, 16);
          memcpy(dst + k, w, 16);
        }
        for (uint32_t k = 64; k < length; k += 16) {
          uint64_t w[2];
          memcpy(w, src + k, 16);
          memcpy(dst + k, w, 16);
        }
        continue;
      }
      if (room >= (size_t)length + 8) {  // room to finish the last word past the match
        if (distance >= 8) {
          for (uint32_t k = 0; k < length; k += 8) {
            uint64_t w;
            memcpy(&w, src + k, 8);
            memcpy(dst + k, &w, 8);
          }
          continue;
        }
        if (distance == 1 || distance == 2 || distance == 4) {  // a period that divides 8: one repeated word
          uint64_t w = 0;
          if (distance == 1) {
            w = 0x0101010101010101ull * src[0];
          } else if (distance == 2) {
            uint16_t h;
            memcpy(&h, src, 2);
            w = 0x0001000100010001ull * h;
          } else {
            uint32_t q;
            memcpy(&q, src, 4);
            w = 0x0000000100000001ull * q;
          }
          for (uint32_t k = 0; k < length; k += 8) memcpy(dst + k, &w, 8);
          continue;
        }
      }
      for (uint32_t k = 0; k < length; ++k) dst[k] = src[k];
      continue;
    }
    out_ref = out;
    return kind == kEndOfBlock && !br.past_end();
  }
}

}  // namespace

// Inflates the raw deflate stream [in, in + in_len) into exactly out_len bytes at out.
bool inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, size_t* consumed) {
  static const Tables* const fixed = [] {
    Tables* t = new Tables;
    return build_fixed_tables(*t) ? t : nullptr;
  }();
  BitReader br(in, in_len);
  uint8_t* const out_begin = out;
  uint8_t* const out_end = out + out_len;
  Tables dyn;
  for (;;) {
    br.refill();
    const uint32_t last = br.take(1), type = br.take(2);
    if (type == 0) {  // stored: skip to the byte boundary, LEN / NLEN, bytes
      br.drop(br.bits & 7);
      br.refill();
      const uint32_t len = br.take(16), nlen = br.take(16);
      if ((len ^ 0xffffu) != nlen || br.past_end()) return false;
      // the bit buffer holds whole bytes now (the top `overrun` of them are padding, not input):
      // give the real ones back to the byte pointer
      const int real = (br.bits >> 3) - br.overrun;
      if (real < 0) return false;
      const uint8_t* p = br.in - real;
      if ((size_t)(br.end - p) < len || (size_t)(out_end - out) < len) return false;
      memcpy(out, p, len);
      out += len;
      br.in = p + len;
      br.buf = 0;
      br.bits = 0;
      br.overrun = 0;
    } else if (type == 1) {
      if (!fixed || !inflate_block(br, *fixed, out_begin, out, out_end)) return false;
    } else if (type == 2) {
      if (!read_dynamic_tables(br, dyn) || !inflate_block(br, dyn, out_begin, out, out_end)) return false;
    } else {
      return false;
    }
    if (last) break;
  }
  if (out != out_end || br.past_end()) return false;
  if (consumed) {  // whole bytes still in the bit buffer (minus the padding shifted in at the end) were not used
    const int64_t unused = (int64_t)(br.bits >> 3) - br.overrun;
    *consumed = (size_t)((br.in - in) - (unused > 0 ? unused : 0));
  }
  return true;
}

// ---------------------------------------------------------------------------------- CRC-32
namespace {

struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = t[0][t[s - 1][i] & 0xffu] ^ (t[s - 1][i] >> 8);
  }
};

// state in, state out (no pre/post inversion)
uint32_t crc_tables_update(uint32_t c, const uint8_t* p, size_t n) {
  static const CrcTables T;
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = T.t[7][w & 0xff] ^ T.t[6][(w >> 8) & 0xff] ^ T.t[5][(w >> 16) & 0xff] ^ T.t[4][(w >> 24) & 0xff] ^
        T.t[3][(w >> 32) & 0xff] ^ T.t[2][(w >> 40) & 0xff] ^ T.t[1][(w >> 48) & 0xff] ^ T.t[0][w >> 56];
    p += 8;
    n -= 8;
  }
  while (n--) c = T.t[0][(c ^ *p++) & 0xffu] ^ (c >> 8);
  return c;
}

#ifdef SAI_X86
#pragma GCC push_options
#pragma GCC target("sse4.1,pclmul")
// n >= 64 and a multiple of 16; state in, state out
uint32_t crc_pclmul_update(uint32_t crc, const uint8_t* buf, size_t n) {
  alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};
  alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};
  alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};
  alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};
  __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
  x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
  x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
  x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
  x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
  x0 = _mm_load_si128((const __m128i*)k1k2);
  buf += 64;
  n -= 64;
  while (n >= 64) {  // four lanes folded by 512 bits
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
    x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
    x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
    x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
    y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
    y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
    y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
    x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
    x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
    x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
    buf += 64;
    n -= 64;
  }
  x0 = _mm_load_si128((const __m128i*)k3k4);  // the four lanes into one, 128 bits at a time
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
  while (n >= 16) {
    x2 = _mm_loadu_si128((const __m128i*)buf);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    buf += 16;
    n -= 16;
  }
  x2 = _mm_clmulepi64_si128(x1, x0, 0x10);  // 128 -> 64 bits
  x3 = _mm_setr_epi32(~0, 0, ~0, 0);
  x1 = _mm_srli_si128(x1, 8);
  x1 = _mm_xor_si128(x1, x2);
  x0 = _mm_loadl_epi64((const __m128i*)k5k0);
  x2 = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, x3);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  x0 = _mm_load_si128((const __m128i*)poly);  // Barrett reduction to 32 bits
  x2 = _mm_and_si128(x1, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
  x2 = _mm_and_si128(x2, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}
#pragma GCC pop_options

bool cpu_has_pclmul() {
  __builtin_cpu_init();
  return __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
}
#endif

}  // namespace

// CRC-32 of [p, p + n) as zlib's crc32(0, p, n).  isa: 0 = best available, 1 = tables (tests).
uint32_t crc32_fast(const uint8_t* p, size_t n, int isa) {
  uint32_t c = 0xffffffffu;
#ifdef SAI_X86
  static const bool has = cpu_has_pclmul();
  if (has && isa != 1 && n >= 64) {
    const size_t body = n & ~(size_t)15;
    c = crc_pclmul_update(c, p, body);
    p += body;
    n -= body;
  }
#else
  (void)isa;
#endif
  return ~crc_tables_update(c, p, n);
}

}  // namespace sai
