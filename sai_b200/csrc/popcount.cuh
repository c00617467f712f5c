// Streaming loads and the carry-save (Harley-Seal) population counter shared by the genotype
// passes (k_site, k_site_hist).
#pragma once
#include "common.cuh"

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
  // four words: one weight-4 carry, absorbed by `fours` (one POPC per 4 words instead of per 8;
  // used by the 3- and 4-plane populations, whose groups come 4 per batch of loads)
  __device__ __forceinline__ void add4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    const uint32_t tA = maj3(ones, w0, w1);
    ones = xor3(ones, w0, w1);
    const uint32_t tB = maj3(ones, w2, w3);
    ones = xor3(ones, w2, w3);
    const uint32_t fA = maj3(twos, tA, tB);
    twos = xor3(twos, tA, tB);
    high += __popc(fours & fA);
    fours ^= fA;
  }
  __device__ __forceinline__ int total() const {
    return 8 * high + 4 * __popc(fours) + 2 * __popc(twos) + __popc(ones);
  }
};

}  // namespace sai
