// Host encoder of one zt record (zt_simd.cpp: plain C++ with a run-time selected AVX-512 path).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/sai_b200.h"

namespace sai {

// scratch sizes for a tile of P pair rows: the record under construction / its data bytes
inline size_t zt_record_cap(int P) { return 4 + (size_t)P * (4 + 32 + 256) + 64; }
inline size_t zt_tmp_cap(int P) { return (size_t)P * 256 + 64; }

// Unpadded length of the record of the dense tile `tile` (P rows of 32 pairs; padc[r] = padding
// constant of row r).  isa: 0 = best available, 1 = portable (tests).
size_t zt_tile_size(const uint64_t* tile, int P, const uint64_t* padc, int isa);
// Writes the record to `rec` (zt_record_cap(P) bytes; `tmp`: zt_tmp_cap(P) bytes) and returns its
// unpadded length.  The caller pads to 8 bytes, or stores the tile raw when that is not smaller.
size_t zt_encode_tile(const uint64_t* tile, int P, const uint64_t* padc, uint8_t* rec, uint8_t* tmp, int isa);
const char* zt_isa();  // "avx512vbmi2" or "portable"

}  // namespace sai
