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

// Per-thread scratch of zt_pack_block_i8: one dense tile, one record under construction (+ the
// < 64 bytes carried over from the previous one) and the record's data bytes -- L1/L2 resident.
struct ZtBlockScratch {
  explicit ZtBlockScratch(int P);
  ~ZtBlockScratch();
  ZtBlockScratch(const ZtBlockScratch&) = delete;
  ZtBlockScratch& operator=(const ZtBlockScratch&) = delete;
  uint8_t *mem, *tilebuf, *stage, *tmp;
};

// int8 matrices -> the zt records of tiles [t0, t1), back to back at `region` (64-byte aligned, room
// for the dense tiles of the block): every tile is packed into the scratch tile, encoded while it
// is in L1 and appended; a record that would not be smaller than the tile is the raw tile.
// tile_off[T] = base_off + offset of tile T's record in the region (| SAI_ZT_RAW).  With
// `nontemporal` the region is written as whole 64-byte lines around the caches (the last line
// zero-padded; a pinned staging buffer the copy engine reads next), else with ordinary stores
// and no padding beyond the records' own 8 bytes.  Returns the bytes written; *bad is set when a
// value does not fit its population's bit-planes.
size_t zt_pack_block_i8(const sai_layout& lay, const int8_t* const* gt, const int64_t* row_stride, int64_t n_sites,
                        int64_t t0, int64_t t1, const uint64_t* padc, uint8_t* region, uint64_t base_off,
                        uint64_t* tile_off, ZtBlockScratch& sc, bool nontemporal, bool* bad);

}  // namespace sai
