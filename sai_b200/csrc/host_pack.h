// Host packer core (pack_simd.cpp: plain C++ with run-time selected x86 vector paths).
#pragma once
#include <stdint.h>

#include "../../include/sai_b200.h"

namespace sai {

// Packs population `pop` of tiles [t0, t1) from the row-major int8 matrix `gt`
// (gt[site * row_stride + individual], n_sites rows) into the tile buffer whose tile `tile_base`
// starts at `packed_base`; true = a value does not fit the population's bit-planes.
// isa: 0 = best available on this CPU, 1 = portable, 2 = sse2, 3 = avx2, 4 = avx512bw, 5 = avx512gfni (tests;
// an unavailable choice falls back to the best available).
bool pack_tiles_i8(const sai_layout& lay, int pop, const int8_t* gt, int64_t n_sites, int64_t row_stride,
                   int64_t t0, int64_t t1, int64_t tile_base, uint8_t* packed_base, int isa);
// All populations, site by site (one sequential stream when the populations are column blocks of
// one matrix): the int8 pipeline's packer.  Full cache lines go out with non-temporal stores when
// the destination is 64-byte aligned, unless `allow_nt` is false (a tile that is consumed again
// by this core, e.g. by the zt encoder).
bool pack_tiles_i8_all(const sai_layout& lay, const int8_t* const* gt, const int64_t* row_stride, int64_t n_sites,
                       int64_t t0, int64_t t1, int64_t tile_base, uint8_t* packed_base, int isa, bool allow_nt = true);
// n_lines 64-byte lines src -> dst (both 64-byte aligned) with non-temporal stores; stream_fence()
// orders them before a later publication of the buffer.
void stream_lines(uint8_t* dst, const uint8_t* src, size_t n_lines);
void stream_fence();
const char* pack_isa();

}  // namespace sai
