// Raw-deflate decoder and CRC-32 of the BGZF reader (inflate_fast.cpp: plain C++).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace sai {
// Inflates the raw deflate stream [in, in + in_len) into exactly out_len bytes at out; false on
// invalid data, a different inflated size, or anything the decoder does not expect (the caller
// falls back to zlib).  Never touches memory outside the two ranges.
// *consumed (optional) = input bytes the stream occupies (the final block's last, partly used byte included).
bool inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, size_t* consumed = nullptr);
// CRC-32 of [p, p + n) == zlib's crc32(0, p, n).  isa: 0 = best available (PCLMULQDQ), 1 = tables.
uint32_t crc32_fast(const uint8_t* p, size_t n, int isa);
}  // namespace sai
