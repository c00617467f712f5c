// Sanitizer fuzz of the int8 pipeline's host code (pack_simd.cpp, zt_simd.cpp): on random layouts
// (1-4 populations of 1-1700 individuals, 2-8 bit-planes, 1-150 sites) the best vector packer --
// aligned output with non-temporal stores and unaligned output -- must write the bytes of the
// portable per-population packer, and the zt block encoder (ordinary and non-temporal stores) must
// produce the portable encoder's record sizes and offsets, all inside exact-size heap buffers.
//   g++ -O1 -g -fsanitize=address,undefined -std=c++17 -pthread -Iinclude tools/pack_fuzz.cpp \
//       sai_b200/csrc/pack_simd.cpp sai_b200/csrc/zt_simd.cpp -o tools/bin/pack_fuzz && tools/bin/pack_fuzz
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../sai_b200/csrc/host_pack.h"
#include "../sai_b200/csrc/zt_simd.h"
int main() {
  srand(11);
  for (int round = 0; round < 300; ++round) {
    sai_layout lay; memset(&lay, 0, sizeof lay);
    lay.n_pops = 1 + rand() % 4; int off = 0, tot = 0;
    int B = 2 + rand() % 7;
    for (int p = 0; p < lay.n_pops; ++p) {
      auto& L = lay.pop[p]; L.n_samples = 1 + rand() % (rand() % 3 ? 100 : 1700); L.ploidy = 2; L.bits = rand() % 3 ? B : 2 + rand() % 7;
      L.n_groups = (L.n_samples + 31) / 32; L.n_pairs = (L.n_groups * L.bits + 1) / 2; L.pair_off = off; off += L.n_pairs; tot += L.n_samples;
    }
    lay.pairs_per_site = off;
    int64_t S = 1 + rand() % 150; int64_t nt = (S + 31) / 32;
    size_t tile_bytes = (size_t)off * 256, bytes = nt * tile_bytes;
    int8_t* g = (int8_t*)malloc((size_t)S * tot);
    for (size_t i = 0; i < (size_t)S * tot; ++i) g[i] = rand() % 12 == 0 ? -1 - rand() % 2 : (rand() % 9 == 0 ? 1 : 0);
    std::vector<const int8_t*> gt; std::vector<int64_t> rs; int at = 0;
    for (int p = 0; p < lay.n_pops; ++p) { gt.push_back(g + at); rs.push_back(tot); at += lay.pop[p].n_samples; }
    uint8_t* a = (uint8_t*)aligned_alloc(64, (bytes + 63) / 64 * 64); uint8_t* b = (uint8_t*)malloc(bytes); uint8_t* c = (uint8_t*)malloc(bytes + 1);
    memset(a, 0xEE, bytes); memset(b, 0xEE, bytes); memset(c + 1, 0xEE, bytes);
    bool bad = sai::pack_tiles_i8_all(lay, gt.data(), rs.data(), S, 0, nt, 0, a, 0, true);       // best ISA, aligned (NT stores)
    sai::pack_tiles_i8_all(lay, gt.data(), rs.data(), S, 0, nt, 0, c + 1, 0, true);               // unaligned
    for (int p = 0; p < lay.n_pops; ++p) sai::pack_tiles_i8(lay, p, gt[p], S, tot, 0, nt, 0, b, 1);  // portable, per population
    if (memcmp(a, b, bytes) || memcmp(c + 1, b, bytes)) { printf("PACK MISMATCH round %d\n", round); return 1; }
    // zt block encoder: records into an exact-size region, both store modes; sizes agree with the portable tile encoder
    std::vector<uint64_t> padc(off, 0);
    for (int p = 0; p < lay.n_pops; ++p) { auto& L = lay.pop[p]; for (int r = 0; r < L.n_pairs; ++r) for (int h = 0; h < 2; ++h) { int w = 2 * r + h; if (w >= L.n_groups * L.bits) continue; int real = L.n_samples - 32 * (w / L.bits); uint32_t m = real >= 32 ? 0u : (0xffffffffu << real); padc[L.pair_off + r] |= (uint64_t)m << (32 * h); } }
    sai::ZtBlockScratch sc(off);
    uint8_t* region = (uint8_t*)aligned_alloc(64, bytes); uint8_t* region2 = (uint8_t*)aligned_alloc(64, bytes);
    std::vector<uint64_t> o1(nt + 1), o2(nt + 1); bool bad1 = false, bad2 = false;
    size_t u1 = sai::zt_pack_block_i8(lay, gt.data(), rs.data(), S, 0, nt, padc.data(), region, 0, o1.data(), sc, false, &bad1);
    size_t u2 = sai::zt_pack_block_i8(lay, gt.data(), rs.data(), S, 0, nt, padc.data(), region2, 0, o2.data(), sc, true, &bad2);
    size_t want = 0; std::vector<uint8_t> rec(sai::zt_record_cap(off)), tmp(sai::zt_tmp_cap(off));
    for (int64_t T = 0; T < nt; ++T) { size_t n = (sai::zt_encode_tile((const uint64_t*)(b + T * tile_bytes), off, padc.data(), rec.data(), tmp.data(), 1) + 7) & ~size_t(7); if (n >= tile_bytes) n = tile_bytes; if ((o1[T] & ~(1ull << 63)) != want || o1[T] != o2[T]) { printf("ZT OFFSET MISMATCH round %d\n", round); return 1; } want += n; }
    if (u1 != want || u2 != ((want + 63) & ~size_t(63)) || memcmp(region, region2, want) || bad1 != bad || bad2 != bad) { printf("ZT MISMATCH round %d: %zu %zu %zu\n", round, u1, u2, want); return 1; }
    free(g); free(a); free(b); free(c); free(region); free(region2);
  }
  printf("packer paths agree and the zt block encoder stays inside its buffers on 300 random layouts\n");
}
