// Sanitizer fuzz of the host ingest (bgzf.cu, vcf_parse.cu, vcf_simd.cpp, inflate_fast.cpp compiled as
// plain C++): random VCFs -- truncated and haploid records, other chromosomes, lines of very
// different lengths, a missing final newline, bgzip blocks of 1 to 5000 bytes -- must give the same
// rows through the fused bgzip read (every group size / thread count) as through the text parser,
// with every buffer an exact-size heap allocation.
//   g++ -O1 -g -fsanitize=address,undefined -std=c++17 -pthread -I/usr/local/cuda/include \
//       -x c++ sai_b200/csrc/bgzf.cu sai_b200/csrc/vcf_parse.cu sai_b200/csrc/vcf_simd.cpp \
//       sai_b200/csrc/inflate_fast.cpp tools/ingest_fuzz.cpp -lz -o tools/bin/ingest_fuzz && tools/bin/ingest_fuzz
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <zlib.h>
#include <cstdarg>
#include "../include/sai_b200.h"
namespace sai {  // what api.cu provides in the library
void set_error(const char* fmt, ...) { (void)fmt; }
int sm_count() { return 1; }
}  // namespace sai

static std::vector<uint8_t> bgzf(const std::string& text, size_t block) {
  std::vector<uint8_t> out;
  for (size_t at = 0; ; at += block) {
    size_t n = at < text.size() ? std::min(block, text.size() - at) : 0;
    std::vector<uint8_t> comp(compressBound(n) + 64);
    z_stream zs; memset(&zs, 0, sizeof zs); deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = (Bytef*)text.data() + (n ? at : 0); zs.avail_in = n; zs.next_out = comp.data(); zs.avail_out = comp.size();
    deflate(&zs, Z_FINISH); size_t cn = zs.total_out; deflateEnd(&zs);
    uint32_t crc = crc32(0, (const Bytef*)text.data() + (n ? at : 0), n);
    size_t bsize = 18 + cn + 8;
    uint8_t hdr[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0, (uint8_t)((bsize - 1) & 255), (uint8_t)((bsize - 1) >> 8)};
    out.insert(out.end(), hdr, hdr + 18); out.insert(out.end(), comp.begin(), comp.begin() + cn);
    for (int k = 0; k < 4; ++k) out.push_back((crc >> (8 * k)) & 255);
    for (int k = 0; k < 4; ++k) out.push_back((n >> (8 * k)) & 255);
    if (n == 0) break;
  }
  return out;
}
int main() {
  srand(3);
  for (int round = 0; round < 60; ++round) {
    int n_smp = 5 + rand() % 80, n_rec = 1 + rand() % 300;
    std::string text = "##meta\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
    for (int i = 0; i < n_smp; ++i) text += "\ts" + std::to_string(i);
    text += "\n";
    size_t skip = text.size();
    const char* tok[5] = {"0|0", "0|1", "1|1", ".|.", "1"};
    for (int r = 0; r < n_rec; ++r) {
      text += (rand() % 9 == 0 ? "2\t" : "1\t") + std::to_string(10 + 7 * r) + "\t.\tA\tG\t.\t.\t" + std::string(rand() % 5 == 0 ? rand() % 400 : 1, 'x') + "\tGT";
      int nf = rand() % 11 == 0 ? 1 + rand() % n_smp : n_smp;
      for (int i = 0; i < nf; ++i) { text += "\t"; text += tok[rand() % 20 == 0 ? 4 : rand() % 4]; }
      if (r + 1 < n_rec || rand() % 2) text += "\n";
    }
    size_t block = 1 + rand() % (rand() % 2 ? 200 : 5000);
    std::vector<uint8_t> file = bgzf(text, block);
    // exact-size heap copy so that reads past the file are seen
    uint8_t* data = (uint8_t*)malloc(file.size()); memcpy(data, file.data(), file.size());
    int64_t nb_cap = file.size() / 26 + 8;
    std::vector<int64_t> boff(nb_cap), ooff(nb_cap + 1); int64_t consumed = 0;
    int64_t n = sai_bgzf_scan(data, file.size(), nb_cap, 1ll << 60, boff.data(), ooff.data(), &consumed);
    if (n <= 0) { printf("scan failed\n"); return 1; }
    std::vector<int32_t> col, pl; int n_out = 1 + rand() % n_smp;
    for (int o = 0; o < n_out; ++o) { col.push_back(rand() % 3 ? o : rand() % n_smp); pl.push_back(rand() % 4 ? 2 : 1 + rand() % 3); }
    // reference: parse the text directly
    int64_t cap = n_rec + 4;
    std::vector<int32_t> p0(cap), p1(cap); std::vector<int8_t> g0(cap * n_out), g1(cap * n_out);
    std::string body = text.substr(skip); if (body.empty() || body.back() != '\n') body += "\n";
    int64_t used = 0;
    int64_t r0 = sai_vcf_parse_gt(body.data(), body.size(), "1", 1, 0, col.data(), pl.data(), n_out, nullptr, nullptr, 0, p0.data(), g0.data(), n_out, cap, &used, 2);
    for (int G : {1, 2, 5, 16}) for (int th : {1, 3}) {
      int64_t r1 = sai_bgzf_parse_gt(data, boff.data(), ooff.data(), n, skip, "1", 1, 0, col.data(), pl.data(), n_out, nullptr, nullptr, 0, p1.data(), g1.data(), n_out, cap, G, th);
      if (r1 != r0 || memcmp(p0.data(), p1.data(), r0 * 4) || memcmp(g0.data(), g1.data(), r0 * n_out)) { printf("MISMATCH round %d G %d th %d rows %ld vs %ld block %zu\n", round, G, th, (long)r1, (long)r0, block); return 1; }
    }
    free(data);
  }
  printf("fused bgzf == text parse on 60 random files x 8 (group, thread) settings\n");
  return 0;
}
