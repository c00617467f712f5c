// Host memory-system probe for the int8 packer's ceiling: N threads stream over private slices of a
// large buffer (a) reading only (AVX-512 OR-reduction), (b) reading + writing 1/4 of the bytes read
// with non-temporal stores -- the packer's traffic mix (int8 in, bit-plane tiles out).
//   g++ -O3 -march=native -pthread tools/host_membw.cpp -o tools/bin/host_membw && tools/bin/host_membw [GB]
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__attribute__((target("avx512f"))) static uint64_t read_slice(const uint8_t* p, size_t n) {
  __m512i a = _mm512_setzero_si512(), b = a, c = a, d = a;
  for (size_t i = 0; i + 256 <= n; i += 256) {
    a = _mm512_or_si512(a, _mm512_loadu_si512(p + i));
    b = _mm512_or_si512(b, _mm512_loadu_si512(p + i + 64));
    c = _mm512_or_si512(c, _mm512_loadu_si512(p + i + 128));
    d = _mm512_or_si512(d, _mm512_loadu_si512(p + i + 192));
  }
  a = _mm512_or_si512(_mm512_or_si512(a, b), _mm512_or_si512(c, d));
  return (uint64_t)_mm512_reduce_or_epi64(a);
}

__attribute__((target("avx512f"))) static uint64_t read_write_slice(const uint8_t* p, size_t n, uint8_t* out) {
  for (size_t i = 0; i + 256 <= n; i += 256) {
    __m512i a = _mm512_or_si512(_mm512_or_si512(_mm512_loadu_si512(p + i), _mm512_loadu_si512(p + i + 64)),
                                _mm512_or_si512(_mm512_loadu_si512(p + i + 128), _mm512_loadu_si512(p + i + 192)));
    _mm512_stream_si512((__m512i*)(out + i / 4), a);
  }
  _mm_sfence();
  return 0;
}

int main(int argc, char** argv) {
  const double gb = argc > 1 ? atof(argv[1]) : 8.0;
  const size_t n = (size_t)(gb * 1e9) / 4096 * 4096;
  uint8_t* in = (uint8_t*)aligned_alloc(4096, n);
  uint8_t* out = (uint8_t*)aligned_alloc(4096, n / 4 + 4096);
  const unsigned hw = std::thread::hardware_concurrency();
  {  // first touch in parallel
    std::vector<std::thread> th;
    for (unsigned t = 0; t < hw; ++t)
      th.emplace_back([&, t] {
        const size_t a = n / hw * t / 4096 * 4096, b = t + 1 == hw ? n : n / hw * (t + 1) / 4096 * 4096;
        memset(in + a, 1, b - a);
        memset(out + a / 4, 0, (b - a) / 4);
      });
    for (auto& x : th) x.join();
  }
  printf("{\"gb\": %.2f, \"cpus\": %u", n / 1e9, hw);
  for (int mode = 0; mode < 2; ++mode)
    for (unsigned T : {1u, 2u, 4u, 8u, 12u, 16u, 32u}) {
      if (T > hw) break;
      double best = 1e9;
      volatile uint64_t sink = 0;
      for (int rep = 0; rep < 3; ++rep) {
        std::vector<std::thread> th;
        const double t0 = now();
        for (unsigned t = 0; t < T; ++t)
          th.emplace_back([&, t] {
            const size_t a = n / T * t / 4096 * 4096, b = t + 1 == T ? n : n / T * (t + 1) / 4096 * 4096;
            sink = sink | (mode ? read_write_slice(in + a, b - a, out + a / 4) : read_slice(in + a, b - a));
          });
        for (auto& x : th) x.join();
        best = std::min(best, now() - t0);
      }
      printf(", \"%s_%u\": %.1f", mode ? "read_plus_quarter_nt_write_gbps_of_read" : "read_gbps", T, n / best / 1e9);
    }
  printf("}\n");
  return 0;
}
