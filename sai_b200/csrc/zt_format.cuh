// zt record format helpers shared by the codec (zt_codec.cu) and the int8 pipeline (engine.cu).
#pragma once
#include "common.cuh"

namespace sai {

// padding constant of row (pair index) r: ones for the unused individuals of the
// population's last group in every plane word of the pair
__host__ __device__ inline uint64_t pad_constant(const sai_layout& lay, int r) {
  int pi = 0;
  while (pi + 1 < lay.n_pops && r >= lay.pop[pi + 1].pair_off) ++pi;
  const sai_pop_layout& L = lay.pop[pi];
  uint64_t c = 0;
  for (int h = 0; h < 2; ++h) {
    const int word = (r - L.pair_off) * 2 + h;
    if (word >= L.n_groups * L.bits) continue;  // zero padding word of an odd word count
    const int real = L.n_samples - 32 * (word / L.bits);
    const uint32_t bits = real >= 32 ? 0u : (0xffffffffu << real);
    c |= (uint64_t)bits << (32 * h);
  }
  return c;
}

}  // namespace sai
