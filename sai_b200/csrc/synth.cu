// Synthetic genotype generator (bench / tests only; not on the product path).
//
// Writes packed tiles directly on the device so that a 1000-Genomes-scale
// matrix (6 M sites x 2504 diploid individuals = 3.8 GB packed, 120 GB in the
// reference's int64 form) never has to exist on the host.  Counter-based RNG:
// every (seed, site, population, individual) draw is a pure hash, so any tile
// range can be regenerated independently and the data do not depend on the
// launch geometry.
//
// Site model (DESIGN.md "Synthetic inputs"):
//   class  = hash(site) : 0.5 % "introgressed", rest "background"
//   background: derived-allele frequency f ~ Beta(0.2, 2.0) (inverse CDF of a
//               uniform hash, so the draw stays a pure function of (seed, site));
//               ref, tgt and src individuals draw alleles ~ Bernoulli(f)
//   introgressed: src fixed derived; ref f = 0.0005; tgt f ~ U(0, 0.8)
//   each individual is missing with probability `missing_rate`
#include "common.cuh"

namespace sai {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct SynthParams {
  sai_layout lay;
  uint2* packed;
  int64_t tile0, n_tiles, n_sites;
  int32_t role[SAI_MAX_POPS];
  uint64_t seed;
  uint32_t miss_thr;  // individual missing iff (hash >> 32) < miss_thr
};

__device__ __forceinline__ void site_model(uint64_t seed, int64_t site, int role, uint32_t& f16) {
  const uint64_t h = mix64(seed ^ (uint64_t)site * 0xD1B54A32D192ED03ull);
  const bool intro = (h & 0xffff) < 328;  // 328/65536 = 0.5 %
  const double u1 = (double)((h >> 16) & 0xffffff) * (1.0 / 16777216.0);
  const double u2 = (double)((h >> 40) & 0xffffff) * (1.0 / 16777216.0);
  double f;
  if (!intro) {
    // Beta(0.2, 2): CDF(x) = 1.2 x^0.2 - 0.2 x^1.2.  With y = x^0.2: 1.2 y - 0.2 y^6 = u, monotone on
    // [0, 1]; 40 bisection steps resolve y to 2^-40, then f = y^5.
    double lo = 0.0, hi = 1.0;
    for (int it = 0; it < 40; ++it) {
      const double y = 0.5 * (lo + hi);
      const double y2 = y * y;
      if (1.2 * y - 0.2 * y2 * y2 * y2 < u1) lo = y; else hi = y;
    }
    const double y = 0.5 * (lo + hi), y2 = y * y;
    f = y2 * y2 * y;
  } else if (role == 0) {
    f = 0.0005;
  } else if (role == 1) {
    f = 0.8 * u2;
  } else {
    f = 1.0;
  }
  f16 = (uint32_t)(f * 65536.0 + 0.5);
  if (f16 > 65536u) f16 = 65536u;
}

// One thread per (tile, pair, site-lane): builds the pair's two words.
__global__ void __launch_bounds__(256) k_synth(const __grid_constant__ SynthParams P) {
  const int64_t pps = P.lay.pairs_per_site;
  const int64_t total = P.n_tiles * pps * kTile;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int lane = (int)(idx % kTile);
    const int64_t rest = idx / kTile;
    const int pair = (int)(rest % pps);
    const int64_t T = P.tile0 + rest / pps;
    const int64_t site = T * kTile + lane;
    // population owning this pair
    int pi = 0;
    while (pi + 1 < P.lay.n_pops && pair >= P.lay.pop[pi + 1].pair_off) ++pi;
    const sai_pop_layout& L = P.lay.pop[pi];
    const int B = L.bits;
    const int n_words = L.n_groups * B;
    uint32_t out[2] = {0u, 0u};
    uint32_t f16 = 0;
    if (site < P.n_sites) site_model(P.seed, site, P.role[pi], f16);
    for (int h = 0; h < 2; ++h) {
      const int word = (pair - L.pair_off) * 2 + h;
      if (word >= n_words) continue;  // zero padding word
      const int g = word / B, plane = word % B;
      uint32_t bits = 0;
      for (int i = 0; i < 32; ++i) {
        const int ind = g * 32 + i;
        int code = (1 << B) - 1;  // missing
        if (site < P.n_sites && ind < L.n_samples) {
          const uint64_t r = mix64(P.seed ^ mix64((uint64_t)site * 0x9E3779B97F4A7C15ull +
                                                  ((uint64_t)pi << 40) + (uint64_t)ind));
          if ((uint32_t)(r >> 32) >= P.miss_thr) {
            int v = 0;
            uint64_t rr = r;
            for (int a = 0; a < L.ploidy; ++a) {
              if (a == 2) rr = mix64(r);  // 16 bits per allele, two per hash half
              v += ((uint32_t)(rr & 0xffff) < f16) ? 1 : 0;
              rr >>= 16;
            }
            code = v;
          }
        }
        bits |= (uint32_t)((code >> plane) & 1) << i;
      }
      out[h] = bits;
    }
    P.packed[(size_t)(T * pps + pair) * kTile + lane] = make_uint2(out[0], out[1]);
  }
}

}  // namespace sai

using namespace sai;

extern "C" int sai_synth_fill(const sai_layout* lay, void* d_packed, int64_t tile0,
                              int64_t n_tiles, int64_t n_sites, const int32_t* role,
                              uint64_t seed, double missing_rate, void* stream) {
  if (int rc = validate_layout(lay)) return rc;
  SAI_REQUIRE(d_packed && role && tile0 >= 0 && n_tiles >= 0, "bad argument");
  SAI_REQUIRE(missing_rate >= 0.0 && missing_rate < 1.0, "missing_rate outside [0,1)");
  for (int p = 0; p < lay->n_pops; ++p)
    SAI_REQUIRE(lay->pop[p].ploidy <= 4, "synthetic generator supports ploidy <= 4");
  if (n_tiles == 0) return SAI_OK;
  SynthParams P{};
  P.lay = *lay;
  P.packed = static_cast<uint2*>(d_packed);
  P.tile0 = tile0;
  P.n_tiles = n_tiles;
  P.n_sites = n_sites;
  for (int p = 0; p < lay->n_pops; ++p) P.role[p] = role[p];
  P.seed = seed;
  P.miss_thr = (uint32_t)(missing_rate * 4294967296.0);
  const int64_t total = n_tiles * lay->pairs_per_site * kTile;
  const int64_t want = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 32;
  k_synth<<<(int)(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(P);
  SAI_CUDA_CHECK(cudaGetLastError());
  return SAI_OK;
}
