// Synthetic genotype generator (device side). Integer-only and counter-based so that it is
// bit-identical to oracle/pf_oracle.c:pfo_synth_code and independent of how the site axis
// is sharded across GPUs.
#include "pf_common.cuh"

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ uint32_t hash3(uint64_t seed, uint64_t site, uint32_t k) {
  uint32_t h = static_cast<uint32_t>(seed) ^ 0x9e3779b9u;
  h = mix32(h ^ static_cast<uint32_t>(site));
  h = mix32(h + static_cast<uint32_t>(site >> 32) * 0x7feb352du + 0x165667b1u);
  h = mix32(h ^ (k * 0x9e3779b1u + static_cast<uint32_t>(seed >> 32)));
  return h;
}

// one thread per genotype; x runs over samples (coalesced byte stores along a site row)
__global__ void __launch_bounds__(256)
synth_kernel(uint8_t* __restrict__ codes, int64_t n_samples, int64_t ld, int64_t site0,
             int64_t n_sites, uint64_t seed, int n_pops, uint32_t miss_ppm) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_samples) return;
  const uint32_t pop = static_cast<uint32_t>((i * static_cast<int64_t>(n_pops)) / n_samples);
  for (int64_t s = blockIdx.y; s < n_sites; s += gridDim.y) {
    const uint64_t site = static_cast<uint64_t>(site0 + s);
    const uint32_t h0 = hash3(seed, site, 0xffffffffu);
    const int32_t p_anc = 3277 + static_cast<int32_t>(h0 % 29491u);
    const uint32_t hk = hash3(seed, site, 0x80000000u | pop);
    const int32_t z = static_cast<int32_t>((hk & 0xffu) + ((hk >> 8) & 0xffu) +
                                           ((hk >> 16) & 0xffu) + (hk >> 24)) - 510;
    int32_t p = p_anc + (p_anc * z) / 1024;
    p = max(655, min(64880, p));
    const uint32_t hg = hash3(seed, site, static_cast<uint32_t>(i));
    uint32_t g = ((hg & 0xffffu) < static_cast<uint32_t>(p)) + ((hg >> 16) < static_cast<uint32_t>(p));
    if (miss_ppm) {
      const uint32_t hm = hash3(seed ^ 0xa5a5a5a5deadbeefULL, site, static_cast<uint32_t>(i));
      if (hm % 1000000u < miss_ppm) g = 3;
    }
    codes[s * ld + i] = static_cast<uint8_t>(g);
  }
}

}  // namespace

PF_API int pf_synth_codes(uint8_t* d_codes, int64_t n_samples, int64_t ld, int64_t site0,
                          int64_t n_sites, uint64_t seed, int n_pops, uint32_t miss_ppm,
                          void* stream) {
  PF_TRY_BEGIN
  if (!d_codes || n_samples <= 0 || n_sites < 0 || ld < n_samples || n_pops <= 0)
    return pf_fail("pf_synth_codes: bad arguments (n_samples=%lld ld=%lld n_sites=%lld n_pops=%d)",
                   (long long)n_samples, (long long)ld, (long long)n_sites, n_pops);
  if (n_sites == 0) return PF_OK;
  dim3 grid(static_cast<unsigned>(pf_ceil_div(n_samples, 256)),
            static_cast<unsigned>(n_sites < 65535 ? n_sites : 65535));
  pf_count_launch();
  synth_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_codes, n_samples, ld, site0,
                                                                     n_sites, seed, n_pops, miss_ppm);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}
