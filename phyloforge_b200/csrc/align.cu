// Alignment characters -> 2-bit genotype codes (the inverse of the reference's re-encoding).
//
// `treebest nj` (treetool/script.py:359-360) and the other tree tools receive the character matrix
// the converters wrote: all_snp.fa (IUPAC characters, treetool/vcftree.py:26-65, 168-171) or
// SV.matrix ('0' / '1' / '.', treetool/svtree.py:122-125). To serve that boundary — a FASTA in, a
// tree out — the columns are mapped back to codes: the (at most two) homozygous symbols of a column
// become 0 and 2 in alphabetical order, the IUPAC code of that pair becomes 1, '-', 'N' and '.'
// become 3. Which allele is called 0 does not matter: V, D1 and D2 are symmetric under 0 <-> 2.
// Columns with more than two alleles or with '*'-derived lower-case characters do not fit 2 bits
// and are flagged (PF_LINE_FITS2BIT clear) for the caller to skip or reject.
#include "pf_common.cuh"

namespace {

// bit mask over A=1, C=2, G=4, T=8 of the bases a character stands for; 0 = not a base/het code
__device__ __forceinline__ uint32_t base_mask(uint32_t c) {
  switch (c) {
    case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8;
    case 'M': return 1 | 2; case 'R': return 1 | 4; case 'W': return 1 | 8;
    case 'S': return 2 | 4; case 'Y': return 2 | 8; case 'K': return 4 | 8;
    default: return 0;
  }
}

// one warp per site row of the site-major character matrix
__global__ void __launch_bounds__(256)
chars_to_codes_kernel(const uint8_t* __restrict__ chars, int64_t ld, int64_t n_sites, int64_t n_samples,
                      uint8_t* __restrict__ codes, int64_t ldc, int32_t* __restrict__ line_info) {
  const int lane = threadIdx.x & 31;
  const int64_t s = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= n_sites) return;
  const uint8_t* row = chars + s * ld;
  uint32_t bases = 0;      // union of the bases of every base / het character
  uint32_t binary = 0;     // bit 0: a '0' seen, bit 1: a '1' seen (SV matrix columns)
  uint32_t bad = 0;
  for (int64_t i = lane; i < n_samples; i += 32) {
    const uint32_t c = row[i];
    const uint32_t m = base_mask(c);
    if (m) bases |= m;
    else if (c == '0') binary |= 1;
    else if (c == '1') binary |= 2;
    else if (c != '-' && c != 'N' && c != '.') bad = 1;
  }
  bases = __reduce_or_sync(0xffffffffu, bases);
  binary = __reduce_or_sync(0xffffffffu, binary);
  bad = __reduce_or_sync(0xffffffffu, bad);
  const bool fits = !bad && __popc(bases) <= 2 && !(bases && binary);
  // the two alleles in alphabetical order; a single allele -> only code 0 occurs
  const uint32_t lo = bases & (0u - bases);          // lowest set bit
  const uint32_t hi = bases & ~lo;                   // the other base (or 0)
  uint8_t* out = codes + s * ldc;
  for (int64_t i = lane; i < n_samples; i += 32) {
    const uint32_t c = row[i];
    uint32_t g = 3;
    if (fits) {
      const uint32_t m = base_mask(c);
      if (m) g = (m == lo) ? 0u : ((m == hi) ? 2u : 1u);   // a het code is the union lo | hi
      else if (c == '0') g = 0;
      else if (c == '1') g = 2;
    }
    out[i] = static_cast<uint8_t>(g);
  }
  if (lane == 0) line_info[s] = 1 | (fits ? 4 : 0);
}

}  // namespace

PF_API int pf_chars_to_codes(const uint8_t* d_chars, int64_t ld, int64_t n_sites, int64_t n_samples,
                             uint8_t* d_codes, int64_t ldc, int32_t* d_line_info, void* stream) {
  PF_TRY_BEGIN
  if (n_sites == 0) return PF_OK;
  if (!d_chars || !d_codes || !d_line_info || n_sites < 0 || n_samples <= 0 || ld < n_samples || ldc < n_samples)
    return pf_fail("pf_chars_to_codes: bad arguments");
  const int64_t grid = pf_ceil_div(n_sites, 8);
  if (grid > 0x7fffffffLL) return pf_fail("pf_chars_to_codes: too many sites");
  chars_to_codes_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_chars, ld, n_sites, n_samples, d_codes, ldc, d_line_info);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}
