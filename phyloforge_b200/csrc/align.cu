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

// Columns that do not fit the 2-bit code (more than two alleles, '*'-derived lower-case characters) are counted
// from the characters the reference writes to all_snp.fa (treetool/vcftree.py:26-65; BASELINE.md section 5: every
// character stands for one unordered allele pair over {A, C, G, T, *}). For a pair of samples and a column where
// neither character is missing ('-', 'N', '.'):
//   D1 = the characters differ (what a mismatch model sees; SURVEY.md 8(c).3, the p-distance),
//   D1 + D2 = 2 - (alleles shared, with multiplicity) - the allele-sharing (IBS) distance, which on a biallelic
//             column is |g_i - g_j| - so D2 = the genotypes share no allele.
// Both come from the ordinary packer and pair-count kernels fed with pseudo genotype codes:
//  * one pass per character STATE s (code 0 where the character is s, 2 where it is another non-missing one, 3
//    where missing): V_s = both non-missing (the same for every s), D1_s = exactly one of the two is s. A column
//    where the characters differ is counted by exactly two states, an equal one by none: mismatches = sum_s D1_s / 2;
//  * one pass per ALLELE a (code = copies of a in the genotype: 0, 1, 2; 3 where missing): D1_a + D2_a =
//    |c_a(i) - c_a(j)|, and sum_a |c_a(i) - c_a(j)| = 4 - 2 shared, so sum_a (D1_a + D2_a) / 2 = 2 - shared.
__global__ void __launch_bounds__(256)
chars_state_codes_kernel(const uint8_t* __restrict__ chars, int64_t ld, const int64_t* __restrict__ rows, int64_t n_rows,
                         int64_t n_samples, int state, uint8_t* __restrict__ codes, int64_t ldc) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t k = idx / ldc, i = idx % ldc;
  if (k >= n_rows) return;
  uint32_t g = 3;
  if (i < n_samples) {
    const uint32_t c = chars[(rows ? rows[k] : k) * ld + i];
    g = (c == static_cast<uint32_t>(state)) ? 0u : ((c == '-' || c == 'N' || c == '.') ? 3u : 2u);
  }
  codes[k * ldc + i] = static_cast<uint8_t>(g);
}

// the allele pair (indices 0..4 = A, C, G, T, *) a character of all_snp.fa stands for; false = missing / not a genotype
__device__ __forceinline__ bool allele_pair(uint32_t c, int& a1, int& a2) {
  switch (c) {
    case 'A': a1 = 0; a2 = 0; return true;  case 'C': a1 = 1; a2 = 1; return true;
    case 'G': a1 = 2; a2 = 2; return true;  case 'T': a1 = 3; a2 = 3; return true;
    case 'M': a1 = 0; a2 = 1; return true;  case 'R': a1 = 0; a2 = 2; return true;
    case 'W': a1 = 0; a2 = 3; return true;  case 'S': a1 = 1; a2 = 2; return true;
    case 'Y': a1 = 1; a2 = 3; return true;  case 'K': a1 = 2; a2 = 3; return true;
    case 'a': a1 = 0; a2 = 4; return true;  case 'c': a1 = 1; a2 = 4; return true;
    case 'g': a1 = 2; a2 = 4; return true;  case 't': a1 = 3; a2 = 4; return true;
    default: return false;
  }
}

__global__ void __launch_bounds__(256)
chars_allele_codes_kernel(const uint8_t* __restrict__ chars, int64_t ld, const int64_t* __restrict__ rows, int64_t n_rows,
                          int64_t n_samples, int allele, uint8_t* __restrict__ codes, int64_t ldc) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t k = idx / ldc, i = idx % ldc;
  if (k >= n_rows) return;
  uint32_t g = 3;
  if (i < n_samples) {
    int a1, a2;
    if (allele_pair(chars[(rows ? rows[k] : k) * ld + i], a1, a2)) g = static_cast<uint32_t>((a1 == allele) + (a2 == allele));
  }
  codes[k * ldc + i] = static_cast<uint8_t>(g);
}

// counts.V += sc.V / n_states, counts.D1 += sc.D1 / 2; with allele counts: counts.D2 += (ac.D1 + ac.D2) / 2 - sc.D1 / 2
// (all divisions are exact, see above)
__global__ void __launch_bounds__(256)
counts_add_extra_kernel(const int32_t* __restrict__ sc, int64_t ld_s, int n_states, const int32_t* __restrict__ ac, int64_t ld_a,
                        int32_t* __restrict__ counts, int64_t n, int64_t ld) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int64_t i = idx / n, j = idx % n;
  const int32_t mism = sc[n * ld_s + i * ld_s + j] / 2;
  counts[i * ld + j] += sc[i * ld_s + j] / n_states;
  counts[n * ld + i * ld + j] += mism;
  if (ac) counts[2 * n * ld + i * ld + j] += (ac[n * ld_a + i * ld_a + j] + ac[2 * n * ld_a + i * ld_a + j]) / 2 - mism;
}

}  // namespace

PF_API int pf_chars_state_codes(const uint8_t* d_chars, int64_t ld, const int64_t* d_rows, int64_t n_rows,
                                int64_t n_samples, int state, uint8_t* d_codes, int64_t ldc, void* stream) {
  PF_TRY_BEGIN
  if (n_rows == 0) return PF_OK;
  if (!d_chars || !d_codes || n_rows < 0 || n_samples <= 0 || ld < n_samples || ldc < n_samples || state <= 0 || state > 255)
    return pf_fail("pf_chars_state_codes: bad arguments");
  const int64_t grid = pf_ceil_div(n_rows * ldc, 256);
  if (grid > 0x7fffffffLL) return pf_fail("pf_chars_state_codes: too many columns");
  pf_count_launch();
  chars_state_codes_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_chars, ld, d_rows, n_rows, n_samples, state, d_codes, ldc);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

PF_API int pf_chars_allele_codes(const uint8_t* d_chars, int64_t ld, const int64_t* d_rows, int64_t n_rows,
                                 int64_t n_samples, int allele, uint8_t* d_codes, int64_t ldc, void* stream) {
  PF_TRY_BEGIN
  if (n_rows == 0) return PF_OK;
  if (!d_chars || !d_codes || n_rows < 0 || n_samples <= 0 || ld < n_samples || ldc < n_samples || allele < 0 || allele > 4)
    return pf_fail("pf_chars_allele_codes: bad arguments");
  const int64_t grid = pf_ceil_div(n_rows * ldc, 256);
  if (grid > 0x7fffffffLL) return pf_fail("pf_chars_allele_codes: too many columns");
  pf_count_launch();
  chars_allele_codes_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_chars, ld, d_rows, n_rows, n_samples, allele, d_codes, ldc);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

PF_API int pf_counts_add_extra(const int32_t* d_state_counts, int64_t ld_state, int n_states, const int32_t* d_allele_counts,
                               int64_t ld_allele, int32_t* d_counts, int64_t n_samples, int64_t ld, void* stream) {
  PF_TRY_BEGIN
  if (!d_state_counts || !d_counts || n_samples <= 0 || ld < n_samples || ld_state < n_samples || n_states <= 0 ||
      (d_allele_counts && ld_allele < n_samples))
    return pf_fail("pf_counts_add_extra: bad arguments");
  const int64_t grid = pf_ceil_div(n_samples * n_samples, 256);
  pf_count_launch();
  counts_add_extra_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_state_counts, ld_state, n_states, d_allele_counts, ld_allele, d_counts, n_samples, ld);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

PF_API int pf_chars_to_codes(const uint8_t* d_chars, int64_t ld, int64_t n_sites, int64_t n_samples,
                             uint8_t* d_codes, int64_t ldc, int32_t* d_line_info, void* stream) {
  PF_TRY_BEGIN
  if (n_sites == 0) return PF_OK;
  if (!d_chars || !d_codes || !d_line_info || n_sites < 0 || n_samples <= 0 || ld < n_samples || ldc < n_samples)
    return pf_fail("pf_chars_to_codes: bad arguments");
  const int64_t grid = pf_ceil_div(n_sites, 8);
  if (grid > 0x7fffffffLL) return pf_fail("pf_chars_to_codes: too many sites");
  pf_count_launch();
  chars_to_codes_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_chars, ld, n_sites, n_samples, d_codes, ldc, d_line_info);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}
