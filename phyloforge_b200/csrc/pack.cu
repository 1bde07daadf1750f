// Genotype packer: site-major byte codes -> tiled bit planes (and back).
//
// This is the byte-transform half of the reference's converters once genotypes are codes
// (treetool/vcftree.py:120-133, treetool/svtree.py:84-120): an HBM-bound kernel. The VCF is
// site-major and the pair-count kernel wants, for each 32-site word, the words of 64
// consecutive samples contiguous — so no transpose through shared memory is needed: a
// thread owns 4 consecutive samples, reads one coalesced 32-bit word (4 codes) from each of
// the 32 site rows of a word, and keeps the bits in registers.
//   algorithmic bytes per genotype: 1 read + 0.25 written  (DESIGN.md "Kernels")
#include "pf_common.cuh"

namespace {

// 4x4 byte transpose: in a[q] byte k belongs to sample k; out w[k] byte q comes from a[q]
__device__ __forceinline__ void transpose4x4(const uint32_t a[4], uint32_t w[4]) {
  const uint32_t t0 = __byte_perm(a[0], a[1], 0x5140);
  const uint32_t t1 = __byte_perm(a[2], a[3], 0x5140);
  const uint32_t t2 = __byte_perm(a[0], a[1], 0x7362);
  const uint32_t t3 = __byte_perm(a[2], a[3], 0x7362);
  w[0] = __byte_perm(t0, t1, 0x5410);
  w[1] = __byte_perm(t0, t1, 0x7632);
  w[2] = __byte_perm(t2, t3, 0x5410);
  w[3] = __byte_perm(t2, t3, 0x7632);
}

// blockDim = (GX, GY): x over groups of 4 samples, y over words
template <bool ALIGNED>
__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t* __restrict__ codes, int64_t ld, const int64_t* __restrict__ rows,
            int64_t n_samples, int64_t n_sites, uint32_t* __restrict__ planes, int64_t n_words,
            int64_t word0, int64_t n_groups) {
  const int64_t g = static_cast<int64_t>(blockIdx.y) * blockDim.x + threadIdx.x;  // sample group
  const int64_t wl = static_cast<int64_t>(blockIdx.x) * blockDim.y + threadIdx.y;  // local word
  const int64_t n_local_words = pf_ceil_div(n_sites, 32);
  if (g >= n_groups || wl >= n_local_words) return;
  const int64_t i0 = g * 4;
  uint32_t accL[4], accH[4];
  if (i0 >= n_samples) {
    // padding samples of the last tile: code 3 everywhere
#pragma unroll
    for (int q = 0; q < 4; ++q) accL[q] = accH[q] = 0xffffffffu;
  } else {
    const bool full = ALIGNED && (i0 + 3 < n_samples);
    const int64_t s0 = wl * 32;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t x[8];
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int64_t s = s0 + q * 8 + b;
        uint32_t v = 0x03030303u;
        if (s < n_sites) {
          const int64_t r = rows ? rows[s] : s;
          const uint8_t* p = codes + r * ld + i0;
          if (full) {
            v = *reinterpret_cast<const uint32_t*>(p);
          } else {
            v = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t c = (i0 + k < n_samples) ? p[k] : 3u;
              v |= c << (8 * k);
            }
          }
        }
        x[b] = v;
      }
      uint32_t l = 0, h = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        l |= (x[b] & 0x01010101u) << b;
        h |= ((x[b] >> 1) & 0x01010101u) << b;
      }
      accL[q] = l;
      accH[q] = h;
    }
  }
  uint32_t wL[4], wH[4];
  if (i0 >= n_samples) {
#pragma unroll
    for (int k = 0; k < 4; ++k) wL[k] = wH[k] = 0xffffffffu;
  } else {
    transpose4x4(accL, wL);
    transpose4x4(accH, wH);
  }
  const int64_t tile = i0 / PF_TILE;
  const int64_t lane = i0 % PF_TILE;
  uint32_t* base = planes + ((tile * n_words + (word0 + wl)) * 2) * PF_TILE + lane;
  *reinterpret_cast<uint4*>(base) = make_uint4(wL[0], wL[1], wL[2], wL[3]);
  *reinterpret_cast<uint4*>(base + PF_TILE) = make_uint4(wH[0], wH[1], wH[2], wH[3]);
}

// ---- site-major 2-bit codes (4 samples per byte, sample 4b+k in bits 2k..2k+1 of byte b) -> planes.
// A thread owns 16 consecutive samples: one 32-bit load per site row, then a 32 x 32 bit transpose in
// registers turns "32 sites x (16 samples x 2 bits)" into the L and H words of the 16 samples.
// plink != 0 reads PLINK .bed genotype codes (00 hom A1, 01 missing, 10 het, 11 hom A2) instead.
//   algorithmic bytes per genotype: 0.25 read + 0.25 written
__device__ __forceinline__ void transpose32(uint32_t (&a)[32]) {
  uint32_t m = 0x0000ffffu;
#pragma unroll
  for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
    for (int k = 0; k < 32; k = ((k | j) + 1) & ~j) {
      const uint32_t t = ((a[k] >> j) ^ a[k | j]) & m;
      a[k] ^= t << j;
      a[k | j] ^= t;
    }
  }
}

__global__ void __launch_bounds__(128)
pack2_kernel(const uint8_t* __restrict__ codes2, int64_t ld2, int64_t n_samples, int64_t n_sites,
             uint32_t* __restrict__ planes, int64_t n_words, int64_t word0, int64_t n_groups, int plink) {
  const int64_t g = static_cast<int64_t>(blockIdx.y) * blockDim.x + threadIdx.x;   // group of 16 samples
  const int64_t wl = static_cast<int64_t>(blockIdx.x) * blockDim.y + threadIdx.y;  // local word
  const int64_t n_local_words = pf_ceil_div(n_sites, 32);
  if (g >= n_groups || wl >= n_local_words) return;
  const int64_t i0 = g * 16;
  const uint32_t missing = plink ? 0x55555555u : 0xffffffffu;
  uint32_t a[32];
  // samples of this group beyond n_samples read as missing
  uint32_t keep = 0xffffffffu;
  if (i0 >= n_samples) keep = 0u;
  else if (i0 + 16 > n_samples) keep = (1u << (2 * static_cast<int>(n_samples - i0))) - 1u;
  const bool in_row = (i0 / 4 + 4) <= ld2;   // the 4 bytes of this group lie inside a row
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    const int64_t sidx = wl * 32 + b;
    uint32_t v = missing;
    if (sidx < n_sites && keep != 0u) {
      const uint8_t* p = codes2 + sidx * ld2 + i0 / 4;
      uint32_t x;
      if (in_row) {
        x = *reinterpret_cast<const uint32_t*>(p);
      } else {
        x = 0;
        for (int k = 0; k < 4; ++k)
          if (i0 / 4 + k < ld2) x |= static_cast<uint32_t>(p[k]) << (8 * k);
      }
      v = (x & keep) | (missing & ~keep);
    }
    a[b] = v;
  }
  transpose32(a);   // a[2j] = low code bits of sample i0+j over the 32 sites, a[2j+1] = high code bits
  const int64_t tile = i0 / PF_TILE;
  const int64_t lane = i0 % PF_TILE;
  uint32_t* base = planes + ((tile * n_words + (word0 + wl)) * 2) * PF_TILE + lane;
  uint32_t L[16], H[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint32_t lo = a[2 * j], hi = a[2 * j + 1];
    L[j] = plink ? (lo ^ hi) : lo;   // PLINK: het = 10, missing = 01 -> low code bit; hom A2 = 11, missing -> high bit
    H[j] = plink ? lo : hi;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    *reinterpret_cast<uint4*>(base + 4 * q) = make_uint4(L[4 * q], L[4 * q + 1], L[4 * q + 2], L[4 * q + 3]);
    *reinterpret_cast<uint4*>(base + PF_TILE + 4 * q) = make_uint4(H[4 * q], H[4 * q + 1], H[4 * q + 2], H[4 * q + 3]);
  }
}

// planes -> sample-major byte codes. One CTA per (tile, 32 words): coalesced 512 B reads per
// word, transposed through shared memory, 1 KB contiguous writes per sample.
__global__ void __launch_bounds__(256)
unpack_kernel(const uint32_t* __restrict__ planes, int64_t n_words, int64_t word0,
              int64_t n_samples, int64_t n_sites, uint8_t* __restrict__ out, int64_t ld_out) {
  __shared__ uint32_t sm[32][2 * PF_TILE + 1];
  const int64_t tile = blockIdx.y;
  const int64_t wl0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int64_t n_local_words = pf_ceil_div(n_sites, 32);
  for (int idx = threadIdx.x; idx < 32 * 2 * PF_TILE; idx += blockDim.x) {
    const int w = idx / (2 * PF_TILE), c = idx % (2 * PF_TILE);
    uint32_t v = 0xffffffffu;
    if (wl0 + w < n_local_words) v = planes[((tile * n_words + word0 + wl0 + w) * 2) * PF_TILE + c];
    sm[w][c] = v;
  }
  __syncthreads();
  // 64 samples x 32 words; thread handles (sample, word) and writes 32 bytes
  for (int idx = threadIdx.x; idx < PF_TILE * 32; idx += blockDim.x) {
    const int l = idx / 32, w = idx % 32;
    const int64_t i = tile * PF_TILE + l;
    if (i >= n_samples || wl0 + w >= n_local_words) continue;
    const uint32_t lo = sm[w][l], hi = sm[w][PF_TILE + l];
    uint8_t* dst = out + i * ld_out + (wl0 + w) * 32;
    const int64_t s0 = (wl0 + w) * 32;
#pragma unroll 8
    for (int b = 0; b < 32; ++b)
      if (s0 + b < n_sites) dst[b] = static_cast<uint8_t>(((lo >> b) & 1u) | (((hi >> b) & 1u) << 1));
  }
}

// out[i * ld_out + k] = in[rows[k] * ld_in + i], 64x64 byte tiles through shared memory
__global__ void __launch_bounds__(256)
transpose_kernel(const uint8_t* __restrict__ in, int64_t ld_in, const int64_t* __restrict__ rows,
                 int64_t n_rows, int64_t n_cols, uint8_t* __restrict__ out, int64_t ld_out) {
  __shared__ uint8_t sm[64][65];
  const int64_t k0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int64_t i0 = static_cast<int64_t>(blockIdx.y) * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int r = ty; r < 64; r += 4) {
    const int64_t k = k0 + r;
    uint8_t v = 0;
    if (k < n_rows && i0 + tx < n_cols) {
      const int64_t src = rows ? rows[k] : k;
      v = in[src * ld_in + i0 + tx];
    }
    sm[r][tx] = v;
  }
  __syncthreads();
  for (int c = ty; c < 64; c += 4) {
    const int64_t i = i0 + c;
    if (i < n_cols && k0 + tx < n_rows) out[i * ld_out + k0 + tx] = sm[tx][c];
  }
}

// Shift a packed matrix by `shift` sites (0..31) towards higher site indices: site s of the source becomes site
// s + shift of the destination; the vacated low bits of word 0 and everything behind site n_sites + shift is
// code 3 (L = H = 1: missing). One thread per (tile, destination word, plane, lane) entry.
__global__ void __launch_bounds__(256)
planes_shift_kernel(const uint32_t* __restrict__ src, int64_t n_words_src, uint32_t* __restrict__ dst,
                    int64_t n_words_dst, int64_t n_tiles, int64_t n_sites, int shift, int64_t words_out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t per_tile = words_out * 2 * PF_TILE;
  if (idx >= n_tiles * per_tile) return;
  const int64_t T = idx / per_tile;
  const int64_t rem = idx - T * per_tile;
  const int64_t w = rem / (2 * PF_TILE);
  const int pl = static_cast<int>(rem - w * 2 * PF_TILE);          // plane * 64 + lane
  const int64_t src_words = (n_sites + 31) / 32;
  auto load = [&](int64_t ws) -> uint32_t {                        // source word with the padding behind n_sites forced to 1
    if (ws < 0 || ws >= src_words) return 0xffffffffu;
    uint32_t v = src[(T * n_words_src + ws) * 2 * PF_TILE + pl];
    const int64_t left = n_sites - ws * 32;
    if (left < 32) v |= ~((1u << left) - 1u);
    return v;
  };
  const uint32_t hi = load(w), lo = load(w - 1);
  const uint32_t v = shift ? ((hi << shift) | (lo >> (32 - shift))) : hi;
  dst[(T * n_words_dst + w) * 2 * PF_TILE + pl] = v;
}


}  // namespace

PF_API int64_t pf_planes_len(int64_t n_samples, int64_t n_words) {
  if (n_samples <= 0 || n_words < 0) return 0;
  return pf_ceil_div(n_samples, PF_TILE) * n_words * 2 * PF_TILE;
}

PF_API int pf_pack_codes_u8(const uint8_t* d_codes, int64_t ld, const int64_t* d_rows,
                            int64_t n_samples, int64_t n_sites, uint32_t* d_planes,
                            int64_t n_words, int64_t word0, void* stream) {
  PF_TRY_BEGIN
  if (!d_codes || !d_planes || n_samples <= 0 || n_sites < 0 || ld < n_samples || word0 < 0)
    return pf_fail("pf_pack_codes_u8: bad arguments");
  const int64_t n_local_words = pf_ceil_div(n_sites, 32);
  if (word0 + n_local_words > n_words)
    return pf_fail("pf_pack_codes_u8: words [%lld, %lld) exceed n_words=%lld", (long long)word0,
                   (long long)(word0 + n_local_words), (long long)n_words);
  if (n_sites == 0) return PF_OK;
  const int64_t n_groups = pf_ceil_div(n_samples, PF_TILE) * (PF_TILE / 4);
  int gx = 16;
  while (gx < 256 && gx < n_groups) gx *= 2;
  const int gy = 256 / gx;
  dim3 block(gx, gy);
  dim3 grid(static_cast<unsigned>(pf_ceil_div(n_local_words, gy)),
            static_cast<unsigned>(pf_ceil_div(n_groups, gx)));
  const bool aligned = (ld % 4 == 0) && (reinterpret_cast<uintptr_t>(d_codes) % 4 == 0);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  pf_count_launch();
  if (aligned)
    pack_kernel<true><<<grid, block, 0, s>>>(d_codes, ld, d_rows, n_samples, n_sites, d_planes,
                                             n_words, word0, n_groups);
  else
    pack_kernel<false><<<grid, block, 0, s>>>(d_codes, ld, d_rows, n_samples, n_sites, d_planes,
                                              n_words, word0, n_groups);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

PF_API int pf_unpack_codes_u8(const uint32_t* d_planes, int64_t n_words, int64_t word0,
                              int64_t n_samples, int64_t n_sites, uint8_t* d_out, int64_t ld_out,
                              void* stream) {
  PF_TRY_BEGIN
  if (!d_planes || !d_out || n_samples <= 0 || n_sites < 0 || ld_out < n_sites || word0 < 0)
    return pf_fail("pf_unpack_codes_u8: bad arguments");
  const int64_t n_local_words = pf_ceil_div(n_sites, 32);
  if (word0 + n_local_words > n_words) return pf_fail("pf_unpack_codes_u8: word range exceeds n_words");
  if (n_sites == 0) return PF_OK;
  dim3 grid(static_cast<unsigned>(pf_ceil_div(n_local_words, 32)),
            static_cast<unsigned>(pf_ceil_div(n_samples, PF_TILE)));
  pf_count_launch();
  unpack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_planes, n_words, word0, n_samples,
                                                                      n_sites, d_out, ld_out);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

PF_API int pf_transpose_u8(const uint8_t* d_in, int64_t ld_in, const int64_t* d_rows, int64_t n_rows,
                           int64_t n_cols, uint8_t* d_out, int64_t ld_out, void* stream) {
  PF_TRY_BEGIN
  if (!d_in || !d_out || n_rows < 0 || n_cols < 0 || ld_in < n_cols || ld_out < n_rows)
    return pf_fail("pf_transpose_u8: bad arguments");
  if (n_rows == 0 || n_cols == 0) return PF_OK;
  const int64_t gy = pf_ceil_div(n_cols, 64);
  if (gy > 65535) return pf_fail("pf_transpose_u8: n_cols too large");
  dim3 grid(static_cast<unsigned>(pf_ceil_div(n_rows, 64)), static_cast<unsigned>(gy));
  pf_count_launch();
  transpose_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_in, ld_in, d_rows, n_rows, n_cols,
                                                                         d_out, ld_out);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

PF_API int pf_pack_codes_2bit(const uint8_t* d_codes2, int64_t ld2, int64_t n_samples, int64_t n_sites,
                              uint32_t* d_planes, int64_t n_words, int64_t word0, int plink, void* stream) {
  PF_TRY_BEGIN
  if (!d_codes2 || !d_planes || n_samples <= 0 || n_sites < 0 || ld2 * 4 < n_samples || word0 < 0)
    return pf_fail("pf_pack_codes_2bit: bad arguments");
  if (ld2 % 4 != 0 || reinterpret_cast<uintptr_t>(d_codes2) % 4 != 0)
    return pf_fail("pf_pack_codes_2bit: rows must be 4-byte aligned (ld2 %% 4 == 0)");
  const int64_t n_local_words = pf_ceil_div(n_sites, 32);
  if (word0 + n_local_words > n_words) return pf_fail("pf_pack_codes_2bit: word range exceeds n_words");
  if (n_sites == 0) return PF_OK;
  const int64_t n_groups = pf_ceil_div(n_samples, PF_TILE) * (PF_TILE / 16);
  int gx = 4;
  while (gx < 128 && gx < n_groups) gx *= 2;
  const int gy = 128 / gx;
  dim3 block(gx, gy);
  dim3 grid(static_cast<unsigned>(pf_ceil_div(n_local_words, gy)), static_cast<unsigned>(pf_ceil_div(n_groups, gx)));
  pf_count_launch();
  pack2_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(d_codes2, ld2, n_samples, n_sites, d_planes,
                                                                       n_words, word0, n_groups, plink);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

PF_API int pf_planes_shift(const uint32_t* d_src, int64_t n_words_src, uint32_t* d_dst, int64_t n_words_dst,
                           int64_t n_samples, int64_t n_sites, int shift, void* stream) {
  PF_TRY_BEGIN
  if (!d_src || !d_dst || d_src == d_dst || n_samples <= 0 || n_sites < 0 || shift < 0 || shift > 31)
    return pf_fail("pf_planes_shift: bad arguments");
  const int64_t words_out = pf_ceil_div(n_sites + shift, 32);
  if (pf_ceil_div(n_sites, 32) > n_words_src || words_out > n_words_dst)
    return pf_fail("pf_planes_shift: word range exceeds a buffer");
  if (words_out == 0) return PF_OK;
  const int64_t n_tiles = pf_ceil_div(n_samples, PF_TILE);
  const int64_t total = n_tiles * words_out * 2 * PF_TILE;
  pf_count_launch();
  planes_shift_kernel<<<static_cast<unsigned>(pf_ceil_div(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_src, n_words_src, d_dst, n_words_dst, n_tiles, n_sites, shift, words_out);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}
