// Block bootstrap of the count matrices (the `-b 1000` of the reference's
// `treebest nj -b 1000`, treetool/script.py:359-360, made reproducible). The site axis is cut into B
// contiguous blocks whose V / D1 / D2 matrices are kept; a replicate draws B blocks with replacement
// (counter-based integer hash, identical to oracle/pf_oracle.c:pfo_bootstrap_weights) and its count
// matrices are the weighted sum of the block matrices — integer arithmetic, so the CPU oracle reproduces
// every replicate bit for bit; distances and NJ then run per replicate with the existing kernels.
#include "pf_common.cuh"

namespace {

__device__ __forceinline__ uint32_t bs_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}
// same function as hash3 of synth.cu / pfo_hash3 of the oracle
__device__ __forceinline__ uint32_t bs_hash3(uint64_t seed, uint64_t a, uint32_t k) {
  uint32_t h = static_cast<uint32_t>(seed) ^ 0x9e3779b9u;
  h = bs_mix32(h ^ static_cast<uint32_t>(a));
  h = bs_mix32(h + static_cast<uint32_t>(a >> 32) * 0x7feb352du + 0x165667b1u);
  h = bs_mix32(h ^ (k * 0x9e3779b1u + static_cast<uint32_t>(seed >> 32)));
  return h;
}

constexpr uint64_t BS_TAG = 0xb007573a9b007573ULL;

// weights[r][b] = how often block b is drawn among the n_blocks draws of replicate replicate0 + r
__global__ void __launch_bounds__(256)
bootstrap_weights_kernel(uint64_t seed, int64_t replicate0, int n_reps, int n_blocks, int32_t* __restrict__ w) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(n_reps) * n_blocks) return;
  const int r = static_cast<int>(idx / n_blocks), t = static_cast<int>(idx % n_blocks);
  const uint32_t h = bs_hash3(seed ^ BS_TAG, static_cast<uint64_t>(replicate0 + r), static_cast<uint32_t>(t));
  atomicAdd(w + static_cast<int64_t>(r) * n_blocks + (h % static_cast<uint32_t>(n_blocks)), 1);
}

constexpr int BS_MAX_REPS = 8;

// out[r][e] = sum_b w[r][b] * blocks[b][e]; one thread per group of four consecutive elements, every
// block matrix is read once for all (up to 8) replicates of the call
template <int R>
__global__ void __launch_bounds__(256)
bootstrap_counts_kernel(const int4* __restrict__ blocks, int n_blocks, int64_t vec_per_block, const int32_t* __restrict__ w,
                        int4* __restrict__ out) {
  extern __shared__ int32_t s_w[];   // [R][n_blocks]
  for (int i = threadIdx.x; i < R * n_blocks; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= vec_per_block) return;
  int4 acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = make_int4(0, 0, 0, 0);
  for (int b = 0; b < n_blocks; ++b) {
    const int4 v = __ldg(blocks + static_cast<int64_t>(b) * vec_per_block + e);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int32_t k = s_w[r * n_blocks + b];
      acc[r].x += k * v.x;
      acc[r].y += k * v.y;
      acc[r].z += k * v.z;
      acc[r].w += k * v.w;
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) out[static_cast<int64_t>(r) * vec_per_block + e] = acc[r];
}

template <int R>
cudaError_t bs_launch(const int32_t* blocks, int n_blocks, int64_t vec, const int32_t* w, int32_t* out, cudaStream_t s) {
  const size_t smem = static_cast<size_t>(R) * n_blocks * sizeof(int32_t);
  cudaError_t e = cudaFuncSetAttribute(bootstrap_counts_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  pf_count_launch();
  bootstrap_counts_kernel<R><<<static_cast<unsigned>(pf_ceil_div(vec, static_cast<int64_t>(256))), 256, smem, s>>>(
      reinterpret_cast<const int4*>(blocks), n_blocks, vec, w, reinterpret_cast<int4*>(out));
  return cudaGetLastError();
}

}  // namespace

// See include/phyloforge_b200.h
PF_API int pf_bootstrap_weights(uint64_t seed, int64_t replicate0, int n_reps, int n_blocks, int32_t* d_weights,
                                void* stream) {
  PF_TRY_BEGIN
  if (!d_weights || n_reps <= 0 || n_blocks <= 0 || replicate0 < 0) return pf_fail("pf_bootstrap_weights: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>(n_reps) * n_blocks;
  PF_CUDA(cudaMemsetAsync(d_weights, 0, static_cast<size_t>(total) * sizeof(int32_t), s));
  pf_count_launch();
  bootstrap_weights_kernel<<<static_cast<unsigned>(pf_ceil_div(total, static_cast<int64_t>(256))), 256, 0, s>>>(
      seed, replicate0, n_reps, n_blocks, d_weights);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

PF_API int pf_bootstrap_counts(const int32_t* d_blocks, int n_blocks, int64_t elems_per_block, const int32_t* d_weights,
                               int n_reps, int32_t* d_out, void* stream) {
  PF_TRY_BEGIN
  if (!d_blocks || !d_weights || !d_out || n_blocks <= 0 || n_blocks > 8192 || elems_per_block <= 0 ||
      elems_per_block % 4 != 0 || n_reps <= 0 || n_reps > BS_MAX_REPS)
    return pf_fail("pf_bootstrap_counts: bad arguments (elements per block must be a multiple of 4, at most %d "
                   "replicates per call, at most 8192 blocks)", BS_MAX_REPS);
  if ((reinterpret_cast<uintptr_t>(d_blocks) | reinterpret_cast<uintptr_t>(d_out)) & 15)
    return pf_fail("pf_bootstrap_counts: buffers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t vec = elems_per_block / 4;
  cudaError_t e = cudaSuccess;
  switch (n_reps) {
    case 1: e = bs_launch<1>(d_blocks, n_blocks, vec, d_weights, d_out, s); break;
    case 2: e = bs_launch<2>(d_blocks, n_blocks, vec, d_weights, d_out, s); break;
    case 3: e = bs_launch<3>(d_blocks, n_blocks, vec, d_weights, d_out, s); break;
    case 4: e = bs_launch<4>(d_blocks, n_blocks, vec, d_weights, d_out, s); break;
    case 5: e = bs_launch<5>(d_blocks, n_blocks, vec, d_weights, d_out, s); break;
    case 6: e = bs_launch<6>(d_blocks, n_blocks, vec, d_weights, d_out, s); break;
    case 7: e = bs_launch<7>(d_blocks, n_blocks, vec, d_weights, d_out, s); break;
    default: e = bs_launch<8>(d_blocks, n_blocks, vec, d_weights, d_out, s); break;
  }
  PF_CUDA(e);
  return PF_OK;
  PF_TRY_END
}
