// Self-test of the tcgen05 plumbing in umma.cuh: C[128][256] = A[128][K] * B[256][K]^T with
// signed 8-bit operands and 32-bit accumulation, one CTA, operands staged into the canonical
// K-major no-swizzle layout by ordinary stores. tests/test_gpu_umma.py checks it against numpy.
#include "pf_common.cuh"
#include "umma.cuh"

namespace {

constexpr int ST_M = 128, ST_N = 256;

// byte offset of element (row r, k-byte kb) inside one 32-byte-K operand tile
__device__ __forceinline__ uint32_t canon_off(int r, int kb) {
  return static_cast<uint32_t>((r >> 3) * 256 + (kb >> 4) * 128 + (r & 7) * 16 + (kb & 15));
}

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ C, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int ksteps = K / 32;
  uint8_t* sa = smem;                          // ksteps tiles of 128 x 32 B
  uint8_t* sb = smem + ksteps * ST_M * 32;     // ksteps tiles of 256 x 32 B
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) umma::tmem_alloc(&s_tmem, 256);
  if (tid == 0) {
    umma::mbar_init(&s_bar, 1);
    umma::fence_mbar_init();
  }
  for (int idx = tid; idx < ST_M * K; idx += blockDim.x) {
    const int r = idx / K, k = idx % K;
    sa[(k >> 5) * ST_M * 32 + canon_off(r, k & 31)] = static_cast<uint8_t>(A[idx]);
  }
  for (int idx = tid; idx < ST_N * K; idx += blockDim.x) {
    const int r = idx / K, k = idx % K;
    sb[(k >> 5) * ST_N * 32 + canon_off(r, k & 31)] = static_cast<uint8_t>(B[idx]);
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;

  if (tid == 0) {
    const uint32_t idesc = umma::idesc_i8(ST_M, ST_N);
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t da = umma::smem_desc_kmajor_noswizzle(umma::smem_addr(sa + ks * ST_M * 32), 128, 256);
      const uint64_t db = umma::smem_desc_kmajor_noswizzle(umma::smem_addr(sb + ks * ST_N * 32), 128, 256);
      umma::mma_i8_ss(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    umma::commit(&s_bar);
  }
  umma::mbar_wait(&s_bar, 0);
  umma::fence_after_thread_sync();

  // warp w owns TMEM lanes [32w, 32w + 32): thread = one accumulator row
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < ST_N; c0 += 32) {
    uint32_t v[32];
    umma::tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    umma::tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) C[row * ST_N + c0 + j] = static_cast<int32_t>(v[j]);
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace

PF_API int pf_selftest_umma(const int8_t* d_a, const int8_t* d_b, int32_t* d_c, int k, void* stream) {
  PF_TRY_BEGIN
  if (!d_a || !d_b || !d_c || k <= 0 || k % 32 != 0 || k > 512) return pf_fail("pf_selftest_umma: bad arguments");
  const size_t smem = static_cast<size_t>(k / 32) * (ST_M + ST_N) * 32;
  PF_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem)));
  umma_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(d_a, d_b, d_c, k);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}
