// Self-test of the tcgen05 plumbing in umma.cuh: C[128][256] = A[128][K] * B[256][K]^T with
// signed 8-bit operands and 32-bit accumulation, one CTA, operands staged into the canonical
// K-major no-swizzle layout by ordinary stores. tests/test_gpu_umma.py checks it against numpy.
#include "pf_common.cuh"
#include "phyloforge_b200_probes.h"
#include "umma.cuh"

namespace {

constexpr int ST_M = 128, ST_N = 256;

// byte offset of element (row r, k-byte kb) inside one 32-byte-K operand tile
__device__ __forceinline__ uint32_t canon_off(int r, int kb) {
  return static_cast<uint32_t>((r >> 3) * 256 + (kb >> 4) * 128 + (r & 7) * 16 + (kb & 15));
}

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ C, int K,
                     int a_in_tmem) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int ksteps = K / 32;
  uint8_t* sa = smem;                          // ksteps tiles of 128 x 32 B
  uint8_t* sb = smem + ksteps * ST_M * 32;     // ksteps tiles of 256 x 32 B
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
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
  if (a_in_tmem) {
    // A operand through TMEM: thread = row; the 32 K-bytes of k-step ks go to columns 256 + 8 ks .. + 7
    const int row = warp * 32 + lane;
    for (int ks = 0; ks < ksteps; ++ks) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t x = 0;
        for (int b = 0; b < 4; ++b) x |= static_cast<uint32_t>(static_cast<uint8_t>(A[row * K + ks * 32 + 4 * j + b])) << (8 * b);
        v[j] = x;
      }
      umma::tmem_st_32x8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 256 + 8 * ks, v);
    }
    umma::tmem_wait_st();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
  }

  if (tid == 0) {
    const uint32_t idesc = umma::idesc_i8(ST_M, ST_N);
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t da = umma::smem_desc_kmajor_noswizzle(umma::smem_addr(sa + ks * ST_M * 32), 128, 256);
      const uint64_t db = umma::smem_desc_kmajor_noswizzle(umma::smem_addr(sb + ks * ST_N * 32), 128, 256);
      if (a_in_tmem) umma::mma_i8_ts(tmem, tmem + 256 + 8 * ks, db, idesc, ks > 0 ? 1u : 0u);
      else umma::mma_i8_ss(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
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
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

PF_API int pf_selftest_umma(const int8_t* d_a, const int8_t* d_b, int32_t* d_c, int k, int a_in_tmem, void* stream) {
  PF_TRY_BEGIN
  if (!d_a || !d_b || !d_c || k <= 0 || k % 32 != 0 || k > 256) return pf_fail("pf_selftest_umma: bad arguments");
  const size_t smem = static_cast<size_t>(k / 32) * (ST_M + ST_N) * 32;
  PF_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem)));
  umma_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(d_a, d_b, d_c, k, a_in_tmem);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

// ---- MMA throughput probe: back-to-back kind::i8 MMAs on resident operands (no loads in the loop)
namespace {
template <int N>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int reps, int a_in_tmem, int group, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
  if (tid == 0) {
    umma::mbar_init(&s_bar, 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < (ST_M + N) * 32 * 2 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    const uint32_t idesc = umma::idesc_i8(ST_M, N);
    const uint32_t sa = umma::smem_addr(smem), sb = sa + ST_M * 32 * 2;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      // two products per iteration, like the pair-count kernel
      const uint64_t da0 = umma::smem_desc_kmajor_noswizzle(sa, 128, 256);
      const uint64_t da1 = umma::smem_desc_kmajor_noswizzle(sa + ST_M * 32, 128, 256);
      const uint64_t db0 = umma::smem_desc_kmajor_noswizzle(sb, 128, 256);
      const uint64_t db1 = umma::smem_desc_kmajor_noswizzle(sb + N * 32, 128, 256);
      if (a_in_tmem && group < 0) {
        // -group independent accumulators used round-robin (A staging sits behind them)
        const int n_acc = -group;
        const uint32_t a_col = tmem + n_acc * N;
        umma::mma_i8_ts(tmem + ((2 * r) % n_acc) * N, a_col, db0, idesc, 1u);
        umma::mma_i8_ts(tmem + ((2 * r + 1) % n_acc) * N, a_col + 8, db1, idesc, 1u);
      } else if (a_in_tmem && group > 1) {
        // `group` consecutive MMAs into the same accumulator before switching to the other one
        const uint32_t d = ((r / group) & 1) ? tmem + N : tmem;
        umma::mma_i8_ts(d, tmem + 2 * N, db0, idesc, 1u);
        umma::mma_i8_ts(d, tmem + 2 * N + 8, db1, idesc, 1u);
      } else if (a_in_tmem) {
        umma::mma_i8_ts(tmem, tmem + 2 * N, db0, idesc, 1u);
        umma::mma_i8_ts(tmem + N, tmem + 2 * N + 8, db1, idesc, 1u);
      } else {
        umma::mma_i8_ss(tmem, da0, db0, idesc, 1u);
        umma::mma_i8_ss(tmem + N, da1, db1, idesc, 1u);
      }
    }
    umma::commit(&s_bar);
    umma::mbar_wait(&s_bar, 0);
    out_cycles[blockIdx.x] = clock64() - t0;
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}
}  // namespace

// n = 128, 192 or 256 accumulator columns per product; reports SM-clock64 ticks per MMA (mean over CTAs)
PF_API int pf_probe_umma_rate(int n, int reps, int a_in_tmem, int group, double* ticks_per_mma, void* stream) {
  PF_TRY_BEGIN
  if ((n != 64 && n != 96 && n != 128 && n != 192 && n != 256) || reps <= 0 || !ticks_per_mma)
    return pf_fail("pf_probe_umma_rate: bad arguments");
  if (group < 0 && (-group) * n + 16 > 512) return pf_fail("pf_probe_umma_rate: accumulators do not fit in TMEM");
  if (a_in_tmem && n == 256) return pf_fail("pf_probe_umma_rate: no TMEM columns left for A at n = 256");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 0;
  PF_CUDA(cudaGetDevice(&dev));
  PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long* d = nullptr;
  PF_CUDA(cudaMalloc(&d, sizeof(long long) * sms));
  const size_t smem = static_cast<size_t>(ST_M + n) * 32 * 2;
  auto launch = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kernel<<<sms, 128, smem, s>>>(reps, a_in_tmem, group, d);
    return cudaGetLastError();
  };
  cudaError_t e = n == 64 ? launch(umma_rate_kernel<64>) : n == 96 ? launch(umma_rate_kernel<96>) : n == 128 ? launch(umma_rate_kernel<128>)
                  : (n == 192 ? launch(umma_rate_kernel<192>) : launch(umma_rate_kernel<256>));
  std::vector<long long> h(sms);
  if (e == cudaSuccess) e = cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  cudaFree(d);
  PF_CUDA(e);
  double mean = 0;
  for (int i = 0; i < sms; ++i) mean += double(h[i]);
  *ticks_per_mma = mean / sms / (2.0 * reps);
  return PF_OK;
  PF_TRY_END
}
