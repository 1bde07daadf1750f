// Probe of the CTA-pair form of the 4-bit tensor-core path: tcgen05.mma.cta_group::2 kind::mxf4
// (E2M1 operands, K = 64, fp32 accumulation, block scale factors all 1.0) on a cluster of two CTAs.
// One instruction computes a 256 x n tile: each CTA of the pair supplies 128 A rows and n / 2 B rows
// from its own shared memory and accumulates its 128 rows x n columns in its own tensor memory.
// pf_selftest_umma_fp4_2cta checks exactness against integer dot products (and the cross-CTA hand-over
// the pair-count kernel uses: generic-proxy operand writes, proxy fence, remote mbarrier arrive with
// cluster-scope release, multicast commit); pf_probe_umma_fp4_2cta_rate measures clocks per instruction.
// Diagnostic only.
#include "pf_common.cuh"
#include "phyloforge_b200_probes.h"
#include "umma.cuh"

namespace {

constexpr int P2_SF_COL = 480;   // TMEM columns [480, 512): scale factors, all 1.0

__host__ __device__ constexpr uint32_t idesc_mxf4_2cta(int m, int n) {
  return (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (1u << 23) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t e2m1_of(int v) {
  const uint32_t s = v < 0 ? 8u : 0u;
  const int a = v < 0 ? -v : v;
  const uint32_t m = a == 0 ? 0u : a == 1 ? 2u : a == 2 ? 4u : a == 3 ? 5u : a == 4 ? 6u : 7u;
  return s | m;
}

__device__ __forceinline__ uint32_t canon_off(int r, int kb) {
  return static_cast<uint32_t>((r >> 3) * 256 + (kb >> 4) * 128 + (r & 7) * 16 + (kb & 15));
}

__device__ __forceinline__ void fill_sf(uint32_t tmem, int warp) {
  uint32_t ones[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ones[j] = 0x7f7f7f7fu;
  for (int c = 0; c < 32; c += 8) umma::tmem_st_32x8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + P2_SF_COL + c, ones);
  umma::tmem_wait_st();
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_fp4_2cta_selftest_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, float* __restrict__ C, int N,
                              int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_ready, s_done;
  const int ksteps = K / 64;
  const int nh = N / 2;
  const uint32_t rank = umma::cluster_ctarank();
  uint8_t* sa = smem;                         // ksteps tiles of 128 x 32 B (this CTA's A rows)
  uint8_t* sb = smem + ksteps * 128 * 32;     // ksteps tiles of nh x 32 B (this CTA's B rows)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) umma::tmem_alloc_2cta(&s_tmem, 512);
  if (tid == 0) {
    umma::mbar_init(&s_ready, 8);             // one arrival per warp of both CTAs
    umma::mbar_init(&s_done, 1);
    umma::fence_mbar_init();
  }
  umma::fence_before_thread_sync();
  umma::cluster_sync();                       // barriers of both CTAs are initialised before any remote arrive
  umma::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;
  fill_sf(tmem, warp);
  const int8_t* Ar = A + static_cast<size_t>(rank) * 128 * K;
  const int8_t* Br = B + static_cast<size_t>(rank) * nh * K;
  for (int idx = tid; idx < 128 * K / 2; idx += blockDim.x) {
    const int r = idx / (K / 2), kb = idx % (K / 2);
    const uint32_t lo = e2m1_of(Ar[r * K + 2 * kb]), hi = e2m1_of(Ar[r * K + 2 * kb + 1]);
    sa[(kb >> 5) * 128 * 32 + canon_off(r, kb & 31)] = static_cast<uint8_t>(lo | (hi << 4));
  }
  for (int idx = tid; idx < nh * K / 2; idx += blockDim.x) {
    const int r = idx / (K / 2), kb = idx % (K / 2);
    const uint32_t lo = e2m1_of(Br[r * K + 2 * kb]), hi = e2m1_of(Br[r * K + 2 * kb + 1]);
    sb[(kb >> 5) * nh * 32 + canon_off(r, kb & 31)] = static_cast<uint8_t>(lo | (hi << 4));
  }
  // hand-over as in the pair-count kernel: operand writes -> proxy fence -> (TMEM stores fenced) ->
  // one arrival per warp on the LEADER's barrier, released at cluster scope
  umma::fence_proxy_async_all();
  umma::fence_before_thread_sync();
  __syncwarp();
  if (lane == 0) umma::mbar_arrive_cluster(umma::mapa(umma::smem_addr(&s_ready), 0));
  if (rank == 0 && tid == 0) {
    umma::mbar_wait_cluster_a(umma::smem_addr(&s_ready), 0);
    umma::fence_after_thread_sync();
    const uint32_t idesc = idesc_mxf4_2cta(256, N);
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t da = umma::smem_desc_kmajor_noswizzle(umma::smem_addr(sa + ks * 128 * 32), 128, 256);
      const uint64_t db = umma::smem_desc_kmajor_noswizzle(umma::smem_addr(sb + ks * nh * 32), 128, 256);
      umma::mma_mxf4_ss_2cta(tmem, da, db, idesc, ks > 0 ? 1u : 0u, tmem + P2_SF_COL, tmem + P2_SF_COL + 16);
    }
    umma::commit_2cta_a(umma::smem_addr(&s_done), 3);
  }
  umma::mbar_wait(&s_done, 0);
  umma::fence_after_thread_sync();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    umma::tmem_ld_32x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    umma::tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) C[(static_cast<size_t>(rank) * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
  }
  umma::fence_before_thread_sync();
  umma::cluster_sync();
  if (warp == 0) umma::tmem_dealloc_2cta(tmem, 512);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_fp4_2cta_rate_kernel(int N, int reps, int n_bufs, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_done;
  const int nh = N / 2;
  const uint32_t rank = umma::cluster_ctarank();
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) umma::tmem_alloc_2cta(&s_tmem, 512);
  if (tid == 0) {
    umma::mbar_init(&s_done, 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < (128 * 2 + nh * n_bufs) * 32 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x22222222u;
  umma::fence_proxy_async_all();
  umma::fence_before_thread_sync();
  umma::cluster_sync();
  umma::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;
  fill_sf(tmem, warp);
  umma::fence_before_thread_sync();
  umma::cluster_sync();
  umma::fence_after_thread_sync();
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = idesc_mxf4_2cta(256, N);
    const uint32_t sa = umma::smem_addr(smem), sb = sa + 128 * 32 * 2;
    const long long t0 = clock64();
    int buf = 0;
    for (int r = 0; r < reps; ++r) {
      const uint64_t da0 = umma::smem_desc_kmajor_noswizzle(sa, 128, 256);
      const uint64_t da1 = umma::smem_desc_kmajor_noswizzle(sa + 128 * 32, 128, 256);
      const uint64_t db0 = umma::smem_desc_kmajor_noswizzle(sb + buf * nh * 32, 128, 256);
      if (++buf == n_bufs) buf = 0;
      const uint64_t db1 = umma::smem_desc_kmajor_noswizzle(sb + buf * nh * 32, 128, 256);
      if (++buf == n_bufs) buf = 0;
      umma::mma_mxf4_ss_2cta(tmem, da0, db0, idesc, 1u, tmem + P2_SF_COL, tmem + P2_SF_COL + 16);
      umma::mma_mxf4_ss_2cta(tmem + N, da1, db1, idesc, 1u, tmem + P2_SF_COL, tmem + P2_SF_COL + 16);
    }
    umma::commit_2cta_a(umma::smem_addr(&s_done), 3);
    umma::mbar_wait(&s_done, 0);
    out_cycles[blockIdx.x >> 1] = clock64() - t0;
  } else {
    umma::mbar_wait(&s_done, 0);
  }
  umma::fence_before_thread_sync();
  umma::cluster_sync();
  if (warp == 0) umma::tmem_dealloc_2cta(tmem, 512);
}

}  // namespace

// C[256][n] (fp32) = A[256][k] * B[n][k]^T on one CTA pair; n a multiple of 16 in [32, 256], k a multiple of 64 <= 1024
PF_API int pf_selftest_umma_fp4_2cta(const int8_t* d_a, const int8_t* d_b, float* d_c, int n, int k, void* stream) {
  PF_TRY_BEGIN
  if (!d_a || !d_b || !d_c || n < 32 || n > 256 || n % 16 != 0 || k <= 0 || k % 64 != 0 || k > 1024)
    return pf_fail("pf_selftest_umma_fp4_2cta: bad arguments");
  const size_t smem = static_cast<size_t>(k / 64) * (128 + n / 2) * 32;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PF_CUDA(cudaFuncSetAttribute(umma_fp4_2cta_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  umma_fp4_2cta_selftest_kernel<<<2, 128, smem, s>>>(d_a, d_b, d_c, n, k);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

// SM clock64 ticks per 256 x n x 64 cta_group::2 instruction, two accumulators alternating, on every CTA pair of the device
PF_API int pf_probe_umma_fp4_2cta_rate(int n, int reps, int n_bufs, double* ticks_per_mma, int* h_pairs, void* stream) {
  PF_TRY_BEGIN
  if (n < 32 || n > 240 || n % 16 != 0 || reps <= 0 || n_bufs < 1 || n_bufs > 16 || !ticks_per_mma)
    return pf_fail("pf_probe_umma_fp4_2cta_rate: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 0;
  PF_CUDA(cudaGetDevice(&dev));
  PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int pairs = sms / 2;
  long long* d = nullptr;
  PF_CUDA(cudaMalloc(&d, sizeof(long long) * pairs));
  const size_t smem = static_cast<size_t>(128 * 2 + (n / 2) * n_bufs) * 32;
  cudaError_t e = cudaFuncSetAttribute(umma_fp4_2cta_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e == cudaSuccess) {
    umma_fp4_2cta_rate_kernel<<<2 * pairs, 128, smem, s>>>(n, reps, n_bufs, d);
    e = cudaGetLastError();
  }
  std::vector<long long> h(pairs);
  if (e == cudaSuccess) e = cudaMemcpy(h.data(), d, sizeof(long long) * pairs, cudaMemcpyDeviceToHost);
  cudaFree(d);
  PF_CUDA(e);
  double mean = 0;
  for (int i = 0; i < pairs; ++i) mean += double(h[i]);
  *ticks_per_mma = mean / pairs / (2.0 * reps);
  if (h_pairs) *h_pairs = pairs;
  return PF_OK;
  PF_TRY_END
}

// the same kernel timed with CUDA events around the launch: the sustained device-wide rate in TOP/s
PF_API int pf_probe_umma_fp4_2cta_peak(int n, int reps, int n_bufs, double* h_tops, double* h_ms, void* stream) {
  PF_TRY_BEGIN
  if (n < 32 || n > 240 || n % 16 != 0 || reps <= 0 || n_bufs < 1 || n_bufs > 16 || !h_tops)
    return pf_fail("pf_probe_umma_fp4_2cta_peak: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 0;
  PF_CUDA(cudaGetDevice(&dev));
  PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int pairs = sms / 2;
  long long* d = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  PF_CUDA(cudaMalloc(&d, sizeof(long long) * pairs));
  const size_t smem = static_cast<size_t>(128 * 2 + (n / 2) * n_bufs) * 32;
  float ms = 0.f;
  cudaError_t e = cudaFuncSetAttribute(umma_fp4_2cta_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) {
    umma_fp4_2cta_rate_kernel<<<2 * pairs, 128, smem, s>>>(n, 64, n_bufs, d);   // warm-up
    cudaEventRecord(e0, s);
    umma_fp4_2cta_rate_kernel<<<2 * pairs, 128, smem, s>>>(n, reps, n_bufs, d);
    cudaEventRecord(e1, s);
    e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(d);
  PF_CUDA(e);
  const double ops = 2.0 * 256.0 * n * 64.0 * 2.0 * reps * pairs;
  *h_tops = ops / (ms * 1e-3) / 1e12;
  if (h_ms) *h_ms = ms;
  return PF_OK;
  PF_TRY_END
}
