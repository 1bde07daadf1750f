// Probe of the 4-bit tensor-core path for the indicator GEMMs: tcgen05.mma kind::mxf4 (E2M1 operands,
// K = 64 per instruction, fp32 accumulation, block scale factors all 1.0). The indicator values
// {-1, 0, +1} are exact in E2M1 and integer sums below 2^24 are exact in fp32, so the products the
// pair-count kernel needs can run at twice the sites per instruction of kind::i8.
// pf_selftest_umma_fp4 checks exactness against integer dot products; pf_probe_umma_fp4_rate
// measures clocks per instruction. Diagnostic only.
#include "pf_common.cuh"
#include "phyloforge_b200_probes.h"
#include "umma.cuh"

namespace {

constexpr int F4_M = 128;

// block-scaled instruction descriptor (CUTLASS cute/arch/mma_sm100_desc.hpp, InstrDescriptorBlockScaled):
// E2M1 A and B (MXF4Format 1), K-major, UE8M0 scale factors, dense K = 64
__host__ __device__ constexpr uint32_t idesc_mxf4(int m, int n, uint32_t sfa_id, uint32_t sfb_id) {
  return (sfb_id << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (1u << 23) |
         (static_cast<uint32_t>(m >> 4) << 24) | (sfa_id << 29);
}

__device__ __forceinline__ void mma_mxf4_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate, uint32_t tsfa, uint32_t tsfb) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb)
      : "memory");
}
__device__ __forceinline__ void mma_mxf4_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate, uint32_t tsfa, uint32_t tsfb) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%5], [%6], p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb)
      : "memory");
}

// E2M1 code of a small integer (0, +-1, +-2, +-3, +-4, +-6)
__device__ __forceinline__ uint32_t e2m1_of(int v) {
  const uint32_t s = v < 0 ? 8u : 0u;
  const int a = v < 0 ? -v : v;
  const uint32_t m = a == 0 ? 0u : a == 1 ? 2u : a == 2 ? 4u : a == 3 ? 5u : a == 4 ? 6u : 7u;
  return s | m;
}

// byte offset of (row r, K-byte kb of 32) inside one operand tile of a K = 64 step
__device__ __forceinline__ uint32_t canon_off(int r, int kb) {
  return static_cast<uint32_t>((r >> 3) * 256 + (kb >> 4) * 128 + (r & 7) * 16 + (kb & 15));
}

constexpr int F4_SF_COL = 448;   // TMEM columns [448, 480): scale factors, all 1.0

__device__ __forceinline__ void fill_scale_factors(uint32_t tmem, int warp) {
  uint32_t ones[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ones[j] = 0x7f7f7f7fu;   // UE8M0 127 = 2^0
  for (int c = 0; c < 32; c += 8) umma::tmem_st_32x8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + F4_SF_COL + c, ones);
  umma::tmem_wait_st();
}

template <int N>
__global__ void __launch_bounds__(128, 1)
umma_fp4_selftest_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, float* __restrict__ C, int K,
                         int a_in_tmem) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int ksteps = K / 64;
  uint8_t* sa = smem;                          // ksteps tiles of 128 x 32 B
  uint8_t* sb = smem + ksteps * F4_M * 32;     // ksteps tiles of N x 32 B
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
  if (tid == 0) {
    umma::mbar_init(&s_bar, 1);
    umma::fence_mbar_init();
  }
  // two elements per byte: even k in the low nibble
  for (int idx = tid; idx < F4_M * K / 2; idx += blockDim.x) {
    const int r = idx / (K / 2), kb = idx % (K / 2);
    const uint32_t lo = e2m1_of(A[r * K + 2 * kb]), hi = e2m1_of(A[r * K + 2 * kb + 1]);
    sa[(kb >> 5) * F4_M * 32 + canon_off(r, kb & 31)] = static_cast<uint8_t>(lo | (hi << 4));
  }
  for (int idx = tid; idx < N * K / 2; idx += blockDim.x) {
    const int r = idx / (K / 2), kb = idx % (K / 2);
    const uint32_t lo = e2m1_of(B[r * K + 2 * kb]), hi = e2m1_of(B[r * K + 2 * kb + 1]);
    sb[(kb >> 5) * N * 32 + canon_off(r, kb & 31)] = static_cast<uint8_t>(lo | (hi << 4));
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;
  fill_scale_factors(tmem, warp);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  constexpr int A_COL = 256;   // A operand staging in TMEM: 8 columns per K = 64 step (a_in_tmem, K <= 1024)
  if (a_in_tmem) {
    // thread = row; the 32 bytes (64 nibbles) of a K step in row order, low nibble = even k
    const int row = warp * 32 + lane;
    for (int ks = 0; ks < ksteps; ++ks) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t x = 0;
        for (int e = 0; e < 8; ++e) x |= e2m1_of(A[row * K + ks * 64 + 8 * j + e]) << (4 * e);
        v[j] = x;
      }
      umma::tmem_st_32x8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + A_COL + 8 * ks, v);
    }
    umma::tmem_wait_st();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
  }
  if (tid == 0) {
    const uint32_t idesc = idesc_mxf4(F4_M, N, 0, 0);
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t da = umma::smem_desc_kmajor_noswizzle(umma::smem_addr(sa + ks * F4_M * 32), 128, 256);
      const uint64_t db = umma::smem_desc_kmajor_noswizzle(umma::smem_addr(sb + ks * N * 32), 128, 256);
      if (a_in_tmem) mma_mxf4_ts(tmem, tmem + A_COL + 8 * ks, db, idesc, ks > 0 ? 1u : 0u, tmem + F4_SF_COL, tmem + F4_SF_COL + 16);
      else mma_mxf4_ss(tmem, da, db, idesc, ks > 0 ? 1u : 0u, tmem + F4_SF_COL, tmem + F4_SF_COL + 16);
    }
    umma::commit(&s_bar);
  }
  umma::mbar_wait(&s_bar, 0);
  umma::fence_after_thread_sync();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    umma::tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    umma::tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) C[row * N + c0 + j] = __uint_as_float(v[j]);
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

template <int N>
__global__ void __launch_bounds__(128, 1) umma_fp4_rate_kernel(int reps, int n_acc, int a_in_tmem, int n_bufs,
                                                               int noise, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ int s_stop;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
  if (tid == 0) {
    s_stop = 0;
    umma::mbar_init(&s_bar, 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < ((F4_M * 2 + N * n_bufs) * 32 + 4096) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x22222222u;
  umma::fence_proxy_async_smem();
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;
  fill_scale_factors(tmem, warp);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  if (tid == 0) {
    const uint32_t idesc = idesc_mxf4(F4_M, N, 0, 0);
    const uint32_t sa = umma::smem_addr(smem), sb = sa + F4_M * 32 * 2;
    const uint32_t a_col = tmem + n_acc * N;        // A staging behind the accumulators (contents irrelevant)
    const long long t0 = clock64();
    int buf = 0;
    for (int r = 0; r < reps; ++r) {
      // two products per iteration, like the pair-count kernel; the B tile changes every instruction
      const uint64_t da0 = umma::smem_desc_kmajor_noswizzle(sa, 128, 256);
      const uint64_t da1 = umma::smem_desc_kmajor_noswizzle(sa + F4_M * 32, 128, 256);
      const uint64_t db0 = umma::smem_desc_kmajor_noswizzle(sb + buf * N * 32, 128, 256);
      if (++buf == n_bufs) buf = 0;
      const uint64_t db1 = umma::smem_desc_kmajor_noswizzle(sb + buf * N * 32, 128, 256);
      if (++buf == n_bufs) buf = 0;
      const uint32_t d0 = tmem + ((2 * r) % n_acc) * N, d1 = tmem + ((2 * r + 1) % n_acc) * N;
      if (a_in_tmem) {
        mma_mxf4_ts(d0, a_col, db0, idesc, 1u, tmem + F4_SF_COL, tmem + F4_SF_COL + 16);
        mma_mxf4_ts(d1, a_col + 8, db1, idesc, 1u, tmem + F4_SF_COL, tmem + F4_SF_COL + 16);
      } else {
        mma_mxf4_ss(d0, da0, db0, idesc, 1u, tmem + F4_SF_COL, tmem + F4_SF_COL + 16);
        mma_mxf4_ss(d1, da1, db1, idesc, 1u, tmem + F4_SF_COL, tmem + F4_SF_COL + 16);
      }
    }
    umma::commit(&s_bar);
    umma::mbar_wait(&s_bar, 0);
    out_cycles[blockIdx.x] = clock64() - t0;
    *reinterpret_cast<volatile int*>(&s_stop) = 1;
  } else if (warp >= 1 && noise > 0) {
    // shared-memory traffic beside the MMAs: every `noise` clocks per warp one LDS.128 + one STS.128
    // (512 B each way per warp), on a region the MMAs do not read
    const uint32_t base = umma::smem_addr(smem) + (F4_M * 2 + N * n_bufs) * 32 + (warp - 1) * 1024 + (tid & 31) * 16;
    uint4 v = make_uint4(tid, 0, 0, 0);
    while (*reinterpret_cast<volatile int*>(&s_stop) == 0) {
      const long long t = clock64();
      const uint4 w = umma::lds128(base + 512);
      umma::sts128(base, v.x ^ w.x, v.y, v.z, v.w);
      while (clock64() - t < noise) {}
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

// C[128][n] (fp32) = A[128][k] * B[n][k]^T with entries from {0, +-1, +-2, +-3, +-4, +-6}; n = 128 or 192
PF_API int pf_selftest_umma_fp4(const int8_t* d_a, const int8_t* d_b, float* d_c, int n, int k, int a_in_tmem,
                                void* stream) {
  PF_TRY_BEGIN
  if (!d_a || !d_b || !d_c || (n != 128 && n != 192) || k <= 0 || k % 64 != 0 || k > 1024 || (a_in_tmem && n + 256 > 448 + 0 && n > 192))
    return pf_fail("pf_selftest_umma_fp4: bad arguments");
  const size_t smem = static_cast<size_t>(k / 64) * (F4_M + n) * 32;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 128) {
    PF_CUDA(cudaFuncSetAttribute(umma_fp4_selftest_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    umma_fp4_selftest_kernel<128><<<1, 128, smem, s>>>(d_a, d_b, d_c, k, a_in_tmem);
  } else {
    PF_CUDA(cudaFuncSetAttribute(umma_fp4_selftest_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    umma_fp4_selftest_kernel<192><<<1, 128, smem, s>>>(d_a, d_b, d_c, k, a_in_tmem);
  }
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

// SM-clock64 ticks per kind::mxf4 MMA (128 x n x 64), operands resident in shared memory, n_acc accumulators round-robin
PF_API int pf_probe_umma_fp4_rate(int n, int reps, int n_acc, int a_in_tmem, int n_bufs, int noise,
                                  double* ticks_per_mma, void* stream) {
  PF_TRY_BEGIN
  if ((n != 128 && n != 192 && n != 256) || reps <= 0 || n_acc < 1 || n_acc * n + (a_in_tmem ? 16 : 0) > F4_SF_COL ||
      n_bufs < 1 || n_bufs > 16 || !ticks_per_mma)
    return pf_fail("pf_probe_umma_fp4_rate: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 0;
  PF_CUDA(cudaGetDevice(&dev));
  PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long* d = nullptr;
  PF_CUDA(cudaMalloc(&d, sizeof(long long) * sms));
  const size_t smem = static_cast<size_t>(F4_M * 2 + n * n_bufs) * 32 + 4096;
  auto launch = [&](auto kernel) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kernel<<<sms, 128, smem, s>>>(reps, n_acc, a_in_tmem, n_bufs, noise, d);
    return cudaGetLastError();
  };
  cudaError_t e = n == 128 ? launch(umma_fp4_rate_kernel<128>) : n == 192 ? launch(umma_fp4_rate_kernel<192>) : launch(umma_fp4_rate_kernel<256>);
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
