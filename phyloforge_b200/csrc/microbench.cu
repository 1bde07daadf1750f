// Integer-pipe microbenchmarks for sm_100a (B200).
//
// The pair-count kernel (pair_counts.cu) is bound by the SM integer pipes, not by
// HBM, and MEASURED_PEAKS.json carries no integer peaks, so the roofline
// denominator for that kernel is measured here, on the box, by bench.py
// ("POPC_peak microbenchmarked on the box").
//
// Each kind runs ILP independent register chains of one PTX instruction (or a
// fixed mix) in `asm volatile` so ptxas can neither drop nor fuse them; one
// launch = 148*2 CTAs of 1024 threads (64 resident warps per SM, one wave).
// Reported: thread-level ops/s from CUDA events, and ops/clk/SM from clock64().

#include "pf_common.cuh"
#include "phyloforge_b200_probes.h"

namespace {

constexpr int MB_ILP = 8;
constexpr int MB_UNROLL = 16;

enum MbKind : int {
  MB_POPC = 0,       // popc.b32
  MB_LOP3 = 1,       // lop3.b32 (3-input)
  MB_IADD = 2,       // add.s32 -> IADD3
  MB_IMAD = 3,       // mad.lo.s32 (fma pipe)
  MB_POPC_LOP3 = 4,  // 1 POPC : 4 LOP3, tests whether the two share a pipe
  MB_POPC_IMAD = 5,  // 1 POPC : 4 IMAD
  MB_LOP3_IMAD = 6,  // 1 LOP3 : 1 IMAD, tests alu+fma dual issue
  MB_MIX_GENERAL = 7,  // 4 LOP3 + 3 POPC + 3 IADD : plain kernel body with missing data
  MB_MIX_NOMISS = 8,   // 2 LOP3 + 2 POPC + 2 IADD : plain kernel body without missing data
  MB_KINDS = 9
};

template <int KIND>
__global__ void __launch_bounds__(1024, 2)
mb_kernel(int iters, uint32_t seed, uint32_t* sink, long long* cycles) {
  uint32_t x[MB_ILP];
  uint32_t y[MB_ILP];
#pragma unroll
  for (int k = 0; k < MB_ILP; ++k) {
    x[k] = seed * 2654435761u + threadIdx.x * 40503u + k * 977u + blockIdx.x;
    y[k] = x[k] ^ 0x5bd1e995u;
  }
  uint32_t m = seed | 3u;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < MB_UNROLL; ++u) {
#pragma unroll
      for (int k = 0; k < MB_ILP; ++k) {
        if (KIND == MB_POPC) {
          asm volatile("popc.b32 %0, %0;" : "+r"(x[k]));
        } else if (KIND == MB_LOP3) {
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(y[k]), "r"(m));
        } else if (KIND == MB_IADD) {
          asm volatile("add.s32 %0, %0, %1;" : "+r"(x[k]) : "r"(y[k]));
        } else if (KIND == MB_IMAD) {
          asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(m), "r"(y[k]));
        } else if (KIND == MB_POPC_LOP3) {
          if ((u & 3) == 0 && (k & 1) == 0) {
            // 1/8 of slots popc, and 4/8 lop3 below: keep ratio 1:4 overall
            asm volatile("popc.b32 %0, %0;" : "+r"(y[k]));
          }
          if ((k & 1) == 1 || (u & 1) == 0) {
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(y[k ^ 1]), "r"(m));
          }
        } else if (KIND == MB_POPC_IMAD) {
          if ((u & 3) == 0 && (k & 1) == 0) {
            asm volatile("popc.b32 %0, %0;" : "+r"(y[k]));
          }
          if ((k & 1) == 1 || (u & 1) == 0) {
            asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(m), "r"(y[k ^ 1]));
          }
        } else if (KIND == MB_LOP3_IMAD) {
          if (k & 1) {
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(y[k]), "r"(m));
          } else {
            asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(m), "r"(y[k]));
          }
        }
      }
      if (KIND == MB_MIX_GENERAL) {
        // one pair-word of the plain kernel with missing data, twice (k=0..3 and 4..7):
        //   w=va&vb; d2=(ai&cj)|(ci&aj); hd=(bi^bj)&w; 3 popc; 3 add
#pragma unroll
        for (int h = 0; h < MB_ILP; h += 4) {
          uint32_t w, t, d2, hd, p0, p1, p2;
          asm volatile("and.b32 %0, %1, %2;" : "=r"(w) : "r"(x[h]), "r"(y[h]));
          asm volatile("and.b32 %0, %1, %2;" : "=r"(t) : "r"(x[h + 1]), "r"(y[h + 2]));
          asm volatile("lop3.b32 %0, %1, %2, %3, 0xf8;" : "=r"(d2) : "r"(t), "r"(x[h + 2]), "r"(y[h + 1]));
          asm volatile("lop3.b32 %0, %1, %2, %3, 0x60;" : "=r"(hd) : "r"(w), "r"(x[h + 3]), "r"(y[h + 3]));
          asm volatile("popc.b32 %0, %1;" : "=r"(p0) : "r"(w));
          asm volatile("popc.b32 %0, %1;" : "=r"(p1) : "r"(d2));
          asm volatile("popc.b32 %0, %1;" : "=r"(p2) : "r"(hd));
          asm volatile("add.s32 %0, %0, %1;" : "+r"(x[h]) : "r"(p0));
          asm volatile("add.s32 %0, %0, %1;" : "+r"(x[h + 1]) : "r"(p1));
          asm volatile("add.s32 %0, %0, %1;" : "+r"(x[h + 2]) : "r"(p2));
        }
      } else if (KIND == MB_MIX_NOMISS) {
#pragma unroll
        for (int h = 0; h < MB_ILP; h += 2) {
          uint32_t t1, d2, p0, p1;
          asm volatile("xor.b32 %0, %1, %2;" : "=r"(t1) : "r"(x[h]), "r"(y[h]));
          asm volatile("lop3.b32 %0, %1, %2, %3, 0x06;" : "=r"(d2) : "r"(t1), "r"(x[h + 1]), "r"(y[h + 1]));
          asm volatile("popc.b32 %0, %1;" : "=r"(p0) : "r"(t1));
          asm volatile("popc.b32 %0, %1;" : "=r"(p1) : "r"(d2));
          asm volatile("add.s32 %0, %0, %1;" : "+r"(x[h]) : "r"(p0));
          asm volatile("add.s32 %0, %0, %1;" : "+r"(x[h + 1]) : "r"(p1));
        }
      }
    }
  }
  long long t1c = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < MB_ILP; ++k) acc ^= x[k] ^ y[k];
  if (acc == 0x12345677u) sink[0] = acc;  // keep the chains live
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1c - t0;
}

// thread-level instructions executed per (iteration, thread) for each kind
double mb_ops_per_iter(int kind) {
  const double slots = double(MB_UNROLL) * MB_ILP;
  switch (kind) {
    case MB_POPC: case MB_LOP3: case MB_IADD: case MB_IMAD: case MB_LOP3_IMAD:
      return slots;
    case MB_POPC_LOP3: case MB_POPC_IMAD:
      // popc: u%4==0 && k even -> 4 of 16 u, 4 of 8 k = 16 ; other op: k odd (4*16) + k even && u even (4*8) = 96
      return 16.0 + 96.0;
    case MB_MIX_GENERAL:
      return double(MB_UNROLL) * 2 * 10;
    case MB_MIX_NOMISS:
      return double(MB_UNROLL) * 4 * 6;
  }
  return 0.0;
}

template <int KIND>
cudaError_t mb_launch(int grid, int iters, uint32_t* sink, long long* cyc, cudaStream_t s) {
  mb_kernel<KIND><<<grid, 1024, 0, s>>>(iters, 12345u, sink, cyc);
  return cudaGetLastError();
}

}  // namespace

// See include/phyloforge_b200.h
PF_API int pf_microbench_int_pipe(int kind, int iters, double* ops_per_s,
                                      double* ops_per_clk_per_sm, void* stream) {
  PF_TRY_BEGIN
  if (kind < 0 || kind >= MB_KINDS) return pf_fail("pf_microbench_int_pipe: bad kind %d", kind);
  if (iters <= 0) return pf_fail("pf_microbench_int_pipe: iters must be > 0");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 0;
  PF_CUDA(cudaGetDevice(&dev));
  PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = sms * 2;
  uint32_t* sink = nullptr;
  long long* cyc = nullptr;
  PF_CUDA(cudaMalloc(&sink, sizeof(uint32_t)));
  PF_CUDA(cudaMalloc(&cyc, sizeof(long long) * grid));
  cudaEvent_t e0, e1;
  PF_CUDA(cudaEventCreate(&e0));
  PF_CUDA(cudaEventCreate(&e1));
  auto launch = [&](int it) -> cudaError_t {
    switch (kind) {
      case MB_POPC: return mb_launch<MB_POPC>(grid, it, sink, cyc, s);
      case MB_LOP3: return mb_launch<MB_LOP3>(grid, it, sink, cyc, s);
      case MB_IADD: return mb_launch<MB_IADD>(grid, it, sink, cyc, s);
      case MB_IMAD: return mb_launch<MB_IMAD>(grid, it, sink, cyc, s);
      case MB_POPC_LOP3: return mb_launch<MB_POPC_LOP3>(grid, it, sink, cyc, s);
      case MB_POPC_IMAD: return mb_launch<MB_POPC_IMAD>(grid, it, sink, cyc, s);
      case MB_LOP3_IMAD: return mb_launch<MB_LOP3_IMAD>(grid, it, sink, cyc, s);
      case MB_MIX_GENERAL: return mb_launch<MB_MIX_GENERAL>(grid, it, sink, cyc, s);
      default: return mb_launch<MB_MIX_NOMISS>(grid, it, sink, cyc, s);
    }
  };
  PF_CUDA(launch(iters / 4 + 1));  // warm-up
  PF_CUDA(cudaEventRecord(e0, s));
  PF_CUDA(launch(iters));
  PF_CUDA(cudaEventRecord(e1, s));
  PF_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  PF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h(grid);
  PF_CUDA(cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  double mean_cyc = 0.0;
  for (int b = 0; b < grid; ++b) mean_cyc += double(h[b]);
  mean_cyc /= grid;
  const double per_thread = mb_ops_per_iter(kind) * iters;
  const double total = per_thread * 1024.0 * grid;
  if (ops_per_s) *ops_per_s = total / (double(ms) * 1e-3);
  // 2 CTAs of 1024 threads are co-resident on each SM for the whole launch
  if (ops_per_clk_per_sm) *ops_per_clk_per_sm = per_thread * 2048.0 / mean_cyc;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  cudaFree(cyc);
  return 0;
  PF_TRY_END
}

// ---------------------------------------------------------------------------------------
// Probe of the legacy warp-level MMA paths on sm_100a ("whether legacy
// mma.sync b1 (AND/XOR + POPC) is usefully fast on sm_100 — probe before investing").
// kind 0: mma.sync m16n8k256 b1 and.popc    kind 1: b1 xor.popc    kind 2: m16n8k32 s8
// Reports multiply-accumulate (bit-AND-popc) operations per second.
namespace {

template <int KIND>
__global__ void __launch_bounds__(256, 4) mma_probe_kernel(int iters, uint32_t seed, int* sink) {
  uint32_t a[4], b[2];
  int c[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) a[k] = seed * (k + 3) + threadIdx.x;
  b[0] = seed ^ 0x9e3779b9u;
  b[1] = seed + 77u * threadIdx.x;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int k = 0; k < 4; ++k) c[u][k] = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // 4 independent accumulator tiles
      if (KIND == 0) {
        asm volatile(
            "mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
            "{%0,%1,%2,%3};"
            : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      } else if (KIND == 1) {
        asm volatile(
            "mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.xor.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
            "{%0,%1,%2,%3};"
            : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      } else {
        asm volatile(
            "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      }
    }
  }
  int acc = 0;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc ^= c[u][k];
  if (acc == 0x7fffffff) sink[0] = acc;
}

}  // namespace

PF_API int pf_microbench_mma(int kind, int iters, double* macs_per_s, void* stream) {
  PF_TRY_BEGIN
  if (kind < 0 || kind > 2 || iters <= 0 || !macs_per_s) return pf_fail("pf_microbench_mma: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 0;
  PF_CUDA(cudaGetDevice(&dev));
  PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = sms * 8;
  int* sink = nullptr;
  PF_CUDA(cudaMalloc(&sink, sizeof(int)));
  cudaEvent_t e0, e1;
  PF_CUDA(cudaEventCreate(&e0));
  PF_CUDA(cudaEventCreate(&e1));
  auto launch = [&](int it) {
    if (kind == 0) mma_probe_kernel<0><<<grid, 256, 0, s>>>(it, 12345u, sink);
    else if (kind == 1) mma_probe_kernel<1><<<grid, 256, 0, s>>>(it, 12345u, sink);
    else mma_probe_kernel<2><<<grid, 256, 0, s>>>(it, 12345u, sink);
  };
  launch(iters / 4 + 1);
  PF_CUDA(cudaEventRecord(e0, s));
  launch(iters);
  PF_CUDA(cudaEventRecord(e1, s));
  PF_CUDA(cudaEventSynchronize(e1));
  PF_CUDA(cudaGetLastError());
  float ms = 0.f;
  PF_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  const double k = (kind == 2) ? 32.0 : 256.0;
  const double macs = double(grid) * 8.0 /*warps*/ * iters * 4.0 * 16.0 * 8.0 * k;
  *macs_per_s = macs / (double(ms) * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return PF_OK;
  PF_TRY_END
}
