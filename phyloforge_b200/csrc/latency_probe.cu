// Latency probes for the dependent-chain parts of the NJ kernel (DESIGN.md "NJ"): what one L2 round
// trip, one gpu-scope fence, one CTA barrier, one tagged-word exchange between CTAs and one cooperative
// grid barrier cost on this GPU. Diagnostic only (tools/probe_latency.py).
#include "pf_common.cuh"
#include "phyloforge_b200_probes.h"
#include "umma.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

__device__ __forceinline__ unsigned long long probe_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// kind 0: dependent ld.global.cg chain over `buf` (a random cycle of indices), one thread
// kind 1: st + fence.acq_rel.gpu per step, one thread
// kind 2: __syncthreads per step, blockDim threads
// kind 3: tagged-word exchange between all CTAs of the grid (every CTA publishes, warp 0 polls all)
// kind 4: cooperative grid.sync per step
// kind 5: as 0, but the chain is first written by another CTA (lines last written remotely)
// kind 6: 16 independent ld.global.cg per step (one batch), one warp
// kind 7: tcgen05.commit (no MMA outstanding) -> mbarrier wait round trip, one thread
// kind 8: cp.async.bulk of 6 KB global -> shared, issue to completion (mbarrier wait), one thread
// kind 9: cp.async.bulk of 6 KB: 7 copies in flight per wait
// kind 10: mbarrier arrive by thread 32 -> try_wait wake-up of thread 0 and back (ping-pong), per hop
__global__ void __launch_bounds__(512, 1)
probe_kernel(int kind, int iters, unsigned* buf, int buf_len, unsigned long long* xchg, unsigned long long* out_ns,
             unsigned* sink) {
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_cta = static_cast<int>(gridDim.x);
  if (kind == 5) {
    // CTA 1 rewrites the chain, CTA 0 chases it afterwards
    if (blockIdx.x == 1 % n_cta)
      for (int k = tid; k < buf_len; k += blockDim.x) buf[k] = buf[k] + 0;
    __threadfence();
  }
  if (kind == 3)
    for (int k = blockIdx.x * blockDim.x + tid; k < 2 * n_cta; k += n_cta * blockDim.x) xchg[k] = 0ull;
  grid.sync();
  unsigned long long t0 = probe_now();
  unsigned acc = 0;
  if (kind == 0 || kind == 5) {
    if (blockIdx.x == 0 && tid == 0) {
      unsigned p = 0;
      for (int i = 0; i < iters; ++i) p = __ldcg(buf + p);
      acc = p;
    }
  } else if (kind == 1) {
    if (blockIdx.x == 0 && tid == 0) {
      for (int i = 0; i < iters; ++i) {
        buf[i & 1023] = i;
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
    }
  } else if (kind == 2) {
    if (blockIdx.x == 0)
      for (int i = 0; i < iters; ++i) __syncthreads();
  } else if (kind == 3) {
    for (int i = 0; i < iters; ++i) {
      const unsigned long long tag = static_cast<unsigned long long>(i + 1) << 32;
      unsigned long long* base = xchg + static_cast<int64_t>(i & 1) * n_cta;
      if (warp == 0) {
        if (lane == 0) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(base + blockIdx.x), "l"(tag | blockIdx.x) : "memory");
        }
        for (int c0 = 0; c0 < n_cta; c0 += 32) {
          const int c = c0 + lane;
          if (c < n_cta) {
            unsigned long long w;
            do {
              asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(base + c) : "memory");
            } while ((w >> 32) != static_cast<unsigned long long>(i + 1));
            acc += static_cast<unsigned>(w);
          }
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
      __syncthreads();
    }
  } else if (kind == 4) {
    for (int i = 0; i < iters; ++i) grid.sync();
  } else if (kind == 6) {
    if (blockIdx.x == 0 && warp == 0) {
      unsigned p = lane;
      for (int i = 0; i < iters; ++i) {
        unsigned v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldcg(buf + ((p + u * 97) & (buf_len - 1)));
        unsigned s = 0;
#pragma unroll
        for (int u = 0; u < 16; ++u) s += v[u];
        p = s & (buf_len - 1);
      }
      acc = p;
    }
  }
  if (kind >= 7 && kind <= 10 && blockIdx.x == 0) {
    __shared__ __align__(1024) unsigned char s_buf[7 * 6144];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_tmem;
    if (kind == 7 && warp == 0) umma::tmem_alloc(&s_tmem, 32);
    if (tid == 0) {
      umma::mbar_init(&s_bar[0], 1);
      umma::mbar_init(&s_bar[1], 1);
      umma::fence_mbar_init();
    }
    __syncthreads();
    t0 = probe_now();
    if (kind == 7 && tid == 0) {
      for (int i = 0; i < iters; ++i) {
        umma::commit(&s_bar[0]);
        umma::mbar_wait(&s_bar[0], i & 1);
      }
    } else if (kind == 8 && tid == 0) {
      for (int i = 0; i < iters; ++i) {
        umma::mbar_expect_tx(&s_bar[0], 6144);
        umma::bulk_g2s(s_buf, reinterpret_cast<const unsigned char*>(buf) + (i & 7) * 6144, 6144, &s_bar[0]);
        umma::mbar_wait(&s_bar[0], i & 1);
      }
    } else if (kind == 9 && tid == 0) {
      for (int i = 0, k = 0; i < iters; i += 7, ++k) {
        umma::mbar_expect_tx(&s_bar[0], 7 * 6144);
        for (int j = 0; j < 7; ++j)
          umma::bulk_g2s(s_buf + j * 6144, reinterpret_cast<const unsigned char*>(buf) + j * 6144, 6144, &s_bar[0]);
        umma::mbar_wait(&s_bar[0], k & 1);
      }
    } else if (kind == 10 && (tid == 0 || tid == 32)) {
      for (int i = 0; i < iters; ++i) {
        if (tid == 0) {
          umma::mbar_arrive(&s_bar[0]);
          umma::mbar_wait(&s_bar[1], i & 1);
        } else {
          umma::mbar_wait(&s_bar[0], i & 1);
          umma::mbar_arrive(&s_bar[1]);
        }
      }
    }
    __syncthreads();
    if (kind == 7 && warp == 0) umma::tmem_dealloc(s_tmem, 32);
  }
  unsigned long long t1 = probe_now();
  if (blockIdx.x == 0 && tid == 0) *out_ns = (kind == 10) ? (t1 - t0) / 2 : (t1 - t0);
  if (acc == 0xdeadbeefu) *sink = acc;
}

}  // namespace

// See include/phyloforge_b200.h
PF_API int pf_probe_latency(int kind, int iters, int n_ctas, double* h_ns_per_step, void* stream) {
  PF_TRY_BEGIN
  if (kind < 0 || kind > 10 || iters <= 0 || n_ctas <= 0 || !h_ns_per_step)
    return pf_fail("pf_probe_latency: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 0;
  PF_CUDA(cudaGetDevice(&dev));
  PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (n_ctas > sms) n_ctas = sms;
  int buf_len = 1 << 16;   // 256 KB: L2-resident, larger than nothing else matters
  unsigned* buf = nullptr;
  unsigned long long *xchg = nullptr, *out = nullptr;
  unsigned* sink = nullptr;
  PF_CUDA(cudaMalloc(&buf, sizeof(unsigned) * buf_len));
  PF_CUDA(cudaMalloc(&xchg, sizeof(unsigned long long) * 2 * sms));
  PF_CUDA(cudaMalloc(&out, sizeof(unsigned long long)));
  PF_CUDA(cudaMalloc(&sink, sizeof(unsigned)));
  {
    // one cycle through all entries with a large odd stride (every step lands in another line)
    unsigned* h = static_cast<unsigned*>(malloc(sizeof(unsigned) * buf_len));
    if (!h) return pf_fail("pf_probe_latency: out of host memory");
    for (int i = 0; i < buf_len; ++i) h[i] = static_cast<unsigned>((i + 40503) & (buf_len - 1));
    cudaError_t e = cudaMemcpyAsync(buf, h, sizeof(unsigned) * buf_len, cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);
    free(h);
    PF_CUDA(e);
  }
  void* args[] = {&kind, &iters, &buf, &buf_len, &xchg, &out, &sink};
  for (int rep = 0; rep < 2; ++rep)   // the first launch warms up
    PF_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(probe_kernel), dim3(n_ctas), dim3(512), args, 0, s));
  unsigned long long ns = 0;
  PF_CUDA(cudaMemcpyAsync(&ns, out, sizeof(ns), cudaMemcpyDeviceToHost, s));
  PF_CUDA(cudaStreamSynchronize(s));
  *h_ns_per_step = static_cast<double>(ns) / iters;
  cudaFree(buf);
  cudaFree(xchg);
  cudaFree(out);
  cudaFree(sink);
  return PF_OK;
  PF_TRY_END
}
