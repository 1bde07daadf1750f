// All-pairs genotype count kernel: the arithmetic the reference delegates to `treebest nj`
// (treetool/script.py:359-360) — for every sample pair, over a range of 32-site words,
//   V  += popc(valid_i & valid_j)
//   D1 += popc(valid & (code_i != code_j))
//   D2 += popc(valid & {code_i, code_j} == {0, 2})
//
// Shape of the work: a dense integer contraction over bit planes, bound by the SM integer
// pipes (POPC 16/clk/SM, LOP3 64/clk/SM measured, profiles/r01_int_pipe.json), not by HBM.
//
// Tiling: one CTA = one (64 x 64)-sample tile pair (upper triangle of the tile grid) x one
// contiguous range of word chunks (split-K; partial counts are combined with integer
// atomics, so the result is order-independent and bit-exact). The planes layout makes a
// tile's chunk of KC words one contiguous KC*512-byte block, staged into shared memory
// with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) through a STAGES-deep
// ring. 256 threads as a 16 x 16 grid, each owning a 4 x 4 block of pairs in registers;
// operand words are fetched with broadcast LDS.128.
//
// Per 32-site chunk the CTA tests whether any real sample of either tile has a missing
// genotype (code 3); chunks without missing data run a shorter body (V is then the chunk's
// site count for every pair).
#include "pf_common.cuh"

namespace {

constexpr int KC = 32;                      // words per chunk
constexpr int STAGES = 3;
constexpr int SIDE_WORDS = KC * 2 * PF_TILE;  // u32 per (stage, side)
constexpr int THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

struct PairParams {
  const uint32_t* planes;
  int64_t n_samples;
  int64_t n_words;     // words per tile in the planes buffer
  int64_t word_begin;  // range to count
  int64_t word_end;
  int32_t* counts;     // [3][n_samples][ld]
  int64_t ld;
  int n_tiles;
  int n_pairs;         // n_tiles * (n_tiles + 1) / 2
  int chunks_per_split;
  int n_chunks;
  int two;             // the constant 2, kept in a register on purpose (see csa())
};

// operand words of one 32-site word for this thread's 4 i-samples and 4 j-samples
struct Operands {
  uint32_t li[4], hi[4], lj[4], hj[4];
};
__device__ __forceinline__ Operands load_operands(const uint32_t* si, const uint32_t* sj, int w, int ty, int tx) {
  const uint4 a = *reinterpret_cast<const uint4*>(si + w * 128 + 4 * ty);
  const uint4 b = *reinterpret_cast<const uint4*>(si + w * 128 + 64 + 4 * ty);
  const uint4 c = *reinterpret_cast<const uint4*>(sj + w * 128 + 4 * tx);
  const uint4 d = *reinterpret_cast<const uint4*>(sj + w * 128 + 64 + 4 * tx);
  Operands o;
  o.li[0] = a.x; o.li[1] = a.y; o.li[2] = a.z; o.li[3] = a.w;
  o.hi[0] = b.x; o.hi[1] = b.y; o.hi[2] = b.z; o.hi[3] = b.w;
  o.lj[0] = c.x; o.lj[1] = c.y; o.lj[2] = c.z; o.lj[3] = c.w;
  o.hj[0] = d.x; o.hj[1] = d.y; o.hj[2] = d.z; o.hj[3] = d.w;
  return o;
}

// LOP3 with a fixed truth table; written as PTX so that ptxas keeps exactly these eight
// logic instructions per pair per word pair instead of re-associating them into more
__device__ __forceinline__ uint32_t lop3_xor2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, 0, 0x3c;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t lop3_xor_andn(uint32_t a, uint32_t b, uint32_t c) {  // (a ^ b) & ~c
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x14;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// carry-save step: (ones, carry) <- ones + x0 + x1 per bit; the carries weigh 2.
// `two` is the constant 2 held in a register (a kernel parameter), so the accumulate is an
// IMAD on the fma pipe and the logic pipe is left to the LOP3s.
__device__ __forceinline__ void csa(uint32_t& ones, uint32_t x0, uint32_t x1, int& acc, int two) {
  const uint32_t carry = lop3_maj(ones, x0, x1);
  ones = lop3_xor3(ones, x0, x1);
  asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc) : "r"(__popc(carry)), "r"(two));
}

// CSA = false: one POPC per stream per word (variant 1).
// CSA = true : words are taken two at a time through a carry-save adder, halving the POPCs
//              at the price of two LOP3 per stream per word pair (variant 2): the quarter-rate
//              POPC pipe and the full-rate logic pipe then carry equal load.
template <bool CSA>
__global__ void __launch_bounds__(THREADS, 2)
pair_counts_kernel(const PairParams P) {
  extern __shared__ __align__(128) uint32_t smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];

  const int tid = threadIdx.x;
  const int pair = blockIdx.x % P.n_pairs;
  const int split = blockIdx.x / P.n_pairs;
  // decode pair -> (ti <= tj) over the upper triangle, row-major
  int ti = 0, rem = pair;
  while (rem >= P.n_tiles - ti) {
    rem -= P.n_tiles - ti;
    ++ti;
  }
  const int tj = ti + rem;
  const bool diag = (ti == tj);

  const int c_begin = split * P.chunks_per_split;
  const int c_end = min(P.n_chunks, c_begin + P.chunks_per_split);
  if (c_begin >= c_end) return;

  const uint32_t* gi = P.planes + (static_cast<int64_t>(ti) * P.n_words) * (2 * PF_TILE);
  const uint32_t* gj = P.planes + (static_cast<int64_t>(tj) * P.n_words) * (2 * PF_TILE);
  const int n_sides = diag ? 1 : 2;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto chunk_words = [&](int c) -> int {
    const int64_t w0 = P.word_begin + static_cast<int64_t>(c) * KC;
    const int64_t left = P.word_end - w0;
    return left < KC ? static_cast<int>(left) : KC;
  };
  auto issue = [&](int c) {  // thread 0 only
    const int stage = (c - c_begin) % STAGES;
    const int kw = chunk_words(c);
    const uint32_t bytes = static_cast<uint32_t>(kw) * 2 * PF_TILE * 4;
    const int64_t w0 = P.word_begin + static_cast<int64_t>(c) * KC;
    uint32_t* dst = smem + stage * 2 * SIDE_WORDS;
    mbar_expect_tx(&full_bar[stage], bytes * n_sides);
    bulk_g2s(dst, gi + w0 * (2 * PF_TILE), bytes, &full_bar[stage]);
    if (!diag) bulk_g2s(dst + SIDE_WORDS, gj + w0 * (2 * PF_TILE), bytes, &full_bar[stage]);
  };

  if (tid == 0) {
    for (int c = c_begin; c < c_end && c < c_begin + STAGES; ++c) issue(c);
  }

  const int ty = tid >> 4, tx = tid & 15;
  const int real_i = static_cast<int>(min(static_cast<long long>(PF_TILE), static_cast<long long>(P.n_samples) - static_cast<long long>(ti) * PF_TILE));
  const int real_j = static_cast<int>(min(static_cast<long long>(PF_TILE), static_cast<long long>(P.n_samples) - static_cast<long long>(tj) * PF_TILE));

  // cH counts valid sites whose low code bits differ (het vs hom), cD the {0,2} sites;
  // D1 = cH + cD, D2 = cD. oH/oD are the carry-save "ones" words (CSA only).
  int cV[4][4], cH[4][4], cD[4][4];
  uint32_t oH[4][4], oD[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      cV[a][b] = cH[a][b] = cD[a][b] = 0;
      oH[a][b] = oD[a][b] = 0u;
    }
  int full_valid_sites = 0;  // sites of chunks that had no missing data

  for (int c = c_begin; c < c_end; ++c) {
    const int it = c - c_begin;
    const int stage = it % STAGES;
    const uint32_t parity = (it / STAGES) & 1;
    const int kw = chunk_words(c);
    const uint32_t* si = smem + stage * 2 * SIDE_WORDS;
    const uint32_t* sj = diag ? si : si + SIDE_WORDS;
    mbar_wait(&full_bar[stage], parity);

    // does any real sample of either tile have a missing genotype in this chunk?
    uint32_t miss = 0;
    for (int idx = tid; idx < kw * PF_TILE; idx += THREADS) {
      const int w = idx >> 6, l = idx & 63;
      const uint32_t mi = si[w * 128 + l] & si[w * 128 + 64 + l];
      const uint32_t mj = sj[w * 128 + l] & sj[w * 128 + 64 + l];
      miss |= (l < real_i ? mi : 0u) | (l < real_j ? mj : 0u);
    }
    const int any_missing = __syncthreads_or(miss != 0);

    if (any_missing) {
#pragma unroll 2
      for (int w = 0; w < kw; ++w) {
        const Operands o = load_operands(si, sj, w, ty, tx);
        uint32_t vi[4], vj[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          vi[a] = ~(o.li[a] & o.hi[a]);
          vj[a] = ~(o.lj[a] & o.hj[a]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const uint32_t W = vi[a] & vj[b];
            const uint32_t hd = (o.li[a] ^ o.lj[b]) & W;   // het vs hom among valid
            const uint32_t u = (o.hi[a] ^ o.hj[b]) & W;
            const uint32_t d2 = u & ~hd;                   // {0,2} pairs
            cV[a][b] += __popc(W);
            cH[a][b] += __popc(hd);
            cD[a][b] += __popc(d2);
          }
      }
    } else {
      if (CSA) {
        int w = 0;
        for (; w + 2 <= kw; w += 2) {
          const Operands p = load_operands(si, sj, w, ty, tx);
          const Operands q = load_operands(si, sj, w + 1, ty, tx);
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const uint32_t h0 = lop3_xor2(p.li[a], p.lj[b]);
              const uint32_t h1 = lop3_xor2(q.li[a], q.lj[b]);
              const uint32_t e0 = lop3_xor_andn(p.hi[a], p.hj[b], h0);
              const uint32_t e1 = lop3_xor_andn(q.hi[a], q.hj[b], h1);
              csa(oH[a][b], h0, h1, cH[a][b], P.two);
              csa(oD[a][b], e0, e1, cD[a][b], P.two);
            }
        }
        if (w < kw) {  // odd tail word of a short chunk
          const Operands p = load_operands(si, sj, w, ty, tx);
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const uint32_t h0 = p.li[a] ^ p.lj[b];
              const uint32_t e0 = (p.hi[a] ^ p.hj[b]) & ~h0;
              cH[a][b] += __popc(h0);
              cD[a][b] += __popc(e0);
            }
        }
      } else {
#pragma unroll 2
        for (int w = 0; w < kw; ++w) {
          const Operands o = load_operands(si, sj, w, ty, tx);
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const uint32_t hd = o.li[a] ^ o.lj[b];
              const uint32_t d2 = (o.hi[a] ^ o.hj[b]) & ~hd;
              cH[a][b] += __popc(hd);
              cD[a][b] += __popc(d2);
            }
        }
      }
      full_valid_sites += kw * 32;
    }
    __syncthreads();  // every warp is done with this stage
    if (tid == 0 && c + STAGES < c_end) issue(c + STAGES);
  }

  // epilogue: integer atomics (split-K partials and repeated calls accumulate)
  const int64_t plane = P.n_samples * P.ld;
  int32_t* V = P.counts;
  int32_t* D1 = P.counts + plane;
  int32_t* D2 = P.counts + 2 * plane;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int64_t i = static_cast<int64_t>(ti) * PF_TILE + 4 * ty + a;
    if (i >= P.n_samples) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t j = static_cast<int64_t>(tj) * PF_TILE + 4 * tx + b;
      if (j >= P.n_samples) continue;
      const int ch = cH[a][b] + (CSA ? __popc(oH[a][b]) : 0);
      const int cd = cD[a][b] + (CSA ? __popc(oD[a][b]) : 0);
      const int v = cV[a][b] + full_valid_sites;
      const int d1 = ch + cd;
      const int d2 = cd;
      atomicAdd(V + i * P.ld + j, v);
      atomicAdd(D1 + i * P.ld + j, d1);
      atomicAdd(D2 + i * P.ld + j, d2);
      if (!diag) {
        atomicAdd(V + j * P.ld + i, v);
        atomicAdd(D1 + j * P.ld + i, d1);
        atomicAdd(D2 + j * P.ld + i, d2);
      }
    }
  }
}

}  // namespace

PF_API int pf_pair_counts(const uint32_t* d_planes, int64_t n_samples, int64_t n_words,
                          int64_t word_begin, int64_t word_end, int32_t* d_counts, int64_t ld,
                          int variant, void* stream) {
  PF_TRY_BEGIN
  if (!d_planes || !d_counts || n_samples <= 0 || ld < n_samples || word_begin < 0 ||
      word_end < word_begin || word_end > n_words)
    return pf_fail("pf_pair_counts: bad arguments (n=%lld words=[%lld,%lld) of %lld ld=%lld)",
                   (long long)n_samples, (long long)word_begin, (long long)word_end, (long long)n_words,
                   (long long)ld);
  if (variant < 0 || variant > 2) return pf_fail("pf_pair_counts: unknown variant %d", variant);
  if ((word_end - word_begin) * 32 >= (int64_t(1) << 31))
    return pf_fail("pf_pair_counts: more than 2^31 sites per call overflow the int32 counts");
  if (word_end == word_begin) return PF_OK;
  if (reinterpret_cast<uintptr_t>(d_planes) % 16 != 0)
    return pf_fail("pf_pair_counts: planes must be 16-byte aligned (TMA bulk copy)");

  int dev = 0, sms = 0;
  PF_CUDA(cudaGetDevice(&dev));
  PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

  PairParams P;
  P.planes = d_planes;
  P.n_samples = n_samples;
  P.n_words = n_words;
  P.word_begin = word_begin;
  P.word_end = word_end;
  P.counts = d_counts;
  P.ld = ld;
  P.n_tiles = static_cast<int>(pf_ceil_div(n_samples, PF_TILE));
  P.n_pairs = P.n_tiles * (P.n_tiles + 1) / 2;
  P.n_chunks = static_cast<int>(pf_ceil_div(word_end - word_begin, KC));
  P.two = 2;
  // split-K so that the grid is >= ~24 waves of 2 CTAs/SM, but keep >= 4 chunks per CTA
  const int64_t target_ctas = static_cast<int64_t>(sms) * 2 * 24;
  int64_t splits = pf_ceil_div(target_ctas, P.n_pairs);
  const int64_t max_splits = pf_ceil_div(P.n_chunks, 4);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  P.chunks_per_split = static_cast<int>(pf_ceil_div(P.n_chunks, splits));
  splits = pf_ceil_div(P.n_chunks, P.chunks_per_split);
  const int64_t grid = splits * P.n_pairs;
  if (grid > 0x7fffffffLL) return pf_fail("pf_pair_counts: grid too large");

  const size_t smem_bytes = static_cast<size_t>(STAGES) * 2 * SIDE_WORDS * sizeof(uint32_t);
  auto kernel = (variant == 1) ? pair_counts_kernel<false> : pair_counts_kernel<true>;
  PF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
  pf_count_launch();
  kernel<<<static_cast<unsigned>(grid), THREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(P);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}
