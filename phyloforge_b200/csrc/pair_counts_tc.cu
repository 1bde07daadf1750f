// Tensor-core variant of the all-pairs genotype counts: an int8 indicator GEMM on tcgen05
// (BASELINE.json north star: "an int8 tcgen05 indicator-GEMM variant kept only if ncu shows it
// beats popcount"). Same arithmetic as pair_counts.cu, restated as two integer matrix products:
//
//   x = g - 1 in {-1, 0, +1} (0 for het and missing),  h = 1 for a homozygous genotype else 0
//   P = sum_s x_i x_j = n00 + n22 - n02          Q = sum_s h_i h_j = n00 + n22 + n02
//   D2 = n02 = (Q - P) / 2        het-vs-hom sites = |h_i| + |h_j| - 2 Q        D1 = that + D2
//
// valid when every site is genotyped in all samples or in none (the no-missing-data case, where V
// is the same number for every pair); pf_pair_counts_tc reports otherwise and the caller uses the
// popcount kernel. Exact: int8 x int8 products accumulate in int32 in TMEM.
//
// Data flow per CTA (one 128 x 192 sample tile pair, a contiguous range of 32-site words):
//   "nib4" operand stream (4 bits per genotype, produced once per call by planes_to_nib4):
//   global --TMA bulk copies--> raw ring --10 expansion warps: PRMT nibble -> int8 (x and h)-->
//   A rows straight into TMEM (tcgen05.st), B rows into canonical K-major UMMA tiles in shared memory
//   --tcgen05.mma kind::i8, A from TMEM, B from shared memory (1 thread)--> 2 x 192 TMEM accumulator
//   columns --tcgen05.ld--> epilogue --> int32 RED atomics (split-K, bit-exact).
// The kernel is bound by shared-memory bandwidth (ncu: LSU wavefronts + tensor operand reads), which
// is why the A operand bypasses shared memory. (Reading the operand stream straight from global memory
// with register prefetch instead of the TMA ring was measured 15% slower; kept as a TMA ring.)
#include "pf_common.cuh"
#include "umma.cuh"

namespace {

constexpr int TC_I = 128;            // rows of the accumulator tile (MMA M); the A operand lives in TMEM
constexpr int TC_J = 192;            // columns (MMA N); the B operand is expanded into shared memory
constexpr int TC_ROWS = TC_I + TC_J; // sample rows expanded per word
constexpr int TC_PF = TC_ROWS / PF_TILE;  // 64-sample tiles of the operand stream per step: 2 (I) + 3 (J)
constexpr int STEP_WORDS = 2;        // words per pipeline step (one raw stage feeds one expanded stage)
constexpr int RAW_WORDS = STEP_WORDS;
constexpr int RAW_STAGES = 6;
constexpr int EXP_STAGES = 4;
constexpr int RAW_STAGE_BYTES = RAW_WORDS * TC_ROWS * 16;      // 10 KB
constexpr int EXP_WORD_BYTES = TC_J * 32 * 2;                  // x and h tiles of B for one word: 12 KB
constexpr int EXP_STAGE_BYTES = STEP_WORDS * EXP_WORD_BYTES;   // 24 KB
constexpr int EXP_BX = 0, EXP_BH = TC_J * 32;
// TMEM columns: P accumulators [0, 192), Q accumulators [192, 384), A operand staging [384, 512):
// per in-flight word 8 columns of x and 8 of h (32 K-bytes per row = 8 x 32-bit columns)
constexpr int TMEM_P = 0, TMEM_Q = TC_J, TMEM_A = 2 * TC_J;
constexpr int TC_A_WARPS = TC_I / 32, TC_B_WARPS = TC_J / 32;
constexpr int TC_EXP_WARPS = TC_A_WARPS + TC_B_WARPS;          // 10: one sample row per thread
constexpr int TC_THREADS = (TC_EXP_WARPS + 2) * 32;
constexpr uint32_t POOL_X = 0x000100ffu;  // code 0 -> -1, 1 -> 0, 2 -> +1, 3 -> 0
constexpr uint32_t POOL_H = 0x00010001u;  // code 0 -> 1, 1 -> 0, 2 -> 1, 3 -> 0
constexpr int HSUM_RUN = 16;         // words per thread of the converter; split boundaries are multiples of it

// ------------------------------------------------------------------ planes -> nib4
__device__ __forceinline__ uint32_t spread8(uint32_t b) {  // 8 bits -> 8 nibbles
  uint32_t x = b;
  x = (x | (x << 12)) & 0x000f000fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x;
}

// nib4[((T * n_local + w) * 64 + l) * 4 + q]: nibble j of word q = code of site 32 w + 8 q + j.
// ref_v[w] = valid-site mask of sample 0; flags[0] is set when any real sample's mask differs from
// it, i.e. when some site is genotyped in some samples and missing in others.
// hsum[split][sample] = homozygous genotypes of the sample in the split's word range (|h_i|).
// One thread = one sample x a run of HSUM_RUN consecutive words (runs never straddle a split).
__global__ void __launch_bounds__(256)
planes_to_nib4_kernel(const uint32_t* __restrict__ planes, int64_t n_words, int64_t word_begin, int64_t n_local,
                      int64_t n_samples, int64_t n_tiles, int64_t n_tiles_padded, uint4* __restrict__ nib4,
                      uint32_t* __restrict__ ref_v, int* __restrict__ flags, int* __restrict__ hsum,
                      int64_t words_per_split, int64_t hsum_ld) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int l = static_cast<int>(idx & 63);
  const int64_t tr = idx >> 6;
  const int64_t n_runs = (n_local + HSUM_RUN - 1) / HSUM_RUN;
  const int64_t run = tr % n_runs;
  const int64_t t = tr / n_runs;
  if (t >= n_tiles_padded) return;
  const int64_t w0 = run * HSUM_RUN;
  const int64_t w1 = min(n_local, w0 + HSUM_RUN);
  const bool real = (t * PF_TILE + l) < n_samples;
  int hom = 0;
  bool mixed = false;
  for (int64_t w = w0; w < w1; ++w) {
    uint32_t L = 0xffffffffu, H = 0xffffffffu;
    if (t < n_tiles) {
      const uint32_t* p = planes + ((t * n_words + word_begin + w) * 2) * PF_TILE;
      L = p[l];
      H = p[PF_TILE + l];
    }
    uint4 out;
    out.x = spread8(L & 0xffu) | (spread8(H & 0xffu) << 1);
    out.y = spread8((L >> 8) & 0xffu) | (spread8((H >> 8) & 0xffu) << 1);
    out.z = spread8((L >> 16) & 0xffu) | (spread8((H >> 16) & 0xffu) << 1);
    out.w = spread8(L >> 24) | (spread8(H >> 24) << 1);
    nib4[(t * n_local + w) * 64 + l] = out;
    const uint32_t* p0 = planes + ((word_begin + w) * 2) * PF_TILE;   // sample 0: tile 0, lane 0
    const uint32_t ref = ~(p0[0] & p0[PF_TILE]);
    if (t == 0 && l == 0) ref_v[w] = ref;
    mixed |= (~(L & H) != ref);
    hom += __popc(~L);            // homozygous <=> low code bit 0 (missing has it set)
  }
  if (real) {
    if (mixed) atomicOr(flags, 1);
    if (hom) atomicAdd(hsum + (w0 / words_per_split) * hsum_ld + t * PF_TILE + l, hom);
  }
}

// ------------------------------------------------------------------ the GEMM kernel
// optional role-level cycle accounting (pf_pair_counts_tc_profile): who waits for whom
__device__ long long g_tc_prof[16];
bool g_tc_prof_enabled = false;

struct TcParams {
  const uint4* nib4;      // [n_tiles_padded][n_local][64] uint4
  const uint32_t* all_v;  // [n_local] valid-site mask shared by all samples
  int64_t n_local;        // words in the range
  int64_t n_samples;
  int32_t* counts;        // [3][n_samples][ld]
  int64_t ld;
  int n_itiles;           // ceil(n_samples / 128)
  int n_jtiles;           // ceil(n_samples / 256)
  int n_tilepairs;
  int words_per_split;
  const int* flags;       // flags[0] != 0: validity is not uniform -> results must not be used
  const int* hsum;        // [splits][hsum_ld] homozygous counts per sample and split
  int64_t hsum_ld;
  int prof;               // accumulate wait cycles per role into g_tc_prof
};

// PRMT as the hardware does it (the __byte_perm intrinsic masks the selector first): byte k of the
// result = byte (nibble k of the low 16 selector bits) of {pool, pool}. Codes are 0..3, so the
// sign-replicate bit of a selector nibble is never set.
__device__ __forceinline__ uint32_t prmt(uint32_t pool, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(r) : "r"(pool), "r"(sel));
  return r;
}

// one sample row x one 32-site word: 16 bytes of nibble codes -> 32 int8 x-values and 32 int8 h-values
__device__ __forceinline__ void expand_word(const uint4& nb, uint32_t (&xs)[8], uint32_t (&hs)[8]) {
  const uint32_t q[4] = {nb.x, nb.y, nb.z, nb.w};
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const uint32_t hi = q[m] >> 16;
    xs[2 * m] = prmt(POOL_X, q[m]);
    xs[2 * m + 1] = prmt(POOL_X, hi);
    hs[2 * m] = prmt(POOL_H, q[m]);
    hs[2 * m + 1] = prmt(POOL_H, hi);
  }
}

// byte offset of the 16-byte unit (row r, K half kh) inside a canonical K-major no-swizzle tile
__device__ __forceinline__ uint32_t unit_off(int r, int kh) {
  return static_cast<uint32_t>((r >> 3) * 256 + kh * 128 + (r & 7) * 16);
}

template <bool PROF>
__global__ void __launch_bounds__(TC_THREADS, 1)
pair_counts_tc_kernel(const TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t raw_full[RAW_STAGES], raw_empty[RAW_STAGES];
  __shared__ __align__(8) uint64_t exp_full[EXP_STAGES], exp_empty[EXP_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t s_tmem;
  __shared__ int s_hcount[TC_ROWS];
  __shared__ int s_valid_sites;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // shared-window addresses, computed once
  const uint32_t raw_a = umma::smem_addr(smem);                                   // RAW_STAGES x RAW_STAGE_BYTES
  const uint32_t exp_a = raw_a + RAW_STAGES * RAW_STAGE_BYTES;                    // EXP_STAGES x EXP_STAGE_BYTES (B tiles)
  const uint32_t raw_full_a = umma::smem_addr(raw_full), raw_empty_a = umma::smem_addr(raw_empty);
  const uint32_t exp_full_a = umma::smem_addr(exp_full), exp_empty_a = umma::smem_addr(exp_empty);
  const uint32_t done_a = umma::smem_addr(&done_bar);

  // work item: (tile pair, split). I-tile ti (128 rows) pairs with the J-tiles (192 columns) that
  // reach its first row: tj >= floor(128 ti / 192)
  const int pair = blockIdx.x % P.n_tilepairs;
  const int split = blockIdx.x / P.n_tilepairs;
  int ti = 0, rem = pair;
  while (true) {
    const int cnt = P.n_jtiles - (2 * ti) / 3;
    if (rem < cnt) break;
    rem -= cnt;
    ++ti;
  }
  const int tj = (2 * ti) / 3 + rem;
  const int64_t w_begin = static_cast<int64_t>(split) * P.words_per_split;
  const int64_t w_end = min(P.n_local, w_begin + P.words_per_split);
  const int n_wordsteps = static_cast<int>(w_end - w_begin);
  if (n_wordsteps <= 0) return;
  const int n_steps = (n_wordsteps + STEP_WORDS - 1) / STEP_WORDS;

  if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
  if (tid == 32) {
    for (int s = 0; s < RAW_STAGES; ++s) {
      umma::mbar_init(&raw_full[s], 1);
      umma::mbar_init(&raw_empty[s], TC_EXP_WARPS);
    }
    for (int e = 0; e < EXP_STAGES; ++e) {
      umma::mbar_init(&exp_full[e], TC_EXP_WARPS);
      umma::mbar_init(&exp_empty[e], 1);
    }
    umma::mbar_init(&done_bar, 1);
    umma::fence_mbar_init();
  }
  if (tid == 0) s_valid_sites = 0;
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;

  if (warp == TC_EXP_WARPS) {
    // ===== TMA producer: raw nib4 words of the 2 + 3 sample tiles of this tile pair =====
    if (lane == 0) {
      long long t_wait = 0;
      const long long t_start = PROF ? clock64() : 0;
      const uint4* src_tile[TC_PF];
#pragma unroll
      for (int k = 0; k < TC_PF; ++k) {
        const int64_t pft = (k < 2) ? (2 * static_cast<int64_t>(ti) + k) : (3 * static_cast<int64_t>(tj) + (k - 2));
        src_tile[k] = P.nib4 + (pft * P.n_local + w_begin) * 64;
      }
      int stage = 0;
      uint32_t phase = 1;        // parity of the "previous use" of a stage: first pass must not block
      for (int st = 0; st < n_steps; ++st) {
        if (st >= RAW_STAGES) {
          const long long t0 = PROF ? clock64() : 0;
          umma::mbar_wait_a(raw_empty_a + stage * 8, phase);
          if (PROF) t_wait += clock64() - t0;
        }
        const int kw = min(RAW_WORDS, n_wordsteps - st * RAW_WORDS);
        const uint32_t bytes_per_tile = static_cast<uint32_t>(kw) * 64 * 16;
        umma::mbar_expect_tx_a(raw_full_a + stage * 8, bytes_per_tile * TC_PF);
        const uint32_t dst = raw_a + stage * RAW_STAGE_BYTES;
        // stage layout: [pf_tile 0..4][word][64 samples][16 B]; pf tiles 0,1 = I rows, 2..4 = J rows
#pragma unroll
        for (int k = 0; k < TC_PF; ++k)
          umma::bulk_g2s_a(dst + k * (RAW_WORDS * 64 * 16), src_tile[k] + static_cast<int64_t>(st) * (RAW_WORDS * 64),
                           bytes_per_tile, raw_full_a + stage * 8);
        if (++stage == RAW_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (PROF) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[5]), static_cast<unsigned long long>(t_wait));
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[6]), static_cast<unsigned long long>(clock64() - t_start));
      }
    }
  } else if (warp == TC_EXP_WARPS + 1) {
    // ===== MMA issuer: D[tmem] += A[tmem] * B[smem], two products per word =====
    if (lane == 0) {
      const uint32_t idesc = umma::idesc_i8(TC_I, TC_J);
      long long t_wait = 0;
      const long long t_start = PROF ? clock64() : 0;
      int e = 0;
      uint32_t phase = 0;
      for (int st = 0; st < n_steps; ++st) {
        const long long t0 = PROF ? clock64() : 0;
        umma::mbar_wait_a(exp_full_a + e * 8, phase);
        if (PROF) t_wait += clock64() - t0;
        umma::fence_after_thread_sync();
        const int kw = min(STEP_WORDS, n_wordsteps - st * STEP_WORDS);
#pragma unroll
        for (int k = 0; k < STEP_WORDS; ++k) {
          if (k < kw) {
            const uint32_t base = exp_a + e * EXP_STAGE_BYTES + k * EXP_WORD_BYTES;
            const uint64_t bx = umma::smem_desc_kmajor_noswizzle(base + EXP_BX, 128, 256);
            const uint64_t bh = umma::smem_desc_kmajor_noswizzle(base + EXP_BH, 128, 256);
            const uint32_t a_col = tmem + TMEM_A + (e * STEP_WORDS + k) * 16;
            const uint32_t acc = (st > 0 || k > 0) ? 1u : 0u;
            umma::mma_i8_ts(tmem + TMEM_P, a_col, bx, idesc, acc);
            umma::mma_i8_ts(tmem + TMEM_Q, a_col + 8, bh, idesc, acc);
          }
        }
        umma::commit_a(exp_empty_a + e * 8);
        if (++e == EXP_STAGES) {
          e = 0;
          phase ^= 1;
        }
      }
      umma::commit_a(done_a);
      if (PROF) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[0]), static_cast<unsigned long long>(t_wait));
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[1]), static_cast<unsigned long long>(clock64() - t_start));
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[8]), static_cast<unsigned long long>(n_steps));
      }
    }
  } else {
    // ===== expansion warps: thread = one sample row. Warps 0-3: the 128 A rows, written straight
    // into TMEM (their lane quadrant); warps 4-9: the 192 B rows, written to shared memory in the
    // canonical K-major layout =====
    const int row = tid;                       // 0 .. 319
    const int pf = row >> 6, l = row & 63;     // raw stage sub-tile and sample inside it
    const bool is_a = row < TC_I;
    const int r = is_a ? row : row - TC_I;
    const uint32_t a_base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + TMEM_A;   // A rows: TMEM lane = row
    const uint32_t raw_row = raw_a + pf * (RAW_WORDS * 64 * 16) + l * 16;
    const uint32_t exp_row = exp_a + unit_off(r, 0);                                          // B rows: canonical smem tile
    long long t_wraw = 0, t_wemp = 0, t_fence = 0;
    const long long t_start = PROF ? clock64() : 0;
    int stage = 0, e = 0;
    uint32_t raw_phase = 0, exp_phase = 1;
    for (int st = 0; st < n_steps; ++st) {
      if (lane == 0) {
        const long long t0 = PROF ? clock64() : 0;
        umma::mbar_wait_a(raw_full_a + stage * 8, raw_phase);
        const long long t1 = PROF ? clock64() : 0;
        if (st >= EXP_STAGES) umma::mbar_wait_a(exp_empty_a + e * 8, exp_phase);
        if (PROF) {
          t_wraw += t1 - t0;
          t_wemp += clock64() - t1;
        }
      }
      __syncwarp();
      if (is_a) umma::fence_after_thread_sync();   // the MMAs that read these TMEM columns have completed
      const int kw = min(STEP_WORDS, n_wordsteps - st * STEP_WORDS);
      const uint32_t src = raw_row + stage * RAW_STAGE_BYTES;
      uint4 nb[STEP_WORDS];
#pragma unroll
      for (int k = 0; k < STEP_WORDS; ++k) nb[k] = umma::lds128(src + k * (64 * 16));   // stale words of a short last step are unused
      if (is_a) {
        const uint32_t a_col = a_base + e * (STEP_WORDS * 16);
#pragma unroll
        for (int k = 0; k < STEP_WORDS; ++k) {
          if (k < kw) {
            uint32_t xs[8], hs[8];
            expand_word(nb[k], xs, hs);
            umma::tmem_st_32x8(a_col + k * 16, xs);
            umma::tmem_st_32x8(a_col + k * 16 + 8, hs);
          }
        }
      } else {
        const uint32_t dst = exp_row + e * EXP_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < STEP_WORDS; ++k) {
          if (k < kw) {
            uint32_t xs[8], hs[8];
            expand_word(nb[k], xs, hs);
            const uint32_t d = dst + k * EXP_WORD_BYTES;
            umma::sts128(d + EXP_BX, xs[0], xs[1], xs[2], xs[3]);
            umma::sts128(d + EXP_BX + 128, xs[4], xs[5], xs[6], xs[7]);
            umma::sts128(d + EXP_BH, hs[0], hs[1], hs[2], hs[3]);
            umma::sts128(d + EXP_BH + 128, hs[4], hs[5], hs[6], hs[7]);
          }
        }
      }
      const long long tf0 = PROF ? clock64() : 0;
      if (is_a) {
        umma::tmem_wait_st();
        umma::fence_before_thread_sync();
      } else {
        umma::fence_proxy_async_smem();
      }
      __syncwarp();
      if (PROF) t_fence += clock64() - tf0;
      if (lane == 0) {
        umma::mbar_arrive_a(exp_full_a + e * 8);
        umma::mbar_arrive_a(raw_empty_a + stage * 8);
      }
      if (++stage == RAW_STAGES) {
        stage = 0;
        raw_phase ^= 1;
      }
      if (++e == EXP_STAGES) {
        e = 0;
        exp_phase ^= 1;
      }
    }
    if (PROF && lane == 0 && (warp == 0 || warp == 4)) {
      const int o = warp == 0 ? 2 : 9;   // [2..4] an A-row warp, [9..11] a B-row warp
      atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[o]), static_cast<unsigned long long>(t_wraw));
      atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[o + 1]), static_cast<unsigned long long>(t_wemp));
      atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[o + 2]), static_cast<unsigned long long>(clock64() - t_start));
      atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[warp == 0 ? 12 : 13]), static_cast<unsigned long long>(t_fence));
    }
    {
      const int64_t sample = is_a ? (static_cast<int64_t>(ti) * TC_I + r) : (static_cast<int64_t>(tj) * TC_J + r);
      s_hcount[row] = (sample < P.n_samples) ? P.hsum[static_cast<int64_t>(split) * P.hsum_ld + sample] : 0;
    }
    // valid (genotyped-in-all-samples) sites of this word range: V for every pair
    int vs = 0;
    for (int64_t w = w_begin + tid; w < w_end; w += TC_EXP_WARPS * 32) vs += __popc(P.all_v[w]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, off);
    if (lane == 0 && vs) atomicAdd(&s_valid_sites, vs);
    // all expansion warps meet before the epilogue (named barrier 1)
    asm volatile("bar.sync 1, %0;" ::"r"(TC_EXP_WARPS * 32) : "memory");

    // ===== epilogue: TMEM -> registers -> counts =====
    umma::mbar_wait_a(done_a, 0);
    umma::fence_after_thread_sync();
    const int lane_grp = warp & 3;               // TMEM lanes [32 g, 32 g + 32) are visible to warps with w % 4 == g
    const int grp_rank = warp >> 2;              // rank of this warp among the warps of its lane group
    const int grp_size = (lane_grp < (TC_EXP_WARPS & 3)) ? (TC_EXP_WARPS / 4 + 1) : (TC_EXP_WARPS / 4);
    const int i_local = lane_grp * 32 + lane;
    const int64_t i = static_cast<int64_t>(ti) * TC_I + i_local;
    const int hi = s_hcount[i_local];
    const int v_sites = s_valid_sites;
    const int64_t plane = P.n_samples * P.ld;
    int32_t* V = P.counts;
    int32_t* D1 = P.counts + plane;
    int32_t* D2 = P.counts + 2 * plane;
    for (int chunk = grp_rank; chunk < TC_J / 32; chunk += grp_size) {
      uint32_t pv[32], qv[32];
      const uint32_t taddr = tmem + (static_cast<uint32_t>(lane_grp * 32) << 16) + chunk * 32;
      umma::tmem_ld_32x32(taddr + TMEM_P, pv);
      umma::tmem_ld_32x32(taddr + TMEM_Q, qv);
      umma::tmem_wait_ld();
      if (i < P.n_samples) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int j_local = chunk * 32 + c;
          const int64_t j = static_cast<int64_t>(tj) * TC_J + j_local;
          if (j >= P.n_samples || j < i) continue;       // each unordered pair once: row <= column
          const int pp = static_cast<int>(pv[c]), qq = static_cast<int>(qv[c]);
          const int d2 = (qq - pp) >> 1;
          const int d1 = hi + s_hcount[TC_I + j_local] - 2 * qq + d2;
          atomicAdd(V + i * P.ld + j, v_sites);
          atomicAdd(D1 + i * P.ld + j, d1);
          atomicAdd(D2 + i * P.ld + j, d2);
          if (j != i) {
            atomicAdd(V + j * P.ld + i, v_sites);
            atomicAdd(D1 + j * P.ld + i, d1);
            atomicAdd(D2 + j * P.ld + i, d2);
          }
        }
      }
    }
    umma::fence_before_thread_sync();
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

constexpr int64_t tc_align(int64_t x) { return (x + 255) / 256 * 256; }

struct TcPlan {
  int64_t n_local, n_tiles, n_tiles_padded;
  int n_itiles, n_jtiles, pairs;
  int64_t splits, words_per_split, hsum_ld;
  int64_t nib_bytes, refv_off, flags_off, hsum_off, total;
};

int tc_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      sms <= 0)
    sms = 148;
  return sms;
}

TcPlan tc_plan(int64_t n_samples, int64_t n_local, int sms) {
  TcPlan L;
  L.n_local = n_local;
  L.n_tiles = pf_ceil_div(n_samples, PF_TILE);
  L.n_jtiles = static_cast<int>(pf_ceil_div(n_samples, TC_J));
  L.n_itiles = static_cast<int>(pf_ceil_div(n_samples, TC_I));
  L.n_tiles_padded = static_cast<int64_t>(L.n_jtiles) * (TC_J / PF_TILE);   // whole J tiles of operand stream
  if (L.n_tiles_padded < 2 * static_cast<int64_t>(L.n_itiles)) L.n_tiles_padded = 2 * static_cast<int64_t>(L.n_itiles);
  L.pairs = 0;
  for (int ti = 0; ti < L.n_itiles; ++ti) L.pairs += L.n_jtiles - (2 * ti) / 3;
  // split-K: ~20 waves of one CTA per SM, at least 64 words per CTA, boundaries multiples of HSUM_RUN
  int64_t splits = pf_ceil_div(static_cast<int64_t>(sms) * 20, L.pairs);
  const int64_t max_splits = pf_ceil_div(n_local, 64);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  L.words_per_split = pf_ceil_div(pf_ceil_div(n_local, splits), HSUM_RUN) * HSUM_RUN;
  L.splits = pf_ceil_div(n_local, L.words_per_split);
  if (L.splits < 1) L.splits = 1;
  L.hsum_ld = L.n_tiles_padded * PF_TILE;
  L.nib_bytes = tc_align(L.n_tiles_padded * n_local * 64 * 16);
  L.refv_off = L.nib_bytes;
  L.flags_off = L.refv_off + tc_align(n_local * 4);
  L.hsum_off = L.flags_off + 256;
  L.total = L.hsum_off + tc_align(L.splits * L.hsum_ld * 4);
  return L;
}

}  // namespace

PF_API int64_t pf_pair_counts_tc_work_bytes(int64_t n_samples, int64_t n_words_range) {
  if (n_samples <= 0 || n_words_range < 0) return 0;
  return tc_plan(n_samples, n_words_range, tc_sm_count()).total;
}

PF_API int pf_pair_counts_tc(const uint32_t* d_planes, int64_t n_samples, int64_t n_words, int64_t word_begin,
                             int64_t word_end, int32_t* d_counts, int64_t ld, void* d_work, int64_t work_bytes,
                             int* h_used, void* stream) {
  PF_TRY_BEGIN
  if (!d_planes || !d_counts || !d_work || !h_used || n_samples <= 0 || ld < n_samples || word_begin < 0 ||
      word_end < word_begin || word_end > n_words)
    return pf_fail("pf_pair_counts_tc: bad arguments");
  *h_used = 0;
  const int64_t n_local = word_end - word_begin;
  if (n_local == 0) {
    *h_used = 1;
    return PF_OK;
  }
  if (n_local * 32 >= (int64_t(1) << 31)) return pf_fail("pf_pair_counts_tc: more than 2^31 sites per call");
  const TcPlan L = tc_plan(n_samples, n_local, tc_sm_count());
  if (work_bytes < L.total)
    return pf_fail("pf_pair_counts_tc: workspace of %lld bytes needed, %lld given", (long long)L.total,
                   (long long)work_bytes);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(d_work);
  uint4* nib4 = reinterpret_cast<uint4*>(w);
  uint32_t* ref_v = reinterpret_cast<uint32_t*>(w + L.refv_off);
  int* flags = reinterpret_cast<int*>(w + L.flags_off);
  int* hsum = reinterpret_cast<int*>(w + L.hsum_off);

  PF_CUDA(cudaMemsetAsync(flags, 0, 256, s));
  PF_CUDA(cudaMemsetAsync(hsum, 0, static_cast<size_t>(L.splits * L.hsum_ld) * 4, s));
  {
    const int64_t n_runs = pf_ceil_div(n_local, HSUM_RUN);
    const int64_t threads = L.n_tiles_padded * n_runs * 64;
    planes_to_nib4_kernel<<<static_cast<unsigned>(pf_ceil_div(threads, 256)), 256, 0, s>>>(
        d_planes, n_words, word_begin, n_local, n_samples, L.n_tiles, L.n_tiles_padded, nib4, ref_v, flags, hsum,
        L.words_per_split, L.hsum_ld);
    PF_CUDA(cudaGetLastError());
  }
  int h_flags = 0;
  PF_CUDA(cudaMemcpyAsync(&h_flags, flags, sizeof(int), cudaMemcpyDeviceToHost, s));
  PF_CUDA(cudaStreamSynchronize(s));
  if (h_flags != 0) return PF_OK;   // mixed missingness: *h_used stays 0, nothing was added to d_counts

  TcParams P;
  P.nib4 = nib4;
  P.all_v = ref_v;
  P.n_local = n_local;
  P.n_samples = n_samples;
  P.counts = d_counts;
  P.ld = ld;
  P.n_itiles = L.n_itiles;
  P.n_jtiles = L.n_jtiles;
  P.n_tilepairs = L.pairs;
  P.flags = flags;
  P.hsum = hsum;
  P.hsum_ld = L.hsum_ld;
  P.words_per_split = static_cast<int>(L.words_per_split);
  P.prof = g_tc_prof_enabled ? 1 : 0;
  const int64_t grid = L.splits * L.pairs;
  if (grid > 0x7fffffffLL) return pf_fail("pf_pair_counts_tc: grid too large");
  const size_t smem = static_cast<size_t>(RAW_STAGES) * RAW_STAGE_BYTES + static_cast<size_t>(EXP_STAGES) * EXP_STAGE_BYTES;
  auto kernel = g_tc_prof_enabled ? pair_counts_tc_kernel<true> : pair_counts_tc_kernel<false>;
  PF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kernel<<<static_cast<unsigned>(grid), TC_THREADS, smem, s>>>(P);
  PF_CUDA(cudaGetLastError());
  *h_used = 1;
  return PF_OK;
  PF_TRY_END
}

// Diagnostic: enable role-level cycle accounting in the tensor-core kernel and read it back.
// h_out16: [0] MMA thread cycles waiting for expanded operands, [1] MMA thread total, [2..4] an A-row
// expansion warp: waiting for raw data / waiting for a free stage / total, [5] TMA producer waiting,
// [6] producer total, [8] pipeline steps, [9..11] a B-row expansion warp (as [2..4]); summed over CTAs.
PF_API int pf_pair_counts_tc_profile(int enable, long long* h_out16) {
  PF_TRY_BEGIN
  if (h_out16) PF_CUDA(cudaMemcpyFromSymbol(h_out16, g_tc_prof, sizeof(long long) * 16));
  long long zero[16] = {0};
  PF_CUDA(cudaMemcpyToSymbol(g_tc_prof, zero, sizeof(zero)));
  g_tc_prof_enabled = enable != 0;
  return PF_OK;
  PF_TRY_END
}
