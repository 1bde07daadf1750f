// Tensor-core variant of the all-pairs genotype counts: 4-bit indicator GEMMs on tcgen05
// (BASELINE.json north star: "an int8 tcgen05 indicator-GEMM variant kept only if ncu shows it
// beats popcount" — it does; the 4-bit form below is 2.2x the int8 form it replaced, see
// profiles/r01_mma_probe.md). Same arithmetic as pair_counts.cu, restated as matrix products over
// the site axis. Per genotype:
//   x = g - 1 in {-1, 0, +1} (0 for het and missing), h = 1 if homozygous, v = 1 if genotyped.
//   XX = sum x_i x_j = n00 + n22 - n02      HH = sum h_i h_j = n00 + n22 + n02
//   VV = sum v_i v_j = V                    TT = sum t_i t_j, t = v - h = 1 if heterozygous
//   D2 = n02 = (HH - XX) / 2                equal genotypes = (XX + HH) / 2 + TT     D1 = VV - that
// The operands are E2M1 (4-bit float) values, tcgen05.mma kind::mxf4 with all block scale factors 1.0:
// 64 sites per instruction instead of the 32 of kind::i8 at the same instruction time. -1, 0, +1 are
// exact in E2M1, products are -1 / 0 / +1, and the fp32 accumulators hold integers below 2^24 exactly
// (a CTA never accumulates more than 2^23 sites), so the counts are bit-exact.
// Two kernel modes share one body:
//   mode 0  XX and HH on 128 x 192 tiles. When every site is genotyped in all samples or in none
//           (no sample-specific missingness) this is all that is needed: the het-vs-hom counts |h_i|, |h_j| and V
//           are per-sample / global numbers the converter already has.
//   mode 1  TT and VV on the same tiles, run in addition when genotypes are missing in some samples only.
//           The two launches add their linear pieces into the same int32 matrices.
//
// Operand streams (produced once per call by planes_to_e2m1; the second one, for mode 1, only when it is
// needed: nibble = E2M1 code of t, het 0x2, missing 0x8, else 0x0; v = "sign bit clear"). Stream 0: one nibble per genotype that is at the
// same time the E2M1 code of x and a tag: hom-REF 0xA (-1), het 0x0 (+0), hom-ALT 0x2 (+1),
// missing 0x8 (-0). h = nibble & 2, v = "nibble != 8", t = "nibble == 0". It is stored step-major as 2 KB chunks (64
// samples x 64 sites) already in the canonical K-major shared-memory order of a UMMA operand, so ONE
// TMA bulk copy of the J rows of a step lands as a ready B_x operand.
// Data flow per CTA (one sample tile pair, a contiguous range of 64-site steps):
//   stream --TMA bulk copy per step--> raw ring (the B rows of a step = the B_x operand in mode 0);
//   the A rows go from the stream to registers of their expansion threads, one stage ahead
//   --expansion warps (thread = sample row): a few logic ops per 8 genotypes--> one A plane straight into
//   TMEM (tcgen05.st), the other A plane and the derived B planes into shared memory
//   --tcgen05.mma kind::mxf4, A from TMEM, B from shared memory (1 thread)--> fp32 accumulators in TMEM
//   --tcgen05.ld--> epilogue --> int32 RED atomics (split-K, bit-exact).
#include "pf_common.cuh"
#include "umma.cuh"

#include <cuda.h>   // CUtensorMap and its enums only: the encoder is fetched through cudaGetDriverEntryPoint (no -lcuda)

namespace {

constexpr int TC_I = 128;            // rows of the accumulator tile (MMA M); the A operand lives in TMEM
constexpr int STEP_WORDS = 2;        // one pipeline step = one MMA K step = 64 sites = two 32-site words
constexpr int CHUNK_BYTES = 2048;    // 64 samples x 64 sites x 4 bits
constexpr int MAX_RAW_STAGES = 4;
constexpr int EXP_STAGES = 2;        // a barrier hand-over costs ~230 clocks (profiles/r01_mma_probe.md): the ring
                                     // must cover MMA time + three hand-overs + the expansion itself
constexpr int TMEM_A_COLS = 8;       // per step: the 8 columns (32 K-bytes per row) of the A plane kept in TMEM;
                                     // the other A plane goes through shared memory so that four stages fit
constexpr int TMEM_SF = 496;         // 16 columns of scale factors (all 1.0) behind accumulators and A staging
constexpr uint32_t NIB_HOM = 0x22222222u;   // bit 1 of a stream nibble: homozygous; as E2M1 that bit alone is +1.0
constexpr int HSUM_RUN = 16;         // words per thread of the converter; split boundaries are multiples of it
constexpr int64_t MAX_WORDS_PER_CTA = 1 << 18;   // 2^23 sites: fp32 accumulators stay exact

// per-mode geometry. GROUP = steps per pipeline stage: the barrier round trips and fences of a stage
// (~500 clocks for an expansion warp) are paid once per GROUP steps; RAW = stages of the TMA ring.
template <int MODE> struct TcCfg;
template <> struct TcCfg<0> {   // stream 0: plane 0 = x (the stream itself), plane 1 = h; products XX (acc 0), HH (acc 1)
  static constexpr int J = 192, NPROD = 2, NEXP = 1;   // NEXP = derived B planes written by the expansion warps
  static constexpr int GROUP = 5, RAW = 3;
};
template <> struct TcCfg<1> {   // stream 1: plane 0 = t (the stream itself), plane 1 = v; products TT (acc 0), VV (acc 1)
  static constexpr int J = 192, NPROD = 2, NEXP = 1;
  static constexpr int GROUP = 5, RAW = 3;
};
template <int MODE> struct TcGeo {
  using C = TcCfg<MODE>;
  static constexpr int J = C::J;                               // columns (MMA N); B operand in shared memory
  static constexpr int GROUP = C::GROUP, RAW = C::RAW;
  static constexpr int ROWS = TC_I + J;                        // sample rows handled per step
  static constexpr int RAW_STEP_BYTES = (J / PF_TILE) * CHUNK_BYTES;   // the B rows of a step, canonical order
  static constexpr int RAW_STAGE_BYTES = GROUP * RAW_STEP_BYTES;
  static constexpr int EXP_PLANE_BYTES = J * 32;
  static constexpr int A_PLANE_BYTES = TC_I * 32;              // the A plane that goes through shared memory, in front
  static constexpr int EXP_STEP_BYTES = A_PLANE_BYTES + C::NEXP * EXP_PLANE_BYTES;
  static constexpr int EXP_STAGE_BYTES = GROUP * EXP_STEP_BYTES;
  static constexpr int TMEM_A = C::NPROD * J;                  // accumulators first, then A staging
  static constexpr int EXP_WARPS = TC_I / 32 + J / 32;         // one sample row per thread
  static constexpr int THREADS = (EXP_WARPS + 2) * 32;         // + TMA producer warp + MMA issuer warp
  static constexpr int RAW_EMPTY_ARRIVALS = J / 32 + 1;   // B-row warps + the MMA, which reads the plane-0 B operand from the raw stage
  static constexpr size_t SMEM = static_cast<size_t>(RAW) * RAW_STAGE_BYTES +
                                 static_cast<size_t>(EXP_STAGES) * EXP_STAGE_BYTES;
  static_assert(TMEM_A + EXP_STAGES * GROUP * TMEM_A_COLS <= TMEM_SF && TMEM_SF + 16 <= 512, "TMEM columns");
  static_assert(RAW <= MAX_RAW_STAGES && SMEM + 2048 <= 227 * 1024, "shared memory");
};

// ------------------------------------------------------------------ planes -> E2M1 stream
__device__ __forceinline__ uint32_t spread8(uint32_t b) {  // 8 bits -> bit 0 of 8 nibbles
  uint32_t x = b;
  x = (x | (x << 12)) & 0x000f000fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x;
}

// 16 bits -> 8 nibbles that each hold one bit PAIR (nibble k = bits 2k, 2k + 1): the selectors of two PRMTs
__device__ __forceinline__ uint32_t pairs16(uint32_t x16) {
  uint32_t t = (x16 | (x16 << 8)) & 0x00ff00ffu;
  t = (t | (t << 4)) & 0x0f0f0f0fu;
  t = (t | (t << 2)) & 0x33333333u;
  return t;
}

// 32 sign-plane bits S and 32 magnitude-plane bits M -> 32 nibbles (S << 3 | M << 1), without shared memory:
// a bit pair selects one byte (two nibbles) of a 4-entry table held in a register, so one PRMT (__byte_perm)
// expands 8 bits; the two tables are the two source registers of the same PRMT (selector 0-3: sign, 4-7:
// magnitude). ~46 ALU instructions per word against 8 shared-memory lookups (whose bank conflicts made the
// LUT form of this kernel LSU-bound at 0.57 of the HBM rate).
__device__ __forceinline__ uint4 expand_e2m1(uint32_t S, uint32_t M) {
  constexpr uint32_t TS = 0x88800800u, TM = 0x22200200u;
  const uint32_t s0 = pairs16(S & 0xffffu), s1 = pairs16(S >> 16);
  const uint32_t m0 = pairs16(M & 0xffffu) | 0x44444444u, m1 = pairs16(M >> 16) | 0x44444444u;
  uint4 o;
  o.x = __byte_perm(TS, TM, s0) | __byte_perm(TS, TM, m0);
  o.y = __byte_perm(TS, TM, s0 >> 16) | __byte_perm(TS, TM, m0 >> 16);
  o.z = __byte_perm(TS, TM, s1) | __byte_perm(TS, TM, m1);
  o.w = __byte_perm(TS, TM, s1 >> 16) | __byte_perm(TS, TM, m1 >> 16);
  return o;
}

// stream[((s * n_tiles_padded + T) * 128 + (l >> 3) * 16 + kh * 8 + (l & 7))] (uint4) = the 32 sites of word
// 2 s + kh of sample 64 T + l, nibble j of 32-bit word q = site 8 q + j: bit 3 = "hom-REF or missing",
// bit 1 = "homozygous" (see the file header; stream 1: bit 3 = "missing", bit 1 = "heterozygous"). A second
// half-step past the last word is all missing.
// ref_v[w] = valid-site mask of sample 0; flags[0] is set when any real sample's mask differs from
// it, i.e. when some site is genotyped in some samples and missing in others.
// hsum[split][sample] = homozygous genotypes of the sample in the split's word range (|h_i|).
// One thread = one sample x a run of HSUM_RUN consecutive words (runs never straddle a split).
template <int STREAM, bool LUT>
__global__ void __launch_bounds__(256)
planes_to_e2m1_kernel(const uint32_t* __restrict__ planes, int64_t n_words, int64_t word_begin, int64_t n_local,
                      int64_t n_samples, int64_t n_tiles, int64_t n_tiles_padded, uint4* __restrict__ stream,
                      uint32_t* __restrict__ ref_v, int* __restrict__ flags, int* __restrict__ hsum,
                      int64_t words_per_split, int64_t hsum_ld) {
  // 8 plane bits -> 8 nibbles through two 256-entry tables (bit 3 for the sign plane, bit 1 for the
  // homozygous plane): 20 instructions per word instead of ~110 of shift-and-mask spreading, which made
  // this kernel ALU-bound
  __shared__ uint32_t s_lut_s[LUT ? 256 : 1], s_lut_m[LUT ? 256 : 1];
  if (STREAM == 1 && __ldg(flags) == 0) return;   // stream 1 is needed only with sample-specific missingness
  if (LUT) {
    const uint32_t sp = spread8(threadIdx.x);
    s_lut_s[threadIdx.x] = sp << 3;
    s_lut_m[threadIdx.x] = sp << 1;
    __syncthreads();
  }
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
  const int unit = (l >> 3) * 16 + (l & 7);
  int hom = 0;
  bool mixed = false;
#pragma unroll(LUT ? 1 : 4)
  for (int64_t w = w0; w < w1; ++w) {
    uint32_t L = 0xffffffffu, H = 0xffffffffu;
    if (t < n_tiles) {
      const uint32_t* p = planes + ((t * n_words + word_begin + w) * 2) * PF_TILE;
      L = p[l];
      H = p[PF_TILE + l];
    }
    // stream 0: code 0 / 3 -> sign bit; code 0 / 2 -> bit 1 (x = -1 / +0 / +1 / -0 for codes 0 / 1 / 2 / 3)
    // stream 1: code 3 -> sign bit; code 1 -> bit 1       (t =  0 / +1 /  0 / -0)
    const uint32_t S = STREAM == 0 ? ~(L ^ H) : (L & H), M = STREAM == 0 ? ~L : (L & ~H);
    uint4 out;
    if (LUT) {
      out.x = s_lut_s[S & 0xffu] | s_lut_m[M & 0xffu];
      out.y = s_lut_s[(S >> 8) & 0xffu] | s_lut_m[(M >> 8) & 0xffu];
      out.z = s_lut_s[(S >> 16) & 0xffu] | s_lut_m[(M >> 16) & 0xffu];
      out.w = s_lut_s[S >> 24] | s_lut_m[M >> 24];
    } else {
      out = expand_e2m1(S, M);
    }
    stream[((w >> 1) * n_tiles_padded + t) * 128 + (w & 1) * 8 + unit] = out;
    if (w == n_local - 1 && (w & 1) == 0)
      stream[((w >> 1) * n_tiles_padded + t) * 128 + 8 + unit] = make_uint4(0x88888888u, 0x88888888u, 0x88888888u, 0x88888888u);
    if (STREAM == 0) {
      const uint32_t* p0 = planes + ((word_begin + w) * 2) * PF_TILE;   // sample 0: tile 0, lane 0
      const uint32_t ref = ~(p0[0] & p0[PF_TILE]);
      if (t == 0 && l == 0) ref_v[w] = ref;
      mixed |= (~(L & H) != ref);
      hom += __popc(M);             // homozygous <=> low code bit 0 (missing has it set)
    }
  }
  if (STREAM == 0 && real) {
    if (mixed) atomicOr(flags, 1);
    if (hom) atomicAdd(hsum + (w0 / words_per_split) * hsum_ld + t * PF_TILE + l, hom);
  }
}

// The same conversion with more work per thread: one thread = FOUR samples of a tile (lanes j, j + 16, j + 32,
// j + 48) x a run of HSUM_RUN words, walked one 64-site step (two words) at a time; the 8 expansions of a step are
// register-only (expand_e2m1). With that interleaving every load and every store instruction of the 16 threads of
// a tile touches whole 32-byte sectors (lanes j = 0..7 -> 8 consecutive 16-byte units of the chunk), and a thread
// keeps 16 independent loads in flight per step instead of 2 per word. (Four CONSECUTIVE samples per thread with
// 16-byte loads was measured first: its store instructions wrote 16 bytes per lane into 32 different sectors and
// the kernel took 16.3 ms on 3,000 x 10M against 6.3 ms for the one-sample-per-thread kernel above.)
// Same outputs bit for bit (stream, ref_v, flags, hsum).
template <int STREAM>
__global__ void __launch_bounds__(256)
planes_to_e2m1_x4_kernel(const uint32_t* __restrict__ planes, int64_t n_words, int64_t word_begin, int64_t n_local,
                         int64_t n_samples, int64_t n_tiles, int64_t n_tiles_padded, uint4* __restrict__ stream,
                         uint32_t* __restrict__ ref_v, int* __restrict__ flags, int* __restrict__ hsum,
                         int64_t words_per_split, int64_t hsum_ld) {
  static_assert(HSUM_RUN % 2 == 0, "runs are whole steps");
  if (STREAM == 1 && __ldg(flags) == 0) return;   // stream 1 is needed only with sample-specific missingness
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int j = static_cast<int>(idx & 15);        // lanes j + 16 k of the tile, k = 0..3
  const int64_t tr = idx >> 4;
  const int64_t n_runs = (n_local + HSUM_RUN - 1) / HSUM_RUN;
  // tile fastest: the CTAs that run at the same time then write ONE contiguous window of the step-major stream
  // (all tiles of a few runs of steps) instead of one 2 KB chunk out of every 96 KB over gigabytes
  const int64_t t = tr % n_tiles_padded;
  const int64_t run = tr / n_tiles_padded;
  if (run >= n_runs) return;
  const int64_t w0 = run * HSUM_RUN;
  const int64_t w1 = min(n_local, w0 + HSUM_RUN);
  const int unit0 = (j >> 3) * 16 + (j & 7);       // unit of lane j; lane j + 16 k is 32 k units further
  const bool tile_real = t < n_tiles;
  bool real[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) real[k] = (t * PF_TILE + j + 16 * k) < n_samples;
  int hom[4] = {0, 0, 0, 0};
  bool mixed = false;
#pragma unroll 2
  for (int64_t w = w0; w < w1; w += 2) {
    const bool has2 = (w + 1) < w1;
    uint32_t L[2][4], H[2][4];
#pragma unroll
    for (int kh = 0; kh < 2; ++kh)
#pragma unroll
      for (int k = 0; k < 4; ++k) L[kh][k] = H[kh][k] = 0xffffffffu;
    if (tile_real) {
      const uint32_t* p = planes + ((t * n_words + word_begin + w) * 2) * PF_TILE + j;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        L[0][k] = p[16 * k];
        H[0][k] = p[PF_TILE + 16 * k];
      }
      if (has2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          L[1][k] = p[2 * PF_TILE + 16 * k];
          H[1][k] = p[3 * PF_TILE + 16 * k];
        }
      }
    }
    uint4* chunk = stream + ((w >> 1) * n_tiles_padded + t) * 128 + unit0;
#pragma unroll
    for (int kh = 0; kh < 2; ++kh) {
      if (kh == 1 && !has2) {
        // a second half-step past the last word is all missing
        if (w == n_local - 1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) chunk[8 + 32 * k] = make_uint4(0x88888888u, 0x88888888u, 0x88888888u, 0x88888888u);
        }
        break;
      }
      uint32_t ref = 0;
      if (STREAM == 0) {
        const uint32_t* p0 = planes + ((word_begin + w + kh) * 2) * PF_TILE;   // sample 0: tile 0, lane 0
        ref = ~(__ldg(p0) & __ldg(p0 + PF_TILE));
        if (t == 0 && j == 0) ref_v[w + kh] = ref;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // stream 0: code 0 / 3 -> sign bit; code 0 / 2 -> bit 1 (x = -1 / +0 / +1 / -0 for codes 0 / 1 / 2 / 3)
        // stream 1: code 3 -> sign bit; code 1 -> bit 1       (t =  0 / +1 /  0 / -0)
        const uint32_t lw = L[kh][k], hw = H[kh][k];
        const uint32_t S = STREAM == 0 ? ~(lw ^ hw) : (lw & hw);
        const uint32_t M = STREAM == 0 ? ~lw : (lw & ~hw);
        chunk[kh * 8 + 32 * k] = expand_e2m1(S, M);
        if (STREAM == 0) {
          mixed |= real[k] && (~(lw & hw) != ref);
          hom[k] += __popc(M);          // homozygous <=> low code bit 0 (missing has it set)
        }
      }
    }
  }
  if (STREAM == 0) {
    if (mixed) atomicOr(flags, 1);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (real[k] && hom[k]) atomicAdd(hsum + (w0 / words_per_split) * hsum_ld + t * PF_TILE + j + 16 * k, hom[k]);
  }
}

// ------------------------------------------------------------------ the GEMM kernel
// optional role-level cycle accounting (pf_pair_counts_tc_profile): who waits for whom
__device__ long long g_tc_prof[16];
bool g_tc_prof_enabled = false;
int g_tc_debug = 0;

struct TcParams {
  const uint4* stream;    // [n_steps][n_tiles_padded] chunks of 2 KB: stream 0 (x) for mode 0, stream 1 (t) for mode 1
  int64_t n_tiles_padded;
  const uint32_t* all_v;  // [n_local] valid-site mask of sample 0 (= of every sample when uniform)
  int64_t n_local;        // words in the range
  int64_t n_steps;        // 64-site steps in the range
  int64_t n_samples;
  int32_t* counts;        // [3][n_samples][ld]
  int64_t ld;
  int n_itiles;           // ceil(n_samples / 128)
  int n_jtiles;           // ceil(n_samples / J)
  int n_tilepairs;
  int words_per_split;    // even
  const int* hsum;        // [splits][hsum_ld] homozygous counts per sample and split (mode 0, uniform)
  int64_t hsum_ld;
  const int* flags;       // flags[0] != 0: sample-specific missingness (set by the converter). Mode 0 then leaves
                          // V and VV - TT to the mode-1 launch; mode 1 exits at once when it is 0.
  int debug;              // single-CTA kernel, profile build only (results are wrong): 1 = MMAs do not wait for operands,
                          // 2 = expansion warps do no work, 4 = no MMAs are issued. CTA-pair kernel: 1 = every hand-over
                          // with a cluster-scope release ("strict", results identical); timing experiments with wrong
                          // results: 16 = expansion warps do no work, 32 = no A-row loads, 64 = profile the peer CTA,
                          // 256 = no wait for the raw data
  // CTA-pair kernel with TMAP: the stream as a 3-D tensor [step][row group of 8 samples][256 bytes]; boxes of
  // GROUP steps x 16 (A) / J / 16 (B) row groups, so that ONE request per operand fetches a whole stage
  alignas(64) CUtensorMap tm_a;
  alignas(64) CUtensorMap tm_b;
};

// block-scaled instruction descriptor (bit layout: CUTLASS cute/arch/mma_sm100_desc.hpp,
// InstrDescriptorBlockScaled): E2M1 A and B, K-major, UE8M0 scale factors with id 0, dense K = 64
__host__ __device__ constexpr uint32_t idesc_mxf4(int m, int n) {
  return (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (1u << 23) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem], E2M1 operands, K = 64, fp32 accumulation; scale factors from TMEM
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

// as above with the A operand from shared memory
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

// the two operand planes of one sample row and one 64-site step from its 8 stream words
template <int MODE>
__device__ __forceinline__ void make_planes(const uint4& lo, const uint4& hi, uint32_t (&p0)[8], uint32_t (&p1)[8]) {
  const uint32_t n[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    p0[m] = n[m];          // x / t: the stream nibble is its E2M1 code
    if (MODE == 0) p1[m] = n[m] & NIB_HOM;                              // h: bit 1 alone is +1.0
    else p1[m] = ((~n[m] >> 3) & 0x11111111u) << 1;                     // v: +1.0 unless the sign bit is set (nibble 0x8)
  }
}

// byte offset of the 16-byte unit (row r, K half kh) inside a canonical K-major no-swizzle tile
__device__ __forceinline__ uint32_t unit_off(int r, int kh) {
  return static_cast<uint32_t>((r >> 3) * 256 + kh * 128 + (r & 7) * 16);
}

template <int MODE, bool PROF>
__global__ void __launch_bounds__(TcGeo<MODE>::THREADS, 1)
pair_counts_tc_kernel(const TcParams P) {
  using G = TcGeo<MODE>;
  constexpr int J = G::J;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t raw_full[MAX_RAW_STAGES], raw_empty[MAX_RAW_STAGES];
  __shared__ __align__(8) uint64_t exp_full[EXP_STAGES], exp_empty[EXP_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t s_tmem;
  __shared__ int s_hcount[G::ROWS];
  __shared__ int s_valid_sites;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool general = __ldg(P.flags) != 0;
  if (MODE == 1 && !general) return;
  // shared-window addresses, computed once
  const uint32_t raw_a = umma::smem_addr(smem);                                   // G::RAW x RAW_STAGE_BYTES
  const uint32_t exp_a = raw_a + G::RAW * G::RAW_STAGE_BYTES;                 // EXP_STAGES x EXP_STAGE_BYTES (derived B planes)
  const uint32_t raw_full_a = umma::smem_addr(raw_full), raw_empty_a = umma::smem_addr(raw_empty);
  const uint32_t exp_full_a = umma::smem_addr(exp_full), exp_empty_a = umma::smem_addr(exp_empty);
  const uint32_t done_a = umma::smem_addr(&done_bar);

  // work item: (tile pair, split). I-tile ti (128 rows) pairs with the J-tiles that reach its first
  // row: tj >= floor(128 ti / J)
  const int pair = blockIdx.x % P.n_tilepairs;
  const int split = blockIdx.x / P.n_tilepairs;
  int ti = 0, rem = pair;
  while (true) {
    const int cnt = P.n_jtiles - (TC_I * ti) / J;
    if (rem < cnt) break;
    rem -= cnt;
    ++ti;
  }
  const int tj = (TC_I * ti) / J + rem;
  const int64_t w_begin = static_cast<int64_t>(split) * P.words_per_split;
  const int64_t w_end = min(P.n_local, w_begin + P.words_per_split);
  if (w_end <= w_begin) return;
  const int n_steps = static_cast<int>((w_end - w_begin + STEP_WORDS - 1) / STEP_WORDS);
  const int n_groups = (n_steps + G::GROUP - 1) / G::GROUP;
  const int64_t s_begin = w_begin / STEP_WORDS;

  if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
  if (tid == 32) {
    for (int s = 0; s < G::RAW; ++s) {
      umma::mbar_init(&raw_full[s], 1);
      umma::mbar_init(&raw_empty[s], G::RAW_EMPTY_ARRIVALS);
    }
    for (int e = 0; e < EXP_STAGES; ++e) {
      umma::mbar_init(&exp_full[e], G::EXP_WARPS);
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
  if (warp < TC_I / 32) {
    // block scale factors: UE8M0 127 = 2^0 everywhere (4 columns for A, 12 for B cover any tile shape)
    uint32_t ones[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ones[j] = 0x7f7f7f7fu;
#pragma unroll
    for (int c = 0; c < 16; c += 8) umma::tmem_st_32x8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + TMEM_SF + c, ones);
    umma::tmem_wait_st();
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();

  if (warp == G::EXP_WARPS) {
    // ===== TMA producer: one bulk copy per step. The stream is step-major, so the 64-sample chunks of
    // the B rows (J / 64) of a step are one contiguous range. (The TMA unit takes ~120 clocks per bulk
    // request whatever its size: five 2 KB requests per step capped the kernel at ~700 clocks per step;
    // the A rows are therefore loaded by their expansion threads directly.) =====
    if (lane == 0) {
      long long t_wait = 0;
      const long long t_start = PROF ? clock64() : 0;
      const int64_t step_stride = P.n_tiles_padded * (CHUNK_BYTES / 16);
      const uint4* src_b = P.stream + s_begin * step_stride + (J / PF_TILE) * static_cast<int64_t>(tj) * (CHUNK_BYTES / 16);
      int stage = 0;
      uint32_t phase = 1;        // parity of the "previous use" of a stage: first pass must not block
      for (int g = 0; g < n_groups; ++g) {
        if (g >= G::RAW) {
          const long long t0 = PROF ? clock64() : 0;
          umma::mbar_wait_a(raw_empty_a + stage * 8, phase);
          if (PROF) t_wait += clock64() - t0;
        }
        const int kw = min(G::GROUP, n_steps - g * G::GROUP);
        umma::mbar_expect_tx_a(raw_full_a + stage * 8, kw * G::RAW_STEP_BYTES);
#pragma unroll
        for (int j = 0; j < G::GROUP; ++j) {
          if (j < kw)
            umma::bulk_g2s_a(raw_a + stage * G::RAW_STAGE_BYTES + j * G::RAW_STEP_BYTES,
                             src_b + (static_cast<int64_t>(g) * G::GROUP + j) * step_stride, G::RAW_STEP_BYTES,
                             raw_full_a + stage * 8);
        }
        if (++stage == G::RAW) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (PROF) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[5]), static_cast<unsigned long long>(t_wait));
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[6]), static_cast<unsigned long long>(clock64() - t_start));
      }
    }
  } else if (warp == G::EXP_WARPS + 1) {
    // ===== MMA issuer: D[tmem] += A[tmem] * B[smem] =====
    if (lane == 0) {
      const uint32_t idesc = idesc_mxf4(TC_I, J);
      const uint32_t tsfa = tmem + TMEM_SF, tsfb = tmem + TMEM_SF + 4;
      long long t_wait = 0;
      const long long t_start = PROF ? clock64() : 0;
      int e = 0, stage = 0;
      uint32_t phase = 0;
      for (int g = 0; g < n_groups; ++g) {
        const long long t0 = PROF ? clock64() : 0;
        umma::mbar_wait_a(exp_full_a + e * 8, phase);
        if (PROF) t_wait += clock64() - t0;
        umma::fence_after_thread_sync();
        const int kw = min(G::GROUP, n_steps - g * G::GROUP);
        if (!(PROF && (P.debug & 4))) {
#pragma unroll
          for (int j = 0; j < G::GROUP; ++j) {
            if (j < kw) {
              const uint32_t a_t = tmem + G::TMEM_A + (e * G::GROUP + j) * TMEM_A_COLS;    // A plane 1 (h / v) in TMEM
              const uint32_t acc = (g > 0 || j > 0) ? 1u : 0u;
              const uint32_t ebase = exp_a + e * G::EXP_STAGE_BYTES + j * G::EXP_STEP_BYTES;
              const uint64_t a_s = umma::smem_desc_kmajor_noswizzle(ebase, 128, 256);    // A plane 0 (x / h) in shared memory
              const uint32_t bbase = ebase + G::A_PLANE_BYTES;
              const uint64_t b0 = umma::smem_desc_kmajor_noswizzle(
                  raw_a + stage * G::RAW_STAGE_BYTES + j * G::RAW_STEP_BYTES, 128, 256);
              const uint64_t b1 = umma::smem_desc_kmajor_noswizzle(bbase, 128, 256);
              mma_mxf4_ss(tmem + 0 * J, a_s, b0, idesc, acc, tsfa, tsfb);   // XX / TT
              mma_mxf4_ts(tmem + 1 * J, a_t, b1, idesc, acc, tsfa, tsfb);   // HH / VV
            }
          }
        }
        umma::commit_a(raw_empty_a + stage * 8);   // the raw stage held the plane-0 B operands
        umma::commit_a(exp_empty_a + e * 8);
        if (++e == EXP_STAGES) {
          e = 0;
          phase ^= 1;
        }
        if (++stage == G::RAW) stage = 0;
      }
      umma::commit_a(done_a);
      if (PROF) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[0]), static_cast<unsigned long long>(t_wait));
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[1]), static_cast<unsigned long long>(clock64() - t_start));
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[8]), static_cast<unsigned long long>(n_steps));
      }
    }
  } else {
    // ===== expansion warps: thread = one sample row. Warps 0-3: the 128 A rows, loaded from the
    // stream one stage ahead into registers, planes written straight into TMEM (their lane quadrant);
    // the others: the B rows, read from the raw ring, derived planes written to shared memory in the
    // canonical K-major layout =====
    const int row = tid;
    const bool is_a = row < TC_I;
    const int r = is_a ? row : row - TC_I;
    const uint32_t a_base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + G::TMEM_A;   // A rows: TMEM lane = row
    const uint32_t raw_row = raw_a + unit_off(r, 0);     // raw and expanded stages are canonical tiles
    const uint32_t exp_row = exp_a + unit_off(r, 0);
    const int64_t step_stride = P.n_tiles_padded * (CHUNK_BYTES / 16);
    // A rows: this thread's two 16-byte units of a step inside the chunk of its 64-sample tile
    const uint4* a_src = P.stream + s_begin * step_stride + (2 * static_cast<int64_t>(ti) + (r >> 6)) * (CHUNK_BYTES / 16) +
                         ((r & 63) >> 3) * 16 + (r & 7);
    long long t_wraw = 0, t_wemp = 0, t_fence = 0;
    const long long t_start = PROF ? clock64() : 0;
    int stage = 0, e = 0;
    uint32_t raw_phase = 0, exp_phase = 1;
    const bool idle = PROF && (P.debug & 2);
    uint4 lo[G::GROUP], hi[G::GROUP];
    auto load_a_rows = [&](int g) {
#pragma unroll
      for (int j = 0; j < G::GROUP; ++j) {
        lo[j] = make_uint4(0, 0, 0, 0);
        hi[j] = lo[j];
        const int st = g * G::GROUP + j;
        if (st < n_steps && !idle) {
          lo[j] = __ldg(a_src + st * step_stride);
          hi[j] = __ldg(a_src + st * step_stride + 8);
        }
      }
    };
    if (is_a) load_a_rows(0);
    for (int g = 0; g < n_groups; ++g) {
      {
        // B rows: lane 0 waits for the raw data, lane 1 for the free expanded stage, concurrently
        const long long t0 = PROF ? clock64() : 0;
        if (!is_a && lane == 0) umma::mbar_wait_a(raw_full_a + stage * 8, raw_phase);
        if (lane == 1 && g >= EXP_STAGES) umma::mbar_wait_a(exp_empty_a + e * 8, exp_phase);
        if (PROF && lane < 2) {
          if (lane == 0) t_wraw += clock64() - t0;
          else t_wemp += clock64() - t0;
        }
      }
      __syncwarp();
      if (is_a) umma::fence_after_thread_sync();   // the MMAs that read these TMEM columns have completed
      const int kw = min(G::GROUP, n_steps - g * G::GROUP);
      if (!is_a) {
        const uint32_t src = raw_row + stage * G::RAW_STAGE_BYTES;
#pragma unroll
        for (int j = 0; j < G::GROUP; ++j) {
          lo[j] = make_uint4(0, 0, 0, 0);
          hi[j] = lo[j];
          if (j < kw && !idle) {
            lo[j] = umma::lds128(src + j * G::RAW_STEP_BYTES);
            hi[j] = umma::lds128(src + j * G::RAW_STEP_BYTES + 128);
          }
        }
      }
      const long long tf0 = PROF ? clock64() : 0;
#pragma unroll
      for (int j = 0; j < G::GROUP; ++j) {
        if (j < kw && !idle) {
          uint32_t p0[8], p1[8];
          make_planes<MODE>(lo[j], hi[j], p0, p1);
          if (is_a) {
            // plane 0 (x / h) to the A tile in shared memory, plane 1 (h / v) to TMEM
            const uint32_t dst = exp_row + e * G::EXP_STAGE_BYTES + j * G::EXP_STEP_BYTES;
            umma::sts128(dst, p0[0], p0[1], p0[2], p0[3]);
            umma::sts128(dst + 128, p0[4], p0[5], p0[6], p0[7]);
            umma::tmem_st_32x8(a_base + (e * G::GROUP + j) * TMEM_A_COLS, p1);
          } else {
            const uint32_t dst = exp_row + e * G::EXP_STAGE_BYTES + j * G::EXP_STEP_BYTES + G::A_PLANE_BYTES;
            umma::sts128(dst, p1[0], p1[1], p1[2], p1[3]);           // h / v (plane 0 is read from the raw stage)
            umma::sts128(dst + 128, p1[4], p1[5], p1[6], p1[7]);
          }
        }
      }
      if (is_a) {
        umma::tmem_wait_st();
        umma::fence_before_thread_sync();
      }
      umma::fence_proxy_async_smem();
      __syncwarp();
      if (PROF) t_fence += clock64() - tf0;
      if (lane == 0) {
        umma::mbar_arrive_a(exp_full_a + e * 8);
        if (!is_a) umma::mbar_arrive_a(raw_empty_a + stage * 8);
      }
      // next stage's rows: in flight while this stage's MMAs run (issued after the fences, which
      // would otherwise wait for these loads)
      if (is_a) load_a_rows(g + 1);
      if (++stage == G::RAW) {
        stage = 0;
        raw_phase ^= 1;
      }
      if (++e == EXP_STAGES) {
        e = 0;
        exp_phase ^= 1;
      }
    }
    t_wemp = __shfl_sync(0xffffffffu, t_wemp, 1);
    if (PROF && lane == 0 && (warp == 0 || warp == 4)) {
      const int o = warp == 0 ? 2 : 9;   // [2..4] an A-row warp, [9..11] a B-row warp
      atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[o]), static_cast<unsigned long long>(t_wraw));
      atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[o + 1]), static_cast<unsigned long long>(t_wemp));
      atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[o + 2]), static_cast<unsigned long long>(clock64() - t_start));
      atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[warp == 0 ? 12 : 13]), static_cast<unsigned long long>(t_fence));
    }
    const bool uniform = (MODE == 0) && !general;
    if (uniform) {
      // |h_i| of this word range from the converter; V = sites genotyped in (all) samples
      const int64_t sample = is_a ? (static_cast<int64_t>(ti) * TC_I + r) : (static_cast<int64_t>(tj) * J + r);
      s_hcount[row] = (sample < P.n_samples) ? P.hsum[static_cast<int64_t>(split) * P.hsum_ld + sample] : 0;
      int vs = 0;
      for (int64_t w = w_begin + tid; w < w_end; w += G::EXP_WARPS * 32) vs += __popc(P.all_v[w]);
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, off);
      if (lane == 0 && vs) atomicAdd(&s_valid_sites, vs);
    }
    // all expansion warps meet before the epilogue (named barrier 1)
    asm volatile("bar.sync 1, %0;" ::"r"(G::EXP_WARPS * 32) : "memory");

    // ===== epilogue: TMEM (fp32, integer-valued) -> registers -> counts =====
    umma::mbar_wait_a(done_a, 0);
    umma::fence_after_thread_sync();
    const int lane_grp = warp & 3;               // TMEM lanes [32 g, 32 g + 32) are visible to warps with w % 4 == g
    const int grp_rank = warp >> 2;              // rank of this warp among the warps of its lane group
    const int grp_size = (lane_grp < (G::EXP_WARPS & 3)) ? (G::EXP_WARPS / 4 + 1) : (G::EXP_WARPS / 4);
    const int i_local = lane_grp * 32 + lane;
    const int64_t i = static_cast<int64_t>(ti) * TC_I + i_local;
    const int hi_cnt = uniform ? s_hcount[i_local] : 0;
    const int v_sites = s_valid_sites;
    const int64_t plane = P.n_samples * P.ld;
    int32_t* V = P.counts;
    int32_t* D1 = P.counts + plane;
    int32_t* D2 = P.counts + 2 * plane;
    for (int chunk = grp_rank; chunk < J / 32; chunk += grp_size) {
      const uint32_t taddr = tmem + (static_cast<uint32_t>(lane_grp * 32) << 16) + chunk * 32;
      if (MODE == 0) {
        uint32_t pv[32], qv[32];
        umma::tmem_ld_32x32(taddr + 0 * J, pv);
        umma::tmem_ld_32x32(taddr + 1 * J, qv);
        umma::tmem_wait_ld();
        if (i < P.n_samples) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int j_local = chunk * 32 + c;
            const int64_t j = static_cast<int64_t>(tj) * J + j_local;
            if (j >= P.n_samples || j < i) continue;       // each unordered pair once: row <= column
            const int xx = __float2int_rn(__uint_as_float(pv[c])), hh = __float2int_rn(__uint_as_float(qv[c]));
            const int d2 = (hh - xx) >> 1;
            // D1 = (valid sites) - (sites with equal genotypes). Uniform: |h_i| + |h_j| - 2 HH + D2.
            // General: VV - TT - (XX + HH) / 2, VV - TT coming from the mode-1 launch; XX + HH is twice the
            // number of equal homozygous sites of this CTA's site range, so the halving is exact.
            const int d1 = uniform ? (hi_cnt + s_hcount[TC_I + j_local] - 2 * hh + d2) : -((xx + hh) >> 1);
            if (uniform) atomicAdd(V + i * P.ld + j, v_sites);
            atomicAdd(D1 + i * P.ld + j, d1);
            atomicAdd(D2 + i * P.ld + j, d2);
            if (j != i) {
              if (uniform) atomicAdd(V + j * P.ld + i, v_sites);
              atomicAdd(D1 + j * P.ld + i, d1);
              atomicAdd(D2 + j * P.ld + i, d2);
            }
          }
        }
      } else {
        uint32_t tt[32], vv[32];
        umma::tmem_ld_32x32(taddr + 0 * J, tt);
        umma::tmem_ld_32x32(taddr + 1 * J, vv);
        umma::tmem_wait_ld();
        if (i < P.n_samples) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int64_t j = static_cast<int64_t>(tj) * J + chunk * 32 + c;
            if (j >= P.n_samples || j < i) continue;
            const int v = __float2int_rn(__uint_as_float(vv[c]));
            const int d1 = v - __float2int_rn(__uint_as_float(tt[c]));
            atomicAdd(V + i * P.ld + j, v);
            atomicAdd(D1 + i * P.ld + j, d1);
            if (j != i) {
              atomicAdd(V + j * P.ld + i, v);
              atomicAdd(D1 + j * P.ld + i, d1);
            }
          }
        }
      }
    }
    umma::fence_before_thread_sync();
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------ the CTA-pair GEMM kernel
// tcgen05.mma.cta_group::2: a cluster of two CTAs (the two SMs of a TPC) works on one 256 x J sample tile pair
// (default J = 208). Each CTA holds 128 A rows and J / 2 B rows of every 64-site step in its own shared memory and
// accumulates its 128 rows x J columns (two products) in its own tensor memory; the leader CTA (cluster rank 0)
// issues the instructions for both. Measured (pf_probe_umma_fp4_2cta_rate): the pair instruction takes 0.5
// clocks per B column (120 clocks for 256 x 240 x 64, the nominal fp4 rate) where the single-CTA one takes 147
// clocks for any N <= 256, and every SM fetches and expands only half of the B tile.
// Default data flow per CTA (the Tc2Cfg flags below select the alternatives that the tests also run):
//   TMA producer (1 thread): two tensor-map copies per stage - the 128 A rows and the J / 2 B rows of GROUP steps -
//     into a ring of RAW raw stages. Plane 0 of both operands (x / t) is read by the MMAs straight from there.
//   A-row warps (2 threads per row): LDS raw -> plane 1 (h / v) -> TENSOR memory (tcgen05.st): no shared-memory
//     write, no proxy fence, no operand read from shared memory for the A side of the second product.
//   B-row warps (2 threads per row): LDS raw -> plane 1 -> shared memory (ring of EXP expanded stages).
//   MMA issuer (leader CTA, 1 thread): per step XX += A_x(raw) B_x(raw), HH += A_h(TMEM) B_h(expanded); one
//     multicast commit per stage.
// Hand-over. Inside a CTA through hardware named barriers: thread 0 polls the mbarriers (raw data landed, stage
// slot free) and releases the expansion warps with bar.sync; after the stage's writes (tcgen05.wait::st +
// tcgen05.fence / fence.proxy.async) every warp bumps a shared-memory counter (acq_rel at CTA scope) and the warp that
// finishes the slot last arrives on the LEADER's exp_full barrier (count 2: one arrival per CTA; the peer's is a
// remote arrive). Stage consumed: ONE multicast tcgen05.commit per stage on
// a ring of stage_done barriers in both CTAs; the expansion warps (reuse of the expanded stage EXP stages later)
// and the TMA producer (reuse of the raw stage RAW stages later) wait on the same ring, whose depth max(RAW, EXP)
// keeps a waiter at most one phase behind.
// Memory-model note: the remote arrive is the default `mbarrier.arrive.shared::cluster` (release at CTA
// scope), as in CUTLASS's ClusterBarrier::arrive. The PTX model would ask for `.release.cluster`; that
// compiles to MEMBAR.ALL.GPU, which stalls the SM's whole memory pipeline (563 instead of 263 clocks per step even
// with one such arrive per CTA and stage). The data in question is shared / tensor memory written by STS /
// tcgen05.st and completed (FENCE.VIEW.ASYNC, tcgen05.wait::st) before the arrive leaves the SM, so nothing is in
// flight that the heavier fence could order; TcParams::debug bit 0 selects the strict form
// (tests/test_gpu_kernels.py runs both and compares them with the oracle bit for bit).
// B tiles of A block ti start at its first row (256 ti + J k), so only the first tile of a block straddles
// the diagonal. profiles/r01_mma_probe.md lists the variants that were measured on the way.
constexpr int TC2_I = 256, TC2_IH = 128;     // rows per pair / per CTA
// Geometry. J = columns of the pair's tile (each CTA supplies J / 2 B rows). With both A planes in shared memory
// (J = 240) the shared-memory data pipe was the bound (ncu: LSU wavefronts 46 % + tensor-core operand reads 27 %
// of the peak, tensor pipe 53 % active at 423 clocks per step), so with AH_TMEM the derived A plane (h / v) goes
// to TENSOR memory: the 2 J accumulator and 16 scale-factor columns leave 512 - 2 J - 16 columns for EXP x GROUP
// steps of 8 columns, which is what limits J to 208 there.
template <int CFG> struct Tc2Cfg;
// SPLIT: two threads per sample row (one per 32-site half of a step) instead of one: the per-stage latency
// chain of an expansion warp (wake-up, loads, stores, fences, arrive) is what the MMAs wait for, and it
// halves when a warp handles half the bytes
// A_TMA: the A rows arrive by TMA too (a second bulk copy per step) and plane 0 of A is read by the MMAs straight
// from the raw ring; the A-row warps then only derive plane 1 (LDS -> tensor or shared memory) and never have
// global loads in flight across a proxy fence (fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC,
// and the MEMBAR waits for the register prefetch of the following stages).
// TMAP: the raw stages are filled by two tensor-map copies per STAGE (cp.async.bulk.tensor.3d, a box of GROUP
// steps) instead of two bulk copies per STEP (the TMA unit retires about one request per 128 clocks whatever its
// size, so two requests per step would cap a step at ~256 clocks; at today's 263 the two forms measure the same).
template <> struct Tc2Cfg<0> { static constexpr int J = 208, GROUP = 5, RAW = 4, EXP = 2; static constexpr bool AH_TMEM = true, SPLIT = true, A_TMA = true, TMAP = true; };
template <> struct Tc2Cfg<1> { static constexpr int J = 240, GROUP = 6, RAW = 3, EXP = 2; static constexpr bool AH_TMEM = false, SPLIT = false, A_TMA = false, TMAP = false; };
template <> struct Tc2Cfg<2> { static constexpr int J = 240, GROUP = 5, RAW = 3, EXP = 2; static constexpr bool AH_TMEM = false, SPLIT = true, A_TMA = true, TMAP = false; };
template <> struct Tc2Cfg<3> { static constexpr int J = 208, GROUP = 5, RAW = 4, EXP = 2; static constexpr bool AH_TMEM = true, SPLIT = true, A_TMA = false, TMAP = false; };
template <> struct Tc2Cfg<4> { static constexpr int J = 208, GROUP = 5, RAW = 4, EXP = 2; static constexpr bool AH_TMEM = true, SPLIT = true, A_TMA = true, TMAP = false; };
// Narrower tiles leave tensor-memory columns for more steps of the A plane. Measured on 3,000 x 10M (tools/ab_tc.py,
// profiles/r02_ab_tc_ring.txt, where the variants are numbered 5..9): a DEEPER ring of expanded stages does not help -
// J = 176 with 3 x 5 or 3 x 6 steps 247 clocks per step, J = 192 with 3 x 4 294, J = 160 with 4 x 5 248, whatever the
// 176 / 192 / 160 clocks of their two instructions: the step is bound by the expansion warps' shared-memory traffic
// and hand-over, not by the depth - but LONGER stages do: J = 192 with 2 x 7 steps 236 clocks per step = 1.23 per B
// column against 1.28 for J = 208 with 2 x 5. That one is kept as cfg 5 for A/B runs; it is not the default: the
// narrower tiles are more tiles (104 against 98 at 3,000 samples) with other split counts, and whole calls measure
// 35.5 against 34.7 ms at 3,000 x 10M, 18.4 against 17.7 at 5,000 x 2M, 194 against 179-194 at 10,000 x 5M, and win
// only on short site ranges (4.42 against 4.63 ms at 3,000 x 1.25M, one GPU's share of eight);
// profiles/r02_ab_tc_auto.txt. A per-call choice from the kernel's cost model was tried and predicted none of this.
template <> struct Tc2Cfg<5> { static constexpr int J = 192, GROUP = 7, RAW = 3, EXP = 2; static constexpr bool AH_TMEM = true, SPLIT = true, A_TMA = true, TMAP = true; };
constexpr int TC2_CFGS = 6;
#define TC2_CFG_FIELD(cfg, F)                                                                                        \
  ((cfg) == 1 ? Tc2Cfg<1>::F : (cfg) == 2 ? Tc2Cfg<2>::F : (cfg) == 3 ? Tc2Cfg<3>::F : (cfg) == 4 ? Tc2Cfg<4>::F     \
   : (cfg) == 5 ? Tc2Cfg<5>::F : Tc2Cfg<0>::F)
constexpr int tc2_cols(int cfg) { return TC2_CFG_FIELD(cfg, J); }
constexpr int tc2_group(int cfg) { return TC2_CFG_FIELD(cfg, GROUP); }
constexpr bool tc2_tmap(int cfg) { return TC2_CFG_FIELD(cfg, TMAP); }
template <int CFG> struct Tc2Geo {
  using C = Tc2Cfg<CFG>;
  static constexpr int J = C::J, JH = C::J / 2;
  static constexpr bool AH_TMEM = C::AH_TMEM, SPLIT = C::SPLIT;
  static constexpr int NU = SPLIT ? 1 : 2;                                  // 16-byte units of a row and step per thread
  static constexpr int GROUP = C::GROUP, RAW = C::RAW, EXP = C::EXP;
  static constexpr int DONE = RAW > EXP ? RAW : EXP;                        // depth of the stage_done ring
  static constexpr bool A_TMA = C::A_TMA;
  static constexpr int A_PLANE_BYTES = TC2_IH * 32;                         // 4096
  static constexpr int B_STEP_BYTES = JH * 32;                              // the B rows of a step
  static constexpr bool TMAP = C::TMAP;
  static constexpr int RAW_B_OFF = A_TMA ? A_PLANE_BYTES : 0;               // bulk copies: a raw step is [A rows (A_TMA)][B rows]
  static constexpr int RAW_STEP_BYTES = RAW_B_OFF + B_STEP_BYTES;
  static constexpr int RAW_STAGE_BYTES = GROUP * RAW_STEP_BYTES;
  // where the A / B rows of step j of a raw stage are: bulk copies interleave them per step, the two tensor-map
  // boxes of a stage land as [A rows of GROUP steps][B rows of GROUP steps]
  static constexpr int RAW_A_STRIDE = TMAP ? A_PLANE_BYTES : RAW_STEP_BYTES;
  static constexpr int RAW_B_BASE = TMAP ? GROUP * A_PLANE_BYTES : RAW_B_OFF;
  static constexpr int RAW_B_STRIDE = TMAP ? B_STEP_BYTES : RAW_STEP_BYTES;
  static_assert(!TMAP || (A_TMA && B_STEP_BYTES % 128 == 0), "tensor-map boxes land on 128-byte boundaries");
  static constexpr int A_PLANES_SMEM = (A_TMA ? 0 : 1) + (AH_TMEM ? 0 : 1);   // A planes the expansion warps write to shared memory
  static constexpr int EXP_A1_OFF = A_TMA ? 0 : A_PLANE_BYTES;              // expanded step: [A plane 0 (!A_TMA)][A plane 1 (!AH_TMEM)][B plane 1]
  static constexpr int EXP_STEP_BYTES = A_PLANES_SMEM * A_PLANE_BYTES + B_STEP_BYTES;
  static constexpr int EXP_STAGE_BYTES = GROUP * EXP_STEP_BYTES;
  static constexpr int TMEM_A = 2 * J;                                       // A plane 1 staging behind the accumulators
  static constexpr int A_WARPS = 4 * (2 / NU);
  static constexpr int A_THREADS = A_WARPS * 32;
  static constexpr int B_WARPS = (JH * (2 / NU) + 31) / 32;
  static constexpr int EXP_WARPS = A_WARPS + B_WARPS;
  static constexpr int THREADS = (EXP_WARPS + 2) * 32;
  static constexpr size_t SMEM = static_cast<size_t>(RAW) * RAW_STAGE_BYTES + static_cast<size_t>(EXP) * EXP_STAGE_BYTES;
  static_assert(J % 16 == 0 && JH % 8 == 0, "tile width");
  static_assert(2 * J + (AH_TMEM ? EXP * GROUP * TMEM_A_COLS : 0) <= TMEM_SF && TMEM_SF + 16 <= 512, "TMEM columns");
  static_assert(RAW <= 8 && EXP <= 8 && SMEM + 4096 <= 227 * 1024, "shared memory");
};

// planes of one 16-byte unit (32 sites of one sample row): p0 = the stream itself (x / t), p1 = h / v
template <int MODE>
__device__ __forceinline__ void make_unit_planes(const uint4& n, uint4& p1) {
  if (MODE == 0) {
    p1 = make_uint4(n.x & NIB_HOM, n.y & NIB_HOM, n.z & NIB_HOM, n.w & NIB_HOM);
  } else {
    p1 = make_uint4(((~n.x >> 3) & 0x11111111u) << 1, ((~n.y >> 3) & 0x11111111u) << 1, ((~n.z >> 3) & 0x11111111u) << 1,
                    ((~n.w >> 3) & 0x11111111u) << 1);
  }
}

// work item -> (A block ti, first column of the B tile)
__device__ __forceinline__ void tc2_tile(int pair, int64_t n_samples, int J, int& ti, int64_t& b_start) {
  ti = 0;
  int rem = pair;
  while (true) {
    const int cnt = static_cast<int>((n_samples - static_cast<int64_t>(TC2_I) * ti + J - 1) / J);
    if (rem < cnt) break;
    rem -= cnt;
    ++ti;
  }
  b_start = static_cast<int64_t>(TC2_I) * ti + static_cast<int64_t>(J) * rem;
}

// position in a ring of N barriers: slot and the parity of the phase that use completes
template <int N> struct RingPos {
  int slot = 0;
  uint32_t parity = 0;
  __device__ __forceinline__ void advance() {
    if (++slot == N) {
      slot = 0;
      parity ^= 1;
    }
  }
};

template <int MODE, int CFG, bool PROF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Tc2Geo<CFG>::THREADS, 1)
pair_counts_tc2_kernel(const __grid_constant__ TcParams P) {
  using G = Tc2Geo<CFG>;
  constexpr int J = G::J, JH = G::JH;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t raw_full[G::RAW], exp_full[G::EXP], stage_done[G::DONE];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t s_tmem;
  __shared__ int s_hcount[TC2_IH + J];
  __shared__ int s_valid_sites;
  __shared__ int s_stage_warps[8];      // per expanded slot: expansion warps that have finished writing it

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool general = __ldg(P.flags) != 0;
  if (MODE == 1 && !general) return;
  const uint32_t rank = umma::cluster_ctarank();
  const uint32_t raw_a = umma::smem_addr(smem);
  const uint32_t exp_a = raw_a + G::RAW * G::RAW_STAGE_BYTES;
  const uint32_t raw_full_a = umma::smem_addr(raw_full), exp_full_a = umma::smem_addr(exp_full);
  const uint32_t stage_done_a = umma::smem_addr(stage_done);
  const uint32_t done_a = umma::smem_addr(&done_bar);
  const uint32_t exp_full_leader = umma::mapa(exp_full_a, 0);      // the leader's barrier, cluster address
  const bool strict = (P.debug & 1) != 0;

  const int item = static_cast<int>(blockIdx.x >> 1);
  const int pair = item % P.n_tilepairs;
  const int split = item / P.n_tilepairs;
  int ti;
  int64_t b_start;
  tc2_tile(pair, P.n_samples, J, ti, b_start);
  const int64_t w_begin = static_cast<int64_t>(split) * P.words_per_split;
  const int64_t w_end = min(P.n_local, w_begin + P.words_per_split);
  if (w_end <= w_begin) return;                 // same decision in both CTAs of the pair
  const int n_steps = static_cast<int>((w_end - w_begin + STEP_WORDS - 1) / STEP_WORDS);
  const int n_groups = (n_steps + G::GROUP - 1) / G::GROUP;
  const int64_t s_begin = w_begin / STEP_WORDS;
  const int64_t a_row0 = static_cast<int64_t>(TC2_I) * ti + TC2_IH * rank;   // first A sample of this CTA
  const int64_t b_row0 = b_start + JH * rank;                                 // first B sample of this CTA

  if (warp == 0) umma::tmem_alloc_2cta(&s_tmem, 512);
  if (tid == 32) {
    for (int s = 0; s < G::RAW; ++s) umma::mbar_init(&raw_full[s], 1);
    for (int e = 0; e < G::EXP; ++e) umma::mbar_init(&exp_full[e], 2);                    // one arrival per CTA (used in the leader)
    for (int d = 0; d < G::DONE; ++d) umma::mbar_init(&stage_done[d], 1);
    umma::mbar_init(&done_bar, 1);
    umma::fence_mbar_init();
  }
  if (tid == 0) s_valid_sites = 0;
  if (tid < 8) s_stage_warps[tid] = 0;
  umma::fence_before_thread_sync();
  umma::cluster_sync();                          // barriers of both CTAs exist before any remote arrival
  umma::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;
  if (warp < 4) {
    uint32_t ones[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ones[j] = 0x7f7f7f7fu;
#pragma unroll
    for (int c = 0; c < 16; c += 8) umma::tmem_st_32x8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + TMEM_SF + c, ones);
    umma::tmem_wait_st();
  }
  umma::fence_before_thread_sync();
  umma::cluster_sync();                          // scale factors of both CTAs are in place
  umma::fence_after_thread_sync();

  const int64_t step_stride = P.n_tiles_padded * (CHUNK_BYTES / 16);
  if (warp == G::EXP_WARPS) {
    // ===== TMA producer: the J / 2 B rows of this CTA (A_TMA: and its 128 A rows), one bulk copy each per step =====
    if (lane == 0) {
      const uint4* src_b = P.stream + s_begin * step_stride + (b_row0 >> 3) * 16;
      const uint4* src_a = P.stream + s_begin * step_stride + (a_row0 >> 3) * 16;
      int stage = 0;
      RingPos<G::DONE> freed;                    // stage g - RAW, whose MMAs must have completed
      for (int g = 0; g < n_groups; ++g) {
        if (g >= G::RAW) {
          umma::mbar_wait_a(stage_done_a + freed.slot * 8, freed.parity);
          freed.advance();
        }
        const int kw = min(G::GROUP, n_steps - g * G::GROUP);
        if (G::TMAP) {
          // one box of GROUP steps per operand (steps past the end of the stream are zero-filled, steps past
          // this CTA's range are loaded and not used): the transaction count is always the full boxes
          const uint32_t dst = raw_a + stage * G::RAW_STAGE_BYTES;
          const int step0 = static_cast<int>(s_begin) + g * G::GROUP;
          umma::mbar_expect_tx_a(raw_full_a + stage * 8, G::RAW_STAGE_BYTES);
          umma::tensor_g2s_3d(dst, &P.tm_a, 0, static_cast<int>(a_row0 >> 3), step0, raw_full_a + stage * 8);
          umma::tensor_g2s_3d(dst + G::RAW_B_BASE, &P.tm_b, 0, static_cast<int>(b_row0 >> 3), step0, raw_full_a + stage * 8);
          if (++stage == G::RAW) stage = 0;
          continue;
        }
        umma::mbar_expect_tx_a(raw_full_a + stage * 8, kw * G::RAW_STEP_BYTES);
#pragma unroll
        for (int j = 0; j < G::GROUP; ++j) {
          if (j < kw) {
            const uint32_t dst = raw_a + stage * G::RAW_STAGE_BYTES + j * G::RAW_STEP_BYTES;
            const int64_t off = (static_cast<int64_t>(g) * G::GROUP + j) * step_stride;
            if (G::A_TMA) umma::bulk_g2s_a(dst, src_a + off, G::A_PLANE_BYTES, raw_full_a + stage * 8);
            umma::bulk_g2s_a(dst + G::RAW_B_OFF, src_b + off, G::B_STEP_BYTES, raw_full_a + stage * 8);
          }
        }
        if (++stage == G::RAW) stage = 0;
      }
    }
  } else if (warp == G::EXP_WARPS + 1) {
    // ===== MMA issuer (leader CTA only): D[tmem of both CTAs] += A * B over the pair =====
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = idesc_mxf4(TC2_I, J);
      const uint32_t tsfa = tmem + TMEM_SF, tsfb = tmem + TMEM_SF + 4;
      long long t_wait = 0;
      const long long t_start = PROF ? clock64() : 0;
      int e = 0, stage = 0, d = 0;
      uint32_t phase = 0;
      // Shared-memory descriptors differ only in the start-address field (bits [0, 14) of the low word, in
      // 16-byte units; all addresses are below 256 KB, so adding an offset never carries out of the field):
      // one 32-bit add per operand instead of rebuilding 64-bit descriptors. The issuing thread's own
      // instruction stream was the limit with the tensor-memory A plane (293 clocks per step for 208 of MMAs).
      const uint64_t desc0 = umma::smem_desc_kmajor_noswizzle(0, 128, 256);
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32), desc_lo = static_cast<uint32_t>(desc0);
      auto desc_at = [&](uint32_t lo) { return (static_cast<uint64_t>(desc_hi) << 32) | lo; };
      auto issue_stage = [&](const bool full, int kw, uint32_t first_acc) {
        const uint32_t e_lo = desc_lo + ((exp_a + e * G::EXP_STAGE_BYTES) >> 4);
        const uint32_t r_lo = desc_lo + ((raw_a + stage * G::RAW_STAGE_BYTES) >> 4);
        const uint32_t ta = tmem + G::TMEM_A + e * G::GROUP * TMEM_A_COLS;
#pragma unroll
        for (int j = 0; j < G::GROUP; ++j) {
          if (full || j < kw) {
            const uint32_t acc = j > 0 ? 1u : first_acc;
            const uint32_t ex = e_lo + j * (G::EXP_STEP_BYTES >> 4);
            const uint32_t ra = r_lo + j * (G::RAW_A_STRIDE >> 4), rb = r_lo + ((G::RAW_B_BASE + j * G::RAW_B_STRIDE) >> 4);
            umma::mma_mxf4_ss_2cta(tmem + 0 * J, desc_at(G::A_TMA ? ra : ex), desc_at(rb), idesc, acc, tsfa, tsfb);   // XX / TT
            const uint64_t b1 = desc_at(ex + ((G::A_PLANES_SMEM * G::A_PLANE_BYTES) >> 4));
            if (G::AH_TMEM) umma::mma_mxf4_ts_2cta(tmem + 1 * J, ta + j * TMEM_A_COLS, b1, idesc, acc, tsfa, tsfb);
            else umma::mma_mxf4_ss_2cta(tmem + 1 * J, desc_at(ex + (G::EXP_A1_OFF >> 4)), b1, idesc, acc, tsfa, tsfb);                   // HH / VV
          }
        }
      };
      for (int g = 0; g < n_groups; ++g) {
        const long long t0 = PROF ? clock64() : 0;
        umma::mbar_wait_cluster_a(exp_full_a + e * 8, phase);
        if (PROF) t_wait += clock64() - t0;
        umma::fence_after_thread_sync();
        const int kw = min(G::GROUP, n_steps - g * G::GROUP);
        if (kw == G::GROUP) issue_stage(true, kw, g > 0 ? 1u : 0u);
        else issue_stage(false, kw, g > 0 ? 1u : 0u);
        umma::commit_2cta_a(stage_done_a + d * 8, 3);          // both CTAs: raw and expanded stage of g are free
        if (++e == G::EXP) {
          e = 0;
          phase ^= 1;
        }
        if (++stage == G::RAW) stage = 0;
        if (++d == G::DONE) d = 0;
      }
      umma::commit_2cta_a(done_a, 3);
      if (PROF) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[0]), static_cast<unsigned long long>(t_wait));
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[1]), static_cast<unsigned long long>(clock64() - t_start));
        atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[8]), static_cast<unsigned long long>(n_steps));
      }
    }
  } else {
    // ===== expansion warps: thread = one sample row of this CTA (SPLIT: one 32-site half of it). The A-row
    // and the B-row warps run separate, branch-free instruction streams (a stage of six steps was ~365
    // instructions per warp when both roles shared one predicated body, and an in-order warp needs ~3 clocks
    // per instruction: the expansion warps, not the memory system, were what the MMAs waited for) =====
    constexpr int NU = G::NU;
    constexpr int NR = G::GROUP * NU;
    long long t_wait = 0, t_fence = 0, t_end = 0;
    const long long t_start = PROF ? clock64() : 0;
    // Stage hand-over inside a CTA goes through hardware named barriers: ONE thread polls the mbarriers (raw
    // data landed, stage slot free) and ONE thread arrives on the leader's exp_full. With every expansion warp
    // polling and arriving on the mbarriers themselves (30 arrivals per stage on one barrier word that the
    // MMA thread polls) an arrive took ~400 clocks.
    // (Releasing the first three steps of a stage slot early - a second commit per stage, the expansion warps
    // refilling that half while the rest is still being multiplied - was measured and is slower: 296 vs 264
    // clocks per step; the extra barrier round costs more than the overlap gains.)
    constexpr int EXP_THREADS = G::EXP_WARPS * 32;
    int stage = 0, e = 0;
    uint32_t raw_phase = 0;
    RingPos<G::DONE> freed;                      // stage g - EXP, whose MMAs must have completed
    auto stage_begin = [&](int g) {
      const long long t0 = PROF ? clock64() : 0;
      if (tid == 0) {
        if (!(P.debug & 256)) umma::mbar_wait_a(raw_full_a + stage * 8, raw_phase);
        if (g >= G::EXP) umma::mbar_wait_a(stage_done_a + freed.slot * 8, freed.parity);
      }
      asm volatile("bar.sync 2, %0;" ::"r"(EXP_THREADS) : "memory");
      if (PROF) t_wait += clock64() - t0;
      if (g >= G::EXP) freed.advance();
    };
    auto stage_end = [&](bool wrote_smem, bool wrote_tmem) {     // operands of the stage in slot e are written
      if (wrote_tmem) {
        umma::tmem_wait_st();
        umma::fence_before_thread_sync();
      }
      const long long tf = PROF ? clock64() : 0;
      if (wrote_smem) umma::fence_proxy_async_smem();    // generic-proxy operand writes -> the tensor cores' async proxy
      const long long tb = PROF ? clock64() : 0;
      // the warp that finishes the slot last hands it over (no second barrier round: a counter in shared memory,
      // acquire-release at CTA scope, orders the other warps' fenced writes before the arrive)
      __syncwarp();
      if (lane == 0) {
        int done;
        asm volatile("atom.acq_rel.cta.shared::cta.add.s32 %0, [%1], 1;" : "=r"(done) : "r"(umma::smem_addr(&s_stage_warps[e])) : "memory");
        if (done == G::EXP_WARPS - 1) {
          s_stage_warps[e] = 0;             // next used EXP stages later, after a named-barrier round of all warps
          if (strict) umma::mbar_arrive_cluster(exp_full_leader + e * 8);
          else if (rank == 0) umma::mbar_arrive_a(exp_full_a + e * 8);
          else umma::mbar_arrive_remote(exp_full_leader + e * 8);
        }
      }
      if (PROF) {
        t_fence += tb - tf;
        t_end += clock64() - tb;
      }
      if (++stage == G::RAW) {
        stage = 0;
        raw_phase ^= 1;
      }
      if (++e == G::EXP) e = 0;
    };
    if (warp < G::A_WARPS && G::A_TMA) {
      // ---- A rows, from the raw ring: plane 1 to tensor or shared memory (plane 0 is read from the raw ring by the MMAs)
      const int r = G::SPLIT ? (warp & 3) * 32 + lane : tid;     // TMEM lane = row: warps w and w + 4 share a lane quadrant
      const int kh = G::SPLIT ? (warp >> 2) : 0;
      const uint32_t a_tmem = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + G::TMEM_A + kh * 4;
      const uint32_t raw_row = raw_a + unit_off(r, kh);
      const uint32_t exp_row = exp_a + unit_off(r, kh);
      auto derive_stage = [&](const bool full, int kw) {
        const uint32_t src0 = raw_row + stage * G::RAW_STAGE_BYTES;
        const uint32_t dst0 = exp_row + e * G::EXP_STAGE_BYTES;
        const uint32_t ta0 = a_tmem + e * G::GROUP * TMEM_A_COLS;
        uint4 u[NR];
#pragma unroll
        for (int j = 0; j < G::GROUP; ++j) {
#pragma unroll
          for (int q = 0; q < NU; ++q)
            if (full || j < kw) u[j * NU + q] = umma::lds128(src0 + j * G::RAW_A_STRIDE + q * 128);
        }
#pragma unroll
        for (int j = 0; j < G::GROUP; ++j) {
          if (full || j < kw) {
            uint4 p1[NU];
#pragma unroll
            for (int q = 0; q < NU; ++q) {
              make_unit_planes<MODE>(u[j * NU + q], p1[q]);
              if (!G::AH_TMEM) umma::sts128(dst0 + j * G::EXP_STEP_BYTES + q * 128, p1[q].x, p1[q].y, p1[q].z, p1[q].w);
            }
            if (G::AH_TMEM) {
              if (NU == 2) {
                const uint32_t v[8] = {p1[0].x, p1[0].y, p1[0].z, p1[0].w, p1[NU - 1].x, p1[NU - 1].y, p1[NU - 1].z, p1[NU - 1].w};
                umma::tmem_st_32x8(ta0 + j * TMEM_A_COLS, v);
              } else {
                umma::tmem_st_32x4(ta0 + j * TMEM_A_COLS, p1[0].x, p1[0].y, p1[0].z, p1[0].w);
              }
            }
          }
        }
      };
      for (int g = 0; g < n_groups; ++g) {
        stage_begin(g);
        if (g >= G::EXP && G::AH_TMEM) umma::fence_after_thread_sync();   // the MMAs that read these TMEM columns have completed
        const int kw = (P.debug & (16 | 512)) ? 0 : min(G::GROUP, n_steps - g * G::GROUP);
        if (kw == G::GROUP) derive_stage(true, kw);
        else derive_stage(false, kw);
        stage_end(!G::AH_TMEM, G::AH_TMEM);     // tensor-memory stores need no proxy fence
      }
    } else if (warp < G::A_WARPS) {
      // ---- A rows: stream -> registers (two stages ahead) -> plane 0 to shared memory, plane 1 to shared or tensor memory
      const int r = G::SPLIT ? (warp & 3) * 32 + lane : tid;     // TMEM lane = row: warps w and w + 4 share a lane quadrant
      const int kh = G::SPLIT ? (warp >> 2) : 0;
      const uint32_t a_tmem = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + G::TMEM_A + kh * 4;
      const uint32_t exp_row = exp_a + unit_off(r, kh);
      const int64_t a_sample = a_row0 + r;
      const int64_t step_bytes = step_stride * 16;
      const char* a_ptr = reinterpret_cast<const char*>(P.stream + s_begin * step_stride + (a_sample >> 3) * 16 + (a_sample & 7) + kh * 8);
      uint4 u0[NR], u1[NR];
      auto load_rows = [&](int g, uint4 (&u)[NR]) {
        const char* p = a_ptr + static_cast<int64_t>(g) * G::GROUP * step_bytes;
        if ((g + 1) * G::GROUP <= n_steps) {
#pragma unroll
          for (int j = 0; j < G::GROUP; ++j) {
#pragma unroll
            for (int q = 0; q < NU; ++q) u[j * NU + q] = __ldg(reinterpret_cast<const uint4*>(p + j * step_bytes) + q * 8);
          }
        } else {
#pragma unroll
          for (int j = 0; j < G::GROUP; ++j) {
#pragma unroll
            for (int q = 0; q < NU; ++q) {
              u[j * NU + q] = make_uint4(0, 0, 0, 0);
              if (g * G::GROUP + j < n_steps) u[j * NU + q] = __ldg(reinterpret_cast<const uint4*>(p + j * step_bytes) + q * 8);
            }
          }
        }
      };
      auto write_stage = [&](const bool full, int kw, uint4 (&u)[NR]) {
        const uint32_t dst0 = exp_row + e * G::EXP_STAGE_BYTES;
        const uint32_t ta0 = a_tmem + e * G::GROUP * TMEM_A_COLS;
#pragma unroll
        for (int j = 0; j < G::GROUP; ++j) {
          if (full || j < kw) {
            const uint32_t dst = dst0 + j * G::EXP_STEP_BYTES;
            uint4 p1[NU];
#pragma unroll
            for (int q = 0; q < NU; ++q) {
              const uint4& x = u[j * NU + q];
              make_unit_planes<MODE>(x, p1[q]);
              umma::sts128(dst + q * 128, x.x, x.y, x.z, x.w);                          // x / t
              if (!G::AH_TMEM) umma::sts128(dst + G::A_PLANE_BYTES + q * 128, p1[q].x, p1[q].y, p1[q].z, p1[q].w);   // h / v
            }
            if (G::AH_TMEM) {
              if (NU == 2) {
                const uint32_t v[8] = {p1[0].x, p1[0].y, p1[0].z, p1[0].w, p1[NU - 1].x, p1[NU - 1].y, p1[NU - 1].z, p1[NU - 1].w};
                umma::tmem_st_32x8(ta0 + j * TMEM_A_COLS, v);
              } else {
                umma::tmem_st_32x4(ta0 + j * TMEM_A_COLS, p1[0].x, p1[0].y, p1[0].z, p1[0].w);
              }
            }
          }
        }
      };
      auto do_stage = [&](int g, uint4 (&u)[NR]) {
        stage_begin(g);
        if (g >= G::EXP && G::AH_TMEM) umma::fence_after_thread_sync();   // the MMAs that read these TMEM columns have completed
        const int kw = (P.debug & 16) ? 0 : min(G::GROUP, n_steps - g * G::GROUP);
        if (kw == G::GROUP) write_stage(true, kw, u);    // inlined twice: the full-stage copy has no per-step predicates
        else write_stage(false, kw, u);
        stage_end(true, G::AH_TMEM);
        if (!(P.debug & 32)) load_rows(g + 2, u);
      };
      load_rows(0, u0);
      load_rows(1, u1);
      for (int g = 0; g < n_groups; g += 2) {
        do_stage(g, u0);
        if (g + 1 < n_groups) do_stage(g + 1, u1);
      }
    } else {
      // ---- B rows: raw ring -> registers -> plane 1 to shared memory (plane 0 is read from the raw ring by the MMAs)
      const int t = tid - G::A_THREADS;
      int r = G::SPLIT ? t % JH : t;
      int kh = G::SPLIT ? t / JH : 0;
      const bool active = G::SPLIT ? (kh < 2) : (r < JH);         // the last B-row warp may have threads to spare
      if (!active) {
        r = 0;
        kh = 0;
      }
      const uint32_t raw_row = raw_a + G::RAW_B_BASE + unit_off(r, kh);
      const uint32_t exp_row = exp_a + unit_off(r, kh) + G::A_PLANES_SMEM * G::A_PLANE_BYTES;
      auto expand_stage = [&](const bool full, int kw) {
        const uint32_t src0 = raw_row + stage * G::RAW_STAGE_BYTES;
        const uint32_t dst0 = exp_row + e * G::EXP_STAGE_BYTES;
        uint4 u[NR];
#pragma unroll
        for (int j = 0; j < G::GROUP; ++j) {
#pragma unroll
          for (int q = 0; q < NU; ++q)
            if (full || j < kw) u[j * NU + q] = umma::lds128(src0 + j * G::RAW_B_STRIDE + q * 128);
        }
#pragma unroll
        for (int j = 0; j < G::GROUP; ++j) {
#pragma unroll
          for (int q = 0; q < NU; ++q) {
            if (full || j < kw) {
              uint4 p1;
              make_unit_planes<MODE>(u[j * NU + q], p1);
              umma::sts128(dst0 + j * G::EXP_STEP_BYTES + q * 128, p1.x, p1.y, p1.z, p1.w);
            }
          }
        }
      };
      for (int g = 0; g < n_groups; ++g) {
        stage_begin(g);
        const int kw = (P.debug & (16 | 1024)) ? 0 : min(G::GROUP, n_steps - g * G::GROUP);
        if (active) {
          if (kw == G::GROUP) expand_stage(true, kw);
          else expand_stage(false, kw);
        }
        stage_end(true, false);
      }
    }
    if (PROF && lane == 0 && (warp == 0 || warp == G::A_WARPS) && rank == ((P.debug >> 6) & 1)) {   // debug bit 6: which CTA of the pair reports
      const int o = warp == 0 ? 2 : 9;           // [3], [4]: an A-row warp waiting / total; [10], [11]: a B-row warp
      atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[o + 1]), static_cast<unsigned long long>(t_wait));
      atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[o + 2]), static_cast<unsigned long long>(clock64() - t_start));
      atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[warp == 0 ? 12 : 14]), static_cast<unsigned long long>(t_fence));   // proxy fence
      atomicAdd(reinterpret_cast<unsigned long long*>(&g_tc_prof[warp == 0 ? 13 : 15]), static_cast<unsigned long long>(t_end));     // end-of-stage barrier
    }
    const bool uniform = (MODE == 0) && !general;
    if (uniform) {
      // |h| of this CTA's 128 A rows and of all J B columns of the pair, for this word range
      for (int k = tid; k < TC2_IH + J; k += G::EXP_WARPS * 32) {
        const int64_t sample = k < TC2_IH ? a_row0 + k : b_start + (k - TC2_IH);
        s_hcount[k] = (sample < P.n_samples) ? P.hsum[static_cast<int64_t>(split) * P.hsum_ld + sample] : 0;
      }
      int vs = 0;
      for (int64_t w = w_begin + tid; w < w_end; w += G::EXP_WARPS * 32) vs += __popc(P.all_v[w]);
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, off);
      if (lane == 0 && vs) atomicAdd(&s_valid_sites, vs);
    }
    asm volatile("bar.sync 1, %0;" ::"r"(G::EXP_WARPS * 32) : "memory");

    // ===== epilogue: this CTA's 128 rows x J columns, TMEM -> registers -> counts =====
    umma::mbar_wait_a(done_a, 0);
    umma::fence_after_thread_sync();
    const int lane_grp = warp & 3;               // TMEM lanes [32 g, 32 g + 32) are visible to warps with w % 4 == g
    const int grp_rank = warp >> 2;              // rank of this warp among the warps of its lane group
    const int grp_size = (lane_grp < (G::EXP_WARPS & 3)) ? (G::EXP_WARPS / 4 + 1) : (G::EXP_WARPS / 4);
    const int i_local = lane_grp * 32 + lane;
    const int64_t i = a_row0 + i_local;
    const int hi_cnt = uniform ? s_hcount[i_local] : 0;
    const int v_sites = s_valid_sites;
    const int64_t plane = P.n_samples * P.ld;
    int32_t* V = P.counts;
    int32_t* D1 = P.counts + plane;
    int32_t* D2 = P.counts + 2 * plane;
    for (int chunk = grp_rank; chunk < J / 16; chunk += grp_size) {
      const uint32_t taddr = tmem + (static_cast<uint32_t>(lane_grp * 32) << 16) + chunk * 16;
      uint32_t pv[16], qv[16];
      umma::tmem_ld_32x16(taddr + 0 * J, pv);
      umma::tmem_ld_32x16(taddr + 1 * J, qv);
      umma::tmem_wait_ld();
      if (i < P.n_samples) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int j_local = chunk * 16 + c;
          const int64_t j = b_start + j_local;
          if (j >= P.n_samples || j < i) continue;       // each unordered pair once: row <= column
          const int p = __float2int_rn(__uint_as_float(pv[c])), q = __float2int_rn(__uint_as_float(qv[c]));
          if (MODE == 0) {
            const int xx = p, hh = q;
            const int d2 = (hh - xx) >> 1;
            const int d1 = uniform ? (hi_cnt + s_hcount[TC2_IH + j_local] - 2 * hh + d2) : -((xx + hh) >> 1);
            if (uniform) atomicAdd(V + i * P.ld + j, v_sites);
            atomicAdd(D1 + i * P.ld + j, d1);
            atomicAdd(D2 + i * P.ld + j, d2);
            if (j != i) {
              if (uniform) atomicAdd(V + j * P.ld + i, v_sites);
              atomicAdd(D1 + j * P.ld + i, d1);
              atomicAdd(D2 + j * P.ld + i, d2);
            }
          } else {
            const int v = q, d1 = q - p;                 // VV, VV - TT
            atomicAdd(V + i * P.ld + j, v);
            atomicAdd(D1 + i * P.ld + j, d1);
            if (j != i) {
              atomicAdd(V + j * P.ld + i, v);
              atomicAdd(D1 + j * P.ld + i, d1);
            }
          }
        }
      }
    }
    umma::fence_before_thread_sync();
  }
  __syncwarp();
  umma::cluster_sync();        // neither CTA leaves (or frees tensor memory) while its peer may still use its memory
  if (warp == 0) umma::tmem_dealloc_2cta(tmem, 512);
}

constexpr int64_t tc_align(int64_t x) { return (x + 255) / 256 * 256; }

struct TcModePlan {
  int n_jtiles, pairs;
  int64_t splits, words_per_split;
};
struct TcPlan {
  int64_t n_local, n_tiles, n_tiles_padded, hsum_ld;
  int n_itiles;
  int cfg;        // tile / pipeline geometry of the CTA-pair kernel this call uses (Tc2Cfg)
  TcModePlan m[2];
  int64_t n_steps, nib_bytes, stream1_off, refv_off, flags_off, hsum_off, total;
};

int tc_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      sms <= 0)
    sms = 148;
  return sms;
}

// optional split timing of a pf_pair_counts_tc call (pf_pair_counts_tc_timing): events around the conversion and the GEMM launches
bool g_tc_timing = false;
cudaEvent_t g_tc_ev[3] = {nullptr, nullptr, nullptr};

// which GEMM kernel pf_pair_counts_tc launches: 1 = one CTA per 128 x 192 tile pair, 2 = a CTA pair (cluster of
// two, tcgen05.mma.cta_group::2) per 256 x 240 tile pair; g_tc_cfg = pipeline geometry of the pair kernel
int g_tc_impl = 2;
int g_tc_cfg = 0;

// CTA pairs of the default pair-kernel geometry the device runs at once (B200: 74, one per TPC); -1 when the
// query itself fails (no device: plans are still computed), 0 when the device cannot co-schedule the cluster
int tc2_cluster_slots() {
  thread_local int cached_dev = -1, cached = -1;   // per host thread: worker threads of one process drive one device each
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (dev == cached_dev) return cached;
  using G = Tc2Geo<0>;
  auto kernel = pair_counts_tc2_kernel<0, 0, false>;
  int n = -1;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(G::SMEM)) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2, 1, 1);
    cfg.blockDim = dim3(G::THREADS, 1, 1);
    cfg.dynamicSmemBytes = G::SMEM;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) n = -1;
  }
  cudaGetLastError();   // a failed query must not poison later calls
  cached_dev = dev;
  cached = n;
  return n;
}
// the GEMM kernel that will run: the CTA-pair kernel unless it was switched off or the device cannot run clusters of two
int tc_effective_impl() { return (g_tc_impl == 2 && tc2_cluster_slots() != 0) ? 2 : 1; }

// The pair kernel's cost model (see tc_plan): waves x (steps per item + per-item overhead), in pipeline steps
double tc2_model_steps(int64_t pairs, int64_t n_local, int slots, int64_t* best_splits) {
  const double overhead_steps = 1400.0;
  const double steps = static_cast<double>((n_local + 1) / 2);
  double best = 1e300;
  int64_t splits = 1;
  for (int64_t c = 1; c <= 64; ++c) {
    const double waves = static_cast<double>(pf_ceil_div(c * pairs, static_cast<int64_t>(slots)));
    const double cost = waves * (steps / static_cast<double>(c) + overhead_steps);
    if (cost < best * 0.999) {
      best = cost;
      splits = c;
    }
  }
  if (best_splits) *best_splits = splits;
  return best;
}
TcPlan tc_plan(int64_t n_samples, int64_t n_local, int sms) {
  TcPlan L;
  L.n_local = n_local;
  L.n_tiles = pf_ceil_div(n_samples, PF_TILE);
  const bool pair_kernel = tc_effective_impl() == 2;
  L.n_itiles = static_cast<int>(pf_ceil_div(n_samples, pair_kernel ? TC2_I : TC_I));
  const int widths[2] = {TcCfg<0>::J, TcCfg<1>::J};
  L.n_tiles_padded = 2 * static_cast<int64_t>(L.n_itiles);
  L.cfg = g_tc_cfg;
  if (pair_kernel) {
    const int slots = tc2_cluster_slots();
    sms = slots > 0 ? slots : sms / 2;           // work items run on CTA pairs
    if (sms < 1) sms = 1;
    // A blocks of 256 rows; the B tiles of block ti start at its first row, so the last one ends before n + J
    const int64_t rows = static_cast<int64_t>(TC2_I) * L.n_itiles;
    const int J2 = tc2_cols(L.cfg);
    L.n_tiles_padded = pf_ceil_div(rows > n_samples + J2 ? rows : n_samples + J2, PF_TILE);
  }
  for (int mode = 0; mode < 2; ++mode) {
    TcModePlan& M = L.m[mode];
    const int J = widths[mode];
    M.pairs = 0;
    if (pair_kernel) {
      const int J2 = tc2_cols(L.cfg);
      M.n_jtiles = static_cast<int>(pf_ceil_div(n_samples, J2));
      for (int ti = 0; ti < L.n_itiles; ++ti)
        M.pairs += static_cast<int>(pf_ceil_div(n_samples - static_cast<int64_t>(TC2_I) * ti, J2));
    } else {
      M.n_jtiles = static_cast<int>(pf_ceil_div(n_samples, J));
      const int64_t pf_tiles = static_cast<int64_t>(M.n_jtiles) * (J / PF_TILE);
      if (pf_tiles > L.n_tiles_padded) L.n_tiles_padded = pf_tiles;      // whole J tiles of operand stream
      for (int ti = 0; ti < L.n_itiles; ++ti) M.pairs += M.n_jtiles - (TC_I * ti) / J;
    }
    // split-K: ~20 waves of one CTA per SM, at least 64 words per CTA, boundaries multiples of HSUM_RUN.
    // All CTAs do the same work, so the time is waves x (work per CTA): among the next few split counts
    // take the one whose last wave is fullest (e.g. 208 tile pairs on 148 SMs: 15 splits = 21.08 waves,
    // 17 splits = 23.9 waves of smaller CTAs, 4 % less time).
    int64_t splits = pf_ceil_div(static_cast<int64_t>(sms) * 20, M.pairs);
    {
      double best = 1e30;
      int64_t best_s = splits;
      for (int64_t c = splits; c < splits + 12; ++c) {
        const double cost = static_cast<double>(pf_ceil_div(c * M.pairs, static_cast<int64_t>(sms))) / static_cast<double>(c);
        if (cost < best * 0.995) {
          best = cost;
          best_s = c;
        }
      }
      splits = best_s;
    }
    if (pair_kernel) {
      // CTA-pair kernel: every work item pays ~0.19 ms outside its main loop (allocation, two cluster barriers, the
      // drain of the last stage and 160 K scattered int32 atomics per CTA), as much as ~1,400 pipeline steps, and
      // 18 splits of 3,000 x 10M spent 14 % of the launch there. Time = waves x (steps per item + that overhead):
      // take the split count that minimises it (3,000 x 10M: 3 splits = 3.97 waves instead of 18 = 23.8).
      tc2_model_steps(M.pairs, n_local, sms, &splits);
    }
    const int64_t max_splits = pf_ceil_div(n_local, 64);
    if (splits > max_splits) splits = max_splits;
    const int64_t min_splits = pf_ceil_div(n_local, MAX_WORDS_PER_CTA);   // fp32 accumulators stay exact
    if (splits < min_splits) splits = min_splits;
    if (splits < 1) splits = 1;
    M.words_per_split = pf_ceil_div(pf_ceil_div(n_local, splits), HSUM_RUN) * HSUM_RUN;
    if (M.words_per_split < HSUM_RUN) M.words_per_split = HSUM_RUN;   // empty word range: no division by zero below
    M.splits = pf_ceil_div(n_local, M.words_per_split);
    if (M.splits < 1) M.splits = 1;
  }
  L.hsum_ld = L.n_tiles_padded * PF_TILE;
  L.n_steps = (n_local + 1) / 2;
  L.nib_bytes = tc_align(L.n_tiles_padded * L.n_steps * CHUNK_BYTES);
  L.stream1_off = L.nib_bytes;          // second operand stream (t), written only with sample-specific missingness
  L.refv_off = 2 * L.nib_bytes;
  L.flags_off = L.refv_off + tc_align(n_local * 4);
  L.hsum_off = L.flags_off + 256;
  L.total = L.hsum_off + tc_align(L.m[0].splits * L.hsum_ld * 4);
  return L;
}

template <int MODE>
cudaError_t tc_launch(const TcParams& P, int64_t grid, cudaStream_t s) {
  using G = TcGeo<MODE>;
  auto kernel = g_tc_prof_enabled ? pair_counts_tc_kernel<MODE, true> : pair_counts_tc_kernel<MODE, false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(G::SMEM));
  if (e != cudaSuccess) return e;
  pf_count_launch();
  kernel<<<static_cast<unsigned>(grid), G::THREADS, G::SMEM, s>>>(P);
  return cudaGetLastError();
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*TcEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TcEncodeTiledFn tc_encode_fn() {
  static const TcEncodeTiledFn fn = [] {   // (initialised once, also when several host threads arrive together)
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<TcEncodeTiledFn>(p);
    return static_cast<TcEncodeTiledFn>(nullptr);
  }();
  return fn;
}
// the operand stream as [step][row group][256 B], box = `steps` x `row_groups` x 256 B
bool tc_make_tensor_map(CUtensorMap* tm, const void* stream, int64_t n_tiles_padded, int64_t n_steps, int row_groups, int steps) {
  TcEncodeTiledFn fn = tc_encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[3] = {256, static_cast<cuuint64_t>(n_tiles_padded * 8), static_cast<cuuint64_t>(n_steps)};
  const cuuint64_t strides[2] = {256, static_cast<cuuint64_t>(n_tiles_padded) * CHUNK_BYTES};
  const cuuint32_t box[3] = {256, static_cast<cuuint32_t>(row_groups), static_cast<cuuint32_t>(steps)};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(stream), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int MODE, int CFG>
cudaError_t tc2_launch_cfg(const TcParams& P, int64_t items, cudaStream_t s) {
  using G = Tc2Geo<CFG>;
  auto kernel = g_tc_prof_enabled ? pair_counts_tc2_kernel<MODE, CFG, true> : pair_counts_tc2_kernel<MODE, CFG, false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(G::SMEM));
  if (e != cudaSuccess) return e;
  pf_count_launch();
  kernel<<<static_cast<unsigned>(2 * items), G::THREADS, G::SMEM, s>>>(P);   // __cluster_dims__(2): one pair per item
  return cudaGetLastError();
}
template <int MODE>
cudaError_t tc2_launch(const TcParams& P, int64_t items, cudaStream_t s, int cfg) {
  switch (cfg) {
    case 1: return tc2_launch_cfg<MODE, 1>(P, items, s);
    case 2: return tc2_launch_cfg<MODE, 2>(P, items, s);
    case 3: return tc2_launch_cfg<MODE, 3>(P, items, s);
    case 4: return tc2_launch_cfg<MODE, 4>(P, items, s);
    case 5: return tc2_launch_cfg<MODE, 5>(P, items, s);
    default: return tc2_launch_cfg<MODE, 0>(P, items, s);
  }
}

}  // namespace

// Diagnostic / A-B switch: impl 1 = single-CTA GEMM kernel (128 x 192 tiles), 2 = CTA-pair kernel (256 x 240
// tiles, default); cfg = tile / pipeline geometry of the pair kernel (Tc2Cfg 0..5). Affects later pf_pair_counts_tc* calls
// (the workspace layout depends on it: query pf_pair_counts_tc_work_bytes again).
PF_API int pf_pair_counts_tc_set_impl(int impl, int cfg) {
  if ((impl != 1 && impl != 2) || cfg < 0 || cfg >= TC2_CFGS) return pf_fail("pf_pair_counts_tc_set_impl: bad arguments");
  g_tc_impl = impl;
  g_tc_cfg = cfg;
  return PF_OK;
}

PF_API int64_t pf_pair_counts_tc_work_bytes(int64_t n_samples, int64_t n_words_range) {
  if (n_samples <= 0 || n_words_range <= 0) return 256;   // an empty range needs no scratch (non-zero: callers allocate it)
  return tc_plan(n_samples, n_words_range, tc_sm_count()).total;
}

PF_API int pf_pair_counts_tc(const uint32_t* d_planes, int64_t n_samples, int64_t n_words, int64_t word_begin,
                             int64_t word_end, int32_t* d_counts, int64_t ld, void* d_work, int64_t work_bytes,
                             int allow_general, int* h_used, void* stream) {
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

  if (g_tc_timing) PF_CUDA(cudaEventRecord(g_tc_ev[0], s));
  PF_CUDA(cudaMemsetAsync(flags, 0, 256, s));
  PF_CUDA(cudaMemsetAsync(hsum, 0, static_cast<size_t>(L.m[0].splits * L.hsum_ld) * 4, s));
  {
    const int64_t n_runs = pf_ceil_div(n_local, HSUM_RUN);
    pf_count_launch();
    // A/B switches: debug bit 2048 = the one-sample-per-thread kernel, bit 4096 = that kernel with the PRMT expansion
    // (default: four samples per thread, PRMT expansion)
    const bool narrow = (g_tc_debug & (2048 | 4096)) != 0, lut = (g_tc_debug & 4096) == 0;
    auto conv0 = !narrow ? planes_to_e2m1_x4_kernel<0> : lut ? planes_to_e2m1_kernel<0, true> : planes_to_e2m1_kernel<0, false>;
    auto conv1 = !narrow ? planes_to_e2m1_x4_kernel<1> : lut ? planes_to_e2m1_kernel<1, true> : planes_to_e2m1_kernel<1, false>;
    const int64_t threads = L.n_tiles_padded * n_runs * (narrow ? 64 : 16);
    conv0<<<static_cast<unsigned>(pf_ceil_div(threads, 256)), 256, 0, s>>>(
        d_planes, n_words, word_begin, n_local, n_samples, L.n_tiles, L.n_tiles_padded, nib4, ref_v, flags, hsum,
        L.m[0].words_per_split, L.hsum_ld);
    PF_CUDA(cudaGetLastError());
    if (allow_general) {
      // returns at once unless the first pass found sample-specific missingness
      pf_count_launch();
      conv1<<<static_cast<unsigned>(pf_ceil_div(threads, 256)), 256, 0, s>>>(
          d_planes, n_words, word_begin, n_local, n_samples, L.n_tiles, L.n_tiles_padded,
          reinterpret_cast<uint4*>(w + L.stream1_off), ref_v, flags, hsum, L.m[0].words_per_split, L.hsum_ld);
      PF_CUDA(cudaGetLastError());
    }
  }
  if (!allow_general) {
    // the caller wants to fall back to the popcount kernel: the decision needs the flag on the host
    int h_flags = 0;
    PF_CUDA(cudaMemcpyAsync(&h_flags, flags, sizeof(int), cudaMemcpyDeviceToHost, s));
    PF_CUDA(cudaStreamSynchronize(s));
    if (h_flags != 0) return PF_OK;  // *h_used stays 0, nothing was added to d_counts
  }

  if (g_tc_timing) PF_CUDA(cudaEventRecord(g_tc_ev[1], s));
  TcParams P;
  P.stream = nib4;
  P.n_tiles_padded = L.n_tiles_padded;
  P.all_v = ref_v;
  P.n_local = n_local;
  P.n_steps = L.n_steps;
  P.n_samples = n_samples;
  P.counts = d_counts;
  P.ld = ld;
  P.n_itiles = L.n_itiles;
  P.hsum = hsum;
  P.hsum_ld = L.hsum_ld;
  P.flags = flags;
  P.debug = g_tc_debug;
  // mode 1 is launched unconditionally (it returns at once unless the converter found sample-specific
  // missingness), so that the call never waits for the device
  for (int mode = 0; mode < (allow_general ? 2 : 1); ++mode) {
    const TcModePlan& M = L.m[mode];
    P.stream = mode == 0 ? nib4 : reinterpret_cast<const uint4*>(w + L.stream1_off);
    P.n_jtiles = M.n_jtiles;
    P.n_tilepairs = M.pairs;
    P.words_per_split = static_cast<int>(M.words_per_split);
    const int64_t grid = M.splits * M.pairs;
    if (grid > 0x3fffffffLL) return pf_fail("pf_pair_counts_tc: grid too large");
    if (tc_effective_impl() == 2 && tc2_tmap(L.cfg)) {
      const int J2 = tc2_cols(L.cfg), G2 = tc2_group(L.cfg);
      if (!tc_make_tensor_map(&P.tm_a, P.stream, L.n_tiles_padded, L.n_steps, TC2_IH / 8, G2) ||
          !tc_make_tensor_map(&P.tm_b, P.stream, L.n_tiles_padded, L.n_steps, J2 / 16, G2))
        return pf_fail_code(PF_ERR_CUDA, "pf_pair_counts_tc: cuTensorMapEncodeTiled failed or is unavailable");
    }
    if (tc_effective_impl() == 2) PF_CUDA(mode == 0 ? tc2_launch<0>(P, grid, s, L.cfg) : tc2_launch<1>(P, grid, s, L.cfg));
    else PF_CUDA(mode == 0 ? tc_launch<0>(P, grid, s) : tc_launch<1>(P, grid, s));
  }
  if (g_tc_timing) PF_CUDA(cudaEventRecord(g_tc_ev[2], s));
  *h_used = 1;
  return PF_OK;
  PF_TRY_END
}

// Diagnostic: enable != 0 makes later pf_pair_counts_tc calls record CUDA events on their stream around the
// operand-stream conversion and around the GEMM launch(es); with h_ms2 != NULL the call first waits for the
// last recorded call and returns h_ms2[0] = conversion ms, h_ms2[1] = GEMM ms (both launches when two ran).
PF_API int pf_pair_counts_tc_timing(int enable, float* h_ms2) {
  PF_TRY_BEGIN
  if (h_ms2) {
    if (!g_tc_ev[0]) return pf_fail("pf_pair_counts_tc_timing: timing was never enabled");
    PF_CUDA(cudaEventSynchronize(g_tc_ev[2]));
    PF_CUDA(cudaEventElapsedTime(&h_ms2[0], g_tc_ev[0], g_tc_ev[1]));
    PF_CUDA(cudaEventElapsedTime(&h_ms2[1], g_tc_ev[1], g_tc_ev[2]));
  }
  if (enable && !g_tc_ev[0])
    for (int i = 0; i < 3; ++i) PF_CUDA(cudaEventCreate(&g_tc_ev[i]));
  g_tc_timing = enable != 0;
  return PF_OK;
  PF_TRY_END
}

// After the work of a pf_pair_counts_tc call on `stream` has finished: did that call find genotypes
// missing in some samples only (1: the TT / VV launch did the work too) or not (0)? Synchronises.
PF_API int pf_pair_counts_tc_was_general(const void* d_work, int64_t n_samples, int64_t n_words_range, int* h_general,
                                         void* stream) {
  PF_TRY_BEGIN
  if (!d_work || !h_general || n_samples <= 0 || n_words_range <= 0) return pf_fail("pf_pair_counts_tc_was_general: bad arguments");
  const TcPlan L = tc_plan(n_samples, n_words_range, tc_sm_count());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int h_flags = 0;
  PF_CUDA(cudaMemcpyAsync(&h_flags, static_cast<const char*>(d_work) + L.flags_off, sizeof(int), cudaMemcpyDeviceToHost, s));
  PF_CUDA(cudaStreamSynchronize(s));
  *h_general = h_flags != 0 ? 1 : 0;
  return PF_OK;
  PF_TRY_END
}

// Diagnostic: enable role-level cycle accounting in the tensor-core kernel and read it back.
// h_out16: [0] MMA thread cycles waiting for expanded operands, [1] MMA thread total, [2..4] an A-row
// expansion warp: waiting for raw data / waiting for a free stage / total, [5] TMA producer waiting,
// [6] producer total, [8] pipeline steps, [9..11] a B-row expansion warp (as [2..4]), [12] / [13] fence
// cycles of the A-row / B-row warp; summed over CTAs.
PF_API int pf_pair_counts_tc_profile(int enable, long long* h_out16) {
  PF_TRY_BEGIN
  if (h_out16) PF_CUDA(cudaMemcpyFromSymbol(h_out16, g_tc_prof, sizeof(long long) * 16));
  long long zero[16] = {0};
  PF_CUDA(cudaMemcpyToSymbol(g_tc_prof, zero, sizeof(zero)));
  g_tc_prof_enabled = (enable & 1) != 0;
  g_tc_debug = enable >> 4;   // timing experiments, see TcParams::debug
  return PF_OK;
  PF_TRY_END
}
