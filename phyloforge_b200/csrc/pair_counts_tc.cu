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
// Data flow per CTA (one 128 x 256 sample tile pair, a contiguous range of 32-site words):
//   "nib4" operand stream (4 bits per genotype, produced once per call by planes_to_nib4):
//   global --TMA bulk copies--> raw ring --12 expansion warps: PRMT nibble -> int8 (x and h)-->
//   canonical K-major UMMA tiles in shared memory --tcgen05.mma kind::i8 (1 thread)--> 2 x 256
//   TMEM columns (all 512) --tcgen05.ld--> epilogue --> int32 RED atomics (split-K, bit-exact).
#include "pf_common.cuh"
#include "umma.cuh"

namespace {

constexpr int TC_I = 128;            // rows of the accumulator tile (MMA M)
constexpr int TC_J = 256;            // columns (MMA N)
constexpr int TC_ROWS = TC_I + TC_J; // sample rows expanded per word
constexpr int STEP_WORDS = 2;        // words per pipeline step (one raw stage feeds one expanded stage)
constexpr int RAW_WORDS = STEP_WORDS;
constexpr int RAW_STAGES = 4;
constexpr int EXP_STAGES = 3;
constexpr int RAW_STAGE_BYTES = RAW_WORDS * TC_ROWS * 16;      // 12 KB
constexpr int EXP_WORD_BYTES = TC_ROWS * 32 * 2;               // x and h tiles of A and B for one word: 24 KB
constexpr int EXP_STAGE_BYTES = STEP_WORDS * EXP_WORD_BYTES;   // 48 KB
constexpr int EXP_AX = 0, EXP_AH = TC_I * 32, EXP_BX = 2 * TC_I * 32, EXP_BH = 2 * TC_I * 32 + TC_J * 32;
constexpr int HSUM_RUN = 16;         // words per thread of the converter; split boundaries are multiples of it
constexpr int TC_EXP_WARPS = 12;     // 384 threads: one sample row each
constexpr int TC_THREADS = (TC_EXP_WARPS + 2) * 32;
constexpr uint32_t POOL_X = 0x000100ffu;  // code 0 -> -1, 1 -> 0, 2 -> +1, 3 -> 0
constexpr uint32_t POOL_H = 0x00010001u;  // code 0 -> 1, 1 -> 0, 2 -> 1, 3 -> 0

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
};

// byte offset of the 16-byte unit (row r, K half kh) inside a canonical K-major no-swizzle tile
__device__ __forceinline__ uint32_t unit_off(int r, int kh) {
  return static_cast<uint32_t>((r >> 3) * 256 + kh * 128 + (r & 7) * 16);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
pair_counts_tc_kernel(const TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* raw = smem;                                        // RAW_STAGES x RAW_STAGE_BYTES
  uint8_t* exp = smem + RAW_STAGES * RAW_STAGE_BYTES;         // EXP_STAGES x EXP_STAGE_BYTES
  __shared__ __align__(8) uint64_t raw_full[RAW_STAGES], raw_empty[RAW_STAGES];
  __shared__ __align__(8) uint64_t exp_full[EXP_STAGES], exp_empty[EXP_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t s_tmem;
  __shared__ int s_hcount[TC_ROWS];
  __shared__ int s_valid_sites;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // work item: (tile pair, split)
  const int pair = blockIdx.x % P.n_tilepairs;
  const int split = blockIdx.x / P.n_tilepairs;
  int ti = 0, rem = pair;  // I-tile ti pairs with J-tiles tj >= ti / 2
  while (true) {
    const int cnt = P.n_jtiles - ti / 2;
    if (rem < cnt) break;
    rem -= cnt;
    ++ti;
  }
  const int tj = ti / 2 + rem;
  const int64_t w_begin = static_cast<int64_t>(split) * P.words_per_split;
  const int64_t w_end = min(P.n_local, w_begin + P.words_per_split);
  const int n_wordsteps = static_cast<int>(w_end - w_begin);
  if (n_wordsteps <= 0) return;
  const int n_rawsteps = (n_wordsteps + RAW_WORDS - 1) / RAW_WORDS;

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
  if (tid < TC_ROWS) s_hcount[tid] = 0;
  if (tid == 0) s_valid_sites = 0;
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = s_tmem;

  if (warp == TC_EXP_WARPS) {
    // ===== TMA producer: raw nib4 words of the 2 + 4 sample tiles of this tile pair =====
    if (lane == 0) {
      for (int rs = 0; rs < n_rawsteps; ++rs) {
        const int stage = rs % RAW_STAGES;
        if (rs >= RAW_STAGES) umma::mbar_wait(&raw_empty[stage], ((rs / RAW_STAGES) - 1) & 1);
        const int64_t w0 = w_begin + static_cast<int64_t>(rs) * RAW_WORDS;
        const int kw = static_cast<int>(min(static_cast<long long>(RAW_WORDS), static_cast<long long>(w_end - w0)));
        const uint32_t bytes_per_tile = static_cast<uint32_t>(kw) * 64 * 16;
        umma::mbar_expect_tx(&raw_full[stage], bytes_per_tile * 6);
        uint8_t* dst = raw + stage * RAW_STAGE_BYTES;
        // stage layout: [pf_tile 0..5][word][64 samples][16 B]; pf tiles 0,1 = I rows, 2..5 = J rows
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int64_t pft = (k < 2) ? (2 * static_cast<int64_t>(ti) + k) : (4 * static_cast<int64_t>(tj) + (k - 2));
          umma::bulk_g2s(dst + k * (RAW_WORDS * 64 * 16), P.nib4 + (pft * P.n_local + w0) * 64, bytes_per_tile,
                         &raw_full[stage]);
        }
      }
    }
  } else if (warp == TC_EXP_WARPS + 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma::idesc_i8(TC_I, TC_J);
      for (int st = 0; st < n_rawsteps; ++st) {
        const int e = st % EXP_STAGES;
        umma::mbar_wait(&exp_full[e], (st / EXP_STAGES) & 1);
        umma::fence_after_thread_sync();
        const int kw = min(STEP_WORDS, n_wordsteps - st * STEP_WORDS);
        for (int k = 0; k < kw; ++k) {
          const uint32_t base = umma::smem_addr(exp + e * EXP_STAGE_BYTES + k * EXP_WORD_BYTES);
          const uint64_t ax = umma::smem_desc_kmajor_noswizzle(base + EXP_AX, 128, 256);
          const uint64_t ah = umma::smem_desc_kmajor_noswizzle(base + EXP_AH, 128, 256);
          const uint64_t bx = umma::smem_desc_kmajor_noswizzle(base + EXP_BX, 128, 256);
          const uint64_t bh = umma::smem_desc_kmajor_noswizzle(base + EXP_BH, 128, 256);
          const uint32_t acc = (st > 0 || k > 0) ? 1u : 0u;
          umma::mma_i8_ss(tmem, ax, bx, idesc, acc);            // P -> columns [0, 256)
          umma::mma_i8_ss(tmem + TC_J, ah, bh, idesc, acc);     // Q -> columns [256, 512)
        }
        umma::commit(&exp_empty[e]);
      }
      umma::commit(&done_bar);
    }
  } else {
    // ===== expansion warps: thread = one sample row (0..127 I rows, 128..383 J rows) =====
    const int row = tid;                       // 0 .. 383
    const int pf = row >> 6, l = row & 63;     // raw stage sub-tile and sample inside it
    const bool is_a = row < TC_I;
    const int r = is_a ? row : row - TC_I;
    const uint32_t off_x = is_a ? EXP_AX : EXP_BX;
    const uint32_t off_h = is_a ? EXP_AH : EXP_BH;
    for (int st = 0; st < n_rawsteps; ++st) {
      const int stage = st % RAW_STAGES;
      const int e = st % EXP_STAGES;
      if (lane == 0) {
        umma::mbar_wait(&raw_full[stage], (st / RAW_STAGES) & 1);
        if (st >= EXP_STAGES) umma::mbar_wait(&exp_empty[e], ((st / EXP_STAGES) - 1) & 1);
      }
      __syncwarp();
      const int kw = min(STEP_WORDS, n_wordsteps - st * STEP_WORDS);
      const uint8_t* src = raw + stage * RAW_STAGE_BYTES + pf * (RAW_WORDS * 64 * 16) + l * 16;
      for (int k = 0; k < kw; ++k) {
        const uint4 nb = *reinterpret_cast<const uint4*>(src + k * (64 * 16));
        uint8_t* dst = exp + e * EXP_STAGE_BYTES + k * EXP_WORD_BYTES;
        const uint32_t q[4] = {nb.x, nb.y, nb.z, nb.w};
        uint32_t xs[8], hs[8];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          xs[2 * m] = __byte_perm(POOL_X, 0u, q[m]);
          xs[2 * m + 1] = __byte_perm(POOL_X, 0u, q[m] >> 16);
          hs[2 * m] = __byte_perm(POOL_H, 0u, q[m]);
          hs[2 * m + 1] = __byte_perm(POOL_H, 0u, q[m] >> 16);
        }
        *reinterpret_cast<uint4*>(dst + off_x + unit_off(r, 0)) = make_uint4(xs[0], xs[1], xs[2], xs[3]);
        *reinterpret_cast<uint4*>(dst + off_x + unit_off(r, 1)) = make_uint4(xs[4], xs[5], xs[6], xs[7]);
        *reinterpret_cast<uint4*>(dst + off_h + unit_off(r, 0)) = make_uint4(hs[0], hs[1], hs[2], hs[3]);
        *reinterpret_cast<uint4*>(dst + off_h + unit_off(r, 1)) = make_uint4(hs[4], hs[5], hs[6], hs[7]);
      }
      umma::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        umma::mbar_arrive(&exp_full[e]);
        umma::mbar_arrive(&raw_empty[stage]);
      }
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
    // all expansion warps meet before the epilogue (named barrier 1, 384 threads)
    asm volatile("bar.sync 1, %0;" ::"r"(TC_EXP_WARPS * 32) : "memory");

    // ===== epilogue: TMEM -> registers -> counts =====
    umma::mbar_wait(&done_bar, 0);
    umma::fence_after_thread_sync();
    const int lane_grp = warp & 3;               // TMEM lanes [32 g, 32 g + 32) are visible to warps with w % 4 == g
    const int col_grp = warp >> 2;               // 3 warps share a lane group: split the 8 column chunks
    const int i_local = lane_grp * 32 + lane;
    const int64_t i = static_cast<int64_t>(ti) * TC_I + i_local;
    const int hi = s_hcount[i_local];
    const int v_sites = s_valid_sites;
    const int64_t plane = P.n_samples * P.ld;
    int32_t* V = P.counts;
    int32_t* D1 = P.counts + plane;
    int32_t* D2 = P.counts + 2 * plane;
    for (int chunk = col_grp; chunk < TC_J / 32; chunk += 3) {
      uint32_t pv[32], qv[32];
      const uint32_t taddr = tmem + (static_cast<uint32_t>(lane_grp * 32) << 16) + chunk * 32;
      umma::tmem_ld_32x32(taddr, pv);
      umma::tmem_ld_32x32(taddr + TC_J, qv);
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
  L.n_tiles_padded = static_cast<int64_t>(L.n_jtiles) * (TC_J / PF_TILE);   // multiple of 4 PF tiles
  L.pairs = 0;
  for (int ti = 0; ti < L.n_itiles; ++ti) L.pairs += L.n_jtiles - ti / 2;
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
  const int64_t grid = L.splits * L.pairs;
  if (grid > 0x7fffffffLL) return pf_fail("pf_pair_counts_tc: grid too large");
  const size_t smem = static_cast<size_t>(RAW_STAGES) * RAW_STAGE_BYTES + static_cast<size_t>(EXP_STAGES) * EXP_STAGE_BYTES;
  PF_CUDA(cudaFuncSetAttribute(pair_counts_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem)));
  pair_counts_tc_kernel<<<static_cast<unsigned>(grid), TC_THREADS, smem, s>>>(P);
  PF_CUDA(cudaGetLastError());
  *h_used = 1;
  return PF_OK;
  PF_TRY_END
}
