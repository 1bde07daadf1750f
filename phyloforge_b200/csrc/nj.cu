// Neighbor joining on the GPU: one persistent cooperative kernel, n-3 dependent iterations.
//
// Stands in for the tree step the reference delegates to `treebest nj`
// (treetool/script.py:359-360) and follows oracle/pf_oracle.c:pfo_nj operation for operation
// so that the join order and the branch lengths are bit-identical:
//   0  r[a] = sum_k D[a][k] once — 32 lane-strided partials in ascending k, xor-butterfly
//   B  argmin over a < b of (n-2)*D[a][b] - r[a] - r[b]; ties -> lowest a, then lowest b
//      (a lexicographic (q, a, b) min is associative and commutative, so the reduction tree
//       does not matter)
//   C  branch lengths; row/column a <- merged node; last active slot moves into slot b;
//      row sums updated incrementally with the oracle's exact operation order
// fp64 throughout with explicit round-to-nearest intrinsics (no FMA contraction).
//
// Memory-bound scan of the live upper triangle: ~ 4 n^2 bytes per iteration from L2/HBM
// (DESIGN.md "Kernels"), plus two grid barriers per iteration.
#include "pf_common.cuh"

#include <cooperative_groups.h>
#include <math_constants.h>

namespace cg = cooperative_groups;

namespace {

constexpr int NJ_THREADS = 1024;
constexpr int NJ_MAX_BLOCKS = 1024;

struct NjRec {
  double q;
  double dab;
  int a;
  int b;
};

__device__ __forceinline__ bool rec_less(double q1, int a1, int b1, double q2, int a2, int b2) {
  return (q1 < q2) || (q1 == q2 && (a1 < a2 || (a1 == a2 && b1 < b2)));
}

__device__ __forceinline__ void warp_argmin(double& q, double& dab, int& a, int& b) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const double q2 = __shfl_xor_sync(0xffffffffu, q, off);
    const double d2 = __shfl_xor_sync(0xffffffffu, dab, off);
    const int a2 = __shfl_xor_sync(0xffffffffu, a, off);
    const int b2 = __shfl_xor_sync(0xffffffffu, b, off);
    if (rec_less(q2, a2, b2, q, a, b)) {
      q = q2; dab = d2; a = a2; b = b2;
    }
  }
}

struct NjParams {
  double* D;
  int64_t ld;
  int n;
  int32_t* merge;
  double* len;
  double* r0;   // row sums, double-buffered across iterations
  double* r1;
  int32_t* node;
  NjRec* rec;
  int r_in_smem;  // the row sums of the live nodes fit in dynamic shared memory
};

constexpr int NJ_SEG = 512;  // columns of one (row, segment) work item of the argmin scan

__global__ void __launch_bounds__(NJ_THREADS, 1)
nj_kernel(const NjParams P) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double s_r[];
  __shared__ NjRec s_rec[NJ_THREADS / 32];
  __shared__ NjRec s_best;

  double* __restrict__ D = P.D;
  const int64_t ld = P.ld;
  const int n = P.n;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int warps_per_block = NJ_THREADS / 32;
  const int gwarp = blockIdx.x * warps_per_block + warp;
  const int nwarps = gridDim.x * warps_per_block;
  const int gthread = blockIdx.x * NJ_THREADS + tid;
  const int nthreads = gridDim.x * NJ_THREADS;

  // initial row sums in the fixed order of pfo_rowsum32 (one warp per row)
  for (int s = gthread; s < n; s += nthreads) P.node[s] = s;
  for (int row = gwarp; row < n; row += nwarps) {
    const double* rp = D + static_cast<int64_t>(row) * ld;
    double acc = 0.0;
#pragma unroll 8
    for (int k = lane; k < n; k += 32) acc = __dadd_rn(acc, __ldcg(rp + k));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
    if (lane == 0) P.r0[row] = acc;
  }
  grid.sync();

  double* cur = P.r0;
  double* nxt = P.r1;
  int na = n;
  int t = 0;
  while (na > 3) {
    // ---- B: Q argmin over the upper triangle, work items = (row, 512-column segment)
    const double* rs = cur;
    if (P.r_in_smem) {
      for (int k = tid; k < na; k += NJ_THREADS) s_r[k] = __ldcg(cur + k);
      __syncthreads();
      rs = s_r;
    }
    const double n2 = static_cast<double>(na - 2);
    double bq = CUDART_INF;
    double bd = 0.0;
    int ba = 0x7fffffff, bb = 0x7fffffff;
    const int nseg = (na + NJ_SEG - 1) / NJ_SEG;
    const int items = na * nseg;
    for (int item = gwarp; item < items; item += nwarps) {
      const int a = item / nseg;
      const int sgm = item - a * nseg;
      const int b_lo = max(a + 1, sgm * NJ_SEG);
      const int b_hi = min(na, (sgm + 1) * NJ_SEG);
      if (b_lo >= b_hi) continue;
      const double* rp = D + static_cast<int64_t>(a) * ld;
      const double ra = P.r_in_smem ? rs[a] : __ldcg(rs + a);
#pragma unroll 8
      for (int b = b_lo + lane; b < b_hi; b += 32) {
        const double d = __ldcg(rp + b);
        const double rb = P.r_in_smem ? rs[b] : __ldcg(rs + b);
        double q = __dmul_rn(n2, d);
        q = __dsub_rn(q, ra);
        q = __dsub_rn(q, rb);
        if (rec_less(q, a, b, bq, ba, bb)) {
          bq = q; bd = d; ba = a; bb = b;
        }
      }
    }
    warp_argmin(bq, bd, ba, bb);
    if (lane == 0) s_rec[warp] = NjRec{bq, bd, ba, bb};
    __syncthreads();
    if (warp == 0) {
      NjRec v = s_rec[lane];  // warps_per_block == 32
      warp_argmin(v.q, v.dab, v.a, v.b);
      if (lane == 0) P.rec[blockIdx.x] = v;
    }
    grid.sync();

    // ---- C: every CTA reduces the per-CTA records to the same winner, then updates
    if (warp == 0) {
      NjRec v{CUDART_INF, 0.0, 0x7fffffff, 0x7fffffff};
      for (int k = lane; k < static_cast<int>(gridDim.x); k += 32) {
        const NjRec* g = P.rec + k;
        const double q2 = __ldcg(&g->q);
        const double d2 = __ldcg(&g->dab);
        const int a2 = __ldcg(&g->a);
        const int b2 = __ldcg(&g->b);
        if (rec_less(q2, a2, b2, v.q, v.a, v.b)) v = NjRec{q2, d2, a2, b2};
      }
      warp_argmin(v.q, v.dab, v.a, v.b);
      if (lane == 0) s_best = v;
    }
    __syncthreads();
    int wa = s_best.a, wb = s_best.b;
    double dab = s_best.dab;
    if (wa == 0x7fffffff) {  // no finite candidate (NaN/inf input): join slots 0 and 1
      wa = 0; wb = 1;
      dab = __ldcg(D + 1);
    }
    const int last = na - 1;
    const double* ra_p = D + static_cast<int64_t>(wa) * ld;
    const double* rb_p = D + static_cast<int64_t>(wb) * ld;
    const double* rl_p = D + static_cast<int64_t>(last) * ld;
    const bool move = (wb != last);
    for (int k = gthread; k < na; k += nthreads) {
      if (k == wb) continue;
      if (k == wa) {
        // merged node u: row sum in closed form, branch lengths, bookkeeping
        const double rra = __ldcg(cur + wa), rrb = __ldcg(cur + wb);
        double ru = __dadd_rn(rra, rrb);
        ru = __dsub_rn(ru, __dmul_rn(static_cast<double>(na), dab));
        nxt[wa] = __dmul_rn(0.5, ru);
        double tt = __dsub_rn(rra, rrb);
        tt = __ddiv_rn(tt, __dmul_rn(2.0, n2));
        double la = __dmul_rn(0.5, dab);
        la = __dadd_rn(la, tt);
        const double lb = __dsub_rn(dab, la);
        P.merge[3 * t + 0] = __ldcg(P.node + wa);
        P.merge[3 * t + 1] = __ldcg(P.node + wb);
        P.merge[3 * t + 2] = 0;
        P.len[3 * t + 0] = la;
        P.len[3 * t + 1] = lb;
        P.len[3 * t + 2] = 0.0;
        if (move) P.node[wb] = __ldcg(P.node + last);
        P.node[wa] = n + t;
        continue;
      }
      const double dak = __ldcg(ra_p + k), dbk = __ldcg(rb_p + k);
      double d = __dadd_rn(dak, dbk);
      d = __dsub_rn(d, dab);
      d = __dmul_rn(0.5, d);
      double rk = __dsub_rn(__ldcg(cur + k), dak);
      rk = __dsub_rn(rk, dbk);
      rk = __dadd_rn(rk, d);
      if (move && k == last) {
        // the last live slot moves into slot b: its distance to u and its row sum go there
        D[static_cast<int64_t>(wa) * ld + wb] = d;
        D[static_cast<int64_t>(wb) * ld + wa] = d;
        nxt[wb] = rk;
      } else {
        D[static_cast<int64_t>(wa) * ld + k] = d;
        D[static_cast<int64_t>(k) * ld + wa] = d;
        nxt[k] = rk;
        if (move) {
          const double mv = __ldcg(rl_p + k);
          D[static_cast<int64_t>(wb) * ld + k] = mv;
          D[static_cast<int64_t>(k) * ld + wb] = mv;
        }
      }
    }
    grid.sync();
    double* tmp = cur; cur = nxt; nxt = tmp;
    --na;
    ++t;
  }

  if (gthread == 0) {
    const double d01 = __ldcg(D + 1), d02 = __ldcg(D + 2), d12 = __ldcg(D + ld + 2);
    double l0 = __dmul_rn(0.5, __dsub_rn(__dadd_rn(d01, d02), d12));
    double l1 = __dmul_rn(0.5, __dsub_rn(__dadd_rn(d01, d12), d02));
    double l2 = __dmul_rn(0.5, __dsub_rn(__dadd_rn(d02, d12), d01));
    P.merge[3 * t + 0] = __ldcg(P.node + 0);
    P.merge[3 * t + 1] = __ldcg(P.node + 1);
    P.merge[3 * t + 2] = __ldcg(P.node + 2);
    P.len[3 * t + 0] = l0;
    P.len[3 * t + 1] = l1;
    P.len[3 * t + 2] = l2;
  }
}

constexpr int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
constexpr int64_t NJ_MAX_SMEM_R = 200 * 1024;

}  // namespace

PF_API int64_t pf_nj_work_bytes(int64_t n) {
  if (n < 0) return 0;
  return 2 * align_up(8 * n, 256) + align_up(4 * n, 256) + align_up(sizeof(NjRec) * NJ_MAX_BLOCKS, 256);
}

PF_API int pf_nj(double* d_dist, int64_t n, int64_t ld, int32_t* d_merge, double* d_len, void* d_work,
                 void* stream) {
  PF_TRY_BEGIN
  if (!d_dist || !d_merge || !d_len || !d_work || n < 3 || ld < n || n > 0x3fffffff)
    return pf_fail("pf_nj: bad arguments (n=%lld, ld=%lld; n must be >= 3)", (long long)n, (long long)ld);
  int dev = 0, sms = 0, coop = 0;
  PF_CUDA(cudaGetDevice(&dev));
  PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PF_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return pf_fail_code(PF_ERR_CUDA, "pf_nj: device does not support cooperative launch");
  const int r_in_smem = (8 * n <= NJ_MAX_SMEM_R) ? 1 : 0;
  const size_t smem = r_in_smem ? static_cast<size_t>(align_up(8 * n, 16)) : 0;
  PF_CUDA(cudaFuncSetAttribute(nj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  PF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nj_kernel, NJ_THREADS, smem));
  if (per_sm < 1) return pf_fail_code(PF_ERR_CUDA, "pf_nj: kernel does not fit on an SM");
  // enough CTAs to give every warp about two (row, segment) items; fewer CTAs make the two grid
  // barriers of an iteration cheaper when the matrix is small
  const int64_t items = n * pf_ceil_div(n, NJ_SEG);
  int64_t blocks = pf_ceil_div(items, (NJ_THREADS / 32) * 2);
  if (blocks > sms) blocks = sms;
  if (blocks > NJ_MAX_BLOCKS) blocks = NJ_MAX_BLOCKS;
  if (blocks < 1) blocks = 1;

  char* w = static_cast<char*>(d_work);
  NjParams P;
  P.D = d_dist;
  P.ld = ld;
  P.n = static_cast<int>(n);
  P.merge = d_merge;
  P.len = d_len;
  P.r0 = reinterpret_cast<double*>(w);
  P.r1 = reinterpret_cast<double*>(w + align_up(8 * n, 256));
  P.node = reinterpret_cast<int32_t*>(w + 2 * align_up(8 * n, 256));
  P.rec = reinterpret_cast<NjRec*>(w + 2 * align_up(8 * n, 256) + align_up(4 * n, 256));
  P.r_in_smem = r_in_smem;
  void* args[] = {&P};
  PF_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(nj_kernel), dim3(static_cast<unsigned>(blocks)),
                                      dim3(NJ_THREADS), args, smem, static_cast<cudaStream_t>(stream)));
  return PF_OK;
  PF_TRY_END
}
