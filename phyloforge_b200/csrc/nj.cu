// Neighbor joining on the GPU: one persistent cooperative kernel, n-3 dependent iterations.
//
// Stands in for the tree step the reference delegates to `treebest nj`
// (treetool/script.py:359-360) and follows oracle/pf_oracle.c:pfo_nj operation for operation
// so that the join order and the branch lengths are bit-identical:
//   A  r[a] = sum_k D[a][k]  — 32 lane-strided partials in ascending k, xor-butterfly
//   B  argmin over a < b of (n-2)*D[a][b] - r[a] - r[b]; ties -> lowest a, then lowest b
//      (a lexicographic (q, a, b) min is associative and commutative, so the reduction tree
//       does not matter)
//   C  branch lengths; row/column a <- merged node; last active slot moves into slot b
// fp64 throughout with explicit round-to-nearest intrinsics (no FMA contraction).
//
// Memory-bound scan of the live matrix: ~ (8 + 4) n^2 bytes per iteration from L2/HBM
// (DESIGN.md "Kernels"), plus three grid barriers per iteration.
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
  double* r;
  int32_t* node;
  NjRec* rec;
};

__global__ void __launch_bounds__(NJ_THREADS, 1)
nj_kernel(const NjParams P) {
  cg::grid_group grid = cg::this_grid();
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

  for (int s = gthread; s < n; s += nthreads) P.node[s] = s;
  grid.sync();

  int na = n;
  int t = 0;
  while (na > 3) {
    // ---- A: row sums in the fixed order of pfo_rowsum32
    for (int row = gwarp; row < na; row += nwarps) {
      const double* rp = D + static_cast<int64_t>(row) * ld;
      double acc = 0.0;
#pragma unroll 4
      for (int k = lane; k < na; k += 32) acc = __dadd_rn(acc, __ldcg(rp + k));
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
      if (lane == 0) P.r[row] = acc;
    }
    grid.sync();

    // ---- B: Q argmin over the upper triangle; rows paired (p, na-1-p) for balance
    const double n2 = static_cast<double>(na - 2);
    double bq = CUDART_INF;
    double bd = 0.0;
    int ba = 0x7fffffff, bb = 0x7fffffff;
    const int half = (na + 1) >> 1;
    for (int p = gwarp; p < half; p += nwarps) {
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {
        const int a = which == 0 ? p : na - 1 - p;
        if (which == 1 && a == p) break;
        const double* rp = D + static_cast<int64_t>(a) * ld;
        const double ra = __ldcg(P.r + a);
#pragma unroll 4
        for (int b = a + 1 + lane; b < na; b += 32) {
          const double d = __ldcg(rp + b);
          double q = __dmul_rn(n2, d);
          q = __dsub_rn(q, ra);
          q = __dsub_rn(q, __ldcg(P.r + b));
          if (rec_less(q, a, b, bq, ba, bb)) {
            bq = q; bd = d; ba = a; bb = b;
          }
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
      if (k == wb || (move && k == last)) continue;
      if (k == wa) {
        if (move) {
          // the moved slot's distance to the merged node
          double d = __dadd_rn(__ldcg(ra_p + last), __ldcg(rb_p + last));
          d = __dsub_rn(d, dab);
          d = __dmul_rn(0.5, d);
          D[static_cast<int64_t>(wa) * ld + wb] = d;
          D[static_cast<int64_t>(wb) * ld + wa] = d;
        }
        continue;
      }
      double d = __dadd_rn(__ldcg(ra_p + k), __ldcg(rb_p + k));
      d = __dsub_rn(d, dab);
      d = __dmul_rn(0.5, d);
      const double mv = move ? __ldcg(rl_p + k) : 0.0;
      D[static_cast<int64_t>(wa) * ld + k] = d;
      D[static_cast<int64_t>(k) * ld + wa] = d;
      if (move) {
        D[static_cast<int64_t>(wb) * ld + k] = mv;
        D[static_cast<int64_t>(k) * ld + wb] = mv;
      }
    }
    if (gthread == 0) {
      const double rra = __ldcg(P.r + wa), rrb = __ldcg(P.r + wb);
      double tt = __dsub_rn(rra, rrb);
      tt = __ddiv_rn(tt, __dmul_rn(2.0, n2));
      double la = __dmul_rn(0.5, dab);
      la = __dadd_rn(la, tt);
      const double lb = __dsub_rn(dab, la);
      P.merge[3 * t + 0] = P.node[wa];
      P.merge[3 * t + 1] = P.node[wb];
      P.merge[3 * t + 2] = 0;
      P.len[3 * t + 0] = la;
      P.len[3 * t + 1] = lb;
      P.len[3 * t + 2] = 0.0;
      if (move) P.node[wb] = P.node[last];
      P.node[wa] = n + t;
    }
    grid.sync();
    --na;
    ++t;
  }

  if (gthread == 0) {
    const double d01 = __ldcg(D + 1), d02 = __ldcg(D + 2), d12 = __ldcg(D + ld + 2);
    double l0 = __dmul_rn(0.5, __dsub_rn(__dadd_rn(d01, d02), d12));
    double l1 = __dmul_rn(0.5, __dsub_rn(__dadd_rn(d01, d12), d02));
    double l2 = __dmul_rn(0.5, __dsub_rn(__dadd_rn(d02, d12), d01));
    P.merge[3 * t + 0] = P.node[0];
    P.merge[3 * t + 1] = P.node[1];
    P.merge[3 * t + 2] = P.node[2];
    P.len[3 * t + 0] = l0;
    P.len[3 * t + 1] = l1;
    P.len[3 * t + 2] = l2;
  }
}

constexpr int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

}  // namespace

PF_API int64_t pf_nj_work_bytes(int64_t n) {
  if (n < 0) return 0;
  return align_up(8 * n, 256) + align_up(4 * n, 256) + align_up(sizeof(NjRec) * NJ_MAX_BLOCKS, 256);
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
  int per_sm = 0;
  PF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nj_kernel, NJ_THREADS, 0));
  if (per_sm < 1) return pf_fail_code(PF_ERR_CUDA, "pf_nj: kernel does not fit on an SM");
  // one warp per row is enough parallelism; fewer CTAs make the grid barriers cheaper
  int64_t blocks = pf_ceil_div(n, NJ_THREADS / 32);
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
  P.r = reinterpret_cast<double*>(w);
  P.node = reinterpret_cast<int32_t*>(w + align_up(8 * n, 256));
  P.rec = reinterpret_cast<NjRec*>(w + align_up(8 * n, 256) + align_up(4 * n, 256));
  void* args[] = {&P};
  PF_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(nj_kernel), dim3(static_cast<unsigned>(blocks)),
                                      dim3(NJ_THREADS), args, 0, static_cast<cudaStream_t>(stream)));
  return PF_OK;
  PF_TRY_END
}
