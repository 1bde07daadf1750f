// Pieces shared by the NJ kernels (nj.cu: full-scan kernels; nj_lazy.cu: the bounded-scan cluster kernel).
#pragma once

#include "pf_common.cuh"

#include <math_constants.h>

namespace njc {

constexpr int NJ_NONE = 0x7fffffff;

struct NjRec {
  double q;
  int a;
  int b;
};

__device__ __forceinline__ bool rec_less(double q1, int a1, int b1, double q2, int a2, int b2) {
  return (q1 < q2) || (q1 == q2 && (a1 < a2 || (a1 == a2 && b1 < b2)));
}

// lexicographic (q, a, b) minimum over the warp with four integer REDUX steps: q through its
// order-preserving 64-bit key (high word, then low word among the lanes that hold the minimal high
// word), then a among the lanes that hold the minimal q, then b. q is never NaN here (a NaN candidate
// never replaces a thread's best); -0.0 is folded into +0.0 so that key equality is double equality.
__device__ __forceinline__ void warp_argmin(double& q, int& a, int& b) {
  const unsigned full = 0xffffffffu;
  const long long bits = __double_as_longlong(__dadd_rn(q, 0.0));
  const unsigned long long key = static_cast<unsigned long long>(bits) ^ ((bits < 0) ? 0xffffffffffffffffull : 0x8000000000000000ull);
  const unsigned hi = static_cast<unsigned>(key >> 32), lo = static_cast<unsigned>(key);
  const unsigned mhi = __reduce_min_sync(full, hi);
  const unsigned mlo = __reduce_min_sync(full, hi == mhi ? lo : 0xffffffffu);
  const bool is_min = hi == mhi && lo == mlo;
  const unsigned ma = __reduce_min_sync(full, is_min ? static_cast<unsigned>(a) : 0xffffffffu);
  const unsigned mb = __reduce_min_sync(full, (is_min && static_cast<unsigned>(a) == ma) ? static_cast<unsigned>(b) : 0xffffffffu);
  const unsigned long long mkey = (static_cast<unsigned long long>(mhi) << 32) | mlo;
  const unsigned long long mbits = (mkey & 0x8000000000000000ull) ? (mkey ^ 0x8000000000000000ull) : ~mkey;
  q = __longlong_as_double(static_cast<long long>(mbits));
  a = static_cast<int>(ma);
  b = static_cast<int>(mb);
}

// Node ids and branch lengths from the join log (one CTA, after the last iteration). During the iterations
// merge[3 t + {0, 1}] holds the joined SLOTS and len[3 t + {0, 1, 2}] holds (r_a, r_b, d_ab); the node ids are
// replayed with a slot -> node map in shared memory (s_node: n ints, s_log: one int2 per iteration) and the
// branch lengths are computed with the oracle's operation order (oracle/pf_oracle.c:pfo_nj_impl).
// d01 / d02 / d12: the distances between the last three live slots (used by thread 0 only).
// The caller has made the log visible to the CTA (__syncthreads after thread 0's last log write).
__device__ __forceinline__ void nj_finish_from_log(int32_t* merge, double* len, int n, int n_iter, double d01,
                                                   double d02, double d12, int2* s_log, int* s_node, int tid,
                                                   int nthreads) {
  for (int i = tid; i < n_iter; i += nthreads) {
    s_log[i] = make_int2(__ldcg(merge + 3 * i), __ldcg(merge + 3 * i + 1));
    const double rra = __ldcg(len + 3 * i), rrb = __ldcg(len + 3 * i + 1), dab = __ldcg(len + 3 * i + 2);
    const double n2 = static_cast<double>(n - i - 2);
    double tt = __dsub_rn(rra, rrb);
    tt = __ddiv_rn(tt, __dmul_rn(2.0, n2));
    double la = __dmul_rn(0.5, dab);
    la = __dadd_rn(la, tt);
    len[3 * i + 0] = la;
    len[3 * i + 1] = __dsub_rn(dab, la);
    len[3 * i + 2] = 0.0;
    merge[3 * i + 2] = 0;
  }
  for (int k = tid; k < n; k += nthreads) s_node[k] = k;
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < n_iter; ++i) {
      const int2 ab = s_log[i];
      const int last = n - i - 1;
      merge[3 * i + 0] = s_node[ab.x];
      merge[3 * i + 1] = s_node[ab.y];
      if (ab.y != last) s_node[ab.y] = s_node[last];
      s_node[ab.x] = n + i;
    }
    double l0 = __dmul_rn(0.5, __dsub_rn(__dadd_rn(d01, d02), d12));
    double l1 = __dmul_rn(0.5, __dsub_rn(__dadd_rn(d01, d12), d02));
    double l2 = __dmul_rn(0.5, __dsub_rn(__dadd_rn(d02, d12), d01));
    merge[3 * n_iter + 0] = s_node[0];
    merge[3 * n_iter + 1] = s_node[1];
    merge[3 * n_iter + 2] = s_node[2];
    len[3 * n_iter + 0] = l0;
    len[3 * n_iter + 1] = l1;
    len[3 * n_iter + 2] = l2;
  }
}

// row sum in the fixed order of pfo_rowsum32 (one warp per row): 32 lane-strided partials in ascending k,
// xor-butterfly. Every lane returns the sum.
__device__ __forceinline__ double nj_row_sum32(const double* __restrict__ rp, int n, int lane) {
  double acc = 0.0;
#pragma unroll 8
  for (int k = lane; k < n; k += 32) acc = __dadd_rn(acc, __ldcg(rp + k));
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
  return acc;
}

}  // namespace njc

// ---- host side of the bounded-scan cluster kernel (nj_lazy.cu), called by nj.cu
// largest matrix the cluster kernel takes (shared-memory bound) and its workspace per matrix
int64_t nj_lazy_max_n();
int64_t nj_lazy_work_bytes(int64_t n);
// Joins n_mat matrices (mat_stride doubles apart, leading dimension ld; left untouched) on one cluster each.
// d_work: n_mat x nj_lazy_work_bytes(n). d_status[m] = 0 when matrix m was joined (outputs written), != 0 when
// the caller must run a full-scan kernel on it (asymmetric input, too many row rescans, no cluster available).
// d_stats (may be null): device counters {row rescans, list re-evaluations, iterations, matrices joined}.
// Returns PF_OK also when nothing was launched (then every status is 1).
int nj_lazy_run(const double* d_dist, int64_t n, int64_t ld, int n_mat, int64_t mat_stride, int32_t* d_merge,
                double* d_len, void* d_work, int* d_status, unsigned long long* d_stats, cudaStream_t s);
