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
// The scan of the live upper triangle (4 n^2 bytes per iteration from L2 / HBM) is the only part
// that scales; everything else is a chain of dependent memory round trips, so the kernel is built to
// keep that chain short. Two kernels:
//
// nj_fast_kernel (n <= NJ_FAST_MAX_N, the row sums and the new node's distances fit in shared memory)
//   ONE grid-wide exchange per iteration and no grid barrier:
//   * every CTA keeps the full row-sum vector, the slot -> physical-row map and the distances of the
//     node merged last (s_du) in shared memory and updates them redundantly (same operations, same bits);
//   * CTAs publish their argmin candidate as four 64-bit words {tag | 32 data bits} (each word is
//     written atomically, the tag validates it: no flag, no fence round trip between data and flag)
//     and poll everybody else's words; the exchange doubles as the barrier that orders this
//     iteration's scan before the update writes;
//   * matrix rows are never overwritten while somebody may still read them: the merged node's row
//     goes to a spare physical row (the row freed one iteration earlier), slot b takes over the last
//     slot's physical row by remapping. Column entries are written in place. These writes become
//     visible with the NEXT exchange; the scan in between takes the affected entries from s_du /
//     the last slot's column instead ("pending" update).
// nj_kernel (any n): row sums in global memory when they do not fit, two grid barriers per iteration.
#include "nj_common.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

using njc::NJ_NONE;
using njc::NjRec;
using njc::rec_less;
using njc::warp_argmin;

constexpr int NJ_MAX_BLOCKS = 1024;

// optional phase accounting (pf_nj_profile): nanoseconds of CTA 0 per phase, summed over iterations
__device__ unsigned long long g_nj_prof[8];
bool g_nj_prof_enabled = false;
__device__ __forceinline__ unsigned long long nj_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define NJ_MARK(slot)                                                  \
  if (PROF && blockIdx.x == 0 && threadIdx.x == 0) {                   \
    const unsigned long long now_ = nj_now();                          \
    g_nj_prof[slot] += now_ - t_prev;                                  \
    t_prev = now_;                                                     \
  }

// initial row sums in the fixed order of pfo_rowsum32 (one warp per row, rows spread over the grid)
__device__ __forceinline__ void nj_initial_row_sums(const double* __restrict__ D, int64_t ld, int n, int gwarp,
                                                    int nwarps, int lane, double* r_out) {
  for (int row = gwarp; row < n; row += nwarps) {
    const double acc = njc::nj_row_sum32(D + static_cast<int64_t>(row) * ld, n, lane);
    if (lane == 0) r_out[row] = acc;
  }
}

// The update of the previous iteration whose global-memory writes are not visible yet (fast kernel):
// slot a = the merged node (its distances are in s_du), slot b = took over slot `last` (-1: it did not).
struct NjPending {
  int a, b, last;
};

// ---- argmin scan. Work items = (row a, segment of 32 U columns) of the upper triangle, enumerated
// row-block major so that every thread meets its candidates in ascending (a, b) order: a strict
// `q < best` then implements the lexicographic tie rule inside a thread. The U loads of an item are
// issued together, and the loads of a warp's NEXT item are issued before the candidates of the current
// one are evaluated (two register buffers), so the memory round trip of an item is hidden behind the
// arithmetic of the previous one. Items that lie wholly inside the triangle and touch no pending
// column (almost all of them) load without per-element predicates.
template <int U, bool VIRT, bool R_SMEM>
__device__ __forceinline__ void nj_scan(const double* __restrict__ D, int64_t ld, const double* spare, int n_phys,
                                        const int* s_phys, const double* rs, const double* s_du, int na, double n2,
                                        NjPending pend, int gwarp, int nwarps, int lane, double& bq, int& ba,
                                        int& bb) {
  constexpr int SEG = 32 * U;
  const int nseg = (na + SEG - 1) / SEG;
  int total = 0;
  for (int k = 0; k < nseg; ++k) total += min(SEG, na - k * SEG) * (nseg - k);
  if (gwarp >= total) return;
  // enumeration state; decode() is called with increasing item numbers
  int rb = 0, base = 0;
  int segs = nseg;
  int cnt = min(SEG, na) * segs;
  // kind: 0 = interior item (no predicates), 1 = general, 2 = the row of the node merged last (in s_du)
  auto decode = [&](int item, int& a, int& c0, int& kind, const double*& rp) {
    while (item >= base + cnt) {
      base += cnt;
      ++rb;
      segs = nseg - rb;
      cnt = min(SEG, na - rb * SEG) * segs;
    }
    const int local = item - base;
    const int ar = local / segs;
    a = rb * SEG + ar;
    c0 = (rb + (local - ar * segs)) * SEG;
    const bool pend_in_seg = VIRT && ((pend.a >= c0 && pend.a < c0 + SEG) || (pend.b >= c0 && pend.b < c0 + SEG));
    int prow = a;
    if (VIRT) prow = s_phys[(a == pend.b) ? pend.last : a];
    rp = (VIRT && prow == n_phys) ? spare : D + static_cast<int64_t>(prow) * ld;
    kind = (VIRT && a == pend.a) ? 2 : ((c0 > a && c0 + SEG <= na && !pend_in_seg) ? 0 : 1);
  };
  auto load = [&](double (&dv)[U], int a, int c0, int kind, const double* rp) {
    if (kind == 0) {
      const double* p = rp + c0 + lane;
#pragma unroll
      for (int u = 0; u < U; ++u) dv[u] = __ldcg(p + u * 32);
    } else if (kind == 1) {
      // the column of the slot that took over the last slot is still in the last slot's column
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int b = c0 + u * 32 + lane;
        const int col = (VIRT && b == pend.b) ? pend.last : b;
        dv[u] = (b > a && b < na) ? __ldcg(rp + col) : CUDART_INF;
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int b = c0 + u * 32 + lane;
        dv[u] = (b > a && b < na) ? s_du[b] : CUDART_INF;
      }
    }
  };
  auto compute = [&](double (&dv)[U], int a, int c0, int kind) {
    if (VIRT && kind == 1 && pend.a > a && pend.a >= c0 && pend.a < c0 + SEG) {
      const double dua = s_du[a];      // the column of the node merged last
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (c0 + u * 32 + lane == pend.a) dv[u] = dua;
    }
    const double ra = R_SMEM ? rs[a] : __ldcg(rs + a);
    int iu = -1;     // index u of this item's best candidate, if it beats bq
    if (kind == 0) {
      const double* rsp = rs + c0 + lane;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double rbv = R_SMEM ? rsp[u * 32] : __ldcg(rsp + u * 32);
        double q = __dmul_rn(n2, dv[u]);
        q = __dsub_rn(q, ra);
        q = __dsub_rn(q, rbv);
        const bool take = q < bq;
        bq = take ? q : bq;
        iu = take ? u : iu;
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int bc = min(c0 + u * 32 + lane, na - 1);
        const double rbv = R_SMEM ? rs[bc] : __ldcg(rs + bc);
        double q = __dmul_rn(n2, dv[u]);
        q = __dsub_rn(q, ra);
        q = __dsub_rn(q, rbv);
        const bool take = q < bq;     // +inf for masked entries and NaN never win
        bq = take ? q : bq;
        iu = take ? u : iu;
      }
    }
    if (iu >= 0) {
      ba = a;
      bb = c0 + iu * 32 + lane;
    }
  };

  double dv0[U], dv1[U];
  int a0, c00, k0, a1 = 0, c01 = 0, k1 = 0;
  const double *p0, *p1 = nullptr;
  int item = gwarp;
  decode(item, a0, c00, k0, p0);
  load(dv0, a0, c00, k0, p0);
  while (true) {
    item += nwarps;
    const bool more1 = item < total;
    if (more1) {
      decode(item, a1, c01, k1, p1);
      load(dv1, a1, c01, k1, p1);
    }
    compute(dv0, a0, c00, k0);
    if (!more1) break;
    item += nwarps;
    const bool more0 = item < total;
    if (more0) {
      decode(item, a0, c00, k0, p0);
      load(dv0, a0, c00, k0, p0);
    }
    compute(dv1, a1, c01, k1);
    if (!more0) break;
  }
}

// ===================================================================== fast kernel
constexpr int NJF_THREADS = 512;
constexpr int NJF_U = 8;             // 256-column scan items, two in flight per warp
constexpr int NJF_U_WIDE = 16;       // 512-column items for matrices of NJF_WIDE_FROM nodes or more
constexpr int NJF_WIDE_FROM = 5000;
constexpr int NJF_K = 8;             // slots per thread and batch of the update
constexpr int64_t NJ_FAST_MAX_N = 11000;   // 20 n bytes of shared memory
constexpr int NJF_SHARDS = 8;        // arrival counters of the exchange, one 128-byte line each
constexpr int NJF_MAX_CTAS = 256;
constexpr int64_t NJ_XCHG_BYTES = 32768;   // >= 128 NJF_SHARDS + 2 x NJF_MAX_CTAS x 16, and the general kernel's records

struct NjFastParams {
  double* D;
  int64_t ld;
  int n;
  int32_t* merge;             // during the iterations: log of (slot a, slot b); node ids at the end
  double* len;                // during the iterations: log of (r_a, r_b, d_ab); branch lengths at the end
  double* r0;                 // initial row sums
  double* spare;              // one extra matrix row (physical row index n)
  unsigned* arrive;           // [NJF_SHARDS][32] arrival counters (monotonic)
  unsigned long long* recs;   // [2][CTAs of the matrix][2]: {q bits, a | b << 32}
  // batch: the grid is n_mat groups of ctas_per_mat CTAs, group g joins matrix g; all pointers above
  // belong to matrix 0 and move by these strides per matrix
  int n_mat, ctas_per_mat;
  int64_t mat_stride;         // doubles between distance matrices
  int64_t work_stride;        // bytes between the workspaces (r0, spare, arrive, recs)
  const int* status;          // per matrix, may be null: 0 = already joined by the bounded-scan kernel (nj_lazy.cu), skip it
};

struct NjWinner {
  int a, b;
  double ra, rb;
};

// per-matrix view of the parameters: for a single matrix the fields stay in the constant bank
// (a rebased copy in registers costs the scan ~5 %)
template <bool BATCH> struct NjFastView;
template <> struct NjFastView<false> {
  const NjFastParams& p;
  __device__ NjFastView(const NjFastParams& q, int) : p(q) {}
  __device__ __forceinline__ double* D() const { return p.D; }
  __device__ __forceinline__ int32_t* merge() const { return p.merge; }
  __device__ __forceinline__ double* len() const { return p.len; }
  __device__ __forceinline__ double* r0() const { return p.r0; }
  __device__ __forceinline__ double* spare() const { return p.spare; }
  __device__ __forceinline__ unsigned* arrive() const { return p.arrive; }
  __device__ __forceinline__ unsigned long long* recs() const { return p.recs; }
};
template <> struct NjFastView<true> {
  const NjFastParams& p;
  int64_t d_off, out_off, w_off;
  __device__ NjFastView(const NjFastParams& q, int mat)
      : p(q), d_off(mat * q.mat_stride), out_off(3 * static_cast<int64_t>(q.n - 2) * mat), w_off(mat * q.work_stride) {}
  template <typename T> __device__ __forceinline__ T* w(T* base) const {
    return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + w_off);
  }
  __device__ __forceinline__ double* D() const { return p.D + d_off; }
  __device__ __forceinline__ int32_t* merge() const { return p.merge + out_off; }
  __device__ __forceinline__ double* len() const { return p.len + out_off; }
  __device__ __forceinline__ double* r0() const { return w(p.r0); }
  __device__ __forceinline__ double* spare() const { return w(p.spare); }
  __device__ __forceinline__ unsigned* arrive() const { return w(p.arrive); }
  __device__ __forceinline__ unsigned long long* recs() const { return w(p.recs); }
};

// U = scan item width in 32-column units. 256-column items (U = 8) are best up to a few thousand nodes; from
// ~5,000 on, where the scan is bandwidth-bound, 512-column items amortise the per-item overhead better (whole runs
// at n = 6,000: 87.6 ms with 256 columns, 82.3 ms with 512; n = 3,000: 22.0 vs 23.1). One width per launch: a
// kernel that switched per iteration held three instantiations of the scan, spilled and was slower everywhere.
template <bool PROF, bool BATCH, int U = NJF_U>
__global__ void __launch_bounds__(NJF_THREADS, 1)
nj_fast_kernel(const NjFastParams P_in) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char nj_dyn[];
  __shared__ NjRec s_rec[NJF_THREADS / 32];
  __shared__ NjWinner s_win;

  const int n = P_in.n;
  // this CTA's matrix (batch) and its rank among the CTAs of that matrix
  const int n_cta = BATCH ? P_in.ctas_per_mat : static_cast<int>(gridDim.x);
  const int mat = BATCH ? static_cast<int>(blockIdx.x) / n_cta : 0;
  const int cta = static_cast<int>(blockIdx.x) - mat * n_cta;
  const NjFastView<BATCH> V(P_in, mat);
  double* __restrict__ D = V.D();
  const int64_t ld = P_in.ld;
  const int n_pad = (n + 1) & ~1;
  double* s_r = reinterpret_cast<double*>(nj_dyn);
  double* s_du = s_r + n_pad;
  int* s_phys = reinterpret_cast<int*>(s_du + n_pad);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int WARPS = NJF_THREADS / 32;
  const int gwarp = cta * WARPS + warp;
  const int nwarps = n_cta * WARPS;
  const int gthread = cta * NJF_THREADS + tid;
  const int nthreads = n_cta * NJF_THREADS;

  if (P_in.status && P_in.status[mat] == 0) {   // nothing to do for this matrix (every CTA still meets the grid barrier)
    grid.sync();
    return;
  }
  for (int k = gthread; k < NJF_SHARDS * 32; k += nthreads) V.arrive()[k] = 0u;
  nj_initial_row_sums(D, ld, n, gwarp, nwarps, lane, V.r0());
  grid.sync();
  for (int k = tid; k < n; k += NJF_THREADS) {
    s_r[k] = __ldcg(V.r0() + k);
    s_phys[k] = k;
  }
  __syncthreads();
  unsigned long long t_prev = PROF ? nj_now() : 0;

  // exchange bookkeeping of warp 0: lane s < NJF_SHARDS watches arrival counter s, which
  // (blockIdx % NJF_SHARDS == s) CTAs bump once per iteration
  unsigned* my_arrive = V.arrive() + (cta % NJF_SHARDS) * 32;
  const unsigned shard_ctas = (lane < NJF_SHARDS) ? static_cast<unsigned>((n_cta - lane + NJF_SHARDS - 1) / NJF_SHARDS) : 0u;

  NjPending pend{-1, -1, -1};
  int spare_row = n;     // physical row index of the spare row (n = the workspace row)
  int na = n, t = 0;
  while (na > 3) {
    // ---- scan
    const double n2 = static_cast<double>(na - 2);
    double bq = CUDART_INF;
    int ba = NJ_NONE, bb = NJ_NONE;
    nj_scan<U, true, true>(D, ld, V.spare(), n, s_phys, s_r, s_du, na, n2, pend, gwarp, nwarps, lane, bq, ba, bb);
    warp_argmin(bq, ba, bb);
    if (lane == 0) s_rec[warp] = NjRec{bq, ba, bb};
    __syncthreads();
    NJ_MARK(0)   // scan + waiting for this CTA's slowest warp
    // ---- exchange: warp 0 publishes this CTA's candidate, waits until every CTA has, reads them all.
    // It doubles as the barrier that orders this iteration's scan before the update writes.
    if (warp == 0) {
      NjRec v = (lane < WARPS) ? s_rec[lane] : NjRec{CUDART_INF, NJ_NONE, NJ_NONE};
      warp_argmin(v.q, v.a, v.b);
      unsigned long long* recs = V.recs() + static_cast<int64_t>(t & 1) * n_cta * 2;
      if (lane == 0) {
        const unsigned long long w0 = static_cast<unsigned long long>(__double_as_longlong(v.q));
        const unsigned long long w1 = static_cast<unsigned>(v.a) | (static_cast<unsigned long long>(static_cast<unsigned>(v.b)) << 32);
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(recs + 2 * cta), "l"(w0), "l"(w1) : "memory");
        // release: the record and this CTA's update writes of the previous iteration (ordered before
        // this point by the __syncthreads above) are visible before the arrival is
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(my_arrive) : "memory");
      }
      NJ_MARK(1)
      const unsigned want = shard_ctas * static_cast<unsigned>(t + 1);
      bool ok;
      do {
        unsigned seen = want;
        if (lane < NJF_SHARDS)
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(V.arrive() + lane * 32) : "memory");
        ok = __all_sync(0xffffffffu, seen >= want);
      } while (!ok);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");   // acquire: reads below are ordered after the arrivals were seen
      NJ_MARK(2)
      double wq = CUDART_INF;
      int wa = NJ_NONE, wb = NJ_NONE;
      for (int c0 = 0; c0 < n_cta; c0 += 32 * 8) {
        unsigned long long w[8][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c0 + j * 32 + lane;
          w[j][0] = 0x7ff0000000000000ull;   // +inf
          w[j][1] = 0x7fffffff7fffffffull;
          if (c < n_cta)
            asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[j][0]), "=l"(w[j][1]) : "l"(recs + 2 * c) : "memory");
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const double q2 = __longlong_as_double(static_cast<long long>(w[j][0]));
          const int a2 = static_cast<int>(static_cast<unsigned>(w[j][1]));
          const int b2 = static_cast<int>(static_cast<unsigned>(w[j][1] >> 32));
          const bool take = rec_less(q2, a2, b2, wq, wa, wb);
          wq = take ? q2 : wq;
          wa = take ? a2 : wa;
          wb = take ? b2 : wb;
        }
      }
      warp_argmin(wq, wa, wb);
      if (lane == 0) {
        if (wa == NJ_NONE) {   // no finite candidate (NaN / inf input): join slots 0 and 1
          wa = 0; wb = 1;
        }
        s_win = NjWinner{wa, wb, s_r[wa], s_r[wb]};
      }
    }
    __syncthreads();
    NJ_MARK(3)   // exchange (includes waiting for the slowest CTA)

    // ---- update (every CTA: shared-memory state; the owner of a slot: global memory)
    const int wa = s_win.a, wb = s_win.b;
    const double rra = s_win.ra, rrb = s_win.rb;
    const int last = na - 1;
    const bool move = (wb != last);
    const int pwa = s_phys[wa], pwb = s_phys[wb], plast = s_phys[last];
    const double* row_a = (pwa == n) ? V.spare() : D + static_cast<int64_t>(pwa) * ld;
    const double* row_b = (pwb == n) ? V.spare() : D + static_cast<int64_t>(pwb) * ld;
    double* row_u = (spare_row == n) ? V.spare() : D + static_cast<int64_t>(spare_row) * ld;
    const double dab = __ldcg(row_a + wb);
    // the global-memory side of slot k belongs to one CTA: contiguous, 32-aligned ranges of slots
    const int own_span = ((na + n_cta - 1) / n_cta + 31) & ~31;
    const int own_lo = cta * own_span, own_hi = own_lo + own_span;
    for (int k0 = tid; k0 < na; k0 += NJF_THREADS * NJF_K) {
      double dak[NJF_K], dbk[NJF_K], mv[NJF_K];
      int prow[NJF_K];
#pragma unroll
      for (int j = 0; j < NJF_K; ++j) {
        const int k = k0 + j * NJF_THREADS;
        const bool live = k < na && k != wa && k != wb;
        const bool own = live && k >= own_lo && k < own_hi;
        dak[j] = live ? __ldcg(row_a + k) : 0.0;
        dbk[j] = live ? __ldcg(row_b + k) : 0.0;
        prow[j] = own ? s_phys[k] : -1;
        mv[j] = 0.0;
        if (own && move && k != last) {
          const double* rk_p = (prow[j] == n) ? V.spare() : D + static_cast<int64_t>(prow[j]) * ld;
          mv[j] = __ldcg(rk_p + last);
        }
      }
#pragma unroll
      for (int j = 0; j < NJF_K; ++j) {
        const int k = k0 + j * NJF_THREADS;
        if (k >= na || k == wb) continue;
        if (k == wa) {
          double ru = __dadd_rn(rra, rrb);
          ru = __dsub_rn(ru, __dmul_rn(static_cast<double>(na), dab));
          s_r[wa] = __dmul_rn(0.5, ru);
          s_du[wa] = 0.0;
          continue;
        }
        double d = __dadd_rn(dak[j], dbk[j]);
        d = __dsub_rn(d, dab);
        d = __dmul_rn(0.5, d);
        double rk = __dsub_rn(s_r[k], dak[j]);
        rk = __dsub_rn(rk, dbk[j]);
        rk = __dadd_rn(rk, d);
        const int slot = (move && k == last) ? wb : k;   // the last live slot moves into slot b
        s_r[slot] = rk;
        s_du[slot] = d;
        if (prow[j] >= 0) {
          double* rk_p = (prow[j] == n) ? V.spare() : D + static_cast<int64_t>(prow[j]) * ld;
          row_u[slot] = d;         // row of the merged node, in the spare physical row
          rk_p[wa] = d;            // column a of this node's row
          if (move && k != last) rk_p[wb] = mv[j];   // column b <- column last
        }
      }
    }
    if (gthread == 0) {
      // log of the join; node ids and branch lengths are derived after the last iteration
      V.merge()[3 * t + 0] = wa;
      V.merge()[3 * t + 1] = wb;
      V.len()[3 * t + 0] = rra;
      V.len()[3 * t + 1] = rrb;
      V.len()[3 * t + 2] = dab;
    }
    __syncthreads();
    if (tid == 0) {
      s_phys[wa] = spare_row;
      if (move) s_phys[wb] = plast;
    }
    // (the scan below does not read s_phys[wa] / s_phys[wb]: those rows are "pending")
    pend = NjPending{wa, move ? wb : -1, last};
    spare_row = pwa;       // free from the next iteration on: everybody has read it by the next exchange
    --na;
    ++t;
    NJ_MARK(4)   // update
  }

  if (cta != 0) return;
  // ---- CTA 0 of the matrix: node ids and branch lengths from the log
  double d01 = 0.0, d02 = 0.0, d12 = 0.0;
  if (tid == 0) {
    // the last three nodes; the final update may still be pending, so read through it
    auto vread = [&](int x, int y) -> double {   // x < y < 3
      if (x == pend.a) return s_du[y];
      if (y == pend.a) return s_du[x];
      const int pr = s_phys[x];   // thread 0 wrote s_phys itself: already remapped
      const double* rp = (pr == n) ? V.spare() : D + static_cast<int64_t>(pr) * ld;
      return __ldcg(rp + ((y == pend.b) ? pend.last : y));
    };
    d01 = vread(0, 1);
    d02 = vread(0, 2);
    d12 = vread(1, 2);
  }
  __syncthreads();   // the log written by thread 0 is visible to the CTA; s_r / s_du / s_phys are free
  njc::nj_finish_from_log(V.merge(), V.len(), n, t, d01, d02, d12, reinterpret_cast<int2*>(nj_dyn), s_phys, tid,
                          NJF_THREADS);   // the log (8 bytes per iteration) over s_r, the node map over s_phys
}

// ===================================================================== general kernel
constexpr int NJ_THREADS = 512;
constexpr int NJ_U = 8;   // 256-column scan items, two in flight per warp

struct NjRecG {   // record of the general kernel: d_ab travels with the candidate, because the
  double q;       // update overwrites D[a][b] in place while other CTAs may not have read it yet
  double dab;
  int a;
  int b;
};

struct NjParams {
  double* D;
  int64_t ld;
  int n;
  int32_t* merge;
  double* len;
  double* r0;   // row sums, double-buffered across iterations
  double* r1;
  int32_t* node;
  NjRecG* rec;
  int r_in_smem;  // the row sums of the live nodes fit in dynamic shared memory
};

template <bool PROF>
__global__ void __launch_bounds__(NJ_THREADS, 1)
nj_kernel(const NjParams P) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double s_r[];
  __shared__ NjRec s_rec[NJ_THREADS / 32];
  __shared__ NjRecG s_best;

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
  nj_initial_row_sums(D, ld, n, gwarp, nwarps, lane, P.r0);
  grid.sync();
  unsigned long long t_prev = PROF ? nj_now() : 0;

  double* cur = P.r0;
  double* nxt = P.r1;
  int na = n;
  int t = 0;
  const NjPending none{-1, -1, -1};
  while (na > 3) {
    // ---- B: Q argmin over the upper triangle
    if (P.r_in_smem) {
      for (int k = tid; k < na; k += NJ_THREADS) s_r[k] = __ldcg(cur + k);
      __syncthreads();
    }
    const double n2 = static_cast<double>(na - 2);
    double bq = CUDART_INF;
    int ba = NJ_NONE, bb = NJ_NONE;
    if (P.r_in_smem)
      nj_scan<NJ_U, false, true>(D, ld, nullptr, 0, nullptr, s_r, nullptr, na, n2, none, gwarp, nwarps, lane, bq, ba,
                                 bb);
    else
      nj_scan<NJ_U, false, false>(D, ld, nullptr, 0, nullptr, cur, nullptr, na, n2, none, gwarp, nwarps, lane, bq, ba,
                                  bb);
    warp_argmin(bq, ba, bb);
    if (lane == 0) s_rec[warp] = NjRec{bq, ba, bb};
    __syncthreads();
    if (warp == 0) {
      NjRec v = (lane < warps_per_block) ? s_rec[lane] : NjRec{CUDART_INF, NJ_NONE, NJ_NONE};
      warp_argmin(v.q, v.a, v.b);
      if (lane == 0) {
        const double dab = (v.a == NJ_NONE) ? 0.0 : __ldcg(D + static_cast<int64_t>(v.a) * ld + v.b);
        P.rec[blockIdx.x] = NjRecG{v.q, dab, v.a, v.b};
      }
    }
    NJ_MARK(0)   // scan + CTA reduction
    grid.sync();
    NJ_MARK(1)   // barrier 1 (includes waiting for the slowest CTA's scan)

    // ---- C: every CTA reduces the per-CTA records to the same winner, then updates
    if (warp == 0) {
      NjRecG v{CUDART_INF, 0.0, NJ_NONE, NJ_NONE};
      for (int k = lane; k < static_cast<int>(gridDim.x); k += 32) {
        const NjRecG* g = P.rec + k;
        const double q2 = __ldcg(&g->q);
        const double d2 = __ldcg(&g->dab);
        const int a2 = __ldcg(&g->a);
        const int b2 = __ldcg(&g->b);
        const bool take = rec_less(q2, a2, b2, v.q, v.a, v.b);
        v.q = take ? q2 : v.q;
        v.dab = take ? d2 : v.dab;
        v.a = take ? a2 : v.a;
        v.b = take ? b2 : v.b;
      }
      // (q, a, b) is unique per candidate, so d_ab can follow the winner through the shuffles
      const double q0 = v.q;
      const int a0 = v.a, b0 = v.b;
      warp_argmin(v.q, v.a, v.b);
      const unsigned src = __ballot_sync(0xffffffffu, q0 == v.q && a0 == v.a && b0 == v.b);
      const double dwin = __shfl_sync(0xffffffffu, v.dab, src ? __ffs(src) - 1 : 0);
      if (lane == 0) s_best = NjRecG{v.q, dwin, v.a, v.b};
    }
    __syncthreads();
    int wa = s_best.a, wb = s_best.b;
    double dab = s_best.dab;
    if (wa == NJ_NONE) {  // no finite candidate (NaN/inf input): join slots 0 and 1
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
      const double dak = __ldcg(ra_p + k), dbk = __ldcg(rb_p + k), rk0 = __ldcg(cur + k);
      // row and column of the last slot separately (identical for a symmetric matrix; the oracle
      // copies both, so an asymmetric input gives the same bits too)
      const double mv = (move && k != last) ? __ldcg(rl_p + k) : 0.0;
      const double mvc = (move && k != last) ? __ldcg(D + static_cast<int64_t>(k) * ld + last) : 0.0;
      double d = __dadd_rn(dak, dbk);
      d = __dsub_rn(d, dab);
      d = __dmul_rn(0.5, d);
      double rk = __dsub_rn(rk0, dak);
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
          D[static_cast<int64_t>(wb) * ld + k] = mv;
          D[static_cast<int64_t>(k) * ld + wb] = mvc;
        }
      }
    }
    NJ_MARK(2)   // winner selection + update
    grid.sync();
    NJ_MARK(3)   // barrier 2
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
int g_nj_force_general = 0;
int g_nj_no_lazy = 0;        // diagnostic: full-scan kernels only
int g_nj_lazy_always = 0;    // diagnostic: bounded scan at every size it can take
constexpr int64_t NJ_LAZY_FROM = 1536;         // single matrix
constexpr int64_t NJ_LAZY_BATCH_FROM = 1536;   // matrices of a batch

// per-matrix workspace: the full-scan kernels' part, the bounded-scan kernel's part, one status line
int64_t nj_classic_bytes(int64_t n) { return 2 * align_up(8 * n, 256) + align_up(4 * n, 256) + NJ_XCHG_BYTES; }
constexpr int64_t NJ_STATUS_BYTES = 256;

template <typename K>
int nj_launch(K kernel, int threads, size_t smem, int64_t want_blocks, int64_t max_blocks, void* params,
              cudaStream_t s) {
  int dev = 0, sms = 0;
  PF_CUDA(cudaGetDevice(&dev));
  PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = 0;
  PF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  if (per_sm < 1) return pf_fail_code(PF_ERR_CUDA, "pf_nj: kernel does not fit on an SM");
  int64_t blocks = want_blocks;
  if (blocks > sms) blocks = sms;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  void* args[] = {params};
  pf_count_launch();
  PF_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), dim3(static_cast<unsigned>(blocks)),
                                      dim3(threads), args, smem, s));
  return PF_OK;
}

}  // namespace

PF_API int64_t pf_nj_work_bytes(int64_t n) {
  if (n < 0) return 0;
  return nj_classic_bytes(n) + nj_lazy_work_bytes(n) + NJ_STATUS_BYTES;
}

namespace {

// n_mat matrices of the same size, mat_stride doubles apart; outputs 3 (n - 2) apart; workspaces
// pf_nj_work_bytes(n) apart
int nj_run(double* d_dist, int64_t n, int64_t ld, int n_mat, int64_t mat_stride, int32_t* d_merge, double* d_len,
           void* d_work, cudaStream_t s) {
  int dev = 0, coop = 0, sms = 0;
  PF_CUDA(cudaGetDevice(&dev));
  PF_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (!coop) return pf_fail_code(PF_ERR_CUDA, "pf_nj: device does not support cooperative launch");
  // workspace of n_mat matrices: [n_mat x full-scan part][n_mat x bounded-scan part][n_mat x status line]
  const int64_t work_stride = nj_classic_bytes(n);
  const int64_t out3 = 3 * (n - 2);
  char* const w_lazy = static_cast<char*>(d_work) + n_mat * work_stride;
  int* const d_status = reinterpret_cast<int*>(w_lazy + n_mat * nj_lazy_work_bytes(n));

  // bounded scan on one cluster per matrix first; it reports per matrix whether a full-scan kernel has to run
  // Measured on B200 (profiles/r02_nj_bounded_scan.md): from ~1,000 nodes on the bounded scan is at least as fast as
  // the full scan on every input tried (genotype distances of a star-like population sample are the hard case for the
  // bounds: 19.0 vs 21.9 ms with p-distances at 3,000 x 10M, 20.8 vs 22.0 ms with IBS), and pulls away with n (47 vs
  // 80 ms at 6,000, 134-165 vs 316 ms at 10,000; in batches 4.3 vs 10.2 ms per tree at 9 x 3,000). Below that the
  // cooperative kernel, which shares the SMs among the matrices of a batch, is up to 25 % faster.
  const bool lazy = n <= NJ_FAST_MAX_N && n <= nj_lazy_max_n() && nj_lazy_work_bytes(n) > 0 && !g_nj_force_general &&
                    !g_nj_no_lazy && !g_nj_prof_enabled &&
                    (g_nj_lazy_always || (n_mat == 1 ? n >= NJ_LAZY_FROM : n >= NJ_LAZY_BATCH_FROM));
  if (lazy) {
    unsigned long long* d_stats = nullptr;
    PF_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&d_stats), g_nj_prof));
    const int rc = nj_lazy_run(d_dist, n, ld, n_mat, mat_stride, d_merge, d_len, w_lazy, d_status, d_stats + 4, s);
    if (rc != PF_OK) return rc;
  }

  if (n <= NJ_FAST_MAX_N && !g_nj_force_general) {
    int done = 0;
    while (done < n_mat) {
      // as many matrices per launch as keep at least 8 CTAs each (one CTA per SM, cooperative launch)
      int batch = n_mat - done;
      if (batch > sms / 8) batch = sms / 8;
      if (batch < 1) batch = 1;
      char* w = static_cast<char*>(d_work) + done * work_stride;
      NjFastParams P;
      P.D = d_dist + done * mat_stride;
      P.ld = ld;
      P.n = static_cast<int>(n);
      P.merge = d_merge + done * out3;
      P.len = d_len + done * out3;
      P.r0 = reinterpret_cast<double*>(w);
      P.spare = reinterpret_cast<double*>(w + align_up(8 * n, 256));
      char* w_rec = w + 2 * align_up(8 * n, 256) + align_up(4 * n, 256);
      P.arrive = reinterpret_cast<unsigned*>(w_rec);
      P.recs = reinterpret_cast<unsigned long long*>(w_rec + 128 * NJF_SHARDS);
      P.n_mat = batch;
      P.mat_stride = mat_stride;
      P.work_stride = work_stride;
      P.status = lazy ? d_status + done : nullptr;
      const int64_t n_pad = (n + 1) & ~int64_t(1);
      const size_t smem = static_cast<size_t>(16 * n_pad + 4 * n_pad);
      // about four 256-column scan items per warp; fewer CTAs make the exchange cheaper for small matrices
      const int64_t items = n * pf_ceil_div(n, 32 * NJF_U) / 2 + 1;
      int64_t per_mat = pf_ceil_div(items, (NJF_THREADS / 32) * 2);
      if (per_mat > sms / batch) per_mat = sms / batch;
      if (per_mat > NJF_MAX_CTAS) per_mat = NJF_MAX_CTAS;
      if (per_mat < 1) per_mat = 1;
      P.ctas_per_mat = static_cast<int>(per_mat);
      static_assert(128 * NJF_SHARDS + 2 * NJF_MAX_CTAS * 16 <= NJ_XCHG_BYTES, "exchange area");
      static_assert(sizeof(NjRecG) * NJ_MAX_BLOCKS <= NJ_XCHG_BYTES, "record area");
      const int64_t blocks = per_mat * batch;
      int rc;
      if (batch == 1 && n >= NJF_WIDE_FROM)
        rc = g_nj_prof_enabled ? nj_launch(nj_fast_kernel<true, false, NJF_U_WIDE>, NJF_THREADS, smem, blocks, blocks, &P, s)
                               : nj_launch(nj_fast_kernel<false, false, NJF_U_WIDE>, NJF_THREADS, smem, blocks, blocks, &P, s);
      else if (batch == 1)
        rc = g_nj_prof_enabled ? nj_launch(nj_fast_kernel<true, false>, NJF_THREADS, smem, blocks, blocks, &P, s)
                               : nj_launch(nj_fast_kernel<false, false>, NJF_THREADS, smem, blocks, blocks, &P, s);
      else
        rc = g_nj_prof_enabled ? nj_launch(nj_fast_kernel<true, true>, NJF_THREADS, smem, blocks, blocks, &P, s)
                               : nj_launch(nj_fast_kernel<false, true>, NJF_THREADS, smem, blocks, blocks, &P, s);
      if (rc != PF_OK) return rc;
      done += batch;
    }
    return PF_OK;
  }
  for (int m = 0; m < n_mat; ++m) {
    char* w = static_cast<char*>(d_work) + m * work_stride;
    NjParams P;
    P.D = d_dist + m * mat_stride;
    P.ld = ld;
    P.n = static_cast<int>(n);
    P.merge = d_merge + m * out3;
    P.len = d_len + m * out3;
    P.r0 = reinterpret_cast<double*>(w);
    P.r1 = reinterpret_cast<double*>(w + align_up(8 * n, 256));
    P.node = reinterpret_cast<int32_t*>(w + 2 * align_up(8 * n, 256));
    P.rec = reinterpret_cast<NjRecG*>(w + 2 * align_up(8 * n, 256) + align_up(4 * n, 256));
    P.r_in_smem = (8 * n <= NJ_MAX_SMEM_R) ? 1 : 0;
    const size_t smem = P.r_in_smem ? static_cast<size_t>(align_up(8 * n, 16)) : 0;
    const int64_t items = n * pf_ceil_div(n, 32 * NJ_U) / 2 + 1;
    const int64_t blocks = pf_ceil_div(items, (NJ_THREADS / 32) * 2);
    const int rc = g_nj_prof_enabled ? nj_launch(nj_kernel<true>, NJ_THREADS, smem, blocks, NJ_MAX_BLOCKS, &P, s)
                                     : nj_launch(nj_kernel<false>, NJ_THREADS, smem, blocks, NJ_MAX_BLOCKS, &P, s);
    if (rc != PF_OK) return rc;
  }
  return PF_OK;
}

}  // namespace

PF_API int pf_nj(double* d_dist, int64_t n, int64_t ld, int32_t* d_merge, double* d_len, void* d_work,
                 void* stream) {
  PF_TRY_BEGIN
  if (!d_dist || !d_merge || !d_len || !d_work || n < 3 || ld < n || n > 0x3fffffff)
    return pf_fail("pf_nj: bad arguments (n=%lld, ld=%lld; n must be >= 3)", (long long)n, (long long)ld);
  return nj_run(d_dist, n, ld, 1, 0, d_merge, d_len, d_work, static_cast<cudaStream_t>(stream));
  PF_TRY_END
}

PF_API int pf_nj_batch(double* d_dist, int64_t n, int64_t ld, int n_mat, int64_t mat_stride, int32_t* d_merge,
                       double* d_len, void* d_work, void* stream) {
  PF_TRY_BEGIN
  if (!d_dist || !d_merge || !d_len || !d_work || n < 3 || ld < n || n > 0x3fffffff || n_mat < 1 ||
      (n_mat > 1 && mat_stride < n * ld))
    return pf_fail("pf_nj_batch: bad arguments (n=%lld, ld=%lld, n_mat=%d, mat_stride=%lld)", (long long)n,
                   (long long)ld, n_mat, (long long)mat_stride);
  return nj_run(d_dist, n, ld, n_mat, mat_stride, d_merge, d_len, d_work, static_cast<cudaStream_t>(stream));
  PF_TRY_END
}

// Diagnostic: phase accounting of the NJ kernel, and a switch that forces the general (two grid
// barriers per iteration) kernel for matrices the fast kernel would take. h_out8 = nanoseconds of CTA 0
// summed over iterations. Fast kernel: [0] scan, [1] CTA reduction + publish, [2] waiting for the arrivals,
// [3] reading the records + winner, [4] update.
// General kernel: [0] scan + CTA reduction, [1] grid barrier 1, [2] winner + update, [3] grid barrier 2.
// [4..7]: counters of the bounded-scan kernel (nj_lazy.cu): row rescans, list evaluations, iterations, matrices
// joined by it (0 when a full-scan kernel did the work).
// enable: bit 0 = instrumented kernels (full-scan), bit 1 = force the general kernel, bit 2 = full-scan kernels
// only (no bounded scan), bit 3 = bounded scan also for batches of small matrices. Reading resets the counters.
PF_API int pf_nj_profile(int enable, unsigned long long* h_out8) {
  PF_TRY_BEGIN
  if (h_out8) PF_CUDA(cudaMemcpyFromSymbol(h_out8, g_nj_prof, sizeof(unsigned long long) * 8));
  unsigned long long zero[8] = {0};
  PF_CUDA(cudaMemcpyToSymbol(g_nj_prof, zero, sizeof(zero)));
  g_nj_prof_enabled = (enable & 1) != 0;
  g_nj_force_general = (enable & 2) != 0;
  g_nj_no_lazy = (enable & 4) != 0;
  g_nj_lazy_always = (enable & 8) != 0;
  return PF_OK;
  PF_TRY_END
}
