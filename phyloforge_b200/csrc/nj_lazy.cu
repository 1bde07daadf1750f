// Neighbor joining with a bounded scan: one thread-block cluster per matrix, exchange through distributed
// shared memory, the distance matrix touched only where a bound says the minimum could be.
//
// Same definition as the full-scan kernels of nj.cu (oracle/pf_oracle.c:pfo_nj_impl, the tree step the reference
// delegates to `treebest nj`, treetool/script.py:359-360): per iteration the lexicographic minimum of
//     q(a, b) = fl(fl(fl((n - 2) d_ab) - r_a) - r_b),  a < b  (slots),
// the same update arithmetic, the same join log. What differs is how the minimum is found. Write
// s_j = r_j / (n - 2); in real numbers q(i, j) / (n - 2) = (d_ij - s_j) - s_i, so inside row i the order of the
// candidates is the order of v_ij = d_ij - s_j, and a lower bound on v over a set of columns bounds q over
// that set. Every row (owned by one CTA of the cluster) keeps
//   * a LIST of up to G tracked columns (column slot j sits in group j mod G) with their exact distances,
//   * a CHAMPION, the list entry that was best when the list was last evaluated,
//   * R1 <= v of every other list entry and R2 <= v of every column that is not in the list,
// R1 / R2 in the frame of a SNAPSHOT of s (snap_j, float): d_ij - s_j(t) >= R - M(t) with
// M(t) = max_j (s_j(t) - snap_j), one number per iteration computed where the row sums are updated. Per
// iteration a CTA evaluates the champions of its rows exactly (T = its best, or the runner-up of the previous
// exchange re-evaluated in the new frame if that is better: any exact q of a live pair is a valid threshold), re-evaluates (exactly, from shared
// memory) the lists whose R1 bound does not clear T, and rescans from the matrix only rows whose R2 bound does
// not clear T. Whatever is skipped is provably worse than T by a margin far above the rounding of the bound, and
// everything that is not skipped is evaluated with the oracle's operations, so the winner and the ties are the
// oracle's. A pair is covered by the row of the node that existed first (a new node starts with an empty list; the
// other rows take it in as a new column).
//
// One exchange per iteration: every CTA writes its candidate into the shared memory of all CTAs of the cluster
// (st.shared::cluster), one cluster barrier (release / acquire, which also publishes the matrix writes of the
// update), every CTA reduces the same records. The row sums, the snapshot and the slot -> row map are replicated
// in every CTA and updated redundantly with the oracle's operation order (same bits everywhere).
//
// The kernel works on a private copy of the matrix ((n + 1) rows: the merged node goes to a spare row, as in
// nj_fast_kernel) and leaves the input alone: when the input is not symmetric, or the bounds do not prune
// (rescans over budget: degenerate inputs), it reports so and nj.cu runs the full-scan kernel on the input.
#include "nj_common.cuh"

#include <cuda_runtime.h>

namespace {

using njc::NJ_NONE;
using njc::NjRec;
using njc::rec_less;
using njc::warp_argmin;

constexpr int NJL_THREADS = 256;                     // few warps: the iteration is a latency chain, and every warp pays its fixed parts
constexpr int NJL_WARPS = NJL_THREADS / 32;
constexpr int NJL_MAX_RPT = 3;                       // rows of a CTA per thread (template parameter RPT): ceil((NJL_MAX_N + 1) / 16 / NJL_THREADS)
constexpr int NJL_MAX_CLUSTER = 16;
constexpr int NJL_EPOCH = 8;                         // iterations between snapshots
constexpr int NJL_MAX_G = 64;
constexpr unsigned short NJL_EMPTY = 0xffffu;
constexpr int64_t NJL_MAX_N = 11000;                 // slots and rows are 16-bit, row indices of a CTA 15-bit
constexpr int NJL_SMEM_LIMIT = 227 * 1024;

struct NjLazyParams {
  double* W;              // private copies: (n + 1) rows of ldw doubles per matrix
  int64_t ldw, w_stride;  // doubles
  const double* r0;       // initial row sums, r_stride doubles per matrix
  int64_t r_stride;
  int32_t* merge;         // outputs: 3 (n - 2) per matrix
  double* len;
  int* status;            // per matrix: in 0 = go, out 0 = joined
  unsigned long long* stats;
  int n, G, rc, ncta, ncta_shift;
};

// dynamic shared memory carve-up (host and device agree through this struct)
struct NjLazySmem {
  int r, snap, s2p, R1, R2, ld, lc, slot, champ, q_resc, total;
  __host__ __device__ NjLazySmem(int n, int G, int rc) {
    const int n_pad = (n + 3) & ~3, rc_pad = (rc + 3) & ~3;
    int o = 0;
    r = o; o += 8 * (n_pad + 2);   // + a scratch element for masked writes
    R1 = o; o += 8 * rc_pad;
    R2 = o; o += 8 * rc_pad;
    ld = o; o += 8 * G * rc;
    snap = o; o += 4 * n_pad;
    s2p = o; o += 2 * n_pad;
    lc = o; o += (2 * G * rc + 7) & ~7;
    slot = o; o += 2 * rc_pad;
    q_resc = o; o += 2 * rc_pad;
    champ = o; o += rc_pad;
    total = (o + 15) & ~15;
  }
};

struct Cand {
  double q;
  int lo, hi;
};

// optional phase accounting (pf_nj_lazy_profile): nanoseconds of thread 0 of cluster rank 0 per phase
__device__ unsigned long long g_njl_prof[16];
bool g_njl_prof_enabled = false;
__device__ __forceinline__ unsigned long long njl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define NJL_MARK(slot)                                  \
  if (PROF && rank == 0 && tid == 0) {                  \
    const unsigned long long now_ = njl_now();          \
    g_njl_prof[slot] += now_ - t_prev;                  \
    t_prev = now_;                                      \
  }

__device__ __forceinline__ void cand_take(Cand& c, double q, int lo, int hi) {
  const bool t = rec_less(q, lo, hi, c.q, c.lo, c.hi);
  c.q = t ? q : c.q;
  c.lo = t ? lo : c.lo;
  c.hi = t ? hi : c.hi;
}

// the oracle's q of the pair of slots (i, j) with distance d
__device__ __forceinline__ double fpq(double n2, const double* s_r, int i, int j, double d, int& lo, int& hi) {
  lo = min(i, j);
  hi = max(i, j);
  double q = __dmul_rn(n2, d);
  q = __dsub_rn(q, s_r[lo]);
  q = __dsub_rn(q, s_r[hi]);
  return q;
}

// warp-wide min / max of doubles through their order-preserving 64-bit keys: two integer REDUX steps (high word,
// then the low word among the lanes that hold the extreme high word). No NaN among the inputs (callers filter them);
// -0.0 and +0.0 are distinct keys, which is harmless for bounds.
__device__ __forceinline__ unsigned long long f64_key(double x) {
  const long long bits = __double_as_longlong(x);
  return static_cast<unsigned long long>(bits) ^ ((bits < 0) ? 0xffffffffffffffffull : 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_unkey(unsigned long long key) {
  const unsigned long long bits = (key & 0x8000000000000000ull) ? (key ^ 0x8000000000000000ull) : ~key;
  return __longlong_as_double(static_cast<long long>(bits));
}
__device__ __forceinline__ double warp_min_f64(double x) {
  const unsigned long long key = f64_key(x);
  const unsigned hi = static_cast<unsigned>(key >> 32), lo = static_cast<unsigned>(key);
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  return f64_unkey((static_cast<unsigned long long>(mhi) << 32) | mlo);
}
__device__ __forceinline__ double warp_max_f64(double x) {
  const unsigned long long key = f64_key(x);
  const unsigned hi = static_cast<unsigned>(key >> 32), lo = static_cast<unsigned>(key);
  const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
  return f64_unkey((static_cast<unsigned long long>(mhi) << 32) | mlo);
}
// min / max of the NJL_WARPS per-warp values in shared memory (every thread, balanced tree)
__device__ __forceinline__ double smem_min_warps(const double* s) {
  static_assert(NJL_WARPS == 8, "tree of eight");
  const double a = fmin(s[0], s[1]), b = fmin(s[2], s[3]), c = fmin(s[4], s[5]), d = fmin(s[6], s[7]);
  return fmin(fmin(a, b), fmin(c, d));
}
__device__ __forceinline__ double smem_max_warps(const double* s) {
  static_assert(NJL_WARPS == 8, "tree of eight");
  const double a = fmax(s[0], s[1]), b = fmax(s[2], s[3]), c = fmax(s[4], s[5]), d = fmax(s[6], s[7]);
  return fmax(fmax(a, b), fmax(c, d));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t local_addr, uint32_t rank, unsigned long long w0, unsigned long long w1) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  asm volatile("st.shared::cluster.v2.u64 [%0], {%1, %2};" ::"r"(remote), "l"(w0), "l"(w1) : "memory");
}

// bound check: true when everything behind the bound R (snapshot frame) of a row with s_i = si is PROVEN worse
// than the threshold T. The margin is ~1e6 times the rounding of the terms (not of their possibly cancelled sum);
// NaN / inf on either side fail the proof.
__device__ __forceinline__ bool bound_clears(double R, double M, double si, double n2, double T) {
  const double lb = n2 * ((R - M) - si);
  const double margin = 1e-9 * (n2 * (fabs(R) + fabs(M) + fabs(si)) + fabs(T)) + 1e-290;
  return (lb - margin) > T;
}

struct RowCtx {   // per-CTA shared-memory views
  double* s_r;
  float* s_snap;
  unsigned short* s_s2p;
  double* s_R1;
  double* s_R2;
  double* s_ld;
  unsigned short* s_lc;
  short* s_slot;
  signed char* s_champ;
  int rc, G;
};

// exact evaluation of the list of own row lr (slot k) by one warp: every entry is a candidate, the best becomes the
// champion, R1 bounds the others. Lane l holds the entries l, l + 32, ...; returns the row's best on all lanes.
template <int G>
__device__ __forceinline__ void list_evaluate(const RowCtx& C, int lr, int k, double n2, int lane, Cand& out) {
  constexpr int EPL = (G + 31) / 32;
  double ve[EPL];
  int elo[EPL], ehi[EPL];
  Cand me{CUDART_INF, NJ_NONE, NJ_NONE};
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int g = lane + 32 * e;
    const int col = (g < G) ? C.s_lc[g * C.rc + lr] : NJL_EMPTY;
    elo[e] = NJ_NONE;
    ehi[e] = NJ_NONE;
    ve[e] = CUDART_INF;
    if (col != NJL_EMPTY) {
      const double d = C.s_ld[g * C.rc + lr];
      int lo, hi;
      const double q = fpq(n2, C.s_r, k, col, d, lo, hi);
      const double v = d - static_cast<double>(C.s_snap[col]);
      ve[e] = (v == v) ? v : CUDART_INF;   // NaN: a distance that can never win (the oracle's q is NaN too)
      if (q < CUDART_INF) {                // NaN / +inf never win
        elo[e] = lo;
        ehi[e] = hi;
        cand_take(me, q, lo, hi);
      }
    }
  }
  warp_argmin(me.q, me.lo, me.hi);
  int gch = -1;
  double r1 = CUDART_INF;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const bool is_champ = me.lo != NJ_NONE && elo[e] == me.lo && ehi[e] == me.hi;   // the pair identifies the entry
    if (is_champ) gch = lane + 32 * e;
    else r1 = fmin(r1, ve[e]);
  }
  const unsigned champ_mask = __ballot_sync(0xffffffffu, gch >= 0);
  r1 = warp_min_f64(r1);
  const int gsel = champ_mask ? __shfl_sync(0xffffffffu, gch, __ffs(champ_mask) - 1) : -1;
  if (lane == 0) {
    C.s_champ[lr] = static_cast<signed char>(gsel);
    C.s_R1[lr] = r1;
  }
  out = me;
  __syncwarp();   // the row's owner is a lane of this warp
}

// two partial results of the same group (v = ve of the tracked candidate, b2 = bound of the others): the smaller
// (v, column) stays, the other one is covered by the bound. Symmetric in its arguments, so both lanes of a shuffle
// pair arrive at the same result.
__device__ __forceinline__ void group_merge(double& v, int& col, double& d, double& b2, double v2, int col2, double d2,
                                            double b22) {
  const bool other = (v2 < v) || (v2 == v && col2 < col);
  const double lv = other ? v : v2;   // ve of the loser (+inf when it is empty)
  b2 = fmin(fmin(b2, b22), lv);
  v = other ? v2 : v;
  col = other ? col2 : col;
  d = other ? d2 : d;
}

// rescan of own row lr by the CTA: per group the column with the smallest ve = d - snap becomes the list entry,
// everything else goes into R2; the list is evaluated by the owner's warp in the next iteration (no champion, R1 =
// -inf). The row's candidate of this iteration does not depend on the list: every column whose q is not PROVEN worse
// than the CTA's threshold T (the same bound and margin as bound_clears, per column) is evaluated with the oracle's
// operations, so ties and near-ties beyond the list's capacity are decided as the oracle decides them.
// The NJL_THREADS / G threads of a group are neighbouring lanes of one warp (a warp reads 32 / TPG consecutive
// columns per 32-column period: full sectors), so a group is finished with shuffles; only R2 and the candidate cross
// the warps (buffers `buf`, alternating between consecutive calls: one CTA barrier per call).
// Returns the row's best candidate on all lanes of warp 0 (other warps: +inf).
template <int G, bool PROF>
__device__ __forceinline__ void rescan_row(const RowCtx& C, const double* __restrict__ rowp, int lr, int na, double n2,
                                           double inv, double M, double T, int tid, NjRec* s_rrec, double* s_rb2,
                                           Cand& out, int rank) {
  constexpr int TPG = NJL_THREADS / G;   // threads per group: 4 ... 32
  constexpr int GPW = 32 / TPG;          // groups per warp
  unsigned long long t_prev = PROF ? njl_now() : 0;
  if (PROF && rank == 0 && tid == 0) {   // one dependent load from the row: the round trip as this kernel sees it
    const double x = __ldcg(rowp + 1);
    if (x == 123.456) g_njl_prof[15] += 1;
    NJL_MARK(13)
  }
  const int k = C.s_slot[lr];
  const int lane = tid & 31, warp = tid >> 5;
  const int sub = lane % TPG, g = warp * GPW + lane / TPG;
  double bv = CUDART_INF, bd = 0.0, b2 = CUDART_INF;
  int bcol = NJL_EMPTY;
  const double rk = C.s_r[k];
  // column j is skipped when n2 ((ve_j - M) - s_k) - margin > T, i.e. ve_j - 1e-9 |ve_j| > theta
  const double sk = rk * inv, Tn = T * inv;
  const double theta = ((Tn + M) + sk) + 1e-9 * ((fabs(M) + fabs(sk)) + fabs(Tn)) + 1e-290;
  Cand ex{CUDART_INF, NJ_NONE, NJ_NONE};
  constexpr int U = 8;
  for (int j0 = g + G * sub; j0 < na; j0 += NJL_THREADS * U) {
    double dv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + u * NJL_THREADS;
      dv[u] = (j < na) ? __ldcg(rowp + j) : 0.0;
    }
    double ve[U];
    bool any_pass = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + u * NJL_THREADS;
      const bool valid = j < na && j != k;
      const double x = dv[u] - static_cast<double>(C.s_snap[min(j, na - 1)]);
      ve[u] = valid ? x : CUDART_INF;
      any_pass |= valid && !(__fma_rn(-1e-9, fabs(x), x) > theta);   // not proven worse than T (NaN: not proven)
    }
    if (any_pass) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * NJL_THREADS;
        if (j < na && j != k && !(__fma_rn(-1e-9, fabs(ve[u]), ve[u]) > theta)) {
          const double rj = C.s_r[j];
          const bool jlo = j < k;
          double q = __dmul_rn(n2, dv[u]);
          q = __dsub_rn(q, jlo ? rj : rk);
          q = __dsub_rn(q, jlo ? rk : rj);
          if (q < CUDART_INF) cand_take(ex, q, jlo ? j : k, jlo ? k : j);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // the smaller ve stays, the other one (the old best, or this column) is covered by the bound
      const bool lt = ve[u] < bv;
      b2 = fmin(b2, lt ? bv : ve[u]);
      bv = lt ? ve[u] : bv;
      bcol = lt ? j0 + u * NJL_THREADS : bcol;
      bd = lt ? dv[u] : bd;
    }
  }
  NJL_MARK(9)   // rescan: loads + scan
#pragma unroll
  for (int off = TPG / 2; off >= 1; off >>= 1) {
    const double v2 = __shfl_xor_sync(0xffffffffu, bv, off);
    const int c2 = __shfl_xor_sync(0xffffffffu, bcol, off);
    const double d2 = __shfl_xor_sync(0xffffffffu, bd, off);
    const double b22 = __shfl_xor_sync(0xffffffffu, b2, off);
    group_merge(bv, bcol, bd, b2, v2, c2, d2, b22);
  }
  if (sub == 0) {
    C.s_lc[g * C.rc + lr] = static_cast<unsigned short>(bcol);
    C.s_ld[g * C.rc + lr] = bd;
  }
  b2 = warp_min_f64(b2);
  warp_argmin(ex.q, ex.lo, ex.hi);
  if (lane == 0) {
    s_rrec[warp] = NjRec{ex.q, ex.lo, ex.hi};
    s_rb2[warp] = b2;
  }
  __syncthreads();
  NJL_MARK(10)   // rescan: group and warp reductions, barrier
  out.q = CUDART_INF;
  out.lo = NJ_NONE;
  out.hi = NJ_NONE;
  if (warp == 0) {
    const double r2 = smem_min_warps(s_rb2);
    if (lane == 0) {
      C.s_R2[lr] = r2;
      C.s_champ[lr] = -1;
      C.s_R1[lr] = -CUDART_INF;   // the owner's warp evaluates the list in the next iteration
    }
    if (lane < NJL_WARPS) cand_take(out, s_rrec[lane].q, s_rrec[lane].a, s_rrec[lane].b);
    warp_argmin(out.q, out.lo, out.hi);
  }
  NJL_MARK(11)   // rescan: tail of warp 0
  if (PROF && rank == 0 && tid == 0) g_njl_prof[12] += 1;
}

// maintenance of the list of own row lr (slot k, distance du to the merged node) after the join of (wa, wb):
// the joined columns leave, column `last` becomes column wb, the merged node enters as column wa.
__device__ __forceinline__ void row_update(const RowCtx& C, int lr, int k, double du, int wa, int wb, int last, bool move,
                                           double sue) {
  const int G = C.G, rc = C.rc;
  int champ = C.s_champ[lr];
  double R1 = C.s_R1[lr], R2 = C.s_R2[lr];
  auto ve_of = [&](int col, double d) { return d - static_cast<double>(C.s_snap[col]); };
  auto fold = [&](double& R, double ve) { if (!(ve >= R)) R = (ve == ve) ? ve : R; };   // NaN bounds nothing away: ignored like the oracle ignores NaN q
  // (1) the joined columns disappear
  {
    const int ga = wa % G, gb = wb % G;
    if (C.s_lc[ga * rc + lr] == wa) {
      C.s_lc[ga * rc + lr] = NJL_EMPTY;
      if (champ == ga) { champ = -1; R1 = -CUDART_INF; }
    }
    if (C.s_lc[gb * rc + lr] == wb) {
      C.s_lc[gb * rc + lr] = NJL_EMPTY;
      if (champ == gb) { champ = -1; R1 = -CUDART_INF; }
    }
  }
  // (2) the column of the last slot becomes column wb
  if (move) {
    const int gl = last % G, gb = wb % G;
    if (C.s_lc[gl * rc + lr] == last) {
      if (gl == gb) {
        C.s_lc[gl * rc + lr] = static_cast<unsigned short>(wb);
      } else {
        const double dl = C.s_ld[gl * rc + lr];
        const bool was_champ = (champ == gl);
        C.s_lc[gl * rc + lr] = NJL_EMPTY;
        if (was_champ) champ = -1;
        const int oc = C.s_lc[gb * rc + lr];
        if (oc == NJL_EMPTY) {
          C.s_lc[gb * rc + lr] = static_cast<unsigned short>(wb);
          C.s_ld[gb * rc + lr] = dl;
          if (was_champ) champ = gb;
        } else {
          // two tracked columns in one group: the better one stays, R2 covers the other. (s_r / s_snap of slot
          // wb already hold the moved node's values.)
          const double od = C.s_ld[gb * rc + lr];
          if (ve_of(wb, dl) < ve_of(oc, od)) {
            fold(R2, ve_of(oc, od));
            if (champ == gb) { champ = -1; R1 = -CUDART_INF; }
            C.s_lc[gb * rc + lr] = static_cast<unsigned short>(wb);
            C.s_ld[gb * rc + lr] = dl;
            if (was_champ) champ = gb;
          } else {
            fold(R2, ve_of(wb, dl));
            if (was_champ) R1 = -CUDART_INF;
          }
        }
      }
    }
  }
  // (3) the merged node is a new column (slot wa)
  {
    const int g = wa % G;
    const double ve_new = du - sue;
    const int oc = C.s_lc[g * rc + lr];
    bool placed = false;
    if (oc == NJL_EMPTY) {
      placed = true;
    } else {
      const double od = C.s_ld[g * rc + lr];
      if (ve_new < ve_of(oc, od)) {
        fold(R2, ve_of(oc, od));
        if (champ == g) champ = -1;
        placed = true;
      } else {
        fold(R2, ve_new);
      }
    }
    if (placed) {
      C.s_lc[g * rc + lr] = static_cast<unsigned short>(wa);
      C.s_ld[g * rc + lr] = du;
      if (champ < 0) {
        R1 = -CUDART_INF;                    // evaluate the list next time
      } else {
        const int cc = C.s_lc[champ * rc + lr];
        const double cd = C.s_ld[champ * rc + lr];
        if (ve_new < ve_of(cc, cd)) {
          fold(R1, ve_of(cc, cd));
          champ = g;
        } else {
          fold(R1, ve_new);
        }
      }
    }
  }
  C.s_champ[lr] = static_cast<signed char>(champ);
  C.s_R1[lr] = R1;
  C.s_R2[lr] = R2;
}

template <int G, int RPT, bool PROF>
__global__ void __launch_bounds__(NJL_THREADS, 1)
nj_lazy_kernel(const NjLazyParams P) {
  extern __shared__ __align__(16) unsigned char nl_dyn[];
  __shared__ double s_wq[NJL_WARPS];          // per-warp best champion q (threshold of the CTA)
  __shared__ NjRec s_wrec[NJL_WARPS];          // per-warp best candidate after the list evaluations
  __shared__ NjRec s_rrec[2][NJL_WARPS];       // rescans: per-warp best exact candidate (alternating buffers)
  __shared__ double s_rb2[2][NJL_WARPS];       // rescans: per-warp bound of the untracked columns
  __shared__ double s_red[3][NJL_WARPS];
  __shared__ __align__(16) unsigned long long s_x[2][NJL_MAX_CLUSTER][4];   // exchange records {q, lo | hi << 16 | flags << 32, r_lo, r_hi}
  __shared__ int s_cnt;                                                     // rows queued for a rescan

  const int n = P.n, rc = P.rc, ncta = P.ncta, shift = P.ncta_shift;
  const int rank = static_cast<int>(cluster_rank());
  const int mat = static_cast<int>(blockIdx.x) >> shift;
  if (P.status[mat] != 0) return;   // the whole cluster leaves: nothing was touched

  const NjLazySmem L(n, G, rc);
  RowCtx C;
  C.s_r = reinterpret_cast<double*>(nl_dyn + L.r);
  C.s_snap = reinterpret_cast<float*>(nl_dyn + L.snap);
  C.s_s2p = reinterpret_cast<unsigned short*>(nl_dyn + L.s2p);
  C.s_R1 = reinterpret_cast<double*>(nl_dyn + L.R1);
  C.s_R2 = reinterpret_cast<double*>(nl_dyn + L.R2);
  C.s_ld = reinterpret_cast<double*>(nl_dyn + L.ld);
  C.s_lc = reinterpret_cast<unsigned short*>(nl_dyn + L.lc);
  C.s_slot = reinterpret_cast<short*>(nl_dyn + L.slot);
  C.s_champ = reinterpret_cast<signed char*>(nl_dyn + L.champ);
  C.rc = rc;
  C.G = G;
  unsigned short* q_resc = reinterpret_cast<unsigned short*>(nl_dyn + L.q_resc);

  double* __restrict__ W = P.W + mat * P.w_stride;
  const int64_t ldw = P.ldw;
  const double* r0 = P.r0 + mat * P.r_stride;
  int32_t* merge = P.merge + 3 * static_cast<int64_t>(n - 2) * mat;
  double* len = P.len + 3 * static_cast<int64_t>(n - 2) * mat;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- initial state: row sums, snapshot, identity maps, empty lists; every live row is scanned once
  int na = n, t = 0;
  double M;
  {
    const double inv0 = 1.0 / static_cast<double>(na - 2);
    double mx = -CUDART_INF;
    for (int k = tid; k < n; k += NJL_THREADS) {
      const double rk = __ldcg(r0 + k);
      C.s_r[k] = rk;
      const double s = rk * inv0;
      const float sn = static_cast<float>(s);
      C.s_snap[k] = sn;
      C.s_s2p[k] = static_cast<unsigned short>(k);
      mx = fmax(mx, s - static_cast<double>(sn));
    }
    mx = warp_max_f64(mx);
    if (lane == 0) s_red[0][warp] = mx;
    for (int lr = tid; lr < rc; lr += NJL_THREADS) {
      const int p = (lr << shift) + rank;
      C.s_slot[lr] = static_cast<short>(p < n ? p : -1);
      C.s_champ[lr] = -1;
      C.s_R1[lr] = CUDART_INF;
      C.s_R2[lr] = CUDART_INF;
      for (int g = 0; g < G; ++g) C.s_lc[g * rc + lr] = NJL_EMPTY;
      q_resc[lr] = static_cast<unsigned short>(lr);
    }
    // rows of this CTA: physical rows rank, rank + ncta, ... < n (the spare row n is not live)
    if (tid == 0) s_cnt = (n - rank + ncta - 1) >> shift;
    __syncthreads();
    M = smem_max_warps(s_red[0]);
  }
  cluster_sync_all();   // every CTA of the cluster is running

  double* own_row[RPT];   // the matrix rows of this thread's list rows
#pragma unroll
  for (int i = 0; i < RPT; ++i) own_row[i] = W + static_cast<int64_t>(((tid + i * NJL_THREADS) << shift) + rank) * ldw;
  unsigned long long n_rescan = 0, n_reeval = 0;   // warp 0 counts for the CTA
  unsigned long long t_prev = PROF ? njl_now() : 0;
  int spare = n;
  bool first = true, bailed = false;
  double T0 = CUDART_INF;   // exact q of the previous exchange's runner-up in the current frame (a valid threshold)
  double inv = 1.0 / static_cast<double>(na - 2);

  while (na > 3) {
    const double n2 = static_cast<double>(na - 2);
    Cand best{CUDART_INF, NJ_NONE, NJ_NONE};
    double T = CUDART_INF;   // the CTA's threshold: its best champion (first iteration: none yet)
    if (!first) {
      // ---- A: champions of the own rows, exactly -> the CTA's threshold T. Lists without a champion that ask for it
      // (R1 = -inf: after a rescan, or when the champion was joined) are evaluated here, by the warp of the owner.
      unsigned fresh = 0;   // bit i: row tid + i * NJL_THREADS was evaluated in this phase
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int lr = tid + i * NJL_THREADS;
        const int k = (lr < rc) ? static_cast<int>(C.s_slot[lr]) : -1;
        const int ch = (k >= 0) ? static_cast<int>(C.s_champ[lr]) : -1;
        const bool need = k >= 0 && ch < 0 && C.s_R1[lr] == -CUDART_INF;
        unsigned mev = __ballot_sync(0xffffffffu, need);
        if (lane == 0) n_reeval += __popc(mev);
        while (mev) {
          const int src = __ffs(mev) - 1;
          mev &= mev - 1;
          const int lr_s = __shfl_sync(0xffffffffu, lr, src), k_s = __shfl_sync(0xffffffffu, k, src);
          Cand c;
          list_evaluate<G>(C, lr_s, k_s, n2, lane, c);
          cand_take(best, c.q, c.lo, c.hi);
        }
        if (need) fresh |= 1u << i;
        if (k < 0 || ch < 0) continue;
        int lo, hi;
        const double q = fpq(n2, C.s_r, k, C.s_lc[ch * rc + lr], C.s_ld[ch * rc + lr], lo, hi);
        if (q < CUDART_INF) cand_take(best, q, lo, hi);   // +inf and NaN never win (as in the full-scan kernels)
      }
      {
        const double wq = warp_min_f64(best.q);
        if (lane == 0) s_wq[warp] = wq;
      }
      __syncthreads();
      T = fmin(smem_min_warps(s_wq), T0);
      NJL_MARK(1)   // A
      // ---- B: bounds that do not clear T. Lists are evaluated on the spot by the warp of the row's owner;
      // rows whose rest-of-row bound fails are queued for a rescan by the whole CTA.
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int lr = tid + i * NJL_THREADS;
        const int k = (lr < rc) ? static_cast<int>(C.s_slot[lr]) : -1;
        bool ev = false;
        if (k >= 0) {
          const double si = C.s_r[k] * inv;
          const double R2 = C.s_R2[lr], R1 = C.s_R1[lr];
          if (!(R2 == CUDART_INF) && !bound_clears(R2, M, si, n2, T)) {
            q_resc[atomicAdd(&s_cnt, 1)] = static_cast<unsigned short>(lr);
          } else if (!((fresh >> i) & 1u) && !(R1 == CUDART_INF) && !bound_clears(R1, M, si, n2, T)) {
            ev = true;
          }
        }
        unsigned mev = __ballot_sync(0xffffffffu, ev);
        if (lane == 0) n_reeval += __popc(mev);
        while (mev) {
          const int src = __ffs(mev) - 1;
          mev &= mev - 1;
          const int lr_s = __shfl_sync(0xffffffffu, lr, src), k_s = __shfl_sync(0xffffffffu, k, src);
          Cand c;
          list_evaluate<G>(C, lr_s, k_s, n2, lane, c);
          cand_take(best, c.q, c.lo, c.hi);
        }
      }
      warp_argmin(best.q, best.lo, best.hi);
      if (lane == 0) s_wrec[warp] = NjRec{best.q, best.lo, best.hi};
      __syncthreads();
      NJL_MARK(2)   // B
    }
    // ---- C: rescans (the whole CTA per row); first iteration: every row
    const int n_resc = s_cnt;
    Cand bres{CUDART_INF, NJ_NONE, NJ_NONE};
    for (int qi = 0; qi < n_resc; ++qi) {
      const int lr = q_resc[qi];
      const int p = (lr << shift) + rank;
      Cand c;
      rescan_row<G, PROF>(C, W + static_cast<int64_t>(p) * ldw, lr, na, n2, inv, M, T, tid, s_rrec[qi & 1], s_rb2[qi & 1], c, rank);
      cand_take(bres, c.q, c.lo, c.hi);
    }
    // ---- D: exchange over the cluster. Warp 0 assembles the CTA's candidate and writes it to every CTA
    const int par = t & 1;
    if (warp == 0) {
      n_rescan += n_resc;
      Cand fin = bres;
      if (!first && lane < NJL_WARPS) cand_take(fin, s_wrec[lane].q, s_wrec[lane].a, s_wrec[lane].b);
      warp_argmin(fin.q, fin.lo, fin.hi);
      // budget of rescans (degenerate inputs: ties everywhere, NaN): beyond it the full-scan kernel is the better tool
      const bool over = n_rescan > 1024ull + 8ull * static_cast<unsigned long long>(t) + 2ull * static_cast<unsigned long long>(rc);
      const bool none = fin.lo == NJ_NONE;
      const unsigned ab = none ? 0xffffffffu : (static_cast<unsigned>(fin.lo) | (static_cast<unsigned>(fin.hi) << 16));
      const unsigned long long w0 = static_cast<unsigned long long>(__double_as_longlong(fin.q));
      const unsigned long long w1 = (static_cast<unsigned long long>(over ? 1u : 0u) << 32) | ab;
      const unsigned long long w2 = static_cast<unsigned long long>(__double_as_longlong(none ? 0.0 : C.s_r[fin.lo]));
      const unsigned long long w3 = static_cast<unsigned long long>(__double_as_longlong(none ? 0.0 : C.s_r[fin.hi]));
      if (lane < ncta) {
        const uint32_t a0 = smem_u32(&s_x[par][rank][0]);
        st_cluster_v2(a0, static_cast<uint32_t>(lane), w0, w1);
        st_cluster_v2(a0 + 16, static_cast<uint32_t>(lane), w2, w3);
      }
    }
    first = false;
    NJL_MARK(3)   // C + publish
    const double inv_new = (na > 3) ? 1.0 / static_cast<double>(na - 3) : 0.0;   // (before the barrier: off the critical path)
    cluster_sync_all();
    NJL_MARK(4)   // cluster barrier (includes waiting for the slowest CTA)
    if (PROF) {
      cluster_sync_all();
      NJL_MARK(8)   // the barrier alone (nobody is late, nothing to publish)
    }
    int wa, wb;
    double rra, rrb;
    int ru_a = NJ_NONE, ru_b = NJ_NONE;   // the runner-up of this exchange: best published pair that survives the join
    {
      double q = CUDART_INF;
      int a = NJ_NONE, b = NJ_NONE;
      unsigned flags = 0;
      double xa = 0.0, xb = 0.0;
      if (lane < ncta) {
        const ulonglong2 u0 = *reinterpret_cast<const ulonglong2*>(&s_x[par][lane][0]);
        const ulonglong2 u1 = *reinterpret_cast<const ulonglong2*>(&s_x[par][lane][2]);
        const unsigned ab = static_cast<unsigned>(u0.y);
        flags = static_cast<unsigned>(u0.y >> 32);
        xa = __longlong_as_double(static_cast<long long>(u1.x));
        xb = __longlong_as_double(static_cast<long long>(u1.y));
        if (ab != 0xffffffffu) {
          q = __longlong_as_double(static_cast<long long>(u0.x));
          a = static_cast<int>(ab & 0xffffu);
          b = static_cast<int>(ab >> 16);
        }
      }
      const int a0 = a, b0 = b;
      const double q0 = q;
      warp_argmin(q, a, b);
      if (__any_sync(0xffffffffu, flags != 0)) { bailed = true; break; }
      if (a == NJ_NONE) {
        // no finite candidate anywhere (NaN / inf input): join slots 0 and 1, as the full-scan kernels do
        wa = 0;
        wb = 1;
        rra = C.s_r[0];
        rrb = C.s_r[1];
        __syncthreads();   // (this path only: r_0 / r_1 are read from shared memory that the update overwrites)
      } else {
        const unsigned src = __ballot_sync(0xffffffffu, a0 == a && b0 == b);   // the pair identifies the record
        const int sl = __ffs(src) - 1;
        wa = a;
        wb = b;
        rra = __shfl_sync(0xffffffffu, xa, sl);
        rrb = __shfl_sync(0xffffffffu, xb, sl);
        // the best record whose nodes are not joined now: after the update its exact q is a threshold every CTA can
        // use before it has seen anybody else's candidates (the CTA-local best is weaker: more rescans)
        const bool alive = a0 != NJ_NONE && a0 != wa && a0 != wb && b0 != wa && b0 != wb;
        double q2 = alive ? q0 : CUDART_INF;
        ru_a = alive ? a0 : NJ_NONE;
        ru_b = alive ? b0 : NJ_NONE;
        warp_argmin(q2, ru_a, ru_b);
      }
    }

    // ---- E: update. Row sums / snapshot / maps: every CTA, redundantly; matrix: the owner of a row
    const int last = na - 1;
    const bool move = (wb != last);
    const int pwa = C.s_s2p[wa], pwb = C.s_s2p[wb], plast = C.s_s2p[last];
    const double* row_a = W + static_cast<int64_t>(pwa) * ldw;
    const double* row_b = W + static_cast<int64_t>(pwb) * ldw;
    double* row_u = W + static_cast<int64_t>(spare) * ldw;
    const bool own_u = (spare & (ncta - 1)) == rank;
    const double dab = __ldcg(row_a + wb);
    const double ru_d = (ru_a != NJ_NONE) ? __ldcg(W + static_cast<int64_t>(C.s_s2p[ru_a]) * ldw + ru_b) : 0.0;
    const bool resnap = ((t + 1) % NJL_EPOCH) == 0;
    // the rows of this CTA: their distances to a and b (and the entry that moves into column b) are requested now and
    // used after the replicated part
    int ok_[RPT];
    double oa_[RPT], ob_[RPT], om_[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int lr = tid + i * NJL_THREADS;
      int k = (lr < rc) ? static_cast<int>(C.s_slot[lr]) : -1;
      if (k == wa || k == wb) k = -1;
      ok_[i] = k;
      oa_[i] = (k >= 0) ? __ldcg(row_a + k) : 0.0;
      ob_[i] = (k >= 0) ? __ldcg(row_b + k) : 0.0;
      om_[i] = (k >= 0 && move && k != last) ? __ldcg(own_row[i] + last) : 0.0;
    }
    NJL_MARK(5)   // winner
    // replicated: r_k <- ((r_k - d_ak) - d_bk) + d_uk for every untouched node, the maximal drift against the snapshot
    double mx = -CUDART_INF;
    const int r_dummy = (n + 3) & ~3;   // the scratch element after the row sums
    constexpr int UK = 8;
    for (int k0 = tid; k0 < na; k0 += NJL_THREADS * UK) {
      double dak[UK], dbk[UK];
#pragma unroll
      for (int u = 0; u < UK; ++u) {
        const int k = k0 + u * NJL_THREADS;
        dak[u] = (k < na) ? __ldcg(row_a + k) : 0.0;
        dbk[u] = (k < na) ? __ldcg(row_b + k) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < UK; ++u) {
        const int k = k0 + u * NJL_THREADS;
        const bool live = k < na && k != wa && k != wb;
        const int kc = min(k, na - 1);
        double d = __dadd_rn(dak[u], dbk[u]);
        d = __dsub_rn(d, dab);
        d = __dmul_rn(0.5, d);
        double rk = __dsub_rn(C.s_r[kc], dak[u]);
        rk = __dsub_rn(rk, dbk[u]);
        rk = __dadd_rn(rk, d);
        const int slot = (move && k == last) ? wb : kc;   // the last live slot moves into slot b
        const double drift = rk * inv_new - static_cast<double>(C.s_snap[kc]);
        mx = live ? fmax(mx, drift) : mx;
        C.s_r[live ? slot : r_dummy] = rk;                // (a scratch element takes the writes of the joined slots)
        if (own_u && live) row_u[slot] = d;               // row of the merged node, in the spare physical row
      }
    }
    if (tid == 0 && move) C.s_snap[wb] = C.s_snap[last];  // (nobody else touches these two elements in this phase)
    // r of the merged node: every thread computes it (same bits), one writes it
    double ru = __dadd_rn(rra, rrb);
    ru = __dsub_rn(ru, __dmul_rn(static_cast<double>(na), dab));
    ru = __dmul_rn(0.5, ru);
    mx = warp_max_f64(mx);
    if (lane == 0) s_red[0][warp] = mx;
    if (tid == 0) {
      C.s_r[wa] = ru;
      if (own_u) row_u[wa] = 0.0;
      if (rank == 0) {   // log of the join; node ids and branch lengths are derived after the last iteration
        merge[3 * t + 0] = wa;
        merge[3 * t + 1] = wb;
        len[3 * t + 0] = rra;
        len[3 * t + 1] = rrb;
        len[3 * t + 2] = dab;
      }
    }
    // the rows of this CTA in the matrix: column a <- distance to the merged node, column b <- column last
    double odu_[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      double d = __dadd_rn(oa_[i], ob_[i]);
      d = __dsub_rn(d, dab);
      d = __dmul_rn(0.5, d);
      odu_[i] = d;
      if (ok_[i] >= 0) {
        double* rk_p = own_row[i];
        rk_p[wa] = d;
        if (move && ok_[i] != last) rk_p[wb] = om_[i];
      }
    }
    __syncthreads();
    NJL_MARK(6)   // E: update of the row sums and the matrix
    const double M_old_frame = smem_max_warps(s_red[0]);   // max_j s_j(t + 1) - snap_j, untouched nodes
    double radj = 0.0;   // added to every stored bound (new snapshot only)
    double Mn = M_old_frame;
    if (resnap) {
      // new snapshot of every untouched node (the merged node's follows below); the stored bounds move from
      // "d - snap_old >= R" to the new frame: d - snap' = (d - s) + (s - snap') >= (R - M_old) + min_j (s_j - snap'_j)
      double mx2 = -CUDART_INF, mn2 = CUDART_INF;
      for (int j = tid; j < na - 1; j += NJL_THREADS) {
        if (j == wa) continue;
        const double sj = C.s_r[j] * inv_new;
        const float sn = static_cast<float>(sj);
        const double df = sj - static_cast<double>(sn);
        C.s_snap[j] = sn;
        mx2 = fmax(mx2, df);
        mn2 = fmin(mn2, df);
      }
      mx2 = warp_max_f64(mx2);
      mn2 = warp_min_f64(mn2);
      if (lane == 0) {
        s_red[1][warp] = mx2;
        s_red[2][warp] = mn2;
      }
      __syncthreads();
      const double Mx2 = smem_max_warps(s_red[1]);
      const double Mn2 = smem_min_warps(s_red[2]);
      radj = Mn2 - M_old_frame;
      Mn = Mx2;
    }
    // the runner-up in the frame of the next iteration (its nodes were not touched, only relabelled and their row sums
    // updated): the exact q of a live pair, hence a valid threshold
    T0 = CUDART_INF;
    if (ru_a != NJ_NONE && na > 4) {
      if (move && ru_a == last) ru_a = wb;
      if (move && ru_b == last) ru_b = wb;
      int lo, hi;
      const double qn = fpq(static_cast<double>(na - 3), C.s_r, ru_a, ru_b, ru_d, lo, hi);
      if (qn < CUDART_INF) T0 = qn;
    }
    // the merged node's column starts at the current maximal drift, so that it does not raise M
    const double su = ru * inv_new;
    const float snu = static_cast<float>(su - Mn);
    const double sue = static_cast<double>(snu);
    M = fmax(Mn, su - sue);
    if (lane == 0) {
      // Lane 0 of EVERY warp writes the same values: the list evaluations of phase A (by the warp of a row's owner,
      // before the next CTA barrier) read s_snap of the merged node's column, and a warp is ordered only against its
      // own writes there (__syncwarp below). The maps are read after the next cluster barrier.
      C.s_snap[wa] = snu;
      C.s_s2p[wa] = static_cast<unsigned short>(spare);
      if (move) C.s_s2p[wb] = static_cast<unsigned short>(plast);
      if (tid == 0) s_cnt = 0;   // (queue of phase B, behind the barrier that ends phase A)
    }

    // ---- F: the rows of this CTA (a thread per row): slot relabelling, list maintenance
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int lr = tid + i * NJL_THREADS;
      if (lr >= rc) continue;
      const int p = (lr << shift) + rank;
      if (p == spare) {
        // this physical row now holds the merged node: nothing tracked yet (the other rows cover its pairs)
        C.s_slot[lr] = static_cast<short>(wa);
        C.s_champ[lr] = -1;
        C.s_R1[lr] = CUDART_INF;
        C.s_R2[lr] = CUDART_INF;
        for (int g = 0; g < G; ++g) C.s_lc[g * rc + lr] = NJL_EMPTY;
        continue;
      }
      int k = ok_[i];
      if (k < 0) {   // dead already, or joined now: the row of a becomes the spare row, the row of b is dead
        if (C.s_slot[lr] >= 0) C.s_slot[lr] = -1;
        continue;
      }
      if (move && k == last) {
        k = wb;
        C.s_slot[lr] = static_cast<short>(wb);
      }
      if (resnap) {
        const double R1 = C.s_R1[lr], R2 = C.s_R2[lr];
        if (!(R1 == CUDART_INF) && !(R1 == -CUDART_INF)) C.s_R1[lr] = R1 + radj;
        if (!(R2 == CUDART_INF)) C.s_R2[lr] = R2 + radj;
      }
      row_update(C, lr, k, odu_[i], wa, wb, last, move, sue);
    }
    spare = pwa;   // free from the next iteration on: everybody has read it by the next cluster barrier
    --na;
    ++t;
    inv = inv_new;
    __syncwarp();
    NJL_MARK(7)   // F: rows
  }

  if (bailed) {
    if (rank == 0 && tid == 0) P.status[mat] = 1;
    return;   // (all CTAs passed the same cluster barrier and read the same flags: nobody touches a peer again)
  }
  cluster_sync_all();   // the matrix writes of the last update are visible; no remote shared-memory access after this
  if (lane == 0 && P.stats) {
    if (warp == 0) atomicAdd(P.stats + 0, n_rescan);
    atomicAdd(P.stats + 1, n_reeval);
    if (rank == 0 && warp == 0) {
      atomicAdd(P.stats + 2, static_cast<unsigned long long>(t));
      atomicAdd(P.stats + 3, 1ull);
    }
  }
  if (rank != 0) return;
  // ---- CTA 0 of the cluster: node ids and branch lengths from the log
  double d01 = 0.0, d02 = 0.0, d12 = 0.0;
  if (tid == 0) {
    auto vread = [&](int x, int y) { return __ldcg(W + static_cast<int64_t>(C.s_s2p[x]) * ldw + y); };
    d01 = vread(0, 1);
    d02 = vread(0, 2);
    d12 = vread(1, 2);
  }
  __syncthreads();   // the log written by thread 0 is visible to the CTA; the shared-memory state is free
  // s_log over s_r (8 (n - 3) <= 8 n bytes), s_node over the list distances (4 n <= 8 G rc bytes: G rc >= n / 2)
  njc::nj_finish_from_log(merge, len, n, t, d01, d02, d12, reinterpret_cast<int2*>(nl_dyn + L.r),
                          reinterpret_cast<int*>(nl_dyn + L.ld), tid, NJL_THREADS);
}

// ---- preparation (full grid): private copy of the matrix + initial row sums; symmetry check
__global__ void __launch_bounds__(256)
nj_lazy_prep_kernel(const double* __restrict__ D, int64_t ld, int64_t mat_stride, int n, int n_mat, double* __restrict__ W,
                    int64_t ldw, int64_t w_stride, double* __restrict__ r0, int64_t r_stride) {
  const int lane = threadIdx.x & 31;
  const int64_t gwarp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t item = gwarp; item < static_cast<int64_t>(n_mat) * n; item += nwarps) {
    const int mat = static_cast<int>(item / n), row = static_cast<int>(item - static_cast<int64_t>(mat) * n);
    const double* rp = D + mat * mat_stride + static_cast<int64_t>(row) * ld;
    double* wp = W + mat * w_stride + static_cast<int64_t>(row) * ldw;
    double acc = 0.0;
#pragma unroll 8
    for (int k = lane; k < n; k += 32) {
      const double v = __ldcg(rp + k);
      acc = __dadd_rn(acc, v);
      wp[k] = v;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
    if (lane == 0) r0[mat * r_stride + row] = acc;
  }
}

// status[mat] = 1 when D[i][j] and D[j][i] differ in any bit pattern that matters (value comparison; NaN counts as
// different): the bounded scan reads each distance from the row of either node
__global__ void __launch_bounds__(256)
nj_lazy_symmetry_kernel(const double* __restrict__ D, int64_t ld, int64_t mat_stride, int n, int n_mat, int* status) {
  __shared__ double tile[32][33];
  const int nt = (n + 31) / 32;
  const int64_t pairs = static_cast<int64_t>(nt) * (nt + 1) / 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int64_t item = blockIdx.x; item < pairs * n_mat; item += gridDim.x) {
    const int mat = static_cast<int>(item / pairs);
    int64_t pi = item - mat * pairs;
    int I = 0;
    while (pi >= nt - I) { pi -= nt - I; ++I; }   // tile row I, tile column J = I + pi
    const int J = I + static_cast<int>(pi);
    const double* Dm = D + mat * mat_stride;
    __syncthreads();
    for (int y = ty; y < 32; y += 8) {
      const int i = J * 32 + y, j = I * 32 + tx;     // tile (J, I), read row-wise
      tile[y][tx] = (i < n && j < n) ? __ldcg(Dm + static_cast<int64_t>(i) * ld + j) : 0.0;
    }
    __syncthreads();
    bool bad = false;
    for (int y = ty; y < 32; y += 8) {
      const int i = I * 32 + y, j = J * 32 + tx;     // tile (I, J)
      if (i < n && j < n) {
        const double v = __ldcg(Dm + static_cast<int64_t>(i) * ld + j);
        const double w = tile[tx][y];
        if (!(v == w)) bad = true;
      }
    }
    if (bad) status[mat] = 1;
  }
}

__global__ void nj_lazy_status_kernel(int* status, int n_mat, int value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_mat) status[i] = value;
}

struct LazyPlan {
  int ncta, shift, G, rc, smem;
  int rpt() const { return (rc + NJL_THREADS - 1) / NJL_THREADS; }
};

// cluster size, list width and shared memory for n nodes; ncta = 0 when the matrix does not fit
LazyPlan lazy_plan(int64_t n, int ncta) {
  LazyPlan pl{0, 0, 0, 0, 0};
  if (n > NJL_MAX_N || n < 8) return pl;
  int shift = 0;
  while ((1 << shift) < ncta) ++shift;
  int rc = static_cast<int>((n + 1 + ncta - 1) / ncta);
  rc |= 1;   // odd: lists are stored [group][row]
  for (int G = NJL_MAX_G; G >= 8; G >>= 1) {
    const NjLazySmem L(static_cast<int>(n), G, rc);
    if (L.total <= NJL_SMEM_LIMIT - 4096 && 2 * G * rc >= n && rc <= NJL_MAX_RPT * NJL_THREADS) {   // (static shared memory: ~3 KB; s_node fits over the lists)
      pl = LazyPlan{ncta, shift, G, rc, L.total};
      return pl;
    }
  }
  return pl;
}

template <int G, int RPT, bool PROF>
int lazy_cluster_slots(int ncta, int smem) {
  if (cudaFuncSetAttribute(nj_lazy_kernel<G, RPT, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (ncta > 8 && cudaFuncSetAttribute(nj_lazy_kernel<G, RPT, PROF>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(ncta));
  cfg.blockDim = dim3(NJL_THREADS);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = static_cast<unsigned>(ncta);
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int slots = 0;
  if (cudaOccupancyMaxActiveClusters(&slots, nj_lazy_kernel<G, RPT, PROF>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return slots;
}

// (G, RPT, PROF) -> instantiation
#define NJL_DISPATCH_RPT(G_, CALL)                                   \
  switch (pl.rpt()) {                                                \
    case 1: CALL(G_, 1); break;                                      \
    case 2: CALL(G_, 2); break;                                      \
    default: CALL(G_, 3); break;                                     \
  }
#define NJL_DISPATCH(CALL)                                           \
  switch (pl.G) {                                                    \
    case 64: NJL_DISPATCH_RPT(64, CALL) break;                       \
    case 32: NJL_DISPATCH_RPT(32, CALL) break;                       \
    case 16: NJL_DISPATCH_RPT(16, CALL) break;                       \
    default: NJL_DISPATCH_RPT(8, CALL) break;                        \
  }

int lazy_slots(const LazyPlan& pl) {
  const bool pr = g_njl_prof_enabled;
  int slots = 0;
#define NJL_CALL_SLOTS(G_, R_) slots = pr ? lazy_cluster_slots<G_, R_, true>(pl.ncta, pl.smem) : lazy_cluster_slots<G_, R_, false>(pl.ncta, pl.smem)
  NJL_DISPATCH(NJL_CALL_SLOTS)
#undef NJL_CALL_SLOTS
  return slots;
}

template <int G, int RPT, bool PROF>
cudaError_t lazy_launch(const LazyPlan& pl, int n_mat, const NjLazyParams& P, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(pl.ncta * n_mat));
  cfg.blockDim = dim3(NJL_THREADS);
  cfg.dynamicSmemBytes = static_cast<size_t>(pl.smem);
  cfg.stream = s;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = static_cast<unsigned>(pl.ncta);
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, nj_lazy_kernel<G, RPT, PROF>, P);
}

int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }
int64_t lazy_ldw(int64_t n) { return (n + 15) / 16 * 16; }   // rows start on 128-byte lines

int g_lazy_cluster = -1;   // cluster size the device schedules for this kernel (0: none), probed once

}  // namespace

int64_t nj_lazy_max_n() { return NJL_MAX_N; }

int64_t nj_lazy_work_bytes(int64_t n) {
  if (n > NJL_MAX_N || n < 8) return 0;
  return align256((n + 1) * lazy_ldw(n) * 8) + align256(8 * n);
}

int nj_lazy_run(const double* d_dist, int64_t n, int64_t ld, int n_mat, int64_t mat_stride, int32_t* d_merge,
                double* d_len, void* d_work, int* d_status, unsigned long long* d_stats, cudaStream_t s) {
  pf_count_launch();
  nj_lazy_status_kernel<<<(n_mat + 255) / 256, 256, 0, s>>>(d_status, n_mat, 1);
  PF_CUDA(cudaGetLastError());
  if (n > NJL_MAX_N || n < 8) return PF_OK;
  // cluster size: 16 when the device schedules it, else 8
  LazyPlan pl{0, 0, 0, 0, 0};
  for (int ncta = (g_lazy_cluster > 0 ? g_lazy_cluster : NJL_MAX_CLUSTER); ncta >= 8; ncta >>= 1) {
    const LazyPlan cand = lazy_plan(n, ncta);
    if (cand.ncta == 0) continue;
    if (lazy_slots(cand) > 0) {
      pl = cand;
      break;
    }
    if (g_lazy_cluster > 0) break;
  }
  if (pl.ncta == 0) return PF_OK;
  if (g_lazy_cluster < 0) g_lazy_cluster = pl.ncta;

  int dev = 0, sms = 0;
  PF_CUDA(cudaGetDevice(&dev));
  PF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t ldw = lazy_ldw(n);
  const int64_t work_stride = nj_lazy_work_bytes(n);
  char* w = static_cast<char*>(d_work);
  NjLazyParams P;
  P.W = reinterpret_cast<double*>(w);
  P.ldw = ldw;
  P.w_stride = work_stride / 8;
  P.r0 = reinterpret_cast<const double*>(w + align256((n + 1) * ldw * 8));
  P.r_stride = work_stride / 8;
  P.merge = d_merge;
  P.len = d_len;
  P.status = d_status;
  P.stats = d_stats;
  P.n = static_cast<int>(n);
  P.G = pl.G;
  P.rc = pl.rc;
  P.ncta = pl.ncta;
  P.ncta_shift = pl.shift;

  pf_count_launch(3);
  nj_lazy_status_kernel<<<(n_mat + 255) / 256, 256, 0, s>>>(d_status, n_mat, 0);
  PF_CUDA(cudaGetLastError());
  {
    const int64_t rows = n * n_mat;
    int64_t blocks = pf_ceil_div(rows, 8);
    if (blocks > 8 * sms) blocks = 8 * sms;
    nj_lazy_prep_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(d_dist, ld, mat_stride, static_cast<int>(n), n_mat, P.W, ldw,
                                                                      P.w_stride, const_cast<double*>(P.r0), P.r_stride);
    PF_CUDA(cudaGetLastError());
    const int64_t nt = (n + 31) / 32;
    int64_t tiles = nt * (nt + 1) / 2 * n_mat;
    if (tiles > 16 * sms) tiles = 16 * sms;
    nj_lazy_symmetry_kernel<<<static_cast<unsigned>(tiles), 256, 0, s>>>(d_dist, ld, mat_stride, static_cast<int>(n), n_mat, d_status);
    PF_CUDA(cudaGetLastError());
  }
  pf_count_launch();
  cudaError_t e;
  const bool pr = g_njl_prof_enabled;
#define NJL_CALL_LAUNCH(G_, R_) e = pr ? lazy_launch<G_, R_, true>(pl, n_mat, P, s) : lazy_launch<G_, R_, false>(pl, n_mat, P, s)
  NJL_DISPATCH(NJL_CALL_LAUNCH)
#undef NJL_CALL_LAUNCH
  if (e != cudaSuccess) {
    // the cluster could not be launched: every matrix goes to the full-scan kernel
    cudaGetLastError();
    nj_lazy_status_kernel<<<(n_mat + 255) / 256, 256, 0, s>>>(d_status, n_mat, 1);
    PF_CUDA(cudaGetLastError());
  }
  return PF_OK;
}

// Diagnostic: phase accounting of the bounded-scan kernel. h_out16 = nanoseconds of thread 0 of the cluster's first
// CTA summed over iterations: [1] champions + CTA threshold, [2] bound checks + list evaluations, [3] rescans +
// publish, [4] cluster barrier (with the wait for the slowest CTA), [5] winner, [6] update of row sums / matrix,
// [7] row lists. enable != 0: instrumented kernel from the next call on. Reading resets the counters.
PF_API int pf_nj_lazy_profile(int enable, unsigned long long* h_out16) {
  PF_TRY_BEGIN
  if (h_out16) PF_CUDA(cudaMemcpyFromSymbol(h_out16, g_njl_prof, sizeof(unsigned long long) * 16));
  unsigned long long zero[16] = {0};
  PF_CUDA(cudaMemcpyToSymbol(g_njl_prof, zero, sizeof(zero)));
  g_njl_prof_enabled = enable != 0;
  return PF_OK;
  PF_TRY_END
}
