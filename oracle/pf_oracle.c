/* TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the population-phylogeny
 * hot path. Nothing in the product (phyloforge_b200/, treetool/) may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do, and only as the checker or the CPU arm.
 *
 * What it restates
 *  - pack:     the 2-bit code of each reference character (treetool/vcftree.py:26-65 for
 *              SNPs, treetool/svtree.py:58-61,84-120 for SVs) -> canonical sample-major
 *              2-bit matrix and the kernel's tiled bit-plane layout.
 *  - counts:   V / D1 / D2 all-pairs integer matrices.
 *  - distance: p-distance and IBS distance, one fp64 divide.
 *  - NJ:       Saitou-Nei neighbor joining.
 *
 * PARITY UNPINNED for counts/distance/NJ: the reference delegates these to the external
 * binary `treebest nj -b 1000` (treetool/script.py:359-360; bioconda package, version
 * unpinned in meta.yaml:24) whose source and binary are absent here, and the reference
 * holds no tests or golden vectors. This file is a documented restatement of the published
 * algorithms, anchored on the reference's call site; it is checked against textbook
 * known-answer cases (tests/test_oracle.py), not against treebest.
 * The pack stage IS pinned: tests compare it with the unmodified reference converters.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define PFO_API __attribute__((visibility("default")))
#define PFO_TILE 64

/* ------------------------------------------------------------------ synthetic genotypes
 * Integer-only, counter-based: code(i, s) is a pure function of (seed, s, i, n_samples,
 * n_pops, miss_ppm). Mirrors phyloforge_b200/csrc/synth.cu bit for bit. Population
 * structure: n_pops contiguous blocks of samples; ancestral allele frequency per site in
 * [0.05, 0.5); each population's frequency drifts multiplicatively. */
static inline uint32_t pfo_mix(uint32_t x) {
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}

static inline uint32_t pfo_hash3(uint64_t seed, uint64_t site, uint32_t k) {
  uint32_t h = (uint32_t)seed ^ 0x9e3779b9u;
  h = pfo_mix(h ^ (uint32_t)site);
  h = pfo_mix(h + (uint32_t)(site >> 32) * 0x7feb352du + 0x165667b1u);
  h = pfo_mix(h ^ (k * 0x9e3779b1u + (uint32_t)(seed >> 32)));
  return h;
}

static inline uint8_t pfo_synth_code(uint64_t seed, int64_t site, int64_t i, int64_t n_samples,
                                     int n_pops, uint32_t miss_ppm) {
  uint32_t h0 = pfo_hash3(seed, (uint64_t)site, 0xffffffffu);
  int32_t p_anc = 3277 + (int32_t)(h0 % 29491u);
  uint32_t pop = (uint32_t)((i * (int64_t)n_pops) / n_samples);
  uint32_t hk = pfo_hash3(seed, (uint64_t)site, 0x80000000u | pop);
  int32_t z = (int32_t)((hk & 0xffu) + ((hk >> 8) & 0xffu) + ((hk >> 16) & 0xffu) + (hk >> 24)) - 510;
  int32_t p = p_anc + (p_anc * z) / 1024;
  if (p < 655) p = 655;
  if (p > 64880) p = 64880;
  uint32_t hg = pfo_hash3(seed, (uint64_t)site, (uint32_t)i);
  uint32_t g = ((hg & 0xffffu) < (uint32_t)p) + ((hg >> 16) < (uint32_t)p);
  if (miss_ppm) {
    uint32_t hm = pfo_hash3(seed ^ 0xa5a5a5a5deadbeefULL, (uint64_t)site, (uint32_t)i);
    if (hm % 1000000u < miss_ppm) g = 3;
  }
  return (uint8_t)g;
}

/* Block bootstrap (mirrors phyloforge_b200/csrc/bootstrap.cu): weights[r * n_blocks + b] = multiplicity of
 * block b among the n_blocks draws with replacement of replicate replicate0 + r. */
PFO_API void pfo_bootstrap_weights(uint64_t seed, int64_t replicate0, int n_reps, int n_blocks, int32_t* weights) {
  for (int64_t k = 0; k < (int64_t)n_reps * n_blocks; ++k) weights[k] = 0;
  for (int r = 0; r < n_reps; ++r)
    for (int t = 0; t < n_blocks; ++t) {
      uint32_t h = pfo_hash3(seed ^ 0xb007573a9b007573ULL, (uint64_t)(replicate0 + r), (uint32_t)t);
      weights[(int64_t)r * n_blocks + (h % (uint32_t)n_blocks)] += 1;
    }
}

/* site-major: codes[(s - site0) * ld + i] */
PFO_API void pfo_synth_codes(uint8_t* codes, int64_t n_samples, int64_t ld, int64_t site0,
                             int64_t n_sites, uint64_t seed, int n_pops, uint32_t miss_ppm) {
#pragma omp parallel for schedule(static)
  for (int64_t s = 0; s < n_sites; ++s)
    for (int64_t i = 0; i < n_samples; ++i)
      codes[s * ld + i] = pfo_synth_code(seed, site0 + s, i, n_samples, n_pops, miss_ppm);
}

/* sample-major rows for a subset of samples: out[r * ld + s] = code(rows[r], site0 + s) */
PFO_API void pfo_synth_rows(uint8_t* out, const int64_t* rows, int64_t n_rows, int64_t ld,
                            int64_t n_samples, int64_t site0, int64_t n_sites, uint64_t seed,
                            int n_pops, uint32_t miss_ppm) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n_rows; ++r)
    for (int64_t s = 0; s < n_sites; ++s)
      out[r * ld + s] = pfo_synth_code(seed, site0 + s, rows[r], n_samples, n_pops, miss_ppm);
}

/* ------------------------------------------------------------------ pack
 * Canonical packed matrix of SURVEY.md §8(c).1: sample-major, 4 codes per byte, code of
 * column s in bits 2(s%4)..2(s%4)+1 of byte s/4; row stride `stride` bytes (a multiple of
 * 128); padding columns are code 3. codes is sample-major codes[i * ld + s]. */
PFO_API void pfo_pack_canonical(const uint8_t* codes, int64_t n_samples, int64_t n_sites,
                                int64_t ld, uint8_t* packed, int64_t stride) {
  for (int64_t i = 0; i < n_samples; ++i) {
    uint8_t* row = packed + i * stride;
    memset(row, 0xff, (size_t)stride);
    for (int64_t s = 0; s < n_sites; ++s) {
      uint8_t c = codes[i * ld + s] & 3u;
      int sh = 2 * (int)(s & 3);
      row[s >> 2] = (uint8_t)((row[s >> 2] & ~(3u << sh)) | ((unsigned)c << sh));
    }
  }
}

/* The kernels' tiled bit-plane layout (include/phyloforge_b200.h): from sample-major codes. */
PFO_API void pfo_pack_planes(const uint8_t* codes, int64_t n_samples, int64_t n_sites,
                             int64_t ld, uint32_t* planes, int64_t n_words) {
  int64_t n_tiles = (n_samples + PFO_TILE - 1) / PFO_TILE;
  for (int64_t t = 0; t < n_tiles; ++t)
    for (int64_t w = 0; w < n_words; ++w)
      for (int l = 0; l < PFO_TILE; ++l) {
        int64_t i = t * PFO_TILE + l;
        uint32_t lo = 0, hi = 0;
        for (int b = 0; b < 32; ++b) {
          int64_t s = w * 32 + b;
          uint8_t c = (i < n_samples && s < n_sites) ? (codes[i * ld + s] & 3u) : 3u;
          lo |= (uint32_t)(c & 1u) << b;
          hi |= (uint32_t)(c >> 1) << b;
        }
        planes[((t * n_words + w) * 2 + 0) * PFO_TILE + l] = lo;
        planes[((t * n_words + w) * 2 + 1) * PFO_TILE + l] = hi;
      }
}

/* ------------------------------------------------------------------ counts
 * Definitional version: one comparison per (pair, site). codes sample-major.
 * counts = [3][n][ld_c] int64 in the order V, D1, D2. */
PFO_API void pfo_pair_counts_bytes(const uint8_t* codes, int64_t n_samples, int64_t n_sites,
                                   int64_t ld, int64_t* counts, int64_t ld_c) {
  int64_t* V = counts;
  int64_t* D1 = counts + n_samples * ld_c;
  int64_t* D2 = counts + 2 * n_samples * ld_c;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < n_samples; ++i) {
    const uint8_t* a = codes + i * ld;
    for (int64_t j = i; j < n_samples; ++j) {
      const uint8_t* b = codes + j * ld;
      int64_t v = 0, d1 = 0, d2 = 0;
      for (int64_t s = 0; s < n_sites; ++s) {
        unsigned ca = a[s], cb = b[s];
        int valid = (ca != 3u) & (cb != 3u);
        v += valid;
        d1 += valid & (ca != cb);
        d2 += valid & ((ca ^ cb) == 2u) & ((ca & 1u) == 0u);
      }
      V[i * ld_c + j] = V[j * ld_c + i] = v;
      D1[i * ld_c + j] = D1[j * ld_c + i] = d1;
      D2[i * ld_c + j] = D2[j * ld_c + i] = d2;
    }
  }
}

/* Bit-parallel version for large inputs and for the CPU baseline: 64 sites per word,
 * hardware popcount, all host threads. Checked against pfo_pair_counts_bytes in tests.
 * lo/hi: sample-major bit planes, n_samples x n_w64 uint64 each (padding bits = 1/1). */
PFO_API void pfo_planes64_from_codes(const uint8_t* codes, int64_t n_samples, int64_t n_sites,
                                     int64_t ld, uint64_t* lo, uint64_t* hi, int64_t n_w64) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_samples; ++i)
    for (int64_t w = 0; w < n_w64; ++w) {
      uint64_t l = 0, h = 0;
      for (int b = 0; b < 64; ++b) {
        int64_t s = w * 64 + b;
        uint64_t c = s < n_sites ? (codes[i * ld + s] & 3u) : 3u;
        l |= (c & 1u) << b;
        h |= (c >> 1) << b;
      }
      lo[i * n_w64 + w] = l;
      hi[i * n_w64 + w] = h;
    }
}

PFO_API void pfo_pair_counts_bits(const uint64_t* lo, const uint64_t* hi, int64_t n_samples,
                                  int64_t n_w64, int64_t* counts, int64_t ld_c) {
  int64_t* V = counts;
  int64_t* D1 = counts + n_samples * ld_c;
  int64_t* D2 = counts + 2 * n_samples * ld_c;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < n_samples; ++i) {
    const uint64_t* li = lo + i * n_w64;
    const uint64_t* hi_i = hi + i * n_w64;
    for (int64_t j = i; j < n_samples; ++j) {
      const uint64_t* lj = lo + j * n_w64;
      const uint64_t* hj = hi + j * n_w64;
      int64_t v = 0, d1 = 0, d2 = 0;
      for (int64_t w = 0; w < n_w64; ++w) {
        uint64_t valid = ~(li[w] & hi_i[w]) & ~(lj[w] & hj[w]);
        uint64_t dl = li[w] ^ lj[w], dh = hi_i[w] ^ hj[w];
        v += __builtin_popcountll(valid);
        d1 += __builtin_popcountll(valid & (dl | dh));
        d2 += __builtin_popcountll(valid & dh & ~(li[w] | lj[w]));
      }
      V[i * ld_c + j] = V[j * ld_c + i] = v;
      D1[i * ld_c + j] = D1[j * ld_c + i] = d1;
      D2[i * ld_c + j] = D2[j * ld_c + i] = d2;
    }
  }
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm of bench.py sets the count explicitly */
PFO_API void pfo_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

PFO_API int pfo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ distances
 * metric 0: p-distance D1/V; metric 1: IBS (D1+D2)/(2V). V == 0 -> 1.0 (returned count).
 * Diagonal 0. One IEEE divide of exactly converted integers. */
PFO_API int64_t pfo_normalise(const int64_t* counts, int64_t n, int64_t ld_c, int metric,
                              double* dist, int64_t ld_d) {
  const int64_t* V = counts;
  const int64_t* D1 = counts + n * ld_c;
  const int64_t* D2 = counts + 2 * n * ld_c;
  int64_t zero_valid = 0;
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < n; ++j) {
      double d;
      if (i == j) {
        d = 0.0;
      } else if (V[i * ld_c + j] == 0) {
        d = 1.0;
        ++zero_valid;
      } else if (metric == 0) {
        d = (double)D1[i * ld_c + j] / (double)V[i * ld_c + j];
      } else {
        d = (double)(D1[i * ld_c + j] + D2[i * ld_c + j]) / (double)(2 * V[i * ld_c + j]);
      }
      dist[i * ld_d + j] = d;
    }
  return zero_valid;
}

/* ------------------------------------------------------------------ neighbor joining
 * Fixed-order row sum shared with the GPU kernel (SURVEY.md §7 "NJ determinism"):
 * 32 strided partials, each summed in ascending index order, then an xor-butterfly. */
static double pfo_rowsum32(const double* row, int64_t n) {
  double p[32];
  for (int l = 0; l < 32; ++l) {
    double acc = 0.0;
    for (int64_t k = l; k < n; k += 32) acc = acc + row[k];
    p[l] = acc;
  }
  for (int off = 16; off >= 1; off >>= 1) {
    double q[32];
    for (int l = 0; l < 32; ++l) q[l] = p[l] + p[l ^ off];
    memcpy(p, q, sizeof(p));
  }
  return p[0];
}

/* dist: n x ld symmetric, zero diagonal; destroyed. merge: 3*(n-2) ints, len: 3*(n-2)
 * doubles, laid out as documented for pf_nj in include/phyloforge_b200.h. n >= 3.
 *
 * Row sums: r is summed once, in the fixed order of pfo_rowsum32, and then maintained
 * incrementally (as most NJ implementations do), each update a fixed sequence of fp64
 * operations so that the GPU kernel can reproduce it bit for bit:
 *     r_k   <- ((r_k - d_ak) - d_bk) + d_uk            for the untouched nodes k
 *     r_u   <- 0.5 * ((r_a + r_b) - n * d_ab)          for the merged node u (= sum_k d_uk)
 * mode 1 (checker of the checker): recompute every row sum from scratch each iteration. */
static int pfo_nj_impl(double* dist, int64_t n, int64_t ld, int32_t* merge, double* len, int from_scratch) {
  if (n < 3) return 1;
  int64_t na = n;
  int32_t* node = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  double* r = (double*)malloc(sizeof(double) * (size_t)n);
  if (!node || !r) { free(node); free(r); return 2; }
  for (int64_t s = 0; s < n; ++s) node[s] = (int32_t)s;
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < na; ++a) r[a] = pfo_rowsum32(dist + a * ld, na);
  int64_t t = 0;
  while (na > 3) {
    if (from_scratch) {
#pragma omp parallel for schedule(static)
      for (int64_t a = 0; a < na; ++a) r[a] = pfo_rowsum32(dist + a * ld, na);
    }
    const double n2 = (double)(na - 2);
    double best = INFINITY;
    int64_t ba = INT64_MAX, bb = INT64_MAX;
    /* minimum of (q, a, b) in lexicographic order == first strict minimum of a sequential
     * scan in (a, b) order; the order-independent form lets the rows be scanned in parallel */
#pragma omp parallel
    {
      double tq = INFINITY;
      int64_t ta = INT64_MAX, tb = INT64_MAX;
#pragma omp for schedule(dynamic, 8) nowait
      for (int64_t a = 0; a < na; ++a) {
        const double* row = dist + a * ld;
        const double ra = r[a];
        for (int64_t b = a + 1; b < na; ++b) {
          double q = n2 * row[b];
          q = q - ra;
          q = q - r[b];
          if (q < tq || (q == tq && (a < ta || (a == ta && b < tb)))) { tq = q; ta = a; tb = b; }
        }
      }
#pragma omp critical
      {
        if (tq < best || (tq == best && (ta < ba || (ta == ba && tb < bb)))) { best = tq; ba = ta; bb = tb; }
      }
    }
    if (ba == INT64_MAX) { ba = 0; bb = 1; }  /* no finite candidate (NaN input) */
    const double dab = dist[ba * ld + bb];
    const double ra = r[ba], rb = r[bb];
    double tt = ra - rb;
    tt = tt / (2.0 * n2);
    double la = 0.5 * dab;
    la = la + tt;
    const double lb = dab - la;
    merge[3 * t + 0] = node[ba]; merge[3 * t + 1] = node[bb]; merge[3 * t + 2] = 0;
    len[3 * t + 0] = la; len[3 * t + 1] = lb; len[3 * t + 2] = 0.0;
    for (int64_t k = 0; k < na; ++k) {
      if (k == ba || k == bb) continue;
      const double dak = dist[ba * ld + k], dbk = dist[bb * ld + k];
      double d = dak + dbk;
      d = d - dab;
      d = 0.5 * d;
      double rk = r[k] - dak;
      rk = rk - dbk;
      rk = rk + d;
      r[k] = rk;
      dist[ba * ld + k] = d;
      dist[k * ld + ba] = d;
    }
    {
      double ru = ra + rb;
      ru = ru - (double)na * dab;
      r[ba] = 0.5 * ru;
    }
    dist[ba * ld + ba] = 0.0;
    const int64_t last = na - 1;
    if (bb != last) {
      for (int64_t k = 0; k < na; ++k) {
        dist[bb * ld + k] = dist[last * ld + k];
        dist[k * ld + bb] = dist[k * ld + last];
      }
      dist[bb * ld + bb] = 0.0;
      node[bb] = node[last];
      r[bb] = r[last];
    }
    node[ba] = (int32_t)(n + t);
    --na;
    ++t;
  }
  const double d01 = dist[0 * ld + 1], d02 = dist[0 * ld + 2], d12 = dist[1 * ld + 2];
  double l0 = d01 + d02; l0 = l0 - d12; l0 = 0.5 * l0;
  double l1 = d01 + d12; l1 = l1 - d02; l1 = 0.5 * l1;
  double l2 = d02 + d12; l2 = l2 - d01; l2 = 0.5 * l2;
  merge[3 * t + 0] = node[0]; merge[3 * t + 1] = node[1]; merge[3 * t + 2] = node[2];
  len[3 * t + 0] = l0; len[3 * t + 1] = l1; len[3 * t + 2] = l2;
  free(node);
  free(r);
  return 0;
}

PFO_API int pfo_nj(double* dist, int64_t n, int64_t ld, int32_t* merge, double* len) {
  return pfo_nj_impl(dist, n, ld, merge, len, 0);
}

/* same algorithm with every row sum recomputed from scratch each iteration: used by the tests
 * to show that the incremental sums do not change the tree on generic (tie-free) inputs */
PFO_API int pfo_nj_from_scratch(double* dist, int64_t n, int64_t ld, int32_t* merge, double* len) {
  return pfo_nj_impl(dist, n, ld, merge, len, 1);
}
