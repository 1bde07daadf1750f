/* CPU model of the bounded-scan neighbor joining of phyloforge_b200/csrc/nj_lazy.cu (test infrastructure and
 * design tool, not product code): the same per-row state (a list of tracked columns indexed by column mod G, a
 * champion, the bounds R1 over the rest of the list and R2 over the untracked columns in the frame of a float snapshot
 * of s_j = r_j / (n - 2), the drift maximum M), the same maintenance after a join, the same decisions (which lists
 * are evaluated, which rows are rescanned against the threshold of the owning "CTA"), run sequentially and checked
 * join by join against oracle/pf_oracle.c:pfo_nj. It prints how many rescans and list evaluations the scheme needs.
 *
 *   gcc -O2 -o nj_bounded_scan_model tools/nj_bounded_scan_model.c -lm
 *   nj_bounded_scan_model n matrix.f64 [G = 32] [ctas = 16] [epoch = 8] [permutation seed = 0]
 *
 * matrix.f64: n x n doubles, row-major, symmetric. Exit status 0 when every join equals the oracle's.
 * tests/test_nj_bounded_scan_model.py runs it on random, tied and star-shaped matrices. */
#include "../oracle/pf_oracle.c"
#include <stdio.h>

#define GMAX 512
typedef struct { int col[GMAX]; double d[GMAX]; int champ; double R1, R2; } Row;
typedef struct { double q; int64_t a, b; } Best;

static int G = 32, NC = 16, E = 8;
static int64_t n, ld, na;
static double* D; static double* r; static Row* rows; static float* snap; static int* owner;
static double M = 0.0;
static int64_t n_rescan = 0, n_reeval = 0;

/* the oracle's q of the pair of slots (i, j), and the lexicographic (q, a, b) minimum */
static inline void consider(Best* B, double n2, int64_t i, int64_t j, double dd) {
  const int64_t lo = i < j ? i : j, hi = i < j ? j : i;
  double q = n2 * dd; q = q - r[lo]; q = q - r[hi];
  if (q < B->q || (q == B->q && (lo < B->a || (lo == B->a && hi < B->b)))) { B->q = q; B->a = lo; B->b = hi; }
}
static inline double ve_of(int col, double d) { return d - (double)snap[col]; }
static inline void fold(double* R, double ve) { if (ve < *R) *R = ve; }

/* exact evaluation of the list of row i: every entry is a candidate, the best becomes the champion, R1 bounds the rest */
static void list_evaluate(int64_t i, double n2, Best* B) {
  Row* R = &rows[i]; ++n_reeval;
  Best rb = {INFINITY, INT64_MAX, INT64_MAX}; int ch = -1;
  for (int g = 0; g < G; ++g) if (R->col[g] >= 0) {
    const Best before = rb; consider(&rb, n2, i, R->col[g], R->d[g]);
    if (rb.q != before.q || rb.a != before.a || rb.b != before.b) ch = g;
  }
  R->champ = ch; R->R1 = INFINITY;
  for (int g = 0; g < G; ++g) if (R->col[g] >= 0 && g != ch) fold(&R->R1, ve_of(R->col[g], R->d[g]));
  if (ch >= 0) consider(B, n2, i, R->col[ch], R->d[ch]);
}

/* rescan of row i: every column is a candidate of this iteration (the kernel skips those it can prove worse than the
 * threshold: the same set of possible winners); per group the column with the smallest ve is tracked, R2 bounds the rest */
static void rescan(int64_t i, double n2, Best* B) {
  Row* R = &rows[i]; ++n_rescan;
  double bv[GMAX], b2 = INFINITY;
  for (int g = 0; g < G; ++g) { R->col[g] = -1; bv[g] = INFINITY; }
  for (int64_t j = 0; j < na; ++j) {
    if (j == i) continue;
    const double d = D[i * ld + j];
    consider(B, n2, i, j, d);
    const double ve = ve_of((int)j, d);
    const int g = (int)(j % G);
    if (ve < bv[g]) { fold(&b2, bv[g]); bv[g] = ve; R->col[g] = (int)j; R->d[g] = d; } else fold(&b2, ve);
  }
  R->R2 = b2;
  --n_reeval; list_evaluate(i, n2, B);
}

static int bound_clears(double R, double si, double n2, double T) {
  const double lb = n2 * ((R - M) - si);
  const double margin = 1e-9 * (n2 * (fabs(R) + fabs(M) + fabs(si)) + fabs(T)) + 1e-290;
  return (lb - margin) > T;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s n matrix.f64 [G] [ctas] [epoch] [perm]\n", argv[0]); return 2; }
  n = atoll(argv[1]); const char* path = argv[2];
  G = argc > 3 ? atoi(argv[3]) : 32; NC = argc > 4 ? atoi(argv[4]) : 16; E = argc > 5 ? atoi(argv[5]) : 8;
  const int perm = argc > 6 ? atoi(argv[6]) : 0;
  /* what-if switch (design tool only; the kernel does not do this): 1 = in an iteration in which some CTA has to rescan,
   * every CTA that does not refreshes its row with the least slack (in the shadow of the wait); 2 = in every iteration */
  const int proactive = argc > 7 ? atoi(argv[7]) : 0;
  int64_t n_proactive = 0;
  if (G > GMAX || G < 1 || n < 4) return 2;
  ld = n;
  D = malloc(n * n * 8); double* D0 = malloc(n * n * 8);
  FILE* f = fopen(path, "rb");
  if (!f || fread(D0, 8, n * n, f) != (size_t)(n * n)) { fprintf(stderr, "cannot read %s\n", path); return 2; }
  fclose(f);
  if (perm) {   /* random relabelling of the samples */
    int64_t* p = malloc(n * 8); for (int64_t i = 0; i < n; ++i) p[i] = i; srand(perm);
    for (int64_t i = n - 1; i > 0; --i) { const int64_t j = rand() % (i + 1); const int64_t x = p[i]; p[i] = p[j]; p[j] = x; }
    for (int64_t i = 0; i < n; ++i) for (int64_t j = 0; j < n; ++j) D[i * ld + j] = D0[p[i] * ld + p[j]];
    memcpy(D0, D, n * n * 8);
  } else memcpy(D, D0, n * n * 8);
  int32_t* mref = malloc(12 * (n - 2)); double* lref = malloc(24 * (n - 2));
  pfo_nj(D0, n, ld, mref, lref);

  r = malloc(n * 8); rows = malloc(n * sizeof(Row)); int32_t* node = malloc(n * 4); snap = malloc(n * 4); owner = malloc(n * 4);
  for (int64_t a = 0; a < n; ++a) { r[a] = pfo_rowsum32(D + a * ld, n); node[a] = (int32_t)a; owner[a] = (int)(a % NC); }
  na = n;
  { const double inv0 = 1.0 / (double)(na - 2); M = -INFINITY;
    for (int64_t j = 0; j < n; ++j) { snap[j] = (float)(r[j] * inv0); const double dd = r[j] * inv0 - (double)snap[j]; if (dd > M) M = dd; }
    Best dummy = {INFINITY, INT64_MAX, INT64_MAX};
    for (int64_t i = 0; i < n; ++i) rescan(i, (double)(na - 2), &dummy);
    n_rescan = 0; n_reeval = 0; }
  int64_t t = 0, bad = 0, iters_resc = 0, max_resc = 0, path_resc = 0;   /* path_resc: sum over iterations of the most
                                                                          rescans any one CTA does (what the others wait for) */
  int64_t* cta_resc = malloc(NC * 8);
  double* du = malloc(n * 8);
  Best* cb = malloc(NC * sizeof(Best)); Best* cb2 = malloc(NC * sizeof(Best));
  double T0 = INFINITY;   /* exact q of the previous exchange's runner-up in the current frame: a threshold for every CTA */
  while (na > 3) {
    const double n2 = (double)(na - 2), inv = 1.0 / n2;
    /* champions -> the threshold of each CTA */
    for (int c = 0; c < NC; ++c) cb[c] = (Best){INFINITY, INT64_MAX, INT64_MAX};
    for (int64_t i = 0; i < na; ++i) {
      Row* R = &rows[i];
      if (R->champ < 0 && R->R1 == -INFINITY) list_evaluate(i, n2, &cb[owner[i]]);   /* lists that ask for it */
      else if (R->champ >= 0) consider(&cb[owner[i]], n2, i, R->col[R->champ], R->d[R->champ]);
    }
    const int64_t resc0 = n_rescan;
    memcpy(cb2, cb, NC * sizeof(Best));
    memset(cta_resc, 0, NC * 8);
    /* bounds that do not clear the threshold: rescan the row, or evaluate its list */
    for (int64_t i = 0; i < na; ++i) {
      Row* R = &rows[i]; const double T = fmin(cb[owner[i]].q, T0), si = r[i] * inv;
      if (!(R->R2 == INFINITY) && !bound_clears(R->R2, si, n2, T)) { rescan(i, n2, &cb2[owner[i]]); ++cta_resc[owner[i]]; }
      else if (!(R->R1 == INFINITY) && !bound_clears(R->R1, si, n2, T)) list_evaluate(i, n2, &cb2[owner[i]]);
    }
    Best W = {INFINITY, INT64_MAX, INT64_MAX};
    for (int c = 0; c < NC; ++c)
      if (cb2[c].q < W.q || (cb2[c].q == W.q && (cb2[c].a < W.a || (cb2[c].a == W.a && cb2[c].b < W.b)))) W = cb2[c];
    if (n_rescan > resc0) { ++iters_resc; if (n_rescan - resc0 > max_resc) max_resc = n_rescan - resc0; }
    { int64_t m = 0; for (int c = 0; c < NC; ++c) if (cta_resc[c] > m) m = cta_resc[c]; path_resc += m;
      if (getenv("NJ_MODEL_TRACE") && m > 0) fprintf(stderr, "t=%lld na=%lld phase=%lld total=%lld busiest=%lld M=%.3e\n", (long long)t,
                                                     (long long)na, (long long)(t % E), (long long)(n_rescan - resc0), (long long)m, M); }
    if (proactive == 2 || (proactive == 1 && n_rescan > resc0)) {
      for (int c = 0; c < NC; ++c) {
        if (cta_resc[c]) continue;
        int64_t pick = -1; double least = INFINITY;
        for (int64_t i = 0; i < na; ++i) {
          if (owner[i] != c || rows[i].R2 == INFINITY) continue;
          const double slack = n2 * ((rows[i].R2 - M) - r[i] * inv);
          if (slack < least) { least = slack; pick = i; }
        }
        if (pick >= 0) { Best dummy = {INFINITY, INT64_MAX, INT64_MAX}; rescan(pick, n2, &dummy); --n_rescan; ++n_proactive; }
      }
    }
    int64_t ba = W.a, bb = W.b;
    if (ba == INT64_MAX) { ba = 0; bb = 1; }
    /* the runner-up: the best published record whose nodes are not joined now */
    Best RU = {INFINITY, INT64_MAX, INT64_MAX};
    for (int c = 0; c < NC; ++c) {
      const Best x = cb2[c];
      if (x.a == INT64_MAX || x.a == ba || x.a == bb || x.b == ba || x.b == bb) continue;
      if (x.q < RU.q || (x.q == RU.q && (x.a < RU.a || (x.a == RU.a && x.b < RU.b)))) RU = x;
    }
    const double ru_d = (RU.a != INT64_MAX) ? D[RU.a * ld + RU.b] : 0.0;
    if (node[ba] != mref[3 * t] || node[bb] != mref[3 * t + 1]) {
      if (bad < 5) fprintf(stderr, "MISMATCH t=%lld: (%d,%d) vs oracle (%d,%d)\n", (long long)t, node[ba], node[bb], mref[3 * t], mref[3 * t + 1]);
      ++bad;
    }
    /* update (the oracle's arithmetic) */
    const double dab = D[ba * ld + bb]; const double ra = r[ba], rb = r[bb];
    const double inv_new = (na - 3 > 0) ? 1.0 / (double)(na - 3) : 0.0;
    double Mn = -INFINITY;
    for (int64_t k = 0; k < na; ++k) {
      if (k == ba || k == bb) continue;
      const double dak = D[ba * ld + k], dbk = D[bb * ld + k];
      double d = dak + dbk; d = d - dab; d = 0.5 * d;
      double rk = r[k] - dak; rk = rk - dbk; rk = rk + d;
      r[k] = rk; du[k] = d;
      D[ba * ld + k] = d; D[k * ld + ba] = d;
      const double dd = rk * inv_new - (double)snap[k]; if (dd > Mn) Mn = dd;
    }
    { double ru = ra + rb; ru = ru - (double)na * dab; r[ba] = 0.5 * ru; }
    D[ba * ld + ba] = 0.0;
    const int64_t last = na - 1; const int move = (bb != last);
    const int resnap = ((t + 1) % E) == 0;
    double radj = 0.0;
    if (resnap) {   /* new snapshot of the untouched nodes: the stored bounds move into its frame */
      double mx2 = -INFINITY, mn2 = INFINITY;
      for (int64_t k = 0; k < na; ++k) {
        if (k == ba || k == bb) continue;
        const double sj = r[k] * inv_new; snap[k] = (float)sj;
        const double df = sj - (double)snap[k]; if (df > mx2) mx2 = df; if (df < mn2) mn2 = df;
      }
      radj = mn2 - Mn; Mn = mx2;
    }
    const double su = r[ba] * inv_new;
    snap[ba] = (float)(su - Mn);            /* the new column starts at the current maximal drift */
    const double sue = (double)snap[ba];
    M = fmax(Mn, su - sue);
    for (int64_t k = 0; k < na; ++k) {
      if (k == ba || k == bb) continue;
      Row* R = &rows[k];
      if (resnap) { if (R->R1 < INFINITY && R->R1 > -INFINITY) R->R1 += radj; if (R->R2 < INFINITY) R->R2 += radj; }
      /* (1) the joined columns disappear */
      for (int w = 0; w < 2; ++w) {
        const int64_t c = w ? bb : ba; const int g = (int)(c % G);
        if (R->col[g] == c) { R->col[g] = -1; if (R->champ == g) { R->champ = -1; R->R1 = -INFINITY; } }
      }
      /* (2) the column of the last slot becomes column b (another group) */
      if (move) {
        const int gl = (int)(last % G), gb = (int)(bb % G);
        if (R->col[gl] == last) {
          if (gl == gb) R->col[gl] = (int)bb;
          else {
            const double dl = R->d[gl]; const int was_champ = (R->champ == gl);
            R->col[gl] = -1; if (was_champ) R->champ = -1;
            if (R->col[gb] < 0) { R->col[gb] = (int)bb; R->d[gb] = dl; if (was_champ) R->champ = gb; }
            else if (ve_of((int)last, dl) < ve_of(R->col[gb], R->d[gb])) {   /* snap[last] is the moved node's */
              fold(&R->R2, ve_of(R->col[gb], R->d[gb]));
              if (R->champ == gb) { R->champ = -1; R->R1 = -INFINITY; }
              R->col[gb] = (int)bb; R->d[gb] = dl; if (was_champ) R->champ = gb;
            } else {
              fold(&R->R2, ve_of((int)last, dl));
              if (was_champ) R->R1 = -INFINITY;
            }
          }
        }
      }
    }
    if (move) {
      for (int64_t k = 0; k < na; ++k) { D[bb * ld + k] = D[last * ld + k]; D[k * ld + bb] = D[k * ld + last]; }
      D[bb * ld + bb] = 0.0; node[bb] = node[last]; r[bb] = r[last]; du[bb] = du[last];
      rows[bb] = rows[last]; snap[bb] = snap[last]; owner[bb] = owner[last];
    }
    /* (3) the merged node is a new column (slot a) */
    for (int64_t k = 0; k < na - 1; ++k) {
      if (k == ba) continue;
      Row* R = &rows[k]; const int g = (int)(ba % G);
      const double ve_new = du[k] - sue;
      int placed = 0;
      if (R->col[g] < 0) placed = 1;
      else if (ve_new < ve_of(R->col[g], R->d[g])) { fold(&R->R2, ve_of(R->col[g], R->d[g])); if (R->champ == g) R->champ = -1; placed = 1; }
      else fold(&R->R2, ve_new);
      if (placed) {
        R->col[g] = (int)ba; R->d[g] = du[k];
        if (R->champ < 0) R->R1 = -INFINITY;
        else if (ve_new < ve_of(R->col[R->champ], R->d[R->champ])) { fold(&R->R1, ve_of(R->col[R->champ], R->d[R->champ])); R->champ = g; }
        else fold(&R->R1, ve_new);
      }
    }
    { Row* R = &rows[ba]; for (int g = 0; g < G; ++g) R->col[g] = -1; R->champ = -1; R->R1 = INFINITY; R->R2 = INFINITY; }
    node[ba] = (int32_t)(n + t);
    --na; ++t;
    T0 = INFINITY;
    if (RU.a != INT64_MAX && na > 3) {   /* relabelled, in the frame of the next iteration */
      if (move && RU.a == last) RU.a = bb;
      if (move && RU.b == last) RU.b = bb;
      Best x = {INFINITY, INT64_MAX, INT64_MAX};
      consider(&x, (double)(na - 2), RU.a, RU.b, ru_d);
      if (x.q < INFINITY) T0 = x.q;
    }
  }
  if (proactive) printf("proactive refreshes (mode %d): %.3f / iteration\n", proactive, (double)n_proactive / t);
  printf("n=%lld G=%d ctas=%d epoch=%d perm=%d: mismatches %lld; rescans %.3f / iteration (iterations with one: %lld of %lld, max %lld); "
         "on the busiest CTA %.3f / iteration; list evaluations %.2f / iteration\n", (long long)n, G, NC, E, perm, (long long)bad,
         (double)n_rescan / t, (long long)iters_resc, (long long)t, (long long)max_resc, (double)path_resc / t, (double)n_reeval / t);
  return bad != 0;
}
