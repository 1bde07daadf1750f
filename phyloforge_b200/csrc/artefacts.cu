// Host-side helpers of the artefacts written beside the GPU tree (no device work): the PHYLIP distance matrix and the
// bootstrap support values.
//
// Writer of the PHYLIP square distance matrix: `n`, then per sample its label, two blanks
// and the n distances as "%.<decimals>f" separated by single blanks - byte for byte what
// phyloforge_b200/phylip.py:write_distance_matrix writes with Python's own formatting, at tens of nanoseconds per
// number on several host threads instead of a microsecond on one. The distance file stands beside the tree the
// reference gets from `treebest nj` (treetool/script.py:359-360); at 3,000 samples it holds 9 M
// numbers, which took the Python writer longer than the whole GPU path takes to compute them.
#include "pf_common.cuh"

#include <cerrno>
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

// 10^k for k <= 18
constexpr unsigned long long kPow10[19] = {1ull,
                                           10ull,
                                           100ull,
                                           1000ull,
                                           10000ull,
                                           100000ull,
                                           1000000ull,
                                           10000000ull,
                                           100000000ull,
                                           1000000000ull,
                                           10000000000ull,
                                           100000000000ull,
                                           1000000000000ull,
                                           10000000000000ull,
                                           100000000000000ull,
                                           1000000000000000ull,
                                           10000000000000000ull,
                                           100000000000000000ull,
                                           1000000000000000000ull};

// "%.<dec>f" of v, correctly rounded (round-half-even on the exact binary value, as glibc's printf and CPython's
// float formatting both do). Fast path: 0 <= v < 1024 and dec <= 10 -> v * 10^dec = m * 10^dec * 2^e exactly in 128-bit
// integers (m < 2^53, 10^dec < 2^34). Everything else goes through snprintf; NaN is written as Python writes it.
void append_fixed(std::string& out, double v, int dec) {
  if (v != v) {
    out += "nan";
    return;
  }
  if (v >= 0.0 && v < 1024.0 && dec <= 10 && !std::signbit(v)) {
    unsigned long long bits;
    memcpy(&bits, &v, sizeof(bits));
    const int be = static_cast<int>(bits >> 52);                // biased exponent (sign bit is clear here)
    const unsigned long long frac = bits & ((1ull << 52) - 1);
    const unsigned long long m = be ? (frac | (1ull << 52)) : frac;   // v = m * 2^-sh
    const int sh = be ? 1075 - be : 1074;                       // >= 43 because v < 2^10
    unsigned long long q = 0;
    if (m != 0 && sh < 128) {                                   // (sh >= 128: far below half a unit of the last place)
      const unsigned __int128 one = 1;
      const unsigned __int128 p = static_cast<unsigned __int128>(m) * kPow10[dec];
      const unsigned __int128 rem = p & ((one << sh) - 1);
      const unsigned __int128 half = one << (sh - 1);
      q = static_cast<unsigned long long>(p >> sh);
      if (rem > half || (rem == half && (q & 1ull))) ++q;
    }
    // q = round(v * 10^dec): digits from the back, the point after `dec` of them, at least one digit in front
    char buf[48];
    int pos = 48;
    for (int d = 0; d < dec; ++d) {
      buf[--pos] = static_cast<char>('0' + q % 10);
      q /= 10;
    }
    if (dec > 0) buf[--pos] = '.';
    do {
      buf[--pos] = static_cast<char>('0' + q % 10);
      q /= 10;
    } while (q);
    out.append(buf + pos, static_cast<size_t>(48 - pos));
    return;
  }
  char buf[400];
  const int k = snprintf(buf, sizeof(buf), "%.*f", dec, v);
  out.append(buf, static_cast<size_t>(k < 0 ? 0 : (k < static_cast<int>(sizeof(buf)) ? k : static_cast<int>(sizeof(buf)) - 1)));
}

}  // namespace

// See include/phyloforge_b200.h
PF_API int pf_write_distance_matrix(const char* path, const char* const* names, int64_t n, const double* h_dist,
                                    int64_t ld, int decimals, int n_threads) {
  PF_TRY_BEGIN
  if (!path || !names || n < 0 || (n > 0 && !h_dist) || ld < n || decimals < 0 || decimals > 17)
    return pf_fail("pf_write_distance_matrix: bad arguments");
  int threads = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  if (threads < 1) threads = 1;
  if (threads > 64) threads = 64;
  if (static_cast<int64_t>(threads) * 64 > n) threads = static_cast<int>(n / 64 > 0 ? n / 64 : 1);   // >= 64 rows per thread
  FILE* f = fopen(path, "wb");
  if (!f) return pf_fail_code(PF_ERR_INPUT, "pf_write_distance_matrix: cannot open %s (%s)", path, strerror(errno));
  bool ok = fprintf(f, "%lld\n", static_cast<long long>(n)) > 0;
  // rows in slabs: every thread formats a contiguous run of rows of the slab into its own buffer, the buffers are
  // written in order; a slab is bounded so that the text in flight stays around 64 MB
  const int64_t per_row = n * (decimals + 3) + 64;
  int64_t slab = (64ll << 20) / (per_row > 0 ? per_row : 1);
  if (slab < threads) slab = threads;
  std::vector<std::string> bufs(static_cast<size_t>(threads));
  for (int64_t r0 = 0; r0 < n && ok; r0 += slab) {
    const int64_t r1 = (r0 + slab < n) ? r0 + slab : n;
    auto work = [&](int t) {
      const int64_t rows = r1 - r0;
      const int64_t b = r0 + rows * t / threads, e = r0 + rows * (t + 1) / threads;
      std::string out = std::move(bufs[static_cast<size_t>(t)]);   // a local object: the strings of `bufs` share cache lines
      out.clear();
      out.reserve(static_cast<size_t>((e - b) * per_row));
      for (int64_t i = b; i < e; ++i) {
        out += names[i] ? names[i] : "";
        out += "  ";
        const double* row = h_dist + i * ld;
        for (int64_t j = 0; j < n; ++j) {
          if (j) out += ' ';
          append_fixed(out, row[j], decimals);
        }
        out += '\n';
      }
      bufs[static_cast<size_t>(t)] = std::move(out);
    };
    if (threads == 1) {
      work(0);
    } else {
      std::vector<std::thread> pool;
      pool.reserve(static_cast<size_t>(threads));
      for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
      for (auto& th : pool) th.join();
    }
    for (int t = 0; t < threads && ok; ++t)
      ok = bufs[static_cast<size_t>(t)].empty() ||
           fwrite(bufs[static_cast<size_t>(t)].data(), 1, bufs[static_cast<size_t>(t)].size(), f) == bufs[static_cast<size_t>(t)].size();
  }
  if (fclose(f) != 0) ok = false;
  if (!ok) return pf_fail_code(PF_ERR_INPUT, "pf_write_distance_matrix: writing %s failed (%s)", path, strerror(errno));
  return PF_OK;
  PF_TRY_END
}

// ---- bootstrap support: which internal edges of the main tree each replicate tree has
// The same split keys as phyloforge_b200/bootstrap.py:split_keys (two SplitMix64 keys per leaf, summed over the leaf set
// of a node with wrap-around, taken on the side without leaf 0, folded to 64 bits), computed per replicate in one pass
// over its join list and looked up in the main tree's sorted keys; replicates are dealt out over host threads. The
// reference asks `treebest nj -b 1000` for these numbers (treetool/script.py:359-360); 1,000 replicates of 3,000
// leaves took the numpy version about as long as the GPUs needed to join them.
namespace {

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct SplitScratch {
  std::vector<uint64_t> h1, h2;
  std::vector<uint8_t> has0;
};

// keys[t] for t < n - 3 of one join list; false when a row refers to a node that does not exist yet
bool split_keys_of(const int32_t* merge, int64_t n, const uint64_t* k1, const uint64_t* k2, uint64_t t1, uint64_t t2,
                   SplitScratch& S, uint64_t* keys) {
  S.h1.resize(static_cast<size_t>(2 * n));
  S.h2.resize(static_cast<size_t>(2 * n));
  S.has0.resize(static_cast<size_t>(2 * n));
  for (int64_t i = 0; i < n; ++i) {
    S.h1[i] = k1[i];
    S.h2[i] = k2[i];
    S.has0[i] = i == 0;
  }
  for (int64_t t = 0; t < n - 3; ++t) {
    const int64_t a = merge[3 * t + 0], b = merge[3 * t + 1];
    if (a < 0 || b < 0 || a >= n + t || b >= n + t) return false;
    const uint64_t s1 = S.h1[a] + S.h1[b], s2 = S.h2[a] + S.h2[b];
    const uint8_t z = S.has0[a] | S.has0[b];
    S.h1[n + t] = s1;
    S.h2[n + t] = s2;
    S.has0[n + t] = z;
    const uint64_t a1 = z ? t1 - s1 : s1, a2 = z ? t2 - s2 : s2;
    keys[t] = a1 ^ (a2 * 0x9E3779B97F4A7C15ull);
  }
  return true;
}

}  // namespace

// See include/phyloforge_b200.h
PF_API int pf_bootstrap_support(const int32_t* h_main_merge, const int32_t* h_rep_merges, int64_t n, int64_t n_reps,
                                double* h_support, int n_threads) {
  PF_TRY_BEGIN
  if (!h_main_merge || (n_reps > 0 && !h_rep_merges) || !h_support || n < 4 || n_reps < 0)
    return pf_fail("pf_bootstrap_support: bad arguments (n must be >= 4)");
  const int64_t ne = n - 3;                       // internal edges of a tree
  std::vector<uint64_t> k1(static_cast<size_t>(n)), k2(static_cast<size_t>(n));
  uint64_t t1 = 0, t2 = 0;
  for (int64_t i = 0; i < n; ++i) {
    k1[i] = splitmix64(static_cast<uint64_t>(2 * i + 1));
    k2[i] = splitmix64(0x5851F42D4C957F2Dull ^ static_cast<uint64_t>(2 * i));
    t1 += k1[i];
    t2 += k2[i];
  }
  SplitScratch S0;
  std::vector<uint64_t> main_keys(static_cast<size_t>(ne));
  if (!split_keys_of(h_main_merge, n, k1.data(), k2.data(), t1, t2, S0, main_keys.data()))
    return pf_fail_code(PF_ERR_INPUT, "pf_bootstrap_support: the main join list refers to a node that does not exist yet");
  std::vector<uint64_t> uniq(main_keys);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  const int64_t nu = static_cast<int64_t>(uniq.size());
  int threads = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  if (threads < 1) threads = 1;
  if (threads > 64) threads = 64;
  if (threads > n_reps) threads = static_cast<int>(n_reps > 0 ? n_reps : 1);
  std::vector<std::vector<int64_t>> hits(static_cast<size_t>(threads));
  std::vector<int> bad(static_cast<size_t>(threads), 0);
  auto work = [&](int w) {
    std::vector<int64_t> mine(static_cast<size_t>(nu), 0);            // local: the threads' vectors share cache lines
    std::vector<int64_t> seen(static_cast<size_t>(nu), -1);           // last replicate that held this key: counted once
    std::vector<uint64_t> keys(static_cast<size_t>(ne));
    SplitScratch S;
    for (int64_t r = n_reps * w / threads; r < n_reps * (w + 1) / threads; ++r) {
      if (!split_keys_of(h_rep_merges + r * 3 * (n - 2), n, k1.data(), k2.data(), t1, t2, S, keys.data())) {
        bad[static_cast<size_t>(w)] = 1;
        return;
      }
      for (int64_t t = 0; t < ne; ++t) {
        const auto it = std::lower_bound(uniq.begin(), uniq.end(), keys[t]);
        if (it != uniq.end() && *it == keys[t]) {
          const int64_t u = it - uniq.begin();
          if (seen[u] != r) {
            seen[u] = r;
            ++mine[u];
          }
        }
      }
    }
    hits[static_cast<size_t>(w)] = std::move(mine);
  };
  if (threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; ++w) pool.emplace_back(work, w);
    for (auto& th : pool) th.join();
  }
  for (int w = 0; w < threads; ++w)
    if (bad[static_cast<size_t>(w)])
      return pf_fail_code(PF_ERR_INPUT, "pf_bootstrap_support: a replicate join list refers to a node that does not exist yet");
  const double denom = static_cast<double>(n_reps > 0 ? n_reps : 1);
  for (int64_t t = 0; t < ne; ++t) {
    const int64_t u = std::lower_bound(uniq.begin(), uniq.end(), main_keys[t]) - uniq.begin();
    int64_t total = 0;
    for (int w = 0; w < threads; ++w)
      if (!hits[static_cast<size_t>(w)].empty()) total += hits[static_cast<size_t>(w)][static_cast<size_t>(u)];
    h_support[t] = 100.0 * static_cast<double>(total) / denom;
  }
  return PF_OK;
  PF_TRY_END
}
