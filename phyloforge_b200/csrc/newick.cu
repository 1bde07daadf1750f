// Host-side serialisation of an NJ join list as Newick text (no device work): the same output as
// phyloforge_b200/newick.py:merges_to_newick with fmt "%.10g", byte for byte, at a few hundred
// nanoseconds per node instead of microseconds. Leaf labels arrive already quoted. The reference gets this text from `treebest nj`
// (treetool/script.py:359-360 -> SNP_treebest.NHX).
#include "pf_common.cuh"

#include <cstdio>
#include <cstring>
#include <string>

namespace {

void append_len(std::string& out, double v) {
  char buf[40];
  const int k = snprintf(buf, sizeof(buf), "%.10g", v);
  out.append(buf, static_cast<size_t>(k));
}

}  // namespace

// See include/phyloforge_b200.h
PF_API int64_t pf_newick_from_merges(const int32_t* h_merge, const double* h_len, int64_t n, const char* const* names,
                                     const double* h_support, char* h_out, int64_t capacity) {
  try {
    if (!h_merge || !h_len || !names || n < 3) {
      pf_fail("pf_newick_from_merges: bad arguments (n must be >= 3)");
      return -1;
    }
    const int64_t n_steps = n - 2;
    for (int64_t t = 0; t < n_steps; ++t)
      for (int k = 0; k < (t == n_steps - 1 ? 3 : 2); ++k) {
        const int64_t c = h_merge[3 * t + k];
        if (c < 0 || c >= n + t) {
          pf_fail("pf_newick_from_merges: row %lld refers to node %lld", (long long)t, (long long)c);
          return -1;
        }
      }
    std::string out;
    out.reserve(static_cast<size_t>(n) * 40);
    // explicit stack of (node, next child) frames: linear in the size of the tree
    struct Frame { int64_t node; int next; };
    std::vector<Frame> st;
    st.reserve(64);
    const int32_t* root = h_merge + 3 * (n_steps - 1);
    const double* rl = h_len + 3 * (n_steps - 1);
    out += '(';
    for (int k = 0; k < 3; ++k) {
      if (k) out += ',';
      st.push_back(Frame{root[k], 0});
      while (!st.empty()) {
        Frame& f = st.back();
        if (f.node < n) {
          out += names[f.node];
          st.pop_back();
          continue;
        }
        const int64_t t = f.node - n;
        if (f.next == 0) {
          out += '(';
          f.next = 1;
          st.push_back(Frame{h_merge[3 * t + 0], 0});
        } else if (f.next == 1) {
          out += ':';
          append_len(out, h_len[3 * t + 0]);
          out += ',';
          f.next = 2;
          st.push_back(Frame{h_merge[3 * t + 1], 0});
        } else {
          out += ':';
          append_len(out, h_len[3 * t + 1]);
          out += ')';
          if (h_support) {
            char buf[40];
            const int k2 = snprintf(buf, sizeof(buf), "%g", h_support[t]);
            out.append(buf, static_cast<size_t>(k2));
          }
          st.pop_back();
        }
      }
      out += ':';
      append_len(out, rl[k]);
    }
    out += ");";
    const int64_t need = static_cast<int64_t>(out.size());
    if (h_out && capacity > need) memcpy(h_out, out.c_str(), static_cast<size_t>(need) + 1);
    return need;
  } catch (const std::exception& e) {
    pf_fail("pf_newick_from_merges: %s", e.what());
    return -1;
  }
}
