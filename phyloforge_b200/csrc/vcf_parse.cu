// VCF text -> genotype characters and 2-bit codes on the GPU.
//
// Replaces the per-line / per-genotype Python loops of the reference:
//   SNP  VcfTree.convert_vcf_to_seq + generate_ambiguous_code  treetool/vcftree.py:96-134, 26-65
//   SV   SvTree.convert_vcf_to_seq                              treetool/svtree.py:37-120
// and the whole-file line splitting of cutFile / readlines (vcftree.py:75-76, svtree.py:35-36).
//
// One warp per line. The warp walks the line 32 bytes at a time; a ballot of the whitespace
// bytes gives the field starts, a popcount prefix gives each start its field index, and the
// lane that owns the first byte of a sample field parses that genotype on its own (GT tokens
// are a handful of bytes). Fields are runs of non-whitespace, as `line.strip().split()`.
// Inputs on which the unmodified reference raises or corrupts its rows are flagged per line
// (PF_LINE_ERR_*) instead of being emulated; the host raises.
#include "pf_common.cuh"

#include <mutex>

namespace {

constexpr int LINES_PER_CTA = 8;  // 256 threads

__device__ __forceinline__ bool is_ws(uint32_t c) {
  return c == ' ' || (c >= 9 && c <= 13);  // '\n' never occurs inside a line
}
__device__ __forceinline__ uint32_t upper(uint32_t c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }
__device__ __forceinline__ bool is_base(uint32_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
__device__ __forceinline__ bool is_exotic_int_char(uint32_t c) { return c == '+' || c == '-' || c == '_' || c >= 0x80; }

// character VcfTree.generate_ambiguous_code returns for two (upper-cased) allele bytes;
// 0 = the reference returns None and crashes
__device__ __forceinline__ uint32_t iupac(uint32_t b1, uint32_t b2) {
  if (b1 == b2) {
    if (is_base(b1)) return b1;
    if (b1 == '.') return '-';
    if (b1 == '*') return 'N';
    return 0;
  }
  const bool in1 = is_base(b1), in2 = is_base(b2);
  if (!in1 && !in2) return 'N';
  if (in1 && in2) {
    const uint32_t lo = b1 < b2 ? b1 : b2, hi = b1 < b2 ? b2 : b1;
    if (lo == 'A') return hi == 'C' ? 'M' : (hi == 'G' ? 'R' : 'W');
    if (lo == 'C') return hi == 'G' ? 'S' : 'Y';
    return 'K';
  }
  const uint32_t base = in1 ? b1 : b2, other = in1 ? b2 : b1;
  if (other == '*') return base + 32;  // lower-case
  return 'N';
}

struct LineSpan {
  int64_t begin;  // first non-whitespace byte
  int64_t end;    // one past the last non-whitespace byte
  bool comment;   // raw line starts with '#'
};

// strip() of line k; all lanes return the same value
__device__ LineSpan line_span(const uint8_t* __restrict__ text, int64_t n_bytes,
                              const int64_t* __restrict__ line_start, int64_t n_lines, int64_t k) {
  LineSpan s;
  int64_t b = line_start[k];
  int64_t e = (k + 1 < n_lines) ? line_start[k + 1] - 1 : n_bytes;  // exclude the '\n'
  if (e > b && k + 1 >= n_lines && text[e - 1] == '\n') --e;         // file ends with '\n'
  s.comment = (e > b) && text[b] == '#';
  while (e > b && is_ws(text[e - 1])) --e;
  while (b < e && is_ws(text[b])) ++b;
  s.begin = b;
  s.end = e;
  return s;
}

// Walk fields until `want` field starts have been seen; records the byte offsets of the
// starts of fields f0, f1, f2 (all < want). Returns number of fields seen (<= want).
__device__ int find_fields(const uint8_t* __restrict__ text, const LineSpan& sp, int lane, int want, int f0,
                           int f1, int f2, int64_t& o0, int64_t& o1, int64_t& o2) {
  int count = 0;
  uint32_t carry = 1;  // previous byte is whitespace / line start
  o0 = o1 = o2 = -1;
  for (int64_t pos = sp.begin; pos < sp.end && count < want; pos += 32) {
    const int64_t p = pos + lane;
    const uint32_t c = p < sp.end ? text[p] : ' ';
    const uint32_t ws = __ballot_sync(0xffffffffu, is_ws(c));
    const uint32_t starts = ~ws & ((ws << 1) | carry);
    carry = ws >> 31;
    const int idx = count + __popc(starts & ((1u << lane) - 1u));
    const bool mine = (starts >> lane) & 1u;
    const uint32_t m0 = __ballot_sync(0xffffffffu, mine && idx == f0);
    const uint32_t m1 = __ballot_sync(0xffffffffu, mine && idx == f1);
    const uint32_t m2 = __ballot_sync(0xffffffffu, mine && idx == f2);
    if (m0) o0 = pos + __ffs(m0) - 1;
    if (m1) o1 = pos + __ffs(m1) - 1;
    if (m2) o2 = pos + __ffs(m2) - 1;
    count += __popc(starts);
  }
  return count;
}

__device__ __forceinline__ int64_t field_end(const uint8_t* __restrict__ text, int64_t p, int64_t end) {
  while (p < end && !is_ws(text[p])) ++p;
  return p;
}

// ------------------------------------------------------------------ SNP
__global__ void __launch_bounds__(32 * LINES_PER_CTA)
snp_parse_kernel(const uint8_t* __restrict__ text, int64_t n_bytes, const int64_t* __restrict__ line_start,
                 int64_t n_lines, int64_t n_samples, uint8_t* __restrict__ chars, uint8_t* __restrict__ codes,
                 int64_t ld, int32_t* __restrict__ line_info) {
  const int lane = threadIdx.x & 31;
  const int64_t k = static_cast<int64_t>(blockIdx.x) * LINES_PER_CTA + (threadIdx.x >> 5);
  if (k >= n_lines) return;
  const LineSpan sp = line_span(text, n_bytes, line_start, n_lines, k);
  if (sp.comment || sp.begin >= sp.end) {  // vcftree.py:104-106
    if (lane == 0) line_info[k] = 0;
    return;
  }
  int64_t o_ref, o_alt, o_gt;
  const int seen = find_fields(text, sp, lane, 10, 3, 4, 9, o_ref, o_alt, o_gt);
  if (seen < 5) {  // fields[3] / fields[4] would raise IndexError
    if (lane == 0) line_info[k] = PF_LINE_ERR_FIELDS << 8;
    return;
  }
  // alleles = [REF] + ALT.split(','), every lane redundantly (a few bytes)   vcftree.py:108-119
  uint32_t alleles[8];
  int n_alleles = 0;
  bool keep = true, too_many = false, empty_allele = false;
  {
    const int64_t e = field_end(text, o_ref, sp.end);
    if (e - o_ref > 1) keep = false;
    alleles[0] = upper(text[o_ref]);
    n_alleles = 1;
    int64_t p = o_alt;
    const int64_t ea = field_end(text, o_alt, sp.end);
    while (keep) {
      int64_t q = p;
      while (q < ea && text[q] != ',') ++q;
      if (q - p > 1) {
        keep = false;
        break;
      }
      if (q == p) empty_allele = true;
      if (n_alleles < 8) alleles[n_alleles] = (q > p) ? upper(text[p]) : 0u;
      else too_many = true;
      ++n_alleles;
      if (q >= ea) break;
      p = q + 1;
    }
  }
  if (!keep) {  // a multi-base allele: the whole site is skipped
    if (lane == 0) line_info[k] = 0;
    return;
  }
  int err = 0;
  if (too_many) err = PF_LINE_ERR_TOO_MANY_ALLELES;
  else if (empty_allele) err = PF_LINE_ERR_ALLELE_CHAR;
  if (n_alleles > 8) n_alleles = 8;
  const uint32_t ref = alleles[0], alt = alleles[1];
  const bool fits = (n_alleles == 2) && is_base(ref) && (is_base(alt) || alt == '.') && ref != alt;

  // sample fields
  int count = 9;
  uint32_t carry = 1;
  uint8_t* crow = chars + k * ld;
  uint8_t* grow = codes + k * ld;
  if (seen >= 10) {
    for (int64_t pos = o_gt; pos < sp.end; pos += 32) {
      const int64_t p = pos + lane;
      const uint32_t c = p < sp.end ? text[p] : ' ';
      const uint32_t ws = __ballot_sync(0xffffffffu, is_ws(c));
      const uint32_t starts = ~ws & ((ws << 1) | carry);
      carry = ws >> 31;
      const int64_t i = count - 9 + __popc(starts & ((1u << lane) - 1u));
      count += __popc(starts);
      if (((starts >> lane) & 1u) && i < n_samples) {
        // GT = text before ':' ; two integers separated by '/' or '|'      vcftree.py:122-129
        int nparts = 1, nd0 = 0, nd1 = 0;
        uint32_t v0 = 0, v1 = 0;
        bool bad = false, exotic = false;
        int64_t q = p;
        for (; q < sp.end; ++q) {
          const uint32_t ch = text[q];
          if (ch == ':' || is_ws(ch)) break;
          if (ch == '/' || ch == '|') {
            ++nparts;
          } else if (ch >= '0' && ch <= '9') {
            if (nparts == 1) { v0 = min(v0 * 10 + (ch - '0'), 100000u); ++nd0; }
            else if (nparts == 2) { v1 = min(v1 * 10 + (ch - '0'), 100000u); ++nd1; }
          } else {
            bad = true;
            if (is_exotic_int_char(ch)) exotic = true;
          }
        }
        uint32_t out_c, out_g;
        if (q == p) {
          err = PF_LINE_ERR_EMPTY_GT;
          out_c = '-'; out_g = 3;
        } else if (exotic) {
          err = PF_LINE_ERR_NEG_INDEX;
          out_c = '-'; out_g = 3;
        } else if (bad || nparts != 2 || nd0 == 0 || nd1 == 0) {
          out_c = '-'; out_g = 3;  // ValueError path -> ('.', '.') -> '-'
        } else if (v0 >= static_cast<uint32_t>(n_alleles) || v1 >= static_cast<uint32_t>(n_alleles)) {
          err = PF_LINE_ERR_ALLELE_INDEX;
          out_c = '-'; out_g = 3;
        } else {
          const uint32_t b1 = alleles[v0], b2 = alleles[v1];
          out_c = iupac(b1, b2);
          if (out_c == 0) {
            err = PF_LINE_ERR_ALLELE_CHAR;
            out_c = '-';
          }
          if (!fits) out_g = 3;
          else if (v0 == v1) out_g = (v0 == 0) ? 0u : (is_base(alt) ? 2u : 3u);
          else out_g = is_base(alt) ? 1u : 3u;
        }
        crow[i] = static_cast<uint8_t>(out_c);
        grow[i] = static_cast<uint8_t>(out_g);
      }
    }
  }
  if (count - 9 != n_samples && err == 0) err = PF_LINE_ERR_FIELDS;
  // combine the lanes' error classes (any non-zero; lowest lane wins)
  const uint32_t em = __ballot_sync(0xffffffffu, err != 0);
  if (em) err = __shfl_sync(0xffffffffu, err, __ffs(em) - 1);
  if (lane == 0) line_info[k] = 1 | (fits ? 4 : 0) | (err << 8);
}

// ------------------------------------------------------------------ SV
__device__ __forceinline__ bool str_eq(const uint8_t* __restrict__ t, int64_t p, int64_t e, const char* lit, int n) {
  if (e - p != n) return false;
  for (int i = 0; i < n; ++i)
    if (t[p + i] != static_cast<uint8_t>(lit[i])) return false;
  return true;
}

// int(allele) against the DEL/INS map: returns '0'/'1'/'.', sets err on IndexError / exotic
__device__ __forceinline__ uint32_t sv_char(const uint8_t* __restrict__ t, int64_t p, int64_t e, bool is_del, int& err) {
  if (e <= p) return '.';
  uint32_t v = 0;
  for (int64_t q = p; q < e; ++q) {
    const uint32_t ch = t[q];
    if (ch >= '0' && ch <= '9') {
      v = min(v * 10 + (ch - '0'), 100000u);
    } else {
      if (is_exotic_int_char(ch)) err = PF_LINE_ERR_NEG_INDEX;
      return '.';
    }
  }
  if (v >= 2) {
    err = PF_LINE_ERR_ALLELE_INDEX;
    return '.';
  }
  return (v == 0) ? (is_del ? '1' : '0') : (is_del ? '0' : '1');   // svtree.py:58-61
}

__global__ void __launch_bounds__(32 * LINES_PER_CTA)
sv_parse_kernel(const uint8_t* __restrict__ text, int64_t n_bytes, const int64_t* __restrict__ line_start,
                int64_t n_lines, int64_t n_samples, uint8_t* __restrict__ chars, uint8_t* __restrict__ codes,
                int64_t ld, int32_t* __restrict__ line_info) {
  const int lane = threadIdx.x & 31;
  const int64_t k = static_cast<int64_t>(blockIdx.x) * LINES_PER_CTA + (threadIdx.x >> 5);
  if (k >= n_lines) return;
  const LineSpan sp = line_span(text, n_bytes, line_start, n_lines, k);
  if (sp.comment || sp.begin >= sp.end) {
    if (lane == 0) line_info[k] = 0;
    return;
  }
  int64_t o_info, o_unused, o_gt;
  const int seen = find_fields(text, sp, lane, 10, 7, -1, 9, o_info, o_unused, o_gt);
  if (seen < 8) {  // fields[7] would raise IndexError
    if (lane == 0) line_info[k] = PF_LINE_ERR_FIELDS << 8;
    return;
  }
  // INFO: items split on ';', each split on '=' must give exactly two parts, until SVTYPE
  int type = 0;  // 0 none/other, 1 DEL, 2 INS                                 svtree.py:47-57
  int err = 0;
  {
    const int64_t e = field_end(text, o_info, sp.end);
    int64_t p = o_info;
    while (true) {
      int64_t q = p, eq = -1;
      int n_eq = 0;
      while (q < e && text[q] != ';') {
        if (text[q] == '=') {
          if (n_eq == 0) eq = q;
          ++n_eq;
        }
        ++q;
      }
      if (n_eq != 1) {
        err = PF_LINE_ERR_SV_INFO;
        break;
      }
      if (str_eq(text, p, eq, "SVTYPE", 6)) {
        if (str_eq(text, eq + 1, q, "DEL", 3)) type = 1;
        else if (str_eq(text, eq + 1, q, "INS", 3)) type = 2;
        break;
      }
      if (q >= e) break;
      p = q + 1;
    }
  }
  if (err) {
    if (lane == 0) line_info[k] = err << 8;
    return;
  }
  if (type == 0) {
    if (lane == 0) line_info[k] = 0;
    return;
  }
  const bool is_del = (type == 1);
  int count = 9;
  uint32_t carry = 1;
  bool hom = true;
  uint8_t* c1row = chars + (2 * k) * ld;
  uint8_t* c2row = chars + (2 * k + 1) * ld;
  uint8_t* g1row = codes + (2 * k) * ld;
  uint8_t* g2row = codes + (2 * k + 1) * ld;
  if (seen >= 10) {
    for (int64_t pos = o_gt; pos < sp.end; pos += 32) {
      const int64_t p = pos + lane;
      const uint32_t c = p < sp.end ? text[p] : ' ';
      const uint32_t ws = __ballot_sync(0xffffffffu, is_ws(c));
      const uint32_t starts = ~ws & ((ws << 1) | carry);
      carry = ws >> 31;
      const int64_t i = count - 9 + __popc(starts & ((1u << lane) - 1u));
      count += __popc(starts);
      if (((starts >> lane) & 1u) && i < n_samples) {
        // token before ':'; '/' has precedence over '|'; pieces 0 and 1    svtree.py:69-77
        int64_t te = p;
        int64_t first_slash = -1, first_bar = -1;
        for (; te < sp.end; ++te) {
          const uint32_t ch = text[te];
          if (ch == ':' || is_ws(ch)) break;
          if (ch == '/' && first_slash < 0) first_slash = te;
          if (ch == '|' && first_bar < 0) first_bar = te;
        }
        uint32_t ch1 = '.', ch2 = '.';
        if (te == p) {
          err = PF_LINE_ERR_EMPTY_GT;
        } else if (first_slash < 0 && first_bar < 0) {
          err = PF_LINE_ERR_SV_HAPLOID;
        } else {
          const uint32_t sep = first_slash >= 0 ? '/' : '|';
          const int64_t s1 = first_slash >= 0 ? first_slash : first_bar;
          int64_t e2 = s1 + 1;
          while (e2 < te && text[e2] != sep) ++e2;
          // allele1 = [p, s1), allele2 = [s1 + 1, e2)
          bool same = (s1 - p) == (e2 - s1 - 1);
          if (same)
            for (int64_t x = 0; x < s1 - p; ++x)
              if (text[p + x] != text[s1 + 1 + x]) { same = false; break; }
          if (!same) hom = false;                                          // svtree.py:79-82
          ch1 = sv_char(text, p, s1, is_del, err);
          ch2 = sv_char(text, s1 + 1, e2, is_del, err);
        }
        c1row[i] = static_cast<uint8_t>(ch1);
        c2row[i] = static_cast<uint8_t>(ch2);
        g1row[i] = ch1 == '0' ? 0 : (ch1 == '1' ? 2 : 3);
        g2row[i] = ch2 == '0' ? 0 : (ch2 == '1' ? 2 : 3);
      }
    }
  }
  if (count - 9 != n_samples && err == 0) err = PF_LINE_ERR_FIELDS;
  const uint32_t em = __ballot_sync(0xffffffffu, err != 0);
  if (em) err = __shfl_sync(0xffffffffu, err, __ffs(em) - 1);
  const bool all_hom = __all_sync(0xffffffffu, hom);
  if (lane == 0) line_info[k] = (all_hom ? 1 : 2) | (err << 8);
}

// ------------------------------------------------------------------ line index
constexpr int LI_THREADS = 256;
constexpr int LI_BYTES_PER_THREAD = 64;
constexpr int LI_CHUNK = LI_THREADS * LI_BYTES_PER_THREAD;

__device__ __forceinline__ int count_nl(const uint8_t* __restrict__ text, int64_t b, int64_t e, int64_t n_bytes) {
  int c = 0;
  for (int64_t p = b; p < e; ++p) c += (text[p] == '\n' && p + 1 < n_bytes);
  return c;
}

__global__ void __launch_bounds__(LI_THREADS)
li_count_kernel(const uint8_t* __restrict__ text, int64_t n_bytes, int64_t* __restrict__ block_counts) {
  __shared__ int s[LI_THREADS / 32];
  const int64_t b = static_cast<int64_t>(blockIdx.x) * LI_CHUNK + static_cast<int64_t>(threadIdx.x) * LI_BYTES_PER_THREAD;
  const int64_t e = min(b + LI_BYTES_PER_THREAD, n_bytes);
  int c = b < n_bytes ? count_nl(text, b, e, n_bytes) : 0;
  for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < LI_THREADS / 32; ++w) t += s[w];
    block_counts[blockIdx.x] = t;
  }
}

// single-CTA exclusive scan of int64 values in place; total -> out_total[0]
__global__ void __launch_bounds__(1024)
scan_i64_kernel(int64_t* __restrict__ v, int64_t n, int64_t* __restrict__ out_total) {
  __shared__ int64_t s_warp[32];
  __shared__ int64_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t idx = base + threadIdx.x;
    const int64_t x = idx < n ? v[idx] : 0;
    int64_t incl = x;
    for (int off = 1; off < 32; off <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int64_t w = s_warp[lane];
      int64_t wi = w;
      for (int off = 1; off < 32; off <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, wi, off);
        if (lane >= off) wi += y;
      }
      s_warp[lane] = wi - w;  // exclusive prefix of warp totals
    }
    __syncthreads();
    const int64_t carry = s_carry;
    const int64_t excl = carry + s_warp[warp] + incl - x;
    if (idx < n) v[idx] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) out_total[0] = s_carry;
}

__global__ void __launch_bounds__(LI_THREADS)
li_write_kernel(const uint8_t* __restrict__ text, int64_t n_bytes, const int64_t* __restrict__ block_offsets,
                int64_t* __restrict__ line_start, int64_t capacity) {
  __shared__ int s[LI_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * LI_CHUNK + static_cast<int64_t>(threadIdx.x) * LI_BYTES_PER_THREAD;
  const int64_t e = min(b + LI_BYTES_PER_THREAD, n_bytes);
  const int c = b < n_bytes ? count_nl(text, b, e, n_bytes) : 0;
  int incl = c;
  for (int off = 1; off < 32; off <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += y;
  }
  if (lane == 31) s[warp] = incl;
  __syncthreads();
  int wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += s[w];
  int64_t out = 1 + block_offsets[blockIdx.x] + wbase + incl - c;  // line 0 starts at byte 0
  if (blockIdx.x == 0 && threadIdx.x == 0 && capacity > 0) line_start[0] = 0;
  for (int64_t p = b; p < e; ++p)
    if (text[p] == '\n' && p + 1 < n_bytes) {
      if (out < capacity) line_start[out] = p + 1;
      ++out;
    }
}

// ------------------------------------------------------------------ row selection
struct SelSummary {
  long long total_cols;
  long long n_err;
  long long n_unfit;
  long long first_err_line;
  long long first_err_class;
};

__global__ void __launch_bounds__(1024)
select_rows_kernel(const int32_t* __restrict__ line_info, int64_t n_lines, int rows_per_line, int require_fit,
                   int64_t* __restrict__ rows, int64_t capacity, SelSummary* __restrict__ summary) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  __shared__ long long s_err, s_unfit;
  __shared__ unsigned long long s_first;
  if (threadIdx.x == 0) {
    s_carry = 0; s_err = 0; s_unfit = 0; s_first = ~0ull;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < n_lines; base += 1024) {
    const int64_t k = base + threadIdx.x;
    int cols = 0;
    if (k < n_lines) {
      const int32_t info = line_info[k];
      cols = PF_LINE_COLS(info);
      if (PF_LINE_ERR(info)) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&s_err), 1ull);
        atomicMin(&s_first, (static_cast<unsigned long long>(k) << 8) | PF_LINE_ERR(info));
      }
      if (require_fit && cols && !PF_LINE_FITS2BIT(info)) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&s_unfit), 1ull);
        cols = 0;
      }
    }
    long long incl = cols;
    for (int off = 1; off < 32; off <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const long long w = s_warp[lane];
      long long wi = w;
      for (int off = 1; off < 32; off <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, wi, off);
        if (lane >= off) wi += y;
      }
      s_warp[lane] = wi - w;
    }
    __syncthreads();
    const long long excl = s_carry + s_warp[warp] + incl - cols;
    for (int c = 0; c < cols; ++c)
      if (excl + c < capacity) rows[excl + c] = k * rows_per_line + c;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + cols;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    summary->total_cols = s_carry;
    summary->n_err = s_err;
    summary->n_unfit = s_unfit;
    summary->first_err_line = (s_first == ~0ull) ? -1 : static_cast<long long>(s_first >> 8);
    summary->first_err_class = (s_first == ~0ull) ? 0 : static_cast<long long>(s_first & 0xff);
  }
}

}  // namespace

PF_API int64_t pf_line_index_work_bytes(int64_t n_bytes) {
  if (n_bytes < 0) return 0;
  return (pf_ceil_div(n_bytes, LI_CHUNK) + 2) * static_cast<int64_t>(sizeof(int64_t));
}

PF_API int pf_line_index(const uint8_t* d_text, int64_t n_bytes, int64_t* d_line_start, int64_t capacity,
                         int64_t* h_n_lines, void* d_work, void* stream) {
  PF_TRY_BEGIN
  if (!h_n_lines) return pf_fail("pf_line_index: h_n_lines is NULL");
  *h_n_lines = 0;
  if (n_bytes == 0) return PF_OK;
  if (!d_text || !d_line_start || !d_work || n_bytes < 0 || capacity < 1)
    return pf_fail("pf_line_index: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n_blocks = pf_ceil_div(n_bytes, LI_CHUNK);
  if (n_blocks > 0x7fffffffLL) return pf_fail("pf_line_index: text too large");
  int64_t* counts = static_cast<int64_t*>(d_work);
  int64_t* total = counts + n_blocks;
  pf_count_launch();
  li_count_kernel<<<static_cast<unsigned>(n_blocks), LI_THREADS, 0, s>>>(d_text, n_bytes, counts);
  PF_CUDA(cudaGetLastError());
  pf_count_launch();
  scan_i64_kernel<<<1, 1024, 0, s>>>(counts, n_blocks, total);
  PF_CUDA(cudaGetLastError());
  int64_t h_total = 0;
  PF_CUDA(cudaMemcpyAsync(&h_total, total, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  PF_CUDA(cudaStreamSynchronize(s));
  const int64_t n_lines = h_total + 1;
  if (n_lines > capacity)
    return pf_fail_code(PF_ERR_INPUT, "pf_line_index: %lld lines exceed capacity %lld", (long long)n_lines,
                        (long long)capacity);
  pf_count_launch();
  li_write_kernel<<<static_cast<unsigned>(n_blocks), LI_THREADS, 0, s>>>(d_text, n_bytes, counts, d_line_start,
                                                                         capacity);
  PF_CUDA(cudaGetLastError());
  *h_n_lines = n_lines;
  return PF_OK;
  PF_TRY_END
}

static int parse_args_ok(const void* t, int64_t n_bytes, const void* ls, int64_t n_lines, int64_t n_samples,
                         const void* ch, const void* co, int64_t ld, const void* li) {
  return t && ls && ch && co && li && n_bytes >= 0 && n_lines >= 0 && n_samples > 0 && ld >= n_samples;
}

PF_API int pf_vcf_snp_parse(const uint8_t* d_text, int64_t n_bytes, const int64_t* d_line_start, int64_t n_lines,
                            int64_t n_samples, uint8_t* d_chars, uint8_t* d_codes, int64_t ld,
                            int32_t* d_line_info, void* stream) {
  PF_TRY_BEGIN
  if (n_lines == 0) return PF_OK;
  if (!parse_args_ok(d_text, n_bytes, d_line_start, n_lines, n_samples, d_chars, d_codes, ld, d_line_info))
    return pf_fail("pf_vcf_snp_parse: bad arguments");
  const int64_t grid = pf_ceil_div(n_lines, LINES_PER_CTA);
  if (grid > 0x7fffffffLL) return pf_fail("pf_vcf_snp_parse: too many lines");
  pf_count_launch();
  snp_parse_kernel<<<static_cast<unsigned>(grid), 32 * LINES_PER_CTA, 0, static_cast<cudaStream_t>(stream)>>>(
      d_text, n_bytes, d_line_start, n_lines, n_samples, d_chars, d_codes, ld, d_line_info);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

PF_API int pf_vcf_sv_parse(const uint8_t* d_text, int64_t n_bytes, const int64_t* d_line_start, int64_t n_lines,
                           int64_t n_samples, uint8_t* d_chars, uint8_t* d_codes, int64_t ld,
                           int32_t* d_line_info, void* stream) {
  PF_TRY_BEGIN
  if (n_lines == 0) return PF_OK;
  if (!parse_args_ok(d_text, n_bytes, d_line_start, n_lines, n_samples, d_chars, d_codes, ld, d_line_info))
    return pf_fail("pf_vcf_sv_parse: bad arguments");
  const int64_t grid = pf_ceil_div(n_lines, LINES_PER_CTA);
  if (grid > 0x7fffffffLL) return pf_fail("pf_vcf_sv_parse: too many lines");
  pf_count_launch();
  sv_parse_kernel<<<static_cast<unsigned>(grid), 32 * LINES_PER_CTA, 0, static_cast<cudaStream_t>(stream)>>>(
      d_text, n_bytes, d_line_start, n_lines, n_samples, d_chars, d_codes, ld, d_line_info);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}

PF_API int pf_vcf_select_rows(const int32_t* d_line_info, int64_t n_lines, int rows_per_line, int require_fit,
                              int64_t* d_rows, int64_t capacity, int64_t* h_summary, void* stream) {
  PF_TRY_BEGIN
  if (!h_summary) return pf_fail("pf_vcf_select_rows: h_summary is NULL");
  h_summary[0] = h_summary[1] = h_summary[2] = 0;
  h_summary[3] = -1;
  h_summary[4] = 0;
  if (n_lines == 0) return PF_OK;
  if (!d_line_info || !d_rows || n_lines < 0 || capacity < 0 || (rows_per_line != 1 && rows_per_line != 2))
    return pf_fail("pf_vcf_select_rows: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // the 40-byte device summary is allocated once per device and kept for the life of the process (a cudaMalloc /
  // cudaFree pair per chunk synchronised the whole device every call); calls on one device are serialised by its lock
  static std::mutex mu[64];   // one per device: worker threads on different devices do not wait for each other
  static SelSummary* d_sum_of[64] = {nullptr};
  int dev = 0;
  PF_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return pf_fail("pf_vcf_select_rows: device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu[dev]);
  if (!d_sum_of[dev]) PF_CUDA(cudaMalloc(&d_sum_of[dev], sizeof(SelSummary)));
  SelSummary* d_sum = d_sum_of[dev];
  pf_count_launch();
  select_rows_kernel<<<1, 1024, 0, s>>>(d_line_info, n_lines, rows_per_line, require_fit, d_rows, capacity, d_sum);
  cudaError_t e = cudaGetLastError();
  SelSummary h{};
  if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d_sum, sizeof(h), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  PF_CUDA(e);
  h_summary[0] = h.total_cols;
  h_summary[1] = h.n_err;
  h_summary[2] = h.n_unfit;
  h_summary[3] = h.first_err_line;
  h_summary[4] = h.first_err_class;
  if (h.total_cols > capacity)
    return pf_fail_code(PF_ERR_INPUT, "pf_vcf_select_rows: %lld columns exceed capacity %lld",
                        (long long)h.total_cols, (long long)capacity);
  return PF_OK;
  PF_TRY_END
}
