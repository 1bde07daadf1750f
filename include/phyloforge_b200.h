/* libphyloforge_b200.so — C ABI of the B200-native population-phylogeny hot path.
 *
 * The reference (wangyayaya/PhyloForge) has no FFI on this path: its boundary is two
 * Python methods and a shell command line. Each entry point below
 * replaces one step of that path and cites the reference lines it stands in for
 * (paths relative to the reference root). INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every function returns 0 (PF_OK) or a non-zero status; pf_last_error() gives the
 *    message of the last failure on the calling thread. There is no CPU fallback.
 *  - pointers named d_* are DEVICE pointers owned by the caller (PyTorch tensors on the
 *    Python side); h_* are host pointers. No entry point allocates memory that outlives
 *    the call unless documented. `stream` is a cudaStream_t passed as void* (NULL = the
 *    default stream). Calls are asynchronous on `stream` unless documented.
 *  - devices and threads: every call acts on the calling thread's CURRENT device (cudaSetDevice); a host that
 *    drives several GPUs calls from one thread per device (phyloforge_b200/multigpu.py: gpu_workers = threads).
 *    Calls from different threads on different devices do not wait for each other; the A/B and profiling switches
 *    (pf_*_set_impl, pf_*_profile, pf_*_timing) are process-wide and meant for single-threaded diagnostic runs.
 *  - genotype codes: 0 hom-REF, 1 HET, 2 hom-ALT, 3 missing.
 *  - packed layout ("planes"): uint32 words; sample tile T = i / PF_TILE (64), site word
 *    w = s / 32, plane p (0 = low code bit L, 1 = high code bit H):
 *        planes[((T * n_words + w) * 2 + p) * 64 + (i % 64)], bit (s % 32) = bit p of code(i, s)
 *    padding samples and padding sites hold code 3 (L = H = 1).
 */
#ifndef PHYLOFORGE_B200_H
#define PHYLOFORGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
#define PF_API extern "C" __attribute__((visibility("default")))
#else
#define PF_API extern
#endif

#define PF_TILE 64

/* ------------------------------------------------------------------ status / info */
PF_API const char* pf_last_error(void);
PF_API const char* pf_version(void);
PF_API int pf_tile(void);
PF_API int pf_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem,
                          int* l2_bytes);

/* ------------------------------------------------------------------ layout helpers */
/* number of uint32 words in a planes buffer for n_samples x n_words */
PF_API int64_t pf_planes_len(int64_t n_samples, int64_t n_words);

/* ------------------------------------------------------------------ synthetic genotypes
 * Counter-based, integer-only generator (bit-identical to oracle/pf_oracle.c:pfo_synth_codes):
 * code(i, s) depends only on (seed, global site s, sample i), so any sharding of the site
 * axis yields the same global matrix (SURVEY.md §8(d) "Synthetic inputs").
 * Writes site-major codes d_codes[(s - site0) * ld + i] for s in [site0, site0 + n_sites). */
PF_API int pf_synth_codes(uint8_t* d_codes, int64_t n_samples, int64_t ld, int64_t site0,
                          int64_t n_sites, uint64_t seed, int n_pops, uint32_t miss_ppm,
                          void* stream);

/* ------------------------------------------------------------------ genotype packer
 * Replaces the per-genotype re-encoding loop of VcfTree.convert_vcf_to_seq
 * (treetool/vcftree.py:96-134) and SvTree.convert_vcf_to_seq (treetool/svtree.py:37-120)
 * once genotypes are byte codes: site-major u8 codes -> bit planes.
 * d_codes[r * ld + i] is the code of sample i in row r. If d_rows != NULL the k-th packed
 * site is row d_rows[k] (gather of surviving sites, vcftree.py:113-119), else row k.
 * Packs n_sites sites into words [word0, word0 + ceil(n_sites/32)) of a planes buffer that
 * has n_words words per tile; sites beyond n_sites in the last word get code 3.
 * Codes > 3 are a caller error (masked to 2 bits). */
PF_API int pf_pack_codes_u8(const uint8_t* d_codes, int64_t ld, const int64_t* d_rows,
                            int64_t n_samples, int64_t n_sites, uint32_t* d_planes,
                            int64_t n_words, int64_t word0, void* stream);

/* Same packer for genotypes that are already 2 bits wide on the host side: site-major rows of ld2
 * bytes (ld2 % 4 == 0), 4 samples per byte, sample 4b+k in bits 2k..2k+1 of byte b. plink = 0: the
 * codes are this library's (0 hom-REF, 1 HET, 2 hom-ALT, 3 missing); plink = 1: PLINK .bed codes
 * (00 hom A1, 01 missing, 10 het, 11 hom A2; the first three magic bytes of a .bed are not data). */
PF_API int pf_pack_codes_2bit(const uint8_t* d_codes2, int64_t ld2, int64_t n_samples, int64_t n_sites,
                              uint32_t* d_planes, int64_t n_words, int64_t word0, int plink,
                              void* stream);

/* d_dst = d_src moved by `shift` (0..31) sites towards higher site indices: site s of the source is site
 * s + shift of the destination; sites [0, shift) and everything behind n_sites + shift are code 3. Used when the
 * site axis of ONE file is parsed by several GPUs (treetool/vcftree.py:77-93 cuts it over worker processes): a rank
 * packs its own columns from site 0 and, once the column counts of the ranks before it are known, moves them onto
 * the global 32-site word grid, so that word ranges (bootstrap blocks) mean the same sites for any number of GPUs.
 * The buffers must not overlap; d_dst needs ceil((n_sites + shift) / 32) <= n_words_dst words per tile. */
PF_API int pf_planes_shift(const uint32_t* d_src, int64_t n_words_src, uint32_t* d_dst, int64_t n_words_dst,
                           int64_t n_samples, int64_t n_sites, int shift, void* stream);

/* planes -> sample-major u8 codes d_out[i * ld_out + s] for s in [0, n_sites) taken from
 * words [word0, ...). Used to emit the reference's on-disk artefacts and by parity tests. */
PF_API int pf_unpack_codes_u8(const uint32_t* d_planes, int64_t n_words, int64_t word0,
                              int64_t n_samples, int64_t n_sites, uint8_t* d_out,
                              int64_t ld_out, void* stream);

/* u8 matrix transpose with row gather: d_out[i * ld_out + k] = d_in[d_rows[k] * ld_in + i]
 * (d_rows NULL -> row k). Site-major characters -> the per-sample sequences the reference
 * writes (vcftree.py:133-143, svtree.py:122-125). */
PF_API int pf_transpose_u8(const uint8_t* d_in, int64_t ld_in, const int64_t* d_rows,
                           int64_t n_rows, int64_t n_cols, uint8_t* d_out, int64_t ld_out,
                           void* stream);

/* ------------------------------------------------------------------ VCF text on the GPU
 * Line index of a text buffer: d_line_start[k] = byte offset of line k (k < *n_lines),
 * d_line_start[n_lines] = n_bytes sentinel is NOT written; line k ends at the '\n' before
 * line k+1 or at n_bytes. Replaces read().splitlines() of VcfTree.cutFile
 * (vcftree.py:75-76) / readlines() of svtree.py:35-36. Synchronous (returns the count).
 * d_work must hold pf_line_index_work_bytes(n_bytes) bytes. capacity = max lines. */
PF_API int64_t pf_line_index_work_bytes(int64_t n_bytes);
PF_API int pf_line_index(const uint8_t* d_text, int64_t n_bytes, int64_t* d_line_start,
                         int64_t capacity, int64_t* h_n_lines, void* d_work, void* stream);

/* Per-line status written by the parsers (d_line_info[k]):
 *  bits 0-1  number of matrix columns the line yields (SNP: 0/1; SV: 0/1/2)
 *  bit  2    SNP: the kept site fits the 2-bit code (biallelic A/C/G/T, or ALT '.')
 *  bits 8-15 error class, 0 = none (see PF_LINE_ERR_*). A line with an error is one on
 *            which the unmodified reference raises (-> "Failed to Convert", exit 1) or
 *            desynchronises its rows; the host raises on any error. */
#define PF_LINE_COLS(x) ((x) & 3)
#define PF_LINE_FITS2BIT(x) (((x) >> 2) & 1)
#define PF_LINE_ERR(x) (((x) >> 8) & 0xff)
#define PF_LINE_ERR_FIELDS 1      /* fewer than 10 fields or sample count != n_samples */
#define PF_LINE_ERR_ALLELE_INDEX 2 /* allele index out of range (IndexError in the reference) */
#define PF_LINE_ERR_ALLELE_CHAR 3  /* homozygous non-ACGT/./'*' allele (reference returns None) */
#define PF_LINE_ERR_EMPTY_GT 4     /* empty first FORMAT subfield (reference desynchronises rows) */
#define PF_LINE_ERR_TOO_MANY_ALLELES 5 /* more than 8 single-base alleles */
#define PF_LINE_ERR_SV_INFO 6      /* INFO item without exactly one '=' before SVTYPE (ValueError) */
#define PF_LINE_ERR_SV_HAPLOID 7   /* GT without '/' or '|' (reference re-uses stale alleles) */
#define PF_LINE_ERR_NEG_INDEX 8    /* signed / exotic integer the reference would accept */

/* SNP lines -> reference characters and 2-bit codes, site-major.
 * For line k (rows of d_chars/d_codes, leading dimension ld >= n_samples):
 *   d_chars[k*ld+i] = the character VcfTree.generate_ambiguous_code yields
 *                     (vcftree.py:26-65 via 120-133), d_codes[k*ld+i] = its 2-bit code.
 * Lines that start with '#', are blank, or have an allele longer than one base
 * (vcftree.py:113-119) get 0 columns. One CTA per line. */
PF_API int pf_vcf_snp_parse(const uint8_t* d_text, int64_t n_bytes, const int64_t* d_line_start,
                            int64_t n_lines, int64_t n_samples, uint8_t* d_chars,
                            uint8_t* d_codes, int64_t ld, int32_t* d_line_info, void* stream);

/* SV lines -> presence/absence characters ('0','1','.') and codes (0, 2, 3), two rows per
 * line: row 2k is the first column of line k, row 2k+1 the second (only meaningful when
 * the line yields 2 columns). svtree.py:37-120. */
PF_API int pf_vcf_sv_parse(const uint8_t* d_text, int64_t n_bytes, const int64_t* d_line_start,
                           int64_t n_lines, int64_t n_samples, uint8_t* d_chars,
                           uint8_t* d_codes, int64_t ld, int32_t* d_line_info, void* stream);

/* Alignment characters -> codes, for callers that hold the character matrix the reference writes
 * (all_snp.fa: IUPAC; SV.matrix: '0' '1' '.') instead of a VCF — the input of `treebest nj`
 * (treetool/script.py:359-360). d_chars is SITE-major (row = alignment column; use pf_transpose_u8 on a
 * sample-major matrix). Per column the two homozygous symbols become 0 and 2 (alphabetical order),
 * their IUPAC code 1, '-' 'N' '.' 3. d_line_info[s] = 1 | (4 if the column fits 2 bits), so that
 * pf_vcf_select_rows / pf_pack_codes_u8 apply unchanged. */
PF_API int pf_chars_to_codes(const uint8_t* d_chars, int64_t ld, int64_t n_sites, int64_t n_samples,
                             uint8_t* d_codes, int64_t ldc, int32_t* d_line_info, void* stream);

/* Counts for the columns that do NOT fit the 2-bit code (more than two alleles, '*'-derived lower-case characters),
 * from the characters of the all_snp.fa the reference hands to `treebest nj` (treetool/vcftree.py:26-65 writes them -
 * each stands for one unordered allele pair over {A, C, G, T, *}, BASELINE.md section 5 - treetool/script.py:359-360
 * runs the tool). For a pair and a column where neither character is missing ('-', 'N', '.'): D1 = the characters
 * differ (mismatch model, p-distance), D1 + D2 = 2 - alleles shared (allele-sharing / IBS distance; on a biallelic
 * column |g_i - g_j|), i.e. D2 = no allele shared.
 * pf_chars_state_codes: for one character state, d_codes[k][i] = 0 where d_chars[rows[k]][i] == state, 2 where it is
 * another non-missing character, 3 where it is missing; d_rows may be NULL (all rows in order).
 * pf_chars_allele_codes: for one allele (0..4 = A, C, G, T, *), d_codes[k][i] = copies of it in the genotype (0, 1, 2),
 * 3 where the character is missing or stands for no genotype.
 * Feeding these codes through pf_pack_codes_u8 + pf_pair_counts, once per state present into one zeroed
 * [3][n][ld_state] buffer and once per allele into another, gives V summed n_states times, D1 = twice the mismatches
 * and (allele buffer) D1 + D2 = twice (2 - shared); pf_counts_add_extra adds V / n_states and D1 / 2 to the V and D1
 * planes of d_counts and, when d_allele_counts is not NULL, (D1 + D2) / 2 - mismatches to its D2 plane (all exact). */
PF_API int pf_chars_state_codes(const uint8_t* d_chars, int64_t ld, const int64_t* d_rows, int64_t n_rows,
                                int64_t n_samples, int state, uint8_t* d_codes, int64_t ldc, void* stream);
PF_API int pf_chars_allele_codes(const uint8_t* d_chars, int64_t ld, const int64_t* d_rows, int64_t n_rows,
                                 int64_t n_samples, int allele, uint8_t* d_codes, int64_t ldc, void* stream);
PF_API int pf_counts_add_extra(const int32_t* d_state_counts, int64_t ld_state, int n_states,
                               const int32_t* d_allele_counts, int64_t ld_allele, int32_t* d_counts, int64_t n_samples,
                               int64_t ld, void* stream);

/* Exclusive scan of PF_LINE_COLS over lines -> gather rows. d_rows[c] = source row of matrix
 * column c (SNP: line index; SV: 2*line + which). If require_fit != 0 SNP lines whose
 * FITS2BIT bit is clear are left out (and counted in h_summary[2]).
 * Synchronous. h_summary[0] = total columns, [1] = number of lines with an error,
 * [2] = kept lines that do not fit 2 bits, [3] = first line with an error (or -1),
 * [4] = its error class. rows_per_line = 1 (SNP) or 2 (SV). Keeps one 40-byte device buffer per
 * device for the life of the process (the only allocation of this library that outlives a call). */
PF_API int pf_vcf_select_rows(const int32_t* d_line_info, int64_t n_lines, int rows_per_line,
                              int require_fit, int64_t* d_rows, int64_t capacity,
                              int64_t* h_summary, void* stream);

/* ------------------------------------------------------------------ all-pairs counts
 * The arithmetic the reference delegates to `treebest nj` (treetool/script.py:359-360):
 * for every sample pair (i, j), over the sites of words [word_begin, word_end):
 *   V  = #sites where both codes != 3
 *   D1 = #valid sites where the codes differ
 *   D2 = #valid sites where {code_i, code_j} = {0, 2}
 * ACCUMULATES (+=, integer atomics, order-independent) into d_counts, three n x ld int32
 * matrices laid out [3][n_samples][ld] in the order V, D1, D2; both triangles and the
 * diagonal are written. The caller zeroes d_counts before the first call; multi-GPU
 * callers sum the matrices of all site shards (NCCL allreduce).
 * variant: 0 = auto, 1 = plain popcount kernel, 2 = carry-save kernel. */
PF_API int pf_pair_counts(const uint32_t* d_planes, int64_t n_samples, int64_t n_words,
                          int64_t word_begin, int64_t word_end, int32_t* d_counts, int64_t ld,
                          int variant, void* stream);

/* Tensor-core variant of pf_pair_counts (tcgen05 kind::mxf4 indicator GEMMs on 4-bit E2M1 operands;
 * csrc/pair_counts_tc.cu). ACCUMULATES the same V / D1 / D2 as pf_pair_counts.
 * allow_general != 0: fully asynchronous; *h_used = 1. A second GEMM launch (TT, VV) does its work only
 * when the converter finds genotypes missing in some samples only (pf_pair_counts_tc_was_general tells).
 * allow_general == 0: synchronises `stream` once to read that flag; *h_used = 0 = declined (such
 * missingness exists): nothing was added and the caller runs pf_pair_counts.
 * d_work: pf_pair_counts_tc_work_bytes(n_samples, word_end - word_begin) bytes of scratch (the 4-bit
 * operand stream). */
PF_API int64_t pf_pair_counts_tc_work_bytes(int64_t n_samples, int64_t n_words_range);
PF_API int pf_pair_counts_tc(const uint32_t* d_planes, int64_t n_samples, int64_t n_words,
                             int64_t word_begin, int64_t word_end, int32_t* d_counts, int64_t ld,
                             void* d_work, int64_t work_bytes, int allow_general, int* h_used,
                             void* stream);
/* After a pf_pair_counts_tc call on `stream` with the same d_work / sizes: *h_general = 1 if that call
 * found sample-specific missingness (the second launch worked), else 0. Synchronises `stream`. */
PF_API int pf_pair_counts_tc_was_general(const void* d_work, int64_t n_samples, int64_t n_words_range,
                                         int* h_general, void* stream);

/* Diagnostic for the tensor-core kernel: enable/disable role-level wait-cycle accounting and fetch
 * (then clear) the 16 counters accumulated since the last call; see csrc/pair_counts_tc.cu. */
PF_API int pf_pair_counts_tc_profile(int enable, long long* h_out16);

/* Diagnostic / A-B switch for pf_pair_counts_tc: impl 1 = the single-CTA GEMM kernel (128 x 192 sample
 * tiles), impl 2 = the CTA-pair kernel (cluster of two CTAs, tcgen05.mma.cta_group::2, 256 x 240 tiles;
 * the default); cfg 0..4 = tile width / pipeline geometry of the pair kernel (0 = default: 256 x 208 tiles, A rows
 * by TMA tensor-map copies, one A plane in tensor memory). The workspace layout depends on impl and cfg:
 * query pf_pair_counts_tc_work_bytes again after changing it. Results are identical. */
PF_API int pf_pair_counts_tc_set_impl(int impl, int cfg);

/* Diagnostic: enable != 0 makes later pf_pair_counts_tc calls record CUDA events on their stream around the
 * operand-stream conversion and around the GEMM launch(es). With h_ms2 != NULL the call first waits for the
 * last recorded call and returns h_ms2[0] = conversion ms, h_ms2[1] = GEMM ms (bench.py's roofline). */
PF_API int pf_pair_counts_tc_timing(int enable, float* h_ms2);

/* counts -> fp64 distances, one IEEE divide of exactly converted integers per pair:
 *   metric 0: p-distance D1 / V;  metric 1: IBS distance (D1 + D2) / (2 V).
 * V == 0 -> 1.0 and the pair is counted in *d_n_zero_valid (int32, caller-zeroed).
 * The diagonal is 0. d_dist is n x ld_dist. */
PF_API int pf_normalise(const int32_t* d_counts, int64_t n_samples, int64_t ld, int metric,
                        double* d_dist, int64_t ld_dist, int32_t* d_n_zero_valid, void* stream);

/* ------------------------------------------------------------------ neighbor joining
 * Saitou-Nei / Studier-Keppler NJ exactly as restated in oracle/pf_oracle.c:pfo_nj
 * (the stand-in for the absent `treebest nj`, script.py:359-360): fp64, no FMA
 * contraction, row sums in the fixed 32-lane strided + butterfly order, Q argmin ties ->
 * lowest slot i then lowest slot j, merged node replaces slot i, last slot moves into j.
 * d_dist (n x ld, symmetric, zero diagonal) may be destroyed. Outputs (device):
 *   d_merge[3*t+0..2] for t < n-3: node ids (a, b) joined at step t and 0;
 *   d_merge[3*(n-3)+0..2]: the three node ids of the final trifurcation;
 *   d_len likewise: branch lengths of those children.
 * Node ids: leaves 0..n-1; the node created at step t is n + t. n >= 3.
 * d_work must hold pf_nj_work_bytes(n) bytes (for n <= 11,000 this includes a private copy of the matrix,
 * 8 (n + 1) n bytes rounded up).
 * Two implementations with identical results (join order, branch lengths, ties): from 1,536 to 11,000 nodes the
 * bounded-scan kernel (csrc/nj_lazy.cu: one thread-block cluster per matrix, per-row candidate lists with lower
 * bounds, the matrix read only where a bound does not exclude the minimum) runs first; it hands a matrix over to
 * the full-scan kernels (csrc/nj.cu: cooperative persistent launch over all SMs) when the input is not symmetric or
 * its bounds do not prune (degenerate inputs: everything ties, NaN). Larger matrices go to the full-scan kernel. */
PF_API int64_t pf_nj_work_bytes(int64_t n);
PF_API int pf_nj(double* d_dist, int64_t n, int64_t ld, int32_t* d_merge, double* d_len,
                 void* d_work, void* stream);
/* n_mat matrices of the same size in one call (bootstrap replicates): matrix m at d_dist + m * mat_stride
 * doubles, outputs at d_merge / d_len + m * 3 (n - 2), workspace n_mat * pf_nj_work_bytes(n) bytes. The
 * matrices are joined concurrently: by one cluster each (bounded scan, from 1,536 nodes on) or each by its own
 * group of CTAs of one cooperative launch (full scan). */
PF_API int pf_nj_batch(double* d_dist, int64_t n, int64_t ld, int n_mat, int64_t mat_stride,
                       int32_t* d_merge, double* d_len, void* d_work, void* stream);

/* Host-side: the join list of pf_nj (copied to the host) as unrooted Newick text with a trifurcating root,
 * branch lengths "%.10g" — what `treebest nj` prints into SNP_treebest.NHX (treetool/script.py:359-360).
 * names[i] = label of leaf i, already quoted where Newick needs it; h_support: NULL or one value
 * per row t < n - 3, written as the label of internal node n + t. Returns the text length (without the
 * terminating 0) or -1 on error; the text is written when h_out != NULL and capacity > length, so a first
 * call with h_out = NULL sizes the buffer. */
PF_API int64_t pf_newick_from_merges(const int32_t* h_merge, const double* h_len, int64_t n,
                                     const char* const* names, const double* h_support, char* h_out,
                                     int64_t capacity);

/* Host-side: the PHYLIP square distance matrix file that stands beside the tree (`{out}/SNP.dist`, SURVEY.md 8(f).2;
 * the reference's tree step is `treebest nj`, treetool/script.py:359-360): the line "n", then per sample its label,
 * two blanks and its n distances as "%.<decimals>f" separated by single blanks (correctly rounded, the digits Python's
 * and glibc's formatting give). h_dist: n rows of ld doubles in host memory. n_threads <= 0: all host threads. */
PF_API int pf_write_distance_matrix(const char* path, const char* const* names, int64_t n, const double* h_dist,
                                    int64_t ld, int decimals, int n_threads);

/* Host-side: bootstrap support of the main tree's internal edges - h_support[t], t < n - 3, = percentage of the n_reps
 * replicate join lists (pf_nj / pf_nj_batch outputs copied to the host, [n_reps][n - 2][3]) whose tree has the split
 * above node n + t of the main join list; the numbers the reference asks `treebest nj -b 1000` for
 * (treetool/script.py:359-360). n >= 4. n_threads <= 0: all host threads. */
PF_API int pf_bootstrap_support(const int32_t* h_main_merge, const int32_t* h_rep_merges, int64_t n, int64_t n_reps,
                                double* h_support, int n_threads);

/* Block bootstrap (csrc/bootstrap.cu; the reproducible counterpart of the `-b 1000` the reference passes to
 * `treebest nj`, treetool/script.py:359-360). The caller keeps the count matrices of n_blocks contiguous
 * site blocks (pf_pair_counts / pf_pair_counts_tc on word ranges into separate [3][n][ld] buffers,
 * elems_per_block = 3 n ld apart).
 * pf_bootstrap_weights: d_weights[r][b] (int32, n_reps x n_blocks) = multiplicity of block b among the
 * n_blocks draws with replacement of replicate replicate0 + r (counter-based hash of seed, replicate, draw).
 * pf_bootstrap_counts: d_out[r][e] = sum_b d_weights[r][b] * d_blocks[b][e] for up to 8 replicates per call;
 * elems_per_block a multiple of 4, buffers 16-byte aligned. Normalise and join each replicate with
 * pf_normalise / pf_nj. */
PF_API int pf_bootstrap_weights(uint64_t seed, int64_t replicate0, int n_reps, int n_blocks,
                                int32_t* d_weights, void* stream);
PF_API int pf_bootstrap_counts(const int32_t* d_blocks, int n_blocks, int64_t elems_per_block,
                               const int32_t* d_weights, int n_reps, int32_t* d_out, void* stream);

/* Diagnostic: phase accounting of the NJ kernels (nanoseconds of CTA 0 summed over iterations) and a
 * switch that forces the general kernel. enable: bit 0 = run the instrumented kernels in the following
 * pf_nj calls, bit 1 = use the general kernel (two grid barriers per iteration) also for matrices the
 * fast kernel would take. h_out8, fast kernel: [0] scan, [1] CTA reduction + publish, [2] exchange,
 * [3] update; general kernel: [0] scan + CTA reduction, [1] grid barrier 1, [2] winner + update,
 * [3] grid barrier 2; [4..7] counters of the bounded-scan kernel: row rescans, list evaluations, iterations, matrices
 * it joined. Further bits of enable: bit 2 = full-scan kernels only, bit 3 = bounded scan at every size it can
 * take (8 ... 11,000 nodes). Reading resets the counters. */
PF_API int pf_nj_profile(int enable, unsigned long long* h_out8);
/* Diagnostic: phase accounting of the bounded-scan NJ kernel (csrc/nj_lazy.cu), 16 counters in nanoseconds (see the
 * definition); enable != 0 switches the instrumented kernel on for the following pf_nj calls. */
PF_API int pf_nj_lazy_profile(int enable, unsigned long long* h_out16);

/* Number of kernels this library has launched in the calling process since it was loaded (every
 * <<< >>> / cudaLaunch* site of the product kernels bumps one counter). bench.py reports the difference
 * over its timed region as `gpu_launches`. */
PF_API int64_t pf_launch_count(void);

#endif /* PHYLOFORGE_B200_H */
