/* libphyloforge_b200_probes.so — microbenchmarks, latency probes and instruction self-tests.
 *
 * NOT part of the drop-in boundary: nothing in phyloforge_b200/ (the product) calls these. They measure
 * the roofline denominators bench.py quotes (integer-pipe rates, the sustained tcgen05 kind::mxf4 rate)
 * and check the tcgen05 / TMEM plumbing against integer dot products (tests/test_gpu_umma.py,
 * tools/probe_*.py). Same conventions as include/phyloforge_b200.h: 0 = ok, pf_last_error() of THIS
 * library gives the message, d_* are device pointers, `stream` is a cudaStream_t.
 */
#ifndef PHYLOFORGE_B200_PROBES_H
#define PHYLOFORGE_B200_PROBES_H

#include <stdint.h>

#ifdef __cplusplus
#define PF_API extern "C" __attribute__((visibility("default")))
#else
#define PF_API extern
#endif

PF_API const char* pf_last_error(void);

/* Integer-pipe microbenchmark (roofline denominator of pf_pair_counts).
 * kind: 0 POPC, 1 LOP3, 2 IADD, 3 IMAD, 4 POPC:LOP3 1:4, 5 POPC:IMAD 1:4, 6 LOP3:IMAD 1:1,
 *       7 plain pair-word body with missing data, 8 plain pair-word body without.
 * Synchronous. Outputs thread-level instructions per second and per clock per SM. */
PF_API int pf_microbench_int_pipe(int kind, int iters, double* ops_per_s,
                                  double* ops_per_clk_per_sm, void* stream);

/* Probe of the legacy warp-level tensor paths: kind 0 mma.sync b1
 * and.popc (m16n8k256), 1 b1 xor.popc, 2 s8 (m16n8k32). Synchronous; multiply-accumulates/s. */
PF_API int pf_microbench_mma(int kind, int iters, double* macs_per_s, void* stream);

/* Self-test of the tcgen05 (kind::i8, TMEM) plumbing used by the tensor-core pair-count variant:
 * d_c[128][256] (int32) = d_a[128][k] * d_b[256][k]^T for int8 operands, k a multiple of 32, <= 256.
 * a_in_tmem != 0 feeds A through tensor memory (tcgen05.st + the TS form of the MMA). */
PF_API int pf_selftest_umma(const int8_t* d_a, const int8_t* d_b, int32_t* d_c, int k, int a_in_tmem,
                            void* stream);

/* Back-to-back kind::i8 MMA issue rate (128 x n x 32 per MMA) with operands resident in shared
 * memory / TMEM: clock64 ticks per MMA, averaged over one CTA per SM. group > 1 (TMEM mode): that many
 * iterations target one accumulator before switching to the other. Synchronous. */
PF_API int pf_probe_umma_rate(int n, int reps, int a_in_tmem, int group, double* ticks_per_mma, void* stream);

/* Diagnostic: latency of the building blocks of the NJ iteration, nanoseconds per step.
 * kind 0 dependent L2 load chain, 1 store + gpu-scope fence, 2 CTA barrier (512 threads), 3 tagged-word
 * exchange between n_ctas CTAs, 4 cooperative grid barrier over n_ctas CTAs, 5 as 0 on lines last
 * written by another SM, 6 a batch of 16 independent L2 loads per step, 7 tcgen05.commit to mbarrier wait,
 * 8 one 6 KB cp.async.bulk from issue to completion, 9 the same with 8 copies in flight (per copy),
 * 10 mbarrier arrive to try_wait wake-up between two warps (per hop). */
PF_API int pf_probe_latency(int kind, int iters, int n_ctas, double* h_ns_per_step, void* stream);

/* Diagnostics for the 4-bit tensor-core path (tcgen05.mma kind::mxf4, E2M1 operands, K = 64 per
 * instruction, all block scale factors 1.0; csrc/umma_fp4_probe.cu).
 * pf_selftest_umma_fp4: d_c[128][n] (fp32) = d_a[128][k] * d_b[n][k]^T for int8 entries from
 * {0, +-1, +-2, +-3, +-4, +-6}; n = 128 or 192, k a multiple of 64, at most 1024; a_in_tmem != 0
 * feeds the A operand from tensor memory (tcgen05.st) instead of shared memory.
 * pf_probe_umma_fp4_rate: SM clock64 ticks per 128 x n x 64 instruction, operands resident in shared
 * memory (a_in_tmem: the A operand in tensor memory), n_acc accumulators used round-robin, the B
 * operand rotating over n_bufs different tiles; noise > 0: three other warps each do one LDS.128 +
 * one STS.128 (512 B each way) every `noise` clocks beside the MMAs. */
PF_API int pf_selftest_umma_fp4(const int8_t* d_a, const int8_t* d_b, float* d_c, int n, int k,
                                int a_in_tmem, void* stream);
PF_API int pf_probe_umma_fp4_rate(int n, int reps, int n_acc, int a_in_tmem, int n_bufs, int noise,
                                  double* ticks_per_mma, void* stream);

/* Diagnostics for the CTA-pair form of that path (tcgen05.mma.cta_group::2 on a cluster of two CTAs;
 * csrc/umma_fp4_2cta_probe.cu). pf_selftest_umma_fp4_2cta: d_c[256][n] (fp32) = d_a[256][k] * d_b[n][k]^T,
 * n a multiple of 16 in [32, 256], k a multiple of 64, at most 1024. pf_probe_umma_fp4_2cta_rate: SM
 * clock64 ticks per 256 x n x 64 instruction (n <= 240, two accumulators), on every CTA pair of the
 * device at once; *h_pairs = number of pairs. */
PF_API int pf_selftest_umma_fp4_2cta(const int8_t* d_a, const int8_t* d_b, float* d_c, int n, int k, void* stream);
PF_API int pf_probe_umma_fp4_2cta_rate(int n, int reps, int n_bufs, double* ticks_per_mma, int* h_pairs,
                                       void* stream);

/* Sustained rate of the CTA-pair instruction on the whole device, timed with CUDA events around the launch
 * (not SM clocks): every CTA pair issues 2 * reps back-to-back 256 x n x 64 kind::mxf4 instructions on operands
 * resident in shared memory. *h_tops = 2 * 256 * n * 64 * 2 * reps * pairs / elapsed / 1e12 (TOP/s, fp4 dense),
 * *h_ms = elapsed. Choose reps so that the launch lasts long enough for the board's power limit to act
 * (bench.py: ~0.5 s). This is the measured denominator next to "4 x bf16" in bench.py's roofline. */
PF_API int pf_probe_umma_fp4_2cta_peak(int n, int reps, int n_bufs, double* h_tops, double* h_ms, void* stream);

#endif /* PHYLOFORGE_B200_PROBES_H */
