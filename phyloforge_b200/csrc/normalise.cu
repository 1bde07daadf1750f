// Integer count matrices -> fp64 distances: one IEEE-754 divide of
// exactly converted integers per pair, so any implementation with the right integers
// agrees to the last bit with oracle/pf_oracle.c:pfo_normalise. Elementwise, HBM-bound:
// 12 B read + 8 B written per pair.
#include "pf_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
normalise_kernel(const int32_t* __restrict__ counts, int64_t n, int64_t ld, int metric,
                 double* __restrict__ dist, int64_t ld_dist, int32_t* __restrict__ n_zero_valid) {
  const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j >= n) return;
  const int64_t plane = n * ld;
  const int32_t v = counts[i * ld + j];
  const int32_t d1 = counts[plane + i * ld + j];
  const int32_t d2 = counts[2 * plane + i * ld + j];
  double d;
  if (i == j) {
    d = 0.0;
  } else if (v == 0) {
    d = 1.0;
    if (n_zero_valid) atomicAdd(n_zero_valid, 1);
  } else if (metric == 0) {
    d = __ddiv_rn(static_cast<double>(d1), static_cast<double>(v));
  } else {
    d = __ddiv_rn(static_cast<double>(static_cast<int64_t>(d1) + d2),
                  static_cast<double>(2 * static_cast<int64_t>(v)));
  }
  dist[i * ld_dist + j] = d;
}

}  // namespace

PF_API int pf_normalise(const int32_t* d_counts, int64_t n_samples, int64_t ld, int metric,
                        double* d_dist, int64_t ld_dist, int32_t* d_n_zero_valid, void* stream) {
  PF_TRY_BEGIN
  if (!d_counts || !d_dist || n_samples <= 0 || ld < n_samples || ld_dist < n_samples)
    return pf_fail("pf_normalise: bad arguments");
  if (metric != 0 && metric != 1) return pf_fail("pf_normalise: metric must be 0 (p) or 1 (ibs)");
  if (n_samples > 65535) return pf_fail("pf_normalise: n_samples > 65535 not supported");
  dim3 grid(static_cast<unsigned>(pf_ceil_div(n_samples, 256)), static_cast<unsigned>(n_samples));
  pf_count_launch();
  normalise_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_counts, n_samples, ld, metric,
                                                                         d_dist, ld_dist, d_n_zero_valid);
  PF_CUDA(cudaGetLastError());
  return PF_OK;
  PF_TRY_END
}
