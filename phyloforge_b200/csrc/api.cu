// Error plumbing and device queries of the C ABI (include/phyloforge_b200.h).
#include "pf_common.cuh"

#include <atomic>
#include <cstring>

namespace {
thread_local char g_err[1024] = "";

int vfail(int code, const char* fmt, va_list ap) {
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  return code;
}
std::atomic<long long> g_launches{0};
}  // namespace

void pf_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int pf_fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  int rc = vfail(PF_ERR_ARG, fmt, ap);
  va_end(ap);
  return rc;
}

int pf_fail_code(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  int rc = vfail(code, fmt, ap);
  va_end(ap);
  return rc;
}

PF_API const char* pf_last_error(void) { return g_err; }

#ifndef PF_PROBES_LIB
PF_API int64_t pf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

PF_API const char* pf_version(void) { return "phyloforge_b200 0.1.0 (sm_100a)"; }

PF_API int pf_tile(void) { return PF_TILE; }

PF_API int pf_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem,
                              int* l2_bytes) {
  PF_TRY_BEGIN
  int dev = 0;
  PF_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  PF_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (total_mem) *total_mem = static_cast<int64_t>(p.totalGlobalMem);
  if (l2_bytes) *l2_bytes = p.l2CacheSize;
  return PF_OK;
  PF_TRY_END
}
#endif  // PF_PROBES_LIB
