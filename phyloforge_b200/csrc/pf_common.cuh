// Shared helpers for libphyloforge_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "phyloforge_b200.h"

#include <cstdarg>
#include <cstdio>
#include <exception>
#include <vector>

// ---- layout constants shared by every kernel (DESIGN.md "Data layout in HBM") ----
// Sites are grouped into 32-site words. Samples are grouped into tiles of PF_TILE.
// For sample tile T, word w, plane p (0 = L = low code bit, 1 = H = high code bit)
// the 32-bit word of sample i lives at
//     planes[((T * n_words + w) * 2 + p) * PF_TILE + (i % PF_TILE)]
// Genotype codes: 0 hom-REF, 1 HET, 2 hom-ALT, 3 missing. Padding (samples beyond
// n_samples in the last tile, sites beyond n_sites in the last word) is code 3.
#define PF_WORD_SITES 32  // PF_TILE (64) comes from phyloforge_b200.h

// status codes of the C ABI
#define PF_OK 0
#define PF_ERR_ARG 1
#define PF_ERR_CUDA 2
#define PF_ERR_INPUT 3
#define PF_ERR_INTERNAL 4

void pf_count_launch(int n = 1);                  // bumps the counter behind pf_launch_count()
int pf_fail(const char* fmt, ...);                 // records message, returns PF_ERR_ARG
int pf_fail_code(int code, const char* fmt, ...);  // records message, returns code

#define PF_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t pf_e_ = (expr);                                                         \
    if (pf_e_ != cudaSuccess)                                                           \
      return pf_fail_code(PF_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,    \
                          cudaGetErrorString(pf_e_));                                   \
  } while (0)

#define PF_TRY_BEGIN try {
#define PF_TRY_END                                                              \
  }                                                                             \
  catch (const std::exception& e) {                                             \
    return pf_fail_code(PF_ERR_INTERNAL, "exception: %s", e.what());            \
  }                                                                             \
  catch (...) {                                                                 \
    return pf_fail_code(PF_ERR_INTERNAL, "unknown exception");                  \
  }

static inline __host__ __device__ int64_t pf_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
