// Shared helpers for libtik.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/tik.h"

namespace tik {

void set_error(const char* fmt, ...);

#define TIK_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::tik::set_error(__VA_ARGS__);             \
      return TIK_ERR_INVALID;                    \
    }                                            \
  } while (0)

#define TIK_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      ::tik::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return TIK_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// launch check without host synchronisation (SURVEY.md section 8b "Error conventions")
#define TIK_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t e_ = cudaPeekAtLastError();                                              \
    if (e_ != cudaSuccess) {                                                             \
      ::tik::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
      (void)cudaGetLastError();                                                          \
      return TIK_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// internal entry points shared between translation units
int rowgemm_f32(const TikRowGemm* d, cudaStream_t s);
int rowgemm_bf16(const TikRowGemm* d, cudaStream_t s);
int launch_batch_rodrigues(const float* aa, float* R9, int64_t M, cudaStream_t s);
int stem_gcn_impl(int dtype, const float* x, const float* in_scale, const float* in_shift, const float* agg,
                  const float* w, const float* bias, void* out, const float* res_w, void* res_out,
                  int res_stride, int64_t N, int T, int V, int Cin, int K, int Cout, int relu,
                  const TikWindowing* winp, int64_t win_n0, cudaStream_t s);

}  // namespace tik
