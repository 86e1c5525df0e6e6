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

// Programmatic dependent launch: a kernel launched this way may start (and run its prologue: barrier init, tensor-
// memory allocation, constant operands to TMEM / shared memory) while its predecessor in the stream is still draining;
// it must execute pdl_wait() before it touches anything the predecessor wrote.  Captured as a programmatic edge in
// CUDA graphs.  TIK_NO_PDL=1 falls back to ordinary stream order.
bool pdl_enabled();
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// Per-launch options the plan sets before each tensor-core launch (thread-local; read when the launch fills its
// parameters).  rev: walk the tiles from the last to the first -- consecutive kernels alternate, so each starts on the
// part of its input the previous kernel wrote last, which is still in L2.  l2: read-once activation loads carry an
// evict-first hint, so that they do not push the freshly written output out of L2.
struct LaunchOpts { int rev; int l2; };
LaunchOpts& launch_opts();

// fp32 implicit-GEMM arguments (stgcn_simt.cu rowgemm_f32_kernel, rowgemm_tf32.cu): TikRowGemm with the K offsets resolved
struct F32Slab { const float* a; int c, t_in, t_mul, t_off, koff; };
struct F32Args {
  F32Slab slabs[TIK_MAX_SLABS];
  int n_slabs;
  const float* w; int ktot;
  const float* bias; int bias_per_node;
  int64_t rows; int v, t_out, c_out, c_out_valid;
  int act; float slope;
  int res_kind; const void* res; const float* res_w; int res_cin, res_t_mul, res_t_in;
  float* out; int out_layout;
  // n / t_out and n / v for n < 2^31 as (n * magic) >> shift (rowgemm_tf32.cu: ~100 integer divisions per tile row otherwise)
  unsigned long long div_t_magic, div_v_magic; int div_t_shift, div_v_shift;
  // optional: the weights already split into their TF32 parts (tf32_split_weights; the plan does it once per weight
  // version in its workspace), both (c_out, ktot) like w.  nullptr: the kernel splits W on the way, like the activations.
  const float* w_big; const float* w_small;
};
bool rowgemm_tf32_supported(const F32Args& a);
int rowgemm_tf32_launch(const F32Args& a, cudaStream_t s);

// internal entry points shared between translation units
int rowgemm_f32(const TikRowGemm* d, cudaStream_t s);
int rowgemm_f32_presplit(const TikRowGemm* d, const float* w_big, const float* w_small, cudaStream_t s);
int tf32_split_weights(const float* w, float* big, float* small, int64_t n, cudaStream_t s);
int rowgemm_bf16(const TikRowGemm* d, cudaStream_t s);
int launch_batch_rodrigues(const float* aa, float* R9, int64_t M, cudaStream_t s);
int stem_gcn_impl(int dtype, const float* x, const float* in_scale, const float* in_shift, const float* agg,
                  const float* w, const float* bias, void* out, const float* res_w, void* res_out,
                  int res_stride, int64_t N, int T, int V, int Cin, int K, int Cout, int relu,
                  const TikWindowing* winp, int64_t win_n0, cudaStream_t s);

}  // namespace tik
