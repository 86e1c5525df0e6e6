// Prepared (tensor maps encoded once) launches of the tcgen05 implicit GEMM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tik.h"

namespace tik {
struct UmmaPrepared;
// Encodes the TMA tensor maps for `d` with room for nv_capacity row groups (host-only work).
int umma_prepare(const TikRowGemm* d, int64_t nv_capacity, UmmaPrepared** out);
// Launch with the current nv / bias / residual / output fields of `d` (slab tensors are those baked at prepare time).
int umma_launch(UmmaPrepared* u, const TikRowGemm* d, cudaStream_t s);
void umma_free(UmmaPrepared* u);
struct GcnFusedPrepared;
bool gcn_fused_supported(int cin, int cout, int V, int K);
int gcn_fused_prepare(const void* x, const void* abd, const void* w, const float* bias, void* out, int64_t n_clips, int T, int V,
                      int cin, int cout, int relu, GcnFusedPrepared** outp);
int gcn_fused_launch(GcnFusedPrepared* g, int64_t n_clips, cudaStream_t s);
void gcn_fused_free(GcnFusedPrepared* g);
}  // namespace tik
