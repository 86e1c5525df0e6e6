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
}  // namespace tik
