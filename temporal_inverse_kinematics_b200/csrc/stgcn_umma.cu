// bf16 implicit-GEMM on Blackwell tensor cores (tcgen05.mma + TMEM accumulators, TMA-fed).
//
//   out[row, :] = act( sum_slabs A_s[row_s, :] . W[:, slab]^T + bias + residual )
//
// A "row" is (row group nv = clip*V + node, frame t).  Because activations are node-major
// (NV, T, C), the rows a 128-row MMA tile needs from one slab form a 3-D box
// (64 channels, TT frames, VV row groups) of the source tensor, so every A operand -- including the
// time-shifted taps of the (kt x 1) temporal convolution, its stride-2 sub-sampling and its zero
// padding at clip ends -- is ONE cp.async.bulk.tensor load: the shift is the box start coordinate,
// the stride is the tensor map's elementStrides and the padding is TMA out-of-bounds zero fill.
// The box lands in shared memory in the 128-byte-swizzled K-major layout that the UMMA shared-memory
// descriptor consumes directly.  Weights (c_out, K_total) arrive the same way as the B operand.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM -> registers -> bias / residual / activation -> global); warp 2 also
// owns the TMEM allocation.  One output tile per CTA; 2-3 CTAs are co-resident per SM so that one
// CTA's epilogue overlaps another's main loop.
//
// Serves: BN-folded 1x1 channel GEMM of ConvTemporalGraphical (gconv_origin.py:59) after the adjacency
// aggregation, the temporal convolution + residual + BN + ReLU of StGcnBlock (st_gcn_aaai18.py:177-214)
// and both Linear layers of the head (pose_trainer.py:89-92).
#include <cuda.h>

#include "tik_common.cuh"
#include "umma_prepared.h"

namespace tik {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a descriptor / byte-count mistake must surface as a trap, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 24)) {
      printf("tik: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols));
}

// D[tmem] (+)= A[smem] . B[smem]^T, bf16 inputs, fp32 accumulate, M=128, K=16
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows at 128 B pitch,
// 8-row swizzle atoms 1024 B apart (SBO), LBO unused for swizzled K-major (encoded 1), version 1.
__device__ __forceinline__ uint64_t make_smem_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (kind::f16): D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ kernel
constexpr int kTileM = 128;
constexpr int kChunkK = 64;                       // bf16 elements per K chunk = one 128 B swizzle row
constexpr int kABytes = kTileM * kChunkK * 2;     // 16 KB
constexpr int kUmmaThreads = 192;

struct UmmaParams {
  CUtensorMap map_a[TIK_MAX_SLABS];
  CUtensorMap map_w;
  int32_t n_slabs;
  int32_t chunks[TIK_MAX_SLABS];   // c_s / 64
  int32_t t_mul[TIK_MAX_SLABS];
  int32_t t_off[TIK_MAX_SLABS];
  int32_t tt, vv, tiles_t;         // tile = tt frames x vv row groups
  int32_t a_box_bytes;             // tt*vv*128
  int64_t nv;
  int32_t v, t_out, c_out, c_out_valid;
  const float* bias; int32_t bias_per_node;
  int32_t act; float slope;
  int32_t res_kind; const void* res; const float* res_w; int32_t res_cin, res_t_mul, res_t_in;
  void* out; int32_t out_layout;
};

template <int BN, int STAGES>
struct UmmaSmem {
  static constexpr int kBBytes = BN * kChunkK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;   // barriers + alignment slack
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == TIK_ACT_RELU) return fmaxf(v, 0.f);
  if (act == TIK_ACT_LEAKY) return v > 0.f ? v : v * slope;
  return v;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kUmmaThreads) rowgemm_umma_kernel(const __grid_constant__ UmmaParams p) {
  using SM = UmmaSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_t = blockIdx.x % p.tiles_t;
  const int64_t tile_nv = blockIdx.x / p.tiles_t;
  const int t0 = tile_t * p.tt;
  const int64_t nv0 = tile_nv * p.vv;
  const int n0 = blockIdx.y * BN;

  int total_chunks = 0;
  for (int s = 0; s < p.n_slabs; ++s) total_chunks += p.chunks[s];

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_slabs; ++s) tma_prefetch_desc(&p.map_a[s]);
    tma_prefetch_desc(&p.map_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; int kw = 0;
      for (int s = 0; s < p.n_slabs; ++s) {
        const int ts = t0 * p.t_mul[s] + p.t_off[s];
        for (int c = 0; c < p.chunks[s]; ++c, ++kw) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * SM::kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(&full_bar[stage], (uint32_t)(p.a_box_bytes + SM::kBBytes));
          tma_load_3d(sa, &p.map_a[s], &full_bar[stage], c * kChunkK, ts, (int)nv0);
          tma_load_2d(sb, &p.map_w, &full_bar[stage], kw * kChunkK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kTileM, BN);
      int stage = 0; uint32_t phase = 0;
      for (int kc = 0; kc < total_chunks; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * SM::kStageBytes);
        const uint64_t da = make_smem_desc_kmajor_sw128(sa);
        const uint64_t db = make_smem_desc_kmajor_sw128(sa + kABytes);
#pragma unroll
        for (int k = 0; k < kChunkK / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the swizzled row: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);          // frees this smem stage once the MMAs have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);                // accumulator complete
    }
  } else {
    // ===================== epilogue =====================
    const int lane_grp = warp & 3;               // TMEM lanes [32*lane_grp, +32) are accessible to this warp
    const int r = lane_grp * 32 + lane;          // tile row == TMEM lane
    const int nv_l = r / p.tt, t_l = r - nv_l * p.tt;
    const int64_t nv = nv0 + nv_l;
    const int t = t0 + t_l;
    const bool valid = (nv_l < p.vv) && (nv < p.nv) && (t < p.t_out);
    const int node = (int)(nv % p.v);
    const int64_t n = nv / p.v;
    const int64_t row = nv * p.t_out + t;

    const float* bias = p.bias + (p.bias_per_node ? (int64_t)node * p.c_out : 0);
    const __nv_bfloat16* res_row = nullptr;
    float xin[8];
    const float* rw = nullptr;
    if (valid && p.res_kind == TIK_RES_IDENTITY) {
      res_row = reinterpret_cast<const __nv_bfloat16*>(p.res) + row * p.c_out;
    } else if (valid && p.res_kind == TIK_RES_STEM) {
      const float* xp = reinterpret_cast<const float*>(p.res) +
                        ((n * p.res_t_in + (int64_t)t * p.res_t_mul) * p.v + node) * p.res_cin;
      for (int ci = 0; ci < p.res_cin; ++ci) xin[ci] = __ldg(xp + ci);
      rw = p.res_w + (int64_t)node * p.c_out * p.res_cin;
    }
    int64_t out_off;
    int ld;
    if (p.out_layout == TIK_OUT_NODE_MAJOR) { out_off = row * p.c_out; ld = p.c_out; }
    else if (p.out_layout == TIK_OUT_TIME_MAJOR) { out_off = ((n * p.t_out + t) * p.v + node) * (int64_t)p.c_out; ld = p.c_out; }
    else { out_off = row * p.c_out_valid; ld = p.c_out_valid; }

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0, acc);
      tmem_ld_wait();
      if (!valid) continue;
      const int cg = n0 + c0;                    // first global output channel of this chunk
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) + __ldg(bias + cg + j);
      if (res_row != nullptr) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u = __ldg(reinterpret_cast<const uint4*>(res_row + cg) + q);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 f = __bfloat1622float2(h[e]);
            v[q * 8 + 2 * e] += f.x;
            v[q * 8 + 2 * e + 1] += f.y;
          }
        }
      } else if (rw != nullptr) {
        for (int j = 0; j < 32; ++j) {
          const float* wj = rw + (int64_t)(cg + j) * p.res_cin;
          float a = v[j];
          for (int ci = 0; ci < p.res_cin; ++ci) a = fmaf(__ldg(wj + ci), xin[ci], a);
          v[j] = a;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act, p.slope);
      if (p.out_layout == TIK_OUT_ROWS_F32) {
        float* o = reinterpret_cast<float*>(p.out) + out_off;
        for (int j = 0; j < 32; ++j)
          if (cg + j < ld) o[cg + j] = v[j];
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_off + cg;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
          reinterpret_cast<uint4*>(o)[q] = u;
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const uint32_t* estr) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return TIK_ERR_CUDA;
  TIK_CHECK_ARG(((uintptr_t)base & 15) == 0, "TMA base address must be 16-byte aligned");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              box[0], box[1], rank > 2 ? box[2] : 0);
    return TIK_ERR_CUDA;
  }
  return TIK_OK;
}

struct UmmaPrepared {
  UmmaParams p;
  int bn, stages;
  int64_t nv_capacity;
};

template <int BN, int STAGES>
static int launch_variant(const UmmaParams& p, dim3 grid, cudaStream_t s) {
  using SM = UmmaSmem<BN, STAGES>;
  static bool attr_done[64] = {};
  int dev = 0;
  TIK_CUDA(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {   // the attribute is per device
    TIK_CUDA(cudaFuncSetAttribute(rowgemm_umma_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
    attr_done[dev & 63] = true;
  }
  rowgemm_umma_kernel<BN, STAGES><<<grid, kUmmaThreads, SM::kTotal, s>>>(p);
  TIK_LAUNCH_CHECK();
  return TIK_OK;
}

int umma_prepare(const TikRowGemm* d, int64_t nv_capacity, UmmaPrepared** out) {
  TIK_CHECK_ARG(d->c_out % 64 == 0, "bf16 path: c_out=%d must be a multiple of 64", d->c_out);
  TIK_CHECK_ARG(nv_capacity >= d->nv && nv_capacity > 0, "nv capacity");
  UmmaPrepared* u = new UmmaPrepared();
  UmmaParams& p = u->p;
  memset(&p, 0, sizeof(p));
  int ktot = 0, max_mul = 1;
  for (int i = 0; i < d->n_slabs; ++i) {
    const TikSlab& sl = d->slabs[i];
    if (!(sl.a_dev && sl.c > 0 && sl.c % 64 == 0 && sl.t_in > 0 && sl.t_mul >= 1 && sl.t_mul <= 8)) {
      delete u;
      set_error("bf16 path: slab %d malformed (c=%d must be a multiple of 64, t_mul=%d in [1,8])", i, sl.c, sl.t_mul);
      return TIK_ERR_INVALID;
    }
    ktot += sl.c;
    if (sl.t_mul > max_mul) max_mul = sl.t_mul;
  }
  int tt = d->t_out < kTileM ? d->t_out : kTileM;
  if (tt * max_mul > 256) tt = 256 / max_mul;     // TMA box extent limit
  int vv = kTileM / tt;
  p.tt = tt; p.vv = vv; p.tiles_t = (int)ceil_div(d->t_out, tt);
  p.a_box_bytes = tt * vv * kChunkK * 2;
  p.n_slabs = d->n_slabs;
  for (int i = 0; i < d->n_slabs; ++i) {
    const TikSlab& sl = d->slabs[i];
    p.chunks[i] = sl.c / 64; p.t_mul[i] = sl.t_mul; p.t_off[i] = sl.t_off;
    uint64_t dims[3] = {(uint64_t)sl.c, (uint64_t)sl.t_in, (uint64_t)nv_capacity};
    uint64_t strides[2] = {(uint64_t)sl.c * 2, (uint64_t)sl.c * 2 * (uint64_t)sl.t_in};
    uint32_t box[3] = {(uint32_t)kChunkK, (uint32_t)(tt * sl.t_mul), (uint32_t)vv};
    uint32_t estr[3] = {1, (uint32_t)sl.t_mul, 1};
    int rc = encode_map(&p.map_a[i], sl.a_dev, 3, dims, strides, box, estr);
    if (rc != TIK_OK) { delete u; return rc; }
  }
  u->bn = d->c_out % 256 == 0 ? 256 : (d->c_out % 128 == 0 ? 128 : 64);
  u->stages = u->bn == 64 ? 4 : (u->bn == 128 ? 3 : (ktot >= 1024 ? 4 : 2));
  {
    uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)d->c_out};
    uint64_t strides[1] = {(uint64_t)ktot * 2};
    uint32_t box[2] = {(uint32_t)kChunkK, (uint32_t)u->bn};
    uint32_t estr[2] = {1, 1};
    int rc = encode_map(&p.map_w, d->w_dev, 2, dims, strides, box, estr);
    if (rc != TIK_OK) { delete u; return rc; }
  }
  u->nv_capacity = nv_capacity;
  *out = u;
  return TIK_OK;
}

int umma_launch(UmmaPrepared* u, const TikRowGemm* d, cudaStream_t s) {
  UmmaParams& p = u->p;
  TIK_CHECK_ARG(d->nv <= u->nv_capacity, "nv exceeds prepared capacity");
  p.nv = d->nv; p.v = d->v; p.t_out = d->t_out; p.c_out = d->c_out; p.c_out_valid = d->c_out_valid;
  p.bias = d->bias_dev; p.bias_per_node = d->bias_per_node;
  p.act = d->act; p.slope = d->slope;
  p.res_kind = d->res_kind; p.res = d->res_dev; p.res_w = d->res_w_dev;
  p.res_cin = d->res_cin; p.res_t_mul = d->res_t_mul; p.res_t_in = d->res_t_in;
  p.out = d->out_dev; p.out_layout = d->out_layout;
  if (d->nv == 0) return TIK_OK;
  int64_t tiles = ceil_div(d->nv, p.vv) * p.tiles_t;
  TIK_CHECK_ARG(tiles < (1ll << 31), "too many tiles");
  dim3 grid((unsigned)tiles, (unsigned)(d->c_out / u->bn));
  if (u->bn == 64) return launch_variant<64, 4>(p, grid, s);
  if (u->bn == 128) return launch_variant<128, 3>(p, grid, s);
  if (u->stages == 4) return launch_variant<256, 4>(p, grid, s);
  return launch_variant<256, 2>(p, grid, s);
}

void umma_free(UmmaPrepared* u) { delete u; }

int rowgemm_bf16(const TikRowGemm* d, cudaStream_t s) {
  UmmaPrepared* u = nullptr;
  int rc = umma_prepare(d, d->nv, &u);
  if (rc != TIK_OK) return rc;
  rc = umma_launch(u, d, s);
  umma_free(u);
  return rc;
}

}  // namespace tik
