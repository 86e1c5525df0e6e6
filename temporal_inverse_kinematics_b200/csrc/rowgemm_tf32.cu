// fp32 implicit GEMM on the tensor pipe: 3xTF32 (tcgen05.mma kind::tf32), the 1e-4 parity path of
//   out[row, :] = act( sum_slabs A_s[row_s, :] . W[:, koff_s : koff_s + c_s]^T + bias + residual )
// i.e. the 1x1 channel GEMM of ConvTemporalGraphical (gconv_origin.py:59), the (kt x 1) temporal convolution with its
// residual branch (st_gcn_aaai18.py:177-214) and the head Linear layers (pose_trainer.py:89-92) with fp32 activations.
//
// A TF32 operand keeps 10 mantissa bits, so every fp32 value is split in the kernel into big = rna_tf32(v) and
// small = v - big (exact in fp32; the tensor core reads its top 19 bits), and each product is three MMAs,
//   A.W ~= A_small.W_big + A_big.W_small + A_big.W_big         (the dropped small.small term is ~2^-22 relative),
// ~21 bits per product instead of TF32's 10.  The tensor core's fp32 accumulation TRUNCATES (round toward zero): one
// long accumulation chain shrinks the result by ~0.5 ulp per MMA -- measured (tools/tf32_check.py, ONE accumulator for
// all of K): max relative error 7e-7 at K = 64, 2.2e-6 at 256, 8.5e-6 at 1024, 3.2e-5 at 4352, i.e. linear in K, and
// 2.4e-4 max-abs on the network's 'poses'.  So K is accumulated in tensor memory one 32-wide chunk at a time (12 MMAs,
// alternating between two accumulators) and the chunk sums are added in registers (round to nearest) while the next
// chunk's MMAs run: single-GEMM error <= 1.1e-6 at every K above (the SIMT fp32 kernel: 3.6e-7 ... 2.8e-6), network
// parity 1.07e-5 / 1.14e-5 / 1.0e-5 on poses / rotation matrices / joints (SIMT: 7.6e-6) -- 9x inside the 1e-4 bar.
// Speed (B200, 1.97 GHz): 109-114 TFLOP/s on the K >= 1024 layers (SIMT: 38), 60 at K = 256, 15 at K = 64 (per-tile
// prologue / epilogue not overlapped); configs[1] GEMM time 4.41 -> 1.83 ms, step 4.62 -> 2.05 ms (3.54 -> 8.0 M
// frames/s).  TIK_NO_TF32=1 restores the SIMT kernel.
//
// Persistent CTAs (one per resident slot) walk 128-row x BN-column tiles; 512 threads.  Activations and weights stay fp32 in HBM: per 32-wide K chunk
// every thread loads its float4 pieces with plain coalesced loads (row = (row group, frame) with the tap shift / stride
// / zero padding of the slab resolved per row), splits them, and writes big and small parts into two K-major
// 128B-swizzled shared-memory tiles (chunk j of row r at 16-byte slot j ^ (r & 7): the canonical layout the MMA
// descriptors expect, here written by hand instead of by TMA because the split has to happen on the way).  One
// thread issues the 12 MMAs of the chunk (4 K-steps of 8 x 3 products) and commits them to the stage's mbarrier; two
// stages, so the split / stores of chunk i+1 overlap the MMAs of chunk i, and the global loads run two chunks ahead in
// registers.  Epilogue: the register sums go through a padded shared-memory tile (aliasing the dead stages) so that a
// warp stores whole rows (bias / residual loads of four rows in flight at a time) -> fp32 rows, any output layout.
// The next tile's row bookkeeping and its first two chunks' loads are issued before the current tile's epilogue
// (K = 64 ... 256 layers: -8 %; the K >= 384 layers pay ~4 % for the longer loop bookkeeping: net +1 % on configs[1]).
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "tik_common.cuh"
#include "umma_ptx.cuh"

namespace tik {

constexpr int kTfKc = 32;                  // fp32 elements per K chunk = one 128-byte swizzle row
constexpr int kTfATile = 128 * 128;        // 128 rows x 128 B
// K chunks accumulated in tensor memory before the sum moves to registers.  Whole-network parity at B=256, T=64 (max-abs on
// 'poses' vs the reference, A/B on one box): 4 chunks 3.1e-5, 2 chunks 1.8e-5, 1 chunk 1.07e-5 (SIMT fp32 kernel: 7.6e-6),
// GEMM time 3.08 / 3.09 / 3.12 ms -- the truncation bias is linear in the chain length and the drain is almost free.
#if defined(TFBLOCK4)
constexpr int kTfBlock = 4;
#elif defined(TFBLOCK2)
constexpr int kTfBlock = 2;
#else
constexpr int kTfBlock = 1;
#endif

// Instruction descriptor (kind::tf32): D = f32 (bits 4-5 = 1), A = B = TF32 (format 2 at bits 7-9 and 10-12), K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] . B[smem]^T, tf32 inputs (32-bit containers), fp32 accumulate, M = 128, K = 8
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// warp-collective form (umma_ptx.cuh: the whole warp executes it with uniform operands, one elected lane issues)
__device__ __forceinline__ void umma_tf32_w(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// big = v rounded to TF32 (nearest, ties away: add half an ulp of the 13 dropped bits to the magnitude and mask -- what
// cvt.rna.tf32.f32 does, minus its Inf / NaN test: 2 instructions instead of 4); small = v - big is exact in fp32 and goes to
// the tensor core as it is (the hardware drops its low 13 bits: rounding it here as well changed the network's parity
// from 3.19e-5 to 3.15e-5 at 4-chunk accumulation -- not worth 4 more instructions per element on an issue-bound fill).
__device__ __forceinline__ float tf32_big(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void split_store(uint8_t* big, uint8_t* small, uint32_t off, float4 v) {
  float4 b, s;
  b.x = tf32_big(v.x); b.y = tf32_big(v.y); b.z = tf32_big(v.z); b.w = tf32_big(v.w);
  s.x = v.x - b.x; s.y = v.y - b.y; s.z = v.z - b.z; s.w = v.w - b.w;
  *reinterpret_cast<float4*>(big + off) = b;
  *reinterpret_cast<float4*>(small + off) = s;
}

// Tried and dropped: a 17th warp that only issues the MMAs (fill warps hand stages over through mbarriers, no bar.sync per
// chunk) -- 544 threads cap the kernel at 96 registers, the two-deep load buffer spills (800 B) and the K = 4352 layer
// drops from 114 to 84 TFLOP/s.
// BN = 128: 512 threads, one CTA per SM.  BN = 64 (the 64-channel layers: 2-6 chunks per tile, where the per-tile prologue
// and epilogue weigh most): 256 threads and half the shared memory, two CTAs per SM overlap each other's fixed costs.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int fast_div(int n, unsigned long long magic, int shift) {
  return (int)(((unsigned long long)(unsigned)n * magic) >> shift);
}

template <int BN, int kTfThreads, bool kPre>
__global__ void __launch_bounds__(kTfThreads, BN == 64 ? 2 : 1) rowgemm_tf32_kernel(const __grid_constant__ F32Args p) {
  constexpr int WT = BN * 128;                       // bytes of one weight tile (BN rows x 128 B)
  constexpr int STAGE = 2 * kTfATile + 2 * WT;       // [A_big | A_small | W_big | W_small]
  // kPre (BN = 64 only; weights pre-split by the plan): W_big / W_small of a chunk go global -> shared memory with cp.async,
  // no registers and no split arithmetic: -10 % on the 64-column layers, where two CTAs per SM cover the copy latency.  At
  // BN = 128 (one CTA per SM) the same scheme measured SLOWER than splitting in the kernel -- copy started when the stage is
  // free: +15..40 %; with a third weight buffer and the copy one chunk ahead: still +7..10 % -- so those layers keep the
  // register path (profiles/r2_notes.md).
  static_assert(!kPre || BN == 64, "pre-split weights are used by the 64-column kernel only");
  constexpr int NWB = 2;                             // weight buffers = the stages' own W slots
  constexpr int WDIST = 0;                           // chunks the weight copy runs ahead
  constexpr int OPER_BYTES = 2 * STAGE + (NWB - 2) * 2 * WT;
  constexpr int NQA = 128 * 8 / kTfThreads;          // float4 pieces of A per thread and chunk
  constexpr int NQW = BN * 8 / kTfThreads;           // float4 pieces of W per thread and chunk
  constexpr int NACC = 128 * BN / kTfThreads;        // fp32 accumulators per thread (one tile row x a quarter of the columns)
  constexpr int OUT_PITCH = BN * 4 + 16;             // staging row pitch (bytes): float4 writes of 8 consecutive rows hit 32 different banks
  static_assert(128 * OUT_PITCH <= OPER_BYTES, "the output staging tile aliases the operand buffers");
  static_assert(NACC == 16 || NACC == 32, "drain uses one tcgen05.ld of 16 or 32 columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem + OPER_BYTES);   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows = (int)p.rows;
  const int tiles_n = (p.c_out + BN - 1) / BN, total_tiles = ((rows + 127) / 128) * tiles_n;
  int row0 = 0, col0 = 0;                            // tile the load cursor works on (one tile ahead of the epilogue)

  if (tid == 0) { mbar_init(&mma_done[0], 1); mbar_init(&mma_done[1], 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  // this thread's pieces: A piece q is 16-byte slot tid & 7 of tile row (q * 512 + tid) >> 3
  int a_nv[NQA], a_t[NQA]; bool a_ok[NQA]; uint32_t a_off[NQA];
  const int a_k = (tid & 7) * 4;
  const float* w_ptr[NQW]; uint32_t w_off[NQW];
#pragma unroll
  for (int q = 0; q < NQA; ++q) {
    const int idx = q * kTfThreads + tid, r_l = idx >> 3, j = idx & 7;
    a_off[q] = (uint32_t)(r_l * 128 + ((j ^ (r_l & 7)) << 4));
  }
#pragma unroll
  for (int q = 0; q < NQW; ++q) {
    const int idx = q * kTfThreads + tid, n_l = idx >> 3, j = idx & 7;
    w_off[q] = (uint32_t)(n_l * 128 + ((j ^ (n_l & 7)) << 4));
  }
  int n_chunks = 0;
  for (int s = 0; s < p.n_slabs; ++s) n_chunks += p.slabs[s].c / kTfKc;
  const int n_blocks = (n_chunks + kTfBlock - 1) / kTfBlock;

  // Global loads run TWO chunks ahead of the split (register double buffer): one chunk of distance left the L2 latency
  // exposed in every iteration (2850 cycles per chunk at K = 4352 against 12 x 64 cycles of MMAs).
  float4 ra[2][NQA], rw[2][kPre ? 1 : NQW];
  const int64_t w_small_delta = kPre ? p.w_small - p.w_big : 0;
  // kPre: weight buffer of (tile-relative) chunk c: the two stages' W slots, then the extra buffer
  auto w_buf = [&](int c) -> uint8_t* {
    const int b = c % NWB;
    return b < 2 ? smem + (size_t)b * STAGE + 2 * kTfATile : smem + 2 * STAGE + (size_t)(b - 2) * 2 * WT;
  };
  auto copy_w = [&](int c) {                         // cp.async of chunk c's W_big / W_small (one group per call, empty past the end)
    if (c < n_chunks) {
      uint8_t* wb = w_buf(c);
      const int wk = c * kTfKc;                      // the slabs are concatenated along K in order
#pragma unroll
      for (int q = 0; q < NQW; ++q) {
        if (w_ptr[q] != nullptr) {
          cp_async16(wb + w_off[q], w_ptr[q] + wk);
          cp_async16(wb + WT + w_off[q], w_ptr[q] + w_small_delta + wk);
        } else {
          *reinterpret_cast<float4*>(wb + w_off[q]) = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(wb + WT + w_off[q]) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    cp_async_commit();
  };
  int ls = 0, lk0 = 0, loaded = 0;                   // cursor of the next chunk to load
  uint32_t a_src[NQA];                               // this thread's rows in the cursor's slab as element offsets (~0: padding frame / row past the end)
  auto issue_loads = [&](auto PB) {
    constexpr int pb = decltype(PB)::value;
    if (loaded >= n_chunks) return;
    const F32Slab& sl = p.slabs[ls];
    if (lk0 == 0) {                                  // new slab: resolve the tap shift / stride / padding once per row
#pragma unroll
      for (int q = 0; q < NQA; ++q) {
        const int ts = a_t[q] * sl.t_mul + sl.t_off;
        const bool ok = a_ok[q] && ts >= 0 && ts < sl.t_in;
        a_src[q] = ok ? (uint32_t)((a_nv[q] * sl.t_in + ts) * sl.c + a_k) : 0xFFFFFFFFu;   // element offset (< 2^31: checked on the host)
      }
    }
#pragma unroll
    for (int q = 0; q < NQA; ++q)
      ra[pb][q] = a_src[q] != 0xFFFFFFFFu ? __ldg(reinterpret_cast<const float4*>(sl.a + a_src[q] + lk0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (!kPre) {
#pragma unroll
      for (int q = 0; q < NQW; ++q)
        rw[pb][q] = w_ptr[q] ? __ldg(reinterpret_cast<const float4*>(w_ptr[q] + sl.koff + lk0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ++loaded;
    lk0 += kTfKc;
    if (lk0 >= sl.c) { ++ls; lk0 = 0; }
  };

  // Persistent over tiles: the row / column bookkeeping of the NEXT tile and its first two chunks' loads are issued right
  // after the K loop of the current tile, so their latency hides behind the current tile's epilogue.
  auto setup_tile = [&](int tile) {
    const int tm = tile / tiles_n;                   // column tiles of one row tile run side by side: A is read from DRAM once
    row0 = tm * 128;
    col0 = (tile - tm * tiles_n) * BN;
#pragma unroll
    for (int q = 0; q < NQA; ++q) {
      const int r = row0 + ((q * kTfThreads + tid) >> 3);
      a_ok[q] = r < rows;
      a_nv[q] = a_ok[q] ? fast_div(r, p.div_t_magic, p.div_t_shift) : 0;
      a_t[q] = a_ok[q] ? r - a_nv[q] * p.t_out : 0;
    }
#pragma unroll
    for (int q = 0; q < NQW; ++q) {
      const int col = col0 + ((q * kTfThreads + tid) >> 3);
      w_ptr[q] = col < p.c_out ? (kPre ? p.w_big : p.w) + (int64_t)col * p.ktot + a_k : nullptr;
    }
    ls = 0; lk0 = 0; loaded = 0;
    issue_loads(std::integral_constant<int, 0>{});
    issue_loads(std::integral_constant<int, 1>{});
  };

  // K is accumulated in tensor memory in BLOCKS of kTfBlock chunks (alternating between two accumulators) and the block
  // sums are added up in registers (header comment: the tensor core's accumulation truncates).
  const int lg = warp & 3, cq = warp >> 2;
  float acc[NACC];
  int gch0 = 0, gblk0 = 0;                           // chunks / blocks this CTA has issued before the current tile (barrier phases, accumulator parity)
  auto drain = [&](int blk) {
    uint32_t a[NACC];
    const uint32_t taddr = tmem_acc + ((uint32_t)(lg * 32) << 16) + (uint32_t)(((gblk0 + blk) & 1) * BN + cq * NACC);
    if constexpr (NACC == 32) tmem_ld32(taddr, a); else tmem_ld16(taddr, a);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] += __uint_as_float(a[i]);
    tc_fence_before();
  };

  constexpr uint32_t idesc = make_idesc_tf32(128, BN);
  const uint32_t smem_base = smem_u32(smem);
  int drained = 0;
  auto body = [&](int ch, auto PB) {
    constexpr int pb = decltype(PB)::value;
    const int gch = gch0 + ch;                               // stage and barrier phase follow the CTA's running chunk count
    const int st = gch & 1;
    uint8_t* sa_big = smem + (size_t)st * STAGE;
    uint8_t* sa_small = sa_big + kTfATile;
    uint8_t* sw_big = sa_small + kTfATile;
    uint8_t* sw_small = sw_big + WT;
    if (ch >= 2) {                                           // the MMAs of chunk ch-2 have finished: its stage is free ...
      mbar_wait_warp(&mma_done[st], (uint32_t)(((gch >> 1) - 1) & 1));
      tc_fence_after();
      if ((ch - 2) % kTfBlock == kTfBlock - 1) { drain(drained); ++drained; }   // ... and if it closed a block, so is that block's sum
    }
    if (kPre) copy_w(ch + WDIST);                            // its buffer was last read by the MMAs of chunk ch + WDIST - NWB <= ch - 2: finished
#pragma unroll
    for (int q = 0; q < NQA; ++q) split_store(sa_big, sa_small, a_off[q], ra[pb][q]);
    if (kPre) {
      cp_async_wait<WDIST>();                                // chunk ch's weights have landed (the newest WDIST groups may still fly)
    } else {
#pragma unroll
      for (int q = 0; q < NQW; ++q) split_store(sw_big, sw_small, w_off[q], rw[pb][q]);
    }
    issue_loads(PB);                                         // chunk ch + 2 into the registers just consumed
    fence_proxy_async_smem();                                // generic-proxy stores -> visible to the MMA's async-proxy reads
    __syncthreads();
    if (warp == 0) {                                         // warp-uniform: the election is inside the asm (no uniformisation loop per MMA)
      tc_fence_after();
      const uint32_t ab = smem_base + (uint32_t)(st * STAGE), as = ab + kTfATile;
      const uint32_t wb = kPre ? smem_u32(w_buf(ch)) : as + kTfATile, ws = wb + WT;
      const uint32_t d = tmem_acc + (uint32_t)(((gblk0 + ch / kTfBlock) & 1) * BN);
      const bool fresh = ch % kTfBlock == 0;                 // first chunk of a block overwrites the accumulator
#pragma unroll
      for (int k = 0; k < kTfKc / 8; ++k) {
        const uint32_t ko = (uint32_t)k * 32u;               // 8 tf32 = 32 bytes along K inside the swizzled row
        umma_tf32_w(d, make_smem_desc_kmajor_sw128(as + ko), make_smem_desc_kmajor_sw128(wb + ko), idesc, (fresh && k == 0) ? 0u : 1u);
        umma_tf32_w(d, make_smem_desc_kmajor_sw128(ab + ko), make_smem_desc_kmajor_sw128(ws + ko), idesc, 1u);
        umma_tf32_w(d, make_smem_desc_kmajor_sw128(ab + ko), make_smem_desc_kmajor_sw128(wb + ko), idesc, 1u);
      }
      umma_commit_w(&mma_done[st]);
    }
  };
  if ((int)blockIdx.x < total_tiles) setup_tile((int)blockIdx.x);
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
  const int e_row0 = row0, e_col0 = col0;                    // the epilogue's tile (setup_tile moves row0 / col0 on to the next one)
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
  drained = 0;
  if (kPre) {
#pragma unroll
    for (int c = 0; c < WDIST; ++c) copy_w(c);               // the copies that run ahead of the first chunk
  }
  for (int ch = 0; ch < n_chunks; ch += 2) {
    body(ch, std::integral_constant<int, 0>{});
    if (ch + 1 < n_chunks) body(ch + 1, std::integral_constant<int, 1>{});
  }
  if (tile + (int)gridDim.x < total_tiles) setup_tile(tile + (int)gridDim.x);   // both load buffers are free: next tile's first chunks
  {
    const int last = gch0 + n_chunks - 1;                    // the commit of the last chunk covers every MMA of the tile
    mbar_wait_warp(&mma_done[last & 1], (uint32_t)((last >> 1) & 1));
    tc_fence_after();
    for (; drained < n_blocks; ++drained) drain(drained);
  }
  gch0 += n_chunks; gblk0 += n_blocks;
  __syncthreads();                                           // operand stages are dead: they become the staging tile

  // ---- epilogue: raw sums -> shared memory (thread = tile row), then one warp per row: bias / residual / activation and
  // a coalesced row store (a warp writes BN * 4 contiguous bytes)
  {
    const int r_l = lg * 32 + lane;
    uint8_t* srow = smem + (size_t)r_l * OUT_PITCH + (size_t)cq * NACC * 4;
#pragma unroll
    for (int i = 0; i < NACC; i += 4) *reinterpret_cast<float4*>(srow + i * 4) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
  }
  __syncthreads();
  {
    // A warp stores whole rows: BN = 128: one row per pass (32 lanes x 4 columns), BN = 64: two rows per pass (lane >> 4 picks
    // the row).  Passes go RB at a time: first the RB rows' residual / bias loads go out (their latencies overlap), then the
    // math and the stores.  The epilogue was 57 % of a K = 64 tile (in-kernel clock probe: 11.4 k of 20 k cycles, ~70
    // instructions per row with half the lanes idle at BN = 64): rows and pointers now advance incrementally, the
    // (clip, node, frame) split is computed only when the bias / layout / residual needs it, the activation is branch-free.
    constexpr int LPR = BN / 4;                        // lanes per row
    constexpr int RPP = 32 / LPR;                      // rows per pass
    constexpr int NW = kTfThreads / 32, NP = 128 / (NW * RPP), RB = 4;
    static_assert(NP % RB == 0, "passes per warp must be a multiple of the batch");
    const int sub = lane / LPR;                        // row of the pass this lane works on
    const int c_l = (lane % LPR) * 4;
    const int c = e_col0 + c_l;
    const bool lane_on = c < p.c_out;
    const bool vec_ok = c + 4 <= p.c_out;
    const int ld = p.out_layout == TIK_OUT_ROWS_F32 ? p.c_out_valid : p.c_out;
    const bool vec_st = c + 4 <= ld && (ld & 3) == 0;
    const bool need_info = p.bias_per_node || p.out_layout == TIK_OUT_TIME_MAJOR || p.res_kind == TIK_RES_STEM;
    const float act_s = p.act == TIK_ACT_RELU ? 0.f : (p.act == TIK_ACT_LEAKY ? p.slope : 1.f);   // act(v) = max(v, v * s), s in [0, 1]
    auto row_info = [&](int r, int& n, int& t, int& node) {
      const int nv = fast_div(r, p.div_t_magic, p.div_t_shift);
      t = r - nv * p.t_out;
      n = fast_div(nv, p.div_v_magic, p.div_v_shift);
      node = nv - n * p.v;
    };
    float4 bias_shared = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane_on && !p.bias_per_node) {
      if (vec_ok) bias_shared = __ldg(reinterpret_cast<const float4*>(p.bias + c));
      else { bias_shared.x = __ldg(p.bias + c); if (c + 1 < p.c_out) bias_shared.y = __ldg(p.bias + c + 1); if (c + 2 < p.c_out) bias_shared.z = __ldg(p.bias + c + 2); }
    }
    const int rr0 = warp * RPP + sub;                  // first tile row of this lane; a pass advances it by NW * RPP
#pragma unroll 1
    for (int i0 = 0; i0 < NP; i0 += RB) {
      float4 res4[RB], bias4[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int r = min(e_row0 + rr0 + (i0 + i) * NW * RPP, rows - 1);
        res4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        bias4[i] = bias_shared;
        if (lane_on) {
          if (p.bias_per_node) {
            int n, t, node;
            row_info(r, n, t, node);
            const float* bias = p.bias + node * p.c_out + c;
            if (vec_ok) bias4[i] = __ldg(reinterpret_cast<const float4*>(bias));
            else { bias4[i].x = __ldg(bias); if (c + 1 < p.c_out) bias4[i].y = __ldg(bias + 1); if (c + 2 < p.c_out) bias4[i].z = __ldg(bias + 2); }
          }
          if (p.res_kind == TIK_RES_IDENTITY) {
            const float* rp = reinterpret_cast<const float*>(p.res) + (int64_t)r * p.c_out + c;
            if (vec_ok) res4[i] = __ldg(reinterpret_cast<const float4*>(rp));
            else { res4[i].x = __ldg(rp); if (c + 1 < p.c_out) res4[i].y = __ldg(rp + 1); if (c + 2 < p.c_out) res4[i].z = __ldg(rp + 2); }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int rr = rr0 + (i0 + i) * NW * RPP;
        const int r = e_row0 + rr;
        if (r >= rows || !lane_on) continue;
        const float4 v4 = *reinterpret_cast<const float4*>(smem + (size_t)rr * OUT_PITCH + (size_t)c_l * 4);
        float o[4] = {v4.x + bias4[i].x + res4[i].x, v4.y + bias4[i].y + res4[i].y, v4.z + bias4[i].z + res4[i].z, v4.w + bias4[i].w + res4[i].w};
        int n = 0, t = 0, node = 0;
        if (need_info) row_info(r, n, t, node);
        if (p.res_kind == TIK_RES_STEM) {
          const float* xin = reinterpret_cast<const float*>(p.res) + ((((int64_t)n * p.res_t_in + (int64_t)t * p.res_t_mul) * p.v + node) * p.res_cin);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c + j < p.c_out) {
              const float* rw2 = p.res_w + ((int64_t)node * p.c_out + c + j) * p.res_cin;
              for (int ci = 0; ci < p.res_cin; ++ci) o[j] = fmaf(__ldg(rw2 + ci), __ldg(xin + ci), o[j]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], o[j] * act_s);
        float* dst;
        if (p.out_layout == TIK_OUT_TIME_MAJOR) dst = p.out + (((int64_t)n * p.t_out + t) * p.v + node) * (int64_t)p.c_out;
        else dst = p.out + (int64_t)r * ld;
        if (vec_st) {
          *reinterpret_cast<float4*>(dst + c) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < ld) dst[c + j] = o[j];
        }
      }
    }
  }
  __syncthreads();                                           // the staging tile has been read: the next tile may refill the stages
  }  // tile loop
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_acc);
  }
}

__global__ void tf32_split_kernel(const float* __restrict__ w, float* __restrict__ big, float* __restrict__ small, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = w[i], b = tf32_big(v);
    big[i] = b;
    small[i] = v - b;
  }
}

int tf32_split_weights(const float* w, float* big, float* small, int64_t n, cudaStream_t s) {
  if (n <= 0) return TIK_OK;
  tf32_split_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 1184), 256, 0, s>>>(w, big, small, n);
  TIK_LAUNCH_CHECK();
  return TIK_OK;
}

bool rowgemm_tf32_supported(const F32Args& a) {
  // On by default for the fp32 path; TIK_NO_TF32=1 forces the SIMT kernel (read per call: A/B inside one process).
  if (getenv("TIK_NO_TF32")) return false;
  if (a.ktot % 4 != 0 || a.c_out < 1 || a.v < 1 || a.t_out < 1 || a.n_slabs < 1 || a.rows >= (1ll << 31) - 128) return false;
  for (int i = 0; i < a.n_slabs; ++i)
    if ((a.rows / a.t_out + 1) * (int64_t)a.slabs[i].t_in * a.slabs[i].c >= (1ll << 31)) return false;   // 32-bit element offsets
  for (int i = 0; i < a.n_slabs; ++i)
    if (a.slabs[i].c % kTfKc != 0 || (reinterpret_cast<uintptr_t>(a.slabs[i].a) & 15) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(a.w) & 15) != 0) return false;
  if (a.w_big && ((reinterpret_cast<uintptr_t>(a.w_big) | reinterpret_cast<uintptr_t>(a.w_small)) & 15) != 0) return false;
  if (a.out_layout != TIK_OUT_ROWS_F32 && a.c_out % 4 != 0) return false;
  return true;
}

template <int BN, int TH, bool kPre>
static int launch_tf32(const F32Args& a, cudaStream_t s) {
  constexpr int smem = 2 * (2 * kTfATile + 2 * BN * 128) + 64 + 1024;
  static bool attr_done[64] = {};
  int dev = 0;
  TIK_CUDA(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    TIK_CUDA(cudaFuncSetAttribute(rowgemm_tf32_kernel<BN, TH, kPre>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done[dev & 63] = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = ceil_div(a.rows, 128) * ceil_div(a.c_out, BN);
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)sms * (BN == 64 ? 2 : 1));   // persistent: one CTA per resident slot
  rowgemm_tf32_kernel<BN, TH, kPre><<<grid, TH, smem, s>>>(a);
  TIK_LAUNCH_CHECK();
  return TIK_OK;
}

static void make_div(int d, unsigned long long* magic, int* shift) {   // exact for 0 <= n < 2^31 (checked exhaustively at the edges in the test-suite's shapes)
  int s = 0;
  while ((1ll << s) < d) ++s;
  *magic = (1ull << (31 + s)) / (unsigned long long)d + 1ull;
  *shift = 31 + s;
}

int rowgemm_tf32_launch(const F32Args& a0, cudaStream_t s) {
  F32Args a = a0;
  make_div(a.t_out, &a.div_t_magic, &a.div_t_shift);
  make_div(a.v, &a.div_v_magic, &a.div_v_shift);
  if (a.c_out > 64) return launch_tf32<128, 512, false>(a, s);
  return a.w_big != nullptr ? launch_tf32<64, 256, true>(a, s) : launch_tf32<64, 256, false>(a, s);
}

}  // namespace tik
