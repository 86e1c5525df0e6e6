// ST-GCN building blocks on CUDA cores: stem graph convolution, adjacency aggregation and the
// fp32 implicit-GEMM used by the 1e-4 parity path.  Layout of every activation tensor is
// node-major (N, V, T, C): channels contiguous, then time, then graph node, then clip.
#include "tik_common.cuh"

namespace tik {

template <class T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }

// ------------------------------------------------------------------------------------------------
// Aggregation: out[k][(n,w),t,:] = sum_v agg[k][v][w] * x[(n,v),t,:]   (einsum of gconv_origin.py:63,
// moved in front of the channel GEMM).  One thread owns one 16-byte channel vector of one (n,t) and
// keeps the V input vectors in registers.
template <class T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) { float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  __device__ __forceinline__ static void store(float* p, const float* a) { *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]); }
};
template <> struct Vec16<__nv_bfloat16> {   // 4 bf16 = 8 bytes per thread: half the registers of a 16-byte vector, twice the warps
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 2; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ __forceinline__ static void store(__nv_bfloat16* p, const float* a) {
    uint2 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 2; ++i) h[i] = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

constexpr int kAggThreads = 256;

template <class T, int V>
__global__ void __launch_bounds__(kAggThreads)
aggregate_kernel(const T* __restrict__ x, const float* __restrict__ agg, T* __restrict__ out,
                 int64_t N, int Tn, int C, int K) {
  __shared__ float s_agg[5 * V * V];
  for (int i = threadIdx.x; i < K * V * V; i += kAggThreads) s_agg[i] = __ldg(agg + i);
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();                                               // x comes from the previous kernel
  constexpr int VN = Vec16<T>::N;
  const int cv = C / VN;                      // channel vectors per row
  const int64_t total = N * Tn * cv;
  const int64_t plane = N * V * (int64_t)Tn * C;
  for (int64_t i = (int64_t)blockIdx.x * kAggThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kAggThreads) {
    const int c0 = (int)(i % cv) * VN;
    const int t = (int)((i / cv) % Tn);
    const int64_t n = i / ((int64_t)cv * Tn);
    const T* px = x + ((n * V) * (int64_t)Tn + t) * C + c0;
    Vec16<T> xv[V];
#pragma unroll
    for (int v = 0; v < V; ++v) xv[v].load(px + (int64_t)v * Tn * C);
    for (int k = 0; k < K; ++k) {
      T* po = out + k * plane + ((n * V) * (int64_t)Tn + t) * C + c0;
#pragma unroll
      for (int w = 0; w < V; ++w) {
        float acc[VN];
#pragma unroll
        for (int j = 0; j < VN; ++j) acc[j] = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float a = s_agg[(k * V + v) * V + w];   // dense: 289 broadcast LDS + FMAs beat 289 dependent branches
#pragma unroll
          for (int j = 0; j < VN; ++j) acc[j] = fmaf(a, xv[v].v[j], acc[j]);
        }
        Vec16<T>::store(po + (int64_t)w * Tn * C, acc);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 implicit GEMM (parity path).  C[M x Nc] = sum_slabs gather(A_s)[M x c_s] . W[:, slab]^T.
// 128x64 tile, BK = 16, 256 threads, 8x4 register micro-tile.  M rows are (row group nv, frame t).
constexpr int GM = 128, GN = 64, GK = 16, GT = 256;


__global__ void __launch_bounds__(GT, 3) rowgemm_f32_kernel(const __grid_constant__ F32Args p) {
  __shared__ __align__(16) float As[GK][GM + 4];
  __shared__ __align__(16) float Bs[GK][GN + 4];
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * GM;
  const int col0 = blockIdx.y * GN;
  const int ty = tid / 16, tx = tid % 16;   // 16 x 16 threads; thread owns rows ty*8.., cols tx*4..
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // A loader: thread -> (row = tid/2, 8 consecutive k = (tid%2)*8 .. +8) as two float4
  const int lr = tid >> 1, lk = (tid & 1) * 8;
  const int64_t grow = row0 + lr;
  const bool row_ok = grow < p.rows;
  const int64_t g_nv = row_ok ? grow / p.t_out : 0;
  const int g_t = row_ok ? (int)(grow % p.t_out) : 0;
  // B loader: thread -> (col = tid/4, 4 consecutive k = (tid%4)*4)
  const int bc = tid >> 2, bk = (tid & 3) * 4;
  const bool col_ok = (col0 + bc) < p.c_out;

  // Software-pipelined K loop over all (slab, k0) chunks: the global loads of chunk i+1 are issued before the FMAs of
  // chunk i and stored to shared memory after them (the first version loaded, waited and computed in turn, which left
  // the FMA pipe idle for one L2 / DRAM latency per 16-wide chunk: 28 of 74 TFLOP/s at B=256).
  int n_chunks = 0;
  for (int s = 0; s < p.n_slabs; ++s) n_chunks += (p.slabs[s].c + GK - 1) / GK;
  float4 a0, a1, b0;
  auto issue = [&](int s, int k0) {
    const F32Slab sl = p.slabs[s];
    const int ts = g_t * sl.t_mul + sl.t_off;
    const bool a_ok = row_ok && ts >= 0 && ts < sl.t_in;
    const float* arow = sl.a + (g_nv * sl.t_in + (a_ok ? ts : 0)) * (int64_t)sl.c;
    const float* wrow = p.w + (int64_t)(col0 + bc) * p.ktot + sl.koff;
    a0 = make_float4(0, 0, 0, 0); a1 = a0; b0 = a0;
    if (a_ok) {
      if (k0 + lk + 8 <= sl.c) {
        a0 = __ldg(reinterpret_cast<const float4*>(arow + k0 + lk));
        a1 = __ldg(reinterpret_cast<const float4*>(arow + k0 + lk + 4));
      } else {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = (k0 + lk + j < sl.c) ? __ldg(arow + k0 + lk + j) : 0.f;
        a0 = make_float4(t[0], t[1], t[2], t[3]);
        a1 = make_float4(t[4], t[5], t[6], t[7]);
      }
    }
    if (col_ok) {
      if (k0 + bk + 4 <= sl.c) {
        b0 = __ldg(reinterpret_cast<const float4*>(wrow + k0 + bk));
      } else {
        float t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = (k0 + bk + j < sl.c) ? __ldg(wrow + k0 + bk + j) : 0.f;
        b0 = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
  };
  int cs = 0, ck0 = 0;
  issue(0, 0);
  for (int ch = 0; ch < n_chunks; ++ch) {
    __syncthreads();
    As[lk + 0][lr] = a0.x; As[lk + 1][lr] = a0.y; As[lk + 2][lr] = a0.z; As[lk + 3][lr] = a0.w;
    As[lk + 4][lr] = a1.x; As[lk + 5][lr] = a1.y; As[lk + 6][lr] = a1.z; As[lk + 7][lr] = a1.w;
    Bs[bk + 0][bc] = b0.x; Bs[bk + 1][bc] = b0.y; Bs[bk + 2][bc] = b0.z; Bs[bk + 3][bc] = b0.w;
    __syncthreads();
    if (ch + 1 < n_chunks) {
      ck0 += GK;
      if (ck0 >= p.slabs[cs].c) { ++cs; ck0 = 0; }
      issue(cs, ck0);
    }
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float4 ra0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      float4 ra1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      float4 rb = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ra[8] = {ra0.x, ra0.y, ra0.z, ra0.w, ra1.x, ra1.y, ra1.z, ra1.w};
      const float rbv[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ra[i], rbv[j], acc[i][j]);
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = row0 + ty * 8 + i;
    if (r >= p.rows) continue;
    const int64_t nv = r / p.t_out;
    const int t = (int)(r % p.t_out);
    const int node = (int)(nv % p.v);
    const int64_t n = nv / p.v;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      float val = acc[i][j];
      if (c < p.c_out) {
        val += __ldg(p.bias + (p.bias_per_node ? node * p.c_out : 0) + c);
        if (p.res_kind == TIK_RES_IDENTITY) {
          val += __ldg(reinterpret_cast<const float*>(p.res) + r * p.c_out + c);
        } else if (p.res_kind == TIK_RES_STEM) {
          const float* xin = reinterpret_cast<const float*>(p.res) +
                             ((n * p.res_t_in + (int64_t)t * p.res_t_mul) * p.v + node) * p.res_cin;
          const float* rw = p.res_w + ((int64_t)node * p.c_out + c) * p.res_cin;
          for (int ci = 0; ci < p.res_cin; ++ci) val = fmaf(__ldg(rw + ci), __ldg(xin + ci), val);
        }
        if (p.act == TIK_ACT_RELU) val = fmaxf(val, 0.f);
        else if (p.act == TIK_ACT_LEAKY) val = val > 0.f ? val : val * p.slope;
      }
      o[j] = val;
    }
    const int c = col0 + tx * 4;
    float* dst;
    int ld;
    if (p.out_layout == TIK_OUT_NODE_MAJOR) { dst = p.out + r * p.c_out; ld = p.c_out; }
    else if (p.out_layout == TIK_OUT_TIME_MAJOR) { dst = p.out + ((n * p.t_out + t) * p.v + node) * (int64_t)p.c_out; ld = p.c_out; }
    else { dst = p.out + r * p.c_out_valid; ld = p.c_out_valid; }
    if (c + 4 <= ld && (ld % 4) == 0) {
      *reinterpret_cast<float4*>(dst + c) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < ld) dst[c + j] = o[j];
    }
  }
}

int rowgemm_f32(const TikRowGemm* d, cudaStream_t s) { return rowgemm_f32_presplit(d, nullptr, nullptr, s); }

int rowgemm_f32_presplit(const TikRowGemm* d, const float* w_big, const float* w_small, cudaStream_t s) {
  F32Args a;
  a.w_big = w_big; a.w_small = w_small;
  int koff = 0;
  for (int i = 0; i < d->n_slabs; ++i) {
    const TikSlab& sl = d->slabs[i];
    TIK_CHECK_ARG(sl.a_dev && sl.c > 0 && sl.t_in > 0, "slab %d malformed", i);
    TIK_CHECK_ARG(sl.c % 4 == 0, "fp32 path needs channel counts divisible by 4 (slab %d has %d)", i, sl.c);
    a.slabs[i] = {reinterpret_cast<const float*>(sl.a_dev), sl.c, sl.t_in, sl.t_mul, sl.t_off, koff};
    koff += sl.c;
  }
  a.n_slabs = d->n_slabs;
  a.w = reinterpret_cast<const float*>(d->w_dev);
  a.ktot = koff;
  a.bias = d->bias_dev;
  a.bias_per_node = d->bias_per_node;
  a.rows = d->nv * d->t_out;
  a.v = d->v; a.t_out = d->t_out; a.c_out = d->c_out; a.c_out_valid = d->c_out_valid;
  a.act = d->act; a.slope = d->slope;
  a.res_kind = d->res_kind; a.res = d->res_dev; a.res_w = d->res_w_dev;
  a.res_cin = d->res_cin; a.res_t_mul = d->res_t_mul; a.res_t_in = d->res_t_in;
  a.out = reinterpret_cast<float*>(d->out_dev); a.out_layout = d->out_layout;
  if (a.rows == 0) return TIK_OK;
  if (rowgemm_tf32_supported(a)) return rowgemm_tf32_launch(a, s);     // 3xTF32 on the tensor pipe (rowgemm_tf32.cu)
  dim3 grid((unsigned)ceil_div(a.rows, GM), (unsigned)ceil_div(d->c_out, GN));
  rowgemm_f32_kernel<<<grid, GT, 0, s>>>(a);
  TIK_LAUNCH_CHECK();
  return TIK_OK;
}

template <class T, int V>
static int launch_agg_v(const void* x, const float* agg, void* out, int64_t N, int Tn, int C, int K, cudaStream_t s) {
  constexpr int VN = Vec16<T>::N;
  int64_t total = N * Tn * (C / VN);
  int64_t blocks = ceil_div(total, kAggThreads);
  if (blocks > 148 * 32) blocks = 148 * 32;
  static bool carve[64] = {};
  int dev = 0;
  TIK_CUDA(cudaGetDevice(&dev));
  if (!carve[dev & 63]) {   // same carveout as the tensor-core kernels (no L1/shared reconfiguration between launches)
    TIK_CUDA(cudaFuncSetAttribute(aggregate_kernel<T, V>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carve[dev & 63] = true;
  }
  TIK_CUDA(launch_pdl(aggregate_kernel<T, V>, (unsigned)blocks, kAggThreads, (size_t)0, s, reinterpret_cast<const T*>(x), agg,
                      reinterpret_cast<T*>(out), N, Tn, C, K));
  return TIK_OK;
}

template <class T>
static int launch_agg(const void* x, const float* agg, void* out, int64_t N, int Tn, int V, int C, int K, cudaStream_t s) {
  switch (V) {
    case 17: return launch_agg_v<T, 17>(x, agg, out, N, Tn, C, K, s);
    case 18: return launch_agg_v<T, 18>(x, agg, out, N, Tn, C, K, s);
    case 24: return launch_agg_v<T, 24>(x, agg, out, N, Tn, C, K, s);
    case 25: return launch_agg_v<T, 25>(x, agg, out, N, Tn, C, K, s);
    default: set_error("aggregate: graph with V=%d nodes is not instantiated (17, 18, 24, 25)", V); return TIK_ERR_UNSUPPORTED;
  }
}

}  // namespace tik

extern "C" {

int tik_aggregate(int dtype, const void* x, const float* agg, void* out, int64_t N, int T, int V, int C, int K, void* stream) {
  using namespace tik;
  TIK_CHECK_ARG(x && agg && out, "null pointer");
  TIK_CHECK_ARG(N >= 0 && T > 0 && C > 0 && K > 0 && K <= 5, "aggregate: bad shape");
  if (N == 0) return TIK_OK;
  if (dtype == TIK_F32) {
    TIK_CHECK_ARG(C % 4 == 0, "aggregate fp32: C=%d must be a multiple of 4", C);
    return launch_agg<float>(x, agg, out, N, T, V, C, K, (cudaStream_t)stream);
  }
  if (dtype == TIK_BF16) {
    TIK_CHECK_ARG(C % 4 == 0, "aggregate bf16: C=%d must be a multiple of 4", C);
    return launch_agg<__nv_bfloat16>(x, agg, out, N, T, V, C, K, (cudaStream_t)stream);
  }
  set_error("bad dtype %d", dtype);
  return TIK_ERR_INVALID;
}

int tik_rowgemm(int dtype, const TikRowGemm* d, void* stream) {
  using namespace tik;
  TIK_CHECK_ARG(d != nullptr, "null descriptor");
  TIK_CHECK_ARG(d->n_slabs >= 1 && d->n_slabs <= TIK_MAX_SLABS, "n_slabs=%d outside [1,%d]", d->n_slabs, TIK_MAX_SLABS);
  TIK_CHECK_ARG(d->w_dev && d->bias_dev && d->out_dev, "null pointer");
  TIK_CHECK_ARG(d->nv >= 0 && d->v > 0 && d->t_out > 0 && d->c_out > 0 && d->c_out_valid > 0 && d->c_out_valid <= d->c_out,
                "rowgemm: bad shape");
  TIK_CHECK_ARG(d->res_kind == TIK_RES_NONE || d->res_dev != nullptr, "residual pointer missing");
  TIK_CHECK_ARG(d->res_kind != TIK_RES_STEM || (d->res_w_dev && d->res_cin > 0 && d->res_cin <= 8), "stem residual malformed");
  if (dtype == TIK_F32) return rowgemm_f32(d, (cudaStream_t)stream);
  if (dtype == TIK_BF16) return rowgemm_bf16(d, (cudaStream_t)stream);
  set_error("bad dtype %d", dtype);
  return TIK_ERR_INVALID;
}

}  // extern "C"
