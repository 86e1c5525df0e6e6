// Stem of the ST-GCN backbone: everything block 0 does to the raw 3-channel keypoints, in one pass.
//
//   X0 = data_bn(x)                                             st_gcn_aaai18.py:119-125 (channel index v*C+c)
//   H0[n,w,t,:] = relu( Wg' . (sum_v A^[k,v,w] X0[n,t,v,:]) + b1[w,:] )   gconv_origin.py:56-65 + tcn.0/tcn.1
//   R0[n,w,t',:] = Wr'[w] . x[n, s*t', w, :]                    residual 1x1 conv + BN (st_gcn_aaai18.py:198-204),
//                                                               data_bn scale folded into Wr', constants into the
//                                                               temporal conv's per-node bias
//
// Input is the reference's (N,T,V,Cin) fp32 layout; outputs are node-major (N,V,T,Cout) so that the temporal
// convolution's TMA boxes are dense.  HBM-bound: Cin*4*V B in, 2 * V*Cout*sizeof(T) B out per frame.
// One CTA handles kStemFrames frames of one clip; each thread produces 8 consecutive channels (one 16 B store).
#include "tik_common.cuh"

namespace tik {

constexpr int kStemFrames = 32;
constexpr int kStemThreads = 256;
constexpr int kStemMaxV = 32;
constexpr int kStemMaxKC = 40;  // K * Cin

template <class T> __device__ __forceinline__ void store8(T* p, const float* a);
template <> __device__ __forceinline__ void store8<float>(float* p, const float* a) {
  reinterpret_cast<float4*>(p)[0] = make_float4(a[0], a[1], a[2], a[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(a[4], a[5], a[6], a[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float* a) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}

template <class OutT>
__global__ void __launch_bounds__(kStemThreads)
stem_gcn_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                const float* __restrict__ agg, const float* __restrict__ w, const float* __restrict__ bias,
                OutT* __restrict__ out, const float* __restrict__ res_w, OutT* __restrict__ res_out, int res_stride,
                int T, int V, int Cin, int K, int Cout, int relu, const TikWindowing win, long long win_n0) {
  extern __shared__ __align__(16) float smem[];
  const int KC = K * Cin, VC = V * Cin;
  float* s_raw = smem;                                // [frames][V*Cin] raw input
  float* s_a = s_raw + kStemFrames * VC;              // [frames][V][K*Cin] aggregated data_bn(x)
  float* s_wT = s_a + kStemFrames * V * KC;           // [K*Cin][Cout]
  float* s_b = s_wT + KC * Cout;                      // [V][Cout]
  float* s_agg = s_b + V * Cout;                      // [K][V][V]
  float* s_rw = s_agg + K * V * V;                    // [V][Cin][Cout] (only if res_w)
  const int tiles_t = (T + kStemFrames - 1) / kStemFrames;
  const int64_t n = blockIdx.x / tiles_t;
  const int t0 = (blockIdx.x % tiles_t) * kStemFrames;
  const int nf = min(kStemFrames, T - t0);

  for (int i = threadIdx.x; i < Cout * KC; i += kStemThreads) {
    const int c = i / KC, kc = i - c * KC;
    s_wT[kc * Cout + c] = __ldg(w + i);
  }
  for (int i = threadIdx.x; i < V * Cout; i += kStemThreads) s_b[i] = __ldg(bias + i);
  for (int i = threadIdx.x; i < K * V * V; i += kStemThreads) s_agg[i] = __ldg(agg + i);
  if (res_w != nullptr) {
    for (int i = threadIdx.x; i < V * Cout * Cin; i += kStemThreads) {   // (V,Cout,Cin) -> [V][Cin][Cout]
      const int ci = i % Cin, c = (i / Cin) % Cout, v = i / (Cin * Cout);
      s_rw[(v * Cin + ci) * Cout + c] = __ldg(res_w + i);
    }
  }
  float* s_bn = s_rw + (res_w != nullptr ? V * Cin * Cout : 0);   // [frames][V*Cin] after data_bn
  const float* gx = x + (n * T + t0) * (int64_t)VC;   // (N,T,V,C): frames contiguous
  for (int i = threadIdx.x; i < nf * VC; i += kStemThreads) {
    const int vc = i % VC;
    float raw;
    if (win.frames > 0) {
      // window mode: clip n is a window of one resident sequence (F,V,C); frame = clamp(n*stride + t + offset),
      // i.e. sample_window's edge padding (data_amass.py:18-42), root-centred on 0.5*(kp[a]+kp[b]) (:232-235)
      long long fi = (win_n0 + (long long)n) * win.stride + (t0 + i / VC) + win.offset;
      fi = fi < 0 ? 0 : (fi >= win.frames ? win.frames - 1 : fi);
      const float* fr = x + fi * VC;
      raw = __ldg(fr + vc);
      if (win.root_a >= 0) {
        const int c = vc % Cin;
        raw -= 0.5f * (__ldg(fr + win.root_a * Cin + c) + __ldg(fr + win.root_b * Cin + c));
      }
    } else {
      raw = __ldg(gx + i);
    }
    s_raw[i] = raw;
    s_bn[i] = fmaf(raw, __ldg(in_scale + vc), __ldg(in_shift + vc));
  }
  __syncthreads();
  // aggregated input: a[f][w][k*Cin+ci] = sum_v agg[k][v][w] * data_bn(x)[f][v][ci]
  for (int i = threadIdx.x; i < nf * V * KC; i += kStemThreads) {
    const int kc = i % KC, wv = (i / KC) % V, f = i / (KC * V);
    const int k = kc / Cin, ci = kc - k * Cin;
    float acc = 0.f;
    for (int v = 0; v < V; ++v) acc = fmaf(s_agg[(k * V + v) * V + wv], s_bn[f * VC + v * Cin + ci], acc);
    s_a[i] = acc;
  }
  __syncthreads();
  // outputs: item = (w, f, 8-channel group); consecutive threads -> consecutive channel groups of one (w,f)
  const int cg = Cout / 8;
  const int total = V * nf * cg;
  for (int i = threadIdx.x; i < total; i += kStemThreads) {
    const int g = i % cg;
    const int pair = i / cg;                        // (w, f) with f fastest -> adjacent threads write adjacent rows
    const int wv = pair / nf, f = pair - wv * nf;
    const int c0 = g * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = s_b[wv * Cout + c0 + j];
    const float* a = s_a + (f * V + wv) * KC;
    for (int kc = 0; kc < KC; ++kc) {
      const float av = a[kc];
      const float4 w0 = *reinterpret_cast<const float4*>(s_wT + kc * Cout + c0);
      const float4 w1 = *reinterpret_cast<const float4*>(s_wT + kc * Cout + c0 + 4);
      acc[0] = fmaf(w0.x, av, acc[0]); acc[1] = fmaf(w0.y, av, acc[1]); acc[2] = fmaf(w0.z, av, acc[2]); acc[3] = fmaf(w0.w, av, acc[3]);
      acc[4] = fmaf(w1.x, av, acc[4]); acc[5] = fmaf(w1.y, av, acc[5]); acc[6] = fmaf(w1.z, av, acc[6]); acc[7] = fmaf(w1.w, av, acc[7]);
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
    }
    store8<OutT>(out + ((n * V + wv) * (int64_t)T + t0 + f) * Cout + c0, acc);
    if (res_w != nullptr && ((t0 + f) % res_stride) == 0) {
      const int T_res = (T - 1) / res_stride + 1;
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = 0.f;
      for (int ci = 0; ci < Cin; ++ci) {
        const float xv = s_raw[f * VC + wv * Cin + ci];
        const float* rw = s_rw + (wv * Cin + ci) * Cout + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = fmaf(rw[j], xv, r[j]);
      }
      store8<OutT>(res_out + ((n * V + wv) * (int64_t)T_res + (t0 + f) / res_stride) * Cout + c0, r);
    }
  }
}

template <class T>
static int launch_stem(const float* x, const float* sc, const float* sh, const float* agg, const float* w,
                       const float* bias, void* out, const float* res_w, void* res_out, int res_stride, int64_t N,
                       int Tn, int V, int Cin, int K, int Cout, int relu, const TikWindowing& win, int64_t win_n0, cudaStream_t s) {
  const size_t KC = (size_t)K * Cin;
  size_t smem = sizeof(float) * (2 * (size_t)kStemFrames * V * Cin + (size_t)kStemFrames * V * KC + KC * Cout +
                                 (size_t)V * Cout + (size_t)K * V * V + (res_w ? (size_t)V * Cin * Cout : 0));
  TIK_CHECK_ARG(smem <= 200 * 1024, "stem shared memory %zu too large", smem);
  static bool attr_set[64] = {};
  int dev = 0;
  TIK_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {   // per device (and per template instantiation: the static is per T)
    TIK_CUDA(cudaFuncSetAttribute(stem_gcn_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set[dev & 63] = true;
  }
  int64_t blocks = N * ceil_div(Tn, kStemFrames);
  TIK_CHECK_ARG(blocks < (1ll << 31), "grid too large");
  stem_gcn_kernel<T><<<(unsigned)blocks, kStemThreads, smem, s>>>(x, sc, sh, agg, w, bias, reinterpret_cast<T*>(out), res_w,
                                                                   reinterpret_cast<T*>(res_out), res_stride, Tn, V, Cin, K,
                                                                   Cout, relu, win, (long long)win_n0);
  TIK_LAUNCH_CHECK();
  return TIK_OK;
}

}  // namespace tik

namespace tik {
int stem_gcn_impl(int dtype, const float* x, const float* in_scale, const float* in_shift, const float* agg,
                  const float* w, const float* bias, void* out, const float* res_w, void* res_out,
                  int res_stride, int64_t N, int T, int V, int Cin, int K, int Cout, int relu,
                  const TikWindowing* winp, int64_t win_n0, cudaStream_t s);
}

extern "C" int tik_stem_gcn(int dtype, const float* x, const float* in_scale, const float* in_shift, const float* agg,
                            const float* w, const float* bias, void* out, const float* res_w, void* res_out,
                            int res_stride, int64_t N, int T, int V, int Cin, int K, int Cout, int relu, void* stream) {
  return tik::stem_gcn_impl(dtype, x, in_scale, in_shift, agg, w, bias, out, res_w, res_out, res_stride, N, T, V, Cin, K, Cout,
                            relu, nullptr, 0, (cudaStream_t)stream);
}

int tik::stem_gcn_impl(int dtype, const float* x, const float* in_scale, const float* in_shift, const float* agg,
                       const float* w, const float* bias, void* out, const float* res_w, void* res_out,
                       int res_stride, int64_t N, int T, int V, int Cin, int K, int Cout, int relu,
                       const TikWindowing* winp, int64_t win_n0, cudaStream_t s) {
  using namespace tik;
  TikWindowing win;
  if (winp) win = *winp; else { win.frames = 0; win.offset = 0; win.stride = 1; win.root_a = -1; win.root_b = -1; }
  TIK_CHECK_ARG(win.frames == 0 || (win.stride >= 1 && win.root_a < V && win.root_b < V && (win.root_a < 0) == (win.root_b < 0)),
                "stem: bad windowing");
  TIK_CHECK_ARG(x && in_scale && in_shift && agg && w && bias && out, "null pointer");
  TIK_CHECK_ARG(N >= 0 && T > 0 && V > 0 && V <= kStemMaxV && Cin > 0 && K > 0 && K <= 5 && K * Cin <= kStemMaxKC &&
                    Cout > 0 && Cout % 8 == 0,
                "stem: unsupported shape N=%lld T=%d V=%d Cin=%d K=%d Cout=%d", (long long)N, T, V, Cin, K, Cout);
  TIK_CHECK_ARG((res_w == nullptr) == (res_out == nullptr), "stem: res_w and res_out go together");
  TIK_CHECK_ARG(res_w == nullptr || res_stride >= 1, "stem: bad residual stride");
  if (N == 0) return TIK_OK;
  if (dtype == TIK_F32)
    return launch_stem<float>(x, in_scale, in_shift, agg, w, bias, out, res_w, res_out, res_stride, N, T, V, Cin, K, Cout, relu, win, win_n0, s);
  if (dtype == TIK_BF16)
    return launch_stem<__nv_bfloat16>(x, in_scale, in_shift, agg, w, bias, out, res_w, res_out, res_stride, N, T, V, Cin, K, Cout, relu, win, win_n0, s);
  set_error("bad dtype %d", dtype);
  return TIK_ERR_INVALID;
}
