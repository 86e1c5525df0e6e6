// Eval-mode weight folding on the device (tik_pack_bn / tik_pack_block): the reference's raw parameters -> the packed
// operands TikBlock describes.  Same algebra and the same rounding points as the host-side packer (engine.PackedNet):
//   s = gamma / sqrt(var + eps), o = beta - mean * s                                        (fp64)
//   agg   = A * importance                                                                   (fp32 product, st_gcn_aaai18.py:128-129)
//   w_gcn[c, k*Cin+ci] = Wg[k*Cout+c, ci] * s1[c]            b_gcn[w,c] = (sum_k bg[k*Cout+c] * colsum_k[w]) * s1[c] + o1[c]
//   w_tcn[c, dt*Cout+c'] = Wt[c,c',dt] * s2[c]  [| I | Wr*s3]   b_tcn[c] = s2[c]*bt[c] + o2[c] [+ s3[c]*br[c] + o3[c]]
//   first block, conv residual: w_res_stem[v,c,ci] = Wr[c,ci]*s3[c]*s0[v,ci], b_tcn[v,c] += sum_ci Wr[c,ci]*s3[c]*o0[v,ci]
// Every output element is one thread; the tensors are a few hundred KB, this runs once per weight update.
#include "tik_common.cuh"

// the ctypes mirrors in _lib.py are checked against these sizes (tests/test_packing_cpu.py)
static_assert(sizeof(TikRawBN) == 40 && sizeof(TikRawBlock) == 216 && sizeof(TikPackBuffers) == 48, "ABI layout changed");

namespace tik {
namespace {

struct BnDev {
  const float *w, *b, *m, *v;
  double eps;
};

__device__ __forceinline__ void bn_fold(const BnDev& bn, int64_t i, double* s, double* o) {
  if (bn.m == nullptr) { *s = 1.0; *o = 0.0; return; }
  const double g = bn.w ? (double)bn.w[i] : 1.0, b = bn.b ? (double)bn.b[i] : 0.0;
  const double sc = __ddiv_rn(g, __dsqrt_rn(__dadd_rn((double)bn.v[i], bn.eps)));
  *s = sc;
  *o = __dsub_rn(b, __dmul_rn((double)bn.m[i], sc));
}

// fp64 -> output element, rounded the way torch's .to() rounds (fp64 -> fp32 -> bf16)
__device__ __forceinline__ void store_elem(void* out, int64_t i, double v, int dtype) {
  const float f = (float)v;
  if (dtype == TIK_BF16) reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(f);
  else reinterpret_cast<float*>(out)[i] = f;
}

__global__ void pack_bn_kernel(BnDev bn, int64_t n, float* scale, float* shift) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s, o;
  bn_fold(bn, i, &s, &o);
  scale[i] = (float)s;
  shift[i] = (float)o;
}

__global__ void pack_agg_kernel(const float* A, const float* imp, int64_t n, float* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = imp ? __fmul_rn(A[i], imp[i]) : A[i];
}

struct GcnArgs {
  const float *w, *b, *agg;   // raw conv weight / bias, packed A * importance
  BnDev bn1;
  int K, V, cin, cout, dtype;
  void* w_out;
  float* b_out;
};

// threads [0, cout*K*cin): weights; [.., + V*cout): bias table
__global__ void pack_gcn_kernel(GcnArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nw = (int64_t)a.cout * a.K * a.cin;
  if (i < nw) {
    const int c = (int)(i / (a.K * a.cin)), rem = (int)(i % (a.K * a.cin)), k = rem / a.cin, ci = rem % a.cin;
    double s, o;
    bn_fold(a.bn1, c, &s, &o);
    store_elem(a.w_out, i, __dmul_rn((double)a.w[((int64_t)k * a.cout + c) * a.cin + ci], s), a.dtype);
  } else if (i < nw + (int64_t)a.V * a.cout) {
    const int j = (int)(i - nw), w = j / a.cout, c = j % a.cout;
    double s, o;
    bn_fold(a.bn1, c, &s, &o);
    double acc = 0.0;
    for (int k = 0; k < a.K; ++k) {
      double colsum = 0.0;
      for (int v = 0; v < a.V; ++v) colsum = __dadd_rn(colsum, (double)a.agg[((int64_t)k * a.V + v) * a.V + w]);
      const double bg = a.b ? (double)a.b[k * a.cout + c] : 0.0;
      acc = __dadd_rn(acc, __dmul_rn(bg, colsum));
    }
    a.b_out[j] = (float)__dadd_rn(__dmul_rn(acc, s), o);
  }
}

struct TcnArgs {
  const float *w, *b, *rw, *rb;   // temporal conv weight / bias, residual conv weight / bias
  BnDev bn2, bnr, bn0;            // bn0 = data_bn (first block)
  int kt, cin, cout, V, dtype;
  int tail;                       // 0 none, 1 identity slab, 2 residual-conv slab
  int stem;                       // first block with a conv residual
  int bias_rows;                  // 1 or V
  void* w_out;
  float* b_out;
  float* w_res_stem;
};

__global__ void pack_tcn_kernel(TcnArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int ktail = a.tail == 1 ? a.cout : (a.tail == 2 ? a.cin : 0);
  const int kw = a.kt * a.cout + ktail;
  const int64_t nw = (int64_t)a.cout * kw, nb = (int64_t)a.bias_rows * a.cout;
  const int64_t ns = a.stem ? (int64_t)a.V * a.cout * a.cin : 0;
  if (i < nw) {
    const int c = (int)(i / kw), col = (int)(i % kw);
    double v;
    if (col < a.kt * a.cout) {
      const int dt = col / a.cout, cp = col % a.cout;
      double s, o;
      bn_fold(a.bn2, c, &s, &o);
      v = __dmul_rn((double)a.w[((int64_t)c * a.cout + cp) * a.kt + dt], s);
    } else if (a.tail == 1) {
      v = (col - a.kt * a.cout) == c ? 1.0 : 0.0;
    } else {
      double s, o;
      bn_fold(a.bnr, c, &s, &o);
      v = __dmul_rn((double)a.rw[(int64_t)c * a.cin + (col - a.kt * a.cout)], s);
    }
    store_elem(a.w_out, i, v, a.dtype);
  } else if (i < nw + nb) {
    const int j = (int)(i - nw), node = j / a.cout, c = j % a.cout;
    double s2, o2;
    bn_fold(a.bn2, c, &s2, &o2);
    double b = __dadd_rn(__dmul_rn(s2, a.b ? (double)a.b[c] : 0.0), o2);
    if (a.rw) {
      double s3, o3;
      bn_fold(a.bnr, c, &s3, &o3);
      b = __dadd_rn(__dadd_rn(b, __dmul_rn(s3, a.rb ? (double)a.rb[c] : 0.0)), o3);
      if (a.stem) {
        double acc = 0.0;
        for (int ci = 0; ci < a.cin; ++ci) {
          double s0, o0;
          bn_fold(a.bn0, (int64_t)node * a.cin + ci, &s0, &o0);
          acc = __dadd_rn(acc, __dmul_rn(__dmul_rn((double)a.rw[(int64_t)c * a.cin + ci], s3), o0));
        }
        b = __dadd_rn(b, acc);
      }
    }
    a.b_out[j] = (float)b;
  } else if (i < nw + nb + ns) {
    const int64_t j = i - nw - nb;
    const int ci = (int)(j % a.cin), c = (int)((j / a.cin) % a.cout), node = (int)(j / ((int64_t)a.cin * a.cout));
    double s3, o3, s0, o0;
    bn_fold(a.bnr, c, &s3, &o3);
    bn_fold(a.bn0, (int64_t)node * a.cin + ci, &s0, &o0);
    a.w_res_stem[j] = (float)__dmul_rn(__dmul_rn((double)a.rw[(int64_t)c * a.cin + ci], s3), s0);
  }
}

BnDev bn_dev(const TikRawBN* b) {
  BnDev d = {nullptr, nullptr, nullptr, nullptr, 0.0};
  if (b) { d.w = b->weight_dev; d.b = b->bias_dev; d.m = b->mean_dev; d.v = b->var_dev; d.eps = b->eps; }
  return d;
}

int bn_ok(const TikRawBN* b, const char* what) {
  TIK_CHECK_ARG((b->mean_dev == nullptr) == (b->var_dev == nullptr), "%s: running_mean and running_var must both be set or both be NULL", what);
  TIK_CHECK_ARG(b->mean_dev != nullptr || (b->weight_dev == nullptr && b->bias_dev == nullptr), "%s: affine parameters without running statistics", what);
  TIK_CHECK_ARG(b->mean_dev == nullptr || b->eps >= 0.0, "%s: negative eps", what);
  return TIK_OK;
}

struct PackShape {
  int tail, stem, bias_rows, res_kind, res_as_slab;
  int64_t bytes[6];
};

int pack_shape(const TikRawBlock* r, int dtype, int first, PackShape* p) {
  TIK_CHECK_ARG(r != nullptr, "tik_pack_block: null block");
  TIK_CHECK_ARG(dtype == TIK_F32 || dtype == TIK_BF16, "tik_pack_block: dtype must be TIK_F32 or TIK_BF16");
  TIK_CHECK_ARG(r->c_in > 0 && r->c_out > 0 && r->kt > 0 && (r->kt & 1) && r->stride > 0 && r->K > 0 && r->V > 0,
                "tik_pack_block: bad sizes (c_in %d, c_out %d, kt %d, stride %d, K %d, V %d)", r->c_in, r->c_out, r->kt, r->stride, r->K, r->V);
  TIK_CHECK_ARG(r->residual == TIK_RES_NONE || r->residual == TIK_RES_IDENTITY || r->residual == TIK_RES_CONV,
                "tik_pack_block: residual must be TIK_RES_NONE, TIK_RES_IDENTITY or TIK_RES_CONV");
  const int64_t es = dtype == TIK_BF16 ? 2 : 4;
  p->tail = 0; p->stem = 0; p->bias_rows = 1; p->res_as_slab = 0; p->res_kind = r->residual;
  if (r->residual == TIK_RES_IDENTITY) {
    if (first) { set_error("tik_pack_block: an identity residual on the first block is not supported"); return TIK_ERR_UNSUPPORTED; }
    TIK_CHECK_ARG(r->c_in == r->c_out && r->stride == 1, "tik_pack_block: identity residual needs c_in == c_out and stride 1");
    if (dtype == TIK_BF16) { p->tail = 1; p->res_as_slab = 1; }
  } else if (r->residual == TIK_RES_CONV) {
    if (first) {
      if (r->c_in > 8) { set_error("tik_pack_block: first-block residual convolution needs c_in <= 8"); return TIK_ERR_UNSUPPORTED; }
      p->stem = 1; p->bias_rows = r->V; p->res_kind = TIK_RES_STEM;
      if (dtype == TIK_BF16) { p->tail = 1; p->res_as_slab = 1; }
    } else {
      p->tail = 2;
    }
  }
  const int ktail = p->tail == 1 ? r->c_out : (p->tail == 2 ? r->c_in : 0);
  p->bytes[0] = (int64_t)r->K * r->V * r->V * 4;
  p->bytes[1] = (int64_t)r->c_out * r->K * r->c_in * (first ? 4 : es);
  p->bytes[2] = (int64_t)r->V * r->c_out * 4;
  p->bytes[3] = (int64_t)r->c_out * (r->kt * r->c_out + ktail) * es;
  p->bytes[4] = (int64_t)p->bias_rows * r->c_out * 4;
  p->bytes[5] = p->stem ? (int64_t)r->V * r->c_out * r->c_in * 4 : 0;
  return TIK_OK;
}

}  // namespace
}  // namespace tik

using namespace tik;

extern "C" int tik_pack_bn(const TikRawBN* bn, int64_t n, float* scale_dev, float* shift_dev, void* stream) {
  TIK_CHECK_ARG(bn && scale_dev && shift_dev && n >= 0, "tik_pack_bn: null argument");
  int rc = bn_ok(bn, "tik_pack_bn");
  if (rc) return rc;
  if (n == 0) return TIK_OK;
  pack_bn_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(bn_dev(bn), n, scale_dev, shift_dev);
  TIK_LAUNCH_CHECK();
  return TIK_OK;
}

extern "C" int tik_pack_block_bytes(const TikRawBlock* raw, int dtype, int first_block, int64_t* bytes) {
  TIK_CHECK_ARG(bytes != nullptr, "tik_pack_block_bytes: null output");
  PackShape p;
  int rc = pack_shape(raw, dtype, first_block, &p);
  if (rc) return rc;
  for (int i = 0; i < 6; ++i) bytes[i] = p.bytes[i];
  return TIK_OK;
}

extern "C" int tik_pack_block(const TikRawBlock* raw, int dtype, int first_block, const TikRawBN* data_bn,
                              const TikPackBuffers* buf, TikBlock* out, void* stream) {
  PackShape p;
  int rc = pack_shape(raw, dtype, first_block, &p);
  if (rc) return rc;
  TIK_CHECK_ARG(buf && out, "tik_pack_block: null output");
  TIK_CHECK_ARG(raw->A_dev && raw->gcn_w_dev && raw->tcn_w_dev, "tik_pack_block: A, gcn_w and tcn_w are required");
  TIK_CHECK_ARG(buf->agg_dev && buf->w_gcn_dev && buf->b_gcn_dev && buf->w_tcn_dev && buf->b_tcn_dev, "tik_pack_block: null output buffer");
  TIK_CHECK_ARG(raw->residual != TIK_RES_CONV || raw->res_w_dev, "tik_pack_block: residual convolution without res_w");
  TIK_CHECK_ARG(!p.stem || buf->w_res_stem_dev, "tik_pack_block: first block with a conv residual needs w_res_stem_dev");
  if ((rc = bn_ok(&raw->bn1, "bn1")) || (rc = bn_ok(&raw->bn2, "bn2")) || (rc = bn_ok(&raw->bn_res, "bn_res"))) return rc;
  if (first_block && data_bn && (rc = bn_ok(data_bn, "data_bn"))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t na = (int64_t)raw->K * raw->V * raw->V;
  pack_agg_kernel<<<(unsigned)ceil_div(na, 256), 256, 0, s>>>(raw->A_dev, raw->importance_dev, na, buf->agg_dev);
  TIK_LAUNCH_CHECK();
  GcnArgs g = {raw->gcn_w_dev, raw->gcn_b_dev, buf->agg_dev, bn_dev(&raw->bn1), raw->K, raw->V, raw->c_in, raw->c_out,
               first_block ? TIK_F32 : dtype, buf->w_gcn_dev, buf->b_gcn_dev};
  const int64_t ng = (int64_t)raw->c_out * raw->K * raw->c_in + (int64_t)raw->V * raw->c_out;
  pack_gcn_kernel<<<(unsigned)ceil_div(ng, 256), 256, 0, s>>>(g);
  TIK_LAUNCH_CHECK();
  const bool conv = raw->residual == TIK_RES_CONV;
  TcnArgs t = {raw->tcn_w_dev, raw->tcn_b_dev, conv ? raw->res_w_dev : nullptr, conv ? raw->res_b_dev : nullptr,
               bn_dev(&raw->bn2), bn_dev(conv ? &raw->bn_res : nullptr), bn_dev(first_block ? data_bn : nullptr),
               raw->kt, raw->c_in, raw->c_out, raw->V, dtype, p.tail, p.stem, p.bias_rows, buf->w_tcn_dev, buf->b_tcn_dev,
               buf->w_res_stem_dev};
  const int64_t nt = p.bytes[3] / (dtype == TIK_BF16 ? 2 : 4) + p.bytes[4] / 4 + p.bytes[5] / 4;
  pack_tcn_kernel<<<(unsigned)ceil_div(nt, 256), 256, 0, s>>>(t);
  TIK_LAUNCH_CHECK();
  out->c_in = raw->c_in; out->c_out = raw->c_out; out->stride = raw->stride; out->kt = raw->kt;
  out->res_kind = p.res_kind; out->res_as_slab = p.res_as_slab;
  out->agg_dev = buf->agg_dev; out->w_gcn_dev = buf->w_gcn_dev; out->b_gcn_dev = buf->b_gcn_dev;
  out->w_tcn_dev = buf->w_tcn_dev; out->b_tcn_dev = buf->b_tcn_dev;
  out->w_res_stem_dev = p.stem ? buf->w_res_stem_dev : nullptr;
  return TIK_OK;
}
