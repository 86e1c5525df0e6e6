// Rotation-representation conversions, fp32, HBM-bandwidth bound.
//
// One thread per rotation; each CTA stages its 256 rotations through shared memory so that
// global traffic is fully coalesced 128-bit loads/stores even though a rotation is 3, 6 or 9
// floats.  Precise libm (no fast-math): parity target is 1e-4 max-abs against the reference
// formulas, cited per function.
#include "tik_common.cuh"

namespace tik {

constexpr int kConvThreads = 256;

template <int IN, int OUT, class F>
__device__ __forceinline__ void staged_map(const float* __restrict__ in, float* __restrict__ out, int64_t M, F f) {
  __shared__ __align__(16) float s_in[kConvThreads * IN];
  __shared__ __align__(16) float s_out[kConvThreads * OUT];
  const int64_t base = (int64_t)blockIdx.x * kConvThreads;
  const int n = (int)min((int64_t)kConvThreads, M - base);
  const float* gin = in + base * IN;
  float* gout = out + base * OUT;
  if (n == kConvThreads) {
    const float4* g4 = reinterpret_cast<const float4*>(gin);
    float4* s4 = reinterpret_cast<float4*>(s_in);
#pragma unroll
    for (int i = threadIdx.x; i < kConvThreads * IN / 4; i += kConvThreads) s4[i] = __ldg(g4 + i);
  } else {
    for (int i = threadIdx.x; i < n * IN; i += kConvThreads) s_in[i] = gin[i];
  }
  __syncthreads();
  if ((int)threadIdx.x < n) {
    float a[IN], r[OUT];
#pragma unroll
    for (int k = 0; k < IN; ++k) a[k] = s_in[threadIdx.x * IN + k];
    f(a, r);
#pragma unroll
    for (int k = 0; k < OUT; ++k) s_out[threadIdx.x * OUT + k] = r[k];
  }
  __syncthreads();
  if (n == kConvThreads) {
    float4* g4 = reinterpret_cast<float4*>(gout);
    const float4* s4 = reinterpret_cast<const float4*>(s_out);
#pragma unroll
    for (int i = threadIdx.x; i < kConvThreads * OUT / 4; i += kConvThreads) g4[i] = s4[i];
  } else {
    for (int i = threadIdx.x; i < n * OUT; i += kConvThreads) gout[i] = s_out[i];
  }
}

// common/geometry.py:330-344 -- a1 = (x0,x2,x4), a2 = (x1,x3,x5); R columns = b1,b2,b3.
__device__ __forceinline__ void rot6d_one(const float* x, float* R) {
  float a1x = x[0], a1y = x[2], a1z = x[4], a2x = x[1], a2y = x[3], a2z = x[5];
  // one reciprocal per normalisation instead of three divisions (the kernels were issue-bound, profiles/r2_hbm_kernels.md)
  float i1 = 1.0f / fmaxf(sqrtf(a1x * a1x + a1y * a1y + a1z * a1z), 1e-6f);
  float b1x = a1x * i1, b1y = a1y * i1, b1z = a1z * i1;
  float d = b1x * a2x + b1y * a2y + b1z * a2z;
  float ux = a2x - d * b1x, uy = a2y - d * b1y, uz = a2z - d * b1z;
  float i2 = 1.0f / fmaxf(sqrtf(ux * ux + uy * uy + uz * uz), 1e-6f);
  float b2x = ux * i2, b2y = uy * i2, b2z = uz * i2;
  float b3x = b1y * b2z - b1z * b2y, b3y = b1z * b2x - b1x * b2z, b3z = b1x * b2y - b1y * b2x;
  R[0] = b1x; R[1] = b2x; R[2] = b3x;
  R[3] = b1y; R[4] = b2y; R[5] = b3y;
  R[6] = b1z; R[7] = b2z; R[8] = b3z;
}

// common/kornia_geometry_conversion.py:125-201
__device__ __forceinline__ void aa_kornia_one(const float* a, float* R) {
  float rx = a[0], ry = a[1], rz = a[2];
  float t2 = rx * rx + ry * ry + rz * rz;
  if (t2 > 1e-6f) {
    float th = sqrtf(t2);
    float inv = 1.0f / (th + 1e-6f);
    float wx = rx * inv, wy = ry * inv, wz = rz * inv;
    float s, c;
    sincosf(th, &s, &c);
    float k = 1.0f - c;
    R[0] = c + wx * wx * k;       R[1] = wx * wy * k - wz * s;  R[2] = wy * s + wx * wz * k;
    R[3] = wz * s + wx * wy * k;  R[4] = c + wy * wy * k;       R[5] = -wx * s + wy * wz * k;
    R[6] = -wy * s + wx * wz * k; R[7] = wx * s + wy * wz * k;  R[8] = c + wz * wz * k;
  } else {
    R[0] = 1.f; R[1] = -rz; R[2] = ry;
    R[3] = rz;  R[4] = 1.f; R[5] = -rx;
    R[6] = -ry; R[7] = rx;  R[8] = 1.f;
  }
}

// common/geometry.py:22-65 (batch_rodrigues + quat2mat)
__device__ __forceinline__ void rodrigues_one(const float* a, float* R) {
  float ex = a[0] + 1e-8f, ey = a[1] + 1e-8f, ez = a[2] + 1e-8f;
  float ang = sqrtf(ex * ex + ey * ey + ez * ez);
  float ia = 1.0f / ang;
  float nx = a[0] * ia, ny = a[1] * ia, nz = a[2] * ia;
  float s, c;
  sincosf(ang * 0.5f, &s, &c);
  float w = c, x = s * nx, y = s * ny, z = s * nz;
  float iq = rsqrtf(w * w + x * x + y * y + z * z);
  w *= iq; x *= iq; y *= iq; z *= iq;
  float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = w2 + x2 - y2 - z2; R[1] = 2.f * xy - 2.f * wz;   R[2] = 2.f * wy + 2.f * xz;
  R[3] = 2.f * wz + 2.f * xy; R[4] = w2 - x2 + y2 - z2;   R[5] = 2.f * yz - 2.f * wx;
  R[6] = 2.f * xz - 2.f * wy; R[7] = 2.f * wx + 2.f * yz; R[8] = w2 - x2 - y2 + z2;
}

// atan2(y, x) for x >= 0 (or NaN): |error| <= 1.1e-7 rad (degree-8 minimax-style fit of atan(z)/z in z^2 on [0,1],
// checked against float64 on 200,001 points in fp32 arithmetic; tools/fit_atan.py).  libm's atan2f costs ~60
// instructions with divergent paths; this is ~20, branch-free.
__device__ __forceinline__ float atan2_xpos(float y, float x) {
  const float ay = fabsf(y);
  const float mx = fmaxf(ay, x), mn = fminf(ay, x);
  const float z = mx > 0.0f ? __fdividef(mn, mx) : (mx == 0.0f ? 0.0f : mx);     // NaN propagates
  const float u = z * z;
  float p = 0.0028340641874819994f;
  p = fmaf(p, u, -0.016005029901862144f);
  p = fmaf(p, u, 0.042587608098983765f);
  p = fmaf(p, u, -0.07495445758104324f);
  p = fmaf(p, u, 0.10636754333972931f);
  p = fmaf(p, u, -0.14202570915222168f);
  p = fmaf(p, u, 0.19992484152317047f);
  p = fmaf(p, u, -0.3333306610584259f);
  p = fmaf(p, u, 1.0f);
  float r = z * p;
  r = ay > x ? 1.5707963267948966f - r : r;
  r = (x != x || y != y) ? x + y : r;                   // fmaxf / fminf drop NaNs: put them back as atan2f would
  return copysignf(r, y);
}

// common/geometry.py:100-150 (w,x,y,z).  The reference picks atan2(-s, -q0) when q0 < 0 and atan2(s, q0) otherwise:
// one atan2 of sign-flipped arguments, whose second argument is never negative.
__device__ __forceinline__ void quat_to_aa(float q0, float q1, float q2, float q3, float* aa) {
  const float s2 = q1 * q1 + q2 * q2 + q3 * q3;
  const float rs = rsqrtf(s2);
  const float s = s2 * rs;                              // sqrt(s2) for s2 > 0 (s2 == 0 takes the k = 2 branch below)
  const bool neg = q0 < 0.0f;
  const float two_theta = 2.0f * atan2_xpos(neg ? -s : s, neg ? -q0 : q0);
  const float k = s2 > 0.0f ? two_theta * rs : 2.0f;
  aa[0] = q1 * k; aa[1] = q2 * k; aa[2] = q3 * k;
}

// common/geometry.py:68-97,153-233: four-case selection on the transposed matrix, NaN -> 0.
// The ncu capture of round 2 (profiles/r2_hbm_kernels.md) showed the first version issue-bound (306 instructions per
// rotation: four divergent cases, two atan2f, five precise divisions): the cases are now selects over shared sums, the
// common factor 0.5 / sqrt(t) is one rsqrt (same inf / NaN behaviour as q / sqrt(t) * 0.5 for t <= 0).
__device__ __forceinline__ void rotmat_to_aa_one(const float* R, float* aa) {
  // m = R^T
  const float m00 = R[0], m01 = R[3], m02 = R[6];
  const float m10 = R[1], m11 = R[4], m12 = R[7];
  const float m20 = R[2], m21 = R[5], m22 = R[8];
  const bool d2 = m22 < 1e-6f, d01 = m00 > m11, d0n1 = m00 < -m11;
  const bool c0 = d2 && d01, c1 = d2 && !d01, c2 = !d2 && d0n1;           // else: c3
  const float s01 = m01 + m10, s20 = m20 + m02, s12 = m12 + m21;
  const float a12 = m12 - m21, a20 = m20 - m02, a01 = m01 - m10;
  const float t = c0 ? 1.f + m00 - m11 - m22 : c1 ? 1.f - m00 + m11 - m22 : c2 ? 1.f - m00 - m11 + m22 : 1.f + m00 + m11 + m22;
  float q0 = c0 ? a12 : c1 ? a20 : c2 ? a01 : t;
  float q1 = c0 ? t : c1 ? s01 : c2 ? s20 : a12;
  float q2 = c0 ? s01 : c1 ? t : c2 ? s12 : a20;
  float q3 = c0 ? s20 : c1 ? s12 : c2 ? t : a01;
  const float h = 0.5f * rsqrtf(t);
  q0 *= h; q1 *= h; q2 *= h; q3 *= h;
  quat_to_aa(q0, q1, q2, q3, aa);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (isnan(aa[k])) aa[k] = 0.f;
}

// common/kornia_geometry_conversion.py:204-307,396-439: (x,y,z,w) quaternion read as (w,x,y,z).
__device__ __forceinline__ void rotmat_to_aa_kornia_quirk_one(const float* R, float* aa) {
  const float tiny = 1.17549435e-38f;
  float m00 = R[0], m01 = R[1], m02 = R[2], m10 = R[3], m11 = R[4], m12 = R[5], m20 = R[6], m21 = R[7], m22 = R[8];
  float tr = m00 + m11 + m22;
  float qx, qy, qz, qw;
  if (tr > 0.f) {
    float sq = sqrtf(tr + 1.0f) * 2.f, dn = fmaxf(sq, tiny);
    qw = 0.25f * sq; qx = (m21 - m12) / dn; qy = (m02 - m20) / dn; qz = (m10 - m01) / dn;
  } else if (m00 > m11 && m00 > m22) {
    float sq = sqrtf(1.0f + m00 - m11 - m22 + 1e-8f) * 2.f, dn = fmaxf(sq, tiny);
    qw = (m21 - m12) / dn; qx = 0.25f * sq; qy = (m01 + m10) / dn; qz = (m02 + m20) / dn;
  } else if (m11 > m22) {
    float sq = sqrtf(1.0f + m11 - m00 - m22 + 1e-8f) * 2.f, dn = fmaxf(sq, tiny);
    qw = (m02 - m20) / dn; qx = (m01 + m10) / dn; qy = 0.25f * sq; qz = (m12 + m21) / dn;
  } else {
    float sq = sqrtf(1.0f + m22 - m00 - m11 + 1e-8f) * 2.f, dn = fmaxf(sq, tiny);
    qw = (m10 - m01) / dn; qx = (m02 + m20) / dn; qy = (m12 + m21) / dn; qz = 0.25f * sq;
  }
  quat_to_aa(qx, qy, qz, qw, aa);   // the quirk: element 0 (= x) is treated as cos(theta/2)
}

__global__ void __launch_bounds__(kConvThreads) rot6d_kernel(const float* in, float* out, int64_t M) {
  staged_map<6, 9>(in, out, M, [](const float* a, float* r) { rot6d_one(a, r); });
}
__global__ void __launch_bounds__(kConvThreads) aa_kornia_kernel(const float* in, float* out, int64_t M) {
  staged_map<3, 9>(in, out, M, [](const float* a, float* r) { aa_kornia_one(a, r); });
}
__global__ void __launch_bounds__(kConvThreads) rodrigues_kernel(const float* in, float* out, int64_t M) {
  staged_map<3, 9>(in, out, M, [](const float* a, float* r) { rodrigues_one(a, r); });
}
__global__ void __launch_bounds__(kConvThreads) rotmat_to_aa_kernel(const float* in, float* out, int64_t M, int quirk) {
  if (quirk)
    staged_map<9, 3>(in, out, M, [](const float* a, float* r) { rotmat_to_aa_kornia_quirk_one(a, r); });
  else
    staged_map<9, 3>(in, out, M, [](const float* a, float* r) { rotmat_to_aa_one(a, r); });
}

template <class K, class... Args>
static int launch_conv(K kernel, const void* in, void* out, int64_t M, cudaStream_t s, Args... extra) {
  TIK_CHECK_ARG(M >= 0, "negative count");
  if (M == 0) return TIK_OK;
  TIK_CHECK_ARG(in && out, "null pointer");
  TIK_CHECK_ARG((((uintptr_t)in | (uintptr_t)out) & 15) == 0, "pointers must be 16-byte aligned");
  int64_t blocks = ceil_div(M, kConvThreads);
  TIK_CHECK_ARG(blocks < (1ll << 31), "too many rotations for one launch");
  kernel<<<(unsigned)blocks, kConvThreads, 0, s>>>((const float*)in, (float*)out, M, extra...);
  TIK_LAUNCH_CHECK();
  return TIK_OK;
}

int launch_batch_rodrigues(const float* aa, float* R9, int64_t M, cudaStream_t s) {
  return launch_conv(rodrigues_kernel, aa, R9, M, s);
}

}  // namespace tik

extern "C" {
int tik_rot6d_to_rotmat(const float* x6, float* R, int64_t M, void* stream) {
  return tik::launch_conv(tik::rot6d_kernel, x6, R, M, (cudaStream_t)stream);
}
int tik_aa_to_rotmat(const float* aa, float* R, int64_t M, void* stream) {
  return tik::launch_conv(tik::aa_kornia_kernel, aa, R, M, (cudaStream_t)stream);
}
int tik_batch_rodrigues(const float* aa, float* R9, int64_t M, void* stream) {
  return tik::launch_conv(tik::rodrigues_kernel, aa, R9, M, (cudaStream_t)stream);
}
int tik_rotmat_to_aa(const float* R, float* aa, int64_t M, int kornia_quirk, void* stream) {
  return tik::launch_conv(tik::rotmat_to_aa_kernel, R, aa, M, (cudaStream_t)stream, kornia_quirk);
}
}
