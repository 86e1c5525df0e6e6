// Body forward kinematics: one warp per frame, one lane per joint.
//
//   G_0 = [R_0 | j_0],  G_i = G_parent(i) . [R_i | j_i - j_parent(i)],  joint_i = translation of G_i (+ transl)
//
// Replaces the joint output of common/smpl_util.py:67-70 (third-party smplx forward).
// The chain is evaluated by pointer jumping over the kinematic tree: in round r every lane composes
// its accumulated transform with that of its 2^r-th ancestor, fetched with warp shuffles, so a tree
// of depth D needs ceil(log2(D+1)) rounds of 12 shuffles (3 for the 22-joint SMPL-X body, depth 7)
// instead of one round per tree level.  Inputs/outputs are staged per warp through shared memory so
// that global traffic is contiguous.  HBM-bound: 264 B in, 264 + 792 B out per frame (aa in,
// joints + local rotmats out).
#include "tik_common.cuh"

namespace tik {

struct FkParams {
  float rest[TIK_MAX_JOINTS * 3];  // offset to parent (root: rest position), precomputed on host
  int8_t parent[TIK_MAX_JOINTS];
  int J;
  int rounds;
};

constexpr int kFkWarps = 8;

__device__ __forceinline__ void rodrigues_q(float ax, float ay, float az, float* R) {
  // common/geometry.py:22-65
  float ex = ax + 1e-8f, ey = ay + 1e-8f, ez = az + 1e-8f;
  float ang = sqrtf(ex * ex + ey * ey + ez * ez);
  float nx = ax / ang, ny = ay / ang, nz = az / ang;
  float s, c;
  sincosf(ang * 0.5f, &s, &c);
  float w = c, x = s * nx, y = s * ny, z = s * nz;
  float qn = sqrtf(w * w + x * x + y * y + z * z);
  w /= qn; x /= qn; y /= qn; z /= qn;
  float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = w2 + x2 - y2 - z2; R[1] = 2.f * xy - 2.f * wz;   R[2] = 2.f * wy + 2.f * xz;
  R[3] = 2.f * wz + 2.f * xy; R[4] = w2 - x2 + y2 - z2;   R[5] = 2.f * yz - 2.f * wx;
  R[6] = 2.f * xz - 2.f * wy; R[7] = 2.f * wx + 2.f * yz; R[8] = w2 - x2 - y2 + z2;
}

template <bool kRotIn>
__global__ void __launch_bounds__(kFkWarps * 32)
fk_kernel(const float* __restrict__ pose, const float* __restrict__ transl, float* __restrict__ joints,
          float* __restrict__ localR, float* __restrict__ globalR, int64_t F, const __grid_constant__ FkParams p) {
  constexpr int IN = kRotIn ? 9 : 3;
  __shared__ __align__(16) float s_io[kFkWarps][TIK_MAX_JOINTS * 9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int J = p.J;
  const bool active = lane < J;
  const int par0 = active ? p.parent[lane] : -1;
  const float dx = active ? p.rest[lane * 3 + 0] : 0.f;
  const float dy = active ? p.rest[lane * 3 + 1] : 0.f;
  const float dz = active ? p.rest[lane * 3 + 2] : 0.f;
  float* st = s_io[warp];
  const int64_t warps_total = (int64_t)gridDim.x * kFkWarps;

  for (int64_t f = (int64_t)blockIdx.x * kFkWarps + warp; f < F; f += warps_total) {
    // ---- coalesced load of this frame's pose into the warp's staging row
    const float* gp = pose + f * (int64_t)J * IN;
    for (int i = lane; i < J * IN; i += 32) st[i] = __ldg(gp + i);
    __syncwarp();
    float R[9];
    if (active) {
      if (kRotIn) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = st[lane * 9 + k];
      } else {
        rodrigues_q(st[lane * 3], st[lane * 3 + 1], st[lane * 3 + 2], R);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.f : 0.f;
    }
    __syncwarp();
    if (localR != nullptr) {
      if (active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) st[lane * 9 + k] = R[k];
      }
      __syncwarp();
      float* go = localR + f * (int64_t)J * 9;
      for (int i = lane; i < J * 9; i += 32) go[i] = st[i];
      __syncwarp();
    }
    // ---- pointer jumping
    float tx = dx, ty = dy, tz = dz;
    int anc = par0;
    for (int r = 0; r < p.rounds; ++r) {
      const int src = anc < 0 ? lane : anc;
      float A[9], ax, ay, az;
#pragma unroll
      for (int k = 0; k < 9; ++k) A[k] = __shfl_sync(0xffffffffu, R[k], src);
      ax = __shfl_sync(0xffffffffu, tx, src);
      ay = __shfl_sync(0xffffffffu, ty, src);
      az = __shfl_sync(0xffffffffu, tz, src);
      const int anc2 = __shfl_sync(0xffffffffu, anc, src);
      if (anc >= 0) {
        float ntx = A[0] * tx + A[1] * ty + A[2] * tz + ax;
        float nty = A[3] * tx + A[4] * ty + A[5] * tz + ay;
        float ntz = A[6] * tx + A[7] * ty + A[8] * tz + az;
        float N[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) N[i * 3 + j] = A[i * 3] * R[j] + A[i * 3 + 1] * R[3 + j] + A[i * 3 + 2] * R[6 + j];
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = N[k];
        tx = ntx; ty = nty; tz = ntz;
        anc = anc2;
      }
    }
    if (transl != nullptr) {
      tx += __ldg(transl + f * 3 + 0);
      ty += __ldg(transl + f * 3 + 1);
      tz += __ldg(transl + f * 3 + 2);
    }
    // ---- coalesced stores
    if (active) { st[lane * 3] = tx; st[lane * 3 + 1] = ty; st[lane * 3 + 2] = tz; }
    __syncwarp();
    float* gj = joints + f * (int64_t)J * 3;
    for (int i = lane; i < J * 3; i += 32) gj[i] = st[i];
    __syncwarp();
    if (globalR != nullptr) {
      if (active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) st[lane * 9 + k] = R[k];
      }
      __syncwarp();
      float* gg = globalR + f * (int64_t)J * 9;
      for (int i = lane; i < J * 9; i += 32) gg[i] = st[i];
      __syncwarp();
    }
  }
}

}  // namespace tik

extern "C" int tik_fk_body(const float* pose_dev, int pose_is_rotmat, const float* rest_host,
                           const int32_t* parents_host, int J, const float* transl_dev, float* joints_dev,
                           float* local_R_dev, float* global_R_dev, int64_t F, void* stream) {
  using namespace tik;
  TIK_CHECK_ARG(J >= 1 && J <= TIK_MAX_JOINTS, "J=%d outside [1,%d]", J, TIK_MAX_JOINTS);
  TIK_CHECK_ARG(F >= 0, "negative frame count");
  if (F == 0) return TIK_OK;
  TIK_CHECK_ARG(pose_dev && rest_host && parents_host && joints_dev, "null pointer");
  FkParams p;
  int depth[TIK_MAX_JOINTS], maxd = 0;
  for (int i = 0; i < J; ++i) {
    int par = parents_host[i];
    TIK_CHECK_ARG(par < i && par >= -1, "parents[%d]=%d must be -1 or an earlier joint", i, par);
    p.parent[i] = (int8_t)par;
    depth[i] = par < 0 ? 0 : depth[par] + 1;
    if (depth[i] > maxd) maxd = depth[i];
    for (int k = 0; k < 3; ++k)
      p.rest[i * 3 + k] = par < 0 ? rest_host[i * 3 + k] : rest_host[i * 3 + k] - rest_host[par * 3 + k];
  }
  for (int i = J; i < TIK_MAX_JOINTS; ++i) { p.parent[i] = -1; p.rest[i * 3] = p.rest[i * 3 + 1] = p.rest[i * 3 + 2] = 0.f; }
  p.J = J;
  p.rounds = 0;
  while ((1 << p.rounds) < maxd + 1) ++p.rounds;
  int sms = 148;
  int64_t blocks = ceil_div(F, kFkWarps);
  int64_t cap = (int64_t)sms * 8 * 4;   // persistent-ish: grid-stride over frames
  if (blocks > cap) blocks = cap;
  cudaStream_t s = (cudaStream_t)stream;
  if (pose_is_rotmat)
    fk_kernel<true><<<(unsigned)blocks, kFkWarps * 32, 0, s>>>(pose_dev, transl_dev, joints_dev, local_R_dev, global_R_dev, F, p);
  else
    fk_kernel<false><<<(unsigned)blocks, kFkWarps * 32, 0, s>>>(pose_dev, transl_dev, joints_dev, local_R_dev, global_R_dev, F, p);
  TIK_LAUNCH_CHECK();
  return TIK_OK;
}
