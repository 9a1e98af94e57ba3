// Human-frame producers: the arithmetic of the reference's loaders AFTER file parsing,
// for a whole batch of frames, writing the solve kernel's input layout (pos [F,nh,3], quat [F,nh,4] wxyz)
// directly instead of per-frame Python dicts:
//   * gmr_bvh_kernel   — utils/lafan1.py:17-35: quat_fk (lafan_vendor/utils.py:88-103), Y-up -> Z-up, cm -> m,
//                        LeftFootMod / RightFootMod (position of one joint, orientation of another);
//   * gmr_smplx_kernel — utils/smpl.py:127-196: resampling to the target rate (SLERP of from_rotvec'd neighbours,
//                        linear joint positions) and the global joint-orientation chain.
// One warp per frame, lanes = the joints the IK table needs plus their ancestors (<= 32, compacted on the host);
// the chain is walked by tree depth with warp shuffles (a lane reads its parent's global pose from the parent's
// lane), no shared memory.  HBM-bound: ~1 KB read + 0.4 KB written per frame.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct GmrChain {
  int32_t n, nlevel, nh, _pad;
  uint8_t orig[32];      // compact lane -> joint index in the source arrays
  int8_t parent[32];     // compact lane -> compact lane of its parent (-1: root)
  uint8_t level[32];     // tree depth of the lane's joint
  uint8_t pos_lane[32];  // output body -> lane (BVH) / source joint index (SMPL-X) its position comes from
  uint8_t rot_lane[32];  // output body -> lane its orientation comes from
};

namespace gmr_prod {

struct Q4 { float w, x, y, z; };
__device__ __forceinline__ Q4 qmul(const Q4& a, const Q4& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
__device__ __forceinline__ float3 qrot(const Q4& q, const float3& v) {
  const float tx = 2.f * (q.y * v.z - q.z * v.y), ty = 2.f * (q.z * v.x - q.x * v.z), tz = 2.f * (q.x * v.y - q.y * v.x);
  return make_float3(v.x + q.w * tx + (q.y * tz - q.z * ty), v.y + q.w * ty + (q.z * tx - q.x * tz), v.z + q.w * tz + (q.x * ty - q.y * tx));
}
__device__ __forceinline__ Q4 shfl(const Q4& q, int src) {
  return {__shfl_sync(0xffffffffu, q.w, src), __shfl_sync(0xffffffffu, q.x, src), __shfl_sync(0xffffffffu, q.y, src), __shfl_sync(0xffffffffu, q.z, src)};
}
__device__ __forceinline__ float3 shfl(const float3& v, int src) {
  return make_float3(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src), __shfl_sync(0xffffffffu, v.z, src));
}
// scipy Rotation.from_rotvec (small-angle series below 1e-3 rad)
__device__ __forceinline__ Q4 from_rotvec(float rx, float ry, float rz) {
  const float a2 = rx * rx + ry * ry + rz * rz, a = sqrtf(a2);
  float s, c;
  sincosf(0.5f * a, &s, &c);
  const float k = a <= 1e-3f ? 0.5f - a2 * (1.f / 48.f) + a2 * a2 * (1.f / 3840.f) : s / a;
  return {c, k * rx, k * ry, k * rz};
}
// utils/smpl.py:77-104
__device__ __forceinline__ Q4 slerp(Q4 a, Q4 b, float t) {
  float n = rsqrtf(a.w * a.w + a.x * a.x + a.y * a.y + a.z * a.z);
  a = {a.w * n, a.x * n, a.y * n, a.z * n};
  n = rsqrtf(b.w * b.w + b.x * b.x + b.y * b.y + b.z * b.z);
  b = {b.w * n, b.x * n, b.y * n, b.z * n};
  float dot = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;
  if (dot < 0.f) { b = {-b.w, -b.x, -b.y, -b.z}; dot = -dot; }
  float s0, s1;
  if (dot > 0.9995f) { s0 = 1.f - t; s1 = t; }
  else {
    const float th0 = acosf(dot), th = th0 * t, st = sinf(th), st0 = sinf(th0);
    s0 = cosf(th) - dot * st / st0; s1 = st / st0;
  }
  Q4 q = {s0 * a.w + s1 * b.w, s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z};
  n = rsqrtf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  if (q.w < 0.f) n = -n;                              // the as_rotvec() / from_rotvec() round trip: canonical sign
  return {q.w * n, q.x * n, q.y * n, q.z * n};
}
// global pose of every lane's joint from local ones, by tree depth
__device__ __forceinline__ void chain(const GmrChain& ch, int lane, Q4& q, float3* p) {
  const int par = lane < ch.n ? ch.parent[lane] : -1, lvl = lane < ch.n ? ch.level[lane] : 0;
  for (int l = 1; l < ch.nlevel; l++) {
    const Q4 pq = shfl(q, par < 0 ? 0 : par);
    if (p) {
      const float3 pp = shfl(*p, par < 0 ? 0 : par);
      if (lvl == l) { const float3 o = qrot(pq, *p); *p = make_float3(pp.x + o.x, pp.y + o.y, pp.z + o.z); }
    }
    if (lvl == l) q = qmul(pq, q);
  }
}

__global__ void __launch_bounds__(256)
gmr_bvh_kernel(const __grid_constant__ GmrChain ch, const float* __restrict__ lrot, const float* __restrict__ lpos, int F, int J,
               float* __restrict__ pos_out, float* __restrict__ quat_out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int f = blockIdx.x * wpb + (threadIdx.x >> 5); f < F; f += gridDim.x * wpb) {
    Q4 q = {1.f, 0.f, 0.f, 0.f};
    float3 p = make_float3(0.f, 0.f, 0.f);
    if (lane < ch.n) {
      const size_t j = (size_t)f * J + ch.orig[lane];
      const float4 t = *reinterpret_cast<const float4*>(lrot + j * 4);
      q = {t.x, t.y, t.z, t.w};
      p = make_float3(lpos[j * 3], lpos[j * 3 + 1], lpos[j * 3 + 2]);
    }
    chain(ch, lane, q, &p);
    const int b = lane < ch.nh ? lane : 0;
    const float3 P = shfl(p, ch.pos_lane[b]);
    const Q4 Q = shfl(q, ch.rot_lane[b]);
    if (lane < ch.nh) {
      // position @ [[1,0,0],[0,0,-1],[0,1,0]]^T / 100, orientation = quat(+90 deg about x) * Q   (lafan1.py:20-28)
      float* po = pos_out + ((size_t)f * ch.nh + lane) * 3;
      po[0] = P.x * 0.01f; po[1] = -P.z * 0.01f; po[2] = P.y * 0.01f;
      const Q4 r = qmul(Q4{0.70710678118654752f, 0.70710678118654752f, 0.f, 0.f}, Q);
      *reinterpret_cast<float4*>(quat_out + ((size_t)f * ch.nh + lane) * 4) = make_float4(r.w, r.x, r.y, r.z);
    }
  }
}

__global__ void __launch_bounds__(256)
gmr_smplx_kernel(const __grid_constant__ GmrChain ch, const float* __restrict__ global_orient, const float* __restrict__ full_pose,
                 const float* __restrict__ joints, int F, int NJ, int NJo, int Fo, int resample,
                 float* __restrict__ pos_out, float* __restrict__ quat_out) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int fo = blockIdx.x * wpb + (threadIdx.x >> 5); fo < Fo; fo += gridDim.x * wpb) {
    int i1 = fo, i2 = fo;
    float al = 0.f;
    if (resample) {                                          // np.linspace(0, F - 1, Fo)[fo]
      const double step = Fo > 1 ? (double)(F - 1) / (double)(Fo - 1) : 0.0, tt = fo * step;
      i1 = (int)floor(tt); i2 = i1 + 1 < F ? i1 + 1 : F - 1; al = (float)(tt - (double)i1);
    }
    Q4 q = {1.f, 0.f, 0.f, 0.f};
    if (lane < ch.n) {
      const int j = ch.orig[lane];
      const float* r1 = j == 0 ? global_orient + (size_t)i1 * 3 : full_pose + ((size_t)i1 * NJ + j) * 3;
      q = from_rotvec(r1[0], r1[1], r1[2]);
      if (resample) {
        const float* r2 = j == 0 ? global_orient + (size_t)i2 * 3 : full_pose + ((size_t)i2 * NJ + j) * 3;
        q = slerp(q, from_rotvec(r2[0], r2[1], r2[2]), al);
      }
    }
    chain(ch, lane, q, nullptr);
    const int b = lane < ch.nh ? lane : 0;
    const Q4 Q = shfl(q, ch.rot_lane[b]);
    if (lane < ch.nh) {
      const int j = ch.pos_lane[lane];                        // source joint index
      const float* a = joints + ((size_t)i1 * NJo + j) * 3;
      const float* c = joints + ((size_t)i2 * NJo + j) * 3;
      float* po = pos_out + ((size_t)fo * ch.nh + lane) * 3;
      po[0] = a[0] + al * (c[0] - a[0]); po[1] = a[1] + al * (c[1] - a[1]); po[2] = a[2] + al * (c[2] - a[2]);
      *reinterpret_cast<float4*>(quat_out + ((size_t)fo * ch.nh + lane) * 4) = make_float4(Q.w, Q.x, Q.y, Q.z);
    }
  }
}

// joints needed = the selected ones and all their ancestors, in source order (parents precede children)
inline int build_chain(const int32_t* parents, int J, const int32_t* rot_joint, int nh, GmrChain* ch, const char** why) {
  if (nh < 1 || nh > 32) { *why = "1..32 output bodies"; return -1; }
  if (J < 1 || J > 255) { *why = "1..255 source joints"; return -1; }
  bool need[256] = {false};
  for (int b = 0; b < nh; b++) {
    if (rot_joint[b] < 0 || rot_joint[b] >= J) { *why = "joint index out of range"; return -1; }
    for (int j = rot_joint[b]; j >= 0; j = parents[j]) { if (parents[j] >= j) { *why = "parents must precede children"; return -1; } need[j] = true; }
  }
  int lane_of[256];
  ch->n = 0; ch->nlevel = 1; ch->nh = nh;
  for (int j = 0; j < J; j++) {
    lane_of[j] = -1;
    if (!need[j]) continue;
    if (ch->n == 32) { *why = "more than 32 joints on the chains of the selected bodies"; return -4; }
    const int k = ch->n++;
    lane_of[j] = k;
    ch->orig[k] = (uint8_t)j;
    ch->parent[k] = (int8_t)(parents[j] < 0 ? -1 : lane_of[parents[j]]);
    ch->level[k] = (uint8_t)(parents[j] < 0 ? 0 : ch->level[lane_of[parents[j]]] + 1);
    if (ch->level[k] + 1 > ch->nlevel) ch->nlevel = ch->level[k] + 1;
  }
  for (int b = 0; b < nh; b++) ch->rot_lane[b] = (uint8_t)lane_of[rot_joint[b]];
  return 0;
}

}  // namespace gmr_prod
