// Motion-space math shared by the alignment and mix/sampler kernels. fp32, mirroring the formula order of
// the reference (utils/rotation_conversions.py, utils/quaternion.py, utils/alignment.py) so that the
// ill-conditioned corners (matrix -> quaternion near 180 degrees) behave like the reference does.
#pragma once
#include "common.cuh"

namespace mmdm {

constexpr int NJ = 22;             // joints with positions / velocities
constexpr int NROT = 21;           // joints with a 6-D rotation
constexpr int NFEAT = 262;         // 66 pos + 66 vel + 126 rot6d + 4 foot contacts
constexpr int OFF_VEL = 66;
constexpr int OFF_ROT = 132;
constexpr int OFF_FC = 258;

struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// utils/quaternion.py:54-73  v + 2 * (w * (q x v) + q x (q x v))
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
  const V3 qv = v3(q.x, q.y, q.z);
  const V3 uv = cross3(qv, v);
  const V3 uuv = cross3(qv, uv);
  return v3(v.x + 2.f * (q.w * uv.x + uuv.x), v.y + 2.f * (q.w * uv.y + uuv.y), v.z + 2.f * (q.w * uv.z + uuv.z));
}

// utils/quaternion.py:386-396 (+ qnormalize :28-30)
__device__ __forceinline__ Q4 qbetween(V3 v0, V3 v1) {
  const V3 v = cross3(v0, v1);
  const float w = sqrtf(dot3(v0, v0) * dot3(v1, v1)) + dot3(v0, v1) + 1e-8f;
  const float n = sqrtf(w * w + v.x * v.x + v.y * v.y + v.z * v.z);
  Q4 q; q.w = w / n; q.x = v.x / n; q.y = v.y / n; q.z = v.z / n;
  return q;
}

__device__ __forceinline__ float sqrt_pos(float x) { return x > 0.f ? sqrtf(x) : 0.f; }
__device__ __forceinline__ float copysign_ref(float a, float b) { return ((a < 0.f) != (b < 0.f)) ? -a : a; }

// ih_to_smpl followed by smpl_to_ih on ONE joint (utils/alignment.py:11-67):
//   rot6d (interleaved a1=d[0,2,4], a2=d[1,3,5]) -> Gram-Schmidt matrix (rotation_conversions.py:513-535)
//   -> quaternion (:98-120) -> axis-angle (:482-510) -> *(-1) -> *(-1) -> quaternion (:450-479)
//   -> matrix (:38-66) -> first two rows, interleaved (:541-571).
__device__ __forceinline__ float sqrt_fast(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Same formula order as the reference; divisions and square roots use the approx instructions (2 ulp), which keeps
// the deviation from the reference's own fp32 evaluation at the 1e-6 level (tested at 1e-4) and takes the round
// trip from ~300 to ~90 instructions per joint - the fused mix kernel was issue-bound on it.
__device__ __forceinline__ void rot6d_roundtrip(const float* d, float* o) {
  V3 a1 = v3(d[0], d[2], d[4]);
  V3 a2 = v3(d[1], d[3], d[5]);
  const float i1 = rcp_approx(fmaxf(sqrt_fast(dot3(a1, a1)), 1e-12f));   // F.normalize eps (rotation_conversions.py:529)
  V3 b1 = v3(a1.x * i1, a1.y * i1, a1.z * i1);
  const float pr = dot3(b1, a2);
  V3 b2 = v3(a2.x - pr * b1.x, a2.y - pr * b1.y, a2.z - pr * b1.z);
  const float i2 = rcp_approx(fmaxf(sqrt_fast(dot3(b2, b2)), 1e-12f));
  b2 = v3(b2.x * i2, b2.y * i2, b2.z * i2);
  const V3 b3 = cross3(b1, b2);
  // matrix rows b1,b2,b3 -> quaternion
  const float m00 = b1.x, m11 = b2.y, m22 = b3.z;
  const float qw = 0.5f * sqrt_fast(fmaxf(1.f + m00 + m11 + m22, 0.f));
  float qx = 0.5f * sqrt_fast(fmaxf(1.f + m00 - m11 - m22, 0.f));
  float qy = 0.5f * sqrt_fast(fmaxf(1.f - m00 + m11 - m22, 0.f));
  float qz = 0.5f * sqrt_fast(fmaxf(1.f - m00 - m11 + m22, 0.f));
  qx = copysign_ref(qx, b3.y - b2.z);
  qy = copysign_ref(qy, b1.z - b3.x);
  qz = copysign_ref(qz, b2.x - b1.y);
  // quaternion -> axis-angle -> (*-1, *-1) -> quaternion (rotation_conversions.py:482-510, 450-479) is the identity
  // on (qw, qv) up to a factor 1 + O(|q|-1) that the normalising quaternion_to_matrix below removes, so the
  // atan2/sin/cos chain is not evaluated (difference to the reference's fp32 evaluation ~1e-6, tested at 1e-4).
  const float r = qw, i = qx, j = qy, k = qz;
  const float two_s = 2.f * rcp_approx(r * r + i * i + j * j + k * k);
  const float r00 = 1.f - two_s * (j * j + k * k), r01 = two_s * (i * j - k * r), r02 = two_s * (i * k + j * r);
  const float r10 = two_s * (i * j + k * r), r11 = 1.f - two_s * (i * i + k * k), r12 = two_s * (j * k - i * r);
  o[0] = r00; o[1] = r10; o[2] = r01; o[3] = r11; o[4] = r02; o[5] = r12;
}

// index of the Mixer weight (0..22) that multiplies per-person feature fi (models/mixermdm.py:753-779):
// joints for positions and velocities (x3 each), joints 0..20 for rot6d (x6), weight 22 for foot contacts.
__device__ __forceinline__ int weight_index(int fi) {
  if (fi < OFF_VEL) return fi / 3;
  if (fi < OFF_ROT) return (fi - OFF_VEL) / 3;
  if (fi < OFF_FC) return (fi - OFF_ROT) / 6;
  return 22;
}

// Alignment of an individual-model motion onto the interaction-model trajectory (utils/alignment.py:69-158).
struct AlignXf { Q4 q; V3 d0; V3 d1; };
// i0/iL: interaction root (joint 0) at frame 0 / last frame; s0/sL: individual root at frame 0 / last frame.
__device__ __forceinline__ AlignXf make_align(V3 i0, V3 iL, V3 s0, V3 sL) {
  AlignXf a;
  a.d0 = i0 - s0;
  const V3 s0t = s0 + a.d0, sLt = sL + a.d0;
  V3 v1 = iL - i0, v2 = sLt - s0t;
  v1.y = 0.f; v2.y = 0.f;
  const float n1 = sqrtf(dot3(v1, v1) + 1e-8f), n2 = sqrtf(dot3(v2, v2) + 1e-8f);
  v1 = v3(v1.x / n1, v1.y / n1, v1.z / n1);
  v2 = v3(v2.x / n2, v2.y / n2, v2.z / n2);
  a.q = qbetween(v2, v1);
  a.d1 = i0 - qrot(a.q, s0t);
  return a;
}

// center_motion (utils/alignment.py:161-222): floor to min-y, frame-0 root to xz origin, face +Z.
struct CenterXf { Q4 q; float floor_y; float sx, sz; };
// root/rhip/lhip: frame-0 joints 0, 2, 1 of the motion BEFORE centring; ymin: min y over all frames/joints.
__device__ __forceinline__ CenterXf make_center(V3 root, V3 rhip, V3 lhip, float ymin) {
  CenterXf c;
  c.floor_y = ymin;
  c.sx = root.x; c.sz = root.z;
  V3 ac = v3(rhip.x - lhip.x, (rhip.y - ymin) - (lhip.y - ymin), rhip.z - lhip.z);
  const float na = sqrtf(dot3(ac, ac));
  ac = v3(ac.x / na, ac.y / na, ac.z / na);
  V3 fw = cross3(v3(0.f, 1.f, 0.f), ac);
  const float nf = sqrtf(dot3(fw, fw));
  fw = v3(fw.x / nf, fw.y / nf, fw.z / nf);
  c.q = qbetween(fw, v3(0.f, 0.f, 1.f));
  return c;
}
__device__ __forceinline__ V3 center_pos(const CenterXf& c, V3 p) {
  p.y -= c.floor_y;
  p.x -= c.sx; p.y -= 0.f; p.z -= c.sz;
  return qrot(c.q, p);
}

}  // namespace mmdm
