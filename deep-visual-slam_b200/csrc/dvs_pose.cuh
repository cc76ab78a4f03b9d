// Pose chain of the VO learner as device functions (shared by the granular operator in dvs_ops.cu and the fused loss in
// dvs_fused.cu): transformation_from_parameters / rot_from_axisangle / get_translation_matrix, vo/learner_func.py:29-104.
// Rodrigues with axis = v / (|v| + 1e-7);  M = T(t) R, or R^T T(-t) when invert.
#pragma once
#include <cuda_runtime.h>

namespace dvs {

struct Rod {
  float x, y, z, ca, sa, C, ang, inv;
  float R[9];
};
__device__ __forceinline__ Rod rodrigues(const float* v) {
  Rod r;
  r.ang = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  r.inv = 1.0f / (r.ang + 1e-7f);
  r.x = v[0] * r.inv; r.y = v[1] * r.inv; r.z = v[2] * r.inv;
  r.ca = cosf(r.ang); r.sa = sinf(r.ang); r.C = 1.f - r.ca;
  float xs = r.x * r.sa, ys = r.y * r.sa, zs = r.z * r.sa;
  float xC = r.x * r.C, yC = r.y * r.C, zC = r.z * r.C;
  float xyC = r.x * yC, yzC = r.y * zC, zxC = r.z * xC;
  r.R[0] = r.x * xC + r.ca; r.R[1] = xyC - zs;          r.R[2] = zxC + ys;
  r.R[3] = xyC + zs;        r.R[4] = r.y * yC + r.ca;   r.R[5] = yzC - xs;
  r.R[6] = zxC - ys;        r.R[7] = yzC + xs;          r.R[8] = r.z * zC + r.ca;
  return r;
}
// (axis-angle v[3], translation t[3]) -> 4x4 row-major m[16]
__device__ __forceinline__ void pose_matrix(const float* v, const float* t, int invert, float* m) {
  const Rod r = rodrigues(v);
  if (!invert) {
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) m[i * 4 + j] = r.R[i * 3 + j];
      m[i * 4 + 3] = t[i];
    }
  } else {
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) m[i * 4 + j] = r.R[j * 3 + i];
      m[i * 4 + 3] = -(r.R[0 * 3 + i] * t[0] + r.R[1 * 3 + i] * t[1] + r.R[2 * 3 + i] * t[2]);
    }
  }
  m[12] = 0.f; m[13] = 0.f; m[14] = 0.f; m[15] = 1.f;
}
// d loss / d m[16] -> d loss / d v[3], d loss / d t[3]
__device__ __forceinline__ void pose_matrix_grad(const float* g, const float* v, const float* t, int invert, float* gv, float* gtr) {
  const Rod r = rodrigues(v);
  float gR[9], gt[3];
  if (!invert) {
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) gR[i * 3 + j] = g[i * 4 + j];
      gt[i] = g[i * 4 + 3];
    }
  } else {
    // M33 = R^T ; M3_i = -sum_j R[j][i] t_j
    for (int j = 0; j < 3; ++j) {
      float a = 0.f;
      for (int i = 0; i < 3; ++i) {
        gR[j * 3 + i] = g[i * 4 + j] - t[j] * g[i * 4 + 3];
        a -= r.R[j * 3 + i] * g[i * 4 + 3];
      }
      gt[j] = a;
    }
  }
  const float x = r.x, y = r.y, z = r.z, C = r.C, sa = r.sa, ca = r.ca;
  float s01 = gR[1] + gR[3], s02 = gR[2] + gR[6], s12 = gR[5] + gR[7];
  float a01 = gR[3] - gR[1], a02 = gR[2] - gR[6], a12 = gR[7] - gR[5];
  float gx = gR[0] * 2.f * x * C + s01 * y * C + s02 * z * C + a12 * sa;
  float gy = gR[4] * 2.f * y * C + s01 * x * C + s12 * z * C + a02 * sa;
  float gz = gR[8] * 2.f * z * C + s02 * x * C + s12 * y * C + a01 * sa;
  float gC = gR[0] * x * x + gR[4] * y * y + gR[8] * z * z + s01 * x * y + s02 * z * x + s12 * y * z;
  float gca = gR[0] + gR[4] + gR[8] - gC;
  float gsa = a01 * z + a02 * y + a12 * x;
  float gang = -sa * gca + ca * gsa - (gx * v[0] + gy * v[1] + gz * v[2]) * r.inv * r.inv;
  float rn = r.ang > 0.f ? 1.0f / r.ang : 0.f;
  gv[0] = gx * r.inv + gang * v[0] * rn;
  gv[1] = gy * r.inv + gang * v[1] * rn;
  gv[2] = gz * r.inv + gang * v[2] * rn;
  gtr[0] = gt[0]; gtr[1] = gt[1]; gtr[2] = gt[2];
}

}  // namespace dvs
