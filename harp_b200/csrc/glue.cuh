// Device-side pieces of the per-frame camera / light glue (utils/visualize.py:266-270; renderer_helper.py:454-468;
// look_at_rotation SURVEY.md A.10), shared by glue.cu and the fused geometry kernels (geom.cu).
#pragma once
#include "common.cuh"

namespace {

constexpr float SHADOW_SPHERE_RADIUS = 1.5f;  // renderer_helper.py:463

// F.normalize(v, eps): v / max(|v|, eps)
template <typename T>
__device__ inline void normalize_eps(T v[3], float eps) {
  T n = dsqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (dval(n) > eps) { v[0] = v[0] / n; v[1] = v[1] / n; v[2] = v[2] / n; }
  else { T k(1.f / eps); v[0] = v[0] * k; v[1] = v[1] * k; v[2] = v[2] * k; }
}

template <typename T>
__device__ inline void cross3(const T a[3], const T b[3], T o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// Light camera on the radius-1.5 sphere around the centroid, looking at it: R (row-major, columns = camera x, y, z axes),
// T = -R^T eye.
template <typename T>
__device__ inline void light_camera(const T L[3], const T c[3], T R[9], T Tt[3]) {
  T off[3] = {L[0] - c[0], L[1] - c[1], L[2] - c[2]};
  T s = T(SHADOW_SPHERE_RADIUS) / dsqrt(off[0] * off[0] + off[1] * off[1] + off[2] * off[2]);
  T eye[3] = {c[0] + off[0] * s, c[1] + off[1] * s, c[2] + off[2] * s};
  T z[3] = {c[0] - eye[0], c[1] - eye[1], c[2] - eye[2]};
  normalize_eps(z, 1e-5f);
  T up[3] = {T(0.f), T(1.f), T(0.f)};
  T x[3], y[3];
  cross3(up, z, x);
  normalize_eps(x, 1e-5f);
  cross3(z, x, y);
  normalize_eps(y, 1e-5f);
  if (fabsf(dval(x[0])) <= 5e-3f && fabsf(dval(x[1])) <= 5e-3f && fabsf(dval(x[2])) <= 5e-3f) {  // light straight above / below
    cross3(y, z, x);
    normalize_eps(x, 1e-5f);
  }
  for (int i = 0; i < 3; ++i) { R[i * 3 + 0] = x[i]; R[i * 3 + 1] = y[i]; R[i * 3 + 2] = z[i]; }
  for (int j = 0; j < 3; ++j) Tt[j] = -(R[0 * 3 + j] * eye[0] + R[1 * 3 + j] * eye[1] + R[2 * 3 + j] * eye[2]);
}

}  // namespace
