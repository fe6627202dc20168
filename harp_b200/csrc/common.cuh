// Shared device/host helpers for libharpb200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/harp_b200.h"

#define HARP_SM_COUNT 148  // B200: 2 dies x 74 SMs

void harp_set_error(const std::string& msg);
void harp_count_launches(int n);
struct harp_ctx;
void harp_prof_begin(harp_ctx* c, int kernel_id, cudaStream_t st);
void harp_prof_end(harp_ctx* c, int kernel_id, cudaStream_t st);

#define HARP_CHECK_CUDA(expr)                                                                       \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      harp_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                           \
      return HARP_ERR_CUDA;                                                                         \
    }                                                                                               \
  } while (0)

#define HARP_REQUIRE(cond, code, msg)                                                               \
  do {                                                                                              \
    if (!(cond)) {                                                                                  \
      harp_set_error(msg);                                                                          \
      return code;                                                                                  \
    }                                                                                               \
  } while (0)

#define HARP_LAUNCH_CHECK_N(n)                                                                      \
  do {                                                                                              \
    harp_count_launches(n);                                                                         \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess) {                                                                        \
      harp_set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));                      \
      return HARP_ERR_CUDA;                                                                         \
    }                                                                                               \
  } while (0)
#define HARP_LAUNCH_CHECK() HARP_LAUNCH_CHECK_N(1)

// Topology tables + model constants + scratch, all device resident.
struct harp_ctx {
  int V0 = 0, F0 = 0, E0 = 0, V = 0, F = 0, E = 0, Vt = 0, NP = 0;  // NP = normal-consistency pairs
  bool subdivide = false, has_model = false, has_uv = false;
  // mesh tables
  int32_t* faces = nullptr;      // (F,3) final faces
  int32_t* edges0 = nullptr;     // (E0,2) coarse edges (subdivision)
  int32_t* v0e_off = nullptr;    // (V0+1) coarse vertex -> incident coarse edges CSR
  int32_t* v0e_idx = nullptr;
  int32_t* vf_off = nullptr;     // (V+1) vertex -> incident (face*4 + corner) CSR
  int32_t* vf_idx = nullptr;
  uint32_t* vf_ring = nullptr;   // per incidence (same order as vf_idx): next | prev << 16, the face's other two vertices in
                                 // cyclic order after this one, so the corner's cross product is (next - self) x (prev - self)
  uint4* vf_ring8 = nullptr;     // (V,2) the same entries at a fixed stride of eight per vertex, padded with 0xffffffff: one aligned
                                 // 32-byte load and no offset lookup (a dependent global access less per gather); null when some
                                 // vertex has more than eight incident faces (the CSR tables above are then used)
  int32_t* vn_off = nullptr;     // (V+1) vertex -> neighbour vertex CSR (final mesh edges)
  int32_t* vn_idx = nullptr;
  int32_t* edges = nullptr;      // (E,2) final edges, PyTorch3D edges_packed order
  int32_t* nc_pairs = nullptr;   // (NP,4): v0, v1 of the shared edge, opposite verts a, b
  float* face_uvs = nullptr;     // (F,3,2)
  // skinning model (MANO: J=16, n_shape=10; SMPL-X arm: J=55, n_shape=10)
  int J = 0, n_shape = 0, rodrigues_kind = 0, center_joint = -1, n_joints_out = 0, n_tips = 0;
  float *shapedirs = nullptr, *posedirs = nullptr;      // transposed: (n_shape, 3 V0), ((J-1)*9, 3 V0)
  float *v_template = nullptr, *weights = nullptr, *pose_mean = nullptr;
  float* weightsT = nullptr;      // (J,V0) transposed skinning weights
  float *J_template = nullptr, *J_shapedirs = nullptr;  // (J,3), (J,3,n_shape) = J_regressor folded
  int32_t *tips = nullptr, *parents = nullptr, *joint_out_src = nullptr;
  // scratch (harp_ctx_reserve)
  int maxB = 0, S = 0, tiles = 0;
  bool defer_fold = false;        // host-side: harp_render_fused leaves d_verts4 / d_normals4 unfolded (harp_step_fused)
  int pins = 0;                   // captured CUDA graphs that hold the scratch pointers (harp_ctx_pin): no reallocation while > 0
  int64_t bin_cap = 0;            // bin list capacity per frame
  float* ndc_cam = nullptr;       // (B,V,3)
  float* ndc_light = nullptr;     // (B,V,3)
  float* d_ndc_cam = nullptr;     // (B,V,4) xyz + pad: 16-byte rows so one vector atomic covers a vertex
  float* d_ndc_light = nullptr;   // (B,V,4)
  float* d_verts4 = nullptr;      // (B,V,4) shade-backward accumulators, folded into the caller's (B,V,3) tensors
  float* d_normals4 = nullptr;    // (B,V,4)
  float* d_tex4 = nullptr;        // (2,texcap,4) texture / normal-map gradient accumulators
  float* tex4 = nullptr;          // (2,texcap,4) packed copies read by the shading taps: albedo (r,g,b,0) and the NORMALISED
                                  // normal map (n.x, n.y, n.z, 1/|m|; negative when |m| <= 1e-12), one 16-byte load per tap
  int64_t texcap = 0;
  int32_t* bin_count = nullptr;   // (2,B,tiles)  [0]=camera view, [1]=light view
  int32_t* bin_start = nullptr;   // (2,B,tiles)
  int32_t* bin_fill = nullptr;    // (2,B,tiles)
  int32_t* bin_list = nullptr;    // (2,B,bin_cap)
  struct FaceTile* face_rec = nullptr;  // (2,B,F) per-face raster records (144 B), written once per frame and view by the setup
                                        // kernel (setup.cu) and staged into shared memory by bulk async copies (TMA)
  int4* work = nullptr;           // (2,B*tiles) non-empty tiles {frame, tile, list start, face count}: heavy tiles from the front,
                                  // light ones from the back of each view's array (one dynamic queue, heavy first)
  int32_t* work_ctl = nullptr;    // (2,4) per view: {heavy count, light count, consumer cursor, unused}
  float* light_zbuf = nullptr;    // (B,S,S)
  int32_t* light_face = nullptr;  // (B,S,S)
  float* d_light_zbuf = nullptr;  // (B,S,S)
  int32_t* hit_list = nullptr;    // (B,S*S,2) compacted (row*S+col, nearest camera-view face) of hit pixels, per frame
  int32_t* hit_count = nullptr;   // (B)
  float* vN = nullptr;            // (B,V,3) unnormalised vertex normals (mesh build scratch)
  float* vtmp = nullptr;          // (B,V,3) scratch
  float* vtmp2 = nullptr;         // (B,V,3) scratch
  int32_t* ovf_list = nullptr;    // pixels (b*S*S + row*S + col) deferred to the exact top-K silhouette kernel
  int32_t* ovf_count = nullptr;   // (1)
  int ovf_cap = 0;
  unsigned long long* counters = nullptr;  // [0] top-K pixels that could not be handled exactly, [1] bin overflow
  struct harp_prof* prof = nullptr;        // host-side event lists (harp_ctx_profile)
  // host-side handles for the fused render's light branch (forked from / joined to the caller's stream; capturable)
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_setup = nullptr;
  cudaStream_t reg_stream = nullptr;       // harp_step_fused's regulariser branch
  cudaEvent_t ev_reg_fork = nullptr, ev_reg_join = nullptr;
};

// ---------------------------------------------------------------------------------------------- small vector math
struct f3 {
  float x, y, z;
};
__host__ __device__ inline f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
__host__ __device__ inline f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ inline f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ inline f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ inline f3 operator*(f3 a, float s) { return mk3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ inline f3 cross(f3 a, f3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ inline f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
__device__ inline void st3(float* p, f3 v) {
  p[0] = v.x;
  p[1] = v.y;
  p[2] = v.z;
}
__device__ inline void atomic_add3(float* p, f3 v) {
  atomicAdd(p + 0, v.x);
  atomicAdd(p + 1, v.y);
  atomicAdd(p + 2, v.z);
}

// one 16-byte vector reduction (sm_90+): 3x fewer L2 atomic operations than three scalar REDs
__device__ inline void atomic_add4(float* p, float x, float y, float z) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(x, y, z, 0.f));
}

__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0. `red` = shared float[32].
__device__ inline float block_sum(float v, float* red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? red[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

// ---------------------------------------------------------------------------------------------- forward-mode duals
// Used for the tiny per-frame chains (Rodrigues, look-at light camera) where a Jacobian of a handful of outputs w.r.t.
// a handful of inputs is cheaper and safer than hand-written reverse mode.
template <int N>
struct Dual {
  float v;
  float d[N];
  __host__ __device__ Dual() : v(0.f) {
#pragma unroll
    for (int i = 0; i < N; ++i) d[i] = 0.f;
  }
  __host__ __device__ Dual(float c) : v(c) {
#pragma unroll
    for (int i = 0; i < N; ++i) d[i] = 0.f;
  }
  __host__ __device__ static Dual var(float c, int k) {
    Dual r(c);
    r.d[k] = 1.f;
    return r;
  }
};
#define DUAL_BIN(op, vexpr, dexpr)                                                   \
  template <int N>                                                                   \
  __host__ __device__ inline Dual<N> operator op(const Dual<N>& a, const Dual<N>& b) { \
    Dual<N> r;                                                                       \
    r.v = vexpr;                                                                     \
    _Pragma("unroll") for (int i = 0; i < N; ++i) r.d[i] = dexpr;                    \
    return r;                                                                        \
  }
DUAL_BIN(+, a.v + b.v, a.d[i] + b.d[i])
DUAL_BIN(-, a.v - b.v, a.d[i] - b.d[i])
DUAL_BIN(*, a.v * b.v, a.d[i] * b.v + a.v * b.d[i])
DUAL_BIN(/, a.v / b.v, (a.d[i] * b.v - a.v * b.d[i]) / (b.v * b.v))
#undef DUAL_BIN
template <int N>
__host__ __device__ inline Dual<N> operator-(const Dual<N>& a) {
  Dual<N> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int N>
__host__ __device__ inline Dual<N> dsqrt(const Dual<N>& a) {
  Dual<N> r;
  r.v = sqrtf(a.v);
  float k = r.v > 0.f ? 0.5f / r.v : 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * k;
  return r;
}
template <int N>
__host__ __device__ inline Dual<N> dsin(const Dual<N>& a) {
  Dual<N> r;
  float s, c;
  sincosf(a.v, &s, &c);
  r.v = s;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * c;
  return r;
}
template <int N>
__host__ __device__ inline Dual<N> dcos(const Dual<N>& a) {
  Dual<N> r;
  float s, c;
  sincosf(a.v, &s, &c);
  r.v = c;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i] * s;
  return r;
}
// plain-float overloads so templated chains work for T = float too
__host__ __device__ inline float dsqrt(float a) { return sqrtf(a); }
__host__ __device__ inline float dsin(float a) { return sinf(a); }
__host__ __device__ inline float dcos(float a) { return cosf(a); }
__host__ __device__ inline float dval(float a) { return a; }
template <int N>
__host__ __device__ inline float dval(const Dual<N>& a) {
  return a.v;
}

// manopth/rodrigues_layer.py:15-54: axis-angle -> quaternion -> rotation matrix (row-major R[9]).
template <typename T>
__host__ __device__ inline void rodrigues_quat(const T a[3], T R[9]) {
  T ax = a[0] + T(1e-8f), ay = a[1] + T(1e-8f), az = a[2] + T(1e-8f);
  T angle = dsqrt(ax * ax + ay * ay + az * az);
  T nx = a[0] / angle, ny = a[1] / angle, nz = a[2] / angle;
  T half = angle * T(0.5f);
  T c = dcos(half), s = dsin(half);
  T qw = c, qx = s * nx, qy = s * ny, qz = s * nz;
  T qn = dsqrt(qw * qw + qx * qx + qy * qy + qz * qz);
  T w = qw / qn, x = qx / qn, y = qy / qn, z = qz / qn;
  T w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  T wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  T two(2.f);
  R[0] = w2 + x2 - y2 - z2;
  R[1] = two * xy - two * wz;
  R[2] = two * wy + two * xz;
  R[3] = two * wz + two * xy;
  R[4] = w2 - x2 + y2 - z2;
  R[5] = two * yz - two * wx;
  R[6] = two * xz - two * wy;
  R[7] = two * wx + two * yz;
  R[8] = w2 - x2 - y2 + z2;
}
