// Per-frame camera / light glue of the fused optimiser step, forward and backward, as two launches.
// Replaces, per minibatch: the weak-perspective camera of render_image (utils/visualize.py:266-270), the shadow pass's light
// camera (process_info_for_shadow, renderer/renderer_helper.py:454-468; look_at_rotation SURVEY.md A.10), hand_verts.mean(1),
// sigmoid(amb_ratio) (renderer_helper.py:420-433) and the autograd graph through all of them (about 60 ATen launches).
// The backward uses forward-mode duals over the six inputs (light position, centroid) of the 12 light-camera outputs.
#include "glue.cuh"

namespace {

__global__ void __launch_bounds__(256) frame_glue_fwd_kernel(int V, int S, float focal, const float* __restrict__ cam_table,
                                                             const int64_t* __restrict__ fid, const float* __restrict__ light_pos,
                                                             const float* __restrict__ amb_ratio, const float* __restrict__ verts,
                                                             float* __restrict__ cam_T, float* __restrict__ light_b,
                                                             float* __restrict__ light_R, float* __restrict__ light_T,
                                                             float* __restrict__ amb, float* __restrict__ centroid) {
  __shared__ float red[32];
  int b = blockIdx.x;
  const float* vb = verts + (size_t)b * V * 3;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) { sx += vb[v * 3]; sy += vb[v * 3 + 1]; sz += vb[v * 3 + 2]; }
  sx = block_sum(sx, red); sy = block_sum(sy, red); sz = block_sum(sz, red);
  if (threadIdx.x != 0) return;
  float c[3] = {sx / (float)V, sy / (float)V, sz / (float)V};
  const float* cam = cam_table + (size_t)(fid ? fid[b] : b) * 3;
  cam_T[b * 3 + 0] = -cam[1];
  cam_T[b * 3 + 1] = -cam[2];
  // utils/visualize.py:268 `2 * focal_length / (img_size * cam[:, 0] + 1e-9)`: torch evaluates scalar / tensor as
  // reciprocal(tensor) * scalar (Tensor.__rtruediv__), two roundings — reproduced so that t_z is bit-identical
  cam_T[b * 3 + 2] = (1.f / ((float)S * cam[0] + 1e-9f)) * (2.f * focal);
  float L[3] = {light_pos[0], light_pos[1], light_pos[2]}, R[9], T[3];
  light_camera<float>(L, c, R, T);
  for (int i = 0; i < 9; ++i) light_R[b * 9 + i] = R[i];
  for (int i = 0; i < 3; ++i) { light_T[b * 3 + i] = T[i]; light_b[b * 3 + i] = L[i]; centroid[b * 3 + i] = c[i]; }
  if (b == 0) amb[0] = 1.f / (1.f + expf(-amb_ratio[0]));
}

__global__ void __launch_bounds__(256) frame_glue_bwd_kernel(int V, int S, float focal, const float* __restrict__ cam_table,
                                                             const int64_t* __restrict__ fid, const float* __restrict__ light_pos,
                                                             const float* __restrict__ amb_ratio, const float* __restrict__ centroid,
                                                             const float* __restrict__ d_cam_T, const float* __restrict__ d_light_R,
                                                             const float* __restrict__ d_light_T, const float* __restrict__ d_light_pos,
                                                             const float* __restrict__ d_amb, float* __restrict__ g_cam_table,
                                                             float* __restrict__ g_light_pos, float* __restrict__ g_amb_ratio,
                                                             float* __restrict__ d_verts) {
  __shared__ float gc[3];
  int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int64_t row = fid ? fid[b] : b;
    const float* cam = cam_table + (size_t)row * 3;
    if (g_cam_table) {
      float den = (float)S * cam[0] + 1e-9f;
      atomicAdd(g_cam_table + row * 3 + 0, d_cam_T[b * 3 + 2] * (-(2.f * focal) * (float)S / (den * den)));
      atomicAdd(g_cam_table + row * 3 + 1, -d_cam_T[b * 3 + 0]);
      atomicAdd(g_cam_table + row * 3 + 2, -d_cam_T[b * 3 + 1]);
    }
    Dual<6> L[3], c[3], R[9], T[3];
    for (int i = 0; i < 3; ++i) { L[i] = Dual<6>::var(light_pos[i], i); c[i] = Dual<6>::var(centroid[b * 3 + i], 3 + i); }
    light_camera<Dual<6>>(L, c, R, T);
    float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 9; ++k) {
      float u = d_light_R[b * 9 + k];
      for (int i = 0; i < 6; ++i) g[i] += u * R[k].d[i];
    }
    for (int k = 0; k < 3; ++k) {
      float u = d_light_T[b * 3 + k];
      for (int i = 0; i < 6; ++i) g[i] += u * T[k].d[i];
    }
    if (g_light_pos)
      for (int i = 0; i < 3; ++i) atomicAdd(g_light_pos + i, g[i] + d_light_pos[b * 3 + i]);  // share_light_position: row 0
    for (int i = 0; i < 3; ++i) gc[i] = g[3 + i] / (float)V;
    if (b == 0 && g_amb_ratio) {
      float s = 1.f / (1.f + expf(-amb_ratio[0]));
      atomicAdd(g_amb_ratio, d_amb[0] * s * (1.f - s));
    }
  }
  __syncthreads();
  float gx = gc[0], gy = gc[1], gz = gc[2];
  if (gx == 0.f && gy == 0.f && gz == 0.f) return;
  float* dv = d_verts + (size_t)b * V * 3;
  for (int v = threadIdx.x; v < V; v += blockDim.x) { dv[v * 3] += gx; dv[v * 3 + 1] += gy; dv[v * 3 + 2] += gz; }
}

}  // namespace

extern "C" int harp_frame_glue_fwd(harp_ctx* c, int B, int S, float focal, const float* cam_table, const int64_t* fid,
                                   const float* light_pos, const float* amb_ratio, const float* verts, float* cam_T,
                                   float* light_pos_b, float* light_R, float* light_T, float* amb, float* centroid, void* stream) {
  HARP_REQUIRE(c && B > 0 && cam_table && light_pos && amb_ratio && verts && cam_T && light_pos_b && light_R && light_T && amb &&
                   centroid, HARP_ERR_INVALID, "harp_frame_glue_fwd: null argument");
  frame_glue_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(c->V, S, focal, cam_table, fid, light_pos, amb_ratio, verts, cam_T,
                                                             light_pos_b, light_R, light_T, amb, centroid);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

extern "C" int harp_frame_glue_bwd(harp_ctx* c, int B, int S, float focal, const float* cam_table, const int64_t* fid,
                                   const float* light_pos, const float* amb_ratio, const float* centroid, const float* d_cam_T,
                                   const float* d_light_R, const float* d_light_T, const float* d_light_pos, const float* d_amb,
                                   float* g_cam_table, float* g_light_pos, float* g_amb_ratio, float* d_verts, void* stream) {
  HARP_REQUIRE(c && B > 0 && cam_table && light_pos && amb_ratio && centroid && d_cam_T && d_light_R && d_light_T && d_light_pos &&
                   d_amb && d_verts, HARP_ERR_INVALID, "harp_frame_glue_bwd: null argument");
  frame_glue_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(c->V, S, focal, cam_table, fid, light_pos, amb_ratio, centroid, d_cam_T,
                                                             d_light_R, d_light_T, d_light_pos, d_amb, g_cam_table, g_light_pos,
                                                             g_amb_ratio, d_verts);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}
