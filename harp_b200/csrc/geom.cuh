// Arguments of the fused per-frame geometry kernels (geom.cu), shared with harp_step_fused (fused_step.cu).
#pragma once
#include "common.cuh"

struct GeoArgs {
  int B, S;
  float focal;
  const int64_t* fid;
  const float *pose, *rot, *trans, *cam, *wrist, *shape, *disps, *light_pos, *amb_ratio;  // parameter tables / shared parameters
  // forward outputs
  float *verts0, *joints, *verts, *normals, *vposed;  // vposed: (B,V0,3) blend-shaped rest vertices, kept for the backward
  float *cam_T, *light_b, *light_R, *light_T, *centroid, *amb;  // (B,3) (B,3) (B,9) (B,3) (B,3) (1)
  float* cam_R;     // (B,9) PyTorch3D's view of the reference camera, R = diag(-1, -1, 1) per frame (utils/visualize.py:269-270)
  int zero_grads;   // the forward also zero-fills the "backward inputs" below (the step's per-frame gradient accumulators)
  // backward inputs
  const float *d_verts, *d_normals, *d_joints;
  const float *d_verts4, *d_normals4;  // optional (B,V,4) padded accumulators of the shading backward, added to d_verts / d_normals
  const float *d_cam_T, *d_light_pos, *d_light_R, *d_light_T, *d_amb;  // (B,3) (B,3) (B,9) (B,3) (1)
  // backward outputs (accumulated with atomics; rows of the tables by fid)
  float *g_pose, *g_cam, *g_rot, *g_wrist, *g_shape, *g_disps, *g_light, *g_amb;
  float *d_rot, *d_trans, *d_wrist;  // (B,3) per minibatch frame, overwritten (nullable)
};

int harp_geo_fwd(harp_ctx* c, const GeoArgs& a, cudaStream_t st);
int harp_geo_bwd(harp_ctx* c, const GeoArgs& a, cudaStream_t st);
