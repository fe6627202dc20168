// harp_step_fused: the body of the reference's per-minibatch loop (optimize_sequence.py:446-579) — prepare_mesh, both renders,
// every loss of 517-553 except VGG, and the whole backward down to the parameter-table gradients — as ONE C-ABI call that
// enqueues a fixed launch sequence on the caller's stream (capturable into a CUDA graph):
//
//   geo_fwd                      (geom.cu)   zero fill of the per-frame gradient accumulators; rows of the tables -> LBS -> mesh
//                                            build -> camera / light glue
//   |- regulariser branch (ctx stream): texture regularisers, Laplacian / normal consistency / ARAP / keypoints / sum disp^2
//   harp_render_fused            (pixel.cu)  raster setup, tile rasteriser, light view, deferred shading, their backward
//   geo_bwd                      (geom.cu)   glue, mesh build and LBS backward -> atomics on the table gradients
//
// The exchange of the shared-parameter gradients between GPUs (harp_peer_allreduce) and the optimiser update (harp_adam_dev)
// follow on the same stream.  Loss weights, stage gating and frame scaling: optimize_sequence.py:411-422, 507-515.
#include "geom.cuh"

namespace {

struct Workspace {
  float *verts0, *joints, *verts, *normals, *vposed, *cam_T, *light_b, *light_R, *light_T, *centroid, *amb, *cam_R;
  float* zero_begin;  // everything from here on is zero-filled by geo_fwd at the start of the step
  float *d_verts, *d_normals, *d_joints, *d_cam_T, *d_light_pos, *d_light_R, *d_light_T, *d_amb;
  size_t zero_floats, total_floats;
};

inline size_t pad4(size_t n) { return (n + 3) & ~(size_t)3; }

Workspace carve(const harp_ctx* c, int B, float* base) {
  Workspace w{};
  size_t o = 0;
  auto take = [&](size_t n) { float* p = base ? base + o : nullptr; o += pad4(n); return p; };
  const size_t nv = (size_t)B * c->V * 3, nj = (size_t)B * c->n_joints_out * 3;
  w.verts0 = take((size_t)B * c->V0 * 3); w.vposed = take((size_t)B * c->V0 * 3); w.joints = take(nj); w.verts = take(nv); w.normals = take(nv);
  w.cam_T = take((size_t)B * 3); w.light_b = take((size_t)B * 3); w.light_R = take((size_t)B * 9); w.light_T = take((size_t)B * 3);
  w.centroid = take((size_t)B * 3); w.amb = take(4); w.cam_R = take((size_t)B * 9);
  size_t z0 = o;
  w.zero_begin = base ? base + o : nullptr;
  w.d_verts = take(nv); w.d_normals = take(nv); w.d_joints = take(nj);
  w.d_cam_T = take((size_t)B * 3); w.d_light_pos = take((size_t)B * 3); w.d_light_R = take((size_t)B * 9); w.d_light_T = take((size_t)B * 3);
  w.d_amb = take(4);
  w.zero_floats = o - z0;
  w.total_floats = o;
  return w;
}

}  // namespace

extern "C" size_t harp_step_workspace_floats(const harp_ctx* c, int B) { return c && B > 0 ? carve(c, B, nullptr).total_floats : 0; }

extern "C" int harp_step_fused(harp_ctx* c, const harp_step_args* p, void* stream) {
  HARP_REQUIRE(c && p, HARP_ERR_INVALID, "harp_step_fused: null argument");
  const int B = p->B, S = p->S;
  HARP_REQUIRE(B > 0 && S > 0 && p->stage >= 1 && p->stage <= 3, HARP_ERR_INVALID, "harp_step_fused: bad B / S / stage");
  HARP_REQUIRE(B <= c->maxB && S == c->S, HARP_ERR_CAPACITY, "harp_step_fused: call harp_ctx_reserve(B, S) first");
  HARP_REQUIRE(c->has_model && c->has_uv, HARP_ERR_NO_MODEL, "harp_step_fused: ctx needs a skinning model and UV tables");
  HARP_REQUIRE(p->pose && p->rot && p->trans && p->cam && p->shape && p->texture && p->normal_map && p->light_position && p->amb_ratio,
               HARP_ERR_INVALID, "harp_step_fused: missing parameter tensor");
  HARP_REQUIRE(p->y_true && p->y_sil && p->y_sil_eroded && p->init_joints && p->ref_verts && p->uv_mask && p->off_albedo && p->off_normal,
               HARP_ERR_INVALID, "harp_step_fused: missing frame / static input");
  HARP_REQUIRE(p->losses && p->alpha && p->rgb && p->workspace, HARP_ERR_INVALID, "harp_step_fused: missing output / workspace");
  HARP_REQUIRE(p->g_pose && p->g_cam && p->g_shape && p->g_texture && p->g_normal_map && p->g_light_position && p->g_amb_ratio,
               HARP_ERR_INVALID, "harp_step_fused: missing gradient tensor");
  HARP_REQUIRE(!p->verts_disps || p->g_verts_disps, HARP_ERR_INVALID, "harp_step_fused: g_verts_disps required with verts_disps");
  cudaStream_t st = (cudaStream_t)stream;
  const bool coarse_on = (p->stage & 1) != 0, app_on = (p->stage & 2) != 0;
  const float fs = p->frame_scale > 0.f ? p->frame_scale : 1.f;
  const float* W = p->weights;  // silhouette, laplacian, normal, arap, kps_anchor, vert_disp_reg, photo, albedo, normal_reg
  Workspace w = carve(c, B, p->workspace);
  int rc;
  // (the zero fill of the per-frame gradient accumulators and the constant camera rotations are written by geo_fwd: two graph
  // nodes less at the head of the step)
  GeoArgs g{};
  g.cam_R = w.cam_R; g.zero_grads = 1;
  g.B = B; g.S = S; g.focal = p->focal; g.fid = p->fid;
  g.pose = p->pose; g.rot = p->rot; g.trans = p->trans; g.cam = p->cam; g.wrist = p->wrist_pose; g.shape = p->shape;
  g.disps = p->verts_disps; g.light_pos = p->light_position; g.amb_ratio = p->amb_ratio;
  g.verts0 = w.verts0; g.vposed = w.vposed; g.joints = w.joints; g.verts = w.verts; g.normals = w.normals;
  g.cam_T = w.cam_T; g.light_b = w.light_b; g.light_R = w.light_R; g.light_T = w.light_T; g.centroid = w.centroid; g.amb = w.amb;
  g.d_verts = w.d_verts; g.d_normals = w.d_normals; g.d_joints = w.d_joints;
  g.d_cam_T = w.d_cam_T; g.d_light_pos = w.d_light_pos; g.d_light_R = w.d_light_R; g.d_light_T = w.d_light_T; g.d_amb = w.d_amb;
  g.g_pose = p->g_pose; g.g_cam = p->g_cam; g.g_rot = p->g_rot; g.g_wrist = p->g_wrist_pose; g.g_shape = p->g_shape;
  g.g_disps = p->g_verts_disps; g.g_light = p->g_light_position; g.g_amb = p->g_amb_ratio;
  g.d_rot = p->d_rot; g.d_trans = p->d_trans; g.d_wrist = p->d_wrist;
  if ((rc = harp_geo_fwd(c, g, st)) != HARP_OK) return rc;
  // ---- regulariser branch: depends only on the mesh / the textures; accumulates with atomics
  if (!c->reg_stream) {
    HARP_CHECK_CUDA(cudaStreamCreateWithFlags(&c->reg_stream, cudaStreamNonBlocking));
    HARP_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_reg_fork, cudaEventDisableTiming));
    HARP_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_reg_join, cudaEventDisableTiming));
  }
  cudaStream_t rs = c->reg_stream;
  HARP_CHECK_CUDA(cudaEventRecord(c->ev_reg_fork, st));
  HARP_CHECK_CUDA(cudaStreamWaitEvent(rs, c->ev_reg_fork, 0));
  harp_prof_begin(c, 10, rs);
  if (app_on) {
    if ((rc = harp_texture_reg(c, 0, p->texture, p->texH, p->texW, p->off_albedo, p->uv_mask, W[7] * fs, p->losses + 7, p->g_texture, rs)) != HARP_OK) return rc;
    if ((rc = harp_texture_reg(c, 1, p->normal_map, p->texH, p->texW, p->off_normal, p->uv_mask, W[8] * fs, p->losses + 8, p->g_normal_map, rs)) != HARP_OK) return rc;
  }
  if (coarse_on) {
    const float w5[5] = {W[1] * fs, W[2] * fs, W[3] * fs, W[4] * fs, W[5] * fs};
    if ((rc = harp_mesh_losses(c, B, w.verts, p->ref_verts, w.joints, p->init_joints, p->verts_disps, w5, p->losses + 1, w.d_verts, w.d_joints,
                               p->g_verts_disps, rs)) != HARP_OK) return rc;
  }
  harp_prof_end(c, 10, rs);
  HARP_CHECK_CUDA(cudaEventRecord(c->ev_reg_join, rs));
  // ---- both renders (they run in every stage; their losses are gated by COARSE_OPT / APP_OPT) in one fused pixel pass
  harp_render_args r{};
  r.verts = w.verts; r.normals = w.normals; r.cam_R = w.cam_R; r.cam_T = w.cam_T; r.light_R = w.light_R; r.light_T = w.light_T;
  r.light_pos = w.light_b; r.amb = w.amb; r.texture = p->texture; r.normal_map = p->normal_map; r.texH = p->texH; r.texW = p->texW;
  r.self_shadow = p->self_shadow; r.faces_per_pixel = p->faces_per_pixel; r.sigma = p->sigma;
  r.y_sil = p->y_sil; r.y_true = p->y_true; r.y_mask = p->y_sil_eroded;
  r.w_sil = coarse_on ? W[0] * fs : 0.f; r.w_photo = app_on ? W[6] * fs : 0.f;
  r.alpha = p->alpha; r.rgb = p->rgb; r.loss_sil = p->losses + 0; r.loss_photo = p->losses + 6;
  r.grads.d_verts = w.d_verts; r.grads.d_normals = w.d_normals; r.grads.d_cam_R = nullptr; r.grads.d_cam_T = w.d_cam_T;
  r.grads.d_light_R = w.d_light_R; r.grads.d_light_T = w.d_light_T; r.grads.d_light_pos = w.d_light_pos; r.grads.d_amb = w.d_amb;
  r.grads.d_texture = p->g_texture; r.grads.d_normal_map = p->g_normal_map;
  c->defer_fold = true;  // the padded per-vertex accumulators of the shading backward are folded inside geo_bwd
  rc = harp_render_fused(c, B, S, p->focal, &r, st);
  c->defer_fold = false;
  if (rc != HARP_OK) return rc;
  if (app_on) { g.d_verts4 = c->d_verts4; g.d_normals4 = c->d_normals4; }
  HARP_CHECK_CUDA(cudaStreamWaitEvent(st, c->ev_reg_join, 0));
  return harp_geo_bwd(c, g, st);
}

// Device pointers of the step's intermediate results inside a workspace (for callers that want the mesh or the joints).
extern "C" int harp_step_workspace_views(const harp_ctx* c, int B, float* workspace, float** verts0, float** joints, float** verts,
                                         float** normals) {
  HARP_REQUIRE(c && B > 0 && workspace, HARP_ERR_INVALID, "harp_step_workspace_views: bad argument");
  Workspace w = carve(c, B, workspace);
  if (verts0) *verts0 = w.verts0;
  if (joints) *joints = w.joints;
  if (verts) *verts = w.verts;
  if (normals) *normals = w.normals;
  return HARP_OK;
}
