/* harp_b200.h — C ABI of libharpb200.so: the B200-native (sm_100a) replacement for the PyTorch3D / manopth hot path of
 * korrawe/harp (per-minibatch body of optimize_sequence.py:446-579).
 *
 * The reference has no FFI of its own: its seam is a set of Python call signatures.  Each entry
 * point below names the reference call it replaces; the Python modules under `harp_b200/` bind them with ctypes and re-exposes the
 * reference's Python signatures (INTEGRATION.md shows the binding a maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (enum below); harp_last_error() gives the message (thread local)
 *   - all tensor arguments are DEVICE pointers to dense row-major fp32 / int32 unless marked [host]
 *   - the caller (PyTorch) owns every tensor; the library owns only the opaque harp_ctx (static topology, model
 *     constants, scratch sized by harp_ctx_reserve); no allocation happens inside hot calls
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it; one ctx per device+stream
 *   - gradient outputs named d_* are ACCUMULATED INTO (+=) unless the comment says "overwritten"
 *   - units: metres (the reference's mm scaling at manolayer.py:291-292 is undone at visualize.py:45-46)
 */
#ifndef HARP_B200_H
#define HARP_B200_H

#if defined(__GNUC__)
#define HARP_API __attribute__((visibility("default")))
#else
#define HARP_API
#endif

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct harp_ctx harp_ctx;

enum {
  HARP_OK = 0,
  HARP_ERR_INVALID = -1,   /* bad argument / shape */
  HARP_ERR_CUDA = -2,      /* CUDA runtime error (message has the string) */
  HARP_ERR_CAPACITY = -3,  /* harp_ctx_reserve too small for this call */
  HARP_ERR_NO_MODEL = -4   /* LBS called on a ctx created without a MANO model */
};

/* Static mesh topology + UV tables.  [host] pointers, copied at create time.
 * faces0: coarse triangle list (manolayer th_faces, manolayer.py:82-83).  If `subdivide` != 0 the ctx applies one
 * PyTorch3D SubdivideMeshes step (optimize_sequence.py:67-89): V = V0 + E0, F = 4*F0, face order f0|f1|f2|f3.
 * faces_uvs (F,3) / verts_uvs (Vt,2): TexturesUV tables (utils/hand_model_utils.py:59-61); may be NULL (no shading). */
typedef struct {
  int32_t V0, F0;
  const int32_t* faces0;
  int32_t subdivide;
  int32_t Vt;
  const int32_t* faces_uvs;
  const float* verts_uvs;
} harp_topology;

/* MANO constants, names as registered at manopth/manolayer.py:70-97.  [host] pointers, fp32, copied. */
typedef struct {
  const float* shapedirs;   /* (V0,3,10) */
  const float* posedirs;    /* (V0,3,135) */
  const float* v_template;  /* (V0,3) metres */
  const float* J_regressor; /* (16,V0) */
  const float* weights;     /* (V0,16) */
  const float* hands_mean;  /* (45) */
  const int32_t* tips;      /* (5) fingertip vertex ids, manolayer.py:272 */
} harp_mano_model;

/* Generic linear-blend-skinning model of the SMPL family restricted to the vertices the caller keeps (the 1026 arm verts
 * of SMPLXARM, hand_models_harp/body_models.py:2100-2127,2383).  [host] pointers, fp32 / int32, copied.
 * The joint regressor is passed FOLDED: J(shape) = J_template + J_shapedirs . shape  (= J_regressor (v_template + S shape)).
 * Output joints: joint_out_src[k] < J selects chain joint k's posed location, J + i selects vertex tips[i] (the
 * VertexJointSelector fingertips).  center_joint >= 0 subtracts that joint from every output (body_models.py:2342-2343). */
#define HARP_MAX_JOINTS 64
#define HARP_MAX_SHAPE 32
typedef struct {
  int32_t J, n_shape;            /* V0 comes from the topology */
  const int32_t* parents;        /* (J) parents[0] = -1, parents precede children */
  const float* shapedirs;        /* (V0,3,n_shape) */
  const float* posedirs;         /* (V0,3,(J-1)*9) */
  const float* v_template;       /* (V0,3) metres */
  const float* J_template;       /* (J,3) */
  const float* J_shapedirs;      /* (J,3,n_shape) */
  const float* weights;          /* (V0,J) */
  const float* pose_mean;        /* (3J) added to the axis-angle pose */
  int32_t rodrigues_kind;        /* 0: manopth quaternion form (rodrigues_layer.py:15-54); 1: smplx.lbs.batch_rodrigues */
  int32_t center_joint;          /* -1: none */
  int32_t n_joints_out;
  const int32_t* joint_out_src;  /* (n_joints_out) */
  int32_t n_tips;
  const int32_t* tips;           /* (n_tips) vertex ids in [0,V0) */
} harp_lbs_model;

HARP_API const char* harp_last_error(void);
HARP_API int harp_version(void);

HARP_API int harp_ctx_create(harp_ctx** out, const harp_topology* topo, const harp_mano_model* model /* nullable */);
/* Same, with a generic skinning model (SMPL-X arm).  harp_ctx_create(topo, mano) is the MANO special case. */
HARP_API int harp_ctx_create_lbs(harp_ctx** out, const harp_topology* topo, const harp_lbs_model* model);
HARP_API void harp_ctx_destroy(harp_ctx* ctx);
/* Size all scratch (bin lists, projected verts, light depth, per-frame Jacobians) for <= max_batch frames of
 * img_size x img_size.  Allocates; call outside the hot loop. */
HARP_API int harp_ctx_reserve(harp_ctx* ctx, int max_batch, int img_size);
/* Kernels receive the ctx by value, so a captured CUDA graph holds the scratch pointers.  A caller that captures calls on this
 * ctx pins it (delta = +1) for the graph's lifetime and unpins it (-1) when the graph is destroyed; while pinned,
 * harp_ctx_reserve refuses (HARP_ERR_CAPACITY) any request that would move the scratch instead of freeing memory a graph still
 * replays into.  The batch capacity never shrinks.  A ctx must not be used from two streams concurrently. */
HARP_API int harp_ctx_pin(harp_ctx* ctx, int delta);
HARP_API int harp_ctx_dims(const harp_ctx* ctx, int32_t* V0, int32_t* V, int32_t* F, int32_t* E);
/* Copy the (F,3) int32 (subdivided) face list to [host] memory, for callers that need Meshes.faces_padded(). */
HARP_API int harp_ctx_faces(const harp_ctx* ctx, int32_t* faces_out);
/* counters written by the kernels since the last call: [0] silhouette pixels whose top-K truncation could not be
 * evaluated exactly (more than 512 passing faces on one pixel, or more than 4M such pixels), [1] bin-list capacity
 * overflows.  Both 0 means every result follows PyTorch3D's semantics.  [host] out. Synchronises. */
HARP_API int harp_ctx_read_counters(harp_ctx* ctx, int64_t out[2], void* stream);
/* Pixels the last silhouette pass on this ctx handed to the exact per-pixel top-K kernel (a pixel with more candidate faces than
 * faces_per_pixel, or a tile whose 256-entry band queue overflowed).  Results are exact either way; the number tells how far the
 * content is from the fast path.  [host] out.  Synchronises. */
HARP_API int harp_ctx_deferred_pixels(harp_ctx* ctx, int64_t* out, void* stream);

/* Number of CUDA kernels this library has launched in this process (all contexts), for launch accounting. */
HARP_API int64_t harp_launch_count(void);
/* Per-kernel device timing.  enable = 1: the two pixel kernels (id 0 = silhouette / pixel_vis, 1 = shade) are bracketed by
 * CUDA events on the launch stream (inside a stream capture: external event-record nodes, re-recorded on every replay);
 * enable = 2 also brackets the step's segments: 2 LBS forward, 3 LBS backward, 4 mesh build forward, 5 mesh build backward,
 * 6 camera projection + binning, 7 the render's backward tail.  harp_ctx_profile_read synchronises, returns the summed
 * elapsed ms and launch count of one id (0..7) since the last read, and clears (events of a captured graph are kept). */
HARP_API int harp_ctx_profile(harp_ctx* ctx, int enable);
HARP_API int harp_ctx_profile_read(harp_ctx* ctx, int kernel_id, float* total_ms, int* launches);
/* Every profiling event in recording order: out[2k] = id, out[2k+1] = ms since the first event (even occurrences of an id open
 * its segment, odd ones close it).  Besides ids 0..7: 8 step prologue, 9 the render's side branch (light view, empty tiles),
 * 10 the regulariser branch.  Shows the gaps and the overlap that the per-id sums hide.  capacity in floats.  Outside a captured graph the
 * call consumes the events. */
HARP_API int harp_ctx_profile_timeline(harp_ctx* ctx, float* out, int capacity, int* n);

/* ---- MANO linear blend skinning: ManoLayer.forward (manopth/manolayer.py:108-296), use_pca=False, axis-angle -------
 * pose48 (B,48) = [rot3 | pose45]; betas (B,10) or one shared row if betas_shared; trans (B,3).
 * verts0 (B,V0,3) m; joints (B,21,3) m. */
HARP_API int harp_lbs_mano_fwd(harp_ctx*, int B, const float* pose48, const float* betas, int betas_shared, const float* trans,
                      float* verts0, float* joints, void* stream);
/* d_pose48 (B,48), d_betas (B,10) [per frame, caller sums for a shared shape], d_trans (B,3): overwritten. */
HARP_API int harp_lbs_mano_bwd(harp_ctx*, int B, const float* pose48, const float* betas, int betas_shared, const float* trans,
                      const float* d_verts0, const float* d_joints, float* d_pose48, float* d_betas, float* d_trans,
                      void* stream);

/* ---- generic LBS (SMPLXARM.forward -> smplx.lbs.lbs, hand_models_harp/body_models.py:2163-2390) ----------------------
 * pose (B,3J) full axis-angle pose (pose_mean is added inside); shape (B,n_shape) or one shared row; trans (B,3).
 * verts0 (B,V0,3) m, joints (B,n_joints_out,3) m.  bwd: d_pose (B,3J), d_shape (B,n_shape), d_trans (B,3) overwritten. */
HARP_API int harp_lbs_fwd(harp_ctx*, int B, const float* pose, const float* shape, int shape_shared, const float* trans,
                          float* verts0, float* joints, void* stream);
HARP_API int harp_lbs_bwd(harp_ctx*, int B, const float* pose, const float* shape, int shape_shared, const float* trans,
                          const float* d_verts0, const float* d_joints, float* d_pose, float* d_shape, float* d_trans,
                          void* stream);

/* ---- mesh build: SubdivideMeshes + normal displacement + vertex normals (utils/visualize.py:48-64) -------------------
 * verts0 (B,V0,3); disps (V) or NULL.  verts (B,V,3), normals (B,V,3) = Meshes.verts_normals_padded() of `verts`. */
HARP_API int harp_mesh_build_fwd(harp_ctx*, int B, const float* verts0, const float* disps, float* verts, float* normals,
                        void* stream);
/* d_verts / d_normals (B,V,3) in (d_normals may be NULL).  d_verts0 (B,V0,3) overwritten; d_disps (V) accumulated. */
HARP_API int harp_mesh_build_bwd(harp_ctx*, int B, const float* verts0, const float* disps, const float* d_verts,
                        const float* d_normals, float* d_verts0, float* d_disps, void* stream);

/* Meshes.verts_normals_padded() of arbitrary verts on the ctx topology (area-weighted, F.normalize eps 1e-6). */
HARP_API int harp_vertex_normals_fwd(harp_ctx*, int B, const float* verts, float* normals, void* stream);
/* d_verts (B,V,3) accumulated. */
HARP_API int harp_vertex_normals_bwd(harp_ctx*, int B, const float* verts, const float* d_normals, float* d_verts, void* stream);

/* ---- mesh regularisers + keypoint anchor, value and gradient in one pass --------------------------------------------
 * mesh_laplacian_smoothing / mesh_normal_consistency (optimize_sequence.py:536-537), arap_loss (loss/arap.py:4-57),
 * kps_loss (loss/kps_loss.py:4-17), sum(disp^2) (optimize_sequence.py:533).
 * w[5] = upstream weights for {laplacian, normal, arap, kps, disp}; losses[5] (device) receives the UNWEIGHTED values
 * (accumulated, zero it first); gradients are weighted; a term with weight 0 still reports its value, a term with a NEGATIVE
 * weight (or without its inputs) is not evaluated.  ref_verts (V,3); joints (B,n_joints_out,3) m of the ctx's model
 * (21 for MANO / no model, 22 for the arm: the first 21 are used); init_joints (B,21,3) mm. */
HARP_API int harp_mesh_losses(harp_ctx*, int B, const float* verts, const float* ref_verts, const float* joints,
                     const float* init_joints_mm, const float* disps, const float w[5], float* losses,
                     float* d_verts, float* d_joints, float* d_disps, void* stream);

/* ---- cameras --------------------------------------------------------------------------------------------------------
 * Per-frame world->view transform in PyTorch3D's row-vector convention X_view = X_world R + T.  R (B,3,3), T (B,3). */

/* ---- soft silhouette: MeshRasterizer(K=50, blur=ln(1/1e-4-1)*sigma) + SoftSilhouetteShader ---------------------------
 * (renderer/renderer_helper.py:37-58, utils/visualize.py:258-285).  alpha (B,S,S).
 * mode 0: forward only.
 * mode 1: backward with upstream gradient g_alpha (B,S,S): d_verts (B,V,3), d_R (B,9), d_T (B,3) accumulated.
 * mode 2: fused L1 loss: target (B,S,S); loss (device scalar, accumulated, unweighted mean|target-alpha|);
 *         gradients of weight*loss accumulated.
 * faces_per_pixel = K of RasterizationSettings (50 in HARP; <= 0 means 50).  Pixels with more candidate faces than K are
 * re-evaluated by an exact per-pixel top-K-by-depth kernel, so results follow PyTorch3D's fragment truncation. */
HARP_API int harp_silhouette(harp_ctx*, int mode, int B, int S, float focal, float sigma, int faces_per_pixel, const float* verts, const float* R,
                    const float* T, float* alpha, const float* g_alpha_or_target, float weight, float* loss,
                    float* d_verts, float* d_R, float* d_T, void* stream);

/* ---- shadowed PBR colour render: MeshRendererShadow + SoftPhongShaderShadow + PBRMaterials.apply_normal_map ----------
 * (renderer/renderer_helper.py:331-412,472-592; renderer/pbr_materials.py:58-124).
 * verts, normals (B,V,3); cam R,T and light R,T (B,3,3)/(B,3) (process_info_for_shadow, renderer_helper.py:454-468);
 * light_pos (B,3); amb (1) = sigmoid(amb_ratio) already applied; texture, normal_map (H,W,3) shared maps, normal_map RAW
 * (normalised per texel inside, visualize.py:98).  self_shadow=0 renders with vis=1.  rgb (B,S,S,3).
 * fixed_lights = {ambient, diffuse, specular} selects the non-shadow Phong twin (SoftPhongShaderPBR, renderer_helper.py:
 * 60-83,106-190: lights .5/.4/.1, shininess 0 => constant specular); NULL = ambient amb, diffuse 1-amb, no specular.
 * mode as above; for mode 1 g = g_rgb (B,S,S,3); for mode 2 g = y_true (B,S,S,3) and `mask` = eroded mask (B,S,S):
 * loss = mean|y*m - rgb*m| (optimize_sequence.py:543). */
typedef struct {
  float *d_verts, *d_normals;        /* (B,V,3) */
  float *d_cam_R, *d_cam_T;          /* (B,9),(B,3) */
  float *d_light_R, *d_light_T;      /* (B,9),(B,3) */
  float *d_light_pos;                /* (B,3) */
  float *d_amb;                      /* (1) w.r.t. the post-sigmoid ambient */
  float *d_texture, *d_normal_map;   /* (H,W,3) */
} harp_shade_grads;

HARP_API int harp_shade(harp_ctx*, int mode, int B, int S, float focal, int self_shadow, const float* verts,
               const float* normals, const float* cam_R, const float* cam_T, const float* light_R, const float* light_T,
               const float* light_pos, const float* amb, const float* texture, const float* normal_map, int texH,
               int texW, const float* fixed_lights /* [host] nullable */, float* rgb, const float* g_or_target,
               const float* mask, float weight, float* loss, const harp_shade_grads* grads, void* stream);

/* Coverage of the last harp_shade / harp_render_fused pass on this ctx: mask (B,S,S) = 1 where a face covers the pixel centre,
 * else 0 — the alpha channel of the reference's K=1 blends (renderer_helper.py:190,592: softmax_rgb_blend at K=1), which the
 * reference's callers use as the foreground mask (utils/visualize.py:298). */
HARP_API int harp_hit_mask(harp_ctx*, int B, int S, float* mask, void* stream);

/* ---- evaluation-pass normal renderer: MeshRasterizer(K=10) + SoftPhongNormalShader (renderer_helper.py:85-101,216-301):
 * colours = (n * (1,-1,-1) + 1) / 2 of the normal-mapped interpolated normals, softmax_rgb_blend over the K nearest
 * faces.  Forward only.  normal_map NULL skips the perturbation.  rgb (B,S,S,3). */
HARP_API int harp_render_normals(harp_ctx*, int B, int S, float focal, int faces_per_pixel, const float* verts,
                                 const float* normals, const float* R, const float* T, const float* normal_map, int texH,
                                 int texW, float* rgb, void* stream);

/* ---- fused pixel pass of the optimiser step --------------------------------------------------------------------------
 * harp_silhouette(mode 2) + harp_shade(mode 2) in one launch sequence that walks every tile's face list ONCE (silhouette
 * factors and nearest-face visibility together), compacts hit pixels per frame and shades them with full warps.
 * w_sil / w_photo: loss weights; 0 renders that image forward-only (the reference renders both in every stage,
 * optimize_sequence.py:469-488).  Losses (device scalars) receive the unweighted means; gradients are accumulated (with
 * atomics: other work of the step may add to the same buffers concurrently).  What only the shading needs — the light-view
 * pass, the packed texel copies, the zero fills of the internal accumulators — runs on a stream owned by the ctx, forked from
 * and joined to `stream` with events, so the call is still one in-order unit of work on `stream` and can be captured. */
typedef struct {
  const float *verts, *normals;                       /* (B,V,3) */
  const float *cam_R, *cam_T, *light_R, *light_T;     /* (B,3,3), (B,3) */
  const float *light_pos, *amb;                       /* (B,3), (1) post-sigmoid */
  const float *texture, *normal_map;                  /* (H,W,3) shared maps, normal_map raw */
  int32_t texH, texW, self_shadow;
  int32_t faces_per_pixel;                            /* silhouette K (50; <= 0 means 50) */
  float sigma;                                        /* silhouette blend sigma (1e-7) */
  const float *y_sil, *y_true, *y_mask;               /* targets: (B,S,S), (B,S,S,3), eroded mask (B,S,S) */
  float w_sil, w_photo;
  float *alpha, *rgb;                                 /* (B,S,S), (B,S,S,3) */
  float *loss_sil, *loss_photo;
  harp_shade_grads grads;
} harp_render_args;

HARP_API int harp_render_fused(harp_ctx*, int B, int S, float focal, const harp_render_args* args, void* stream);

/* ---- per-frame camera / light glue of the fused step -----------------------------------------------------------------------
 * Forward: cam_T[b] = (-tx, -ty, 2 f / (S s + 1e-9)) from the weak-perspective row cam_table[fid[b]] = (s, tx, ty)
 * (utils/visualize.py:266-270; fid NULL = identity); centroid[b] = mean_v verts[b] (renderer_helper.py:455 hand_verts.mean(1));
 * the shadow pass's light camera on the radius-1.5 sphere about the centroid looking at it, light_R (B,3,3) / light_T (B,3)
 * (process_info_for_shadow, renderer_helper.py:454-468, with PyTorch3D's look_at_rotation); light_pos_b (B,3) = the shared
 * light position repeated (optimize_sequence.py:453-454); amb (1) = sigmoid(amb_ratio) (renderer_helper.py:420-433).
 * Backward: given d_cam_T, d_light_R, d_light_T, d_light_pos (B,..) and d_amb (1) as harp_render_fused accumulates them,
 * ADDS the gradients to g_cam_table rows fid[b], g_light_pos (3), g_amb_ratio (1) and the centroid's share to d_verts (B,V,3).
 * Replaces the ~60 ATen launches of these tiny chains and their autograd graph with two launches. */
HARP_API int harp_frame_glue_fwd(harp_ctx*, int B, int S, float focal, const float* cam_table, const int64_t* fid,
                        const float* light_pos, const float* amb_ratio, const float* verts, float* cam_T, float* light_pos_b,
                        float* light_R, float* light_T, float* amb, float* centroid, void* stream);
HARP_API int harp_frame_glue_bwd(harp_ctx*, int B, int S, float focal, const float* cam_table, const int64_t* fid,
                        const float* light_pos, const float* amb_ratio, const float* centroid, const float* d_cam_T,
                        const float* d_light_R, const float* d_light_T, const float* d_light_pos, const float* d_amb,
                        float* g_cam_table, float* g_light_pos, float* g_amb_ratio, float* d_verts, void* stream);

/* ---- texture regularisers: albedo_reg / normal_reg (loss/texture_reg.py:5-66) ----------------------------------------
 * offsets (H,W,2) int32 = the reference's CPU-RNG neighbour offsets, passed in.  kind 0: albedo_reg; 1: normal_reg
 * (adds 0.2*close_to_z with the reference's dim=2 norm).  loss accumulated (unweighted); d_map accumulated. */
HARP_API int harp_texture_reg(harp_ctx*, int kind, const float* map, int H, int W, const int32_t* offsets, const float* uv_mask,
                     float weight, float* loss, float* d_map, void* stream);

/* ---- vertex MSE of the upstream per-frame MANO fit (metro_modifications/hand_utils.py:16-133, optimize_for_mano_param) -----
 * loss += mean_i (scale x_i - target_i)^2 over n elements (torch.nn.MSELoss, 64/74/93); d_x_i = 2 scale (scale x_i - target_i) / n
 * (overwritten; may be NULL).  With x = harp_lbs_mano_fwd's verts in metres, scale = 1000 and target in millimetres this is
 * the fit's criterion on ManoLayer's mm output. */
HARP_API int harp_mse(int64_t n, const float* x, const float* target, float scale, float* loss, float* d_x, void* stream);

/* ---- evaluation-pass image metrics: l1_diff and sil_iou of utils/eval_util.py:34-49 ------------------------------------------
 * ref_image / pred_image (B,S,S,3) in [0,1], ref_mask / pred_mask (B,S,S).  ACCUMULATES l1_sum += sum |ref - pred| (the caller
 * divides by 3 B S^2 for torch.mean, 38-39) and, per frame, inter[b] / uni[b] += the pixel counts of (ref >= 0.5) AND / OR
 * (pred >= 0.5) (42-47); the caller zeroes the three outputs first and forms mean_b inter / uni (48-49).
 * (LPIPS and MS-SSIM, 51-60, need pretrained networks: out of scope like the VGG term.) */
HARP_API int harp_image_metrics(int B, int S, const float* ref_image, const float* pred_image, const float* ref_mask,
                       const float* pred_mask, float* l1_sum, int32_t* inter, int32_t* uni, void* stream);

/* ---- uint8 frame ingest: load_img of utils/data_util.py:11-30 on the device ---------------------------------------------
 * rgb_u8 (B,S,S,3), mask_u8 (B,S,S) as decoded from the JPEG / PNG files -> y_true = rgb/255 (B,S,S,3) fp32,
 * y_sil = mask/255 (B,S,S), y_sil_eroded = cv2.erode(mask/255, ones(3,3), iterations=2) (border pixels ignore the outside).
 * Moves 4 bytes per pixel over PCIe instead of the 20 the reference's fp32 DataLoader tensors need. */
HARP_API int harp_ingest_u8(int B, int S, const uint8_t* rgb_u8, const uint8_t* mask_u8, float* y_true, float* y_sil,
                            float* y_sil_eroded, void* stream);

/* ---- Adam (torch.optim.Adam defaults, optimize_sequence.py:253-310,567-573), fused over a flat buffer ----------------*/
HARP_API int harp_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
              float beta2, float eps, int step, void* stream);

/* Same update with the step counter on the device, so the launch can be captured in a CUDA graph and replayed:
 * state = {int32 step, uint32 scratch, uint32 scratch[, float lr]} (3 or 4 x 4 bytes, zero-initialised by the caller); each call
 * updates with the bias corrections of step + 1 and leaves state.step incremented (ONE launch: the last CTA to finish advances
 * the counter).  lr < 0: the learning rate is read from state[3] on the device, so a
 * scheduler (ReduceLROnPlateau, optimize_sequence.py:312-321) can change it without re-capturing the graph. */
HARP_API int harp_adam_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                           float beta2, float eps, void* state, void* stream);

/* ---- the whole minibatch step behind one call (optimize_sequence.py:446-579) ----------------------------
 * prepare_mesh, both renders, every loss of 517-553 except VGG and the complete backward, enqueued on `stream` as a fixed
 * launch sequence (capturable).  Parameters are the reference's tensors (optimize_sequence.py:181-250) as device tables; the
 * frame rows a minibatch uses are selected by `fid` INSIDE the kernels, and their gradients are accumulated (+=, atomics) into
 * the same rows of the g_* tables.  `weights` / `losses` order: silhouette, laplacian, normal, arap, kps_anchor, vert_disp_reg,
 * photo, albedo, normal_reg (unweighted loss values are ACCUMULATED into `losses`: zero it per step).
 * stage: 1 = COARSE_OPT, 2 = APP_OPT, 3 = both (optimize_sequence.py:507-515).  frame_scale = 1 / world_size under frame
 * sharding (frame losses are means over the global batch; frame-independent terms are split over ranks).
 * MANO ctx (J = 16): pose (T,45), rot (T,3); SMPL-X-arm ctx (J = 55): also wrist_pose (T,3) (utils/visualize.py:36-44). */
typedef struct {
  int32_t B, S, stage, self_shadow, faces_per_pixel, texH, texW;
  float focal, sigma, frame_scale;
  float weights[9];
  const int64_t* fid;                                   /* (B) rows of the per-frame tables */
  const float *pose, *rot, *trans, *cam, *wrist_pose;   /* (T,45) (T,3) (T,3) (T,3) (T,3 | NULL) */
  const float *shape, *verts_disps;                     /* (10), (V | NULL) */
  const float *texture, *normal_map;                    /* (texH,texW,3) */
  const float *light_position, *amb_ratio;              /* (3) shared light (row 0), () pre-sigmoid ambient */
  const float *y_true, *y_sil, *y_sil_eroded;           /* minibatch frames: (B,S,S,3) (B,S,S) (B,S,S) */
  const float *init_joints, *ref_verts, *uv_mask;       /* (B,21,3) mm, (V,3) ARAP reference, (texH,texW) */
  const int32_t *off_albedo, *off_normal;               /* (texH,texW,2) neighbour offsets of the texture regularisers */
  float *losses, *alpha, *rgb;                          /* (9) accumulated; (B,S,S); (B,S,S,3) */
  float *g_pose, *g_cam, *g_rot, *g_wrist_pose;         /* table gradients, rows by fid; g_rot / g_wrist_pose nullable */
  float *g_shape, *g_verts_disps, *g_texture, *g_normal_map, *g_light_position, *g_amb_ratio;
  float *d_rot, *d_trans, *d_wrist;                     /* (B,3) per minibatch frame, overwritten; nullable */
  float* workspace;                                     /* harp_step_workspace_floats(ctx, B) floats, caller-allocated */
} harp_step_args;
HARP_API size_t harp_step_workspace_floats(const harp_ctx* ctx, int B);
HARP_API int harp_step_fused(harp_ctx* ctx, const harp_step_args* args, void* stream);
/* where the step left its intermediate results inside `workspace`: (B,V0,3) (B,n_joints,3) (B,V,3) (B,V,3) */
HARP_API int harp_step_workspace_views(const harp_ctx* ctx, int B, float* workspace, float** verts0, float** joints, float** verts,
                                       float** normals);

/* ---- data-parallel exchange over NVLink peer memory (the reference is single-process) -------------------
 * One process per GPU of ONE node.  Every rank's gradient buffer lives in an IPC-shareable allocation that all ranks map; the
 * sum of the shared-parameter gradients is a reduce-scatter + all-gather kernel written against the peers' buffers and
 * synchronised by epoch flags in the same allocations: an ordinary, CUDA-graph-capturable kernel on the caller's stream.
 *   harp_peer_create   allocates `data_bytes` (+ flags) on the current device; returns the device pointer of the data region
 *                      (the caller builds its gradient tensors on it) and this rank's 64-byte IPC handle [host]
 *   harp_peer_connect  `handles`: world x 64 bytes [host], rank-major, gathered by the host layer (torch.distributed)
 *   harp_peer_allreduce  data[start:stop) (float elements) <- sum over ranks, in place on every rank; elements of
 *                      [scale_lo, scale_hi) are then multiplied by `scale`.  Every rank launches it once per step. */
typedef struct harp_peer harp_peer;
HARP_API int harp_peer_create(harp_peer** out, int world, int rank, size_t data_bytes, void** data_ptr, void* ipc_handle_out);
HARP_API int harp_peer_connect(harp_peer* peer, const void* handles);
HARP_API int harp_peer_allreduce(harp_peer* peer, int64_t start, int64_t stop, int64_t scale_lo, int64_t scale_hi, float scale,
                                 void* stream);
HARP_API void harp_peer_destroy(harp_peer* peer);

#ifdef __cplusplus
}
#endif
#endif /* HARP_B200_H */
