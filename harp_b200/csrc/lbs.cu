// Linear blend skinning of a kinematic-tree model (MANO, SMPL-X arm sub-mesh), forward and analytic backward:
// one CTA per frame.
// Replaces ManoLayer.forward (manopth/manolayer.py:108-296) + batch_rodrigues (rodrigues_layer.py:43-54), and
// SMPLXARM.forward (hand_models_harp/body_models.py:2163-2390) -> smplx.lbs.lbs, with their autograd graphs
// (~40 ATen launches, 2 host syncs for MANO; a 10 475-vertex skin of which 90 % is discarded for the arm) by one launch
// each way over exactly the vertices the caller keeps.  FP32 FFMA throughout: the blendshape contractions are
// [(J-1)*9 x 3V] per frame (135 x 2334 for MANO, 486 x 3078 for the arm) — far too small for tcgen05 tiles and needed
// at full fp32 accuracy (last bullet).
#include "lbs.cuh"

namespace {

__global__ void __launch_bounds__(LBS_THREADS) lbs_fwd_kernel(harp_ctx c, const float* __restrict__ pose,
                                                                const float* __restrict__ shape, int shape_shared,
                                                                const float* __restrict__ trans, float* __restrict__ verts0,
                                                                float* __restrict__ joints) {
  __shared__ LbsShared s;
  int b = blockIdx.x, J = c.J;
  lbs_prologue(c, pose + (size_t)b * J * 3, shape + (shape_shared ? 0 : (size_t)b * c.n_shape), s);
  lbs_forward_core(c, s, ld3(trans + (size_t)b * 3), verts0 + (size_t)b * c.V0 * 3, nullptr, joints + (size_t)b * c.n_joints_out * 3);
}

__global__ void __launch_bounds__(LBS_THREADS) lbs_bwd_kernel(harp_ctx c, const float* __restrict__ pose,
                                                                const float* __restrict__ shape, int shape_shared,
                                                                const float* __restrict__ d_verts0,
                                                                const float* __restrict__ d_joints, float* __restrict__ d_pose,
                                                                float* __restrict__ d_shape, float* __restrict__ d_trans) {
  extern __shared__ float dyn[];
  LbsShared& s = *reinterpret_cast<LbsShared*>(dyn);
  LbsBwdShared& g = *reinterpret_cast<LbsBwdShared*>(dyn + (sizeof(LbsShared) + 3) / 4);
  float* dvp = dyn + (sizeof(LbsShared) + 3) / 4 + (sizeof(LbsBwdShared) + 3) / 4;  // (V0*3) gradient w.r.t. v_posed
  float* dTs = dvp + c.V0 * 3;      // (12, V0) per-vertex dT = g (x) [v_posed;1], k-major (bank-conflict free)
  int b = blockIdx.x, J = c.J, NS = c.n_shape, NJO = c.n_joints_out;
  lbs_prologue(c, pose + (size_t)b * J * 3, shape + (shape_shared ? 0 : (size_t)b * NS), s);
  lbs_backward_core(c, s, g, dvp, dTs, d_verts0 + (size_t)b * c.V0 * 3, d_joints ? d_joints + (size_t)b * NJO * 3 : nullptr,
                    d_pose + (size_t)b * J * 3, d_shape + (size_t)b * NS, d_trans ? d_trans + (size_t)b * 3 : nullptr);
}

size_t lbs_bwd_smem(const harp_ctx* c) {
  return ((sizeof(LbsShared) + 3) / 4 + (sizeof(LbsBwdShared) + 3) / 4 + lbs_bwd_scratch_floats(c)) * sizeof(float);
}

}  // namespace

extern "C" int harp_lbs_fwd(harp_ctx* c, int B, const float* pose, const float* shape, int shape_shared, const float* trans,
                            float* verts0, float* joints, void* stream) {
  HARP_REQUIRE(c && B > 0 && pose && shape && trans && verts0 && joints, HARP_ERR_INVALID, "harp_lbs_fwd: null argument");
  HARP_REQUIRE(c->has_model, HARP_ERR_NO_MODEL, "harp_lbs_fwd: ctx has no skinning model");
  harp_prof_begin(c, 2, (cudaStream_t)stream);
  lbs_fwd_kernel<<<B, LBS_THREADS, 0, (cudaStream_t)stream>>>(*c, pose, shape, shape_shared, trans, verts0, joints);
  harp_prof_end(c, 2, (cudaStream_t)stream);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

extern "C" int harp_lbs_bwd(harp_ctx* c, int B, const float* pose, const float* shape, int shape_shared, const float* trans,
                            const float* d_verts0, const float* d_joints, float* d_pose, float* d_shape, float* d_trans,
                            void* stream) {
  (void)trans;
  HARP_REQUIRE(c && B > 0 && pose && shape && d_verts0 && d_pose && d_shape, HARP_ERR_INVALID, "harp_lbs_bwd: null argument");
  HARP_REQUIRE(c->has_model, HARP_ERR_NO_MODEL, "harp_lbs_bwd: ctx has no skinning model");
  size_t dyn = lbs_bwd_smem(c);
  static size_t configured[64] = {};  // the attribute is per device
  int dev = 0;
  HARP_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || dyn > configured[dev]) {
    HARP_CHECK_CUDA(cudaFuncSetAttribute(lbs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    if (dev >= 0 && dev < 64) configured[dev] = dyn;
  }
  harp_prof_begin(c, 3, (cudaStream_t)stream);
  lbs_bwd_kernel<<<B, LBS_THREADS, dyn, (cudaStream_t)stream>>>(*c, pose, shape, shape_shared, d_verts0, d_joints, d_pose,
                                                                d_shape, d_trans);
  harp_prof_end(c, 3, (cudaStream_t)stream);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

// MANO entry points: the MANO layout (16 joints, pose48 = [rot | pose45], 10 betas) is one instance of the generic model.
extern "C" int harp_lbs_mano_fwd(harp_ctx* c, int B, const float* pose48, const float* betas, int betas_shared,
                                 const float* trans, float* verts0, float* joints, void* stream) {
  HARP_REQUIRE(c && c->has_model && c->J == 16 && c->n_shape == 10, HARP_ERR_NO_MODEL, "harp_lbs_mano_fwd: ctx has no MANO model");
  return harp_lbs_fwd(c, B, pose48, betas, betas_shared, trans, verts0, joints, stream);
}

extern "C" int harp_lbs_mano_bwd(harp_ctx* c, int B, const float* pose48, const float* betas, int betas_shared,
                                 const float* trans, const float* d_verts0, const float* d_joints, float* d_pose48,
                                 float* d_betas, float* d_trans, void* stream) {
  HARP_REQUIRE(c && c->has_model && c->J == 16 && c->n_shape == 10, HARP_ERR_NO_MODEL, "harp_lbs_mano_bwd: ctx has no MANO model");
  return harp_lbs_bwd(c, B, pose48, betas, betas_shared, trans, d_verts0, d_joints, d_pose48, d_betas, d_trans, stream);
}
