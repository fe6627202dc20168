// MANO linear blend skinning, forward and analytic backward: one CTA per frame.
// Replaces ManoLayer.forward (manopth/manolayer.py:108-296) + batch_rodrigues (rodrigues_layer.py:43-54) and their
// autograd graph (~40 ATen launches, 2 host syncs) with one launch each way.  FP32 FFMA throughout: the blendshape
// contractions are [135 x 2334] per frame — far too small for tcgen05 tiles and needed at full fp32 accuracy
// (last bullet).
#include "common.cuh"

namespace {

constexpr int LBS_THREADS = 512;
constexpr int NJ = 16;

// joint order of the returned 21 joints (manolayer.py:279) indexing [16 chain joints | 5 tips]
__constant__ int c_joint_reorder[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};

__device__ __forceinline__ int mano_parent(int j) { return ((j - 1) % 3 == 0) ? 0 : j - 1; }

struct LbsShared {
  float beta[10];
  float pose[48];        // full pose (hands_mean added)
  float R[NJ][9];
  float J[NJ][3];
  float G[NJ][12];       // 3x4 [R|t] global transforms
  float A[NJ][12];       // with rest joint removed
  float pose_map[135];
};

// Recompute everything that depends only on (pose, beta): shared by forward and backward.
__device__ void lbs_prologue(const harp_ctx& c, const float* pose48, const float* beta, LbsShared& s) {
  int t = threadIdx.x;
  if (t < 10) s.beta[t] = beta[t];
  if (t < 48) s.pose[t] = pose48[t] + (t >= 3 ? c.hands_mean[t - 3] : 0.f);
  __syncthreads();
  if (t < NJ) rodrigues_quat<float>(&s.pose[3 * t], s.R[t]);
  if (t >= 32 && t < 32 + 48) {
    int i = t - 32;
    float a = c.J_template[i];
#pragma unroll
    for (int b = 0; b < 10; ++b) a += c.J_shapedirs[i * 10 + b] * s.beta[b];
    s.J[i / 3][i % 3] = a;
  }
  __syncthreads();
  if (t < 135) {
    int j = t / 9 + 1, k = t % 9;
    s.pose_map[t] = s.R[j][k] - ((k == 0 || k == 4 || k == 8) ? 1.f : 0.f);
  }
  if (t == 0) {
    for (int k = 0; k < 9; ++k) s.G[0][(k / 3) * 4 + (k % 3)] = s.R[0][k];
    for (int k = 0; k < 3; ++k) s.G[0][k * 4 + 3] = s.J[0][k];
  }
  __syncthreads();
  if (t < 5) {  // one finger per thread, three levels
    for (int lev = 0; lev < 3; ++lev) {
      int j = 1 + 3 * t + lev, p = mano_parent(j);
      float rel[3] = {s.J[j][0] - s.J[p][0], s.J[j][1] - s.J[p][1], s.J[j][2] - s.J[p][2]};
      for (int r = 0; r < 3; ++r) {
        for (int cc = 0; cc < 3; ++cc) {
          float a = 0.f;
          for (int k = 0; k < 3; ++k) a += s.G[p][r * 4 + k] * s.R[j][k * 3 + cc];
          s.G[j][r * 4 + cc] = a;
        }
        float a = s.G[p][r * 4 + 3];
        for (int k = 0; k < 3; ++k) a += s.G[p][r * 4 + k] * rel[k];
        s.G[j][r * 4 + 3] = a;
      }
    }
  }
  __syncthreads();
  if (t < NJ) {
    for (int r = 0; r < 3; ++r) {
      float tj = 0.f;
      for (int k = 0; k < 3; ++k) {
        s.A[t][r * 4 + k] = s.G[t][r * 4 + k];
        tj += s.G[t][r * 4 + k] * s.J[t][k];
      }
      s.A[t][r * 4 + 3] = s.G[t][r * 4 + 3] - tj;
    }
  }
  __syncthreads();
}

// v_posed for vertex v (template + shape + pose blendshapes); blendshape tables are stored transposed
// (k, V0*3) so that consecutive threads read consecutive addresses.
__device__ __forceinline__ f3 lbs_v_posed(const harp_ctx& c, const LbsShared& s, int v) {
  int n3 = c.V0 * 3, i = v * 3;
  float x = c.v_template[i], y = c.v_template[i + 1], z = c.v_template[i + 2];
#pragma unroll
  for (int b = 0; b < 10; ++b) {
    const float* p = c.shapedirs + (size_t)b * n3 + i;
    float w = s.beta[b];
    x += p[0] * w; y += p[1] * w; z += p[2] * w;
  }
#pragma unroll 5
  for (int k = 0; k < 135; ++k) {
    const float* p = c.posedirs + (size_t)k * n3 + i;
    float w = s.pose_map[k];
    x += p[0] * w; y += p[1] * w; z += p[2] * w;
  }
  return mk3(x, y, z);
}

__device__ __forceinline__ void lbs_blend(const harp_ctx& c, const LbsShared& s, int v, float T[12]) {
#pragma unroll
  for (int k = 0; k < 12; ++k) T[k] = 0.f;
  const float* w = c.weights + (size_t)v * NJ;
#pragma unroll 4
  for (int j = 0; j < NJ; ++j) {
    float wj = w[j];
    if (wj != 0.f) {
#pragma unroll
      for (int k = 0; k < 12; ++k) T[k] += wj * s.A[j][k];
    }
  }
}

__global__ void __launch_bounds__(LBS_THREADS) lbs_mano_fwd_kernel(harp_ctx c, const float* __restrict__ pose48,
                                                                     const float* __restrict__ betas, int betas_shared,
                                                                     const float* __restrict__ trans,
                                                                     float* __restrict__ verts0, float* __restrict__ joints) {
  __shared__ LbsShared s;
  int b = blockIdx.x;
  lbs_prologue(c, pose48 + (size_t)b * 48, betas + (betas_shared ? 0 : (size_t)b * 10), s);
  f3 tr = ld3(trans + (size_t)b * 3);
  float* vout = verts0 + (size_t)b * c.V0 * 3;
  for (int v = threadIdx.x; v < c.V0; v += blockDim.x) {
    f3 vp = lbs_v_posed(c, s, v);
    float T[12];
    lbs_blend(c, s, v, T);
    f3 o = mk3(T[0] * vp.x + T[1] * vp.y + T[2] * vp.z + T[3], T[4] * vp.x + T[5] * vp.y + T[6] * vp.z + T[7],
               T[8] * vp.x + T[9] * vp.y + T[10] * vp.z + T[11]);
    st3(vout + (size_t)v * 3, o + tr);
  }
  __syncthreads();
  if (threadIdx.x < 21) {
    int src = c_joint_reorder[threadIdx.x];
    f3 p;
    if (src < NJ) p = mk3(s.G[src][3], s.G[src][7], s.G[src][11]) + tr;
    else p = ld3(vout + (size_t)c.tips[src - NJ] * 3);  // already translated
    st3(joints + ((size_t)b * 21 + threadIdx.x) * 3, p);
  }
}

struct LbsBwdShared {
  float dA[NJ][12];
  float dR[NJ][9];
  float dJ[NJ][3];
  float djtr[NJ][3];
  float dpm[135];
  float dbeta[10];
  float red[32];
};

__global__ void __launch_bounds__(LBS_THREADS) lbs_mano_bwd_kernel(harp_ctx c, const float* __restrict__ pose48,
                                                                     const float* __restrict__ betas, int betas_shared,
                                                                     const float* __restrict__ d_verts0,
                                                                     const float* __restrict__ d_joints,
                                                                     float* __restrict__ d_pose48, float* __restrict__ d_betas,
                                                                     float* __restrict__ d_trans) {
  __shared__ LbsShared s;
  __shared__ LbsBwdShared g;
  extern __shared__ float dyn[];
  float* dvp = dyn;                 // (V0*3) gradient w.r.t. v_posed
  float* dTs = dyn + c.V0 * 3;      // (12, V0) per-vertex dT = g (x) [v_posed;1], k-major (bank-conflict free)
  int b = blockIdx.x, t = threadIdx.x;
  lbs_prologue(c, pose48 + (size_t)b * 48, betas + (betas_shared ? 0 : (size_t)b * 10), s);
  const float* dj = d_joints ? d_joints + (size_t)b * 63 : nullptr;
  if (t < NJ * 3) (&g.djtr[0][0])[t] = 0.f;
  __syncthreads();
  if (t < 21 && dj) {
    int src = c_joint_reorder[t];
    if (src < NJ)
      for (int k = 0; k < 3; ++k) g.djtr[src][k] = dj[t * 3 + k];
  }
  __syncthreads();
  const float* gv = d_verts0 + (size_t)b * c.V0 * 3;
  f3 tsum = mk3(0.f, 0.f, 0.f);
  for (int v = t; v < c.V0; v += blockDim.x) {
    f3 gg = ld3(gv + (size_t)v * 3);
    if (dj) {
#pragma unroll
      for (int k = 0; k < 5; ++k)
        if (c.tips[k] == v) {
          // output slot of tip k: position of (16+k) in the reorder table
          for (int o = 0; o < 21; ++o)
            if (c_joint_reorder[o] == NJ + k) gg = gg + ld3(dj + o * 3);
        }
    }
    tsum = tsum + gg;
    f3 vp = lbs_v_posed(c, s, v);
    float T[12];
    lbs_blend(c, s, v, T);
    // d v_posed = T.R^T g
    dvp[v * 3 + 0] = T[0] * gg.x + T[4] * gg.y + T[8] * gg.z;
    dvp[v * 3 + 1] = T[1] * gg.x + T[5] * gg.y + T[9] * gg.z;
    dvp[v * 3 + 2] = T[2] * gg.x + T[6] * gg.y + T[10] * gg.z;
    // dT = g (x) [vp;1]  ->  dA_j += w_j dT
    float dT[12] = {gg.x * vp.x, gg.x * vp.y, gg.x * vp.z, gg.x, gg.y * vp.x, gg.y * vp.y, gg.y * vp.z, gg.y,
                    gg.z * vp.x, gg.z * vp.y, gg.z * vp.z, gg.z};
#pragma unroll
    for (int k = 0; k < 12; ++k) dTs[k * c.V0 + v] = dT[k];
  }
  // d_trans = sum_v g_v + sum of chain-joint grads (tips were folded into g_v)
  float tx = block_sum(tsum.x, g.red), ty = block_sum(tsum.y, g.red), tz = block_sum(tsum.z, g.red);
  __syncthreads();
  if (t == 0) {
    for (int j = 0; j < NJ; ++j) { tx += g.djtr[j][0]; ty += g.djtr[j][1]; tz += g.djtr[j][2]; }
    if (d_trans) { d_trans[b * 3] = tx; d_trans[b * 3 + 1] = ty; d_trans[b * 3 + 2] = tz; }
  }
  int warp = t >> 5, lane = t & 31, nw = blockDim.x >> 5, n3 = c.V0 * 3;
  // dA (16x12) = W^T (16 x V0) . dT (V0 x 12): one warp per output, lanes over vertices (no shared-memory float atomics:
  // those compile to CAS spin loops)
  for (int o = warp; o < NJ * 12; o += nw) {
    int j = o / 12, k = o % 12;
    const float* wrow = c.weightsT + (size_t)j * c.V0;
    const float* drow = dTs + k * c.V0;
    float a = 0.f;
    for (int v = lane; v < c.V0; v += 32) a += wrow[v] * drow[v];
    a = warp_sum(a);
    if (lane == 0) g.dA[j][k] = a;
  }
  // d pose_map (135) and the direct d beta (10): transposed-table mat-vec, one warp per output
  for (int o = warp; o < 145; o += nw) {
    const float* row = o < 135 ? c.posedirs + (size_t)o * n3 : c.shapedirs + (size_t)(o - 135) * n3;
    float a = 0.f;
    for (int i = lane; i < n3; i += 32) a += row[i] * dvp[i];
    a = warp_sum(a);
    if (lane == 0) { if (o < 135) g.dpm[o] = a; else g.dbeta[o - 135] = a; }
  }
  __syncthreads();
  if (t == 0) {
    // reverse-mode through A_j = [G.R | G.t - G.R J_j] and the kinematic chain G_j = G_p [R_j | J_j - J_p]
    float dGR[NJ][9], dGt[NJ][3];
    for (int j = 0; j < NJ; ++j) {
      for (int r = 0; r < 3; ++r) {
        float dat = g.dA[j][r * 4 + 3];
        dGt[j][r] = dat + g.djtr[j][r];
        for (int k = 0; k < 3; ++k) dGR[j][r * 3 + k] = g.dA[j][r * 4 + k] - dat * s.J[j][k];
      }
      for (int k = 0; k < 3; ++k) {
        float a = 0.f;
        for (int r = 0; r < 3; ++r) a += s.G[j][r * 4 + k] * g.dA[j][r * 4 + 3];
        g.dJ[j][k] = -a;
      }
    }
    for (int j = NJ - 1; j >= 1; --j) {
      int p = mano_parent(j);
      float rel[3] = {s.J[j][0] - s.J[p][0], s.J[j][1] - s.J[p][1], s.J[j][2] - s.J[p][2]};
      for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k) {
          // dG_p.R[r][k] += sum_c dG_j.R[r][c] R_j[k][c] + dG_j.t[r] rel[k]
          float a = dGt[j][r] * rel[k];
          for (int cc = 0; cc < 3; ++cc) a += dGR[j][r * 3 + cc] * s.R[j][k * 3 + cc];
          dGR[p][r * 3 + k] += a;
        }
      for (int k = 0; k < 3; ++k)
        for (int cc = 0; cc < 3; ++cc) {
          // dR_j[k][c] = sum_r G_p.R[r][k] dG_j.R[r][c]
          float a = 0.f;
          for (int r = 0; r < 3; ++r) a += s.G[p][r * 4 + k] * dGR[j][r * 3 + cc];
          g.dR[j][k * 3 + cc] = a;
        }
      for (int k = 0; k < 3; ++k) {
        float a = 0.f;
        for (int r = 0; r < 3; ++r) a += s.G[p][r * 4 + k] * dGt[j][r];
        g.dJ[j][k] += a;
        g.dJ[p][k] -= a;
      }
      for (int r = 0; r < 3; ++r) dGt[p][r] += dGt[j][r];
    }
    for (int k = 0; k < 9; ++k) g.dR[0][k] = dGR[0][k];
    for (int k = 0; k < 3; ++k) g.dJ[0][k] += dGt[0][k];
  }
  __syncthreads();
  if (t < 135) g.dR[t / 9 + 1][t % 9] += g.dpm[t];  // pose_map = R - I
  __syncthreads();
  if (t < NJ) {  // Rodrigues Jacobian by forward-mode duals
    Dual<3> a[3] = {Dual<3>::var(s.pose[3 * t], 0), Dual<3>::var(s.pose[3 * t + 1], 1), Dual<3>::var(s.pose[3 * t + 2], 2)};
    Dual<3> R[9];
    rodrigues_quat<Dual<3>>(a, R);
    for (int i = 0; i < 3; ++i) {
      float acc = 0.f;
      for (int k = 0; k < 9; ++k) acc += R[k].d[i] * g.dR[t][k];
      d_pose48[(size_t)b * 48 + 3 * t + i] = acc;
    }
  }
  if (t >= 32 && t < 42) {
    int bi = t - 32;
    float a = g.dbeta[bi];
    for (int i = 0; i < 48; ++i) a += c.J_shapedirs[i * 10 + bi] * (&g.dJ[0][0])[i];
    d_betas[(size_t)b * 10 + bi] = a;
  }
}

}  // namespace

extern "C" int harp_lbs_mano_fwd(harp_ctx* c, int B, const float* pose48, const float* betas, int betas_shared,
                                 const float* trans, float* verts0, float* joints, void* stream) {
  HARP_REQUIRE(c && B > 0 && pose48 && betas && trans && verts0 && joints, HARP_ERR_INVALID, "harp_lbs_mano_fwd: null argument");
  HARP_REQUIRE(c->has_model, HARP_ERR_NO_MODEL, "harp_lbs_mano_fwd: ctx has no MANO model");
  lbs_mano_fwd_kernel<<<B, LBS_THREADS, 0, (cudaStream_t)stream>>>(*c, pose48, betas, betas_shared, trans, verts0, joints);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

extern "C" int harp_lbs_mano_bwd(harp_ctx* c, int B, const float* pose48, const float* betas, int betas_shared,
                                 const float* trans, const float* d_verts0, const float* d_joints, float* d_pose48,
                                 float* d_betas, float* d_trans, void* stream) {
  (void)trans;
  HARP_REQUIRE(c && B > 0 && pose48 && betas && d_verts0 && d_pose48 && d_betas, HARP_ERR_INVALID, "harp_lbs_mano_bwd: null argument");
  HARP_REQUIRE(c->has_model, HARP_ERR_NO_MODEL, "harp_lbs_mano_bwd: ctx has no MANO model");
  size_t dyn = (size_t)c->V0 * 15 * sizeof(float);
  static size_t configured = 0;
  if (dyn > configured) {
    HARP_CHECK_CUDA(cudaFuncSetAttribute(lbs_mano_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    configured = dyn;
  }
  lbs_mano_bwd_kernel<<<B, LBS_THREADS, dyn, (cudaStream_t)stream>>>(*c, pose48, betas, betas_shared, d_verts0, d_joints,
                                                                     d_pose48, d_betas, d_trans);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}
