// Device-side pieces of the kinematic-tree LBS (forward and analytic backward of ONE frame by one CTA), shared by the
// stand-alone kernels (lbs.cu) and the fused per-frame geometry kernels of the optimiser step (geom.cu).
// (manopth/manolayer.py:108-296, rodrigues_layer.py:43-54; hand_models_harp/body_models.py:2163-2390 -> smplx.lbs.lbs)
#pragma once
#include "common.cuh"

namespace {


constexpr int LBS_THREADS = 1024;  // one vertex per thread for MANO (778); 64 registers x 1024 threads = one SM's register file
constexpr int MAXJ = HARP_MAX_JOINTS;      // 64
constexpr int MAXS = HARP_MAX_SHAPE;       // 32
constexpr int MAXPB = (MAXJ - 1) * 9;      // pose-blendshape basis size

struct LbsShared {
  float beta[MAXS];
  float pose[MAXJ * 3];   // full axis-angle pose (pose_mean added)
  float R[MAXJ][9];
  float J[MAXJ][3];
  float G[MAXJ][12];      // 3x4 [R|t] global transforms
  float A[MAXJ][12];      // with rest joint removed
  float pose_map[MAXPB];
  int parent[MAXJ];
  int depth[MAXJ];        // distance to the root: joints of one depth are independent in the chain (forward and reverse)
  int max_depth;
};

// smplx.lbs.batch_rodrigues: R = I + sin(a) K + (1 - cos(a)) K^2, a = |r + 1e-8|   
template <typename T>
__device__ inline void rodrigues_smplx(const T r[3], T R[9]) {
  T ax = r[0] + T(1e-8f), ay = r[1] + T(1e-8f), az = r[2] + T(1e-8f);
  T angle = dsqrt(ax * ax + ay * ay + az * az);
  T x = r[0] / angle, y = r[1] / angle, z = r[2] / angle;
  T s = dsin(angle), c1 = T(1.f) - dcos(angle);
  // K = [[0,-z,y],[z,0,-x],[-y,x,0]];  K^2 = [[-(y2+z2), xy, xz],[xy, -(x2+z2), yz],[xz, yz, -(x2+y2)]]
  T xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = T(1.f) - c1 * (yy + zz); R[1] = c1 * xy - s * z;          R[2] = c1 * xz + s * y;
  R[3] = c1 * xy + s * z;          R[4] = T(1.f) - c1 * (xx + zz); R[5] = c1 * yz - s * x;
  R[6] = c1 * xz - s * y;          R[7] = c1 * yz + s * x;          R[8] = T(1.f) - c1 * (xx + yy);
}

template <typename T>
__device__ inline void rodrigues_any(int kind, const T r[3], T R[9]) {
  if (kind == 0) rodrigues_quat<T>(r, R);
  else rodrigues_smplx<T>(r, R);
}

// Everything that depends only on (pose, shape): shared by forward and backward.
__device__ void lbs_prologue(const harp_ctx& c, const float* pose, const float* beta, LbsShared& s) {
  int t = threadIdx.x, J = c.J, NS = c.n_shape;
  if (t < NS) s.beta[t] = beta[t];
  if (t < J * 3) s.pose[t] = pose[t] + c.pose_mean[t];
  if (t < J) s.parent[t] = c.parents[t];
  if (t == 0) s.max_depth = 0;
  __syncthreads();
  if (t < J) {
    int d = 0;
    for (int q = t; q > 0; q = s.parent[q]) ++d;  // parents precede children (SMPL family ordering): the root is joint 0
    s.depth[t] = d;
    atomicMax(&s.max_depth, d);
  }
  if (t < J) rodrigues_any<float>(c.rodrigues_kind, &s.pose[3 * t], s.R[t]);
  if (t >= 64 && t < 64 + J * 3) {
    int i = t - 64;
    float a = c.J_template[i];
    for (int b = 0; b < NS; ++b) a += c.J_shapedirs[i * NS + b] * s.beta[b];
    s.J[i / 3][i % 3] = a;
  }
  __syncthreads();
  for (int i = t; i < (J - 1) * 9; i += blockDim.x) {
    int j = i / 9 + 1, k = i % 9;
    s.pose_map[i] = s.R[j][k] - ((k == 0 || k == 4 || k == 8) ? 1.f : 0.f);
  }
  // kinematic chain G_j = G_parent [R_j | J_j - J_parent], one tree level per barrier, a thread per entry of the 3x4 transform
  // (the same operation order per entry as a serial walk: bit-identical, ~4 us shorter per launch for MANO's 16 joints)
  if (t < 12) s.G[0][t] = (t & 3) == 3 ? s.J[0][t >> 2] : s.R[0][(t >> 2) * 3 + (t & 3)];
  for (int lvl = 1; lvl <= s.max_depth; ++lvl) {
    __syncthreads();
    for (int i = t; i < J * 12; i += blockDim.x) {
      const int j = i / 12, e = i % 12, r = e >> 2, cc = e & 3;
      if (s.depth[j] == lvl) {
        const int p = s.parent[j];
        if (cc < 3) {
          float a = 0.f;
          for (int k = 0; k < 3; ++k) a += s.G[p][r * 4 + k] * s.R[j][k * 3 + cc];
          s.G[j][e] = a;
        } else {
          float a = s.G[p][r * 4 + 3];
          for (int k = 0; k < 3; ++k) a += s.G[p][r * 4 + k] * (s.J[j][k] - s.J[p][k]);
          s.G[j][e] = a;
        }
      }
    }
  }
  __syncthreads();
  if (t < J) {
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
  // The blendshape offsets (|offset| ~ 1e-3 m) are summed on their own, in double, and added to the template once — as the
  // reference does (v_shaped = template + blend_shapes(...), v_posed = v_shaped + matmul(pose_feature, posedirs):
  // manolayer.py:181-200, smplx.lbs.lbs).  Accumulating them one by one INTO the coordinate (|x| ~ 0.5 m, ulp 6e-8) loses
  // ~1e-6 m over the arm's 486 terms, which the sigma = 1e-7 silhouette turns into 5e-3 of gradient error.
  int n3 = c.V0 * 3, i = v * 3, NS = c.n_shape, NP = (c.J - 1) * 9;
  double sx = 0.0, sy = 0.0, sz = 0.0;
  for (int b = 0; b < NS; ++b) {
    const float* p = c.shapedirs + (size_t)b * n3 + i;
    double w = s.beta[b];
    sx += p[0] * w; sy += p[1] * w; sz += p[2] * w;
  }
  float x = c.v_template[i] + (float)sx, y = c.v_template[i + 1] + (float)sy, z = c.v_template[i + 2] + (float)sz;
  double px = 0.0, py = 0.0, pz = 0.0;
#pragma unroll 9
  for (int k = 0; k < NP; ++k) {
    const float* p = c.posedirs + (size_t)k * n3 + i;
    double w = s.pose_map[k];
    px += p[0] * w; py += p[1] * w; pz += p[2] * w;
  }
  return mk3(x + (float)px, y + (float)py, z + (float)pz);
}

__device__ __forceinline__ void lbs_blend(const harp_ctx& c, const LbsShared& s, int v, float T[12]) {
#pragma unroll
  for (int k = 0; k < 12; ++k) T[k] = 0.f;
  const float* w = c.weights + (size_t)v * c.J;
  for (int j = 0; j < c.J; ++j) {
    float wj = w[j];
    if (wj != 0.f) {
#pragma unroll
      for (int k = 0; k < 12; ++k) T[k] += wj * s.A[j][k];
    }
  }
}

// LBS of one frame by the whole CTA, after lbs_prologue: skinned vertices to `vout` (V0,3) [and `vout2` if given], the output
// joints to `jout` (n_joints_out,3).  The pointers may be shared or global memory.
__device__ __forceinline__ void lbs_forward_core(const harp_ctx& c, const LbsShared& s, f3 trans, float* vout, float* vout2, float* jout,
                                                 float* vposed_out = nullptr) {
  const int J = c.J;
  f3 off = trans;
  if (c.center_joint >= 0)  // wrist-centring of SMPLXARM (body_models.py:2342-2343)
    off = off - mk3(s.G[c.center_joint][3], s.G[c.center_joint][7], s.G[c.center_joint][11]);
  for (int v = threadIdx.x; v < c.V0; v += blockDim.x) {
    f3 vp = lbs_v_posed(c, s, v);
    if (vposed_out) st3(vposed_out + (size_t)v * 3, vp);  // kept for the backward (saves its pass over the pose blendshapes)
    float T[12];
    lbs_blend(c, s, v, T);
    f3 o = mk3(T[0] * vp.x + T[1] * vp.y + T[2] * vp.z + T[3], T[4] * vp.x + T[5] * vp.y + T[6] * vp.z + T[7],
               T[8] * vp.x + T[9] * vp.y + T[10] * vp.z + T[11]);
    st3(vout + (size_t)v * 3, o + off);
    if (vout2) st3(vout2 + (size_t)v * 3, o + off);
  }
  __syncthreads();
  if (threadIdx.x < c.n_joints_out) {
    int src = c.joint_out_src[threadIdx.x];
    f3 p;
    if (src < J) p = mk3(s.G[src][3], s.G[src][7], s.G[src][11]) + off;
    else p = ld3(vout + (size_t)c.tips[src - J] * 3);  // vertex-joints (fingertips): already offset
    st3(jout + (size_t)threadIdx.x * 3, p);
  }
}

struct LbsBwdShared {
  float dA[MAXJ][12];
  float dR[MAXJ][9];
  float dJ[MAXJ][3];
  float djtr[MAXJ][3];
  float dpm[MAXPB];
  float dbeta[MAXS];
  float red[32];
};

// Backward of one frame's LBS, after lbs_prologue, in two parts so that several CTAs (a thread-block cluster) can share a frame:
//
//   lbs_backward_accumulate  over the vertices [vb, ve): d v_posed and dT = g (x) [v_posed; 1] per vertex, then the contractions
//                            against the skinning weights and the blendshape tables restricted to that range ->
//                            PARTIAL sums  out = [dA (J*12) | d pose_map ((J-1)*9) | d beta (n_shape) | sum of g (3)]
//   lbs_backward_finish      sums the partial results of all ranges, then the reverse-mode pass through the kinematic chain,
//                            the Rodrigues Jacobians and the joint regressor -> d_pose (J*3), d_shape (n_shape), d_trans (3)
//
// gv: (V0,3) gradient w.r.t. the skinned vertices; dj: (n_joints_out,3) gradient w.r.t. the output joints or null; dvp (V0*3)
// and dTs (12*V0) are scratch indexed by the global vertex id.  All pointers may be shared or global memory.
__host__ __device__ inline int lbs_partial_floats(int J, int NS) { return J * 12 + (J - 1) * 9 + NS + 3; }

__device__ __forceinline__ void lbs_backward_accumulate(const harp_ctx& c, const LbsShared& s, float* red, float* dvp, float* dTs,
                                                        const float* gv, const float* dj, const float* vposed, int vb, int ve,
                                                        float* out) {
  const int t = threadIdx.x, J = c.J, NS = c.n_shape, NP = (J - 1) * 9, NJO = c.n_joints_out;
  f3 tsum = mk3(0.f, 0.f, 0.f);
  for (int v = vb + t; v < ve; v += blockDim.x) {
    f3 gg = ld3(gv + (size_t)v * 3);
    if (dj) {
      for (int k = 0; k < c.n_tips; ++k)
        if (c.tips[k] == v) {
          for (int o = 0; o < NJO; ++o)
            if (c.joint_out_src[o] == J + k) gg = gg + ld3(dj + o * 3);
        }
    }
    tsum = tsum + gg;
    f3 vp = vposed ? ld3(vposed + (size_t)v * 3) : lbs_v_posed(c, s, v);
    float T[12];
    lbs_blend(c, s, v, T);
    // d v_posed = T.R^T g
    dvp[v * 3 + 0] = T[0] * gg.x + T[4] * gg.y + T[8] * gg.z;
    dvp[v * 3 + 1] = T[1] * gg.x + T[5] * gg.y + T[9] * gg.z;
    dvp[v * 3 + 2] = T[2] * gg.x + T[6] * gg.y + T[10] * gg.z;
    float dT[12] = {gg.x * vp.x, gg.x * vp.y, gg.x * vp.z, gg.x, gg.y * vp.x, gg.y * vp.y, gg.y * vp.z, gg.y,
                    gg.z * vp.x, gg.z * vp.y, gg.z * vp.z, gg.z};
#pragma unroll
    for (int k = 0; k < 12; ++k) dTs[k * c.V0 + v] = dT[k];
  }
  float tx = block_sum(tsum.x, red), ty = block_sum(tsum.y, red), tz = block_sum(tsum.z, red);  // (block_sum synchronises)
  __syncthreads();
  if (t == 0) { out[J * 12 + NP + NS] = tx; out[J * 12 + NP + NS + 1] = ty; out[J * 12 + NP + NS + 2] = tz; }
  const int warp = t >> 5, lane = t & 31, nw = blockDim.x >> 5, n3 = c.V0 * 3;
  // dA (J x 12) = W^T (J x V0) . dT (V0 x 12): one warp per output, lanes over vertices (no shared-memory float atomics
  // in the hot loop: those compile to CAS spin loops)
  for (int o = warp; o < J * 12; o += nw) {
    int j = o / 12, k = o % 12;
    const float* wrow = c.weightsT + (size_t)j * c.V0;
    const float* drow = dTs + k * c.V0;
    float a = 0.f;
#pragma unroll 8
    for (int v = vb + lane; v < ve; v += 32) a += __ldg(wrow + v) * drow[v];  // unrolled: eight table loads in flight per lane
    a = warp_sum(a);
    if (lane == 0) out[o] = a;
  }
  // d pose_map and the direct d shape: transposed-table mat-vec, one warp per group of five outputs (145 = 29 x 5: one round
  // for MANO), 8-byte loads, four iterations unrolled: 20 table loads in flight per lane — the loop is bound by the L2
  // latency of the 1.3 MB blendshape table, not by arithmetic
  {
    constexpr int RW = 5;
    const int ib = 3 * vb, ie = 3 * ve;
    const bool vec2 = ((n3 | ib | ie) & 1) == 0 && (reinterpret_cast<size_t>(dvp) & 7) == 0;  // rows and range 8-byte aligned
    const float2* d2 = reinterpret_cast<const float2*>(dvp);
    for (int o0 = warp * RW; o0 < NP + NS; o0 += nw * RW) {
      const float* row[RW];
      float a[RW];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        int o = min(o0 + r, NP + NS - 1);
        row[r] = o < NP ? c.posedirs + (size_t)o * n3 : c.shapedirs + (size_t)(o - NP) * n3;
        a[r] = 0.f;
      }
      if (vec2) {
#pragma unroll 4
        for (int i = (ib >> 1) + lane; i < (ie >> 1); i += 32) {
          float2 d = d2[i];
#pragma unroll
          for (int r = 0; r < RW; ++r) {
            float2 x = __ldg(reinterpret_cast<const float2*>(row[r]) + i);
            a[r] += x.x * d.x + x.y * d.y;
          }
        }
      } else {
        for (int i = ib + lane; i < ie; i += 32) {
          float d = dvp[i];
#pragma unroll
          for (int r = 0; r < RW; ++r) a[r] += row[r][i] * d;
        }
      }
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        float v = warp_sum(a[r]);
        int o = o0 + r;
        if (lane == 0 && o < NP + NS) out[J * 12 + o] = v;
      }
    }
  }
}

// parts: `nparts` partial-result blocks of `stride` floats each (lbs_backward_accumulate); dTs: scratch of >= MAXJ*12 floats
__device__ __forceinline__ void lbs_backward_finish(const harp_ctx& c, LbsShared& s, LbsBwdShared& g, float* dTs, const float* parts,
                                                    int nparts, int stride, const float* dj, float* d_pose, float* d_shape,
                                                    float* d_trans) {
  const int t = threadIdx.x, J = c.J, NS = c.n_shape, NP = (J - 1) * 9, NJO = c.n_joints_out;
  if (t < J * 3) (&g.djtr[0][0])[t] = 0.f;
  __syncthreads();
  if (t < NJO && dj) {
    int src = c.joint_out_src[t];
    if (src < J)
      for (int k = 0; k < 3; ++k) atomicAdd(&g.djtr[src][k], dj[t * 3 + k]);  // (a chain joint may be listed twice)
  }
  for (int i = t; i < J * 12 + NP + NS; i += blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < nparts; ++r) a += parts[(size_t)r * stride + i];
    if (i < J * 12) g.dA[i / 12][i % 12] = a;
    else if (i < J * 12 + NP) g.dpm[i - J * 12] = a;
    else g.dbeta[i - J * 12 - NP] = a;
  }
  __syncthreads();
  if (t == 0) {
    // total translation gradient = sum_v g_v + sum of chain-joint grads (vertex-joints were folded into g_v)
    float tx = 0.f, ty = 0.f, tz = 0.f;
    for (int r = 0; r < nparts; ++r) {
      const float* q = parts + (size_t)r * stride + J * 12 + NP + NS;
      tx += q[0]; ty += q[1]; tz += q[2];
    }
    for (int j = 0; j < J; ++j) { tx += g.djtr[j][0]; ty += g.djtr[j][1]; tz += g.djtr[j][2]; }
    if (d_trans) { d_trans[0] = tx; d_trans[1] = ty; d_trans[2] = tz; }
    if (c.center_joint >= 0) {  // every output was shifted by -G[center].t
      g.djtr[c.center_joint][0] -= tx; g.djtr[c.center_joint][1] -= ty; g.djtr[c.center_joint][2] -= tz;
    }
  }
  __syncthreads();
  // reverse-mode through A_j = [G.R | G.t - G.R J_j] and the kinematic chain G_j = G_p [R_j | J_j - J_p]; dG accumulators reuse
  // dTs (dead after the contraction above).  One tree level per barrier: first every joint of the level turns its (final) dG
  // into dR_j and its share of dJ_j, then every parent one level up GATHERS from its children in descending joint order — the
  // order a serial walk from the last joint to the first adds them in, so the sums are bit-identical and deterministic (the
  // single-thread walk was 20 us of this kernel's 130).
  float(*dGR)[9] = reinterpret_cast<float(*)[9]>(dTs);
  float(*dGt)[3] = reinterpret_cast<float(*)[3]>(dTs + MAXJ * 9);
  for (int i = t; i < J * 12; i += blockDim.x) {
    const int j = i / 12, e = i % 12;
    if (e < 9) {
      const int r = e / 3, k = e % 3;
      dGR[j][e] = g.dA[j][r * 4 + k] - g.dA[j][r * 4 + 3] * s.J[j][k];
    } else {
      const int k = e - 9;
      dGt[j][k] = g.dA[j][k * 4 + 3] + g.djtr[j][k];
      float a = 0.f;
      for (int r = 0; r < 3; ++r) a += s.G[j][r * 4 + k] * g.dA[j][r * 4 + 3];
      g.dJ[j][k] = -a;
    }
  }
  for (int lvl = s.max_depth; lvl >= 1; --lvl) {
    __syncthreads();
    for (int i = t; i < J * 12; i += blockDim.x) {  // joints of this level: dG_j is final
      const int j = i / 12, e = i % 12;
      if (s.depth[j] != lvl) continue;
      const int p = s.parent[j];
      if (e < 9) {  // dR_j[k][c] = sum_r G_p.R[r][k] dG_j.R[r][c]
        const int k = e / 3, cc = e % 3;
        float a = 0.f;
        for (int r = 0; r < 3; ++r) a += s.G[p][r * 4 + k] * dGR[j][r * 3 + cc];
        g.dR[j][e] = a;
      } else {
        const int k = e - 9;
        float a = 0.f;
        for (int r = 0; r < 3; ++r) a += s.G[p][r * 4 + k] * dGt[j][r];
        g.dJ[j][k] += a;
      }
    }
    __syncthreads();
    for (int i = t; i < J * 15; i += blockDim.x) {  // parents one level up gather from their children
      const int p = i / 15, e = i % 15;
      if (s.depth[p] != lvl - 1) continue;
      for (int j = J - 1; j > p; --j) {
        if (s.parent[j] != p) continue;
        if (e < 9) {  // dG_p.R[r][k] += sum_c dG_j.R[r][c] R_j[k][c] + dG_j.t[r] rel[k]
          const int r = e / 3, k = e % 3;
          float a = dGt[j][r] * (s.J[j][k] - s.J[p][k]);
          for (int cc = 0; cc < 3; ++cc) a += dGR[j][r * 3 + cc] * s.R[j][k * 3 + cc];
          dGR[p][e] += a;
        } else if (e < 12) {
          const int k = e - 9;
          float a = 0.f;
          for (int r = 0; r < 3; ++r) a += s.G[p][r * 4 + k] * dGt[j][r];
          g.dJ[p][k] -= a;
        }
      }
      if (e >= 12) {  // (after the loop above in program order only for this entry: dGt[p] is read by nobody on this level)
        const int r = e - 12;
        for (int j = J - 1; j > p; --j)
          if (s.parent[j] == p) dGt[p][r] += dGt[j][r];
      }
    }
  }
  __syncthreads();
  if (t < 9) g.dR[0][t] = dGR[0][t];
  if (t >= 32 && t < 35) g.dJ[0][t - 32] += dGt[0][t - 32];
  __syncthreads();
  for (int i = t; i < NP; i += blockDim.x) g.dR[i / 9 + 1][i % 9] += g.dpm[i];  // pose_map = R - I
  __syncthreads();
  if (t < J) {  // Rodrigues Jacobian by forward-mode duals
    Dual<3> a[3] = {Dual<3>::var(s.pose[3 * t], 0), Dual<3>::var(s.pose[3 * t + 1], 1), Dual<3>::var(s.pose[3 * t + 2], 2)};
    Dual<3> R[9];
    rodrigues_any<Dual<3>>(c.rodrigues_kind, a, R);
    for (int i = 0; i < 3; ++i) {
      float acc = 0.f;
      for (int k = 0; k < 9; ++k) acc += R[k].d[i] * g.dR[t][k];
      d_pose[3 * t + i] = acc;
    }
  }
  if (t >= 64 && t < 64 + NS) {
    int bi = t - 64;
    float a = g.dbeta[bi];
    for (int i = 0; i < J * 3; ++i) a += c.J_shapedirs[i * NS + bi] * (&g.dJ[0][0])[i];
    d_shape[bi] = a;
  }
}

// One CTA, all vertices (the stand-alone kernel): accumulate into the scratch's tail, then finish.
__device__ __forceinline__ void lbs_backward_core(const harp_ctx& c, LbsShared& s, LbsBwdShared& g, float* dvp, float* dTs, const float* gv,
                                                  const float* dj, float* d_pose, float* d_shape, float* d_trans,
                                                  const float* vposed = nullptr) {
  float* part = dTs + (size_t)12 * c.V0 + (c.V0 * 12 < MAXJ * 12 ? MAXJ * 12 : 0);
  lbs_backward_accumulate(c, s, g.red, dvp, dTs, gv, dj, vposed, 0, c.V0, part);
  __syncthreads();
  lbs_backward_finish(c, s, g, dTs, part, 1, 0, dj, d_pose, d_shape, d_trans);
}

inline size_t lbs_bwd_scratch_floats(const harp_ctx* c) {
  size_t words = (size_t)c->V0 * 15;
  size_t need_dg = (size_t)MAXJ * 12;  // dG accumulators alias dTs
  if ((size_t)c->V0 * 12 < need_dg) words += need_dg;
  return words + (size_t)lbs_partial_floats(c->J, c->n_shape);  // + one block of partial sums behind dTs
}

}  // namespace
