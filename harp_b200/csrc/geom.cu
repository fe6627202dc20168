// The per-frame geometry of the optimiser step as TWO kernels (one CTA per frame, the frame's vertices in shared memory):
//
//   geo_fwd   gather the frame's rows of the parameter tables -> LBS -> SubdivideMeshes + displacement along vertex normals
//             -> vertex normals -> centroid, weak-perspective camera, look-at light camera, sigmoid(amb)
//   geo_bwd   backward of all of the above, down to atomics on the parameter-table gradients
//
// Replaces, per minibatch (optimize_sequence.py:446-579): prepare_mesh (utils/visualize.py:16-88: ManoLayer / SMPLXARM forward,
// SubdivideMeshes, verts_normals, displacement), the camera of render_image (visualize.py:266-270), process_info_for_shadow
// (renderer_helper.py:454-468), sigmoid(amb_ratio) and their autograd graph — formerly 5 + 9 launches of this package plus
// a dozen ATen gathers / scatters, each a few microseconds of work behind tens of microseconds of dependent-launch latency.
#include <cooperative_groups.h>

#include <algorithm>

#include "geom.cuh"
#include "glue.cuh"
#include "lbs.cuh"
#include "mesh.cuh"

namespace {

namespace cg = cooperative_groups;

constexpr int GEO_THREADS = LBS_THREADS;
constexpr int POSE_PAD = MAXJ * 3;
constexpr int PART_PAD = (MAXJ * 12 + MAXPB + MAXS + 3 + 3) & ~3;  // one CTA's block of partial LBS sums

__host__ __device__ inline size_t words(size_t bytes) { return (bytes + 3) / 4; }

// full axis-angle pose of the model from the frame's table rows: MANO [rot | pose45] (utils/visualize.py:42-44); SMPL-X arm
// [global_orient | 20 body joints (zero) | right wrist (joint 21) | jaw, eyes, left hand (zero) | right hand] (body_models.py:2278-2296)
__device__ __forceinline__ float assemble_pose(const harp_ctx& c, const GeoArgs& a, int64_t row, int i) {
  if (c.J == 16) return i < 3 ? a.rot[row * 3 + i] : a.pose[row * 45 + (i - 3)];
  if (i < 3) return a.rot[row * 3 + i];
  if (i >= 63 && i < 66) return a.wrist ? a.wrist[row * 3 + (i - 63)] : 0.f;
  if (i >= 120) return a.pose[row * 45 + (i - 120)];
  return 0.f;
}

// A frame is shared by a thread-block CLUSTER of 1, 2 or 4 CTAs (picked so that B x cluster fills the 148 SMs): every CTA keeps
// complete copies of the frame's vertex arrays in its own shared memory, computes a quarter / half / all of each phase's
// vertices and writes its results into every CTA's copy through distributed shared memory; a cluster barrier separates the
// phases.  (One CTA per frame left 84 of 148 SMs idle at 64 frames and made the step's cost constant below that.)
struct Cluster {
  cg::cluster_group g;
  int rank, n;
  __device__ Cluster() : g(cg::this_cluster()) { rank = (int)g.block_rank(); n = (int)g.num_blocks(); }
  __device__ __forceinline__ void st3_all(float* local, f3 v) const {
    for (int r = 0; r < n; ++r) st3(g.map_shared_rank(local, r), v);
  }
  __device__ __forceinline__ void sync() const { g.sync(); }
  // A peer's shared memory may be written only once that CTA has started: every CTA arrives at kernel entry and waits right
  // before its first remote store (split barrier: the work in between hides it).
  __device__ __forceinline__ void entered_arrive() const { g.barrier_arrive(); }
  __device__ __forceinline__ void entered_wait() const { g.barrier_wait(); }
};

__global__ void __launch_bounds__(GEO_THREADS) geo_fwd_kernel(harp_ctx c, GeoArgs a) {
  extern __shared__ float dyn[];
  LbsShared& s = *reinterpret_cast<LbsShared*>(dyn);
  float* s_pose = dyn + words(sizeof(LbsShared));
  float* v0s = s_pose + POSE_PAD;          // (V0,3) skinned coarse vertices
  float* vs = v0s + (size_t)c.V0 * 3;      // (V,3) final vertices
  float* vsub = vs + (size_t)c.V * 3;      // (V,3) subdivided, undisplaced vertices
  __shared__ float red[32];
  __shared__ float s_csum[4][3];           // centroid partial sums of the cluster's CTAs (rank 0's copy is used)
  const Cluster cl;
  cl.entered_arrive();
  const int b = blockIdx.x / cl.n, t = threadIdx.x, J = c.J;
  const int64_t row = a.fid ? a.fid[b] : b;
  for (int i = t; i < J * 3; i += GEO_THREADS) s_pose[i] = assemble_pose(c, a, row, i);
  if (a.zero_grads) {  // this frame's slices of the step's gradient accumulators (was a memset node at the head of the step)
    const int n = c.V * 3, nj = c.n_joints_out * 3;
    float* zv = const_cast<float*>(a.d_verts) + (size_t)b * n;
    float* zn = const_cast<float*>(a.d_normals) + (size_t)b * n;
    for (int i = cl.rank * GEO_THREADS + t; i < n; i += GEO_THREADS * cl.n) { zv[i] = 0.f; zn[i] = 0.f; }
    if (cl.rank == 0) {
      if (t < nj) const_cast<float*>(a.d_joints)[(size_t)b * nj + t] = 0.f;
      if (t < 9) const_cast<float*>(a.d_light_R)[b * 9 + t] = 0.f;
      if (t < 3) {
        const_cast<float*>(a.d_cam_T)[b * 3 + t] = 0.f;
        const_cast<float*>(a.d_light_pos)[b * 3 + t] = 0.f;
        const_cast<float*>(a.d_light_T)[b * 3 + t] = 0.f;
      }
      if (b == 0 && t == 0) const_cast<float*>(a.d_amb)[0] = 0.f;
    }
  }
  if (a.cam_R && cl.rank == 0 && t < 9) a.cam_R[b * 9 + t] = (t == 0 || t == 4) ? -1.f : (t == 8 ? 1.f : 0.f);
  __syncthreads();
  lbs_prologue(c, s_pose, a.shape, s);  // (every CTA of the cluster: a few hundred flops)
  // ---- LBS: vertices rank, rank + n, ... (interleaved by thread block) -> every CTA's v0s, and HBM
  {
    f3 off = ld3(a.trans + row * 3);
    if (c.center_joint >= 0)  // wrist-centring of SMPLXARM (body_models.py:2342-2343)
      off = off - mk3(s.G[c.center_joint][3], s.G[c.center_joint][7], s.G[c.center_joint][11]);
    float* v0g = a.verts0 + (size_t)b * c.V0 * 3;
    float* vpg = a.vposed ? a.vposed + (size_t)b * c.V0 * 3 : nullptr;
    cl.entered_wait();  // every CTA of the cluster is running: its shared memory may be written from here on
    for (int v = cl.rank * GEO_THREADS + t; v < c.V0; v += GEO_THREADS * cl.n) {
      f3 vp = lbs_v_posed(c, s, v);
      if (vpg) st3(vpg + (size_t)v * 3, vp);  // kept for the backward (saves its pass over the pose blendshapes)
      float T[12];
      lbs_blend(c, s, v, T);
      f3 o = mk3(T[0] * vp.x + T[1] * vp.y + T[2] * vp.z + T[3], T[4] * vp.x + T[5] * vp.y + T[6] * vp.z + T[7],
                 T[8] * vp.x + T[9] * vp.y + T[10] * vp.z + T[11]) + off;
      cl.st3_all(v0s + (size_t)v * 3, o);
      st3(v0g + (size_t)v * 3, o);
    }
    cl.sync();
    if (cl.rank == 0 && t < c.n_joints_out) {
      int src = c.joint_out_src[t];
      f3 p = src < J ? mk3(s.G[src][3], s.G[src][7], s.G[src][11]) + off : ld3(v0s + (size_t)c.tips[src - J] * 3);
      st3(a.joints + ((size_t)b * c.n_joints_out + t) * 3, p);
    }
  }
  // ---- SubdivideMeshes (materialised once in shared memory, by every CTA for itself: two smem reads per vertex) +
  //      displacement along the normals of the undisplaced mesh (utils/visualize.py:48-60)
  SubdivAccessor sub{v0s, c.edges0, c.V0};
  for (int v = t; v < c.V; v += GEO_THREADS) st3(vsub + (size_t)v * 3, sub(v));
  __syncthreads();
  DirectAccessor und{vsub};
  float* vg = a.verts + (size_t)b * c.V * 3;
  for (int v = cl.rank * GEO_THREADS + t; v < c.V; v += GEO_THREADS * cl.n) {
    f3 p = und(v);
    if (a.disps) p = p + a.disps[v] * normalize_eps(vertex_normal_sum(c, und, v), 1e-6f);
    cl.st3_all(vs + (size_t)v * 3, p);
    st3(vg + (size_t)v * 3, p);
  }
  cl.sync();
  DirectAccessor fin{vs};
  float* ng = a.normals + (size_t)b * c.V * 3;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int v = cl.rank * GEO_THREADS + t; v < c.V; v += GEO_THREADS * cl.n) {
    st3(ng + (size_t)v * 3, normalize_eps(vertex_normal_sum(c, fin, v), 1e-6f));
    f3 p = fin(v);
    sx += p.x; sy += p.y; sz += p.z;
  }
  sx = block_sum(sx, red); sy = block_sum(sy, red); sz = block_sum(sz, red);
  if (t == 0) st3(cl.g.map_shared_rank(&s_csum[cl.rank][0], 0), mk3(sx, sy, sz));
  cl.sync();
  if (cl.rank != 0 || t != 0) return;
  // camera / light glue (as frame_glue_fwd_kernel, glue.cu)
  sx = sy = sz = 0.f;
  for (int r = 0; r < cl.n; ++r) { sx += s_csum[r][0]; sy += s_csum[r][1]; sz += s_csum[r][2]; }
  float ctr[3] = {sx / (float)c.V, sy / (float)c.V, sz / (float)c.V};
  const float* cam = a.cam + row * 3;
  a.cam_T[b * 3 + 0] = -cam[1];
  a.cam_T[b * 3 + 1] = -cam[2];
  a.cam_T[b * 3 + 2] = (1.f / ((float)a.S * cam[0] + 1e-9f)) * (2.f * a.focal);  // reciprocal * scalar: torch's scalar / tensor (visualize.py:268)
  float L[3] = {a.light_pos[0], a.light_pos[1], a.light_pos[2]}, R[9], T[3];
  light_camera<float>(L, ctr, R, T);
  for (int i = 0; i < 3; ++i) { a.light_b[b * 3 + i] = L[i]; a.light_T[b * 3 + i] = T[i]; a.centroid[b * 3 + i] = ctr[i]; }
  for (int i = 0; i < 9; ++i) a.light_R[b * 9 + i] = R[i];
  if (b == 0) a.amb[0] = 1.f / (1.f + expf(-a.amb_ratio[0]));
}

__global__ void __launch_bounds__(GEO_THREADS) geo_bwd_kernel(harp_ctx c, GeoArgs a, int scratch_floats) {
  extern __shared__ float dyn[];
  LbsShared& s = *reinterpret_cast<LbsShared*>(dyn);
  LbsBwdShared& g = *reinterpret_cast<LbsBwdShared*>(dyn + words(sizeof(LbsShared)));
  float* s_pose = dyn + words(sizeof(LbsShared)) + words(sizeof(LbsBwdShared));
  float* s_out = s_pose + POSE_PAD;        // d_pose (J*3) | d_shape (n_shape) | d_trans (3)
  float* s_part = s_out + POSE_PAD + MAXS + 4;  // the cluster's partial LBS sums (rank 0's copy is used): 4 x PART_PAD
  float* v0s = s_part + 4 * PART_PAD;      // (V0,3)
  float* dv0 = v0s + (size_t)c.V0 * 3;     // (V0,3) gradient w.r.t. the skinned coarse vertices
  float* A = dv0 + (size_t)c.V0 * 3;       // three (V,3) arrays, later the LBS backward's scratch
  float* Bf = A + (size_t)c.V * 3;
  float* C = Bf + (size_t)c.V * 3;
  __shared__ float gc[3];
  const Cluster cl;
  cl.entered_arrive();
  const int b = blockIdx.x / cl.n, t = threadIdx.x, J = c.J, NS = c.n_shape, V = c.V;
  const int64_t row = a.fid ? a.fid[b] : b;
  for (int i = t; i < J * 3; i += GEO_THREADS) s_pose[i] = assemble_pose(c, a, row, i);
  const float* v0g = a.verts0 + (size_t)b * c.V0 * 3;
  for (int i = t; i < c.V0 * 3; i += GEO_THREADS) v0s[i] = v0g[i];
  __syncthreads();
  lbs_prologue(c, s_pose, a.shape, s);
  cl.entered_wait();  // every CTA of the cluster is running: its shared memory may be written from here on
  // ---- camera / light glue backward (as frame_glue_bwd_kernel, glue.cu): rank 0, result (the centroid's gradient) to everyone
  if (cl.rank == 0 && t < 6) {
    // thread i carries the tangent of input i (light position xyz, centroid xyz) through the look-at construction: the same
    // per-component operations as one thread with six tangents, a third of the dependent chain
    Dual<1> L[3], ct[3], R[9], T[3];
    for (int i = 0; i < 3; ++i) {
      L[i] = Dual<1>(a.light_pos[i]); ct[i] = Dual<1>(a.centroid[b * 3 + i]);
      if (t == i) L[i].d[0] = 1.f;
      if (t == 3 + i) ct[i].d[0] = 1.f;
    }
    light_camera<Dual<1>>(L, ct, R, T);
    float gg = 0.f;
    for (int k = 0; k < 9; ++k) gg += a.d_light_R[b * 9 + k] * R[k].d[0];
    for (int k = 0; k < 3; ++k) gg += a.d_light_T[b * 3 + k] * T[k].d[0];
    if (t < 3 && a.g_light) atomicAdd(a.g_light + t, gg + a.d_light_pos[b * 3 + t]);  // share_light_position: row 0
    const float g3 = __shfl_sync(0x3fu, gg, 3), g4 = __shfl_sync(0x3fu, gg, 4), g5 = __shfl_sync(0x3fu, gg, 5);
    if (t == 0) {
      cl.st3_all(gc, mk3(g3 / (float)V, g4 / (float)V, g5 / (float)V));
      const float* cam = a.cam + row * 3;
      const float* dct = a.d_cam_T + (size_t)b * 3;
      float den = (float)a.S * cam[0] + 1e-9f;
      atomicAdd(a.g_cam + row * 3 + 0, dct[2] * (-(2.f * a.focal) * (float)a.S / (den * den)));
      atomicAdd(a.g_cam + row * 3 + 1, -dct[0]);
      atomicAdd(a.g_cam + row * 3 + 2, -dct[1]);
      if (b == 0 && a.g_amb) {
        float sg = 1.f / (1.f + expf(-a.amb_ratio[0]));
        atomicAdd(a.g_amb, a.d_amb[0] * sg * (1.f - sg));
      }
    }
  }
  // Three (V,3) arrays in shared memory, re-used as the phases go (every gather reads materialised vertices: computing the
  // edge midpoints on the fly inside the gathers was a third of this kernel's instructions).  "split" phases are shared by the
  // cluster's CTAs (results written to every CTA's copy), "local" ones are cheap enough for every CTA to do for itself:
  //   P1  local  A = subdivided vertices;   split  Bf = final (displaced) vertices
  //   P2  split  C = dN = normalize_bwd(N(final), d_normals)
  //   P3  split  A = g1 = d_verts + centroid gradient + gather(dN over final)
  //   P4  local  Bf = subdivided vertices again;   split  C = dNs = normalize_bwd(N(subdivided), disp * g1);  d_disps += n_s . g1
  //   P5  split  A += gather(dNs over subdivided)     (in place: only the owner of a row reads it in this phase)
  //   P6  local  dv0 = subdivision backward of A
  const int vfirst = cl.rank * GEO_THREADS + t, vstep = GEO_THREADS * cl.n;
  SubdivAccessor sub{v0s, c.edges0, c.V0};
  for (int v = t; v < V; v += GEO_THREADS) st3(A + (size_t)v * 3, sub(v));
  __syncthreads();
  DirectAccessor undA{A};
  for (int v = vfirst; v < V; v += vstep) {
    f3 p = undA(v);
    if (a.disps) p = p + a.disps[v] * normalize_eps(vertex_normal_sum(c, undA, v), 1e-6f);
    cl.st3_all(Bf + (size_t)v * 3, p);
  }
  cl.sync();
  DirectAccessor fin{Bf};
  const float* dng = a.d_normals + (size_t)b * V * 3;
  const float4* dn4 = a.d_normals4 ? reinterpret_cast<const float4*>(a.d_normals4) + (size_t)b * V : nullptr;
  for (int v = vfirst; v < V; v += vstep) {
    f3 dn = ld3(dng + (size_t)v * 3);
    if (dn4) { float4 q = dn4[v]; dn = dn + mk3(q.x, q.y, q.z); }
    cl.st3_all(C + (size_t)v * 3, normalize_bwd(vertex_normal_sum(c, fin, v), dn, 1e-6f));
  }
  cl.sync();
  const float* dvg = a.d_verts + (size_t)b * V * 3;
  const f3 gcv = mk3(gc[0], gc[1], gc[2]);
  const float4* dv4 = a.d_verts4 ? reinterpret_cast<const float4*>(a.d_verts4) + (size_t)b * V : nullptr;
  for (int v = vfirst; v < V; v += vstep) {
    f3 dv = ld3(dvg + (size_t)v * 3) + gcv;
    if (dv4) { float4 q = dv4[v]; dv = dv + mk3(q.x, q.y, q.z); }
    cl.st3_all(A + (size_t)v * 3, dv + normal_grad_gather(c, fin, C, v));
  }
  cl.sync();
  if (a.disps) {
    for (int v = t; v < V; v += GEO_THREADS) st3(Bf + (size_t)v * 3, sub(v));
    __syncthreads();
    DirectAccessor undB{Bf};
    for (int v = vfirst; v < V; v += vstep) {
      f3 N = vertex_normal_sum(c, undB, v);
      f3 gv = ld3(A + (size_t)v * 3);
      if (a.g_disps) atomicAdd(a.g_disps + v, dot(normalize_eps(N, 1e-6f), gv));
      cl.st3_all(C + (size_t)v * 3, normalize_bwd(N, a.disps[v] * gv, 1e-6f));
    }
    cl.sync();
    for (int v = vfirst; v < V; v += vstep)
      cl.st3_all(A + (size_t)v * 3, ld3(A + (size_t)v * 3) + normal_grad_gather(c, undB, C, v));
    cl.sync();
  }
  const float* g2 = A;
  for (int v = t; v < c.V0; v += GEO_THREADS) {
    f3 acc = ld3(g2 + (size_t)v * 3);
    if (c.subdivide)
      for (int i = c.v0e_off[v]; i < c.v0e_off[v + 1]; ++i) acc = acc + 0.5f * ld3(g2 + (size_t)(c.V0 + c.v0e_idx[i]) * 3);
    st3(dv0 + (size_t)v * 3, acc);
  }
  __syncthreads();
  // ---- LBS backward: every CTA accumulates over its contiguous (even-aligned) range of coarse vertices into rank 0's block of
  //      partial sums; rank 0 finishes.  Scratch: the three vertex arrays are dead.
  (void)scratch_floats;
  const float* dj = a.d_joints ? a.d_joints + (size_t)b * c.n_joints_out * 3 : nullptr;
  {
    const int per = ((c.V0 + cl.n - 1) / cl.n + 1) & ~1;
    const int vb = min(c.V0, cl.rank * per), ve = min(c.V0, vb + per);
    float* part0 = cl.g.map_shared_rank(s_part, 0) + (size_t)cl.rank * PART_PAD;
    lbs_backward_accumulate(c, s, g.red, A, A + (size_t)c.V0 * 3, dv0, dj, a.vposed ? a.vposed + (size_t)b * c.V0 * 3 : nullptr, vb, ve, part0);
  }
  cl.sync();
  if (cl.rank != 0) return;
  lbs_backward_finish(c, s, g, A + (size_t)c.V0 * 3, s_part, cl.n, PART_PAD, dj, s_out, s_out + POSE_PAD, s_out + POSE_PAD + MAXS);
  __syncthreads();
  // ---- scatter into the parameter-table gradients
  if (J == 16) {
    if (t < 45) atomicAdd(a.g_pose + row * 45 + t, s_out[3 + t]);
  } else {
    if (t < 45) atomicAdd(a.g_pose + row * 45 + t, s_out[120 + t]);
    if (t >= 64 && t < 67) {
      float w = s_out[63 + (t - 64)];
      if (a.g_wrist) atomicAdd(a.g_wrist + row * 3 + (t - 64), w);
      if (a.d_wrist) a.d_wrist[(size_t)b * 3 + (t - 64)] = w;
    }
  }
  if (t >= 96 && t < 99) {
    float r = s_out[t - 96];
    if (a.g_rot) atomicAdd(a.g_rot + row * 3 + (t - 96), r);
    if (a.d_rot) a.d_rot[(size_t)b * 3 + (t - 96)] = r;
    if (a.d_trans) a.d_trans[(size_t)b * 3 + (t - 96)] = s_out[POSE_PAD + MAXS + (t - 96)];
  }
  if (t >= 128 && t < 128 + NS) atomicAdd(a.g_shape + (t - 128), s_out[POSE_PAD + (t - 128)]);
}

size_t geo_fwd_smem(const harp_ctx* c) { return (words(sizeof(LbsShared)) + POSE_PAD + (size_t)c->V0 * 3 + (size_t)c->V * 6) * 4; }
size_t geo_bwd_scratch(const harp_ctx* c) { return std::max((size_t)c->V * 9, (size_t)c->V0 * 3 + lbs_bwd_scratch_floats(c)); }
size_t geo_bwd_smem(const harp_ctx* c) {
  return (words(sizeof(LbsShared)) + words(sizeof(LbsBwdShared)) + 2 * POSE_PAD + MAXS + 4 + 4 * PART_PAD + (size_t)c->V0 * 6 + geo_bwd_scratch(c)) * 4;
}

// cluster size: the largest of 4, 2, 1 with B x size <= #SMs
int geo_cluster(int B) { return B * 4 <= HARP_SM_COUNT ? 4 : (B * 2 <= HARP_SM_COUNT ? 2 : 1); }

template <typename... Args>
int launch_clustered(void (*kernel)(Args...), int B, int cl, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * cl); cfg.blockDim = dim3(GEO_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  HARP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
  return HARP_OK;
}

int configure(const void* fn, size_t bytes, size_t* configured) {
  int dev = 0;
  HARP_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || bytes > configured[dev]) {
    HARP_REQUIRE(bytes <= 227 * 1024, HARP_ERR_CAPACITY, "fused geometry kernels: mesh too large for one CTA's shared memory");
    HARP_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (dev >= 0 && dev < 64) configured[dev] = bytes;
  }
  return HARP_OK;
}

}  // namespace

int harp_geo_fwd(harp_ctx* c, const GeoArgs& a, cudaStream_t st) {
  HARP_REQUIRE(c->has_model && (c->J == 16 || c->J == 55), HARP_ERR_NO_MODEL, "harp_step_fused: ctx needs a MANO or SMPL-X-arm model");
  static size_t configured[64] = {};
  int rc = configure((const void*)geo_fwd_kernel, geo_fwd_smem(c), configured);
  if (rc != HARP_OK) return rc;
  harp_prof_begin(c, 2, st);
  if ((rc = launch_clustered(geo_fwd_kernel, a.B, geo_cluster(a.B), geo_fwd_smem(c), st, *c, a)) != HARP_OK) return rc;
  harp_prof_end(c, 2, st);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

int harp_geo_bwd(harp_ctx* c, const GeoArgs& a, cudaStream_t st) {
  static size_t configured[64] = {};
  int rc = configure((const void*)geo_bwd_kernel, geo_bwd_smem(c), configured);
  if (rc != HARP_OK) return rc;
  harp_prof_begin(c, 3, st);
  if ((rc = launch_clustered(geo_bwd_kernel, a.B, geo_cluster(a.B), geo_bwd_smem(c), st, *c, a, (int)geo_bwd_scratch(c))) != HARP_OK) return rc;
  harp_prof_end(c, 3, st);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}
