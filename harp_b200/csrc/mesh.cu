// Mesh build (midpoint subdivision, area-weighted vertex normals, displacement along normals) and the mesh
// regularisers, forward + analytic backward, all in gather form over static CSR tables (no scatter atomics in the
// build; the regularisers use a handful of global atomics per element).
// Replaces utils/visualize.py:48-64 (SubdivideMeshes, Meshes.verts_normals_padded), optimize_sequence.py:533-539
// (mesh_laplacian_smoothing, mesh_normal_consistency, arap_loss) and loss/kps_loss.py.
#include "mesh.cuh"

namespace {

constexpr int TPB = 256;

// ---- forward ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) build_displace_kernel(harp_ctx c, int B, const float* __restrict__ verts0,
                                                              const float* __restrict__ disps, float* __restrict__ verts) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * c.V) return;
  int b = idx / c.V, v = idx % c.V;
  SubdivAccessor acc{verts0 + (size_t)b * c.V0 * 3, c.edges0, c.V0};
  f3 p = acc(v);
  if (disps) {
    f3 n = normalize_eps(vertex_normal_sum(c, acc, v), 1e-6f);
    p = p + disps[v] * n;
  }
  st3(verts + (size_t)idx * 3, p);
}

__global__ void __launch_bounds__(TPB) vertex_normals_kernel(harp_ctx c, int B, const float* __restrict__ verts,
                                                              float* __restrict__ normals) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * c.V) return;
  int b = idx / c.V, v = idx % c.V;
  DirectAccessor acc{verts + (size_t)b * c.V * 3};
  st3(normals + (size_t)idx * 3, normalize_eps(vertex_normal_sum(c, acc, v), 1e-6f));
}

// ---- backward -----------------------------------------------------------------------------------------------------
// dN = normalize_bwd(N, scale * g)  for the final mesh (scale = 1, g = d_normals)
__global__ void __launch_bounds__(TPB) normals_bwd_dN_kernel(harp_ctx c, int B, const float* __restrict__ verts,
                                                              const float* __restrict__ d_normals, float* __restrict__ dN) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * c.V) return;
  int b = idx / c.V, v = idx % c.V;
  DirectAccessor acc{verts + (size_t)b * c.V * 3};
  st3(dN + (size_t)idx * 3, normalize_bwd(vertex_normal_sum(c, acc, v), ld3(d_normals + (size_t)idx * 3), 1e-6f));
}

// out = d_verts + gather(dN)   on the final mesh
__global__ void __launch_bounds__(TPB) normals_bwd_gather_kernel(harp_ctx c, int B, const float* __restrict__ verts,
                                                                  const float* __restrict__ dN, const float* d_verts, float* out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * c.V) return;
  int b = idx / c.V, v = idx % c.V;
  DirectAccessor acc{verts + (size_t)b * c.V * 3};
  f3 g = ld3(d_verts + (size_t)idx * 3) + normal_grad_gather(c, acc, dN + (size_t)b * c.V * 3, v);
  st3(out + (size_t)idx * 3, g);
}

// displacement backward: d_disps[v] += n_s . g ;  dNs = normalize_bwd(N_s, disp * g)
__global__ void __launch_bounds__(TPB) displace_bwd_kernel(harp_ctx c, int B, const float* __restrict__ verts0,
                                                            const float* __restrict__ disps, const float* __restrict__ g_verts,
                                                            float* __restrict__ dNs, float* __restrict__ d_disps) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * c.V) return;
  int b = idx / c.V, v = idx % c.V;
  SubdivAccessor acc{verts0 + (size_t)b * c.V0 * 3, c.edges0, c.V0};
  f3 N = vertex_normal_sum(c, acc, v);
  f3 g = ld3(g_verts + (size_t)idx * 3);
  f3 n = normalize_eps(N, 1e-6f);
  if (d_disps) atomicAdd(d_disps + v, dot(n, g));
  st3(dNs + (size_t)idx * 3, normalize_bwd(N, disps[v] * g, 1e-6f));
}

// d vs = g_verts + gather(dNs) on the subdivided (undisplaced) mesh
__global__ void __launch_bounds__(TPB) displace_bwd_gather_kernel(harp_ctx c, int B, const float* __restrict__ verts0,
                                                                   const float* __restrict__ dNs, const float* __restrict__ g_verts,
                                                                   float* __restrict__ d_vs) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * c.V) return;
  int b = idx / c.V, v = idx % c.V;
  SubdivAccessor acc{verts0 + (size_t)b * c.V0 * 3, c.edges0, c.V0};
  f3 g = ld3(g_verts + (size_t)idx * 3) + normal_grad_gather(c, acc, dNs + (size_t)b * c.V * 3, v);
  st3(d_vs + (size_t)idx * 3, g);
}

// subdivision backward: d verts0[v] = d vs[v] + 0.5 * sum_{e incident} d vs[V0 + e]
__global__ void __launch_bounds__(TPB) subdivide_bwd_kernel(harp_ctx c, int B, const float* __restrict__ d_vs,
                                                             float* __restrict__ d_verts0) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * c.V0) return;
  int b = idx / c.V0, v = idx % c.V0;
  const float* g = d_vs + (size_t)b * c.V * 3;
  f3 a = ld3(g + (size_t)v * 3);
  if (c.subdivide)
    for (int i = c.v0e_off[v]; i < c.v0e_off[v + 1]; ++i) a = a + 0.5f * ld3(g + (size_t)(c.V0 + c.v0e_idx[i]) * 3);
  st3(d_verts0 + (size_t)idx * 3, a);
}

// ---- regularisers -------------------------------------------------------------------------------------------------
// uniform Laplacian: mean_b mean_v || sum_j v_j / deg - v_i ||   
__global__ void __launch_bounds__(TPB) laplacian_kernel(harp_ctx c, int B, const float* __restrict__ verts, float w,
                                                         float* __restrict__ loss, float* __restrict__ d_verts) {
  __shared__ float red[32];
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  float val = 0.f;
  if (idx < B * c.V) {
    int b = idx / c.V, v = idx % c.V;
    const float* vb = verts + (size_t)b * c.V * 3;
    int s = c.vn_off[v], e = c.vn_off[v + 1];
    float inv = e > s ? 1.f / (float)(e - s) : 0.f;
    f3 L = mk3(0.f, 0.f, 0.f);
    for (int i = s; i < e; ++i) L = L + inv * ld3(vb + (size_t)c.vn_idx[i] * 3);
    L = L - ld3(vb + (size_t)v * 3);
    float n = sqrtf(dot(L, L));
    float scale = 1.f / ((float)c.V * (float)B);
    val = n * scale;
    if (d_verts && n > 0.f) {
      f3 g = (w * scale / n) * L;
      float* db = d_verts + (size_t)b * c.V * 3;
      atomic_add3(db + (size_t)v * 3, -1.f * g);
      f3 gj = inv * g;
      for (int i = s; i < e; ++i) atomic_add3(db + (size_t)c.vn_idx[i] * 3, gj);
    }
  }
  float tot = block_sum(val, red);
  if (threadIdx.x == 0 && tot != 0.f) atomicAdd(loss, tot);
}

// ARAP edge-length preservation vs. a constant reference mesh, mm^2 (loss/arap.py:54-57)
__global__ void __launch_bounds__(TPB) arap_kernel(harp_ctx c, int B, const float* __restrict__ verts,
                                                    const float* __restrict__ ref, float w, float* __restrict__ loss,
                                                    float* __restrict__ d_verts) {
  __shared__ float red[32];
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  float val = 0.f;
  if (idx < B * c.E) {
    int b = idx / c.E, e = idx % c.E;
    int i0 = c.edges[2 * e], i1 = c.edges[2 * e + 1];
    const float* vb = verts + (size_t)b * c.V * 3;
    f3 d = ld3(vb + (size_t)i0 * 3) - ld3(vb + (size_t)i1 * 3);
    f3 r = ld3(ref + (size_t)i0 * 3) - ld3(ref + (size_t)i1 * 3);
    float len = sqrtf(dot(d, d)), rl = sqrtf(dot(r, r));
    float diff = len * 1000.f - rl * 1000.f;
    float scale = 1.f / ((float)c.E * (float)B);
    val = diff * diff * scale;
    if (d_verts && len > 0.f) {
      f3 g = (w * scale * 2.f * diff * 1000.f / len) * d;
      float* db = d_verts + (size_t)b * c.V * 3;
      atomic_add3(db + (size_t)i0 * 3, g);
      atomic_add3(db + (size_t)i1 * 3, -1.f * g);
    }
  }
  float tot = block_sum(val, red);
  if (threadIdx.x == 0 && tot != 0.f) atomicAdd(loss, tot);
}

// normal consistency over every pair of faces sharing an edge: 1 - cos(n0, -n1)   
__global__ void __launch_bounds__(TPB) normal_consistency_kernel(harp_ctx c, int B, const float* __restrict__ verts, float w,
                                                                  float* __restrict__ loss, float* __restrict__ d_verts) {
  __shared__ float red[32];
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  float val = 0.f;
  if (idx < B * c.NP) {
    int b = idx / c.NP, p = idx % c.NP;
    const int32_t* q = c.nc_pairs + (size_t)p * 4;
    const float* vb = verts + (size_t)b * c.V * 3;
    f3 v0 = ld3(vb + (size_t)q[0] * 3), v1 = ld3(vb + (size_t)q[1] * 3);
    f3 pa = ld3(vb + (size_t)q[2] * 3) - v0, pb = ld3(vb + (size_t)q[3] * 3) - v0;
    f3 e = v1 - v0;
    f3 n0 = cross(e, pa), n1 = -1.f * cross(e, pb);
    float l0 = fmaxf(sqrtf(dot(n0, n0)), 1e-8f), l1 = fmaxf(sqrtf(dot(n1, n1)), 1e-8f);
    float cs = dot(n0, n1) / (l0 * l1);
    float scale = 1.f / ((float)c.NP * (float)B);
    val = (1.f - cs) * scale;
    if (d_verts) {
      float gc = -w * scale;  // d/dcos
      f3 g0 = gc * ((1.f / (l0 * l1)) * n1 - (cs / (l0 * l0)) * n0);
      f3 g1 = gc * ((1.f / (l0 * l1)) * n0 - (cs / (l1 * l1)) * n1);
      f3 g1n = -1.f * g1;  // w.r.t. e x pb
      // c = e x p :  dc/de = p x G,  dc/dp = G x e
      f3 de = cross(pa, g0) + cross(pb, g1n);
      f3 dpa = cross(g0, e), dpb = cross(g1n, e);
      float* db = d_verts + (size_t)b * c.V * 3;
      atomic_add3(db + (size_t)q[1] * 3, de);
      atomic_add3(db + (size_t)q[0] * 3, -1.f * (de + dpa + dpb));
      atomic_add3(db + (size_t)q[2] * 3, dpa);
      atomic_add3(db + (size_t)q[3] * 3, dpb);
    }
  }
  float tot = block_sum(val, red);
  if (threadIdx.x == 0 && tot != 0.f) atomicAdd(loss, tot);
}

// root-aligned keypoint anchor, (mm/100)^2 (loss/kps_loss.py:4-17): one warp per frame
// `nj`: joints per frame in joints_m / d_joints (21 for MANO; 22 for the SMPL-X arm, whose first 21 are compared:
// kps_loss(use_arm=True) slices pred_kps[:, :21], loss/kps_loss.py:7-8)
__global__ void kps_kernel(int B, int nj, const float* __restrict__ joints_m, const float* __restrict__ init_mm, float w,
                           float* __restrict__ loss, float* __restrict__ d_joints) {
  int b = blockIdx.x, j = threadIdx.x;
  float val = 0.f;
  f3 g = mk3(0.f, 0.f, 0.f);
  const float* J = joints_m + (size_t)b * nj * 3;
  const float* G = init_mm + (size_t)b * 63;
  float scale = 1.f / (21.f * (float)B);
  if (j < 21) {
    f3 gt = ld3(G + j * 3) - ld3(G);
    f3 pr = 1000.f * (ld3(J + j * 3) - ld3(J));
    f3 d = gt - pr;
    float n = sqrtf(dot(d, d)) / 100.f;
    val = n * n * scale;
    g = (-2.f * w * scale / 1e4f * 1000.f) * d;  // d loss / d J_j (metres)
  }
  float gx = warp_sum(g.x), gy = warp_sum(g.y), gz = warp_sum(g.z);
  val = warp_sum(val);
  if (d_joints && j < 21) {
    f3 o = j == 0 ? g - mk3(gx, gy, gz) : g;
    float* dj = d_joints + ((size_t)b * nj + j) * 3;
    dj[0] += o.x; dj[1] += o.y; dj[2] += o.z;
  }
  if (j == 0) atomicAdd(loss, val);
}

__global__ void __launch_bounds__(TPB) disp_reg_kernel(int V, const float* __restrict__ disps, float w,
                                                        float* __restrict__ loss, float* __restrict__ d_disps) {
  __shared__ float red[32];
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  float val = 0.f;
  if (v < V) {
    float d = disps[v];
    val = d * d;
    if (d_disps) atomicAdd(d_disps + v, 2.f * w * d);
  }
  float tot = block_sum(val, red);
  if (threadIdx.x == 0 && tot != 0.f) atomicAdd(loss, tot);
}

inline int nblk(size_t n) { return (int)((n + TPB - 1) / TPB); }

}  // namespace

extern "C" int harp_mesh_build_fwd(harp_ctx* c, int B, const float* verts0, const float* disps, float* verts,
                                   float* normals, void* stream) {
  HARP_REQUIRE(c && B > 0 && verts0 && verts, HARP_ERR_INVALID, "harp_mesh_build_fwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  harp_prof_begin(c, 4, st);
  build_displace_kernel<<<nblk((size_t)B * c->V), TPB, 0, st>>>(*c, B, verts0, disps, verts);
  if (normals) vertex_normals_kernel<<<nblk((size_t)B * c->V), TPB, 0, st>>>(*c, B, verts, normals);
  harp_prof_end(c, 4, st);
  HARP_LAUNCH_CHECK_N(normals ? 2 : 1);
  return HARP_OK;
}

extern "C" int harp_vertex_normals_fwd(harp_ctx* c, int B, const float* verts, float* normals, void* stream) {
  HARP_REQUIRE(c && B > 0 && verts && normals, HARP_ERR_INVALID, "harp_vertex_normals_fwd: null argument");
  vertex_normals_kernel<<<nblk((size_t)B * c->V), TPB, 0, (cudaStream_t)stream>>>(*c, B, verts, normals);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

extern "C" int harp_vertex_normals_bwd(harp_ctx* c, int B, const float* verts, const float* d_normals, float* d_verts,
                                       void* stream) {
  HARP_REQUIRE(c && B > 0 && verts && d_normals && d_verts, HARP_ERR_INVALID, "harp_vertex_normals_bwd: null argument");
  HARP_REQUIRE(B <= c->maxB, HARP_ERR_CAPACITY, "harp_vertex_normals_bwd: batch exceeds harp_ctx_reserve");
  cudaStream_t st = (cudaStream_t)stream;
  int nb = nblk((size_t)B * c->V);
  normals_bwd_dN_kernel<<<nb, TPB, 0, st>>>(*c, B, verts, d_normals, c->vtmp2);
  normals_bwd_gather_kernel<<<nb, TPB, 0, st>>>(*c, B, verts, c->vtmp2, d_verts, d_verts);  // in place: each thread owns its row
  HARP_LAUNCH_CHECK_N(2);
  return HARP_OK;
}

extern "C" int harp_mesh_build_bwd(harp_ctx* c, int B, const float* verts0, const float* disps, const float* d_verts,
                                   const float* d_normals, float* d_verts0, float* d_disps, void* stream) {
  HARP_REQUIRE(c && B > 0 && verts0 && d_verts && d_verts0, HARP_ERR_INVALID, "harp_mesh_build_bwd: null argument");
  HARP_REQUIRE(B <= c->maxB, HARP_ERR_CAPACITY, "harp_mesh_build_bwd: batch exceeds harp_ctx_reserve");
  cudaStream_t st = (cudaStream_t)stream;
  int nb = nblk((size_t)B * c->V);
  const float* g = d_verts;  // running gradient w.r.t. the final verts
  harp_prof_begin(c, 5, st);
  if (d_normals) {
    build_displace_kernel<<<nb, TPB, 0, st>>>(*c, B, verts0, disps, c->vtmp);  // recompute final verts
    normals_bwd_dN_kernel<<<nb, TPB, 0, st>>>(*c, B, c->vtmp, d_normals, c->vtmp2);
    normals_bwd_gather_kernel<<<nb, TPB, 0, st>>>(*c, B, c->vtmp, c->vtmp2, d_verts, c->vN);
    g = c->vN;
  }
  if (disps) {
    displace_bwd_kernel<<<nb, TPB, 0, st>>>(*c, B, verts0, disps, g, c->vtmp2, d_disps);
    displace_bwd_gather_kernel<<<nb, TPB, 0, st>>>(*c, B, verts0, c->vtmp2, g, c->vtmp);
    g = c->vtmp;
  }
  subdivide_bwd_kernel<<<nblk((size_t)B * c->V0), TPB, 0, st>>>(*c, B, g, d_verts0);
  harp_prof_end(c, 5, st);
  HARP_LAUNCH_CHECK_N(1 + (d_normals ? 3 : 0) + (disps ? 2 : 0));
  return HARP_OK;
}

extern "C" int harp_mesh_losses(harp_ctx* c, int B, const float* verts, const float* ref_verts, const float* joints,
                                const float* init_joints_mm, const float* disps, const float w[5], float* losses,
                                float* d_verts, float* d_joints, float* d_disps, void* stream) {
  HARP_REQUIRE(c && B > 0 && verts && w && losses, HARP_ERR_INVALID, "harp_mesh_losses: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  // a term runs when its inputs are there and its weight is not negative (weight 0: value only)
  const bool on[5] = {w[0] >= 0.f, c->NP > 0 && w[1] >= 0.f, ref_verts && w[2] >= 0.f, joints && init_joints_mm && w[3] >= 0.f,
                      disps && w[4] >= 0.f};
  if (on[0]) laplacian_kernel<<<nblk((size_t)B * c->V), TPB, 0, st>>>(*c, B, verts, w[0], losses + 0, d_verts);
  if (on[1]) normal_consistency_kernel<<<nblk((size_t)B * c->NP), TPB, 0, st>>>(*c, B, verts, w[1], losses + 1, d_verts);
  if (on[2]) arap_kernel<<<nblk((size_t)B * c->E), TPB, 0, st>>>(*c, B, verts, ref_verts, w[2], losses + 2, d_verts);
  if (on[3])
    kps_kernel<<<B, 32, 0, st>>>(B, c->n_joints_out >= 21 ? c->n_joints_out : 21, joints, init_joints_mm, w[3], losses + 3, d_joints);
  if (on[4]) disp_reg_kernel<<<nblk(c->V), TPB, 0, st>>>(c->V, disps, w[4], losses + 4, d_disps);
  HARP_LAUNCH_CHECK_N(on[0] + on[1] + on[2] + on[3] + on[4]);
  return HARP_OK;
}
