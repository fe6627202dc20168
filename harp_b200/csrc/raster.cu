// Backward of the projection, the exact top-K silhouette pass for pixels the tile rasteriser (pixel.cu) defers, and the
// fold of the padded gradient accumulators.  With pixel.cu this replaces MeshRasterizer.transform + _C.rasterize_meshes
// (K=50) + sigmoid_alpha_blend and their autograd (renderer/renderer_helper.py:37-58).
#include "raster.cuh"

int harp_raster_setup(harp_ctx* c, int view, int B, int S, float focal, float radius, const float* verts, const float* R,
                      const float* T, cudaStream_t st, bool zero_grads = false);

namespace {

constexpr int TPB = 256;

// backward of the projection (setup.cu, phase 0): d_ndc (B,V,4: xyz + pad) -> d_verts += , d_R += , d_T +=
struct ProjView {
  const float *R, *T, *d_ndc;
  float *d_R, *d_T;
};

// blockIdx.z selects the view (camera / light): both backward projections of the shading step are one launch
__global__ void __launch_bounds__(TPB) project_bwd_kernel(int B, int V, const float* __restrict__ verts, ProjView va, ProjView vb,
                                                           float focal, float half_s, float* __restrict__ d_verts) {
  __shared__ float red[32];
  const ProjView pv = blockIdx.z == 0 ? va : vb;
  const float* __restrict__ d_ndc = pv.d_ndc;
  float* __restrict__ d_R = pv.d_R;
  float* __restrict__ d_T = pv.d_T;
  int b = blockIdx.y;
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  const float* r = pv.R + (size_t)b * 9;
  const float* t = pv.T + (size_t)b * 3;
  float acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.f;
  if (v < V) {
    size_t idx = (size_t)b * V + v;
    f3 g = ld3(d_ndc + idx * 4);
    if (g.x != 0.f || g.y != 0.f || g.z != 0.f) {
      f3 p = ld3(verts + idx * 3);
      float X = p.x * r[0] + p.y * r[3] + p.z * r[6] + t[0];
      float Y = p.x * r[1] + p.y * r[4] + p.z * r[7] + t[1];
      float Z = p.x * r[2] + p.y * r[5] + p.z * r[8] + t[2];
      float k = focal / half_s;
      float gX = g.x * k / Z, gY = g.y * k / Z;
      float gZ = g.z - (g.x * k * X + g.y * k * Y) / (Z * Z);
      if (d_verts) {
        float* dv = d_verts + idx * 3;
        atomicAdd(dv + 0, r[0] * gX + r[1] * gY + r[2] * gZ);
        atomicAdd(dv + 1, r[3] * gX + r[4] * gY + r[5] * gZ);
        atomicAdd(dv + 2, r[6] * gX + r[7] * gY + r[8] * gZ);
      }
      acc[0] = p.x * gX; acc[1] = p.x * gY; acc[2] = p.x * gZ;
      acc[3] = p.y * gX; acc[4] = p.y * gY; acc[5] = p.y * gZ;
      acc[6] = p.z * gX; acc[7] = p.z * gY; acc[8] = p.z * gZ;
      acc[9] = gX; acc[10] = gY; acc[11] = gZ;
    }
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    if (i < 9 && !d_R) continue;
    if (i >= 9 && !d_T) continue;
    float s = block_sum(acc[i], red);
    if (threadIdx.x == 0 && s != 0.f) atomicAdd(i < 9 ? d_R + (size_t)b * 9 + i : d_T + (size_t)b * 3 + (i - 9), s);
  }
}

// Exact PyTorch3D semantics for pixels with more candidate faces than faces_per_pixel: keep the K nearest in depth
// (ties to the lower face id), blend only those.  One warp per deferred pixel; entries live in shared memory.
// Reached only when faces are sub-pixel (tiny images) or very many layers overlap; never at the reference's 448^2.
constexpr int TOPK_CAP = 512;   // passing faces per pixel handled exactly
constexpr int TOPK_WARPS = 4;

template <int MODE>
__global__ void __launch_bounds__(32 * TOPK_WARPS) sil_topk_kernel(harp_ctx c, int S, int K, float sigma, float blur, float radius,
                                                                    const float* __restrict__ ndc_all,
                                                                    const int32_t* __restrict__ bin_count,
                                                                    const int32_t* __restrict__ bin_start,
                                                                    const int32_t* __restrict__ bin_list, float* __restrict__ alpha,
                                                                    const float* __restrict__ gin, float gscale, float inv_n,
                                                                    float* __restrict__ loss, float* __restrict__ d_ndc_all) {
  __shared__ float e_pz[TOPK_WARPS][TOPK_CAP], e_sd[TOPK_WARPS][TOPK_CAP], e_t[TOPK_WARPS][TOPK_CAP];
  __shared__ int e_fid[TOPK_WARPS][TOPK_CAP];
  __shared__ unsigned char e_edge[TOPK_WARPS][TOPK_CAP], e_keep[TOPK_WARPS][TOPK_CAP];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = min(*c.ovf_count, c.ovf_cap);
  const int tw = S / HARP_TILE;
  for (int item = blockIdx.x * TOPK_WARPS + w; item < total; item += gridDim.x * TOPK_WARPS) {
    int pix = c.ovf_list[item];
    int b = pix / (S * S), rem = pix % (S * S), row = rem / S, col = rem % S;
    size_t tglob = (size_t)b * tw * tw + (row / HARP_TILE) * tw + (col / HARP_TILE);
    int n = bin_count[tglob], start = bin_start[tglob];
    if (start + n > c.bin_cap) n = max(0, (int)c.bin_cap - start);
    const int32_t* list = bin_list + (size_t)b * c.bin_cap + start;
    const float* ndc = ndc_all + (size_t)b * c.V * 3;
    float px = pix_to_ndc(col, S), py = pix_to_ndc(row, S);
    int ne = 0;
    for (int base = 0; base < n; base += 32) {
      int i = base + lane;
      bool pass = false;
      PairEval e;
      FaceRec f;
      if (i < n) {
        load_face(c, ndc, list[i], radius, f);
        if (!(px > f.xmax || px < f.xmin || py > f.ymax || py < f.ymin)) {
          eval_pair_full<true, true>(f, px, py, blur, e);
          pass = e.pass;
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, pass);
      int slot = ne + __popc(m & ((1u << lane) - 1u));
      if (pass && slot < TOPK_CAP) {
        e_pz[w][slot] = e.pz; e_sd[w][slot] = e.inside ? -e.dist : e.dist; e_t[w][slot] = e.t;
        e_fid[w][slot] = f.fid; e_edge[w][slot] = (unsigned char)e.edge;
      }
      ne += __popc(m);
    }
    if (ne > TOPK_CAP) { if (lane == 0) atomicAdd(c.counters, 1ULL); ne = TOPK_CAP; }
    __syncwarp();
    // rank by (depth, face id); keep rank < K
    for (int i = lane; i < ne; i += 32) {
      float z = e_pz[w][i]; int fid = e_fid[w][i], rank = 0;
      for (int j = 0; j < ne; ++j) rank += (e_pz[w][j] < z || (e_pz[w][j] == z && e_fid[w][j] < fid)) ? 1 : 0;
      e_keep[w][i] = rank < K ? 1 : 0;
    }
    __syncwarp();
    // product in double: the float result then does not depend on the order of the factors, so this kernel and the tile
    // kernels (pixel.cu) give the same alpha for the same set of faces
    double dprod = 1.0; int nzero = 0;
    for (int i = lane; i < ne; i += 32) {
      if (!e_keep[w][i]) continue;
      float om = 1.f - 1.f / (1.f + expf(fdiv(e_sd[w][i], sigma)));
      if (om == 0.f) ++nzero; else dprod *= (double)om;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { dprod *= __shfl_xor_sync(0xffffffffu, dprod, o); nzero += __shfl_xor_sync(0xffffffffu, nzero, o); }
    float prod = (float)dprod;
    float a = nzero > 0 ? 1.f : 1.f - prod;
    if (MODE != 1 && lane == 0) alpha[pix] = a;
    if (MODE == 0) continue;
    float g;
    if (MODE == 1) g = gin[pix];
    else {
      float diff = a - gin[pix];
      g = diff > 0.f ? gscale : (diff < 0.f ? -gscale : 0.f);
      if (lane == 0 && diff != 0.f) atomicAdd(loss, fabsf(diff) * inv_n);
    }
    if (g == 0.f || nzero > 0 || !(prod > 0.f)) continue;
    float* d_ndc = d_ndc_all + (size_t)b * c.V * 4;
    for (int i = lane; i < ne; i += 32) {
      if (!e_keep[w][i]) continue;
      FaceRec f;
      load_face(c, ndc, e_fid[w][i], radius, f);
      sil_pair_gradient(c, f, px, py, e_sd[w][i], e_edge[w][i], e_t[w][i], g, prod, sigma, d_ndc);
    }
  }
}

// dst (n,3) += src (n,4).xyz; rows n .. 2n-1 of src go to dst_b (the texture and the normal map fold in one launch)
__global__ void __launch_bounds__(TPB) fold4_kernel(size_t n, size_t total, const float4* __restrict__ src, float* __restrict__ dst_a,
                                                    float* __restrict__ dst_b) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float4 v = src[i];
  float* dst = dst_a;
  if (i >= n) { dst = dst_b; i -= n; }
  if (v.x != 0.f || v.y != 0.f || v.z != 0.f) {  // atomics: other branches of the step may accumulate into dst concurrently
    atomicAdd(dst + i * 3, v.x); atomicAdd(dst + i * 3 + 1, v.y); atomicAdd(dst + i * 3 + 2, v.z);
  }
}

inline int nblk(size_t n) { return (int)((n + TPB - 1) / TPB); }

}  // namespace

// Exact top-K pass over the pixels the tile kernels deferred (c->ovf_list); mode as harp_silhouette.
int harp_sil_topk(harp_ctx* c, int mode, int S, int K, float sigma, float blur, float radius, float* alpha, const float* gin,
                  float gscale, float inv_n, float* loss, cudaStream_t st) {
  const int32_t *cnt = c->bin_count, *bst = c->bin_start, *lst = c->bin_list;
  int grid = 2 * HARP_SM_COUNT;
  if (mode == 0) sil_topk_kernel<0><<<grid, 32 * TOPK_WARPS, 0, st>>>(*c, S, K, sigma, blur, radius, c->ndc_cam, cnt, bst, lst, alpha, nullptr, 0.f, 0.f, nullptr, nullptr);
  else if (mode == 1) sil_topk_kernel<1><<<grid, 32 * TOPK_WARPS, 0, st>>>(*c, S, K, sigma, blur, radius, c->ndc_cam, cnt, bst, lst, alpha, gin, 0.f, 0.f, nullptr, c->d_ndc_cam);
  else sil_topk_kernel<2><<<grid, 32 * TOPK_WARPS, 0, st>>>(*c, S, K, sigma, blur, radius, c->ndc_cam, cnt, bst, lst, alpha, gin, gscale, inv_n, loss, c->d_ndc_cam);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

int harp_fold4(size_t n, const float* src4, float* dst3, cudaStream_t st) {
  fold4_kernel<<<nblk(n), TPB, 0, st>>>(n, n, reinterpret_cast<const float4*>(src4), dst3, nullptr);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

// src4 holds two (n,4) blocks back to back: the first folds into dst_a, the second into dst_b
int harp_fold4_pair(size_t n, const float* src4, float* dst_a, float* dst_b, cudaStream_t st) {
  fold4_kernel<<<nblk(2 * n), TPB, 0, st>>>(n, 2 * n, reinterpret_cast<const float4*>(src4), dst_a, dst_b);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

int harp_project_bwd(harp_ctx* c, int B, int S, float focal, const float* verts, const float* R, const float* T,
                     const float* d_ndc, float* d_verts, float* d_R, float* d_T, cudaStream_t st) {
  dim3 grid((c->V + TPB - 1) / TPB, B, 1);
  ProjView v{R, T, d_ndc, d_R, d_T};
  project_bwd_kernel<<<grid, TPB, 0, st>>>(B, c->V, verts, v, v, focal, 0.5f * (float)S, d_verts);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

// camera and light view in one launch (both accumulate into d_verts with atomics)
int harp_project_bwd2(harp_ctx* c, int B, int S, float focal, const float* verts, const float* R0, const float* T0, const float* d_ndc0,
                      float* d_R0, float* d_T0, const float* R1, const float* T1, const float* d_ndc1, float* d_R1, float* d_T1,
                      float* d_verts, cudaStream_t st) {
  dim3 grid((c->V + TPB - 1) / TPB, B, 2);
  project_bwd_kernel<<<grid, TPB, 0, st>>>(B, c->V, verts, ProjView{R0, T0, d_ndc0, d_R0, d_T0}, ProjView{R1, T1, d_ndc1, d_R1, d_T1}, focal,
                                           0.5f * (float)S, d_verts);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}
