// Projection, tile binning and the soft-silhouette render (forward, backward, fused L1 loss).
// Replaces MeshRasterizer.transform + _C.rasterize_meshes (K=50) + sigmoid_alpha_blend and their autograd
// (renderer/renderer_helper.py:37-58).  The reference materialises 28 B x 50 fragments per pixel
// (367 MB per 512^2 frame); here a pixel's candidate faces live in shared memory and nothing but alpha is written.
#include "raster.cuh"

int harp_raster_setup(harp_ctx* c, int view, int B, int S, float focal, float radius, const float* verts, const float* R,
                      const float* T, cudaStream_t st, bool zero_grads = false);

namespace {

constexpr int TPB = 256;
constexpr int CHUNK = 64;  // faces staged per shared-memory round

// backward of the projection (setup.cu, phase 0): d_ndc (B,V,4: xyz + pad) -> d_verts += , d_R += , d_T +=
__global__ void __launch_bounds__(TPB) project_bwd_kernel(int B, int V, const float* __restrict__ verts, const float* __restrict__ R,
                                                           const float* __restrict__ T, float focal, float half_s,
                                                           const float* __restrict__ d_ndc, float* __restrict__ d_verts,
                                                           float* __restrict__ d_R, float* __restrict__ d_T) {
  __shared__ float red[32];
  int b = blockIdx.y;
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  const float* r = R + (size_t)b * 9;
  const float* t = T + (size_t)b * 3;
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

// ---- soft silhouette -------------------------------------------------------------------------------------------------
// MODE 0 forward; 1 backward with upstream g_alpha; 2 fused L1(target, alpha) loss (forward + backward in one pass).
// The +-1/3 px band pairs that need the full distance computation are a few percent of all (pixel, face) candidates but
// sit in nearly every face a warp visits, so evaluating them in the pixel loop leaves ~6 of 32 lanes active.  Instead the
// convergent classification pass queues them per tile in shared memory and all 64 threads drain the queue.
constexpr int QCAP = 512;

template <int MODE>
__global__ void __launch_bounds__(HARP_TILE* HARP_TILE) sil_kernel(harp_ctx c, int S, int K, float sigma, float blur, float radius,
                                                                   const float* __restrict__ ndc_all,
                                                                   const int32_t* __restrict__ bin_count,
                                                                   const int32_t* __restrict__ bin_start,
                                                                   const int32_t* __restrict__ bin_list, float* __restrict__ alpha,
                                                                   const float* __restrict__ gin, float gscale, float inv_n,
                                                                   float* __restrict__ loss, float* __restrict__ d_ndc_all) {
  constexpr int NT = HARP_TILE * HARP_TILE;
  __shared__ FaceRec faces[CHUNK];
  __shared__ float red[32];
  __shared__ unsigned char q_pix[QCAP], q_face[QCAP], q_edge[QCAP];  // q_edge == 255: pair did not pass
  __shared__ float q_sd[QCAP], q_t[QCAP];
  __shared__ int q_n, any_bwd;
  __shared__ float s_g[NT], s_prod[NT];
  __shared__ unsigned char s_need[NT];
  int b = blockIdx.z, tw = S / HARP_TILE, tid = threadIdx.x;
  int tile = blockIdx.y * tw + blockIdx.x;
  int lx = tid % HARP_TILE, ly = tid / HARP_TILE;
  int col = blockIdx.x * HARP_TILE + lx, row = blockIdx.y * HARP_TILE + ly;
  size_t pix = ((size_t)b * S + row) * S + col;
  size_t tglob = (size_t)b * tw * tw + tile;
  int n = bin_count[tglob];
  int start = bin_start[tglob];
  if (start + n > c.bin_cap) n = max(0, (int)c.bin_cap - start);
  const int32_t* list = bin_list + (size_t)b * c.bin_cap + start;
  const float* ndc = ndc_all + (size_t)b * c.V * 3;
  float px = pix_to_ndc(col, S), py = pix_to_ndc(row, S);
  float prod = 1.f;
  int nzero = 0, npass = 0, ncand = 0;
  bool q_overflow = false;
  const float sat = 17.5f * sigma;
  if (tid == 0) { any_bwd = 0; q_n = 0; }
  for (int base = 0; base < n; base += CHUNK) {
    int m = min(CHUNK, n - base);
    __syncthreads();
    if (tid < m) load_face(c, ndc, list[base + tid], radius, faces[tid]);
    if (tid == 0) q_n = 0;
    __syncthreads();
    // A: convergent classification; a pixel with one exactly-zero factor is final (alpha = 1, zero gradient)
    {
      for (int i = 0; i < m; ++i) {
        int cls = classify_pair<true>(faces[i], px, py, blur, sat);
        if (cls == PAIR_REJECT) continue;
        ++ncand;  // upper bound on the faces PyTorch3D would consider for this pixel (exact top-K decision below)
        if (nzero > 0) continue;
        if (cls == PAIR_SATURATED) { ++npass; ++nzero; continue; }
        int slot = atomicAdd(&q_n, 1);
        if (slot < QCAP) {
          q_pix[slot] = (unsigned char)tid; q_face[slot] = (unsigned char)i;
        } else {  // queue full (never seen at the reference's scales): evaluate in place
          q_overflow = true;
          PairEval e;
          eval_pair_full<true, true>(faces[i], px, py, blur, e);
          if (!e.pass) continue;
          ++npass;
          float om = 1.f - 1.f / (1.f + expf(fdiv(e.inside ? -e.dist : e.dist, sigma)));
          if (om == 0.f) ++nzero; else prod *= om;
        }
      }
    }
    __syncthreads();
    // B: drain the queue with every thread busy
    int qn = min(q_n, QCAP);
    for (int s = tid; s < qn; s += NT) {
      int p = q_pix[s];
      float ppx = pix_to_ndc(blockIdx.x * HARP_TILE + (p % HARP_TILE), S), ppy = pix_to_ndc(blockIdx.y * HARP_TILE + (p / HARP_TILE), S);
      PairEval e;
      eval_pair_full<true, true>(faces[q_face[s]], ppx, ppy, blur, e);
      if (e.pass) { q_sd[s] = e.inside ? -e.dist : e.dist; q_edge[s] = (unsigned char)e.edge; q_t[s] = e.t; }
      else q_edge[s] = 255;
    }
    __syncthreads();
    // C: every pixel folds its own results in queue (= face list) order
    if (nzero == 0) {
      for (int s = 0; s < qn; ++s) {
        if (q_pix[s] != tid || q_edge[s] == 255) continue;
        ++npass;
        float om = 1.f - 1.f / (1.f + expf(fdiv(q_sd[s], sigma)));  // 1 - sigmoid(-sd/sigma)
        if (om == 0.f) { ++nzero; break; }
        prod *= om;
      }
    }
  }
  float a = nzero > 0 ? 1.f : 1.f - prod;
  // more candidates than faces_per_pixel: PyTorch3D keeps only the K nearest in depth -> exact per-pixel kernel
  const bool deferred = ncand > K;
  if (deferred) sil_defer_overflow(c, (int)pix);
  if (MODE != 1 && !deferred) alpha[pix] = a;
  if (MODE == 0) return;
  float g;  // dL/d alpha
  if (MODE == 1) {
    g = deferred ? 0.f : gin[pix];
  } else {
    float tgt = gin[pix];
    float diff = deferred ? 0.f : a - tgt;
    g = diff > 0.f ? gscale : (diff < 0.f ? -gscale : 0.f);
    float s = block_sum(fabsf(diff), red);
    if (tid == 0 && s != 0.f) atomicAdd(loss, s * inv_n);
  }
  bool need = (g != 0.f) && nzero == 0 && prod > 0.f && npass > 0;
  if (need) any_bwd = 1;
  s_g[tid] = g; s_prod[tid] = prod; s_need[tid] = need ? 1 : 0;
  int any_overflow = __syncthreads_or(q_overflow ? 1 : 0);
  if (!any_bwd) return;
  float* d_ndc = d_ndc_all + (size_t)b * c.V * 4;
  if (n <= CHUNK && !any_overflow) {
    // single-chunk tile (the common case): the queue and the staged faces of pass 1 are still valid
    int qn = min(q_n, QCAP);
    for (int s = tid; s < qn; s += NT) {
      int p = q_pix[s];
      if (!s_need[p] || q_edge[s] == 255) continue;
      float ppx = pix_to_ndc(blockIdx.x * HARP_TILE + (p % HARP_TILE), S), ppy = pix_to_ndc(blockIdx.y * HARP_TILE + (p / HARP_TILE), S);
      sil_pair_gradient(c, faces[q_face[s]], ppx, ppy, q_sd[s], q_edge[s], q_t[s], s_g[p], s_prod[p], sigma, d_ndc);
    }
    return;
  }
  for (int base = 0; base < n; base += CHUNK) {  // generic path: re-stage every chunk, pixel-parallel
    int m = min(CHUNK, n - base);
    __syncthreads();
    if (tid < m) load_face(c, ndc, list[base + tid], radius, faces[tid]);
    __syncthreads();
    if (!need) continue;
    for (int i = 0; i < m; ++i) {
      PairEval e;
      eval_pair<true, true>(faces[i], px, py, blur, sat, e);
      if (!e.pass || e.saturated) continue;
      sil_pair_gradient(c, faces[i], px, py, e.inside ? -e.dist : e.dist, e.edge, e.t, g, prod, sigma, d_ndc);
    }
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

// dst (n,3) += src (n,4).xyz
__global__ void __launch_bounds__(TPB) fold4_kernel(size_t n, const float4* __restrict__ src, float* __restrict__ dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v = src[i];
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
  fold4_kernel<<<nblk(n), TPB, 0, st>>>(n, reinterpret_cast<const float4*>(src4), dst3);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

int harp_project_bwd(harp_ctx* c, int B, int S, float focal, const float* verts, const float* R, const float* T,
                     const float* d_ndc, float* d_verts, float* d_R, float* d_T, cudaStream_t st) {
  dim3 grid((c->V + TPB - 1) / TPB, B);
  project_bwd_kernel<<<grid, TPB, 0, st>>>(B, c->V, verts, R, T, focal, 0.5f * (float)S, d_ndc, d_verts, d_R, d_T);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

extern "C" int harp_silhouette(harp_ctx* c, int mode, int B, int S, float focal, float sigma, int faces_per_pixel, const float* verts,
                               const float* R, const float* T, float* alpha, const float* gin, float weight, float* loss,
                               float* d_verts, float* d_R, float* d_T, void* stream) {
  HARP_REQUIRE(c && B > 0 && verts && R && T, HARP_ERR_INVALID, "harp_silhouette: null argument");
  HARP_REQUIRE(mode >= 0 && mode <= 2, HARP_ERR_INVALID, "harp_silhouette: mode must be 0, 1 or 2");
  HARP_REQUIRE(B <= c->maxB && S == c->S, HARP_ERR_CAPACITY, "harp_silhouette: call harp_ctx_reserve(B, S) first");
  HARP_REQUIRE(mode == 1 || alpha, HARP_ERR_INVALID, "harp_silhouette: alpha output required");
  HARP_REQUIRE(mode == 0 || gin, HARP_ERR_INVALID, "harp_silhouette: upstream gradient / target required");
  HARP_REQUIRE(mode != 2 || loss, HARP_ERR_INVALID, "harp_silhouette: loss output required in fused mode");
  cudaStream_t st = (cudaStream_t)stream;
  // blur_radius = ln(1/1e-4 - 1) * sigma (renderer_helper.py:46), evaluated in double like the reference's numpy call
  double blur_d = 9.210240366975849 * (double)sigma;
  float blur = (float)blur_d, radius = (float)sqrt(blur_d);
  int rc;
  if ((rc = harp_raster_setup(c, 0, B, S, focal, radius, verts, R, T, st)) != HARP_OK) return rc;
  if (mode != 0) HARP_CHECK_CUDA(cudaMemsetAsync(c->d_ndc_cam, 0, (size_t)B * c->V * 4 * sizeof(float), st));
  HARP_CHECK_CUDA(cudaMemsetAsync(c->ovf_count, 0, sizeof(int32_t), st));
  const int K = faces_per_pixel > 0 ? faces_per_pixel : 50;
  dim3 grid(S / HARP_TILE, S / HARP_TILE, B);
  float inv_n = 1.f / ((float)B * (float)S * (float)S);
  float gscale = weight * inv_n;
  const int32_t *cnt = c->bin_count, *bst = c->bin_start, *lst = c->bin_list;
  harp_prof_begin(c, 0, st);
  if (mode == 0)
    sil_kernel<0><<<grid, HARP_TILE * HARP_TILE, 0, st>>>(*c, S, K, sigma, blur, radius, c->ndc_cam, cnt, bst, lst, alpha, nullptr, 0.f, 0.f, nullptr, nullptr);
  else if (mode == 1)
    sil_kernel<1><<<grid, HARP_TILE * HARP_TILE, 0, st>>>(*c, S, K, sigma, blur, radius, c->ndc_cam, cnt, bst, lst, alpha, gin, 0.f, 0.f, nullptr, c->d_ndc_cam);
  else
    sil_kernel<2><<<grid, HARP_TILE * HARP_TILE, 0, st>>>(*c, S, K, sigma, blur, radius, c->ndc_cam, cnt, bst, lst, alpha, gin, gscale, inv_n, loss, c->d_ndc_cam);
  harp_prof_end(c, 0, st);
  HARP_LAUNCH_CHECK();
  if ((rc = harp_sil_topk(c, mode, S, K, sigma, blur, radius, alpha, gin, gscale, inv_n, loss, st)) != HARP_OK) return rc;
  if (mode != 0) {
    if ((rc = harp_project_bwd(c, B, S, focal, verts, R, T, c->d_ndc_cam, d_verts, d_R, d_T, st)) != HARP_OK) return rc;
  }
  return HARP_OK;
}
