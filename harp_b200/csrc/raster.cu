// Projection, tile binning and the soft-silhouette render (forward, backward, fused L1 loss).
// Replaces MeshRasterizer.transform + _C.rasterize_meshes (K=50) + sigmoid_alpha_blend and their autograd
// (renderer/renderer_helper.py:37-58).  The reference materialises 28 B x 50 fragments per pixel
// (367 MB per 512^2 frame); here a pixel's candidate faces live in shared memory and nothing but alpha is written.
#include "raster.cuh"

namespace {

constexpr int TPB = 256;
constexpr int CHUNK = 64;  // faces staged per shared-memory round

// ---- projection: X_view = X R + T ; ndc_xy = (f X/Z) / (S/2) ; keep view z   ---------------------------
__global__ void __launch_bounds__(TPB) project_kernel(int B, int V, const float* __restrict__ verts, const float* __restrict__ R,
                                                       const float* __restrict__ T, float focal, float half_s,
                                                       float* __restrict__ ndc) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * V) return;
  int b = idx / V;
  const float* r = R + (size_t)b * 9;
  const float* t = T + (size_t)b * 3;
  f3 p = ld3(verts + (size_t)idx * 3);
  float X = p.x * r[0] + p.y * r[3] + p.z * r[6] + t[0];
  float Y = p.x * r[1] + p.y * r[4] + p.z * r[7] + t[1];
  float Z = p.x * r[2] + p.y * r[5] + p.z * r[8] + t[2];
  st3(ndc + (size_t)idx * 3, mk3((focal * X / Z) / half_s, (focal * Y / Z) / half_s, Z));
}

// backward of project_kernel: d_ndc (B,V,4: xyz + pad) -> d_verts += , d_R += , d_T +=
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

// ---- binning ---------------------------------------------------------------------------------------------------------
// Pixel column j has x_ndc = 1 - (2j+1)/S (decreasing); a face with inflated bbox [lo,hi] can touch columns
// ceil((1-hi) S/2 - 0.5) .. floor((1-lo) S/2 - 0.5); a 0.01 px guard absorbs rounding (the exact test runs per pixel).
// Per-face binning record: the tile range of the inflated bbox plus what the per-tile edge test needs.
struct BinFace {
  int tx0, tx1, ty0, ty1;
  bool edge_test;            // all three view depths > 0: inside/outside is decided by the edge functions alone
  float x0, y0, x1, y1, x2, y2, sgn;
  float thr12, thr20, thr01;  // squared edge length x (1.05 blur + guard): a tile farther than this outside an edge line is skipped
};

__device__ __forceinline__ bool face_tile_range(const harp_ctx& c, const float* ndc, int f, float radius, int S, BinFace& r) {
  const int32_t* fv = c.faces + (size_t)f * 3;
  const float *a = ndc + (size_t)fv[0] * 3, *b = ndc + (size_t)fv[1] * 3, *d = ndc + (size_t)fv[2] * 3;
  float zmax = fmaxf(fmaxf(a[2], b[2]), d[2]);
  float area = edge_fn(a[0], a[1], b[0], b[1], d[0], d[1]);  // EdgeFunctionForward(v0, v1, v2)
  if (zmax < 0.f || (area <= HARP_K_EPS && area >= -HARP_K_EPS)) return false;
  if (!(area == area)) return false;  // NaN guard
  float xmin = fminf(fminf(a[0], b[0]), d[0]) - radius, xmax = fmaxf(fmaxf(a[0], b[0]), d[0]) + radius;
  float ymin = fminf(fminf(a[1], b[1]), d[1]) - radius, ymax = fmaxf(fmaxf(a[1], b[1]), d[1]) + radius;
  float hs = 0.5f * (float)S;
  float jlo = ceilf((1.f - xmax) * hs - 0.5f - 0.01f), jhi = floorf((1.f - xmin) * hs - 0.5f + 0.01f);
  float ilo = ceilf((1.f - ymax) * hs - 0.5f - 0.01f), ihi = floorf((1.f - ymin) * hs - 0.5f + 0.01f);
  jlo = fmaxf(jlo, 0.f); ilo = fmaxf(ilo, 0.f);
  jhi = fminf(jhi, (float)(S - 1)); ihi = fminf(ihi, (float)(S - 1));
  if (jlo > jhi || ilo > ihi) return false;
  r.tx0 = (int)jlo / HARP_TILE; r.tx1 = (int)jhi / HARP_TILE;
  r.ty0 = (int)ilo / HARP_TILE; r.ty1 = (int)ihi / HARP_TILE;
  // Per-tile edge test (only worth it when the bbox spans several tiles).  Pixel pairs of a skipped tile are exactly those
  // classify_pair rejects: every pixel centre lies outside one edge LINE by more than sqrt(1.05 blur) (+1e-5 NDC guard, four
  // orders above the fp32 error of an edge function), and the distance to the triangle is at least the distance to that line.
  r.edge_test = fminf(fminf(a[2], b[2]), d[2]) > 0.f && (r.tx1 > r.tx0 || r.ty1 > r.ty0);
  if (r.edge_test) {
    r.x0 = a[0]; r.y0 = a[1]; r.x1 = b[0]; r.y1 = b[1]; r.x2 = d[0]; r.y2 = d[1];
    r.sgn = fadd(edge_fn(r.x2, r.y2, r.x0, r.y0, r.x1, r.y1), HARP_K_EPS) > 0.f ? 1.f : -1.f;  // sign of FaceRec::area
    float m = 1.05f * radius * radius + 1e-10f;
    r.thr12 = m * ((r.x2 - r.x1) * (r.x2 - r.x1) + (r.y2 - r.y1) * (r.y2 - r.y1));
    r.thr20 = m * ((r.x0 - r.x2) * (r.x0 - r.x2) + (r.y0 - r.y2) * (r.y0 - r.y2));
    r.thr01 = m * ((r.x1 - r.x0) * (r.x1 - r.x0) + (r.y1 - r.y0) * (r.y1 - r.y0));
  }
  return true;
}

// largest signed edge function over the four corner pixel centres of a tile (the function is affine in the pixel position)
__device__ __forceinline__ float tile_max_edge(float sgn, float xa, float xb, float ya, float yb, float ax, float ay, float bx,
                                               float by) {
  float dx = bx - ax, dy = by - ay;
  float ea = (xa - ax) * dy, eb = (xb - ax) * dy, fa = (ya - ay) * dx, fb = (yb - ay) * dx;
  return fmaxf(fmaxf(sgn * (ea - fa), sgn * (ea - fb)), fmaxf(sgn * (eb - fa), sgn * (eb - fb)));
}

__device__ __forceinline__ bool tile_outside_face(const BinFace& r, int tx, int ty, int S) {
  float xa = pix_to_ndc(tx * HARP_TILE, S), xb = pix_to_ndc(tx * HARP_TILE + HARP_TILE - 1, S);
  float ya = pix_to_ndc(ty * HARP_TILE, S), yb = pix_to_ndc(ty * HARP_TILE + HARP_TILE - 1, S);
  float m0 = tile_max_edge(r.sgn, xa, xb, ya, yb, r.x1, r.y1, r.x2, r.y2);  // e0 = EdgeFunction(p, v1, v2)
  if (m0 < 0.f && m0 * m0 >= r.thr12) return true;
  float m1 = tile_max_edge(r.sgn, xa, xb, ya, yb, r.x2, r.y2, r.x0, r.y0);
  if (m1 < 0.f && m1 * m1 >= r.thr20) return true;
  float m2 = tile_max_edge(r.sgn, xa, xb, ya, yb, r.x0, r.y0, r.x1, r.y1);
  return m2 < 0.f && m2 * m2 >= r.thr01;
}

template <bool FILL>
__global__ void __launch_bounds__(TPB) bin_kernel(harp_ctx c, int B, int S, float radius, const float* __restrict__ ndc,
                                                   int32_t* __restrict__ count, const int32_t* __restrict__ start,
                                                   int32_t* __restrict__ fill, int32_t* __restrict__ list) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * c.F) return;
  int b = idx / c.F, f = idx % c.F;
  BinFace r;
  if (!face_tile_range(c, ndc + (size_t)b * c.V * 3, f, radius, S, r)) return;
  int tw = S / HARP_TILE;
  size_t tb = (size_t)b * tw * tw;
  for (int ty = r.ty0; ty <= r.ty1; ++ty)
    for (int tx = r.tx0; tx <= r.tx1; ++tx) {
      if (r.edge_test && tile_outside_face(r, tx, ty, S)) continue;  // deterministic: count and fill passes agree
      size_t t = tb + (size_t)ty * tw + tx;
      if (!FILL) {
        atomicAdd(count + t, 1);
      } else {
        int slot = start[t] + atomicAdd(fill + t, 1);
        if (slot < c.bin_cap) list[(size_t)b * c.bin_cap + slot] = f;
      }
    }
}

// exclusive scan of one frame's tile counts (one CTA per frame)
__global__ void __launch_bounds__(1024) bin_scan_kernel(harp_ctx c, int tiles, const int32_t* __restrict__ count,
                                                         int32_t* __restrict__ start) {
  __shared__ int wsum[32];
  int b = blockIdx.x, t = threadIdx.x;
  int per = (tiles + 1023) / 1024;
  const int32_t* cnt = count + (size_t)b * tiles;
  int32_t* st = start + (size_t)b * tiles;
  int lo = t * per, hi = min(lo + per, tiles);
  int local = 0;
  for (int i = lo; i < hi; ++i) local += cnt[i];
  int lane = t & 31, w = t >> 5;
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    int v = wsum[lane], iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += n;
    }
    wsum[lane] = iv - v;
    if (lane == 31 && iv > c.bin_cap) atomicAdd(c.counters + 1, 1ULL);
  }
  __syncthreads();
  int run = wsum[w] + incl - local;
  for (int i = lo; i < hi; ++i) {
    st[i] = run;
    run += cnt[i];
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
    float prod = 1.f; int nzero = 0;
    for (int i = lane; i < ne; i += 32) {
      if (!e_keep[w][i]) continue;
      float om = 1.f - 1.f / (1.f + expf(fdiv(e_sd[w][i], sigma)));
      if (om == 0.f) ++nzero; else prod *= om;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { prod *= __shfl_xor_sync(0xffffffffu, prod, o); nzero += __shfl_xor_sync(0xffffffffu, nzero, o); }
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

// Build the bin lists of one view (0 = camera, 1 = light) for B frames from already-projected verts.
int harp_build_bins(harp_ctx* c, int view, int B, int S, float radius, const float* ndc, cudaStream_t st) {
  size_t nt = (size_t)B * c->tiles;
  int32_t* count = c->bin_count + (size_t)view * c->maxB * c->tiles;
  int32_t* start = c->bin_start + (size_t)view * c->maxB * c->tiles;
  int32_t* fill = c->bin_fill + (size_t)view * c->maxB * c->tiles;
  int32_t* list = c->bin_list + (size_t)view * c->maxB * c->bin_cap;
  HARP_CHECK_CUDA(cudaMemsetAsync(count, 0, nt * sizeof(int32_t), st));
  HARP_CHECK_CUDA(cudaMemsetAsync(fill, 0, nt * sizeof(int32_t), st));
  bin_kernel<false><<<nblk((size_t)B * c->F), TPB, 0, st>>>(*c, B, S, radius, ndc, count, start, fill, list);
  bin_scan_kernel<<<B, 1024, 0, st>>>(*c, c->tiles, count, start);
  bin_kernel<true><<<nblk((size_t)B * c->F), TPB, 0, st>>>(*c, B, S, radius, ndc, count, start, fill, list);
  HARP_LAUNCH_CHECK_N(3);
  return HARP_OK;
}

int harp_project(harp_ctx* c, int B, int S, float focal, const float* verts, const float* R, const float* T, float* ndc,
                 cudaStream_t st) {
  project_kernel<<<nblk((size_t)B * c->V), TPB, 0, st>>>(B, c->V, verts, R, T, focal, 0.5f * (float)S, ndc);
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
  if ((rc = harp_project(c, B, S, focal, verts, R, T, c->ndc_cam, st)) != HARP_OK) return rc;
  if ((rc = harp_build_bins(c, 0, B, S, radius, c->ndc_cam, st)) != HARP_OK) return rc;
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
