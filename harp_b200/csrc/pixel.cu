// The fused pixel pass of the optimiser step: ONE walk over each tile's face list yields the soft silhouette (value, L1
// loss, gradient) AND the nearest covering face of every pixel; hit pixels are compacted per frame and shaded (forward,
// photometric L1, backward) by a second kernel whose warps are full and barrier-free.
// Replaces, per minibatch: the K=50 and the K=1 camera-view rasterisations (renderer_helper.py:44-58,353), both shaders
// and blends, both image losses (optimize_sequence.py:519,543) and their autograd.
#include "shade_pixel.cuh"

int harp_build_bins(harp_ctx* c, int view, int B, int S, float radius, const float* ndc, cudaStream_t st);
int harp_project(harp_ctx* c, int B, int S, float focal, const float* verts, const float* R, const float* T, float* ndc,
                 cudaStream_t st);
int harp_project_bwd(harp_ctx* c, int B, int S, float focal, const float* verts, const float* R, const float* T,
                     const float* d_ndc, float* d_verts, float* d_R, float* d_T, cudaStream_t st);
int harp_fold4(size_t n, const float* src4, float* dst3, cudaStream_t st);
int harp_tex_prepare(harp_ctx* c, const float* texture, const float* normal_map, int texH, int texW, cudaStream_t st);
int harp_light_depth(harp_ctx* c, int B, int S, cudaStream_t st);
int harp_light_depth_bwd(harp_ctx* c, int B, int S, cudaStream_t st);
int harp_sil_topk(harp_ctx* c, int mode, int S, int K, float sigma, float blur, float radius, float* alpha, const float* gin,
                  float gscale, float inv_n, float* loss, cudaStream_t st);

namespace {

#ifndef HARP_PV_CHUNK
#define HARP_PV_CHUNK 64
#endif
#ifndef HARP_PV_QCAP
#define HARP_PV_QCAP 352
#endif
#ifndef HARP_PV_MINCTAS
#define HARP_PV_MINCTAS 16
#endif
#ifndef HARP_SHADE_GRID
#define HARP_SHADE_GRID 96
#endif
constexpr int CHUNK = HARP_PV_CHUNK;  // faces staged per shared-memory round (tuning knobs: tools/build_variants.sh)
constexpr int NT = HARP_TILE * HARP_TILE;
constexpr int QCAP = HARP_PV_QCAP;  // 352 with the 128-byte face records: 13.5 KB per CTA -> 16 CTAs per SM

struct VisArgs {
  int S, K;
  float sigma, blur, radius;
  const float* ndc;
  const int32_t *bin_count, *bin_start, *bin_list;
  // silhouette
  float* alpha;
  const float* y_sil;
  float sil_gscale, sil_inv_n;
  float* sil_loss;
  float* d_ndc;
  // visibility
  int32_t *hit_list, *hit_count;  // hit_list: (B, S*S) int2 records (row*S+col, face id)
  // background pixels of the colour image
  float* rgb;
  const float *y_true, *y_mask;
  float photo_inv_n;
  float* photo_loss;
};

// Block-wide sum of two values at once (64-thread CTAs: two warps); results valid in thread 0.
__device__ __forceinline__ void block_sum2(float& u, float& v, float* red) {
  u = warp_sum(u); v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[2 * w] = u; red[2 * w + 1] = v; }
  __syncthreads();
  if (threadIdx.x == 0) { u = red[0] + red[2]; v = red[1] + red[3]; }
}

// SIL_BWD: accumulate the silhouette L1 loss and its gradient (COARSE_OPT stages); PHOTO: accumulate the photometric loss
// of background pixels (APP_OPT stages; hit pixels are handled by the deferred shading kernel).
template <bool SIL_BWD, bool PHOTO>
__global__ void __launch_bounds__(NT, HARP_PV_MINCTAS) pixel_vis_kernel(harp_ctx c, VisArgs a) {
  __shared__ FaceTile faces[CHUNK];
  __shared__ float red[32];
  __shared__ unsigned char q_pix[QCAP], q_face[QCAP], q_edge[QCAP];
  __shared__ unsigned short q_next[QCAP];  // per-pixel chain through the queue, in face-list order
  __shared__ float q_sd[QCAP], q_t[QCAP];
  __shared__ int q_n, any_bwd;
  __shared__ float s_g[NT], s_prod[NT];
  __shared__ unsigned char s_need[NT];
  __shared__ float s_px[HARP_TILE], s_py[HARP_TILE];  // NDC centres of the tile's columns / rows (queue drains look pixels up)
  const int S = a.S;
  int b = blockIdx.z, tw = S / HARP_TILE, tid = threadIdx.x;
  int lx = tid % HARP_TILE, ly = tid / HARP_TILE;
  int col = blockIdx.x * HARP_TILE + lx, row = blockIdx.y * HARP_TILE + ly;
  size_t pix = ((size_t)b * S + row) * S + col;
  size_t tglob = (size_t)b * tw * tw + blockIdx.y * tw + blockIdx.x;
  // The tile's targets are requested together with its face count so that, for an empty tile (about 4 of 5 at the
  // reference's framing), one memory latency covers everything it needs; a tile with faces drops them and re-reads at the end.
  float ys = SIL_BWD ? __ldcs(a.y_sil + pix) : 0.f;
  float mk = 0.f;
  f3 y = mk3(0.f, 0.f, 0.f);
  if (PHOTO) { mk = __ldcs(a.y_mask + pix); y = mk3(__ldcs(a.y_true + pix * 3), __ldcs(a.y_true + pix * 3 + 1), __ldcs(a.y_true + pix * 3 + 2)); }
  int n = a.bin_count[tglob];
  if (n == 0) {
    // ---- empty tile: nothing but streaming
    a.alpha[pix] = 0.f;
    st3(a.rgb + pix * 3, mk3(1.f, 1.f, 1.f));
    float s_sil = fabsf(ys), s_ph = PHOTO ? fabsf(mk - y.x * mk) + fabsf(mk - y.y * mk) + fabsf(mk - y.z * mk) : 0.f;
    if (!SIL_BWD && !PHOTO) return;
    if (__ballot_sync(0xffffffffu, s_sil != 0.f || s_ph != 0.f)) {  // rare: target foreground over an empty tile (per warp, no barrier)
      s_sil = warp_sum(s_sil); s_ph = warp_sum(s_ph);
      if ((tid & 31) == 0) {
        if (SIL_BWD && s_sil != 0.f) atomicAdd(a.sil_loss, s_sil * a.sil_inv_n);
        if (PHOTO && s_ph != 0.f) atomicAdd(a.photo_loss, s_ph * a.photo_inv_n);
      }
    }
    return;
  }
  int start = a.bin_start[tglob];
  if (start + n > c.bin_cap) n = max(0, (int)c.bin_cap - start);
  const int32_t* list = a.bin_list + (size_t)b * c.bin_cap + start;
  const float* ndc = a.ndc + (size_t)b * c.V * 3;
  const float sigma = a.sigma, blur = a.blur;
  float px = pix_to_ndc(col, S), py = pix_to_ndc(row, S);
  if (ly == 0) s_px[lx] = px;
  if (lx == 0) s_py[ly] = py;
  float prod = 1.f;
  int nzero = 0, npass = 0, ncand = 0;
  int best_fid = -1;
  float best_q = 0.f;
  bool q_overflow = false;
  const float sat = 17.5f * sigma, far = 1.02f * blur;  // margins of the affine classification (raster.cuh: FaceTile)
  if (tid == 0) { any_bwd = 0; q_n = 0; }
  for (int base = 0; base < n; base += CHUNK) {
    int m = min(CHUNK, n - base);
    __syncthreads();
    if (tid < m) stage_face(c, ndc, list[base + tid], a.radius, faces[tid]);
    if (tid == 0) q_n = 0;
    __syncthreads();
    // A: classification of every (pixel, face) pair, branch-free up to the class: visibility (nearest inside face by
    //    1/depth) for all pixels; silhouette factors until the pixel saturates (one exactly-zero factor => alpha = 1,
    //    zero gradient).  Band pairs go to the tile's queue, chained per pixel.
    int q_first = -1, q_last = -1, q_mine = 0;
    for (int i = 0; i < m; ++i) {
      const FaceTile& f = faces[i];
      bool inside;
      float e0, e1, e2;
      bool inb = in_face_bbox(f, px, py);
      if (!__any_sync(0xffffffffu, inb)) continue;  // the face's bbox misses this warp's 8x4 half of the tile
      int cls = classify_fast<true>(f, px, py, far, sat, inb, inside, e0, e1, e2);
      if (cls == PAIR_REJECT) continue;
      float q;
      if (cls == PAIR_EXACT) {  // a vertex behind the camera: decide with the exact operation sequence
        PairEval e;
        eval_pair_full<false, false>(f, px, py, 0.f, e);
        inside = e.pass;
        q = e.pass ? 1.f / e.pz : 0.f;
        cls = PAIR_FULL;
      } else {
        q = fast_inv_depth(f, e0, e1, e2);
      }
      if (inside) vis_update(c, ndc, px, py, f.fid, q, best_fid, best_q);
      ++ncand;  // upper bound on the faces PyTorch3D would consider (exact top-K decision after the loop)
      if (nzero > 0) continue;
      if (cls == PAIR_SATURATED) { ++npass; ++nzero; continue; }
      int slot = atomicAdd(&q_n, 1);
      if (slot < QCAP) {
        q_pix[slot] = (unsigned char)tid; q_face[slot] = (unsigned char)i;
        if (q_last >= 0) q_next[q_last] = (unsigned short)slot; else q_first = slot;
        q_last = slot;
        ++q_mine;
      } else {
        q_overflow = true;
        PairEval e;
        eval_pair_full<true, true>(f, px, py, blur, e);
        if (!e.pass) continue;
        ++npass;
        float om = 1.f - 1.f / (1.f + expf(fdiv(e.inside ? -e.dist : e.dist, sigma)));
        if (om == 0.f) ++nzero; else prod *= om;
      }
    }
    __syncthreads();
    // B: drain the queue of band pairs with every thread busy
    int qn = min(q_n, QCAP);
    for (int s = tid; s < qn; s += NT) {
      int p = q_pix[s];
      float ppx = s_px[p % HARP_TILE], ppy = s_py[p / HARP_TILE];
      PairEval e;
      eval_pair_full<true, true>(faces[q_face[s]], ppx, ppy, blur, e);
      if (e.pass) { q_sd[s] = e.inside ? -e.dist : e.dist; q_edge[s] = (unsigned char)e.edge; q_t[s] = e.t; }
      else q_edge[s] = 255;
    }
    __syncthreads();
    // C: every pixel folds its own chain in face-list order
    if (nzero == 0) {
      for (int s = q_first; q_mine > 0; --q_mine, s = q_next[s]) {
        if (q_edge[s] == 255) continue;
        ++npass;
        float om = 1.f - 1.f / (1.f + expf(fdiv(q_sd[s], sigma)));
        if (om == 0.f) { ++nzero; break; }
        prod *= om;
      }
    }
  }
  // ---- visibility outputs: face id image + per-frame compacted hit list (warp-aggregated append)
  bool hit = best_fid >= 0;
  {
    unsigned mask = __ballot_sync(0xffffffffu, hit);
    int lane = tid & 31, base = 0;
    if (mask) {
      if (lane == 0) base = atomicAdd(a.hit_count + b, __popc(mask));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (hit) reinterpret_cast<int2*>(a.hit_list)[(size_t)b * S * S + base + __popc(mask & ((1u << lane) - 1u))] = make_int2(row * S + col, best_fid);
    }
  }
  float photo_abs = 0.f;
  if (!hit) {
    st3(a.rgb + pix * 3, mk3(1.f, 1.f, 1.f));  // background colour (renderer_helper.py:430)
    if (PHOTO) {
      float mk = a.y_mask[pix];
      f3 y = ld3(a.y_true + pix * 3);
      photo_abs = fabsf(mk - y.x * mk) + fabsf(mk - y.y * mk) + fabsf(mk - y.z * mk);
    }
  }
  // ---- silhouette value, loss and gradient
  float al = nzero > 0 ? 1.f : 1.f - prod;
  const bool deferred = ncand > a.K;  // more candidates than faces_per_pixel -> exact top-K kernel
  if (deferred) sil_defer_overflow(c, (int)pix);
  else a.alpha[pix] = al;
  float diff = (deferred || !SIL_BWD) ? 0.f : al - a.y_sil[pix];
  if (SIL_BWD || PHOTO) {
    float s_sil = fabsf(diff), s_ph = photo_abs;
    __syncthreads();  // red[] is free
    block_sum2(s_sil, s_ph, red);
    if (tid == 0) {
      if (SIL_BWD && s_sil != 0.f) atomicAdd(a.sil_loss, s_sil * a.sil_inv_n);
      if (PHOTO && s_ph != 0.f) atomicAdd(a.photo_loss, s_ph * a.photo_inv_n);
    }
  }
  if (!SIL_BWD) return;
  float g = diff > 0.f ? a.sil_gscale : (diff < 0.f ? -a.sil_gscale : 0.f);
  bool need = (g != 0.f) && nzero == 0 && prod > 0.f && npass > 0;
  if (need) any_bwd = 1;
  s_g[tid] = g; s_prod[tid] = prod; s_need[tid] = need ? 1 : 0;
  int any_overflow = __syncthreads_or(q_overflow ? 1 : 0);
  if (!any_bwd) return;
  float* d_ndc = a.d_ndc + (size_t)b * c.V * 4;
  if (n <= CHUNK && !any_overflow) {
    int qn = min(q_n, QCAP);
    for (int s = tid; s < qn; s += NT) {
      int p = q_pix[s];
      if (!s_need[p] || q_edge[s] == 255) continue;
      float ppx = s_px[p % HARP_TILE], ppy = s_py[p / HARP_TILE];
      sil_pair_gradient(c, faces[q_face[s]], ppx, ppy, q_sd[s], q_edge[s], q_t[s], s_g[p], s_prod[p], sigma, d_ndc);
    }
    return;
  }
  for (int base = 0; base < n; base += CHUNK) {  // generic path (multi-chunk tiles): re-stage and evaluate per pixel
    int m = min(CHUNK, n - base);
    __syncthreads();
    if (tid < m) stage_face(c, ndc, list[base + tid], a.radius, faces[tid]);
    __syncthreads();
    if (!need) continue;
    for (int i = 0; i < m; ++i) {
      bool inside;
      float e0, e1, e2;
      int cls = classify_fast<true>(faces[i], px, py, far, sat, in_face_bbox(faces[i], px, py), inside, e0, e1, e2);
      if (cls == PAIR_REJECT || cls == PAIR_SATURATED) continue;
      PairEval e;
      eval_pair_full<true, true>(faces[i], px, py, blur, e);
      if (!e.pass) continue;
      sil_pair_gradient(c, faces[i], px, py, e.inside ? -e.dist : e.dist, e.edge, e.t, g, prod, sigma, d_ndc);
    }
  }
}

// Shade the compacted hit pixels of one frame per blockIdx.y; grid-stride over the frame's hit list.
#ifndef HARP_SHADE_MIN_CTAS
#define HARP_SHADE_MIN_CTAS 4
#endif
template <int MODE>
__global__ void __launch_bounds__(128, HARP_SHADE_MIN_CTAS) shade_deferred_kernel(harp_ctx c, ShadeArgs a, const int2* __restrict__ hit_list,
                                                             const int32_t* __restrict__ hit_count) {
  const int S = a.S;
  int b = blockIdx.y;
  int n = hit_count[b];
  const float* ndc = a.ndc_cam + (size_t)b * c.V * 3;
  const int2* list = hit_list + (size_t)b * S * S;
  // per-thread partial sums of the frame scalars live in shared memory (column per thread: conflict-free), not registers
  __shared__ float s_fr[9][128];
#pragma unroll
  for (int k = 0; k < 9; ++k) s_fr[k][threadIdx.x] = 0.f;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    int2 rec = list[idx];
    int rc = rec.x;
    int row = rc / S, col = rc % S;
    size_t pix = (size_t)b * S * S + rc;
    float px = pix_to_ndc(col, S), py = pix_to_ndc(row, S);
    Hit h;
    h.fid = rec.y;
    FaceRec f;
    load_face(c, ndc, h.fid, 0.f, f);
    ShadePre pre;  // vertex attributes, uvs and targets: in flight while the barycentrics are evaluated
    shade_preload<MODE>(c, a, b, pix, h.fid, f.v0, f.v1, f.v2, pre);
    PairEval e;
    eval_pair_full<false, false>(f, px, py, 0.f, e);
    if (e.pass) { h.pz = e.pz; h.b0 = e.b0; h.b1 = e.b1; h.b2 = e.b2; h.w0 = e.w0; h.w1 = e.w1; h.w2 = e.w2; }
    else h.fid = -1;  // the exact test disagrees with the fast classification on a boundary pixel: background
    ShadeOut o;
    shade_pixel<MODE>(c, a, b, pix, px, py, h, o, &pre);
    if (MODE != 1) st3(a.rgb + pix * 3, o.rgb);
    if (MODE != 0) {
      s_fr[8][threadIdx.x] += o.loss_abs;
#pragma unroll
      for (int k = 0; k < 8; ++k) s_fr[k][threadIdx.x] += o.frame[k];
    }
  }
  if (MODE == 0) return;
  float fr[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) fr[k] = s_fr[k][threadIdx.x];
  if (MODE == 2) {
    float s = warp_sum(s_fr[8][threadIdx.x]);
    if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(a.loss, s * a.inv_n);
  }
  shade_flush_frame_scalars(a, b, fr);
}

}  // namespace

extern "C" int harp_render_fused(harp_ctx* c, int B, int S, float focal, const harp_render_args* r, void* stream) {
  HARP_REQUIRE(c && r && B > 0, HARP_ERR_INVALID, "harp_render_fused: null argument");
  HARP_REQUIRE(r->verts && r->normals && r->cam_R && r->cam_T && r->light_pos && r->amb && r->texture && r->normal_map &&
                   r->alpha && r->rgb, HARP_ERR_INVALID, "harp_render_fused: missing input / output tensor");
  HARP_REQUIRE(c->has_uv, HARP_ERR_INVALID, "harp_render_fused: ctx was created without UV tables");
  HARP_REQUIRE(B <= c->maxB && S == c->S, HARP_ERR_CAPACITY, "harp_render_fused: call harp_ctx_reserve(B, S) first");
  HARP_REQUIRE(!r->self_shadow || (r->light_R && r->light_T), HARP_ERR_INVALID, "harp_render_fused: light camera required");
  const bool sil_bwd = r->w_sil != 0.f, photo_bwd = r->w_photo != 0.f;
  HARP_REQUIRE(!sil_bwd || (r->y_sil && r->loss_sil && r->grads.d_verts && r->grads.d_cam_T), HARP_ERR_INVALID,
               "harp_render_fused: silhouette loss needs y_sil, loss_sil, d_verts, d_cam_T");
  HARP_REQUIRE(!photo_bwd || (r->y_true && r->y_mask && r->loss_photo && r->grads.d_verts && r->grads.d_normals), HARP_ERR_INVALID,
               "harp_render_fused: photometric loss needs y_true, y_mask, loss_photo, d_verts, d_normals");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const size_t npx = (size_t)B * S * S, nv4 = (size_t)B * c->V * 4 * sizeof(float), ntex = (size_t)r->texH * r->texW;
  double blur_d = 9.210240366975849 * (double)r->sigma;  // ln(1/1e-4 - 1) sigma (renderer_helper.py:46)
  float blur = (float)blur_d, radius = (float)sqrt(blur_d);
  // Camera branch first (it is the critical path: its kernels are submitted before the side branch's so that they are
  // dispatched first).
  if (!c->side_stream) {
    int lo = 0, hi = 0;
    HARP_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // lo = least priority
    HARP_CHECK_CUDA(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, lo));
    HARP_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    HARP_CHECK_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  }
  cudaStream_t ss = c->side_stream;
  HARP_CHECK_CUDA(cudaEventRecord(c->ev_fork, st));
  harp_prof_begin(c, 6, st);
  if ((rc = harp_project(c, B, S, focal, r->verts, r->cam_R, r->cam_T, c->ndc_cam, st)) != HARP_OK) return rc;
  if ((rc = harp_build_bins(c, 0, B, S, radius, c->ndc_cam, st)) != HARP_OK) return rc;
  if (sil_bwd || photo_bwd) HARP_CHECK_CUDA(cudaMemsetAsync(c->d_ndc_cam, 0, nv4, st));
  harp_prof_end(c, 6, st);
  HARP_CHECK_CUDA(cudaMemsetAsync(c->hit_count, 0, (size_t)B * sizeof(int32_t), st));
  HARP_CHECK_CUDA(cudaMemsetAsync(c->ovf_count, 0, sizeof(int32_t), st));
  VisArgs v{};
  v.S = S; v.K = r->faces_per_pixel > 0 ? r->faces_per_pixel : 50; v.sigma = r->sigma; v.blur = blur; v.radius = radius; v.ndc = c->ndc_cam;
  v.bin_count = c->bin_count; v.bin_start = c->bin_start; v.bin_list = c->bin_list;
  v.alpha = r->alpha; v.y_sil = r->y_sil; v.sil_inv_n = 1.f / ((float)B * S * S); v.sil_gscale = r->w_sil * v.sil_inv_n;
  v.sil_loss = r->loss_sil; v.d_ndc = c->d_ndc_cam;
  v.hit_list = c->hit_list; v.hit_count = c->hit_count;
  v.rgb = r->rgb; v.y_true = r->y_true; v.y_mask = r->y_mask; v.photo_inv_n = 1.f / ((float)B * S * S * 3.f);
  v.photo_loss = r->loss_photo;
  dim3 grid(S / HARP_TILE, S / HARP_TILE, B);
  harp_prof_begin(c, 0, st);
  if (sil_bwd && photo_bwd) pixel_vis_kernel<true, true><<<grid, NT, 0, st>>>(*c, v);
  else if (sil_bwd) pixel_vis_kernel<true, false><<<grid, NT, 0, st>>>(*c, v);
  else if (photo_bwd) pixel_vis_kernel<false, true><<<grid, NT, 0, st>>>(*c, v);
  else pixel_vis_kernel<false, false><<<grid, NT, 0, st>>>(*c, v);
  harp_prof_end(c, 0, st);
  HARP_LAUNCH_CHECK();
  if ((rc = harp_sil_topk(c, sil_bwd ? 2 : 0, S, v.K, r->sigma, blur, radius, r->alpha, r->y_sil, v.sil_gscale, v.sil_inv_n,
                          r->loss_sil, st)) != HARP_OK) return rc;
  // Side branch on the ctx's own (least-priority) stream, concurrent with the camera branch above and joined before the
  // deferred shading: everything only the shading needs — the light view (projection, bins, nearest-depth raster), the
  // packed texel copies, and the zero fills of the shading's gradient accumulators (92 MB at 64 x 512^2).
  // Event fork / join is capturable into a CUDA graph.
  HARP_CHECK_CUDA(cudaStreamWaitEvent(ss, c->ev_fork, 0));
  if (r->self_shadow) {
    if ((rc = harp_project(c, B, S, focal, r->verts, r->light_R, r->light_T, c->ndc_light, ss)) != HARP_OK) return rc;
    if ((rc = harp_build_bins(c, 1, B, S, 0.f, c->ndc_light, ss)) != HARP_OK) return rc;
    if ((rc = harp_light_depth(c, B, S, ss)) != HARP_OK) return rc;
  }
  if ((rc = harp_tex_prepare(c, r->texture, r->normal_map, r->texH, r->texW, ss)) != HARP_OK) return rc;
  if (photo_bwd) {
    HARP_CHECK_CUDA(cudaMemsetAsync(c->d_verts4, 0, nv4, ss));
    HARP_CHECK_CUDA(cudaMemsetAsync(c->d_normals4, 0, nv4, ss));
    HARP_CHECK_CUDA(cudaMemsetAsync(c->d_tex4, 0, 2 * ntex * 4 * sizeof(float), ss));
    if (r->self_shadow) {
      HARP_CHECK_CUDA(cudaMemsetAsync(c->d_ndc_light, 0, nv4, ss));
      HARP_CHECK_CUDA(cudaMemsetAsync(c->d_light_zbuf, 0, npx * sizeof(float), ss));
    }
  }
  HARP_CHECK_CUDA(cudaEventRecord(c->ev_join, ss));
  HARP_CHECK_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));
  ShadeArgs a{};
  a.B = B; a.S = S; a.self_shadow = r->self_shadow; a.texH = r->texH; a.texW = r->texW; a.focal = focal;
  a.inv_n = v.photo_inv_n; a.gscale = r->w_photo * a.inv_n;
  a.verts = r->verts; a.normals = r->normals; a.ndc_cam = c->ndc_cam;
  a.light_R = r->light_R ? r->light_R : r->cam_R; a.light_T = r->light_T ? r->light_T : r->cam_T;
  a.light_pos = r->light_pos; a.amb = r->amb;
  a.tex4 = reinterpret_cast<const float4*>(c->tex4); a.nm4 = a.tex4 + ntex;
  a.light_zbuf = c->light_zbuf; a.rgb = r->rgb; a.gin = r->y_true; a.mask = r->y_mask; a.loss = r->loss_photo;
  if (photo_bwd) {
    a.d_ndc_cam = c->d_ndc_cam; a.d_verts = c->d_verts4; a.d_normals = c->d_normals4;
    a.d_light_R = r->grads.d_light_R; a.d_light_T = r->grads.d_light_T; a.d_light_pos = r->grads.d_light_pos;
    a.d_amb = r->grads.d_amb; a.d_texture = c->d_tex4; a.d_normal_map = c->d_tex4 + ntex * 4; a.d_light_zbuf = c->d_light_zbuf;
  }
  dim3 sgrid(HARP_SHADE_GRID, B);
  harp_prof_begin(c, 1, st);
  if (photo_bwd) shade_deferred_kernel<2><<<sgrid, 128, 0, st>>>(*c, a, reinterpret_cast<const int2*>(c->hit_list), c->hit_count);
  else shade_deferred_kernel<0><<<sgrid, 128, 0, st>>>(*c, a, reinterpret_cast<const int2*>(c->hit_list), c->hit_count);
  harp_prof_end(c, 1, st);
  HARP_LAUNCH_CHECK();
  harp_prof_begin(c, 7, st);
  if (photo_bwd) {
    if ((rc = harp_fold4((size_t)B * c->V, c->d_verts4, r->grads.d_verts, st)) != HARP_OK) return rc;
    if ((rc = harp_fold4((size_t)B * c->V, c->d_normals4, r->grads.d_normals, st)) != HARP_OK) return rc;
    if (r->grads.d_texture && (rc = harp_fold4(ntex, c->d_tex4, r->grads.d_texture, st)) != HARP_OK) return rc;
    if (r->grads.d_normal_map && (rc = harp_fold4(ntex, c->d_tex4 + ntex * 4, r->grads.d_normal_map, st)) != HARP_OK) return rc;
    if (r->self_shadow) {
      if ((rc = harp_light_depth_bwd(c, B, S, st)) != HARP_OK) return rc;
      if ((rc = harp_project_bwd(c, B, S, focal, r->verts, r->light_R, r->light_T, c->d_ndc_light, r->grads.d_verts,
                                 r->grads.d_light_R, r->grads.d_light_T, st)) != HARP_OK) return rc;
    }
  }
  if (sil_bwd || photo_bwd) {
    if ((rc = harp_project_bwd(c, B, S, focal, r->verts, r->cam_R, r->cam_T, c->d_ndc_cam, r->grads.d_verts, r->grads.d_cam_R,
                               r->grads.d_cam_T, st)) != HARP_OK) return rc;
  }
  harp_prof_end(c, 7, st);
  return HARP_OK;
}
