// Shadowed PBR colour render, forward + analytic backward (+ fused photometric L1 loss).
// Replaces MeshRendererShadow.forward (renderer/renderer_helper.py:331-412): two K=1 rasterisations, hit-point
// interpolation, light-space depth, 3x3 PCF taps; SoftPhongShaderShadow / phong_shading_with_shadow (472-592): texture
// + normal-map bilinear fetches (TexturesUV), PBRMaterials.apply_normal_map (pbr_materials.py:58-124),
// Lambert lighting (A.7), softmax_rgb_blend at K=1 (A.4); and their autograd graph.
// One pass per pixel, nothing staged in HBM except the light-view depth/face buffers the PCF taps must read.
#include "shade_pixel.cuh"

int harp_raster_setup(harp_ctx* c, int view, int B, int S, float focal, float radius, const float* verts, const float* R,
                      const float* T, cudaStream_t st, bool zero_grads = false);
int harp_project_bwd(harp_ctx* c, int B, int S, float focal, const float* verts, const float* R, const float* T,
                     const float* d_ndc, float* d_verts, float* d_R, float* d_T, cudaStream_t st);
int harp_fold4(size_t n, const float* src4, float* dst3, cudaStream_t st);

int harp_tex_prepare(harp_ctx* c, const float* texture, const float* normal_map, int texH, int texW, cudaStream_t st);

namespace {

constexpr int CHUNK = 64;

// texture (r,g,b) -> (r,g,b,0); normal map m -> (m / max(|m|, 1e-12), +-1 / max(|m|, 1e-12)): prepare_materials'
// F.normalize (utils/visualize.py:99) done once per texel instead of once per tap, and one 16-byte load per tap
__global__ void __launch_bounds__(256) pack_tex_kernel(size_t n, const float* __restrict__ tex, const float* __restrict__ nm,
                                                        float4* __restrict__ tex4, float4* __restrict__ nm4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  tex4[i] = make_float4(tex[i * 3], tex[i * 3 + 1], tex[i * 3 + 2], 0.f);
  float x = nm[i * 3], y = nm[i * 3 + 1], z = nm[i * 3 + 2];
  float l = sqrtf(x * x + y * y + z * z);
  float inv = 1.f / fmaxf(l, 1e-12f);
  nm4[i] = make_float4(x * inv, y * inv, z * inv, l > 1e-12f ? inv : -inv);
}

constexpr int NT = HARP_TILE * HARP_TILE;

__device__ __forceinline__ void nearest_face(const harp_ctx& c, const float* ndc, const int32_t* list, int n, float px,
                                             float py, FaceTile* faces, Hit& h) {
  // ordering by the fast inverse depth q (raster.cuh: classify_fast); the winner is re-evaluated exactly below
  int best = -1;
  float best_q = 0.f;
  for (int base = 0; base < n; base += CHUNK) {
    int m = min(CHUNK, n - base);
    __syncthreads();
    if (threadIdx.x < m) stage_face(c, ndc, list[base + threadIdx.x], 0.f, faces[threadIdx.x]);
    __syncthreads();
    for (int i = 0; i < m; ++i) {
      const FaceTile& f = faces[i];
      bool inside;
      float e0, e1, e2, q;
      bool inb = in_face_bbox(f, px, py);
      if (!__any_sync(0xffffffffu, inb)) continue;  // the face's bbox misses this warp's 8x4 half of the tile
      int cls = classify_fast<false>(f, px, py, 0.f, 0.f, inb, inside, e0, e1, e2);
      if (cls == PAIR_REJECT) continue;
      if (cls == PAIR_EXACT) {
        PairEval e;
        eval_pair_full<false, false>(f, px, py, 0.f, e);
        inside = e.pass;
        q = e.pass ? 1.f / e.pz : 0.f;
      } else {
        q = fast_inv_depth(f, e0, e1, e2);
      }
      if (inside) vis_update(c, ndc, px, py, f.fid, q, best, best_q);
    }
  }
  h.fid = best;
  h.pz = 3.4e38f;
  if (best >= 0) {
    FaceRec f;
    load_face(c, ndc, best, 0.f, f);
    PairEval e;
    eval_pair_full<false, false>(f, px, py, 0.f, e);
    if (e.pass) { h.pz = e.pz; h.b0 = e.b0; h.b1 = e.b1; h.b2 = e.b2; h.w0 = e.w0; h.w1 = e.w1; h.w2 = e.w2; }
    else h.fid = -1;
  }
}

// ---- light-view depth pass (renderer_helper.py:344): nearest z + face id per light pixel ------------------------------
__global__ void __launch_bounds__(NT) light_depth_kernel(harp_ctx c, int S, const float* __restrict__ ndc_all,
                                                         const int32_t* __restrict__ bin_count, const int32_t* __restrict__ bin_start,
                                                         const int32_t* __restrict__ bin_list, float* __restrict__ zbuf,
                                                         int32_t* __restrict__ face) {
  __shared__ FaceTile faces[CHUNK];
  int b = blockIdx.z, tw = S / HARP_TILE;
  int lx = threadIdx.x % HARP_TILE, ly = threadIdx.x / HARP_TILE;
  int col = blockIdx.x * HARP_TILE + lx, row = blockIdx.y * HARP_TILE + ly;
  size_t pix = ((size_t)b * S + row) * S + col;
  size_t tglob = (size_t)b * tw * tw + blockIdx.y * tw + blockIdx.x;
  int n = bin_count[tglob];
  if (n == 0) {  // empty light tile
    zbuf[pix] = -1.f;
    face[pix] = -1;
    return;
  }
  int start = bin_start[tglob];
  if (start + n > c.bin_cap) n = max(0, (int)c.bin_cap - start);
  Hit h;
  nearest_face(c, ndc_all + (size_t)b * c.V * 3, bin_list + (size_t)b * c.bin_cap + start, n, pix_to_ndc(col, S),
               pix_to_ndc(row, S), faces, h);
  zbuf[pix] = h.fid >= 0 ? h.pz : -1.f;
  face[pix] = h.fid;
}

// backward of the light-view depth buffer: zbuf = sum_i b_i z_i at the stored face.
// The gradient image is sparse (only light pixels some PCF tap touched): each CTA scans 1024 pixels with 16-byte loads,
// compacts the active ones into shared memory and then processes them with full warps.
constexpr int LDB_THREADS = 256, LDB_PIX = 4 * LDB_THREADS;

__global__ void __launch_bounds__(LDB_THREADS) light_depth_bwd_kernel(harp_ctx c, int B, int S, const float* __restrict__ ndc_all,
                                                                      const int32_t* __restrict__ face,
                                                                      const float* __restrict__ d_zbuf,
                                                                      float* __restrict__ d_ndc_all) {
  __shared__ unsigned short s_idx[LDB_PIX];
  __shared__ int s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const size_t total = (size_t)B * S * S, base = (size_t)blockIdx.x * LDB_PIX;  // S*S is a multiple of 4: whole float4s
  size_t p0 = base + (size_t)threadIdx.x * 4;
  if (p0 < total) {
    float4 g4 = *reinterpret_cast<const float4*>(d_zbuf + p0);
    float gs[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (gs[k] != 0.f) s_idx[atomicAdd(&s_n, 1)] = (unsigned short)(threadIdx.x * 4 + k);
  }
  __syncthreads();
  const int n = s_n;
  for (int s = threadIdx.x; s < n; s += LDB_THREADS) {
    size_t pix = base + s_idx[s];
    int fid = face[pix];
    if (fid < 0) continue;
    float g = d_zbuf[pix];
    int b = (int)(pix / ((size_t)S * S));
    int rem = (int)(pix % ((size_t)S * S));
    int row = rem / S, col = rem % S;
    const float* ndc = ndc_all + (size_t)b * c.V * 3;
    FaceRec f;
    load_face(c, ndc, fid, 0.f, f);
    float px = pix_to_ndc(col, S), py = pix_to_ndc(row, S);
    PairEval e;
    eval_pair<false, false>(f, px, py, 0.f, 0.f, e);
    if (!e.pass) continue;
    Hit h; h.fid = fid; h.pz = e.pz; h.b0 = e.b0; h.b1 = e.b1; h.b2 = e.b2; h.w0 = e.w0; h.w1 = e.w1; h.w2 = e.w2;
    float gb[3] = {g * f.z0, g * f.z1, g * f.z2};
    float gz[3] = {g * e.b0, g * e.b1, g * e.b2};
    bary_backward(tri_from_ndc(ndc, f.v0, f.v1, f.v2), f.v0, f.v1, f.v2, px, py, h, gb, gz, d_ndc_all + (size_t)b * c.V * 4);
  }
}

// MODE 0 forward; 1 backward with upstream g_rgb; 2 fused photometric L1 on the eroded mask
template <int MODE>
__global__ void __launch_bounds__(NT, (MODE == 0 ? 16 : 10)) shade_kernel(harp_ctx c, ShadeArgs a) {
  __shared__ FaceTile faces[CHUNK];
  __shared__ float red[32];
  int b = blockIdx.z, S = a.S, tw = S / HARP_TILE;
  int lx = threadIdx.x % HARP_TILE, ly = threadIdx.x / HARP_TILE;
  int col = blockIdx.x * HARP_TILE + lx, row = blockIdx.y * HARP_TILE + ly;
  size_t pix = ((size_t)b * S + row) * S + col;
  size_t tglob = (size_t)b * tw * tw + blockIdx.y * tw + blockIdx.x;
  int n = a.bin_count[tglob], start = a.bin_start[tglob];
  if (start + n > c.bin_cap) n = max(0, (int)c.bin_cap - start);
  const float* ndc = a.ndc_cam + (size_t)b * c.V * 3;
  float px = pix_to_ndc(col, S), py = pix_to_ndc(row, S);
  Hit h;
  nearest_face(c, ndc, a.bin_list + (size_t)b * c.bin_cap + start, n, px, py, faces, h);

  ShadeOut o;
  shade_pixel<MODE>(c, a, b, pix, px, py, h, o);
  if (MODE != 1) st3(a.rgb + pix * 3, o.rgb);
  if (MODE == 0) return;
  if (MODE == 2) {
    float s = block_sum(o.loss_abs, red);
    if (threadIdx.x == 0 && s != 0.f) atomicAdd(a.loss, s * a.inv_n);
  }
  int any = __syncthreads_or(o.active ? 1 : 0);
  if (!any) return;
  shade_flush_frame_scalars(a, b, o.frame);
}

// ---- normal renderer: K nearest faces per pixel, normal-map-perturbed normals as colours, softmax_rgb_blend ------------
// (renderer_helper.py:85-101, 216-301; evaluation pass, forward only)
constexpr int NORMAL_KMAX = 16;

__global__ void __launch_bounds__(NT) normal_render_kernel(harp_ctx c, int S, int K, const float* __restrict__ ndc_all,
                                                           const float* __restrict__ normals_all,
                                                           const float* __restrict__ normal_map, int texH, int texW,
                                                           const int32_t* __restrict__ bin_count, const int32_t* __restrict__ bin_start,
                                                           const int32_t* __restrict__ bin_list, float* __restrict__ rgb) {
  __shared__ FaceRec faces[CHUNK];
  int b = blockIdx.z, tw = S / HARP_TILE;
  int lx = threadIdx.x % HARP_TILE, ly = threadIdx.x / HARP_TILE;
  int col = blockIdx.x * HARP_TILE + lx, row = blockIdx.y * HARP_TILE + ly;
  size_t pix = ((size_t)b * S + row) * S + col;
  size_t tglob = (size_t)b * tw * tw + blockIdx.y * tw + blockIdx.x;
  int n = bin_count[tglob], start = bin_start[tglob];
  if (start + n > c.bin_cap) n = max(0, (int)c.bin_cap - start);
  const int32_t* list = bin_list + (size_t)b * c.bin_cap + start;
  const float* ndc = ndc_all + (size_t)b * c.V * 3;
  const float* nb = normals_all + (size_t)b * c.V * 3;
  float px = pix_to_ndc(col, S), py = pix_to_ndc(row, S);
  // K smallest depths, kept sorted ascending (ties: lower face id first, as the z-sorted fragments)
  float kz[NORMAL_KMAX]; int kf[NORMAL_KMAX]; int cnt = 0;
  for (int base = 0; base < n; base += CHUNK) {
    int m = min(CHUNK, n - base);
    __syncthreads();
    if (threadIdx.x < m) load_face(c, ndc, list[base + threadIdx.x], 0.f, faces[threadIdx.x]);
    __syncthreads();
    for (int i = 0; i < m; ++i) {
      PairEval e;
      eval_pair<false, false>(faces[i], px, py, 0.f, 0.f, e);
      if (!e.pass) continue;
      int fid = faces[i].fid;
      int pos = cnt;
      while (pos > 0 && (kz[pos - 1] > e.pz || (kz[pos - 1] == e.pz && kf[pos - 1] > fid))) --pos;
      if (pos >= K) continue;
      int last = min(cnt, K - 1);
      for (int j = last; j > pos; --j) { kz[j] = kz[j - 1]; kf[j] = kf[j - 1]; }
      kz[pos] = e.pz; kf[pos] = fid;
      cnt = min(cnt + 1, K);
    }
  }
  // softmax_rgb_blend: background white, sigma = gamma = 1e-4, znear 1, zfar 100
  f3 acc = mk3(0.f, 0.f, 0.f);
  float wsum = 0.f, zinv_max = 1e-10f;
  for (int k = 0; k < cnt; ++k) zinv_max = fmaxf(zinv_max, (100.f - kz[k]) / 99.f);
  for (int k = 0; k < cnt; ++k) {
    FaceRec f;
    load_face(c, ndc, kf[k], 0.f, f);
    PairEval e;
    eval_pair_full<false, true>(f, px, py, 0.f, e);
    f3 N = e.b0 * ld3(nb + (size_t)f.v0 * 3) + e.b1 * ld3(nb + (size_t)f.v1 * 3) + e.b2 * ld3(nb + (size_t)f.v2 * 3);
    if (normal_map) {
      const float* fu = c.face_uvs + (size_t)f.fid * 6;
      Taps tp;
      make_taps(e.b0 * fu[0] + e.b1 * fu[2] + e.b2 * fu[4], e.b0 * fu[1] + e.b1 * fu[3] + e.b2 * fu[5], texH, texW, tp);
      f3 m = mk3(0.f, 0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 4; ++t) m = m + tp.w[t] * normalize12(ld3(normal_map + (size_t)tp.idx[t] * 3));
      float s = N.z >= 0.f ? 1.f : -1.f, aa = -1.f / (s + N.z), bb = N.x * N.y * aa;
      f3 u = mk3(1.f + s * N.x * N.x * aa, s * bb, -s * N.x), v = mk3(bb, s + N.y * N.y * aa, -N.y);
      N = normalize12((-m.x) * u + (-m.y) * v + m.z * N);
    }
    f3 colr = mk3((N.x + 1.f) / 2.f, (-N.y + 1.f) / 2.f, (-N.z + 1.f) / 2.f);
    float prob = 1.f / (1.f + expf(-e.dist / 1e-4f));  // signed distance of an inside pixel is -dist
    float w = prob * expf(((100.f - kz[k]) / 99.f - zinv_max) / 1e-4f);
    acc = acc + w * colr;
    wsum += w;
  }
  float delta = fmaxf(expf((1e-10f - zinv_max) / 1e-4f), 1e-10f);
  float den = wsum + delta;
  st3(rgb + pix * 3, mk3((acc.x + delta) / den, (acc.y + delta) / den, (acc.z + delta) / den));
}

}  // namespace

extern "C" int harp_render_normals(harp_ctx* c, int B, int S, float focal, int K, const float* verts, const float* normals,
                                   const float* R, const float* T, const float* normal_map, int texH, int texW, float* rgb,
                                   void* stream) {
  HARP_REQUIRE(c && B > 0 && verts && normals && R && T && rgb, HARP_ERR_INVALID, "harp_render_normals: null argument");
  HARP_REQUIRE(K >= 1 && K <= NORMAL_KMAX, HARP_ERR_INVALID, "harp_render_normals: faces_per_pixel must be in [1,16]");
  HARP_REQUIRE(!normal_map || c->has_uv, HARP_ERR_INVALID, "harp_render_normals: a normal map needs a ctx with UV tables");
  HARP_REQUIRE(B <= c->maxB && S == c->S, HARP_ERR_CAPACITY, "harp_render_normals: call harp_ctx_reserve(B, S) first");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = harp_raster_setup(c, 0, B, S, focal, 0.f, verts, R, T, st)) != HARP_OK) return rc;
  dim3 grid(S / HARP_TILE, S / HARP_TILE, B);
  normal_render_kernel<<<grid, NT, 0, st>>>(*c, S, K, c->ndc_cam, normals, normal_map, texH, texW, c->bin_count, c->bin_start,
                                            c->bin_list, rgb);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

// (Re)allocate the packed texel buffers for a texH x texW map (one-time) and refresh the packed copies.
int harp_tex_prepare(harp_ctx* c, const float* texture, const float* normal_map, int texH, int texW, cudaStream_t st) {
  size_t ntex = (size_t)texH * texW;
  if ((int64_t)ntex > c->texcap) {
    if (c->d_tex4) cudaFree(c->d_tex4);
    if (c->tex4) cudaFree(c->tex4);
    c->d_tex4 = c->tex4 = nullptr; c->texcap = 0;
    HARP_CHECK_CUDA(cudaMalloc((void**)&c->d_tex4, 2 * ntex * 4 * sizeof(float)));
    HARP_CHECK_CUDA(cudaMalloc((void**)&c->tex4, 2 * ntex * 4 * sizeof(float)));
    c->texcap = (int64_t)ntex;
  }
  pack_tex_kernel<<<(int)((ntex + 255) / 256), 256, 0, st>>>(ntex, texture, normal_map, reinterpret_cast<float4*>(c->tex4),
                                                             reinterpret_cast<float4*>(c->tex4) + ntex);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

// light-view depth pass / its backward on the ctx scratch (projected light-view verts and bins must be current)
int harp_light_depth(harp_ctx* c, int B, int S, cudaStream_t st) {
  dim3 grid(S / HARP_TILE, S / HARP_TILE, B);
  size_t voff = (size_t)c->maxB * c->tiles;
  light_depth_kernel<<<grid, NT, 0, st>>>(*c, S, c->ndc_light, c->bin_count + voff, c->bin_start + voff,
                                          c->bin_list + (size_t)c->maxB * c->bin_cap, c->light_zbuf, c->light_face);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

int harp_light_depth_bwd(harp_ctx* c, int B, int S, cudaStream_t st) {
  size_t npx = (size_t)B * S * S;
  light_depth_bwd_kernel<<<(int)((npx + LDB_PIX - 1) / LDB_PIX), LDB_THREADS, 0, st>>>(*c, B, S, c->ndc_light, c->light_face, c->d_light_zbuf,
                                                                   c->d_ndc_light);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

extern "C" int harp_shade(harp_ctx* c, int mode, int B, int S, float focal, int self_shadow, const float* verts,
                          const float* normals, const float* cam_R, const float* cam_T, const float* light_R,
                          const float* light_T, const float* light_pos, const float* amb, const float* texture,
                          const float* normal_map, int texH, int texW, const float* fixed_lights, float* rgb, const float* gin,
                          const float* mask, float weight, float* loss, const harp_shade_grads* grads, void* stream) {
  HARP_REQUIRE(c && B > 0 && verts && normals && cam_R && cam_T && light_pos && amb && texture && normal_map,
               HARP_ERR_INVALID, "harp_shade: null argument");
  HARP_REQUIRE(c->has_uv, HARP_ERR_INVALID, "harp_shade: ctx was created without UV tables");
  HARP_REQUIRE(mode >= 0 && mode <= 2, HARP_ERR_INVALID, "harp_shade: mode must be 0, 1 or 2");
  HARP_REQUIRE(B <= c->maxB && S == c->S, HARP_ERR_CAPACITY, "harp_shade: call harp_ctx_reserve(B, S) first");
  HARP_REQUIRE(!self_shadow || (light_R && light_T), HARP_ERR_INVALID, "harp_shade: light_R/light_T required for self_shadow");
  HARP_REQUIRE(mode == 1 || rgb, HARP_ERR_INVALID, "harp_shade: rgb output required");
  HARP_REQUIRE(mode == 0 || (gin && grads && grads->d_verts && grads->d_normals), HARP_ERR_INVALID,
               "harp_shade: upstream gradient / target and d_verts, d_normals required");
  HARP_REQUIRE(mode != 2 || (mask && loss), HARP_ERR_INVALID, "harp_shade: mask and loss required in fused mode");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  size_t npx = (size_t)B * S * S;
  if ((rc = harp_raster_setup(c, 0, B, S, focal, 0.f, verts, cam_R, cam_T, st)) != HARP_OK) return rc;
  dim3 grid(S / HARP_TILE, S / HARP_TILE, B);
  if (self_shadow) {
    if ((rc = harp_raster_setup(c, 1, B, S, focal, 0.f, verts, light_R, light_T, st)) != HARP_OK) return rc;
    size_t voff = (size_t)c->maxB * c->tiles;
    light_depth_kernel<<<grid, NT, 0, st>>>(*c, S, c->ndc_light, c->bin_count + voff, c->bin_start + voff,
                                            c->bin_list + (size_t)c->maxB * c->bin_cap, c->light_zbuf, c->light_face);
    HARP_LAUNCH_CHECK();
  }
  ShadeArgs a{};
  a.B = B; a.S = S; a.self_shadow = self_shadow; a.texH = texH; a.texW = texW; a.focal = focal;
  if (fixed_lights) { a.fixed_lights = 1; a.ka = fixed_lights[0]; a.kd = fixed_lights[1]; a.ks = fixed_lights[2]; }
  a.inv_n = 1.f / ((float)B * (float)S * (float)S * 3.f);
  a.gscale = weight * a.inv_n;
  a.verts = verts; a.normals = normals; a.ndc_cam = c->ndc_cam; a.light_R = light_R ? light_R : cam_R;
  a.light_T = light_T ? light_T : cam_T; a.light_pos = light_pos; a.amb = amb;
  a.light_zbuf = c->light_zbuf; a.bin_count = c->bin_count; a.bin_start = c->bin_start; a.bin_list = c->bin_list;
  a.rgb = rgb; a.gin = gin; a.mask = mask; a.loss = loss;
  size_t ntex = (size_t)texH * texW;
  if ((rc = harp_tex_prepare(c, texture, normal_map, texH, texW, st)) != HARP_OK) return rc;
  a.tex4 = reinterpret_cast<const float4*>(c->tex4); a.nm4 = a.tex4 + ntex;
  if (mode != 0) {
    size_t nv4 = (size_t)B * c->V * 4 * sizeof(float);
    HARP_CHECK_CUDA(cudaMemsetAsync(c->d_ndc_cam, 0, nv4, st));
    HARP_CHECK_CUDA(cudaMemsetAsync(c->d_verts4, 0, nv4, st));
    HARP_CHECK_CUDA(cudaMemsetAsync(c->d_normals4, 0, nv4, st));
    HARP_CHECK_CUDA(cudaMemsetAsync(c->d_tex4, 0, 2 * ntex * 4 * sizeof(float), st));
    if (self_shadow) {
      HARP_CHECK_CUDA(cudaMemsetAsync(c->d_ndc_light, 0, nv4, st));
      HARP_CHECK_CUDA(cudaMemsetAsync(c->d_light_zbuf, 0, npx * sizeof(float), st));
    }
    a.d_ndc_cam = c->d_ndc_cam; a.d_verts = c->d_verts4; a.d_normals = c->d_normals4;
    a.d_light_R = grads->d_light_R; a.d_light_T = grads->d_light_T; a.d_light_pos = grads->d_light_pos;
    a.d_amb = grads->d_amb; a.d_texture = c->d_tex4; a.d_normal_map = c->d_tex4 + ntex * 4;
    a.d_light_zbuf = c->d_light_zbuf;
  }
  harp_prof_begin(c, 1, st);
  if (mode == 0) shade_kernel<0><<<grid, NT, 0, st>>>(*c, a);
  else if (mode == 1) shade_kernel<1><<<grid, NT, 0, st>>>(*c, a);
  else shade_kernel<2><<<grid, NT, 0, st>>>(*c, a);
  harp_prof_end(c, 1, st);
  HARP_LAUNCH_CHECK();
  if (mode != 0) {
    if ((rc = harp_fold4((size_t)B * c->V, c->d_verts4, grads->d_verts, st)) != HARP_OK) return rc;
    if ((rc = harp_fold4((size_t)B * c->V, c->d_normals4, grads->d_normals, st)) != HARP_OK) return rc;
    if (grads->d_texture && (rc = harp_fold4(ntex, c->d_tex4, grads->d_texture, st)) != HARP_OK) return rc;
    if (grads->d_normal_map && (rc = harp_fold4(ntex, c->d_tex4 + ntex * 4, grads->d_normal_map, st)) != HARP_OK) return rc;
    if (self_shadow) {
      light_depth_bwd_kernel<<<(int)((npx + LDB_PIX - 1) / LDB_PIX), LDB_THREADS, 0, st>>>(*c, B, S, c->ndc_light, c->light_face, c->d_light_zbuf,
                                                                       c->d_ndc_light);
      HARP_LAUNCH_CHECK();
      if ((rc = harp_project_bwd(c, B, S, focal, verts, light_R, light_T, c->d_ndc_light, grads->d_verts, grads->d_light_R,
                                 grads->d_light_T, st)) != HARP_OK) return rc;
    }
    if ((rc = harp_project_bwd(c, B, S, focal, verts, cam_R, cam_T, c->d_ndc_cam, grads->d_verts, grads->d_cam_R,
                               grads->d_cam_T, st)) != HARP_OK) return rc;
  }
  return HARP_OK;
}
