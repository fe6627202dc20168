// What the colour renders need besides the per-pixel shading (shade_pixel.cuh, run by shade_deferred_kernel in pixel.cu): the
// packed texel copies (prepare_materials' per-texel F.normalize, utils/visualize.py:98-99, done once per texel) and the
// evaluation pass's normal renderer — MeshRasterizer(K=10) + SoftPhongNormalShader + softmax_rgb_blend
// (renderer/renderer_helper.py:85-101, 216-301; forward only, the reference never differentiates through it).
#include "shade_pixel.cuh"

int harp_raster_setup(harp_ctx* c, int view, int B, int S, float focal, float radius, const float* verts, const float* R,
                      const float* T, cudaStream_t st, bool zero_grads = false);
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
