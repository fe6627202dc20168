// The fused pixel pass of the optimiser step: ONE walk over each tile's face list yields the soft silhouette (value, L1
// loss, gradient) AND the nearest covering face of every pixel; hit pixels are compacted per frame and shaded (forward,
// photometric L1, backward) by a second kernel whose warps are full and barrier-free.
// Replaces, per minibatch: the K=50 and the K=1 camera-view rasterisations (renderer_helper.py:44-58,353), the light-view
// rasterisation (344), both shaders and blends, both image losses (optimize_sequence.py:519,543) and their autograd.
// Entry points: harp_render_fused (both images, fused losses), harp_silhouette and harp_shade (one image each; forward,
// upstream-gradient backward or fused loss) — all three are render_pass with different plans — and harp_hit_mask.
#include "shade_pixel.cuh"

int harp_raster_setup(harp_ctx* c, int view, int B, int S, float focal, float radius, const float* verts, const float* R,
                      const float* T, cudaStream_t st, bool zero_grads = false);
int harp_project_bwd(harp_ctx* c, int B, int S, float focal, const float* verts, const float* R, const float* T,
                     const float* d_ndc, float* d_verts, float* d_R, float* d_T, cudaStream_t st);
int harp_fold4(size_t n, const float* src4, float* dst3, cudaStream_t st);
int harp_fold4_pair(size_t n, const float* src4, float* dst_a, float* dst_b, cudaStream_t st);
int harp_project_bwd2(harp_ctx* c, int B, int S, float focal, const float* verts, const float* R0, const float* T0, const float* d_ndc0,
                      float* d_R0, float* d_T0, const float* R1, const float* T1, const float* d_ndc1, float* d_R1, float* d_T1,
                      float* d_verts, cudaStream_t st);
int harp_tex_prepare(harp_ctx* c, const float* texture, const float* normal_map, int texH, int texW, cudaStream_t st);
int harp_light_depth_tiles(harp_ctx* c, int B, int S, bool zero_grad, cudaStream_t st);
int harp_light_view_setup(harp_ctx* c, int B, int S, float focal, const float* verts, const float* R, const float* T, bool zero_grad,
                          cudaStream_t st);
int harp_light_depth_bwd_tiles(harp_ctx* c, int B, int S, cudaStream_t st);
int harp_sil_topk(harp_ctx* c, int mode, int S, int K, float sigma, float blur, float radius, float* alpha, const float* gin,
                  float gscale, float inv_n, float* loss, cudaStream_t st);

namespace {

#ifndef HARP_PV_MINCTAS
#define HARP_PV_MINCTAS 14
#endif
#ifndef HARP_SHADE_GRID
#define HARP_SHADE_GRID 96
#endif
#ifndef HARP_PV_CHUNK
#define HARP_PV_CHUNK 20  // 15.4 KB of shared memory and <= 72 registers: 14 resident CTAs per SM (32 faces: 11 CTAs, 0.499 -> 0.479 ms)
#endif
constexpr int CHUNK = HARP_PV_CHUNK;   // faces staged per bulk-copy round (one lane of warp 0 per face: at most 32)
static_assert(CHUNK >= 8 && CHUNK <= 32, "one lane of warp 0 per staged face");
constexpr int NT = HARP_TILE * HARP_TILE;
#ifndef HARP_PV_QCAP
#define HARP_PV_QCAP 256
#endif
constexpr int QCAP = HARP_PV_QCAP;   // band pairs of one tile (all chunks); a tile that needs more goes to the exact per-pixel kernel
constexpr int TIECAP = 128; // visibility candidates within 1e-5 of each other (resolved with the exact depth)

// ---- Blackwell/Hopper async-copy plumbing: mbarrier + bulk global->shared copies (TMA unit; SASS: UBLKCP / SYNCS) ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct VisArgs {
  int S, K;
  float sigma, blur, radius;
  const float* ndc;
  const FaceTile* rec;       // (B,F) records of this view
  const int32_t* bin_list;   // (B,bin_cap)
  const int4* work;          // non-empty tiles: heavy from the front, light from the back
  int32_t* work_ctl;         // {heavy, light, cursor, -}
  int work_cap;
  // silhouette
  float* alpha;          // may be null (gradient-only pass)
  const float* y_sil;    // target of the fused L1 loss ...
  const float* g_alpha;  // ... or an upstream dL/dalpha image (then no loss is accumulated)
  float sil_gscale, sil_inv_n;
  float* sil_loss;
  float* d_ndc;
  // visibility
  int32_t *hit_list, *hit_count;  // hit_list: (B, S*S) int2 records (row << 16 | col, face id); null: silhouette only
  // background pixels of the colour image (null: silhouette only / gradient-only pass)
  float* rgb;
  const float *y_true, *y_mask;
  float photo_inv_n;
  float* photo_loss;
  // light view (MODE_LIGHT): nearest depth / face id, and the zero fill of the depth-gradient image on non-empty tiles
  float* zbuf;
  int32_t* face;
  float* d_zbuf;
};

struct __align__(16) BandEntry {
  int fid;
  float sd;              // signed squared distance (inside negative)
  float t;               // clamped parameter on the closest edge (-1: degenerate edge)
  short next;            // chain through the pixel's entries (-1 ends)
  unsigned char pix;
  unsigned char edge;    // queued: chunk-local face index | 0x40 inside | 0x80 exact path; evaluated: 0..2, or 255 = does not pass
};
// s_om[slot] = 1 - sigmoid(-sd / sigma) of an evaluated entry, computed with full warps in the drain
// a visibility near-tie candidate: pixel (6 bits) | face id << 6

// Block-wide sum of two values at once (64-thread CTAs: two warps); results valid in thread 0.
__device__ __forceinline__ void block_sum2(float& u, float& v, float* red) {
  u = warp_sum(u); v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[2 * w] = u; red[2 * w + 1] = v; }
  __syncthreads();
  if (threadIdx.x == 0) { u = red[0] + red[2]; v = red[1] + red[3]; }
}

// Distance part of CheckPixelInsideFace for a face in front of the camera (pz > 0 is implied, strict-inside is known from the
// bit-exact edge functions): PointTriangleDistanceForward, operation by operation.
__device__ __forceinline__ void band_dist(const FaceTile& f, float px, float py, float& dist, int& edge, float& t) {
  float ta, tb, tc;
  float d01 = point_seg_dist(px, py, f.x0, f.y0, f.x1, f.y1, &ta);
  float d02 = point_seg_dist(px, py, f.x0, f.y0, f.x2, f.y2, &tb);
  float d12 = point_seg_dist(px, py, f.x1, f.y1, f.x2, f.y2, &tc);
  dist = d01; edge = 0; t = ta;
  if (d02 < dist) { dist = d02; edge = 1; t = tb; }
  if (d12 < dist) { dist = d12; edge = 2; t = tc; }
}

// sil_pair_gradient for a face known by id only (the tile's staged records are gone by the time gradients are scattered)
__device__ __forceinline__ void sil_pair_gradient_fid(const harp_ctx& c, const float* ndc, int fid, float px, float py, float sd, int edge,
                                                      float t, float g, float prod, float sigma, float* d_ndc) {
  float prob = 1.f / (1.f + expf(fdiv(sd, sigma)));
  float gsd = g * (-prod * prob / sigma);
  float gd = sd < 0.f ? -gsd : gsd;
  if (gd == 0.f) return;
  const int32_t* fv = c.faces + (size_t)fid * 3;
  int va = edge == 2 ? fv[1] : fv[0], vb = edge == 0 ? fv[1] : fv[2];
  const float *pa = ndc + (size_t)va * 3, *pb = ndc + (size_t)vb * 3;
  float ax = pa[0], ay = pa[1], bx = pb[0], by = pb[1];
  if (t < 0.f) {  // degenerate edge: dist = |p - b|^2
    atomicAdd(reinterpret_cast<float2*>(d_ndc + (size_t)vb * 4), make_float2(gd * 2.f * (bx - px), gd * 2.f * (by - py)));
    return;
  }
  float qx = ax + t * (bx - ax), qy = ay + t * (by - ay);
  float dx = 2.f * gd * (qx - px), dy = 2.f * gd * (qy - py);
  atomicAdd(reinterpret_cast<float2*>(d_ndc + (size_t)va * 4), make_float2((1.f - t) * dx, (1.f - t) * dy));
  atomicAdd(reinterpret_cast<float2*>(d_ndc + (size_t)vb * 4), make_float2(t * dx, t * dy));
}

// exact perspective-correct depth of pixel (px,py) on face `fid`, from the face's record in HBM/L2
__device__ __forceinline__ float exact_pz_rec(const FaceTile* rec, int fid, float px, float py, PairEval* out = nullptr) {
  const float4* r = reinterpret_cast<const float4*>(rec + fid);
  float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r7 = __ldg(r + 7);
  FaceRec g;
  g.x0 = r0.x; g.y0 = r0.y; g.x1 = r0.z; g.y1 = r0.w; g.x2 = r1.x; g.y2 = r1.y;
  g.area = r2.w; g.z0 = r7.x; g.z1 = r7.y; g.z2 = r7.z;
  PairEval e;
  eval_pair_full<false, false>(g, px, py, 0.f, e);
  if (out) *out = e;
  return e.pass ? e.pz : 3.4e38f;
}

// The persistent tile rasteriser.  One 64-thread CTA takes non-empty 8x8 tiles from the view's work queue (heavy tiles first).
// A tile's face records arrive by bulk async copies, 32 faces per round, the NEXT tile's first round being requested before
// the current tile is processed.  Work is spread over (pixel, face) PAIRS, not pixels: each face contributes only the pixels
// of its bbox inside the tile (a 5-7 px square for the reference's meshes), the pairs of a round are numbered by a prefix
// sum and every lane classifies one pair per step, so lanes do useful tests only (the former pixel-major walk tested all 64
// pixels of the tile against every face).  Visibility is a 64-bit shared-memory atomicMax on (1/depth, ~face id); silhouette
// band pairs are queued, evaluated with full warps, chained per pixel and folded in ascending face id (deterministic alpha).
//   SIL     soft silhouette (band pairs) next to the visibility;  CAM: camera-view outputs (hit compaction, background
//           colour);  !SIL && !CAM: light view, nearest depth + face id only.  The fitting step runs <SIL, CAM>; the
//           stand-alone Silhouette op runs the same with no hit list / colour image, the stand-alone Shade op <!SIL, CAM>
//   SIL_BWD the silhouette gradient: of the fused L1 loss (COARSE_OPT stages) or, when a.g_alpha is set, of an upstream
//           dL/dalpha image (the autograd drop-in)
//   PHOTO   accumulate the photometric loss of background pixels (APP_OPT stages)
template <bool SIL, bool SIL_BWD, bool PHOTO, bool CAM>
__global__ void __launch_bounds__(NT, HARP_PV_MINCTAS) tile_raster_kernel(harp_ctx c, VisArgs a) {
  __shared__ FaceTile s_rec[2][CHUNK];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ unsigned long long s_vis[NT];
  __shared__ BandEntry s_q[SIL ? QCAP : 1];
  __shared__ float s_om[SIL ? QCAP : 1];
  __shared__ int s_tie[TIECAP];
  __shared__ int s_cnt[NT], s_head[NT];
  __shared__ int s_pre[33], s_rect[32];
  __shared__ unsigned short s_span[CHUNK * HARP_TILE], s_upre[CHUNK * HARP_TILE + 2], s_tmp[CHUNK * HARP_TILE];  // (face, row) spans of the round
  __shared__ unsigned char s_win[CHUNK * 2 + 2];  // unit of the first pair of every 32-pair window
  __shared__ int s_scan0, s_nu;
  __shared__ float s_g[NT], s_prod[NT];
  // a pixel with one exactly-zero factor has alpha = 1 and no gradient.  s_sat is set by the classification (phase A) and read
  // by the drain (phase B) after a barrier; s_sat2 is set by the drain; the epilogue reads both.  Phase A also PEEKS at s_sat
  // while other lanes set it, to keep band pairs of already-saturated pixels out of the queue (without it, tiles where fingers
  // overlap the palm overflow the queue and fall back to the exact top-K kernel: +0.4 ms per step).  That read is a deliberate,
  // benign race — the flag only moves 0 -> 1 and a stale 0 merely queues an entry the drain then skips — and the only access
  // compute-sanitizer's racecheck reports in this library (profiles/r02_compute_sanitizer.txt).
  __shared__ unsigned char s_need[NT], s_sat[NT], s_sat2[NT];
  __shared__ float s_px[HARP_TILE], s_py[HARP_TILE];
  __shared__ float red[4];
  __shared__ int s_qn, s_tn, s_next;
  __shared__ unsigned char s_ovfp[NT];  // pixels with a band pair that did not fit the queue: only THEY go to the exact top-K kernel
  const int S = a.S, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, tw = S / HARP_TILE;
  const float sigma = a.sigma, blur = a.blur;
  const float sat = 17.5f * sigma, far = 1.02f * blur;  // margins of the affine classification (raster.cuh: FaceTile)
  const int n_heavy = a.work_ctl[0], total = n_heavy + a.work_ctl[1];
  if (tid == 0) {
    mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_next = atomicAdd(a.work_ctl + 2, 1);
  }
  __syncthreads();
  int idx = s_next;
  __syncthreads();  // (thread 0 overwrites s_next at the top of the tile loop)
  if (idx >= total) return;
  auto load_item = [&](int i) { return a.work[i < n_heavy ? i : a.work_cap - 1 - (i - n_heavy)]; };
  // one bulk copy per face of the round, issued by warp 0; completion is counted in bytes on the buffer's mbarrier
  auto issue = [&](const int4& it, int base, int buf) {
    if (warp == 0) {
      int m = min(CHUNK, it.w - base);
      if (lane == 0) mbar_expect_tx(&s_bar[buf], (uint32_t)(m * sizeof(FaceTile)));
      __syncwarp();
      if (lane < m) {
        int f = a.bin_list[(size_t)it.x * c.bin_cap + it.z + base + lane];
        bulk_g2s(&s_rec[buf][lane], a.rec + (size_t)it.x * c.F + f, (uint32_t)sizeof(FaceTile), &s_bar[buf]);
      }
    }
  };
  int4 item = load_item(idx);
  int cur = 0;
  uint32_t phase0 = 0, phase1 = 0;
  issue(item, 0, 0);
  while (true) {
    // ---- per-tile state; fetch the next work item and request its first round of records
    s_vis[tid] = 0ULL; s_cnt[tid] = 0; s_head[tid] = -1; s_sat[tid] = 0; s_sat2[tid] = 0;
    const int b = item.x, tile = item.y, n = item.w;
    const int tx0 = (tile % tw) * HARP_TILE, ty0 = (tile / tw) * HARP_TILE;
    if (tid < HARP_TILE) s_px[tid] = pix_to_ndc(tx0 + tid, S);
    else if (tid < 2 * HARP_TILE) s_py[tid - HARP_TILE] = pix_to_ndc(ty0 + tid - HARP_TILE, S);
    s_ovfp[tid] = 0;
    if (tid == 0) { s_qn = 0; s_tn = 0; s_next = atomicAdd(a.work_ctl + 2, 1); }
    __syncthreads();
    const int nidx = s_next;
    const bool have_next = nidx < total;
    int4 nitem = make_int4(0, 0, 0, 0);
    if (have_next) { nitem = load_item(nidx); issue(nitem, 0, cur ^ 1); }
    const float* ndc = a.ndc + (size_t)b * c.V * 3;
    const bool count_cand = SIL && n > a.K;  // a pixel can have more than K candidates only if the tile has more than K faces
    int q_lo = 0;
    for (int base = 0; base < n; base += CHUNK) {
      const int m = min(CHUNK, n - base);
      if (base > 0) { __syncthreads(); issue(item, base, cur); }  // the buffer is free: everyone is past the previous round
      if (cur == 0) { mbar_wait(&s_bar[0], phase0); phase0 ^= 1; } else { mbar_wait(&s_bar[1], phase1); phase1 ^= 1; }
      const FaceTile* recs = s_rec[cur];
      // ---- (1) each face's pixel rectangle inside the tile: its rows are the UNITS of this round (warp 0, lane = face)
      if (warp == 0) {
        int h = 0, rc = 0;
        if (lane < m) {
          int cols = recs[lane].cols, rows = recs[lane].rows;
          int x0 = max(cols & 0xffff, tx0) - tx0, x1 = min(cols >> 16, tx0 + HARP_TILE - 1) - tx0;
          int y0 = max(rows & 0xffff, ty0) - ty0, y1 = min(rows >> 16, ty0 + HARP_TILE - 1) - ty0;
          int w = x1 - x0 + 1;
          h = (w > 0 && y1 >= y0) ? y1 - y0 + 1 : 0;
          rc = x0 | (y0 << 4) | (max(w, 0) << 8);
        }
        int incl = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int x = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += x;
        }
        s_pre[lane + 1] = incl;
        if (lane == 0) s_pre[0] = 0;
        s_rect[lane] = rc;
      }
      __syncthreads();
      // ---- (2) per (face, row): the span of columns that can matter — inside every edge line inflated by the band half-width
      //      (exactly the pairs the classification below would not reject; triangles fill a quarter of their bounding boxes,
      //      so this removes two thirds of the pair tests).  Each thread owns `per` consecutive units; non-empty spans are
      //      compacted and their pixel counts prefix-summed in one block scan (pairs in the low, units in the high half-word).
      {
        const int U = s_pre[m];
        const int per = (U + NT - 1) / NT;  // <= 4: at most 32 faces x 8 rows
        const int u0 = tid * per, u1 = min(U, u0 + per);
        int packed = 0;
        if (u0 < u1) {
          int lo = 0, hi = m;  // face of unit u0: the last face whose unit offset is <= u0
          while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pre[mid] <= u0) lo = mid; else hi = mid; }
          int fi = lo;
          const float hs = 0.5f * (float)S, sq = SIL ? sqrtf(far) * 1.001f : 0.f;
#pragma unroll 1
          for (int u = u0; u < u1; ++u) {
            while (s_pre[fi + 1] <= u) ++fi;  // (faces without rows have equal offsets and are skipped)
            const FaceTile& f = recs[fi];
            const int rc = s_rect[fi], rx0 = rc & 15, rw = rc >> 8;
            const int ry = ((rc >> 4) & 15) + (u - s_pre[fi]);
            int c0 = rx0, c1 = rx0 + rw - 1;  // tile-local column range, inclusive
            if (f.sgn != 0.f) {               // (a face crossing the camera plane keeps its whole rectangle)
              const float py = s_py[ry];
              float xlo = f.xmin, xhi = f.xmax;
              bool empty = false;
#pragma unroll 1
              for (int e = 0; e < 3; ++e) {
                // keep x with  a x + (b py + c) > -M,  M = sqrt(far) |edge| (+ the affine form's rounding bound)
                const float4 ed = reinterpret_cast<const float4*>(&f)[3 + e];  // a, b, c, 1 / |edge|^2
                float base = fmaf(ed.y, py, ed.z) + sq * rsqrtf(ed.w) + 2.f * f.eps;
                if (fabsf(ed.x) > 1e-12f) {
                  float xb = __fdividef(-base, ed.x);
                  if (ed.x > 0.f) xlo = fmaxf(xlo, xb); else xhi = fminf(xhi, xb);
                } else if (base <= 0.f) {
                  empty = true;
                }
              }
              // pixel column j has x = 1 - (2j+1)/S (decreasing): columns ceil((1-xhi) S/2 - 0.5) .. floor((1-xlo) S/2 - 0.5),
              // widened by 0.01 px (the exact tests run per pair)
              float fj0 = ceilf((1.f - xhi) * hs - 0.5f - 0.01f), fj1 = floorf((1.f - xlo) * hs - 0.5f + 0.01f);
              fj0 = fmaxf(fj0, (float)(tx0 + c0)); fj1 = fminf(fj1, (float)(tx0 + c1));
              if (empty || !(fj0 <= fj1)) { c1 = c0 - 1; } else { c0 = (int)fj0 - tx0; c1 = (int)fj1 - tx0; }
            }
            const int len = max(0, c1 - c0 + 1);
            s_tmp[u] = (unsigned short)(fi | (ry << 5) | (c0 << 8) | (len << 11));
            packed += len > 0 ? (len | (1 << 16)) : 0;
          }
        }
        int incl = packed;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int x = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += x;
        }
        if (tid == 31) s_scan0 = incl;
        __syncthreads();
        int excl = incl - packed + (warp == 1 ? s_scan0 : 0);
        int pp = excl & 0xffff, uu = excl >> 16;
#pragma unroll 1
        for (int u = u0; u < u1; ++u) {
          const int cd = s_tmp[u], len = cd >> 11;
          if (len > 0) {
            s_span[uu] = (unsigned short)cd;
            const int w = (pp + 31) & ~31;  // a unit has at most 8 pairs: it contains at most one multiple of 32
            if (w < pp + len) s_win[w >> 5] = (unsigned char)uu;
            pp += len;
            s_upre[++uu] = (unsigned short)pp;
          }
        }
        if (tid == NT - 1) { s_upre[0] = 0; s_nu = uu; }
      }
      __syncthreads();
      // ---- A: one (pixel, face) pair per lane and step
      {
        const int NU = s_nu;
        const int P = NU > 0 ? s_upre[NU] : 0;
        for (int p0 = warp * 32; p0 < P; p0 += 2 * 32) {
          const int p = p0 + lane;
          const bool valid = p < P;
          // unit of pair p: the unit holding the window's first pair (tabulated during the compaction) plus the number of
          // units that end at or before p — unit ends inside the window are one bit each of a warp-wide OR
          const int ua = s_win[p0 >> 5];
          int e = 64;
          if (ua + 1 + lane <= NU) e = (int)s_upre[ua + 1 + lane] - p0;
          const unsigned ends = __reduce_or_sync(0xffffffffu, e < 32 ? (1u << e) : 0u);
          const int lo = valid ? ua + __popc(ends & ((2u << lane) - 1u)) : 0;
          const int cd = s_span[lo];
          const int fi = cd & 31, ly = (cd >> 5) & 7, lx = ((cd >> 8) & 7) + (valid ? p - (int)s_upre[lo] : 0);
          const FaceTile& f = recs[fi];
          const int w = valid ? 1 : 0;
          const int pix = ly * HARP_TILE + lx;
          const float px = s_px[lx], py = s_py[ly];
          bool inb = valid && w != 0 && in_face_bbox(f, px, py);
          bool inside;
          float e0, e1, e2, q = 0.f;
          int cls = classify_fast<SIL>(f, px, py, far, sat, inb, inside, e0, e1, e2);
          bool exact = false;
          if (cls == PAIR_EXACT) {  // a vertex behind the camera: decide with the exact operation sequence
            PairEval e;
            eval_pair_full<false, false>(f, px, py, 0.f, e);
            inside = e.pass;
            q = e.pass ? 1.f / e.pz : 0.f;
            cls = SIL ? PAIR_FULL : (inside ? PAIR_SATURATED : PAIR_REJECT);
            exact = true;
          } else if (inside) {
            q = fast_inv_depth(f, e0, e1, e2);
          }
          if (cls != PAIR_REJECT && inside) {
            // nearest face = largest 1/depth; equal depths prefer the lower face id.  Candidates within 1e-5 (relative) of the
            // running maximum are remembered and settled with the exact depth at the end.
            unsigned long long key = ((unsigned long long)__float_as_uint(q) << 32) | (unsigned long long)(0xffffffffu - (unsigned)f.fid);
            unsigned long long old = atomicMax(&s_vis[pix], key);
            if (old != 0ULL) {
              float qo = __uint_as_float((unsigned)(old >> 32));
              if (fabsf(q - qo) <= 1e-5f * fmaxf(q, qo)) {
                int slot = atomicAdd(&s_tn, 2);
                if (slot + 1 < TIECAP) {
                  s_tie[slot] = pix | (f.fid << 6);
                  s_tie[slot + 1] = pix | ((int)(0xffffffffu - (unsigned)(old & 0xffffffffu)) << 6);
                }
              }
            }
          }
          if (SIL) {
            if (cls != PAIR_REJECT && count_cand) atomicAdd(&s_cnt[pix], 1);  // upper bound on PyTorch3D's candidates
            if (cls == PAIR_SATURATED) s_sat[pix] = 1;  // one exactly-zero factor: alpha = 1, zero gradient
            bool band = cls == PAIR_FULL && !s_sat[pix];  // (benign race, see s_sat)
            unsigned bm = __ballot_sync(0xffffffffu, band);
            if (bm) {
              int qb = 0;
              if (lane == 0) qb = atomicAdd(&s_qn, __popc(bm));
              qb = __shfl_sync(0xffffffffu, qb, 0);
              if (band) {
                int slot = qb + __popc(bm & ((1u << lane) - 1u));
                if (slot < QCAP) {
                  BandEntry e;
                  e.fid = f.fid; e.sd = 0.f; e.t = 0.f; e.pix = (unsigned char)pix;
                  e.edge = (unsigned char)(fi | (inside ? 0x40 : 0) | (exact ? 0x80 : 0));
                  e.next = (short)atomicExch(&s_head[pix], slot);
                  s_q[slot] = e;
                } else {
                  s_ovfp[pix] = 1;
                }
              }
            }
          }
        }
      }
      if (SIL) {
        __syncthreads();
        // ---- B: evaluate this round's band pairs with every thread busy (pixels already known saturated are skipped)
        const int qn = min(s_qn, QCAP);
        for (int s = q_lo + tid; s < qn; s += NT) {
          BandEntry e = s_q[s];
          int pix = e.pix;
          if (s_sat[pix]) { s_q[s].edge = 255; continue; }
          const FaceTile& f = recs[e.edge & 31];
          float ppx = s_px[pix % HARP_TILE], ppy = s_py[pix / HARP_TILE];
          bool inside = (e.edge & 0x40) != 0, pass;
          float dist, t;
          int edge;
          if (e.edge & 0x80) {
            PairEval r;
            eval_pair_full<true, true>(f, ppx, ppy, blur, r);
            pass = r.pass; inside = r.inside; dist = r.dist; edge = r.edge; t = r.t;
          } else {
            band_dist(f, ppx, ppy, dist, edge, t);
            pass = inside || dist < blur;
          }
          if (pass) {
            e.sd = inside ? -dist : dist; e.t = t; e.edge = (unsigned char)edge;
            float om = 1.f - 1.f / (1.f + expf(fdiv(e.sd, sigma)));  // 1 - sigmoid(-sd / sigma)
            s_om[s] = om;
            if (om == 0.f) s_sat2[pix] = 1;  // an exactly-zero factor: alpha = 1, zero gradient
          } else {
            e.edge = 255;
          }
          s_q[s] = e;
        }
        q_lo = qn;
      }
    }
    __syncthreads();
    // ================================================================== per-pixel epilogue (thread = pixel)
    const int lx = tid % HARP_TILE, ly = tid / HARP_TILE;
    const int col = tx0 + lx, row = ty0 + ly;
    const size_t pix = ((size_t)b * S + row) * S + col;
    const float px = s_px[lx], py = s_py[ly];
    // ---- visibility: the running maximum, unless a remembered near-tie has the smaller exact depth
    unsigned long long key = s_vis[tid];
    int best_fid = key != 0ULL ? (int)(0xffffffffu - (unsigned)(key & 0xffffffffu)) : -1;
    {
      const int nt = min(s_tn, TIECAP);
      if (s_tn > TIECAP && tid == 0) atomicAdd(c.counters, 1ULL);  // reported by harp_ctx_read_counters (never seen)
      if (nt > 0 && best_fid >= 0) {
        // a remembered candidate that has since fallen far behind simply loses the exact comparison
        float best_pz = 0.f;
        bool have_pz = false;
        for (int i = 0; i < nt; ++i) {
          int te = s_tie[i], tf = te >> 6;
          if ((te & 63) != tid || tf == best_fid) continue;
          if (!have_pz) { best_pz = exact_pz_rec(a.rec + (size_t)b * c.F, best_fid, px, py); have_pz = true; }
          float pz = exact_pz_rec(a.rec + (size_t)b * c.F, tf, px, py);
          if (pz < best_pz || (pz == best_pz && tf < best_fid)) { best_fid = tf; best_pz = pz; }
        }
      }
    }
    const bool hit = best_fid >= 0;
    if (!CAM) {
      // light view: nearest exact depth + face id (renderer_helper.py:344)
      float z = -1.f;
      int fo = -1;
      if (hit) {
        PairEval e;
        exact_pz_rec(a.rec + (size_t)b * c.F, best_fid, px, py, &e);
        if (e.pass) { z = e.pz; fo = best_fid; }
      }
      a.zbuf[pix] = z;
      a.face[pix] = fo;
      if (a.d_zbuf) a.d_zbuf[pix] = 0.f;  // the gradient image is only ever touched on tiles some face covers
    } else {
      if (a.hit_list) {  // per-frame compacted hit list (warp-aggregated append)
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        int hb = 0;
        if (mask) {
          if (lane == 0) hb = atomicAdd(a.hit_count + b, __popc(mask));
          hb = __shfl_sync(0xffffffffu, hb, 0);
          if (hit) reinterpret_cast<int2*>(a.hit_list)[(size_t)b * S * S + hb + __popc(mask & ((1u << lane) - 1u))] = make_int2((row << 16) | col, best_fid);
        }
      }
      float photo_abs = 0.f;
      if (!hit) {
        if (a.rgb) st3(a.rgb + pix * 3, mk3(1.f, 1.f, 1.f));  // background colour (renderer_helper.py:430)
        if (PHOTO) {
          float mk = a.y_mask[pix];
          if (mk != 0.f) {
            f3 y = ld3(a.y_true + pix * 3);
            photo_abs = fabsf(mk - y.x * mk) + fabsf(mk - y.y * mk) + fabsf(mk - y.z * mk);
          }
        }
      }
      // ---- silhouette: the product of the pixel's band factors.  The chain is in arrival order (shared-memory atomics), so
      // the product is taken in double: re-ordering then changes it by ~1e-16, i.e. the float result is the same however the
      // pairs were scheduled (alpha is reproducible run to run)
      const int qn = min(s_qn, QCAP);
      float prod = 1.f, diff = 0.f;
      int nzero = 0, npass = 0;
      bool deferred = false;
      if (SIL) {
        nzero = (s_sat[tid] | s_sat2[tid]) ? 1 : 0; npass = nzero;
        if (nzero == 0) {
          double dp = 1.0;
          for (int s = s_head[tid]; s >= 0; s = s_q[s].next) {
            if (s_q[s].edge == 255) continue;
            ++npass;
            dp *= (double)s_om[s];
          }
          prod = (float)dp;
        }
        float al = nzero > 0 ? 1.f : 1.f - prod;
        // more candidates than faces_per_pixel, or one of the pixel's band pairs did not fit the queue -> exact per-pixel top-K
        // kernel (the other pixels of an overflowing tile have complete chains and stay on this path)
        deferred = (count_cand && s_cnt[tid] > a.K) || s_ovfp[tid] != 0;
        if (deferred) sil_defer_overflow(c, (int)pix);
        else if (a.alpha) a.alpha[pix] = al;
        if (SIL_BWD && !deferred && !a.g_alpha) diff = al - a.y_sil[pix];
      }
      if (SIL_BWD || PHOTO) {
        float s_sil = fabsf(diff), s_ph = photo_abs;
        block_sum2(s_sil, s_ph, red);
        if (tid == 0) {
          if (SIL_BWD && s_sil != 0.f) atomicAdd(a.sil_loss, s_sil * a.sil_inv_n);
          if (PHOTO && s_ph != 0.f) atomicAdd(a.photo_loss, s_ph * a.photo_inv_n);
        }
      }
      if (SIL_BWD) {
        float g = diff > 0.f ? a.sil_gscale : (diff < 0.f ? -a.sil_gscale : 0.f);
        if (a.g_alpha) g = deferred ? 0.f : a.g_alpha[pix];  // an upstream dL/dalpha image instead of the fused loss
        bool need = (g != 0.f) && nzero == 0 && prod > 0.f && npass > 0;
        s_g[tid] = g; s_prod[tid] = prod; s_need[tid] = need ? 1 : 0;
        if (__syncthreads_or(need ? 1 : 0)) {
          float* d_ndc = a.d_ndc + (size_t)b * c.V * 4;
          for (int s = tid; s < qn; s += NT) {
            BandEntry e = s_q[s];
            if (e.edge == 255 || !s_need[e.pix]) continue;
            sil_pair_gradient_fid(c, ndc, e.fid, s_px[e.pix % HARP_TILE], s_py[e.pix / HARP_TILE], e.sd, e.edge, e.t, s_g[e.pix],
                                  s_prod[e.pix], sigma, d_ndc);
          }
        }
      }
    }
    if (!have_next) break;
    item = nitem;
    idx = nidx;
    cur ^= 1;
    __syncthreads();  // everyone is done with this tile's shared state
  }
}

// Tiles no face touches (4 of 5 at the reference's framing): pure streaming, 16-byte vectors, one thread per 4 pixels of a row.
//   camera view: alpha = 0, colour = background white, and the two L1 losses of those pixels (the colour target is read only
//   where the eroded target mask is non-zero);  light view: depth = -1, face = -1.
template <bool LIGHT, bool SIL_BWD, bool PHOTO>
__global__ void __launch_bounds__(256) empty_tiles_kernel(int B, int S, const int32_t* __restrict__ bin_count, VisArgs a) {
  const int q4 = S / 4;  // float4 groups per row; a group never straddles an 8-pixel tile
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * S * q4;
  bool work = gid < total;
  float s_sil = 0.f, s_ph = 0.f;
  if (work) {
    int g4 = (int)(gid % q4);
    size_t br = gid / q4;
    int row = (int)(br % S), b = (int)(br / S), tw = S / HARP_TILE;
    int n = __ldg(bin_count + (size_t)b * tw * tw + (row / HARP_TILE) * tw + (g4 * 4) / HARP_TILE);
    if (n == 0) {
      size_t pix = br * S + (size_t)g4 * 4;
      if (LIGHT) {
        __stcs(reinterpret_cast<float4*>(a.zbuf + pix), make_float4(-1.f, -1.f, -1.f, -1.f));
        __stcs(reinterpret_cast<int4*>(a.face + pix), make_int4(-1, -1, -1, -1));
      } else {
        if (a.alpha) __stcs(reinterpret_cast<float4*>(a.alpha + pix), make_float4(0.f, 0.f, 0.f, 0.f));
        if (a.rgb) {
          float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
          float4* o = reinterpret_cast<float4*>(a.rgb + pix * 3);
          __stcs(o, one); __stcs(o + 1, one); __stcs(o + 2, one);
        }
        if (SIL_BWD && a.y_sil) {
          float4 ys = __ldcs(reinterpret_cast<const float4*>(a.y_sil + pix));
          s_sil = fabsf(ys.x) + fabsf(ys.y) + fabsf(ys.z) + fabsf(ys.w);
        }
        if (PHOTO) {
          float4 mk = __ldcs(reinterpret_cast<const float4*>(a.y_mask + pix));
          if (mk.x != 0.f || mk.y != 0.f || mk.z != 0.f || mk.w != 0.f) {  // rare: target foreground over an empty tile
            const float4* yt = reinterpret_cast<const float4*>(a.y_true + pix * 3);
            float4 y0 = __ldcs(yt), y1 = __ldcs(yt + 1), y2 = __ldcs(yt + 2);
            s_ph = fabsf(mk.x - y0.x * mk.x) + fabsf(mk.x - y0.y * mk.x) + fabsf(mk.x - y0.z * mk.x) +
                   fabsf(mk.y - y0.w * mk.y) + fabsf(mk.y - y1.x * mk.y) + fabsf(mk.y - y1.y * mk.y) +
                   fabsf(mk.z - y1.z * mk.z) + fabsf(mk.z - y1.w * mk.z) + fabsf(mk.z - y2.x * mk.z) +
                   fabsf(mk.w - y2.y * mk.w) + fabsf(mk.w - y2.z * mk.w) + fabsf(mk.w - y2.w * mk.w);
          }
        }
      }
    }
  }
  if (LIGHT || (!SIL_BWD && !PHOTO)) return;
  if (__ballot_sync(0xffffffffu, s_sil != 0.f || s_ph != 0.f)) {
    s_sil = warp_sum(s_sil); s_ph = warp_sum(s_ph);
    if ((threadIdx.x & 31) == 0) {
      if (SIL_BWD && s_sil != 0.f) atomicAdd(a.sil_loss, s_sil * a.sil_inv_n);
      if (PHOTO && s_ph != 0.f) atomicAdd(a.photo_loss, s_ph * a.photo_inv_n);
    }
  }
}

// Shade the compacted hit pixels of one frame per blockIdx.y; grid-stride over the frame's hit list.
#ifndef HARP_SHADE_MIN_CTAS
#define HARP_SHADE_MIN_CTAS 3  // 168 registers, no spills: the kernel is bound by per-thread instruction latency, so registers beat warps (0.549 -> 0.535 ms)
#endif
template <int MODE>
__global__ void __launch_bounds__(128, HARP_SHADE_MIN_CTAS) shade_deferred_kernel(harp_ctx c, ShadeArgs a, const int2* __restrict__ hit_list,
                                                             const int32_t* __restrict__ hit_count) {
  const int S = a.S;
  int b = blockIdx.y;
  int n = hit_count[b];
  const FaceTile* frec = c.face_rec + (size_t)b * c.F;  // camera view
  const int2* list = hit_list + (size_t)b * S * S;
  // per-thread partial sums of the frame scalars live in shared memory (column per thread: conflict-free), not registers
  __shared__ float s_fr[9][128];
#pragma unroll
  for (int k = 0; k < 9; ++k) s_fr[k][threadIdx.x] = 0.f;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    int2 rec = list[idx];
    const int row = rec.x >> 16, col = rec.x & 0xffff;
    size_t pix = ((size_t)b * S + row) * S + col;
    float px = pix_to_ndc(col, S), py = pix_to_ndc(row, S);
    Hit h;
    h.fid = rec.y;
    // the face's raster record (one contiguous 144-byte row written by the setup kernel) and its vertex ids are requested
    // together: one dependent hop less than face table -> three projected vertices
    const float4* r4 = reinterpret_cast<const float4*>(frec + h.fid);
    const int32_t* fv = c.faces + (size_t)h.fid * 3;
    float4 r0 = __ldg(r4), r1 = __ldg(r4 + 1), r2 = __ldg(r4 + 2), r7 = __ldg(r4 + 7);
    const int v0 = fv[0], v1 = fv[1], v2 = fv[2];
    FaceRec f;
    f.x0 = r0.x; f.y0 = r0.y; f.x1 = r0.z; f.y1 = r0.w; f.x2 = r1.x; f.y2 = r1.y; f.area = r2.w; f.z0 = r7.x; f.z1 = r7.y; f.z2 = r7.z;
    ShadePre pre;  // vertex attributes, uvs and targets: in flight while the barycentrics are evaluated
    shade_preload<MODE>(c, a, b, pix, h.fid, v0, v1, v2, pre);
    PairEval e;
    eval_pair_full<false, false>(f, px, py, 0.f, e);
    if (e.pass) { h.pz = e.pz; h.b0 = e.b0; h.b1 = e.b1; h.b2 = e.b2; h.w0 = e.w0; h.w1 = e.w1; h.w2 = e.w2; }
    else h.fid = -1;  // the exact test disagrees with the fast classification on a boundary pixel: background
    const Tri2 tri{f.x0, f.y0, f.z0, f.x1, f.y1, f.z1, f.x2, f.y2, f.z2};
    ShadeOut o;
    shade_pixel<MODE>(c, a, b, pix, px, py, h, o, &pre, &tri);
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


// One render pass over B frames: which branches run and how their gradients are driven.
struct RenderPlan {
  bool sil = true, shade = true;   // silhouette branch / colour branch
  int sil_mode = 0, shade_mode = 0;  // 0 forward; 1 backward of an upstream gradient image; 2 fused L1 loss + its gradient
  const float* g_alpha = nullptr;  // sil_mode 1: dL/dalpha (B,S,S)
  const float* fixed_lights = nullptr;  // host {ka, kd, ks}: the Phong twin with constant light colours
  const char* who = "harp_render";
};

// The schedule every render entry point shares (harp_render_fused, harp_silhouette, harp_shade).  In shade mode 1 the
// upstream dL/drgb image travels in r->y_true.
int render_pass(harp_ctx* c, int B, int S, float focal, const harp_render_args* r, const RenderPlan& p, cudaStream_t st) {
  const bool sil_bwd = p.sil && p.sil_mode != 0, shade_bwd = p.shade && p.shade_mode != 0, photo = p.shade && p.shade_mode == 2;
  const bool shadow = p.shade && r->self_shadow;
  int rc;
  const size_t nv4 = (size_t)B * c->V * 4 * sizeof(float), ntex = (size_t)r->texH * r->texW;
  double blur_d = 9.210240366975849 * (double)r->sigma;  // ln(1/1e-4 - 1) sigma (renderer_helper.py:46), in double like numpy
  float blur = p.sil ? (float)blur_d : 0.f, radius = p.sil ? (float)sqrt(blur_d) : 0.f;
  // Two branches, forked from and joined to the caller's stream with events (capturable into a CUDA graph):
  //   main   camera setup -> tile rasteriser (persistent: it holds every SM until its queue drains) -> top-K -> shading -> tail
  //   side   (least priority) what only the shading needs — light-view setup and the light view's empty tiles, packed texel
  //          copies, zero fills of the gradient accumulators, then the light-view rasteriser, whose CTAs move in as the
  //          camera rasteriser's CTAs retire (they fill its tail) — and the HBM-bound streaming pass over the camera view's
  //          empty tiles; joined before the shading.
  // (Measured with harp_ctx_profile_timeline: the GPU is busy end to end, the branches only fill each other's tails.  Running
  // the empty-tile pass UNDER the shading instead, on a high-priority stream, lengthened the shading by what the pass takes.)
  if (!c->side_stream) {
    int lo = 0, hi = 0;
    HARP_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // lo = least priority
    HARP_CHECK_CUDA(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, lo));
    for (cudaEvent_t* e : {&c->ev_fork, &c->ev_join, &c->ev_setup})
      HARP_CHECK_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  cudaStream_t ss = c->side_stream;
  HARP_CHECK_CUDA(cudaEventRecord(c->ev_fork, st));
  harp_prof_begin(c, 6, st);
  // (the setup kernels also zero the projected-vertex gradient accumulator, the hit counters and the top-K counter)
  if ((rc = harp_raster_setup(c, 0, B, S, focal, radius, r->verts, r->cam_R, r->cam_T, st, true)) != HARP_OK) return rc;
  HARP_CHECK_CUDA(cudaEventRecord(c->ev_setup, st));
  harp_prof_end(c, 6, st);
  VisArgs v{};
  v.S = S; v.K = r->faces_per_pixel > 0 ? r->faces_per_pixel : 50; v.sigma = r->sigma; v.blur = blur; v.radius = radius; v.ndc = c->ndc_cam;
  v.rec = c->face_rec; v.bin_list = c->bin_list; v.work = c->work; v.work_ctl = c->work_ctl; v.work_cap = c->maxB * c->tiles;
  v.sil_inv_n = 1.f / ((float)B * S * S); v.sil_gscale = r->w_sil * v.sil_inv_n;
  if (p.sil) {
    v.alpha = p.sil_mode == 1 ? nullptr : r->alpha;
    v.y_sil = p.sil_mode == 2 ? r->y_sil : nullptr; v.g_alpha = p.sil_mode == 1 ? p.g_alpha : nullptr;
    v.sil_loss = r->loss_sil; v.d_ndc = c->d_ndc_cam;
  }
  v.photo_inv_n = 1.f / ((float)B * S * S * 3.f);
  if (p.shade) {
    v.hit_list = c->hit_list; v.hit_count = c->hit_count;
    v.rgb = p.shade_mode == 1 ? nullptr : r->rgb;
    if (photo) { v.y_true = r->y_true; v.y_mask = r->y_mask; v.photo_loss = r->loss_photo; }
  }
  const int grid = HARP_SM_COUNT * HARP_PV_MINCTAS;  // persistent: every CTA pulls tiles until the queue is empty
  harp_prof_begin(c, 0, st);
  if (p.sil) {
    if (sil_bwd && photo) tile_raster_kernel<true, true, true, true><<<grid, NT, 0, st>>>(*c, v);
    else if (sil_bwd) tile_raster_kernel<true, true, false, true><<<grid, NT, 0, st>>>(*c, v);
    else if (photo) tile_raster_kernel<true, false, true, true><<<grid, NT, 0, st>>>(*c, v);
    else tile_raster_kernel<true, false, false, true><<<grid, NT, 0, st>>>(*c, v);
  } else {
    if (photo) tile_raster_kernel<false, false, true, true><<<grid, NT, 0, st>>>(*c, v);
    else tile_raster_kernel<false, false, false, true><<<grid, NT, 0, st>>>(*c, v);
  }
  harp_prof_end(c, 0, st);
  HARP_LAUNCH_CHECK();
  if (p.sil && (rc = harp_sil_topk(c, p.sil_mode, S, v.K, r->sigma, blur, radius, r->alpha, p.sil_mode == 1 ? p.g_alpha : r->y_sil,
                                   v.sil_gscale, v.sil_inv_n, r->loss_sil, st)) != HARP_OK) return rc;
  // ---- side branch
  HARP_CHECK_CUDA(cudaStreamWaitEvent(ss, c->ev_fork, 0));
  harp_prof_begin(c, 9, ss);
  if (shadow && (rc = harp_light_view_setup(c, B, S, focal, r->verts, r->light_R, r->light_T, shade_bwd, ss)) != HARP_OK) return rc;
  if (p.shade && (rc = harp_tex_prepare(c, r->texture, r->normal_map, r->texH, r->texW, ss)) != HARP_OK) return rc;
  if (shade_bwd) {
    HARP_CHECK_CUDA(cudaMemsetAsync(c->d_verts4, 0, nv4, ss));
    HARP_CHECK_CUDA(cudaMemsetAsync(c->d_normals4, 0, nv4, ss));
    HARP_CHECK_CUDA(cudaMemsetAsync(c->d_tex4, 0, 2 * ntex * 4 * sizeof(float), ss));
  }
  if (shadow && (rc = harp_light_depth_tiles(c, B, S, shade_bwd, ss)) != HARP_OK) return rc;
  if (v.alpha || v.rgb || sil_bwd || photo) {  // (in stream order after the camera setup: ev_setup)
    HARP_CHECK_CUDA(cudaStreamWaitEvent(ss, c->ev_setup, 0));
    const int eg = (int)(((size_t)B * S * (S / 4) + 255) / 256);
    if (sil_bwd && photo) empty_tiles_kernel<false, true, true><<<eg, 256, 0, ss>>>(B, S, c->bin_count, v);
    else if (sil_bwd) empty_tiles_kernel<false, true, false><<<eg, 256, 0, ss>>>(B, S, c->bin_count, v);
    else if (photo) empty_tiles_kernel<false, false, true><<<eg, 256, 0, ss>>>(B, S, c->bin_count, v);
    else empty_tiles_kernel<false, false, false><<<eg, 256, 0, ss>>>(B, S, c->bin_count, v);
    HARP_LAUNCH_CHECK();
  }
  harp_prof_end(c, 9, ss);
  HARP_CHECK_CUDA(cudaEventRecord(c->ev_join, ss));
  HARP_CHECK_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));
  harp_prof_begin(c, 1, st);
  if (p.shade) {
    ShadeArgs a{};
    a.B = B; a.S = S; a.self_shadow = r->self_shadow; a.texH = r->texH; a.texW = r->texW; a.focal = focal;
    if (p.fixed_lights) { a.fixed_lights = 1; a.ka = p.fixed_lights[0]; a.kd = p.fixed_lights[1]; a.ks = p.fixed_lights[2]; }
    a.inv_n = v.photo_inv_n; a.gscale = r->w_photo * a.inv_n;
    a.verts = r->verts; a.normals = r->normals; a.ndc_cam = c->ndc_cam;
    a.light_R = r->light_R ? r->light_R : r->cam_R; a.light_T = r->light_T ? r->light_T : r->cam_T;
    a.light_pos = r->light_pos; a.amb = r->amb;
    a.tex4 = reinterpret_cast<const float4*>(c->tex4); a.nm4 = a.tex4 + ntex;
    a.light_zbuf = c->light_zbuf; a.rgb = r->rgb; a.gin = r->y_true; a.mask = r->y_mask; a.loss = r->loss_photo;
    if (shade_bwd) {
      a.d_ndc_cam = c->d_ndc_cam; a.d_verts = c->d_verts4; a.d_normals = c->d_normals4;
      a.d_light_R = r->grads.d_light_R; a.d_light_T = r->grads.d_light_T; a.d_light_pos = r->grads.d_light_pos;
      a.d_amb = r->grads.d_amb; a.d_texture = c->d_tex4; a.d_normal_map = c->d_tex4 + ntex * 4; a.d_light_zbuf = c->d_light_zbuf;
    }
    dim3 sgrid(HARP_SHADE_GRID, B);
    const int2* hits = reinterpret_cast<const int2*>(c->hit_list);
    if (p.shade_mode == 2) shade_deferred_kernel<2><<<sgrid, 128, 0, st>>>(*c, a, hits, c->hit_count);
    else if (p.shade_mode == 1) shade_deferred_kernel<1><<<sgrid, 128, 0, st>>>(*c, a, hits, c->hit_count);
    else shade_deferred_kernel<0><<<sgrid, 128, 0, st>>>(*c, a, hits, c->hit_count);
    HARP_LAUNCH_CHECK();
  }
  harp_prof_end(c, 1, st);
  harp_prof_begin(c, 7, st);
  if (shade_bwd) {
    if (!c->defer_fold) {  // harp_step_fused folds the padded vertex accumulators inside its geometry backward
      if ((rc = harp_fold4((size_t)B * c->V, c->d_verts4, r->grads.d_verts, st)) != HARP_OK) return rc;
      if ((rc = harp_fold4((size_t)B * c->V, c->d_normals4, r->grads.d_normals, st)) != HARP_OK) return rc;
    }
    if (r->grads.d_texture && r->grads.d_normal_map) {
      if ((rc = harp_fold4_pair(ntex, c->d_tex4, r->grads.d_texture, r->grads.d_normal_map, st)) != HARP_OK) return rc;
    } else {
      if (r->grads.d_texture && (rc = harp_fold4(ntex, c->d_tex4, r->grads.d_texture, st)) != HARP_OK) return rc;
      if (r->grads.d_normal_map && (rc = harp_fold4(ntex, c->d_tex4 + ntex * 4, r->grads.d_normal_map, st)) != HARP_OK) return rc;
    }
    if (shadow && (rc = harp_light_depth_bwd_tiles(c, B, S, st)) != HARP_OK) return rc;
  }
  if (shade_bwd && shadow) {  // both views' projection backward in one launch
    if ((rc = harp_project_bwd2(c, B, S, focal, r->verts, r->cam_R, r->cam_T, c->d_ndc_cam, r->grads.d_cam_R, r->grads.d_cam_T, r->light_R,
                                r->light_T, c->d_ndc_light, r->grads.d_light_R, r->grads.d_light_T, r->grads.d_verts, st)) != HARP_OK) return rc;
  } else if (sil_bwd || shade_bwd) {
    if ((rc = harp_project_bwd(c, B, S, focal, r->verts, r->cam_R, r->cam_T, c->d_ndc_cam, r->grads.d_verts, r->grads.d_cam_R,
                               r->grads.d_cam_T, st)) != HARP_OK) return rc;
  }
  harp_prof_end(c, 7, st);
  return HARP_OK;
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
  RenderPlan p;
  p.sil_mode = sil_bwd ? 2 : 0; p.shade_mode = photo_bwd ? 2 : 0;
  return render_pass(c, B, S, focal, r, p, (cudaStream_t)stream);
}

// Soft silhouette alone (the drop-in for SoftSilhouetteShader rendering, renderer_helper.py:40-62): the same setup + tile
// rasteriser + top-K kernels as the fitting step, without hit list or colour image.
extern "C" int harp_silhouette(harp_ctx* c, int mode, int B, int S, float focal, float sigma, int faces_per_pixel, const float* verts,
                               const float* R, const float* T, float* alpha, const float* gin, float weight, float* loss,
                               float* d_verts, float* d_R, float* d_T, void* stream) {
  HARP_REQUIRE(c && B > 0 && verts && R && T, HARP_ERR_INVALID, "harp_silhouette: null argument");
  HARP_REQUIRE(mode >= 0 && mode <= 2, HARP_ERR_INVALID, "harp_silhouette: mode must be 0, 1 or 2");
  HARP_REQUIRE(B <= c->maxB && S == c->S, HARP_ERR_CAPACITY, "harp_silhouette: call harp_ctx_reserve(B, S) first");
  HARP_REQUIRE(mode == 1 || alpha, HARP_ERR_INVALID, "harp_silhouette: alpha output required");
  HARP_REQUIRE(mode == 0 || gin, HARP_ERR_INVALID, "harp_silhouette: upstream gradient / target required");
  HARP_REQUIRE(mode != 2 || loss, HARP_ERR_INVALID, "harp_silhouette: loss output required in fused mode");
  HARP_REQUIRE(mode == 0 || d_verts, HARP_ERR_INVALID, "harp_silhouette: d_verts required");
  harp_render_args r{};
  r.verts = verts; r.cam_R = R; r.cam_T = T; r.sigma = sigma; r.faces_per_pixel = faces_per_pixel;
  r.alpha = alpha; r.y_sil = gin; r.w_sil = weight; r.loss_sil = loss;
  r.grads.d_verts = d_verts; r.grads.d_cam_R = d_R; r.grads.d_cam_T = d_T;
  RenderPlan p;
  p.shade = false; p.sil_mode = mode; p.g_alpha = gin;
  return render_pass(c, B, S, focal, &r, p, (cudaStream_t)stream);
}

// Shadowed, textured colour image alone (the drop-in for the hard-Phong / NeuralShadow renders, renderer_helper.py:64-120,
// 327-430): setup + visibility-only tile rasteriser + deferred shading, as in the fitting step.
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
  harp_render_args r{};
  r.verts = verts; r.normals = normals; r.cam_R = cam_R; r.cam_T = cam_T; r.light_R = light_R; r.light_T = light_T;
  r.light_pos = light_pos; r.amb = amb; r.texture = texture; r.normal_map = normal_map; r.texH = texH; r.texW = texW;
  r.self_shadow = self_shadow; r.rgb = rgb; r.y_true = gin; r.y_mask = mask; r.w_photo = weight; r.loss_photo = loss;
  if (grads) r.grads = *grads;
  RenderPlan p;
  p.sil = false; p.shade_mode = mode; p.fixed_lights = fixed_lights;
  return render_pass(c, B, S, focal, &r, p, (cudaStream_t)stream);
}

namespace {

__global__ void __launch_bounds__(256) hit_mask_kernel(int S, const int2* __restrict__ hit_list, const int32_t* __restrict__ hit_count,
                                                       float* __restrict__ mask) {
  const int b = blockIdx.y, n = hit_count[b];
  const int2* list = hit_list + (size_t)b * S * S;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int rc = list[i].x;
    mask[((size_t)b * S + (rc >> 16)) * S + (rc & 0xffff)] = 1.f;
  }
}

// backward of the light-view depth buffer (zbuf = sum_i b_i z_i at the stored face) over the non-empty light tiles only: the
// PCF taps of the shading can only have touched pixels of tiles some face covers.  Four tiles per CTA round; the pixels with a
// gradient are compacted in shared memory and then processed by full warps.
__global__ void __launch_bounds__(256) light_depth_bwd_tiles_kernel(harp_ctx c, int S, const int4* __restrict__ work, const int32_t* __restrict__ ctl,
                                                                    int work_cap, const FaceTile* __restrict__ rec_all,
                                                                    const int32_t* __restrict__ face, const float* __restrict__ d_zbuf,
                                                                    float* __restrict__ d_ndc_all) {
  __shared__ int s_pix[256];
  __shared__ float s_g[256];
  __shared__ int s_n;
  const int n_heavy = ctl[0], total = n_heavy + ctl[1], tw = S / HARP_TILE, t = threadIdx.x;
  for (int i0 = blockIdx.x * 4; i0 < total; i0 += gridDim.x * 4) {
    if (t == 0) s_n = 0;
    __syncthreads();
    int i = i0 + (t >> 6);
    if (i < total) {
      int4 it = work[i < n_heavy ? i : work_cap - 1 - (i - n_heavy)];
      int lx = t & 7, ly = (t >> 3) & 7;
      int row = (it.y / tw) * HARP_TILE + ly, col = (it.y % tw) * HARP_TILE + lx;
      int pix = (it.x * S + row) * S + col;
      float g = d_zbuf[pix];
      if (g != 0.f) { int s = atomicAdd(&s_n, 1); s_pix[s] = pix; s_g[s] = g; }
    }
    __syncthreads();
    const int n = s_n;
    if (t < n) {
      int pix = s_pix[t];
      int fid = face[pix];
      if (fid >= 0) {
        float g = s_g[t];
        int b = pix / (S * S), rem = pix % (S * S), row = rem / S, col = rem % S;
        float px = pix_to_ndc(col, S), py = pix_to_ndc(row, S);
        const float4* r4 = reinterpret_cast<const float4*>(rec_all + (size_t)b * c.F + fid);
        float4 r0 = __ldg(r4), r1 = __ldg(r4 + 1), r2 = __ldg(r4 + 2), r7 = __ldg(r4 + 7);
        FaceRec f;
        f.x0 = r0.x; f.y0 = r0.y; f.x1 = r0.z; f.y1 = r0.w; f.x2 = r1.x; f.y2 = r1.y; f.area = r2.w; f.z0 = r7.x; f.z1 = r7.y; f.z2 = r7.z;
        PairEval e;
        eval_pair_full<false, false>(f, px, py, 0.f, e);
        if (e.pass) {
          const int32_t* fv = c.faces + (size_t)fid * 3;
          Hit h; h.fid = fid; h.pz = e.pz; h.b0 = e.b0; h.b1 = e.b1; h.b2 = e.b2; h.w0 = e.w0; h.w1 = e.w1; h.w2 = e.w2;
          float gb[3] = {g * f.z0, g * f.z1, g * f.z2};
          float gz[3] = {g * e.b0, g * e.b1, g * e.b2};
          bary_backward(Tri2{f.x0, f.y0, f.z0, f.x1, f.y1, f.z1, f.x2, f.y2, f.z2}, fv[0], fv[1], fv[2], px, py, h, gb, gz,
                        d_ndc_all + (size_t)b * c.V * 4);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int harp_hit_mask(harp_ctx* c, int B, int S, float* mask, void* stream) {
  HARP_REQUIRE(c && mask && B > 0, HARP_ERR_INVALID, "harp_hit_mask: null argument");
  HARP_REQUIRE(B <= c->maxB && S == c->S, HARP_ERR_CAPACITY, "harp_hit_mask: no render pass of this size on the ctx");
  cudaStream_t st = (cudaStream_t)stream;
  HARP_CHECK_CUDA(cudaMemsetAsync(mask, 0, (size_t)B * S * S * sizeof(float), st));
  hit_mask_kernel<<<dim3(32, B), 256, 0, st>>>(S, reinterpret_cast<const int2*>(c->hit_list), c->hit_count, mask);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

int harp_light_depth_bwd_tiles(harp_ctx* c, int B, int S, cudaStream_t st) {
  (void)B;
  const int cap = c->maxB * c->tiles;
  light_depth_bwd_tiles_kernel<<<HARP_SM_COUNT * 4, 256, 0, st>>>(*c, S, c->work + cap, c->work_ctl + 4, cap, c->face_rec + (size_t)c->maxB * c->F,
                                                               c->light_face, c->d_light_zbuf, c->d_ndc_light);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

// Light view (renderer_helper.py:344), part 1: setup of view 1 and the streaming fill of its empty tiles (depth -1, face -1).
static VisArgs light_view_args(harp_ctx* c, int S, bool zero_grad) {
  VisArgs v{};
  v.S = S; v.K = 1; v.sigma = 0.f; v.blur = 0.f; v.radius = 0.f; v.ndc = c->ndc_light;
  v.rec = c->face_rec + (size_t)c->maxB * c->F;
  v.bin_list = c->bin_list + (size_t)c->maxB * c->bin_cap;
  v.work_cap = c->maxB * c->tiles;
  v.work = c->work + v.work_cap;
  v.work_ctl = c->work_ctl + 4;
  v.zbuf = c->light_zbuf; v.face = c->light_face; v.d_zbuf = zero_grad ? c->d_light_zbuf : nullptr;
  return v;
}

int harp_light_view_setup(harp_ctx* c, int B, int S, float focal, const float* verts, const float* R, const float* T, bool zero_grad,
                          cudaStream_t st) {
  int rc = harp_raster_setup(c, 1, B, S, focal, 0.f, verts, R, T, st, zero_grad);
  if (rc != HARP_OK) return rc;
  const int eg = (int)(((size_t)B * S * (S / 4) + 255) / 256);
  empty_tiles_kernel<true, false, false><<<eg, 256, 0, st>>>(B, S, c->bin_count + (size_t)c->maxB * c->tiles, light_view_args(c, S, zero_grad));
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

// Part 2: nearest depth + face id on the covered tiles with the persistent tile rasteriser.
int harp_light_depth_tiles(harp_ctx* c, int B, int S, bool zero_grad, cudaStream_t st) {
  (void)B;
  // (62 registers and 9 KB of shared memory without the silhouette state: 16 CTAs fit an SM)
  tile_raster_kernel<false, false, false, false><<<HARP_SM_COUNT * 16, NT, 0, st>>>(*c, light_view_args(c, S, zero_grad));
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}
