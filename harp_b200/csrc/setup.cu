// Raster setup of one view: projection, per-face raster records, tile binning and the work queue of non-empty tiles, as three
// grid-wide launches (a one-CTA-per-frame version that kept everything in shared memory used 64 of 148 SMs: 0.21 ms against 0.08).
// Replaces MeshRasterizer.transform and the coarse pass of _C.rasterize_meshes (renderer/renderer_helper.py:44-58,344,353).
//
//   face_setup_kernel  thread per (frame, face): vertices -> NDC (threads f < V also write the projected vertices
//                      and zero their gradient rows), the 144-byte FaceTile record (raster.cuh) written ONCE to HBM with the face's
//                      pixel rectangle, tile range and kept-tile mask in its tail, tile counts with global atomics
//   bin_scan_kernel    CTA per frame: exclusive scan of the tile counts -> starts / cursors; the work queue of non-empty tiles
//                      (heavy tiles from the front, light ones from the back); zeroes the hit and overflow counters
//   bin_fill_kernel    thread per (frame, face): the face id into the lists of the tiles its mask keeps
#include <algorithm>

#include "raster.cuh"

namespace {

constexpr int SETUP_THREADS = 1024;
constexpr int HEAVY_FACES = 40;  // tiles with more faces than this are consumed first by the persistent pixel kernels

struct TileRange {
  int tx0, tx1, ty0, ty1;
  bool edge_test;  // all three view depths > 0 and the bbox spans several tiles: skip tiles wholly outside one edge line
  float sgn, thr12, thr20, thr01;
};

// Pixel column j has x_ndc = 1 - (2j+1)/S (decreasing); a face with inflated bbox [lo,hi] can touch columns
// ceil((1-hi) S/2 - 0.5) .. floor((1-lo) S/2 - 0.5); a 0.01 px guard absorbs rounding (the exact test runs per pixel).
__device__ __forceinline__ bool face_pixel_rect(const float* a, const float* b, const float* d, float radius, int S, int& jlo, int& jhi,
                                                int& ilo, int& ihi) {
  float zmax = fmaxf(fmaxf(a[2], b[2]), d[2]);
  float area = edge_fn(a[0], a[1], b[0], b[1], d[0], d[1]);  // EdgeFunctionForward(v0, v1, v2)
  if (zmax < 0.f || (area <= HARP_K_EPS && area >= -HARP_K_EPS)) return false;
  if (!(area == area)) return false;  // NaN guard
  float xmin = fminf(fminf(a[0], b[0]), d[0]) - radius, xmax = fmaxf(fmaxf(a[0], b[0]), d[0]) + radius;
  float ymin = fminf(fminf(a[1], b[1]), d[1]) - radius, ymax = fmaxf(fmaxf(a[1], b[1]), d[1]) + radius;
  float hs = 0.5f * (float)S;
  float fjlo = ceilf((1.f - xmax) * hs - 0.5f - 0.01f), fjhi = floorf((1.f - xmin) * hs - 0.5f + 0.01f);
  float filo = ceilf((1.f - ymax) * hs - 0.5f - 0.01f), fihi = floorf((1.f - ymin) * hs - 0.5f + 0.01f);
  fjlo = fmaxf(fjlo, 0.f); filo = fmaxf(filo, 0.f);
  fjhi = fminf(fjhi, (float)(S - 1)); fihi = fminf(fihi, (float)(S - 1));
  if (!(fjlo <= fjhi && filo <= fihi)) return false;
  jlo = (int)fjlo; jhi = (int)fjhi; ilo = (int)filo; ihi = (int)fihi;
  return true;
}

// Per-tile edge test.  Pixel pairs of a skipped tile are exactly those the classification rejects: every pixel centre lies
// outside one edge LINE by more than sqrt(1.05 blur) (+1e-5 NDC guard, four orders above the fp32 error of an edge function),
// and the distance to the triangle is at least the distance to that line.
__device__ __forceinline__ void tile_range(const float* a, const float* b, const float* d, float radius, int jlo, int jhi, int ilo,
                                           int ihi, TileRange& r) {
  r.tx0 = jlo / HARP_TILE; r.tx1 = jhi / HARP_TILE; r.ty0 = ilo / HARP_TILE; r.ty1 = ihi / HARP_TILE;
  r.edge_test = fminf(fminf(a[2], b[2]), d[2]) > 0.f && (r.tx1 > r.tx0 || r.ty1 > r.ty0);
  if (r.edge_test) {
    r.sgn = fadd(edge_fn(d[0], d[1], a[0], a[1], b[0], b[1]), HARP_K_EPS) > 0.f ? 1.f : -1.f;  // sign of FaceTile::area
    float m = 1.05f * radius * radius + 1e-10f;
    r.thr12 = m * ((d[0] - b[0]) * (d[0] - b[0]) + (d[1] - b[1]) * (d[1] - b[1]));
    r.thr20 = m * ((a[0] - d[0]) * (a[0] - d[0]) + (a[1] - d[1]) * (a[1] - d[1]));
    r.thr01 = m * ((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]));
  }
}

// largest signed edge function over the four corner pixel centres of a tile (the function is affine in the pixel position)
__device__ __forceinline__ float tile_max_edge(float sgn, float xa, float xb, float ya, float yb, float ax, float ay, float bx,
                                               float by) {
  float dx = bx - ax, dy = by - ay;
  float ea = (xa - ax) * dy, eb = (xb - ax) * dy, fa = (ya - ay) * dx, fb = (yb - ay) * dx;
  return fmaxf(fmaxf(sgn * (ea - fa), sgn * (ea - fb)), fmaxf(sgn * (eb - fa), sgn * (eb - fb)));
}

__device__ __forceinline__ bool tile_outside_face(const TileRange& r, const float* a, const float* b, const float* d, float xa, float xb,
                                                  float ya, float yb) {
  float m0 = tile_max_edge(r.sgn, xa, xb, ya, yb, b[0], b[1], d[0], d[1]);  // e0 = EdgeFunction(p, v1, v2)
  if (m0 < 0.f && m0 * m0 >= r.thr12) return true;
  float m1 = tile_max_edge(r.sgn, xa, xb, ya, yb, d[0], d[1], a[0], a[1]);
  if (m1 < 0.f && m1 * m1 >= r.thr20) return true;
  float m2 = tile_max_edge(r.sgn, xa, xb, ya, yb, a[0], a[1], b[0], b[1]);
  return m2 < 0.f && m2 * m2 >= r.thr01;
}

struct SetupArgs {
  int B, S, tiles;
  float focal, radius;
  const float *verts, *R, *T;   // (B,V,3), (B,3,3), (B,3)
  float* ndc;                   // (B,V,3) out
  FaceTile* rec;                // (B,F) out
  int32_t *bin_count, *bin_start, *bin_cursor, *bin_list;  // (B,tiles) x3, (B,bin_cap)
  int4* work;                   // (B*tiles) out: heavy items from the front, light from the back
  int32_t* work_ctl;            // {heavy, light, cursor, -}
  int work_cap;
  // optional zero fills folded into these kernels (saves their memset nodes on the step's critical path)
  float* zero_d_ndc;            // (B,V,4) projected-vertex gradient accumulator of this view
  int32_t *zero_hit_count, *zero_ovf_count;  // (B), (1)
};

// approximate NDC of a pixel centre (the conservative tile test has a 1e-5 NDC guard; the exact value is pix_to_ndc)
__device__ __forceinline__ float pix_to_ndc_fast(int idx, float inv_s) { return 1.f - (2.f * (float)idx + 1.f) * inv_s; }

// ---- K1: one thread per (frame, face): projection of its three vertices, the 144-byte raster record, the tile range with
// the per-tile edge test folded into a bit mask (kept in the record for K3), tile counts by global atomics.
// Threads with face index < V also write the projected vertex of that index for the kernels that read NDC by vertex.
__global__ void __launch_bounds__(256) face_setup_kernel(harp_ctx c, SetupArgs a) {
  const int n = max(c.F, c.V);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B * n) return;
  const int b = idx / n, f = idx % n, S = a.S;
  const float* r = a.R + (size_t)b * 9;
  const float* tt = a.T + (size_t)b * 3;
  const float r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3], r4 = r[4], r5 = r[5], r6 = r[6], r7 = r[7], r8 = r[8];
  const float t0 = tt[0], t1 = tt[1], t2 = tt[2], half_s = 0.5f * (float)S;
  const float* vb = a.verts + (size_t)b * c.V * 3;
  auto project = [&](int v, float* o) {  // X_view = X R + T ; ndc_xy = (f X/Z) / (S/2) ; keep view z  
    f3 p = ld3(vb + (size_t)v * 3);
    float X = p.x * r0 + p.y * r3 + p.z * r6 + t0;
    float Y = p.x * r1 + p.y * r4 + p.z * r7 + t1;
    float Z = p.x * r2 + p.y * r5 + p.z * r8 + t2;
    o[0] = (a.focal * X / Z) / half_s; o[1] = (a.focal * Y / Z) / half_s; o[2] = Z;
  };
  if (f < c.V) {
    float o[3];
    project(f, o);
    st3(a.ndc + ((size_t)b * c.V + f) * 3, mk3(o[0], o[1], o[2]));
    if (a.zero_d_ndc) reinterpret_cast<float4*>(a.zero_d_ndc)[(size_t)b * c.V + f] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (f >= c.F) return;
  const int32_t* fv = c.faces + (size_t)f * 3;
  float pa[3], pb[3], pd[3];
  project(fv[0], pa); project(fv[1], pb); project(fv[2], pd);
  FaceTile ft;
  make_face_record(pa[0], pa[1], pa[2], pb[0], pb[1], pb[2], pd[0], pd[1], pd[2], f, a.radius, ft);
  int jlo, jhi, ilo, ihi;
  const bool ok = face_pixel_rect(pa, pb, pd, a.radius, S, jlo, jhi, ilo, ihi);
  ft.cols = ok ? (jlo | (jhi << 16)) : 1;  // jlo = 1 > jhi = 0: skipped face
  ft.rows = ok ? (ilo | (ihi << 16)) : 1;
  unsigned mask = 0u;
  int packed = 0;
  if (ok) {
    TileRange tr;
    tile_range(pa, pb, pd, a.radius, jlo, jhi, ilo, ihi, tr);
    const int ntx = tr.tx1 - tr.tx0 + 1, nty = tr.ty1 - tr.ty0 + 1, tw = S / HARP_TILE;
    packed = tr.tx0 | (tr.ty0 << 8) | (ntx << 16) | (nty << 24);
    const bool small = ntx * nty <= 32;  // larger ranges (a hand zoomed past the image) are re-tested by the fill kernel
    int32_t* cnt = a.bin_count + (size_t)b * a.tiles;
    const float inv_s = 1.f / (float)S;
    int k = 0;
    for (int ty = tr.ty0; ty <= tr.ty1; ++ty)
      for (int tx = tr.tx0; tx <= tr.tx1; ++tx, ++k) {
        if (tr.edge_test) {
          float xa = pix_to_ndc_fast(tx * HARP_TILE, inv_s), xb = pix_to_ndc_fast(tx * HARP_TILE + HARP_TILE - 1, inv_s);
          float ya = pix_to_ndc_fast(ty * HARP_TILE, inv_s), yb = pix_to_ndc_fast(ty * HARP_TILE + HARP_TILE - 1, inv_s);
          if (tile_outside_face(tr, pa, pb, pd, xa, xb, ya, yb)) continue;
        }
        if (small) mask |= 1u << k;
        atomicAdd(cnt + ty * tw + tx, 1);
      }
  }
  ft.pad0 = packed;
  ft.pad1 = (int)mask;
  const float4* src = reinterpret_cast<const float4*>(&ft);
  float4* dst = reinterpret_cast<float4*>(a.rec + (size_t)b * c.F + f);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(FaceTile) / 16); ++k) dst[k] = src[k];
}

// ---- K2: one CTA per frame: exclusive scan of the tile counts -> list starts / fill cursors, and the frame's non-empty
// tiles appended to the view's work queue (heavy tiles from the front, light from the back).
__global__ void __launch_bounds__(SETUP_THREADS) bin_scan_kernel(harp_ctx c, SetupArgs a) {
  __shared__ int s_wsum[32], s_wne[32], s_whv[32];
  __shared__ int s_base_heavy, s_base_light;
  const int b = blockIdx.x, t = threadIdx.x, tiles = a.tiles, lane = t & 31, w = t >> 5;
  if (t == 0) {
    if (a.zero_hit_count) a.zero_hit_count[b] = 0;
    if (a.zero_ovf_count && b == 0) a.zero_ovf_count[0] = 0;
  }
  const int32_t* gcount = a.bin_count + (size_t)b * tiles;
  int32_t* gstart = a.bin_start + (size_t)b * tiles;
  int32_t* gcur = a.bin_cursor + (size_t)b * tiles;
  const int per = (tiles + SETUP_THREADS - 1) / SETUP_THREADS;
  const int lo = t * per, hi = min(lo + per, tiles);
  int local = 0, ne = 0, hv = 0;
  for (int i = lo; i < hi; ++i) {
    int n = gcount[i];
    local += n; ne += n > 0 ? 1 : 0; hv += n > HEAVY_FACES ? 1 : 0;
  }
  int incl = local, ine = ne, ihv = hv;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int x = __shfl_up_sync(0xffffffffu, incl, o), y = __shfl_up_sync(0xffffffffu, ine, o), z = __shfl_up_sync(0xffffffffu, ihv, o);
    if (lane >= o) { incl += x; ine += y; ihv += z; }
  }
  if (lane == 31) { s_wsum[w] = incl; s_wne[w] = ine; s_whv[w] = ihv; }
  __syncthreads();
  if (w == 0) {
    int v = s_wsum[lane], y = s_wne[lane], z = s_whv[lane], iv = v, iy = y, iz = z;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n1 = __shfl_up_sync(0xffffffffu, iv, o), n2 = __shfl_up_sync(0xffffffffu, iy, o), n3 = __shfl_up_sync(0xffffffffu, iz, o);
      if (lane >= o) { iv += n1; iy += n2; iz += n3; }
    }
    s_wsum[lane] = iv - v; s_wne[lane] = iy - y; s_whv[lane] = iz - z;
    if (lane == 31) {
      if (iv > c.bin_cap) atomicAdd(c.counters + 1, 1ULL);
      s_base_heavy = atomicAdd(a.work_ctl + 0, iz);
      s_base_light = atomicAdd(a.work_ctl + 1, iy - iz);
    }
  }
  __syncthreads();
  int run = s_wsum[w] + incl - local;
  int k_hv = s_whv[w] + ihv - hv, k_lt = (s_wne[w] + ine - ne) - k_hv;  // exclusive ranks among the frame's heavy / light tiles
  for (int i = lo; i < hi; ++i) {
    int n = gcount[i];
    gstart[i] = run;
    gcur[i] = run;
    if (n > 0) {
      int nn = run + n > c.bin_cap ? max(0, (int)c.bin_cap - run) : n;
      int4 item = make_int4(b, i, run, nn);
      if (n > HEAVY_FACES) { int slot = s_base_heavy + k_hv++; if (slot < a.work_cap) a.work[slot] = item; }
      else { int slot = a.work_cap - 1 - (s_base_light + k_lt++); if (slot >= 0) a.work[slot] = item; }
    }
    run += n;
  }
}

// ---- K3: one thread per (frame, face): append the face to the list of every tile its record's mask keeps
__global__ void __launch_bounds__(256) bin_fill_kernel(harp_ctx c, SetupArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B * c.F) return;
  const int b = idx / c.F, f = idx % c.F, S = a.S, tw = S / HARP_TILE;
  const FaceTile* rec = a.rec + (size_t)b * c.F + f;
  const int4 tail = __ldg(reinterpret_cast<const int4*>(rec) + 8);  // cols, rows, packed tile range, tile mask
  if ((tail.x & 0xffff) > (tail.x >> 16)) return;                     // skipped face
  const int tx0 = tail.z & 0xff, ty0 = (tail.z >> 8) & 0xff, ntx = (tail.z >> 16) & 0xff, nty = (tail.z >> 24) & 0xff;
  int32_t* cur = a.bin_cursor + (size_t)b * a.tiles;
  int32_t* list = a.bin_list + (size_t)b * c.bin_cap;
  if (ntx * nty <= 32) {
    unsigned mask = (unsigned)tail.w;
    while (mask) {
      int k = __ffs(mask) - 1;
      mask &= mask - 1;
      int slot = atomicAdd(cur + (ty0 + k / ntx) * tw + tx0 + k % ntx, 1);
      if (slot < c.bin_cap) list[slot] = f;
    }
    return;
  }
  // a face spanning many tiles: repeat K1's per-tile test (same arithmetic, so counts and fills agree)
  const float4* r4 = reinterpret_cast<const float4*>(rec);
  float4 q0 = __ldg(r4), q1 = __ldg(r4 + 1), q7 = __ldg(r4 + 7);
  float pa[3] = {q0.x, q0.y, q7.x}, pb[3] = {q0.z, q0.w, q7.y}, pd[3] = {q1.x, q1.y, q7.z};
  TileRange tr;
  tile_range(pa, pb, pd, a.radius, tail.x & 0xffff, tail.x >> 16, tail.y & 0xffff, tail.y >> 16, tr);
  const float inv_s = 1.f / (float)S;
  for (int ty = tr.ty0; ty <= tr.ty1; ++ty)
    for (int tx = tr.tx0; tx <= tr.tx1; ++tx) {
      if (tr.edge_test) {
        float xa = pix_to_ndc_fast(tx * HARP_TILE, inv_s), xb = pix_to_ndc_fast(tx * HARP_TILE + HARP_TILE - 1, inv_s);
        float ya = pix_to_ndc_fast(ty * HARP_TILE, inv_s), yb = pix_to_ndc_fast(ty * HARP_TILE + HARP_TILE - 1, inv_s);
        if (tile_outside_face(tr, pa, pb, pd, xa, xb, ya, yb)) continue;
      }
      int slot = atomicAdd(cur + ty * tw + tx, 1);
      if (slot < c.bin_cap) list[slot] = f;
    }
}

}  // namespace

// Project `verts` with (R, T) and build everything the tile kernels of one view consume (0 = camera, 1 = light).
int harp_raster_setup(harp_ctx* c, int view, int B, int S, float focal, float radius, const float* verts, const float* R,
                      const float* T, cudaStream_t st, bool zero_grads) {
  HARP_REQUIRE(S / HARP_TILE <= 255, HARP_ERR_CAPACITY, "harp_raster_setup: images up to 2040 x 2040");
  const size_t voff = (size_t)view * c->maxB * c->tiles;
  SetupArgs a{};
  a.B = B; a.S = S; a.tiles = c->tiles; a.focal = focal; a.radius = radius; a.verts = verts; a.R = R; a.T = T;
  a.ndc = view == 0 ? c->ndc_cam : c->ndc_light;
  a.rec = c->face_rec + (size_t)view * c->maxB * c->F;
  a.bin_count = c->bin_count + voff; a.bin_start = c->bin_start + voff; a.bin_cursor = c->bin_fill + voff;
  a.bin_list = c->bin_list + (size_t)view * c->maxB * c->bin_cap;
  a.work_cap = c->maxB * c->tiles;
  a.work = c->work + (size_t)view * a.work_cap;
  a.work_ctl = c->work_ctl + view * 4;
  if (zero_grads) {
    a.zero_d_ndc = view == 0 ? c->d_ndc_cam : c->d_ndc_light;
    if (view == 0) { a.zero_hit_count = c->hit_count; a.zero_ovf_count = c->ovf_count; }
  }
  // one zero fill for the tile counters and the view's queue control words (allocated back to back: ctx.cu)
  HARP_CHECK_CUDA(cudaMemsetAsync(a.work_ctl, 0, 4 * sizeof(int32_t), st));
  HARP_CHECK_CUDA(cudaMemsetAsync(a.bin_count, 0, (size_t)B * c->tiles * sizeof(int32_t), st));
  const int n = std::max(c->F, c->V);
  face_setup_kernel<<<(int)(((size_t)B * n + 255) / 256), 256, 0, st>>>(*c, a);
  bin_scan_kernel<<<B, SETUP_THREADS, 0, st>>>(*c, a);
  bin_fill_kernel<<<(int)(((size_t)B * c->F + 255) / 256), 256, 0, st>>>(*c, a);
  HARP_LAUNCH_CHECK_N(3);
  return HARP_OK;
}
