// Shared rasterisation machinery: projection, 8x8-pixel tile binning, and the per-(pixel, face) evaluation that restates
// PyTorch3D v0.6.2's CheckPixelInsideFace operation by operation.
//
// Numerics: the silhouette's sigmoid has sigma = 1e-7 in squared-NDC units, so alpha in the +-1/3 px band around every
// edge is sensitive to single-ulp differences in the distance computation.  The per-pair arithmetic therefore uses
// explicitly rounded, *unfused* fp32 operations in the oracle's operation order (the reference's ATen/CUDA kernels do not
// contract across these statements either); everything else may use FMA.
#pragma once
#include "common.cuh"

#define HARP_TILE 8
#define HARP_K_EPS 1e-8f

__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// pixel index (0 = left/top) -> NDC coordinate of its centre; PyTorch3D reverses both axes (+X left, +Y up)
__device__ __forceinline__ float pix_to_ndc(int idx, int S) {
  float i = (float)(S - 1 - idx);
  return fadd(-1.0f, fdiv(fadd(fmul(2.0f, i), 1.0f), (float)S));
}

// EdgeFunctionForward(p, a, b) = (p.x-a.x)*(b.y-a.y) - (p.y-a.y)*(b.x-a.x)
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  return fsub(fmul(fsub(px, ax), fsub(by, ay)), fmul(fsub(py, ay), fsub(bx, ax)));
}

// PointLineDistanceForward: squared distance from p to the segment (a,b); also returns the clamped parameter
__device__ __forceinline__ float point_seg_dist(float px, float py, float ax, float ay, float bx, float by, float* t_out) {
  float bax = fsub(bx, ax), bay = fsub(by, ay);
  float l2 = fadd(fmul(bax, bax), fmul(bay, bay));
  if (l2 <= HARP_K_EPS) {
    float ex = fsub(px, bx), ey = fsub(py, by);
    *t_out = -1.f;  // degenerate marker
    return fadd(fmul(ex, ex), fmul(ey, ey));
  }
  float t = fdiv(fadd(fmul(bax, fsub(px, ax)), fmul(bay, fsub(py, ay))), l2);
  t = fminf(fmaxf(t, 0.f), 1.f);
  float qx = fadd(ax, fmul(t, bax)), qy = fadd(ay, fmul(t, bay));
  float dx = fsub(qx, px), dy = fsub(qy, py);
  *t_out = t;
  return fadd(fmul(dx, dx), fmul(dy, dy));
}

struct FaceRec {  // staged in shared memory, one per face of the current chunk (struct-of-arrays)
  float x0, y0, z0, x1, y1, z1, x2, y2, z2;
  float xmin, xmax, ymin, ymax;  // bbox inflated by sqrt(blur)
  float area;                    // EdgeFunction(v2, v0, v1) + kEpsilon
  float l12, l20, l01;           // squared lengths of the edges opposite v0, v1, v2
  float zmin;
  float inv_area, iz0, iz1, iz2;  // reciprocals for the fast 1/depth ordering (visibility only; never for output values)
  int v0, v1, v2, fid;
};

struct PairEval {
  bool pass, inside, saturated;
  float pz, dist;      // dist = unsigned squared distance to the triangle boundary
  float b0, b1, b2;    // perspective-correct (and, for the silhouette, clipped) barycentrics
  float w0, w1, w2;    // screen-space barycentrics
  int edge;            // 0: (v0,v1)  1: (v0,v2)  2: (v1,v2)  closest edge
  float t;             // clamped parameter on that edge (-1: degenerate edge)
};

// One (pixel, face) pair, in two stages.
//
// classify_pair — cheap, fully convergent (exactness argued in DESIGN.md section 4): with all three view depths > 0 the sign
// of every perspective-correct barycentric equals the sign of e_i / area, so strict-inside is decided by the edge functions
// alone.  A point's distance to the triangle boundary is >= its distance to any edge LINE, |e_i| / |edge_i|, hence
//   * inside and min_i e_i^2/l_i >= sat  (sat = 17.5 sigma; 1 - sigmoid(x) == 0 in fp32 for x >= 16.64)
//       => the silhouette factor is exactly 0: PAIR_SATURATED, no distance computation needed;
//   * outside and e_i^2/l_i >= 1.001 blur for a violated edge => dist >= blur_radius => PyTorch3D skips the pair.
// eval_pair_full — everything in between, operation by operation as PyTorch3D's CheckPixelInsideFace.
enum { PAIR_REJECT = 0, PAIR_SATURATED = 1, PAIR_FULL = 2 };

template <bool NEED_DIST>
__device__ __forceinline__ int classify_pair(const FaceRec& f, float px, float py, float blur, float sat) {
  if (px > f.xmax || px < f.xmin || py > f.ymax || py < f.ymin) return PAIR_REJECT;
  if (f.zmin > 0.f) {
    float e0 = edge_fn(px, py, f.x1, f.y1, f.x2, f.y2);
    float e1 = edge_fn(px, py, f.x2, f.y2, f.x0, f.y0);
    float e2 = edge_fn(px, py, f.x0, f.y0, f.x1, f.y1);
    bool pos = f.area > 0.f;
    bool i0 = pos ? e0 > 0.f : e0 < 0.f, i1 = pos ? e1 > 0.f : e1 < 0.f, i2 = pos ? e2 > 0.f : e2 < 0.f;
    if (i0 && i1 && i2) {
      if (NEED_DIST && e0 * e0 >= sat * f.l12 && e1 * e1 >= sat * f.l20 && e2 * e2 >= sat * f.l01) return PAIR_SATURATED;
    } else {
      if (!NEED_DIST) return PAIR_REJECT;
      float far = 1.001f * blur;
      if ((!i0 && e0 * e0 >= far * f.l12) || (!i1 && e1 * e1 >= far * f.l20) || (!i2 && e2 * e2 >= far * f.l01)) return PAIR_REJECT;
    }
  }
  return PAIR_FULL;
}

// The fast inverse depth q = 1 / (perspective-correct depth) the tile rasteriser orders inside pairs by (fast_inv_depth below):
// with w_i = e_i / area and b_i ~ w_i / z_i, pz = sum b_i z_i = sum(w_i) / sum(w_i / z_i) = sum(e_i) / sum(e_i / z_i) (the
// screen-space weights of a sliver face need not sum to 1 in fp32, hence the normalisation).  q orders the faces covering a pixel
// (nearest = largest q); the winner's barycentrics and depth are then recomputed with the exact operation sequence.

// CLIP: clip_barycentric_coords (silhouette); blur: blur_radius (squared NDC).  The bounding-box test is classify_pair's.
template <bool CLIP, bool NEED_DIST, typename F>
__device__ __forceinline__ void eval_pair_full(const F& f, float px, float py, float blur, PairEval& r) {
  r.pass = false;
  r.saturated = false;
  float w0 = fdiv(edge_fn(px, py, f.x1, f.y1, f.x2, f.y2), f.area);
  float w1 = fdiv(edge_fn(px, py, f.x2, f.y2, f.x0, f.y0), f.area);
  float w2 = fdiv(edge_fn(px, py, f.x0, f.y0, f.x1, f.y1), f.area);
  // BarycentricPerspectiveCorrectionForward
  float t0 = fmul(fmul(w0, f.z1), f.z2), t1 = fmul(fmul(f.z0, w1), f.z2), t2 = fmul(fmul(f.z0, f.z1), w2);
  float den = fmaxf(fadd(fadd(t0, t1), t2), HARP_K_EPS);
  float b0 = fdiv(t0, den), b1 = fdiv(t1, den), b2 = fdiv(t2, den);
  bool inside = b0 > 0.f && b1 > 0.f && b2 > 0.f;
  float c0 = b0, c1 = b1, c2 = b2;
  if (CLIP) {  // BarycentricClipForward
    c0 = fmaxf(b0, 0.f); c1 = fmaxf(b1, 0.f); c2 = fmaxf(b2, 0.f);
    float s = fmaxf(fadd(fadd(c0, c1), c2), 1e-5f);
    c0 = fdiv(c0, s); c1 = fdiv(c1, s); c2 = fdiv(c2, s);
  }
  float pz = fadd(fadd(fmul(c0, f.z0), fmul(c1, f.z1)), fmul(c2, f.z2));
  if (pz < 0.f) return;
  r.inside = inside;
  r.pz = pz;
  r.b0 = c0; r.b1 = c1; r.b2 = c2;
  r.w0 = w0; r.w1 = w1; r.w2 = w2;
  if (NEED_DIST) {
    float ta, tb, tc;
    float d01 = point_seg_dist(px, py, f.x0, f.y0, f.x1, f.y1, &ta);
    float d02 = point_seg_dist(px, py, f.x0, f.y0, f.x2, f.y2, &tb);
    float d12 = point_seg_dist(px, py, f.x1, f.y1, f.x2, f.y2, &tc);
    // torch.min over the stacked [d01, d02, d12]: first minimal index (matches PointTriangleDistanceBackward's order)
    float d = d01; int e = 0; float t = ta;
    if (d02 < d) { d = d02; e = 1; t = tb; }
    if (d12 < d) { d = d12; e = 2; t = tc; }
    r.dist = d; r.edge = e; r.t = t;
    if (!inside && d >= blur) return;
  } else {
    if (!inside) return;
  }
  r.pass = true;
}

template <bool CLIP, bool NEED_DIST>
__device__ __forceinline__ void eval_pair(const FaceRec& f, float px, float py, float blur, float sat, PairEval& r) {
  r.pass = false;
  r.saturated = false;
  int cls = classify_pair<NEED_DIST>(f, px, py, blur, sat);
  if (cls == PAIR_REJECT) return;
  if (cls == PAIR_SATURATED) { r.pass = true; r.inside = true; r.saturated = true; return; }
  eval_pair_full<CLIP, NEED_DIST>(f, px, py, blur, r);
}

// Load face `fid` of frame-local projected verts into a FaceRec.  Returns false for faces PyTorch3D skips outright.
__device__ __forceinline__ void load_face(const harp_ctx& c, const float* ndc, int fid, float radius, FaceRec& f) {
  const int32_t* fv = c.faces + (size_t)fid * 3;
  f.v0 = fv[0]; f.v1 = fv[1]; f.v2 = fv[2]; f.fid = fid;
  const float *a = ndc + (size_t)f.v0 * 3, *b = ndc + (size_t)f.v1 * 3, *d = ndc + (size_t)f.v2 * 3;
  f.x0 = a[0]; f.y0 = a[1]; f.z0 = a[2];
  f.x1 = b[0]; f.y1 = b[1]; f.z1 = b[2];
  f.x2 = d[0]; f.y2 = d[1]; f.z2 = d[2];
  f.xmin = fsub(fminf(fminf(f.x0, f.x1), f.x2), radius);
  f.xmax = fadd(fmaxf(fmaxf(f.x0, f.x1), f.x2), radius);
  f.ymin = fsub(fminf(fminf(f.y0, f.y1), f.y2), radius);
  f.ymax = fadd(fmaxf(fmaxf(f.y0, f.y1), f.y2), radius);
  f.area = fadd(edge_fn(f.x2, f.y2, f.x0, f.y0, f.x1, f.y1), HARP_K_EPS);
  f.l12 = (f.x2 - f.x1) * (f.x2 - f.x1) + (f.y2 - f.y1) * (f.y2 - f.y1);
  f.l20 = (f.x0 - f.x2) * (f.x0 - f.x2) + (f.y0 - f.y2) * (f.y0 - f.y2);
  f.l01 = (f.x1 - f.x0) * (f.x1 - f.x0) + (f.y1 - f.y0) * (f.y1 - f.y0);
  f.zmin = fminf(fminf(f.z0, f.z1), f.z2);
  f.inv_area = 1.f / f.area;
  f.iz0 = 1.f / f.z0; f.iz1 = 1.f / f.z1; f.iz2 = 1.f / f.z2;
}

// The shared-memory record of the tile kernels' face walk: eight 16-byte groups (one LDS.128 each, broadcast to the warp)
// holding per-face constants so that the per-(pixel, face) classification is ~35 branch-free instructions:
//   * the three edge functions in affine form, already multiplied by sign(area):  t_i(p) = a_i p.x + b_i p.y + c_i  (2 FMAs);
//     t_i > 0 for all i  <=>  strictly inside.  The affine value differs from PyTorch3D's own EdgeFunctionForward rounding
//     by at most ~9 ulp of  m_i = |a_i| (1 + |x_anchor|) + |b_i| (1 + |y_anchor|);  eps = 1e-6 max_i m_i (16 ulp) is stored and
//     every pair with min_i |t_i| < eps is re-classified with the bit-exact edge functions (about 1 in-bbox pair in 1700, tests/test_classification_bounds.py), so the
//     strict-inside decision is PyTorch3D's bit for bit.  A face with a vertex behind the camera has eps = +inf;
//   * il_i = 1 / |edge_i|^2: the squared distance to an edge line is t_i^2 il_i, compared against uniform thresholds with
//     margins (2 % / 5 %) an order of magnitude above its relative error at those thresholds; never used for output values;
//   * iz_i = 1 / z_i for the fast inverse depth that orders visibility candidates; the vertices, depths and area for the
//     exact evaluation of the pairs that need it.
struct __align__(16) FaceTile {
  float x0, y0, x1, y1;
  float x2, y2, xmin, xmax;
  float ymin, ymax, eps, area;
  float a0, b0, c0, il12;   // t0 = sgn * EdgeFunction(p, v1, v2)
  float a1, b1, c1, il20;   // t1 = sgn * EdgeFunction(p, v2, v0)
  float a2, b2, c2, il01;   // t2 = sgn * EdgeFunction(p, v0, v1)
  float iz0, iz1, iz2, sgn;
  float z0, z1, z2;
  int fid;
  // pixel rectangle of the inflated bbox (0.01 px guard; the exact float bbox test still runs per pair): columns
  // jlo | jhi << 16, rows ilo | ihi << 16; jlo > jhi marks a face the rasteriser skips outright (behind the camera, zero area,
  // off screen).  Written by the per-frame setup kernel (setup.cu), which owns the binning; 
  int cols, rows;
  int pad0, pad1;
};  // 144 B = 9 x 16 B: consecutive records start 4 banks apart, so the 2-3 records a warp touches per LDS.128 do not collide.
    // The vertex ids (needed by the rare gradient path only) are re-read from the face table.

__device__ __forceinline__ void make_face_record(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2,
                                                 float z2, int fid, float radius, FaceTile& f) {
  f.x0 = x0; f.y0 = y0; f.x1 = x1; f.y1 = y1; f.x2 = x2; f.y2 = y2;
  f.z0 = z0; f.z1 = z1; f.z2 = z2; f.fid = fid;
  f.xmin = fsub(fminf(fminf(x0, x1), x2), radius);
  f.xmax = fadd(fmaxf(fmaxf(x0, x1), x2), radius);
  f.ymin = fsub(fminf(fminf(y0, y1), y2), radius);
  f.ymax = fadd(fmaxf(fmaxf(y0, y1), y2), radius);
  float area = fadd(edge_fn(x2, y2, x0, y0, x1, y1), HARP_K_EPS);
  f.area = area;
  bool front = fminf(fminf(z0, z1), z2) > 0.f;
  float sg = front ? (area > 0.f ? 1.f : -1.f) : 0.f;
  f.sgn = sg;
  // edge deltas rounded exactly as edge_fn rounds them: EdgeFunction(p, A, B) = (p.x - A.x) dy - (p.y - A.y) dx
  float d12x = fsub(x2, x1), d12y = fsub(y2, y1);  // anchor v1
  float d20x = fsub(x0, x2), d20y = fsub(y0, y2);  // anchor v2
  float d01x = fsub(x1, x0), d01y = fsub(y1, y0);  // anchor v0
  f.a0 = sg * d12y; f.b0 = -sg * d12x; f.c0 = sg * (y1 * d12x - x1 * d12y);
  f.a1 = sg * d20y; f.b1 = -sg * d20x; f.c1 = sg * (y2 * d20x - x2 * d20y);
  f.a2 = sg * d01y; f.b2 = -sg * d01x; f.c2 = sg * (y0 * d01x - x0 * d01y);
  float m0 = fabsf(d12y) * (1.f + fabsf(x1)) + fabsf(d12x) * (1.f + fabsf(y1));
  float m1 = fabsf(d20y) * (1.f + fabsf(x2)) + fabsf(d20x) * (1.f + fabsf(y2));
  float m2 = fabsf(d01y) * (1.f + fabsf(x0)) + fabsf(d01x) * (1.f + fabsf(y0));
  f.eps = front ? 1e-6f * fmaxf(fmaxf(m0, m1), m2) : __int_as_float(0x7f800000);
  f.il12 = 1.f / (d12x * d12x + d12y * d12y);
  f.il20 = 1.f / (d20x * d20x + d20y * d20y);
  f.il01 = 1.f / (d01x * d01x + d01y * d01y);
  f.iz0 = 1.f / z0; f.iz1 = 1.f / z1; f.iz2 = 1.f / z2;
  f.cols = f.rows = f.pad0 = f.pad1 = 0;
}

// Classification of one (pixel, face) pair against a staged FaceTile; same decisions as classify_pair (the exactness
// argument is above it): returns PAIR_REJECT / PAIR_SATURATED / PAIR_FULL, or PAIR_EXACT for a face with a
// vertex behind the camera.  `inside` = strict inside as PyTorch3D's bit-exact edge functions decide it; t[] = signed edge
// functions (for the fast inverse depth).  With NEED_DIST = false (visibility only) inside pairs report PAIR_SATURATED, all
// others REJECT.  `far` must be >= 1.02 blur and `sat` >= 17.5 sigma (margins, see FaceTile).
enum { PAIR_EXACT = 3 };

__device__ __forceinline__ bool in_face_bbox(const FaceTile& f, float px, float py) {
  return !((px > f.xmax) | (px < f.xmin) | (py > f.ymax) | (py < f.ymin));  // bitwise: no short-circuit branches
}

template <bool NEED_DIST>
__device__ __forceinline__ int classify_from_t(const FaceTile& f, float t0, float t1, float t2, float far, float sat, bool inb,
                                               bool& inside) {
  bool i0 = t0 > 0.f, i1 = t1 > 0.f, i2 = t2 > 0.f;
  inside = inb & i0 & i1 & i2;
  if (!NEED_DIST) return inside ? PAIR_SATURATED : PAIR_REJECT;
  float u0 = t0 * t0 * f.il12, u1 = t1 * t1 * f.il20, u2 = t2 * t2 * f.il01;
  bool sat_all = (u0 >= sat) & (u1 >= sat) & (u2 >= sat);
  bool far_any = (!i0 & (u0 >= far)) | (!i1 & (u1 >= far)) | (!i2 & (u2 >= far));
  int cls = inside ? (sat_all ? PAIR_SATURATED : PAIR_FULL) : (far_any ? PAIR_REJECT : PAIR_FULL);
  return inb ? cls : PAIR_REJECT;
}

template <bool NEED_DIST>
__device__ __forceinline__ int classify_fast(const FaceTile& f, float px, float py, float far, float sat, bool inb, bool& inside,
                                             float& t0, float& t1, float& t2) {
  t0 = fmaf(f.a0, px, fmaf(f.b0, py, f.c0));
  t1 = fmaf(f.a1, px, fmaf(f.b1, py, f.c1));
  t2 = fmaf(f.a2, px, fmaf(f.b2, py, f.c2));
  int cls = classify_from_t<NEED_DIST>(f, t0, t1, t2, far, sat, inb, inside);
  bool ambiguous = inb & (fminf(fminf(fabsf(t0), fabsf(t1)), fabsf(t2)) < f.eps);
  if (ambiguous) {  // rare: a sign within rounding distance of zero (or a face crossing the camera plane)
    if (f.sgn == 0.f) return PAIR_EXACT;
    t0 = f.sgn * edge_fn(px, py, f.x1, f.y1, f.x2, f.y2);
    t1 = f.sgn * edge_fn(px, py, f.x2, f.y2, f.x0, f.y0);
    t2 = f.sgn * edge_fn(px, py, f.x0, f.y0, f.x1, f.y1);
    cls = classify_from_t<NEED_DIST>(f, t0, t1, t2, far, sat, inb, inside);
  }
  return cls;
}

// fast inverse depth of an inside pair: sum(w_i / z_i) / sum(w_i) with w_i ~ e_i (independent of the area)
__device__ __forceinline__ float fast_inv_depth(const FaceTile& f, float e0, float e1, float e2) {
  return __fdividef(e0 * f.iz0 + e1 * f.iz1 + e2 * f.iz2, e0 + e1 + e2);
}

// Backward of one silhouette factor: alpha = 1 - prod_k(1 - p_k), p = sigmoid(-sd/sigma)  =>  d alpha / d sd_k = -prod p_k / sigma
// (only reached when no factor is exactly zero), then PointLineDistanceBackward of the closest edge:
// q = a + t (b - a), d = q - p ; grad_a = g (1-t) 2 d ; grad_b = g t 2 d.   d_ndc rows are (x, y, z, pad).
__device__ __forceinline__ void face_verts(const harp_ctx&, const FaceRec& f, int& v0, int& v1, int& v2) { v0 = f.v0; v1 = f.v1; v2 = f.v2; }
__device__ __forceinline__ void face_verts(const harp_ctx& c, const FaceTile& f, int& v0, int& v1, int& v2) {
  const int32_t* fv = c.faces + (size_t)f.fid * 3;
  v0 = fv[0]; v1 = fv[1]; v2 = fv[2];
}

template <typename F>
__device__ __forceinline__ void sil_pair_gradient(const harp_ctx& c, const F& f, float px, float py, float sd, int edge, float t, float g,
                                                  float prod, float sigma, float* d_ndc) {
  float prob = 1.f / (1.f + expf(fdiv(sd, sigma)));
  float gsd = g * (-prod * prob / sigma);
  float gd = sd < 0.f ? -gsd : gsd;  // w.r.t. the unsigned squared distance (inside <=> signed distance negative)
  if (gd == 0.f) return;
  float ax, ay, bx, by; int va, vb, fv0, fv1, fv2;
  face_verts(c, f, fv0, fv1, fv2);
  if (edge == 0) { ax = f.x0; ay = f.y0; bx = f.x1; by = f.y1; va = fv0; vb = fv1; }
  else if (edge == 1) { ax = f.x0; ay = f.y0; bx = f.x2; by = f.y2; va = fv0; vb = fv2; }
  else { ax = f.x1; ay = f.y1; bx = f.x2; by = f.y2; va = fv1; vb = fv2; }
  if (t < 0.f) {  // degenerate edge: dist = |p - b|^2
    atomicAdd(reinterpret_cast<float2*>(d_ndc + (size_t)vb * 4), make_float2(gd * 2.f * (bx - px), gd * 2.f * (by - py)));
    return;
  }
  float qx = ax + t * (bx - ax), qy = ay + t * (by - ay);
  float dx = 2.f * gd * (qx - px), dy = 2.f * gd * (qy - py);
  atomicAdd(reinterpret_cast<float2*>(d_ndc + (size_t)va * 4), make_float2((1.f - t) * dx, (1.f - t) * dy));
  atomicAdd(reinterpret_cast<float2*>(d_ndc + (size_t)vb * 4), make_float2(t * dx, t * dy));
}

// A pixel whose candidate count may exceed faces_per_pixel is deferred to the exact top-K kernel (raster.cu).
__device__ __forceinline__ void sil_defer_overflow(const harp_ctx& c, int pixel_id) {
  int slot = atomicAdd(c.ovf_count, 1);
  if (slot < c.ovf_cap) c.ovf_list[slot] = pixel_id;
  else atomicAdd(c.counters, 1ULL);  // list full: this pixel keeps the all-faces product (reported by harp_ctx_read_counters)
}
