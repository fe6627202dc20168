// Per-pixel shading + analytic backward of the shadowed PBR colour render, as a device function.
// (renderer/renderer_helper.py:331-412,472-592; renderer/pbr_materials.py:58-124)
#pragma once
#include "raster.cuh"

struct ShadeArgs {
  int B, S, self_shadow, texH, texW;
  int fixed_lights;      // 1: Phong twin with constant light colours ka, kd, ks (renderer_helper.py:70-73); 0: ka = amb, kd = 1 - amb, ks = 0
  float ka, kd, ks;
  float focal, gscale, inv_n;
  const float *verts, *normals, *ndc_cam, *light_R, *light_T, *light_pos, *amb;
  const float4 *tex4, *nm4;  // packed albedo / normalised normal map (harp_tex_prepare)
  const float* light_zbuf;
  const int32_t *bin_count, *bin_start, *bin_list;
  float* rgb;
  const float *gin, *mask;
  float* loss;
  float *d_ndc_cam, *d_verts, *d_normals, *d_light_R, *d_light_T, *d_light_pos, *d_amb, *d_texture, *d_normal_map;
  float* d_light_zbuf;
};

// nearest face covering the pixel centre: strict inside, smallest perspective-correct depth, ties -> lower face id
struct Hit {
  int fid;
  float pz, b0, b1, b2, w0, w1, w2;
};

// ---- barycentric backward: grads of the perspective-correct barycentrics -> NDC xy and view z of the 3 face verts ------
// b_i = t_i / sum(t), t_0 = w0 z1 z2 ..., w_i = e_i / A  (BarycentricPerspectiveCorrection / BarycentricCoords backward)
struct Tri2 {  // projected triangle: NDC xy and view z of the three vertices
  float x0, y0, z0, x1, y1, z1, x2, y2, z2;
};
__device__ __forceinline__ Tri2 tri_from_ndc(const float* ndc, int v0, int v1, int v2) {
  const float *A0 = ndc + (size_t)v0 * 3, *A1 = ndc + (size_t)v1 * 3, *A2 = ndc + (size_t)v2 * 3;
  return Tri2{A0[0], A0[1], A0[2], A1[0], A1[1], A1[2], A2[0], A2[1], A2[2]};
}

__device__ __forceinline__ void bary_backward(const Tri2& t, int v0, int v1, int v2, float px, float py, const Hit& h,
                                              const float gb[3], float gz_direct[3], float* d_ndc) {
  float x0 = t.x0, y0 = t.y0, z0 = t.z0, x1 = t.x1, y1 = t.y1, z1 = t.z1, x2 = t.x2, y2 = t.y2, z2 = t.z2;
  float den = h.w0 * z1 * z2 + z0 * h.w1 * z2 + z0 * z1 * h.w2;
  float gz0 = gz_direct[0], gz1 = gz_direct[1], gz2 = gz_direct[2];
  float gw0 = 0.f, gw1 = 0.f, gw2 = 0.f;
  if (den > HARP_K_EPS) {
    float dotgb = gb[0] * h.b0 + gb[1] * h.b1 + gb[2] * h.b2;
    float iden = __fdividef(1.f, den);
    float gt0 = (gb[0] - dotgb) * iden, gt1 = (gb[1] - dotgb) * iden, gt2 = (gb[2] - dotgb) * iden;
    gw0 = gt0 * z1 * z2; gz1 += gt0 * h.w0 * z2; gz2 += gt0 * h.w0 * z1;
    gw1 = gt1 * z0 * z2; gz0 += gt1 * h.w1 * z2; gz2 += gt1 * z0 * h.w1;
    gw2 = gt2 * z0 * z1; gz0 += gt2 * z1 * h.w2; gz1 += gt2 * z0 * h.w2;
  }
  float area = (x2 - x0) * (y1 - y0) - (y2 - y0) * (x1 - x0) + HARP_K_EPS;
  float iarea = __fdividef(1.f, area);
  float ge0 = gw0 * iarea, ge1 = gw1 * iarea, ge2 = gw2 * iarea;
  float gA = -(gw0 * h.w0 + gw1 * h.w1 + gw2 * h.w2) * iarea;
  // E(p,a,b) = (px-ax)(by-ay) - (py-ay)(bx-ax):  dE/da = (py-by, bx-px),  dE/db = (ay-py, px-ax)
  // e0 = E(p,v1,v2), e1 = E(p,v2,v0), e2 = E(p,v0,v1);  A = E(v2,v0,v1) + eps
  float gx0 = ge1 * (y2 - py) + ge2 * (py - y1) + gA * (y2 - y1);
  float gy0 = ge1 * (px - x2) + ge2 * (x1 - px) + gA * (x1 - x2);
  float gx1 = ge0 * (py - y2) + ge2 * (y0 - py) + gA * (y0 - y2);
  float gy1 = ge0 * (x2 - px) + ge2 * (px - x0) + gA * (x2 - x0);
  float gx2 = ge0 * (y1 - py) + ge1 * (py - y0) + gA * (y1 - y0);
  float gy2 = ge0 * (px - x1) + ge1 * (x0 - px) + gA * (x0 - x1);
  atomic_add4(d_ndc + (size_t)v0 * 4, gx0, gy0, gz0);
  atomic_add4(d_ndc + (size_t)v1 * 4, gx1, gy1, gz1);
  atomic_add4(d_ndc + (size_t)v2 * 4, gx2, gy2, gz2);
}

// ---- texture fetch: F.grid_sample(bilinear, align_corners=True, border) on the H-flipped map ------------
struct Taps {
  int idx[4];     // texel linear indices (row*W + col) in the ORIGINAL (unflipped) map: nw, ne, sw, se
  float w[4];
  float fx, fy;
  float gmx, gmy; // d(ix)/du , d(iy)/dv including the border-clip mask
};

__device__ __forceinline__ void make_taps(float u, float v, int H, int W, Taps& t) {
  float gx = u * 2.f - 1.f, gy = v * 2.f - 1.f;
  float ix = ((gx + 1.f) / 2.f) * (float)(W - 1), iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
  t.gmx = (float)(W - 1); t.gmy = (float)(H - 1);
  if (ix <= 0.f) { ix = 0.f; t.gmx = 0.f; } else if (ix >= (float)(W - 1)) { ix = (float)(W - 1); t.gmx = 0.f; }
  if (iy <= 0.f) { iy = 0.f; t.gmy = 0.f; } else if (iy >= (float)(H - 1)) { iy = (float)(H - 1); t.gmy = 0.f; }
  float x0 = floorf(ix), y0 = floorf(iy);
  t.fx = ix - x0; t.fy = iy - y0;
  int c0 = (int)x0, r0 = (int)y0;
  int c1 = min(c0 + 1, W - 1), r1 = min(r0 + 1, H - 1);  // out-of-range taps carry zero weight
  int R0 = H - 1 - r0, R1 = H - 1 - r1;                  // undo the flip
  t.idx[0] = R0 * W + c0; t.idx[1] = R0 * W + c1; t.idx[2] = R1 * W + c0; t.idx[3] = R1 * W + c1;
  t.w[0] = (1.f - t.fx) * (1.f - t.fy); t.w[1] = t.fx * (1.f - t.fy); t.w[2] = (1.f - t.fx) * t.fy; t.w[3] = t.fx * t.fy;
}

__device__ __forceinline__ f3 normalize12(f3 x) {  // F.normalize default eps = 1e-12
  float inv = 1.f / fmaxf(sqrtf(dot(x, x)), 1e-12f);
  return mk3(x.x * inv, x.y * inv, x.z * inv);
}
// backward of F.normalize at a packed normal-map texel (n = m / max(|m|, eps), w = 1 / max(|m|, eps), negative in the eps case)
__device__ __forceinline__ f3 normalize_bwd_packed(float4 t, f3 gy) {
  f3 n = mk3(t.x, t.y, t.z);
  return t.w > 0.f ? t.w * (gy - dot(n, gy) * n) : (-t.w) * gy;
}
__device__ __forceinline__ f3 normalize_bwd_eps(f3 x, f3 gy, float eps) {
  float l2 = dot(x, x);
  if (l2 > eps * eps) {
    float il = rsqrtf(l2);
    f3 y = il * x;
    return il * (gy - dot(y, gy) * y);
  }
  return (1.f / eps) * gy;
}

// Per-vertex attributes and targets of one hit pixel, requested by the caller as soon as the face id is known so that their
// latency overlaps the exact barycentric evaluation (the deferred kernel's dependent-load chain is its limiter).
struct ShadePre {
  f3 P0, P1, P2, N0, N1, N2;
  float uv[6];
  float mk;
  f3 y;
};

template <int MODE>
__device__ __forceinline__ void shade_preload(const harp_ctx& c, const ShadeArgs& a, int b, size_t pix, int fid, int v0, int v1, int v2,
                                              ShadePre& p) {
  const float* vb = a.verts + (size_t)b * c.V * 3;
  const float* nb = a.normals + (size_t)b * c.V * 3;
  p.P0 = ld3(vb + (size_t)v0 * 3); p.P1 = ld3(vb + (size_t)v1 * 3); p.P2 = ld3(vb + (size_t)v2 * 3);
  p.N0 = ld3(nb + (size_t)v0 * 3); p.N1 = ld3(nb + (size_t)v1 * 3); p.N2 = ld3(nb + (size_t)v2 * 3);
  const float* fu = c.face_uvs + (size_t)fid * 6;
#pragma unroll
  for (int k = 0; k < 6; ++k) p.uv[k] = fu[k];
  if (MODE == 2) { p.mk = a.mask[pix]; p.y = ld3(a.gin + pix * 3); }
  else if (MODE == 1) { p.mk = 0.f; p.y = ld3(a.gin + pix * 3); }
}

struct ShadeOut {
  f3 rgb;
  float loss_abs;   // sum_c |rgb_c m - y_c m| of this pixel (MODE 2)
  float frame[8];   // per-frame scalar gradient partials: light_pos xyz, amb, light_R[:,2] xyz, light_T.z
  bool active;
};

// Per-frame scalar partials of one warp -> global accumulators (one atomic per warp and value).
__device__ __forceinline__ void shade_flush_frame_scalars(const ShadeArgs& a, int b, const float vals[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float s = warp_sum(vals[k]);
    if ((threadIdx.x & 31) == 0 && s != 0.f) {
      if (k < 3) { if (a.d_light_pos) atomicAdd(a.d_light_pos + (size_t)b * 3 + k, s); }
      else if (k == 3) { if (a.d_amb) atomicAdd(a.d_amb, s); }
      else if (k < 7) { if (a.d_light_R) atomicAdd(a.d_light_R + (size_t)b * 9 + (k - 4) * 3 + 2, s); }
      else { if (a.d_light_T) atomicAdd(a.d_light_T + (size_t)b * 3 + 2, s); }
    }
  }
}

// Shade one pixel whose nearest face and barycentrics are known (h.fid < 0: background), and, for MODE 1 / 2, scatter
// its gradients.  Shared by the tile kernel (shade.cu) and the deferred kernel over compacted hit pixels (pixel.cu).
// MODE 0 forward; 1 backward with upstream g_rgb; 2 fused photometric L1 on the eroded mask.
template <int MODE>
__device__ __forceinline__ void shade_pixel(const harp_ctx& c, const ShadeArgs& a, int b, size_t pix, float px, float py,
                                            const Hit& h, ShadeOut& o, const ShadePre* pre = nullptr, const Tri2* tri = nullptr) {
  const int S = a.S;
  const float* ndc = a.ndc_cam + (size_t)b * c.V * 3;
  f3 rgb = mk3(1.f, 1.f, 1.f);  // background (renderer_helper.py:430)
  // forward state kept for the backward half
  int v0 = 0, v1 = 0, v2 = 0;
  f3 P, N, tex, m, npre, n1, dvec, dh, n2;  // per-vertex attributes are re-read in the backward half (L1 hits) to keep registers free
  Taps tp;
  float amb = a.amb[0], cosang = 0.f, vis = 1.f, sh = 0.f, zL = 0.f, s_sign = 1.f, aa = 0.f;
  f3 uvec, vvec;
  int tap_xk = 0, tap_yk = 0;  // centre of the 3x3 PCF window (tap addresses are recomputed in the backward half)
  float tapsig[9];
  const float* lR = a.light_R + (size_t)b * 9;
  const float* lT = a.light_T + (size_t)b * 3;
  f3 L = ld3(a.light_pos + (size_t)b * 3);
  bool hit = h.fid >= 0;
  if (hit) {
    const int32_t* fv = c.faces + (size_t)h.fid * 3;
    v0 = fv[0]; v1 = fv[1]; v2 = fv[2];
    const float* vb = a.verts + (size_t)b * c.V * 3;
    const float* nb = a.normals + (size_t)b * c.V * 3;
    float u, v;
    if (pre) {
      P = h.b0 * pre->P0 + h.b1 * pre->P1 + h.b2 * pre->P2;
      N = h.b0 * pre->N0 + h.b1 * pre->N1 + h.b2 * pre->N2;
      u = h.b0 * pre->uv[0] + h.b1 * pre->uv[2] + h.b2 * pre->uv[4];
      v = h.b0 * pre->uv[1] + h.b1 * pre->uv[3] + h.b2 * pre->uv[5];
    } else {
      P = h.b0 * ld3(vb + (size_t)v0 * 3) + h.b1 * ld3(vb + (size_t)v1 * 3) + h.b2 * ld3(vb + (size_t)v2 * 3);
      N = h.b0 * ld3(nb + (size_t)v0 * 3) + h.b1 * ld3(nb + (size_t)v1 * 3) + h.b2 * ld3(nb + (size_t)v2 * 3);
      const float* fu = c.face_uvs + (size_t)h.fid * 6;
      u = h.b0 * fu[0] + h.b1 * fu[2] + h.b2 * fu[4];
      v = h.b0 * fu[1] + h.b1 * fu[3] + h.b2 * fu[5];
    }
    // the 9 PCF taps of the light-view depth buffer depend only on P: request them before the texture taps
    float ztap[9];
    if (a.self_shadow) {
      // hit point in light view space (row-vector: p R + T), its light-screen pixel, 3x3 clamped taps
      float xl = P.x * lR[0] + P.y * lR[3] + P.z * lR[6] + lT[0];
      float yl = P.x * lR[1] + P.y * lR[4] + P.z * lR[7] + lT[1];
      zL = P.x * lR[2] + P.y * lR[5] + P.z * lR[8] + lT[2];
      float hs = 0.5f * (float)S;
      float xs = hs - hs * ((a.focal * xl / zL) / hs), ys = hs - hs * ((a.focal * yl / zL) / hs);
      // .round().long(): round-half-even, then saturate like a float->int64 cast would never need to here
      float xr = rintf(xs), yr = rintf(ys);
      xr = fminf(fmaxf(xr, -2.f), (float)S + 1.f); yr = fminf(fmaxf(yr, -2.f), (float)S + 1.f);
      tap_xk = (int)xr; tap_yk = (int)yr;
      const float* zb = a.light_zbuf + (size_t)b * S * S;
#pragma unroll
      for (int ii = -1; ii <= 1; ++ii)
#pragma unroll
        for (int jj = -1; jj <= 1; ++jj) {
          int ry = min(max(tap_yk + ii, 0), S - 1), rx = min(max(tap_xk + jj, 0), S - 1);
          ztap[(ii + 1) * 3 + (jj + 1)] = zb[ry * S + rx];
        }
    }
    make_taps(u, v, a.texH, a.texW, tp);
    tex = mk3(0.f, 0.f, 0.f);
    m = mk3(0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 t = __ldg(a.tex4 + tp.idx[k]), nm = __ldg(a.nm4 + tp.idx[k]);
      tex = tex + tp.w[k] * mk3(t.x, t.y, t.z);
      m = m + tp.w[k] * mk3(nm.x, nm.y, nm.z);
    }
    // PBRMaterials.compute_tangent / apply_normal_map on the (un-normalised) interpolated normal
    s_sign = N.z >= 0.f ? 1.f : -1.f;
    aa = -1.f / (s_sign + N.z);
    float bb = N.x * N.y * aa;
    uvec = mk3(1.f + s_sign * N.x * N.x * aa, s_sign * bb, -s_sign * N.x);
    vvec = mk3(bb, s_sign + N.y * N.y * aa, -N.y);
    npre = (-m.x) * uvec + (-m.y) * vvec + m.z * N;
    n1 = normalize12(npre);
    // lighting: F.normalize(eps=1e-6) on both, relu(dot)
    float l1 = fmaxf(sqrtf(dot(n1, n1)), 1e-6f);
    n2 = (1.f / l1) * n1;
    dvec = L - P;
    float ld = fmaxf(sqrtf(dot(dvec, dvec)), 1e-6f);
    dh = (1.f / ld) * dvec;
    cosang = fmaxf(dot(n2, dh), 0.f);
    if (a.self_shadow) {
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float x = (ztap[t] - (zL - 0.008f)) * 1000.f;
        float sg = __fdividef(1.f, 1.f + __expf(-x));  // sigmoid; 2-ulp exp and division are far inside the 1e-4 image tolerance
        tapsig[t] = sg;
        acc += sg;
      }
      vis = acc / 9.f;
    }
    const float ka = a.fixed_lights ? a.ka : amb, kd = a.fixed_lights ? a.kd : 1.f - amb;
    sh = ka + kd * cosang * vis;
    rgb = sh * tex;
    if (a.fixed_lights) rgb = rgb + mk3(a.ks, a.ks, a.ks);  // specular^0 == 1 on every fragment
    // softmax_rgb_blend with K=1: (prob c + delta bg)/(prob + delta), delta = 1e-10 -> equals c to < 1 ulp; kept implicit
  }
  o.rgb = rgb;
  o.loss_abs = 0.f;
  o.active = false;
#pragma unroll
  for (int k = 0; k < 8; ++k) o.frame[k] = 0.f;
  if (MODE == 0) return;

  f3 g;  // dL/d rgb
  if (MODE == 1) {
    g = pre ? pre->y : ld3(a.gin + pix * 3);
  } else {
    float mk = pre ? pre->mk : a.mask[pix];
    f3 y = pre ? pre->y : ld3(a.gin + pix * 3);
    f3 d = mk3(rgb.x * mk - y.x * mk, rgb.y * mk - y.y * mk, rgb.z * mk - y.z * mk);
    o.loss_abs = fabsf(d.x) + fabsf(d.y) + fabsf(d.z);
    auto sgn = [](float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); };
    g = mk3(a.gscale * mk * sgn(d.x), a.gscale * mk * sgn(d.y), a.gscale * mk * sgn(d.z));
  }
  float r_light[3] = {0.f, 0.f, 0.f}, r_amb = 0.f, r_lR[3] = {0.f, 0.f, 0.f}, r_lT = 0.f;
  bool active = hit && (g.x != 0.f || g.y != 0.f || g.z != 0.f);
  if (active) {
    f3 g_tex = sh * g;
    float g_sh = g.x * tex.x + g.y * tex.y + g.z * tex.z;
    const float kd = a.fixed_lights ? a.kd : 1.f - amb;
    r_amb = a.fixed_lights ? 0.f : g_sh * (1.f - cosang * vis);
    float g_cos = cosang > 0.f ? g_sh * kd * vis : 0.f;
    float g_vis = g_sh * kd * cosang;
    f3 gP = mk3(0.f, 0.f, 0.f), gN = mk3(0.f, 0.f, 0.f);
    // --- lighting ---
    f3 g_n2 = g_cos * dh, g_dh = g_cos * n2;
    f3 g_n1 = normalize_bwd_eps(n1, g_n2, 1e-6f);
    f3 g_npre = normalize_bwd_eps(npre, g_n1, 1e-12f);
    f3 g_d = normalize_bwd_eps(dvec, g_dh, 1e-6f);
    r_light[0] = g_d.x; r_light[1] = g_d.y; r_light[2] = g_d.z;
    gP = gP - g_d;
    // --- normal map / TBN ---
    f3 g_m = mk3(-dot(uvec, g_npre), -dot(vvec, g_npre), dot(N, g_npre));
    gN = gN + m.z * g_npre;
    f3 g_u = (-m.x) * g_npre, g_v = (-m.y) * g_npre;
    {
      float s = s_sign, x = N.x, y = N.y, a2 = aa * aa;
      // u = (1 + s x^2 a, s x y a, -s x), v = (x y a, s + y^2 a, -y), a = -1/(s+z), da/dz = a^2
      gN.x += g_u.x * (2.f * s * x * aa) + g_u.y * (s * y * aa) + g_u.z * (-s) + g_v.x * (y * aa);
      gN.y += g_u.y * (s * x * aa) + g_v.x * (x * aa) + g_v.y * (2.f * y * aa) + g_v.z * (-1.f);
      gN.z += g_u.x * (s * x * x * a2) + g_u.y * (s * x * y * a2) + g_v.x * (x * y * a2) + g_v.y * (y * y * a2);
    }
    // --- shadow ---
    if (a.self_shadow && g_vis != 0.f) {
      float gz = 0.f;
      float* dzb = a.d_light_zbuf + (size_t)b * S * S;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float ga = g_vis / 9.f * tapsig[t] * (1.f - tapsig[t]) * 1000.f;
        if (ga != 0.f) {
          int ry = min(max(tap_yk + t / 3 - 1, 0), S - 1), rx = min(max(tap_xk + t % 3 - 1, 0), S - 1);
          atomicAdd(dzb + ry * S + rx, ga);
          gz -= ga;
        }
      }
      gP = gP + mk3(gz * lR[2], gz * lR[5], gz * lR[8]);
      r_lR[0] = gz * P.x; r_lR[1] = gz * P.y; r_lR[2] = gz * P.z;
      r_lT = gz;
    }
    // --- textures: value grads to the 4 taps, coordinate grads to uv ---
    float g_u_tex = 0.f, g_v_tex = 0.f;
    {
      f3 t4[4], m4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 t = __ldg(a.tex4 + tp.idx[k]), nm = __ldg(a.nm4 + tp.idx[k]);
        t4[k] = mk3(t.x, t.y, t.z);
        m4[k] = mk3(nm.x, nm.y, nm.z);
        if (tp.w[k] != 0.f) {
          f3 gt = tp.w[k] * g_tex, gm = normalize_bwd_packed(nm, tp.w[k] * g_m);
          atomic_add4(a.d_texture + (size_t)tp.idx[k] * 4, gt.x, gt.y, gt.z);
          atomic_add4(a.d_normal_map + (size_t)tp.idx[k] * 4, gm.x, gm.y, gm.z);
        }
      }
      // d out / d ix = (ne - nw)(1-fy) + (se - sw) fy ;  d out / d iy = (sw - nw)(1-fx) + (se - ne) fx
      f3 dtx = (1.f - tp.fy) * (t4[1] - t4[0]) + tp.fy * (t4[3] - t4[2]);
      f3 dty = (1.f - tp.fx) * (t4[2] - t4[0]) + tp.fx * (t4[3] - t4[1]);
      f3 dmx = (1.f - tp.fy) * (m4[1] - m4[0]) + tp.fy * (m4[3] - m4[2]);
      f3 dmy = (1.f - tp.fx) * (m4[2] - m4[0]) + tp.fx * (m4[3] - m4[1]);
      g_u_tex = tp.gmx * (dot(g_tex, dtx) + dot(g_m, dmx));
      g_v_tex = tp.gmy * (dot(g_tex, dty) + dot(g_m, dmy));
    }
    // --- interpolation: P, N, uv = sum_i b_i attr_i ---
    const float* vb = a.verts + (size_t)b * c.V * 3;
    const float* nb = a.normals + (size_t)b * c.V * 3;
    const float* fu = c.face_uvs + (size_t)h.fid * 6;
    float gb[3] = {dot(ld3(vb + (size_t)v0 * 3), gP) + dot(ld3(nb + (size_t)v0 * 3), gN) + fu[0] * g_u_tex + fu[1] * g_v_tex,
                   dot(ld3(vb + (size_t)v1 * 3), gP) + dot(ld3(nb + (size_t)v1 * 3), gN) + fu[2] * g_u_tex + fu[3] * g_v_tex,
                   dot(ld3(vb + (size_t)v2 * 3), gP) + dot(ld3(nb + (size_t)v2 * 3), gN) + fu[4] * g_u_tex + fu[5] * g_v_tex};
    float* dvb = a.d_verts + (size_t)b * c.V * 4;   // padded (B,V,4) accumulators, folded after the kernel
    float* dnb = a.d_normals + (size_t)b * c.V * 4;
    atomic_add4(dvb + (size_t)v0 * 4, h.b0 * gP.x, h.b0 * gP.y, h.b0 * gP.z);
    atomic_add4(dvb + (size_t)v1 * 4, h.b1 * gP.x, h.b1 * gP.y, h.b1 * gP.z);
    atomic_add4(dvb + (size_t)v2 * 4, h.b2 * gP.x, h.b2 * gP.y, h.b2 * gP.z);
    atomic_add4(dnb + (size_t)v0 * 4, h.b0 * gN.x, h.b0 * gN.y, h.b0 * gN.z);
    atomic_add4(dnb + (size_t)v1 * 4, h.b1 * gN.x, h.b1 * gN.y, h.b1 * gN.z);
    atomic_add4(dnb + (size_t)v2 * 4, h.b2 * gN.x, h.b2 * gN.y, h.b2 * gN.z);
    float gz[3] = {0.f, 0.f, 0.f};
    bary_backward(tri ? *tri : tri_from_ndc(ndc, v0, v1, v2), v0, v1, v2, px, py, h, gb, gz, a.d_ndc_cam + (size_t)b * c.V * 4);
  }
  o.active = active;
  o.frame[0] = r_light[0]; o.frame[1] = r_light[1]; o.frame[2] = r_light[2]; o.frame[3] = r_amb;
  o.frame[4] = r_lR[0]; o.frame[5] = r_lR[1]; o.frame[6] = r_lR[2]; o.frame[7] = r_lT;
}
