// Device-side pieces of the mesh build (PyTorch3D SubdivideMeshes + displacement along vertex normals + vertex normals and
// their backward: utils/visualize.py:45-64), shared by mesh.cu and the fused geometry kernels (geom.cu).
// The vertex accessors take any address space: the fused kernels keep a frame's vertices in shared memory.
#pragma once
#include "common.cuh"

namespace {

// Vertex accessors ------------------------------------------------------------------------------------------------
struct SubdivAccessor {  // subdivided vertex v computed on the fly from the coarse verts of one frame
  const float* v0;
  const int32_t* edges0;
  int V0;
  __device__ __forceinline__ f3 operator()(int v) const {
    if (v < V0) return ld3(v0 + (size_t)v * 3);
    int e = v - V0;
    f3 a = ld3(v0 + (size_t)edges0[2 * e] * 3), b = ld3(v0 + (size_t)edges0[2 * e + 1] * 3);
    return mk3((a.x + b.x) / 2.f, (a.y + b.y) / 2.f, (a.z + b.z) / 2.f);  // mean(dim=2) of two
  }
};
struct DirectAccessor {
  const float* v;
  __device__ __forceinline__ f3 operator()(int i) const { return ld3(v + (size_t)i * 3); }
};

// Meshes._compute_vertex_normals: sum over incident faces of the corner-specific cross product
// cross(next - self, prev - self) (PyTorch3D adds cross(v1 - v0, v2 - v0) at corner 0, cross(v2 - v1, v0 - v1) at corner 1, ...).
// The ring table gives (next, prev) per incidence in one coalesced load; incidences are fetched eight at a time before any of
// them is used, so the loads are in flight together (the gather is bound by memory latency, not by arithmetic).
template <typename Acc>
__device__ __forceinline__ f3 vertex_normal_sum(const harp_ctx& c, const Acc& acc, int v) {
  f3 n = mk3(0.f, 0.f, 0.f);
  const f3 a = acc(v);
  if (c.vf_ring8) {  // fixed-stride table: one aligned 32-byte load, entries in the same order as the CSR walk below
    const uint4 r0 = __ldg(c.vf_ring8 + 2 * (size_t)v), r1 = __ldg(c.vf_ring8 + 2 * (size_t)v + 1);
    const uint32_t ring[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (ring[k] == 0xffffffffu) continue;
      f3 b = acc((int)(ring[k] & 0xffffu)), d = acc((int)(ring[k] >> 16));
      n = n + cross(b - a, d - a);
    }
    return n;
  }
  const int i0 = c.vf_off[v], i1 = c.vf_off[v + 1];
  for (int base = i0; base < i1; base += 8) {
    uint32_t ring[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) ring[k] = base + k < i1 ? c.vf_ring[base + k] : 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (ring[k] == 0xffffffffu) continue;
      f3 b = acc((int)(ring[k] & 0xffffu)), d = acc((int)(ring[k] >> 16));
      n = n + cross(b - a, d - a);
    }
  }
  return n;
}

__device__ __forceinline__ f3 normalize_eps(f3 n, float eps, float* len_out = nullptr) {
  float l = sqrtf(dot(n, n));
  float d = fmaxf(l, eps);
  if (len_out) *len_out = l;
  return mk3(n.x / d, n.y / d, n.z / d);
}

// backward of y = x / max(|x|, eps)
__device__ __forceinline__ f3 normalize_bwd(f3 x, f3 gy, float eps) {
  float l = sqrtf(dot(x, x));
  if (l > eps) {
    f3 y = (1.f / l) * x;
    return (1.f / l) * (gy - dot(y, gy) * y);
  }
  return (1.f / eps) * gy;
}

// gradient of sum_f (c_f . G_f) w.r.t. vertex u, c_f = (next - u) x (prev - u) (the same vector at all three corners of f),
// G_f = dN summed over the three corners of f:  d/du = (next - prev) x G
template <typename Acc>
__device__ __forceinline__ f3 normal_grad_gather(const harp_ctx& c, const Acc& acc, const float* dN, int u) {
  f3 g = mk3(0.f, 0.f, 0.f);
  const f3 Gu = ld3(dN + (size_t)u * 3);
  if (c.vf_ring8) {
    const uint4 r0 = __ldg(c.vf_ring8 + 2 * (size_t)u), r1 = __ldg(c.vf_ring8 + 2 * (size_t)u + 1);
    const uint32_t ring[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (ring[k] == 0xffffffffu) continue;
      int nx = (int)(ring[k] & 0xffffu), pv = (int)(ring[k] >> 16);
      f3 G = Gu + ld3(dN + (size_t)nx * 3) + ld3(dN + (size_t)pv * 3);
      g = g + cross(acc(nx) - acc(pv), G);
    }
    return g;
  }
  const int i0 = c.vf_off[u], i1 = c.vf_off[u + 1];
  for (int base = i0; base < i1; base += 8) {
    uint32_t ring[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) ring[k] = base + k < i1 ? c.vf_ring[base + k] : 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (ring[k] == 0xffffffffu) continue;
      int nx = (int)(ring[k] & 0xffffu), pv = (int)(ring[k] >> 16);
      f3 G = Gu + ld3(dN + (size_t)nx * 3) + ld3(dN + (size_t)pv * 3);
      g = g + cross(acc(nx) - acc(pv), G);
    }
  }
  return g;
}

}  // namespace
