// Context: static topology tables (PyTorch3D Meshes/SubdivideMeshes layouts), folded MANO
// constants, and scratch buffers.  Host-side construction runs once per (model, topology).
#include <algorithm>
#include <map>
#include <mutex>

#include "raster.cuh"

static thread_local std::string g_last_error;
void harp_set_error(const std::string& msg) { g_last_error = msg; }
extern "C" const char* harp_last_error(void) { return g_last_error.c_str(); }
extern "C" int harp_version(void) { return 100; }

#include <atomic>
static std::atomic<int64_t> g_launches{0};
void harp_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int64_t harp_launch_count(void) { return g_launches.load(); }

struct harp_prof {
  int level = 0;          // 0 off; 1: the two pixel kernels (ids 0, 1; bench roofline timing); 2: also the step's segments (ids 2..7)
  bool in_graph = false;  // events belong to a captured graph: keep them across reads
  std::vector<cudaEvent_t> ev[12];  // start/stop pairs per id: 0 tile rasteriser, 1 shade, 2 geometry fwd, 3 geometry bwd,
                                    // 4, 5 stand-alone mesh build fwd / bwd, 6 camera project + bins, 7 render backward tail; timeline only:
                                    // 8 step prologue, 9 the render's side branch, 10 the regulariser branch
  std::vector<std::pair<int, cudaEvent_t>> order;  // every event in recording order (harp_ctx_profile_timeline)
};
void harp_prof_begin(harp_ctx* c, int id, cudaStream_t st) {
  if (!c->prof || c->prof->level == 0 || (id >= 2 && c->prof->level < 2)) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return; }
  // inside a stream capture the record must become an EXTERNAL event node: it is re-recorded on every graph replay and
  // stays usable with cudaEventElapsedTime
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  if (cap == cudaStreamCaptureStatusActive) c->prof->in_graph = true;
  cudaError_t rc = cap == cudaStreamCaptureStatusActive ? cudaEventRecordWithFlags(e, st, cudaEventRecordExternal)
                                                        : cudaEventRecord(e, st);
  if (rc != cudaSuccess) { cudaGetLastError(); cudaEventDestroy(e); return; }
  c->prof->ev[id].push_back(e);
  c->prof->order.emplace_back(id, e);
}
void harp_prof_end(harp_ctx* c, int id, cudaStream_t st) { harp_prof_begin(c, id, st); }
extern "C" int harp_ctx_profile(harp_ctx* c, int enable) {
  HARP_REQUIRE(c, HARP_ERR_INVALID, "harp_ctx_profile: null ctx");
  if (!c->prof) c->prof = new harp_prof();
  c->prof->level = enable;
  return HARP_OK;
}
extern "C" int harp_ctx_profile_read(harp_ctx* c, int id, float* total_ms, int* launches) {
  HARP_REQUIRE(c && id >= 0 && id < 8 && total_ms && launches, HARP_ERR_INVALID, "harp_ctx_profile_read: bad argument");
  *total_ms = 0.f; *launches = 0;
  if (!c->prof) return HARP_OK;
  auto& v = c->prof->ev[id];
  HARP_CHECK_CUDA(cudaDeviceSynchronize());
  for (size_t i = 0; i + 1 < v.size(); i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, v[i], v[i + 1]) == cudaSuccess) { *total_ms += ms; ++*launches; }
  }
  if (!c->prof->in_graph) {
    for (cudaEvent_t e : v) cudaEventDestroy(e);
    v.clear();
    auto& o = c->prof->order;
    for (size_t i = 0; i < o.size();) { if (o[i].first == id) o.erase(o.begin() + i); else ++i; }
  }
  return HARP_OK;
}
// Every profiling event in recording order as (id, ms since the first event) pairs; even occurrences of an id open its
// segment, odd ones close it.  For the events of a captured graph these are the times of the last replay.
extern "C" int harp_ctx_profile_timeline(harp_ctx* c, float* out, int capacity, int* n) {
  HARP_REQUIRE(c && out && n && capacity >= 0, HARP_ERR_INVALID, "harp_ctx_profile_timeline: bad argument");
  *n = 0;
  if (!c->prof || c->prof->order.empty()) return HARP_OK;
  HARP_CHECK_CUDA(cudaDeviceSynchronize());
  auto& o = c->prof->order;
  for (size_t i = 0; i < o.size() && *n < capacity / 2; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, o[0].second, o[i].second) != cudaSuccess) { cudaGetLastError(); continue; }
    out[2 * *n] = (float)o[i].first; out[2 * *n + 1] = ms;
    ++*n;
  }
  if (!c->prof->in_graph) {  // eager launches: reading the timeline consumes every event (a captured graph keeps its nodes)
    for (auto& e : o) cudaEventDestroy(e.second);
    o.clear();
    for (auto& v : c->prof->ev) v.clear();
  }
  return HARP_OK;
}

namespace {

struct EdgeTables {
  std::vector<int32_t> edges;  // (E,2) sorted unique (min,max), ascending key min*V+max  == Meshes.edges_packed
  std::vector<int32_t> f2e;    // (F,3) edge opposite v0 | v1 | v2  == faces_packed_to_edges_packed
};

EdgeTables build_edges(const std::vector<int32_t>& faces, int V) {
  int F = (int)faces.size() / 3;
  std::map<int64_t, int32_t> ids;
  auto key = [&](int a, int b) { return (int64_t)std::min(a, b) * V + std::max(a, b); };
  for (int f = 0; f < F; ++f) {
    int a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
    ids[key(b, c)] = 0;
    ids[key(c, a)] = 0;
    ids[key(a, b)] = 0;
  }
  EdgeTables t;
  int32_t n = 0;
  for (auto& kv : ids) {
    kv.second = n++;
    t.edges.push_back((int32_t)(kv.first / V));
    t.edges.push_back((int32_t)(kv.first % V));
  }
  t.f2e.resize(3 * F);
  for (int f = 0; f < F; ++f) {
    int a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
    t.f2e[3 * f + 0] = ids[key(b, c)];
    t.f2e[3 * f + 1] = ids[key(c, a)];
    t.f2e[3 * f + 2] = ids[key(a, b)];
  }
  return t;
}

template <typename T>
int upload(T** dst, const std::vector<T>& src) {
  size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
  HARP_CHECK_CUDA(cudaMalloc((void**)dst, bytes));
  if (!src.empty()) HARP_CHECK_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return HARP_OK;
}

void csr_from_lists(const std::vector<std::vector<int32_t>>& lists, std::vector<int32_t>& off, std::vector<int32_t>& idx) {
  off.assign(lists.size() + 1, 0);
  for (size_t i = 0; i < lists.size(); ++i) off[i + 1] = off[i] + (int32_t)lists[i].size();
  idx.clear();
  for (auto& l : lists) idx.insert(idx.end(), l.begin(), l.end());
}

}  // namespace

extern "C" int harp_ctx_create_lbs(harp_ctx** out, const harp_topology* topo, const harp_lbs_model* model) {
  HARP_REQUIRE(out && topo && topo->faces0 && topo->V0 > 0 && topo->F0 > 0, HARP_ERR_INVALID, "harp_ctx_create: bad topology");
  harp_ctx* c = new harp_ctx();
  c->V0 = topo->V0;
  c->F0 = topo->F0;
  c->subdivide = topo->subdivide != 0;
  std::vector<int32_t> faces0(topo->faces0, topo->faces0 + 3 * topo->F0);
  for (int32_t v : faces0) HARP_REQUIRE(v >= 0 && v < c->V0, HARP_ERR_INVALID, "harp_ctx_create: face index out of range");
  EdgeTables e0 = build_edges(faces0, c->V0);
  c->E0 = (int)e0.edges.size() / 2;
  std::vector<int32_t> faces;
  if (c->subdivide) {  // SubdivideMeshes.subdivide_faces
    c->V = c->V0 + c->E0;
    c->F = 4 * c->F0;
    faces.resize(3 * c->F);
    for (int f = 0; f < c->F0; ++f) {
      int v0 = faces0[3 * f], v1 = faces0[3 * f + 1], v2 = faces0[3 * f + 2];
      int a = e0.f2e[3 * f] + c->V0, b = e0.f2e[3 * f + 1] + c->V0, d = e0.f2e[3 * f + 2] + c->V0;  // e0,e1,e2
      int32_t* f0 = &faces[3 * f];
      int32_t* f1 = &faces[3 * (f + c->F0)];
      int32_t* f2 = &faces[3 * (f + 2 * c->F0)];
      int32_t* f3 = &faces[3 * (f + 3 * c->F0)];
      f0[0] = v0; f0[1] = d; f0[2] = b;
      f1[0] = v1; f1[1] = a; f1[2] = d;
      f2[0] = v2; f2[1] = b; f2[2] = a;
      f3[0] = a;  f3[1] = b; f3[2] = d;
    }
  } else {
    c->V = c->V0;
    c->F = c->F0;
    faces = faces0;
  }
  EdgeTables e = build_edges(faces, c->V);
  c->E = (int)e.edges.size() / 2;
  // vertex -> (face, corner)
  std::vector<std::vector<int32_t>> vf(c->V), vn(c->V), v0e(c->V0);
  for (int f = 0; f < c->F; ++f)
    for (int k = 0; k < 3; ++k) vf[faces[3 * f + k]].push_back(f * 4 + k);
  for (int i = 0; i < c->E; ++i) {
    vn[e.edges[2 * i]].push_back(e.edges[2 * i + 1]);
    vn[e.edges[2 * i + 1]].push_back(e.edges[2 * i]);
  }
  for (int i = 0; i < c->E0; ++i) {
    v0e[e0.edges[2 * i]].push_back(i);
    v0e[e0.edges[2 * i + 1]].push_back(i);
  }
  // normal-consistency pairs: every pair of faces sharing an edge (mesh_normal_consistency.py)
  std::vector<std::vector<int32_t>> opp(c->E);
  for (int f = 0; f < c->F; ++f)
    for (int k = 0; k < 3; ++k) opp[e.f2e[3 * f + k]].push_back(faces[3 * f + k]);
  std::vector<int32_t> pairs;
  for (int i = 0; i < c->E; ++i)
    for (size_t p = 0; p < opp[i].size(); ++p)
      for (size_t q = p + 1; q < opp[i].size(); ++q) {
        pairs.push_back(e.edges[2 * i]);
        pairs.push_back(e.edges[2 * i + 1]);
        pairs.push_back(opp[i][p]);
        pairs.push_back(opp[i][q]);
      }
  c->NP = (int)pairs.size() / 4;
  std::vector<int32_t> off, idx;
  int rc;
#define UP(dst, vec) if ((rc = upload(&c->dst, vec)) != HARP_OK) { harp_ctx_destroy(c); return rc; }
  UP(faces, faces);
  UP(edges0, e0.edges);
  UP(edges, e.edges);
  UP(nc_pairs, pairs);
  csr_from_lists(vf, off, idx);  UP(vf_off, off);  UP(vf_idx, idx);
  {
    HARP_REQUIRE(c->V < 65536, HARP_ERR_INVALID, "harp_ctx_create: meshes up to 65535 vertices");
    std::vector<uint32_t> ring(idx.size());
    for (size_t i = 0; i < idx.size(); ++i) {
      int f = idx[i] >> 2, k = idx[i] & 3;
      ring[i] = (uint32_t)faces[3 * f + (k + 1) % 3] | ((uint32_t)faces[3 * f + (k + 2) % 3] << 16);
    }
    UP(vf_ring, ring);
    int max_valence = 0;
    for (int v = 0; v < c->V; ++v) max_valence = std::max(max_valence, off[v + 1] - off[v]);
    if (max_valence <= 8) {  // the MANO / SMPL-X-arm meshes and their subdivisions: at most 8
      std::vector<uint32_t> ring8((size_t)c->V * 8, 0xffffffffu);
      for (int v = 0; v < c->V; ++v)
        for (int i = off[v]; i < off[v + 1]; ++i) ring8[(size_t)v * 8 + (i - off[v])] = ring[i];
      uint32_t* dev = nullptr;
      if ((rc = upload(&dev, ring8)) != HARP_OK) { harp_ctx_destroy(c); return rc; }
      c->vf_ring8 = reinterpret_cast<uint4*>(dev);
    }
  }
  csr_from_lists(vn, off, idx);  UP(vn_off, off);  UP(vn_idx, idx);
  csr_from_lists(v0e, off, idx); UP(v0e_off, off); UP(v0e_idx, idx);
  if (topo->faces_uvs && topo->verts_uvs && topo->Vt > 0) {
    c->Vt = topo->Vt;
    std::vector<float> fuv(6 * (size_t)c->F);
    for (int f = 0; f < c->F; ++f)
      for (int k = 0; k < 3; ++k) {
        int t = topo->faces_uvs[3 * f + k];
        if (t < 0 || t >= c->Vt) { harp_ctx_destroy(c); harp_set_error("harp_ctx_create: faces_uvs out of range"); return HARP_ERR_INVALID; }
        fuv[6 * f + 2 * k] = topo->verts_uvs[2 * t];
        fuv[6 * f + 2 * k + 1] = topo->verts_uvs[2 * t + 1];
      }
    UP(face_uvs, fuv);
    c->has_uv = true;
  }
  if (model) {
    int V0 = c->V0, J = model->J, NS = model->n_shape, NP = (J - 1) * 9;
    bool ok = J >= 1 && J <= HARP_MAX_JOINTS && NS >= 0 && NS <= HARP_MAX_SHAPE && model->parents && model->shapedirs &&
              model->posedirs && model->v_template && model->J_template && model->J_shapedirs && model->weights &&
              model->pose_mean && model->n_joints_out >= 0 && model->n_joints_out <= 512 && (model->n_tips == 0 || model->tips);
    for (int j = 1; ok && j < J; ++j) ok = model->parents[j] >= 0 && model->parents[j] < j;
    for (int i = 0; ok && i < model->n_tips; ++i) ok = model->tips[i] >= 0 && model->tips[i] < V0;
    for (int i = 0; ok && i < model->n_joints_out; ++i) ok = model->joint_out_src[i] >= 0 && model->joint_out_src[i] < J + model->n_tips;
    if (!ok) { harp_ctx_destroy(c); harp_set_error("harp_ctx_create: invalid skinning model"); return HARP_ERR_INVALID; }
    c->J = J; c->n_shape = NS; c->rodrigues_kind = model->rodrigues_kind; c->center_joint = model->center_joint;
    c->n_joints_out = model->n_joints_out; c->n_tips = model->n_tips;
    std::vector<float> vt(model->v_template, model->v_template + (size_t)V0 * 3);
    std::vector<float> w(model->weights, model->weights + (size_t)V0 * J);
    std::vector<float> pm(model->pose_mean, model->pose_mean + (size_t)J * 3);
    std::vector<float> Jt(model->J_template, model->J_template + (size_t)J * 3);
    std::vector<float> Js(model->J_shapedirs, model->J_shapedirs + (size_t)J * 3 * NS);
    std::vector<int32_t> tips(model->tips, model->tips + model->n_tips), par(model->parents, model->parents + J);
    std::vector<int32_t> jsrc(model->joint_out_src, model->joint_out_src + model->n_joints_out);
    // blendshape tables are stored transposed, (k, V0*3), so that a thread-per-vertex loop reads coalesced
    std::vector<float> sdT((size_t)V0 * 3 * NS), pdT((size_t)V0 * 3 * NP), wT((size_t)V0 * J);
    for (int i = 0; i < V0 * 3; ++i) {
      for (int b = 0; b < NS; ++b) sdT[(size_t)b * V0 * 3 + i] = model->shapedirs[(size_t)i * NS + b];
      for (int k = 0; k < NP; ++k) pdT[(size_t)k * V0 * 3 + i] = model->posedirs[(size_t)i * NP + k];
    }
    for (int v = 0; v < V0; ++v)
      for (int j = 0; j < J; ++j) wT[(size_t)j * V0 + v] = w[(size_t)v * J + j];
    UP(weightsT, wT); UP(shapedirs, sdT); UP(posedirs, pdT); UP(v_template, vt); UP(weights, w); UP(pose_mean, pm);
    UP(tips, tips); UP(parents, par); UP(joint_out_src, jsrc); UP(J_template, Jt); UP(J_shapedirs, Js);
    c->has_model = true;
  }
#undef UP
  if (cudaMalloc((void**)&c->counters, 2 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMemset(c->counters, 0, 2 * sizeof(unsigned long long)) != cudaSuccess) {
    harp_ctx_destroy(c);
    harp_set_error("harp_ctx_create: counters alloc failed");
    return HARP_ERR_CUDA;
  }
  *out = c;
  return HARP_OK;
}

static void free_scratch(harp_ctx* c) {
  void* ptrs[] = {c->ndc_cam, c->ndc_light, c->d_ndc_cam, c->d_ndc_light, c->bin_count, c->bin_start, c->bin_fill,
                  c->bin_list, c->light_zbuf, c->light_face, c->d_light_zbuf, c->vN, c->vtmp, c->vtmp2, c->d_verts4,
                  c->d_normals4, c->hit_list, c->hit_count, c->ovf_list, c->ovf_count, c->face_rec, c->work, c->work_ctl};
  c->ovf_list = c->ovf_count = nullptr;
  c->face_rec = nullptr; c->work = nullptr; c->work_ctl = nullptr;
  c->d_verts4 = c->d_normals4 = nullptr;
  c->hit_list = c->hit_count = nullptr;
  for (void* p : ptrs) if (p) cudaFree(p);
  c->ndc_cam = c->ndc_light = c->d_ndc_cam = c->d_ndc_light = c->light_zbuf = c->d_light_zbuf = c->vN = c->vtmp = c->vtmp2 = nullptr;
  c->bin_count = c->bin_start = c->bin_fill = c->bin_list = c->light_face = nullptr;
  c->maxB = 0;
}

extern "C" void harp_ctx_destroy(harp_ctx* c) {
  if (!c) return;
  free_scratch(c);
  void* ptrs[] = {c->faces, c->edges0, c->v0e_off, c->v0e_idx, c->vf_off, c->vf_idx, c->vn_off, c->vn_idx, c->edges,
                  c->nc_pairs, c->face_uvs, c->vf_ring, c->vf_ring8, c->shapedirs, c->posedirs, c->v_template, c->weights,
                  c->J_template, c->J_shapedirs, c->tips, c->counters, c->d_tex4, c->tex4, c->weightsT, c->pose_mean, c->parents,
                  c->joint_out_src};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_setup) cudaEventDestroy(c->ev_setup);
  if (c->ev_reg_fork) cudaEventDestroy(c->ev_reg_fork);
  if (c->ev_reg_join) cudaEventDestroy(c->ev_reg_join);
  if (c->reg_stream) cudaStreamDestroy(c->reg_stream);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  if (c->prof) {
    for (auto& v : c->prof->ev) for (cudaEvent_t e : v) cudaEventDestroy(e);
    delete c->prof;
  }
  delete c;
}

extern "C" int harp_ctx_reserve(harp_ctx* c, int max_batch, int img_size) {
  HARP_REQUIRE(c && max_batch > 0 && img_size > 0 && img_size % 8 == 0, HARP_ERR_INVALID,
               "harp_ctx_reserve: need max_batch>0 and img_size a positive multiple of 8");
  if (max_batch <= c->maxB && img_size == c->S) return HARP_OK;
  // Kernels take the ctx by value, so a captured CUDA graph bakes the scratch pointers in: while any graph is alive
  // (harp_ctx_pin) the scratch must not move.  The batch capacity never shrinks.
  HARP_REQUIRE(c->pins == 0, HARP_ERR_CAPACITY,
               "harp_ctx_reserve: the scratch is pinned by a captured CUDA graph (harp_ctx_pin); release it or use another context");
  int B = std::max(max_batch, c->maxB), S = img_size;
  free_scratch(c);
  int tiles = (S / 8) * (S / 8);
  size_t nv = (size_t)B * c->V * 3 * sizeof(float);
  size_t npx = (size_t)B * S * S;
  // a face overlaps ~1-8 8x8 tiles at the reference's scales (bbox pairs measured on the synthetic hand: 0.2-0.35 x 24 F at
  // 512^2); the capacity per face grows with the tile count so that a hand zoomed far beyond the image, or crossing the
  // camera plane (huge projected faces: 3 x 24 F at 512^2, 13 x at 1024^2), still fits: 24 F up to 256^2, 64 F at 512^2,
  // 256 F from 1024^2 (0.2 / 0.5 GB for both views at 64 x 512^2 / 32 x 1024^2).  Overflow is counted (counters[1]), never silent.
  int per_face = std::min(256, std::max(24, tiles / 64));
  c->bin_cap = (int64_t)c->F * per_face;
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->ndc_cam, nv));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->ndc_light, nv));
  size_t nv4 = (size_t)B * c->V * 4 * sizeof(float);
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->d_ndc_cam, nv4));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->d_ndc_light, nv4));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->d_verts4, nv4));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->d_normals4, nv4));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->vN, nv));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->vtmp, nv));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->vtmp2, nv));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->bin_count, 2 * (size_t)B * tiles * sizeof(int32_t)));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->bin_start, 2 * (size_t)B * tiles * sizeof(int32_t)));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->bin_fill, 2 * (size_t)B * tiles * sizeof(int32_t)));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->bin_list, 2 * (size_t)B * c->bin_cap * sizeof(int32_t)));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->light_zbuf, npx * sizeof(float)));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->light_face, npx * sizeof(int32_t)));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->d_light_zbuf, npx * sizeof(float)));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->hit_list, 2 * npx * sizeof(int32_t)));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->hit_count, (size_t)B * sizeof(int32_t)));
  c->ovf_cap = (int)std::min<size_t>(npx, (size_t)1 << 22);
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->ovf_list, (size_t)c->ovf_cap * sizeof(int32_t)));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->ovf_count, sizeof(int32_t)));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->face_rec, 2 * (size_t)B * c->F * sizeof(FaceTile)));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->work, 2 * (size_t)B * tiles * sizeof(int4)));
  HARP_CHECK_CUDA(cudaMalloc((void**)&c->work_ctl, 8 * sizeof(int32_t)));
  HARP_CHECK_CUDA(cudaMemset(c->work_ctl, 0, 8 * sizeof(int32_t)));
  c->maxB = B;
  c->S = S;
  c->tiles = tiles;
  return HARP_OK;
}

extern "C" int harp_ctx_pin(harp_ctx* c, int delta) {
  HARP_REQUIRE(c, HARP_ERR_INVALID, "harp_ctx_pin: null ctx");
  c->pins = std::max(0, c->pins + delta);
  return HARP_OK;
}

extern "C" int harp_ctx_dims(const harp_ctx* c, int32_t* V0, int32_t* V, int32_t* F, int32_t* E) {
  HARP_REQUIRE(c, HARP_ERR_INVALID, "harp_ctx_dims: null ctx");
  if (V0) *V0 = c->V0;
  if (V) *V = c->V;
  if (F) *F = c->F;
  if (E) *E = c->E;
  return HARP_OK;
}

extern "C" int harp_ctx_faces(const harp_ctx* c, int32_t* out) {
  HARP_REQUIRE(c && out, HARP_ERR_INVALID, "harp_ctx_faces: null");
  HARP_CHECK_CUDA(cudaMemcpy(out, c->faces, (size_t)c->F * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return HARP_OK;
}

extern "C" int harp_ctx_read_counters(harp_ctx* c, int64_t out[2], void* stream) {
  HARP_REQUIRE(c && out, HARP_ERR_INVALID, "harp_ctx_read_counters: null");
  unsigned long long h[2];
  cudaStream_t st = (cudaStream_t)stream;
  HARP_CHECK_CUDA(cudaMemcpyAsync(h, c->counters, sizeof(h), cudaMemcpyDeviceToHost, st));
  HARP_CHECK_CUDA(cudaMemsetAsync(c->counters, 0, sizeof(h), st));
  HARP_CHECK_CUDA(cudaStreamSynchronize(st));
  out[0] = (int64_t)h[0];
  out[1] = (int64_t)h[1];
  return HARP_OK;
}

// Pixels the LAST silhouette pass on this ctx handed to the exact per-pixel top-K kernel (more candidate faces than
// faces_per_pixel in a tile with more than K faces, or a tile whose band queue overflowed).  Results are exact either way; the
// number says how far the content is from the fast path (the fallback costs ~0.1 us per pixel).
extern "C" int harp_ctx_deferred_pixels(harp_ctx* c, int64_t* out, void* stream) {
  HARP_REQUIRE(c && out, HARP_ERR_INVALID, "harp_ctx_deferred_pixels: null");
  int32_t h = 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (c->ovf_count) {
    HARP_CHECK_CUDA(cudaMemcpyAsync(&h, c->ovf_count, sizeof(h), cudaMemcpyDeviceToHost, st));
    HARP_CHECK_CUDA(cudaStreamSynchronize(st));
  }
  *out = h;
  return HARP_OK;
}

// MANO (manopth/manolayer.py:70-97) as an instance of the generic model: fold the joint regressor,
// J(beta) = Jreg v_template + (Jreg shapedirs) beta (manolayer.py:185-188), kinematic tree of manolayer.py:209-242,
// 21 output joints = 16 chain joints + 5 fingertip vertices in the order of manolayer.py:279.
extern "C" int harp_ctx_create(harp_ctx** out, const harp_topology* topo, const harp_mano_model* m) {
  if (!m) return harp_ctx_create_lbs(out, topo, nullptr);
  HARP_REQUIRE(out && topo && topo->V0 > 0, HARP_ERR_INVALID, "harp_ctx_create: bad topology");
  HARP_REQUIRE(m->shapedirs && m->posedirs && m->v_template && m->J_regressor && m->weights && m->hands_mean && m->tips,
               HARP_ERR_INVALID, "harp_ctx_create: incomplete MANO model");
  int V0 = topo->V0;
  std::vector<double> Jt(48, 0.0), Js(480, 0.0);
  for (int j = 0; j < 16; ++j)
    for (int v = 0; v < V0; ++v) {
      double r = m->J_regressor[(size_t)j * V0 + v];
      if (r == 0.0) continue;
      for (int k = 0; k < 3; ++k) {
        Jt[j * 3 + k] += r * m->v_template[v * 3 + k];
        for (int b = 0; b < 10; ++b) Js[(j * 3 + k) * 10 + b] += r * m->shapedirs[((size_t)v * 3 + k) * 10 + b];
      }
    }
  std::vector<float> Jtf(Jt.begin(), Jt.end()), Jsf(Js.begin(), Js.end()), pm(48, 0.f);
  for (int i = 0; i < 45; ++i) pm[3 + i] = m->hands_mean[i];
  const int32_t parents[16] = {-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14};
  const int32_t jsrc[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};
  harp_lbs_model g{};
  g.J = 16; g.n_shape = 10; g.parents = parents; g.shapedirs = m->shapedirs; g.posedirs = m->posedirs;
  g.v_template = m->v_template; g.J_template = Jtf.data(); g.J_shapedirs = Jsf.data(); g.weights = m->weights;
  g.pose_mean = pm.data(); g.rodrigues_kind = 0; g.center_joint = -1; g.n_joints_out = 21; g.joint_out_src = jsrc;
  g.n_tips = 5; g.tips = m->tips;
  return harp_ctx_create_lbs(out, topo, &g);
}
