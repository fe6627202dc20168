// Data-parallel exchange of the shared-parameter gradients as ONE kernel over NVLink peer memory.
//
// The reference is single-process (optimize_sequence.py has no distributed code); north_star shards the frames of a
// sequence over the GPUs of one box and sums the shared-parameter gradients once per optimiser step.  NCCL does that in a
// separate eager launch sequence after the step's CUDA graph.  Here every rank's gradient buffer lives in an IPC-shareable
// allocation that all ranks of the node map; the sum is a reduce-scatter + all-gather written directly against the peers'
// buffers, synchronised by epoch flags in the same allocations, so the exchange is an ordinary kernel node INSIDE the
// captured step graph (followed by Adam in the same graph): no host round trip, no staging copy.
//
//   1. arrive: rank r publishes "gradients of step k are final" to every peer (release store, system scope)
//   2. rank r owns slice r of the shared span: it waits for all arrivals, reads slice r from every rank (16-byte loads over
//      NVLink), sums in rank order (so every rank computes bit-identical totals), and writes the total into slice r of EVERY
//      rank's buffer — slices are disjoint, so the in-place writes never touch data another rank still reads
//   3. done: the last CTA of rank r publishes "slice r is written everywhere"; every CTA waits for all ranks' done flags
//      before exiting, so the kernels that follow in the stream (Adam) see complete sums.
#include "common.cuh"

namespace {

constexpr int PEER_MAX = 16;
constexpr int PEER_THREADS = 256;
constexpr int PEER_CTAS = 48;  // far below one wave: every CTA of the kernel is resident, the waits cannot starve anyone

struct PeerTable {
  float* data[PEER_MAX];           // every rank's gradient buffer (element 0 of the flat buffer), indexed by rank
  unsigned int* flags[PEER_MAX];   // every rank's flag block: [0..R) arrive epochs, [R..2R) done epochs, [2R] local CTA counter
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// `epoch_ptr` is a device counter of this rank incremented by the kernel itself (so a replayed graph needs no new argument).
__global__ void __launch_bounds__(PEER_THREADS) peer_allreduce_kernel(PeerTable t, int R, int me, long long start, long long stop,
                                                                       long long scale_lo, long long scale_hi, float scale,
                                                                       unsigned int* epoch_ptr) {
  __shared__ unsigned int s_epoch;
  unsigned int* my = t.flags[me];
  if (threadIdx.x == 0) s_epoch = *epoch_ptr + 1u;  // every CTA reads the same value: it is bumped only by the last CTA, below
  __syncthreads();
  const unsigned int epoch = s_epoch;
  // ---- 1. arrive (one thread of CTA 0 tells every rank, itself included)
  if (blockIdx.x == 0 && threadIdx.x < R) {
    __threadfence_system();
    st_release_sys(t.flags[threadIdx.x] + me, epoch);
  }
  // ---- 2. wait for everyone's gradients, then reduce my slice and broadcast it
  if (threadIdx.x < R) {
    while (ld_acquire_sys(my + threadIdx.x) < epoch) __nanosleep(64);
  }
  __syncthreads();
  // the span [start, stop) in float4 units from the first 16-byte boundary; the unaligned head / tail go to rank 0 as scalars
  const long long a0 = (start + 3) & ~3LL, a1 = stop & ~3LL;
  const long long nvec = a1 > a0 ? (a1 - a0) / 4 : 0;
  const long long per = (nvec + R - 1) / R;
  const long long v0 = min(nvec, per * me), v1 = min(nvec, v0 + per);
  for (long long i = v0 + (long long)blockIdx.x * PEER_THREADS + threadIdx.x; i < v1; i += (long long)gridDim.x * PEER_THREADS) {
    const long long e = a0 + 4 * i;
    float4 s = ld_sys4(t.data[0] + e);
    for (int r = 1; r < R; ++r) {
      float4 v = ld_sys4(t.data[r] + e);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float k[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (e + j >= scale_lo && e + j < scale_hi) k[j] *= scale;
    s = make_float4(k[0], k[1], k[2], k[3]);
    for (int r = 0; r < R; ++r) *reinterpret_cast<float4*>(t.data[r] + e) = s;
  }
  if (me == 0 && blockIdx.x == 0) {  // scalar head and tail
    const long long nh = min(a0, stop) - start, nt = a1 > a0 ? stop - a1 : 0;
    for (long long i = threadIdx.x; i < nh + nt; i += PEER_THREADS) {
      const long long e = i < nh ? start + i : a1 + (i - nh);
      float s = 0.f;
      for (int r = 0; r < R; ++r) s += ld_sys1(t.data[r] + e);
      if (e >= scale_lo && e < scale_hi) s *= scale;
      for (int r = 0; r < R; ++r) t.data[r][e] = s;
    }
  }
  // ---- 3. done: the last CTA of this rank to finish publishes the slice; everybody waits for all slices
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int n = atomicAdd(my + 2 * R, 1u);
    if (n == gridDim.x - 1) {
      my[2 * R] = 0u;            // re-armed for the next step (nobody reads it before the next kernel)
      *epoch_ptr = epoch;
      __threadfence_system();
      for (int r = 0; r < R; ++r) st_release_sys(t.flags[r] + R + me, epoch);
    }
  }
  if (threadIdx.x < R) {
    while (ld_acquire_sys(my + R + threadIdx.x) < epoch) __nanosleep(64);
  }
  __syncthreads();
}

}  // namespace

struct harp_peer {
  int R = 0, me = 0;
  void* base = nullptr;                 // this rank's allocation: [data bytes | flag block]
  size_t data_bytes = 0;
  void* peer_base[PEER_MAX] = {nullptr};
  PeerTable table{};
  unsigned int* epoch = nullptr;        // device counter (local)
};

static size_t flag_bytes() { return 256; }

extern "C" int harp_peer_create(harp_peer** out, int world, int rank, size_t data_bytes, void** data_ptr, void* ipc_handle_out /* 64 B */) {
  HARP_REQUIRE(out && world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world && data_ptr && ipc_handle_out, HARP_ERR_INVALID,
               "harp_peer_create: bad argument (up to 16 ranks of one node)");
  harp_peer* p = new harp_peer();
  p->R = world; p->me = rank;
  p->data_bytes = (data_bytes + 255) & ~(size_t)255;
  HARP_CHECK_CUDA(cudaMalloc(&p->base, p->data_bytes + flag_bytes()));
  HARP_CHECK_CUDA(cudaMemset(p->base, 0, p->data_bytes + flag_bytes()));
  HARP_CHECK_CUDA(cudaMalloc((void**)&p->epoch, sizeof(unsigned int)));
  HARP_CHECK_CUDA(cudaMemset(p->epoch, 0, sizeof(unsigned int)));
  cudaIpcMemHandle_t h;
  HARP_CHECK_CUDA(cudaIpcGetMemHandle(&h, p->base));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(ipc_handle_out, &h, 64);
  *data_ptr = p->base;
  *out = p;
  return HARP_OK;
}

// handles: world x 64 bytes, rank-major (every rank's handle as produced by harp_peer_create, gathered by the host layer)
extern "C" int harp_peer_connect(harp_peer* p, const void* handles) {
  HARP_REQUIRE(p && handles, HARP_ERR_INVALID, "harp_peer_connect: null argument");
  for (int r = 0; r < p->R; ++r) {
    void* b = p->base;
    if (r != p->me) {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)handles + 64 * r, 64);
      HARP_CHECK_CUDA(cudaIpcOpenMemHandle(&b, h, cudaIpcMemLazyEnablePeerAccess));
      p->peer_base[r] = b;
    }
    p->table.data[r] = reinterpret_cast<float*>(b);
    p->table.flags[r] = reinterpret_cast<unsigned int*>((char*)b + p->data_bytes);
  }
  return HARP_OK;
}

// In-place sum over ranks of data[start:stop) (float elements of the buffer harp_peer_create returned); elements in
// [scale_lo, scale_hi) are multiplied by `scale` after the sum (the loss scalars are means over ranks).  Asynchronous on
// `stream`, capturable; every rank must launch it once per step.
extern "C" int harp_peer_allreduce(harp_peer* p, int64_t start, int64_t stop, int64_t scale_lo, int64_t scale_hi, float scale,
                                   void* stream) {
  HARP_REQUIRE(p && start >= 0 && stop >= start && (size_t)stop * 4 <= p->data_bytes, HARP_ERR_INVALID, "harp_peer_allreduce: bad span");
  HARP_REQUIRE(p->table.data[p->R - 1] != nullptr, HARP_ERR_INVALID, "harp_peer_allreduce: call harp_peer_connect first");
  peer_allreduce_kernel<<<PEER_CTAS, PEER_THREADS, 0, (cudaStream_t)stream>>>(p->table, p->R, p->me, start, stop, scale_lo, scale_hi,
                                                                             scale, p->epoch);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

extern "C" void harp_peer_destroy(harp_peer* p) {
  if (!p) return;
  cudaDeviceSynchronize();
  for (int r = 0; r < p->R; ++r)
    if (p->peer_base[r]) cudaIpcCloseMemHandle(p->peer_base[r]);
  if (p->base) cudaFree(p->base);
  if (p->epoch) cudaFree(p->epoch);
  delete p;
}
