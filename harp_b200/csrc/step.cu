// Per-step (not per-frame) pieces: texture regularisers and fused Adam.
// Replaces loss/texture_reg.py:5-66 and torch.optim.Adam.step on the CPU-resident texture / normal map
// (optimize_sequence.py:253-310,567-573).
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int TPB = 256;

// smooth_texture_reg / albedo_reg: mean_{ij} mask_ij * ||t_ij - t_{clamp(ij + off_ij)}||_1 / 3
__global__ void __launch_bounds__(TPB) tex_smooth_kernel(const float* __restrict__ map, int H, int W,
                                                          const int32_t* __restrict__ off, const float* __restrict__ mask,
                                                          float w, float* __restrict__ loss, float* __restrict__ d_map) {
  __shared__ float red[32];
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  float val = 0.f;
  if (idx < H * W) {
    int i = idx / W, j = idx % W;
    int ti = min(max(i + off[2 * idx], 0), H - 1), tj = min(max(j + off[2 * idx + 1], 0), W - 1);
    int tidx = ti * W + tj;
    float m = mask ? mask[idx] : 1.f;
    float inv = 1.f / ((float)H * (float)W);
    float s = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float d = map[idx * 3 + ch] - map[tidx * 3 + ch];
      s += fabsf(d);
      if (d_map && d != 0.f && m != 0.f && tidx != idx) {
        float g = w * inv * m / 3.f * (d > 0.f ? 1.f : -1.f);
        atomicAdd(d_map + idx * 3 + ch, g);
        atomicAdd(d_map + tidx * 3 + ch, -g);
      }
    }
    val = s / 3.f * m * inv;
  }
  float tot = block_sum(val, red);
  if (threadIdx.x == 0 && tot != 0.f) atomicAdd(loss, tot);
}

// close_to_z_reg (loss/texture_reg.py:40-45): torch.norm(normal_map - [0,0,1], p=2, dim=2) on the 4-D (1,H,W,3) tensor,
// i.e. the L2 norm runs along the W axis per (row, channel) — reproduced as written.  One CTA per row.
__global__ void __launch_bounds__(TPB) close_to_z_kernel(const float* __restrict__ map, int H, int W, float w,
                                                          float* __restrict__ loss, float* __restrict__ d_map) {
  __shared__ float red[32];
  __shared__ float nrm[3];
  int i = blockIdx.x;
  const float* row = map + (size_t)i * W * 3;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float d = row[j * 3 + ch] - (ch == 2 ? 1.f : 0.f);
      acc[ch] += d * d;
    }
  }
  for (int ch = 0; ch < 3; ++ch) {
    float s = block_sum(acc[ch], red);
    if (threadIdx.x == 0) nrm[ch] = sqrtf(s);
  }
  __syncthreads();
  float inv = 1.f / (3.f * (float)H * 3.f);  // (/3.0).mean() over (1,H,3)
  if (threadIdx.x == 0) atomicAdd(loss, 0.2f * (nrm[0] + nrm[1] + nrm[2]) * inv);
  if (d_map) {
    float* drow = d_map + (size_t)i * W * 3;
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        if (nrm[ch] > 0.f) {
          float d = row[j * 3 + ch] - (ch == 2 ? 1.f : 0.f);
          atomicAdd(drow + j * 3 + ch, w * 0.2f * inv * d / nrm[ch]);
        }
      }
    }
  }
}

// torch.optim.Adam (amsgrad=False, weight_decay=0): one thread per element
__global__ void __launch_bounds__(TPB) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                    float bc1, float bc2_sqrt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  float mi = m[i] + (1.f - b1) * (gi - m[i]);  // lerp, as torch._single_tensor_adam
  float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

// load_img (utils/data_util.py:11-30): /255 in double then float32 like numpy -> torch.Tensor; two 3x3 erosions = one
// 5x5 minimum with out-of-image taps ignored (cv2.erode's default border value is +inf)
__global__ void __launch_bounds__(TPB) ingest_u8_kernel(int B, int S, const uint8_t* __restrict__ rgb, const uint8_t* __restrict__ mask,
                                                         float* __restrict__ y_true, float* __restrict__ y_sil,
                                                         float* __restrict__ y_er) {
  size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (size_t)B * S * S) return;
  int rem = (int)(pix % ((size_t)S * S)), row = rem / S, col = rem % S;
  const uint8_t* mb = mask + (pix - rem);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) y_true[pix * 3 + ch] = (float)((double)rgb[pix * 3 + ch] / 255.0);
  y_sil[pix] = (float)((double)mask[pix] / 255.0);
  int mn = 255;
  for (int dy = -2; dy <= 2; ++dy) {
    int r = row + dy;
    if (r < 0 || r >= S) continue;
    for (int dx = -2; dx <= 2; ++dx) {
      int cc = col + dx;
      if (cc < 0 || cc >= S) continue;
      mn = min(mn, (int)mb[r * S + cc]);
    }
  }
  y_er[pix] = (float)((double)mn / 255.0);
}

// torch.nn.MSELoss()(scale * x, target) and its gradient w.r.t. x, in one pass (metro_modifications/hand_utils.py:64,74,93)
__global__ void __launch_bounds__(TPB) mse_kernel(int64_t n, const float* __restrict__ x, const float* __restrict__ target,
                                                   float scale, float* __restrict__ loss, float* __restrict__ d_x) {
  __shared__ float red[32];
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float val = 0.f;
  if (i < n) {
    float d = scale * x[i] - target[i];
    val = d * d / (float)n;
    if (d_x) d_x[i] = 2.f * scale * d / (float)n;
  }
  float tot = block_sum(val, red);
  if (threadIdx.x == 0 && tot != 0.f) atomicAdd(loss, tot);
}

// Evaluation metrics of utils/eval_util.py:34-49 in one pass: sum |ref - pred| over all channels (l1_diff) and, per frame, the
// intersection / union pixel counts of the masks thresholded at 0.5 (sil_iou).  One CTA per 1024 pixels of one frame.
__global__ void __launch_bounds__(TPB) image_metrics_kernel(int S2, const float* __restrict__ ref_img, const float* __restrict__ pred_img,
                                                             const float* __restrict__ ref_mask, const float* __restrict__ pred_mask,
                                                             float* __restrict__ l1_sum, int32_t* __restrict__ inter,
                                                             int32_t* __restrict__ uni) {
  __shared__ float red[32];
  __shared__ int cnt[2];
  int b = blockIdx.y;
  if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
  __syncthreads();
  float acc = 0.f;
  int ni = 0, nu = 0;
  for (int k = 0; k < 4; ++k) {
    int p = (blockIdx.x * 4 + k) * TPB + threadIdx.x;
    if (p >= S2) break;
    size_t pix = (size_t)b * S2 + p;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) acc += fabsf(ref_img[pix * 3 + ch] - pred_img[pix * 3 + ch]);
    bool r = ref_mask[pix] >= 0.5f, q = pred_mask[pix] >= 0.5f;
    ni += (r && q) ? 1 : 0;
    nu += (r || q) ? 1 : 0;
  }
  float tot = block_sum(acc, red);
  if (ni) atomicAdd(&cnt[0], ni);
  if (nu) atomicAdd(&cnt[1], nu);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (tot != 0.f) atomicAdd(l1_sum, tot);
    if (cnt[0]) atomicAdd(inter + b, cnt[0]);
    if (cnt[1]) atomicAdd(uni + b, cnt[1]);
  }
}

struct AdamState {
  int step;
  unsigned int done;  // CTAs of the running update that have finished (the last one advances `step` and re-arms this)
  float unused;
  float lr;           // read only when the call passes lr < 0
};

// One launch per parameter group: every thread derives the bias corrections of step + 1 itself (two powf against a 20-byte
// read-modify-write per element), and the last CTA to finish advances the step counter — the separate one-thread "tick"
// launch this replaces was a graph node on the step's critical path.
__global__ void __launch_bounds__(TPB) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                        AdamState* __restrict__ st) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int s = st->step + 1;  // (nobody writes `step` before every CTA has read it: see below)
  if (i < n) {
    float bc1 = 1.f - powf(b1, (float)s), bc2s = sqrtf(1.f - powf(b2, (float)s));
    if (lr < 0.f) lr = st->lr;  // learning rate on the device: a scheduler can change it under a captured graph
    float gi = g[i];
    float mi = m[i] + (1.f - b1) * (gi - m[i]);  // lerp, as torch._single_tensor_adam
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= (lr / bc1) * (mi / (sqrtf(vi) / bc2s + eps));
  }
  __syncthreads();  // every thread of this CTA has read `step`
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&st->done, 1u) == gridDim.x - 1) {  // the last CTA: all others are past their read of `step`
      st->done = 0u;
      st->step = s;
    }
  }
}

}  // namespace

extern "C" int harp_adam_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, void* state, void* stream) {
  HARP_REQUIRE(param && grad && exp_avg && exp_avg_sq && state && n >= 0, HARP_ERR_INVALID, "harp_adam_dev: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  adam_dev_kernel<<<(int)((std::max<int64_t>(n, 1) + TPB - 1) / TPB), TPB, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                                   eps, (AdamState*)state);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

extern "C" int harp_mse(int64_t n, const float* x, const float* target, float scale, float* loss, float* d_x, void* stream) {
  HARP_REQUIRE(n > 0 && x && target && loss, HARP_ERR_INVALID, "harp_mse: bad argument");
  mse_kernel<<<(int)((n + TPB - 1) / TPB), TPB, 0, (cudaStream_t)stream>>>(n, x, target, scale, loss, d_x);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

extern "C" int harp_image_metrics(int B, int S, const float* ref_image, const float* pred_image, const float* ref_mask,
                                  const float* pred_mask, float* l1_sum, int32_t* inter, int32_t* uni, void* stream) {
  HARP_REQUIRE(B > 0 && S > 0 && ref_image && pred_image && ref_mask && pred_mask && l1_sum && inter && uni, HARP_ERR_INVALID,
               "harp_image_metrics: bad argument");
  int S2 = S * S;
  dim3 grid((S2 + 4 * TPB - 1) / (4 * TPB), B);
  image_metrics_kernel<<<grid, TPB, 0, (cudaStream_t)stream>>>(S2, ref_image, pred_image, ref_mask, pred_mask, l1_sum, inter, uni);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

extern "C" int harp_ingest_u8(int B, int S, const uint8_t* rgb_u8, const uint8_t* mask_u8, float* y_true, float* y_sil,
                              float* y_sil_eroded, void* stream) {
  HARP_REQUIRE(B > 0 && S > 0 && rgb_u8 && mask_u8 && y_true && y_sil && y_sil_eroded, HARP_ERR_INVALID, "harp_ingest_u8: bad argument");
  size_t n = (size_t)B * S * S;
  ingest_u8_kernel<<<(int)((n + TPB - 1) / TPB), TPB, 0, (cudaStream_t)stream>>>(B, S, rgb_u8, mask_u8, y_true, y_sil, y_sil_eroded);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

extern "C" int harp_texture_reg(harp_ctx* c, int kind, const float* map, int H, int W, const int32_t* offsets,
                                const float* uv_mask, float weight, float* loss, float* d_map, void* stream) {
  (void)c;
  HARP_REQUIRE(map && offsets && loss && H > 0 && W > 0, HARP_ERR_INVALID, "harp_texture_reg: null argument");
  HARP_REQUIRE(kind == 0 || kind == 1, HARP_ERR_INVALID, "harp_texture_reg: kind must be 0 (albedo) or 1 (normal)");
  cudaStream_t st = (cudaStream_t)stream;
  tex_smooth_kernel<<<(H * W + TPB - 1) / TPB, TPB, 0, st>>>(map, H, W, offsets, uv_mask, weight, loss, d_map);
  if (kind == 1) close_to_z_kernel<<<H, TPB, 0, st>>>(map, H, W, weight, loss, d_map);
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}

extern "C" int harp_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                         float beta1, float beta2, float eps, int step, void* stream) {
  HARP_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, HARP_ERR_INVALID, "harp_adam: bad argument");
  if (n == 0) return HARP_OK;
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<(int)((n + TPB - 1) / TPB), TPB, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                             beta2, eps, bc1, sqrtf(bc2));
  HARP_LAUNCH_CHECK();
  return HARP_OK;
}
