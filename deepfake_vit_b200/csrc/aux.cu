// Small operators either side of the hot path:
//   clip_aggregate   per-clip mean logit, mean fake probability and label of the video-scoring rule (task.ipynb:434-442)
//   global_avg_pool  NHWC feature map -> [B][C] fp32 (extract_multi_scale_features, feature_extractor.py:119-154)
//   l2_normalize     F.normalize(p=2, dim=1) of the pooled features (get_embedding, feature_extractor.py:156-178)
//   grad_sqnorm / clip_adamw   the optimizer step around the path (trainer.py:158-167): global-norm
//                    clip_grad_norm_(max_norm) + AdamW over the FLAT fp32 parameter / gradient / moment buffers
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace dfv {

// one warp per clip
__global__ void clip_aggregate_kernel(const float* __restrict__ logits, int n_clips, int frames, int n_classes,
                                      float* __restrict__ mean_logits, float* __restrict__ fake_prob,
                                      int* __restrict__ labels, float threshold) {
  pdl_prologue();
  const int clip = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (clip >= n_clips) return;
  const float* l = logits + (size_t)clip * frames * n_classes;
  float prob = 0.f;
  for (int f = lane; f < frames; f += 32) {
    // softmax(logits)[1]
    float mx = -INFINITY;
    for (int c = 0; c < n_classes; ++c) mx = fmaxf(mx, l[f * n_classes + c]);
    float den = 0.f;
    for (int c = 0; c < n_classes; ++c) den += expf(l[f * n_classes + c] - mx);
    prob += expf(l[f * n_classes + 1] - mx) / den;
  }
  prob = warp_sum(prob) / (float)frames;
  for (int c = 0; c < n_classes; ++c) {
    float s = 0.f;
    for (int f = lane; f < frames; f += 32) s += l[f * n_classes + c];
    s = warp_sum(s);
    if (lane == 0) mean_logits[(size_t)clip * n_classes + c] = s / (float)frames;
  }
  if (lane == 0) {
    fake_prob[clip] = prob;
    labels[clip] = prob >= threshold ? 1 : 0;
  }
}

// thread = 8 channels of one image; rows are walked with 4 loads in flight
template <typename T>
__global__ void global_avg_pool_kernel(const T* __restrict__ x, float* __restrict__ out, int B, long long rows, int C) {
  pdl_prologue();
  const int CV = C >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * CV) return;
  const int cv = (int)(i % CV), b = (int)(i / CV);
  const T* base = x + (size_t)b * rows * C + cv * 8;
  float s[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = 0.f;
#pragma unroll 4
  for (long long r = 0; r < rows; ++r) {
    float v[8];
    load8(base + (size_t)r * C, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] += v[e];
  }
  const float inv = 1.0f / (float)rows;
#pragma unroll
  for (int e = 0; e < 8; ++e) out[(size_t)b * C + cv * 8 + e] = s[e] * inv;
}

// one warp per row: x / max(||x||_2, eps)
__global__ void l2_normalize_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int D, float eps) {
  pdl_prologue();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = x[(size_t)row * D + d];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  const float inv = 1.0f / fmaxf(sqrtf(s), eps);
  for (int d = lane; d < D; d += 32) y[(size_t)row * D + d] = x[(size_t)row * D + d] * inv;
}

// sum of squares of the flat gradient buffer: per-CTA partial -> one double atomic
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  pdl_prologue();
  __shared__ double red[8];
  double s = 0.0;       // double accumulation: the clip coefficient inherits this sum's relative error
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) s += (double)(g[i] * g[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

// torch.nn.utils.clip_grad_norm_ (clip_coef = max_norm / (total_norm + 1e-6), clamped to 1) followed by
// torch.optim.AdamW (decoupled weight decay, bias-corrected moments), one pass over the flat buffers.
__global__ void __launch_bounds__(256) clip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, const double* __restrict__ sqnorm,
                                                        float max_norm, float grad_scale, float decay, float beta1, float omb1,
                                                        float beta2, float omb2, float eps, float step_size, float bc2_sqrt,
                                                        float* __restrict__ total_norm_out, const unsigned char* __restrict__ frozen) {
  pdl_prologue();
  const float total = sqrtf((float)*sqnorm) * grad_scale;
  float coef = grad_scale;
  if (max_norm > 0.f) coef *= fminf(max_norm / (total + 1e-6f), 1.0f);
  if (total_norm_out && blockIdx.x == 0 && threadIdx.x == 0) *total_norm_out = total;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (frozen && frozen[i >> 6]) continue;    // a parameter the optimizer must not touch (requires_grad False / no gradient)
    const float gi = g[i] * coef;
    float pi = p[i];
    pi *= decay;                                           // param.mul_(1 - lr * weight_decay)
    const float mi = fmaf(omb1, gi - m[i], m[i]);          // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(omb2 * gi, gi, beta2 * v[i]);    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

}  // namespace dfv

using namespace dfv;

extern "C" {

int dfv_clip_aggregate(const float* logits, int n_clips, int frames_per_clip, int n_classes, float* mean_logits, float* fake_prob,
                       int* labels, float threshold, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(logits && mean_logits && fake_prob && labels, "dfv_clip_aggregate: null pointer");
  DFV_REQUIRE(n_clips > 0 && frames_per_clip > 0 && n_classes >= 2, "dfv_clip_aggregate: bad shape");
  DFV_PDL((clip_aggregate_kernel), (n_clips + 3) / 4, 128, 0, as_stream(stream), logits, n_clips, frames_per_clip, n_classes, mean_logits,
                                                                        fake_prob, labels, threshold);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_global_avg_pool(const void* x, int dtype, float* out, int B, long long rows, int C, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(x && out && valid_dtype(dtype) && B > 0 && rows > 0 && C > 0 && C % 8 == 0, "dfv_global_avg_pool: bad arguments");
  const long long n = (long long)B * (C / 8);
  if (dtype == DFV_BF16)
    DFV_PDL((global_avg_pool_kernel<__nv_bfloat16>), (unsigned)((n + 127) / 128), 128, 0, as_stream(stream), (const __nv_bfloat16*)x, out, B, rows, C);
  else
    DFV_PDL((global_avg_pool_kernel<float>), (unsigned)((n + 127) / 128), 128, 0, as_stream(stream), (const float*)x, out, B, rows, C);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_l2_normalize(const float* x, float* y, int B, int D, float eps, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(x && y && B > 0 && D > 0, "dfv_l2_normalize: bad arguments");
  DFV_PDL((l2_normalize_kernel), (B + 3) / 4, 128, 0, as_stream(stream), x, y, B, D, eps);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

/* Optimizer step over flat fp32 buffers of n elements (parameters, gradients, exp_avg, exp_avg_sq):
 * total_norm = ||grad_scale * g||_2, clip to max_norm (<= 0: no clipping), then AdamW at `step` (1-based).
 * sqnorm_ws: one double of device scratch.  total_norm_out: optional device float (the value clip_grad_norm_ returns). */
int dfv_clip_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, double* sqnorm_ws,
                        double max_norm, double grad_scale, double lr, double beta1, double beta2, double eps, double weight_decay,
                        long long step, float* total_norm_out, const unsigned char* frozen_chunks, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(params && grads && exp_avg && exp_avg_sq && sqnorm_ws && n > 0 && step >= 1, "dfv_clip_adamw_step: bad arguments");
  cudaStream_t st = as_stream(stream);
  DFV_CUDA(cudaMemsetAsync(sqnorm_ws, 0, sizeof(double), st));
  const unsigned blocks = (unsigned)std::min<long long>((n / 4 + 255) / 256 + 1, 4LL * num_sms());
  DFV_PDL((grad_sqnorm_kernel), blocks, 256, 0, st, grads, n, sqnorm_ws);
  DFV_LAUNCH_CHECK();
  // hyper-parameter arithmetic in double on the host, as the Python optimizer does it
  const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
  DFV_PDL((clip_adamw_kernel), blocks, 256, 0, st, params, grads, exp_avg, exp_avg_sq, n, sqnorm_ws, (float)max_norm, (float)grad_scale,
                                           (float)(1.0 - lr * weight_decay), (float)beta1, (float)(1.0 - beta1), (float)beta2,
                                           (float)(1.0 - beta2), (float)eps, (float)(lr / bc1), (float)std::sqrt(bc2), total_norm_out, frozen_chunks);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

}  // extern "C"
