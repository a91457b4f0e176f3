// Landmark heat-map and the fused HybridAttention + global-average-pool kernel.
//
// Heat-map (LandmarkAttention._create_attention_map, landmark_attention.py:76-130): the fp32
// operation order of the eager reference is kept instruction for instruction (explicit
// __fmul_rn/__fadd_rn/__fdiv_rn so the compiler cannot contract into FMAs), so the scaled
// coordinates are bit-identical and the map agrees to the last ulp or two of expf.
//
// HybridAttention (landmark -> channel -> spatial) + adaptive_avg_pool2d: one CTA per image,
// three streaming passes over the (H*W) x C map (0.5 MB in bf16, L2 resident after the head
// GEMM): channel statistics -> spatial statistics -> gated pooled features.  The attended map
// (three full read+write round trips in the reference) is never written.  All arithmetic fp32,
// as the autocast reference promotes at the landmark product.
#include "common.cuh"

namespace dfv {

__global__ void heat_raw_kernel(const float* __restrict__ lm, const float* __restrict__ w5, float* __restrict__ raw,
                                uint32_t* __restrict__ gmax, float* __restrict__ scaled_xy, int B, int H, int W,
                                float sx, float sy, float denom, int group) {
  pdl_prologue();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (idx >= B * HW) return;
  const int b = idx / HW, pos = idx % HW;
  const float yv = (float)(pos / W), xv = (float)(pos % W);
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const float lx = __fmul_rn(lm[(b * 5 + i) * 2 + 0], sx);
    const float ly = __fmul_rn(lm[(b * 5 + i) * 2 + 1], sy);
    if (scaled_xy != nullptr && pos == 0) {
      scaled_xy[(b * 5 + i) * 2 + 0] = lx;
      scaled_xy[(b * 5 + i) * 2 + 1] = ly;
    }
    const float dx = __fsub_rn(xv, lx), dy = __fsub_rn(yv, ly);
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const float g = expf(__fdiv_rn(-d2, denom));
    a = __fadd_rn(a, __fmul_rn(g, w5[i]));
  }
  raw[idx] = a;
  // group maximum: values with positive weights are >= 0, where float order == uint order;
  // negative sums (negative learnt weights) are handled by an order-preserving key.
  uint32_t key = __float_as_uint(a);
  key = (key & 0x80000000u) ? ~key : (key | 0x80000000u);
  atomicMax(&gmax[b / group], key);
}

__global__ void heat_norm_kernel(const float* __restrict__ raw, const uint32_t* __restrict__ gmax,
                                 float* __restrict__ heat, int B, int HW, int group) {
  pdl_prologue();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * HW) return;
  uint32_t key = gmax[(idx / HW) / group];
  key = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
  const float mx = __uint_as_float(key);
  float v = __fdiv_rn(raw[idx], __fadd_rn(mx, 1e-8f));
  heat[idx] = fminf(fmaxf(v, 0.1f), 1.0f);
}

// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) hybrid_attention_kernel(
    const T* __restrict__ fmap, const float* __restrict__ heat, const float* __restrict__ w1,
    const float* __restrict__ w2t, const float* __restrict__ sa_w, float* __restrict__ features,
    float* __restrict__ channel_gate, float* __restrict__ spatial_gate, int H, int W, int C, int hidden,
    int use_channel, int use_spatial) {
  pdl_prologue();
  extern __shared__ float sm[];
  const int HW = H * W;
  float* a_lm = sm;                 // [HW]   landmark gate (1 if absent)
  float* avg_c = a_lm + HW;         // [C]
  float* max_c = avg_c + C;         // [C]
  float* gate_c = max_c + C;        // [C]    sigmoid channel gate
  float* hid = gate_c + C;          // [hidden]
  float* sp_mean = hid + hidden;    // [HW]
  float* sp_max = sp_mean + HW;     // [HW]
  float* gate_p = sp_max + HW;      // [HW]   landmark * spatial gate

  const int b = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const T* fb = fmap + (size_t)b * HW * C;
  const int CV = C >> 3;            // 8-channel vectors
  const float inv_hw = 1.0f / (float)HW;

  for (int p = tid; p < HW; p += blockDim.x) a_lm[p] = heat ? heat[(size_t)b * HW + p] : 1.0f;
  __syncthreads();

  if (use_channel) {
    // pass 1: per-channel mean / max over positions of x * A
    for (int cv = tid; cv < CV; cv += blockDim.x) {
      float s[8], m[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] = 0.f; m[e] = -INFINITY; }
#pragma unroll 8
      for (int p = 0; p < HW; ++p) {      // unrolled: eight independent 16-byte loads in flight per thread
        float v[8];
        load8(fb + (size_t)p * C + cv * 8, v);
        const float a = a_lm[p];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = v[e] * a;
          s[e] += x;
          m[e] = fmaxf(m[e], x);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { avg_c[cv * 8 + e] = s[e] * inv_hw; max_c[cv * 8 + e] = m[e]; }
    }
    __syncthreads();
    // shared MLP: relu(W1 avg) + relu(W1 max), then W2 (linear, no bias) applied once to the sum
    for (int j = warp; j < hidden; j += nwarps) {
      const float* wr = w1 + (size_t)j * C;
      float sa = 0.f, sx = 0.f;
#pragma unroll 8
      for (int c = lane; c < C; c += 32) {
        const float wv = wr[c];
        sa = fmaf(wv, avg_c[c], sa);
        sx = fmaf(wv, max_c[c], sx);
      }
      sa = warp_sum(sa);
      sx = warp_sum(sx);
      if (lane == 0) hid[j] = fmaxf(sa, 0.f) + fmaxf(sx, 0.f);
    }
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {
      float s = 0.f;
#pragma unroll 16
      for (int j = 0; j < hidden; ++j) s = fmaf(w2t[(size_t)j * C + c], hid[j], s);
      const float g = sigmoid_exact(s);
      gate_c[c] = g;
      if (channel_gate) channel_gate[(size_t)b * C + c] = g;
    }
  } else {
    for (int c = tid; c < C; c += blockDim.x) gate_c[c] = 1.0f;
  }
  __syncthreads();

  if (use_spatial) {
    // pass 2: per-position mean / max over channels of x * A * gate_c
    for (int p = warp; p < HW; p += nwarps) {
      const float a = a_lm[p];
      float s = 0.f, m = -INFINITY;
#pragma unroll 8
      for (int cv = lane; cv < CV; cv += 32) {
        float v[8];
        load8(fb + (size_t)p * C + cv * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = (v[e] * a) * gate_c[cv * 8 + e];
          s += x;
          m = fmaxf(m, x);
        }
      }
      s = warp_sum(s);
      m = warp_max(m);
      if (lane == 0) { sp_mean[p] = s / (float)C; sp_max[p] = m; }
    }
    __syncthreads();
    // 7x7 conv (2 -> 1, zero pad 3, no bias) + sigmoid
    for (int p = tid; p < HW; p += blockDim.x) {
      const int y = p / W, x = p % W;
      float s = 0.f;
      for (int ky = 0; ky < 7; ++ky) {
        const int yy = y + ky - 3;
        if (yy < 0 || yy >= H) continue;
        for (int kx = 0; kx < 7; ++kx) {
          const int xx = x + kx - 3;
          if (xx < 0 || xx >= W) continue;
          s = fmaf(sa_w[ky * 7 + kx], sp_mean[yy * W + xx], s);
          s = fmaf(sa_w[49 + ky * 7 + kx], sp_max[yy * W + xx], s);
        }
      }
      const float g = sigmoid_exact(s);
      if (spatial_gate) spatial_gate[(size_t)b * HW + p] = g;
      gate_p[p] = g;
    }
  } else {
    for (int p = tid; p < HW; p += blockDim.x) gate_p[p] = 1.0f;
  }
  __syncthreads();

  // pass 3: pooled features = mean_p ((x * A) * gate_c) * gate_p
  for (int cv = tid; cv < CV; cv += blockDim.x) {
    float s[8], gc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; gc[e] = gate_c[cv * 8 + e]; }
#pragma unroll 8
    for (int p = 0; p < HW; ++p) {
      float v[8];
      load8(fb + (size_t)p * C + cv * 8, v);
      const float a = a_lm[p], g = gate_p[p];
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] += ((v[e] * a) * gc[e]) * g;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) features[(size_t)b * C + cv * 8 + e] = s[e] * inv_hw;
  }
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_landmark_heatmap_fwd(const float* landmarks, const float* weights5, float* heat, float* raw_ws,
                                        uint32_t* max_ws, float* scaled_xy, int B, int H, int W, float ref_size,
                                        float sigma, int group, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(landmarks && weights5 && heat && raw_ws && max_ws, "dfv_landmark_heatmap_fwd: null pointer");
  DFV_REQUIRE(B > 0 && H > 0 && W > 0 && ref_size > 0.f && sigma > 0.f, "dfv_landmark_heatmap_fwd: bad shape");
  if (group <= 0 || group > B) group = B;
  const int n_groups = (B + group - 1) / group;
  cudaStream_t st = as_stream(stream);
  DFV_CUDA(cudaMemsetAsync(max_ws, 0, sizeof(uint32_t) * n_groups, st));
  // python: scale_x = W / 224.0 (double) multiplied into an fp32 tensor -> rounded to fp32 first
  const float sx = (float)((double)W / (double)ref_size), sy = (float)((double)H / (double)ref_size);
  const float denom = (float)(2.0 * (double)sigma * (double)sigma);
  const int total = B * H * W;
  ProfScope prof(PK_HEATMAP, 4.0 * (B * 10.0 + 3.0 * total), 60.0 * total, st);
  DFV_PDL((heat_raw_kernel), (total + 255) / 256, 256, 0, st, landmarks, weights5, raw_ws, max_ws, scaled_xy, B, H, W, sx, sy,
                                                      denom, group);
  DFV_LAUNCH_CHECK();
  DFV_PDL((heat_norm_kernel), (total + 255) / 256, 256, 0, st, raw_ws, max_ws, heat, B, H * W, group);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

extern "C" int dfv_hybrid_attention_fwd(const void* fmap, const float* heat, const float* ca_w1, const float* ca_w2_t,
                                        const float* sa_w, float* features, float* channel_gate, float* spatial_gate,
                                        int dtype, int B, int H, int W, int C, int hidden, int use_channel,
                                        int use_spatial, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(fmap && features, "dfv_hybrid_attention_fwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype), "dfv_hybrid_attention_fwd: bad dtype %d", dtype);
  DFV_REQUIRE(!use_channel || (ca_w1 && ca_w2_t && hidden > 0), "dfv_hybrid_attention_fwd: channel attention needs weights");
  DFV_REQUIRE(!use_spatial || sa_w, "dfv_hybrid_attention_fwd: spatial attention needs weights");
  DFV_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "dfv_hybrid_attention_fwd: bad shape (C %% 8 == 0)");
  const size_t smem = sizeof(float) * ((size_t)4 * H * W + 3 * (size_t)C + (size_t)(hidden > 0 ? hidden : 0));
  DFV_REQUIRE(smem <= 200 * 1024, "dfv_hybrid_attention_fwd: map too large for one CTA (H*W=%d C=%d)", H * W, C);
  cudaStream_t st = as_stream(stream);
  // algorithmic minimum: channel stats need all positions, spatial stats need all channels after the
  // channel gate -> 2 reads of the map; this kernel does 3 (L2-resident)
  ProfScope prof(PK_ATTENTION, 2.0 * (double)B * H * W * C * dtype_size(dtype) + 4.0 * B * C, 8.0 * (double)B * H * W * C, st);
  if (dtype == DFV_BF16) {
    auto k = hybrid_attention_kernel<__nv_bfloat16>;
    if (smem > 48 * 1024) DFV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DFV_PDL((k), B, 256, smem, st, (const __nv_bfloat16*)fmap, heat, ca_w1, ca_w2_t, sa_w, features, channel_gate, spatial_gate, H, W,
                            C, hidden, use_channel, use_spatial);
  } else {
    auto k = hybrid_attention_kernel<float>;
    if (smem > 48 * 1024) DFV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DFV_PDL((k), B, 256, smem, st, (const float*)fmap, heat, ca_w1, ca_w2_t, sa_w, features, channel_gate, spatial_gate, H, W, C,
                            hidden, use_channel, use_spatial);
  }
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}
