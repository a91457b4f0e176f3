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
#include "small_linear.cuh"

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
  // one atomic per (warp, group) instead of per thread: with one batch-wide group every thread hit the same word
  // (36864 serialised atomics = 25 us for a 147 KB map)
  const int gi = b / group;
  const unsigned active = __activemask();
  const unsigned same = __match_any_sync(active, gi);
  if (active == 0xffffffffu && same == active) {      // warp-uniform: the whole warp is in one group (the common case)
    uint32_t m = key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(&gmax[gi], m);
  } else {
    atomicMax(&gmax[gi], key);
  }
}

__global__ void heat_norm_kernel(const float* __restrict__ raw, const uint32_t* __restrict__ gmax,
                                 float* __restrict__ heat, int B, int HW, int group, const uint32_t* __restrict__ floor_key) {
  pdl_prologue();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * HW) return;
  uint32_t key = gmax[(idx / HW) / group];
  if (floor_key != nullptr) key = max(key, floor_key[0]);      // keys are order-preserving: the larger key is the larger float
  key = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
  const float mx = __uint_as_float(key);
  float v = __fdiv_rn(raw[idx], __fadd_rn(mx, 1e-8f));
  heat[idx] = fminf(fmaxf(v, 0.1f), 1.0f);
}

// ------------------------------------------------------------------------------------
// HybridAttention + global average pool, inference.  The one-CTA-per-image kernel ran the channel-attention MLP inside
// every CTA: each of B CTAs walked both FC matrices (0.8 MB each) from L2 in ~150 dependent load batches -- 270 us for
// 132 MB of map.  Now the map passes are per-image streaming kernels with 16 loads in flight per thread, and the MLP is
// the batch-wide small-linear chain (small_linear.cuh) shared with the squeeze-excite gates.

// pass 1: per-channel mean / max over positions of x * A
// grid = B; thread = (8-channel vector, position phase): kPhases threads share a vector and take every kPhases-th
// position (more loads in flight per SM than one thread per vector), partial results meet in shared memory.
constexpr int kAttnPhases = 4;
template <typename T>
__global__ void __launch_bounds__(1024) attn_pool_kernel(const T* __restrict__ fmap, const float* __restrict__ heat,
                                                         float* __restrict__ avg, float* __restrict__ mx, int HW, int C) {
  pdl_prologue();
  extern __shared__ float sm[];
  float* a_lm = sm;                           // [HW]   landmark gate (1 if absent)
  float* red = sm + HW;                       // [kAttnPhases - 1][2][C]
  const int b = blockIdx.x, tid = threadIdx.x;
  const T* fb = fmap + (size_t)b * HW * C;
  const int CV = C >> 3;
  for (int p = tid; p < HW; p += blockDim.x) a_lm[p] = heat ? heat[(size_t)b * HW + p] : 1.0f;
  __syncthreads();
  const int nvec = blockDim.x / kAttnPhases;  // vectors per sweep
  const int ph = tid / nvec, lv = tid % nvec;
  const float inv_hw = 1.0f / (float)HW;
  for (int cv0 = 0; cv0 < CV; cv0 += nvec) {
    const int cv = cv0 + lv;
    float s[8], m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; m[e] = -INFINITY; }
    if (cv < CV) {
#pragma unroll 12
      for (int p = ph; p < HW; p += kAttnPhases) {      // unrolled: twelve independent 16-byte loads in flight per thread
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
      if (ph > 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { red[((size_t)(ph - 1) * 2 + 0) * C + cv * 8 + e] = s[e]; red[((size_t)(ph - 1) * 2 + 1) * C + cv * 8 + e] = m[e]; }
      }
    }
    __syncthreads();
    if (ph == 0 && cv < CV) {
#pragma unroll
      for (int q = 0; q < kAttnPhases - 1; ++q)      // fixed order: phase 0 + 1 + 2 + 3
#pragma unroll
        for (int e = 0; e < 8; ++e) { s[e] += red[((size_t)q * 2 + 0) * C + cv * 8 + e]; m[e] = fmaxf(m[e], red[((size_t)q * 2 + 1) * C + cv * 8 + e]); }
#pragma unroll
      for (int e = 0; e < 8; ++e) { avg[(size_t)b * C + cv * 8 + e] = s[e] * inv_hw; mx[(size_t)b * C + cv * 8 + e] = m[e]; }
    }
    __syncthreads();
  }
}

// pass 2: per-position mean / max over channels of x * A * gate_c      (one warp per (image, position))
template <typename T>
__global__ void __launch_bounds__(256) attn_spatial_kernel(const T* __restrict__ fmap, const float* __restrict__ heat,
                                                           const float* __restrict__ gate_c, float* __restrict__ sp_mean,
                                                           float* __restrict__ sp_max, int B, int HW, int C) {
  pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wp = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (wp >= (long long)B * HW) return;
  const int b = (int)(wp / HW);
  const T* row = fmap + (size_t)wp * C;
  const float* gc = gate_c ? gate_c + (size_t)b * C : nullptr;
  const float a = heat ? heat[wp] : 1.0f;
  const int CV = C >> 3;
  float s = 0.f, m = -INFINITY;
#pragma unroll 8
  for (int cv = lane; cv < CV; cv += 32) {
    float v[8];
    load8(row + cv * 8, v);
    float g[8];
    if (gc) {
      const float4 g0 = *reinterpret_cast<const float4*>(gc + cv * 8), g1 = *reinterpret_cast<const float4*>(gc + cv * 8 + 4);
      g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = 1.0f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x = (v[e] * a) * g[e];
      s += x;
      m = fmaxf(m, x);
    }
  }
  s = warp_sum(s);
  m = warp_max(m);
  if (lane == 0) { sp_mean[wp] = s / (float)C; sp_max[wp] = m; }
}

// pass 3: 7x7 conv (2 -> 1, zero pad 3, no bias) + sigmoid on the spatial statistics, then
// pooled features = mean_p ((x * A) * gate_c) * gate_p      (grid = B, thread = 8-channel vector)
template <typename T>
__global__ void __launch_bounds__(1024) attn_feature_kernel(const T* __restrict__ fmap, const float* __restrict__ heat,
                                                           const float* __restrict__ gate_c, const float* __restrict__ sp_mean,
                                                           const float* __restrict__ sp_max, const float* __restrict__ sa_w,
                                                           float* __restrict__ features, float* __restrict__ spatial_gate, int H,
                                                           int W, int C, int use_spatial) {
  pdl_prologue();
  extern __shared__ float sm[];
  const int HW = H * W;
  float* wgt = sm;                  // [HW]   spatial gate
  float* spm = wgt + HW;            // [HW]
  float* spx = spm + HW;            // [HW]
  float* alm = spx + HW;            // [HW]   landmark gate (1 if absent)
  float* red = alm + HW;            // [kAttnPhases - 1][C]
  const int b = blockIdx.x, tid = threadIdx.x;
  const T* fb = fmap + (size_t)b * HW * C;
  const int CV = C >> 3;
  if (use_spatial) {
    for (int p = tid; p < HW; p += blockDim.x) { spm[p] = sp_mean[(size_t)b * HW + p]; spx[p] = sp_max[(size_t)b * HW + p]; }
    __syncthreads();
  }
  for (int p = tid; p < HW; p += blockDim.x) {
    float g = 1.0f;
    if (use_spatial) {
      const int y = p / W, x = p % W;
      float s = 0.f;
      for (int ky = 0; ky < 7; ++ky) {
        const int yy = y + ky - 3;
        if (yy < 0 || yy >= H) continue;
        for (int kx = 0; kx < 7; ++kx) {
          const int xx = x + kx - 3;
          if (xx < 0 || xx >= W) continue;
          s = fmaf(sa_w[ky * 7 + kx], spm[yy * W + xx], s);
          s = fmaf(sa_w[49 + ky * 7 + kx], spx[yy * W + xx], s);
        }
      }
      g = sigmoid_exact(s);
      if (spatial_gate) spatial_gate[(size_t)b * HW + p] = g;
    }
    wgt[p] = g;      // the landmark gate is applied separately below: ((x * A) * gate_c) * gate_p, as the reference orders it
    alm[p] = heat ? heat[(size_t)b * HW + p] : 1.0f;
  }
  __syncthreads();
  const float inv_hw = 1.0f / (float)HW;
  const int nvec = blockDim.x / kAttnPhases;
  const int ph = tid / nvec, lv = tid % nvec;
  for (int cv0 = 0; cv0 < CV; cv0 += nvec) {
    const int cv = cv0 + lv;
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    if (cv < CV) {
      float gc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) gc[e] = gate_c ? gate_c[(size_t)b * C + cv * 8 + e] : 1.0f;
#pragma unroll 12
      for (int p = ph; p < HW; p += kAttnPhases) {
        float v[8];
        load8(fb + (size_t)p * C + cv * 8, v);
        const float a = alm[p], g = wgt[p];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += ((v[e] * a) * gc[e]) * g;
      }
      if (ph > 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) red[(size_t)(ph - 1) * C + cv * 8 + e] = s[e];
      }
    }
    __syncthreads();
    if (ph == 0 && cv < CV) {
#pragma unroll
      for (int q = 0; q < kAttnPhases - 1; ++q)
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += red[(size_t)q * C + cv * 8 + e];
#pragma unroll
      for (int e = 0; e < 8; ++e) features[(size_t)b * C + cv * 8 + e] = s[e] * inv_hw;
    }
    __syncthreads();
  }
}

// scratch layout (floats): avg [B][C] | max [B][C] | gate_c [B][C] | partA [ks][B][hid] | partM [ks][B][hid] | hid [B][hid]
//                          | sp_mean [B][HW] | sp_max [B][HW]
static size_t attention_scratch_floats(int B, int HW, int C, int hidden) {
  const size_t h = hidden > 0 ? hidden : 0;
  return (size_t)3 * B * C + ((size_t)2 * sl_ksplit_rowmajor(C) + 1) * B * h + (size_t)2 * B * HW;
}

template <typename T>
static int launch_attention(const T* fmap, const float* heat, const float* ca_w1, const float* ca_w2_t, const float* sa_w,
                            float* features, float* channel_gate, float* spatial_gate, float* scratch, int B, int H, int W, int C,
                            int hidden, int use_channel, int use_spatial, cudaStream_t st) {
  const int HW = H * W;
  const int ks = sl_ksplit_rowmajor(C);
  const size_t h = hidden > 0 ? hidden : 0;
  // threads: kAttnPhases position phases x enough 8-channel vectors to cover C in one sweep (<= 1024)
  int attn_threads = ((C / 8 + 31) / 32 * 32) * kAttnPhases;
  if (attn_threads > 1024) attn_threads = 1024;
  const size_t smem_pool = sizeof(float) * ((size_t)HW + (size_t)(kAttnPhases - 1) * 2 * C);
  const size_t smem_feat = sizeof(float) * ((size_t)4 * HW + (size_t)(kAttnPhases - 1) * C);
  {
    static thread_local bool configured = false;
    if (!configured) {
      DFV_CUDA(cudaFuncSetAttribute(attn_pool_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      DFV_CUDA(cudaFuncSetAttribute(attn_feature_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      configured = true;
    }
  }
  DFV_REQUIRE(smem_pool <= 160 * 1024 && smem_feat <= 160 * 1024, "dfv_hybrid_attention_fwd: map too large (H*W=%d C=%d)", HW, C);
  float* avg = scratch;
  float* mx = avg + (size_t)B * C;
  float* gate_c = channel_gate ? channel_gate : mx + (size_t)B * C;
  float* part_a = scratch + (size_t)3 * B * C;
  float* part_m = part_a + (size_t)ks * B * h;
  float* hid = part_m + (size_t)ks * B * h;
  float* sp_mean = hid + (size_t)B * h;
  float* sp_max = sp_mean + (size_t)B * HW;
  if (use_channel) {
    DFV_PDL((attn_pool_kernel<T>), B, attn_threads, smem_pool, st, fmap, heat, avg, mx, HW, C);
    // shared MLP: relu(W1 avg) + relu(W1 max), then W2 (linear, no bias) applied once to the sum, sigmoid
    static thread_local bool configured = false;
    if (!configured) {
      DFV_CUDA(cudaFuncSetAttribute(sl_rowmajor_partial_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      DFV_CUDA(cudaFuncSetAttribute(sl_kmajor_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      configured = true;
    }
    DFV_REQUIRE(sl_rowmajor_smem(hidden) <= 160 * 1024 && sl_kmajor_smem(hidden) <= 100 * 1024,
                "dfv_hybrid_attention_fwd: channel-attention hidden width %d too large", hidden);
    const dim3 grid_a((unsigned)ks, (unsigned)((B + kSlRows - 1) / kSlRows), 2);      // z: the mean and the max statistics
    DFV_PDL((sl_rowmajor_partial_kernel<false>), grid_a, kSlThreads, sl_rowmajor_smem(hidden), st, (const float*)avg, 1, 1.0f, ca_w1, part_a,
            (float*)nullptr, B, C, hidden, (size_t)B * C, (size_t)ks * B * h);
    DFV_PDL(sl_combine_kernel, (unsigned)(((size_t)B * hidden + kSlThreads - 1) / kSlThreads), kSlThreads, 0, st, (const float*)part_a,
            (const float*)part_m, ks, (const float*)nullptr, hid, B, hidden, 1);
    const dim3 grid_b((unsigned)((C + kSlCols - 1) / kSlCols), (unsigned)((B + kSlRows - 1) / kSlRows));
    DFV_PDL((sl_kmajor_kernel<float, false>), grid_b, kSlThreads, sl_kmajor_smem(hidden), st, (const float*)hid, ca_w2_t, (const float*)nullptr,
            gate_c, (float*)nullptr, (float*)nullptr, B, C, hidden, hidden, 0, SL_OUT_SIGMOID, (const long long*)nullptr, (const float*)nullptr, 1.0f);
    count_launch(4);
  }
  const float* gc = use_channel ? gate_c : nullptr;
  if (use_spatial) {
    const long long warps = (long long)B * HW;
    DFV_PDL((attn_spatial_kernel<T>), (unsigned)((warps + 7) / 8), 256, 0, st, fmap, heat, gc, sp_mean, sp_max, B, HW, C);
    count_launch(1);
  }
  DFV_PDL((attn_feature_kernel<T>), B, attn_threads, smem_feat, st, fmap, heat, gc, (const float*)sp_mean, (const float*)sp_max, sa_w,
          features, spatial_gate, H, W, C, use_spatial);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_landmark_heatmap_fwd(const float* landmarks, const float* weights5, float* heat, float* raw_ws,
                                        uint32_t* max_ws, float* scaled_xy, int B, int H, int W, float ref_size,
                                        float sigma, int group, dfv_stream_t stream) {
  return dfv_landmark_heatmap_fwd_ex(landmarks, weights5, heat, raw_ws, max_ws, scaled_xy, B, H, W, ref_size, sigma, group, nullptr, stream);
}

extern "C" int dfv_landmark_heatmap_fwd_ex(const float* landmarks, const float* weights5, float* heat, float* raw_ws,
                                           uint32_t* max_ws, float* scaled_xy, int B, int H, int W, float ref_size,
                                           float sigma, int group, const uint32_t* max_floor, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(landmarks && weights5 && heat && raw_ws && max_ws, "dfv_landmark_heatmap_fwd: null pointer");
  DFV_REQUIRE(B > 0 && H > 0 && W > 0 && ref_size > 0.f && sigma > 0.f, "dfv_landmark_heatmap_fwd: bad shape");
  if (group <= 0 || group > B) group = B;
  DFV_REQUIRE(max_floor == nullptr || group == B, "dfv_landmark_heatmap_fwd_ex: a maximum floor needs the whole-call group");
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
  DFV_PDL((heat_norm_kernel), (total + 255) / 256, 256, 0, st, raw_ws, max_ws, heat, B, H * W, group, max_floor);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

extern "C" size_t dfv_attention_scratch_floats(int B, int H, int W, int C, int hidden) {
  return B > 0 && H > 0 && W > 0 && C > 0 ? attention_scratch_floats(B, H * W, C, hidden) : 0;
}

extern "C" int dfv_hybrid_attention_fwd(const void* fmap, const float* heat, const float* ca_w1, const float* ca_w2_t,
                                        const float* sa_w, float* features, float* channel_gate, float* spatial_gate,
                                        float* scratch, int dtype, int B, int H, int W, int C, int hidden, int use_channel,
                                        int use_spatial, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(fmap && features && scratch, "dfv_hybrid_attention_fwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype), "dfv_hybrid_attention_fwd: bad dtype %d", dtype);
  DFV_REQUIRE(!use_channel || (ca_w1 && ca_w2_t && hidden > 0), "dfv_hybrid_attention_fwd: channel attention needs weights");
  DFV_REQUIRE(!use_spatial || sa_w, "dfv_hybrid_attention_fwd: spatial attention needs weights");
  DFV_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "dfv_hybrid_attention_fwd: bad shape (C %% 8 == 0)");
  cudaStream_t st = as_stream(stream);
  // algorithmic minimum: channel stats need all positions, spatial stats need all channels after the
  // channel gate -> 2 reads of the map; three passes run here (the map is L2-resident between them)
  ProfScope prof(PK_ATTENTION, 2.0 * (double)B * H * W * C * dtype_size(dtype) + 4.0 * B * C, 8.0 * (double)B * H * W * C, st);
  if (dtype == DFV_BF16)
    return launch_attention((const __nv_bfloat16*)fmap, heat, ca_w1, ca_w2_t, sa_w, features, channel_gate, spatial_gate, scratch, B, H, W, C,
                            hidden, use_channel, use_spatial, st);
  return launch_attention((const float*)fmap, heat, ca_w1, ca_w2_t, sa_w, features, channel_gate, spatial_gate, scratch, B, H, W, C, hidden,
                          use_channel, use_spatial, st);
}
