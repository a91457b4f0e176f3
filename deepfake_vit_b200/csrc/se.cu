// Squeeze-excite gate: finishes the depthwise kernel's pool partial sums, then
// C -> squeeze (bias, swish) -> C (bias, sigmoid).  One CTA per image; all fp32.
// Negligible bytes (pool partials + the two small FC matrices, L2-resident).
#include "common.cuh"

namespace dfv {

// IMG images per CTA share every weight load (the two FC matrices are the only real traffic).
template <int IMG, typename GT>
__global__ void __launch_bounds__(512) se_gate_kernel(const float* __restrict__ partial, int parts, float inv_hw,
                                                     const float* __restrict__ w1, const float* __restrict__ b1,
                                                     const float* __restrict__ w2t, const float* __restrict__ b2,
                                                     GT* __restrict__ gate, int B, int C, int sq) {
  extern __shared__ float sm[];
  float* pooled = sm;               // [IMG][C]
  float* hidden = sm + IMG * C;     // [IMG][sq]
  const int b0 = blockIdx.x * IMG, tid = threadIdx.x;
  for (int i = tid; i < IMG * C; i += blockDim.x) {
    const int im = i / C, c = i % C;
    float s = 0.f;
    if (b0 + im < B) {
      const float* pb = partial + (size_t)(b0 + im) * parts * C + c;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int t = 0;
      for (; t + 4 <= parts; t += 4) {
        s0 += pb[(size_t)t * C];
        s1 += pb[(size_t)(t + 1) * C];
        s2 += pb[(size_t)(t + 2) * C];
        s3 += pb[(size_t)(t + 3) * C];
      }
      for (; t < parts; ++t) s0 += pb[(size_t)t * C];
      s = (s0 + s1) + (s2 + s3);
    }
    pooled[i] = s * inv_hw;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < sq; j += nwarps) {
    const float* wr = w1 + (size_t)j * C;
    float s[IMG];
#pragma unroll
    for (int im = 0; im < IMG; ++im) s[im] = 0.f;
#pragma unroll 4
    for (int c = lane; c < C; c += 32) {
      const float wv = wr[c];
#pragma unroll
      for (int im = 0; im < IMG; ++im) s[im] = fmaf(wv, pooled[im * C + c], s[im]);
    }
#pragma unroll
    for (int im = 0; im < IMG; ++im) {
      float v = warp_sum(s[im]);
      if (lane == 0) {
        v += b1[j];
        hidden[im * sq + j] = v * sigmoid_exact(v);
      }
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    float s[IMG];
    const float bv = b2[c];
#pragma unroll
    for (int im = 0; im < IMG; ++im) s[im] = bv;
#pragma unroll 8
    for (int j = 0; j < sq; ++j) {
      const float wv = w2t[(size_t)j * C + c];
#pragma unroll
      for (int im = 0; im < IMG; ++im) s[im] = fmaf(wv, hidden[im * sq + j], s[im]);
    }
#pragma unroll
    for (int im = 0; im < IMG; ++im)
      if (b0 + im < B) {
        const float gv = sigmoid_exact(s[im]);
        if constexpr (sizeof(GT) == 2) gate[(size_t)(b0 + im) * C + c] = __float2bfloat16_rn(gv);
        else gate[(size_t)(b0 + im) * C + c] = gv;
      }
  }
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_se_gate_fwd(const float* pool_partial, int parts, float inv_hw, const float* w_reduce,
                               const float* b_reduce, const float* w_expand_t, const float* b_expand, void* gate,
                               int gate_dtype, int B, int C, int squeeze, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(pool_partial && w_reduce && b_reduce && w_expand_t && b_expand && gate, "dfv_se_gate_fwd: null pointer");
  DFV_REQUIRE(B > 0 && C > 0 && squeeze > 0 && parts > 0 && valid_dtype(gate_dtype), "dfv_se_gate_fwd: bad shape / dtype");
  if (debug_flags() & 2) return DFV_OK;
  const int img = B >= 2 * num_sms() ? 2 : 1;
  const size_t smem = (size_t)img * (C + squeeze) * sizeof(float);
  DFV_REQUIRE(smem <= 48 * 1024, "dfv_se_gate_fwd: C + squeeze too large (%d + %d)", C, squeeze);
  ProfScope prof(PK_SE_GATE, 4.0 * ((double)B * parts * C + (double)B * C + 2.0 * C * squeeze), 4.0 * B * (double)C * squeeze,
                 as_stream(stream));
#define SE_LAUNCH(IMG_, GT_)                                                                                     \
  se_gate_kernel<IMG_, GT_><<<(B + IMG_ - 1) / IMG_, 512, smem, as_stream(stream)>>>(                            \
      pool_partial, parts, inv_hw, w_reduce, b_reduce, w_expand_t, b_expand, (GT_*)gate, B, C, squeeze)
  if (gate_dtype == DFV_BF16) {
    if (img == 2) SE_LAUNCH(2, __nv_bfloat16); else SE_LAUNCH(1, __nv_bfloat16);
  } else {
    if (img == 2) SE_LAUNCH(2, float); else SE_LAUNCH(1, float);
  }
#undef SE_LAUNCH
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}
