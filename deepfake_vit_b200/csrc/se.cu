// Squeeze-excite gate: finishes the depthwise kernel's pool partial sums, then
// C -> squeeze (bias, swish) -> C (bias, sigmoid).  One CTA per image; all fp32.
// Negligible bytes (pool partials + the two small FC matrices, L2-resident).
#include "common.cuh"

namespace dfv {

__global__ void __launch_bounds__(256) se_gate_kernel(const float* __restrict__ partial, int parts, float inv_hw,
                                                     const float* __restrict__ w1, const float* __restrict__ b1,
                                                     const float* __restrict__ w2t, const float* __restrict__ b2,
                                                     float* __restrict__ gate, int C, int sq) {
  extern __shared__ float sm[];
  float* pooled = sm;        // [C]
  float* hidden = sm + C;    // [sq]
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* pb = partial + (size_t)b * parts * C;
  for (int c = tid; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < parts; ++t) s += pb[(size_t)t * C + c];
    pooled[c] = s * inv_hw;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < sq; j += nwarps) {
    const float* wr = w1 + (size_t)j * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(wr[c], pooled[c], s);
    s = warp_sum(s);
    if (lane == 0) {
      s += b1[j];
      hidden[j] = s * sigmoid_exact(s);
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    float s = b2[c];
    for (int j = 0; j < sq; ++j) s = fmaf(w2t[(size_t)j * C + c], hidden[j], s);
    gate[(size_t)b * C + c] = sigmoid_exact(s);
  }
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_se_gate_fwd(const float* pool_partial, int parts, float inv_hw, const float* w_reduce,
                               const float* b_reduce, const float* w_expand_t, const float* b_expand, float* gate, int B,
                               int C, int squeeze, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(pool_partial && w_reduce && b_reduce && w_expand_t && b_expand && gate, "dfv_se_gate_fwd: null pointer");
  DFV_REQUIRE(B > 0 && C > 0 && squeeze > 0 && parts > 0, "dfv_se_gate_fwd: bad shape");
  const size_t smem = (size_t)(C + squeeze) * sizeof(float);
  DFV_REQUIRE(smem <= 48 * 1024, "dfv_se_gate_fwd: C + squeeze too large (%d + %d)", C, squeeze);
  ProfScope prof(PK_SE_GATE, 4.0 * ((double)B * parts * C + (double)B * C + 2.0 * C * squeeze), 4.0 * B * (double)C * squeeze,
                 as_stream(stream));
  se_gate_kernel<<<B, 256, smem, as_stream(stream)>>>(pool_partial, parts, inv_hw, w_reduce, b_reduce, w_expand_t,
                                                      b_expand, gate, C, squeeze);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}
