// Classifier head (feature_extractor.py:223-238), eval mode: a chain of small Linear layers
// with BatchNorm1d folded in and ReLU between them, run in one kernel.  Each CTA handles
// kRows feature rows so the (L2-resident) weights are re-read B/kRows times, not B times.
#include "common.cuh"

namespace dfv {

constexpr int kHeadMaxLayers = 8;
constexpr int kLinRows = 8;     // feature rows per CTA
constexpr int kLinCols = 128;   // output columns per CTA (4 per lane)

// One Linear (+ folded BatchNorm1d) (+ ReLU) layer: out[r][n] = act(b[n] + sum_k in[r][k] * w_t[k][n]).
// CTA = 8 rows x 128 columns; its 8 warps split K (each keeps 8 x 4 accumulators per lane and has 8 weight
// loads in flight), then reduce through shared memory.  The previous single-kernel chain walked K = 1792
// sequentially per thread (~0.4 ms of pure L2 latency at batch 256).
__global__ void __launch_bounds__(256) linear_splitk_kernel(const float* __restrict__ in, const float* __restrict__ w_t,
                                                           const float* __restrict__ bias, float* __restrict__ out, int B,
                                                           int din, int dout, int relu) {
  pdl_prologue();
  extern __shared__ float sm[];
  float* xin = sm;                                  // [kLinRows][din]
  float* red = sm + (size_t)kLinRows * din;         // [8 warps][kLinRows][kLinCols]
  const int row0 = blockIdx.x * kLinRows, n0 = blockIdx.y * kLinCols;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kLinRows * din; i += blockDim.x) {
    const int r = i / din, k = i % din;
    xin[i] = row0 + r < B ? in[(size_t)(row0 + r) * din + k] : 0.f;
  }
  __syncthreads();
  const int kper = (din + 7) / 8;
  const int k0 = min(din, warp * kper), k1 = min(din, k0 + kper);
  float acc[kLinRows][4];
#pragma unroll
  for (int r = 0; r < kLinRows; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;
  bool ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) ok[i] = n0 + lane + 32 * i < dout;
#pragma unroll 2
  for (int k = k0; k < k1; ++k) {
    float wv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wv[i] = ok[i] ? __ldg(w_t + (size_t)k * dout + n0 + lane + 32 * i) : 0.f;
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) {
      const float xv = xin[r * din + k];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[r][i] = fmaf(xv, wv[i], acc[r][i]);
    }
  }
#pragma unroll
  for (int r = 0; r < kLinRows; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) red[((size_t)warp * kLinRows + r) * kLinCols + lane + 32 * i] = acc[r][i];
  __syncthreads();
  for (int o = tid; o < kLinRows * kLinCols; o += blockDim.x) {
    const int r = o / kLinCols, c = o % kLinCols;
    if (row0 + r < B && n0 + c < dout) {
      float s = bias[n0 + c];
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[((size_t)w * kLinRows + r) * kLinCols + c];
      out[(size_t)(row0 + r) * dout + n0 + c] = relu ? fmaxf(s, 0.f) : s;
    }
  }
}

}  // namespace dfv

using namespace dfv;

/* The hidden activations ping-pong through `scratch` (fp32, >= 2 * B * max(dims[1..n_layers-1]) floats). */
extern "C" size_t dfv_mlp_head_scratch_floats(const int* dims, int n_layers, int B) {
  int m = 1;
  for (int l = 1; l < n_layers; ++l) m = dims[l] > m ? dims[l] : m;
  return (size_t)2 * B * m;
}

extern "C" int dfv_mlp_head_fwd(const float* features, const float* const* w_t, const float* const* b, const int* dims,
                                int n_layers, float* logits, float* scratch, int B, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(features && w_t && b && dims && logits, "dfv_mlp_head_fwd: null pointer");
  DFV_REQUIRE(n_layers >= 1 && n_layers <= kHeadMaxLayers && B > 0, "dfv_mlp_head_fwd: bad n_layers %d", n_layers);
  DFV_REQUIRE(n_layers == 1 || scratch, "dfv_mlp_head_fwd: scratch missing");
  int max_hidden = 1;
  for (int l = 0; l <= n_layers; ++l) DFV_REQUIRE(dims[l] > 0, "dfv_mlp_head_fwd: bad dim");
  for (int l = 1; l < n_layers; ++l) max_hidden = dims[l] > max_hidden ? dims[l] : max_hidden;
  double wbytes = 0;
  for (int l = 0; l < n_layers; ++l) wbytes += 4.0 * dims[l] * dims[l + 1];
  ProfScope prof(PK_MLP_HEAD, wbytes + 4.0 * B * (dims[0] + dims[n_layers]), 2.0 * B * wbytes / 4.0, as_stream(stream));
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(linear_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  const float* in = features;
  for (int l = 0; l < n_layers; ++l) {
    DFV_REQUIRE(w_t[l] && b[l], "dfv_mlp_head_fwd: null weight");
    const bool last = l == n_layers - 1;
    float* out = last ? logits : scratch + (size_t)(l & 1) * B * max_hidden;
    const size_t smem = sizeof(float) * ((size_t)kLinRows * dims[l] + (size_t)8 * kLinRows * kLinCols);
    DFV_REQUIRE(smem <= 160 * 1024, "dfv_mlp_head_fwd: layer too wide (%d)", dims[l]);
    dim3 grid((unsigned)((B + kLinRows - 1) / kLinRows), (unsigned)((dims[l + 1] + kLinCols - 1) / kLinCols));
    DFV_PDL((linear_splitk_kernel), grid, 256, smem, as_stream(stream), in, w_t[l], b[l], out, B, dims[l], dims[l + 1], last ? 0 : 1);
    DFV_LAUNCH_CHECK();
    in = out;
  }
  return DFV_OK;
}
