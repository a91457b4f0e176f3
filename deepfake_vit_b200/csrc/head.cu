// Classifier head (feature_extractor.py:223-238), eval mode: a chain of small Linear layers
// with BatchNorm1d folded in and ReLU between them, run in one kernel.  Each CTA handles
// kRows feature rows so the (L2-resident) weights are re-read B/kRows times, not B times.
#include "common.cuh"

namespace dfv {

constexpr int kHeadRows = 2;
constexpr int kHeadMaxLayers = 8;

struct HeadParams {
  const float* w_t[kHeadMaxLayers];  // [din][dout]
  const float* b[kHeadMaxLayers];
  int dims[kHeadMaxLayers + 1];
  int n_layers;
  int max_dim;
};

__global__ void __launch_bounds__(256) mlp_head_kernel(const float* __restrict__ features, float* __restrict__ logits,
                                                      HeadParams hp, int B) {
  extern __shared__ float sm[];
  float* buf0 = sm;                              // [kHeadRows][max_dim]
  float* buf1 = sm + kHeadRows * hp.max_dim;
  const int row0 = blockIdx.x * kHeadRows;
  const int rows = min(kHeadRows, B - row0);
  const int d0 = hp.dims[0];
  for (int i = threadIdx.x; i < kHeadRows * d0; i += blockDim.x) {
    const int r = i / d0, k = i % d0;
    buf0[r * hp.max_dim + k] = r < rows ? features[(size_t)(row0 + r) * d0 + k] : 0.f;
  }
  __syncthreads();
  float* in = buf0;
  float* out = buf1;
  for (int l = 0; l < hp.n_layers; ++l) {
    const int din = hp.dims[l], dout = hp.dims[l + 1];
    const float* wt = hp.w_t[l];
    const bool last = l == hp.n_layers - 1;
    for (int n = threadIdx.x; n < dout; n += blockDim.x) {
      float acc[kHeadRows];
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) acc[r] = 0.f;
#pragma unroll 16
      for (int k = 0; k < din; ++k) {
        const float wv = __ldg(wt + (size_t)k * dout + n);
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r) acc[r] = fmaf(in[r * hp.max_dim + k], wv, acc[r]);
      }
      const float bv = hp.b[l][n];
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) {
        float v = acc[r] + bv;
        if (last) {
          if (r < rows) logits[(size_t)(row0 + r) * dout + n] = v;
        } else {
          out[r * hp.max_dim + n] = fmaxf(v, 0.f);
        }
      }
    }
    __syncthreads();
    float* t = in; in = out; out = t;
  }
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_mlp_head_fwd(const float* features, const float* const* w_t, const float* const* b, const int* dims,
                                int n_layers, float* logits, int B, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(features && w_t && b && dims && logits, "dfv_mlp_head_fwd: null pointer");
  DFV_REQUIRE(n_layers >= 1 && n_layers <= kHeadMaxLayers && B > 0, "dfv_mlp_head_fwd: bad n_layers %d", n_layers);
  HeadParams hp;
  hp.n_layers = n_layers;
  hp.max_dim = 0;
  for (int l = 0; l <= n_layers; ++l) {
    DFV_REQUIRE(dims[l] > 0, "dfv_mlp_head_fwd: bad dim");
    hp.dims[l] = dims[l];
    if (dims[l] > hp.max_dim) hp.max_dim = dims[l];
  }
  for (int l = 0; l < n_layers; ++l) {
    DFV_REQUIRE(w_t[l] && b[l], "dfv_mlp_head_fwd: null weight");
    hp.w_t[l] = w_t[l];
    hp.b[l] = b[l];
  }
  const size_t smem = sizeof(float) * 2 * kHeadRows * hp.max_dim;
  DFV_REQUIRE(smem <= 160 * 1024, "dfv_mlp_head_fwd: layer too wide (%d)", hp.max_dim);
  if (smem > 48 * 1024) DFV_CUDA(cudaFuncSetAttribute(mlp_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  double wbytes = 0;
  for (int l = 0; l < n_layers; ++l) wbytes += 4.0 * dims[l] * dims[l + 1];
  ProfScope prof(PK_MLP_HEAD, wbytes + 4.0 * B * (dims[0] + dims[n_layers]), 2.0 * B * wbytes / 4.0, as_stream(stream));
  mlp_head_kernel<<<(B + kHeadRows - 1) / kHeadRows, 256, smem, as_stream(stream)>>>(features, logits, hp, B);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}
