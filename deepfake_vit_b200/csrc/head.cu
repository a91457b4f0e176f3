// Classifier head (feature_extractor.py:223-238), eval mode: a chain of small Linear layers with BatchNorm1d folded in
// and ReLU between them.  Each layer is the weight-stationary small-linear kernel of small_linear.cuh over K slices of
// 256 inputs (64 outputs x 16 rows per CTA, its weight slab and inputs requested up front into shared memory), and when
// the layer has several K slices a combine kernel adds the partial sums in fixed order (+bias, ReLU).  The previous
// split-K-across-warps kernel kept 8 weight loads in flight per thread: 112 dependent batches against HBM-cold weights,
// 120 us for 0.5 GFLOP.
#include "small_linear.cuh"

namespace dfv {

constexpr int kHeadMaxLayers = 8;
constexpr int kHeadKc = 256;     // inputs per K slice

static int head_ksplit(int din) { return (din + kHeadKc - 1) / kHeadKc; }

}  // namespace dfv

using namespace dfv;

/* scratch: two ping-pong activation buffers [B][max hidden] plus the K-slice partial sums of the widest layer. */
extern "C" size_t dfv_mlp_head_scratch_floats(const int* dims, int n_layers, int B) {
  size_t m = 1, part = 0;
  for (int l = 1; l < n_layers; ++l) m = (size_t)dims[l] > m ? (size_t)dims[l] : m;
  for (int l = 0; l < n_layers; ++l) {
    const int ks = head_ksplit(dims[l]);
    if (ks > 1) part = std::max(part, (size_t)ks * B * dims[l + 1]);
  }
  return (size_t)2 * B * m + part;
}

extern "C" int dfv_mlp_head_fwd(const float* features, const float* const* w_t, const float* const* b, const int* dims,
                                int n_layers, float* logits, float* scratch, int B, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(features && w_t && b && dims && logits, "dfv_mlp_head_fwd: null pointer");
  DFV_REQUIRE(n_layers >= 1 && n_layers <= kHeadMaxLayers && B > 0, "dfv_mlp_head_fwd: bad n_layers %d", n_layers);
  size_t max_hidden = 1;
  for (int l = 0; l <= n_layers; ++l) DFV_REQUIRE(dims[l] > 0, "dfv_mlp_head_fwd: bad dim");
  for (int l = 1; l < n_layers; ++l) max_hidden = (size_t)dims[l] > max_hidden ? (size_t)dims[l] : max_hidden;
  bool need_scratch = n_layers > 1;
  for (int l = 0; l < n_layers; ++l) need_scratch = need_scratch || head_ksplit(dims[l]) > 1;
  DFV_REQUIRE(!need_scratch || scratch, "dfv_mlp_head_fwd: scratch missing");
  double wbytes = 0;
  for (int l = 0; l < n_layers; ++l) wbytes += 4.0 * dims[l] * dims[l + 1];
  cudaStream_t st = as_stream(stream);
  ProfScope prof(PK_MLP_HEAD, wbytes + 4.0 * B * (dims[0] + dims[n_layers]), 2.0 * B * wbytes / 4.0, st);
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(sl_kmajor_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  float* part = scratch ? scratch + 2 * (size_t)B * max_hidden : nullptr;
  const float* in = features;
  for (int l = 0; l < n_layers; ++l) {
    DFV_REQUIRE(w_t[l] && b[l], "dfv_mlp_head_fwd: null weight");
    const bool last = l == n_layers - 1;
    const int din = dims[l], dout = dims[l + 1], ks = head_ksplit(din);
    float* out = last ? logits : scratch + (size_t)(l & 1) * B * max_hidden;
    const int kc = ks > 1 ? kHeadKc : din;
    const dim3 grid((unsigned)((dout + kSlCols - 1) / kSlCols), (unsigned)((B + kSlRows - 1) / kSlRows), (unsigned)ks);
    // one K slice: bias and ReLU in the kernel; several: raw partial sums, finished by the combine pass
    DFV_PDL((sl_kmajor_kernel<float, false>), grid, kSlThreads, sl_kmajor_smem(kc), st, in, w_t[l], b[l], out, (float*)nullptr, part, B, dout,
            din, kc, 0, last ? 0 : SL_OUT_RELU, (const long long*)nullptr, (const float*)nullptr, 1.0f);
    DFV_LAUNCH_CHECK();
    if (ks > 1) {
      DFV_PDL(sl_combine_kernel, (unsigned)(((size_t)B * dout + kSlThreads - 1) / kSlThreads), kSlThreads, 0, st, (const float*)part,
              (const float*)nullptr, ks, b[l], out, B, dout, last ? 0 : 1);
      DFV_LAUNCH_CHECK();
    }
    in = out;
  }
  return DFV_OK;
}

/* One fp32 Linear layer with few rows (the classifier inside the training step, forward and input gradient):
 *   out[b][n] = bias[n] + sum_k in[b][k] * W,   W = w[n][k] (torch layout, w_kmajor = 0) or w[k][n] (w_kmajor = 1, i.e. the
 *   SAME torch tensor read as the transposed layer: the input gradient needs no transposed weight copy).
 * The small-linear kernels over K slices of 256 (+ a fixed-order combine): ~230 CTAs for 1792 -> 512 at batch 64 instead of
 * the 8 CTAs of the tiled SIMT GEMM walking K in 112 dependent steps (202 us).  scratch: dfv_linear_f32_scratch_floats(). */
extern "C" size_t dfv_linear_f32_scratch_floats(int B, int K, int N) {
  const int ks = head_ksplit(K);
  return ks > 1 ? (size_t)ks * B * N : 0;
}

extern "C" int dfv_linear_f32_fwd(const float* in, const float* w, const float* bias, float* out, float* scratch, size_t scratch_floats,
                                  int B, int K, int N, int w_kmajor, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(in && w && out && B > 0 && K > 0 && N > 0, "dfv_linear_f32_fwd: bad arguments");
  const int ks = head_ksplit(K);
  const int kc = ks > 1 ? kHeadKc : K;
  DFV_REQUIRE(ks == 1 || (scratch && scratch_floats >= (size_t)ks * B * N), "dfv_linear_f32_fwd: scratch too small");
  cudaStream_t st = as_stream(stream);
  ProfScope prof(PK_MLP_HEAD, 4.0 * ((double)K * N + (double)B * (K + N)), 2.0 * (double)B * K * N, st);
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(sl_kmajor_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    DFV_CUDA(cudaFuncSetAttribute(sl_kmajor_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  const dim3 grid((unsigned)((N + kSlCols - 1) / kSlCols), (unsigned)((B + kSlRows - 1) / kSlRows), (unsigned)ks);
  if (w_kmajor)
    DFV_PDL((sl_kmajor_kernel<float, false>), grid, kSlThreads, sl_kmajor_smem(kc), st, in, w, bias, out, (float*)nullptr, scratch, B, N, K, kc, 0,
            0, (const long long*)nullptr, (const float*)nullptr, 1.0f);
  else
    DFV_PDL((sl_kmajor_kernel<float, true>), grid, kSlThreads, sl_kmajor_smem(kc), st, in, w, bias, out, out, scratch, B, N, K, kc, 0, 0,
            (const long long*)nullptr, (const float*)nullptr, 1.0f);
  DFV_LAUNCH_CHECK();
  if (ks > 1) {
    DFV_PDL(sl_combine_kernel, (unsigned)(((size_t)B * N + kSlThreads - 1) / kSlThreads), kSlThreads, 0, st, (const float*)scratch,
            (const float*)nullptr, ks, bias, out, B, N, 0);
    DFV_LAUNCH_CHECK();
  }
  return DFV_OK;
}
