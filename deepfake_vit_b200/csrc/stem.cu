// Stem: static pad (0,1,0,1) + conv3x3 stride 2 (3 -> 48) + folded BN + swish.
// Reads the reference's NCHW fp32 input contract directly and writes NHWC.
// HBM-bound by bytes (1.7 MB in / 3.5 MB out per image at 380 in bf16) but carries
// 1296 FMA per output pixel, so it sits near the FP32-pipe / HBM crossover.
#include "common.cuh"

namespace dfv {

constexpr int kStemC = 48;

template <typename T, bool kFast>
__global__ void __launch_bounds__(128) stem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                  const float* __restrict__ bias, T* __restrict__ y, int B, int H,
                                                  int W, int Ho, int Wo) {
  __shared__ __align__(16) float ws[27 * kStemC];
  __shared__ __align__(16) float bs[kStemC];
  for (int i = threadIdx.x; i < 27 * kStemC; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < kStemC) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();

  const long long total = (long long)B * Ho * Wo;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int wo = (int)(p % Wo);
  const int ho = (int)((p / Wo) % Ho);
  const int b = (int)(p / ((long long)Wo * Ho));

  float acc[kStemC];
#pragma unroll
  for (int i = 0; i < kStemC; ++i) acc[i] = 0.f;

  const float* xb = x + (size_t)b * 3 * H * W;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int hi = 2 * ho + kh;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int wi = 2 * wo + kw;
      const bool ok = (hi < H) && (wi < W);
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float v = ok ? __ldg(xb + ((size_t)ci * H + hi) * W + wi) : 0.f;
        const float4* wr = reinterpret_cast<const float4*>(ws + ((kh * 3 + kw) * 3 + ci) * kStemC);
#pragma unroll
        for (int q = 0; q < kStemC / 4; ++q) {
          const float4 w4 = wr[q];
          acc[4 * q + 0] = fmaf(v, w4.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
        }
      }
    }
  }
  T* yo = y + (size_t)p * kStemC;
#pragma unroll
  for (int q = 0; q < kStemC / 8; ++q) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = silu<kFast>(acc[8 * q + j] + bs[8 * q + j]);
    store8(yo + 8 * q, o);
  }
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_stem_conv_fwd(const float* x, const float* w, const float* bias, void* y, int dtype, int B, int H,
                                 int W, int C, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(x && w && bias && y, "dfv_stem_conv_fwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype), "dfv_stem_conv_fwd: bad dtype %d", dtype);
  DFV_REQUIRE(C == kStemC, "dfv_stem_conv_fwd: C must be %d (EfficientNet-B4 stem), got %d", kStemC, C);
  DFV_REQUIRE(B > 0 && H >= 3 && W >= 3, "dfv_stem_conv_fwd: bad shape B=%d H=%d W=%d", B, H, W);
  const int Ho = (H + 1 - 3) / 2 + 1, Wo = (W + 1 - 3) / 2 + 1;
  const long long total = (long long)B * Ho * Wo;
  const unsigned grid = (unsigned)((total + 127) / 128);
  ProfScope prof(PK_STEM, (double)B * 3 * H * W * 4 + (double)total * kStemC * dtype_size(dtype),
                 2.0 * 27 * kStemC * (double)total, as_stream(stream));
  if (dtype == DFV_BF16)
    stem_kernel<__nv_bfloat16, true><<<grid, 128, 0, as_stream(stream)>>>(x, w, bias, (__nv_bfloat16*)y, B, H, W, Ho, Wo);
  else
    stem_kernel<float, false><<<grid, 128, 0, as_stream(stream)>>>(x, w, bias, (float*)y, B, H, W, Ho, Wo);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}
