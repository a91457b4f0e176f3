// Stem: static pad (0,1,0,1) + conv3x3 stride 2 (3 -> 48) + folded BN + swish.
// Reads the reference's NCHW fp32 input contract directly and writes NHWC.
// HBM-bound by bytes (1.7 MB in / 3.5 MB out per image at 380 in bf16) but carries
// 1296 FMA per output pixel, so it sits near the FP32-pipe / HBM crossover.
#include "common.cuh"

namespace dfv {

constexpr int kStemC = 48;

// Register blocking: each thread computes 4 adjacent output pixels x 12 channels, so every
// 16-byte weight read from shared memory feeds 16 FMAs (the 1 pixel x 48 channel version was
// bound by shared-memory weight broadcasts).  Four threads (channel quarters) share each pixel quad.
__device__ __forceinline__ void stem_store12(__nv_bfloat16* p, const float o[12]) {
  uint2* q = reinterpret_cast<uint2*>(p);   // 24 bytes, 8-byte aligned
  q[0] = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
  q[1] = make_uint2(pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
  q[2] = make_uint2(pack_bf16(o[8], o[9]), pack_bf16(o[10], o[11]));
}
__device__ __forceinline__ void stem_store12(float* p, const float o[12]) {
  float4* q = reinterpret_cast<float4*>(p);
  q[0] = make_float4(o[0], o[1], o[2], o[3]);
  q[1] = make_float4(o[4], o[5], o[6], o[7]);
  q[2] = make_float4(o[8], o[9], o[10], o[11]);
}

template <typename T, bool kFast>
__global__ void __launch_bounds__(256, 3) stem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                  const float* __restrict__ bias, T* __restrict__ y, int B, int H,
                                                  int W, int Ho, int Wo, int act) {
  __shared__ __align__(16) float ws[27 * kStemC];
  __shared__ __align__(16) float bs[kStemC];
  for (int i = threadIdx.x; i < 27 * kStemC; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < kStemC) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();

  const int quads_w = (Wo + 3) / 4;
  const long long total = (long long)B * Ho * quads_w * 4;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cq = (int)(gid & 3);                 // channel quarter: channels 12*cq .. 12*cq+11
  const long long pq = gid >> 2;
  const int wq = (int)(pq % quads_w);
  const int ho = (int)((pq / quads_w) % Ho);
  const int b = (int)(pq / ((long long)quads_w * Ho));
  const int wo0 = wq * 4;

  float acc[4][12];
#pragma unroll
  for (int pxl = 0; pxl < 4; ++pxl)
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[pxl][i] = 0.f;

  const float* xb = x + (size_t)b * 3 * H * W;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int hi = 2 * ho + kh;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      float in[9];
      const float* xr = xb + ((size_t)ci * H + hi) * W + 2 * wo0;
#pragma unroll
      for (int i = 0; i < 9; ++i) in[i] = (hi < H && 2 * wo0 + i < W) ? __ldg(xr + i) : 0.f;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4* wr = reinterpret_cast<const float4*>(ws + ((kh * 3 + kw) * 3 + ci) * kStemC + cq * 12);
        const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2];
        const float wv[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
        for (int pxl = 0; pxl < 4; ++pxl) {
          const float v = in[2 * pxl + kw];
#pragma unroll
          for (int i = 0; i < 12; ++i) acc[pxl][i] = fmaf(v, wv[i], acc[pxl][i]);
        }
      }
    }
  }
#pragma unroll
  for (int pxl = 0; pxl < 4; ++pxl) {
    if (wo0 + pxl < Wo) {
      float o[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const float v = acc[pxl][i] + bs[cq * 12 + i];
        o[i] = act ? silu<kFast>(v) : v;
      }
      stem_store12(y + (((size_t)b * Ho + ho) * Wo + wo0 + pxl) * kStemC + cq * 12, o);
    }
  }
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_stem_conv_fwd(const float* x, const float* w, const float* bias, void* y, int dtype, int B, int H,
                                 int W, int C, int act, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(x && w && bias && y, "dfv_stem_conv_fwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype), "dfv_stem_conv_fwd: bad dtype %d", dtype);
  DFV_REQUIRE(C == kStemC, "dfv_stem_conv_fwd: C must be %d (EfficientNet-B4 stem), got %d", kStemC, C);
  DFV_REQUIRE(B > 0 && H >= 3 && W >= 3, "dfv_stem_conv_fwd: bad shape B=%d H=%d W=%d", B, H, W);
  if (debug_flags() & 4) return DFV_OK;
  const int Ho = (H + 1 - 3) / 2 + 1, Wo = (W + 1 - 3) / 2 + 1;
  const long long total = (long long)B * Ho * Wo;
  const long long threads = (long long)B * Ho * ((Wo + 3) / 4) * 4;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  ProfScope prof(PK_STEM, (double)B * 3 * H * W * 4 + (double)total * kStemC * dtype_size(dtype),
                 2.0 * 27 * kStemC * (double)total, as_stream(stream));
  if (dtype == DFV_BF16)
    stem_kernel<__nv_bfloat16, true><<<grid, 256, 0, as_stream(stream)>>>(x, w, bias, (__nv_bfloat16*)y, B, H, W, Ho, Wo, act);
  else
    stem_kernel<float, false><<<grid, 256, 0, as_stream(stream)>>>(x, w, bias, (float*)y, B, H, W, Ho, Wo, act);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}
