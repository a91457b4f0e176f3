// Backward kernels of the convolutions (training path).
//   pw_wgrad      dW[N][K] = sum_m g[m][n] * a[m][k] (* gate[image(m)][k])      1x1 conv weight gradient, split over M
//   dw_dgrad      depthwise input gradient (gather form; any stride).  Stride-1 layers instead reuse the
//                 forward TMA kernel with flipped taps (train.cu); this kernel serves the four stride-2 layers.
//   dw_wgrad      depthwise weight gradient  dw[kh][kw][c] = sum g[b,ho,wo,c] * x[b,ho*S-p+kh,wo*S-p+kw,c]
//   stem_wgrad    stem 3x3/s2 weight gradient from the NCHW fp32 images
// The 1x1 data gradients are plain GEMMs on a transposed weight copy and use dfv_pw_gemm_fwd (tcgen05).
#include <algorithm>

#include "common.cuh"

namespace dfv {

int launch_wgrad_tc(const void* g, const void* a, const void* a_scale, int rows_per_image, float* dw, long long M, int K,
                    int N, cudaStream_t st);   // wgrad.cu

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// ------------------------------------------------------------------------------------ pw_wgrad (SIMT, fp32 accumulate)
template <typename T>
__global__ void __launch_bounds__(256) pw_wgrad_simt_kernel(const T* __restrict__ g, const T* __restrict__ a,
                                                           const T* __restrict__ a_scale, int rows_per_image,
                                                           float* __restrict__ dw, long long M, int K, int N,
                                                           long long rows_per_split) {
  constexpr int TN = 64, TK = 64, TM = 16;
  __shared__ __align__(16) float Gs[TM][TN + 4];
  __shared__ __align__(16) float As[TM][TK + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * TN, k0 = blockIdx.y * TK;
  const long long m_begin = (long long)blockIdx.z * rows_per_split;
  const long long m_end = min(m_begin + rows_per_split, M);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long m0 = m_begin; m0 < m_end; m0 += TM) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int r = idx >> 6, c = idx & 63;
      const long long m = m0 + r;
      float gv = 0.f, av = 0.f;
      if (m < m_end) {
        if (n0 + c < N) gv = ldf(g + (size_t)m * N + n0 + c);
        if (k0 + c < K) {
          av = ldf(a + (size_t)m * K + k0 + c);
          if (a_scale) {
            av *= ldf(a_scale + (size_t)(m / rows_per_image) * K + k0 + c);
            if constexpr (sizeof(T) == 2) av = __bfloat162float(__float2bfloat16_rn(av));   // as the forward operand was rounded
          }
        }
      }
      Gs[r][c] = gv;
      As[r][c] = av;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < TM; ++mm) {
      const float4 g4 = *reinterpret_cast<const float4*>(&Gs[mm][ty * 4]);
      const float4 a4 = *reinterpret_cast<const float4*>(&As[mm][tx * 4]);
      const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], av[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(dw + (size_t)n * K + k, acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------------------------ dw_dgrad (gather)
template <typename T>
__global__ void __launch_bounds__(256) dw_dgrad_kernel(const T* __restrict__ g, const float* __restrict__ w,
                                                      T* __restrict__ dx, int B, int H, int W, int C, int Ho, int Wo,
                                                      int K, int S, int pad) {
  const int CV = C >> 3;
  const long long total = (long long)B * H * W * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long r = i / CV;
    const int wi = (int)(r % W);
    r /= W;
    const int hi = (int)(r % H);
    const int b = (int)(r / H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int kh = 0; kh < K; ++kh) {
      const int th = hi + pad - kh;
      if (th < 0 || th % S) continue;
      const int ho = th / S;
      if (ho >= Ho) continue;
      for (int kw = 0; kw < K; ++kw) {
        const int tw = wi + pad - kw;
        if (tw < 0 || tw % S) continue;
        const int wo = tw / S;
        if (wo >= Wo) continue;
        float gv[8], wv[8];
        load8(g + (((size_t)b * Ho + ho) * Wo + wo) * C + cv * 8, gv);
        load8(w + (size_t)(kh * K + kw) * C + cv * 8, wv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(gv[e], wv[e], acc[e]);
      }
    }
    store8(dx + (size_t)i * 8, acc);
  }
}

// ------------------------------------------------------------------------------------ dw_wgrad
// block = 8 channel groups x K kernel rows x LANES pixel lanes; every thread keeps K x 8 accumulators.
template <typename T, int K>
__global__ void __launch_bounds__(256) dw_wgrad_kernel(const T* __restrict__ g, const T* __restrict__ x,
                                                      float* __restrict__ dw, int B, int H, int W, int C, int Ho, int Wo,
                                                      int S, int pad, long long pix_per_cta) {
  constexpr int G = 8, LANES = 256 / (G * K);
  __shared__ float sm[LANES * K * G * K * 8];
  const int tid = threadIdx.x;
  const int gi = tid % G, kh = (tid / G) % K, lane = tid / (G * K);
  const int c = blockIdx.x * (G * 8) + gi * 8;
  const long long npix = (long long)B * Ho * Wo;
  const long long p0 = (long long)blockIdx.y * pix_per_cta, p1 = min(p0 + pix_per_cta, npix);
  float acc[K][8];
#pragma unroll
  for (int kw = 0; kw < K; ++kw)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[kw][e] = 0.f;
  if (lane < LANES && c < C) {
    for (long long p = p0 + lane; p < p1; p += LANES) {
      const int wo = (int)(p % Wo);
      const long long r = p / Wo;
      const int ho = (int)(r % Ho), b = (int)(r / Ho);
      const int hi = ho * S - pad + kh;
      if (hi < 0 || hi >= H) continue;
      float gv[8];
      load8(g + (size_t)p * C + c, gv);
      const T* xr = x + (((size_t)b * H + hi) * W) * C + c;
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        const int wi = wo * S - pad + kw;
        if (wi < 0 || wi >= W) continue;
        float xv[8];
        load8(xr + (size_t)wi * C, xv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[kw][e] = fmaf(gv[e], xv[e], acc[kw][e]);
      }
    }
  }
  if (lane < LANES) {
    float* s = sm + ((size_t)(lane * K + kh) * G + gi) * K * 8;
#pragma unroll
    for (int kw = 0; kw < K; ++kw)
#pragma unroll
      for (int e = 0; e < 8; ++e) s[kw * 8 + e] = acc[kw][e];
  }
  __syncthreads();
  // K * G * K * 8 outputs per CTA
  for (int o = tid; o < K * G * K * 8; o += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < LANES; ++l) s += sm[(size_t)l * K * G * K * 8 + o];
    const int e = o % 8, kw = (o / 8) % K, g2 = (o / (8 * K)) % G, kh2 = o / (8 * K * G);
    const int cc = blockIdx.x * (G * 8) + g2 * 8 + e;
    if (cc < C) atomicAdd(dw + (size_t)(kh2 * K + kw) * C + cc, s);
  }
}

// ------------------------------------------------------------------------------------ stem_wgrad
// thread = (tap in 0..26, channel group of 8 in 0..5); CTA loops over a range of output rows (b, ho).
template <typename T>
__global__ void __launch_bounds__(192) stem_wgrad_kernel(const T* __restrict__ g, const float* __restrict__ x,
                                                        float* __restrict__ dw, int B, int H, int W, int Ho, int Wo,
                                                        long long rows_per_cta) {
  constexpr int CO = 48;
  const int tid = threadIdx.x;
  const bool active = tid < 27 * 6;
  const int tap = tid / 6, cg = tid % 6;            // tap = ci * 9 + kh * 3 + kw  (torch weight layout [co][ci][kh][kw])
  const int ci = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
  const long long nrows = (long long)B * Ho;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, nrows);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (active) {
    for (long long r = r0; r < r1; ++r) {
      const int ho = (int)(r % Ho), b = (int)(r / Ho);
      const int hi = 2 * ho + kh;
      if (hi >= H) continue;
      const float* xr = x + (((size_t)b * 3 + ci) * H + hi) * W;
      const T* gr = g + (size_t)r * Wo * CO + cg * 8;
      for (int wo = 0; wo < Wo; ++wo) {
        const int wi = 2 * wo + kw;
        if (wi >= W) continue;
        const float xv = __ldg(xr + wi);
        float gv[8];
        load8(gr + (size_t)wo * CO, gv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(xv, gv[e], acc[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(dw + (size_t)(cg * 8 + e) * 27 + tap, acc[e]);
  }
}

}  // namespace dfv

using namespace dfv;

extern "C" {

/* dW[N][K] += sum over rows of g[m][n] * a[m][k] * a_scale[m / rows_per_image][k]; dw is fp32 in torch layout
 * ([out][in] of the 1x1 conv) and must be zeroed (or hold the value to accumulate onto) by the caller. */
int dfv_pw_wgrad(const void* g, const void* a, const void* a_scale, int rows_per_image, float* dw, int dtype, long long M,
                 int K, int N, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(g && a && dw, "dfv_pw_wgrad: null pointer");
  DFV_REQUIRE(valid_dtype(dtype) && M > 0 && K > 0 && N > 0, "dfv_pw_wgrad: bad shape");
  DFV_REQUIRE(!a_scale || (rows_per_image > 0 && M % rows_per_image == 0), "dfv_pw_wgrad: a_scale needs rows_per_image dividing M");
  cudaStream_t st = as_stream(stream);
  const int nt = (N + 63) / 64, kt = (K + 63) / 64;
  long long splits = (4LL * num_sms() + (long long)nt * kt - 1) / ((long long)nt * kt);
  if (splits < 1) splits = 1;
  long long rps = (M + splits - 1) / splits;
  rps = (rps + 15) / 16 * 16;
  splits = (M + rps - 1) / rps;
  if (splits > 65535) {
    splits = 65535;
    rps = ((M + splits - 1) / splits + 15) / 16 * 16;
    splits = (M + rps - 1) / rps;
  }
  dim3 grid((unsigned)nt, (unsigned)kt, (unsigned)splits);
  const double es = (double)dtype_size(dtype);
  ProfScope prof(PK_WGRAD, es * ((double)M * K + (double)M * N) + 4.0 * N * K, 2.0 * (double)M * K * N, st);
  if (dtype == DFV_BF16 && K % 8 == 0 && N % 8 == 0 && !force_simt_gemm())
    return launch_wgrad_tc(g, a, a_scale, rows_per_image, dw, M, K, N, st);
  if (dtype == DFV_BF16)
    pw_wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)g, (const __nv_bfloat16*)a, (const __nv_bfloat16*)a_scale,
                                                             rows_per_image > 0 ? rows_per_image : 1, dw, M, K, N, rps);
  else
    pw_wgrad_simt_kernel<float><<<grid, 256, 0, st>>>((const float*)g, (const float*)a, (const float*)a_scale,
                                                     rows_per_image > 0 ? rows_per_image : 1, dw, M, K, N, rps);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

/* Depthwise input gradient.  g: [B][Ho][Wo][C], w: fp32 [k*k][C] (NOT flipped), dx: [B][H][W][C]. */
int dfv_dwconv_dgrad(const void* g, const float* w_kkc, void* dx, int dtype, int B, int H, int W, int C, int kernel,
                     int stride, int pad_lo, int pad_hi, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(g && w_kkc && dx, "dfv_dwconv_dgrad: null pointer");
  DFV_REQUIRE(valid_dtype(dtype) && B > 0 && C > 0 && C % 8 == 0, "dfv_dwconv_dgrad: bad shape");
  const int Ho = (H + pad_lo + pad_hi - kernel) / stride + 1, Wo = (W + pad_lo + pad_hi - kernel) / stride + 1;
  DFV_REQUIRE(Ho > 0 && Wo > 0, "dfv_dwconv_dgrad: bad shape");
  cudaStream_t st = as_stream(stream);
  const long long total = (long long)B * H * W * (C / 8);
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 16LL * num_sms());
  const double es = (double)dtype_size(dtype);
  ProfScope prof(PK_DWCONV_BWD, es * ((double)B * H * W * C + (double)B * Ho * Wo * C), 2.0 * kernel * kernel * (double)B * Ho * Wo * C, st);
  if (dtype == DFV_BF16)
    dw_dgrad_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)g, w_kkc, (__nv_bfloat16*)dx, B, H, W, C, Ho, Wo, kernel, stride, pad_lo);
  else
    dw_dgrad_kernel<float><<<blocks, 256, 0, st>>>((const float*)g, w_kkc, (float*)dx, B, H, W, C, Ho, Wo, kernel, stride, pad_lo);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

/* Depthwise weight gradient, accumulated (atomically) into dw_kkc fp32 [k*k][C], which the caller zeroes. */
int dfv_dwconv_wgrad(const void* g, const void* x, float* dw_kkc, int dtype, int B, int H, int W, int C, int kernel,
                     int stride, int pad_lo, int pad_hi, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(g && x && dw_kkc, "dfv_dwconv_wgrad: null pointer");
  DFV_REQUIRE(valid_dtype(dtype) && B > 0 && C > 0 && C % 8 == 0 && (kernel == 3 || kernel == 5), "dfv_dwconv_wgrad: bad shape");
  const int Ho = (H + pad_lo + pad_hi - kernel) / stride + 1, Wo = (W + pad_lo + pad_hi - kernel) / stride + 1;
  DFV_REQUIRE(Ho > 0 && Wo > 0, "dfv_dwconv_wgrad: bad shape");
  cudaStream_t st = as_stream(stream);
  const int chunks = (C + 63) / 64;
  const long long npix = (long long)B * Ho * Wo;
  long long splits = std::max<long long>(1, 4LL * num_sms() / chunks);
  splits = std::min<long long>(splits, std::min<long long>(npix, 65535));
  const long long ppc = (npix + splits - 1) / splits;
  dim3 grid((unsigned)chunks, (unsigned)((npix + ppc - 1) / ppc));
  const double es = (double)dtype_size(dtype);
  ProfScope prof(PK_DWCONV_BWD, es * ((double)B * H * W * C + (double)B * Ho * Wo * C), 2.0 * kernel * kernel * (double)B * Ho * Wo * C, st);
#define DWW(T_, K_)                                                                                                       \
  dw_wgrad_kernel<T_, K_><<<grid, 256, 0, st>>>((const T_*)g, (const T_*)x, dw_kkc, B, H, W, C, Ho, Wo, stride, pad_lo, ppc)
  if (dtype == DFV_BF16) {
    if (kernel == 3) DWW(__nv_bfloat16, 3); else DWW(__nv_bfloat16, 5);
  } else {
    if (kernel == 3) DWW(float, 3); else DWW(float, 5);
  }
#undef DWW
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

/* Stem weight gradient, accumulated into dw fp32 [48][3][3][3] (torch layout), which the caller zeroes.
 * g: [B][Ho][Wo][48] gradient wrt the raw (pre-BN) stem output; x: the NCHW fp32 images. */
int dfv_stem_wgrad(const void* g, const float* x_nchw, float* dw, int dtype, int B, int H, int W, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(g && x_nchw && dw && valid_dtype(dtype) && B > 0 && H >= 3 && W >= 3, "dfv_stem_wgrad: bad arguments");
  const int Ho = (H + 1 - 3) / 2 + 1, Wo = (W + 1 - 3) / 2 + 1;
  const long long nrows = (long long)B * Ho;
  const long long ctas = std::min<long long>(nrows, 8LL * num_sms());
  const long long rpc = (nrows + ctas - 1) / ctas;
  cudaStream_t st = as_stream(stream);
  ProfScope prof(PK_STEM, (double)B * 3 * H * W * 4 + (double)nrows * Wo * 48 * dtype_size(dtype), 2.0 * 27 * 48 * (double)nrows * Wo, st);
  if (dtype == DFV_BF16)
    stem_wgrad_kernel<__nv_bfloat16><<<(unsigned)((nrows + rpc - 1) / rpc), 192, 0, st>>>((const __nv_bfloat16*)g, x_nchw, dw, B, H, W, Ho, Wo, rpc);
  else
    stem_wgrad_kernel<float><<<(unsigned)((nrows + rpc - 1) / rpc), 192, 0, st>>>((const float*)g, x_nchw, dw, B, H, W, Ho, Wo, rpc);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

}  // extern "C"
