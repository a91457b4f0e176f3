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

// ------------------------------------------------------------------------------------ dw_dgrad, stride 2 (TMA-tiled)
// dx[b][hi][wi][c] = sum over (kh, kw) with (hi + pad - kh) and (wi + pad - kw) even of
//                    g[b][(hi + pad - kh) / 2][(wi + pad - kw) / 2][c] * w[kh][kw][c]
// Persistent CTAs over contiguous (image, tile) ranges of one channel chunk; the g tile that a TH x TW tile of dx
// needs ((TH + K - 1) / 2 + 1 rows) arrives by a double-buffered 4-D TMA load whose out-of-bounds zero fill supplies
// the "no contribution" border.  thread = (8-channel group, strip of 4 dx pixels, row); only the parity-matching
// taps are visited.  Replaces the per-pixel gather kernel for the four stride-2 layers.
struct DwDgParams {
  int C, CB, G;
  int H, W, Ho, Wo;
  int TH, TW, THg, TWg;
  int tiles_w, tiles_h;
  int chunks, ctas_per_chunk;
  long long per_chunk, tiles_per_cta;
  int pad;
};

// One dx row of a thread's strip (4 pixels x 8 channels) with the tap parities known at COMPILE time: KH0 = (hi + pad) & 1 picks
// the kernel rows, P = pad & 1 the kernel columns of every pixel (the strip starts at a multiple of 4).  Per kernel row the
// (at most 5) g columns the strip touches are loaded ONCE into registers and every weight vector once; the first version
// computed each tap's tile address and loaded g and w per (pixel, tap): ~27 instructions per 8 FMAs, 308 us for the 144-channel
// 190 -> 95 layer at batch 64.
template <typename T, int K, int KH0, int P>
__device__ __forceinline__ void dw_dgrad_s2_row(const T* __restrict__ gt, const float* __restrict__ wrow0, int TWg, int CB, int ho_base,
                                                int col_base, float (&acc)[4][8]) {
  constexpr int OFF_LO = -((K - 1) / 2), NOFF = 2 - OFF_LO + 1;      // column offsets (l + P - kw) / 2, l < 4, kw < K
#pragma unroll
  for (int t = 0; t < (K + 1) / 2; ++t) {
    constexpr int dummy = 0; (void)dummy;
    const int kh = KH0 + 2 * t;
    if (kh >= K) break;
    // tile-local g row of this kernel row: ((hi + pad - kh) >> 1) - glo_h = ho_base - t
    const T* grow = gt + ((size_t)(ho_base - t) * TWg + col_base) * CB;
    float gv[NOFF][8];
#pragma unroll
    for (int o = 0; o < NOFF; ++o) {
      bool used = false;
#pragma unroll
      for (int l = 0; l < 4; ++l)
#pragma unroll
        for (int kw = 0; kw < K; ++kw) used = used || (((l + P - kw) & 1) == 0 && (l + P - kw) / 2 == OFF_LO + o && (l + P - kw) >= 2 * OFF_LO);
      if (used) load8(grow + (OFF_LO + o) * CB, gv[o]);
    }
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      float wv[8];
      load8(wrow0 + (size_t)(kh * K + kw) * CB, wv);
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        if (((l + P - kw) & 1) != 0) continue;
        constexpr int unused = 0; (void)unused;
        const int o = (l + P - kw + 2 * (K - 1)) / 2 - (K - 1) - OFF_LO;      // (l + P - kw) / 2 - OFF_LO without dividing a negative
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[l][e] = fmaf(gv[o][e], wv[e], acc[l][e]);
      }
    }
  }
}

template <typename T, int K>
__global__ void __launch_bounds__(256, 2) dw_dgrad_s2_kernel(const __grid_constant__ CUtensorMap tm_g, const float* __restrict__ w,
                                                            T* __restrict__ dx, DwDgParams p) {
  constexpr int L = 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const size_t g_bytes = (size_t)p.THg * p.TWg * p.CB * sizeof(T);
  const size_t g_stride = (g_bytes + 127) / 128 * 128;
  float* wsm = reinterpret_cast<float*>(smem_raw + 2 * g_stride);            // [K*K][CB] fp32
  uint64_t* mbar = reinterpret_cast<uint64_t*>(wsm + (size_t)K * K * p.CB);

  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % p.chunks, slot = blockIdx.x / p.chunks;
  const int c0 = chunk * p.CB;
  const int n_tiles = p.tiles_w * p.tiles_h;
  const long long t_begin = (long long)slot * p.tiles_per_cta;
  const long long t_end = min(t_begin + p.tiles_per_cta, p.per_chunk);

  // first g row / column a tile starting at dx row h0 / column w0 can touch (floor division, may be negative)
  auto g_lo = [&](int x0) { const int a = x0 + p.pad - (K - 1); return a >= 0 ? a / 2 : -((1 - a) / 2); };
  auto issue = [&](long long t, int buf) {
    const int tw = (int)(t % p.tiles_w), th = (int)((t / p.tiles_w) % p.tiles_h), b = (int)(t / n_tiles);
    mbar_expect_tx(&mbar[buf], (uint32_t)g_bytes);
    tma_load_4d(smem_raw + buf * g_stride, &tm_g, &mbar[buf], c0, g_lo(tw * p.TW), g_lo(th * p.TH), b);
  };
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
    if (t_begin < t_end) issue(t_begin, 0);
  }
  for (int i = tid; i < K * K * p.CB; i += blockDim.x) {
    const int c = c0 + i % p.CB;
    wsm[i] = c < p.C ? w[(size_t)(i / p.CB) * p.C + c] : 0.f;
  }
  __syncthreads();

  const int G = p.G, strips = p.TW / L;
  const int g = tid % G, j = (tid / G) % strips, r = tid / (G * strips);
  const bool active = r < p.TH && c0 + g * 8 < p.C;

  int it = 0;
  for (long long t = t_begin; t < t_end; ++t, ++it) {
    const int buf = it & 1;
    if (tid == 0 && t + 1 < t_end) issue(t + 1, buf ^ 1);
    const int tw = (int)(t % p.tiles_w), th = (int)((t / p.tiles_w) % p.tiles_h), b = (int)(t / n_tiles);
    const int h0 = th * p.TH, w0 = tw * p.TW;
    const int hi = h0 + r, wi0 = w0 + j * L;
    mbar_wait(&mbar[buf], (it >> 1) & 1, 22);
    if (active && hi < p.H && wi0 < p.W) {
      const T* gt = reinterpret_cast<const T*>(smem_raw + buf * g_stride) + g * 8;
      const int glo_h = g_lo(h0), glo_w = g_lo(w0);
      float acc[L][8];
#pragma unroll
      for (int l = 0; l < L; ++l)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[l][e] = 0.f;
      // taps with (hi + pad - kh) even and (wi + pad - kw) even: both parities resolved at compile time (dw_dgrad_s2_row)
      const int kh0 = (hi + p.pad) & 1, pw = p.pad & 1;
      const int ho_base = ((hi + p.pad - kh0) >> 1) - glo_h;          // tile-local g row of kernel row kh0
      const int col_base = ((wi0 + p.pad - pw) >> 1) - glo_w;         // tile-local g column of column offset 0 (wi0 is a multiple of 4)
      const float* wrow0 = wsm + g * 8;
      if (kh0 == 0) {
        if (pw == 0) dw_dgrad_s2_row<T, K, 0, 0>(gt, wrow0, p.TWg, p.CB, ho_base, col_base, acc);
        else dw_dgrad_s2_row<T, K, 0, 1>(gt, wrow0, p.TWg, p.CB, ho_base, col_base, acc);
      } else {
        if (pw == 0) dw_dgrad_s2_row<T, K, 1, 0>(gt, wrow0, p.TWg, p.CB, ho_base, col_base, acc);
        else dw_dgrad_s2_row<T, K, 1, 1>(gt, wrow0, p.TWg, p.CB, ho_base, col_base, acc);
      }
      T* out = dx + (((size_t)b * p.H + hi) * p.W + wi0) * p.C + c0 + g * 8;
#pragma unroll
      for (int l = 0; l < L; ++l)
        if (wi0 + l < p.W) store8(out + (size_t)l * p.C, acc[l]);
    }
    __syncthreads();   // everyone is done with tile[buf] before it is refilled
  }
}

// ------------------------------------------------------------------------------------ dw_wgrad
// dw[kh][kw][c] = sum over (b, ho, wo) of g[b,ho,wo,c] * x[b, ho*S - pad + kh, wo*S - pad + kw, c]
// Persistent CTAs over (image, tile) for ONE channel chunk, 2-deep TMA pipeline of an x tile (with halo;
// out-of-bounds zero fill is the static pad) and the matching g tile (zero fill beyond Ho / Wo removes
// every bounds check).  thread = (8-channel group, kernel row kh, tile row, strip): it slides a K-wide
// register window along its strip and keeps K x 8 fp32 accumulators for the whole kernel; one shared-memory
// reduction and one set of atomic adds per CTA at the end.
struct DwWgParams {
  int C, CB, G;
  int TH, TW, THI, TWI, SL, NS;
  int tiles_w, tiles_h;
  int chunks, ctas_per_chunk;
  long long per_chunk;
  int pad;
  int units;     // TH * NS
};

__device__ __forceinline__ void wg_fma8(const uint4& x, const uint4& g, float acc[8]) {
  auto pair = [](uint32_t a, uint32_t b, float& a0, float& a1) {
    asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\t"
        "mov.b32 {xl, xh}, %2;\n\t"
        "mov.b32 {wl, wh}, %3;\n\t"
        "fma.rn.f32.bf16 %0, xl, wl, %0;\n\t"
        "fma.rn.f32.bf16 %1, xh, wh, %1;\n\t}"
        : "+f"(a0), "+f"(a1)
        : "r"(a), "r"(b));
  };
  pair(x.x, g.x, acc[0], acc[1]);
  pair(x.y, g.y, acc[2], acc[3]);
  pair(x.z, g.z, acc[4], acc[5]);
  pair(x.w, g.w, acc[6], acc[7]);
}
struct F8 { float v[8]; };
__device__ __forceinline__ void wg_fma8(const F8& x, const F8& g, float acc[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = fmaf(x.v[e], g.v[e], acc[e]);
}
__device__ __forceinline__ void wg_ld(const __nv_bfloat16* p, uint4& o) { o = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void wg_ld(const float* p, F8& o) { load8(p, o.v); }
template <typename T> struct WgVec;
template <> struct WgVec<__nv_bfloat16> { using type = uint4; };
template <> struct WgVec<float> { using type = F8; };

template <typename T, int K, int S>
__global__ void __launch_bounds__(256, 2) dw_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                             const __grid_constant__ CUtensorMap tm_g,
                                                             float* __restrict__ dw, DwWgParams p) {
  using V = typename WgVec<T>::type;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const size_t x_bytes = (size_t)p.THI * p.TWI * p.CB * sizeof(T);
  const size_t g_bytes = (size_t)p.TH * p.TW * p.CB * sizeof(T);
  const size_t x_stride = (x_bytes + 127) / 128 * 128, g_stride = (g_bytes + 127) / 128 * 128;
  const size_t buf_stride = x_stride + g_stride;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + 2 * buf_stride);

  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % p.chunks, slot = blockIdx.x / p.chunks;
  const int c0 = chunk * p.CB;
  const int n_tiles = p.tiles_w * p.tiles_h;

  auto issue = [&](long long t, int buf) {
    const int tw = (int)(t % p.tiles_w), th = (int)((t / p.tiles_w) % p.tiles_h), b = (int)(t / n_tiles);
    unsigned char* base = smem_raw + buf * buf_stride;
    mbar_expect_tx(&mbar[buf], (uint32_t)(x_bytes + g_bytes));
    tma_load_4d(base, &tm_x, &mbar[buf], c0, tw * p.TW * S - p.pad, th * p.TH * S - p.pad, b);
    tma_load_4d(base + x_stride, &tm_g, &mbar[buf], c0, tw * p.TW, th * p.TH, b);
  };
  long long t = slot;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
    if (t < p.per_chunk) issue(t, 0);
  }
  __syncthreads();

  const int G = p.G;
  const int g = tid % G, kh = (tid / G) % K, unit = tid / (G * K);
  const bool active = unit < p.units && c0 + g * 8 < p.C;
  const int r = unit / p.NS, j = unit % p.NS;
  const int x_off = ((r * S + kh) * p.TWI + j * p.SL * S) * p.CB + g * 8;
  const int g_off = (r * p.TW + j * p.SL) * p.CB + g * 8;
  float acc[K][8];
#pragma unroll
  for (int kw = 0; kw < K; ++kw)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[kw][e] = 0.f;

  int it = 0;
  for (; t < p.per_chunk; t += p.ctas_per_chunk, ++it) {
    const int buf = it & 1;
    if (tid == 0 && t + p.ctas_per_chunk < p.per_chunk) issue(t + p.ctas_per_chunk, buf ^ 1);
    mbar_wait(&mbar[buf], (it >> 1) & 1, 21);
    if (active) {
      const T* xrow = reinterpret_cast<const T*>(smem_raw + buf * buf_stride) + x_off;
      const T* grow = reinterpret_cast<const T*>(smem_raw + buf * buf_stride + x_stride) + g_off;
      if constexpr (S == 1) {
        V xw[K];
#pragma unroll
        for (int u = 0; u < K - 1; ++u) wg_ld(xrow + u * p.CB, xw[u]);
        for (int wo0 = 0; wo0 < p.SL; wo0 += K) {
#pragma unroll
          for (int u = 0; u < K; ++u) {
            const int wo = wo0 + u;
            V gv;
            wg_ld(grow + wo * p.CB, gv);
            wg_ld(xrow + (wo + K - 1) * p.CB, xw[(u + K - 1) % K]);
#pragma unroll
            for (int kw = 0; kw < K; ++kw) wg_fma8(xw[(u + kw) % K], gv, acc[kw]);
          }
        }
      } else {
        for (int wo = 0; wo < p.SL; ++wo) {
          V gv;
          wg_ld(grow + wo * p.CB, gv);
#pragma unroll
          for (int kw = 0; kw < K; ++kw) {
            V xv;
            wg_ld(xrow + (wo * S + kw) * p.CB, xv);
            wg_fma8(xv, gv, acc[kw]);
          }
        }
      }
    }
    __syncthreads();   // everyone is done with tile[buf] before it is refilled
  }

  // CTA reduction over units (the tile buffers are free now), then one atomic per (tap, channel)
  float* red = reinterpret_cast<float*>(smem_raw);
  const int per_unit = K * G * K * 8;
  if (unit < p.units) {
    float* s = red + (size_t)unit * per_unit + ((size_t)kh * G + g) * K * 8;
#pragma unroll
    for (int kw = 0; kw < K; ++kw)
#pragma unroll
      for (int e = 0; e < 8; ++e) s[kw * 8 + e] = acc[kw][e];
  }
  __syncthreads();
  for (int o = tid; o < per_unit; o += blockDim.x) {
    float s = 0.f;
    for (int u = 0; u < p.units; ++u) s += red[(size_t)u * per_unit + o];
    const int e = o % 8, kw = (o / 8) % K, g2 = (o / (8 * K)) % G, kh2 = o / (8 * K * G);
    const int cc = c0 + g2 * 8 + e;
    if (cc < p.C) atomicAdd(dw + (size_t)(kh2 * K + kw) * p.C + cc, s);
  }
}

// ------------------------------------------------------------------------------------ stem_wgrad
// dW[co][ci][kh][kw] = sum over (b, ho, wo) of g[b][ho][wo][co] * x[b][ci][2 ho + kh][2 wo + kw]
// thread = (tap in 0..26, row lane): it keeps all 48 output-channel accumulators, so one pixel costs one x load plus
// six 16-byte g loads (the same six for all 27 taps of a lane: served by one L1 line) for 48 FMAs.  The first version
// (thread = tap x 8 channels) issued two loads per 8 FMAs and ran at the load-issue limit (0.9 ms at batch 64).
constexpr int kSwLanes = 9;                       // output rows in flight per CTA
constexpr int kSwThreads = 27 * kSwLanes;         // 243

// acc[0..1] += x.{lo,hi} * w.{lo,hi}, bf16 operands from the packed registers, fp32 accumulate (SASS FHFMA.BF16)
__device__ __forceinline__ void stem_fma_pair(uint32_t x, uint32_t w, float& a0, float& a1) {
  asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\t"
      "mov.b32 {xl, xh}, %2;\n\t"
      "mov.b32 {wl, wh}, %3;\n\t"
      "fma.rn.f32.bf16 %0, xl, wl, %0;\n\t"
      "fma.rn.f32.bf16 %1, xh, wh, %1;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "r"(x), "r"(w));
}

template <typename T>
__global__ void __launch_bounds__(kSwThreads) stem_wgrad_kernel(const T* __restrict__ g, const float* __restrict__ x,
                                                               float* __restrict__ dw, int B, int H, int W, int Ho, int Wo,
                                                               long long rows_per_cta) {
  constexpr int CO = 48;
  __shared__ float red[kSwLanes][27][CO + 1];
  const int tid = threadIdx.x;
  const int tap = tid % 27, lane = tid / 27;        // tap = ci * 9 + kh * 3 + kw  (torch weight layout [co][ci][kh][kw])
  const int ci = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
  const long long nrows = (long long)B * Ho;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, nrows);
  float acc[CO];
#pragma unroll
  for (int e = 0; e < CO; ++e) acc[e] = 0.f;
  for (long long r = r0 + lane; r < r1; r += kSwLanes) {
    const int ho = (int)(r % Ho), b = (int)(r / Ho);
    const int hi = 2 * ho + kh;
    if (hi >= H) continue;
    const float* xr = x + (((size_t)b * 3 + ci) * H + hi) * W + kw;
    const T* gr = g + (size_t)r * Wo * CO;
    const int wmax = min(Wo, (W - kw + 1) / 2);     // 2 wo + kw < W
#pragma unroll 2
    for (int wo = 0; wo < wmax; ++wo) {
      const float xv = __ldg(xr + 2 * wo);
      if constexpr (sizeof(T) == 2) {
        // bf16 path: the forward stem multiplies bf16(x) on the tensor cores, so does its weight gradient -- and with both
        // operands bf16 the sm_100 mixed-precision FMA takes g straight from the packed registers (no 48 unpack
        // instructions per 48 FMAs: the kernel ran at ~40 % FMA density, 0.40 ms at batch 64)
        const uint32_t x2 = pack_bf16(xv, xv);
#pragma unroll
        for (int c8 = 0; c8 < CO / 8; ++c8) {
          const uint4 gw = *reinterpret_cast<const uint4*>(gr + (size_t)wo * CO + c8 * 8);
          stem_fma_pair(x2, gw.x, acc[c8 * 8 + 0], acc[c8 * 8 + 1]);
          stem_fma_pair(x2, gw.y, acc[c8 * 8 + 2], acc[c8 * 8 + 3]);
          stem_fma_pair(x2, gw.z, acc[c8 * 8 + 4], acc[c8 * 8 + 5]);
          stem_fma_pair(x2, gw.w, acc[c8 * 8 + 6], acc[c8 * 8 + 7]);
        }
      } else {
#pragma unroll
        for (int c8 = 0; c8 < CO / 8; ++c8) {
          float gv[8];
          load8(gr + (size_t)wo * CO + c8 * 8, gv);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[c8 * 8 + e] = fmaf(xv, gv[e], acc[c8 * 8 + e]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < CO; ++e) red[lane][tap][e] = acc[e];
  __syncthreads();
  for (int o = tid; o < 27 * CO; o += kSwThreads) {
    const int t = o / CO, co = o % CO;
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < kSwLanes; ++l) s += red[l][t][co];
    atomicAdd(dw + (size_t)co * 27 + t, s);
  }
}

}  // namespace dfv

namespace dfv {
template <typename T, int K, int S>
static int launch_dw_wgrad(const CUtensorMap& tmx, const CUtensorMap& tmg, float* dw, const DwWgParams& p, size_t smem, unsigned grid,
                           cudaStream_t st) {
  auto kern = dw_wgrad_tma_kernel<T, K, S>;
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DFV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  DFV_TRY(init_timeout_word_tu());
  kern<<<grid, 256, smem, st>>>(tmx, tmg, dw, p);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
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
  if (dtype == DFV_BF16 && K % 8 == 0 && N % 8 == 0)
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
  const double es = (double)dtype_size(dtype);
  ProfScope prof(PK_DWCONV_BWD, es * ((double)B * H * W * C + (double)B * Ho * Wo * C), 2.0 * kernel * kernel * (double)B * Ho * Wo * C, st);
  if (stride == 2 && (kernel == 3 || kernel == 5)) {
    DwDgParams p;
    p.C = C; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.pad = pad_lo;
    const int cap = dtype == DFV_BF16 ? 64 : 32;
    p.CB = 8;
    for (int cb = 8; cb <= cap; cb += 8)
      if (C % cb == 0) p.CB = cb;
    p.G = p.CB / 8;
    p.TW = 16;
    p.TH = std::max(1, std::min(256 / (p.G * (p.TW / 4)), 16));
    p.THg = (p.TH + kernel - 1) / 2 + 1;
    p.TWg = (p.TW + kernel - 1) / 2 + 1;
    p.tiles_w = (W + p.TW - 1) / p.TW;
    p.tiles_h = (H + p.TH - 1) / p.TH;
    p.chunks = (C + p.CB - 1) / p.CB;
    p.per_chunk = (long long)B * p.tiles_w * p.tiles_h;
    long long cpc = std::max<long long>(1, (long long)num_sms() * 2 / p.chunks);
    cpc = std::min(cpc, p.per_chunk);
    p.tiles_per_cta = (p.per_chunk + cpc - 1) / cpc;
    cpc = (p.per_chunk + p.tiles_per_cta - 1) / p.tiles_per_cta;
    p.ctas_per_chunk = (int)cpc;
    const size_t g_stride = align_up((size_t)p.THg * p.TWg * p.CB * dtype_size(dtype), 128);
    const size_t smem = 2 * g_stride + (size_t)kernel * kernel * p.CB * 4 + 64;
    CUtensorMap tmg;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)C * dtype_size(dtype), (uint64_t)Wo * C * dtype_size(dtype), (uint64_t)Ho * Wo * C * dtype_size(dtype)};
    uint32_t box[4] = {(uint32_t)p.CB, (uint32_t)p.TWg, (uint32_t)p.THg, 1};
    DFV_TRY(make_tensor_map(&tmg, dtype, 4, g, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE));
    DFV_TRY(init_timeout_word_tu());
    const unsigned grid = (unsigned)(cpc * p.chunks);
#define DGL(T_, K_)                                                                                                              \
  do {                                                                                                                           \
    static thread_local bool configured = false;                                                                                 \
    if (!configured) {                                                                                                           \
      DFV_CUDA(cudaFuncSetAttribute(dw_dgrad_s2_kernel<T_, K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));       \
      configured = true;                                                                                                         \
    }                                                                                                                            \
    dw_dgrad_s2_kernel<T_, K_><<<grid, 256, smem, st>>>(tmg, w_kkc, (T_*)dx, p);                                                 \
  } while (0)
    if (dtype == DFV_BF16) { if (kernel == 3) DGL(__nv_bfloat16, 3); else DGL(__nv_bfloat16, 5); }
    else { if (kernel == 3) DGL(float, 3); else DGL(float, 5); }
#undef DGL
    DFV_LAUNCH_CHECK();
    return DFV_OK;
  }
  const long long total = (long long)B * H * W * (C / 8);
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 16LL * num_sms());
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
  DFV_REQUIRE(valid_dtype(dtype) && B > 0 && C > 0 && C % 8 == 0 && (kernel == 3 || kernel == 5) && (stride == 1 || stride == 2),
              "dfv_dwconv_wgrad: bad shape");
  const int Ho = (H + pad_lo + pad_hi - kernel) / stride + 1, Wo = (W + pad_lo + pad_hi - kernel) / stride + 1;
  DFV_REQUIRE(Ho > 0 && Wo > 0, "dfv_dwconv_wgrad: bad shape");
  cudaStream_t st = as_stream(stream);
  const int K = kernel, S = stride;
  const size_t es = dtype_size(dtype);
  DwWgParams p;
  p.C = C;
  p.pad = pad_lo;
  const int cap = dtype == DFV_BF16 ? 32 : 16;
  p.CB = 8;
  for (int cb = 8; cb <= cap; cb += 8)
    if (C % cb == 0) p.CB = cb;
  p.G = p.CB / 8;
  const int units_max = 256 / (p.G * K);
  if (S == 1) {
    p.SL = K == 3 ? 12 : (Wo <= 15 ? 15 : 25);
    p.NS = (K == 3 && Wo > 12) ? 2 : 1;
  } else {
    p.SL = 8;
    p.NS = Wo > 8 ? 2 : 1;
  }
  p.TW = p.SL * p.NS;
  p.TH = std::min(std::min(units_max / p.NS, S == 1 ? 16 : 8), Ho);
  auto smem_of = [&](int th) {
    const size_t xb = align_up((size_t)((th - 1) * S + K) * ((p.TW - 1) * S + K) * p.CB * es, 128);
    const size_t gb = align_up((size_t)th * p.TW * p.CB * es, 128);
    return 2 * (xb + gb) + 64;
  };
  while (p.TH > 1 && smem_of(p.TH) > 100 * 1024) --p.TH;
  p.THI = (p.TH - 1) * S + K;
  p.TWI = (p.TW - 1) * S + K;
  p.units = p.TH * p.NS;
  p.tiles_w = (Wo + p.TW - 1) / p.TW;
  p.tiles_h = (Ho + p.TH - 1) / p.TH;
  p.chunks = (C + p.CB - 1) / p.CB;
  p.per_chunk = (long long)B * p.tiles_w * p.tiles_h;
  const size_t red_bytes = (size_t)p.units * K * p.G * K * 8 * sizeof(float);
  const size_t smem = std::max(smem_of(p.TH), red_bytes + 64);
  DFV_REQUIRE(smem <= 200 * 1024 && p.TWI <= 256 && p.THI <= 256, "dfv_dwconv_wgrad: cannot tile H=%d W=%d C=%d k=%d s=%d", H, W, C, K, S);
  long long cpc = (long long)num_sms() * 2 / p.chunks;
  if (cpc < 1) cpc = 1;
  if (cpc > p.per_chunk) cpc = p.per_chunk;
  p.ctas_per_chunk = (int)cpc;
  const unsigned grid = (unsigned)(cpc * p.chunks);

  CUtensorMap tmx, tmg;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)C * es, (uint64_t)W * C * es, (uint64_t)H * W * C * es};
    uint32_t box[4] = {(uint32_t)p.CB, (uint32_t)p.TWI, (uint32_t)p.THI, 1};
    DFV_TRY(make_tensor_map(&tmx, dtype, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE));
  }
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)C * es, (uint64_t)Wo * C * es, (uint64_t)Ho * Wo * C * es};
    uint32_t box[4] = {(uint32_t)p.CB, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    DFV_TRY(make_tensor_map(&tmg, dtype, 4, g, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE));
  }
  ProfScope prof(PK_DWCONV_BWD, es * ((double)B * H * W * C + (double)B * Ho * Wo * C), 2.0 * kernel * kernel * (double)B * Ho * Wo * C, st);
#define DWW(T_, K_, S_) return launch_dw_wgrad<T_, K_, S_>(tmx, tmg, dw_kkc, p, smem, grid, st)
  if (dtype == DFV_BF16) {
    if (K == 3 && S == 1) DWW(__nv_bfloat16, 3, 1);
    if (K == 3 && S == 2) DWW(__nv_bfloat16, 3, 2);
    if (K == 5 && S == 1) DWW(__nv_bfloat16, 5, 1);
    DWW(__nv_bfloat16, 5, 2);
  } else {
    if (K == 3 && S == 1) DWW(float, 3, 1);
    if (K == 3 && S == 2) DWW(float, 3, 2);
    if (K == 5 && S == 1) DWW(float, 5, 1);
    DWW(float, 5, 2);
  }
#undef DWW
}

/* Stem weight gradient, accumulated into dw fp32 [48][3][3][3] (torch layout), which the caller zeroes.
 * g: [B][Ho][Wo][48] gradient wrt the raw (pre-BN) stem output; x: the NCHW fp32 images. */
int dfv_stem_wgrad(const void* g, const float* x_nchw, float* dw, int dtype, int B, int H, int W, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(g && x_nchw && dw && valid_dtype(dtype) && B > 0 && H >= 3 && W >= 3, "dfv_stem_wgrad: bad arguments");
  const int Ho = (H + 1 - 3) / 2 + 1, Wo = (W + 1 - 3) / 2 + 1;
  const long long nrows = (long long)B * Ho;
  const long long ctas = std::min<long long>((nrows + kSwLanes - 1) / kSwLanes, 2LL * num_sms());
  const long long rpc = (nrows + ctas - 1) / ctas;
  cudaStream_t st = as_stream(stream);
  ProfScope prof(PK_STEM, (double)B * 3 * H * W * 4 + (double)nrows * Wo * 48 * dtype_size(dtype), 2.0 * 27 * 48 * (double)nrows * Wo, st);
  if (dtype == DFV_BF16)
    stem_wgrad_kernel<__nv_bfloat16><<<(unsigned)((nrows + rpc - 1) / rpc), kSwThreads, 0, st>>>((const __nv_bfloat16*)g, x_nchw, dw, B, H, W, Ho, Wo, rpc);
  else
    stem_wgrad_kernel<float><<<(unsigned)((nrows + rpc - 1) / rpc), kSwThreads, 0, st>>>((const float*)g, x_nchw, dw, B, H, W, Ho, Wo, rpc);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

}  // extern "C"
