// Depthwise k x k convolution (k in {3,5}, stride in {1,2}), NHWC, fused with folded
// BatchNorm, swish and the squeeze-excite global-average-pool partial sums.
//
// Data movement: one 4-D TMA tile load per CTA (channels x W x H x image) straight into
// shared memory.  The tile origin may be negative / the tile may overhang the tensor: TMA
// zero-fills out-of-bounds elements, which IS the reference's static asymmetric "SAME"
// padding (efficientnet-pytorch Conv2dStaticSamePadding pads with an explicit ZeroPad2d
// copy; here no padded tensor is ever materialised).
// Compute: each thread owns 8 consecutive channels (one 16-byte vector) and a strip of L
// output pixels along W; the input row window slides through registers so each shared-memory
// vector is read once per kernel row.  fp32 accumulation.
// Bound: HBM (read C*H*W + write C*Ho*Wo elements per image), with the 5x5 layers close to
// the FP32-FMA limit (25 FMA per output).
#include "common.cuh"

namespace dfv {

struct DwParams {
  int C, Ho, Wo;
  int CB;          // channels per CTA (multiple of 8)
  int TH, TW;      // output tile
  int THI, TWI;    // input tile = (T-1)*S + K
  int tiles_w, tiles_h;
  int pad;         // pad_lo (top == left)
  int act;
  int nthreads;
};

template <typename T>
struct WVec;  // 8 weights of type T in shared memory -> 8 floats

template <typename T, int K, int S, int L, bool kFast>
__global__ void __launch_bounds__(384) dwconv_kernel(const __grid_constant__ CUtensorMap tmap,
                                                    const float* __restrict__ w, const float* __restrict__ bias,
                                                    T* __restrict__ y, float* __restrict__ pool_partial, DwParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // [tile: THI*TWI*CB T][weights: K*K*CB T][bias: CB f32][red: nthreads*8 f32][mbar]
  T* tile = reinterpret_cast<T*>(smem_raw);
  const size_t tile_bytes = (size_t)p.THI * p.TWI * p.CB * sizeof(T);
  T* wsm = reinterpret_cast<T*>(smem_raw + ((tile_bytes + 127) / 128) * 128);
  float* bsm = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(wsm) + (((size_t)K * K * p.CB * sizeof(T) + 15) / 16) * 16);
  float* red = bsm + p.CB;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(red + (size_t)p.nthreads * 8);

  const int tid = threadIdx.x;
  const int tile_id = blockIdx.x;
  const int tw_i = tile_id % p.tiles_w, th_i = tile_id / p.tiles_w;
  const int c0 = blockIdx.y * p.CB;
  const int b = blockIdx.z;
  const int h0 = th_i * p.TH, w0 = tw_i * p.TW;

  if (tid == 0) {
    mbar_init(mbar, 1);
    fence_mbar_init();
    mbar_expect_tx(mbar, (uint32_t)tile_bytes);
    tma_load_4d(tile, &tmap, mbar, c0, w0 * S - p.pad, h0 * S - p.pad, b);
  }
  // Stage this chunk's weights (BN scale already folded in) and bias while the tile lands.
  for (int i = tid; i < K * K * p.CB; i += blockDim.x) {
    const int c = c0 + i % p.CB;
    const float v = (c < p.C) ? w[(size_t)(i / p.CB) * p.C + c] : 0.f;
    if constexpr (sizeof(T) == 2) wsm[i] = __float2bfloat16_rn(v); else wsm[i] = v;
  }
  for (int i = tid; i < p.CB; i += blockDim.x) bsm[i] = (c0 + i < p.C) ? bias[c0 + i] : 0.f;
  __syncthreads();
  mbar_wait(mbar, 0);

  const int G = p.CB >> 3;
  const int strips = (p.TW + L - 1) / L;
  const int n_items = p.TH * strips * G;
  const int g = tid % G;  // blockDim is a multiple of G: the channel group is fixed per thread
  const int c = c0 + g * 8;
  float psum[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) psum[e] = 0.f;
  float bv[8];
  load8(bsm + g * 8, bv);

  constexpr int NI = (L - 1) * S + K;  // input window per kernel row
  if (c < p.C) {
    for (int item = tid; item < n_items; item += blockDim.x) {
      const int rest = item / G;
      const int j = rest % strips, r = rest / strips;
      const int ho = h0 + r, wo0 = w0 + j * L;
      if (ho >= p.Ho || wo0 >= p.Wo) continue;
      float acc[L][8];
#pragma unroll
      for (int l = 0; l < L; ++l)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[l][e] = 0.f;

#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        float wk[K][8];
#pragma unroll
        for (int kw = 0; kw < K; ++kw) load8(wsm + (kh * K + kw) * p.CB + g * 8, wk[kw]);
        const T* row = tile + ((size_t)(r * S + kh) * p.TWI + j * L * S) * p.CB + g * 8;
#pragma unroll
        for (int iw = 0; iw < NI; ++iw) {
          // the last strip of a tile may overhang TWI; those outputs are masked below
          if (j * L * S + iw < p.TWI) {
            float v[8];
            load8(row + (size_t)iw * p.CB, v);
#pragma unroll
            for (int l = 0; l < L; ++l) {
              const int kw = iw - l * S;
              if (kw >= 0 && kw < K) {
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[l][e] = fmaf(v[e], wk[kw][e], acc[l][e]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int l = 0; l < L; ++l) {
        const int wo = wo0 + l;
        if (wo < p.Wo && j * L + l < p.TW) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float t = acc[l][e] + bv[e];
            o[e] = p.act ? silu<kFast>(t) : t;
            psum[e] += o[e];
          }
          store8(y + (((size_t)b * p.Ho + ho) * p.Wo + wo) * p.C + c, o);
        }
      }
    }
  }

  if (pool_partial != nullptr) {
    // deterministic CTA reduction: threads sharing a channel group are tid = g + G*i
#pragma unroll
    for (int e = 0; e < 8; ++e) red[tid * 8 + e] = psum[e];
    __syncthreads();
    if (tid < p.CB && c0 + tid < p.C) {
      const int gg = tid >> 3, e = tid & 7;
      float s = 0.f;
      for (int t = gg; t < (int)blockDim.x; t += G) s += red[t * 8 + e];
      const int parts = p.tiles_w * p.tiles_h;
      pool_partial[((size_t)b * parts + tile_id) * p.C + c0 + tid] = s;
    }
  }
}

// ------------------------------------------------------------------------------------
struct DwPlan {
  DwParams p;
  int L;
  size_t smem;
  int chunks;
};

static int pick_cb(int C, int dtype) {
  const int cap = dtype == DFV_BF16 ? 64 : 32;
  int best = 8;
  for (int cb = 8; cb <= cap; cb += 8)
    if (C % cb == 0) best = cb;
  return best;
}

static int make_plan(DwPlan* pl, int dtype, int H, int W, int C, int K, int S, int pad_lo, int pad_hi) {
  DwParams& p = pl->p;
  p.C = C;
  p.Ho = (H + pad_lo + pad_hi - K) / S + 1;
  p.Wo = (W + pad_lo + pad_hi - K) / S + 1;
  if (p.Ho <= 0 || p.Wo <= 0) return DFV_ERR_INVALID;
  p.CB = pick_cb(C, dtype);
  p.pad = pad_lo;
  int L, TW, TH;
  if (S == 2) {
    L = 4;
    TW = p.Wo >= 16 ? 16 : ((p.Wo + 3) / 4) * 4;
    TH = p.Ho >= 8 ? 8 : p.Ho;
  } else if (p.Wo > 24) {
    L = 8;
    TW = p.Wo >= 32 ? 32 : 24;
    if (p.Wo > 32 && p.Wo <= 48) TW = 24;
    TH = 8;
  } else if (p.Wo > 12) {
    L = 8;
    TW = ((p.Wo + 7) / 8) * 8;
    TH = p.Ho >= 12 ? 12 : p.Ho;
  } else {
    L = p.Wo > 8 ? 6 : 4;
    TW = ((p.Wo + L - 1) / L) * L;
    TH = p.Ho;
  }
  pl->L = L;
  p.TW = TW;
  p.TH = TH;
  p.TWI = (TW - 1) * S + K;
  p.THI = (TH - 1) * S + K;
  p.tiles_w = (p.Wo + TW - 1) / TW;
  p.tiles_h = (p.Ho + TH - 1) / TH;
  pl->chunks = (C + p.CB - 1) / p.CB;
  const int G = p.CB / 8;
  const int items = TH * (TW / L) * G;
  const int rounds = (items + 319) / 320;  // aim for <= 320 threads, balanced rounds
  int nt = (items + rounds - 1) / rounds;
  nt = ((nt + G - 1) / G) * G;
  if (nt > 384) nt = (384 / G) * G;
  if (nt < G) nt = G;
  p.nthreads = nt;
  const size_t ts = dtype_size(dtype);
  size_t tile_bytes = (size_t)p.THI * p.TWI * p.CB * ts;
  pl->smem = align_up(tile_bytes, 128) + align_up((size_t)K * K * p.CB * ts, 16) + (size_t)p.CB * 4 + (size_t)nt * 32 + 16;
  if (p.TWI > 256 || p.THI > 256 || p.CB > 256 || pl->smem > 200 * 1024) return DFV_ERR_INVALID;
  return DFV_OK;
}

template <typename T, int K, int S, int L, bool kFast>
static int launch(const CUtensorMap& tm, const float* w, const float* bias, void* y, float* pool, const DwPlan& pl, int B,
                  cudaStream_t st) {
  auto kern = dwconv_kernel<T, K, S, L, kFast>;
  static thread_local size_t configured = 0;
  if (pl.smem > 48 * 1024 && pl.smem > configured) {
    DFV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  dim3 grid(pl.p.tiles_w * pl.p.tiles_h, pl.chunks, B);
  kern<<<grid, pl.p.nthreads, pl.smem, st>>>(tm, w, bias, (T*)y, pool, pl.p);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

template <typename T, bool kFast>
static int dispatch(int K, int S, int L, const CUtensorMap& tm, const float* w, const float* bias, void* y, float* pool,
                    const DwPlan& pl, int B, cudaStream_t st) {
#define DW_CASE(k, s, l) \
  if (K == k && S == s && L == l) return launch<T, k, s, l, kFast>(tm, w, bias, y, pool, pl, B, st);
  DW_CASE(3, 1, 8) DW_CASE(3, 1, 6) DW_CASE(3, 1, 4) DW_CASE(5, 1, 8) DW_CASE(5, 1, 6) DW_CASE(5, 1, 4)
  DW_CASE(3, 2, 4) DW_CASE(5, 2, 4)
#undef DW_CASE
  set_error("dfv_dwconv_fwd: unsupported kernel/stride/strip combination k=%d s=%d L=%d", K, S, L);
  return DFV_ERR_INVALID;
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_dwconv_pool_parts(int dtype, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi) {
  DwPlan pl;
  if (!valid_dtype(dtype) || make_plan(&pl, dtype, H, W, C, kernel, stride, pad_lo, pad_hi) != DFV_OK) {
    set_error("dfv_dwconv_pool_parts: bad shape");
    return DFV_ERR_INVALID;
  }
  return pl.p.tiles_w * pl.p.tiles_h;
}

extern "C" int dfv_dwconv_fwd(const void* x, const float* w, const float* bias, void* y, float* pool_partial, int dtype,
                              int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, int act,
                              dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(x && w && bias && y, "dfv_dwconv_fwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype), "dfv_dwconv_fwd: bad dtype %d", dtype);
  DFV_REQUIRE((kernel == 3 || kernel == 5) && (stride == 1 || stride == 2), "dfv_dwconv_fwd: k=%d s=%d unsupported",
              kernel, stride);
  DFV_REQUIRE(B > 0 && B <= 65535 && C > 0 && C % 8 == 0, "dfv_dwconv_fwd: need 0 < B <= 65535 and C %% 8 == 0 (B=%d C=%d)", B, C);
  DFV_REQUIRE(pad_lo >= 0 && pad_hi >= 0 && pad_lo < kernel && pad_hi < kernel, "dfv_dwconv_fwd: bad pad");
  DwPlan pl;
  DFV_REQUIRE(make_plan(&pl, dtype, H, W, C, kernel, stride, pad_lo, pad_hi) == DFV_OK,
              "dfv_dwconv_fwd: cannot tile H=%d W=%d C=%d k=%d s=%d", H, W, C, kernel, stride);
  pl.p.act = act;
  const size_t es = dtype_size(dtype);
  CUtensorMap tm;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)C * es, (uint64_t)W * C * es, (uint64_t)H * W * C * es};
  uint32_t box[4] = {(uint32_t)pl.p.CB, (uint32_t)pl.p.TWI, (uint32_t)pl.p.THI, 1};
  DFV_TRY(make_tensor_map(&tm, dtype, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE));
  ProfScope prof(PK_DWCONV, ((double)B * H * W * C + (double)B * pl.p.Ho * pl.p.Wo * C) * es,
                 2.0 * kernel * kernel * (double)B * pl.p.Ho * pl.p.Wo * C, as_stream(stream));
  if (dtype == DFV_BF16)
    return dispatch<__nv_bfloat16, true>(kernel, stride, pl.L, tm, w, bias, y, pool_partial, pl, B, as_stream(stream));
  return dispatch<float, false>(kernel, stride, pl.L, tm, w, bias, y, pool_partial, pl, B, as_stream(stream));
}
