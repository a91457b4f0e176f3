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
#include <algorithm>

#include "common.cuh"

namespace dfv {

struct DwParams {
  int C, Ho, Wo;
  int CB;          // channels per CTA (multiple of 8)
  int TH, TW;      // output tile
  int THI, TWI;    // input tile = (T-1)*S + K
  int tiles_w, tiles_h;
  int chunks;      // channel chunks; gridDim.x = chunks * ctas_per_chunk, so a CTA's chunk is fixed
  long long per_chunk;   // B * tiles_w * tiles_h tiles per channel chunk
  int ctas_per_chunk;
  int d_tw, d_th, d_b;   // tile-coordinate increments of a step of ctas_per_chunk tiles
  int pad;         // pad_lo (top == left)
  int act;
  int nthreads;    // = G * strips * rpr: thread -> (channel group, strip, row) is fixed for the whole kernel
  int strips;      // TW / L
  int rpr;         // tile rows per round
  int rounds;      // ceil(TH / rpr)
  int red_parts;   // second-stage partials of the pool reduction
};

// 8 channels of activations / weights as they sit in shared memory.
template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> { uint4 r; };
template <> struct Vec8<float> { float v[8]; };

__device__ __forceinline__ void ldvec(const __nv_bfloat16* p, Vec8<__nv_bfloat16>& o) { o.r = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void ldvec(const float* p, Vec8<float>& o) { load8(p, o.v); }

// acc[0..1] += x.{lo,hi} * w.{lo,hi}: sm_100 mixed-precision FMA (SASS FHFMA.BF16) -- bf16 operands are
// taken straight from the packed registers (half selectors), fp32 accumulate; the product of two
// bf16 values is exact in fp32, so this equals convert-then-FFMA bit for bit without the 2 unpack
// instructions per pair.
__device__ __forceinline__ void fma_pair(uint32_t x, uint32_t w, float& a0, float& a1) {
  asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\t"
      "mov.b32 {xl, xh}, %2;\n\t"
      "mov.b32 {wl, wh}, %3;\n\t"
      "fma.rn.f32.bf16 %0, xl, wl, %0;\n\t"
      "fma.rn.f32.bf16 %1, xh, wh, %1;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "r"(x), "r"(w));
}
__device__ __forceinline__ void fma8(const Vec8<__nv_bfloat16>& x, const Vec8<__nv_bfloat16>& w, float acc[8]) {
  fma_pair(x.r.x, w.r.x, acc[0], acc[1]);
  fma_pair(x.r.y, w.r.y, acc[2], acc[3]);
  fma_pair(x.r.z, w.r.z, acc[4], acc[5]);
  fma_pair(x.r.w, w.r.w, acc[6], acc[7]);
}
__device__ __forceinline__ void fma8(const Vec8<float>& x, const Vec8<float>& w, float acc[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = fmaf(x.v[e], w.v[e], acc[e]);
}

// Persistent CTA: loops over the tiles (image, tile row, tile column) of ONE channel chunk with a 2-deep
// TMA pipeline: the tile for step i+1 is in flight while step i is computed.  All index arithmetic that
// does not depend on the tile is hoisted: a thread keeps its (channel group, strip, first row) for the
// whole kernel and tile coordinates advance by precomputed increments (no divisions in the loops).
template <typename T, int K, int S, int L, bool kFast>
__global__ void __launch_bounds__(256, 2) dwconv_kernel(const __grid_constant__ CUtensorMap tmap,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       T* __restrict__ y, float* __restrict__ pool_partial, DwParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // [tile0][tile1][weights: K*K*CB T][bias: CB f32][red: nthreads*8 f32][red2: 4*CB f32][mbar x2]
  const size_t tile_bytes = (size_t)p.THI * p.TWI * p.CB * sizeof(T);
  const size_t tile_stride = ((tile_bytes + 127) / 128) * 128;
  T* wsm = reinterpret_cast<T*>(smem_raw + 2 * tile_stride);
  float* bsm = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(wsm) + (((size_t)K * K * p.CB * sizeof(T) + 15) / 16) * 16);
  float* red = bsm + p.CB;
  float* red2 = red + (size_t)p.nthreads * 8;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(red2 + 4 * p.CB);

  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % p.chunks, slot = blockIdx.x / p.chunks;
  const int c0 = chunk * p.CB;
  const int n_tiles = p.tiles_w * p.tiles_h;

  // tile coordinates of this CTA's current and next tile
  long long t = slot;
  int tw_i = (int)(t % p.tiles_w), th_i = (int)((t / p.tiles_w) % p.tiles_h), b = (int)(t / n_tiles);
  auto advance = [&](int& tw, int& th, int& bb) {
    tw += p.d_tw;
    if (tw >= p.tiles_w) { tw -= p.tiles_w; ++th; }
    th += p.d_th;
    if (th >= p.tiles_h) { th -= p.tiles_h; ++bb; }
    bb += p.d_b;
  };
  auto issue = [&](int tw, int th, int bb, int buf) {
    mbar_expect_tx(&mbar[buf], (uint32_t)tile_bytes);
    tma_load_4d(smem_raw + buf * tile_stride, &tmap, &mbar[buf], c0, tw * p.TW * S - p.pad, th * p.TH * S - p.pad, bb);
  };

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
    if (t < p.per_chunk) issue(tw_i, th_i, b, 0);
  }
  // Stage this chunk's weights (BN scale already folded in) and bias while the first tile lands.
  for (int i = tid; i < K * K * p.CB; i += blockDim.x) {
    const int c = c0 + i % p.CB;
    const float v = (c < p.C) ? w[(size_t)(i / p.CB) * p.C + c] : 0.f;
    if constexpr (sizeof(T) == 2) wsm[i] = __float2bfloat16_rn(v); else wsm[i] = v;
  }
  for (int i = tid; i < p.CB; i += blockDim.x) bsm[i] = (c0 + i < p.C) ? bias[c0 + i] : 0.f;
  __syncthreads();

  // fixed per-thread role
  const int G = p.CB >> 3;
  const int g = tid % G;
  const int j = (tid / G) % p.strips;
  const int r0 = tid / (G * p.strips);
  const int c = c0 + g * 8;
  const bool chan_ok = c < p.C;
  const int in_off0 = ((r0 * S) * p.TWI + j * L * S) * p.CB + g * 8;
  const int in_step = p.rpr * S * p.TWI * p.CB;
  const int out_off0 = (r0 * p.Wo + j * L) * p.C + g * 8;
  const int out_step = p.rpr * p.Wo * p.C;
  const int row_stride = p.TWI * p.CB;
  const T* wbase = wsm + g * 8;
  float bv[8];
  load8(bsm + g * 8, bv);
  constexpr int NI = (L - 1) * S + K;  // input window per kernel row
  const int nth = blockDim.x;

  int it = 0;
  for (; t < p.per_chunk; t += p.ctas_per_chunk, ++it) {
    const int buf = it & 1;
    int ntw = tw_i, nth_i = th_i, nb = b;
    advance(ntw, nth_i, nb);
    if (tid == 0 && t + p.ctas_per_chunk < p.per_chunk) issue(ntw, nth_i, nb, buf ^ 1);   // prefetch the next tile
    const int h0 = th_i * p.TH, w0 = tw_i * p.TW;
    const int hrem = p.Ho - h0;                    // valid rows in this tile
    const int wrem = p.Wo - (w0 + j * L);          // valid pixels from this thread's strip start
    const T* tile = reinterpret_cast<const T*>(smem_raw + buf * tile_stride);
    T* out_tile = y + (((size_t)b * p.Ho + h0) * p.Wo + w0) * p.C + c0;
    mbar_wait(&mbar[buf], (it >> 1) & 1, 6);

    float psum[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) psum[e] = 0.f;

    if (chan_ok && wrem > 0) {
      int r = r0, in_off = in_off0, out_off = out_off0;
      for (int rr = 0; rr < p.rounds; ++rr, r += p.rpr, in_off += in_step, out_off += out_step) {
        if (r >= p.TH || r >= hrem) break;
        float acc[L][8];
#pragma unroll
        for (int l = 0; l < L; ++l)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[l][e] = bv[e];

        const T* in = tile + in_off;
#pragma unroll
        for (int kh = 0; kh < K; ++kh) {
          Vec8<T> wk[K];
#pragma unroll
          for (int kw = 0; kw < K; ++kw) ldvec(wbase + (kh * K + kw) * p.CB, wk[kw]);
          const T* row = in + kh * row_stride;
#pragma unroll
          for (int iw = 0; iw < NI; ++iw) {
            Vec8<T> v;
            ldvec(row + iw * p.CB, v);
#pragma unroll
            for (int l = 0; l < L; ++l) {
              const int kw = iw - l * S;
              if (kw >= 0 && kw < K) fma8(v, wk[kw], acc[l]);
            }
          }
        }
        T* out = out_tile + out_off;
        if (wrem >= L) {
#pragma unroll
          for (int l = 0; l < L; ++l) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              o[e] = p.act ? silu<kFast>(acc[l][e]) : acc[l][e];
              psum[e] += o[e];
            }
            store8(out + (size_t)l * p.C, o);
          }
        } else {
#pragma unroll
          for (int l = 0; l < L; ++l) {
            if (l < wrem) {
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                o[e] = p.act ? silu<kFast>(acc[l][e]) : acc[l][e];
                psum[e] += o[e];
              }
              store8(out + (size_t)l * p.C, o);
            }
          }
        }
      }
    }

    if (pool_partial != nullptr) {
      // deterministic two-stage CTA reduction over the threads that share a channel group (tid = g + G*i)
#pragma unroll
      for (int e = 0; e < 8; ++e) red[tid * 8 + e] = psum[e];
      __syncthreads();   // also: everyone is done with tile[buf] before it is refilled
      const int R = p.red_parts;
      for (int idx = tid; idx < p.CB * R; idx += nth) {
        const int o = idx % p.CB, q = idx / p.CB;
        const int gg = o >> 3, e = o & 7;
        float s = 0.f;
        for (int u = gg + G * q; u < nth; u += G * R) s += red[u * 8 + e];
        red2[idx] = s;
      }
      __syncthreads();
      for (int o = tid; o < p.CB; o += nth) {
        if (c0 + o < p.C) {
          float s = 0.f;
          for (int q = 0; q < R; ++q) s += red2[q * p.CB + o];
          pool_partial[((size_t)b * n_tiles + th_i * p.tiles_w + tw_i) * p.C + c0 + o] = s;
        }
      }
    } else {
      __syncthreads();
    }
    tw_i = ntw;
    th_i = nth_i;
    b = nb;
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
  // strip length: 8 for 3x3 (weights + 64 accumulators fit 128 registers), 4 for 5x5 and stride 2
  int L = (S == 1) ? 8 : 4;
  if (p.Wo <= 12 && L == 8) L = p.Wo > 8 ? 6 : 4;
  int TW, TH;
  if (S == 2) {
    TW = p.Wo >= 16 ? 16 : ((p.Wo + L - 1) / L) * L;
    TH = p.Ho >= 8 ? 8 : p.Ho;
  } else if (p.Wo > 24) {
    TW = p.Wo >= 32 ? 32 : 24;
    if (p.Wo > 32 && p.Wo <= 48) TW = 24;
    TH = 8;
  } else {
    TW = ((p.Wo + L - 1) / L) * L;
    TH = p.Ho > 12 ? 8 : p.Ho;
  }
  pl->L = L;
  p.TW = TW;
  p.TH = TH;
  p.TWI = (TW - 1) * S + K;
  p.THI = (TH - 1) * S + K;
  p.tiles_w = (p.Wo + TW - 1) / TW;
  p.tiles_h = (p.Ho + TH - 1) / TH;
  p.chunks = pl->chunks = (C + p.CB - 1) / p.CB;
  const int G = p.CB / 8;
  p.strips = TW / L;
  const int per_row = G * p.strips;               // threads per tile row (<= 64)
  const int rpr_max = std::max(1, 256 / per_row);
  p.rounds = (TH + rpr_max - 1) / rpr_max;
  p.rpr = (TH + p.rounds - 1) / p.rounds;
  const int nt = per_row * p.rpr;
  p.nthreads = nt;
  p.red_parts = std::max(1, std::min(4, nt / p.CB));
  const size_t ts = dtype_size(dtype);
  size_t tile_bytes = (size_t)p.THI * p.TWI * p.CB * ts;
  pl->smem = 2 * align_up(tile_bytes, 128) + align_up((size_t)K * K * p.CB * ts, 16) + (size_t)p.CB * 4 + (size_t)nt * 32 +
             (size_t)p.CB * 16 + 32;
  if (p.TWI > 256 || p.THI > 256 || p.CB > 256 || pl->smem > 220 * 1024) return DFV_ERR_INVALID;
  return DFV_OK;
}

template <typename T, int K, int S, int L, bool kFast>
static int launch(const CUtensorMap& tm, const float* w, const float* bias, void* y, float* pool, DwPlan& pl, int B,
                  cudaStream_t st) {
  auto kern = dwconv_kernel<T, K, S, L, kFast>;
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    DFV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  DFV_TRY(init_timeout_word_tu());
  const long long per_chunk = (long long)B * pl.p.tiles_w * pl.p.tiles_h;
  pl.p.per_chunk = per_chunk;
  // persistent grid: a multiple of `chunks` (fixed chunk per CTA), about two CTAs per SM
  const int occ = pl.smem <= 110 * 1024 ? 2 : 1;
  long long ctas_per_chunk = (long long)num_sms() * occ / pl.chunks;
  if (ctas_per_chunk < 1) ctas_per_chunk = 1;
  if (ctas_per_chunk > per_chunk) ctas_per_chunk = per_chunk;
  pl.p.ctas_per_chunk = (int)ctas_per_chunk;
  pl.p.d_tw = (int)(ctas_per_chunk % pl.p.tiles_w);
  pl.p.d_th = (int)((ctas_per_chunk / pl.p.tiles_w) % pl.p.tiles_h);
  pl.p.d_b = (int)(ctas_per_chunk / ((long long)pl.p.tiles_w * pl.p.tiles_h));
  const unsigned grid = (unsigned)(ctas_per_chunk * pl.chunks);
  kern<<<grid, pl.p.nthreads, pl.smem, st>>>(tm, w, bias, (T*)y, pool, pl.p);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

template <typename T, bool kFast>
static int dispatch(int K, int S, int L, const CUtensorMap& tm, const float* w, const float* bias, void* y, float* pool,
                    DwPlan& pl, int B, cudaStream_t st) {
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
  DFV_REQUIRE(B > 0 && C > 0 && C % 8 == 0, "dfv_dwconv_fwd: need B > 0 and C %% 8 == 0 (B=%d C=%d)", B, C);
  DFV_REQUIRE(pad_lo >= 0 && pad_hi >= 0 && pad_lo < kernel && pad_hi < kernel, "dfv_dwconv_fwd: bad pad");
  if (debug_flags() & 1) return DFV_OK;
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
