// EXPERIMENT (round 1), not part of the product library -- see DESIGN.md section 8 for the result (correct, 2.3x slower than the
// SIMT kernel: shared-memory operand bandwidth).  To rebuild it: copy next to csrc/common.cuh, add to SRCS in csrc/Makefile, declare
// dfv_dwconv_tc_pool_parts / dfv_dwconv_tc_fwd in include/dfvit.h and _lib.py; scripts/experiments/check_dwconv_tc.py compares it
// with the SIMT kernel and torch.
//
// Depthwise k x k stride-1 convolution (+ folded BN bias, swish, SE pool partial sums) on the TENSOR CORES.
//
// A depthwise conv has no reduction across channels, so it is usually written off as FMA-pipe work: 25 FMAs per output
// for 5 x 5, one per lane per instruction -- the SIMT kernel (dwconv.cu) runs the FMA pipe at 60 % and still reaches only
// ~0.45 of the HBM roofline on those layers.  But with the NHWC tile in shared memory as 128-byte pixel rows (64 channels
// = one SWIZZLE_128B atom row, exactly what the 4-D TMA load writes), tap (ky, kx) of a 16-channel block IS a
// tcgen05.mma M128 x N16 x K16:
//     A = the tile viewed from pixel row ky * TWI + kx   (a descriptor start address -- no im2col; see dw_tc_probe.cu:
//         any 128-byte row offset works with the plain descriptor),
//     B = diag(w[tap][16 channels])                       (k*k small diagonal matrices, 2 KB per tap and 64-channel chunk),
//     D = 128 linearised output positions (pitch TWI) x 16 channels, fp32 in TMEM.
// 15/16 of the MACs multiply zeros; the tensor pipe does them at ~20 outputs/clk/SM against 3 on the FMA pipe.
// Positions with x >= TW (the halo columns of the linearisation) are computed and dropped.
//
// Persistent, warp-specialised: one producer / MMA-issuer thread, eight epilogue warps, two input stages and two TMEM
// accumulator buffers (the one-tile-per-CTA first version spent 54 % of its samples waiting on the load and the MMAs).
#include "common.cuh"

namespace dfv {

struct DwTcParams {
  int C, H, W, Ho, Wo;
  int K, pad;
  int TW, TH, TWI, THI;
  int tiles_w, tiles_h;
  int NP, MG;          // linearised positions per tile (TH * TWI), M groups of 128
  int act;
};

__device__ __forceinline__ uint64_t dwtc_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

constexpr int kDwTcEpiWarps = 8;
constexpr int kDwTcIssuers = 8;       // MMA-issuing warps, one per (16-channel block, M group) (a single thread issues ~1 MMA per 90 cycles)
constexpr int kDwTcThreads = (kDwTcEpiWarps + kDwTcIssuers) * 32;      // warps 0..7 epilogue, warps 8..11 (lane 0) MMA issuers; warp 8 also the TMA producer

struct __align__(8) DwTcBars {
  uint64_t full[2];        // input tile landed (TMA)
  uint64_t empty[2];       // the MMAs that read the stage completed
  uint64_t tmem_full[2];   // accumulators of a tile complete
  uint64_t tmem_empty[2];  // epilogue drained them
  uint32_t tmem_base;
};

// Persistent: each CTA owns one 64-channel chunk (its diagonal weight tiles are built once) and a contiguous range of
// that chunk's tiles; the input tile of tile i+1 is in flight while tile i is multiplied and tile i-1 is written out.
__global__ void __launch_bounds__(kDwTcThreads, 1)
    dwconv_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const float* __restrict__ w, const float* __restrict__ bias,
                     __nv_bfloat16* __restrict__ y, float* __restrict__ pool_partial, DwTcParams p, int ctas_per_chunk,
                     int tiles_per_cta, int n_tiles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int taps = p.K * p.K;
  const size_t xt_bytes = (size_t)(((p.THI * p.TWI + 8) * 128 + 1023) / 1024) * 1024;
  unsigned char* xt = smem;                          // [2 stages][THI * TWI (+8)][128 B]   input tiles, one row per pixel
  unsigned char* bt = xt + 2 * xt_bytes;             // [taps][16 rows][128 B]               diagonal weight blocks
  float* red = reinterpret_cast<float*>(bt + (size_t)taps * 2048);     // [2][8 warps][32] pool partials
  float* bias_sm = red + 2 * 8 * 32;                 // [64]
  DwTcBars* bars = reinterpret_cast<DwTcBars*>(bias_sm + 64);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x / ctas_per_chunk, cta = blockIdx.x % ctas_per_chunk;
  const int c0 = chunk * 64;
  const int t_begin = cta * tiles_per_cta, t_end = min(t_begin + tiles_per_cta, n_tiles);
  const int per_img = p.tiles_w * p.tiles_h;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], kDwTcIssuers);
      mbar_init(&bars->tmem_full[s], kDwTcIssuers);
      mbar_init(&bars->tmem_empty[s], kDwTcEpiWarps);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_x);
  }
  // zero the weight tiles (16-byte stores), then drop the diagonals in
  for (int i = tid; i < taps * 128; i += kDwTcThreads) reinterpret_cast<uint4*>(bt)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 64) bias_sm[tid] = c0 + tid < p.C ? bias[c0 + tid] : 0.f;
  __syncthreads();
  for (int i = tid; i < taps * 64; i += kDwTcThreads) {
    const int t = i >> 6, c = i & 63, blk = c >> 4, n = c & 15;
    const int k = blk * 16 + n;                       // column of the [16][64] tile that holds this block's diagonal
    const float wv = c0 + c < p.C ? w[(size_t)t * p.C + c0 + c] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(bt + (size_t)t * 2048 + (n >> 3) * 1024 + (n & 7) * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) =
        __float2bfloat16_rn(wv);
  }
  fence_proxy_async();
  if (warp == kDwTcEpiWarps) tmem_alloc(&bars->tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp >= kDwTcEpiWarps) {
    if (lane == 0) {
      const int blk = (warp - kDwTcEpiWarps) & 3;            // this issuer's 16-channel block
      const int g_first = (warp - kDwTcEpiWarps) >> 2;       // and its M groups: g_first, g_first + 2, ...
      const bool producer = warp == kDwTcEpiWarps;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t ba = smem_u32(bt);
      auto issue_load = [&](int t, int it) {
        const int s = it & 1;
        mbar_wait(&bars->empty[s], (uint32_t)(((it >> 1) & 1) ^ 1), 1);
        const int b = t / per_img, tin = t % per_img;
        const int ty = tin / p.tiles_w, tx = tin % p.tiles_w;
        mbar_expect_tx(&bars->full[s], (uint32_t)(p.THI * p.TWI) * 128u);
        tma_load_4d(xt + (size_t)s * xt_bytes, &tm_x, &bars->full[s], c0, tx * p.TW - p.pad, ty * p.TH - p.pad, b);
      };
      if (producer && t_begin < t_end) issue_load(t_begin, 0);
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int s = it & 1;
        const uint32_t ph = (uint32_t)((it >> 1) & 1);
        if (producer && t + 1 < t_end) issue_load(t + 1, it + 1);
        mbar_wait(&bars->full[s], ph, 2);
        mbar_wait(&bars->tmem_empty[s], ph ^ 1, 3);
        tc_fence_after();
        const uint32_t xa = smem_u32(xt + (size_t)s * xt_bytes);
        for (int ky = 0, tp = 0; ky < p.K; ++ky) {
          for (int kx = 0; kx < p.K; ++kx, ++tp) {
            const uint64_t db = dwtc_sw128_desc(ba + (uint32_t)tp * 2048u) + (uint64_t)(blk * 2);
            for (int g = g_first; g < p.MG; g += 2) {
              const uint64_t da = dwtc_sw128_desc(xa + (uint32_t)(g * 128 + ky * p.TWI + kx) * 128u) + (uint64_t)(blk * 2);
              umma_bf16(tmem_base + (uint32_t)(s * 128 + g * 64 + blk * 16), da, db, idesc, tp != 0);
            }
          }
        }
        umma_commit(&bars->empty[s]);
        umma_commit(&bars->tmem_full[s]);
      }
    }
  } else {
    // epilogue: warp = (TMEM lane quarter q, channel half); lane = position
    const int q = warp & 3, half = warp >> 2;
    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const int s = it & 1;
      const uint32_t ph = (uint32_t)((it >> 1) & 1);
      const int b = t / per_img, tin = t % per_img;
      const int ty = tin / p.tiles_w, tx = tin % p.tiles_w;
      const int y0 = ty * p.TH, x0 = tx * p.TW;
      mbar_wait(&bars->tmem_full[s], ph, 4);
      tc_fence_after();
      float psum[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) psum[j] = 0.f;
      for (int g = 0; g < p.MG; ++g) {
        const int pos = g * 128 + q * 32 + lane;
        const int py = pos / p.TWI, px = pos % p.TWI;
        const int oy = y0 + py, ox = x0 + px;
        const bool valid = pos < p.NP && px < p.TW && py < p.TH && oy < p.Ho && ox < p.Wo;
        uint32_t v0[16], v1[16];
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 128 + g * 64 + half * 32);
        tmem_ld16(ta, v0);
        tmem_ld16(ta + 16, v1);
        tmem_ld_wait();
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float a = __uint_as_float(j < 16 ? v0[j] : v1[j - 16]) + bias_sm[half * 32 + j];
          if (p.act == DFV_ACT_SILU) {
            const float h = 0.5f * a;
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
            a = fmaf(h, th, h);
          }
          o[j] = valid ? a : 0.f;
        }
        if (valid) {
          __nv_bfloat16* dst = y + (((size_t)b * p.Ho + oy) * p.Wo + ox) * p.C + c0 + half * 32;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            if (c0 + half * 32 + j8 * 8 < p.C) {
              uint4 pk;
              pk.x = pack_bf16(o[j8 * 8 + 0], o[j8 * 8 + 1]); pk.y = pack_bf16(o[j8 * 8 + 2], o[j8 * 8 + 3]);
              pk.z = pack_bf16(o[j8 * 8 + 4], o[j8 * 8 + 5]); pk.w = pack_bf16(o[j8 * 8 + 6], o[j8 * 8 + 7]);
              *reinterpret_cast<uint4*>(dst + j8 * 8) = pk;
            }
          }
        }
        if (pool_partial) {
#pragma unroll
          for (int j = 0; j < 32; ++j) psum[j] += __bfloat162float(__float2bfloat16_rn(o[j]));   // pool what the next layer reads
        }
      }
      // the accumulators are in registers: hand the TMEM buffer back before the (shuffle-heavy) pool reduction
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[s]);
      if (pool_partial) {
        float* rd = red + (size_t)s * 8 * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float sj = psum[j];
#pragma unroll
          for (int o_ = 16; o_ > 0; o_ >>= 1) sj += __shfl_xor_sync(0xffffffffu, sj, o_);
          if (lane == j) rd[warp * 32 + j] = sj;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");      // the eight epilogue warps
        if (tid < 64) {
          const int hf = tid >> 5, c = tid & 31;
          const float sm_ = (rd[(hf * 4 + 0) * 32 + c] + rd[(hf * 4 + 1) * 32 + c]) + (rd[(hf * 4 + 2) * 32 + c] + rd[(hf * 4 + 3) * 32 + c]);
          if (c0 + tid < p.C) pool_partial[((size_t)b * per_img + tin) * p.C + c0 + tid] = sm_;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kDwTcEpiWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static int dwtc_plan(DwTcParams& p, int H, int W, int C, int kernel, int pad_lo, int pad_hi) {
  p.C = C; p.H = H; p.W = W; p.K = kernel; p.pad = pad_lo;
  p.Ho = H + pad_lo + pad_hi - kernel + 1;
  p.Wo = W + pad_lo + pad_hi - kernel + 1;
  if (p.Ho <= 0 || p.Wo <= 0) return DFV_ERR_INVALID;
  p.TW = p.Wo + kernel - 1 <= 32 ? p.Wo : 12;
  p.TWI = p.TW + kernel - 1;
  p.TH = std::min(p.Ho, 256 / p.TWI);
  p.THI = p.TH + kernel - 1;
  p.tiles_w = (p.Wo + p.TW - 1) / p.TW;
  p.tiles_h = (p.Ho + p.TH - 1) / p.TH;
  p.NP = p.TH * p.TWI;
  p.MG = (p.NP + 127) / 128;
  return (p.MG >= 1 && p.MG <= 2 && p.TWI <= 256 && p.THI <= 256) ? DFV_OK : DFV_ERR_INVALID;
}

}  // namespace dfv

using namespace dfv;

/* Pool-partial slots per image of dfv_dwconv_tc_fwd (one per tile), or a negative error code. */
extern "C" int dfv_dwconv_tc_pool_parts(int H, int W, int C, int kernel, int pad_lo, int pad_hi) {
  DwTcParams p;
  if (dwtc_plan(p, H, W, C, kernel, pad_lo, pad_hi) != DFV_OK) return DFV_ERR_INVALID;
  return p.tiles_w * p.tiles_h;
}

/* Tensor-core depthwise convolution, stride 1, bf16 (experimental entry; see the header of this file).
 * x [B][H][W][C], w_kkc fp32 [k*k][C] (BN scale folded in), bias fp32 [C], y [B][Ho][Wo][C],
 * pool_partial fp32 [B][dfv_dwconv_tc_pool_parts][C] or NULL. */
extern "C" int dfv_dwconv_tc_fwd(const void* x, const float* w_kkc, const float* bias, void* y, float* pool_partial, int B, int H,
                                 int W, int C, int kernel, int pad_lo, int pad_hi, int act, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(x && w_kkc && bias && y, "dfv_dwconv_tc_fwd: null pointer");
  DFV_REQUIRE((kernel == 3 || kernel == 5) && B > 0 && C > 0 && C % 8 == 0, "dfv_dwconv_tc_fwd: bad shape");
  DwTcParams p;
  DFV_REQUIRE(dwtc_plan(p, H, W, C, kernel, pad_lo, pad_hi) == DFV_OK, "dfv_dwconv_tc_fwd: cannot tile H=%d W=%d k=%d", H, W, kernel);
  p.act = act;
  CUtensorMap tm;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {64, (uint32_t)p.TWI, (uint32_t)p.THI, 1};
  DFV_TRY(make_tensor_map(&tm, DFV_BF16, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  DFV_TRY(init_timeout_word_tu());
  const size_t smem = 2 * ((size_t)(((p.THI * p.TWI + 8) * 128 + 1023) / 1024) * 1024) + (size_t)kernel * kernel * 2048 + (2 * 8 * 32 + 64) * 4 +
                      sizeof(DwTcBars) + 64 + 1024;
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(dwconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  DFV_REQUIRE(smem <= 227 * 1024, "dfv_dwconv_tc_fwd: tile too large");
  ProfScope prof(PK_DWCONV, ((double)B * H * W * C + (double)B * p.Ho * p.Wo * C) * 2.0, 2.0 * kernel * kernel * (double)B * p.Ho * p.Wo * C,
                 as_stream(stream));
  // persistent grid: one CTA per SM, a fixed 64-channel chunk per CTA, contiguous tile ranges
  const int chunks = (C + 63) / 64;
  const int n_tiles = B * p.tiles_w * p.tiles_h;
  int cpc = std::max(1, num_sms() / chunks);
  if (cpc > n_tiles) cpc = n_tiles;
  const int tpc = (n_tiles + cpc - 1) / cpc;
  cpc = (n_tiles + tpc - 1) / tpc;
  dwconv_tc_kernel<<<(unsigned)(chunks * cpc), kDwTcThreads, smem, as_stream(stream)>>>(tm, w_kkc, bias, (__nv_bfloat16*)y, pool_partial, p, cpc,
                                                                                      tpc, n_tiles);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}
