// Stem: static pad (0,1,0,1) + conv3x3 stride 2 (3 -> 48) + folded BN + swish.
// Reads the reference's NCHW fp32 input contract directly and writes NHWC.
// HBM-bound by bytes (1.7 MB in / 3.5 MB out per image at 380 in bf16) but carries
// 1296 FMA per output pixel, so it sits near the FP32-pipe / HBM crossover.
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace dfv {

constexpr int kStemC = 48;

// Input contract of the reference (src/data/dataset.py:82-116): NCHW fp32, already normalised -- or, with kU8, the raw
// crop: uint8 RGB in HWC order, normalised here on the operand path with the reference's own arithmetic
//   v = (u8 / 255 - mean[c]) / std[c]        (fp32, IEEE division; dataset.py:95-98, task.ipynb:386)
// Only 3 x 256 values exist, so every CTA tabulates them once (exact, op for op) and the conv reads the table.
struct StemNorm {
  float mean[3], std[3];
};
__device__ __forceinline__ void stem_build_lut(float* lut, const StemNorm& nm) {
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int c = i >> 8;
    lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.0f), nm.mean[c]), nm.std[c]);
  }
  if (threadIdx.x == 0) lut[768] = 0.f;      // pad taps of the tensor-core stem index this entry
}
template <bool kU8>
__device__ __forceinline__ float stem_load(const void* x, const float* lut, int b, int ci, int hi, int wi, int H, int W) {
  if constexpr (kU8) {
    const unsigned char u = __ldg(reinterpret_cast<const unsigned char*>(x) + (((size_t)b * H + hi) * W + wi) * 3 + ci);
    return lut[ci * 256 + u];
  } else {
    return __ldg(reinterpret_cast<const float*>(x) + (((size_t)b * 3 + ci) * H + hi) * W + wi);
  }
}

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

template <typename T, bool kFast, bool kU8>
__global__ void __launch_bounds__(256, 3) stem_kernel(const void* __restrict__ x, const StemNorm nm, const float* __restrict__ w,
                                                  const float* __restrict__ bias, T* __restrict__ y, int B, int H,
                                                  int W, int Ho, int Wo, int act) {
  __shared__ __align__(16) float ws[27 * kStemC];
  __shared__ __align__(16) float bs[kStemC];
  __shared__ float lut[kU8 ? 772 : 1];
  for (int i = threadIdx.x; i < 27 * kStemC; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < kStemC) bs[threadIdx.x] = bias[threadIdx.x];
  if constexpr (kU8) stem_build_lut(lut, nm);
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

#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int hi = 2 * ho + kh;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      float in[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) in[i] = (hi < H && 2 * wo0 + i < W) ? stem_load<kU8>(x, lut, b, ci, hi, 2 * wo0 + i, H, W) : 0.f;
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

// ------------------------------------------------------------------------------------ tcgen05 stem (bf16)
// The stem is the network's first dense contraction: out[pixel][48] = im2col(x)[pixel][27] . W[27][48].
//   warps 0-3  build the im2col A tile (128 output pixels x 32 taps, bf16, 128-byte-swizzled K-major rows) straight
//              from the NCHW fp32 images (static pad (0,1,0,1) = bounds checks), two stages
//   warp  8    issues two tcgen05.mma (M 128, N 48, K 16) per tile, fp32 accumulators in TMEM (two stages)
//   warps 4-7  epilogue: tcgen05.ld -> bias -> swish -> bf16 -> shared staging -> one 12 KB bulk copy per tile
//              (128 consecutive NHWC pixels are contiguous in global memory)
// The SIMT kernel above needs 1296 FFMA per pixel (0.33 ms of FP32 pipe at batch 256 before any memory traffic).
constexpr int kStemTcThreads = 288;
constexpr int kStemTile = 128;
constexpr uint32_t kStemTmemCols = 128;
constexpr int kLutPitch = 264;

struct __align__(8) StemBars {
  uint64_t a_full[2], a_empty[2], t_full[2], t_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t stem_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <bool kAct, bool kU8>
__global__ void __launch_bounds__(kStemTcThreads, 2)
    stem_tc_kernel(const void* __restrict__ x, const StemNorm nm, const float* __restrict__ w, const float* __restrict__ bias,
                   __nv_bfloat16* __restrict__ y, int B, int H, int W, int Ho, int Wo, long long tiles_per_cta) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* a_tiles = smem;                           // 2 x 16 KB
  unsigned char* b_tile = smem + 2 * 16384;                // 48 rows x 128 B (padded to 8 KB)
  unsigned char* staging = b_tile + 8192;                  // 2 x 12 KB
  float* bias_sm = reinterpret_cast<float*>(staging + 2 * 12288);
  StemBars* bars = reinterpret_cast<StemBars*>(bias_sm + 64);
  float* lut = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + ((sizeof(StemBars) + 15) / 16) * 16);   // [3][256], kU8 only
  // [3][kLutPitch]: entries 0..255 the normalised values of channel c, entry 256 = 0 for pad taps -- the builders keep the raw
  // byte (or 256) in a register and the channel offset rides in the load's immediate
  if constexpr (kU8) {
    for (int i = threadIdx.x; i < 3 * kLutPitch; i += blockDim.x) {
      const int c = i / kLutPitch, u = i % kLutPitch;
      lut[i] = u < 256 ? __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), nm.mean[c]), nm.std[c]) : 0.f;
    }
  }

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long total = (long long)B * Ho * Wo;
  const long long n_tiles = (total + kStemTile - 1) / kStemTile;
  const long long t_begin = (long long)blockIdx.x * tiles_per_cta;
  const long long t_end = min(t_begin + tiles_per_cta, n_tiles);

  // B operand: W^T [48][32] bf16 (k = (kh*3 + kw)*3 + ci, as the weight blob is laid out), taps 27..31 zero
  for (int i = tid; i < kStemC * 4; i += blockDim.x) {
    const int n = i >> 2, c = i & 3;
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k0 = c * 8 + 2 * j;
      const float lo = k0 < 27 ? w[k0 * kStemC + n] : 0.f, hi = k0 + 1 < 27 ? w[(k0 + 1) * kStemC + n] : 0.f;
      pk[j] = pack_bf16(lo, hi);
    }
    *reinterpret_cast<uint4*>(b_tile + n * 128 + ((c ^ (n & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  if (tid < kStemC) bias_sm[tid] = kAct ? 0.5f * bias[tid] : bias[tid];
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->a_full[s], 128);
      mbar_init(&bars->a_empty[s], 1);
      mbar_init(&bars->t_full[s], 1);
      mbar_init(&bars->t_empty[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(&bars->tmem_base, kStemTmemCols);
  fence_proxy_async();      // the weight tile was written with ordinary stores; the tensor core reads it through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp < 4) {
    // ------------------------------------------------------------- im2col builders
    // One thread per output pixel: 27 gathered taps -> one 64-byte (32 x bf16, taps 27..31 zero) swizzled A row.
    // The gather of tile t+1 is ISSUED before tile t is packed (two register sets, loop unrolled by two): with the loads
    // and the pack in sequence the builders sat on the load latency for ~40 % of all samples (long_scoreboard at the
    // first use) and the kernel ran at 0.41 of the HBM roofline.  uint8 input: the registers hold the raw bytes
    // (256 = the zero entry, for pad taps); the look-up happens at pack time.
    using RawT = typename std::conditional<kU8, uint32_t, float>::type;
    // (b, ho, wo) of this thread's pixel in the tile about to be gathered: ONE 64-bit division at the start, then advanced
    // by 128 pixels per tile (the four 64-bit divisions per pixel and tile were ~250 of the builders' ~400 instructions,
    // and the builders' issue rate is the kernel's pace: two threads per pixel ran 1.6x SLOWER, an FFMA instead of the table
    // look-up 1.17x slower)
    int g_wo, g_ho, g_b;
    {
      const long long p0 = t_begin * kStemTile + tid;
      g_wo = (int)(p0 % Wo);
      const long long r0 = p0 / Wo;
      g_ho = (int)(r0 % Ho);
      g_b = (int)(r0 / Ho);
    }
    auto gather = [&](long long t, RawT (&v)[27]) {
      const long long p = t * kStemTile + tid;
#pragma unroll
      for (int i = 0; i < 27; ++i) v[i] = kU8 ? (RawT)256 : (RawT)0;
      const int wo = g_wo, ho = g_ho, b = g_b;
      g_wo += kStemTile;                                // the next tile's pixel
      while (g_wo >= Wo) {
        g_wo -= Wo;
        if (++g_ho == Ho) { g_ho = 0; ++g_b; }
      }
      if (p < total) {
        if (2 * ho + 2 < H && 2 * wo + 2 < W) {
          // interior pixel (all but the last output row / column): no bounds checks, one row pointer per (channel, kernel row)
          // and immediate offsets for the taps -- the per-tap 64-bit address arithmetic and predicates were ~130 of the ~260
          // instructions a builder thread spent per tile
          if constexpr (kU8) {
            const unsigned char* p0 = reinterpret_cast<const unsigned char*>(x) + (((size_t)b * H + 2 * ho) * W + 2 * wo) * 3;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const unsigned char* pr = p0 + (size_t)kh * W * 3;
#pragma unroll
              for (int j = 0; j < 9; ++j) v[kh * 9 + j] = (RawT)__ldg(pr + j);      // j = kw * 3 + ci
            }
          } else {
            const size_t plane = (size_t)H * W;
            const float* p0 = reinterpret_cast<const float*>(x) + ((size_t)b * 3 * H + 2 * ho) * W + 2 * wo;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
                const float* pr = p0 + ci * plane + (size_t)kh * W;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) v[(kh * 3 + kw) * 3 + ci] = __ldg(pr + kw);
              }
          }
        } else {
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const int hi = 2 * ho + kh;
#pragma unroll
              for (int kw = 0; kw < 3; ++kw)
                if (hi < H && 2 * wo + kw < W) {
                  if constexpr (kU8)
                    v[(kh * 3 + kw) * 3 + ci] = (RawT)__ldg(reinterpret_cast<const unsigned char*>(x) + (((size_t)b * H + hi) * W + 2 * wo + kw) * 3 + ci);
                  else
                    v[(kh * 3 + kw) * 3 + ci] = __ldg(reinterpret_cast<const float*>(x) + (((size_t)b * 3 + ci) * H + hi) * W + 2 * wo + kw);
                }
            }
        }
      }
    };
    auto emit = [&](int it, const RawT (&v)[27]) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&bars->a_empty[s], ph ^ 1, 31);
      unsigned char* arow = a_tiles + s * 16384 + tid * 128;
      auto val = [&](int k) -> float {
        if (k >= 27) return 0.f;
        if constexpr (kU8) return lut[(k % 3) * kLutPitch + v[k < 27 ? k : 0]]; else return v[k < 27 ? k : 0];
      };
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 pk;
        pk.x = pack_bf16(val(c * 8 + 0), val(c * 8 + 1));
        pk.y = pack_bf16(val(c * 8 + 2), val(c * 8 + 3));
        pk.z = pack_bf16(val(c * 8 + 4), val(c * 8 + 5));
        pk.w = pack_bf16(val(c * 8 + 6), val(c * 8 + 7));
        *reinterpret_cast<uint4*>(arow + ((c ^ (tid & 7)) << 4)) = pk;
      }
      fence_proxy_async();
      mbar_arrive(&bars->a_full[s]);
    };
    RawT va[27], vb[27];
    long long t = t_begin;
    int it = 0;
    if (t < t_end) gather(t, va);
    while (t < t_end) {
      if (t + 1 < t_end) gather(t + 1, vb);
      emit(it, va);
      ++t; ++it;
      if (t >= t_end) break;
      if (t + 1 < t_end) gather(t + 1, va);
      emit(it, vb);
      ++t; ++it;
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kStemC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t db = stem_sw128_desc(smem_u32(b_tile));
      int it = 0;
      for (long long t = t_begin; t < t_end; ++t, ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&bars->t_empty[s], ph ^ 1, 32);
        mbar_wait(&bars->a_full[s], ph, 33);
        tc_fence_after();
        const uint64_t da = stem_sw128_desc(smem_u32(a_tiles + s * 16384));
        const uint32_t d_tmem = tmem_base + (uint32_t)s * 64;
        umma_bf16(d_tmem, da, db, idesc, 0);
        umma_bf16(d_tmem, da + 2, db + 2, idesc, 1);
        umma_commit(&bars->a_empty[s]);
        umma_commit(&bars->t_full[s]);
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue (warps 4..7 <-> TMEM lane quarters 0..3)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool leader = (warp == 4 && lane == 0);
    int it = 0;
    for (long long t = t_begin; t < t_end; ++t, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      unsigned char* stg = staging + s * 12288;
      if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // staging[s]'s previous copy has been read
      __syncwarp();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&bars->t_full[s], ph, 34);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)s * 64;
      uint32_t v[3][16];
      tmem_ld16(tbase, v[0]);
      tmem_ld16(tbase + 16, v[1]);
      tmem_ld16(tbase + 32, v[2]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->t_empty[s]);
#pragma unroll
      for (int g = 0; g < 6; ++g) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = __uint_as_float(v[g >> 1][(g & 1) * 8 + j]);
          if constexpr (kAct) {
            const float h = fmaf(a, 0.5f, bias_sm[g * 8 + j]);
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
            o[j] = fmaf(h, th, h);
          } else {
            o[j] = a + bias_sm[g * 8 + j];
          }
        }
        uint4 pk;
        pk.x = pack_bf16(o[0], o[1]); pk.y = pack_bf16(o[2], o[3]);
        pk.z = pack_bf16(o[4], o[5]); pk.w = pack_bf16(o[6], o[7]);
        *reinterpret_cast<uint4*>(stg + row * 96 + g * 16) = pk;
      }
      fence_proxy_async();
      __syncwarp();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (leader) {
        const long long p0 = t * kStemTile;
        const long long npx = min((long long)kStemTile, total - p0);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(reinterpret_cast<uint64_t>(y + (size_t)p0 * kStemC)), "r"(smem_u32(stg)), "r"((uint32_t)(npx * kStemC * 2))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      __syncwarp();
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kStemTmemCols);
  }
}

template <bool kU8>
static int launch_stem_tc(const void* x, const StemNorm& nm, const float* w, const float* bias, void* y, int B, int H, int W, int Ho, int Wo,
                          int act, cudaStream_t st) {
  const long long total = (long long)B * Ho * Wo;
  const long long n_tiles = (total + kStemTile - 1) / kStemTile;
  long long grid = std::min<long long>(n_tiles, 2LL * num_sms());
  const long long tpc = (n_tiles + grid - 1) / grid;
  grid = (n_tiles + tpc - 1) / tpc;
  const size_t smem = 2 * 16384 + 8192 + 2 * 12288 + 256 + sizeof(StemBars) + 16 + (kU8 ? 3 * kLutPitch * 4 : 0) + 1024;
  DFV_TRY(init_timeout_word_tu());
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(stem_tc_kernel<true, kU8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    DFV_CUDA(cudaFuncSetAttribute(stem_tc_kernel<false, kU8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  if (act)
    stem_tc_kernel<true, kU8><<<(unsigned)grid, kStemTcThreads, smem, st>>>(x, nm, w, bias, (__nv_bfloat16*)y, B, H, W, Ho, Wo, tpc);
  else
    stem_tc_kernel<false, kU8><<<(unsigned)grid, kStemTcThreads, smem, st>>>(x, nm, w, bias, (__nv_bfloat16*)y, B, H, W, Ho, Wo, tpc);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

// uint8 HWC -> normalised fp32 NCHW (the reference's input contract), same table as the fused stem.
// Thread = 4 consecutive pixels of one image: three 32-bit loads (12 bytes), three float4 stores (one per channel plane).
__global__ void __launch_bounds__(256) u8_to_nchw_kernel(const unsigned char* __restrict__ x, const StemNorm nm, float* __restrict__ y,
                                                        int B, int H, int W) {
  __shared__ float lut[772];
  stem_build_lut(lut, nm);
  __syncthreads();
  const long long hw = (long long)H * W;
  if ((hw & 3) == 0) {
    const long long quads = (long long)B * (hw >> 2);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
      const long long b = q / (hw >> 2), p0 = (q % (hw >> 2)) << 2;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(x + (b * hw + p0) * 3);   // 12-byte groups: 4-byte aligned
      const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
      unsigned char by[12];
#pragma unroll
      for (int i = 0; i < 4; ++i) { by[i] = (w0 >> (8 * i)) & 255; by[4 + i] = (w1 >> (8 * i)) & 255; by[8 + i] = (w2 >> (8 * i)) & 255; }
#pragma unroll
      for (int c = 0; c < 3; ++c)
        *reinterpret_cast<float4*>(y + (b * 3 + c) * hw + p0) =
            make_float4(lut[c * 256 + by[c]], lut[c * 256 + by[3 + c]], lut[c * 256 + by[6 + c]], lut[c * 256 + by[9 + c]]);
    }
  } else {
    const long long total = (long long)B * hw;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
      const long long b = p / hw, r = p % hw;
      const unsigned char* px = x + p * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) y[(b * 3 + c) * hw + r] = lut[c * 256 + px[c]];
    }
  }
}

template <bool kU8>
static int stem_entry(const void* x, const float* norm6, const float* w, const float* bias, void* y, int dtype, int B, int H, int W, int C,
                      int act, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(x && w && bias && y && (!kU8 || norm6), "dfv_stem_conv_fwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype), "dfv_stem_conv_fwd: bad dtype %d", dtype);
  DFV_REQUIRE(C == kStemC, "dfv_stem_conv_fwd: C must be %d (EfficientNet-B4 stem), got %d", kStemC, C);
  DFV_REQUIRE(B > 0 && H >= 3 && W >= 3, "dfv_stem_conv_fwd: bad shape B=%d H=%d W=%d", B, H, W);
  if (debug_flags() & 4) return DFV_OK;
  StemNorm nm = {{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
  if (kU8) {
    for (int c = 0; c < 3; ++c) {
      nm.mean[c] = norm6[c];
      nm.std[c] = norm6[3 + c];
      DFV_REQUIRE(nm.std[c] > 0.f, "dfv_stem_conv_u8_fwd: std[%d] must be positive", c);
    }
  }
  const int Ho = (H + 1 - 3) / 2 + 1, Wo = (W + 1 - 3) / 2 + 1;
  const long long total = (long long)B * Ho * Wo;
  const long long threads = (long long)B * Ho * ((Wo + 3) / 4) * 4;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  ProfScope prof(PK_STEM, (double)B * 3 * H * W * (kU8 ? 1 : 4) + (double)total * kStemC * dtype_size(dtype),
                 2.0 * 27 * kStemC * (double)total, as_stream(stream));
  if (dtype == DFV_BF16) return launch_stem_tc<kU8>(x, nm, w, bias, y, B, H, W, Ho, Wo, act, as_stream(stream));
  stem_kernel<float, false, kU8><<<grid, 256, 0, as_stream(stream)>>>(x, nm, w, bias, (float*)y, B, H, W, Ho, Wo, act);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_stem_conv_fwd(const float* x, const float* w, const float* bias, void* y, int dtype, int B, int H,
                                 int W, int C, int act, dfv_stream_t stream) {
  return stem_entry<false>(x, nullptr, w, bias, y, dtype, B, H, W, C, act, stream);
}

extern "C" int dfv_stem_conv_u8_fwd(const uint8_t* x, const float* norm6, const float* w, const float* bias, void* y, int dtype,
                                    int B, int H, int W, int C, int act, dfv_stream_t stream) {
  return stem_entry<true>(x, norm6, w, bias, y, dtype, B, H, W, C, act, stream);
}

extern "C" int dfv_u8_to_nchw_f32(const uint8_t* x, const float* norm6, float* y, int B, int H, int W, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(x && norm6 && y && B > 0 && H > 0 && W > 0, "dfv_u8_to_nchw_f32: bad arguments");
  StemNorm nm;
  for (int c = 0; c < 3; ++c) {
    nm.mean[c] = norm6[c];
    nm.std[c] = norm6[3 + c];
    DFV_REQUIRE(nm.std[c] > 0.f, "dfv_u8_to_nchw_f32: std[%d] must be positive", c);
  }
  DFV_REQUIRE((reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "dfv_u8_to_nchw_f32: unaligned buffers");
  const long long total = (long long)B * H * W;
  const unsigned grid = (unsigned)std::min<long long>((total / 4 + 255) / 256 + 1, 148 * 8);
  u8_to_nchw_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, nm, y, B, H, W);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}
