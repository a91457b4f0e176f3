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
#include <cstdlib>

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
  long long tiles_per_cta;   // contiguous tile range per CTA
  int parts;                 // pool-partial slots per image
  int pad;         // pad_lo (top == left)
  int act;
  int nthreads;    // = G * strips * rpr: thread -> (channel group, strip, row) is fixed for the whole kernel
  int strips;      // TW / L
  int rpr;         // tile rows per round
  int rounds;      // ceil(TH / rpr)
  int red_parts;   // second-stage partials of the pool reduction
};

// Squeeze-excite, first half, fused into the depthwise kernel's tail (inference): every CTA turns the pool sums of ITS
// channel chunk and ITS images into partial hidden sums  sum_{c in chunk} w1[j][c] * mean[b][c]  (the squeeze layer is
// linear, so partial channel sums are enough) and ADDS them into hid[b][j] with 64-bit FIXED-POINT atomics (2^-30
// resolution): integer addition is associative, so the result does not depend on the order in which CTAs arrive --
// bit-reproducible without tickets, finalize passes or any CTA waiting on another.  What is left of the gate is one
// small launch (the excite layer, se.cu) instead of three dependent ones.  The accumulator of the NEXT layer is zeroed
// in this kernel's prologue (two buffers alternate; nobody reads the other one while this kernel runs).
struct SeFuse {
  const float* w1;          // [sq][C] squeeze weight; nullptr = not fused
  long long* hid_fix;       // [B][sq] fixed-point accumulators, zero on entry
  long long* zero_next;     // buffer to zero for the next fused layer (zero_count entries), may be nullptr
  long long zero_count;
  int sq;
  float inv_hw;
};
constexpr int kSeGroup = 32;             // images per tail round (bounds the tail's shared memory)
constexpr float kSeFixScale = 1073741824.0f;   // 2^30

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

__device__ __forceinline__ float tanh_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}

// Activation + pool sum + store of one output pixel (8 channels).  kHalf: the accumulators hold x / 2 (the 1/2
// of swish(x) = h * tanh(h) + h, h = x / 2, is folded into the staged weights and bias).
template <typename T, bool kAct, bool kHalf, bool kStats>
__device__ __forceinline__ void finish_pixel(const float acc[8], float2 psum[4], float2 psq[4], T* out) {
  if constexpr (kAct && kHalf) {
    uint4 pk;
    uint32_t* pw = &pk.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 h = make_float2(acc[2 * j], acc[2 * j + 1]);
      const float2 t = make_float2(tanh_fast(h.x), tanh_fast(h.y));
      const float2 o = ffma2(h, t, h);
      psum[j] = fadd2(psum[j], o);
      pw[j] = pack_bf16(o.x, o.y);
    }
    if constexpr (sizeof(T) == 2) *reinterpret_cast<uint4*>(out) = pk;
  } else {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = kAct ? silu<false>(acc[e]) : acc[e];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      psum[j].x += o[2 * j];
      psum[j].y += o[2 * j + 1];
      if constexpr (kStats) {
        psq[j].x = fmaf(o[2 * j], o[2 * j], psq[j].x);
        psq[j].y = fmaf(o[2 * j + 1], o[2 * j + 1], psq[j].y);
      }
    }
    store8(out, o);
  }
}

// The fused squeeze tail (see SeFuse), deliberately NOT inlined: it recomputes the CTA's tile range from blockIdx so that
// none of its state is live across the depthwise main loop (inlined, it cost that loop 7-8 % in spills).
// The tile buffers are free when it runs: [ws: sq x (CB+1)] [ps: kSeGroup x CB].
struct SeTailArgs {       // passed BY VALUE (scalars only): taking the kernel parameter structs' addresses would move them to local memory
  const float* w1;
  long long* hid_fix;
  const float* pool_partial;
  long long tiles_per_cta, per_chunk;
  int sq, C, chunks, parts, n_tiles, CB;
  float inv_hw;
};
__device__ __noinline__ void se_tail(unsigned char* smem_raw, const SeTailArgs a) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const int CB = a.CB;
  const int chunk = blockIdx.x % a.chunks, slot = blockIdx.x / a.chunks;
  const int c0 = chunk * CB;
  const int n_tiles = a.n_tiles;
  const long long t_begin = (long long)slot * a.tiles_per_cta;
  const long long t_end = min(t_begin + a.tiles_per_cta, a.per_chunk);
  if (t_begin >= t_end) return;
  __syncthreads();
  const float* __restrict__ pool_partial = a.pool_partial;
  struct { const float* w1; long long* hid_fix; int sq; float inv_hw; } se = {a.w1, a.hid_fix, a.sq, a.inv_hw};
  struct { int C, parts; long long tiles_per_cta; } p = {a.C, a.parts, a.tiles_per_cta};
  const int sq = se.sq, WP = CB + 1;
  float* ws = reinterpret_cast<float*>(smem_raw);
  float* ps = ws + (((size_t)sq * WP + 3) & ~(size_t)3);      // 16-byte aligned: read as float4
  for (int i = tid; i < sq * CB; i += nth) {
    const int jj = i / CB, o = i % CB;
    ws[jj * WP + o] = (c0 + o < p.C) ? __ldg(se.w1 + (size_t)jj * p.C + c0 + o) : 0.f;
  }
  const int b_first = (int)(t_begin / n_tiles), b_last = (int)((t_end - 1) / n_tiles);
  for (int g0 = b_first; g0 <= b_last; g0 += kSeGroup) {
    const int ng = min(kSeGroup, b_last - g0 + 1);
    for (int i = tid; i < ng * CB; i += nth) {       // this CTA's own pool sums (written by the main loop, visible after the barrier)
      const int li = i / CB, o = i % CB, bb = g0 + li;
      const long long first_cta = ((long long)bb * n_tiles) / p.tiles_per_cta;
      ps[i] = (c0 + o < p.C) ? pool_partial[((size_t)bb * p.parts + (size_t)(slot - first_cta)) * p.C + c0 + o] * se.inv_hw : 0.f;
    }
    __syncthreads();
    // thread = (squeeze row jj, image pair): the weight row is read once for two images; the pool means are
    // 16-byte broadcast reads
    const int npair = (ng + 1) >> 1;
    for (int idx = tid; idx < npair * sq; idx += nth) {
      const int lp = idx / sq, jj = idx - lp * sq;
      const int l0 = 2 * lp, l1 = min(2 * lp + 1, ng - 1);
      const float* wr = ws + jj * WP;
      const float4* p0 = reinterpret_cast<const float4*>(ps + l0 * CB);
      const float4* p1 = reinterpret_cast<const float4*>(ps + l1 * CB);
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 4
      for (int o = 0; o < CB; o += 4) {
        const float4 u = p0[o >> 2], v = p1[o >> 2];
        const float w0 = wr[o], w1 = wr[o + 1], w2 = wr[o + 2], w3 = wr[o + 3];
        a0 = fmaf(w0, u.x, a0); a1 = fmaf(w1, u.y, a1); a0 = fmaf(w2, u.z, a0); a1 = fmaf(w3, u.w, a1);
        b0 = fmaf(w0, v.x, b0); b1 = fmaf(w1, v.y, b1); b0 = fmaf(w2, v.z, b0); b1 = fmaf(w3, v.w, b1);
      }
      atomicAdd(reinterpret_cast<unsigned long long*>(se.hid_fix + (size_t)(g0 + l0) * sq + jj),
                (unsigned long long)__float2ll_rn((a0 + a1) * kSeFixScale));
      if (l1 != l0)
        atomicAdd(reinterpret_cast<unsigned long long*>(se.hid_fix + (size_t)(g0 + l1) * sq + jj),
                  (unsigned long long)__float2ll_rn((b0 + b1) * kSeFixScale));
    }
    __syncthreads();
  }
}

// Persistent CTA: a CONTIGUOUS range of the (image, tile row, tile column) tiles of ONE channel chunk, with a
// 2-deep TMA pipeline: the tile for step i+1 is in flight while step i is computed.  A thread keeps its
// (channel group, strip, first row) for the whole kernel; consecutive tiles belong to the same image for long
// runs, so the SE pool sums stay in registers across tiles and are reduced (shared memory, deterministic)
// only when the image changes.  Image b's partial sums land in slot (cta - first cta touching b); the
// last CTA of an image zero-fills the unused slots, so the SE-gate kernel simply sums `parts` rows.
// kStats (training, kAct = false): per-channel sum / sum of squares of the raw conv output, accumulated in registers
// over the whole kernel and added (double atomics, one per channel per CTA) into stats[2C] -- the BatchNorm
// batch statistics without a separate pass over the output.
template <typename T, int K, int S, int L, bool kFast, bool kAct, int kCB, bool kStats>
__global__ void __launch_bounds__(256, 2) dwconv_kernel(const __grid_constant__ CUtensorMap tmap,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       T* __restrict__ y, float* __restrict__ pool_partial,
                                                       double* __restrict__ stats, DwParams p, SeFuse se) {
  constexpr bool kHalf = kFast && kAct && sizeof(T) == 2;
  const int CB = kCB ? kCB : p.CB;     // compile-time for the full-width chunk: shared-memory offsets become immediates
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // [tile0][tile1][weights: K*K*CB T][bias: CB f32][red: nthreads*8 f32][red2: 4*CB f32][mbar x2]
  const size_t tile_bytes = (size_t)p.THI * p.TWI * CB * sizeof(T);
  const size_t tile_stride = ((tile_bytes + 127) / 128) * 128;
  T* wsm = reinterpret_cast<T*>(smem_raw + 2 * tile_stride);
  float* bsm = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(wsm) + (((size_t)K * K * CB * sizeof(T) + 15) / 16) * 16);
  float* red = bsm + CB;
  float* red2 = red + (size_t)p.nthreads * 8;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(red2 + 4 * CB);

  const int tid = threadIdx.x;
  const int chunk = blockIdx.x % p.chunks, slot = blockIdx.x / p.chunks;
  const int c0 = chunk * CB;
  const int n_tiles = p.tiles_w * p.tiles_h;
  const long long t_begin = (long long)slot * p.tiles_per_cta;
  const long long t_end = min(t_begin + p.tiles_per_cta, p.per_chunk);

  // tile coordinates of this CTA's current tile
  int b = (int)(t_begin / n_tiles);
  int th_i = (int)((t_begin % n_tiles) / p.tiles_w), tw_i = (int)(t_begin % p.tiles_w);
  auto issue = [&](int tw, int th, int bb, int buf) {
    mbar_expect_tx(&mbar[buf], (uint32_t)tile_bytes);
    tma_load_4d(smem_raw + buf * tile_stride, &tmap, &mbar[buf], c0, tw * p.TW * S - p.pad, th * p.TH * S - p.pad, bb);
  };

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
    if (t_begin < t_end) issue(tw_i, th_i, b, 0);
  }
  // Stage this chunk's weights (BN scale already folded in) and bias while the first tile lands.
  const float fold = kHalf ? 0.5f : 1.0f;
  for (int i = tid; i < K * K * CB; i += blockDim.x) {
    const int c = c0 + i % CB;
    const float v = (c < p.C) ? fold * w[(size_t)(i / CB) * p.C + c] : 0.f;
    if constexpr (sizeof(T) == 2) wsm[i] = __float2bfloat16_rn(v); else wsm[i] = v;
  }
  for (int i = tid; i < CB; i += blockDim.x) bsm[i] = (c0 + i < p.C) ? fold * bias[c0 + i] : 0.f;
  if constexpr (kAct && !kStats) {
    if (se.zero_next != nullptr)
      for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < se.zero_count; i += (long long)gridDim.x * blockDim.x) se.zero_next[i] = 0;
    if (se.w1 != nullptr) {      // the squeeze weights of this chunk are HBM-cold: pull them into L2 now, the tail reads them
      for (int i = tid; i < se.sq * ((CB * 4 + 127) / 128); i += blockDim.x) {
        const int jj = i / ((CB * 4 + 127) / 128), l = i % ((CB * 4 + 127) / 128);
        if (c0 + l * 32 < p.C) asm volatile("prefetch.global.L2 [%0];" ::"l"(se.w1 + (size_t)jj * p.C + c0 + l * 32));
      }
    }
  }
  __syncthreads();

  // fixed per-thread role
  const int G = CB >> 3;
  const int g = tid % G;
  const int j = (tid / G) % p.strips;
  const int r0 = tid / (G * p.strips);
  const int c = c0 + g * 8;
  const bool chan_ok = c < p.C;
  const int in_off0 = ((r0 * S) * p.TWI + j * L * S) * CB + g * 8;
  const int in_step = p.rpr * S * p.TWI * CB;
  const int out_off0 = (r0 * p.Wo + j * L) * p.C + g * 8;
  const int out_step = p.rpr * p.Wo * p.C;
  const int row_stride = p.TWI * CB;
  const T* wbase = wsm + g * 8;
  float bv[8];
  load8(bsm + g * 8, bv);
  constexpr int NI = (L - 1) * S + K;  // input window per kernel row
  const int nth = blockDim.x;

  float2 psum[4], psq[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) psum[e] = psq[e] = make_float2(0.f, 0.f);

  int it = 0;
  for (long long t = t_begin; t < t_end; ++t, ++it) {
    const int buf = it & 1;
    int ntw = tw_i + 1, nth_i = th_i, nb = b;
    if (ntw == p.tiles_w) { ntw = 0; ++nth_i; }
    if (nth_i == p.tiles_h) { nth_i = 0; ++nb; }
    if (tid == 0 && t + 1 < t_end) issue(ntw, nth_i, nb, buf ^ 1);   // prefetch the next tile
    const int h0 = th_i * p.TH, w0 = tw_i * p.TW;
    const int hrem = p.Ho - h0;                    // valid rows in this tile
    const int wrem = p.Wo - (w0 + j * L);          // valid pixels from this thread's strip start
    const T* tile = reinterpret_cast<const T*>(smem_raw + buf * tile_stride);
    T* out_tile = y + (((size_t)b * p.Ho + h0) * p.Wo + w0) * p.C + c0;
    mbar_wait(&mbar[buf], (it >> 1) & 1, 6);

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
          for (int kw = 0; kw < K; ++kw) ldvec(wbase + (kh * K + kw) * CB, wk[kw]);
          const T* row = in + kh * row_stride;
#pragma unroll
          for (int iw = 0; iw < NI; ++iw) {
            Vec8<T> v;
            ldvec(row + iw * CB, v);
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
          for (int l = 0; l < L; ++l) finish_pixel<T, kAct, kHalf, kStats>(acc[l], psum, psq, out + (size_t)l * p.C);
        } else {
#pragma unroll
          for (int l = 0; l < L; ++l)
            if (l < wrem) finish_pixel<T, kAct, kHalf, kStats>(acc[l], psum, psq, out + (size_t)l * p.C);
        }
      }
    }

    const bool flush = pool_partial != nullptr && (nb != b || t + 1 == t_end);
    if (flush) {
      // deterministic two-stage CTA reduction over the threads that share a channel group (tid = g + G*i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        red[tid * 8 + 2 * e] = psum[e].x;
        red[tid * 8 + 2 * e + 1] = psum[e].y;
        psum[e] = make_float2(0.f, 0.f);
      }
      __syncthreads();   // also: everyone is done with tile[buf] before it is refilled
      const int R = p.red_parts;
      for (int idx = tid; idx < CB * R; idx += nth) {
        const int o = idx % CB, q = idx / CB;
        const int gg = o >> 3, e = o & 7;
        float s = 0.f;
        for (int u = gg + G * q; u < nth; u += G * R) s += red[u * 8 + e];
        red2[idx] = s;
      }
      __syncthreads();
      const long long first_cta = ((long long)b * n_tiles) / p.tiles_per_cta;
      const long long last_cta = ((long long)(b + 1) * n_tiles - 1) / p.tiles_per_cta;
      const int my_slot = (int)(slot - first_cta);
      for (int o = tid; o < CB; o += nth) {
        if (c0 + o < p.C) {
          float s = 0.f;
          for (int q = 0; q < R; ++q) s += red2[q * CB + o];
          float* dst = pool_partial + (size_t)b * p.parts * p.C + c0 + o;
          dst[(size_t)my_slot * p.C] = s;
          if (slot == last_cta)
            for (int z = my_slot + 1; z < p.parts; ++z) dst[(size_t)z * p.C] = 0.f;
        }
      }
    } else {
      __syncthreads();
    }
    tw_i = ntw;
    th_i = nth_i;
    b = nb;
  }
  if constexpr (kAct && !kStats) {
    if (se.w1 != nullptr) {      // out of line: nothing of it lives in the main loop's registers
      SeTailArgs ta;
      ta.w1 = se.w1; ta.hid_fix = se.hid_fix; ta.pool_partial = pool_partial; ta.tiles_per_cta = p.tiles_per_cta; ta.per_chunk = p.per_chunk;
      ta.sq = se.sq; ta.C = p.C; ta.chunks = p.chunks; ta.parts = p.parts; ta.n_tiles = p.tiles_w * p.tiles_h; ta.CB = CB; ta.inv_hw = se.inv_hw;
      se_tail(smem_raw, ta);
    }
  }
  if constexpr (kStats) {
    // CTA reduction of the two statistics (same two-stage scheme as the pool flush), one double atomic per channel
    const int R = p.red_parts;
    for (int pass = 0; pass < 2; ++pass) {
      const float2* v = pass == 0 ? psum : psq;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        red[tid * 8 + 2 * e] = v[e].x;
        red[tid * 8 + 2 * e + 1] = v[e].y;
      }
      __syncthreads();
      for (int idx = tid; idx < CB * R; idx += nth) {
        const int o = idx % CB, q = idx / CB;
        const int gg = o >> 3, e = o & 7;
        float s = 0.f;
        for (int u = gg + G * q; u < nth; u += G * R) s += red[u * 8 + e];
        red2[idx] = s;
      }
      __syncthreads();
      for (int o = tid; o < CB; o += nth) {
        if (c0 + o < p.C) {
          float s = 0.f;
          for (int q = 0; q < R; ++q) s += red2[q * CB + o];
          atomicAdd(stats + (size_t)pass * p.C + c0 + o, (double)s);
        }
      }
      __syncthreads();
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

// Tile plan: search (channel chunk, strip length, tile width / height) for the configuration that keeps the most
// threads busy (<= 256 per CTA, two CTAs per SM), wastes the fewest tile cells on the image edge and re-reads the
// smallest halo.  Every candidate is a valid launch; the score only ranks them.
static int make_plan(DwPlan* pl, int dtype, int H, int W, int C, int K, int S, int pad_lo, int pad_hi,
                     const dfv_dwconv_tuning* tuning = nullptr) {
  DwParams& p = pl->p;
  p.C = C;
  p.Ho = (H + pad_lo + pad_hi - K) / S + 1;
  p.Wo = (W + pad_lo + pad_hi - K) / S + 1;
  if (p.Ho <= 0 || p.Wo <= 0) return DFV_ERR_INVALID;
  p.pad = pad_lo;
  const size_t ts = dtype_size(dtype);
  const int cb_cap = dtype == DFV_BF16 ? 64 : 32;
  const int Ls1[3] = {8, 6, 4}, Ls2[1] = {4};
  const int* Ls = S == 1 ? Ls1 : Ls2;
  const int nL = S == 1 ? 3 : 1;
  double best = -1.0;
  // a caller-supplied dfv_dwconv_tuning (0 = free) restricts the search for stride-1 layers
  int fL = tuning ? tuning->L : 0, fTW = tuning ? tuning->TW : 0, fTH = tuning ? tuning->TH : 0, fCB = tuning ? tuning->CB : 0;
  if (S != 1) fL = fTW = fTH = fCB = 0;
  for (int cb = 8; cb <= cb_cap; cb += 8) {
    if (fCB && cb != fCB) continue;
    // a chunk width that does not divide C is allowed for the full-width chunk only (TMA zero-fills the missing
    // channels of the last chunk; its threads are masked): 8 channel groups = one 128-byte shared-memory row per
    // pixel, the only width whose quarter-warp vector loads never collide on a bank
    if (C % cb && !(cb == cb_cap && C > cb)) continue;
    const int G = cb / 8;
    const double waste = (double)C / ((double)((C + cb - 1) / cb) * cb);
    const double banks = (cb * ts == 128) ? 1.0 : (cb * ts > 64 ? 0.8 : 0.65);
    for (int li = 0; li < nL; ++li) {
      const int L = Ls[li];
      if (fL && L != fL) continue;
      for (int strips = 1; strips * L <= 48; ++strips) {
        const int TW = strips * L;
        if (fTW && TW != fTW) continue;
        if (TW - L >= p.Wo) break;                 // a whole strip beyond the image: never better
        const int per_row = G * strips;
        if (per_row > 256) break;
        for (int TH = 1; TH <= 16 && TH - 1 < p.Ho; ++TH) {
          if (fTH && TH != fTH) continue;
          if (S == 2) {   // stride-2 layers (four of them): the hand-picked 16 x 8 tile with the widest channel chunk measured 4.5-4.9 TB/s
            int cbw = 8;
            for (int q = 8; q <= cb_cap; q += 8) if (C % q == 0) cbw = q;
            const int tw2 = p.Wo >= 16 ? 16 : ((p.Wo + L - 1) / L) * L, th2 = p.Ho >= 8 ? 8 : p.Ho;
            if (cb != cbw || TW != tw2 || TH != th2) continue;
          }
          const int rpr_max = std::max(1, 256 / per_row);
          const int rounds = (TH + rpr_max - 1) / rpr_max;
          const int rpr = (TH + rounds - 1) / rounds;
          const int nt = per_row * rpr;
          const int TWI = (TW - 1) * S + K, THI = (TH - 1) * S + K;
          const size_t tile_bytes = (size_t)THI * TWI * cb * ts;
          const size_t smem = 2 * align_up(tile_bytes, 128) + align_up((size_t)K * K * cb * ts, 16) + (size_t)cb * 4 + (size_t)nt * 32 +
                              (size_t)cb * 16 + 32;
          if (TWI > 256 || THI > 256 || smem > (size_t)(S == 2 ? 220 : 110) * 1024) continue;
          const int tiles_w = (p.Wo + TW - 1) / TW, tiles_h = (p.Ho + TH - 1) / TH;
          const double cover = (double)p.Ho * p.Wo / ((double)tiles_w * TW * tiles_h * TH);
          const double busy = (double)TH / (rounds * rpr);              // rows actually computed per round slot
          const double threads = std::min(1.0, nt / 256.0) * ((nt % 32 == 0) ? 1.0 : (double)nt / ((nt + 31) / 32 * 32));
          const double halo = (double)(TH * S) * (TW * S) / ((double)THI * TWI);
          const double regs = L == 8 ? 0.9 : 1.0;                      // 64 accumulators + a row of weight vectors: 128 registers, spills
          const double per_tile = (double)TH * TW / (TH * TW + 24.0);  // fixed per-tile cost (barrier, TMA issue)
          const double seg = std::min(1.0, (double)cb * ts / 128.0);   // contiguous bytes per pixel the TMA box fetches
          const double score = cover * busy * threads * (0.6 + 0.4 * halo) * regs * per_tile * (0.3 + 0.7 * seg) * waste * banks;
          if (score > best) {
            best = score;
            p.CB = cb;
            pl->L = L;
            p.TW = TW;
            p.TH = TH;
            p.TWI = TWI;
            p.THI = THI;
            p.tiles_w = tiles_w;
            p.tiles_h = tiles_h;
            p.strips = strips;
            p.rounds = rounds;
            p.rpr = rpr;
            p.nthreads = nt;
            pl->smem = smem;
          }
        }
      }
    }
  }
  if (best < 0.0) return DFV_ERR_INVALID;
  p.chunks = pl->chunks = (C + p.CB - 1) / p.CB;
  p.red_parts = std::max(1, std::min(4, p.nthreads / p.CB));
  return DFV_OK;
}

// Grid plan shared by the launcher and dfv_dwconv_pool_parts (the SE-gate kernel must sum the same slot count).
static void plan_grid(DwPlan& pl, int B) {
  const long long per_chunk = (long long)B * pl.p.tiles_w * pl.p.tiles_h;
  pl.p.per_chunk = per_chunk;
  // persistent grid: a multiple of `chunks` (fixed chunk per CTA), about two CTAs per SM
  const int occ = pl.smem <= 110 * 1024 ? 2 : 1;
  long long ctas_per_chunk = (long long)num_sms() * occ / pl.chunks;
  if (ctas_per_chunk < 1) ctas_per_chunk = 1;
  if (ctas_per_chunk > per_chunk) ctas_per_chunk = per_chunk;
  const long long tpc = (per_chunk + ctas_per_chunk - 1) / ctas_per_chunk;
  ctas_per_chunk = (per_chunk + tpc - 1) / tpc;
  pl.p.ctas_per_chunk = (int)ctas_per_chunk;
  pl.p.tiles_per_cta = tpc;
  const long long n = (long long)pl.p.tiles_w * pl.p.tiles_h;
  pl.p.parts = (int)((n + tpc - 2) / tpc + 1);
}

// The fused squeeze tail reuses the two tile buffers: [sq x (CB+1)] weights + [kSeGroup x CB] pool means + flags.
static bool se_tail_fits(const DwPlan& pl, int sq) {
  const size_t ts_bytes = pl.smem;     // >= 2 tile buffers + the rest; recompute the tile part exactly as the kernel does
  (void)ts_bytes;
  const size_t tile_bytes_f32 = (size_t)pl.p.THI * pl.p.TWI * pl.p.CB;   // elements
  const size_t two_tiles_min = 2 * ((tile_bytes_f32 * 2 + 127) / 128) * 128;    // bf16 (the smaller of the two dtypes)
  const size_t need = ((size_t)sq * (pl.p.CB + 1) + 4 + (size_t)kSeGroup * pl.p.CB) * sizeof(float);
  return sq > 0 && sq <= 256 && need <= two_tiles_min;
}

template <typename T, int K, int S, int L, bool kFast, bool kAct, int kCB, bool kStats>
static int launch(const CUtensorMap& tm, const float* w, const float* bias, void* y, float* pool, double* stats, DwPlan& pl, int B,
                  cudaStream_t st, const SeFuse& se) {
  auto kern = dwconv_kernel<T, K, S, L, kFast, kAct, kCB, kStats>;
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    DFV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  DFV_TRY(init_timeout_word_tu());
  plan_grid(pl, B);
  const unsigned grid = (unsigned)((long long)pl.p.ctas_per_chunk * pl.chunks);
  kern<<<grid, pl.p.nthreads, pl.smem, st>>>(tm, w, bias, (T*)y, pool, stats, pl.p, se);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

template <typename T, bool kFast>
static int dispatch(int K, int S, int L, int act, const CUtensorMap& tm, const float* w, const float* bias, void* y, float* pool,
                    double* stats, DwPlan& pl, int B, cudaStream_t st, const SeFuse& se) {
#define DW_CASE(k, s, l)                                                                          \
  if (K == k && S == s && L == l) {                                                               \
    if (stats) return launch<T, k, s, l, kFast, false, 0, true>(tm, w, bias, y, pool, stats, pl, B, st, se);    \
    if (sizeof(T) == 2 && pl.p.CB == 64) {                                                                  \
      if (act) return launch<T, k, s, l, kFast, true, 64, false>(tm, w, bias, y, pool, stats, pl, B, st, se);   \
      return launch<T, k, s, l, kFast, false, 64, false>(tm, w, bias, y, pool, stats, pl, B, st, se);           \
    }                                                                                                       \
    if (act) return launch<T, k, s, l, kFast, true, 0, false>(tm, w, bias, y, pool, stats, pl, B, st, se);      \
    return launch<T, k, s, l, kFast, false, 0, false>(tm, w, bias, y, pool, stats, pl, B, st, se);              \
  }
  DW_CASE(3, 1, 8) DW_CASE(3, 1, 6) DW_CASE(3, 1, 4) DW_CASE(5, 1, 8) DW_CASE(5, 1, 6) DW_CASE(5, 1, 4)
  DW_CASE(3, 2, 4) DW_CASE(5, 2, 4)
#undef DW_CASE
  set_error("dfv_dwconv_fwd: unsupported kernel/stride/strip combination k=%d s=%d L=%d", K, S, L);
  return DFV_ERR_INVALID;
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_dwconv_pool_parts(int dtype, int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi) {
  return dfv_dwconv_pool_parts_tuned(dtype, B, H, W, C, kernel, stride, pad_lo, pad_hi, nullptr);
}

extern "C" int dfv_dwconv_pool_parts_tuned(int dtype, int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi,
                                           const dfv_dwconv_tuning* tuning) {
  DwPlan pl;
  if (!valid_dtype(dtype) || B <= 0 || make_plan(&pl, dtype, H, W, C, kernel, stride, pad_lo, pad_hi, tuning) != DFV_OK) {
    set_error("dfv_dwconv_pool_parts: bad shape");
    return DFV_ERR_INVALID;
  }
  plan_grid(pl, B);
  return pl.p.parts;
}

/* Plan introspection (host only): the tile plan chosen for a layer.
 * out[0..9] = CB, L, TW, TH, threads, smem bytes, tiles_w, tiles_h, parts, grid. */
extern "C" int dfv_dwconv_plan_info(int dtype, int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, int* out) {
  DwPlan pl;
  if (!out || !valid_dtype(dtype) || B <= 0 || make_plan(&pl, dtype, H, W, C, kernel, stride, pad_lo, pad_hi) != DFV_OK) {
    set_error("dfv_dwconv_plan_info: bad shape");
    return DFV_ERR_INVALID;
  }
  plan_grid(pl, B);
  out[0] = pl.p.CB; out[1] = pl.L; out[2] = pl.p.TW; out[3] = pl.p.TH; out[4] = pl.p.nthreads; out[5] = (int)pl.smem;
  out[6] = pl.p.tiles_w; out[7] = pl.p.tiles_h; out[8] = pl.p.parts; out[9] = pl.p.ctas_per_chunk * pl.chunks;
  return DFV_OK;
}

static int dwconv_entry(const void* x, const float* w, const float* bias, void* y, float* pool_partial, double* stats, int dtype,
                        int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, int act, dfv_stream_t stream,
                        const dfv_dwconv_tuning* tuning = nullptr, const SeFuse* se_in = nullptr) {
  DFV_TRY(check_device());
  DFV_REQUIRE(x && w && bias && y, "dfv_dwconv_fwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype), "dfv_dwconv_fwd: bad dtype %d", dtype);
  DFV_REQUIRE((kernel == 3 || kernel == 5) && (stride == 1 || stride == 2), "dfv_dwconv_fwd: k=%d s=%d unsupported",
              kernel, stride);
  DFV_REQUIRE(B > 0 && C > 0 && C % 8 == 0, "dfv_dwconv_fwd: need B > 0 and C %% 8 == 0 (B=%d C=%d)", B, C);
  DFV_REQUIRE(pad_lo >= 0 && pad_hi >= 0 && pad_lo < kernel && pad_hi < kernel, "dfv_dwconv_fwd: bad pad");
  if (debug_flags() & 1) return DFV_OK;
  DwPlan pl;
  DFV_REQUIRE(make_plan(&pl, dtype, H, W, C, kernel, stride, pad_lo, pad_hi, tuning) == DFV_OK,
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
  SeFuse se = {};
  if (se_in) {
    DFV_REQUIRE(act == DFV_ACT_SILU && pool_partial && !stats && se_tail_fits(pl, se_in->sq), "dfv_dwconv_se_fwd: layer cannot fuse the squeeze (see dfv_dwconv_se_supported)");
    se = *se_in;
  }
  if (dtype == DFV_BF16)
    return dispatch<__nv_bfloat16, true>(kernel, stride, pl.L, act, tm, w, bias, y, pool_partial, stats, pl, B, as_stream(stream), se);
  return dispatch<float, false>(kernel, stride, pl.L, act, tm, w, bias, y, pool_partial, stats, pl, B, as_stream(stream), se);
}

extern "C" int dfv_dwconv_fwd(const void* x, const float* w, const float* bias, void* y, float* pool_partial, int dtype,
                              int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, int act,
                              dfv_stream_t stream) {
  return dwconv_entry(x, w, bias, y, pool_partial, nullptr, dtype, B, H, W, C, kernel, stride, pad_lo, pad_hi, act, stream);
}

extern "C" int dfv_dwconv_fwd_tuned(const void* x, const float* w, const float* bias, void* y, float* pool_partial, int dtype,
                                    int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, int act,
                                    const dfv_dwconv_tuning* tuning, dfv_stream_t stream) {
  return dwconv_entry(x, w, bias, y, pool_partial, nullptr, dtype, B, H, W, C, kernel, stride, pad_lo, pad_hi, act, stream, tuning);
}

/* Depthwise conv (+ folded BN + swish + SE pool sums) with the squeeze layer of the SE block fused into the kernel's tail:
 * hid_fix[b][j] += round(2^30 * sum_c w_reduce[j][c] * mean_hw(y[b][..][c]))  (64-bit fixed-point accumulators, zero on entry;
 * the bias is added by dfv_se_excite_fwd).  zero_next / zero_count: a buffer this launch zeroes for the next fused layer. */
extern "C" int dfv_dwconv_se_supported(int dtype, int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, int squeeze) {
  DwPlan pl;
  if (!valid_dtype(dtype) || B <= 0 || make_plan(&pl, dtype, H, W, C, kernel, stride, pad_lo, pad_hi) != DFV_OK) return 0;
  return se_tail_fits(pl, squeeze) ? 1 : 0;
}

/* Policy, measured on a B200 at batch 256 inside the forward (round 2): up to C x squeeze ~ 20k (blocks 0-16 of B4) the tail
 * costs 2-3 us and the gate drops from 18 us (three launches) to 9 us (one).  Beyond that the tail (4 us at C = 960, 12 us at
 * C = 1632) plus the HBM-cold excite launch cost as much as the three-launch gate: fusing every layer measured 0.09 ms
 * SLOWER per forward than fusing blocks 0-16 only. */
extern "C" int dfv_dwconv_se_profitable(int dtype, int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, int squeeze) {
  return dfv_dwconv_se_supported(dtype, B, H, W, C, kernel, stride, pad_lo, pad_hi, squeeze) && (long long)C * squeeze <= 20000 ? 1 : 0;
}

extern "C" int dfv_dwconv_se_fwd(const void* x, const float* w, const float* bias, void* y, float* pool_partial, const float* w_reduce,
                                 long long* hid_fix, long long* zero_next, long long zero_count, int squeeze, int dtype, int B, int H, int W,
                                 int C, int kernel, int stride, int pad_lo, int pad_hi, dfv_stream_t stream) {
  DFV_REQUIRE(w_reduce && hid_fix && squeeze > 0 && pool_partial, "dfv_dwconv_se_fwd: null pointer");
  DFV_REQUIRE(zero_next != hid_fix && zero_count >= 0, "dfv_dwconv_se_fwd: the buffer to zero must not be the live accumulator");
  const int Ho = (H + pad_lo + pad_hi - kernel) / stride + 1, Wo = (W + pad_lo + pad_hi - kernel) / stride + 1;
  SeFuse se;
  se.w1 = w_reduce; se.hid_fix = hid_fix; se.zero_next = zero_next; se.zero_count = zero_next ? zero_count : 0; se.sq = squeeze;
  se.inv_hw = 1.0f / (float)((long long)Ho * Wo);
  return dwconv_entry(x, w, bias, y, pool_partial, nullptr, dtype, B, H, W, C, kernel, stride, pad_lo, pad_hi, DFV_ACT_SILU, stream, nullptr, &se);
}

/* Training forward: raw depthwise conv (no activation) that also ADDS the per-channel sum and sum of squares of its
 * output into stats[2C] (double, zeroed by the caller): train-mode BatchNorm statistics without a second pass
 * (finish with dfv_bn_stats_from_sums). */
extern "C" int dfv_dwconv_stats_fwd(const void* x, const float* w, const float* bias, void* y, double* stats, int dtype, int B,
                                    int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, dfv_stream_t stream) {
  DFV_REQUIRE(stats != nullptr, "dfv_dwconv_stats_fwd: null stats");
  return dwconv_entry(x, w, bias, y, nullptr, stats, dtype, B, H, W, C, kernel, stride, pad_lo, pad_hi, DFV_ACT_NONE, stream);
}
