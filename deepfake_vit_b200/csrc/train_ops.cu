// Training-path building blocks over channels-last [M][C] tensors (M = B * rows_per_image):
//   bn_stats        per-channel batch mean / inverse std (+ running-stat update)      nn.BatchNorm2d/1d, train mode
//   bn_act          y = act(bn(raw)) [* mask] [* rowscale[b]] [+ residual], optional SE pool partials
//   act_bn_bwd      du = dy' * act'(u), per-channel sum(du), sum(du * xhat) -> dgamma, dbeta
//   bn_bwd_apply    d raw = gamma * invstd * (du - mean(du) - xhat * mean(du * xhat))
//   se_*            squeeze-excite forward (saving what backward needs) and backward
//   small helpers   weight cast / transpose, depthwise tap flip, dropout masks, column sums
// All HBM-bound streaming kernels: 16-byte vectors of 8 channels, fixed channel group per thread,
// fp32 accumulation, per-CTA partials finished by a tiny second kernel (deterministic).
#include <algorithm>
#include <type_traits>

#include <cooperative_groups.h>

#include "common.cuh"

namespace dfv {

constexpr int kNT = 256;

// Thread -> (column vector, row lane) mapping shared by the streaming kernels.
struct ColMap {
  int CV, cpp, rpp, col_l, row_l;
  __device__ __forceinline__ ColMap(int C) {
    CV = C >> 3;
    cpp = CV < kNT ? CV : kNT;
    rpp = kNT / cpp;
    col_l = threadIdx.x % cpp;
    row_l = threadIdx.x / cpp;
  }
};

template <bool kFast>
__device__ __forceinline__ float act_fwd(float u, int act) {
  if (act == DFV_ACT_SILU) return silu<kFast>(u);
  if (act == DFV_ACT_RELU) return fmaxf(u, 0.f);
  return u;
}
// d act(u) / du.  swish: sigma * (1 + u * (1 - sigma))  (SwishImplementation.backward of efficientnet-pytorch)
// kFast (bf16 tensors): sigma = 0.5 tanh(u / 2) + 0.5 with one MUFU.TANH.
template <bool kFast = false>
__device__ __forceinline__ float act_grad(float u, int act) {
  if (act == DFV_ACT_SILU) {
    float s;
    if constexpr (kFast) {
      float t;
      asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * u));
      s = fmaf(0.5f, t, 0.5f);
    } else {
      s = sigmoid_exact(u);
    }
    return s * (1.f + u * (1.f - s));
  }
  if (act == DFV_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  return 1.f;
}

// Reduce 16 (or 8) per-thread floats across the row lanes of a column; result valid for row_l == 0.
template <int N>
__device__ __forceinline__ void reduce_rows(float* sm, const ColMap& m, float v[N]) {
  if (m.rpp > 1) {
#pragma unroll
    for (int e = 0; e < N; ++e) sm[threadIdx.x * N + e] = v[e];
    __syncthreads();
    if (m.row_l == 0) {
      for (int rl = 1; rl < m.rpp; ++rl)
#pragma unroll
        for (int e = 0; e < N; ++e) v[e] += sm[(rl * m.cpp + m.col_l) * N + e];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------ streaming helpers
// Raw 16-byte (bf16) / 32-byte (fp32) vector of 8 channels: loads are issued for several rows before any
// of them is unpacked, so every thread keeps U independent requests in flight (these kernels are pure
// HBM streams; with one request per thread they ran at ~35% of the copy bandwidth).
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
  uint4 r;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { r = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void unpack(float v[8]) const {
    v[0] = bf16_lo(r.x); v[1] = bf16_hi(r.x); v[2] = bf16_lo(r.y); v[3] = bf16_hi(r.y);
    v[4] = bf16_lo(r.z); v[5] = bf16_hi(r.z); v[6] = bf16_lo(r.w); v[7] = bf16_hi(r.w);
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void unpack(float v[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <typename T> struct Unroll { static constexpr int U = sizeof(T) == 2 ? 4 : 2; };
template <typename T> struct UnrollStats { static constexpr int U = sizeof(T) == 2 ? 8 : 2; };   // read-only stream: deeper

// cp.async (LDGSTS) staging for the compute-heavy streams: every thread owns private 16-byte shared-memory slots
// that it fills for the NEXT round while it computes the current one, so the memory system always has
// U x (tensors) requests per thread in flight (plain loads drained during the ~17-instruction-per-element math
// and the kernels sat at ~45% of the copy bandwidth).  No block barrier: a thread only reads its own slots.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
constexpr int kPipeU = 4;

// ------------------------------------------------------------------------------------ bn_stats
// Per-CTA column sums -> one partial row [2C] per CTA, finished (in double) by bn_stats_finalize_kernel.
template <typename T>
__global__ void __launch_bounds__(kNT) bn_stats_kernel(const T* __restrict__ raw, long long rows_per_image, int C,
                                                      long long rows_per_chunk, float* __restrict__ partial) {
  pdl_prologue();
  constexpr int U = UnrollStats<T>::U;
  __shared__ float sm[kNT * 16];
  const ColMap m(C);
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = min(r0 + rows_per_chunk, rows_per_image);
  const T* base = raw + (size_t)blockIdx.y * rows_per_image * C;
  float* pout = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C;
  for (int cb = 0; cb < m.CV; cb += m.cpp) {
    const int cv = cb + m.col_l;
    float acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.f;
    if (cv < m.CV && m.row_l < m.rpp) {
      for (long long r = r0 + m.row_l; r < r1; r += (long long)U * m.rpp) {
        Raw8<T> x[U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const long long ri = r + (long long)i * m.rpp;
          if (ri < r1) x[i].load(base + (size_t)ri * C + cv * 8); else x[i].zero();
        }
#pragma unroll
        for (int i = 0; i < U; ++i) {
          float v[8];
          x[i].unpack(v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            acc[e] += v[e];
            acc[8 + e] = fmaf(v[e], v[e], acc[8 + e]);
          }
        }
      }
    }
    reduce_rows<16>(sm, m, acc);
    if (m.row_l == 0 && cv < m.CV) {
      *reinterpret_cast<float4*>(pout + cv * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(pout + cv * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      *reinterpret_cast<float4*>(pout + C + cv * 8) = make_float4(acc[8], acc[9], acc[10], acc[11]);
      *reinterpret_cast<float4*>(pout + C + cv * 8 + 4) = make_float4(acc[12], acc[13], acc[14], acc[15]);
    }
  }
}

// Sum of the per-CTA partial rows: 32 channels x 32 row lanes per CTA, 16 independent loads per thread in flight,
// double accumulation.  Result in sh[0][0][cl] (sum) and sh[1][0][cl] (second moment) for lane 0 threads.
constexpr int kFinLanes = 32;      // row lanes per channel in the finalize kernels (CTA = 32 channels x 32 lanes)
__device__ __forceinline__ void reduce_partials(const float* __restrict__ partial, int n_partial, int C, int c, int lane,
                                                int cl, double (*sh)[kFinLanes][32], double& s, double& q) {
  s = 0.0;
  q = 0.0;
  if (c < C) {
    // the partial rows are read in ONE or two dependent batches (the 8-lane version walked 576 rows in 18 batches of
    // 8 loads: 20 us per finalize, 166 finalizes per training step)
    for (int i = lane; i < n_partial; i += kFinLanes * 8) {
      float a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = i + u * kFinLanes;
        a[u] = r < n_partial ? partial[(size_t)r * 2 * C + c] : 0.f;
        b[u] = r < n_partial ? partial[(size_t)r * 2 * C + C + c] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { s += (double)a[u]; q += (double)b[u]; }
    }
  }
  sh[0][lane][cl] = s;
  sh[1][lane][cl] = q;
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < kFinLanes; ++l) { s += sh[0][l][cl]; q += sh[1][l][cl]; }
  }
}

__global__ void __launch_bounds__(kFinLanes * 32) bn_stats_finalize_kernel(const float* __restrict__ partial, int n_partial, int C,
                                                               double count, float eps, float momentum,
                                                               float* __restrict__ mean, float* __restrict__ invstd,
                                                               float* __restrict__ running_mean,
                                                               float* __restrict__ running_var) {
  pdl_prologue();
  __shared__ double sh[2][kFinLanes][32];
  const int cl = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s, q;
  reduce_partials(partial, n_partial, C, c, lane, cl, sh, s, q);
  if (lane != 0 || c >= C) return;
  const double mu = s / count;
  double var = q / count - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  invstd[c] = 1.0f / sqrtf((float)var + eps);
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// Finish statistics that a producer kernel accumulated as per-channel double sums (dfv_dwconv_stats_fwd).
__global__ void __launch_bounds__(256) bn_stats_from_sums_kernel(const double* __restrict__ acc, int C, double count, float eps,
                                                                float momentum, float* __restrict__ mean,
                                                                float* __restrict__ invstd, float* __restrict__ running_mean,
                                                                float* __restrict__ running_var) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = acc[c] / count;
  double var = acc[C + c] / count - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  invstd[c] = 1.0f / sqrtf((float)var + eps);
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// ------------------------------------------------------------------------------------ bn_act
// kSpec: the backbone's expanded tensors -- swish, no mask, no row scale, no residual -- with those four run-time switches
// folded at compile time (these streams co-limit on issue slots: ~20 instructions per element beside one MUFU)
template <typename T, bool kFast, bool kSpec = false>
__global__ void __launch_bounds__(kNT) bn_act_kernel(const T* __restrict__ raw, const float* __restrict__ mean,
                                                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, int act_rt,
                                                    const float* __restrict__ rowscale_rt, const T* __restrict__ residual_rt,
                                                    const float* __restrict__ mask_rt, T* __restrict__ out,
                                                    float* __restrict__ pool_partial, long long rows_per_image, int C,
                                                    long long rows_per_chunk) {
  pdl_prologue();
  const int act = kSpec ? (int)DFV_ACT_SILU : act_rt;
  const float* rowscale = kSpec ? nullptr : rowscale_rt;
  const T* residual = kSpec ? nullptr : residual_rt;
  const float* mask = kSpec ? nullptr : mask_rt;
  constexpr int U = Unroll<T>::U;
  __shared__ float sm[kNT * 8];
  const ColMap m(C);
  const int b = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = min(r0 + rows_per_chunk, rows_per_image);
  const size_t img = (size_t)b * rows_per_image * C;
  const float rs = rowscale ? rowscale[b] : 1.f;
  for (int cb = 0; cb < m.CV; cb += m.cpp) {
    const int cv = cb + m.col_l;
    float ps[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) ps[e] = 0.f;
    if (cv < m.CV && m.row_l < m.rpp) {
      float sc[8], sh[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = cv * 8 + e;
        const float is = invstd ? invstd[c] : 1.f;
        sc[e] = gamma ? gamma[c] * is : is;
        sh[e] = (beta ? beta[c] : 0.f) - (mean ? mean[c] : 0.f) * sc[e];
      }
      auto body = [&](const Raw8<T>& xi, const Raw8<T>& ri, size_t off) {
        float v[8];
        xi.unpack(v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = act_fwd<kFast>(fmaf(v[e], sc[e], sh[e]), act);
        if (mask) {
          float mk[8];
          load8(mask + off, mk);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] *= mk[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) ps[e] += v[e];
        if (rowscale) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] *= rs;
        }
        if (residual) {
          float q[8];
          ri.unpack(q);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] += q[e];
        }
        store8(out + off, v);
      };
      if constexpr (sizeof(T) == 2) {
        // [2 buffers][U rows][1 or 2 tensors][kNT threads] = 16 slots per thread either way: without a residual (every
        // layer but the block outputs) the residual's slots carry four more rows of `raw` -- with 4 x 16 bytes per thread in
        // flight the one-input forward streams ran at 3.8-4.8 TB/s (28.63 -> 28.08 ms per training step with 8 x 16)
        // kSpec: swish on channel PAIRS with packed fp32 instructions (h = u / 2 = x * sc / 2 + sh / 2, y = h tanh(h) + h)
        float2 sch[4], shh[4], ps2[4];
        if constexpr (kSpec) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            sch[j] = make_float2(0.5f * sc[2 * j], 0.5f * sc[2 * j + 1]);
            shh[j] = make_float2(0.5f * sh[2 * j], 0.5f * sh[2 * j + 1]);
            ps2[j] = make_float2(0.f, 0.f);
          }
        }
        auto body2 = [&](const uint4& xr, size_t off) {
          const uint32_t xw[4] = {xr.x, xr.y, xr.z, xr.w};
          uint32_t ow[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 h = ffma2(make_float2(bf16_lo(xw[j]), bf16_hi(xw[j])), sch[j], shh[j]);
            float2 t;
            asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
            const float2 y = ffma2(h, t, h);
            ps2[j] = fadd2(ps2[j], y);
            ow[j] = pack_bf16(y.x, y.y);
          }
          *reinterpret_cast<uint4*>(out + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        };
        extern __shared__ uint4 stage[];
        auto stream_rows = [&](auto u_c, auto res_c) {
          constexpr int U = decltype(u_c)::value;
          constexpr bool kRes = decltype(res_c)::value;
          constexpr int NTen = kRes ? 2 : 1;
          auto slot = [&](int buf, int i, int t) { return stage + ((buf * U + i) * NTen + t) * kNT + threadIdx.x; };
          auto prefetch = [&](long long r, int buf) {
#pragma unroll
            for (int i = 0; i < U; ++i) {
              const long long ri = r + (long long)i * m.rpp;
              if (ri < r1) {
                const size_t off = img + (size_t)ri * C + cv * 8;
                cp_async16(slot(buf, i, 0), raw + off);
                if constexpr (kRes) cp_async16(slot(buf, i, 1), residual + off);
              }
            }
            cp_async_commit();
          };
          long long r = r0 + m.row_l;
          int buf = 0;
          prefetch(r, 0);
          for (; r < r1; r += (long long)U * m.rpp, buf ^= 1) {
            prefetch(r + (long long)U * m.rpp, buf ^ 1);
            cp_async_wait1();
#pragma unroll
            for (int i = 0; i < U; ++i) {
              const long long ri = r + (long long)i * m.rpp;
              if (ri < r1) {
                if constexpr (kSpec) {
                  body2(*slot(buf, i, 0), img + (size_t)ri * C + cv * 8);
                } else {
                  Raw8<T> xi, qi;
                  xi.r = *slot(buf, i, 0);
                  if constexpr (kRes) qi.r = *slot(buf, i, 1); else qi.zero();
                  body(xi, qi, img + (size_t)ri * C + cv * 8);
                }
              }
            }
          }
        };
        if (residual) stream_rows(std::integral_constant<int, kPipeU>{}, std::true_type{});
        else stream_rows(std::integral_constant<int, 2 * kPipeU>{}, std::false_type{});
        if constexpr (kSpec) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { ps[2 * j] = ps2[j].x; ps[2 * j + 1] = ps2[j].y; }
        }
      } else {
        for (long long r = r0 + m.row_l; r < r1; r += (long long)U * m.rpp) {
          Raw8<T> x[U], rr[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const long long ri = r + (long long)i * m.rpp;
            if (ri < r1) {
              const size_t off = img + (size_t)ri * C + cv * 8;
              x[i].load(raw + off);
              if (residual) rr[i].load(residual + off); else rr[i].zero();
            }
          }
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const long long ri = r + (long long)i * m.rpp;
            if (ri < r1) body(x[i], rr[i], img + (size_t)ri * C + cv * 8);
          }
        }
      }
    }
    if (pool_partial) {
      reduce_rows<8>(sm, m, ps);
      if (m.row_l == 0 && cv < m.CV) {
        float* po = pool_partial + ((size_t)b * gridDim.x + blockIdx.x) * C + cv * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) po[e] = ps[e];
      }
    }
  }
}

// ------------------------------------------------------------------------------------ act_bn_bwd
// kApply: the SECOND pass of the pair (dfv_act_bn_bwd_apply): same pipelined loads of (g, raw), recomputes du and writes
//   d raw = gamma * invstd * (du - coef[0] - xhat * coef[1])  into `du` (may alias g); no sums.
// kDefer (gated swish layers, bf16, reduction pass only): the SE backward needs sum_hw(g * d) BEFORE this layer's input gradient
// exists (its dpool enters gi = g * gate + dpool / HW), which used to cost a separate pass over (g, d).  The reduction is
// linear in (gate, dpool) per image, so this pass accumulates per (image, channel)
//   S1 = sum g act'(u),  S2 = sum act'(u),  S3 = sum g act'(u) x,  S4 = sum act'(u) x,  D = sum g swish(u)
// (swish(u) = d is recomputed from raw: one more FFMA2 per pair), the SE backward runs on D, and
// bn_bwd_finalize_defer_kernel forms sum du = sum_b (gate S1 + dp S2), sum du xhat = is (sum_b (gate S3 + dp S4)) + nm sum du.
template <typename T, bool kGate, int U, int MINB, bool kApply = false, bool kSpec = false, bool kDefer = false>
__global__ void __launch_bounds__(kNT, MINB) act_bn_bwd_kernel(const T* __restrict__ g, const T* __restrict__ raw,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           int act_rt, const T* __restrict__ gate, const float* __restrict__ dpool,
                                                           float inv_hw, const float* __restrict__ rowscale_rt,
                                                           const float* __restrict__ mask_rt, T* __restrict__ du,
                                                           float* __restrict__ partial, long long rows_per_image, int C,
                                                           long long rows_per_chunk, const float* __restrict__ coef = nullptr,
                                                           float* __restrict__ partial_d = nullptr) {
  static_assert(!kDefer || (kGate && kSpec && !kApply && sizeof(T) == 2), "deferred gate: gated swish reduction pass, bf16");
  pdl_prologue();
  // kSpec: swish, no dropout mask, no drop-connect row scale (the backbone's expanded tensors) folded at compile time
  const int act = kSpec ? (int)DFV_ACT_SILU : act_rt;
  const float* rowscale = kSpec ? nullptr : rowscale_rt;
  const float* mask = kSpec ? nullptr : mask_rt;
  __shared__ float sm[kApply ? 1 : kNT * 16];
  const ColMap m(C);
  const int b = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = min(r0 + rows_per_chunk, rows_per_image);
  const size_t img = (size_t)b * rows_per_image * C;
  const float rs = rowscale ? rowscale[b] : 1.f;
  float* pout = partial + ((size_t)b * gridDim.x + blockIdx.x) * (kDefer ? 4 : 2) * C;
  for (int cb = 0; cb < m.CV; cb += m.cpp) {
    const int cv = cb + m.col_l;
    float acc[16], acc_b[kDefer ? 24 : 1];      // kDefer: acc = S1 | S2, acc_b = S3 | S4 | D
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.f;
    if constexpr (kDefer) {
#pragma unroll
      for (int e = 0; e < 24; ++e) acc_b[e] = 0.f;
    }
    if (cv < m.CV && m.row_l < m.rpp) {
      // xhat = x * is + nm,  u = xhat * ga + be
      float is[8], nm[8], ga[8], be[8], gt[8], dp[8], pq[kApply ? 8 : 1], pr[kApply ? 8 : 1];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = cv * 8 + e;
        is[e] = invstd ? invstd[c] : 1.f;
        nm[e] = -(mean ? mean[c] : 0.f) * is[e];
        ga[e] = gamma ? gamma[c] : 1.f;
        be[e] = beta ? beta[c] : 0.f;
        if constexpr (kGate) dp[e] = dpool ? dpool[(size_t)b * C + c] * inv_hw * rs : 0.f;
        if constexpr (kApply) {      // out = (ga * is) * d - pq * x - pr
          const float pa = ga[e] * is[e];
          pq[e] = pa * coef[C + c] * is[e];
          pr[e] = pa * fmaf(coef[C + c], nm[e], coef[c]);
        }
      }
      if constexpr (kGate) {
        if (gate) load8(gate + (size_t)b * C + cv * 8, gt);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) gt[e] = 1.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) gt[e] *= rs;
      }
      auto body = [&](const Raw8<T>& gxi, const Raw8<T>& xxi, size_t off) {
        float gv[8], x[8];
        gxi.unpack(gv);
        xxi.unpack(x);
        float mk[8];
        if (mask) load8(mask + off, mk);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float gi;
          if constexpr (kGate) gi = fmaf(gv[e], gt[e], dp[e]);
          else gi = gv[e] * rs;
          if (mask) gi *= mk[e];
          const float xh = fmaf(x[e], is[e], nm[e]);
          const float u = fmaf(xh, ga[e], be[e]);
          const float d = gi * act_grad<sizeof(T) == 2>(u, act);
          if constexpr (kApply) {
            gv[e] = fmaf(ga[e] * is[e], d, -fmaf(pq[e], x[e], pr[e]));
          } else {
            gv[e] = d;
            acc[e] += d;
            acc[8 + e] = fmaf(d, xh, acc[8 + e]);
          }
        }
        if (kApply || du) store8(du + off, gv);      // du = NULL: reduction only (the apply pass recomputes du)
      };
      if constexpr (sizeof(T) == 2) {
        // kSpec (swish, bf16): the same arithmetic on PAIRS of channels with packed fp32 instructions (FFMA2 / FMUL2 / FADD2
        // take one issue slot for two lanes' operations; these streams co-limit on issue slots, ~16 scalar instructions per
        // element beside the MUFU).  With h = u / 2 = x * c1 + c2, t = tanh(h), s = sigma(u) = t / 2 + 1 / 2:
        //   act'(u) = s + u s (1 - s) = s - h (t^2 - 1) / 2
        float2 c1p[4], c2p[4], isp[4], nmp[4], gtp[kGate ? 4 : 1], dpp[kGate ? 4 : 1], pap[kApply ? 4 : 1], npq[kApply ? 4 : 1],
            npr[kApply ? 4 : 1], accd[4], accx[4], acc2[kDefer ? 4 : 1], acc4[kDefer ? 4 : 1], accg[kDefer ? 4 : 1];
        if constexpr (kDefer) {
#pragma unroll
          for (int j = 0; j < 4; ++j) acc2[j] = acc4[j] = accg[j] = make_float2(0.f, 0.f);
        }
        if constexpr (kSpec) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e0 = 2 * j, e1 = 2 * j + 1;
            c1p[j] = make_float2(0.5f * is[e0] * ga[e0], 0.5f * is[e1] * ga[e1]);
            c2p[j] = make_float2(0.5f * fmaf(nm[e0], ga[e0], be[e0]), 0.5f * fmaf(nm[e1], ga[e1], be[e1]));
            isp[j] = make_float2(is[e0], is[e1]);
            nmp[j] = make_float2(nm[e0], nm[e1]);
            accd[j] = accx[j] = make_float2(0.f, 0.f);
            if constexpr (kGate) { gtp[j] = make_float2(gt[e0], gt[e1]); dpp[j] = make_float2(dp[e0], dp[e1]); }
            if constexpr (kApply) {
              pap[j] = make_float2(ga[e0] * is[e0], ga[e1] * is[e1]);
              npq[j] = make_float2(-pq[e0], -pq[e1]);
              npr[j] = make_float2(-pr[e0], -pr[e1]);
            }
          }
        }
        auto body2 = [&](const uint4& gr, const uint4& xr, size_t off) {
          const uint32_t gw[4] = {gr.x, gr.y, gr.z, gr.w}, xw[4] = {xr.x, xr.y, xr.z, xr.w};
          uint32_t ow[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 g2 = make_float2(bf16_lo(gw[j]), bf16_hi(gw[j])), x2 = make_float2(bf16_lo(xw[j]), bf16_hi(xw[j]));
            float2 gi;
            if constexpr (kGate && !kDefer) gi = ffma2(g2, gtp[j], dpp[j]); else gi = g2;
            const float2 h = ffma2(x2, c1p[j], c2p[j]);
            float2 t;
            asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
            const float2 sg = ffma2(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
            const float2 w = ffma2(t, t, make_float2(-1.f, -1.f));
            const float2 z = fmul2(h, w);
            const float2 gr2 = ffma2(z, make_float2(-0.5f, -0.5f), sg);
            const float2 d = fmul2(gi, gr2);
            if constexpr (kDefer) {      // d = g act'(u) here; accd = S1, acc2 = S2, accx = S3, acc4 = S4, accg = D
              accd[j] = fadd2(accd[j], d);
              acc2[j] = fadd2(acc2[j], gr2);
              accx[j] = ffma2(d, x2, accx[j]);
              acc4[j] = ffma2(gr2, x2, acc4[j]);
              accg[j] = ffma2(g2, ffma2(h, t, h), accg[j]);
            } else if constexpr (kApply) {
              const float2 o = ffma2(pap[j], d, ffma2(npq[j], x2, npr[j]));
              ow[j] = pack_bf16(o.x, o.y);
            } else {
              const float2 xh = ffma2(x2, isp[j], nmp[j]);
              accd[j] = fadd2(accd[j], d);
              accx[j] = ffma2(d, xh, accx[j]);
              if (du) ow[j] = pack_bf16(d.x, d.y);
            }
          }
          if constexpr (!kDefer) {
            if (kApply || du) *reinterpret_cast<uint4*>(du + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          }
        };
        extern __shared__ uint4 stage[];   // [2 buffers][kPipeU rows][2 tensors][kNT threads]
        auto slot = [&](int buf, int i, int t) { return stage + ((buf * kPipeU + i) * 2 + t) * kNT + threadIdx.x; };
        auto prefetch = [&](long long r, int buf) {
#pragma unroll
          for (int i = 0; i < kPipeU; ++i) {
            const long long ri = r + (long long)i * m.rpp;
            if (ri < r1) {
              const size_t off = img + (size_t)ri * C + cv * 8;
              cp_async16(slot(buf, i, 0), g + off);
              cp_async16(slot(buf, i, 1), raw + off);
            }
          }
          cp_async_commit();
        };
        long long r = r0 + m.row_l;
        int buf = 0;
        prefetch(r, 0);
        for (; r < r1; r += (long long)kPipeU * m.rpp, buf ^= 1) {
          prefetch(r + (long long)kPipeU * m.rpp, buf ^ 1);   // rows >= r1 are skipped inside; the (empty) group is still committed
          cp_async_wait1();
#pragma unroll
          for (int i = 0; i < kPipeU; ++i) {
            const long long ri = r + (long long)i * m.rpp;
            if (ri < r1) {
              if constexpr (kSpec) {
                body2(*slot(buf, i, 0), *slot(buf, i, 1), img + (size_t)ri * C + cv * 8);
              } else {
                Raw8<T> gxi, xxi;
                gxi.r = *slot(buf, i, 0);
                xxi.r = *slot(buf, i, 1);
                body(gxi, xxi, img + (size_t)ri * C + cv * 8);
              }
            }
          }
        }
        if constexpr (kDefer) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[2 * j] = accd[j].x; acc[2 * j + 1] = accd[j].y;
            acc[8 + 2 * j] = acc2[j].x; acc[8 + 2 * j + 1] = acc2[j].y;
            acc_b[2 * j] = accx[j].x; acc_b[2 * j + 1] = accx[j].y;
            acc_b[8 + 2 * j] = acc4[j].x; acc_b[8 + 2 * j + 1] = acc4[j].y;
            acc_b[16 + 2 * j] = accg[j].x; acc_b[16 + 2 * j + 1] = accg[j].y;
          }
        } else if constexpr (kSpec && !kApply) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[2 * j] = accd[j].x; acc[2 * j + 1] = accd[j].y;
            acc[8 + 2 * j] = accx[j].x; acc[8 + 2 * j + 1] = accx[j].y;
          }
        }
      } else {
        for (long long r = r0 + m.row_l; r < r1; r += (long long)U * m.rpp) {
          Raw8<T> gx[U], xx[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const long long ri = r + (long long)i * m.rpp;
            if (ri < r1) {
              const size_t off = img + (size_t)ri * C + cv * 8;
              gx[i].load(g + off);
              xx[i].load(raw + off);
            }
          }
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const long long ri = r + (long long)i * m.rpp;
            if (ri < r1) body(gx[i], xx[i], img + (size_t)ri * C + cv * 8);
          }
        }
      }
    }
    if constexpr (kApply) continue;
    reduce_rows<16>(sm, m, acc);
    if (m.row_l == 0 && cv < m.CV) {
      *reinterpret_cast<float4*>(pout + cv * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(pout + cv * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      *reinterpret_cast<float4*>(pout + C + cv * 8) = make_float4(acc[8], acc[9], acc[10], acc[11]);
      *reinterpret_cast<float4*>(pout + C + cv * 8 + 4) = make_float4(acc[12], acc[13], acc[14], acc[15]);
    }
    if constexpr (kDefer) {
      reduce_rows<16>(sm, m, acc_b);
      if (m.row_l == 0 && cv < m.CV) {
        *reinterpret_cast<float4*>(pout + 2 * C + cv * 8) = make_float4(acc_b[0], acc_b[1], acc_b[2], acc_b[3]);
        *reinterpret_cast<float4*>(pout + 2 * C + cv * 8 + 4) = make_float4(acc_b[4], acc_b[5], acc_b[6], acc_b[7]);
        *reinterpret_cast<float4*>(pout + 3 * C + cv * 8) = make_float4(acc_b[8], acc_b[9], acc_b[10], acc_b[11]);
        *reinterpret_cast<float4*>(pout + 3 * C + cv * 8 + 4) = make_float4(acc_b[12], acc_b[13], acc_b[14], acc_b[15]);
      }
      reduce_rows<8>(sm, m, acc_b + 16);
      if (m.row_l == 0 && cv < m.CV) {
        float* pd = partial_d + ((size_t)b * gridDim.x + blockIdx.x) * C + cv * 8;
        *reinterpret_cast<float4*>(pd) = make_float4(acc_b[16], acc_b[17], acc_b[18], acc_b[19]);
        *reinterpret_cast<float4*>(pd + 4) = make_float4(acc_b[20], acc_b[21], acc_b[22], acc_b[23]);
      }
    }
  }
}

// Deferred-gate finalize: sums the per-(image, chunk) rows [S1 | S2 | S3 | S4] of the kDefer reduction pass with the image's gate
// and dpool (see act_bn_bwd_kernel), in double, CTA = 32 channels x 32 row lanes.
template <typename GT>
__global__ void __launch_bounds__(kFinLanes * 32) bn_bwd_finalize_defer_kernel(const float* __restrict__ partial, int chunks, int B, int C,
                                                                   const GT* __restrict__ gate, const float* __restrict__ dpool,
                                                                   float inv_hw, const float* __restrict__ mean,
                                                                   const float* __restrict__ invstd, double count,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                   float* __restrict__ coef) {
  pdl_prologue();
  __shared__ double sh[2][kFinLanes][32];
  const int cl = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int n = B * chunks;
  double s1 = 0.0, sx = 0.0;
  if (c < C) {
    for (int i = lane; i < n; i += kFinLanes * 4) {
      float v[4][4], gv[4], dv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = i + u * kFinLanes;
        const bool ok = r < n;
        const int bi = ok ? r / chunks : 0;
        const float* row = partial + (size_t)(ok ? r : 0) * 4 * C + c;
#pragma unroll
        for (int q = 0; q < 4; ++q) v[u][q] = ok ? row[(size_t)q * C] : 0.f;
        gv[u] = ok ? (float)gate[(size_t)bi * C + c] : 0.f;
        dv[u] = ok ? dpool[(size_t)bi * C + c] * inv_hw : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s1 += (double)gv[u] * (double)v[u][0] + (double)dv[u] * (double)v[u][1];
        sx += (double)gv[u] * (double)v[u][2] + (double)dv[u] * (double)v[u][3];
      }
    }
  }
  sh[0][lane][cl] = s1;
  sh[1][lane][cl] = sx;
  __syncthreads();
  if (lane != 0 || c >= C) return;
  for (int l = 1; l < kFinLanes; ++l) { s1 += sh[0][l][cl]; sx += sh[1][l][cl]; }
  const double is = (double)invstd[c], nm = -(double)mean[c] * is;
  const double s2 = is * sx + nm * s1;              // sum du * xhat,  xhat = x * is + nm
  if (dbeta) dbeta[c] = (float)s1;
  if (dgamma) dgamma[c] = (float)s2;
  coef[c] = (float)(s1 / count);
  coef[C + c] = (float)(s2 / count);
}

__global__ void __launch_bounds__(kFinLanes * 32) bn_bwd_finalize_kernel(const float* __restrict__ partial, int n_partial, int C,
                                                             double count, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, float* __restrict__ coef) {
  pdl_prologue();
  __shared__ double sh[2][kFinLanes][32];
  const int cl = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s1, s2;
  reduce_partials(partial, n_partial, C, c, lane, cl, sh, s1, s2);
  if (lane != 0 || c >= C) return;
  if (dbeta) dbeta[c] = (float)s1;
  if (dgamma) dgamma[c] = (float)s2;
  coef[c] = (float)(s1 / count);
  coef[C + c] = (float)(s2 / count);
}

template <typename T>
__global__ void __launch_bounds__(kNT) bn_bwd_apply_kernel(const T* __restrict__ du, const T* __restrict__ raw,
                                                          const float* __restrict__ mean, const float* __restrict__ invstd,
                                                          const float* __restrict__ gamma, const float* __restrict__ coef,
                                                          T* __restrict__ draw, long long M, int C, long long rows_per_chunk) {
  pdl_prologue();
  constexpr int U = Unroll<T>::U;
  const ColMap m(C);
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = min(r0 + rows_per_chunk, M);
  if (m.row_l >= m.rpp) return;
  for (int cb = 0; cb < m.CV; cb += m.cpp) {
    const int cv = cb + m.col_l;
    if (cv >= m.CV) continue;
    // d raw = gi * du - gk * x - k0:   gi = gamma * is,  gk = gi * is * coef2,  k0 = gi * (coef1 - mu * is * coef2)
    float gi[8], gk[8], k0[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cv * 8 + e;
      const float is = invstd[c];
      gi[e] = (gamma ? gamma[c] : 1.f) * is;
      const float k2 = is * coef[C + c];
      gk[e] = gi[e] * k2;
      k0[e] = gi[e] * (coef[c] - mean[c] * k2);
    }
    for (long long r = r0 + m.row_l; r < r1; r += (long long)U * m.rpp) {
      Raw8<T> dd[U], xx[U];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const long long ri = r + (long long)i * m.rpp;
        if (ri < r1) {
          const size_t off = (size_t)ri * C + cv * 8;
          dd[i].load(du + off);
          xx[i].load(raw + off);
        }
      }
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const long long ri = r + (long long)i * m.rpp;
        if (ri < r1) {
          float d[8], x[8];
          dd[i].unpack(d);
          xx[i].unpack(x);
#pragma unroll
          for (int e = 0; e < 8; ++e) d[e] = fmaf(gi[e], d[e], -fmaf(gk[e], x[e], k0[e]));
          store8(draw + (size_t)ri * C + cv * 8, d);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------ SE (train)
// dgate partials: sum over positions of dA[pos][c] * d[pos][c]
template <typename T>
__global__ void __launch_bounds__(kNT) dot_rows_kernel(const T* __restrict__ a, const T* __restrict__ d,
                                                      long long rows_per_image, int C, long long rows_per_chunk,
                                                      float* __restrict__ partial) {
  pdl_prologue();
  __shared__ float sm[kNT * 8];
  const ColMap m(C);
  const int b = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = min(r0 + rows_per_chunk, rows_per_image);
  const size_t img = (size_t)b * rows_per_image * C;
  for (int cb = 0; cb < m.CV; cb += m.cpp) {
    const int cv = cb + m.col_l;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    if (cv < m.CV && m.row_l < m.rpp) {
      for (long long r = r0 + m.row_l; r < r1; r += m.rpp) {
        const size_t off = img + (size_t)r * C + cv * 8;
        float x[8], y[8];
        load8(a + off, x);
        load8(d + off, y);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(x[e], y[e], acc[e]);
      }
    }
    reduce_rows<8>(sm, m, acc);
    if (m.row_l == 0 && cv < m.CV) {
      float* po = partial + ((size_t)b * gridDim.x + blockIdx.x) * C + cv * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) po[e] = acc[e];
    }
  }
}

// Per image group: dgate (finish partials) -> dz -> dh1 -> dpool, on a CLUSTER of 8 CTAs (same scheme as the forward
// gate kernel in se.cu): every CTA owns a channel slice; the squeeze-wide dh1 vector is reduced across the cluster
// through distributed shared memory.  The one-CTA-per-image version walked w2 with a 272-byte lane stride and took
// ~36 us per layer at batch 64.
constexpr int kSeBwdCluster = 8;

template <int IMG>
__global__ void __cluster_dims__(kSeBwdCluster, 1, 1) __launch_bounds__(256, 2)
    se_bwd_image_kernel(const float* __restrict__ dgate_partial, int parts, const float* __restrict__ gate_f32,
                        const float* __restrict__ h1, const float* __restrict__ w1, const float* __restrict__ w2,
                        float* __restrict__ dz_out, float* __restrict__ dh1_out, float* __restrict__ dpool, int B, int C, int sq) {
  pdl_prologue();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float sm[];
  const int cper = (C + kSeBwdCluster - 1) / kSeBwdCluster;
  float* dz = sm;                                   // [IMG][cper]      this CTA's channel slice
  float* part = dz + (size_t)IMG * cper;            // [IMG][sq]        partial dh1 of the slice (read by peers)
  float* dh1 = part + (size_t)IMG * sq;             // [IMG][sq]        full dh1
  float* red = dh1 + (size_t)IMG * sq;              // [8 warps][IMG][sq]
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / kSeBwdCluster) * IMG, tid = threadIdx.x;
  const int c0 = min(C, rank * cper), c1 = min(C, c0 + cper), cn = c1 - c0;
  const int warp = tid >> 5, lane = tid & 31;

  for (int cl = tid; cl < cper; cl += blockDim.x) {
#pragma unroll
    for (int im = 0; im < IMG; ++im) {
      float v = 0.f;
      if (cl < cn && b0 + im < B) {
        const size_t bc = (size_t)(b0 + im) * C + c0 + cl;
        const float* pb = dgate_partial + (size_t)(b0 + im) * parts * C + c0 + cl;
        float s = 0.f;
        for (int t0 = 0; t0 < parts; t0 += 8) {      // eight partial rows per round trip (a plain loop chained them)
          float tv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) tv[u] = t0 + u < parts ? __ldg(pb + (size_t)(t0 + u) * C) : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) s += tv[u];
        }
        const float gv = gate_f32[bc];
        v = s * gv * (1.f - gv);
        dz_out[bc] = v;
      }
      dz[im * cper + cl] = v;
    }
  }
  __syncthreads();
  // partial dh1[im][j] over this slice: warp -> channels (stride 8), lane -> j (coalesced rows of w2 [C][sq])
  {
    float acc[4][IMG];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int im = 0; im < IMG; ++im) acc[u][im] = 0.f;
#pragma unroll 4
    for (int cl = warp; cl < cn; cl += 8) {
      const float* wr = w2 + (size_t)(c0 + cl) * sq;
      float wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) wv[u] = lane + 32 * u < sq ? __ldg(wr + lane + 32 * u) : 0.f;
#pragma unroll
      for (int im = 0; im < IMG; ++im) {
        const float z = dz[im * cper + cl];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u][im] = fmaf(z, wv[u], acc[u][im]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (lane + 32 * u < sq)
#pragma unroll
        for (int im = 0; im < IMG; ++im) red[((size_t)warp * IMG + im) * sq + lane + 32 * u] = acc[u][im];
  }
  __syncthreads();
  for (int i = tid; i < IMG * sq; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[(size_t)w * IMG * sq + i];
    part[i] = s;
  }
  cluster.sync();
  for (int j = tid; j < sq; j += blockDim.x) {
#pragma unroll
    for (int im = 0; im < IMG; ++im) {
      const int i = im * sq + j;
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < kSeBwdCluster; ++r) s += cluster.map_shared_rank(part, r)[i];
      float v = 0.f;
      if (b0 + im < B) {
        v = s * act_grad(h1[(size_t)(b0 + im) * sq + j], DFV_ACT_SILU);
        if (rank == 0) dh1_out[(size_t)(b0 + im) * sq + j] = v;
      }
      dh1[i] = v;
    }
  }
  cluster.sync();   // peers have finished reading `part`; dh1 complete in every CTA
  for (int cl = tid; cl < cn; cl += blockDim.x) {
    float s[IMG];
#pragma unroll
    for (int im = 0; im < IMG; ++im) s[im] = 0.f;
    for (int jb = 0; jb < sq; jb += 32) {
      float wv[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) wv[u] = jb + u < sq ? __ldg(w1 + (size_t)(jb + u) * C + c0 + cl) : 0.f;
#pragma unroll
      for (int u = 0; u < 32; ++u)
        if (jb + u < sq)
#pragma unroll
          for (int im = 0; im < IMG; ++im) s[im] = fmaf(dh1[im * sq + jb + u], wv[u], s[im]);
    }
#pragma unroll
    for (int im = 0; im < IMG; ++im)
      if (b0 + im < B) dpool[(size_t)(b0 + im) * C + c0 + cl] = s[im];
  }
}

// Weight gradients, reduced over the batch.  grid = (ceil(C / 128), squeeze slices), thread = (channel, quarter of the batch):
// the first version (thread = channel, 128 threads) walked the batch in four dependent groups of 16 load pairs, with the db2
// column sum as four more in front of them on the first slice -- 33 us per layer at batch 64 for 1 MB of operands (ncu:
// 20 warps waiting on long_scoreboard per issue).  Now every thread's 2 x B/4 loads are in flight at once, db2 rides the same
// loads, and the four quarters meet in shared memory in a fixed order.
constexpr int kSeWgQ = 4;
__global__ void __launch_bounds__(128 * kSeWgQ) se_bwd_weights_kernel(const float* __restrict__ dz, const float* __restrict__ dh1,
                                                                     const float* __restrict__ pooled, const float* __restrict__ h1,
                                                                     float* __restrict__ dw1, float* __restrict__ db1,
                                                                     float* __restrict__ dw2, float* __restrict__ db2, int B, int C,
                                                                     int sq) {
  pdl_prologue();
  extern __shared__ float sm[];   // hidden [B][jn], dh1 [B][jn] of this slice
  __shared__ float red[kSeWgQ - 1][9][128];
  const int j0 = (int)((long long)sq * blockIdx.y / gridDim.y), j1 = (int)((long long)sq * (blockIdx.y + 1) / gridDim.y);
  const int jn = j1 - j0;
  float* hid = sm;
  float* dh = sm + (size_t)B * jn;
  for (int i = threadIdx.x; i < B * jn; i += blockDim.x) {
    const int b = i / jn, j = j0 + i % jn;
    const float v = h1[(size_t)b * sq + j];
    hid[i] = v * sigmoid_exact(v);
    dh[i] = dh1[(size_t)b * sq + j];
  }
  __syncthreads();
  const int cl = threadIdx.x & 127, q = threadIdx.x >> 7;
  const int c = blockIdx.x * 128 + cl;
  const int b0 = (int)((long long)B * q / kSeWgQ), b1 = (int)((long long)B * (q + 1) / kSeWgQ);
  for (int jj = 0; jj < jn; jj += 4) {
    float s2[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, sb = 0.f;
    if (c < C) {
#pragma unroll 16
      for (int b = b0; b < b1; ++b) {
        const float z = __ldg(dz + (size_t)b * C + c), pv = __ldg(pooled + (size_t)b * C + c);
        sb += z;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (jj + u < jn) {
            s2[u] = fmaf(z, hid[b * jn + jj + u], s2[u]);
            s1[u] = fmaf(dh[b * jn + jj + u], pv, s1[u]);
          }
        }
      }
    }
    if (q > 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) { red[q - 1][u][cl] = s2[u]; red[q - 1][4 + u][cl] = s1[u]; }
      red[q - 1][8][cl] = sb;
    }
    __syncthreads();
    if (q == 0 && c < C) {
#pragma unroll
      for (int r = 0; r < kSeWgQ - 1; ++r) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { s2[u] += red[r][u][cl]; s1[u] += red[r][4 + u][cl]; }
        sb += red[r][8][cl];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (jj + u < jn) {
          dw2[(size_t)c * sq + j0 + jj + u] = s2[u];
          dw1[(size_t)(j0 + jj + u) * C + c] = s1[u];
        }
      }
      if (blockIdx.y == 0 && jj == 0) db2[c] = sb;
    }
    __syncthreads();
  }
  if (blockIdx.x == 0) {
    for (int j = threadIdx.x; j < jn; j += blockDim.x) {
      float s = 0.f;
#pragma unroll 8
      for (int b = 0; b < B; ++b) s += dh[b * jn + j];
      db1[j0 + j] = s;
    }
  }
}

// ------------------------------------------------------------------------------------ helpers
// dst[r][c] (T) = src (fp32): same layout, or transposed (src is [cols][rows]).
template <typename T>
__global__ void cast_weight_kernel(const float* __restrict__ src, T* __restrict__ dst, int rows, int cols, int transpose) {
  pdl_prologue();
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float v = transpose ? src[(size_t)c * rows + r] : src[i];
    if constexpr (sizeof(T) == 2) dst[i] = __float2bfloat16_rn(v); else dst[i] = v;
  }
}

// Depthwise weights: torch [C][1][k][k] -> [k*k][C] fp32, optionally with the taps flipped (dgrad of a stride-1 conv).
__global__ void dw_weight_pack_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int kk, int flip) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * kk) return;
  const int t = i / C, c = i % C;
  dst[i] = src[(size_t)c * kk + (flip ? kk - 1 - t : t)];
}
// and back: gradient [k*k][C] -> torch layout [C][k*k]
__global__ void dw_weight_unpack_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int kk) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * kk) return;
  const int c = i / kk, t = i % kk;
  dst[i] = src[(size_t)t * C + c];
}

// Counter-based uniform hash (splitmix64 finaliser): out[i] = u(seed, i) >= p ? 1 / (1 - p) : 0.
// seed_dev (optional): a device word ADDED to the seed -- the launch arguments of a captured CUDA graph are frozen, a device
// word is not, so a replayed training step still draws fresh masks.
__global__ void dropout_mask_kernel(float* __restrict__ out, long long n, float p, unsigned long long seed,
                                    const unsigned long long* __restrict__ seed_dev) {
  pdl_prologue();
  if (seed_dev) seed += *seed_dev;
  const float scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
    out[i] = u >= p ? scale : 0.f;
  }
}

// out[c] = sum over rows of a[r][c]  (any C; small matrices: classifier bias gradients)
__global__ void colsum_kernel(const float* __restrict__ a, int rows, int C, float* __restrict__ out) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += a[(size_t)r * C + c];
  out[c] = s;
}

// out = (a + b) * mask   (fp32; b / mask optional)
__global__ void add_mul_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mask,
                               float* __restrict__ out, long long n) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = a[i];
    if (b) v += b[i];
    if (mask) v *= mask[i];
    out[i] = v;
  }
}

// fp32 [M][C] -> T [M][C] (and back): gradient hand-over between the fp32 attention block and the bf16 backbone.
template <typename S, typename D>
__global__ void convert_kernel(const S* __restrict__ src, D* __restrict__ dst, long long n) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v;
    if constexpr (sizeof(S) == 2) v = __bfloat162float(src[i]); else v = src[i];
    if constexpr (sizeof(D) == 2) dst[i] = __float2bfloat16_rn(v); else dst[i] = v;
  }
}

// Row chunks per image: B * chunks CTAs fill two waves of a 2-CTA/SM grid without a ragged third wave, and a CTA
// keeps at least 32 rows (small late-layer tensors: fewer, fatter CTAs and fewer partial rows to finish).
static inline long long chunks_for(int B, long long rows_per_image) {
  long long chunks = (4LL * num_sms()) / B;
  if (chunks > rows_per_image / 32) chunks = rows_per_image / 32;
  if (chunks < 1) chunks = 1;
  return chunks;
}

}  // namespace dfv

using namespace dfv;

extern "C" {

int dfv_rows_chunks(int B, long long rows_per_image) {
  if (B <= 0 || rows_per_image <= 0) return DFV_ERR_INVALID;
  return (int)chunks_for(B, rows_per_image);
}

size_t dfv_bn_ws_floats(int B, long long rows_per_image, int C) {
  if (B <= 0 || rows_per_image <= 0 || C <= 0) return 0;
  /* >= 2C doubles (the per-channel sum / sum-of-squares accumulators); 4C floats per partial row: dfv_act_bn_bwd_gated_reduce */
  return std::max<size_t>((size_t)B * chunks_for(B, rows_per_image) * 4 * C, 4 * (size_t)C + 16);
}

int dfv_bn_stats_fwd(const void* raw, int dtype, int B, long long rows_per_image, int C, float eps, float momentum,
                     float* mean, float* invstd, float* running_mean, float* running_var, float* ws,
                     dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(raw && mean && invstd && ws, "dfv_bn_stats_fwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype) && B > 0 && rows_per_image > 0 && C > 0 && C % 8 == 0, "dfv_bn_stats_fwd: bad shape (C %% 8)");
  cudaStream_t st = as_stream(stream);
  const long long chunks = chunks_for(B, rows_per_image);
  const long long rpc = (rows_per_image + chunks - 1) / chunks;
  dim3 grid((unsigned)chunks, (unsigned)B);
  ProfScope prof(PK_BN, (double)B * rows_per_image * C * dtype_size(dtype), 3.0 * B * rows_per_image * C, st);
  if (dtype == DFV_BF16) DFV_PDL((bn_stats_kernel<__nv_bfloat16>), grid, kNT, 0, st, (const __nv_bfloat16*)raw, rows_per_image, C, rpc, ws);
  else DFV_PDL((bn_stats_kernel<float>), grid, kNT, 0, st, (const float*)raw, rows_per_image, C, rpc, ws);
  DFV_LAUNCH_CHECK();
  DFV_PDL(bn_stats_finalize_kernel, (C + 31) / 32, kFinLanes * 32, 0, st, ws, (int)(chunks * B), C, (double)B * (double)rows_per_image, eps,
                                                         momentum, mean, invstd, running_mean, running_var);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

/* mean / invstd (+ running-stat update) from per-channel double sums acc[0..C) = sum x, acc[C..2C) = sum x^2. */
int dfv_bn_stats_from_sums(const double* acc, int C, double count, float eps, float momentum, float* mean, float* invstd,
                           float* running_mean, float* running_var, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(acc && mean && invstd && C > 0 && count > 0, "dfv_bn_stats_from_sums: bad arguments");
  DFV_PDL(bn_stats_from_sums_kernel, (C + 255) / 256, 256, 0, as_stream(stream), acc, C, count, eps, momentum, mean, invstd, running_mean,
                                                                          running_var);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_bn_act_fwd(const void* raw, const float* mean, const float* invstd, const float* gamma, const float* beta, int act,
                   const float* rowscale, const void* residual, const float* mask, void* out, float* pool_partial,
                   int dtype, int B, long long rows_per_image, int C, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(raw && out, "dfv_bn_act_fwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype) && B > 0 && rows_per_image > 0 && C > 0 && C % 8 == 0, "dfv_bn_act_fwd: bad shape (C %% 8)");
  DFV_REQUIRE(act >= DFV_ACT_NONE && act <= DFV_ACT_RELU, "dfv_bn_act_fwd: bad act %d", act);
  cudaStream_t st = as_stream(stream);
  const long long chunks = chunks_for(B, rows_per_image);
  const long long rpc = (rows_per_image + chunks - 1) / chunks;
  dim3 grid((unsigned)chunks, (unsigned)B);
  ProfScope prof(PK_BN, (double)B * rows_per_image * C * dtype_size(dtype) * (residual ? 3.0 : 2.0), 8.0 * B * rows_per_image * C, st);
  if (dtype == DFV_BF16) {
    constexpr int kStage = 2 * kPipeU * 2 * kNT * 16;
    static thread_local bool configured = false;
    if (!configured) {
      DFV_CUDA(cudaFuncSetAttribute(bn_act_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStage));
      configured = true;
    }
    if (act == DFV_ACT_SILU && !rowscale && !residual && !mask) {
      static thread_local bool configured_s = false;
      if (!configured_s) {
        DFV_CUDA(cudaFuncSetAttribute(bn_act_kernel<__nv_bfloat16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStage));
        configured_s = true;
      }
      DFV_PDL((bn_act_kernel<__nv_bfloat16, true, true>), grid, kNT, kStage, st, (const __nv_bfloat16*)raw, mean, invstd, gamma, beta, act,
                                                                        rowscale, (const __nv_bfloat16*)residual, mask, (__nv_bfloat16*)out,
                                                                        pool_partial, rows_per_image, C, rpc);
    } else {
      DFV_PDL((bn_act_kernel<__nv_bfloat16, true>), grid, kNT, kStage, st, (const __nv_bfloat16*)raw, mean, invstd, gamma, beta, act, rowscale,
                                                                  (const __nv_bfloat16*)residual, mask, (__nv_bfloat16*)out,
                                                                  pool_partial, rows_per_image, C, rpc);
    }
  } else {
    DFV_PDL((bn_act_kernel<float, false>), grid, kNT, 0, st, (const float*)raw, mean, invstd, gamma, beta, act, rowscale,
                                                    (const float*)residual, mask, (float*)out, pool_partial, rows_per_image, C, rpc);
  }
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_act_bn_bwd(const void* g, const void* raw, const float* mean, const float* invstd, const float* gamma,
                   const float* beta, int act, const void* gate, const float* dpool, float inv_hw, const float* rowscale,
                   const float* mask, void* du, float* dgamma, float* dbeta, float* coef, float* ws, int dtype, int B,
                   long long rows_per_image, int C, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(g && raw && coef && ws, "dfv_act_bn_bwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype) && B > 0 && rows_per_image > 0 && C > 0 && C % 8 == 0, "dfv_act_bn_bwd: bad shape (C %% 8)");
  cudaStream_t st = as_stream(stream);
  const long long chunks = chunks_for(B, rows_per_image);
  const long long rpc = (rows_per_image + chunks - 1) / chunks;
  dim3 grid((unsigned)chunks, (unsigned)B);
  ProfScope prof(PK_BN, (du ? 3.0 : 2.0) * B * rows_per_image * C * dtype_size(dtype), 12.0 * B * rows_per_image * C, st);
  const bool gated = gate != nullptr || dpool != nullptr;
#define ABB(T_, G_, U_, M_, SMEM_)                                                                                                   \
  do {                                                                                                                              \
    static thread_local bool configured = false;                                                                                    \
    if (SMEM_ > 0 && !configured) {                                                                                                 \
      DFV_CUDA(cudaFuncSetAttribute(act_bn_bwd_kernel<T_, G_, U_, M_>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_));        \
      configured = true;                                                                                                            \
    }                                                                                                                               \
    DFV_PDL((act_bn_bwd_kernel<T_, G_, U_, M_>), grid, kNT, SMEM_, st, (const T_*)g, (const T_*)raw, mean, invstd, gamma, beta, act,       \
                                                               (const T_*)gate, dpool, inv_hw, rowscale, mask, (T_*)du, ws,        \
                                                               rows_per_image, C, rpc, (const float*)nullptr, (float*)nullptr);    \
  } while (0)
#define ABB_SPEC(G_, SMEM_)                                                                                                          \
  do {                                                                                                                              \
    static thread_local bool configured = false;                                                                                    \
    if (!configured) {                                                                                                              \
      DFV_CUDA(cudaFuncSetAttribute(act_bn_bwd_kernel<__nv_bfloat16, G_, 4, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_)); \
      configured = true;                                                                                                            \
    }                                                                                                                               \
    DFV_PDL((act_bn_bwd_kernel<__nv_bfloat16, G_, 4, 2, false, true>), grid, kNT, SMEM_, st, (const __nv_bfloat16*)g, (const __nv_bfloat16*)raw, mean, \
            invstd, gamma, beta, act, (const __nv_bfloat16*)gate, dpool, inv_hw, rowscale, mask, (__nv_bfloat16*)du, ws, rows_per_image, C, rpc,     \
            (const float*)nullptr, (float*)nullptr);                                                                                \
  } while (0)
  constexpr int kStageBytes = 2 * kPipeU * 2 * kNT * 16;   // cp.async staging of the bf16 kernels
  const bool spec = act == DFV_ACT_SILU && !rowscale && !mask;
  if (dtype == DFV_BF16) {
    if (spec) { if (gated) ABB_SPEC(true, kStageBytes); else ABB_SPEC(false, kStageBytes); }
    else if (gated) ABB(__nv_bfloat16, true, 4, 2, kStageBytes); else ABB(__nv_bfloat16, false, 4, 2, kStageBytes);
  } else {
    if (gated) ABB(float, true, 2, 2, 0); else ABB(float, false, 2, 2, 0);
  }
#undef ABB
#undef ABB_SPEC
  DFV_LAUNCH_CHECK();
  DFV_PDL(bn_bwd_finalize_kernel, (C + 31) / 32, kFinLanes * 32, 0, st, ws, (int)(chunks * B), C, (double)B * (double)rows_per_image, dgamma, dbeta, coef);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_bn_bwd_apply(const void* du, const void* raw, const float* mean, const float* invstd, const float* gamma,
                     const float* coef, void* draw, int dtype, long long M, int C, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(du && raw && mean && invstd && coef && draw, "dfv_bn_bwd_apply: null pointer");
  DFV_REQUIRE(valid_dtype(dtype) && M > 0 && C > 0 && C % 8 == 0, "dfv_bn_bwd_apply: bad shape (C %% 8)");
  cudaStream_t st = as_stream(stream);
  long long blocks = std::min<long long>(M, 8LL * num_sms());
  const long long rpc = (M + blocks - 1) / blocks;
  blocks = (M + rpc - 1) / rpc;
  ProfScope prof(PK_BN, 3.0 * M * C * dtype_size(dtype), 6.0 * M * C, st);
  if (dtype == DFV_BF16)
    DFV_PDL((bn_bwd_apply_kernel<__nv_bfloat16>), (unsigned)blocks, kNT, 0, st, (const __nv_bfloat16*)du, (const __nv_bfloat16*)raw, mean, invstd,
                                                                        gamma, coef, (__nv_bfloat16*)draw, M, C, rpc);
  else
    DFV_PDL((bn_bwd_apply_kernel<float>), (unsigned)blocks, kNT, 0, st, (const float*)du, (const float*)raw, mean, invstd, gamma, coef, (float*)draw, M, C, rpc);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_act_bn_bwd_apply(const void* g, const void* raw, const float* mean, const float* invstd, const float* gamma, const float* beta,
                         int act, const void* gate, const float* dpool, float inv_hw, const float* rowscale, const float* mask,
                         const float* coef, void* draw, int dtype, int B, long long rows_per_image, int C, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(g && raw && coef && draw, "dfv_act_bn_bwd_apply: null pointer");
  DFV_REQUIRE(valid_dtype(dtype) && B > 0 && rows_per_image > 0 && C > 0 && C % 8 == 0, "dfv_act_bn_bwd_apply: bad shape (C %% 8)");
  cudaStream_t st = as_stream(stream);
  const long long chunks = chunks_for(B, rows_per_image);
  const long long rpc = (rows_per_image + chunks - 1) / chunks;
  dim3 grid((unsigned)chunks, (unsigned)B);
  ProfScope prof(PK_BN, 3.0 * B * rows_per_image * C * dtype_size(dtype), 14.0 * B * rows_per_image * C, st);
  const bool gated = gate != nullptr || dpool != nullptr;
#define ABA(T_, G_, U_, M_, SMEM_)                                                                                                  \
  do {                                                                                                                              \
    static thread_local bool configured = false;                                                                                    \
    if (SMEM_ > 0 && !configured) {                                                                                                 \
      DFV_CUDA(cudaFuncSetAttribute(act_bn_bwd_kernel<T_, G_, U_, M_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_));  \
      configured = true;                                                                                                            \
    }                                                                                                                               \
    DFV_PDL((act_bn_bwd_kernel<T_, G_, U_, M_, true>), grid, kNT, SMEM_, st, (const T_*)g, (const T_*)raw, mean, invstd, gamma, beta, act,  \
                                                               (const T_*)gate, dpool, inv_hw, rowscale, mask, (T_*)draw, (float*)nullptr, \
                                                               rows_per_image, C, rpc, coef, (float*)nullptr);                     \
  } while (0)
  constexpr int kStageBytes = 2 * kPipeU * 2 * kNT * 16;   // cp.async staging of the bf16 kernels
#define ABA_SPEC(G_, SMEM_)                                                                                                          \
  do {                                                                                                                              \
    static thread_local bool configured = false;                                                                                    \
    if (!configured) {                                                                                                              \
      DFV_CUDA(cudaFuncSetAttribute(act_bn_bwd_kernel<__nv_bfloat16, G_, 4, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_)); \
      configured = true;                                                                                                            \
    }                                                                                                                               \
    DFV_PDL((act_bn_bwd_kernel<__nv_bfloat16, G_, 4, 2, true, true>), grid, kNT, SMEM_, st, (const __nv_bfloat16*)g, (const __nv_bfloat16*)raw, mean, \
            invstd, gamma, beta, act, (const __nv_bfloat16*)gate, dpool, inv_hw, rowscale, mask, (__nv_bfloat16*)draw, (float*)nullptr,             \
            rows_per_image, C, rpc, coef, (float*)nullptr);                                                                         \
  } while (0)
  const bool spec = act == DFV_ACT_SILU && !rowscale && !mask;
  if (dtype == DFV_BF16) {
    if (spec) { if (gated) ABA_SPEC(true, kStageBytes); else ABA_SPEC(false, kStageBytes); }
    else if (gated) ABA(__nv_bfloat16, true, 4, 2, kStageBytes); else ABA(__nv_bfloat16, false, 4, 2, kStageBytes);
  } else {
    if (gated) ABA(float, true, 2, 2, 0); else ABA(float, false, 2, 2, 0);
  }
#undef ABA
#undef ABA_SPEC
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

/* Gated swish layers (the depthwise output), bf16: the BatchNorm-backward reduction with the SE gate / dpool terms DEFERRED
 * (act_bn_bwd_kernel kDefer).  One pass over (g, raw) writes ws4 [B][chunks][4][C] and the SE backward's dot partials
 * partial_d [B][chunks][C]; after dfv_se_bwd_from_partials has produced dpool, dfv_bn_bwd_gated_finalize forms dgamma / dbeta /
 * coef -- exactly what dfv_act_bn_bwd(gate, dpool) returns, without the separate pass over (g, d) of dfv_se_bwd. */
int dfv_act_bn_bwd_gated_reduce(const void* g, const void* raw, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                float* ws4, float* partial_d, int dtype, int B, long long rows_per_image, int C, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(g && raw && mean && invstd && ws4 && partial_d, "dfv_act_bn_bwd_gated_reduce: null pointer");
  DFV_REQUIRE(dtype == DFV_BF16 && B > 0 && rows_per_image > 0 && C > 0 && C % 8 == 0, "dfv_act_bn_bwd_gated_reduce: bf16 tensors, C %% 8 == 0");
  cudaStream_t st = as_stream(stream);
  const long long chunks = chunks_for(B, rows_per_image);
  const long long rpc = (rows_per_image + chunks - 1) / chunks;
  dim3 grid((unsigned)chunks, (unsigned)B);
  ProfScope prof(PK_BN, 2.0 * B * rows_per_image * C * dtype_size(dtype), 16.0 * B * rows_per_image * C, st);
  constexpr int kStageBytes = 2 * kPipeU * 2 * kNT * 16;
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(act_bn_bwd_kernel<__nv_bfloat16, true, 4, 2, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kStageBytes));
    configured = true;
  }
  DFV_PDL((act_bn_bwd_kernel<__nv_bfloat16, true, 4, 2, false, true, true>), grid, kNT, kStageBytes, st, (const __nv_bfloat16*)g,
          (const __nv_bfloat16*)raw, mean, invstd, gamma, beta, (int)DFV_ACT_SILU, (const __nv_bfloat16*)nullptr, (const float*)nullptr, 0.f,
          (const float*)nullptr, (const float*)nullptr, (__nv_bfloat16*)nullptr, ws4, rows_per_image, C, rpc, (const float*)nullptr, partial_d);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_bn_bwd_gated_finalize(const float* ws4, const void* gate, const float* dpool, float inv_hw, const float* mean, const float* invstd,
                              float* dgamma, float* dbeta, float* coef, int dtype, int B, long long rows_per_image, int C,
                              dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(ws4 && gate && dpool && mean && invstd && coef, "dfv_bn_bwd_gated_finalize: null pointer");
  DFV_REQUIRE(dtype == DFV_BF16 && B > 0 && rows_per_image > 0 && C > 0, "dfv_bn_bwd_gated_finalize: bad arguments");
  const long long chunks = chunks_for(B, rows_per_image);
  DFV_PDL((bn_bwd_finalize_defer_kernel<__nv_bfloat16>), (C + 31) / 32, kFinLanes * 32, 0, as_stream(stream), ws4, (int)chunks, B, C,
          (const __nv_bfloat16*)gate, dpool, inv_hw, mean, invstd, (double)B * (double)rows_per_image, dgamma, dbeta, coef);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

/* SE backward.  da: gradient wrt the gated tensor (d (.) gate), d: the activated depthwise output; both [B][rows][C].
 * Outputs: dpool [B][C] (to be folded into the depthwise-activation gradient: + dpool / HW) and the four
 * parameter gradients in torch layout.  ws: fp32, dfv_se_bwd_ws_floats(). */
size_t dfv_se_bwd_ws_floats(int B, long long rows_per_image, int C, int squeeze) {
  if (B <= 0 || rows_per_image <= 0 || C <= 0 || squeeze <= 0) return 0;
  return (size_t)B * chunks_for(B, rows_per_image) * C + (size_t)B * C + (size_t)B * squeeze;
}

int dfv_se_bwd(const void* da, const void* d, int dtype, const float* gate_f32, const float* pooled, const float* h1,
               const float* w_reduce, const float* w_expand, float* dpool, float* dw_reduce, float* db_reduce,
               float* dw_expand, float* db_expand, float* ws, int B, long long rows_per_image, int C, int squeeze,
               dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(da && d && ws, "dfv_se_bwd: null pointer");
  DFV_REQUIRE(valid_dtype(dtype) && B > 0 && rows_per_image > 0 && C > 0 && C % 8 == 0 && squeeze > 0, "dfv_se_bwd: bad shape");
  cudaStream_t st = as_stream(stream);
  const long long chunks = chunks_for(B, rows_per_image);
  const long long rpc = (rows_per_image + chunks - 1) / chunks;
  dim3 grid((unsigned)chunks, (unsigned)B);
  {
    ProfScope prof(PK_SE_GATE, 2.0 * B * rows_per_image * C * dtype_size(dtype), 2.0 * B * rows_per_image * C, st);
    if (dtype == DFV_BF16)
      DFV_PDL((dot_rows_kernel<__nv_bfloat16>), grid, kNT, 0, st, (const __nv_bfloat16*)da, (const __nv_bfloat16*)d, rows_per_image, C, rpc, ws);
    else
      DFV_PDL((dot_rows_kernel<float>), grid, kNT, 0, st, (const float*)da, (const float*)d, rows_per_image, C, rpc, ws);
    DFV_LAUNCH_CHECK();
  }
  return dfv_se_bwd_from_partials(gate_f32, pooled, h1, w_reduce, w_expand, dpool, dw_reduce, db_reduce, dw_expand, db_expand, ws, B,
                                  rows_per_image, C, squeeze, stream);
}

/* dfv_se_bwd without its streaming pass: ws[0 .. B * dfv_rows_chunks * C) already holds the per-(image, chunk) partial sums of
 * da * d (written by dfv_act_bn_bwd_gated_reduce, which forms them while it reads the same tensors for the BatchNorm reduction). */
int dfv_se_bwd_from_partials(const float* gate_f32, const float* pooled, const float* h1, const float* w_reduce, const float* w_expand,
                             float* dpool, float* dw_reduce, float* db_reduce, float* dw_expand, float* db_expand, float* ws, int B,
                             long long rows_per_image, int C, int squeeze, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(gate_f32 && pooled && h1 && w_reduce && w_expand && dpool && dw_reduce && db_reduce && dw_expand && db_expand && ws,
              "dfv_se_bwd: null pointer");
  DFV_REQUIRE(B > 0 && rows_per_image > 0 && C > 0 && C % 8 == 0 && squeeze > 0, "dfv_se_bwd: bad shape");
  cudaStream_t st = as_stream(stream);
  const long long chunks = chunks_for(B, rows_per_image);
  float* partial = ws;
  float* dz = partial + (size_t)B * chunks * C;
  float* dh1 = dz + (size_t)B * C;
  ProfScope prof(PK_SE_GATE, 4.0 * ((double)B * chunks * C + 3.0 * B * C + 4.0 * C * squeeze), 6.0 * B * (double)C * squeeze, st);
  DFV_REQUIRE(squeeze <= 128, "dfv_se_bwd: squeeze width %d > 128", squeeze);
  {
    const int img = B >= 16 ? 4 : 1;
    const int cper = (C + kSeBwdCluster - 1) / kSeBwdCluster;
    const size_t smem = sizeof(float) * ((size_t)img * cper + (size_t)2 * img * squeeze + (size_t)8 * img * squeeze);
    const unsigned grid_i = (unsigned)((B + img - 1) / img) * kSeBwdCluster;
    if (img == 4)
      DFV_PDL((se_bwd_image_kernel<4>), grid_i, 256, smem, st, partial, (int)chunks, gate_f32, h1, w_reduce, w_expand, dz, dh1, dpool, B, C, squeeze);
    else
      DFV_PDL((se_bwd_image_kernel<1>), grid_i, 256, smem, st, partial, (int)chunks, gate_f32, h1, w_reduce, w_expand, dz, dh1, dpool, B, C, squeeze);
    DFV_LAUNCH_CHECK();
  }
  const int slices = std::max(1, std::min(squeeze / 4, (2 * num_sms() * 128 + C - 1) / C));
  const int jn_max = (squeeze + slices - 1) / slices + 1;
  const size_t smem2 = (size_t)2 * B * jn_max * sizeof(float);
  DFV_REQUIRE(smem2 <= 160 * 1024, "dfv_se_bwd: batch too large for the weight-gradient kernel (B * squeeze slice = %d)", B * jn_max);
  if (smem2 > 48 * 1024) DFV_CUDA(cudaFuncSetAttribute(se_bwd_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  DFV_PDL(se_bwd_weights_kernel, dim3((C + 127) / 128, slices), 128 * kSeWgQ, smem2, st, dz, dh1, pooled, h1, dw_reduce, db_reduce, dw_expand, db_expand, B, C, squeeze);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_cast_weight(const float* src, void* dst, int dtype, int rows, int cols, int transpose, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(src && dst && valid_dtype(dtype) && rows > 0 && cols > 0, "dfv_cast_weight: bad arguments");
  const long long n = (long long)rows * cols;
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, 4LL * num_sms());
  if (dtype == DFV_BF16) DFV_PDL((cast_weight_kernel<__nv_bfloat16>), blocks, 256, 0, as_stream(stream), src, (__nv_bfloat16*)dst, rows, cols, transpose);
  else DFV_PDL((cast_weight_kernel<float>), blocks, 256, 0, as_stream(stream), src, (float*)dst, rows, cols, transpose);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_dw_weight_pack(const float* src_ckk, float* dst_kkc, int C, int kernel, int flip, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(src_ckk && dst_kkc && C > 0 && kernel > 0, "dfv_dw_weight_pack: bad arguments");
  const int n = C * kernel * kernel;
  DFV_PDL(dw_weight_pack_kernel, (n + 255) / 256, 256, 0, as_stream(stream), src_ckk, dst_kkc, C, kernel * kernel, flip);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_dw_weight_unpack(const float* src_kkc, float* dst_ckk, int C, int kernel, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(src_kkc && dst_ckk && C > 0 && kernel > 0, "dfv_dw_weight_unpack: bad arguments");
  const int n = C * kernel * kernel;
  DFV_PDL(dw_weight_unpack_kernel, (n + 255) / 256, 256, 0, as_stream(stream), src_kkc, dst_ckk, C, kernel * kernel);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_dropout_mask(float* out, long long n, float p, unsigned long long seed, dfv_stream_t stream) {
  return dfv_dropout_mask_dev(out, n, p, seed, nullptr, stream);
}

int dfv_dropout_mask_dev(float* out, long long n, float p, unsigned long long seed, const unsigned long long* seed_dev, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(out && n > 0 && p >= 0.f && p <= 1.f, "dfv_dropout_mask: bad arguments");
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, 4LL * num_sms());
  DFV_PDL(dropout_mask_kernel, blocks, 256, 0, as_stream(stream), out, n, p, seed, seed_dev);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_colsum(const float* a, int rows, int C, float* out, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(a && out && rows > 0 && C > 0, "dfv_colsum: bad arguments");
  DFV_PDL(colsum_kernel, (C + 127) / 128, 128, 0, as_stream(stream), a, rows, C, out);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_add_mul(const float* a, const float* b, const float* mask, float* out, long long n, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(a && out && n > 0, "dfv_add_mul: bad arguments");
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, 4LL * num_sms());
  DFV_PDL(add_mul_kernel, blocks, 256, 0, as_stream(stream), a, b, mask, out, n);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

int dfv_convert(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(src && dst && n > 0 && valid_dtype(src_dtype) && valid_dtype(dst_dtype), "dfv_convert: bad arguments");
  cudaStream_t st = as_stream(stream);
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, 8LL * num_sms());
  if (src_dtype == dst_dtype) {
    DFV_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * dtype_size(src_dtype), cudaMemcpyDeviceToDevice, st));
    return DFV_OK;
  }
  if (src_dtype == DFV_F32) DFV_PDL((convert_kernel<float, __nv_bfloat16>), blocks, 256, 0, st, (const float*)src, (__nv_bfloat16*)dst, n);
  else DFV_PDL((convert_kernel<__nv_bfloat16, float>), blocks, 256, 0, st, (const __nv_bfloat16*)src, (float*)dst, n);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

}  // extern "C"
