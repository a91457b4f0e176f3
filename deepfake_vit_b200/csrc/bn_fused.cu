// Cluster-fused BatchNorm kernels for the small, L2-resident late-stage tensors (24x24 / 12x12 stages at batch 64:
// 30-50 MB per tensor).  The streaming versions (train_ops.cu) need two launches per direction with a per-CTA
// partial-row hand-off in between and, on these sizes, are bound by launch / latency (25-47 us per launch for
// 4-14 us of HBM time).  Here one launch does both phases:
//   cluster of 8 CTAs  = one 32-channel slice of the tensor, the rows split 8 ways
//   phase 1            per-channel sums over the slice (registers -> shared memory -> DISTRIBUTED shared memory)
//   cluster.sync
//   phase 2            the same rows again (now L2 hits) -> normalised / gradient output
// No global partials, no finalize kernel, each tensor crosses HBM once per direction.
#include <cooperative_groups.h>

#include "common.cuh"

namespace dfv {

constexpr int kFbCluster = 8;
constexpr int kFbCh = 32;        // channels per cluster (4 vectors of 8 -> 64-byte row segments)
constexpr int kFbLanes = 64;     // row lanes per CTA (256 threads = 4 vectors x 64 lanes)

__device__ __forceinline__ float fb_act_grad(float u, int act) {      // as act_grad<true> of train_ops.cu
  if (act == DFV_ACT_SILU) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * u));
    const float s = fmaf(0.5f, t, 0.5f);
    return s * (1.f + u * (1.f - s));
  }
  if (act == DFV_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  return 1.f;
}

// Block reduction of N per-thread values over the 64 row lanes of each vector column, then over the cluster.
// Result (all threads): tot[0..N) for this thread's vector column.
template <int N>
__device__ __forceinline__ void fb_reduce(cooperative_groups::cluster_group& cluster, float* red, float* part, const float v[N],
                                          float tot[N], int vec, int lane) {
  // stage 1: lanes -> 8 partial groups (red: [8][4][N])
#pragma unroll
  for (int e = 0; e < N; ++e) {
    float x = v[e];
    // lanes of one vector column sit at tid = lane * 4 + vec: stride 4 inside a warp -> 8 lanes per warp share a column
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    x += __shfl_xor_sync(0xffffffffu, x, 8);
    x += __shfl_xor_sync(0xffffffffu, x, 16);
    if ((lane & 7) == 0) red[((lane >> 3) * 4 + vec) * N + e] = x;
  }
  __syncthreads();
  if (threadIdx.x < 4 * N) {
    const int vv = threadIdx.x / N, e = threadIdx.x % N;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[(w * 4 + vv) * N + e];
    part[vv * N + e] = s;
  }
  cluster.sync();
#pragma unroll
  for (int e = 0; e < N; ++e) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kFbCluster; ++r) s += cluster.map_shared_rank(part, r)[vec * N + e];
    tot[e] = s;
  }
  cluster.sync();     // every peer has read `part` (it may be reused, and no CTA exits while it is being read)
}

// ------------------------------------------------------------------------------------ forward: stats + normalise + act
template <bool kFast>
__global__ void __cluster_dims__(kFbCluster, 1, 1) __launch_bounds__(256, 2)
    fused_bn_act_kernel(const __nv_bfloat16* __restrict__ raw, const float* __restrict__ gamma, const float* __restrict__ beta,
                        int act, float eps, float momentum, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                        float* __restrict__ running_mean, float* __restrict__ running_var, __nv_bfloat16* __restrict__ out,
                        long long M, int C) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float red[8 * 4 * 16];
  __shared__ float part[4 * 16];
  const int rank = (int)cluster.block_rank();
  const int vec = threadIdx.x & 3, lane = threadIdx.x >> 2;
  const int c = (blockIdx.x / kFbCluster) * kFbCh + vec * 8;
  const long long rows_per = (M + kFbCluster - 1) / kFbCluster;
  const long long m0 = rank * rows_per, m1 = min(M, m0 + rows_per);

  float acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.f;
  for (long long r = m0 + lane; r < m1; r += 4 * kFbLanes) {
    uint4 x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      x[i] = r + i * kFbLanes < m1 ? *reinterpret_cast<const uint4*>(raw + (size_t)(r + i * kFbLanes) * C + c) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t w[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = bf16_lo(w[j]), b = bf16_hi(w[j]);
        acc[2 * j] += a;
        acc[2 * j + 1] += b;
        acc[8 + 2 * j] = fmaf(a, a, acc[8 + 2 * j]);
        acc[8 + 2 * j + 1] = fmaf(b, b, acc[8 + 2 * j + 1]);
      }
    }
  }
  float tot[16];
  fb_reduce<16>(cluster, red, part, acc, tot, vec, lane);

  float sc[8], sh[8];
  const double count = (double)M;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const double mu = (double)tot[e] / count;
    double var = (double)tot[8 + e] / count - mu * mu;
    if (var < 0.0) var = 0.0;
    const float is = 1.0f / sqrtf((float)var + eps);
    sc[e] = gamma[c + e] * is;
    sh[e] = beta[c + e] - (float)mu * sc[e];
    if (rank == 0 && lane == 0) {
      mean_out[c + e] = (float)mu;
      invstd_out[c + e] = is;
      if (running_mean) running_mean[c + e] = (1.f - momentum) * running_mean[c + e] + momentum * (float)mu;
      if (running_var) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_var[c + e] = (1.f - momentum) * running_var[c + e] + momentum * (float)unbiased;
      }
    }
  }
  for (long long r = m0 + lane; r < m1; r += 4 * kFbLanes) {
    uint4 x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (r + i * kFbLanes < m1) x[i] = *reinterpret_cast<const uint4*>(raw + (size_t)(r + i * kFbLanes) * C + c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (r + i * kFbLanes < m1) {
        float v[8];
        v[0] = bf16_lo(x[i].x); v[1] = bf16_hi(x[i].x); v[2] = bf16_lo(x[i].y); v[3] = bf16_hi(x[i].y);
        v[4] = bf16_lo(x[i].z); v[5] = bf16_hi(x[i].z); v[6] = bf16_lo(x[i].w); v[7] = bf16_hi(x[i].w);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float u = fmaf(v[e], sc[e], sh[e]);
          v[e] = act == DFV_ACT_SILU ? silu<kFast>(u) : (act == DFV_ACT_RELU ? fmaxf(u, 0.f) : u);
        }
        store8(out + (size_t)(r + i * kFbLanes) * C + c, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ backward: act' + BN backward
// d = (g * gate[b] + dpool[b] / HW) * act'(u);  d raw = gamma * invstd * (d - mean(d) - xhat * mean(d * xhat))
template <bool kGate>
__global__ void __cluster_dims__(kFbCluster, 1, 1) __launch_bounds__(256, 2)
    fused_act_bn_bwd_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ raw, const float* __restrict__ mean,
                            const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                            const __nv_bfloat16* __restrict__ gate, const float* __restrict__ dpool, float inv_hw,
                            __nv_bfloat16* __restrict__ draw, float* __restrict__ dgamma, float* __restrict__ dbeta, long long M,
                            long long rows_per_image, int C) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float red[8 * 4 * 16];
  __shared__ float part[4 * 16];
  const int rank = (int)cluster.block_rank();
  const int vec = threadIdx.x & 3, lane = threadIdx.x >> 2;
  const int c = (blockIdx.x / kFbCluster) * kFbCh + vec * 8;
  const long long rows_per = (M + kFbCluster - 1) / kFbCluster;
  const long long m0 = rank * rows_per, m1 = min(M, m0 + rows_per);

  float is[8], nm[8], ga[8], be[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    is[e] = invstd[c + e];
    nm[e] = -mean[c + e] * is[e];
    ga[e] = gamma ? gamma[c + e] : 1.f;
    be[e] = beta ? beta[c + e] : 0.f;
  }
  // d and xhat of one row vector
  auto row_d = [&](long long r, const uint4& gx, const uint4& xx, float d[8], float xh[8]) {
    float gv[8], x[8];
    gv[0] = bf16_lo(gx.x); gv[1] = bf16_hi(gx.x); gv[2] = bf16_lo(gx.y); gv[3] = bf16_hi(gx.y);
    gv[4] = bf16_lo(gx.z); gv[5] = bf16_hi(gx.z); gv[6] = bf16_lo(gx.w); gv[7] = bf16_hi(gx.w);
    x[0] = bf16_lo(xx.x); x[1] = bf16_hi(xx.x); x[2] = bf16_lo(xx.y); x[3] = bf16_hi(xx.y);
    x[4] = bf16_lo(xx.z); x[5] = bf16_hi(xx.z); x[6] = bf16_lo(xx.w); x[7] = bf16_hi(xx.w);
    float gt[8], dp[8];
    if constexpr (kGate) {
      const long long b = r / rows_per_image;
      load8(gate + (size_t)b * C + c, gt);
      load8(dpool + (size_t)b * C + c, dp);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float gi = gv[e];
      if constexpr (kGate) gi = fmaf(gv[e], gt[e], dp[e] * inv_hw);
      xh[e] = fmaf(x[e], is[e], nm[e]);
      const float u = fmaf(xh[e], ga[e], be[e]);
      d[e] = gi * fb_act_grad(u, act);
    }
  };

  float acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.f;
  for (long long r = m0 + lane; r < m1; r += 2 * kFbLanes) {
    uint4 gx[2], xx[2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (r + i * kFbLanes < m1) {
        gx[i] = *reinterpret_cast<const uint4*>(g + (size_t)(r + i * kFbLanes) * C + c);
        xx[i] = *reinterpret_cast<const uint4*>(raw + (size_t)(r + i * kFbLanes) * C + c);
      }
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (r + i * kFbLanes < m1) {
        float d[8], xh[8];
        row_d(r + i * kFbLanes, gx[i], xx[i], d, xh);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          acc[e] += d[e];
          acc[8 + e] = fmaf(d[e], xh[e], acc[8 + e]);
        }
      }
  }
  float tot[16];
  fb_reduce<16>(cluster, red, part, acc, tot, vec, lane);

  float gi[8], c1[8], c2[8];
  const float inv_count = 1.0f / (float)M;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    gi[e] = ga[e] * is[e];
    c1[e] = tot[e] * inv_count;
    c2[e] = tot[8 + e] * inv_count;
    if (rank == 0 && lane == 0) {
      if (dbeta) dbeta[c + e] = tot[e];
      if (dgamma) dgamma[c + e] = tot[8 + e];
    }
  }
  for (long long r = m0 + lane; r < m1; r += 2 * kFbLanes) {
    uint4 gx[2], xx[2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (r + i * kFbLanes < m1) {
        gx[i] = *reinterpret_cast<const uint4*>(g + (size_t)(r + i * kFbLanes) * C + c);
        xx[i] = *reinterpret_cast<const uint4*>(raw + (size_t)(r + i * kFbLanes) * C + c);
      }
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (r + i * kFbLanes < m1) {
        float d[8], xh[8];
        row_d(r + i * kFbLanes, gx[i], xx[i], d, xh);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] = gi[e] * (d[e] - c1[e] - xh[e] * c2[e]);
        store8(draw + (size_t)(r + i * kFbLanes) * C + c, d);
      }
  }
}

}  // namespace dfv

using namespace dfv;

extern "C" {

/* 1 if the cluster-fused BatchNorm kernels apply: bf16, C a multiple of 32, and a tensor small enough that the second
 * phase re-reads it (and its gradient) from L2. */
int dfv_bn_fused_applicable(int dtype, long long M, int C) {
  return dtype == DFV_BF16 && C % kFbCh == 0 && M >= 8 * kFbLanes && (double)M * C * 2.0 <= 52.0e6 && !(debug_flags() & 64);
}

/* Train-mode BatchNorm forward in one launch: batch statistics (+ running-stat update), normalise, activation.
 * raw / out: bf16 [M][C]; mean / invstd: saved for backward. */
int dfv_bn_fused_fwd(const void* raw, const float* gamma, const float* beta, int act, float eps, float momentum, float* mean,
                     float* invstd, float* running_mean, float* running_var, void* out, long long M, int C, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(raw && gamma && beta && mean && invstd && out && M > 0 && C > 0 && C % kFbCh == 0, "dfv_bn_fused_fwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  ProfScope prof(PK_BN, 2.0 * (double)M * C * 2.0, 10.0 * (double)M * C, st);
  fused_bn_act_kernel<true><<<(unsigned)(C / kFbCh) * kFbCluster, 256, 0, st>>>((const __nv_bfloat16*)raw, gamma, beta, act, eps, momentum, mean,
                                                                                invstd, running_mean, running_var, (__nv_bfloat16*)out, M, C);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

/* Activation + BatchNorm backward in one launch (dfv_act_bn_bwd + dfv_bn_bwd_apply): g, raw -> d raw (may alias g),
 * dgamma, dbeta.  gate [B][C] bf16 and dpool [B][C] fp32 (both or neither) fold the SE gate / pool gradient in. */
int dfv_bn_fused_bwd(const void* g, const void* raw, const float* mean, const float* invstd, const float* gamma, const float* beta,
                     int act, const void* gate, const float* dpool, float inv_hw, void* draw, float* dgamma, float* dbeta,
                     long long M, long long rows_per_image, int C, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(g && raw && mean && invstd && draw && M > 0 && rows_per_image > 0 && C > 0 && C % kFbCh == 0, "dfv_bn_fused_bwd: bad arguments");
  DFV_REQUIRE((gate == nullptr) == (dpool == nullptr), "dfv_bn_fused_bwd: gate and dpool go together");
  cudaStream_t st = as_stream(stream);
  ProfScope prof(PK_BN, 3.0 * (double)M * C * 2.0, 24.0 * (double)M * C, st);
  const unsigned grid = (unsigned)(C / kFbCh) * kFbCluster;
  if (gate)
    fused_act_bn_bwd_kernel<true><<<grid, 256, 0, st>>>((const __nv_bfloat16*)g, (const __nv_bfloat16*)raw, mean, invstd, gamma, beta, act,
                                                       (const __nv_bfloat16*)gate, dpool, inv_hw, (__nv_bfloat16*)draw, dgamma, dbeta, M,
                                                       rows_per_image, C);
  else
    fused_act_bn_bwd_kernel<false><<<grid, 256, 0, st>>>((const __nv_bfloat16*)g, (const __nv_bfloat16*)raw, mean, invstd, gamma, beta, act,
                                                        nullptr, nullptr, 0.f, (__nv_bfloat16*)draw, dgamma, dbeta, M, rows_per_image, C);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

}  // extern "C"
