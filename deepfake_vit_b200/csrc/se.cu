// Squeeze-excite gate: finishes the depthwise kernel's pool partial sums, then
// C -> squeeze (bias, swish) -> C (bias, sigmoid).  One CTA per image; all fp32.
// Negligible bytes (pool partials + the two small FC matrices, L2-resident).
#include <cooperative_groups.h>

#include "common.cuh"

namespace dfv {

// One thread-block CLUSTER of 8 CTAs per group of IMG images.  Phase 1 splits the squeeze rows over the cluster's
// CTAs (each weight row is read once per image group and reused for IMG images), the hidden vector is exchanged
// through distributed shared memory, phase 2 splits the channels.  The one-CTA-per-image version spent 80-170 us
// per late layer walking both FC matrices from L2 with a single CTA's worth of loads in flight.
constexpr int kSeCluster = 8;

// kTrain: the expand weight arrives in torch layout [C][sq] (not transposed) and the kernel also saves what the
// backward pass needs: pooled [B][C], h1 [B][sq] (pre-swish) and the fp32 gate.
template <int IMG, typename GT, bool kTrain>
__global__ void __cluster_dims__(kSeCluster, 1, 1) __launch_bounds__(256, 2)
    se_gate_kernel(const float* __restrict__ partial, int parts, float inv_hw, const float* __restrict__ w1,
                   const float* __restrict__ b1, const float* __restrict__ w2t, const float* __restrict__ b2,
                   GT* __restrict__ gate, int B, int C, int sq, float* __restrict__ pooled_out,
                   float* __restrict__ h1_out, float* __restrict__ gate_f32) {
  pdl_prologue();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float sm[];
  float* pooled = sm;                         // [IMG][C]
  float* hidden = sm + (size_t)IMG * C;       // [IMG][sq]   (all squeeze rows, gathered)
  float* mine = hidden + (size_t)IMG * sq;    // [IMG][jn]   (this CTA's slice)
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / kSeCluster) * IMG, tid = threadIdx.x;
  const int jper = (sq + kSeCluster - 1) / kSeCluster;
  const int j0 = min(sq, rank * jper), j1 = min(sq, j0 + jper);

  // pooled[im][c] = mean over positions (finishing the producer's partial sums).  No integer division in the
  // loops (the first version spent half its instructions on i / C and i % C): channels advance by blockDim,
  // images are a compile-time inner dimension, IMG x parts independent loads in flight per thread.
  for (int c = tid; c < C; c += blockDim.x) {
    float s[IMG];
#pragma unroll
    for (int im = 0; im < IMG; ++im) s[im] = 0.f;
    for (int t = 0; t < parts; ++t) {
#pragma unroll
      for (int im = 0; im < IMG; ++im)
        if (b0 + im < B) s[im] += __ldg(partial + ((size_t)(b0 + im) * parts + t) * C + c);
    }
#pragma unroll
    for (int im = 0; im < IMG; ++im) {
      const float v = s[im] * inv_hw;
      pooled[im * C + c] = v;
      if constexpr (kTrain) {
        if (rank == 0 && b0 + im < B) pooled_out[(size_t)(b0 + im) * C + c] = v;
      }
    }
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  for (int j = j0 + warp; j < j1; j += nwarps) {
    const float* wr = w1 + (size_t)j * C;
    float s[IMG];
#pragma unroll
    for (int im = 0; im < IMG; ++im) s[im] = 0.f;
    if ((C & 3) == 0 && (reinterpret_cast<uintptr_t>(w1) & 15) == 0) {
      // 16-byte weight loads, the whole row slice of a lane (up to 12 loads = 1536 channels) requested before use
      for (int cb = lane * 4; cb < C; cb += 12 * 128) {
        float4 wv[12];
#pragma unroll
        for (int u = 0; u < 12; ++u)
          wv[u] = cb + u * 128 < C ? __ldg(reinterpret_cast<const float4*>(wr + cb + u * 128)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 12; ++u) {
          const int c = cb + u * 128;
          if (c < C) {
#pragma unroll
            for (int im = 0; im < IMG; ++im) {
              const float4 pv = *reinterpret_cast<const float4*>(pooled + im * C + c);
              s[im] = fmaf(wv[u].x, pv.x, fmaf(wv[u].y, pv.y, fmaf(wv[u].z, pv.z, fmaf(wv[u].w, pv.w, s[im]))));
            }
          }
        }
      }
    } else {
#pragma unroll 8
      for (int c = lane; c < C; c += 32) {
        const float wv = __ldg(wr + c);
#pragma unroll
        for (int im = 0; im < IMG; ++im) s[im] = fmaf(wv, pooled[im * C + c], s[im]);
      }
    }
#pragma unroll
    for (int im = 0; im < IMG; ++im) {
      float v = warp_sum(s[im]);
      if (lane == 0) {
        v += b1[j];
        mine[im * jper + (j - j0)] = v * sigmoid_exact(v);
        if constexpr (kTrain) {
          if (b0 + im < B) h1_out[(size_t)(b0 + im) * sq + j] = v;
        }
      }
    }
  }
  cluster.sync();
  // gather every CTA's slice of the hidden vector through distributed shared memory
  for (int r = 0; r < kSeCluster; ++r) {
    const float* remote = cluster.map_shared_rank(mine, r);
    const int jr0 = r * jper, jn = min(sq, jr0 + jper) - jr0;
    for (int jj = tid; jj < jn; jj += blockDim.x)
#pragma unroll
      for (int im = 0; im < IMG; ++im) hidden[im * sq + jr0 + jj] = remote[im * jper + jj];
  }
  cluster.sync();   // also keeps every CTA's shared memory alive until all peers have read it

  const int cper = (C + kSeCluster - 1) / kSeCluster;
  const int c0 = rank * cper, c1 = min(C, c0 + cper);
  for (int c = c0 + tid; c < c1; c += blockDim.x) {
    float s[IMG];
    const float bv = b2[c];
#pragma unroll
    for (int im = 0; im < IMG; ++im) s[im] = bv;
    for (int jb = 0; jb < sq; jb += 32) {     // 32 weight loads in flight per thread
      float wv[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        if constexpr (kTrain) wv[u] = jb + u < sq ? __ldg(w2t + (size_t)c * sq + jb + u) : 0.f;   // torch layout [C][sq]
        else wv[u] = jb + u < sq ? __ldg(w2t + (size_t)(jb + u) * C + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        if (jb + u < sq) {
#pragma unroll
          for (int im = 0; im < IMG; ++im) s[im] = fmaf(wv[u], hidden[im * sq + jb + u], s[im]);
        }
      }
    }
#pragma unroll
    for (int im = 0; im < IMG; ++im)
      if (b0 + im < B) {
        const float gv = sigmoid_exact(s[im]);
        if constexpr (sizeof(GT) == 2) gate[(size_t)(b0 + im) * C + c] = __float2bfloat16_rn(gv);
        else gate[(size_t)(b0 + im) * C + c] = gv;
        if constexpr (kTrain) gate_f32[(size_t)(b0 + im) * C + c] = gv;
      }
  }
}

// shared launcher of the inference and training entry points
template <bool kTrain>
static int launch_se(const float* pool_partial, int parts, float inv_hw, const float* w_reduce, const float* b_reduce,
                     const float* w_expand, const float* b_expand, void* gate, int gate_dtype, int B, int C, int squeeze,
                     float* pooled, float* h1, float* gate_f32, cudaStream_t st) {
  const int img = B >= 128 ? 8 : (B >= 16 ? 4 : 1);
  const int jper = (squeeze + kSeCluster - 1) / kSeCluster;
  const size_t smem = (size_t)img * ((size_t)C + squeeze + jper) * sizeof(float);
  DFV_REQUIRE(smem <= 160 * 1024, "squeeze-excite: C + squeeze too large (%d + %d)", C, squeeze);
  const unsigned grid = (unsigned)((B + img - 1) / img) * kSeCluster;
#define SE_LAUNCH(IMG_, GT_)                                                                                                  \
  do {                                                                                                                        \
    static thread_local bool configured = false;                                                                              \
    if (!configured) {                                                                                                        \
      DFV_CUDA(cudaFuncSetAttribute(se_gate_kernel<IMG_, GT_, kTrain>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); \
      configured = true;                                                                                                      \
    }                                                                                                                         \
    DFV_PDL((se_gate_kernel<IMG_, GT_, kTrain>), grid, 256, smem, st, pool_partial, parts, inv_hw, w_reduce, b_reduce, w_expand, b_expand, \
                                                              (GT_*)gate, B, C, squeeze, pooled, h1, gate_f32);               \
  } while (0)
  if (gate_dtype == DFV_BF16) {
    if (img == 8) SE_LAUNCH(8, __nv_bfloat16); else if (img == 4) SE_LAUNCH(4, __nv_bfloat16); else SE_LAUNCH(1, __nv_bfloat16);
  } else {
    if (img == 8) SE_LAUNCH(8, float); else if (img == 4) SE_LAUNCH(4, float); else SE_LAUNCH(1, float);
  }
#undef SE_LAUNCH
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

}  // namespace dfv

using namespace dfv;

extern "C" int dfv_se_gate_fwd(const float* pool_partial, int parts, float inv_hw, const float* w_reduce,
                               const float* b_reduce, const float* w_expand_t, const float* b_expand, void* gate,
                               int gate_dtype, int B, int C, int squeeze, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(pool_partial && w_reduce && b_reduce && w_expand_t && b_expand && gate, "dfv_se_gate_fwd: null pointer");
  DFV_REQUIRE(B > 0 && C > 0 && squeeze > 0 && parts > 0 && valid_dtype(gate_dtype), "dfv_se_gate_fwd: bad shape / dtype");
  if (debug_flags() & 2) return DFV_OK;
  ProfScope prof(PK_SE_GATE, 4.0 * ((double)B * parts * C + (double)B * C + 2.0 * C * squeeze), 4.0 * B * (double)C * squeeze,
                 as_stream(stream));
  return launch_se<false>(pool_partial, parts, inv_hw, w_reduce, b_reduce, w_expand_t, b_expand, gate, gate_dtype, B, C, squeeze,
                          nullptr, nullptr, nullptr, as_stream(stream));
}

/* Training variant (declared in dfvit.h next to the other training entry points): torch-layout weights, saves pooled / h1 /
 * fp32 gate for dfv_se_bwd. */
extern "C" int dfv_se_train_fwd(const float* pool_partial, int parts, float inv_hw, const float* w_reduce, const float* b_reduce,
                                const float* w_expand, const float* b_expand, void* gate, int gate_dtype, float* pooled, float* h1,
                                float* gate_f32, int B, int C, int squeeze, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(pool_partial && w_reduce && b_reduce && w_expand && b_expand && gate && pooled && h1 && gate_f32,
              "dfv_se_train_fwd: null pointer");
  DFV_REQUIRE(B > 0 && C > 0 && squeeze > 0 && parts > 0 && valid_dtype(gate_dtype), "dfv_se_train_fwd: bad shape / dtype");
  ProfScope prof(PK_SE_GATE, 4.0 * ((double)B * parts * C + 3.0 * B * C + 2.0 * C * squeeze), 4.0 * B * (double)C * squeeze,
                 as_stream(stream));
  return launch_se<true>(pool_partial, parts, inv_hw, w_reduce, b_reduce, w_expand, b_expand, gate, gate_dtype, B, C, squeeze, pooled,
                         h1, gate_f32, as_stream(stream));
}

