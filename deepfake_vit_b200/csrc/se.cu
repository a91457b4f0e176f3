// Squeeze-excite gate: finishes the depthwise kernel's pool partial sums, then
// C -> squeeze (bias, swish) -> C (bias, sigmoid).  All fp32.
//
// The gate is two tiny GEMMs with M = B, and what it costs is dependent memory round trips, not bytes or flops
// (measured inside the forward by cutting the previous 8-CTA-cluster kernel short: launch 1.6 us, pooling 8.5 us,
// squeeze 3.8 us, cluster exchange + excite 20 us per layer on average -- every phase a chain of load batches, the
// weights cold in HBM since the previous forward).  So: three kernels, each ONE load phase (every byte the CTA needs is
// requested up front into shared memory) followed by arithmetic on shared memory:
//   squeeze: CTA = 128 channels (a K slice) x 16 images -> partial hidden sums for all squeeze rows;
//   hidden : one thread per (image, squeeze row) adds the K-slice partials in fixed order (+bias);
//   excite : CTA = 64 channels x 16 images: swish of the hidden vectors, then the FC.
// Each kernel requests its weights BEFORE waiting on its predecessor (programmatic dependent launch).
// Summation orders are fixed by (C, squeeze) alone: per-image results do not depend on batch size or scheduling.
#include "common.cuh"

namespace dfv {

// 16-byte global -> shared copy without a register round trip; src_bytes = 0 zero-fills (out-of-range chunks)
__device__ __forceinline__ void se_cp16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void se_cp_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

constexpr int kSeImgs = 16;      // images per CTA
constexpr int kSeKSlice = 128;   // channels per squeeze CTA
constexpr int kSeChans = 64;     // channels per excite CTA
constexpr int kSeThreads = 256;

// part[ks][b][j] = sum_{c in slice ks} w1[j][c] * pooled[b][c],  pooled = inv_hw * sum_t partial[b][t][c]
template <bool kTrain>
__global__ void __launch_bounds__(kSeThreads)
    se_squeeze_kernel(const float* __restrict__ partial, int parts, float inv_hw, const float* __restrict__ w1,
                      float* __restrict__ part, float* __restrict__ pooled_out, int B, int C, int sq) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(16) float sm[];
  float* ps = sm;                                  // [kSeKSlice][kSeImgs]   pooled, image-minor
  float* ws = sm + kSeKSlice * kSeImgs;            // [sq][kSeKSlice + 4]    weight slice (padded rows: conflict-free float4 reads)
  constexpr int WP = kSeKSlice + 4;
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * kSeKSlice, nc = min(kSeKSlice, C - c0);
  const int b0 = blockIdx.y * kSeImgs, nimg = min(kSeImgs, B - b0);
  const bool vec = (C & 3) == 0 && ((reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(partial)) & 15) == 0;
  // ---- load phase: everything requested before anything is used.  The weights do not depend on the producer kernel:
  // their (HBM-cold) fetch is issued before this grid waits for it.
  if (vec) {
    const int nc4 = nc >> 2;
    for (int i = tid; i < sq * (kSeKSlice / 4); i += kSeThreads) {          // weight slice: asynchronous copies
      const int jj = i / (kSeKSlice / 4), q = i % (kSeKSlice / 4);
      se_cp16(ws + (size_t)jj * WP + q * 4, w1 + (size_t)jj * C + c0 + (q < nc4 ? q * 4 : 0), q < nc4 ? 16 : 0);
    }
  } else {
    for (int i = tid; i < sq * kSeKSlice; i += kSeThreads) {
      const int jj = i / kSeKSlice, k = i % kSeKSlice;
      ws[(size_t)jj * WP + k] = k < nc ? __ldg(w1 + (size_t)jj * C + c0 + k) : 0.f;
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (vec) {
    const int nc4 = nc >> 2;
    // pooled slice of 16 images: 512 float4 items, two per thread, their (up to 2 x 4) loads issued together
    for (int t0 = 0; t0 < parts; t0 += 4) {
      float4 u[2][4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = tid + e * kSeThreads, im_ = i / (kSeKSlice / 4), q = i % (kSeKSlice / 4);
        const bool ok = im_ < nimg && q < nc4;
        const float4* src = reinterpret_cast<const float4*>(partial + (size_t)(b0 + (ok ? im_ : 0)) * parts * C + c0) + (ok ? q : 0);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          u[e][t] = ok && t0 + t < parts ? __ldg(src + (size_t)(t0 + t) * (C >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = tid + e * kSeThreads, im_ = i / (kSeKSlice / 4), q = i % (kSeKSlice / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 4; ++t) { v.x += u[e][t].x; v.y += u[e][t].y; v.z += u[e][t].z; v.w += u[e][t].w; }
        float* d0 = ps + (q * 4) * kSeImgs + im_;
        if (t0 == 0) { d0[0] = v.x; d0[kSeImgs] = v.y; d0[2 * kSeImgs] = v.z; d0[3 * kSeImgs] = v.w; }
        else { d0[0] += v.x; d0[kSeImgs] += v.y; d0[2 * kSeImgs] += v.z; d0[3 * kSeImgs] += v.w; }
      }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {      // scale (each thread re-reads only what it wrote)
      const int i = tid + e * kSeThreads, im_ = i / (kSeKSlice / 4), q = i % (kSeKSlice / 4);
      float* d0 = ps + (q * 4) * kSeImgs + im_;
      float4 v = make_float4(d0[0] * inv_hw, d0[kSeImgs] * inv_hw, d0[2 * kSeImgs] * inv_hw, d0[3 * kSeImgs] * inv_hw);
      d0[0] = v.x; d0[kSeImgs] = v.y; d0[2 * kSeImgs] = v.z; d0[3 * kSeImgs] = v.w;
      if constexpr (kTrain) {
        if (im_ < nimg && q < nc4) *reinterpret_cast<float4*>(pooled_out + (size_t)(b0 + im_) * C + c0 + q * 4) = v;
      }
    }
    se_cp_wait_all();
  } else {
    for (int i = tid; i < kSeImgs * kSeKSlice; i += kSeThreads) {
      const int im = i / kSeKSlice, k = i % kSeKSlice;
      float v = 0.f;
      if (im < nimg && k < nc) {
        for (int t = 0; t < parts; ++t) v += __ldg(partial + ((size_t)(b0 + im) * parts + t) * C + c0 + k);
        v *= inv_hw;
        if constexpr (kTrain) pooled_out[(size_t)(b0 + im) * C + c0 + k] = v;
      }
      ps[k * kSeImgs + im] = v;
    }
  }
  __syncthreads();
  // ---- thread = (image, row lane): rows rl, rl + 16, ...  (<= 8 rows per pass)
  const int im = tid & (kSeImgs - 1), rl = tid >> 4;
  for (int jb = rl; jb < sq; jb += 16 * 8) {
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    for (int k = 0; k < kSeKSlice; k += 4) {
      const float p0 = ps[(k + 0) * kSeImgs + im], p1 = ps[(k + 1) * kSeImgs + im], p2 = ps[(k + 2) * kSeImgs + im],
                  p3 = ps[(k + 3) * kSeImgs + im];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int jj = jb + r * 16;
        if (jj < sq) {
          const float4 w = *reinterpret_cast<const float4*>(ws + (size_t)jj * WP + k);
          acc[r] = fmaf(w.x, p0, fmaf(w.y, p1, fmaf(w.z, p2, fmaf(w.w, p3, acc[r]))));
        }
      }
    }
    if (im < nimg) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int jj = jb + r * 16;
        if (jj < sq) part[((size_t)blockIdx.x * B + b0 + im) * sq + jj] = acc[r];
      }
    }
  }
}

// h1[b][j] = b1[j] + sum_ks part[ks][b][j]  (fixed order); one thread per element, all its loads in one batch
__global__ void __launch_bounds__(kSeThreads)
    se_hidden_kernel(const float* __restrict__ part, int ksplit, const float* __restrict__ b1, float* __restrict__ h1, int B, int sq) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int i = blockIdx.x * kSeThreads + threadIdx.x;
  const size_t n = (size_t)B * sq;
  const float bj = (size_t)i < n ? b1[i % sq] : 0.f;       // a weight: loaded before waiting on the squeeze kernel
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if ((size_t)i >= n) return;
  float v = bj;
  for (int ks0 = 0; ks0 < ksplit; ks0 += 8) {
    float t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = ks0 + u < ksplit ? part[(size_t)(ks0 + u) * n + i] : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) v += t[u];
  }
  h1[i] = v;
}

// gate[b][c] = sigmoid(b2[c] + sum_j w2[c][j] * swish(h1[b][j]))
// kTrain: the expand weight arrives in torch layout [C][sq] (not transposed) and the fp32 gate is saved as well.
template <typename GT, bool kTrain>
__global__ void __launch_bounds__(kSeThreads)
    se_excite_kernel(const float* __restrict__ h1, const float* __restrict__ w2, const float* __restrict__ b2,
                     GT* __restrict__ gate, float* __restrict__ gate_f32, int B, int C, int sq) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(16) float sm[];
  float* hs = sm;                               // [sq][kSeImgs]       swish(h1), image-minor
  float* ws = sm + (size_t)sq * kSeImgs;        // [sq][kSeChans + 4]  weight slice, channel-minor
  constexpr int WP = kSeChans + 4;
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * kSeChans, nc = min(kSeChans, C - c0);
  const int b0 = blockIdx.y * kSeImgs, nimg = min(kSeImgs, B - b0);
  const int cc = tid & (kSeChans - 1), ig = tid >> 6;             // compute role: 64 channels x 4 image groups
  // ---- load phase; weights and bias first (independent of the producer kernels), then wait, then the hidden vectors
  const bool vec = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0;
  if constexpr (kTrain) {
    const int total = nc * sq;                                   // contiguous [nc][sq] block of the torch-layout weight
    for (int base = tid; base < total; base += kSeThreads * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int i = base + u * kSeThreads; v[u] = i < total ? __ldg(w2 + (size_t)c0 * sq + i) : 0.f; }
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int i = base + u * kSeThreads; if (i < total) ws[(size_t)(i % sq) * WP + i / sq] = v[u]; }
    }
  } else if (vec) {
    const int nc4 = nc >> 2;
    for (int i = tid; i < sq * (kSeChans / 4); i += kSeThreads) {
      const int jj = i / (kSeChans / 4), q = i % (kSeChans / 4);
      se_cp16(ws + (size_t)jj * WP + q * 4, w2 + (size_t)jj * C + c0 + (q < nc4 ? q * 4 : 0), q < nc4 ? 16 : 0);
    }
  } else {
    for (int i = tid; i < sq * kSeChans; i += kSeThreads) {
      const int jj = i / kSeChans, c_ = i % kSeChans;
      ws[(size_t)jj * WP + c_] = c_ < nc ? __ldg(w2 + (size_t)jj * C + c0 + c_) : 0.f;
    }
  }
  const float bv = cc < nc ? b2[c0 + cc] : 0.f;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int base = tid; base < sq * kSeImgs; base += kSeThreads * 8) {      // <= 7 independent loads per thread, one batch
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * kSeThreads, im = i / sq, jj = i % sq;        // consecutive threads -> consecutive j (coalesced)
      v[u] = i < sq * kSeImgs && im < nimg ? h1[(size_t)(b0 + im) * sq + jj] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * kSeThreads, im = i / sq, jj = i % sq;
      if (i < sq * kSeImgs) hs[jj * kSeImgs + im] = v[u] * sigmoid_exact(v[u]);
    }
  }
  se_cp_wait_all();
  __syncthreads();
  // ---- thread = (channel, group of four images)
  if (cc >= nc) return;
  float a0 = bv, a1 = bv, a2 = bv, a3 = bv;
  for (int jj = 0; jj < sq; ++jj) {
    const float w = ws[(size_t)jj * WP + cc];
    const float4 h = *reinterpret_cast<const float4*>(hs + jj * kSeImgs + ig * 4);
    a0 = fmaf(w, h.x, a0); a1 = fmaf(w, h.y, a1); a2 = fmaf(w, h.z, a2); a3 = fmaf(w, h.w, a3);
  }
  const float av[4] = {a0, a1, a2, a3};
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int im = ig * 4 + u;
    if (im < nimg) {
      const float gv = sigmoid_exact(av[u]);
      const size_t o = (size_t)(b0 + im) * C + c0 + cc;
      if constexpr (sizeof(GT) == 2) gate[o] = __float2bfloat16_rn(gv);
      else gate[o] = gv;
      if constexpr (kTrain) gate_f32[o] = gv;
    }
  }
}

// [ksplit][B][sq] partial sums + [B][sq] hidden pre-activations
static size_t se_scratch_floats(int B, int C, int squeeze) { return (size_t)((C + kSeKSlice - 1) / kSeKSlice + 1) * B * squeeze; }

// shared launcher of the inference and training entry points
template <bool kTrain>
static int launch_se(const float* pool_partial, int parts, float inv_hw, const float* w_reduce, const float* b_reduce,
                     const float* w_expand, const float* b_expand, void* gate, int gate_dtype, int B, int C, int squeeze,
                     float* pooled, float* h1, float* gate_f32, float* scratch, cudaStream_t st) {
  const int ksplit = (C + kSeKSlice - 1) / kSeKSlice;
  const size_t smem_a = ((size_t)kSeKSlice * kSeImgs + (size_t)squeeze * (kSeKSlice + 4)) * sizeof(float);
  const size_t smem_b = ((size_t)squeeze * kSeImgs + (size_t)squeeze * (kSeChans + 4)) * sizeof(float);
  DFV_REQUIRE(smem_a <= 160 * 1024, "squeeze-excite: squeeze width %d too large", squeeze);
  const dim3 grid_a((unsigned)ksplit, (unsigned)((B + kSeImgs - 1) / kSeImgs));
  const dim3 grid_b((unsigned)((C + kSeChans - 1) / kSeChans), (unsigned)((B + kSeImgs - 1) / kSeImgs));
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(se_squeeze_kernel<kTrain>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DFV_CUDA(cudaFuncSetAttribute(se_excite_kernel<__nv_bfloat16, kTrain>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    DFV_CUDA(cudaFuncSetAttribute(se_excite_kernel<float, kTrain>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  float* hid = h1 ? h1 : scratch + (size_t)ksplit * B * squeeze;      // training keeps h1 for the backward pass
  DFV_PDL((se_squeeze_kernel<kTrain>), grid_a, kSeThreads, smem_a, st, pool_partial, parts, inv_hw, w_reduce, scratch, pooled, B, C, squeeze);
  DFV_PDL(se_hidden_kernel, (unsigned)(((size_t)B * squeeze + kSeThreads - 1) / kSeThreads), kSeThreads, 0, st, (const float*)scratch, ksplit,
          b_reduce, hid, B, squeeze);
  if (gate_dtype == DFV_BF16)
    DFV_PDL((se_excite_kernel<__nv_bfloat16, kTrain>), grid_b, kSeThreads, smem_b, st, (const float*)hid, w_expand, b_expand,
            (__nv_bfloat16*)gate, gate_f32, B, C, squeeze);
  else
    DFV_PDL((se_excite_kernel<float, kTrain>), grid_b, kSeThreads, smem_b, st, (const float*)hid, w_expand, b_expand, (float*)gate,
            gate_f32, B, C, squeeze);
  count_launch(2);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

}  // namespace dfv

using namespace dfv;

extern "C" size_t dfv_se_scratch_floats(int B, int C, int squeeze) {
  return B > 0 && C > 0 && squeeze > 0 ? se_scratch_floats(B, C, squeeze) : 0;
}

extern "C" int dfv_se_gate_fwd(const float* pool_partial, int parts, float inv_hw, const float* w_reduce,
                               const float* b_reduce, const float* w_expand_t, const float* b_expand, void* gate,
                               int gate_dtype, float* scratch, int B, int C, int squeeze, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(pool_partial && w_reduce && b_reduce && w_expand_t && b_expand && gate && scratch, "dfv_se_gate_fwd: null pointer");
  DFV_REQUIRE(B > 0 && C > 0 && squeeze > 0 && parts > 0 && valid_dtype(gate_dtype), "dfv_se_gate_fwd: bad shape / dtype");
  if (debug_flags() & 2) return DFV_OK;
  ProfScope prof(PK_SE_GATE, 4.0 * ((double)B * parts * C + (double)B * C + 2.0 * C * squeeze), 4.0 * B * (double)C * squeeze,
                 as_stream(stream));
  return launch_se<false>(pool_partial, parts, inv_hw, w_reduce, b_reduce, w_expand_t, b_expand, gate, gate_dtype, B, C, squeeze,
                          nullptr, nullptr, nullptr, scratch, as_stream(stream));
}

/* Training variant (declared in dfvit.h next to the other training entry points): torch-layout weights, saves pooled / h1 /
 * fp32 gate for dfv_se_bwd. */
extern "C" int dfv_se_train_fwd(const float* pool_partial, int parts, float inv_hw, const float* w_reduce, const float* b_reduce,
                                const float* w_expand, const float* b_expand, void* gate, int gate_dtype, float* pooled, float* h1,
                                float* gate_f32, float* scratch, int B, int C, int squeeze, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(pool_partial && w_reduce && b_reduce && w_expand && b_expand && gate && pooled && h1 && gate_f32 && scratch,
              "dfv_se_train_fwd: null pointer");
  DFV_REQUIRE(B > 0 && C > 0 && squeeze > 0 && parts > 0 && valid_dtype(gate_dtype), "dfv_se_train_fwd: bad shape / dtype");
  ProfScope prof(PK_SE_GATE, 4.0 * ((double)B * parts * C + 3.0 * B * C + 2.0 * C * squeeze), 4.0 * B * (double)C * squeeze,
                 as_stream(stream));
  return launch_se<true>(pool_partial, parts, inv_hw, w_reduce, b_reduce, w_expand, b_expand, gate, gate_dtype, B, C, squeeze, pooled,
                         h1, gate_f32, scratch, as_stream(stream));
}

