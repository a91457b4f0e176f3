// Small fp32 fully-connected layers over the whole batch (M = B rows): the squeeze-excite gates, the channel-attention
// MLP and the classifier head.  What these cost is dependent memory round trips, not bytes or flops (their weights are
// HBM-cold -- last touched one forward pass ago -- and a chain of load batches pays that latency per link), so every
// kernel here has ONE load phase: all the bytes a CTA needs are requested up front into shared memory, the weights even
// before the grid waits on its producer (programmatic dependent launch), then arithmetic runs on shared memory.
//   sl_rowmajor_partial_kernel : weights [N][K] (torch layout); CTA = 128-wide K slice x 16 rows -> partial sums
//   sl_combine_kernel          : one thread per output adds the K-slice partials in fixed order (+bias, relu)
//   sl_kmajor_kernel           : weights [K][N] (transposed); CTA = 64 outputs x 16 rows (x K slice)
// Summation orders are fixed by the layer shape alone: results do not depend on batch size or scheduling.
#pragma once
#include "common.cuh"

namespace dfv {

// 16-byte global -> shared copy without a register round trip; src_bytes = 0 zero-fills (out-of-range chunks)
__device__ __forceinline__ void sl_cp16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void sl_cp_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

constexpr int kSlRows = 16;      // images per CTA
constexpr int kSlKSlice = 128;   // channels per squeeze CTA
constexpr int kSlCols = 64;     // channels per excite CTA
constexpr int kSlThreads = 256;

// part[ks][b][j] = sum_{c in slice ks} w1[j][c] * pooled[b][c],  pooled = inv_hw * sum_t partial[b][t][c]
template <bool kTrain>
__global__ void __launch_bounds__(kSlThreads)
    sl_rowmajor_partial_kernel(const float* __restrict__ partial, int parts, float inv_hw, const float* __restrict__ w1,
                      float* __restrict__ part, float* __restrict__ pooled_out, int B, int C, int sq, size_t x_zstride,
                      size_t part_zstride) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  partial += blockIdx.z * x_zstride;      // gridDim.z independent input sets against the same weights
  part += blockIdx.z * part_zstride;
  extern __shared__ __align__(16) float sm[];
  float* ps = sm;                                  // [kSlKSlice][kSlRows]   pooled, image-minor
  float* ws = sm + kSlKSlice * kSlRows;            // [sq][kSlKSlice + 4]    weight slice (padded rows: conflict-free float4 reads)
  constexpr int WP = kSlKSlice + 4;
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * kSlKSlice, nc = min(kSlKSlice, C - c0);
  const int b0 = blockIdx.y * kSlRows, nimg = min(kSlRows, B - b0);
  const bool vec = (C & 3) == 0 && ((reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(partial)) & 15) == 0;
  // ---- load phase: everything requested before anything is used.  The weights do not depend on the producer kernel:
  // their (HBM-cold) fetch is issued before this grid waits for it.
  if (vec) {
    const int nc4 = nc >> 2;
    for (int i = tid; i < sq * (kSlKSlice / 4); i += kSlThreads) {          // weight slice: asynchronous copies
      const int jj = i / (kSlKSlice / 4), q = i % (kSlKSlice / 4);
      sl_cp16(ws + (size_t)jj * WP + q * 4, w1 + (size_t)jj * C + c0 + (q < nc4 ? q * 4 : 0), q < nc4 ? 16 : 0);
    }
  } else {
    for (int i = tid; i < sq * kSlKSlice; i += kSlThreads) {
      const int jj = i / kSlKSlice, k = i % kSlKSlice;
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
        const int i = tid + e * kSlThreads, im_ = i / (kSlKSlice / 4), q = i % (kSlKSlice / 4);
        const bool ok = im_ < nimg && q < nc4;
        const float4* src = reinterpret_cast<const float4*>(partial + (size_t)(b0 + (ok ? im_ : 0)) * parts * C + c0) + (ok ? q : 0);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          u[e][t] = ok && t0 + t < parts ? __ldg(src + (size_t)(t0 + t) * (C >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = tid + e * kSlThreads, im_ = i / (kSlKSlice / 4), q = i % (kSlKSlice / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 4; ++t) { v.x += u[e][t].x; v.y += u[e][t].y; v.z += u[e][t].z; v.w += u[e][t].w; }
        float* d0 = ps + (q * 4) * kSlRows + im_;
        if (t0 == 0) { d0[0] = v.x; d0[kSlRows] = v.y; d0[2 * kSlRows] = v.z; d0[3 * kSlRows] = v.w; }
        else { d0[0] += v.x; d0[kSlRows] += v.y; d0[2 * kSlRows] += v.z; d0[3 * kSlRows] += v.w; }
      }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {      // scale (each thread re-reads only what it wrote)
      const int i = tid + e * kSlThreads, im_ = i / (kSlKSlice / 4), q = i % (kSlKSlice / 4);
      float* d0 = ps + (q * 4) * kSlRows + im_;
      float4 v = make_float4(d0[0] * inv_hw, d0[kSlRows] * inv_hw, d0[2 * kSlRows] * inv_hw, d0[3 * kSlRows] * inv_hw);
      d0[0] = v.x; d0[kSlRows] = v.y; d0[2 * kSlRows] = v.z; d0[3 * kSlRows] = v.w;
      if constexpr (kTrain) {
        if (im_ < nimg && q < nc4) *reinterpret_cast<float4*>(pooled_out + (size_t)(b0 + im_) * C + c0 + q * 4) = v;
      }
    }
    sl_cp_wait_all();
  } else {
    for (int i = tid; i < kSlRows * kSlKSlice; i += kSlThreads) {
      const int im = i / kSlKSlice, k = i % kSlKSlice;
      float v = 0.f;
      if (im < nimg && k < nc) {
        for (int t = 0; t < parts; ++t) v += __ldg(partial + ((size_t)(b0 + im) * parts + t) * C + c0 + k);
        v *= inv_hw;
        if constexpr (kTrain) pooled_out[(size_t)(b0 + im) * C + c0 + k] = v;
      }
      ps[k * kSlRows + im] = v;
    }
  }
  __syncthreads();
  // ---- thread = (image, row lane): rows rl, rl + 16, ...  (<= 8 rows per pass)
  const int im = tid & (kSlRows - 1), rl = tid >> 4;
  for (int jb = rl; jb < sq; jb += 16 * 8) {
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    for (int k = 0; k < kSlKSlice; k += 4) {
      const float p0 = ps[(k + 0) * kSlRows + im], p1 = ps[(k + 1) * kSlRows + im], p2 = ps[(k + 2) * kSlRows + im],
                  p3 = ps[(k + 3) * kSlRows + im];
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

// out[b][j] = f(bias[j] + sum_ks part[ks][b][j]) (+ f(sum_ks part2[ks][b][j])),  f = relu or identity.  Fixed order;
// one thread per element, its loads in batches of eight.  bias and part2 may be NULL.
static __global__ void __launch_bounds__(kSlThreads)
    sl_combine_kernel(const float* __restrict__ part, const float* __restrict__ part2, int ksplit, const float* __restrict__ bias,
                      float* __restrict__ out, int B, int N, int relu) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int i = blockIdx.x * kSlThreads + threadIdx.x;
  const size_t n = (size_t)B * N;
  const float bj = bias && (size_t)i < n ? bias[i % N] : 0.f;       // a weight: loaded before waiting on the producer kernel
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
  if (relu) v = fmaxf(v, 0.f);
  if (part2) {
    float v2 = 0.f;
    for (int ks0 = 0; ks0 < ksplit; ks0 += 8) {
      float t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = ks0 + u < ksplit ? part2[(size_t)(ks0 + u) * n + i] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) v2 += t[u];
    }
    v += relu ? fmaxf(v2, 0.f) : v2;
  }
  out[i] = v;
}

constexpr int SL_IN_SWISH = 1, SL_OUT_SIGMOID = 1, SL_OUT_RELU = 2;

// out[b][c] = oact(bias[c] + sum_k w[k][c] * iact(h[b][k])),  iact = swish or identity, oact = sigmoid, relu or identity.
// CTA = 64 outputs x 16 rows x one K slice of kc (blockIdx.z); with several K slices the CTA writes its raw partial sum
// to part_out[z][b][c] instead (sl_combine_kernel finishes).  kTrain: w arrives in torch layout [C][K] (not
// transposed) and the fp32 output is saved as well (squeeze-excite training).
template <typename GT, bool kTrain>
__global__ void __launch_bounds__(kSlThreads)
    sl_kmajor_kernel(const float* __restrict__ h1, const float* __restrict__ w2, const float* __restrict__ b2,
                     GT* __restrict__ gate, float* __restrict__ gate_f32, float* __restrict__ part_out, int B, int C, int sq, int kc,
                     int in_act, int out_act, const long long* __restrict__ h_fix = nullptr, const float* __restrict__ h_bias = nullptr,
                     float h_fix_scale = 1.0f) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(16) float sm[];
  float* hs = sm;                               // [kc][kSlRows]       iact(h), row-minor
  float* ws = sm + (size_t)kc * kSlRows;        // [kc][kSlCols + 4]   weight slice, output-minor
  constexpr int WP = kSlCols + 4;
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * kSlCols, nc = min(kSlCols, C - c0);
  const int b0 = blockIdx.y * kSlRows, nimg = min(kSlRows, B - b0);
  const int k0 = blockIdx.z * kc, nk = min(kc, sq - k0);
  const int cc = tid & (kSlCols - 1), ig = tid >> 6;             // compute role: 64 outputs x 4 groups of four rows
  // ---- load phase; weights and bias first (independent of the producer kernels), then wait, then the input vectors
  const bool vec = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0;
  if constexpr (kTrain) {
    const int total = nc * nk;
    for (int base = tid; base < total; base += kSlThreads * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * kSlThreads;
        v[u] = i < total ? __ldg(w2 + (size_t)(c0 + i / nk) * sq + k0 + i % nk) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int i = base + u * kSlThreads; if (i < total) ws[(size_t)(i % nk) * WP + i / nk] = v[u]; }
    }
  } else if (vec) {
    const int nc4 = nc >> 2;
    for (int i = tid; i < nk * (kSlCols / 4); i += kSlThreads) {
      const int jj = i / (kSlCols / 4), q = i % (kSlCols / 4);
      sl_cp16(ws + (size_t)jj * WP + q * 4, w2 + (size_t)(k0 + jj) * C + c0 + (q < nc4 ? q * 4 : 0), q < nc4 ? 16 : 0);
    }
  } else {
    for (int i = tid; i < nk * kSlCols; i += kSlThreads) {
      const int jj = i / kSlCols, c_ = i % kSlCols;
      ws[(size_t)jj * WP + c_] = c_ < nc ? __ldg(w2 + (size_t)(k0 + jj) * C + c0 + c_) : 0.f;
    }
  }
  const float bv = b2 && cc < nc && gridDim.z == 1 ? b2[c0 + cc] : 0.f;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int base = tid; base < nk * kSlRows; base += kSlThreads * 8) {      // independent loads, batches of eight per thread
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * kSlThreads, im = i / nk, jj = i % nk;        // consecutive threads -> consecutive k (coalesced)
      if (h_fix != nullptr)      // fixed-point accumulators of the fused squeeze (dwconv.cu SeFuse) + the squeeze bias
        v[u] = i < nk * kSlRows && im < nimg ? (float)((double)h_fix[(size_t)(b0 + im) * sq + k0 + jj] * (double)h_fix_scale) + h_bias[k0 + jj] : 0.f;
      else
        v[u] = i < nk * kSlRows && im < nimg ? h1[(size_t)(b0 + im) * sq + k0 + jj] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * kSlThreads, im = i / nk, jj = i % nk;
      if (i < nk * kSlRows) hs[jj * kSlRows + im] = in_act == SL_IN_SWISH ? v[u] * sigmoid_exact(v[u]) : v[u];
    }
  }
  sl_cp_wait_all();
  __syncthreads();
  // ---- thread = (output, group of four rows)
  if (cc >= nc) return;
  float a0 = bv, a1 = bv, a2 = bv, a3 = bv;
  for (int jj = 0; jj < nk; ++jj) {
    const float w = ws[(size_t)jj * WP + cc];
    const float4 h = *reinterpret_cast<const float4*>(hs + jj * kSlRows + ig * 4);
    a0 = fmaf(w, h.x, a0); a1 = fmaf(w, h.y, a1); a2 = fmaf(w, h.z, a2); a3 = fmaf(w, h.w, a3);
  }
  const float av[4] = {a0, a1, a2, a3};
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int im = ig * 4 + u;
    if (im < nimg) {
      const size_t o = (size_t)(b0 + im) * C + c0 + cc;
      if (gridDim.z > 1) {
        part_out[(size_t)blockIdx.z * B * C + o] = av[u];
      } else {
        const float gv = out_act == SL_OUT_SIGMOID ? sigmoid_exact(av[u]) : (out_act == SL_OUT_RELU ? fmaxf(av[u], 0.f) : av[u]);
        if constexpr (sizeof(GT) == 2) gate[o] = __float2bfloat16_rn(gv);
        else gate[o] = gv;
        if constexpr (kTrain) gate_f32[o] = gv;
      }
    }
  }
}

// host helpers shared by the launchers
inline size_t sl_rowmajor_smem(int N) { return ((size_t)kSlKSlice * kSlRows + (size_t)N * (kSlKSlice + 4)) * sizeof(float); }
inline size_t sl_kmajor_smem(int kc) { return ((size_t)kc * kSlRows + (size_t)kc * (kSlCols + 4)) * sizeof(float); }
inline int sl_ksplit_rowmajor(int K) { return (K + kSlKSlice - 1) / kSlKSlice; }

}  // namespace dfv
