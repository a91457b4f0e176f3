// 1x1-convolution weight gradient on the tensor cores (bf16 tensors, fp32 accumulate):
//     dW[n][k] += sum_m g[m][n] * (a[m][k] * gate[image(m)][k])
// The reduction runs over the ROW index m of two channels-last tensors, so both UMMA operands are
// "MN-major": a TMA box of 64 rows x 64 channels (128-byte swizzle) is, as it lands, the canonical
// MN-major SWIZZLE_128B atom stack (64 channels contiguous, 8-row groups 1024 B apart), and 16 rows
// are one tcgen05.mma K step.  No transposes, no staging copies.
//   P operand (UMMA M = 128, TMEM lanes)   : the narrower of g / a, 128 channels per tile (two boxes)
//   Q operand (UMMA N = QT <= 256, columns): the wider one, QT channels per tile (QT / 64 boxes)
//   grid = (P tiles x Q tiles) x row splits; every CTA streams its row range once through a
//   TMA -> (SE-gate transform) -> tcgen05.mma pipeline and finishes with fp32 atomic adds into dW.
//   warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..9 gate transform, then epilogue.
// HBM-bound by construction: one stage (64 rows) carries 24-48 KB and costs 4 MMAs (130-520 clk).
#include <algorithm>

#include "common.cuh"

namespace dfv {

constexpr int kWgRows = 64;          // rows (the MMA reduction dimension) per pipeline stage
constexpr int kWgBox = kWgRows * 128;  // bytes of one 64-row x 64-channel box
constexpr int kWgMaxStages = 8;
constexpr int kWgThreads = 320;      // 10 warps
constexpr int kWgWorkers = 256;      // warps 2..9

struct WgParams {
  long long M;
  int Pdim, Qdim;        // channel counts of the two operands
  int p_tiles, q_tiles;
  int QT, qboxes;
  int stages;
  long long rows_per_split;
  int rows_per_image;
  int p_is_g;            // 1: P = g (lanes index n), Q = a (columns index k); 0: the other way round
  int K;                 // row stride of dW
  int a_boxes;           // boxes of the a operand per stage (2 or qboxes)
  uint32_t tmem_cols;
};

struct __align__(8) WgBarriers {
  uint64_t full[kWgMaxStages];
  uint64_t ready[kWgMaxStages];
  uint64_t empty[kWgMaxStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

// MN-major SWIZZLE_128B operand: 64-channel chunks kWgBox bytes apart (LBO), 8-row groups 1024 B apart (SBO).
__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(kWgBox >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <bool kHasScale>
__global__ void __launch_bounds__(kWgThreads, 1)
    pw_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_q,
                       const __nv_bfloat16* __restrict__ a_scale, float* __restrict__ dw, WgParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t stage_bytes = (uint32_t)(2 + p.qboxes) * kWgBox;
  WgBarriers* bars = reinterpret_cast<WgBarriers*>(smem + (size_t)p.stages * stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x % (p.p_tiles * p.q_tiles);
  const int split = blockIdx.x / (p.p_tiles * p.q_tiles);
  const int pt = tile % p.p_tiles, qt = tile / p.p_tiles;
  const long long m_begin = (long long)split * p.rows_per_split;
  const long long m_end = min(m_begin + p.rows_per_split, p.M);
  const int n_iters = (int)((m_end - m_begin + kWgRows - 1) / kWgRows);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->ready[s], kWgWorkers);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->tmem_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_p);
    tma_prefetch_desc(&tm_q);
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(&bars->empty[stage], phase ^ 1, 11);
        unsigned char* sp = smem + (size_t)stage * stage_bytes;
        const int m0 = (int)(m_begin + (long long)it * kWgRows);
        mbar_expect_tx(&bars->full[stage], stage_bytes);
        for (int j = 0; j < 2; ++j) tma_load_2d(sp + j * kWgBox, &tm_p, &bars->full[stage], pt * 128 + j * 64, m0);
        for (int j = 0; j < p.qboxes; ++j)
          tma_load_2d(sp + (2 + j) * kWgBox, &tm_q, &bars->full[stage], qt * p.QT + j * 64, m0);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      // D = f32, A = B = bf16, both operands MN-major, N = QT, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.QT >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(kHasScale ? &bars->ready[stage] : &bars->full[stage], phase, 12);
        tc_fence_after();
        const uint32_t sp = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t sq = sp + 2 * kWgBox;
#pragma unroll
        for (int ks = 0; ks < kWgRows / 16; ++ks)
          umma_bf16(tmem_base, make_mn_sw128_desc(sp + ks * 2048), make_mn_sw128_desc(sq + ks * 2048), idesc, (it | ks) != 0);
        umma_commit(&bars->empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&bars->tmem_full);
    }
  } else {
    const int tx = threadIdx.x - 64;   // 0..255
    if constexpr (kHasScale) {
      // ----------------------------------------------------------- SE-gate transform of the a operand
      // thread = (row, chunk pair): two 16-byte chunks of the row in every box of a.
      const int row = tx & 63, cs = tx >> 6;
      const int a_box0 = p.p_is_g ? 2 : 0;
      const int ch_tile = p.p_is_g ? qt * p.QT : pt * 128;
      const int C = p.p_is_g ? p.Qdim : p.Pdim;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_iters; ++it) {
        const long long m = m_begin + (long long)it * kWgRows + row;
        const bool valid = m < m_end;
        const __nv_bfloat16* srow = a_scale + (size_t)(valid ? m / p.rows_per_image : 0) * C;
        mbar_wait(&bars->full[stage], phase, 13);
        if (valid) {
          unsigned char* base = smem + (size_t)stage * stage_bytes + (size_t)a_box0 * kWgBox + (size_t)row * 128;
          for (int j0 = 0; j0 < p.a_boxes; j0 += 2) {
            uint4 u[4], gt[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = j0 + (i >> 1), c = cs * 2 + (i & 1);
              const int ch = ch_tile + j * 64 + c * 8;
              if (j < p.a_boxes && ch < C) {
                u[i] = *reinterpret_cast<const uint4*>(base + (size_t)j * kWgBox + ((c ^ (row & 7)) << 4));
                gt[i] = __ldg(reinterpret_cast<const uint4*>(srow + ch));
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = j0 + (i >> 1), c = cs * 2 + (i & 1);
              const int ch = ch_tile + j * 64 + c * 8;
              if (j < p.a_boxes && ch < C) {
                uint4 v;
                v.x = hmul2_bf16(u[i].x, gt[i].x);
                v.y = hmul2_bf16(u[i].y, gt[i].y);
                v.z = hmul2_bf16(u[i].z, gt[i].z);
                v.w = hmul2_bf16(u[i].w, gt[i].w);
                *reinterpret_cast<uint4*>(base + (size_t)j * kWgBox + ((c ^ (row & 7)) << 4)) = v;
              }
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&bars->ready[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    // ------------------------------------------------------------- epilogue: TMEM -> fp32 atomic adds
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // column half
    const int pidx = pt * 128 + q * 32 + lane;
    const int cols_half = p.QT >> 1;
    mbar_wait(&bars->tmem_full, 0, 14);
    tc_fence_after();
    const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * cols_half);
    for (int c0 = 0; c0 < cols_half; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tbase + (uint32_t)c0, v);
      tmem_ld_wait();
      if (pidx < p.Pdim) {
        const int q0 = qt * p.QT + half * cols_half + c0;
        if (p.p_is_g) {
          // a lane owns 16 consecutive k of one weight row: four 16-byte vector reductions (K % 8 == 0 keeps them
          // aligned) instead of 16 scalar ones -- the split-M epilogue is bound by L2 atomic operations on the
          // late layers (7 splits x 444k weights)
          float* dst = dw + (size_t)pidx * p.K + q0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (q0 + j < p.Qdim)      // Qdim % 8 == 0: a group of 4 is either entirely valid or entirely out of range
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                           : "memory");
          }
        } else {
          // lanes = consecutive k of the same weight row: every scalar reduction of the warp is one coalesced line
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (q0 + j < p.Qdim) atomicAdd(dw + (size_t)(q0 + j) * p.K + pidx, __uint_as_float(v[j]));
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// Host launcher (bf16 tensors; K % 8 == 0 and N % 8 == 0).
int launch_wgrad_tc(const void* g, const void* a, const void* a_scale, int rows_per_image, float* dw, long long M, int K,
                    int N, cudaStream_t st) {
  auto cost = [](int Pd, int Qd) {   // channel columns streamed per row, re-reads included
    const int pt = (Pd + 127) / 128, qt = (Qd + 255) / 256;
    return (long long)qt * Pd + (long long)pt * Qd;
  };
  WgParams p;
  p.M = M;
  p.K = K;
  p.p_is_g = cost(N, K) < cost(K, N) ? 1 : 0;   // tie: lanes over k (coalesced reductions)
  p.Pdim = p.p_is_g ? N : K;
  p.Qdim = p.p_is_g ? K : N;
  p.p_tiles = (p.Pdim + 127) / 128;
  p.q_tiles = (p.Qdim + 255) / 256;
  p.QT = (((p.Qdim + p.q_tiles - 1) / p.q_tiles) + 63) / 64 * 64;
  p.qboxes = p.QT / 64;
  p.a_boxes = p.p_is_g ? p.qboxes : 2;
  p.rows_per_image = rows_per_image > 0 ? rows_per_image : 1;
  p.tmem_cols = p.QT <= 64 ? 64 : (p.QT <= 128 ? 128 : 256);
  const int tiles = p.p_tiles * p.q_tiles;
  long long splits = std::max(1, num_sms() / tiles);
  long long rps = (M + splits - 1) / splits;
  rps = (rps + kWgRows - 1) / kWgRows * kWgRows;
  splits = (M + rps - 1) / rps;
  p.rows_per_split = rps;
  const size_t stage_bytes = (size_t)(2 + p.qboxes) * kWgBox;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + sizeof(WgBarriers) + 1024;

  CUtensorMap tm_g, tm_a;
  {
    uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)N * 2};
    uint32_t box[2] = {64, (uint32_t)kWgRows};
    DFV_TRY(make_tensor_map(&tm_g, DFV_BF16, 2, g, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {64, (uint32_t)kWgRows};
    DFV_TRY(make_tensor_map(&tm_a, DFV_BF16, 2, a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  DFV_TRY(init_timeout_word_tu());
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(pw_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DFV_CUDA(cudaFuncSetAttribute(pw_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const unsigned grid = (unsigned)(splits * tiles);
  const CUtensorMap& tp = p.p_is_g ? tm_g : tm_a;
  const CUtensorMap& tq = p.p_is_g ? tm_a : tm_g;
  if (a_scale)
    pw_wgrad_tc_kernel<true><<<grid, kWgThreads, smem, st>>>(tp, tq, (const __nv_bfloat16*)a_scale, dw, p);
  else
    pw_wgrad_tc_kernel<false><<<grid, kWgThreads, smem, st>>>(tp, tq, nullptr, dw, p);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

}  // namespace dfv
