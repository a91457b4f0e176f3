// Probe for the planned tensor-core depthwise kernel -- NOT on the product path.
//
// Question it answers on the GPU: can a tcgen05.mma A operand start at ANY 128-byte pixel row of a SWIZZLE_128B tile
// that TMA wrote (not only at 1024-byte atom boundaries), so that tap (ky, kx) of a depthwise convolution is "the same
// tile, viewed from pixel offset ky*TWI + kx" with no im2col?  And which descriptor encoding does it need (the
// matrix-base-offset field, bits 49..51, = (start address >> 7) & 7, or none)?
//
//   out[p][c] = sum_t x[p + off[t]][c] * w[t][c]        p < 128, c < 64   (one 64-channel chunk, linearised pixels)
//
// One CTA.  x: [P][64] bf16 by TMA (SWIZZLE_128B, one 128-byte row per pixel).  Per tap, four MMAs M128 x N16 x K16, one
// per 16-channel block: A = x rows from off[t] on, columns blk*16.. (the usual 32-byte k-step advance inside the atom),
// B = diag(w[t][blk*16..+15]) stored as rows of a [16][64] SWIZZLE_128B tile, D = TMEM columns blk*16..+15.
#include "common.cuh"

namespace dfv {

constexpr int kProbeMaxTaps = 32;

__device__ __forceinline__ uint64_t probe_sw128_desc(uint32_t smem_addr, int with_base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;      // SBO: 8-row groups are 1024 B apart
  d |= (uint64_t)1 << 46;                // descriptor version (Blackwell)
  if (with_base_offset) d |= (uint64_t)((smem_addr >> 7) & 7) << 49;
  d |= (uint64_t)2 << 61;                // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(128) dw_tc_probe_kernel(const __grid_constant__ CUtensorMap tm_x, const __nv_bfloat16* __restrict__ w,
                                                          const int* __restrict__ offs, int taps, int P, int mode,
                                                          float* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* xt = smem;                                       // [P][128 B]
  unsigned char* bt = xt + (size_t)((P * 128 + 1023) / 1024) * 1024;   // [taps][16 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bt + (size_t)taps * 2048);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  __shared__ int s_off[kProbeMaxTaps];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < taps) s_off[tid] = offs[tid];
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  // B tiles: element (n, k) of tap t at  t*2048 + (n>>3)*1024 + (n&7)*128 + (((k>>3) ^ (n&7)) << 4) + (k&7)*2
  for (int i = tid; i < taps * 16 * 64; i += blockDim.x) {
    const int t = i / 1024, n = (i / 64) % 16, k = i % 64;
    const int blk = k >> 4, kk = k & 15;
    const __nv_bfloat16 v = kk == n ? w[(size_t)t * 64 + blk * 16 + n] : __float2bfloat16_rn(0.f);
    *reinterpret_cast<__nv_bfloat16*>(bt + (size_t)t * 2048 + (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ (n & 7))) << 4) + (k & 7) * 2) = v;
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(&bars[0], (uint32_t)P * 128);
    tma_load_2d(xt, &tm_x, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0, 1);
    tc_fence_after();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int t = 0; t < taps; ++t) {
      const uint32_t a_row = smem_u32(xt) + (uint32_t)s_off[t] * 128u;
      const uint32_t b_row = smem_u32(bt) + (uint32_t)t * 2048u;
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        const uint64_t da = probe_sw128_desc(a_row, mode) + (uint64_t)(blk * 2);     // 32-byte k-step inside the atom
        const uint64_t db = probe_sw128_desc(b_row, 0) + (uint64_t)(blk * 2);
        umma_bf16(tmem_base + (uint32_t)blk * 16, da, db, idesc, t != 0);
      }
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0, 2);
  tc_fence_after();
  // warp w reads TMEM lanes 32w..32w+31 (= output pixels), 64 columns (= channels)
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cc * 16, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) out[(size_t)(warp * 32 + lane) * 64 + cc * 16 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace dfv

using namespace dfv;

/* Debug probe (see the header of this file): x [P][64] bf16, w [taps][64] bf16, offs: DEVICE int[taps] pixel offsets with
 * off + 128 <= P, mode 0 / 1 = A descriptor without / with the matrix-base-offset field, out [128][64] fp32. */
extern "C" int dfv_debug_dwconv_tc_probe(const void* x, const void* w, const int* offs, int taps, int P, int mode, float* out,
                                         dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(x && w && offs && out && taps > 0 && taps <= kProbeMaxTaps && P >= 128 && P <= 256, "dfv_debug_dwconv_tc_probe: bad arguments");
  CUtensorMap tm;
  uint64_t dims[2] = {64, (uint64_t)P};
  uint64_t strides[1] = {128};
  uint32_t box[2] = {64, (uint32_t)P};
  DFV_TRY(make_tensor_map(&tm, DFV_BF16, 2, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  DFV_TRY(init_timeout_word_tu());
  const size_t smem = (size_t)((P * 128 + 1023) / 1024) * 1024 + (size_t)taps * 2048 + 64 + 1024;
  DFV_CUDA(cudaFuncSetAttribute(dw_tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dw_tc_probe_kernel<<<1, 128, smem, as_stream(stream)>>>(tm, (const __nv_bfloat16*)w, offs, taps, P, mode, out);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}
