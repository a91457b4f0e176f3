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
#include "small_linear.cuh"

namespace dfv {

// [ksplit][B][sq] partial sums + [B][sq] hidden pre-activations
static size_t se_scratch_floats(int B, int C, int squeeze) { return (size_t)((C + kSlKSlice - 1) / kSlKSlice + 1) * B * squeeze; }

// shared launcher of the inference and training entry points
template <bool kTrain>
static int launch_se(const float* pool_partial, int parts, float inv_hw, const float* w_reduce, const float* b_reduce,
                     const float* w_expand, const float* b_expand, void* gate, int gate_dtype, int B, int C, int squeeze,
                     float* pooled, float* h1, float* gate_f32, float* scratch, cudaStream_t st) {
  const int ksplit = sl_ksplit_rowmajor(C);
  const size_t smem_a = sl_rowmajor_smem(squeeze), smem_b = sl_kmajor_smem(squeeze);
  DFV_REQUIRE(smem_a <= 160 * 1024, "squeeze-excite: squeeze width %d too large", squeeze);
  const dim3 grid_a((unsigned)ksplit, (unsigned)((B + kSlRows - 1) / kSlRows));
  const dim3 grid_b((unsigned)((C + kSlCols - 1) / kSlCols), (unsigned)((B + kSlRows - 1) / kSlRows));
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(sl_rowmajor_partial_kernel<kTrain>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DFV_CUDA(cudaFuncSetAttribute(sl_kmajor_kernel<__nv_bfloat16, kTrain>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    DFV_CUDA(cudaFuncSetAttribute(sl_kmajor_kernel<float, kTrain>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  float* hid = h1 ? h1 : scratch + (size_t)ksplit * B * squeeze;      // training keeps h1 for the backward pass
  DFV_PDL((sl_rowmajor_partial_kernel<kTrain>), grid_a, kSlThreads, smem_a, st, pool_partial, parts, inv_hw, w_reduce, scratch, pooled, B, C, squeeze,
          (size_t)0, (size_t)0);
  DFV_PDL(sl_combine_kernel, (unsigned)(((size_t)B * squeeze + kSlThreads - 1) / kSlThreads), kSlThreads, 0, st, (const float*)scratch,
          (const float*)nullptr, ksplit, b_reduce, hid, B, squeeze, 0);
  if (gate_dtype == DFV_BF16)
    DFV_PDL((sl_kmajor_kernel<__nv_bfloat16, kTrain>), grid_b, kSlThreads, smem_b, st, (const float*)hid, w_expand, b_expand,
            (__nv_bfloat16*)gate, gate_f32, (float*)nullptr, B, C, squeeze, squeeze, SL_IN_SWISH, SL_OUT_SIGMOID, (const long long*)nullptr, (const float*)nullptr, 1.0f);
  else
    DFV_PDL((sl_kmajor_kernel<float, kTrain>), grid_b, kSlThreads, smem_b, st, (const float*)hid, w_expand, b_expand, (float*)gate,
            gate_f32, (float*)nullptr, B, C, squeeze, squeeze, SL_IN_SWISH, SL_OUT_SIGMOID, (const long long*)nullptr, (const float*)nullptr, 1.0f);
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

/* Second half of the gate when the squeeze layer ran inside the depthwise kernel (dfv_dwconv_se_fwd):
 * hid = b_reduce + 2^-30 * hid_fix;  gate[b][c] = sigmoid(b_expand[c] + sum_j w_expand_t[j][c] * swish(hid[b][j])). */
extern "C" int dfv_se_excite_fwd(const long long* hid_fix, const float* b_reduce, const float* w_expand_t, const float* b_expand, void* gate,
                                 int gate_dtype, int B, int C, int squeeze, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(hid_fix && b_reduce && w_expand_t && b_expand && gate, "dfv_se_excite_fwd: null pointer");
  DFV_REQUIRE(B > 0 && C > 0 && squeeze > 0 && valid_dtype(gate_dtype), "dfv_se_excite_fwd: bad shape / dtype");
  if (debug_flags() & 2) return DFV_OK;
  cudaStream_t st = as_stream(stream);
  const size_t smem_b = sl_kmajor_smem(squeeze);
  DFV_REQUIRE(smem_b <= 100 * 1024, "dfv_se_excite_fwd: squeeze width %d too large", squeeze);
  static thread_local bool configured = false;
  if (!configured) {
    DFV_CUDA(cudaFuncSetAttribute(sl_kmajor_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    DFV_CUDA(cudaFuncSetAttribute(sl_kmajor_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  ProfScope prof(PK_SE_GATE, 8.0 * B * squeeze + 4.0 * (double)C * squeeze + (double)dtype_size(gate_dtype) * B * C, 2.0 * B * (double)C * squeeze, st);
  const dim3 grid_b((unsigned)((C + kSlCols - 1) / kSlCols), (unsigned)((B + kSlRows - 1) / kSlRows));
  const float inv_scale = 1.0f / 1073741824.0f;    // dwconv.cu kSeFixScale
  if (gate_dtype == DFV_BF16)
    DFV_PDL((sl_kmajor_kernel<__nv_bfloat16, false>), grid_b, kSlThreads, smem_b, st, (const float*)nullptr, w_expand_t, b_expand,
            (__nv_bfloat16*)gate, (float*)nullptr, (float*)nullptr, B, C, squeeze, squeeze, SL_IN_SWISH, SL_OUT_SIGMOID, hid_fix, b_reduce, inv_scale);
  else
    DFV_PDL((sl_kmajor_kernel<float, false>), grid_b, kSlThreads, smem_b, st, (const float*)nullptr, w_expand_t, b_expand, (float*)gate,
            (float*)nullptr, (float*)nullptr, B, C, squeeze, squeeze, SL_IN_SWISH, SL_OUT_SIGMOID, hid_fix, b_reduce, inv_scale);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
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

