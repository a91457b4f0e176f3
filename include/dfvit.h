/*
 * dfvit.h -- C ABI of libdfvit.so: the B200 (sm_100a) kernels behind the reference's
 * DeepfakeDetectionModel / CombinedLoss hot path.
 *
 * The reference (Ji-Hyeon212/Deepfake-ViT) has no FFI layer of its own: the path sits
 * behind a torch.nn.Module API (src/feature_extraction/feature_extractor.py:184-299) and
 * every GPU instruction is issued by stock PyTorch ops.  The Python drop-in
 * (deepfake_vit_b200/model.py) keeps that nn.Module API and state_dict layout and binds
 * these entry points through ctypes.  Each entry point cites the reference code it
 * replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's allocator) unless
 *     it says "host"; the library never allocates or frees device memory and keeps no
 *     pointer after a call returns
 *   - all work is enqueued on the passed stream; no implicit synchronisation, no use of
 *     the default stream; every call is CUDA-graph capturable
 *   - activations are NHWC ("channels last", C contiguous): [B][H][W][C]
 *   - dtype selects the activation storage type: DFV_F32 (parity mode, fp32 SIMT GEMMs)
 *     or DFV_BF16 (bf16 storage, fp32 accumulation, tcgen05 tensor-core GEMMs)
 *   - return value: DFV_OK (0) or a negative error code; dfv_last_error() gives the
 *     message (thread local).  No exceptions, no aborts.
 *   - sm_100 only: every launch entry refuses other devices (DFV_ERR_DEVICE). There is no
 *     CPU path.
 */
#ifndef DFVIT_H_
#define DFVIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dfv_stream_t; /* cudaStream_t */

enum { DFV_F32 = 0, DFV_BF16 = 1 };
enum { DFV_ACT_NONE = 0, DFV_ACT_SILU = 1 };

enum {
  DFV_OK = 0,
  DFV_ERR_INVALID = -1, /* bad argument (shape, alignment, null pointer) */
  DFV_ERR_CUDA = -2,    /* a CUDA runtime / driver call failed */
  DFV_ERR_DEVICE = -3,  /* current device is not sm_100 */
  DFV_ERR_WORKSPACE = -4 /* workspace too small */
};

int dfv_version(void);
const char* dfv_last_error(void);
/* DFV_OK if the current CUDA device can run the kernels (compute capability 10.x). */
int dfv_device_check(void);
/* Debug/localisation switch: route bf16 1x1 convolutions through the SIMT kernel that the
 * fp32 mode uses instead of tcgen05.  Tests only; default 0. */
void dfv_debug_force_simt_gemm(int on);
/* Debug: nonzero if a bounded in-kernel barrier wait starved (the kernel then traps): bit 31 set,
 * bits 24-30 = which wait, bits 0-23 = block index.  Readable after a launch failure. */
unsigned int dfv_debug_last_timeout(void);
/* Number of kernels launched by this thread since the last reset (bench.py's gpu_launches). */
long long dfv_launch_count(int reset);

/* Per-launch profiler (CUDA events on the launching stream around every operator launch), used
 * by bench.py for the roofline leg.  enable(1) clears and starts recording, enable(0) stops.
 * get(): kind (0 stem, 1 expand/head GEMM, 2 depthwise, 3 SE gate, 4 project GEMM, 5 heat-map,
 * 6 attention, 7 MLP head, 8 loss, 9 SIMT GEMM), the launch's ALGORITHMIC bytes and flops, and
 * its duration in ms (synchronises on the launch's stop event). */
int dfv_profile_enable(int on);
int dfv_profile_count(void);
int dfv_profile_get(int idx, int* kind, double* bytes, double* flops, float* ms);

/* ------------------------------------------------------------------------------------
 * EfficientNet-B4 topology (efficientnet-pytorch 0.7.1 `from_name('efficientnet-b4')`,
 * called at src/feature_extraction/efficientnet.py:42-45).
 * The static "SAME" pads are those of the 380 chain regardless of the live input size.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int32_t c_in, c_mid, c_out; /* c_mid = expand_ratio * c_in */
  int32_t kernel, stride;
  int32_t pad_lo, pad_hi;     /* static pad: left = top = pad_lo, right = bottom = pad_hi */
  int32_t se_squeeze;         /* max(1, c_in / 4) */
  int32_t has_expand;         /* 0 for blocks 0 and 1 */
  int32_t has_skip;           /* stride == 1 && c_in == c_out */
} dfv_block_info;

int dfv_b4_num_blocks(void);                      /* 32 */
int dfv_b4_block(int idx, dfv_block_info* out);
int dfv_b4_stem_channels(void);                   /* 48 */
int dfv_b4_head_channels(void);                   /* 1792 */
/* Output spatial size of the backbone for an H x W input under the static pads. */
int dfv_b4_output_hw(int H, int W, int* Ho, int* Wo);

/* ------------------------------------------------------------------------------------
 * Packed ("folded") weight blob for inference.  Eval-mode BatchNorm is folded into the
 * preceding convolution: w' = w * gamma / sqrt(var + eps), bias = beta - mean * scale.
 * The caller asks for the offset of each tensor, writes the folded values there, and
 * hands the blob to dfv_infer_fwd().  Slot layouts (row-major):
 *   DFV_W_STEM         fp32 [3][3][3][48]   (kh, kw, c_in, c_out)
 *   DFV_W_EXPAND/PROJECT/HEAD   T [N][K]    (T = bf16 or fp32 by dtype; K contiguous)
 *   DFV_W_DW           fp32 [k*k][C]
 *   DFV_W_SE_REDUCE    fp32 [sq][C]         DFV_W_SE_EXPAND  fp32 [sq][C] (transposed)
 *   *_BIAS             fp32 [N]
 * block index: 0..31 for block tensors, -1 for stem / head.
 * ---------------------------------------------------------------------------------- */
enum {
  DFV_W_STEM = 0, DFV_W_STEM_BIAS,
  DFV_W_EXPAND, DFV_W_EXPAND_BIAS,
  DFV_W_DW, DFV_W_DW_BIAS,
  DFV_W_SE_REDUCE, DFV_W_SE_REDUCE_BIAS, DFV_W_SE_EXPAND, DFV_W_SE_EXPAND_BIAS,
  DFV_W_PROJECT, DFV_W_PROJECT_BIAS,
  DFV_W_HEAD, DFV_W_HEAD_BIAS,
  DFV_W_KINDS
};
size_t dfv_blob_bytes(int dtype);
/* Byte offset and element count of a slot; returns <0 on a bad (block, kind). */
int dfv_blob_slot(int dtype, int block, int kind, size_t* offset, size_t* elems);

/* ------------------------------------------------------------------------------------
 * Per-operator entry points.
 * ---------------------------------------------------------------------------------- */

/* Stem: ZeroPad2d(0,1,0,1) + conv3x3/s2 3->C + BN + swish.
 * Replaces `_swish(_bn0(_conv_stem(x)))` in EfficientNet.extract_features (third-party
 * 0.7.1, reached from efficientnet.py:163).  x is the reference's input contract
 * (src/data/dataset.py:82-116): NCHW fp32.  y: NHWC [B][Ho][Wo][C]. */
int dfv_stem_conv_fwd(const float* x_nchw, const float* w_khwc, const float* bias, void* y, int dtype,
                      int B, int H, int W, int C, dfv_stream_t stream);

/* Depthwise k x k conv (k in {3,5}, stride in {1,2}) with the static asymmetric pad
 * applied by TMA out-of-bounds zero fill (no padded copy), folded BN, swish, and the SE
 * global-average-pool partial sums emitted by the same kernel.
 * Replaces `_swish(_bn1(_depthwise_conv(x)))` + `F.adaptive_avg_pool2d(x, 1)` of
 * MBConvBlock.forward.  pool_partial: fp32 [B][parts][C] (may be NULL), parts =
 * dfv_dwconv_pool_parts(...).  x: [B][H][W][C], y: [B][Ho][Wo][C]. */
int dfv_dwconv_pool_parts(int dtype, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi);
int dfv_dwconv_fwd(const void* x, const float* w_kkc, const float* bias, void* y, float* pool_partial,
                   int dtype, int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi,
                   int act, dfv_stream_t stream);

/* Squeeze-excite gate: mean over H*W (finishing the partial sums) -> 1x1 conv + bias ->
 * swish -> 1x1 conv + bias -> sigmoid.  Replaces `_se_expand(_swish(_se_reduce(pool)))`
 * and `torch.sigmoid` of MBConvBlock.forward.  gate: [B][C] of type gate_dtype (fp32 in
 * parity mode; bf16 in bf16 mode, where the autocast reference's sigmoid output is a bf16
 * tensor too).  The channel rescale itself is fused into the project GEMM's A-operand path
 * (dfv_pw_gemm_fwd a_scale). */
int dfv_se_gate_fwd(const float* pool_partial, int parts, float inv_hw, const float* w_reduce,
                    const float* b_reduce, const float* w_expand_t, const float* b_expand, void* gate,
                    int gate_dtype, int B, int C, int squeeze, dfv_stream_t stream);

/* Pointwise (1x1) convolution as a GEMM  out[M][N] = act((a[M][K] * a_scale) . w[N][K]^T + bias) + residual.
 * M = B*H*W rows.  a_scale: [M / rows_per_image][K] per-image channel scale (SE gate) of the
 * activation dtype, or NULL; residual: [M][N] or NULL (the MBConv identity skip).  Replaces
 * `_expand_conv+_bn0+_swish`, `sigmoid(se) * x` + `_project_conv+_bn2` + `x + inputs`, and
 * `_conv_head+_bn1+_swish`.
 * bf16: TMA -> smem -> tcgen05.mma (fp32 accumulators in TMEM) -> tcgen05.ld epilogue -> swizzled
 * smem staging -> TMA store.  K % 8 == 0 and N % 8 == 0 required. */
int dfv_pw_gemm_fwd(const void* a, const void* w, const float* bias, const void* a_scale, int rows_per_image,
                    const void* residual, void* out, int dtype, long long M, int K, int N, int act,
                    dfv_stream_t stream);

/* Landmark heat-map of LandmarkAttention._create_attention_map
 * (src/feature_extraction/landmark_attention.py:76-130), op order of SURVEY.md B.2:
 * coordinates scaled by W/ref_size (ref_size = 224.0 hard-coded at :97-98), five weighted
 * Gaussians (sigma 1.5), division by the maximum over each group of `group` consecutive
 * images (+1e-8) -- group = B reproduces the reference's whole-call maximum (:125) --
 * clamp to [0.1, 1].  heat: fp32 [B][H][W]; raw_ws: fp32 [B*H*W] scratch; max_ws:
 * uint32 [ceil(B/group)] scratch.  scaled_xy (optional, fp32 [B][5][2]) receives the
 * scaled coordinates for the bit-exactness test. */
int dfv_landmark_heatmap_fwd(const float* landmarks, const float* weights5, float* heat, float* raw_ws,
                             uint32_t* max_ws, float* scaled_xy, int B, int H, int W, float ref_size,
                             float sigma, int group, dfv_stream_t stream);

/* HybridAttention (landmark -> channel -> spatial, landmark_attention.py:283-310) fused
 * with the global average pool of DeepfakeFeatureExtractor.forward
 * (feature_extractor.py:108-109).  fmap: [B][H*W][C]; heat: fp32 [B][H*W] or NULL
 * (landmarks=None skips the stage, :299); ca_w1: fp32 [hidden][C] (fc.0.weight); ca_w2_t:
 * fp32 [hidden][C] (fc.2.weight transposed); sa_w: fp32 [2][7][7] (spatial_attn.conv.weight).
 * use_channel / use_spatial mirror attention_config.  features: fp32 [B][C].  The attended
 * map is never written.  Optional debug outputs (may be NULL): channel_gate fp32 [B][C],
 * spatial_gate fp32 [B][H*W]. */
int dfv_hybrid_attention_fwd(const void* fmap, const float* heat, const float* ca_w1, const float* ca_w2_t,
                             const float* sa_w, float* features, float* channel_gate, float* spatial_gate,
                             int dtype, int B, int H, int W, int C, int hidden, int use_channel,
                             int use_spatial, dfv_stream_t stream);

/* Classifier head `nn.Sequential` of DeepfakeDetectionModel.__init__
 * (feature_extractor.py:223-238), eval mode: n_layers Linear layers, BatchNorm1d folded
 * into each hidden Linear, ReLU after every layer but the last; Dropout is identity.
 * dims: host int[n_layers + 1]; w_t[l]: fp32 [dims[l]][dims[l+1]] (transposed), b[l]: fp32
 * [dims[l+1]].  w_t / b are HOST arrays of n_layers DEVICE pointers.  n_layers <= 8. */
int dfv_mlp_head_fwd(const float* features, const float* const* w_t, const float* const* b, const int* dims,
                     int n_layers, float* logits, int B, dfv_stream_t stream);

/* CombinedLoss.forward (src/training/losses.py:192-247) and its gradient in one pass:
 * weighted-mean CE + focal(gamma=2, alpha=class weights) + contrastive (consecutive pairs,
 * euclidean, margin 1, pairwise_distance eps 1e-6).  features may be NULL (Evaluator's call,
 * evaluator.py:96) -> no contrastive term.  class_weights: fp32 [C] or NULL.  losses: fp32
 * [4] = {ce, focal, contrastive, total}; has_contrastive (host out) tells whether the term
 * exists (B >= 2).  dlogits fp32 [B][C] / dfeatures fp32 [B][D]: d total / d input (may be
 * NULL to skip the gradient).  targets: int64 [B]. */
int dfv_combined_loss_fwd_bwd(const float* logits, const int64_t* targets, const float* features,
                              const float* class_weights, float w_ce, float w_focal, float w_contrastive,
                              float* losses, float* dlogits, float* dfeatures, int B, int C, int D,
                              int* has_contrastive, dfv_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Whole-path inference: DeepfakeDetectionModel.forward in eval mode
 * (feature_extractor.py:242-269 -> :74-117 -> efficientnet.py:153-163 -> 32 MBConv blocks
 * -> HybridAttention -> pool -> classifier), one call, ~110 kernel launches on `stream`.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int32_t dtype;
  int32_t B, H, W;
  int32_t use_attention, use_landmark, use_channel, use_spatial;
  int32_t heat_group;            /* images per heat-map max group; 0 = whole batch */
  float landmark_ref_size;       /* 224.0 (landmark_attention.py:97-98) */
  const void* blob;              /* folded backbone weights, dfv_blob_* layout */
  const float* images_nchw;      /* [B][3][H][W] fp32 */
  const float* landmarks;        /* [B][5][2] fp32 or NULL */
  const float* lm_weights;       /* [5] */
  const float* ca_w1;            /* [hidden][1792] */
  const float* ca_w2_t;          /* [hidden][1792] */
  int32_t ca_hidden;
  const float* sa_w;             /* [2][7][7] */
  const float* const* head_w_t;  /* host array of device pointers */
  const float* const* head_b;
  const int32_t* head_dims;      /* host int[n+1] */
  int32_t head_layers;
  void* workspace;
  size_t workspace_bytes;
  float* logits;                 /* [B][num_classes] */
  float* features;               /* [B][1792] */
  float* heat;                   /* optional out: fp32 [B][Ho*Wo] (may be NULL) */
  void* const* taps;             /* optional: host array of 34 device pointers (stem, block0..31,
                                    head) receiving NHWC copies of each stage output; entries may
                                    be NULL; NULL array = no taps */
} dfv_infer_args;

size_t dfv_infer_workspace_bytes(int dtype, int B, int H, int W);
int dfv_infer_fwd(const dfv_infer_args* args, dfv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DFVIT_H_ */
