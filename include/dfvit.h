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
enum { DFV_ACT_NONE = 0, DFV_ACT_SILU = 1, DFV_ACT_RELU = 2 };

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
/* Diagnostics: nonzero if a bounded in-kernel barrier wait starved (the kernel then traps): bit 31 set,
 * bits 24-30 = which wait, bits 0-23 = block index.  Readable after a launch failure (appended to the
 * Python wrapper's error message). */
unsigned int dfv_last_timeout_word(void);
/* Number of kernels launched by this process since the last reset (bench.py's gpu_launches). */
long long dfv_launch_count(int reset);

/* Per-launch profiler (CUDA events on the launching stream around every operator launch), used
 * by bench.py for the roofline leg.  enable(1) clears and starts recording, enable(0) stops.
 * get(): kind (0 stem, 1 expand/head GEMM, 2 depthwise, 3 SE gate, 4 project GEMM, 5 heat-map,
 * 6 attention, 7 MLP head, 8 loss, 9 SIMT GEMM, 10 BatchNorm/activation pass, 11 1x1 weight gradient,
 * 12 depthwise backward), the launch's ALGORITHMIC bytes and flops, and its duration in ms
 * (synchronises on the launch's stop event). */
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
                      int B, int H, int W, int C, int act, dfv_stream_t stream);

/* The same stem fed by the RAW crop: uint8 RGB, HWC ([B][H][W][3], what cv2.imread + cvtColor gives before the
 * reference's `image.transpose(2,0,1)`), with the input normalisation of the reference's data path fused into
 * the operand load: v = (u8 / 255 - mean[c]) / std[c] in fp32 with IEEE division, exactly the op order of
 * PreprocessedFaceDataset.__getitem__ (src/data/dataset.py:92-98; constants :59-60; task.ipynb:386 does the same
 * on the device).  The 3 x 256 possible values are tabulated per CTA, so the fp32 path is bit-identical to
 * feeding dfv_stem_conv_fwd the normalised NCHW fp32 tensor.  norm6: HOST float[6] = mean[3], std[3].
 * 4x fewer input bytes from HBM and over PCIe than the fp32 NCHW contract. */
int dfv_stem_conv_u8_fwd(const uint8_t* x_hwc, const float* norm6, const float* w_khwc, const float* bias, void* y,
                         int dtype, int B, int H, int W, int C, int act, dfv_stream_t stream);
/* The normalisation alone: uint8 HWC -> fp32 NCHW (the reference's input contract), same arithmetic.  Used by the
 * training path, whose stem weight gradient reads the fp32 image. */
int dfv_u8_to_nchw_f32(const uint8_t* x_hwc, const float* norm6, float* y_nchw, int B, int H, int W, dfv_stream_t stream);

/* Depthwise k x k conv (k in {3,5}, stride in {1,2}) with the static asymmetric pad
 * applied by TMA out-of-bounds zero fill (no padded copy), folded BN, swish, and the SE
 * global-average-pool partial sums emitted by the same kernel.
 * Replaces `_swish(_bn1(_depthwise_conv(x)))` + `F.adaptive_avg_pool2d(x, 1)` of
 * MBConvBlock.forward.  pool_partial: fp32 [B][parts][C] (may be NULL), parts =
 * dfv_dwconv_pool_parts(...) (depends on B: one slot per persistent CTA that touches an image).  x: [B][H][W][C], y: [B][Ho][Wo][C]. */
int dfv_dwconv_pool_parts(int dtype, int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi);
int dfv_dwconv_fwd(const void* x, const float* w_kkc, const float* bias, void* y, float* pool_partial,
                   int dtype, int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi,
                   int act, dfv_stream_t stream);

/* Squeeze-excite gate: mean over H*W (finishing the partial sums) -> 1x1 conv + bias ->
 * swish -> 1x1 conv + bias -> sigmoid.  Replaces `_se_expand(_swish(_se_reduce(pool)))`
 * and `torch.sigmoid` of MBConvBlock.forward.  gate: [B][C] of type gate_dtype (fp32 in
 * parity mode; bf16 in bf16 mode, where the autocast reference's sigmoid output is a bf16
 * tensor too).  The channel rescale itself is fused into the project GEMM's A-operand path
 * (dfv_pw_gemm_fwd a_scale).  scratch: dfv_se_scratch_floats(B, C, squeeze) floats of device memory (the squeeze
 * layer's per-channel-slice partial sums, handed from the squeeze kernel to the excite kernel). */
size_t dfv_se_scratch_floats(int B, int C, int squeeze);
int dfv_se_gate_fwd(const float* pool_partial, int parts, float inv_hw, const float* w_reduce,
                    const float* b_reduce, const float* w_expand_t, const float* b_expand, void* gate,
                    int gate_dtype, float* scratch, int B, int C, int squeeze, dfv_stream_t stream);

/* The same depthwise convolution with the SQUEEZE layer of the SE block fused into its tail, and the excite layer as the
 * one remaining launch (three dependent launches -> one; replaces the same MBConvBlock lines as the two entries above):
 *   dfv_dwconv_se_fwd   y and pool_partial as dfv_dwconv_fwd (swish), and every CTA adds the partial hidden sums of its
 *                       channel chunk and images into 64-bit FIXED-POINT accumulators (2^-30 resolution)
 *                         hid_fix[b][j] += round(2^30 * sum_{c in chunk} w_reduce[j][c] * mean_hw(y[b][.][c]))
 *                       Integer addition is associative: the sums are bit-reproducible whatever the CTA arrival order,
 *                       without tickets or waiting.  hid_fix: int64 [B][squeeze], ZERO on entry.  zero_next (optional,
 *                       != hid_fix): zero_count int64 entries this launch zeroes for the NEXT fused layer (two buffers
 *                       alternate through the network, so no per-layer memset is needed).
 *   dfv_se_excite_fwd   hid = b_reduce + 2^-30 * hid_fix;  gate[b][c] = sigmoid(b_expand[c] + sum_j w_expand_t[j][c] * swish(hid[b][j]))
 * dfv_dwconv_se_supported() says whether the layer's tile plan leaves room for the tail (else use the three-launch path);
 * dfv_dwconv_se_profitable() is the measured policy dfv_infer_fwd follows. */
int dfv_dwconv_se_supported(int dtype, int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, int squeeze);
int dfv_dwconv_se_profitable(int dtype, int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, int squeeze);
int dfv_dwconv_se_fwd(const void* x, const float* w_kkc, const float* bias, void* y, float* pool_partial, const float* w_reduce,
                      long long* hid_fix, long long* zero_next, long long zero_count, int squeeze, int dtype, int B, int H, int W,
                      int C, int kernel, int stride, int pad_lo, int pad_hi, dfv_stream_t stream);
int dfv_se_excite_fwd(const long long* hid_fix, const float* b_reduce, const float* w_expand_t, const float* b_expand, void* gate,
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

/* The same 1x1 convolution, with thin bf16 layers (K <= 48 channels: the 190x190-stage project / expand convs) run
 * ROW-FOLDED: [M][K] read as [M/f][f*K] against the block-diagonal weight diag(W,...,W) gives the same output memory
 * [M/f][f*N] with f x fewer, f x fatter tensor-core tiles.  fold_ws: dfv_pw_fold_ws_bytes(B) bytes of device scratch
 * (NULL = never fold); B = number of images (rows of a_scale). */
size_t dfv_pw_fold_ws_bytes(int B);
int dfv_pw_conv_fwd(const void* a, const void* w, const float* bias, const void* a_scale, int rows_per_image,
                    const void* residual, void* out, int dtype, int B, long long M, int K, int N, int act,
                    void* fold_ws, dfv_stream_t stream);

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
/* The same with a floor on the maximum: the `global` normaliser mode of data-parallel runs (SURVEY.md 8(e) caveat 1).
 * max_ws holds ORDER-PRESERVING uint32 keys of the group maxima (unsigned order == float order); max_floor (device, one
 * key, may be NULL) is the element-wise maximum of every rank's max_ws[0] (a 1-word all-reduce(MAX) run by the host
 * between two calls): the map is divided by max(local maximum, floor) -- for the global maximum that IS the value a
 * single-GPU run over the concatenated batch divides by (landmark_attention.py:125).  Whole-call group only. */
int dfv_landmark_heatmap_fwd_ex(const float* landmarks, const float* weights5, float* heat, float* raw_ws,
                                uint32_t* max_ws, float* scaled_xy, int B, int H, int W, float ref_size,
                                float sigma, int group, const uint32_t* max_floor, dfv_stream_t stream);

/* HybridAttention (landmark -> channel -> spatial, landmark_attention.py:283-310) fused
 * with the global average pool of DeepfakeFeatureExtractor.forward
 * (feature_extractor.py:108-109).  fmap: [B][H*W][C]; heat: fp32 [B][H*W] or NULL
 * (landmarks=None skips the stage, :299); ca_w1: fp32 [hidden][C] (fc.0.weight); ca_w2_t:
 * fp32 [hidden][C] (fc.2.weight transposed); sa_w: fp32 [2][7][7] (spatial_attn.conv.weight).
 * use_channel / use_spatial mirror attention_config.  features: fp32 [B][C].  The attended
 * map is never written.  Optional debug outputs (may be NULL): channel_gate fp32 [B][C],
 * spatial_gate fp32 [B][H*W].  scratch: dfv_attention_scratch_floats(B, H, W, C, hidden) floats of device memory
 * (channel statistics, the channel-attention MLP's intermediates, spatial statistics). */
size_t dfv_attention_scratch_floats(int B, int H, int W, int C, int hidden);
int dfv_hybrid_attention_fwd(const void* fmap, const float* heat, const float* ca_w1, const float* ca_w2_t,
                             const float* sa_w, float* features, float* channel_gate, float* spatial_gate,
                             float* scratch, int dtype, int B, int H, int W, int C, int hidden, int use_channel,
                             int use_spatial, dfv_stream_t stream);

/* Classifier head `nn.Sequential` of DeepfakeDetectionModel.__init__
 * (feature_extractor.py:223-238), eval mode: n_layers Linear layers, BatchNorm1d folded
 * into each hidden Linear, ReLU after every layer but the last; Dropout is identity.
 * dims: host int[n_layers + 1]; w_t[l]: fp32 [dims[l]][dims[l+1]] (transposed), b[l]: fp32
 * [dims[l+1]].  w_t / b are HOST arrays of n_layers DEVICE pointers.  n_layers <= 8.
 * One split-K kernel per layer; the hidden activations ping-pong through `scratch` (fp32 device buffer of
 * dfv_mlp_head_scratch_floats(...) floats; may be NULL when n_layers == 1). */
size_t dfv_mlp_head_scratch_floats(const int* dims, int n_layers, int B);
int dfv_mlp_head_fwd(const float* features, const float* const* w_t, const float* const* b, const int* dims,
                     int n_layers, float* logits, float* scratch, int B, dfv_stream_t stream);

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
/* The same with the weighted cross-entropy's normaliser supplied by the caller: ce_norm (device, one float, may be NULL
 * = the local sum of w[y_i], nn.CrossEntropyLoss(weight), losses.py:188,217).  Data-parallel exact form (SURVEY.md 8(e)
 * caveat 3): every rank passes the MEAN over ranks of its dfv_class_weight_sum -- then the mean over ranks of the
 * returned ce (and of its gradient, which the gradient all-reduce forms) is the weighted CE of the global batch. */
int dfv_combined_loss_fwd_bwd_ex(const float* logits, const int64_t* targets, const float* features,
                                 const float* class_weights, float w_ce, float w_focal, float w_contrastive,
                                 float* losses, float* dlogits, float* dfeatures, int B, int C, int D,
                                 int* has_contrastive, const float* ce_norm, dfv_stream_t stream);
/* out[0] = sum_i class_weights[targets[i]] (B if class_weights is NULL): fixed summation order. */
int dfv_class_weight_sum(const int64_t* targets, const float* class_weights, float* out, int B, int C,
                         dfv_stream_t stream);

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
  const float* images_nchw;      /* [B][3][H][W] fp32 (NULL when images_u8 is given) */
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
  const uint8_t* images_u8;      /* optional: [B][H][W][3] uint8 RGB crops; normalised inside the stem
                                    (dfv_stem_conv_u8_fwd) with u8_norm = mean[3], std[3] */
  float u8_norm[6];
  const uint32_t* heat_max_floor; /* optional (device, one key): floor of the heat-map maximum, see dfv_landmark_heatmap_fwd_ex */
} dfv_infer_args;

size_t dfv_infer_workspace_bytes(int dtype, int B, int H, int W);
int dfv_infer_fwd(const dfv_infer_args* args, dfv_stream_t stream);


/* ====================================================================================
 * Training path: train-mode forward (batch-statistics BatchNorm,
 * dropout, drop-connect) and the backward pass that `loss.backward()` runs in the reference
 * (src/training/trainer.py:140-153).  Building blocks first, then the whole-path sequencer.
 * Tensors are channels-last [B][rows_per_image][C]; C % 8 == 0.
 * ==================================================================================== */

/* Number of row chunks per image the streaming kernels use (= SE pool `parts` of dfv_bn_act_fwd). */
int dfv_rows_chunks(int B, long long rows_per_image);
/* fp32 scratch the BatchNorm statistics / backward-reduce kernels need. */
size_t dfv_bn_ws_floats(int B, long long rows_per_image, int C);

/* nn.BatchNorm2d / BatchNorm1d in train mode, statistics half: per-channel mean and 1/sqrt(biased var + eps)
 * of raw; running_mean / running_var (may be NULL) updated with `momentum` and the UNBIASED variance. */
int dfv_bn_stats_fwd(const void* raw, int dtype, int B, long long rows_per_image, int C, float eps, float momentum,
                     float* mean, float* invstd, float* running_mean, float* running_var, float* ws, dfv_stream_t stream);
/* out = act(gamma * (raw - mean) * invstd + beta) [* mask] [* rowscale[image]] [+ residual].
 * mean/invstd/gamma/beta may be NULL (identity).  mask: fp32 [B][rows][C] (dropout keep-scale) or NULL.
 * rowscale: fp32 [B] (drop-connect) or NULL.  pool_partial: fp32 [B][dfv_rows_chunks()][C] per-chunk sums of the
 * activated values (SE squeeze) or NULL.  `_swish(_bnX(conv))`, `_bn2(..)` + drop_connect + skip of MBConvBlock.forward. */
int dfv_bn_act_fwd(const void* raw, const float* mean, const float* invstd, const float* gamma, const float* beta, int act,
                   const float* rowscale, const void* residual, const float* mask, void* out, float* pool_partial,
                   int dtype, int B, long long rows_per_image, int C, dfv_stream_t stream);
/* Backward through [mask, rowscale, SE gate] -> activation -> BatchNorm, reduction half:
 *   gin = (g * gate[image][c] + dpool[image][c] * inv_hw) * rowscale[image] * mask;  du = gin * act'(u)
 * writes du (may alias g; NULL = do not write it), dgamma = sum du * xhat, dbeta = sum du (may be NULL), coef [2][C] = the two means. */
int dfv_act_bn_bwd(const void* g, const void* raw, const float* mean, const float* invstd, const float* gamma,
                   const float* beta, int act, const void* gate, const float* dpool, float inv_hw, const float* rowscale,
                   const float* mask, void* du, float* dgamma, float* dbeta, float* coef, float* ws, int dtype, int B,
                   long long rows_per_image, int C, dfv_stream_t stream);
/* The two halves without the du round trip through HBM: call dfv_act_bn_bwd with du = NULL (reduction only: dgamma, dbeta,
 * coef), then this -- it recomputes du from (g, raw) and writes  d raw = gamma * invstd * (du - coef[0] - xhat * coef[1])
 * (draw may alias g).  One write pass less per BatchNorm backward; du stays fp32 between the halves. */
int dfv_act_bn_bwd_apply(const void* g, const void* raw, const float* mean, const float* invstd, const float* gamma, const float* beta,
                         int act, const void* gate, const float* dpool, float inv_hw, const float* rowscale, const float* mask,
                         const float* coef, void* draw, int dtype, int B, long long rows_per_image, int C, dfv_stream_t stream);
/* d raw = gamma * invstd * (du - coef[0] - xhat * coef[1])   (draw may alias du) */
int dfv_bn_bwd_apply(const void* du, const void* raw, const float* mean, const float* invstd, const float* gamma,
                     const float* coef, void* draw, int dtype, long long M, int C, dfv_stream_t stream);

/* Squeeze-excite forward that also saves pooled [B][C], h1 [B][sq] (pre-swish) and the fp32 gate; weights in
 * torch layout: w_reduce [sq][C], w_expand [C][sq].  scratch: dfv_se_scratch_floats(B, C, squeeze) floats. */
int dfv_se_train_fwd(const float* pool_partial, int parts, float inv_hw, const float* w_reduce, const float* b_reduce,
                     const float* w_expand, const float* b_expand, void* gate, int gate_dtype, float* pooled, float* h1,
                     float* gate_f32, float* scratch, int B, int C, int squeeze, dfv_stream_t stream);
size_t dfv_se_bwd_ws_floats(int B, long long rows_per_image, int C, int squeeze);
int dfv_se_bwd(const void* da, const void* d, int dtype, const float* gate_f32, const float* pooled, const float* h1,
               const float* w_reduce, const float* w_expand, float* dpool, float* dw_reduce, float* db_reduce,
               float* dw_expand, float* db_expand, float* ws, int B, long long rows_per_image, int C, int squeeze,
               dfv_stream_t stream);

/* One fp32 Linear layer with few rows (the classifier inside the training step): out[b][n] = bias[n] (may be NULL) +
 * sum_k in[b][k] * W, with W = w[n][k] (torch layout) or, w_kmajor = 1, w[k][n] -- the same torch tensor read as the transposed
 * layer, which is the input gradient of nn.Linear (feature_extractor.py:223-238) without a transposed copy.  scratch:
 * dfv_linear_f32_scratch_floats(B, K, N) floats (K-slice partial sums, added in fixed order). */
size_t dfv_linear_f32_scratch_floats(int B, int K, int N);
int dfv_linear_f32_fwd(const float* in, const float* w, const float* bias, float* out, float* scratch, size_t scratch_floats,
                       int B, int K, int N, int w_kmajor, dfv_stream_t stream);

/* The same backward with one streaming pass less (bf16 tensors, swish): the SE backward needs sum_hw(da * d) before the layer's
 * input gradient exists (its dpool enters that gradient), but the BatchNorm reduction is LINEAR in (gate, dpool) per image.
 * dfv_act_bn_bwd_gated_reduce reads (da, d_raw) once and writes per-(image, chunk) rows ws4 [B][chunks][4][C] (sums of da act'(u),
 * act'(u), and both times x) plus the SE dot partials [B][chunks][C] at the start of se_ws (d = swish(u) recomputed from d_raw);
 * dfv_se_bwd_from_partials is dfv_se_bwd without its own pass over (da, d); dfv_bn_bwd_gated_finalize combines the rows with
 * the gate (activation dtype, as the forward applied it) and dpool into dgamma / dbeta (may be NULL) and coef [2][C] -- the values
 * dfv_act_bn_bwd(gate, dpool) returns.  ws4: dfv_bn_ws_floats() floats. */
int dfv_act_bn_bwd_gated_reduce(const void* g, const void* raw, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                float* ws4, float* partial_d, int dtype, int B, long long rows_per_image, int C, dfv_stream_t stream);
int dfv_se_bwd_from_partials(const float* gate_f32, const float* pooled, const float* h1, const float* w_reduce, const float* w_expand,
                             float* dpool, float* dw_reduce, float* db_reduce, float* dw_expand, float* db_expand, float* ws, int B,
                             long long rows_per_image, int C, int squeeze, dfv_stream_t stream);
int dfv_bn_bwd_gated_finalize(const float* ws4, const void* gate, const float* dpool, float inv_hw, const float* mean, const float* invstd,
                              float* dgamma, float* dbeta, float* coef, int dtype, int B, long long rows_per_image, int C,
                              dfv_stream_t stream);

/* 1x1 conv weight gradient: dw[N][K] (fp32, torch layout, caller zeroes) += sum_m g[m][n] a[m][k] a_scale[m/rpi][k]. */
int dfv_pw_wgrad(const void* g, const void* a, const void* a_scale, int rows_per_image, float* dw, int dtype, long long M,
                 int K, int N, dfv_stream_t stream);
/* Depthwise conv gradients (H, W = INPUT size of the forward conv; w / dw are fp32 [k*k][C]). */
int dfv_dwconv_dgrad(const void* g, const float* w_kkc, void* dx, int dtype, int B, int H, int W, int C, int kernel,
                     int stride, int pad_lo, int pad_hi, dfv_stream_t stream);
int dfv_dwconv_wgrad(const void* g, const void* x, float* dw_kkc, int dtype, int B, int H, int W, int C, int kernel,
                     int stride, int pad_lo, int pad_hi, dfv_stream_t stream);
int dfv_stem_wgrad(const void* g, const float* x_nchw, float* dw, int dtype, int B, int H, int W, dfv_stream_t stream);

/* HybridAttention train forward / backward and the heat-map backward (attention_train.cu). */
size_t dfv_attention_saved_floats(int B, int H, int W, int C, int hidden);
int dfv_hybrid_attention_train_fwd(const void* fmap, const float* heat, const float* ca_w1, const float* ca_w2,
                                   const float* sa_w, float* features, float* saved, int dtype, int B, int H, int W, int C,
                                   int hidden, int use_channel, int use_spatial, dfv_stream_t stream);
int dfv_hybrid_attention_bwd(const void* fmap, const float* heat, const float* ca_w1, const float* ca_w2,
                             const float* sa_w, const float* dfeatures, const float* saved, void* dfmap, float* dheat,
                             float* dca_w1, float* dca_w2, float* dsa_w, float* ws, int dtype, int B, int H, int W, int C,
                             int hidden, int use_channel, int use_spatial, dfv_stream_t stream);
int dfv_landmark_heatmap_bwd(const float* landmarks, const float* weights5, const float* raw_ws, const uint32_t* max_ws,
                             const float* dheat, float* dweights5, int B, int H, int W, float ref_size, float sigma,
                             int group, dfv_stream_t stream);

/* Small helpers. */
int dfv_cast_weight(const float* src, void* dst, int dtype, int rows, int cols, int transpose, dfv_stream_t stream);
int dfv_dw_weight_pack(const float* src_ckk, float* dst_kkc, int C, int kernel, int flip, dfv_stream_t stream);
int dfv_dw_weight_unpack(const float* src_kkc, float* dst_ckk, int C, int kernel, dfv_stream_t stream);
int dfv_dropout_mask(float* out, long long n, float p, unsigned long long seed, dfv_stream_t stream);
/* The same with a DEVICE word added to the seed (may be NULL): launch arguments are frozen inside a captured CUDA graph, a
 * device word is not. */
int dfv_dropout_mask_dev(float* out, long long n, float p, unsigned long long seed, const unsigned long long* seed_dev,
                         dfv_stream_t stream);
int dfv_colsum(const float* a, int rows, int C, float* out, dfv_stream_t stream);
int dfv_add_mul(const float* a, const float* b, const float* mask, float* out, long long n, dfv_stream_t stream);
int dfv_convert(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, dfv_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Whole-path training step.  `params` / `grads` are host arrays of dfv_train_table_size() device
 * pointers to fp32 tensors in torch layout (the module's own parameter / buffer storage):
 *   block tensors   index = dfv_train_index(block 0..31, DFV_T_*)
 *   global tensors  index = dfv_train_index(-1, DFV_TG_*)
 *   classifier      index = dfv_train_cls_index(layer, 0 weight | 1 bias | 2 bn.weight | 3 bn.bias |
 *                                                      4 bn.running_mean | 5 bn.running_var)
 * Entries of tensors a configuration lacks are NULL.  dfv_train_fwd updates the BatchNorm running
 * statistics in place (num_batches_tracked is the host's to bump).  dfv_train_bwd ADDS every parameter
 * gradient into grads[] (the caller zeroes the buffer); entries for running statistics are ignored.
 * arena: saved activations, written by fwd and read by bwd; scratch: backward temporaries.
 * ---------------------------------------------------------------------------------- */
enum {
  DFV_T_EXPAND_W = 0, DFV_T_BN0_G, DFV_T_BN0_B, DFV_T_BN0_RM, DFV_T_BN0_RV,
  DFV_T_DW_W, DFV_T_BN1_G, DFV_T_BN1_B, DFV_T_BN1_RM, DFV_T_BN1_RV,
  DFV_T_SE_R_W, DFV_T_SE_R_B, DFV_T_SE_E_W, DFV_T_SE_E_B,
  DFV_T_PROJ_W, DFV_T_BN2_G, DFV_T_BN2_B, DFV_T_BN2_RM, DFV_T_BN2_RV,
  DFV_T_PER_BLOCK
};
enum {
  DFV_TG_STEM_W = 0, DFV_TG_STEM_G, DFV_TG_STEM_B, DFV_TG_STEM_RM, DFV_TG_STEM_RV,
  DFV_TG_HEAD_W, DFV_TG_HEAD_G, DFV_TG_HEAD_B, DFV_TG_HEAD_RM, DFV_TG_HEAD_RV,
  DFV_TG_LM_W, DFV_TG_SA_W, DFV_TG_CA_W1, DFV_TG_CA_W2,
  DFV_TG_COUNT
};
#define DFV_MAX_CLS_LAYERS 8
int dfv_train_table_size(void);
int dfv_train_index(int block, int kind);
int dfv_train_cls_index(int layer, int kind);

typedef struct {
  int32_t dtype;
  int32_t B, H, W;
  int32_t use_attention, use_landmark, use_channel, use_spatial;
  int32_t heat_group;              /* images per heat-map max group; 0 = whole batch */
  float landmark_ref_size;         /* 224.0 */
  float bn_eps, bn_momentum;       /* backbone BatchNorm2d: 1e-3, 0.01 (efficientnet-pytorch B4 global params) */
  float cls_bn_eps, cls_bn_momentum; /* classifier BatchNorm1d: torch defaults 1e-5, 0.1 */
  float drop_connect_rate;         /* 0.2; block i uses rate * i / 32 (EfficientNet.extract_features) */
  float feat_dropout, cls_dropout; /* feature_extractor.backbone.dropout p, classifier Dropout p */
  uint64_t seed;                   /* dropout / drop-connect masks are a function of (seed, position) */
  const float* const* params;
  float* const* grads;             /* bwd only */
  const float* images_nchw;        /* [B][3][H][W] fp32 */
  const float* landmarks;          /* [B][5][2] fp32 or NULL */
  int32_t ca_hidden;
  const int32_t* head_dims;        /* host int[head_layers + 1] */
  int32_t head_layers;
  void* arena;
  size_t arena_bytes;
  void* scratch;
  size_t scratch_bytes;
  float* logits;                   /* fwd out [B][num_classes] */
  float* features;                 /* fwd out [B][1792]: post-dropout pooled features (what CombinedLoss receives) */
  const float* dlogits;            /* bwd in  [B][num_classes] */
  const float* dfeatures;          /* bwd in  [B][1792] or NULL */
  void* const* taps;               /* optional debug: host array of 34 device pointers (stem, block0..31, head)
                                      receiving NHWC copies of each stage output (fwd) */
  int32_t freeze_bn;               /* 1: backbone BatchNorm2d layers run in EVAL mode inside the training step (running
                                      statistics, not updated; no gamma / beta gradients) -- EfficientNetB4Backbone
                                      freeze_bn=True (src/feature_extraction/efficientnet.py:84-90,165-170).  The
                                      classifier's BatchNorm1d layers stay in train mode, as in the reference. */
  void* const* grad_events;        /* bwd, optional: host array of DFV_GRAD_UNITS cudaEvent_t handles (entries may be
                                      NULL).  Event u is recorded on the stream once every gradient of unit u is
                                      complete: unit 0 = classifier + attention + head conv, unit 1 + j = block 31 - j,
                                      unit 33 = stem.  Lets the caller start the data-parallel all-reduce of a
                                      finished gradient bucket on a side stream while the rest of the backward runs. */
  const unsigned long long* seed_dev; /* optional device word added to `seed` by every mask kernel: lets a CUDA-graph replay of
                                      the step (GraphedTrainStep) draw fresh dropout / drop-connect masks */
} dfv_train_args;
#define DFV_GRAD_UNITS 34

size_t dfv_train_arena_bytes(int dtype, int B, int H, int W, const int32_t* head_dims, int head_layers, int ca_hidden);
size_t dfv_train_scratch_bytes(int dtype, int B, int H, int W, const int32_t* head_dims, int head_layers, int ca_hidden);
int dfv_train_fwd(const dfv_train_args* args, dfv_stream_t stream);
int dfv_train_bwd(const dfv_train_args* args, dfv_stream_t stream);

/* Training forward of the depthwise conv with the train-mode BatchNorm statistics fused in: raw conv output (no
 * activation) plus per-channel sum / sum of squares ADDED into stats[2C] (double; the caller zeroes it), finished by
 * dfv_bn_stats_from_sums (mean, 1/sqrt(var + eps), running statistics with momentum and the unbiased variance). */
int dfv_dwconv_stats_fwd(const void* x, const float* w_kkc, const float* bias, void* y, double* stats, int dtype, int B,
                         int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, dfv_stream_t stream);
int dfv_bn_stats_from_sums(const double* acc, int C, double count, float eps, float momentum, float* mean, float* invstd,
                           float* running_mean, float* running_var, dfv_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Operators either side of the hot path.
 * ---------------------------------------------------------------------------------- */

/* Video-frame scoring rule of the competition notebook (task.ipynb:434-442): the batch holds n_clips groups of
 * frames_per_clip consecutive frames; per clip: mean logit [n_clips][n_classes], mean softmax(logits)[:, 1]
 * (fake_prob) and label = fake_prob >= threshold.  All fp32 / int32 device buffers. */
int dfv_clip_aggregate(const float* logits, int n_clips, int frames_per_clip, int n_classes, float* mean_logits,
                       float* fake_prob, int* labels, float threshold, dfv_stream_t stream);

/* F.adaptive_avg_pool2d(x, 1).flatten(1) of an NHWC map -> fp32 [B][C]
 * (DeepfakeFeatureExtractor.extract_multi_scale_features, feature_extractor.py:119-154). */
int dfv_global_avg_pool(const void* x, int dtype, float* out, int B, long long rows_per_image, int C, dfv_stream_t stream);

/* F.normalize(x, p=2, dim=1) (DeepfakeFeatureExtractor.get_embedding, feature_extractor.py:156-178). */
int dfv_l2_normalize(const float* x, float* y, int B, int D, float eps, dfv_stream_t stream);

/* The optimizer step of the reference training loop (trainer.py:158-167; scripts/train.py:96-102):
 * torch.nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.AdamW, over FLAT fp32 buffers of n
 * elements (parameters, gradients, exp_avg, exp_avg_sq).  grad_scale multiplies the gradients first (1/N of a
 * summed all-reduce, or a loss-scale inverse); max_norm <= 0 disables clipping; step is 1-based.
 * sqnorm_ws: one double of device scratch; total_norm_out: optional device float (pre-clip global norm).
 * frozen_chunks: optional device bytes, one per 64-element chunk of the flat buffers (tensors start on 64-element
 * boundaries): nonzero = leave the chunk's parameters and moments untouched, as torch.optim.AdamW skips parameters
 * without a gradient (freeze_bn, requires_grad = False). */
int dfv_clip_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                        double* sqnorm_ws, double max_norm, double grad_scale, double lr, double beta1, double beta2,
                        double eps, double weight_decay, long long step, float* total_norm_out,
                        const unsigned char* frozen_chunks, dfv_stream_t stream);

/* Plan introspection (host only, pure functions of the shape): the depthwise tile plan chosen for a layer.
 * out[0..9] = CB, L, TW, TH, threads, smem bytes, tiles_w, tiles_h, pool parts, grid. */
int dfv_dwconv_plan_info(int dtype, int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, int* out);
/* Tile plan of a bf16 tensor-core 1x1-conv GEMM (host only): out[0..9] = N tile, weight-stationary flag, pipeline
 * stages, staging buffers, grid, tiles per CTA, shared-memory bytes, N tiles, CTAs per cluster (2 = CTA-pair plan,
 * cta_group::2), N tiles per A stage (2 = shared-A plan: both N tiles of an M tile accumulate from one A stage). */
int dfv_gemm_plan_info(long long M, int K, int N, int scaled, int* out);

/* Tuning entry points: the same operators with the tile plan restricted by the CALLER (a per-call argument: no
 * process state, no environment variables).  Fields left 0 (-1 for weight_stationary) are chosen by the planner.
 * Used by the kernel-sweep scripts and by the tests that exercise every plan family at full size. */
typedef struct {
  int32_t weight_stationary;   /* -1 auto, 0 streaming, 1 weight-stationary */
  int32_t bn;                  /* N tile (0 = auto) */
  int32_t cluster;             /* streaming plans: 2 = CTA pair (cta_group::2: M = 256 over two SMs, half of the weight tile per SM), -1 = single CTA, 0 = auto */
  int32_t share_a;             /* gated CTA-pair plans with two N tiles: 1 = both N tiles share each A stage (two accumulators side by side in TMEM), -1 = off, 0 = auto */
  int32_t rotate;              /* streaming plans: -1 = every CTA walks the N tiles in the same order (off), 0 / 1 = order rotated by the cluster index */
} dfv_gemm_tuning;
int dfv_pw_gemm_fwd_tuned(const void* a, const void* w, const float* bias, const void* a_scale, int rows_per_image,
                          const void* residual, void* out, int dtype, long long M, int K, int N, int act,
                          const dfv_gemm_tuning* tuning, dfv_stream_t stream);
typedef struct {
  int32_t L, TW, TH, CB;       /* strip length, tile width / height, channel chunk (0 = auto); stride-1 layers only */
} dfv_dwconv_tuning;
int dfv_dwconv_fwd_tuned(const void* x, const float* w_kkc, const float* bias, void* y, float* pool_partial, int dtype,
                         int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi, int act,
                         const dfv_dwconv_tuning* tuning, dfv_stream_t stream);
int dfv_dwconv_pool_parts_tuned(int dtype, int B, int H, int W, int C, int kernel, int stride, int pad_lo, int pad_hi,
                                const dfv_dwconv_tuning* tuning);

/* ------------------------------------------------------------------------------------
 * Weight packing for inference, from the module's own parameter storage: `params` is the same host table of
 * device pointers to fp32 torch-layout tensors as dfv_train_args.params (dfv_train_index / dfv_train_cls_index,
 * running statistics included).  Folds eval-mode BatchNorm (w' = w * gamma / sqrt(var + eps),
 * b' = beta - mean * scale [+ bias * scale]) and writes
 *   blob        the dfv_blob_* slots (GEMM weights in `dtype`, the rest fp32)
 *   head_w_t[l] fp32 [dims[l]][dims[l+1]]  transposed folded classifier weights, head_b[l] fp32 [dims[l+1]]
 *   ca_w2_t     fp32 [ca_hidden][1792]     channel_attn.fc.2.weight transposed (NULL: no channel attention)
 * in a handful of launches -- no arithmetic of the path is left to the host framework.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int32_t dtype;
  float bn_eps;                    /* backbone BatchNorm2d eps (1e-3) */
  float cls_bn_eps;                /* classifier BatchNorm1d eps (1e-5) */
  const float* const* params;      /* host table, dfv_train_table_size() entries */
  void* blob;                      /* dfv_blob_bytes(dtype) bytes */
  int32_t head_layers;
  const int32_t* head_dims;        /* host int[head_layers + 1] */
  float* const* head_w_t;          /* host array of device pointers (out) */
  float* const* head_b;
  int32_t ca_hidden;
  float* ca_w2_t;
} dfv_pack_args;
int dfv_pack_weights(const dfv_pack_args* args, dfv_stream_t stream);


#ifdef __cplusplus
}
#endif
#endif /* DFVIT_H_ */
