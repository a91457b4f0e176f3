// Whole-path inference sequencer: DeepfakeDetectionModel.forward in eval mode as one native
// call (stem -> 32 MBConv blocks -> head conv -> heat-map -> HybridAttention+pool -> MLP head).
// Pure host code: it owns no memory, only carves the caller's workspace and enqueues kernels.
#include <vector>

#include "common.cuh"

namespace dfv {

struct Shapes {
  int Hs, Ws;                 // after the stem
  int Hin[64], Win[64];       // block input spatial size
  int Hout[64], Wout[64];
  int Hf, Wf;
  size_t act_elems;           // per image: max stage output (ping-pong buffers)
  size_t exp_elems;           // per image: max expanded tensor
  size_t dw_elems;            // per image: max depthwise output
  size_t pool_floats;         // per image: max parts * c_mid
  int max_cmid;
  size_t se_scratch_floats;   // whole batch: max over blocks of dfv_se_scratch_floats
  int max_sq;
  bool se_fused[64];          // block runs the squeeze layer inside its depthwise kernel
};

static int compute_shapes(Shapes* s, int dtype, int B, int H, int W) {
  int n;
  const dfv_block_info* blk = topo_blocks(&n);
  s->Hs = (H + 1 - 3) / 2 + 1;
  s->Ws = (W + 1 - 3) / 2 + 1;
  int h = s->Hs, w = s->Ws;
  s->act_elems = (size_t)h * w * topo_stem_c();
  s->exp_elems = s->dw_elems = s->pool_floats = 0;
  s->max_cmid = 0;
  s->se_scratch_floats = 0;
  s->max_sq = 1;
  for (int i = 0; i < n; ++i) {
    const dfv_block_info& b = blk[i];
    s->Hin[i] = h;
    s->Win[i] = w;
    const int ho = (h + b.pad_lo + b.pad_hi - b.kernel) / b.stride + 1;
    const int wo = (w + b.pad_lo + b.pad_hi - b.kernel) / b.stride + 1;
    if (ho <= 0 || wo <= 0) return DFV_ERR_INVALID;
    s->Hout[i] = ho;
    s->Wout[i] = wo;
    if (b.has_expand) s->exp_elems = std::max(s->exp_elems, (size_t)h * w * b.c_mid);
    s->dw_elems = std::max(s->dw_elems, (size_t)ho * wo * b.c_mid);
    const int parts = dfv_dwconv_pool_parts(dtype, B, h, w, b.c_mid, b.kernel, b.stride, b.pad_lo, b.pad_hi);
    if (parts <= 0) return DFV_ERR_INVALID;
    s->pool_floats = std::max(s->pool_floats, (size_t)parts * b.c_mid);
    s->act_elems = std::max(s->act_elems, (size_t)ho * wo * b.c_out);
    s->max_cmid = std::max(s->max_cmid, b.c_mid);
    s->se_scratch_floats = std::max(s->se_scratch_floats, dfv_se_scratch_floats(B, b.c_mid, b.se_squeeze));
    s->se_fused[i] = dfv_dwconv_se_profitable(dtype, B, h, w, b.c_mid, b.kernel, b.stride, b.pad_lo, b.pad_hi, b.se_squeeze) != 0;
    s->max_sq = std::max(s->max_sq, b.se_squeeze);
    h = ho;
    w = wo;
  }
  s->Hf = h;
  s->Wf = w;
  s->act_elems = std::max(s->act_elems, (size_t)h * w * topo_head_c());
  return DFV_OK;
}

struct Workspace {
  char* act[2];
  char* expand;
  char* dw;
  float* pool;
  void* gate;
  float* se_scratch;       // squeeze partial sums between the two SE kernels
  long long* se_hid[2];    // fused squeeze: fixed-point hidden accumulators [B][squeeze], two buffers alternating by layer
  float* heat;
  float* heat_raw;
  uint32_t* heat_max;
  float* attn_scratch;     // HybridAttention scratch (dfv_attention_scratch_floats at hidden = kAttnHiddenMax)
  float* head_scratch;     // classifier scratch: kHeadScratchPerRow floats per image (dfv_mlp_head_scratch_floats must fit)
  char* fold_ws;           // scratch of dfv_pw_conv_fwd (row-folded thin 1x1 convolutions)
  size_t bytes;
};
constexpr int kHeadHiddenMax = 2048;
constexpr int kAttnHiddenMax = 288;    // channel-attention MLP width the workspace (and the small-linear kernels' shared memory) allow
constexpr size_t kHeadScratchPerRow = (size_t)10 * kHeadHiddenMax;   // two activation buffers + K-slice partial sums


static void carve(Workspace* ws, char* base, const Shapes& s, int dtype, int B) {
  const size_t es = dtype_size(dtype);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  ws->act[0] = take((size_t)B * s.act_elems * es);
  ws->act[1] = take((size_t)B * s.act_elems * es);
  ws->expand = take((size_t)B * s.exp_elems * es);
  ws->dw = take((size_t)B * s.dw_elems * es);
  ws->pool = (float*)take((size_t)B * s.pool_floats * 4);
  ws->gate = take((size_t)B * s.max_cmid * 4);
  ws->se_scratch = (float*)take(s.se_scratch_floats * 4);
  ws->se_hid[0] = (long long*)take((size_t)B * s.max_sq * 8);
  ws->se_hid[1] = (long long*)take((size_t)B * s.max_sq * 8);
  ws->heat = (float*)take((size_t)B * s.Hf * s.Wf * 4);
  ws->heat_raw = (float*)take((size_t)B * s.Hf * s.Wf * 4);
  ws->heat_max = (uint32_t*)take((size_t)B * 4);
  ws->attn_scratch = (float*)take(dfv_attention_scratch_floats(B, s.Hf, s.Wf, topo_head_c(), kAttnHiddenMax) * 4);
  ws->head_scratch = (float*)take((size_t)B * kHeadScratchPerRow * 4);
  ws->fold_ws = take(dfv_pw_fold_ws_bytes(B));
  ws->bytes = off;
}

}  // namespace dfv

using namespace dfv;

extern "C" size_t dfv_infer_workspace_bytes(int dtype, int B, int H, int W) {
  Shapes s;
  if (!valid_dtype(dtype) || B <= 0 || H < 32 || W < 32 || compute_shapes(&s, dtype, B, H, W) != DFV_OK) {
    set_error("dfv_infer_workspace_bytes: bad arguments");
    return 0;
  }
  Workspace ws;
  carve(&ws, nullptr, s, dtype, B);
  return ws.bytes;
}

extern "C" int dfv_infer_fwd(const dfv_infer_args* a, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_REQUIRE(a != nullptr, "dfv_infer_fwd: null args");
  DFV_REQUIRE(valid_dtype(a->dtype), "dfv_infer_fwd: bad dtype %d", a->dtype);
  DFV_REQUIRE(a->B > 0 && a->H >= 32 && a->W >= 32, "dfv_infer_fwd: bad shape B=%d H=%d W=%d", a->B, a->H, a->W);
  DFV_REQUIRE(a->blob && (a->images_nchw || a->images_u8) && a->workspace && a->logits && a->features, "dfv_infer_fwd: null pointer");
  DFV_REQUIRE(a->head_w_t && a->head_b && a->head_dims && a->head_layers >= 1, "dfv_infer_fwd: classifier head missing");
  const int dtype = a->dtype, B = a->B;
  Shapes s;
  DFV_REQUIRE(compute_shapes(&s, dtype, a->B, a->H, a->W) == DFV_OK, "dfv_infer_fwd: input %dx%d too small for the backbone", a->H, a->W);
  Workspace ws;
  carve(&ws, (char*)a->workspace, s, dtype, B);
  if (ws.bytes > a->workspace_bytes) {
    set_error("dfv_infer_fwd: workspace too small (%zu < %zu bytes)", a->workspace_bytes, ws.bytes);
    return DFV_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const size_t es = dtype_size(dtype);
  int n;
  const dfv_block_info* blk = topo_blocks(&n);
  const int stem_c = topo_stem_c(), head_c = topo_head_c();
  auto W_ = [&](int block, int kind) { return blob_ptr(a->blob, dtype, block, kind); };
  auto tap = [&](int idx, const void* src, size_t elems) -> int {
    if (a->taps && a->taps[idx]) DFV_CUDA(cudaMemcpyAsync(a->taps[idx], src, elems * es, cudaMemcpyDeviceToDevice, st));
    return DFV_OK;
  };

  // fixed-point accumulators of the fused squeeze: the first layer's buffer is zeroed here, every later one by the
  // depthwise kernel of the layer before it
  DFV_CUDA(cudaMemsetAsync(ws.se_hid[0], 0, sizeof(long long) * (size_t)B * s.max_sq, st));
  int se_buf = 0;
  int cur = 0;
  if (a->images_u8)     // raw uint8 HWC crops: the reference's input normalisation runs inside the stem's operand load
    DFV_TRY(dfv_stem_conv_u8_fwd(a->images_u8, a->u8_norm, (const float*)W_(-1, DFV_W_STEM), (const float*)W_(-1, DFV_W_STEM_BIAS),
                                 ws.act[cur], dtype, B, a->H, a->W, stem_c, DFV_ACT_SILU, stream));
  else
    DFV_TRY(dfv_stem_conv_fwd(a->images_nchw, (const float*)W_(-1, DFV_W_STEM), (const float*)W_(-1, DFV_W_STEM_BIAS),
                              ws.act[cur], dtype, B, a->H, a->W, stem_c, DFV_ACT_SILU, stream));
  DFV_TRY(tap(0, ws.act[cur], (size_t)B * s.Hs * s.Ws * stem_c));

  for (int i = 0; i < n; ++i) {
    const dfv_block_info& b = blk[i];
    const int h = s.Hin[i], w = s.Win[i], ho = s.Hout[i], wo = s.Wout[i];
    const void* x = ws.act[cur];
    const void* dw_in = x;
    if (b.has_expand) {
      DFV_TRY(dfv_pw_conv_fwd(x, W_(i, DFV_W_EXPAND), (const float*)W_(i, DFV_W_EXPAND_BIAS), nullptr, h * w, nullptr, ws.expand,
                              dtype, B, (long long)B * h * w, b.c_in, b.c_mid, DFV_ACT_SILU, ws.fold_ws, stream));
      dw_in = ws.expand;
    }
    const int parts = dfv_dwconv_pool_parts(dtype, B, h, w, b.c_mid, b.kernel, b.stride, b.pad_lo, b.pad_hi);
    if (s.se_fused[i]) {      // squeeze layer in the depthwise kernel's tail; one launch left for the gate
      DFV_TRY(dfv_dwconv_se_fwd(dw_in, (const float*)W_(i, DFV_W_DW), (const float*)W_(i, DFV_W_DW_BIAS), ws.dw, ws.pool,
                                (const float*)W_(i, DFV_W_SE_REDUCE), ws.se_hid[se_buf], ws.se_hid[se_buf ^ 1], (long long)B * s.max_sq,
                                b.se_squeeze, dtype, B, h, w, b.c_mid, b.kernel, b.stride, b.pad_lo, b.pad_hi, stream));
      DFV_TRY(dfv_se_excite_fwd(ws.se_hid[se_buf], (const float*)W_(i, DFV_W_SE_REDUCE_BIAS), (const float*)W_(i, DFV_W_SE_EXPAND),
                                (const float*)W_(i, DFV_W_SE_EXPAND_BIAS), ws.gate, dtype, B, b.c_mid, b.se_squeeze, stream));
      se_buf ^= 1;
    } else {
      DFV_TRY(dfv_dwconv_fwd(dw_in, (const float*)W_(i, DFV_W_DW), (const float*)W_(i, DFV_W_DW_BIAS), ws.dw, ws.pool, dtype, B, h,
                             w, b.c_mid, b.kernel, b.stride, b.pad_lo, b.pad_hi, DFV_ACT_SILU, stream));
      DFV_TRY(dfv_se_gate_fwd(ws.pool, parts, 1.0f / (float)(ho * wo), (const float*)W_(i, DFV_W_SE_REDUCE),
                              (const float*)W_(i, DFV_W_SE_REDUCE_BIAS), (const float*)W_(i, DFV_W_SE_EXPAND),
                              (const float*)W_(i, DFV_W_SE_EXPAND_BIAS), ws.gate, dtype, ws.se_scratch, B, b.c_mid, b.se_squeeze, stream));
    }
    DFV_TRY(dfv_pw_conv_fwd(ws.dw, W_(i, DFV_W_PROJECT), (const float*)W_(i, DFV_W_PROJECT_BIAS), ws.gate, ho * wo,
                            b.has_skip ? x : nullptr, ws.act[cur ^ 1], dtype, B, (long long)B * ho * wo, b.c_mid, b.c_out,
                            DFV_ACT_NONE, ws.fold_ws, stream));
    cur ^= 1;
    DFV_TRY(tap(1 + i, ws.act[cur], (size_t)B * ho * wo * b.c_out));
  }

  DFV_TRY(dfv_pw_gemm_fwd(ws.act[cur], W_(-1, DFV_W_HEAD), (const float*)W_(-1, DFV_W_HEAD_BIAS), nullptr, 0, nullptr,
                          ws.act[cur ^ 1], dtype, (long long)B * s.Hf * s.Wf, blk[n - 1].c_out, head_c, DFV_ACT_SILU, stream));
  cur ^= 1;
  DFV_TRY(tap(1 + n, ws.act[cur], (size_t)B * s.Hf * s.Wf * head_c));

  const float* heat = nullptr;
  if (a->use_attention && a->use_landmark && a->landmarks != nullptr) {
    DFV_REQUIRE(a->lm_weights, "dfv_infer_fwd: landmark attention weights missing");
    DFV_TRY(dfv_landmark_heatmap_fwd_ex(a->landmarks, a->lm_weights, ws.heat, ws.heat_raw, ws.heat_max, nullptr, B, s.Hf, s.Wf,
                                        a->landmark_ref_size > 0.f ? a->landmark_ref_size : 224.0f, 1.5f, a->heat_group, a->heat_max_floor,
                                        stream));
    heat = ws.heat;
    if (a->heat) DFV_CUDA(cudaMemcpyAsync(a->heat, ws.heat, sizeof(float) * (size_t)B * s.Hf * s.Wf, cudaMemcpyDeviceToDevice, st));
  }
  const int use_c = a->use_attention && a->use_channel, use_s = a->use_attention && a->use_spatial;
  DFV_REQUIRE(!use_c || a->ca_hidden <= kAttnHiddenMax, "dfv_infer_fwd: channel-attention hidden width %d > %d", a->ca_hidden, kAttnHiddenMax);
  DFV_TRY(dfv_hybrid_attention_fwd(ws.act[cur], heat, a->ca_w1, a->ca_w2_t, a->sa_w, a->features, nullptr, nullptr, ws.attn_scratch, dtype, B,
                                   s.Hf, s.Wf, head_c, use_c ? a->ca_hidden : 0, use_c, use_s, stream));
  for (int l = 1; l < a->head_layers; ++l)
    DFV_REQUIRE(a->head_dims[l] <= kHeadHiddenMax, "dfv_infer_fwd: classifier hidden width %d > %d", a->head_dims[l], kHeadHiddenMax);
  DFV_REQUIRE(dfv_mlp_head_scratch_floats(a->head_dims, a->head_layers, B) <= (size_t)B * kHeadScratchPerRow,
              "dfv_infer_fwd: classifier head too wide for the workspace");
  DFV_TRY(dfv_mlp_head_fwd(a->features, a->head_w_t, a->head_b, a->head_dims, a->head_layers, a->logits, ws.head_scratch, B, stream));
  return DFV_OK;
}
