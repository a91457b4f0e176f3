// Whole-path training sequencer: DeepfakeDetectionModel.forward in train mode and its backward
// (feature_extractor.py:242-269 under autograd; trainer.py:140-153).  Pure host code: carves the
// caller's arena / scratch and enqueues kernels on the caller's stream.
//
// Per convolution:   raw = conv(x)  ->  bn_stats(raw)  ->  act = bn_act(raw)      (forward)
//                    du  = act_bn_bwd(dy)  ->  d raw = bn_bwd_apply(du)  ->  wgrad, dgrad   (backward)
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace dfv {

namespace {

struct TShapes {
  int Hs, Ws;
  int Hin[64], Win[64], Hout[64], Wout[64];
  int Hf, Wf;
};

int train_shapes(TShapes* s, int H, int W) {
  int n;
  const dfv_block_info* blk = topo_blocks(&n);
  s->Hs = (H + 1 - 3) / 2 + 1;
  s->Ws = (W + 1 - 3) / 2 + 1;
  int h = s->Hs, w = s->Ws;
  for (int i = 0; i < n; ++i) {
    const dfv_block_info& b = blk[i];
    s->Hin[i] = h;
    s->Win[i] = w;
    const int ho = (h + b.pad_lo + b.pad_hi - b.kernel) / b.stride + 1;
    const int wo = (w + b.pad_lo + b.pad_hi - b.kernel) / b.stride + 1;
    if (ho <= 0 || wo <= 0) return DFV_ERR_INVALID;
    s->Hout[i] = ho;
    s->Wout[i] = wo;
    h = ho;
    w = wo;
  }
  s->Hf = h;
  s->Wf = w;
  return DFV_OK;
}

struct BlockArena {
  char *e_raw, *e, *d_raw, *d, *p_raw, *out;
  float *m0, *i0, *m1, *i1, *m2, *i2;
  float *pooled, *h1, *gate_f32, *dc;
  char* gate;
  char *wE, *wEt, *wP, *wPt;
  float *wD, *wDf;
};

struct ClsArena {
  float *lin, *mean, *invstd, *act, *mask;
};

struct Arena {
  float* zero_bias;
  float* stem_w;
  char *s_raw, *a0;
  float *sm, *si;
  BlockArena blk[64];
  char *h_raw, *h;
  float *hm, *hi;
  char *wH, *wHt;
  float *heat, *heat_raw;
  uint32_t* heat_max;
  float* attn_saved;
  float* feat_pre;
  float* mask_f;
  float* pool_ws;
  float* se_scratch;   // squeeze partial sums between the two SE kernels
  float* bn_ws;
  double* bn_acc;      // per-channel sum / sum of squares a producer kernel accumulates (fused BatchNorm statistics)
  char* fold_ws;
  ClsArena cls[DFV_MAX_CLS_LAYERS];
  size_t bytes;
};

struct Scratch {
  char* gout[2];
  char *gP, *gA, *gE;
  float *bn_ws, *coef, *se_ws, *dpool, *dwkkc, *attn_ws, *dheat;
  float* gcls[2];
  float* wT;
  float* dfeat;
  char* fold_ws;
  size_t bytes;
};

constexpr int kZeroBias = 4096;

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  char* take(size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes ? bytes : 1, 1024);
    return p;
  }
  float* takef(size_t n) { return reinterpret_cast<float*>(take(n * 4)); }
};

size_t max_bn_ws(const TShapes& s, int B, const int32_t* dims, int layers) {
  int n;
  const dfv_block_info* blk = topo_blocks(&n);
  size_t m = dfv_bn_ws_floats(B, (long long)s.Hs * s.Ws, topo_stem_c());
  for (int i = 0; i < n; ++i) {
    m = std::max(m, dfv_bn_ws_floats(B, (long long)s.Hin[i] * s.Win[i], blk[i].c_mid));
    m = std::max(m, dfv_bn_ws_floats(B, (long long)s.Hout[i] * s.Wout[i], blk[i].c_mid));
    m = std::max(m, dfv_bn_ws_floats(B, (long long)s.Hout[i] * s.Wout[i], blk[i].c_out));
  }
  m = std::max(m, dfv_bn_ws_floats(B, (long long)s.Hf * s.Wf, topo_head_c()));
  for (int l = 1; l <= layers; ++l) m = std::max(m, dfv_bn_ws_floats(B, 1, (dims[l] + 7) / 8 * 8));
  return m;
}

void carve_arena(Arena* a, void* base, const TShapes& s, int dtype, int B, const int32_t* dims, int layers, int hidden) {
  const size_t es = dtype_size(dtype);
  int n;
  const dfv_block_info* blk = topo_blocks(&n);
  Carver c(base);
  a->zero_bias = c.takef(kZeroBias);
  a->stem_w = c.takef(27 * (size_t)topo_stem_c());
  const size_t stem_elems = (size_t)B * s.Hs * s.Ws * topo_stem_c();
  a->s_raw = c.take(stem_elems * es);
  a->a0 = c.take(stem_elems * es);
  a->sm = c.takef(topo_stem_c());
  a->si = c.takef(topo_stem_c());
  size_t pool_max = 0, se_max = 0;
  for (int i = 0; i < n; ++i) {
    const dfv_block_info& b = blk[i];
    BlockArena& ba = a->blk[i];
    const size_t min_ = (size_t)B * s.Hin[i] * s.Win[i], mout = (size_t)B * s.Hout[i] * s.Wout[i];
    if (b.has_expand) {
      ba.e_raw = c.take(min_ * b.c_mid * es);
      ba.e = c.take(min_ * b.c_mid * es);
      ba.m0 = c.takef(b.c_mid);
      ba.i0 = c.takef(b.c_mid);
      ba.wE = c.take((size_t)b.c_mid * b.c_in * es);
      ba.wEt = c.take((size_t)b.c_mid * b.c_in * es);
    } else {
      ba.e_raw = ba.e = ba.wE = ba.wEt = nullptr;
      ba.m0 = ba.i0 = nullptr;
    }
    ba.d_raw = c.take(mout * b.c_mid * es);
    ba.d = c.take(mout * b.c_mid * es);
    ba.p_raw = c.take(mout * b.c_out * es);
    ba.out = c.take(mout * b.c_out * es);
    ba.m1 = c.takef(b.c_mid);
    ba.i1 = c.takef(b.c_mid);
    ba.m2 = c.takef(b.c_out);
    ba.i2 = c.takef(b.c_out);
    ba.pooled = c.takef((size_t)B * b.c_mid);
    ba.h1 = c.takef((size_t)B * b.se_squeeze);
    ba.gate_f32 = c.takef((size_t)B * b.c_mid);
    ba.gate = c.take((size_t)B * b.c_mid * es);
    ba.dc = c.takef(B);
    ba.wP = c.take((size_t)b.c_out * b.c_mid * es);
    ba.wPt = c.take((size_t)b.c_out * b.c_mid * es);
    ba.wD = c.takef((size_t)b.kernel * b.kernel * b.c_mid);
    ba.wDf = c.takef((size_t)b.kernel * b.kernel * b.c_mid);
    pool_max = std::max(pool_max, (size_t)B * dfv_rows_chunks(B, (long long)s.Hout[i] * s.Wout[i]) * b.c_mid);
    se_max = std::max(se_max, dfv_se_scratch_floats(B, b.c_mid, b.se_squeeze));
  }
  const size_t head_elems = (size_t)B * s.Hf * s.Wf * topo_head_c();
  a->h_raw = c.take(head_elems * es);
  a->h = c.take(head_elems * es);
  a->hm = c.takef(topo_head_c());
  a->hi = c.takef(topo_head_c());
  a->wH = c.take((size_t)topo_head_c() * blk[n - 1].c_out * es);
  a->wHt = c.take((size_t)topo_head_c() * blk[n - 1].c_out * es);
  a->heat = c.takef((size_t)B * s.Hf * s.Wf);
  a->heat_raw = c.takef((size_t)B * s.Hf * s.Wf);
  a->heat_max = reinterpret_cast<uint32_t*>(c.takef(B));
  a->attn_saved = c.takef(dfv_attention_saved_floats(B, s.Hf, s.Wf, topo_head_c(), hidden));
  a->feat_pre = c.takef((size_t)B * topo_head_c());
  a->mask_f = c.takef((size_t)B * topo_head_c());
  a->pool_ws = c.takef(pool_max);
  a->se_scratch = c.takef(se_max);
  a->bn_ws = c.takef(max_bn_ws(s, B, dims, layers));
  a->fold_ws = c.take(dfv_pw_fold_ws_bytes(B));
  a->bn_acc = reinterpret_cast<double*>(c.take(sizeof(double) * 2 * 4096));
  for (int l = 0; l < layers && l < DFV_MAX_CLS_LAYERS; ++l) {
    const size_t d = dims[l + 1];
    a->cls[l].lin = c.takef((size_t)B * d);
    a->cls[l].mean = c.takef(d);
    a->cls[l].invstd = c.takef(d);
    a->cls[l].act = c.takef((size_t)B * d);
    a->cls[l].mask = c.takef((size_t)B * d);
  }
  a->bytes = c.off;
}

void carve_scratch(Scratch* sc, void* base, const TShapes& s, int dtype, int B, const int32_t* dims, int layers, int hidden) {
  const size_t es = dtype_size(dtype);
  int n;
  const dfv_block_info* blk = topo_blocks(&n);
  size_t io = (size_t)s.Hs * s.Ws * topo_stem_c(), pmax = 0, amax = 0, emax = (size_t)s.Hf * s.Wf * topo_head_c();
  size_t se_ws = 0, cmax = topo_head_c(), kkc = 0;
  for (int i = 0; i < n; ++i) {
    const dfv_block_info& b = blk[i];
    const size_t hw_in = (size_t)s.Hin[i] * s.Win[i], hw_out = (size_t)s.Hout[i] * s.Wout[i];
    io = std::max(io, std::max(hw_in * b.c_in, hw_out * b.c_out));
    pmax = std::max(pmax, hw_out * b.c_out);
    amax = std::max(amax, hw_out * b.c_mid);
    emax = std::max(emax, hw_in * b.c_mid);
    se_ws = std::max(se_ws, dfv_se_bwd_ws_floats(B, (long long)hw_out, b.c_mid, b.se_squeeze));
    cmax = std::max<size_t>(cmax, b.c_mid);
    kkc = std::max(kkc, (size_t)b.kernel * b.kernel * b.c_mid);
  }
  size_t dmax = 0, wmax = 0;
  for (int l = 0; l <= layers; ++l) dmax = std::max<size_t>(dmax, dims[l]);
  for (int l = 0; l < layers; ++l) wmax = std::max(wmax, (size_t)dims[l] * dims[l + 1]);
  Carver c(base);
  sc->gout[0] = c.take((size_t)B * io * es);
  sc->gout[1] = c.take((size_t)B * io * es);
  sc->gP = c.take((size_t)B * pmax * es);
  sc->gA = c.take((size_t)B * amax * es);
  sc->gE = c.take((size_t)B * emax * es);
  sc->bn_ws = c.takef(max_bn_ws(s, B, dims, layers));
  sc->coef = c.takef(2 * cmax);
  sc->se_ws = c.takef(se_ws);
  sc->dpool = c.takef((size_t)B * cmax);
  sc->dwkkc = c.takef(kkc);
  sc->attn_ws = c.takef((size_t)B * topo_head_c() + (size_t)B * 2 * std::max(hidden, 1));
  sc->dheat = c.takef((size_t)B * s.Hf * s.Wf);
  sc->gcls[0] = c.takef((size_t)B * dmax);
  sc->gcls[1] = c.takef((size_t)B * dmax);
  sc->wT = c.takef(wmax);
  sc->dfeat = c.takef((size_t)B * topo_head_c());
  sc->fold_ws = c.take(dfv_pw_fold_ws_bytes(B));
  sc->bytes = c.off;
}

int validate(const dfv_train_args* a, bool bwd) {
  DFV_REQUIRE(a != nullptr, "dfv_train: null args");
  DFV_REQUIRE(valid_dtype(a->dtype), "dfv_train: bad dtype %d", a->dtype);
  DFV_REQUIRE(a->B > 1 && a->H >= 32 && a->W >= 32, "dfv_train: bad shape B=%d H=%d W=%d (train-mode BatchNorm needs B > 1)", a->B, a->H, a->W);
  DFV_REQUIRE(a->params && a->images_nchw && a->arena && a->logits && a->features, "dfv_train: null pointer");
  DFV_REQUIRE(a->head_dims && a->head_layers >= 1 && a->head_layers <= DFV_MAX_CLS_LAYERS, "dfv_train: classifier head missing");
  DFV_REQUIRE(a->head_dims[0] == topo_head_c(), "dfv_train: classifier input must be %d wide", topo_head_c());
  for (int l = 1; l < a->head_layers; ++l)
    DFV_REQUIRE(a->head_dims[l] % 8 == 0, "dfv_train: hidden classifier widths must be multiples of 8 (got %d)", a->head_dims[l]);
  if (bwd) DFV_REQUIRE(a->grads && a->scratch && a->dlogits, "dfv_train_bwd: null pointer");
  return DFV_OK;
}

}  // namespace

// torch [48][3][3][3] -> stem kernel layout [kh][kw][ci][co]
__global__ void stem_weight_pack_kernel(const float* __restrict__ src, float* __restrict__ dst, int CO) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 27 * CO) return;
  const int co = i % CO, t = i / CO;           // t = (kh * 3 + kw) * 3 + ci
  const int ci = t % 3, kw = (t / 3) % 3, kh = t / 9;
  dst[i] = src[(size_t)co * 27 + ci * 9 + kh * 3 + kw];
}

// freeze_bn: eval-mode BatchNorm inside the training step -- "statistics" are the running ones, nothing is updated
__global__ void bn_frozen_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, float eps, float* __restrict__ mean,
                                       float* __restrict__ invstd, int C) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = rm[c];
  invstd[c] = 1.0f / sqrtf(rv[c] + eps);
}

static int bn_frozen_stats(const float* rm, const float* rv, float eps, float* mean, float* invstd, int C, cudaStream_t st) {
  DFV_REQUIRE(rm && rv, "dfv_train_fwd: freeze_bn needs the running statistics");
  DFV_PDL((bn_frozen_stats_kernel), (C + 127) / 128, 128, 0, st, rm, rv, eps, mean, invstd, C);
  DFV_LAUNCH_CHECK();
  return DFV_OK;
}

}  // namespace dfv

using namespace dfv;

extern "C" {

int dfv_train_table_size(void) { return 32 * DFV_T_PER_BLOCK + DFV_TG_COUNT + 6 * DFV_MAX_CLS_LAYERS; }
int dfv_train_index(int block, int kind) {
  if (block >= 0 && block < 32 && kind >= 0 && kind < DFV_T_PER_BLOCK) return block * DFV_T_PER_BLOCK + kind;
  if (block == -1 && kind >= 0 && kind < DFV_TG_COUNT) return 32 * DFV_T_PER_BLOCK + kind;
  set_error("dfv_train_index: bad (block %d, kind %d)", block, kind);
  return DFV_ERR_INVALID;
}
int dfv_train_cls_index(int layer, int kind) {
  if (layer >= 0 && layer < DFV_MAX_CLS_LAYERS && kind >= 0 && kind < 6) return 32 * DFV_T_PER_BLOCK + DFV_TG_COUNT + layer * 6 + kind;
  set_error("dfv_train_cls_index: bad (layer %d, kind %d)", layer, kind);
  return DFV_ERR_INVALID;
}

size_t dfv_train_arena_bytes(int dtype, int B, int H, int W, const int32_t* head_dims, int head_layers, int ca_hidden) {
  TShapes s;
  if (!valid_dtype(dtype) || B <= 0 || H < 32 || W < 32 || !head_dims || head_layers < 1 || head_layers > DFV_MAX_CLS_LAYERS ||
      train_shapes(&s, H, W) != DFV_OK) {
    set_error("dfv_train_arena_bytes: bad arguments");
    return 0;
  }
  Arena a;
  carve_arena(&a, nullptr, s, dtype, B, head_dims, head_layers, ca_hidden);
  return a.bytes;
}

size_t dfv_train_scratch_bytes(int dtype, int B, int H, int W, const int32_t* head_dims, int head_layers, int ca_hidden) {
  TShapes s;
  if (!valid_dtype(dtype) || B <= 0 || H < 32 || W < 32 || !head_dims || head_layers < 1 || head_layers > DFV_MAX_CLS_LAYERS ||
      train_shapes(&s, H, W) != DFV_OK) {
    set_error("dfv_train_scratch_bytes: bad arguments");
    return 0;
  }
  Scratch sc;
  carve_scratch(&sc, nullptr, s, dtype, B, head_dims, head_layers, ca_hidden);
  return sc.bytes;
}

int dfv_train_fwd(const dfv_train_args* a, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_TRY(validate(a, false));
  const int dtype = a->dtype, B = a->B;
  TShapes s;
  DFV_REQUIRE(train_shapes(&s, a->H, a->W) == DFV_OK, "dfv_train_fwd: input %dx%d too small for the backbone", a->H, a->W);
  Arena ar;
  carve_arena(&ar, a->arena, s, dtype, B, a->head_dims, a->head_layers, a->ca_hidden);
  if (ar.bytes > a->arena_bytes) {
    set_error("dfv_train_fwd: arena too small (%zu < %zu bytes)", a->arena_bytes, ar.bytes);
    return DFV_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const size_t es = dtype_size(dtype);
  int n;
  const dfv_block_info* blk = topo_blocks(&n);
  const int stem_c = topo_stem_c(), head_c = topo_head_c();
  auto P = [&](int block, int kind) -> const float* { return a->params[dfv_train_index(block, kind)]; };
  auto PW = [&](int block, int kind) -> float* { return const_cast<float*>(a->params[dfv_train_index(block, kind)]); };
  auto tap = [&](int idx, const void* src, size_t elems) -> int {
    if (a->taps && a->taps[idx]) DFV_CUDA(cudaMemcpyAsync(a->taps[idx], src, elems * es, cudaMemcpyDeviceToDevice, st));
    return DFV_OK;
  };
  const float eps = a->bn_eps, mom = a->bn_momentum;

  DFV_CUDA(cudaMemsetAsync(ar.zero_bias, 0, kZeroBias * sizeof(float), st));

  // ---- stem
  DFV_REQUIRE(P(-1, DFV_TG_STEM_W) && P(-1, DFV_TG_STEM_G) && P(-1, DFV_TG_STEM_B), "dfv_train_fwd: stem parameters missing");
  DFV_PDL((stem_weight_pack_kernel), (27 * stem_c + 255) / 256, 256, 0, st, P(-1, DFV_TG_STEM_W), ar.stem_w, stem_c);
  DFV_LAUNCH_CHECK();
  DFV_TRY(dfv_stem_conv_fwd(a->images_nchw, ar.stem_w, ar.zero_bias, ar.s_raw, dtype, B, a->H, a->W, stem_c, DFV_ACT_NONE, stream));
  const bool frozen = a->freeze_bn != 0;
  // batch statistics (+ running-statistics update), or -- freeze_bn -- the running statistics themselves
  auto bn_stats = [&](const void* raw, long long rows, int C, float* mean, float* invstd, float* rm, float* rv) -> int {
    if (frozen) return bn_frozen_stats(rm, rv, eps, mean, invstd, C, st);
    return dfv_bn_stats_fwd(raw, dtype, B, rows, C, eps, mom, mean, invstd, rm, rv, ar.bn_ws, stream);
  };
  DFV_TRY(bn_stats(ar.s_raw, (long long)s.Hs * s.Ws, stem_c, ar.sm, ar.si, PW(-1, DFV_TG_STEM_RM), PW(-1, DFV_TG_STEM_RV)));
  DFV_TRY(dfv_bn_act_fwd(ar.s_raw, ar.sm, ar.si, P(-1, DFV_TG_STEM_G), P(-1, DFV_TG_STEM_B), DFV_ACT_SILU, nullptr, nullptr, nullptr,
                         ar.a0, nullptr, dtype, B, (long long)s.Hs * s.Ws, stem_c, stream));
  DFV_TRY(tap(0, ar.a0, (size_t)B * s.Hs * s.Ws * stem_c));

  // ---- 32 MBConv blocks
  const void* x = ar.a0;
  for (int i = 0; i < n; ++i) {
    const dfv_block_info& b = blk[i];
    BlockArena& ba = ar.blk[i];
    const int h = s.Hin[i], w = s.Win[i], ho = s.Hout[i], wo = s.Wout[i];
    const long long hw_in = (long long)h * w, hw_out = (long long)ho * wo;
    DFV_REQUIRE(P(i, DFV_T_DW_W) && P(i, DFV_T_PROJ_W) && P(i, DFV_T_SE_R_W) && P(i, DFV_T_SE_E_W), "dfv_train_fwd: block %d parameters missing", i);
    const void* dw_in = x;
    if (b.has_expand) {
      DFV_TRY(dfv_cast_weight(P(i, DFV_T_EXPAND_W), ba.wE, dtype, b.c_mid, b.c_in, 0, stream));
      DFV_TRY(dfv_cast_weight(P(i, DFV_T_EXPAND_W), ba.wEt, dtype, b.c_in, b.c_mid, 1, stream));
      DFV_TRY(dfv_pw_conv_fwd(x, ba.wE, ar.zero_bias, nullptr, (int)hw_in, nullptr, ba.e_raw, dtype, B, B * hw_in, b.c_in, b.c_mid, DFV_ACT_NONE,
                              ar.fold_ws, stream));
      DFV_TRY(bn_stats(ba.e_raw, hw_in, b.c_mid, ba.m0, ba.i0, PW(i, DFV_T_BN0_RM), PW(i, DFV_T_BN0_RV)));
      DFV_TRY(dfv_bn_act_fwd(ba.e_raw, ba.m0, ba.i0, P(i, DFV_T_BN0_G), P(i, DFV_T_BN0_B), DFV_ACT_SILU, nullptr, nullptr, nullptr, ba.e,
                             nullptr, dtype, B, hw_in, b.c_mid, stream));
      dw_in = ba.e;
    }
    DFV_TRY(dfv_dw_weight_pack(P(i, DFV_T_DW_W), ba.wD, b.c_mid, b.kernel, 0, stream));
    DFV_TRY(dfv_dw_weight_pack(P(i, DFV_T_DW_W), ba.wDf, b.c_mid, b.kernel, 1, stream));
    // depthwise conv with the BatchNorm batch statistics accumulated in its epilogue (no separate pass over d_raw)
    DFV_REQUIRE(b.c_mid <= 4096, "dfv_train_fwd: c_mid %d > 4096", b.c_mid);
    if (frozen) {
      DFV_TRY(dfv_dwconv_fwd(dw_in, ba.wD, ar.zero_bias, ba.d_raw, nullptr, dtype, B, h, w, b.c_mid, b.kernel, b.stride, b.pad_lo, b.pad_hi,
                             DFV_ACT_NONE, stream));
      DFV_TRY(bn_frozen_stats(PW(i, DFV_T_BN1_RM), PW(i, DFV_T_BN1_RV), eps, ba.m1, ba.i1, b.c_mid, st));
    } else {
      DFV_CUDA(cudaMemsetAsync(ar.bn_acc, 0, sizeof(double) * 2 * (size_t)b.c_mid, st));
      DFV_TRY(dfv_dwconv_stats_fwd(dw_in, ba.wD, ar.zero_bias, ba.d_raw, ar.bn_acc, dtype, B, h, w, b.c_mid, b.kernel, b.stride, b.pad_lo,
                                   b.pad_hi, stream));
      DFV_TRY(dfv_bn_stats_from_sums(ar.bn_acc, b.c_mid, (double)B * (double)hw_out, eps, mom, ba.m1, ba.i1, PW(i, DFV_T_BN1_RM),
                                     PW(i, DFV_T_BN1_RV), stream));
    }
    DFV_TRY(dfv_bn_act_fwd(ba.d_raw, ba.m1, ba.i1, P(i, DFV_T_BN1_G), P(i, DFV_T_BN1_B), DFV_ACT_SILU, nullptr, nullptr, nullptr, ba.d,
                           ar.pool_ws, dtype, B, hw_out, b.c_mid, stream));
    DFV_TRY(dfv_se_train_fwd(ar.pool_ws, dfv_rows_chunks(B, hw_out), 1.0f / (float)hw_out, P(i, DFV_T_SE_R_W), P(i, DFV_T_SE_R_B),
                             P(i, DFV_T_SE_E_W), P(i, DFV_T_SE_E_B), ba.gate, dtype, ba.pooled, ba.h1, ba.gate_f32, ar.se_scratch, B, b.c_mid, b.se_squeeze, stream));
    DFV_TRY(dfv_cast_weight(P(i, DFV_T_PROJ_W), ba.wP, dtype, b.c_out, b.c_mid, 0, stream));
    DFV_TRY(dfv_cast_weight(P(i, DFV_T_PROJ_W), ba.wPt, dtype, b.c_mid, b.c_out, 1, stream));
    DFV_TRY(dfv_pw_conv_fwd(ba.d, ba.wP, ar.zero_bias, ba.gate, (int)hw_out, nullptr, ba.p_raw, dtype, B, B * hw_out, b.c_mid, b.c_out, DFV_ACT_NONE,
                            ar.fold_ws, stream));
    DFV_TRY(bn_stats(ba.p_raw, hw_out, b.c_out, ba.m2, ba.i2, PW(i, DFV_T_BN2_RM), PW(i, DFV_T_BN2_RV)));
    const float rate = a->drop_connect_rate * (float)i / (float)n;
    const float* rowscale = nullptr;
    if (b.has_skip && rate > 0.f) {
      DFV_TRY(dfv_dropout_mask_dev(ba.dc, B, rate, a->seed + 1000 + (unsigned long long)i, a->seed_dev, stream));
      rowscale = ba.dc;
    }
    DFV_TRY(dfv_bn_act_fwd(ba.p_raw, ba.m2, ba.i2, P(i, DFV_T_BN2_G), P(i, DFV_T_BN2_B), DFV_ACT_NONE, rowscale, b.has_skip ? x : nullptr, nullptr,
                           ba.out, nullptr, dtype, B, hw_out, b.c_out, stream));
    x = ba.out;
    DFV_TRY(tap(1 + i, x, (size_t)B * hw_out * b.c_out));
  }

  // ---- head conv
  const long long hw_f = (long long)s.Hf * s.Wf;
  const int c_last = blk[n - 1].c_out;
  DFV_REQUIRE(P(-1, DFV_TG_HEAD_W), "dfv_train_fwd: head parameters missing");
  DFV_TRY(dfv_cast_weight(P(-1, DFV_TG_HEAD_W), ar.wH, dtype, head_c, c_last, 0, stream));
  DFV_TRY(dfv_cast_weight(P(-1, DFV_TG_HEAD_W), ar.wHt, dtype, c_last, head_c, 1, stream));
  DFV_TRY(dfv_pw_gemm_fwd(x, ar.wH, ar.zero_bias, nullptr, 0, nullptr, ar.h_raw, dtype, B * hw_f, c_last, head_c, DFV_ACT_NONE, stream));
  DFV_TRY(bn_stats(ar.h_raw, hw_f, head_c, ar.hm, ar.hi, PW(-1, DFV_TG_HEAD_RM), PW(-1, DFV_TG_HEAD_RV)));
  DFV_TRY(dfv_bn_act_fwd(ar.h_raw, ar.hm, ar.hi, P(-1, DFV_TG_HEAD_G), P(-1, DFV_TG_HEAD_B), DFV_ACT_SILU, nullptr, nullptr, nullptr, ar.h, nullptr,
                         dtype, B, hw_f, head_c, stream));
  DFV_TRY(tap(1 + n, ar.h, (size_t)B * hw_f * head_c));

  // ---- attention + pool
  const float* heat = nullptr;
  if (a->use_attention && a->use_landmark && a->landmarks != nullptr) {
    DFV_REQUIRE(P(-1, DFV_TG_LM_W), "dfv_train_fwd: landmark attention weights missing");
    DFV_TRY(dfv_landmark_heatmap_fwd(a->landmarks, P(-1, DFV_TG_LM_W), ar.heat, ar.heat_raw, ar.heat_max, nullptr, B, s.Hf, s.Wf,
                                     a->landmark_ref_size > 0.f ? a->landmark_ref_size : 224.0f, 1.5f, a->heat_group, stream));
    heat = ar.heat;
  }
  const int use_c = a->use_attention && a->use_channel, use_s = a->use_attention && a->use_spatial;
  DFV_TRY(dfv_hybrid_attention_train_fwd(ar.h, heat, P(-1, DFV_TG_CA_W1), P(-1, DFV_TG_CA_W2), P(-1, DFV_TG_SA_W), ar.feat_pre, ar.attn_saved,
                                         dtype, B, s.Hf, s.Wf, head_c, use_c ? a->ca_hidden : 0, use_c, use_s, stream));
  if (a->feat_dropout > 0.f) {
    DFV_TRY(dfv_dropout_mask_dev(ar.mask_f, (long long)B * head_c, a->feat_dropout, a->seed + 1, a->seed_dev, stream));
    DFV_TRY(dfv_bn_act_fwd(ar.feat_pre, nullptr, nullptr, nullptr, nullptr, DFV_ACT_NONE, nullptr, nullptr, ar.mask_f, a->features, nullptr,
                           DFV_F32, B, 1, head_c, stream));
  } else {
    DFV_CUDA(cudaMemcpyAsync(a->features, ar.feat_pre, sizeof(float) * (size_t)B * head_c, cudaMemcpyDeviceToDevice, st));
  }

  // ---- classifier
  const float* in = a->features;
  for (int l = 0; l < a->head_layers; ++l) {
    const int din = a->head_dims[l], dout = a->head_dims[l + 1];
    const bool last = l == a->head_layers - 1;
    const float* wl = a->params[dfv_train_cls_index(l, 0)];
    const float* bl = a->params[dfv_train_cls_index(l, 1)];
    DFV_REQUIRE(wl && bl, "dfv_train_fwd: classifier layer %d parameters missing", l);
    float* out = last ? a->logits : ar.cls[l].lin;
    DFV_TRY(dfv_linear_f32_fwd(in, wl, bl, out, ar.bn_ws, max_bn_ws(s, B, a->head_dims, a->head_layers), B, din, dout, 0, stream));
    if (!last) {
      ClsArena& ca = ar.cls[l];
      DFV_TRY(dfv_bn_stats_fwd(ca.lin, DFV_F32, B, 1, dout, a->cls_bn_eps, a->cls_bn_momentum, ca.mean, ca.invstd,
                               const_cast<float*>(a->params[dfv_train_cls_index(l, 4)]), const_cast<float*>(a->params[dfv_train_cls_index(l, 5)]),
                               ar.bn_ws, stream));
      const float* mask = nullptr;
      if (a->cls_dropout > 0.f) {
        DFV_TRY(dfv_dropout_mask_dev(ca.mask, (long long)B * dout, a->cls_dropout, a->seed + 2000 + (unsigned long long)l, a->seed_dev, stream));
        mask = ca.mask;
      }
      DFV_TRY(dfv_bn_act_fwd(ca.lin, ca.mean, ca.invstd, a->params[dfv_train_cls_index(l, 2)], a->params[dfv_train_cls_index(l, 3)], DFV_ACT_RELU,
                             nullptr, nullptr, mask, ca.act, nullptr, DFV_F32, B, 1, dout, stream));
      in = ca.act;
    }
  }
  return DFV_OK;
}

int dfv_train_bwd(const dfv_train_args* a, dfv_stream_t stream) {
  DFV_TRY(check_device());
  DFV_TRY(validate(a, true));
  const int dtype = a->dtype, B = a->B;
  TShapes s;
  DFV_REQUIRE(train_shapes(&s, a->H, a->W) == DFV_OK, "dfv_train_bwd: bad input size");
  Arena ar;
  carve_arena(&ar, a->arena, s, dtype, B, a->head_dims, a->head_layers, a->ca_hidden);
  Scratch sc;
  carve_scratch(&sc, a->scratch, s, dtype, B, a->head_dims, a->head_layers, a->ca_hidden);
  if (ar.bytes > a->arena_bytes || sc.bytes > a->scratch_bytes) {
    set_error("dfv_train_bwd: arena / scratch too small (%zu < %zu or %zu < %zu bytes)", a->arena_bytes, ar.bytes, a->scratch_bytes, sc.bytes);
    return DFV_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  int n;
  const dfv_block_info* blk = topo_blocks(&n);
  const int stem_c = topo_stem_c(), head_c = topo_head_c();
  auto P = [&](int block, int kind) -> const float* { return a->params[dfv_train_index(block, kind)]; };
  auto G = [&](int block, int kind) -> float* { return a->grads[dfv_train_index(block, kind)]; };
  const bool frozen = a->freeze_bn != 0;
  // freeze_bn: eval-mode BatchNorm has no batch-statistics terms in its input gradient (d raw = gamma * invstd * du)
  // and no gamma / beta gradients: run the same two kernels with the two coefficient means zeroed in between.
  auto bn_grad = [&](float* g_ptr) -> float* { return frozen ? nullptr : g_ptr; };
  // Backward through [mask, rowscale, SE gate] -> activation -> BatchNorm of one backbone layer, two streaming passes over
  // (g, raw): the reduction (dgamma, dbeta, the two coefficient means; du is NOT written) and the apply pass, which
  // recomputes du and writes d raw into `out` (may alias g).
  auto bn_backward = [&](const void* gin, const void* raw, const float* mean, const float* invstd, const float* gamma, const float* beta, int act,
                         const void* gate, const float* dpool, float inv_hw, const float* rowscale, void* out, float* dgamma, float* dbeta,
                         long long rows, int C) -> int {
    DFV_TRY(dfv_act_bn_bwd(gin, raw, mean, invstd, gamma, beta, act, gate, dpool, inv_hw, rowscale, nullptr, nullptr, bn_grad(dgamma), bn_grad(dbeta),
                           sc.coef, sc.bn_ws, dtype, B, rows, C, stream));
    if (frozen) DFV_CUDA(cudaMemsetAsync(sc.coef, 0, sizeof(float) * 2 * (size_t)C, st));
    return dfv_act_bn_bwd_apply(gin, raw, mean, invstd, gamma, beta, act, gate, dpool, inv_hw, rowscale, nullptr, sc.coef, out, dtype, B, rows, C,
                                stream);
  };
  auto unit_done = [&](int unit) -> int {
    if (a->grad_events && a->grad_events[unit]) DFV_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(a->grad_events[unit]), st));
    return DFV_OK;
  };

  // ---- classifier
  const float* g = a->dlogits;
  int cur = 0;
  for (int l = a->head_layers - 1; l >= 0; --l) {
    const int din = a->head_dims[l], dout = a->head_dims[l + 1];
    const bool last = l == a->head_layers - 1;
    const float* in_l = l == 0 ? a->features : ar.cls[l - 1].act;
    const float* gl = g;
    if (!last) {
      ClsArena& ca = ar.cls[l];
      float* du = sc.gcls[cur];   // g already lives in gcls[cur] (written by the previous dgrad): in place
      DFV_TRY(dfv_act_bn_bwd(g, ca.lin, ca.mean, ca.invstd, a->params[dfv_train_cls_index(l, 2)], a->params[dfv_train_cls_index(l, 3)],
                             DFV_ACT_RELU, nullptr, nullptr, 0.f, nullptr, a->cls_dropout > 0.f ? ca.mask : nullptr, du,
                             a->grads[dfv_train_cls_index(l, 2)], a->grads[dfv_train_cls_index(l, 3)], sc.coef, sc.bn_ws, DFV_F32, B, 1, dout, stream));
      DFV_TRY(dfv_bn_bwd_apply(du, ca.lin, ca.mean, ca.invstd, a->params[dfv_train_cls_index(l, 2)], sc.coef, du, DFV_F32, B, dout, stream));
      gl = du;
    }
    DFV_TRY(dfv_pw_wgrad(gl, in_l, nullptr, 0, a->grads[dfv_train_cls_index(l, 0)], DFV_F32, B, din, dout, stream));
    DFV_TRY(dfv_colsum(gl, B, dout, a->grads[dfv_train_cls_index(l, 1)], stream));
    // input gradient: the torch weight [dout][din] read K-major -- no transposed copy
    DFV_TRY(dfv_linear_f32_fwd(gl, a->params[dfv_train_cls_index(l, 0)], nullptr, sc.gcls[cur ^ 1], sc.bn_ws,
                               max_bn_ws(s, B, a->head_dims, a->head_layers), B, dout, din, 1, stream));
    cur ^= 1;
    g = sc.gcls[cur];
  }
  // d features: classifier path + the loss's direct feature gradient, through the feature dropout
  DFV_TRY(dfv_add_mul(g, a->dfeatures, a->feat_dropout > 0.f ? ar.mask_f : nullptr, sc.dfeat, (long long)B * head_c, stream));

  // ---- attention + head conv
  const long long hw_f = (long long)s.Hf * s.Wf;
  const int c_last = blk[n - 1].c_out;
  const bool has_heat = a->use_attention && a->use_landmark && a->landmarks != nullptr;
  const int use_c = a->use_attention && a->use_channel, use_s = a->use_attention && a->use_spatial;
  DFV_TRY(dfv_hybrid_attention_bwd(ar.h, has_heat ? ar.heat : nullptr, P(-1, DFV_TG_CA_W1), P(-1, DFV_TG_CA_W2), P(-1, DFV_TG_SA_W), sc.dfeat,
                                   ar.attn_saved, sc.gE, has_heat ? sc.dheat : nullptr, G(-1, DFV_TG_CA_W1), G(-1, DFV_TG_CA_W2),
                                   G(-1, DFV_TG_SA_W), sc.attn_ws, dtype, B, s.Hf, s.Wf, head_c, use_c ? a->ca_hidden : 0, use_c, use_s, stream));
  if (has_heat && G(-1, DFV_TG_LM_W))
    DFV_TRY(dfv_landmark_heatmap_bwd(a->landmarks, P(-1, DFV_TG_LM_W), ar.heat_raw, ar.heat_max, sc.dheat, G(-1, DFV_TG_LM_W), B, s.Hf, s.Wf,
                                     a->landmark_ref_size > 0.f ? a->landmark_ref_size : 224.0f, 1.5f, a->heat_group, stream));
  DFV_TRY(bn_backward(sc.gE, ar.h_raw, ar.hm, ar.hi, P(-1, DFV_TG_HEAD_G), P(-1, DFV_TG_HEAD_B), DFV_ACT_SILU, nullptr, nullptr, 0.f, nullptr, sc.gE,
                      G(-1, DFV_TG_HEAD_G), G(-1, DFV_TG_HEAD_B), hw_f, head_c));
  DFV_TRY(dfv_pw_wgrad(sc.gE, ar.blk[n - 1].out, nullptr, 0, G(-1, DFV_TG_HEAD_W), dtype, B * hw_f, c_last, head_c, stream));
  int gc = 0;
  DFV_TRY(dfv_pw_gemm_fwd(sc.gE, ar.wHt, ar.zero_bias, nullptr, 0, nullptr, sc.gout[gc], dtype, B * hw_f, head_c, c_last, DFV_ACT_NONE, stream));
  DFV_TRY(unit_done(0));

  // ---- blocks, last to first.  sc.gout[gc] = gradient wrt the block's output.
  for (int i = n - 1; i >= 0; --i) {
    const dfv_block_info& b = blk[i];
    BlockArena& ba = ar.blk[i];
    const int h = s.Hin[i], w = s.Win[i], ho = s.Hout[i], wo = s.Wout[i];
    const long long hw_in = (long long)h * w, hw_out = (long long)ho * wo;
    const void* x = i == 0 ? (const void*)ar.a0 : (const void*)ar.blk[i - 1].out;
    const void* gy = sc.gout[gc];
    const float rate = a->drop_connect_rate * (float)i / (float)n;
    const float* rowscale = (b.has_skip && rate > 0.f) ? ba.dc : nullptr;
    // bn2 (no activation); the skip branch keeps gy
    DFV_TRY(bn_backward(gy, ba.p_raw, ba.m2, ba.i2, P(i, DFV_T_BN2_G), P(i, DFV_T_BN2_B), DFV_ACT_NONE, nullptr, nullptr, 0.f, rowscale, sc.gP,
                        G(i, DFV_T_BN2_G), G(i, DFV_T_BN2_B), hw_out, b.c_out));
    // project conv
    DFV_TRY(dfv_pw_wgrad(sc.gP, ba.d, ba.gate, (int)hw_out, G(i, DFV_T_PROJ_W), dtype, B * hw_out, b.c_mid, b.c_out, stream));
    DFV_TRY(dfv_pw_conv_fwd(sc.gP, ba.wPt, ar.zero_bias, nullptr, (int)hw_out, nullptr, sc.gA, dtype, B, B * hw_out, b.c_out, b.c_mid, DFV_ACT_NONE,
                            sc.fold_ws, stream));
    if (dtype == DFV_BF16) {
      // squeeze-excite + the reduction half of [gate, swish, bn1] in ONE pass over (gA, d_raw): the reduction is linear in
      // (gate, dpool) per image, so it runs before the SE backward (which it also feeds: sum_hw gA * swish(u)) and is
      // combined with gate / dpool afterwards (dfv_act_bn_bwd_gated_reduce).  Saves the SE backward's own pass over (gA, d).
      DFV_TRY(dfv_act_bn_bwd_gated_reduce(sc.gA, ba.d_raw, ba.m1, ba.i1, P(i, DFV_T_BN1_G), P(i, DFV_T_BN1_B), sc.bn_ws, sc.se_ws, dtype, B, hw_out,
                                          b.c_mid, stream));
      DFV_TRY(dfv_se_bwd_from_partials(ba.gate_f32, ba.pooled, ba.h1, P(i, DFV_T_SE_R_W), P(i, DFV_T_SE_E_W), sc.dpool, G(i, DFV_T_SE_R_W),
                                       G(i, DFV_T_SE_R_B), G(i, DFV_T_SE_E_W), G(i, DFV_T_SE_E_B), sc.se_ws, B, hw_out, b.c_mid, b.se_squeeze,
                                       stream));
      DFV_TRY(dfv_bn_bwd_gated_finalize(sc.bn_ws, ba.gate, sc.dpool, 1.0f / (float)hw_out, ba.m1, ba.i1, bn_grad(G(i, DFV_T_BN1_G)),
                                        bn_grad(G(i, DFV_T_BN1_B)), sc.coef, dtype, B, hw_out, b.c_mid, stream));
      if (frozen) DFV_CUDA(cudaMemsetAsync(sc.coef, 0, sizeof(float) * 2 * (size_t)b.c_mid, st));
      DFV_TRY(dfv_act_bn_bwd_apply(sc.gA, ba.d_raw, ba.m1, ba.i1, P(i, DFV_T_BN1_G), P(i, DFV_T_BN1_B), DFV_ACT_SILU, ba.gate, sc.dpool,
                                   1.0f / (float)hw_out, nullptr, nullptr, sc.coef, sc.gA, dtype, B, hw_out, b.c_mid, stream));
    } else {
      // squeeze-excite
      DFV_TRY(dfv_se_bwd(sc.gA, ba.d, dtype, ba.gate_f32, ba.pooled, ba.h1, P(i, DFV_T_SE_R_W), P(i, DFV_T_SE_E_W), sc.dpool, G(i, DFV_T_SE_R_W),
                         G(i, DFV_T_SE_R_B), G(i, DFV_T_SE_E_W), G(i, DFV_T_SE_E_B), sc.se_ws, B, hw_out, b.c_mid, b.se_squeeze, stream));
      // gate, swish, bn1
      DFV_TRY(bn_backward(sc.gA, ba.d_raw, ba.m1, ba.i1, P(i, DFV_T_BN1_G), P(i, DFV_T_BN1_B), DFV_ACT_SILU, ba.gate, sc.dpool, 1.0f / (float)hw_out,
                          nullptr, sc.gA, G(i, DFV_T_BN1_G), G(i, DFV_T_BN1_B), hw_out, b.c_mid));
    }
    // depthwise conv
    const void* dw_in = b.has_expand ? (const void*)ba.e : x;
    const int kk = b.kernel * b.kernel;
    DFV_CUDA(cudaMemsetAsync(sc.dwkkc, 0, sizeof(float) * (size_t)kk * b.c_mid, st));
    DFV_TRY(dfv_dwconv_wgrad(sc.gA, dw_in, sc.dwkkc, dtype, B, h, w, b.c_mid, b.kernel, b.stride, b.pad_lo, b.pad_hi, stream));
    DFV_TRY(dfv_dw_weight_unpack(sc.dwkkc, G(i, DFV_T_DW_W), b.c_mid, b.kernel, stream));
    if (b.stride == 1) {   // dgrad of a stride-1 conv = the forward TMA kernel with flipped taps and mirrored pads
      DFV_TRY(dfv_dwconv_fwd(sc.gA, ba.wDf, ar.zero_bias, sc.gE, nullptr, dtype, B, ho, wo, b.c_mid, b.kernel, 1, b.kernel - 1 - b.pad_lo,
                             b.kernel - 1 - b.pad_hi, DFV_ACT_NONE, stream));
    } else {
      DFV_TRY(dfv_dwconv_dgrad(sc.gA, ba.wD, sc.gE, dtype, B, h, w, b.c_mid, b.kernel, b.stride, b.pad_lo, b.pad_hi, stream));
    }
    void* gx = sc.gout[gc ^ 1];
    if (b.has_expand) {
      DFV_TRY(bn_backward(sc.gE, ba.e_raw, ba.m0, ba.i0, P(i, DFV_T_BN0_G), P(i, DFV_T_BN0_B), DFV_ACT_SILU, nullptr, nullptr, 0.f, nullptr, sc.gE,
                          G(i, DFV_T_BN0_G), G(i, DFV_T_BN0_B), hw_in, b.c_mid));
      DFV_TRY(dfv_pw_wgrad(sc.gE, x, nullptr, 0, G(i, DFV_T_EXPAND_W), dtype, B * hw_in, b.c_in, b.c_mid, stream));
      DFV_TRY(dfv_pw_conv_fwd(sc.gE, ba.wEt, ar.zero_bias, nullptr, (int)hw_in, b.has_skip ? gy : nullptr, gx, dtype, B, B * hw_in, b.c_mid, b.c_in,
                              DFV_ACT_NONE, sc.fold_ws, stream));
    } else {
      // no expand conv: gE is already the gradient wrt the block input (c_mid == c_in); add the skip gradient
      DFV_TRY(dfv_bn_act_fwd(sc.gE, nullptr, nullptr, nullptr, nullptr, DFV_ACT_NONE, nullptr, b.has_skip ? gy : nullptr, nullptr, gx, nullptr, dtype,
                             B, hw_in, b.c_in, stream));
    }
    gc ^= 1;
    DFV_TRY(unit_done(1 + (n - 1 - i)));
  }

  // ---- stem
  DFV_TRY(bn_backward(sc.gout[gc], ar.s_raw, ar.sm, ar.si, P(-1, DFV_TG_STEM_G), P(-1, DFV_TG_STEM_B), DFV_ACT_SILU, nullptr, nullptr, 0.f, nullptr,
                      sc.gout[gc], G(-1, DFV_TG_STEM_G), G(-1, DFV_TG_STEM_B), (long long)s.Hs * s.Ws, stem_c));
  DFV_TRY(dfv_stem_wgrad(sc.gout[gc], a->images_nchw, G(-1, DFV_TG_STEM_W), dtype, B, a->H, a->W, stream));
  DFV_TRY(unit_done(DFV_GRAD_UNITS - 1));
  return DFV_OK;
}

}  // extern "C"
