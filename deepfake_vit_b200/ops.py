"""Per-operator Python wrappers over the C ABI (torch tensors in, torch tensors out).

PyTorch is plumbing only here: it owns device memory and the stream; every computation is a
libdfvit kernel.  Activations are NHWC tensors of dtype float32 or bfloat16.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import DFV_ACT_NONE, DFV_ACT_SILU, DFV_BF16, DFV_F32, check, lib  # noqa: F401


def dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return DFV_F32
    if t == torch.bfloat16:
        return DFV_BF16
    raise TypeError(f"unsupported activation dtype {t} (float32 or bfloat16)")


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "libdfvit needs contiguous CUDA tensors"
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t):
    assert t.dtype == torch.float32
    return _ptr(t)


def stem_conv(x_nchw, w_khwc, bias, out_dtype, act=DFV_ACT_SILU):
    B, _, H, W = x_nchw.shape
    C_ = bias.numel()
    Ho, Wo = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1
    y = torch.empty(B, Ho, Wo, C_, device=x_nchw.device, dtype=out_dtype)
    check(lib.dfv_stem_conv_fwd(_f32(x_nchw), _f32(w_khwc), _f32(bias), _ptr(y), dtype_code(out_dtype), B, H, W, C_, act, _stream()))
    return y


def stem_conv_u8(x_hwc, mean, std, w_khwc, bias, out_dtype, act=DFV_ACT_SILU):
    """x_hwc: [B,H,W,3] uint8 RGB crops; normalised (u8 / 255 - mean) / std inside the kernel."""
    B, H, W, _ = x_hwc.shape
    C_ = bias.numel()
    Ho, Wo = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1
    y = torch.empty(B, Ho, Wo, C_, device=x_hwc.device, dtype=out_dtype)
    assert x_hwc.dtype == torch.uint8
    norm = (C.c_float * 6)(*mean, *std)
    check(lib.dfv_stem_conv_u8_fwd(_ptr(x_hwc), norm, _f32(w_khwc), _f32(bias), _ptr(y), dtype_code(out_dtype), B, H, W, C_, act, _stream()))
    return y


def u8_to_nchw(x_hwc, mean, std):
    B, H, W, _ = x_hwc.shape
    y = torch.empty(B, 3, H, W, device=x_hwc.device, dtype=torch.float32)
    check(lib.dfv_u8_to_nchw_f32(_ptr(x_hwc), (C.c_float * 6)(*mean, *std), _f32(y), B, H, W, _stream()))
    return y


def dwconv(x, w_kkc, bias, kernel, stride, pad_lo, pad_hi, act=DFV_ACT_SILU, want_pool=True, tuning=None):
    """x: [B,H,W,C].  Returns (y [B,Ho,Wo,C], pool_partial [B,parts,C] fp32 or None).
    tuning: optional (L, TW, TH, CB) restriction of the tile plan (0 = free)."""
    B, H, W, C_ = x.shape
    dt = dtype_code(x.dtype)
    Ho = (H + pad_lo + pad_hi - kernel) // stride + 1
    Wo = (W + pad_lo + pad_hi - kernel) // stride + 1
    y = torch.empty(B, Ho, Wo, C_, device=x.device, dtype=x.dtype)
    pool = None
    tn = C.byref(_lib.DwconvTuning(*tuning)) if tuning is not None else None
    if want_pool:
        parts = lib.dfv_dwconv_pool_parts_tuned(dt, B, H, W, C_, kernel, stride, pad_lo, pad_hi, tn)
        if parts <= 0:
            check(parts)
        pool = torch.empty(B, parts, C_, device=x.device, dtype=torch.float32)
    check(lib.dfv_dwconv_fwd_tuned(_ptr(x), _f32(w_kkc), _f32(bias), _ptr(y), _ptr(pool), dt, B, H, W, C_, kernel, stride,
                                   pad_lo, pad_hi, act, tn, _stream()))
    return y, pool


def dwconv_se(x, w_kkc, bias, kernel, stride, pad_lo, pad_hi, w_reduce, zero_next=None):
    """Depthwise conv + folded BN + swish with the SE squeeze layer fused into the kernel's tail.
    Returns (y, pool_partial, hid_fix [B, squeeze] int64 fixed-point hidden sums, 2^-30 resolution, bias not added)."""
    B, H, W, C_ = x.shape
    dt = dtype_code(x.dtype)
    sq = w_reduce.shape[0]
    args = (dt, B, H, W, C_, kernel, stride, pad_lo, pad_hi)
    if not lib.dfv_dwconv_se_supported(*args, sq):
        raise _lib.DfvError("this layer's tile plan cannot host the fused squeeze")
    Ho = (H + pad_lo + pad_hi - kernel) // stride + 1
    Wo = (W + pad_lo + pad_hi - kernel) // stride + 1
    y = torch.empty(B, Ho, Wo, C_, device=x.device, dtype=x.dtype)
    pool = torch.empty(B, lib.dfv_dwconv_pool_parts(*args), C_, device=x.device, dtype=torch.float32)
    hid = torch.zeros(B, sq, device=x.device, dtype=torch.int64)
    check(lib.dfv_dwconv_se_fwd(_ptr(x), _f32(w_kkc), _f32(bias), _ptr(y), _f32(pool), _f32(w_reduce), _ptr(hid), _ptr(zero_next),
                                zero_next.numel() if zero_next is not None else 0, sq, dt, B, H, W, C_, kernel, stride, pad_lo, pad_hi,
                                _stream()))
    return y, pool, hid


def se_excite(hid_fix, b_reduce, w_expand_t, b_expand, gate_dtype=torch.float32):
    B, sq = hid_fix.shape
    C_ = b_expand.numel()
    gate = torch.empty(B, C_, device=hid_fix.device, dtype=gate_dtype)
    check(lib.dfv_se_excite_fwd(_ptr(hid_fix), _f32(b_reduce), _f32(w_expand_t), _f32(b_expand), _ptr(gate), dtype_code(gate_dtype), B, C_,
                                sq, _stream()))
    return gate


def se_gate(pool_partial, hw, w_reduce, b_reduce, w_expand_t, b_expand, gate_dtype=torch.float32):
    B, parts, C_ = pool_partial.shape
    sq = b_reduce.numel()
    gate = torch.empty(B, C_, device=pool_partial.device, dtype=gate_dtype)
    scratch = _f32buf(lib.dfv_se_scratch_floats(B, C_, sq), pool_partial.device)
    check(lib.dfv_se_gate_fwd(_f32(pool_partial), parts, 1.0 / hw, _f32(w_reduce), _f32(b_reduce), _f32(w_expand_t),
                              _f32(b_expand), _ptr(gate), dtype_code(gate_dtype), _f32(scratch), B, C_, sq, _stream()))
    return gate


def pw_gemm(a, w, bias, act=DFV_ACT_NONE, a_scale=None, rows_per_image=0, residual=None, tuning=None):
    """a: [..., K] (NHWC activations), w: [N, K].  Returns [..., N].
    tuning: optional (weight_stationary in {-1, 0, 1}, N tile) restriction of the tile plan."""
    K = a.shape[-1]
    N = w.shape[0]
    M = a.numel() // K
    assert w.shape[1] == K and w.dtype == a.dtype
    assert a_scale is None or a_scale.dtype == a.dtype, "the SE gate has the activation dtype"
    out = torch.empty(*a.shape[:-1], N, device=a.device, dtype=a.dtype)
    tn = C.byref(_lib.GemmTuning(*tuning)) if tuning is not None else None
    check(lib.dfv_pw_gemm_fwd_tuned(_ptr(a), _ptr(w), _f32(bias), _ptr(a_scale), rows_per_image, _ptr(residual), _ptr(out),
                                    dtype_code(a.dtype), M, K, N, act, tn, _stream()))
    return out


def landmark_heatmap(landmarks, weights5, H, W, ref_size=224.0, sigma=1.5, group=0, return_scaled=False, max_floor=None,
                     return_max_key=False):
    """max_floor: optional int32 tensor holding ONE order-preserving key (see landmark_max_key) the group maximum is raised
    to -- the data-parallel `global` normaliser.  return_max_key: also return this call's own key(s) [groups] int32."""
    B = landmarks.shape[0]
    dev = landmarks.device
    heat = torch.empty(B, H, W, device=dev, dtype=torch.float32)
    raw = torch.empty(B * H * W, device=dev, dtype=torch.float32)
    mx = torch.empty(B, device=dev, dtype=torch.int32)
    scaled = torch.empty(B, 5, 2, device=dev, dtype=torch.float32) if return_scaled else None
    check(lib.dfv_landmark_heatmap_fwd_ex(_f32(landmarks), _f32(weights5), _f32(heat), _f32(raw), _ptr(mx), _ptr(scaled),
                                          B, H, W, ref_size, sigma, group, _ptr(max_floor), _stream()))
    out = (heat, scaled) if return_scaled else heat
    if return_max_key:
        g = B if group <= 0 or group > B else group
        return out, mx[:(B + g - 1) // g]
    return out


def landmark_max_key(landmarks, weights5, H, W, ref_size=224.0, sigma=1.5):
    """The whole-call heat-map maximum of this batch as an order-preserving key (int32 storage of the library's uint32 key:
    UNSIGNED order == float order).  parallel.allreduce_max_key() turns every rank's key into the global one."""
    _, key = landmark_heatmap(landmarks, weights5, H, W, ref_size, sigma, 0, return_max_key=True)
    return key[:1]


def hybrid_attention(fmap, heat, ca_w1, ca_w2_t, sa_w, use_channel=True, use_spatial=True, return_gates=False):
    """fmap: [B,H,W,C]; heat: [B,H,W] fp32 or None.  Returns pooled features [B,C] fp32."""
    B, H, W, C_ = fmap.shape
    dev = fmap.device
    feats = torch.empty(B, C_, device=dev, dtype=torch.float32)
    cg = torch.empty(B, C_, device=dev, dtype=torch.float32) if return_gates and use_channel else None
    sg = torch.empty(B, H * W, device=dev, dtype=torch.float32) if return_gates and use_spatial else None
    hidden = ca_w1.shape[0] if use_channel else 0
    scratch = _f32buf(lib.dfv_attention_scratch_floats(B, H, W, C_, hidden), dev)
    check(lib.dfv_hybrid_attention_fwd(_ptr(fmap), _ptr(heat), _ptr(ca_w1), _ptr(ca_w2_t), _ptr(sa_w), _f32(feats),
                                       _ptr(cg), _ptr(sg), _f32(scratch), dtype_code(fmap.dtype), B, H, W, C_, hidden,
                                       int(use_channel), int(use_spatial), _stream()))
    return (feats, cg, sg) if return_gates else feats


class HeadPack:
    """Host-side pointer tables for the classifier head (kept alive with their tensors)."""

    def __init__(self, w_t, b):
        self.w_t, self.b = list(w_t), list(b)
        n = len(self.w_t)
        self.dims = (C.c_int32 * (n + 1))(*([self.w_t[0].shape[0]] + [w.shape[1] for w in self.w_t]))
        self.wp = (C.c_void_p * n)(*[_f32(w) for w in self.w_t])
        self.bp = (C.c_void_p * n)(*[_f32(x) for x in self.b])
        self.n = n


def mlp_head(features, pack: HeadPack):
    B = features.shape[0]
    logits = torch.empty(B, pack.dims[pack.n], device=features.device, dtype=torch.float32)
    dims = C.cast(pack.dims, C.POINTER(C.c_int))
    scratch = _f32buf(lib.dfv_mlp_head_scratch_floats(dims, pack.n, B), features.device)
    check(lib.dfv_mlp_head_fwd(_f32(features), pack.wp, pack.bp, dims, pack.n, _f32(logits), _f32(scratch), B, _stream()))
    return logits


def class_weight_sum(targets, class_weights, n_classes):
    """sum_i class_weights[targets[i]] (B when class_weights is None) as a 1-element fp32 device tensor."""
    out = torch.empty(1, device=targets.device, dtype=torch.float32)
    assert targets.dtype == torch.int64
    check(lib.dfv_class_weight_sum(_ptr(targets), _ptr(class_weights), _f32(out), targets.numel(), n_classes, _stream()))
    return out


def combined_loss(logits, targets, features, class_weights, w_ce, w_focal, w_con, want_grad=True, ce_norm=None):
    """Returns (losses[4] = ce, focal, contrastive, total; has_contrastive; dlogits; dfeatures).
    ce_norm: optional 1-element fp32 device tensor replacing the weighted CE's local normaliser."""
    B, Cn = logits.shape
    D = features.shape[1] if features is not None else 0
    dev = logits.device
    losses = torch.empty(4, device=dev, dtype=torch.float32)
    dlogits = torch.empty_like(logits) if want_grad else None
    dfeat = torch.empty_like(features) if (want_grad and features is not None) else None
    has = C.c_int(0)
    assert targets.dtype == torch.int64
    check(lib.dfv_combined_loss_fwd_bwd_ex(_f32(logits), _ptr(targets), _ptr(features), _ptr(class_weights), w_ce, w_focal,
                                           w_con, _f32(losses), _ptr(dlogits), _ptr(dfeat), B, Cn, D, C.byref(has),
                                           _ptr(ce_norm), _stream()))
    return losses, bool(has.value), dlogits, dfeat


# --------------------------------------------------------------------------------------------------
# Training-path operators (train_ops.cu, conv_bwd.cu, attention_train.cu).  Channels-last [B, rows.., C].
# --------------------------------------------------------------------------------------------------
def _rows(x):
    B, C_ = x.shape[0], x.shape[-1]
    return B, x.numel() // (B * C_), C_


_GUARD = None          # test aid: a list collects (buffer, n, band) while guard mode is on
_GUARD_VALUE = 12345.678


def _f32buf(n, dev):
    """fp32 scratch / output buffer of n floats.  In guard mode (tests) it sits between two sentinel bands so that a
    kernel writing outside the size its *_scratch_floats() function published is caught (guards_intact())."""
    n = max(int(n), 1)
    if _GUARD is None:
        return torch.empty(n, device=dev, dtype=torch.float32)
    band = 1024
    full = torch.full((n + 2 * band,), _GUARD_VALUE, device=dev, dtype=torch.float32)
    _GUARD.append((full, n, band))
    return full[band:band + n]


def guard_mode(on: bool):
    global _GUARD
    _GUARD = [] if on else None


def guards_intact() -> bool:
    torch.cuda.synchronize()
    ref = torch.tensor(_GUARD_VALUE, dtype=torch.float32).item()
    return all(bool((full[:band] == ref).all()) and bool((full[band + n:] == ref).all()) for full, n, band in (_GUARD or []))


def bn_stats(raw, eps, momentum=0.0, running_mean=None, running_var=None):
    """Batch statistics of raw [B, ..., C] -> (mean, invstd); running stats updated in place if given."""
    B, rows, C_ = _rows(raw)
    mean, invstd = _f32buf(C_, raw.device), _f32buf(C_, raw.device)
    ws = _f32buf(lib.dfv_bn_ws_floats(B, rows, C_), raw.device)
    check(lib.dfv_bn_stats_fwd(_ptr(raw), dtype_code(raw.dtype), B, rows, C_, eps, momentum, _f32(mean), _f32(invstd),
                               _ptr(running_mean), _ptr(running_var), _f32(ws), _stream()))
    return mean, invstd


def bn_act(raw, mean, invstd, gamma, beta, act=DFV_ACT_NONE, rowscale=None, residual=None, mask=None, want_pool=False):
    B, rows, C_ = _rows(raw)
    out = torch.empty_like(raw)
    pool = None
    if want_pool:
        pool = torch.empty(B, lib.dfv_rows_chunks(B, rows), C_, device=raw.device, dtype=torch.float32)
    check(lib.dfv_bn_act_fwd(_ptr(raw), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(beta), act, _ptr(rowscale), _ptr(residual),
                             _ptr(mask), _ptr(out), _ptr(pool), dtype_code(raw.dtype), B, rows, C_, _stream()))
    return (out, pool) if want_pool else out


def act_bn_bwd(g, raw, mean, invstd, gamma, beta, act=DFV_ACT_NONE, gate=None, dpool=None, inv_hw=0.0, rowscale=None,
               mask=None):
    """Returns (d raw, dgamma, dbeta)."""
    B, rows, C_ = _rows(raw)
    dev = raw.device
    du = torch.empty_like(raw)
    dgamma, dbeta, coef = _f32buf(C_, dev), _f32buf(C_, dev), _f32buf(2 * C_, dev)
    ws = _f32buf(lib.dfv_bn_ws_floats(B, rows, C_), dev)
    code = dtype_code(raw.dtype)
    check(lib.dfv_act_bn_bwd(_ptr(g), _ptr(raw), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(beta), act, _ptr(gate), _ptr(dpool),
                             inv_hw, _ptr(rowscale), _ptr(mask), _ptr(du), _f32(dgamma), _f32(dbeta), _f32(coef), _f32(ws), code,
                             B, rows, C_, _stream()))
    check(lib.dfv_bn_bwd_apply(_ptr(du), _ptr(raw), _f32(mean), _f32(invstd), _ptr(gamma), _f32(coef), _ptr(du), code, B * rows,
                               C_, _stream()))
    return du, dgamma, dbeta


def pw_wgrad(g, a, a_scale=None, rows_per_image=0):
    """dW [N, K] fp32 = sum_m g[m, n] * a[m, k] * a_scale[m // rows_per_image, k]."""
    N, K = g.shape[-1], a.shape[-1]
    M = a.numel() // K
    dw = torch.zeros(N, K, device=g.device, dtype=torch.float32)
    check(lib.dfv_pw_wgrad(_ptr(g), _ptr(a), _ptr(a_scale), rows_per_image, _f32(dw), dtype_code(g.dtype), M, K, N, _stream()))
    return dw


def dwconv_dgrad(g, w_kkc, H, W, kernel, stride, pad_lo, pad_hi):
    B, _, _, C_ = g.shape
    dx = torch.empty(B, H, W, C_, device=g.device, dtype=g.dtype)
    check(lib.dfv_dwconv_dgrad(_ptr(g), _f32(w_kkc), _ptr(dx), dtype_code(g.dtype), B, H, W, C_, kernel, stride, pad_lo, pad_hi,
                               _stream()))
    return dx


def dwconv_wgrad(g, x, kernel, stride, pad_lo, pad_hi):
    B, H, W, C_ = x.shape
    dw = torch.zeros(kernel * kernel, C_, device=g.device, dtype=torch.float32)
    check(lib.dfv_dwconv_wgrad(_ptr(g), _ptr(x), _f32(dw), dtype_code(g.dtype), B, H, W, C_, kernel, stride, pad_lo, pad_hi,
                               _stream()))
    return dw


def stem_wgrad(g, x_nchw):
    B, _, H, W = x_nchw.shape
    dw = torch.zeros(48, 3, 3, 3, device=g.device, dtype=torch.float32)
    check(lib.dfv_stem_wgrad(_ptr(g), _f32(x_nchw), _f32(dw), dtype_code(g.dtype), B, H, W, _stream()))
    return dw


def se_train_fwd(pool_partial, hw, w_reduce, b_reduce, w_expand, b_expand, gate_dtype=torch.float32):
    """Torch layouts: w_reduce [sq, C], w_expand [C, sq].  Returns (gate, pooled, h1, gate_f32)."""
    B, parts, C_ = pool_partial.shape
    sq, dev = b_reduce.numel(), pool_partial.device
    gate = torch.empty(B, C_, device=dev, dtype=gate_dtype)
    pooled, h1, g32 = _f32buf(B * C_, dev).view(B, C_), _f32buf(B * sq, dev).view(B, sq), _f32buf(B * C_, dev).view(B, C_)
    scratch = _f32buf(lib.dfv_se_scratch_floats(B, C_, sq), dev)
    check(lib.dfv_se_train_fwd(_f32(pool_partial), parts, 1.0 / hw, _f32(w_reduce), _f32(b_reduce), _f32(w_expand), _f32(b_expand),
                               _ptr(gate), dtype_code(gate_dtype), _f32(pooled), _f32(h1), _f32(g32), _f32(scratch), B, C_, sq, _stream()))
    return gate, pooled, h1, g32


def se_bwd(da, d, gate_f32, pooled, h1, w_reduce, w_expand):
    """Returns (dpool [B, C], dw_reduce, db_reduce, dw_expand, db_expand)."""
    B, rows, C_ = _rows(d)
    sq, dev = h1.shape[1], d.device
    dpool = _f32buf(B * C_, dev).view(B, C_)
    dw1, db1 = torch.zeros_like(w_reduce), _f32buf(sq, dev)
    dw2, db2 = torch.zeros_like(w_expand), _f32buf(C_, dev)
    ws = _f32buf(lib.dfv_se_bwd_ws_floats(B, rows, C_, sq), dev)
    check(lib.dfv_se_bwd(_ptr(da), _ptr(d), dtype_code(d.dtype), _f32(gate_f32), _f32(pooled), _f32(h1), _f32(w_reduce),
                         _f32(w_expand), _f32(dpool), _f32(dw1), _f32(db1), _f32(dw2), _f32(db2), _f32(ws), B, rows, C_, sq,
                         _stream()))
    return dpool, dw1, db1, dw2, db2


def gated_bn_se_bwd(da, raw, mean, invstd, gamma, beta, gate, gate_f32, pooled, h1, w_reduce, w_expand):
    """The depthwise-output backward of a training block in ONE reduction pass over (da, raw) (bf16):
    dfv_act_bn_bwd_gated_reduce -> dfv_se_bwd_from_partials -> dfv_bn_bwd_gated_finalize -> dfv_act_bn_bwd_apply.
    Returns (d raw, dgamma, dbeta, dpool, dw_reduce, db_reduce, dw_expand, db_expand)."""
    B, rows, C_ = _rows(raw)
    sq, dev = h1.shape[1], raw.device
    code = dtype_code(raw.dtype)
    ws4 = _f32buf(lib.dfv_bn_ws_floats(B, rows, C_), dev)
    se_ws = _f32buf(lib.dfv_se_bwd_ws_floats(B, rows, C_, sq), dev)
    dpool = _f32buf(B * C_, dev).view(B, C_)
    dw1, db1 = torch.zeros_like(w_reduce), _f32buf(sq, dev)
    dw2, db2 = torch.zeros_like(w_expand), _f32buf(C_, dev)
    dgamma, dbeta, coef = _f32buf(C_, dev), _f32buf(C_, dev), _f32buf(2 * C_, dev)
    out = torch.empty_like(raw)
    check(lib.dfv_act_bn_bwd_gated_reduce(_ptr(da), _ptr(raw), _f32(mean), _f32(invstd), _f32(gamma), _f32(beta), _f32(ws4), _f32(se_ws),
                                          code, B, rows, C_, _stream()))
    check(lib.dfv_se_bwd_from_partials(_f32(gate_f32), _f32(pooled), _f32(h1), _f32(w_reduce), _f32(w_expand), _f32(dpool), _f32(dw1),
                                       _f32(db1), _f32(dw2), _f32(db2), _f32(se_ws), B, rows, C_, sq, _stream()))
    check(lib.dfv_bn_bwd_gated_finalize(_f32(ws4), _ptr(gate), _f32(dpool), 1.0 / rows, _f32(mean), _f32(invstd), _f32(dgamma), _f32(dbeta),
                                        _f32(coef), code, B, rows, C_, _stream()))
    check(lib.dfv_act_bn_bwd_apply(_ptr(da), _ptr(raw), _f32(mean), _f32(invstd), _f32(gamma), _f32(beta), DFV_ACT_SILU, _ptr(gate),
                                   _f32(dpool), 1.0 / rows, None, None, _f32(coef), _ptr(out), code, B, rows, C_, _stream()))
    return out, dgamma, dbeta, dpool, dw1, db1, dw2, db2


def hybrid_attention_train(fmap, heat, ca_w1, ca_w2, sa_w, use_channel=True, use_spatial=True):
    """Torch layouts (ca_w2 = fc.2.weight [C, hidden]).  Returns (features, saved)."""
    B, H, W, C_ = fmap.shape
    hidden = ca_w1.shape[0] if use_channel else 0
    feats = torch.empty(B, C_, device=fmap.device, dtype=torch.float32)
    saved = _f32buf(lib.dfv_attention_saved_floats(B, H, W, C_, hidden), fmap.device)
    check(lib.dfv_hybrid_attention_train_fwd(_ptr(fmap), _ptr(heat), _ptr(ca_w1), _ptr(ca_w2), _ptr(sa_w), _f32(feats), _f32(saved),
                                             dtype_code(fmap.dtype), B, H, W, C_, hidden, int(use_channel), int(use_spatial),
                                             _stream()))
    return feats, saved


def hybrid_attention_bwd(fmap, heat, ca_w1, ca_w2, sa_w, dfeatures, saved, use_channel=True, use_spatial=True):
    """Returns (dfmap, dheat | None, dca_w1, dca_w2, dsa_w)."""
    B, H, W, C_ = fmap.shape
    dev = fmap.device
    hidden = ca_w1.shape[0] if use_channel else 0
    dfmap = torch.empty_like(fmap)
    dheat = torch.empty(B, H, W, device=dev, dtype=torch.float32) if heat is not None else None
    dw1 = torch.zeros_like(ca_w1) if use_channel else None
    dw2 = torch.zeros_like(ca_w2) if use_channel else None
    dsa = torch.zeros(98, device=dev, dtype=torch.float32) if use_spatial else None
    ws = _f32buf(B * C_ + 2 * B * max(hidden, 1), dev)
    check(lib.dfv_hybrid_attention_bwd(_ptr(fmap), _ptr(heat), _ptr(ca_w1), _ptr(ca_w2), _ptr(sa_w), _f32(dfeatures), _f32(saved),
                                       _ptr(dfmap), _ptr(dheat), _ptr(dw1), _ptr(dw2), _ptr(dsa), _f32(ws), dtype_code(fmap.dtype),
                                       B, H, W, C_, hidden, int(use_channel), int(use_spatial), _stream()))
    return dfmap, dheat, dw1, dw2, dsa


def landmark_heatmap_train(landmarks, weights5, H, W, ref_size=224.0, sigma=1.5, group=0):
    """Returns (heat, raw, max_ws) -- the scratch buffers are what landmark_heatmap_bwd needs."""
    B, dev = landmarks.shape[0], landmarks.device
    heat = torch.empty(B, H, W, device=dev, dtype=torch.float32)
    raw = torch.empty(B * H * W, device=dev, dtype=torch.float32)
    mx = torch.empty(B, device=dev, dtype=torch.int32)
    check(lib.dfv_landmark_heatmap_fwd(_f32(landmarks), _f32(weights5), _f32(heat), _f32(raw), _ptr(mx), None, B, H, W, ref_size,
                                       sigma, group, _stream()))
    return heat, raw, mx


def landmark_heatmap_bwd(landmarks, weights5, raw, mx, dheat, H, W, ref_size=224.0, sigma=1.5, group=0):
    B = landmarks.shape[0]
    dw = torch.empty(5, device=landmarks.device, dtype=torch.float32)
    check(lib.dfv_landmark_heatmap_bwd(_f32(landmarks), _f32(weights5), _f32(raw), _ptr(mx), _f32(dheat), _f32(dw), B, H, W,
                                       ref_size, sigma, group, _stream()))
    return dw


def dropout_mask(shape, p, seed, device):
    out = torch.empty(shape, device=device, dtype=torch.float32)
    check(lib.dfv_dropout_mask(_f32(out), out.numel(), p, seed, _stream()))
    return out


# ------------------------------------------------------------------ side operators
def clip_aggregate(logits, frames_per_clip, threshold=0.5):
    """Video scoring rule of task.ipynb:434-442 over consecutive groups of `frames_per_clip` frames:
    returns (mean_logits [n_clips, n_classes], fake_prob [n_clips] = mean softmax[:, 1], labels [n_clips] int32)."""
    B, nc = logits.shape
    assert B % frames_per_clip == 0, "batch must hold whole clips"
    n = B // frames_per_clip
    dev = logits.device
    mean_logits = torch.empty(n, nc, device=dev, dtype=torch.float32)
    prob = torch.empty(n, device=dev, dtype=torch.float32)
    labels = torch.empty(n, device=dev, dtype=torch.int32)
    check(lib.dfv_clip_aggregate(_f32(logits), n, frames_per_clip, nc, _f32(mean_logits), _f32(prob), _ptr(labels), threshold, _stream()))
    return mean_logits, prob, labels


def global_avg_pool(x_nhwc):
    """[B, H, W, C] (fp32 / bf16) -> [B, C] fp32 mean over positions."""
    B, H, W, C_ = x_nhwc.shape
    out = torch.empty(B, C_, device=x_nhwc.device, dtype=torch.float32)
    check(lib.dfv_global_avg_pool(_ptr(x_nhwc), dtype_code(x_nhwc.dtype), _f32(out), B, H * W, C_, _stream()))
    return out


def l2_normalize(x, eps=1e-12):
    y = torch.empty_like(x)
    check(lib.dfv_l2_normalize(_f32(x), _f32(y), x.shape[0], x.shape[1], eps, _stream()))
    return y
