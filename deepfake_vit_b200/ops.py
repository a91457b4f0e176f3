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


def stem_conv(x_nchw, w_khwc, bias, out_dtype):
    B, _, H, W = x_nchw.shape
    C_ = bias.numel()
    Ho, Wo = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1
    y = torch.empty(B, Ho, Wo, C_, device=x_nchw.device, dtype=out_dtype)
    check(lib.dfv_stem_conv_fwd(_f32(x_nchw), _f32(w_khwc), _f32(bias), _ptr(y), dtype_code(out_dtype), B, H, W, C_, _stream()))
    return y


def dwconv(x, w_kkc, bias, kernel, stride, pad_lo, pad_hi, act=DFV_ACT_SILU, want_pool=True):
    """x: [B,H,W,C].  Returns (y [B,Ho,Wo,C], pool_partial [B,parts,C] fp32 or None)."""
    B, H, W, C_ = x.shape
    dt = dtype_code(x.dtype)
    Ho = (H + pad_lo + pad_hi - kernel) // stride + 1
    Wo = (W + pad_lo + pad_hi - kernel) // stride + 1
    y = torch.empty(B, Ho, Wo, C_, device=x.device, dtype=x.dtype)
    pool = None
    if want_pool:
        parts = lib.dfv_dwconv_pool_parts(dt, H, W, C_, kernel, stride, pad_lo, pad_hi)
        if parts <= 0:
            check(parts)
        pool = torch.empty(B, parts, C_, device=x.device, dtype=torch.float32)
    check(lib.dfv_dwconv_fwd(_ptr(x), _f32(w_kkc), _f32(bias), _ptr(y), _ptr(pool), dt, B, H, W, C_, kernel, stride,
                             pad_lo, pad_hi, act, _stream()))
    return y, pool


def se_gate(pool_partial, hw, w_reduce, b_reduce, w_expand_t, b_expand, gate_dtype=torch.float32):
    B, parts, C_ = pool_partial.shape
    sq = b_reduce.numel()
    gate = torch.empty(B, C_, device=pool_partial.device, dtype=gate_dtype)
    check(lib.dfv_se_gate_fwd(_f32(pool_partial), parts, 1.0 / hw, _f32(w_reduce), _f32(b_reduce), _f32(w_expand_t),
                              _f32(b_expand), _ptr(gate), dtype_code(gate_dtype), B, C_, sq, _stream()))
    return gate


def pw_gemm(a, w, bias, act=DFV_ACT_NONE, a_scale=None, rows_per_image=0, residual=None):
    """a: [..., K] (NHWC activations), w: [N, K].  Returns [..., N]."""
    K = a.shape[-1]
    N = w.shape[0]
    M = a.numel() // K
    assert w.shape[1] == K and w.dtype == a.dtype
    assert a_scale is None or a_scale.dtype == a.dtype, "the SE gate has the activation dtype"
    out = torch.empty(*a.shape[:-1], N, device=a.device, dtype=a.dtype)
    check(lib.dfv_pw_gemm_fwd(_ptr(a), _ptr(w), _f32(bias), _ptr(a_scale), rows_per_image, _ptr(residual), _ptr(out),
                              dtype_code(a.dtype), M, K, N, act, _stream()))
    return out


def landmark_heatmap(landmarks, weights5, H, W, ref_size=224.0, sigma=1.5, group=0, return_scaled=False):
    B = landmarks.shape[0]
    dev = landmarks.device
    heat = torch.empty(B, H, W, device=dev, dtype=torch.float32)
    raw = torch.empty(B * H * W, device=dev, dtype=torch.float32)
    mx = torch.empty(B, device=dev, dtype=torch.int32)
    scaled = torch.empty(B, 5, 2, device=dev, dtype=torch.float32) if return_scaled else None
    check(lib.dfv_landmark_heatmap_fwd(_f32(landmarks), _f32(weights5), _f32(heat), _f32(raw), _ptr(mx), _ptr(scaled),
                                       B, H, W, ref_size, sigma, group, _stream()))
    return (heat, scaled) if return_scaled else heat


def hybrid_attention(fmap, heat, ca_w1, ca_w2_t, sa_w, use_channel=True, use_spatial=True, return_gates=False):
    """fmap: [B,H,W,C]; heat: [B,H,W] fp32 or None.  Returns pooled features [B,C] fp32."""
    B, H, W, C_ = fmap.shape
    dev = fmap.device
    feats = torch.empty(B, C_, device=dev, dtype=torch.float32)
    cg = torch.empty(B, C_, device=dev, dtype=torch.float32) if return_gates and use_channel else None
    sg = torch.empty(B, H * W, device=dev, dtype=torch.float32) if return_gates and use_spatial else None
    hidden = ca_w1.shape[0] if use_channel else 0
    check(lib.dfv_hybrid_attention_fwd(_ptr(fmap), _ptr(heat), _ptr(ca_w1), _ptr(ca_w2_t), _ptr(sa_w), _f32(feats),
                                       _ptr(cg), _ptr(sg), dtype_code(fmap.dtype), B, H, W, C_, hidden,
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
    check(lib.dfv_mlp_head_fwd(_f32(features), pack.wp, pack.bp, C.cast(pack.dims, C.POINTER(C.c_int)), pack.n,
                               _f32(logits), B, _stream()))
    return logits


def combined_loss(logits, targets, features, class_weights, w_ce, w_focal, w_con, want_grad=True):
    """Returns (losses[4] = ce, focal, contrastive, total; has_contrastive; dlogits; dfeatures)."""
    B, Cn = logits.shape
    D = features.shape[1] if features is not None else 0
    dev = logits.device
    losses = torch.empty(4, device=dev, dtype=torch.float32)
    dlogits = torch.empty_like(logits) if want_grad else None
    dfeat = torch.empty_like(features) if (want_grad and features is not None) else None
    has = C.c_int(0)
    assert targets.dtype == torch.int64
    check(lib.dfv_combined_loss_fwd_bwd(_f32(logits), _ptr(targets), _ptr(features), _ptr(class_weights), w_ce, w_focal,
                                        w_con, _f32(losses), _ptr(dlogits), _ptr(dfeat), B, Cn, D, C.byref(has),
                                        _stream()))
    return losses, bool(has.value), dlogits, dfeat
