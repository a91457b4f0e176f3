"""Per-kernel parity tests (B200 only): each libdfvit operator against a plain PyTorch fp32
statement of the same reference op, through the C ABI (deepfake_vit_b200.ops -> ctypes).

Tolerances: fp32 mode 1e-5 relative L2 (different summation order only); bf16 mode 1e-2
relative L2 against the fp32 result computed from the same bf16-rounded inputs.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    import deepfake_vit_b200 as d
    return d.ops


def _blocks():
    import deepfake_vit_b200 as d
    return d._lib.b4_blocks()


def silu(x):
    return x * torch.sigmoid(x)


# ------------------------------------------------------------------------------- stem
@pytest.mark.parametrize("shape", [(2, 64, 64), (1, 381, 379), (3, 224, 224)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stem(ops, shape, dtype):
    B, H, W = shape
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(48, 3, 3, 3, generator=g) * 0.3
    bias = torch.randn(48, generator=g) * 0.1
    ref = silu(F.conv2d(F.pad(x, (0, 1, 0, 1)), w, bias, stride=2))            # static pad (0,1,0,1)
    y = ops.stem_conv(x.to(DEV), w.permute(2, 3, 1, 0).contiguous().to(DEV), bias.to(DEV), dtype)
    assert y.shape == (B, ref.shape[2], ref.shape[3], 48)
    r = rel(y.float().permute(0, 3, 1, 2), ref)
    assert r < (2e-6 if dtype == torch.float32 else 8e-3), r   # bf16: images and weights are bf16 tensor-core operands, as under autocast


# ------------------------------------------------------------------------------- depthwise
def _dw_cases():
    seen, out = set(), []
    size = 190
    for b in _blocks():
        key = (b.kernel, b.stride, b.c_mid, size)
        if key not in seen:
            seen.add(key)
            out.append((b.kernel, b.stride, b.c_mid, size, b.pad_lo, b.pad_hi))
        size = (size + b.stride - 1) // b.stride
    return out


def _dw_ref(x_nhwc, w, bias, k, s, plo, phi):
    x = x_nhwc.permute(0, 3, 1, 2)
    C = x.shape[1]
    y = F.conv2d(F.pad(x, (plo, phi, plo, phi)), w.view(C, 1, k, k), bias, stride=s, groups=C)
    return silu(y)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dwconv_all_b4_shapes(ops, dtype):
    """All 14 distinct (k, stride, C, H) depthwise shapes of B4 at 380, batch 2."""
    g = torch.Generator().manual_seed(2)
    for (k, s, C, H, plo, phi) in _dw_cases():
        x = torch.randn(2, H, H, C, generator=g).to(dtype)
        w = torch.randn(C, k * k, generator=g) * 0.2
        bias = torch.randn(C, generator=g) * 0.1
        if dtype == torch.bfloat16:
            w = w.bfloat16().float()      # the bf16 kernel stages weights in bf16 (as autocast does)
        ref = _dw_ref(x.float(), w, bias, k, s, plo, phi)
        y, pool = ops.dwconv(x.to(DEV), w.t().contiguous().to(DEV), bias.to(DEV), k, s, plo, phi)
        assert y.shape == (2, ref.shape[2], ref.shape[3], C), (k, s, C, H)
        r = rel(y.float().permute(0, 3, 1, 2), ref)
        assert r < (2e-6 if dtype == torch.float32 else 5e-3), (k, s, C, H, r)
        pooled = pool.sum(1).cpu() / (ref.shape[2] * ref.shape[3])
        rp = rel(pooled, ref.mean((2, 3)))
        assert rp < (1e-5 if dtype == torch.float32 else 5e-3), (k, s, C, H, rp)


@pytest.mark.parametrize("case", [(3, 1, 24, 7, 9, 1, 1), (5, 1, 40, 13, 5, 2, 2), (5, 2, 16, 9, 11, 1, 2),
                                  (3, 2, 8, 33, 17, 0, 1), (5, 1, 72, 50, 50, 2, 2)])
def test_dwconv_ragged_shapes(ops, case):
    """Non-square / tiny / non-tile-multiple maps, fp32, no activation."""
    k, s, C, H, W, plo, phi = case
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, H, W, C, generator=g)
    w = torch.randn(C, k * k, generator=g)
    bias = torch.randn(C, generator=g)
    xr = x.permute(0, 3, 1, 2)
    ref = F.conv2d(F.pad(xr, (plo, phi, plo, phi)), w.view(C, 1, k, k), bias, stride=s, groups=C)
    y, pool = ops.dwconv(x.to(DEV), w.t().contiguous().to(DEV), bias.to(DEV), k, s, plo, phi, act=0)
    assert rel(y.permute(0, 3, 1, 2), ref) < 2e-6
    assert rel(pool.sum(1), ref.sum((2, 3))) < 1e-5


def test_dwconv_and_stem_full_size_every_element(ops):
    """Benchmark-size (batch 256) launches compared element by element with torch on the GPU, twice each: the
    persistent tile loops, pipeline wrap-around and pool-slot flushes only run long at this size, and a sporadic
    race there is invisible to a norm-based check (see test_pw_gemm_full_size_every_element)."""
    g = torch.Generator(device=DEV).manual_seed(62)
    B = 256
    for (k, s, C, H, plo, phi) in ((3, 1, 192, 95, 1, 1), (5, 1, 336, 48, 2, 2), (3, 2, 336, 48, 0, 1), (5, 2, 960, 24, 1, 2),
                                   (5, 1, 1632, 12, 2, 2), (3, 1, 2688, 12, 1, 1)):
        x = torch.randn(B, H, H, C, device=DEV, generator=g).bfloat16()
        w = (torch.randn(C, k * k, device=DEV, generator=g) * 0.2).bfloat16().float()
        bias = torch.randn(C, device=DEV, generator=g) * 0.1
        ref = _dw_ref(x.float(), w, bias, k, s, plo, phi).permute(0, 2, 3, 1)
        for rep in range(2):
            y, pool = ops.dwconv(x, w.t().contiguous(), bias, k, s, plo, phi)
            bad = ((y.float() - ref).abs() > 0.03 * (ref.abs() + 1.0)).sum().item()
            assert bad == 0, (k, s, C, H, rep, bad)
            ps = pool.sum(1)
            rs = ref.sum((1, 2))
            assert ((ps - rs).abs() <= 0.02 * (rs.abs() + ref.abs().sum((1, 2)) * 0.01 + 1.0)).all(), (k, s, C, H, rep)
        del ref, x
    x = torch.randn(B, 3, 380, 380, device=DEV, generator=g)
    w = torch.randn(48, 3, 3, 3, device=DEV, generator=g) * 0.3
    bias = torch.randn(48, device=DEV, generator=g) * 0.1
    ref = silu(F.conv2d(F.pad(x.bfloat16().float(), (0, 1, 0, 1)), w.bfloat16().float(), bias, stride=2)).permute(0, 2, 3, 1)
    for rep in range(2):
        y = ops.stem_conv(x, w.permute(2, 3, 1, 0).contiguous(), bias, torch.bfloat16)
        bad = ((y.float() - ref).abs() > 0.03 * (ref.abs() + 1.0)).sum().item()
        assert bad == 0, ("stem", rep, bad)


# ------------------------------------------------------------------------------- SE gate
def test_se_gate(ops):
    g = torch.Generator().manual_seed(4)
    for C, sq, parts in ((48, 12, 7), (960, 40, 2), (2688, 112, 1)):
        pool = torch.randn(5, parts, C, generator=g)
        w1, b1 = torch.randn(sq, C, generator=g) * 0.1, torch.randn(sq, generator=g)
        w2, b2 = torch.randn(C, sq, generator=g) * 0.1, torch.randn(C, generator=g)
        hw = 37.0
        m = pool.sum(1) / hw
        ref = torch.sigmoid(silu(m @ w1.t() + b1) @ w2.t() + b2)
        out = ops.se_gate(pool.to(DEV), hw, w1.to(DEV), b1.to(DEV), w2.t().contiguous().to(DEV), b2.to(DEV))
        assert rel(out, ref) < 1e-5
        out16 = ops.se_gate(pool.to(DEV), hw, w1.to(DEV), b1.to(DEV), w2.t().contiguous().to(DEV), b2.to(DEV),
                            torch.bfloat16)
        assert out16.dtype == torch.bfloat16 and rel(out16.float(), ref) < 3e-3


# ------------------------------------------------------------------------------- pointwise GEMM
def _gemm_cases():
    seen = []
    for b in _blocks():
        if b.has_expand:
            seen.append((b.c_in, b.c_mid))
        seen.append((b.c_mid, b.c_out))
    seen.append((448, 1792))
    return sorted(set(seen))


def _gemm_ref(a, w, bias, act, scale, rpi, res):
    a = a.float()
    if scale is not None:
        a = a * scale.float().repeat_interleave(rpi, 0)
        if w.dtype == torch.bfloat16:
            a = a.bfloat16().float()      # the kernel rounds the gated operand to bf16
    y = a @ w.float().t() + bias
    if act:
        y = silu(y)
    if res is not None:
        y = y + res.float()
    return y


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pw_gemm_all_b4_shapes(ops, dtype):
    """Every distinct (K, N) of B4's 63 1x1 convolutions, ragged M, plain / SiLU / gated+residual."""
    g = torch.Generator().manual_seed(5)
    tol = 2e-6 if dtype == torch.float32 else 4e-3
    for (K, N) in _gemm_cases():
        rpi = 36                       # "rows per image" for the SE gate
        M = rpi * 11                   # 396 rows = 3 full 128-row tiles + a 12-row tail
        a = torch.randn(M, K, generator=g).to(dtype)
        w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dtype)
        bias = torch.randn(N, generator=g) * 0.1
        scale = torch.rand(M // rpi, K, generator=g).to(dtype)
        res = torch.randn(M, N, generator=g).to(dtype)
        for act, use_scale, use_res in ((1, False, False), (0, True, True), (0, True, False)):
            ref = _gemm_ref(a, w, bias, act, scale if use_scale else None, rpi, res if use_res else None)
            out = ops.pw_gemm(a.to(DEV), w.to(DEV), bias.to(DEV), act, scale.to(DEV) if use_scale else None,
                              rpi if use_scale else 0, res.to(DEV) if use_res else None)
            r = rel(out.float(), ref)
            assert r < tol, (K, N, act, use_scale, use_res, r)


def test_pw_gemm_tcgen05_matches_fp32_large(ops):
    """bf16 tcgen05 path vs an fp32 matmul of the same bf16 operands on a multi-wave problem (persistent loop, pipeline wrap)."""
    g = torch.Generator().manual_seed(6)
    for (M, K, N) in ((128 * 700 + 5, 32, 192), (128 * 160, 672, 112), (144 * 64, 1632, 272), (144 * 40, 448, 1792)):
        a = torch.randn(M, K, generator=g).bfloat16().to(DEV)
        w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16().to(DEV)
        bias = (torch.randn(N, generator=g) * 0.1).to(DEV)
        out_tc = ops.pw_gemm(a, w, bias, 1)
        r = a.float() @ w.float().t() + bias
        ref = r * torch.sigmoid(r)
        torch.cuda.synchronize()
        assert rel(out_tc.float(), ref) < 3e-3, (M, K, N)


def test_pw_gemm_full_size_every_element(ops):
    """Benchmark-size launches (batch 256) checked element by element, three times each.  Regression: with N = 336 and
    192-column tiles the last column part of every second tile lies beyond N; its epilogue warps skipped those tiles but
    picked their staging buffer by tile parity, so a TMA store still reading the buffer could be overwritten -- a few
    thousand wrong values per launch in the first tiles of some CTAs, invisible to sampled / norm-based checks.
    Also covers the weight-stationary plan (small K, several N tiles) and the streaming plan (large K) at full size."""
    g = torch.Generator(device=DEV).manual_seed(61)
    for (M, K, N, act, gated, rpi) in ((589824, 56, 336, 1, False, 0), (147456, 160, 960, 1, False, 0),
                                       (36864, 1632, 272, 0, True, 144), (2310400 // 4, 96, 576, 1, False, 0),
                                       (577600, 96, 96, 0, True, 9025 // 1), (147461, 672, 112, 0, True, 147461),
                                       (145 * 255, 2688, 448, 0, True, 145), (73728, 272, 448, 1, False, 0), (36864 + 40, 272, 1632, 1, False, 0),
                                       (36864, 448, 1792, 1, False, 0)):
        a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
        w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
        bias = torch.randn(N, device=DEV, generator=g) * 0.1
        sc = torch.rand(M // rpi, K, device=DEV, generator=g).bfloat16() if gated else None
        res = torch.randn(M, N, device=DEV, generator=g).bfloat16() if gated else None     # project GEMMs carry the skip
        ref = torch.empty(M, N, device=DEV)
        for i in range(0, M, 65536):
            av = a[i:i + 65536]
            if gated:
                av = av * sc[torch.arange(i, min(i + 65536, M), device=DEV) // rpi]
            r = av.float() @ w.float().t() + bias
            r = r * torch.sigmoid(r) if act else r
            ref[i:i + 65536] = r + res[i:i + 65536].float() if gated else r
        # planner's choice, forced single-CTA streaming plans, forced CTA-pair plans (cta_group::2; ragged M: the last pair
        # has a phantom tile) -- per-call tuning argument (weight_stationary, N tile, cluster)
        # ... and shared-A plans (two N tiles accumulate from one A stage, 4th field) where a gated shape has exactly two N
        # tiles, against the same CTA-pair plans without sharing
        for forced in (None, (0, 192, -1), (0, 128, -1), (0, 128, 2), (0, 0, 2), (0, 0, 2, 1), (0, 0, 2, -1), (0, 192, 2, 1), (0, 0, 0, 0, -1), (0, 0, -1, 0, -1)):     # 5th field -1: N tiles walked in the same order by every CTA
            for rep in range(3):
                y = ops.pw_gemm(a, w, bias, act, sc, rpi, res, tuning=forced)
                bad = ((y.float() - ref).abs() > 0.03 * (ref.abs() + 1.0)).sum().item()
                assert bad == 0, (M, K, N, forced, rep, bad)


def test_pw_gemm_rejects_bad_shapes(ops):
    import deepfake_vit_b200 as d
    a = torch.zeros(16, 12, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(8, 12, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(d._lib.DfvError):
        ops.pw_gemm(a, w, torch.zeros(8, device=DEV))      # bf16 (TMA / tcgen05) needs K % 8 == 0
    out = ops.pw_gemm(a.float() + 1, w.float() + 1, torch.zeros(8, device=DEV))   # the fp32 SIMT path takes any K, N
    assert torch.allclose(out, torch.full((16, 8), 12.0, device=DEV))


# ------------------------------------------------------------------------------- attention
def test_landmark_heatmap_bit_exact_coords(ops):
    from oracle import refmodel
    g = torch.Generator().manual_seed(7)
    for (B, H, W, size) in ((8, 12, 12, 380), (5, 7, 7, 224), (3, 9, 12, 300)):
        lm = torch.rand(B, 5, 2, generator=g) * size
        la = refmodel.LandmarkAttention()
        with torch.no_grad():
            la.attention_weights.copy_(torch.tensor([1.0, 0.8, 1.2, 0.9, 1.1]))
            ref = la._create_attention_map(lm, (H, W), "cpu")
        heat, scaled = ops.landmark_heatmap(lm.to(DEV), la.attention_weights.detach().to(DEV), H, W, return_scaled=True)
        sx, sy = W / 224.0, H / 224.0
        exp = lm.clone()
        exp[:, :, 0] *= sx
        exp[:, :, 1] *= sy
        assert torch.equal(scaled.cpu(), exp), "scaled landmark coordinates must be bit-exact"
        assert (heat.cpu() - ref[:, 0]).abs().max().item() < 5e-7
        # per-image groups == running the reference one image at a time
        heat1 = ops.landmark_heatmap(lm.to(DEV), la.attention_weights.detach().to(DEV), H, W, group=1)
        with torch.no_grad():
            ref1 = torch.cat([la._create_attention_map(lm[i:i + 1], (H, W), "cpu") for i in range(B)])
        assert (heat1.cpu() - ref1[:, 0]).abs().max().item() < 5e-7


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("flags", [(True, True, True), (False, True, True), (True, False, True), (True, True, False),
                                   (False, False, False)])
def test_hybrid_attention_and_pool(ops, dtype, flags):
    from oracle import refmodel
    use_lm, use_ch, use_sp = flags
    torch.manual_seed(8)
    B, C, H, W = 3, 1792, 12, 12
    att = refmodel.HybridAttention(C, use_landmark=True, use_spatial=True, use_channel=True).eval()
    x = torch.randn(B, C, H, W).to(dtype)
    lm = torch.rand(B, 5, 2) * 380
    with torch.no_grad():
        y = x.float()
        if use_lm:
            y = att.landmark_attn(y, lm)
        if use_ch:
            y = att.channel_attn(y)
        if use_sp:
            y = att.spatial_attn(y)
        ref = F.adaptive_avg_pool2d(y, 1).flatten(1)
    heat = ops.landmark_heatmap(lm.to(DEV), att.landmark_attn.attention_weights.detach().to(DEV), H, W) if use_lm else None
    out = ops.hybrid_attention(x.permute(0, 2, 3, 1).contiguous().to(DEV), heat,
                               att.channel_attn.fc[0].weight.detach().to(DEV).contiguous(),
                               att.channel_attn.fc[2].weight.detach().t().contiguous().to(DEV),
                               att.spatial_attn.conv.weight.detach().reshape(-1).to(DEV), use_ch, use_sp)
    assert rel(out, ref) < 2e-5


def test_mlp_head(ops):
    from oracle import calibrate, refmodel
    m = calibrate.build(refmodel.get_oracle(), "default")
    with torch.no_grad():
        for bn in (m.classifier[1], m.classifier[5], m.classifier[9]):
            bn.running_mean.normal_(0, 0.1)
            bn.running_var.uniform_(0.5, 1.5)
            bn.weight.uniform_(0.8, 1.2)
            bn.bias.normal_(0, 0.1)
    feats = torch.randn(7, 1792) * 0.5
    with torch.no_grad():
        ref = m.classifier(feats)
    import deepfake_vit_b200 as d
    dm = d.DeepfakeDetectionModel(**refmodel.MODEL_CONFIG)
    dm.load_state_dict(m.state_dict())
    dm = dm.to(DEV).eval()
    pk = dm._pack(torch.float32, torch.device(DEV, torch.cuda.current_device()))
    out = ops.mlp_head(feats.to(DEV), pk.head)
    assert rel(out, ref) < 1e-5


# ------------------------------------------------------------------------------- loss
@pytest.mark.parametrize("B", [1, 2, 5, 8, 64])
@pytest.mark.parametrize("cw", [None, [1.0, 1.5]])
def test_combined_loss_values_and_grads(B, cw):
    from oracle import refmodel
    import deepfake_vit_b200 as d
    g = torch.Generator().manual_seed(9)
    logits = (torch.randn(B, 2, generator=g) * 2).requires_grad_()
    feats = (torch.randn(B, 1792, generator=g) * 0.05).requires_grad_()
    y = torch.randint(0, 2, (B,), generator=g)
    cwt = torch.tensor(cw) if cw else None
    weights = {"ce": 1.0, "focal": 0.5, "contrastive": 0.2}
    ref = refmodel.CombinedLoss(weights, cwt)(logits, y, feats)
    ref["total"].backward()
    lo = logits.detach().to(DEV).requires_grad_()
    fe = feats.detach().to(DEV).requires_grad_()
    out = d.CombinedLoss(weights, cwt)(lo, y.to(DEV), fe)
    assert set(out) == set(ref)
    for k in ref:
        assert abs(out[k].item() - float(ref[k])) < 2e-6 * max(1.0, abs(float(ref[k]))), k
    out["total"].backward()
    assert rel(lo.grad, logits.grad) < 1e-5
    if B >= 2:
        assert rel(fe.grad, feats.grad) < 1e-5
    # Evaluator's call: no features -> no contrastive key (evaluator.py:96)
    out2 = d.CombinedLoss(weights, cwt)(lo.detach(), y.to(DEV))
    ref2 = refmodel.CombinedLoss(weights, cwt)(logits.detach(), y)
    assert set(out2) == set(ref2) == {"ce", "focal", "total"}
    assert abs(out2["total"].item() - float(ref2["total"])) < 2e-6


# ------------------------------------------------------------------------------- scratch sizes
def test_scratch_buffers_are_not_overrun(ops):
    """Every wrapper that hands a kernel a scratch / fp32 output buffer sized by a dfv_*_scratch_floats() function runs
    with sentinel bands around the buffer (ops.guard_mode): odd batches, K-split and single-slice layers, with and without
    the optional stages.  (compute-sanitizer is not available on the pool; this is the bounds check of our own.)"""
    g = torch.Generator(device=DEV).manual_seed(81)
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g)
    ops.guard_mode(True)
    try:
        for (B, C, sq, parts) in ((1, 48, 12, 3), (5, 336, 14, 2), (17, 1632, 68, 1), (64, 2688, 112, 1), (33, 200, 9, 4)):
            pool = rn(B, parts, C)
            w1, b1, w2t, b2 = rn(sq, C) * 0.1, rn(sq) * 0.1, rn(sq, C) * 0.1, rn(C) * 0.1
            gate = ops.se_gate(pool, 144, w1, b1, w2t, b2, torch.float32)
            pooled = pool.sum(1) / 144
            h = pooled @ w1.t() + b1
            ref = torch.sigmoid((h * torch.sigmoid(h)) @ w2t + b2)
            assert rel(gate, ref) < 1e-5, (B, C, sq, parts)
            gate_t, pooled_t, h1_t, g32 = ops.se_train_fwd(pool, 144, w1, b1, w2t.t().contiguous(), b2, torch.float32)
            assert rel(gate_t, ref) < 1e-5 and rel(pooled_t, pooled) < 1e-6 and rel(h1_t, h) < 1e-5 and torch.equal(gate_t, g32)
        for (B, H, W, C, hid, uc, us) in ((1, 12, 12, 1792, 112, True, True), (7, 5, 7, 256, 16, True, False), (19, 12, 12, 1792, 112, False, True),
                                          (4, 3, 3, 64, 8, True, True)):
            fmap = rn(B, H, W, C).bfloat16()
            heat = torch.rand(B, H, W, device=DEV, generator=g)
            ops.hybrid_attention(fmap, heat, rn(hid, C) * 0.1, rn(hid, C) * 0.1, rn(98) * 0.1, uc, us)
        for (B, dims) in ((1, (1792, 512, 256, 2)), (37, (1792, 512, 256, 2)), (9, (300, 2)), (64, (2048, 1000, 7))):
            pack = ops.HeadPack([rn(dims[i], dims[i + 1]) * 0.05 for i in range(len(dims) - 1)], [rn(dims[i + 1]) * 0.1 for i in range(len(dims) - 1)])
            ops.mlp_head(rn(B, dims[0]), pack)
        assert ops.guards_intact()
    finally:
        ops.guard_mode(False)
