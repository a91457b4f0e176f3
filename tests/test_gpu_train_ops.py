"""Per-kernel parity tests of the TRAINING path (B200 only): each libdfvit training operator against
PyTorch autograd of the reference op it replaces (fp32 on CPU), through the C ABI.

Tolerances: fp32 mode 2e-5 relative L2 (summation order only); bf16 mode 2e-2 relative L2 against the
fp32 result computed from the same bf16-rounded inputs.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"
DT = [torch.float32, torch.bfloat16]


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def tol(dtype, f32=2e-5, bf16=2e-2):
    return f32 if dtype == torch.float32 else bf16


@pytest.fixture(scope="module")
def ops():
    import deepfake_vit_b200 as d
    return d.ops


def rnd(dtype, *shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g) * scale
    return x.to(dtype).float()          # values exactly representable in `dtype`


def silu(x):
    return x * torch.sigmoid(x)


ACTS = {0: lambda u: u, 1: silu, 2: torch.relu}


# ------------------------------------------------------------------------------- BatchNorm forward
@pytest.mark.parametrize("shape", [(4, 9, 9, 24), (3, 7, 5, 336), (2, 4, 4, 2688), (16, 1, 1, 512), (5, 31, 17, 48)])
@pytest.mark.parametrize("dtype", DT)
def test_bn_stats_and_act(ops, shape, dtype):
    B, H, W, C = shape
    x = (rnd(dtype, B, H, W, C, seed=1) * 1.7 + 0.4).to(dtype).float()
    gamma, beta = rnd(torch.float32, C, seed=2) * 0.3 + 1.0, rnd(torch.float32, C, seed=3) * 0.2
    bn = torch.nn.BatchNorm2d(C, eps=1e-3, momentum=0.01)
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
        bn.running_mean.uniform_(-1, 1)
        bn.running_var.uniform_(0.5, 2)
    rm, rv = bn.running_mean.clone().to(DEV), bn.running_var.clone().to(DEV)
    bn.train()
    ref = silu(bn(x.permute(0, 3, 1, 2))).permute(0, 2, 3, 1)
    xd = x.to(DEV, dtype)
    mean, invstd = ops.bn_stats(xd, 1e-3, 0.01, rm, rv)
    xs = x.reshape(-1, C)
    assert rel(mean, xs.mean(0)) < 1e-5
    assert rel(invstd, 1 / torch.sqrt(xs.var(0, unbiased=False) + 1e-3)) < 1e-5
    assert rel(rm, bn.running_mean) < 1e-5 and rel(rv, bn.running_var) < 1e-5
    y, pool = ops.bn_act(xd, mean, invstd, gamma.to(DEV), beta.to(DEV), act=1, want_pool=True)
    assert rel(y, ref) < tol(dtype, 1e-5, 6e-3)
    assert rel(pool.sum(1) / (H * W), ref.mean((1, 2))) < tol(dtype, 1e-5, 3e-3)


@pytest.mark.parametrize("dtype", DT)
def test_bn_act_residual_rowscale_mask(ops, dtype):
    B, H, W, C = 4, 6, 6, 56
    x, res = rnd(dtype, B, H, W, C, seed=1), rnd(dtype, B, H, W, C, seed=2)
    gamma, beta = rnd(torch.float32, C, seed=3) + 1.0, rnd(torch.float32, C, seed=4)
    rs = torch.tensor([0.0, 1.25, 1.25, 0.0])
    mask = (torch.rand(B, H, W, C, generator=torch.Generator().manual_seed(5)) > 0.4).float() / 0.6
    xd = x.to(DEV, dtype)
    mean, invstd = ops.bn_stats(xd, 1e-3)
    xs = x.reshape(-1, C)
    u = (x - xs.mean(0)) / torch.sqrt(xs.var(0, unbiased=False) + 1e-3) * gamma + beta
    y = ops.bn_act(xd, mean, invstd, gamma.to(DEV), beta.to(DEV), act=0, rowscale=rs.to(DEV), residual=res.to(DEV, dtype))
    assert rel(y, u * rs.view(B, 1, 1, 1) + res) < tol(dtype, 1e-5, 6e-3)
    y = ops.bn_act(xd, mean, invstd, gamma.to(DEV), beta.to(DEV), act=2, mask=mask.to(DEV))
    assert rel(y, torch.relu(u) * mask) < tol(dtype, 1e-5, 6e-3)


# ------------------------------------------------------------------------------- BatchNorm backward
@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("shape", [(4, 9, 9, 24), (3, 5, 7, 336), (8, 1, 1, 128), (2, 3, 3, 2688)])
@pytest.mark.parametrize("dtype", DT)
def test_act_bn_bwd(ops, shape, act, dtype):
    B, H, W, C = shape
    x = (rnd(dtype, B, H, W, C, seed=1) * 1.3 + 0.2).to(dtype).float().requires_grad_(True)
    gamma = (rnd(torch.float32, C, seed=2) * 0.3 + 1.0).requires_grad_(True)
    beta = (rnd(torch.float32, C, seed=3) * 0.3).requires_grad_(True)
    g = rnd(dtype, B, H, W, C, seed=4)
    y = ACTS[act](F.batch_norm(x.permute(0, 3, 1, 2), None, None, gamma, beta, True, 0.0, 1e-3)).permute(0, 2, 3, 1)
    y.backward(g)
    xd = x.detach().to(DEV, dtype)
    mean, invstd = ops.bn_stats(xd, 1e-3)
    dx, dgamma, dbeta = ops.act_bn_bwd(g.to(DEV, dtype), xd, mean, invstd, gamma.detach().to(DEV), beta.detach().to(DEV), act=act)
    assert rel(dx, x.grad) < tol(dtype, 5e-5, 2e-2)
    assert rel(dgamma, gamma.grad) < tol(dtype, 2e-5, 1e-2)
    assert rel(dbeta, beta.grad) < tol(dtype, 2e-5, 1e-2)


@pytest.mark.parametrize("dtype", DT)
def test_act_bn_bwd_gate_dpool_rowscale_mask(ops, dtype):
    """The full prologue: gin = (g * gate + dpool / HW) * rowscale * mask, as the depthwise-output gradient needs it."""
    B, H, W, C = 4, 6, 5, 48
    x = (rnd(dtype, B, H, W, C, seed=1) + 0.1).to(dtype).float().requires_grad_(True)
    gamma = (rnd(torch.float32, C, seed=2) * 0.3 + 1.0).requires_grad_(True)
    beta = (rnd(torch.float32, C, seed=3) * 0.3).requires_grad_(True)
    g = rnd(dtype, B, H, W, C, seed=4)
    gate = torch.sigmoid(rnd(dtype, B, C, seed=5)).to(dtype).float()
    dpool = rnd(torch.float32, B, C, seed=6)
    rs = torch.tensor([1.25, 0.0, 1.25, 1.25])
    mask = (torch.rand(B, H, W, C, generator=torch.Generator().manual_seed(7)) > 0.3).float() / 0.7
    y = silu(F.batch_norm(x.permute(0, 3, 1, 2), None, None, gamma, beta, True, 0.0, 1e-3)).permute(0, 2, 3, 1)
    gin = (g * gate.view(B, 1, 1, C) + dpool.view(B, 1, 1, C) / (H * W)) * rs.view(B, 1, 1, 1) * mask
    y.backward(gin)
    xd = x.detach().to(DEV, dtype)
    mean, invstd = ops.bn_stats(xd, 1e-3)
    dx, dgamma, dbeta = ops.act_bn_bwd(g.to(DEV, dtype), xd, mean, invstd, gamma.detach().to(DEV), beta.detach().to(DEV), act=1,
                                       gate=gate.to(DEV, dtype), dpool=dpool.to(DEV), inv_hw=1.0 / (H * W), rowscale=rs.to(DEV),
                                       mask=mask.to(DEV))
    assert rel(dx, x.grad) < tol(dtype, 5e-5, 2e-2)
    assert rel(dgamma, gamma.grad) < tol(dtype, 2e-5, 1e-2)
    assert rel(dbeta, beta.grad) < tol(dtype, 2e-5, 1e-2)


# ------------------------------------------------------------------------------- 1x1 conv weight gradient
@pytest.mark.parametrize("mkn", [(4 * 81, 24, 144), (2 * 49, 960, 160), (777, 272, 1632), (64, 1792, 512), (16, 32, 2), (3 * 36, 48, 24),
                                 (40000, 32, 192), (30011, 192, 32), (9000, 448, 2688), (5000, 672, 112)])
@pytest.mark.parametrize("dtype", DT)
def test_pw_wgrad(ops, mkn, dtype):
    M, K, N = mkn
    if dtype == torch.bfloat16 and (K % 8 or N % 8):
        pytest.skip("bf16 tensors are channel-vectorised")
    a, g = rnd(dtype, M, K, seed=1), rnd(dtype, M, N, seed=2)
    dw = ops.pw_wgrad(g.to(DEV, dtype), a.to(DEV, dtype))
    assert rel(dw, g.t() @ a) < 2e-5


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape", [(3, 36, 336, 56), (5, 2304, 336, 56), (7, 361, 32, 192), (4, 100, 1632, 272)])
def test_pw_wgrad_gated(ops, dtype, shape):
    B, HW, K, N = shape
    a, g = rnd(dtype, B, HW, K, seed=1), rnd(dtype, B, HW, N, seed=2)
    gate = torch.sigmoid(rnd(dtype, B, K, seed=3)).to(dtype).float()
    ag = (a * gate.view(B, 1, K)).to(dtype).float()          # the forward operand is rounded to the activation dtype
    dw = ops.pw_wgrad(g.to(DEV, dtype), a.to(DEV, dtype), gate.to(DEV, dtype), HW)
    assert rel(dw, g.reshape(-1, N).t() @ ag.reshape(-1, K)) < 2e-5 * (1 if dtype == torch.float32 else 2)


# ------------------------------------------------------------------------------- depthwise backward
DW_CASES = [(3, 1, 1, 1, 48, 20, 20), (3, 2, 0, 1, 144, 22, 22), (5, 2, 2, 2, 192, 19, 19), (5, 1, 2, 2, 336, 12, 12),
            (5, 2, 1, 2, 960, 8, 8), (3, 1, 1, 1, 2688, 4, 4), (3, 2, 0, 1, 336, 9, 11)]


@pytest.mark.parametrize("case", DW_CASES)
@pytest.mark.parametrize("dtype", DT)
def test_dwconv_backward(ops, case, dtype):
    k, s, pl, ph, C, H, W = case
    B = 3
    x = rnd(dtype, B, C, H, W, seed=1).requires_grad_(True)
    w = (rnd(torch.float32, C, 1, k, k, seed=2) * 0.3).requires_grad_(True)
    y = F.conv2d(F.pad(x, (pl, ph, pl, ph)), w, None, stride=s, groups=C)
    g = rnd(dtype, *y.shape, seed=3)
    y.backward(g)
    gd = g.permute(0, 2, 3, 1).contiguous().to(DEV, dtype)
    xd = x.detach().permute(0, 2, 3, 1).contiguous().to(DEV, dtype)
    w_kkc = w.detach().view(C, k * k).t().contiguous().to(DEV)
    dx = ops.dwconv_dgrad(gd, w_kkc, H, W, k, s, pl, ph)
    assert rel(dx.float().permute(0, 3, 1, 2), x.grad) < tol(dtype, 1e-5, 6e-3)
    dw = ops.dwconv_wgrad(gd, xd, k, s, pl, ph)
    assert rel(dw.t().reshape(C, 1, k, k), w.grad) < 2e-5
    if s == 1:      # the sequencer's stride-1 route: forward TMA kernel, flipped taps, mirrored pads
        wf = w.detach().view(C, k * k).flip(1).t().contiguous().to(DEV)
        zero = torch.zeros(C, device=DEV)
        dx2, _ = ops.dwconv(gd, wf, zero, k, 1, k - 1 - pl, k - 1 - ph, act=0, want_pool=False)
        assert dx2.shape == dx.shape
        assert rel(dx2.float().permute(0, 3, 1, 2), x.grad) < tol(dtype, 1e-5, 6e-3)


@pytest.mark.parametrize("shape", [(2, 64, 64), (3, 97, 95)])
@pytest.mark.parametrize("dtype", DT)
def test_stem_wgrad(ops, shape, dtype):
    B, H, W = shape
    x = rnd(torch.float32, B, 3, H, W, seed=1)
    w = (rnd(torch.float32, 48, 3, 3, 3, seed=2) * 0.3).requires_grad_(True)
    # bf16 path: the forward stem feeds bf16(x) to the tensor cores (as autocast does), so its weight gradient is g^T bf16(x)
    xr = x.to(dtype).float()
    y = F.conv2d(F.pad(xr, (0, 1, 0, 1)), w, None, stride=2)
    g = rnd(dtype, *y.shape, seed=3)
    y.backward(g)
    dw = ops.stem_wgrad(g.permute(0, 2, 3, 1).contiguous().to(DEV, dtype), x.to(DEV))
    assert rel(dw, w.grad) < 2e-5


# ------------------------------------------------------------------------------- squeeze-excite
@pytest.mark.parametrize("shape", [(4, 5, 5, 48, 12), (3, 6, 4, 672, 28), (2, 3, 3, 2688, 112)])
@pytest.mark.parametrize("dtype", DT)
def test_se_train(ops, shape, dtype):
    B, H, W, C, sq = shape
    d = rnd(dtype, B, H, W, C, seed=1).requires_grad_(True)
    w1 = (rnd(torch.float32, sq, C, seed=2) * 0.2).requires_grad_(True)
    b1 = (rnd(torch.float32, sq, seed=3) * 0.2).requires_grad_(True)
    w2 = (rnd(torch.float32, C, sq, seed=4) * 0.2).requires_grad_(True)
    b2 = (rnd(torch.float32, C, seed=5) * 0.2).requires_grad_(True)
    pooled = d.mean((1, 2))
    gate = torch.sigmoid(F.linear(silu(F.linear(pooled, w1, b1)), w2, b2))
    if dtype == torch.bfloat16:     # forward applies the bf16-rounded gate (autocast: sigmoid output is bf16)
        gate_used = gate + (gate.detach().to(dtype).float() - gate.detach())
    else:
        gate_used = gate
    out = d * gate_used.view(B, 1, 1, C)
    da = rnd(dtype, B, H, W, C, seed=6)
    out.backward(da)

    dd = d.detach().to(DEV, dtype)
    # pool partials through bn_act with identity statistics
    _, pool = ops.bn_act(dd, None, None, None, None, act=0, want_pool=True)
    g, pooled_k, h1, g32 = ops.se_train_fwd(pool, H * W, w1.detach().to(DEV), b1.detach().to(DEV), w2.detach().to(DEV),
                                            b2.detach().to(DEV), dtype)
    assert rel(pooled_k, pooled) < 1e-5
    assert rel(g32, gate) < 1e-5
    assert rel(g, gate) < tol(dtype, 1e-5, 4e-3)
    dpool, dw1, db1, dw2, db2 = ops.se_bwd(da.to(DEV, dtype), dd, g32, pooled_k, h1, w1.detach().to(DEV), w2.detach().to(DEV))
    for got, want in ((dw1, w1.grad), (db1, b1.grad), (dw2, w2.grad), (db2, b2.grad)):
        assert rel(got, want) < 5e-5
    # d d = da * gate + dpool / HW
    dd_ref = d.grad
    got = da * gate_used.detach().view(B, 1, 1, C) + dpool.cpu().view(B, 1, 1, C) / (H * W)
    assert rel(got, dd_ref) < 5e-5


@pytest.mark.parametrize("shape", [(4, 12, 12, 1632, 68), (3, 24, 24, 672, 28), (2, 48, 47, 336, 14), (64, 5, 5, 48, 12)])
def test_gated_bn_se_bwd_in_one_reduction_pass(ops, shape):
    """bf16 training path of a block's depthwise output: BatchNorm(batch statistics) -> swish -> SE gate.  The fused sequence
    (reduction with the gate / dpool terms deferred, SE backward on its dot partials, finalize, apply) against autograd in fp32 on
    the same bf16 tensors -- d raw, gamma / beta gradients, dpool and the four SE parameter gradients."""
    B, H, W, C, sq = shape
    dtype = torch.bfloat16
    x = (rnd(dtype, B, H, W, C, seed=1) * 1.3 + 0.2).to(dtype).float().requires_grad_(True)
    gamma = (rnd(torch.float32, C, seed=2) * 0.3 + 1.0).requires_grad_(True)
    beta = (rnd(torch.float32, C, seed=3) * 0.3).requires_grad_(True)
    w1 = (rnd(torch.float32, sq, C, seed=4) * 0.2).requires_grad_(True)
    b1 = (rnd(torch.float32, sq, seed=5) * 0.2).requires_grad_(True)
    w2 = (rnd(torch.float32, C, sq, seed=6) * 0.2).requires_grad_(True)
    b2 = (rnd(torch.float32, C, seed=7) * 0.2).requires_grad_(True)
    da = rnd(dtype, B, H, W, C, seed=8)
    d = silu(F.batch_norm(x.permute(0, 3, 1, 2), None, None, gamma, beta, True, 0.0, 1e-3)).permute(0, 2, 3, 1)
    pooled = d.mean((1, 2))
    gate = torch.sigmoid(F.linear(silu(F.linear(pooled, w1, b1)), w2, b2))
    gate_used = gate + (gate.detach().to(dtype).float() - gate.detach())          # the forward applies the bf16-rounded gate
    (d * gate_used.view(B, 1, 1, C)).backward(da)

    xd = x.detach().to(DEV, dtype)
    mean, invstd = ops.bn_stats(xd, 1e-3)
    dk, pool = ops.bn_act(xd, mean, invstd, gamma.detach().to(DEV), beta.detach().to(DEV), act=1, want_pool=True)
    g, pooled_k, h1, g32 = ops.se_train_fwd(pool, H * W, w1.detach().to(DEV), b1.detach().to(DEV), w2.detach().to(DEV), b2.detach().to(DEV), dtype)
    dx, dgamma, dbeta, dpool, dw1, db1, dw2, db2 = ops.gated_bn_se_bwd(da.to(DEV, dtype), xd, mean, invstd, gamma.detach().to(DEV),
                                                                     beta.detach().to(DEV), g, g32, pooled_k, h1, w1.detach().to(DEV),
                                                                     w2.detach().to(DEV))
    assert rel(dx, x.grad) < 2e-2
    assert rel(dgamma, gamma.grad) < 1e-2 and rel(dbeta, beta.grad) < 1e-2
    for got, want in ((dw1, w1.grad), (db1, b1.grad), (dw2, w2.grad), (db2, b2.grad)):
        assert rel(got, want) < 1e-2          # d is bf16-rounded in the stored tensor, recomputed in fp32 here
    # and against the two-pass path it replaces (the SE backward reading the stored bf16 d): same numbers to bf16 noise
    dpool2, dw1b, db1b, dw2b, db2b = ops.se_bwd(da.to(DEV, dtype), dk, g32, pooled_k, h1, w1.detach().to(DEV), w2.detach().to(DEV))
    dx2, dgamma2, dbeta2 = ops.act_bn_bwd(da.to(DEV, dtype), xd, mean, invstd, gamma.detach().to(DEV), beta.detach().to(DEV), act=1, gate=g,
                                          dpool=dpool2, inv_hw=1.0 / (H * W))
    assert rel(dpool, dpool2) < 5e-3 and rel(dw2, dw2b) < 5e-3 and rel(dw1, dw1b) < 5e-3
    assert rel(dgamma, dgamma2) < 5e-3 and rel(dbeta, dbeta2) < 5e-3 and rel(dx, dx2) < 1e-2


# ------------------------------------------------------------------------------- HybridAttention
def _oracle_attention(C):
    from oracle import refmodel
    torch.manual_seed(3)
    att = refmodel.get_oracle().HybridAttention(channels=C, feature_size=(7, 7))
    with torch.no_grad():
        att.landmark_attn.attention_weights.copy_(torch.tensor([1.0, 0.8, 1.2, 0.9, 1.1]))
        att.spatial_attn.conv.weight.mul_(3.0)
    return att


@pytest.mark.parametrize("cfg", [(True, True, True), (False, True, True), (True, False, True), (True, True, False), (False, False, False)])
@pytest.mark.parametrize("dtype", DT)
def test_hybrid_attention_train_fwd_bwd(ops, cfg, dtype):
    _attention_train_case(ops, cfg, dtype, 3, 6, 6, 256)


@pytest.mark.parametrize("dtype", DT)
def test_hybrid_attention_train_head_shape(ops, dtype):
    """The shape the training step runs (12 x 12 x 1792 map, hidden 112): every position phase / vector sweep of the 1024-thread
    forward and the two-phase S3 of the backward carries work, and the argmax of a channel falls in any of the four phases."""
    _attention_train_case(ops, (True, True, True), dtype, 5, 12, 12, 1792)


def _attention_train_case(ops, cfg, dtype, B, H, W, C):
    use_lm, use_c, use_s = cfg
    att = _oracle_attention(C)
    att.use_channel, att.use_spatial = use_c, use_s
    x = (rnd(dtype, B, C, H, W, seed=1) * 0.8).to(dtype).float().requires_grad_(True)
    lm = torch.rand(B, 5, 2, generator=torch.Generator().manual_seed(2)) * 190
    feats = att(x, lm if use_lm else None).mean((2, 3))
    df = rnd(torch.float32, B, C, seed=5)
    feats.backward(df)

    xd = x.detach().permute(0, 2, 3, 1).contiguous().to(DEV, dtype)
    w5 = att.landmark_attn.attention_weights.detach().to(DEV)
    heat = raw = mx = None
    if use_lm:
        heat, raw, mx = ops.landmark_heatmap_train(lm.to(DEV), w5, H, W)
    w1, w2 = att.channel_attn.fc[0].weight.detach().to(DEV), att.channel_attn.fc[2].weight.detach().to(DEV)
    sa = att.spatial_attn.conv.weight.detach().reshape(-1).to(DEV)
    f, saved = ops.hybrid_attention_train(xd, heat, w1, w2, sa, use_c, use_s)
    assert rel(f, feats) < 2e-5
    dx, dheat, dw1, dw2, dsa = ops.hybrid_attention_bwd(xd, heat, w1, w2, sa, df.to(DEV), saved, use_c, use_s)
    assert rel(dx.float().permute(0, 3, 1, 2), x.grad) < tol(dtype, 5e-5, 6e-3)
    if use_c:
        assert rel(dw1, att.channel_attn.fc[0].weight.grad) < 5e-5
        assert rel(dw2, att.channel_attn.fc[2].weight.grad) < 5e-5
    if use_s:
        assert rel(dsa, att.spatial_attn.conv.weight.grad.reshape(-1)) < 5e-5
    if use_lm:
        dw5 = ops.landmark_heatmap_bwd(lm.to(DEV), w5, raw, mx, dheat, H, W)
        assert rel(dw5, att.landmark_attn.attention_weights.grad) < 1e-4


def test_heatmap_bwd_groups_and_ties(ops):
    """Group-wise maxima (heat_group) and duplicated samples (tied maxima share the gradient evenly)."""
    from oracle import refmodel
    la = refmodel.get_oracle().LandmarkAttention()
    with torch.no_grad():
        la.attention_weights.copy_(torch.tensor([1.0, 0.8, 1.2, 0.9, 1.1]))
    H = W = 12
    lm = torch.rand(4, 5, 2, generator=torch.Generator().manual_seed(9)) * 380
    lm[1] = lm[0]                                   # tie inside group 0
    dA = rnd(torch.float32, 4, H, W, seed=3)
    w5 = la.attention_weights.detach().to(DEV)
    want = torch.zeros(5)
    for grp in (slice(0, 2), slice(2, 4)):          # group = 2 == two reference calls
        la.attention_weights.grad = None
        a = la._create_attention_map(lm[grp], (H, W), torch.device("cpu"))
        a.backward(dA[grp].unsqueeze(1))
        want += la.attention_weights.grad
    heat, raw, mx = ops.landmark_heatmap_train(lm.to(DEV), w5, H, W, group=2)
    got = ops.landmark_heatmap_bwd(lm.to(DEV), w5, raw, mx, dA.to(DEV), H, W, group=2)
    assert rel(got, want) < 1e-4


# ------------------------------------------------------------------------------- dropout masks
def test_dropout_mask_statistics(ops):
    n, p = 1 << 20, 0.4
    m = ops.dropout_mask((n,), p, 1234, DEV)
    vals = torch.unique(m).cpu()
    assert torch.allclose(vals, torch.tensor([0.0, 1 / 0.6]))
    keep = (m > 0).float().mean().item()
    assert abs(keep - 0.6) < 3e-3
    assert abs(m.mean().item() - 1.0) < 5e-3
    m2 = ops.dropout_mask((n,), p, 1234, DEV)
    assert torch.equal(m, m2)                       # counter-based: same (seed, position) -> same mask
    m3 = ops.dropout_mask((n,), p, 1235, DEV)
    assert (m != m3).float().mean().item() > 0.3
    assert torch.all(ops.dropout_mask((1000,), 0.0, 7, DEV) == 1.0)


# ------------------------------------------------------------------------------- full-size, every element
def test_train_streams_full_size_every_element(ops):
    """Training-batch (64) launches of the streaming kernels checked element by element against torch on the GPU, twice
    each: multi-wave grids, the cp.async slot pipeline's wrap-around and the TMA tile loops only run long at this size,
    and a sporadic race is invisible to a norm (the GEMM epilogue had one, see tests/test_gpu_ops.py)."""
    dt = torch.bfloat16
    g = torch.Generator(device=DEV).manual_seed(71)

    def bad_count(y, ref, t=0.03):
        return ((y.float() - ref).abs() > t * (ref.abs() + 1.0)).sum().item()

    for (B, H, W, C) in ((64, 95, 95, 192), (64, 12, 12, 1632), (64, 48, 48, 56)):
        x = (torch.randn(B, H, W, C, device=DEV, generator=g) * 1.3 + 0.2).to(dt)
        gy = torch.randn(B, H, W, C, device=DEV, generator=g).to(dt)
        gamma = torch.randn(C, device=DEV, generator=g) * 0.3 + 1.0
        beta = torch.randn(C, device=DEV, generator=g) * 0.3
        xf = x.float().requires_grad_(True)
        gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        ref = silu(F.batch_norm(xf.permute(0, 3, 1, 2), None, None, gm, bt, True, 0.0, 1e-3)).permute(0, 2, 3, 1)
        ref.backward(gy.float())
        for rep in range(2):
            mean, invstd = ops.bn_stats(x, 1e-3)
            y, pool = ops.bn_act(x, mean, invstd, gamma, beta, act=1, want_pool=True)
            assert bad_count(y, ref.detach()) == 0, ("bn_act", B, H, W, C, rep)
            dx, dgamma, dbeta = ops.act_bn_bwd(gy, x, mean, invstd, gamma, beta, act=1)
            assert bad_count(dx, xf.grad) == 0, ("act_bn_bwd", B, H, W, C, rep)
            assert rel(dgamma, gm.grad) < 1e-2 and rel(dbeta, bt.grad) < 1e-2
        del ref, xf
    for (k, s, pl, ph, C, H) in ((5, 2, 1, 2, 192, 95), (3, 2, 0, 1, 336, 48), (5, 1, 2, 2, 960, 24), (3, 2, 0, 1, 144, 190)):
        B = 64
        x = torch.randn(B, C, H, H, device=DEV, generator=g).to(dt).float().requires_grad_(True)
        w = (torch.randn(C, 1, k, k, device=DEV, generator=g) * 0.3).requires_grad_(True)
        y = F.conv2d(F.pad(x, (pl, ph, pl, ph)), w, None, stride=s, groups=C)
        gy = torch.randn(*y.shape, device=DEV, generator=g).to(dt)
        y.backward(gy.float())
        gd = gy.permute(0, 2, 3, 1).contiguous()
        xd = x.detach().permute(0, 2, 3, 1).contiguous().to(dt)
        w_kkc = w.detach().view(C, k * k).t().contiguous()
        for rep in range(2):
            dx = ops.dwconv_dgrad(gd, w_kkc, H, H, k, s, pl, ph)
            assert bad_count(dx.permute(0, 3, 1, 2), x.grad) == 0, ("dw dgrad", k, s, C, H, rep)
            dw = ops.dwconv_wgrad(gd, xd, k, s, pl, ph)
            assert rel(dw.t().reshape(C, 1, k, k), w.grad) < 2e-3, ("dw wgrad", k, s, C, H, rep)
        del x, y


def test_pw_wgrad_full_size_every_element(ops):
    """Training-batch weight gradients of the tensor-core kernel, every dW element, three launches each (the sum over
    147k..578k rows runs the TMA / transform / MMA pipeline through hundreds of stages per CTA)."""
    dt = torch.bfloat16
    g_ = torch.Generator(device=DEV).manual_seed(72)
    for (B, HW, K, N, gated) in ((64, 2304, 336, 56, True), (64, 9025, 32, 192, False), (64, 144, 1632, 272, True),
                                 (64, 576, 160, 960, False), (64, 144, 448, 1792, False)):
        a = torch.randn(B, HW, K, device=DEV, generator=g_).to(dt)
        g = torch.randn(B, HW, N, device=DEV, generator=g_).to(dt)
        gate = torch.sigmoid(torch.randn(B, K, device=DEV, generator=g_)).to(dt) if gated else None
        ag = (a * gate.view(B, 1, K)).to(dt).float() if gated else a.float()
        ref = (g.float().reshape(-1, N).double().t() @ ag.reshape(-1, K).double()).float()
        scale = ref.abs().mean().item()
        for rep in range(3):
            dw = ops.pw_wgrad(g, a, gate, HW if gated else 0)
            bad = ((dw - ref).abs() > 2e-3 * (ref.abs() + scale)).sum().item()
            assert bad == 0, (B, HW, K, N, gated, rep, bad)
