"""Pins the CPU oracle (no GPU needed).

1. the efficientnet_pytorch restatement == HF transformers' independent EfficientNet-B4
   (bit-exact eval features), and has torchvision-B4's parameter count;
2. the restated wrapper (oracle/refmodel.py) == the real reference import, where
   /root/reference is mounted (keys, seeded init, forward, loss, gradients: bit-exact);
3. the restated oracle reproduces the committed golden vectors that were generated from
   the real reference import (oracle/make_golden.py) -- this is what runs on the GPU box.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import calibrate, refmodel
from oracle.load_reference import load_reference, reference_available

torch.set_num_threads(min(8, os.cpu_count() or 1))


def _port_ns():
    import types
    return types.SimpleNamespace(**{k: getattr(refmodel, k) for k in (
        "DeepfakeDetectionModel", "CombinedLoss", "HybridAttention", "LandmarkAttention")}, kind="port")


def test_param_counts_and_keys():
    m = calibrate.build(_port_ns(), "default")
    bb = m.feature_extractor.backbone.backbone
    assert sum(p.numel() for p in bb.parameters()) == 17_548_616
    assert sum(p.numel() for p in m.feature_extractor.attention.parameters()) == 401_511
    assert sum(p.numel() for p in m.classifier.parameters()) == 989_218
    assert sum(p.numel() for p in m.parameters()) == 18_939_345
    sd = m.state_dict()
    assert len(sd) == 731 and not any("_fc" in k for k in sd)
    assert sum(v.numel() for k, v in m.named_buffers()) == 126_643
    # block table of SURVEY A.2
    blocks = bb._blocks
    assert len(blocks) == 32
    pads = {2: (0, 1, 0, 1), 6: (2, 2, 2, 2), 10: (0, 1, 0, 1), 22: (1, 2, 1, 2), 0: (1, 1, 1, 1), 17: (2, 2, 2, 2)}
    for i, p in pads.items():
        assert tuple(blocks[i]._depthwise_conv.static_padding.padding) == p
    assert tuple(bb._conv_stem.static_padding.padding) == (0, 1, 0, 1)


def _hf_model():
    from transformers import EfficientNetConfig, EfficientNetModel
    cfg = EfficientNetConfig(width_coefficient=1.4, depth_coefficient=1.8, image_size=380, hidden_dim=1792,
                             depthwise_padding=[6], dropout_rate=0.4)
    return EfficientNetModel(cfg).eval()


def _remap_to_hf(sd):
    out = {}
    for k, v in sd.items():
        k2 = k
        k2 = k2.replace("_conv_stem.", "embeddings.convolution.").replace("_conv_head.", "encoder.top_conv.")
        if k2.startswith("_bn0."):
            k2 = "embeddings.batchnorm." + k2[5:]
        elif k2.startswith("_bn1."):
            k2 = "encoder.top_bn." + k2[5:]
        elif k2.startswith("_blocks."):
            _, i, rest = k2.split(".", 2)
            rest = (rest.replace("_expand_conv.", "expansion.expand_conv.").replace("_bn0.", "expansion.expand_bn.")
                    .replace("_depthwise_conv.", "depthwise_conv.depthwise_conv.").replace("_bn1.", "depthwise_conv.depthwise_norm.")
                    .replace("_se_reduce.", "squeeze_excite.reduce.").replace("_se_expand.", "squeeze_excite.expand.")
                    .replace("_project_conv.", "projection.project_conv.").replace("_bn2.", "projection.project_bn."))
            k2 = f"encoder.blocks.{i}.{rest}"
        out[k2] = v
    return out


@pytest.mark.parametrize("size", [96, 224])
def test_shim_matches_hf_transformers(size):
    """Independent implementation cross-check."""
    from efficientnet_pytorch import EfficientNet
    torch.manual_seed(0)
    ours = EfficientNet.from_name("efficientnet-b4", num_classes=1000)
    ours._fc = nn.Identity()
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for m in ours.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(0.8 + 0.8 * torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.bias.shape, generator=g))
    ours.eval()
    ours.set_swish(memory_efficient=False)
    hf = _hf_model()
    missing, unexpected = hf.load_state_dict(_remap_to_hf(ours.state_dict()), strict=False)
    assert not unexpected and all("pooler" in k for k in missing), (missing, unexpected)
    x = torch.randn(2, 3, size, size, generator=g)
    with torch.no_grad():
        a = ours.extract_features(x)
        b = hf(pixel_values=x).last_hidden_state
    assert a.shape == b.shape
    # HF uses F.silu, the shim i*sigmoid(i): identical function, may differ in the last ulp
    rel = ((a - b).norm() / b.norm()).item()
    assert rel < 1e-5, rel


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")
def test_port_is_bit_identical_to_reference_import():
    ref = load_reference()
    port = _port_ns()
    for ws in ("default", "calibrated"):
        a, b = calibrate.build(ref, ws), calibrate.build(port, ws)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        assert all(torch.equal(sa[k], sb[k]) for k in sa)
    x, lm, y = calibrate.synthetic_batch(4, 160, landmarks="template")
    with torch.no_grad():
        for lmk in (lm, None):
            (la, fa), (lb, fb) = a(x, lmk, return_features=True), b(x, lmk, return_features=True)
            assert torch.equal(la, lb) and torch.equal(fa, fb)
        assert torch.equal(a.predict(x, lm), b.predict(x, lm))
        ea = a.feature_extractor.get_embedding(x, lm)
        assert torch.allclose(ea.norm(dim=1), torch.ones(4), atol=1e-5)     # feature_extractor.py:338
        ha = a.feature_extractor(x, lm, return_attention=True)[1]
        hb = b.feature_extractor(x, lm, return_attention=True)[1]
        assert ha.shape == (4, 1, 7, 7) and torch.equal(ha, hb)
        ma, mb = a.feature_extractor.extract_multi_scale_features(x, lm), b.feature_extractor.extract_multi_scale_features(x, lm)
        assert set(ma) == {"reduction_2", "reduction_4", "reduction_5", "final"}
        assert all(torch.equal(ma[k], mb[k]) for k in ma)
    # train-mode fwd + CombinedLoss + bwd (stochastic parts off)
    outs = []
    for ns, m in ((ref, a), (port, b)):
        bb = m.feature_extractor.backbone.backbone
        bb._global_params = bb._global_params._replace(drop_connect_rate=0.0)
        for mod in m.modules():
            if isinstance(mod, nn.Dropout):
                mod.p = 0.0
        m.train()
        crit = ns.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, torch.tensor([1.0, 1.5]))
        lo, fe = m(x, lm, return_features=True)
        losses = crit(lo, y, fe)
        losses["total"].backward()
        outs.append((losses, {n: p.grad.clone() for n, p in m.named_parameters()}, m.state_dict()))
    (l1, g1, s1), (l2, g2, s2) = outs
    assert set(l1) == set(l2) == {"ce", "focal", "contrastive", "total"}
    assert all(torch.equal(l1[k], l2[k]) for k in l1)
    assert all(torch.equal(g1[k], g2[k]) for k in g1)
    assert all(torch.equal(s1[k], s2[k]) for k in s1)


def _check_golden(path, rtol):
    g = np.load(path, allow_pickle=False)
    port = _port_ns()
    model = calibrate.build(port, str(g["weight_set"]))
    x, lm, y = calibrate.synthetic_batch(int(g["batch"]), int(g["size"]), landmarks=str(g["landmarks"]))
    taps, remove = calibrate.block_taps(model)
    with torch.no_grad():
        logits, feats = model(x, lm, return_features=True)
        hw = taps["block31"].shape[-1]
        heat = model.feature_extractor.attention.landmark_attn._create_attention_map(lm, (hw, hw), x.device)
        logits_nolm, _ = model(x, None)
    remove()
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=rtol, atol=rtol)
    np.testing.assert_allclose(logits_nolm.numpy(), g["logits_no_landmarks"], rtol=rtol, atol=rtol)
    np.testing.assert_allclose(feats.numpy(), g["features"], rtol=rtol, atol=rtol * np.abs(g["features"]).max())
    np.testing.assert_allclose(heat.numpy(), g["heatmap"], rtol=1e-6, atol=1e-7)
    for i, n in enumerate(g["tap_names"]):
        t = taps[str(n)]
        assert abs(t.mean().item() - g["tap_mean"][i]) <= rtol * (abs(g["tap_mean"][i]) + g["tap_std"][i] + 1e-12)
        assert abs(t.std().item() - g["tap_std"][i]) <= rtol * (g["tap_std"][i] + 1e-12)
    return g, model, (x, lm, y)


def test_golden_default_b8_380(golden_dir):
    """BASELINE.json configs[0]: batch 8, 380x380, default init (degenerate: rows equal)."""
    g, _, _ = _check_golden(os.path.join(golden_dir, "fwd_default_b8_380.npz"), 2e-4)
    assert g["heatmap"].shape == (8, 1, 12, 12)
    assert g["heatmap"].min() >= 0.1 and g["heatmap"].max() <= 1.0


def test_golden_calibrated_b2_380(golden_dir):
    _check_golden(os.path.join(golden_dir, "fwd_calibrated_b2_380.npz"), 2e-3)


def test_golden_calibrated_b4_224_with_training_step(golden_dir):
    g, model, (x, lm, y) = _check_golden(os.path.join(golden_dir, "fwd_calibrated_b4_224.npz"), 2e-3)
    bb = model.feature_extractor.backbone.backbone
    bb._global_params = bb._global_params._replace(drop_connect_rate=0.0)
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    model.train()
    crit = refmodel.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, None)
    lo, fe = model(x, lm, return_features=True)
    losses = crit(lo, y, fe)
    losses["total"].backward()
    for k in ("ce", "focal", "contrastive", "total"):
        assert abs(losses[k].item() - float(g[f"loss_{k}"])) < 1e-4
    for key in g.files:
        if key.startswith("gradnorm:"):
            name = key.split(":", 1)[1]
            p = dict(model.named_parameters())[name]
            assert abs(p.grad.norm().item() - float(g[key])) <= 2e-3 * float(g[key]) + 1e-9, name
    np.testing.assert_allclose(
        model.state_dict()["feature_extractor.backbone.backbone._bn0.running_mean"].numpy(),
        g["bn0_running_mean_after"], rtol=1e-4, atol=1e-6)


def test_golden_combined_loss(golden_dir):
    g = np.load(os.path.join(golden_dir, "combined_loss.npz"))
    for B in (1, 2, 5, 8):
        for tag, cw in (("", None), ("_cw", torch.tensor([1.0, 1.5]))):
            crit = refmodel.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, cw)
            out = crit(torch.from_numpy(g[f"B{B}{tag}_in_logits"]), torch.from_numpy(g[f"B{B}{tag}_in_y"]),
                       torch.from_numpy(g[f"B{B}{tag}_in_feats"]))
            assert ("contrastive" in out) == (B >= 2)
            for k, v in out.items():
                assert abs(float(v) - float(g[f"B{B}{tag}_{k}"])) < 1e-6
