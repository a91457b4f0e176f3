"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/*.npz from the REAL reference import.

Run in the build container (where /root/reference is mounted):
    python -m oracle.make_golden
The reference's own wrapper / attention / head / loss code runs verbatim
(oracle/load_reference.py); only efficientnet_pytorch is the restated shim.  The GPU
box has no /root/reference, so tests there check the restated oracle (and the CUDA
path) against these files instead.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import calibrate
from .load_reference import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SAMPLE_IDX = [0, 1, 7, 101, 1009, 5003, 10007, 20011]


def block_stats(taps):
    names = ["stem_prebn_act"] + [f"block{i}" for i in range(32)] + ["head_prebn_act"]
    mean = np.array([taps[n].mean().item() for n in names], dtype=np.float64)
    std = np.array([taps[n].std().item() for n in names], dtype=np.float64)
    samples = np.stack([taps[n].flatten()[[i % taps[n].numel() for i in SAMPLE_IDX]].numpy() for n in names])
    return names, mean, std, samples


def forward_record(ns, weight_set, batch, size, landmarks):
    model = calibrate.build(ns, weight_set)
    x, lm, y = calibrate.synthetic_batch(batch, size, landmarks=landmarks)
    taps, remove = calibrate.block_taps(model)
    with torch.no_grad():
        logits, feats = model(x, lm, return_features=True)
        fm_hw = taps["block31"].shape[-1]
        la = model.feature_extractor.attention.landmark_attn
        heat = la._create_attention_map(lm, (fm_hw, fm_hw), x.device)
        logits_nolm, _ = model(x, None)
    remove()
    names, mean, std, samples = block_stats(taps)
    rec = dict(logits=logits.numpy(), features=feats.numpy(), heatmap=heat.numpy(),
               logits_no_landmarks=logits_nolm.numpy(), tap_names=np.array(names),
               tap_mean=mean, tap_std=std, tap_samples=samples,
               batch=batch, size=size, weight_set=weight_set, landmarks=landmarks)
    return model, (x, lm, y), rec


def train_record(ns, model, data):
    """fwd + CombinedLoss + bwd in train mode with all stochastic parts switched off."""
    x, lm, y = data
    bb = model.feature_extractor.backbone.backbone
    bb._global_params = bb._global_params._replace(drop_connect_rate=0.0)
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    model.train()
    rec = {}
    for tag, cw in (("", None), ("_cw", torch.tensor([1.0, 1.5]))):
        model.zero_grad(set_to_none=True)
        crit = ns.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, cw)
        sd0 = {k: v.clone() for k, v in model.state_dict().items() if "running" in k}
        logits, feats = model(x, lm, return_features=True)
        losses = crit(logits, y, feats)
        losses["total"].backward()
        for k, v in losses.items():
            rec[f"loss_{k}{tag}"] = v.item()
        rec[f"train_logits{tag}"] = logits.detach().numpy()
        for name, p in model.named_parameters():
            if name.endswith(("_conv_stem.weight", "_blocks.17._depthwise_conv.weight", "_blocks.17._bn1.weight",
                              "_blocks.3._project_conv.weight", "_blocks.30._se_reduce.bias", "_conv_head.weight",
                              "attention_weights", "spatial_attn.conv.weight", "channel_attn.fc.0.weight",
                              "classifier.0.weight", "classifier.12.bias")):
                rec[f"gradnorm{tag}:{name}"] = p.grad.norm().item()
        if tag == "":
            rec["bn0_running_mean_after"] = model.state_dict()[
                "feature_extractor.backbone.backbone._bn0.running_mean"].numpy().copy()
        # restore the running stats so both passes see the same model
        model.load_state_dict({**model.state_dict(), **sd0})
    return rec


def loss_record(ns):
    g = torch.Generator().manual_seed(77)
    rec = {}
    for B in (1, 2, 5, 8):
        logits = torch.randn(B, 2, generator=g) * 2
        feats = torch.randn(B, 1792, generator=g) * 0.05
        y = torch.randint(0, 2, (B,), generator=g)
        for tag, cw in (("", None), ("_cw", torch.tensor([1.0, 1.5]))):
            out = ns.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, cw)(logits, y, feats)
            rec[f"B{B}{tag}_in_logits"] = logits.numpy()
            rec[f"B{B}{tag}_in_feats"] = feats.numpy()
            rec[f"B{B}{tag}_in_y"] = y.numpy()
            for k, v in out.items():
                rec[f"B{B}{tag}_{k}"] = float(v)
    return rec


def main():
    ns = load_reference()
    assert ns.kind == "reference"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)

    _, _, rec = forward_record(ns, "default", 8, 380, "uniform")          # BASELINE.json configs[0]
    np.savez_compressed(os.path.join(OUT, "fwd_default_b8_380.npz"), **rec)
    print("default 380", rec["logits"][:2])

    model, data, rec = forward_record(ns, "calibrated", 4, 224, "template")
    rec.update(train_record(ns, model, data))
    np.savez_compressed(os.path.join(OUT, "fwd_calibrated_b4_224.npz"), **rec)
    print("calibrated 224", rec["logits"][:2], rec["loss_total"])

    _, _, rec = forward_record(ns, "calibrated", 2, 380, "uniform")
    np.savez_compressed(os.path.join(OUT, "fwd_calibrated_b2_380.npz"), **rec)
    print("calibrated 380", rec["logits"][:2])

    np.savez_compressed(os.path.join(OUT, "combined_loss.npz"), **loss_record(ns))


if __name__ == "__main__":
    main()
