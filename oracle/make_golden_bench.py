"""TEST INFRASTRUCTURE ONLY.  Goldens for the BENCHMARKED configurations, from the REAL reference import.

Run in the build container (where /root/reference is mounted; ~5 min of CPU, ~40 GB of RAM):
    python -m oracle.make_golden_bench
Writes
  tests/golden/fwd_bench_b256_380_{calibrated,default}.npz   BASELINE.json configs[1]: the bench's exact batch
      (synthetic_batch(256, 380)) through the reference in fp32 AND under torch.autocast(bf16): logits, features,
      per-block statistics and the per-block relative error of the reference's own autocast run (the bf16 yardstick).
  tests/golden/train_bench_b64_380.npz                        BASELINE.json configs[2]: one training step
      (fwd + class-weighted CombinedLoss + bwd) at batch 64, 380 x 380, stochastic parts off: loss dict, logits,
      features, and for EVERY parameter the gradient norm and a signature (dot product with a fixed pseudo-random
      vector), the BatchNorm buffers after the step likewise, plus a few small gradients in full.

The batch-64 step does not fit this container's RAM with stock autograd (~2.5 GB of saved activations per image), so
every MBConv block is wrapped in torch.utils.checkpoint (non-reentrant): the saved set shrinks to the block inputs, the
arithmetic of forward and backward is unchanged (CPU kernels are deterministic), and the BatchNorm buffers -- which the
recomputation would update a second time -- are snapshotted after the forward and restored after the backward.
The GPU tests run the restated oracle on the B200 at these sizes (fp32, TF32 off), pin IT to these files, and then
compare the CUDA path with it element by element.
"""
import os
import time

import numpy as np
import torch
import torch.nn as nn
from torch.utils.checkpoint import checkpoint

from . import calibrate
from .load_reference import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SAMPLE_IDX = [0, 1, 7, 101, 1009, 5003, 10007, 20011, 300007, 1000003]
TAP_NAMES = ["stem_prebn_act"] + [f"block{i}" for i in range(32)] + ["head_prebn_act"]
LOSS_W = {"ce": 1.0, "focal": 0.5, "contrastive": 0.2}
CLASS_W = [1.0, 1.5]


def signature_vector(n, device="cpu"):
    """Fixed pseudo-random direction (float64) used to fingerprint a tensor of n elements."""
    i = torch.arange(n, dtype=torch.float64, device=device)
    return torch.cos(0.6180339887498949 * i + 0.3)


def fingerprint(t):
    t = t.detach().double().flatten()
    return float(t.norm()), float(t @ signature_vector(t.numel(), t.device))


def tap_stats(taps):
    mean = np.array([taps[n].float().mean().item() for n in TAP_NAMES], dtype=np.float64)
    std = np.array([taps[n].float().std().item() for n in TAP_NAMES], dtype=np.float64)
    norm = np.array([taps[n].double().norm().item() for n in TAP_NAMES], dtype=np.float64)
    samples = np.stack([taps[n].flatten()[[i % taps[n].numel() for i in SAMPLE_IDX]].float().numpy() for n in TAP_NAMES])
    return mean, std, norm, samples


def fwd_bench(ns, weight_set, batch=256, size=380):
    model = calibrate.build(ns, weight_set)
    x, lm, _ = calibrate.synthetic_batch(batch, size)
    taps, remove = calibrate.block_taps(model)
    t0 = time.time()
    with torch.no_grad():
        logits, feats = model(x, lm, return_features=True)
    t32 = time.time() - t0
    taps32 = dict(taps)
    mean, std, norm, samples = tap_stats(taps32)
    t0 = time.time()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        logits_ac, feats_ac = model(x, lm, return_features=True)
    tac = time.time() - t0
    ac_rel = np.array([((taps[n].double() - taps32[n].double()).norm() / (taps32[n].double().norm() + 1e-300)).item()
                       for n in TAP_NAMES])
    remove()
    print(f"{weight_set}: fp32 {t32:.1f}s autocast {tac:.1f}s  logits[0] {logits[0].tolist()}  "
          f"autocast rel err: block0 {ac_rel[1]:.3e} block31 {ac_rel[32]:.3e} "
          f"logits {((logits_ac.float() - logits).norm() / logits.norm()).item():.3e}")
    return dict(logits=logits.numpy(), features=feats.numpy(), logits_autocast=logits_ac.float().numpy(),
                features_autocast=feats_ac.float().numpy(), tap_names=np.array(TAP_NAMES), tap_mean=mean, tap_std=std,
                tap_norm=norm, tap_samples=samples, sample_idx=np.array(SAMPLE_IDX), autocast_block_rel=ac_rel,
                batch=batch, size=size, weight_set=weight_set, landmarks="uniform")


def checkpoint_blocks(model):
    """Wrap every MBConv block of the (shim) backbone in a non-reentrant checkpoint; returns an undo()."""
    blocks = list(model.feature_extractor.backbone.backbone._blocks)
    for blk in blocks:
        orig = blk.forward
        blk.forward = (lambda x, drop_connect_rate=None, _o=orig:
                       checkpoint(_o, x, drop_connect_rate, use_reentrant=False))

    def undo():
        for blk in blocks:
            del blk.forward
    return undo


def no_stochastic(model):
    bb = model.feature_extractor.backbone.backbone
    bb._global_params = bb._global_params._replace(drop_connect_rate=0.0)
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0


def train_step(ns, model, x, lm, y, class_weights, checkpointed=True):
    """One fwd + CombinedLoss + bwd of the oracle `model` (already .train(), stochastic parts off).  Returns
    (logits, features, losses dict); gradients are left in .grad, BatchNorm buffers hold the post-step values."""
    undo = checkpoint_blocks(model) if checkpointed else (lambda: None)
    try:
        model.zero_grad(set_to_none=True)
        logits, feats = model(x, lm, return_features=True)
        buffers = {k: v.clone() for k, v in model.named_buffers()}
        cw = None if class_weights is None else torch.tensor(class_weights, device=x.device)
        losses = ns.CombinedLoss(LOSS_W, cw)(logits, y, feats)
        losses["total"].backward()
        with torch.no_grad():
            for k, v in model.named_buffers():     # the recomputation updated them a second time
                v.copy_(buffers[k])
    finally:
        undo()
    return logits.detach(), feats.detach(), {k: float(v.detach()) if torch.is_tensor(v) else float(v) for k, v in losses.items()}


FULL_GRADS = ("feature_extractor.attention.landmark_attn.attention_weights",
              "feature_extractor.attention.spatial_attn.conv.weight",
              "feature_extractor.backbone.backbone._bn0.weight", "feature_extractor.backbone.backbone._bn0.bias",
              "feature_extractor.backbone.backbone._blocks.31._bn2.weight",
              "feature_extractor.backbone.backbone._blocks.22._se_reduce.bias",
              "classifier.12.weight", "classifier.12.bias", "classifier.9.weight")


def train_bench(ns, batch=64, size=380):
    model = calibrate.build(ns, "calibrated")
    no_stochastic(model)
    model.train()
    x, lm, y = calibrate.synthetic_batch(batch, size)
    t0 = time.time()
    logits, feats, losses = train_step(ns, model, x, lm, y, CLASS_W)
    print(f"train step B={batch} @{size}: {time.time() - t0:.1f}s  losses {losses}")
    names = [n for n, _ in model.named_parameters()]
    fp = np.array([fingerprint(p.grad) for _, p in model.named_parameters()])
    bnames = [n for n, b in model.named_buffers() if b.dtype.is_floating_point]
    bfp = np.array([fingerprint(b) for n, b in model.named_buffers() if b.dtype.is_floating_point])
    rec = dict(logits=logits.numpy(), features=feats.numpy(), labels=y.numpy(), param_names=np.array(names),
               grad_norm=fp[:, 0], grad_signature=fp[:, 1], buffer_names=np.array(bnames), buffer_norm=bfp[:, 0],
               buffer_signature=bfp[:, 1], batch=batch, size=size, class_weights=np.array(CLASS_W))
    for k, v in losses.items():
        rec[f"loss_{k}"] = v
    params = dict(model.named_parameters())
    for n in FULL_GRADS:
        rec[f"grad:{n}"] = params[n].grad.numpy().copy()
    return rec


def main():
    ns = load_reference()
    assert ns.kind == "reference"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 8)
    np.savez_compressed(os.path.join(OUT, "train_bench_b64_380.npz"), **train_bench(ns))
    for ws in ("calibrated", "default"):
        rec = fwd_bench(ns, ws)
        if ws == "default":      # degenerate set: keep the small fields only
            rec = {k: v for k, v in rec.items() if not k.startswith("features")}
        np.savez_compressed(os.path.join(OUT, f"fwd_bench_b256_380_{ws}.npz"), **rec)


if __name__ == "__main__":
    main()
