"""TEST INFRASTRUCTURE ONLY.  The weight sets and synthetic inputs of SURVEY.md 7.1-2 / 8(d).

Default random init in eval mode is a degenerate parity case (the signal vanishes
through 32 identity-BN blocks and every logit row equals the head biases), so
parity is also run on a *BN-calibrated* model: gamma/beta lightly randomised, then
running statistics populated by a few train-mode passes (dropout and drop-connect
off, cumulative-average momentum), then ``.eval()``.
"""
import torch
import torch.nn as nn

SEED_MODEL = 42      # config/model_config.yaml:124
SEED_DATA = 1234
TEMPLATE_5PT = [[0.31, 0.32], [0.69, 0.32], [0.50, 0.55], [0.35, 0.75], [0.65, 0.75]]  # preprocessing_config.yaml:20-26


def synthetic_batch(batch, size, seed=SEED_DATA, landmarks="uniform"):
    """images ~ N(0,1) NCHW fp32, 5-pt landmarks in crop pixels, labels in {0,1}."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, size, size, generator=g)
    if landmarks == "uniform":           # scripts/test_feature_extraction.py:55,78 style
        lm = torch.rand(batch, 5, 2, generator=g) * size
    else:                                # face template +- 3 px
        lm = torch.tensor(TEMPLATE_5PT).expand(batch, 5, 2) * size + 3.0 * torch.randn(batch, 5, 2, generator=g)
    labels = torch.randint(0, 2, (batch,), generator=g)
    return images, lm, labels


def build(ns, weight_set="default", calib_size=224, calib_batches=4, calib_batch=4):
    """ns: namespace from oracle.refmodel.get_oracle() (reference or port)."""
    from .load_reference import quiet
    from .refmodel import MODEL_CONFIG
    torch.manual_seed(SEED_MODEL)
    with quiet():
        model = ns.DeepfakeDetectionModel(**MODEL_CONFIG)
    if weight_set == "default":
        return model.eval()
    assert weight_set == "calibrated", weight_set
    g = torch.Generator().manual_seed(SEED_MODEL + 1)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.weight.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        la = model.feature_extractor.attention.landmark_attn
        la.attention_weights.copy_(torch.tensor([1.0, 0.8, 1.2, 0.9, 1.1]))
    calibrate_bn(model, calib_size, calib_batches, calib_batch)
    return model.eval()


def calibrate_bn(model, size, batches, batch):
    bb = model.feature_extractor.backbone.backbone
    saved_gp = bb._global_params
    bb._global_params = saved_gp._replace(drop_connect_rate=0.0)
    drops = [m for m in model.modules() if isinstance(m, nn.Dropout)]
    saved_p = [d.p for d in drops]
    bns = [m for m in model.modules() if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d))]
    saved_m = [b.momentum for b in bns]
    for d in drops:
        d.p = 0.0
    for b in bns:
        b.momentum = None
        b.reset_running_stats()
    model.train()
    with torch.no_grad():
        for i in range(batches):
            x, lm, _ = synthetic_batch(batch, size, seed=9000 + i)
            model(x, lm)
    for d, p in zip(drops, saved_p):
        d.p = p
    for b, m in zip(bns, saved_m):
        b.momentum = m
    bb._global_params = saved_gp
    model.eval()


def block_taps(model):
    """Registers hooks; returns (dict filled on forward, remove()).  Keys: stem, block0..31, head."""
    bb = model.feature_extractor.backbone.backbone
    taps, handles = {}, []

    def keep(name):
        def hook(_m, _i, out):
            taps[name] = out.detach()
        return hook

    for i, blk in enumerate(bb._blocks):
        handles.append(blk.register_forward_hook(keep(f"block{i}")))
    # stem / head activations are produced by the parent's _swish; tap its inputs instead
    handles.append(bb._bn0.register_forward_hook(keep("stem_prebn_act")))
    handles.append(bb._bn1.register_forward_hook(keep("head_prebn_act")))

    def remove():
        for h in handles:
            h.remove()
    return taps, remove
