"""The optional `global` data-parallel modes (SURVEY.md 8(e) caveats 1 and 3), checked on ONE GPU by playing both ranks:
the batch is cut in two shards, the 4-byte exchanges are done by hand, and the shard results must reproduce the
single-GPU result on the whole batch (what the reference computes when it sees the global batch in one call)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

MIN32 = -2 ** 31


def _umax(a, b):        # maximum of two int32-stored uint32 keys in UNSIGNED order (what parallel.allreduce_max_key does)
    return torch.maximum(a ^ MIN32, b ^ MIN32) ^ MIN32


def test_global_heatmap_maximum_matches_the_whole_batch():
    from deepfake_vit_b200 import ops
    g = torch.Generator().manual_seed(5)
    lm = (torch.rand(8, 5, 2, generator=g) * 380).cuda()
    lm[6:] += 150.0                     # the second shard's landmarks fall partly off the map: its own maximum is smaller
    w5 = torch.tensor([1.0, 0.8, 1.2, 0.5, 0.9]).cuda()
    whole = ops.landmark_heatmap(lm, w5, 12, 12)
    k0, k1 = ops.landmark_max_key(lm[:4], w5, 12, 12), ops.landmark_max_key(lm[4:], w5, 12, 12)
    assert int(k0[0]) != int(k1[0])
    kg = _umax(k0, k1)
    parts = [ops.landmark_heatmap(lm[:4].contiguous(), w5, 12, 12, max_floor=kg), ops.landmark_heatmap(lm[4:].contiguous(), w5, 12, 12, max_floor=kg)]
    assert torch.equal(torch.cat(parts), whole), "shards normalised by the exchanged maximum differ from the whole-batch map"
    local = torch.cat([ops.landmark_heatmap(lm[:4].contiguous(), w5, 12, 12), ops.landmark_heatmap(lm[4:].contiguous(), w5, 12, 12)])
    assert not torch.equal(local, whole), "the case must distinguish the per-rank normaliser from the global one"
    # negative learnt weights: negative maxima keep their order through the key
    wneg = -w5
    whole = ops.landmark_heatmap(lm, wneg, 12, 12)
    kg = _umax(ops.landmark_max_key(lm[:4], wneg, 12, 12), ops.landmark_max_key(lm[4:], wneg, 12, 12))
    parts = [ops.landmark_heatmap(lm[:4].contiguous(), wneg, 12, 12, max_floor=kg), ops.landmark_heatmap(lm[4:].contiguous(), wneg, 12, 12, max_floor=kg)]
    assert torch.equal(torch.cat(parts), whole)


def test_model_global_scope_reproduces_the_whole_batch(monkeypatch):
    """model.landmark_max_scope = 'global': each shard's logits equal the whole-batch call's rows (fp32 mode, bit for bit on
    the heat map -> identical everywhere), with the other rank's key supplied through a patched exchange."""
    import deepfake_vit_b200 as d
    from deepfake_vit_b200 import parallel
    from oracle import calibrate, refmodel
    om = calibrate.build(refmodel.get_oracle(), "calibrated", calib_size=96, calib_batches=1)
    m = d.DeepfakeDetectionModel(**refmodel.MODEL_CONFIG)
    m.load_state_dict(om.state_dict(), strict=True)
    m = m.cuda().eval().set_compute_dtype(torch.float32)
    x, lm, _ = calibrate.synthetic_batch(4, 96)
    x, lm = x.cuda(), lm.cuda()
    lm[2:] += 40.0
    whole, _ = m(x, lm)
    keys = {}
    monkeypatch.setattr(parallel, "allreduce_max_key", lambda k, group=None: keys.setdefault("mine", k) if "other" not in keys else _umax(k, keys["other"]))
    m.landmark_max_scope = "global"
    m(x[:2], lm[:2]); k0 = keys.pop("mine")
    m(x[2:], lm[2:]); k1 = keys.pop("mine")
    keys["other"] = k1
    a, _ = m(x[:2], lm[:2])
    keys["other"] = k0
    b, _ = m(x[2:], lm[2:])
    assert torch.equal(torch.cat([a, b]), whole)
    m.train()
    with pytest.raises(RuntimeError, match="inference mode"):
        m(x, lm)


@pytest.mark.parametrize("w_focal,w_con", [(0.0, 0.0), (0.5, 0.2)])
def test_global_weighted_ce_matches_the_whole_batch(w_focal, w_con):
    from deepfake_vit_b200 import ops
    from oracle import refmodel
    g = torch.Generator().manual_seed(11)
    logits = torch.randn(8, 2, generator=g) * 2
    feats = torch.randn(8, 64, generator=g)
    y = torch.tensor([0, 0, 0, 1, 1, 1, 1, 0])          # shard class mixes differ: 3:1 and 1:3
    cw = torch.tensor([1.0, 1.5])
    # the reference on the global batch (CPU oracle)
    lo = logits.clone().requires_grad_(True)
    ref = refmodel.CombinedLoss({"ce": 1.0, "focal": w_focal, "contrastive": w_con}, cw)(lo, y, feats if w_con else None)
    ref["total"].backward()
    L, F, Y, CW = logits.cuda(), feats.cuda(), y.cuda(), cw.cuda()
    n = [ops.class_weight_sum(Y[:4].contiguous(), CW, 2), ops.class_weight_sum(Y[4:].contiguous(), CW, 2)]
    assert abs(n[0].item() - 4.5) < 1e-6 and abs(n[1].item() - 5.5) < 1e-6
    mean_norm = (n[0] + n[1]) / 2                        # what parallel.allreduce_mean_ leaves on every rank
    outs = [ops.combined_loss(L[s].contiguous(), Y[s].contiguous(), F[s].contiguous() if w_con else None, CW, 1.0, w_focal, w_con, ce_norm=mean_norm)
            for s in (slice(0, 4), slice(4, 8))]
    ce = (outs[0][0][0] + outs[1][0][0]).item() / 2
    total = (outs[0][0][3] + outs[1][0][3]).item() / 2
    assert abs(ce - ref["ce"].item()) < 1e-6 * max(1.0, abs(ref["ce"].item()))
    assert abs(total - ref["total"].item()) < 2e-6 * max(1.0, abs(ref["total"].item()))
    dl = torch.cat([outs[0][2], outs[1][2]]).cpu() / 2   # the gradient all-reduce averages the ranks
    assert torch.allclose(dl, lo.grad, atol=2e-7, rtol=1e-5)
    # and the per-rank normaliser (DDP semantics) is a different number on this class mix
    loc = [ops.combined_loss(L[s].contiguous(), Y[s].contiguous(), None, CW, 1.0, 0.0, 0.0) for s in (slice(0, 4), slice(4, 8))]
    assert abs((loc[0][0][0] + loc[1][0][0]).item() / 2 - ref["ce"].item()) > 1e-4
