"""B200 only: the operators either side of the hot path against the CPU oracle / stock torch:
video-clip scoring (BASELINE.json configs[3]), the feature extractor's side APIs, and the fused
clip_grad_norm_ + AdamW step with a stock-format optimizer state_dict."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _pair(weight_set="calibrated"):
    import deepfake_vit_b200 as d
    from oracle import calibrate, refmodel
    om = calibrate.build(refmodel.get_oracle(), weight_set, calib_size=128, calib_batches=2)
    m = d.DeepfakeDetectionModel(**refmodel.MODEL_CONFIG)
    m.load_state_dict(om.state_dict(), strict=True)
    return om.eval(), m.to(DEV).eval().set_compute_dtype(torch.float32)


def test_score_clips_matches_one_reference_call_per_clip():
    """task.ipynb:434-442: one model call per file, softmax, mean fake probability, >= 0.5."""
    from oracle import calibrate
    om, m = _pair()
    clips, frames, size = 3, 4, 128
    x, lm, _ = calibrate.synthetic_batch(clips * frames, size)
    out = m.score_clips(x.to(DEV), lm.to(DEV), frames_per_clip=frames)
    for c in range(clips):
        sl = slice(c * frames, (c + 1) * frames)
        with torch.no_grad():
            logits, _ = om(x[sl], lm[sl])          # the heat-map normaliser sees this clip only
        prob = torch.softmax(logits, dim=1)[:, 1].mean()
        assert rel(out["mean_logits"][c], logits.mean(0)) < 1e-4
        assert abs(out["fake_prob"][c].item() - prob.item()) < 1e-4
        assert int(out["labels"][c].item()) == int(prob.item() >= 0.5)


def test_clip_aggregate_kernel():
    import deepfake_vit_b200 as d
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(6 * 32, 2, generator=g) * 3
    ml, p, lab = d.ops.clip_aggregate(logits.to(DEV), 32)
    ref_p = torch.softmax(logits, 1)[:, 1].view(6, 32).mean(1)
    assert torch.allclose(ml.cpu(), logits.view(6, 32, 2).mean(1), atol=1e-6)
    assert torch.allclose(p.cpu(), ref_p, atol=1e-6)
    assert torch.equal(lab.cpu().long(), (ref_p >= 0.5).long())


def test_feature_extractor_side_apis():
    from oracle import calibrate
    om, m = _pair()
    x, lm, _ = calibrate.synthetic_batch(3, 128)
    with torch.no_grad():
        f_ref, a_ref = om.feature_extractor(x, lm, return_attention=True)
        ms_ref = om.feature_extractor.extract_multi_scale_features(x, lm)
        e_ref = om.feature_extractor.get_embedding(x, lm)
    f, a = m.feature_extractor(x.to(DEV), lm.to(DEV), return_attention=True)
    assert rel(f, f_ref) < 1e-4
    assert a.shape == a_ref.shape == (3, 1, 7, 7) and rel(a, a_ref) < 1e-6
    ms = m.feature_extractor.extract_multi_scale_features(x.to(DEV), lm.to(DEV))
    assert set(ms) == set(ms_ref)
    for k in ms_ref:
        assert ms[k].shape == ms_ref[k].shape and rel(ms[k], ms_ref[k]) < 1e-4, k
    e = m.feature_extractor.get_embedding(x.to(DEV), lm.to(DEV))
    assert rel(e, e_ref) < 1e-4
    assert torch.allclose(e.norm(dim=1).cpu(), torch.ones(3), atol=1e-5)
    assert m.feature_extractor(x.to(DEV), None)[1] is None


@pytest.mark.parametrize("max_norm", [1.0, None])
def test_fused_clip_adamw_matches_torch(max_norm):
    import deepfake_vit_b200 as d
    g = torch.Generator().manual_seed(5)
    shapes = [(48, 3, 3, 3), (48,), (1792, 448, 1, 1), (5,), (2, 32), (7, 13)]
    ref_p = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref_p]
    ref_opt = torch.optim.AdamW(ref_p, lr=1e-3, weight_decay=1e-2)
    our_opt = d.FusedAdamW(our_p, lr=1e-3, weight_decay=1e-2, max_grad_norm=max_norm)
    for step in range(4):
        grads = [torch.randn(*s, generator=g) * (10.0 if step == 1 else 0.3) for s in shapes]
        for p, q, gr in zip(ref_p, our_p, grads):
            p.grad = gr.clone()
            q.grad = gr.clone().to(DEV)
        norm_ref = torch.nn.utils.clip_grad_norm_(ref_p, max_norm) if max_norm else torch.sqrt(sum((x ** 2).sum() for x in grads))
        ref_opt.step()
        norm = our_opt.step()
        assert abs(norm.item() - float(norm_ref)) < 1e-4 * float(norm_ref)
        for p, q in zip(ref_p, our_p):
            assert rel(q.detach(), p.detach()) < 2e-6
            # moments scale with the clip coefficient, i.e. with each side's fp32 norm reduction order
            assert rel(our_opt.state[q]["exp_avg"], ref_opt.state[p]["exp_avg"]) < (2e-5 if max_norm else 2e-6)
    # stock-format state_dict: loads into torch.optim.AdamW and back
    sd = our_opt.state_dict()
    ref2_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref2 = torch.optim.AdamW(ref2_p, lr=1e-3, weight_decay=1e-2)
    ref2.load_state_dict(copy.deepcopy(sd))
    for (p, q) in zip(ref_p, ref2_p):
        assert rel(ref2.state[q]["exp_avg"], ref_opt.state[p]["exp_avg"]) < (2e-5 if max_norm else 2e-6)
        assert rel(ref2.state[q]["exp_avg_sq"], ref_opt.state[p]["exp_avg_sq"]) < (4e-5 if max_norm else 2e-6)
        assert int(ref2.state[q]["step"]) == 4
    our2_p = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref_p]
    our2 = d.FusedAdamW(our2_p, lr=5e-4, weight_decay=0.0, max_grad_norm=max_norm)
    our2.load_state_dict(ref_opt.state_dict())
    assert our2.param_groups[0]["lr"] == 1e-3 and our2._step == 4
    grads = [torch.randn(*s, generator=g) for s in shapes]
    for p, q, gr in zip(ref_p, our2_p, grads):
        p.grad, q.grad = gr.clone(), gr.clone().to(DEV)
    if max_norm:
        torch.nn.utils.clip_grad_norm_(ref_p, max_norm)
    ref_opt.step()
    our2.step()
    for p, q in zip(ref_p, our2_p):
        assert rel(q.detach(), p.detach()) < 2e-6


def test_train_step_with_fused_optimizer_reads_the_flat_gradient():
    """model backward leaves one flat gradient buffer; FusedAdamW steps from it without gathering."""
    import deepfake_vit_b200 as d
    from oracle import calibrate, refmodel
    torch.manual_seed(0)
    m = d.DeepfakeDetectionModel(**refmodel.MODEL_CONFIG).to(DEV).train().set_compute_dtype(torch.bfloat16)
    opt = d.FusedAdamW(m.parameters(), lr=1e-4, weight_decay=1e-4, max_grad_norm=1.0, grad_source=m)
    crit = d.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, None)
    x, lm, y = calibrate.synthetic_batch(4, 96)
    before = [p.detach().clone() for p in m.parameters()]
    lo, fe = m(x.to(DEV), lm.to(DEV), return_features=True)
    crit(lo, y.to(DEV), fe)["total"].backward()
    assert opt._flat_grad().data_ptr() == m._last_flat_grad.data_ptr()
    norm = opt.step()
    assert torch.isfinite(norm).all() and norm.item() > 0
    changed = sum(int(not torch.equal(a, b.detach())) for a, b in zip(before, m.parameters()))
    assert changed > 0.9 * len(before)
    lo2, _ = m(x.to(DEV), lm.to(DEV))          # the packed / cast weights follow the in-place update
    assert torch.isfinite(lo2).all()


def test_graphed_inference_replays_the_eager_forward_bit_for_bit():
    """GraphedInference (one CUDA-graph replay per batch) returns exactly what the eager forward returns, follows new
    inputs copied into its static buffers, and rejects a different shape."""
    import deepfake_vit_b200 as d
    torch.manual_seed(3)
    m = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG).cuda().eval().set_compute_dtype(torch.bfloat16)
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(8, 3, 224, 224, device="cuda", generator=g)
    lm = torch.rand(8, 5, 2, device="cuda", generator=g) * 224
    gi = d.GraphedInference(m, x, lm, return_features=True)
    for trial in range(2):
        x2 = torch.randn(8, 3, 224, 224, device="cuda", generator=g)
        lm2 = torch.rand(8, 5, 2, device="cuda", generator=g) * 224
        with torch.no_grad():
            ref_lo, ref_fe = m(x2, lm2, return_features=True)
        lo, fe = gi(x2, lm2)
        torch.cuda.synchronize()
        assert torch.equal(lo, ref_lo) and torch.equal(fe, ref_fe), trial
    with pytest.raises(ValueError):
        gi(x[:4], lm[:4])
