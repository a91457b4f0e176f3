"""Same-process A/B: batch-256 bf16 forward replayed from a CUDA graph with the large kernels chained by programmatic
dependent launch (default) vs plain stream order (model.stream_ordered_launches = True).  Prints ms per forward for both,
twice, and checks that the logits / features are bit-identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import deepfake_vit_b200 as d
from oracle import calibrate, refmodel

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B = 256
torch.manual_seed(42)
om = calibrate.build(refmodel.get_oracle(), "calibrated", calib_size=128, calib_batches=2)
m = d.DeepfakeDetectionModel(**refmodel.MODEL_CONFIG)
m.load_state_dict(om.state_dict(), strict=True)
m = m.cuda().eval().set_compute_dtype(torch.bfloat16)
x = torch.randn(B, 3, 380, 380, device="cuda")
lm = torch.rand(B, 5, 2, device="cuda") * 380


def timed(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


outs = {}
for rnd in range(2):
    for so in (True, False):
        m.stream_ordered_launches = so
        gi = d.GraphedInference(m, x, lm, return_features=True)
        t = timed(gi.replay, steps)
        te = timed(lambda: m(x, lm), steps)
        lo, fe = gi.replay()
        torch.cuda.synchronize()
        outs[so] = (lo.clone(), fe.clone())
        print(f"round {rnd} {'stream-ordered' if so else 'PDL chain     '}: graph {t:.3f} ms   eager {te:.3f} ms   ({B / t * 1e3:.0f} img/s)")
        del gi
print("bit-identical logits:", torch.equal(outs[True][0], outs[False][0]), " features:", torch.equal(outs[True][1], outs[False][1]))
# repeatability of the PDL chain itself (a race would show as run-to-run differences)
m.stream_ordered_launches = False
gi = d.GraphedInference(m, x, lm, return_features=True)
ref = [t.clone() for t in gi.replay()]
bad = 0
for i in range(50):
    lo, fe = gi.replay()
    torch.cuda.synchronize()
    bad += int(not (torch.equal(lo, ref[0]) and torch.equal(fe, ref[1])))
print("replays differing from the first:", bad, "of 50")
