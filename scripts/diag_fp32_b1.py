"""fp32 parity at batch 1 @ 380x380 (the case that sits closest to the 1e-4 bar): our fp32 path and the fp32 CPU oracle, each
against the oracle run in float64, plus run-to-run determinism of our path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import copy
import torch
import deepfake_vit_b200 as d
from oracle import calibrate, refmodel


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


om = calibrate.build(refmodel.get_oracle(), "calibrated", calib_size=128, calib_batches=2).eval()
m = d.DeepfakeDetectionModel(**refmodel.MODEL_CONFIG)
m.load_state_dict(om.state_dict(), strict=True)
m = m.cuda().eval()
m.set_compute_dtype(torch.float32)
om64 = copy.deepcopy(om).double()
print("cpu threads", torch.get_num_threads())
for shape in [(1, 380, 380), (1, 96, 96), (3, 250, 190), (2, 129, 161), (5, 64, 64), (8, 380, 380)]:
    B, H, W = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 3, H, W, generator=g)
    lm = torch.rand(B, 5, 2, generator=g) * min(H, W)
    with torch.no_grad():
        ref, fref = om(x, lm, return_features=True)
        r64, f64 = om64(x.double(), lm.double(), return_features=True)
        torch.set_num_threads(1)
        ref1, fref1 = om(x, lm, return_features=True)
        torch.set_num_threads(os.cpu_count())
    runs = [m(x.cuda(), lm.cuda(), return_features=True) for _ in range(3)]
    det = all(torch.equal(runs[0][1], r[1]) for r in runs[1:])
    out, f = runs[0]
    print(shape, f"ours-vs-ref32 f {rel(f, fref):.3e} l {rel(out, ref):.3e} | ours-vs-ref64 f {rel(f, f64):.3e} l {rel(out, r64):.3e} | "
          f"ref32-vs-ref64 f {rel(fref, f64):.3e} l {rel(ref, r64):.3e} | ref32(1 thread)-vs-ref32 f {rel(fref1, fref):.3e} | deterministic {det}", flush=True)
