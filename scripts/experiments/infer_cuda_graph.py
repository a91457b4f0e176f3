"""Experiment: batch-256 eval forward, eager launches vs one CUDA-graph replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import deepfake_vit_b200 as d
B = 256
torch.manual_seed(0)
m = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG).cuda().eval().set_compute_dtype(torch.bfloat16)
x = torch.randn(B, 3, 380, 380, device="cuda"); lm = torch.rand(B, 5, 2, device="cuda") * 380
def fwd():
    with torch.no_grad():
        return m(x, lm)[0]
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): fwd()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
def timeit(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(f"eager: {timeit(fwd):.3f} ms")
try:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fwd()
    ref = fwd(); g.replay(); torch.cuda.synchronize()
    print(f"graph replay: {timeit(g.replay):.3f} ms; identical logits: {torch.equal(out, ref)}")
except Exception as e:
    print("capture failed:", type(e).__name__, str(e)[:300])
