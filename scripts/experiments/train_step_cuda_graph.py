"""Experiment: does one training step (forward + CombinedLoss + backward through the native sequencer) capture into a
CUDA graph, and what does a replay cost next to the eager step?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import deepfake_vit_b200 as d
B = 64
torch.manual_seed(42)
m = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG).cuda().train().set_compute_dtype(torch.bfloat16)
crit = d.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, torch.tensor([1.0, 1.5], device="cuda"))
x = torch.randn(B, 3, 380, 380, device="cuda"); lm = torch.rand(B, 5, 2, device="cuda") * 380; y = torch.randint(0, 2, (B,), device="cuda")
def step():
    lo, fe = m(x, lm, return_features=True)
    loss = crit(lo, y, fe)["total"]
    loss.backward()
    return loss
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        m.zero_grad(set_to_none=True)
        step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
def timeit(fn, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def eager():
    m.zero_grad(set_to_none=True); step()
print(f"eager: {timeit(eager):.2f} ms/step")
try:
    g = torch.cuda.CUDAGraph()
    m.zero_grad(set_to_none=True)
    with torch.cuda.graph(g):
        loss = step()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    eager_loss = None
    print(f"graph replay: {timeit(g.replay):.2f} ms/step, loss {loss.item():.6f}")
    m.zero_grad(set_to_none=True); ref = step(); torch.cuda.synchronize()
    print(f"eager loss {ref.item():.6f}")
except Exception as e:
    print("capture failed:", type(e).__name__, str(e)[:400])
