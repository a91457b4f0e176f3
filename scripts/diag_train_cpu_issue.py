"""How long does the host take to ENQUEUE one training step, next to how long the GPU takes to run it?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d
B = 64
torch.manual_seed(42)
m = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG).cuda().train().set_compute_dtype(torch.bfloat16)
crit = d.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, torch.tensor([1.0, 1.5], device="cuda"))
x = torch.randn(B, 3, 380, 380, device="cuda"); lm = torch.rand(B, 5, 2, device="cuda") * 380; y = torch.randint(0, 2, (B,), device="cuda")
def step():
    m.zero_grad(set_to_none=True)
    lo, fe = m(x, lm, return_features=True)
    crit(lo, y, fe)["total"].backward()
for _ in range(3): step()
torch.cuda.synchronize()
for trial in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); step(); e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"single step: host enqueue {1e3*(t1-t0):.1f} ms, GPU {e0.elapsed_time(e1):.1f} ms")
for n in (5, 10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): step()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{n} steps back to back: host enqueue {1e3*(t1-t0)/n:.1f} ms/step, GPU {e0.elapsed_time(e1)/n:.1f} ms/step")
print("cpu count", os.cpu_count(), "load", os.getloadavg())
