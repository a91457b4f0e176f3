import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d
B = 8
torch.manual_seed(42)
m = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG).cuda().train().set_compute_dtype(torch.bfloat16)
crit = d.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, torch.tensor([1.0, 1.5], device="cuda"))
x = torch.randn(B, 3, 96, 96, device="cuda"); lm = torch.rand(B, 5, 2, device="cuda") * 96; y = torch.randint(0, 2, (B,), device="cuda")
def step():
    lo, fe = m(x, lm, return_features=True)
    loss = crit(lo, y, fe)["total"]
    loss.backward()
    return loss
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        m.zero_grad(set_to_none=True); step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
for mode in ("global", "thread_local", "relaxed"):
    for seed_dev in (False, True):
        m._seed_dev = torch.zeros(1, dtype=torch.int64, device="cuda") if seed_dev else None
        try:
            g = torch.cuda.CUDAGraph()
            m.zero_grad(set_to_none=True)
            with torch.cuda.graph(g, capture_error_mode=mode):
                loss = step()
            torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
            print(mode, "seed_dev", seed_dev, "OK loss", loss.item(), flush=True)
        except Exception as e:
            print(mode, "seed_dev", seed_dev, "FAILED", type(e).__name__, str(e)[:160].replace("\n", " "), flush=True)
            torch.cuda.synchronize()
