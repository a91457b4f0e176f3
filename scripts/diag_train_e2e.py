"""Where does a training step's wall time go when the loss is read back every step?  (host enqueue vs GPU)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG).to(dev).train().set_compute_dtype(torch.bfloat16)
crit = d.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, torch.tensor([1.0, 1.5], device=dev))
B = 64
x = torch.randn(B, 3, 380, 380, device=dev); lm = torch.rand(B, 5, 2, device=dev) * 380; y = torch.randint(0, 2, (B,), device=dev)
u8 = torch.randint(0, 256, (B, 380, 380, 3), device=dev, dtype=torch.uint8)
def step(xx, sync):
    m.zero_grad(set_to_none=True)
    t0 = time.perf_counter()
    lo, fe = m(xx, lm, return_features=True)
    t1 = time.perf_counter()
    loss = crit(lo, y, fe)["total"]
    loss.backward()
    t2 = time.perf_counter()
    if sync:
        loss.item()
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2
for name, xx in (("fp32", x), ("uint8", u8)):
    for sync in (False, True):
        for _ in range(3): step(xx, sync)
        torch.cuda.synchronize(); t = time.perf_counter(); acc = [0, 0, 0]
        for _ in range(10):
            r = step(xx, sync); acc = [a + b for a, b in zip(acc, r)]
        torch.cuda.synchronize(); tot = (time.perf_counter() - t) / 10
        print(f"{name} sync={sync}: {tot*1e3:.1f} ms/step; host fwd {acc[0]*100:.1f} ms, host loss+bwd {acc[1]*100:.1f} ms, item {acc[2]*100:.1f} ms", flush=True)
