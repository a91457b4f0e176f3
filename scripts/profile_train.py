"""A few batch-64 bf16 training steps for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(42)
m = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG).cuda().train()
crit = d.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, torch.tensor([1.0, 1.5], device="cuda"))
x = torch.randn(B, 3, 380, 380, device="cuda")
lm = torch.rand(B, 5, 2, device="cuda") * 380
y = torch.randint(0, 2, (B,), device="cuda")
for i in range(n):
    m.zero_grad(set_to_none=True)
    lo, fe = m(x, lm, return_features=True)
    crit(lo, y, fe)["total"].backward()
torch.cuda.synchronize()
print("ok", d._lib.lib.dfv_launch_count(0))
