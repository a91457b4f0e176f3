"""One batch-256 bf16 forward (after N warm-up forwards) for ncu captures.  argv: batch, forwards, 'u8' for uint8 crops."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
u8 = len(sys.argv) > 3 and sys.argv[3] == "u8"
torch.manual_seed(42)
m = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG).cuda().eval()
x = torch.randint(0, 256, (B, 380, 380, 3), device="cuda", dtype=torch.uint8) if u8 else torch.randn(B, 3, 380, 380, device="cuda")
lm = torch.rand(B, 5, 2, device="cuda") * 380
import time
try:
    for i in range(n):
        t0 = time.time()
        out = m(x, lm)
        if n > 8:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
except Exception as e:
    print("FAILED at iteration", i, "after", round(time.time() - t0, 2), "s:", str(e)[:300])
    print("timeout word: 0x%08x" % d._lib.lib.dfv_last_timeout_word())
    sys.exit(1)
print("ok", out[0].shape, d._lib.lib.dfv_launch_count(0))
