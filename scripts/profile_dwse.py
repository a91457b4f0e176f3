"""One depthwise launch with the fused squeeze tail on a late-stage shape (for ncu --set full)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d
ops = d.ops
B, C_, H, k, sq = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 1632, int(sys.argv[2]) if len(sys.argv) > 2 else 12, 5, int(sys.argv[3]) if len(sys.argv) > 3 else 68
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, H, H, C_, device="cuda", generator=g).bfloat16()
w = torch.randn(k * k, C_, device="cuda", generator=g) * 0.2
bias = torch.randn(C_, device="cuda", generator=g) * 0.1
w1 = torch.randn(sq, C_, device="cuda", generator=g) / C_ ** 0.5
for _ in range(3):
    ops.dwconv(x, w, bias, k, 1, 2, 2)
    ops.dwconv_se(x, w, bias, k, 1, 2, 2, w1)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(10):
    ops.dwconv(x, w, bias, k, 1, 2, 2)
e[1].record()
for _ in range(10):
    ops.dwconv_se(x, w, bias, k, 1, 2, 2, w1)
e[2].record()
torch.cuda.synchronize()
print(f"C{C_} {H}x{H} sq{sq}: plain {e[0].elapsed_time(e[1])*100:.1f} us, with fused squeeze {e[1].elapsed_time(e[2])*100:.1f} us (incl. the wrapper's allocations)")
