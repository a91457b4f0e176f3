"""Stem alone at batch 256 @ 380x380 (bf16 out): fp32 NCHW input and raw uint8 HWC input, us per launch (CUDA events, 30 launches
after 5 warm-ups; the 0.44 / 0.11 GB inputs + 0.89 GB output exceed the L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import deepfake_vit_b200 as d
ops = d.ops
B = 256
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, 3, 380, 380, device="cuda", generator=g)
u8 = torch.randint(0, 256, (B, 380, 380, 3), device="cuda", dtype=torch.uint8, generator=g)
w = torch.randn(27, 48, device="cuda", generator=g) * 0.2
b = torch.randn(48, device="cuda", generator=g) * 0.1


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


t32 = timed(lambda: ops.stem_conv(x, w, b, torch.bfloat16))
t8 = timed(lambda: ops.stem_conv_u8(u8, d.model.IMAGENET_MEAN, d.model.IMAGENET_STD, w, b, torch.bfloat16))
out_b = B * 190 * 190 * 48 * 2
print(f"stem fp32 NCHW in: {t32:.1f} us  ({(x.numel() * 4 + out_b) / t32 / 1e3:.0f} GB/s)   uint8 HWC in: {t8:.1f} us  ({(u8.numel() + out_b) / t8 / 1e3:.0f} GB/s)")
