"""Blocks 0 / 1 depthwise (C = 48 / 24 at 190x190, 3x3): the planner's chunk (CB = C, shared-memory pixel pitch 96 / 48 bytes, bank
conflicts) against a forced 64-channel chunk (pitch 128 bytes, conflict-free, 25 % / 62 % of the lanes masked)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import deepfake_vit_b200 as d
ops = d.ops
g = torch.Generator(device="cuda").manual_seed(0)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for C in (48, 24, 144, 192):
    H = 190 if C <= 48 else 95
    x = torch.randn(256, H, H, C, device="cuda", generator=g).bfloat16()
    w = torch.randn(9, C, device="cuda", generator=g) * 0.3
    b = torch.randn(C, device="cuda", generator=g) * 0.1
    ref, pref = ops.dwconv(x, w, b, 3, 1, 1, 1)
    t0 = timed(lambda: ops.dwconv(x, w, b, 3, 1, 1, 1))
    line = f"C{C} {H}x{H} k3: planner {t0:.1f} us"
    for tn in ((0, 0, 0, 64), (4, 0, 0, 64), (6, 0, 0, 64), (8, 0, 0, 64)):
        try:
            y, pool = ops.dwconv(x, w, b, 3, 1, 1, 1, tuning=tn)
            same = torch.equal(y, ref)
            t = timed(lambda: ops.dwconv(x, w, b, 3, 1, 1, 1, tuning=tn))
            line += f" | L{tn[0]} CB64 {t:.1f} us {'==' if same else '!='}"
        except Exception as e:
            line += f" | L{tn[0]} CB64 failed ({str(e)[:40]})"
    print(line)
