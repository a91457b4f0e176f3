"""Time the depthwise kernel for one layer shape under forced tile plans (DFV_DW_FORCE is read per call)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d
ops = d.ops
B = 256
layers = [(336, 48, 5), (960, 24, 5), (1632, 12, 5), (192, 95, 3)]
for (C, H, K) in layers:
    x = torch.randn(B, H, H, C, device="cuda").bfloat16()
    w = torch.randn(K * K, C, device="cuda") * 0.1
    b = torch.randn(C, device="cuda") * 0.1
    pad = K // 2
    res = []
    for cfg in ["0,0,0,0"] + [f"{L},{TW},{TH},64" for L in (4, 6, 8) for TW in (12, 16, 24, 32, 48) for TH in (4, 6, 8, 12, 16)]:
        os.environ["DFV_DW_FORCE"] = cfg
        try:
            for _ in range(2): ops.dwconv(x, w, b, K, 1, pad, pad)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): ops.dwconv(x, w, b, K, 1, pad, pad)
            e1.record(); torch.cuda.synchronize()
            res.append((e0.elapsed_time(e1) / 5 * 1000, cfg))
        except Exception as ex:
            pass
    res.sort()
    base = [r for r in res if r[1] == "0,0,0,0"]
    print(f"C={C} H={H} k={K}: auto {base[0][0]:.0f} us; best:", [(round(t), c) for t, c in res[:6]])
