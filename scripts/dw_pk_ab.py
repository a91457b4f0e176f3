"""A/B of the depthwise strip variants on the 5x5 stride-1 layers of B4 at batch 256 (bf16): the 4-channel packed-fp32 (FFMA2)
strips (L = 12) against the 8-channel FHFMA strips (L = 6 / 4), bit-equality of outputs and pool sums, time per launch.
Activations of one layer (0.4-0.8 GB) exceed L2, so back-to-back launches are HBM-cold."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d

ops, lib, DEV = d.ops, d._lib.lib, "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
layers = [(336, 48), (672, 24), (960, 24), (1632, 12)]


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


g = torch.Generator(device=DEV).manual_seed(0)
report = {}
for (Cc, H) in layers:
    x = torch.randn(B, H, H, Cc, device=DEV, generator=g).bfloat16()
    w = torch.randn(25, Cc, device=DEV, generator=g) * 0.2
    bias = torch.randn(Cc, device=DEV, generator=g) * 0.1
    info = (C.c_int * 10)()
    lib.dfv_dwconv_plan_info(1, B, H, H, Cc, 5, 1, 2, 2, info)
    nbytes = 4.0 * B * Cc * H * H
    row = {"planner": list(info)}
    ref = None
    for name, tn in (("planner", None), ("L6", (6, 0, 0, 0)), ("L4", (4, 0, 0, 0)), ("L12", (12, 0, 0, 0))):
        try:
            y, pool = ops.dwconv(x, w, bias, 5, 1, 2, 2, tuning=tn)
        except Exception as e:
            print(name, "failed:", str(e)[:100]); continue
        if name == "L6":
            ref = (y, pool.sum(1))
        ms = timed(lambda: ops.dwconv(x, w, bias, 5, 1, 2, 2, tuning=tn))
        row[name] = dict(us=ms * 1e3, gbs=nbytes / ms / 1e6)
        if ref is not None and name != "L6":
            row[name]["equal_y"] = bool(torch.equal(y, ref[0]))
            row[name]["pool_rel"] = ((pool.sum(1) - ref[1]).norm() / ref[1].norm()).item()
    for act in (0,):
        y0, _ = ops.dwconv(x, w, bias, 5, 1, 2, 2, act=act, want_pool=False, tuning=(6, 0, 0, 0))
        y1, _ = ops.dwconv(x, w, bias, 5, 1, 2, 2, act=act, want_pool=False, tuning=(12, 0, 0, 0))
        row["noact_equal"] = bool(torch.equal(y0, y1))
    report[f"C{Cc} {H}x{H}"] = row
    print(f"C{Cc} {H}x{H}:", json.dumps(row), flush=True)
    del x
print(json.dumps(report))
