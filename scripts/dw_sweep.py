"""Measured tile-plan search for the depthwise kernel: every distinct stride-1 layer of B4 at batch 256 (bf16) under the
planner's choice and under restricted plans (per-call dfv_dwconv_tuning: L, TW, TH, CB).  Prints the best few per layer."""
import ctypes as C, itertools, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d

ops, lib, DEV = d.ops, d._lib.lib, "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seen, layers, h = set(), [], 190
for b in d._lib.b4_blocks():
    ho = (h + b.pad_lo + b.pad_hi - b.kernel) // b.stride + 1
    key = (b.c_mid, h, b.kernel, b.stride)
    if b.stride == 1 and key not in seen:
        seen.add(key); layers.append((b.c_mid, h, b.kernel, b.pad_lo, b.pad_hi))
    h = ho


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


g = torch.Generator(device=DEV).manual_seed(0)
report = {}
for (Cc, H, K, pl, ph) in layers:
    x = torch.randn(B, H, H, Cc, device=DEV, generator=g).bfloat16()
    w = torch.randn(K * K, Cc, device=DEV, generator=g) * 0.2
    bias = torch.randn(Cc, device=DEV, generator=g) * 0.1
    ref, ref_pool = ops.dwconv(x, w, bias, K, 1, pl, ph)
    info = (C.c_int * 10)()
    lib.dfv_dwconv_plan_info(1, B, H, H, Cc, K, 1, pl, ph, info)
    base = timed(lambda: ops.dwconv(x, w, bias, K, 1, pl, ph))
    nbytes = 4.0 * B * Cc * H * H
    res = []
    tws = sorted({t for t in (8, 12, 16, 24, 32, 36, 40, 48) if t - 8 < H})
    cbs = [cb for cb in (64, 56, 48, 40, 32, 24) if Cc % cb == 0 or (cb == 64 and Cc > 64)]
    for L, TW, TH, CB in itertools.product((4, 6, 8), tws, (4, 6, 8, 10, 12, 16), cbs):
        if TW % L:
            continue
        try:
            y, pool = ops.dwconv(x, w, bias, K, 1, pl, ph, tuning=(L, TW, TH, CB))
        except Exception:
            continue
        if not torch.equal(y, ref):
            print("MISMATCH", (Cc, H, K), (L, TW, TH, CB), flush=True)
            continue
        ms = timed(lambda: ops.dwconv(x, w, bias, K, 1, pl, ph, tuning=(L, TW, TH, CB)), 3)
        res.append((ms, (L, TW, TH, CB)))
    res.sort()
    report[f"C{Cc} {H}x{H} k{K}"] = dict(planner=dict(plan=list(info)[:4], us=base * 1e3, gbs=nbytes / base / 1e6),
                                         best=[dict(plan=p, us=ms * 1e3, gbs=nbytes / ms / 1e6) for ms, p in res[:5]])
    print(f"C{Cc} {H}x{H} k{K}: planner {list(info)[:4]} {base*1e3:.0f} us | best " + ", ".join(f"{p} {ms*1e3:.0f}" for ms, p in res[:4]), flush=True)
    del x, ref
print(json.dumps(report))
