"""Measured tile-plan search for the tcgen05 1x1-conv GEMM: every distinct expand / project / head GEMM of B4 at batch 256
under the planner's choice and under restricted plans (per-call dfv_gemm_tuning: weight-stationary flag, N tile, 2 = CTA pair)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d

ops, lib, DEV = d.ops, d._lib.lib, "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shapes, seen, h = [], set(), 190
for b in d._lib.b4_blocks():
    ho = (h + b.pad_lo + b.pad_hi - b.kernel) // b.stride + 1
    cand = []
    if b.has_expand:
        cand.append((B * h * h, b.c_in, b.c_mid, False, h * h))
    cand.append((B * ho * ho, b.c_mid, b.c_out, True, ho * ho))
    for c in cand:
        if c not in seen and c[1] > 48:          # thin layers run row-folded (another plan family)
            seen.add(c); shapes.append(c)
    h = ho
shapes.append((B * 144, 448, 1792, False, 144))


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
for (M, K, N, gated, rpi) in shapes:
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    sc = torch.rand(M // rpi, K, device=DEV, generator=g).bfloat16() if gated else None
    act = 0 if gated else 1
    run = lambda tn=None: ops.pw_gemm(a, w, bias, act, sc, rpi if gated else 0, None, tuning=tn)
    ref = run()
    info = (C.c_int * 10)()
    lib.dfv_gemm_plan_info(C.c_longlong(M), K, N, int(gated), info)
    base = timed(run)
    res = []
    plans = [(ws, bn, cl) for ws in (0, 1) for bn in ((32, 64, 96, 128, 192, 256) if gated else (64, 128, 192, 256)) for cl in ((-1, 2) if ws == 0 else (-1,))]
    for (ws, bn, cl) in plans:
        if True:
            try:
                y = run((ws, bn, cl))
            except Exception:
                continue
            chk = (C.c_int * 8)()
            if not torch.equal(y, ref) and (y.float() - ref.float()).abs().max().item() > 0.05:
                print("MISMATCH", (M, K, N, gated), (ws, bn, cl), flush=True)
                continue
            res.append((timed(lambda: run((ws, bn, cl)), 3), (ws, bn, cl)))
    res.sort()
    nbytes = 2.0 * (M * K + M * N + N * K)
    key = f"M{M} K{K} N{N} {'gated' if gated else 'silu'}"
    report[key] = dict(planner=dict(bn=info[0], ws=info[1], cl=info[8], us=base * 1e3, gbs=nbytes / base / 1e6, tflops=2.0 * M * K * N / base / 1e9),
                       best=[dict(plan=p, us=ms * 1e3) for ms, p in res[:4]])
    print(f"{key}: planner bn={info[0]} ws={info[1]} cl={info[8]} {base*1e3:.0f} us | " + ", ".join(f"{p} {ms*1e3:.0f}" for ms, p in res[:4]), flush=True)
    del a, ref
print(json.dumps(report))
