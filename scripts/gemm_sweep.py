"""Time the tensor-core 1x1-conv GEMM for the B4 layer shapes under forced tile plans and check every plan against a
torch fp32 matmul (DFV_GEMM_FORCE="weight_stationary,BN" is read per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d
ops = d.ops
B = 256
# (name, M, K, N, gated, rows_per_image, act, residual) after row folding
L = [("b1 project", 2310400, 96, 96, 1, 9025, 0, 1), ("b2 expand", 2310400, 96, 576, 0, 0, 1, 0), ("b3 expand", 1155200, 64, 384, 0, 0, 1, 0),
     ("b3 project", 2310400, 192, 32, 1, 9025, 0, 1), ("b7 expand", 589824, 56, 336, 0, 0, 1, 0), ("b7 project", 589824, 336, 56, 1, 2304, 0, 1),
     ("b11 expand", 147456, 112, 672, 0, 0, 1, 0), ("b11 project", 147456, 672, 112, 1, 576, 0, 1), ("b17 expand", 147456, 160, 960, 0, 0, 1, 0),
     ("b17 project", 147456, 960, 160, 1, 576, 0, 1), ("b23 expand", 36864, 272, 1632, 0, 0, 1, 0), ("b23 project", 36864, 1632, 272, 1, 144, 0, 1),
     ("b31 expand", 36864, 448, 2688, 0, 0, 1, 0), ("b31 project", 36864, 2688, 448, 1, 144, 0, 1), ("head", 36864, 448, 1792, 0, 0, 1, 0)]
only = sys.argv[1:]
for (name, M, K, N, g, rpi, act, res) in L:
    if only and not any(o in name for o in only): continue
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda") * 0.1
    sc = torch.rand(M // rpi, K, device="cuda").bfloat16() if g else None
    r = torch.randn(M, N, device="cuda").bfloat16() if res else None
    # reference on a row sample
    idx = torch.cat([torch.arange(0, min(M, 300), device="cuda"), torch.randint(0, M, (1500,), device="cuda"), torch.arange(M - 300, M, device="cuda")])
    av = a[idx].float()
    if g: av = (a[idx] * sc[idx // rpi]).float()
    ref = av @ w.float().t() + bias
    if act: ref = ref * torch.sigmoid(ref)
    if res: ref = ref + r[idx].float()
    parts = 2 if g else 4
    bns = [parts * c for c in (16, 32, 48, 64, 96, 128) if parts * c <= 256]
    out = []
    os.environ["DFV_GEMM_FORCE"] = "-1,0"
    for _ in range(10): ops.pw_gemm(a, w, bias, act=act, a_scale=sc, rows_per_image=rpi, residual=r)     # clocks up
    for cfg in ["-1,0"] + [f"0,{bn}" for bn in bns] + [f"1,{bn}" for bn in bns]:
        os.environ["DFV_GEMM_FORCE"] = cfg
        try:
            y = ops.pw_gemm(a, w, bias, act=act, a_scale=sc, rows_per_image=rpi, residual=r)
            err = ((y[idx].float() - ref).abs() / (ref.abs() + 1.0)).max().item()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): ops.pw_gemm(a, w, bias, act=act, a_scale=sc, rows_per_image=rpi, residual=r)
            e1.record(); torch.cuda.synchronize()
            out.append((e0.elapsed_time(e1) / 5 * 1000, cfg, err))
        except Exception as ex:
            if "fit" not in str(ex): print("   ", cfg, "failed:", str(ex)[:120])
    auto = out[0]
    out.sort()
    bad = [(c, round(e, 4)) for t, c, e in out if e > 2e-2]
    print(f"{name:12s} M={M} K={K} N={N}: auto {auto[0]:.0f} us; best:", [(round(t), c) for t, c, e in out[:5]], "BAD" if bad else "ok", bad)
