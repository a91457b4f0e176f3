"""CTA-pair (cta_group::2) GEMM plans on the GPU: for a list of shapes run planner / single-CTA / pair plans, print the plan,
the error (if any) and the worst element error against an fp32 matmul of the same bf16 operands, and the time."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d
ops, lib, DEV = d.ops, d._lib.lib, "cuda"
shapes = [(396, 672, 112, 0, 0), (396, 1632, 272, 1, 36), (36864, 1632, 272, 1, 144), (36864, 272, 1632, 0, 0), (147456, 672, 112, 1, 576),
          (147461, 960, 160, 1, 147461), (36864, 448, 1792, 0, 0), (36864, 2688, 448, 1, 144), (36864, 1632, 448, 1, 144), (36864, 960, 272, 1, 144), (36864, 448, 2688, 0, 0), (147456, 160, 960, 0, 0)]
if len(sys.argv) > 1:      # comma-separated shape indices; argv[2] = timing iterations (0: one call per plan, for ncu)
    shapes = [shapes[int(i)] for i in sys.argv[1].split(",")]
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g = torch.Generator(device=DEV).manual_seed(3)
info = (C.c_int * 10)()
for (M, K, N, gated, rpi) in shapes:
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    sc = torch.rand(M // rpi, K, device=DEV, generator=g).bfloat16() if gated else None
    res = torch.randn(M, N, device=DEV, generator=g).bfloat16() if gated else None
    av = a
    if gated:
        av = (a * sc[torch.arange(M, device=DEV) // rpi])
    ref = av.float() @ w.float().t() + bias
    ref = ref * torch.sigmoid(ref) if not gated else ref + res.float()
    lib.dfv_gemm_plan_info(C.c_longlong(M), K, N, gated, info)
    print(f"M={M} K={K} N={N} gated={gated}: planner {list(info)}", flush=True)
    for tn in ((0, 0, 2, -1, -1), (0, 0, 2, -1, 1), (0, 0, 2, 1, 1), (0, 0, -1, -1, -1), (0, 0, -1, -1, 1)):     # (ws, bn, cluster, share_a, rotate)
        try:
            run = lambda: ops.pw_gemm(a, w, bias, 0 if gated else 1, sc, rpi, res, tuning=tn)
            y = run(); torch.cuda.synchronize()
            bad = ((y.float() - ref).abs() > 0.03 * (ref.abs() + 1.0)).sum().item()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(ITERS): run()
            e1.record(); torch.cuda.synchronize()
            print(f"   tuning {tn}: bad elements {bad} of {y.numel()}, {e0.elapsed_time(e1) / max(ITERS, 1) * 1e3:.1f} us", flush=True)
        except Exception as e:
            print(f"   tuning {tn}: ERROR {str(e)[:300]}", flush=True)
