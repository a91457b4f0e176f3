"""Stress the tcgen05 GEMM: many back-to-back launches per shape; reports the first failing shape."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d
ops = d.ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
shapes = []  # (M, K, N, scale, residual, act)
size = 190
for b in d._lib.b4_blocks():
    hin = size; size = (size + b.stride - 1) // b.stride
    if b.has_expand: shapes.append((B * hin * hin, b.c_in, b.c_mid, False, False, 1))
    shapes.append((B * size * size, b.c_mid, b.c_out, True, bool(b.has_skip), 0))
shapes.append((B * 144, 448, 1792, False, False, 1))
seen = set()
for (M, K, N, sc, res, act) in shapes:
    key = (M, K, N, sc, res)
    if key in seen: continue
    seen.add(key)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda")
    rpi = M // B
    scale = torch.rand(B, K, device="cuda").bfloat16() if sc else None
    r = torch.randn(M, N, device="cuda").bfloat16() if res else None
    try:
        n = max(20, min(iters, int(iters * 2e8 / (M * (K + N)))))
        for i in range(n):
            out = ops.pw_gemm(a, w, bias, act, scale, rpi if sc else 0, r)
        torch.cuda.synchronize()
        print("ok  ", key, n, flush=True)
    except Exception as e:
        print("FAIL", key, str(e)[:200], "word=0x%08x" % d._lib.lib.dfv_debug_last_timeout(), flush=True)
        break
    del a, w, out, r, scale
