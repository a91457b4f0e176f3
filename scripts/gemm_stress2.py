import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d
ops = d.ops
B, HW, K, N, sq = 256, 576, 960, 160, 40
M = B * HW
def run(tag, fn, n=1500):
    try:
        for i in range(n):
            fn()
        torch.cuda.synchronize()
        print("ok  ", tag, flush=True); return True
    except Exception as e:
        print("FAIL", tag, "iter", i, str(e)[:160], flush=True); return False
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
bias = torch.randn(N, device="cuda")
res = torch.randn(M, N, device="cuda").bfloat16()
pool = torch.randn(B, 2, K, device="cuda")
w1 = torch.randn(sq, K, device="cuda") * 0.1; b1 = torch.randn(sq, device="cuda")
w2t = torch.randn(sq, K, device="cuda") * 0.1; b2 = torch.randn(K, device="cuda")
gate_rand = torch.rand(B, K, device="cuda").bfloat16()
mode = sys.argv[1]
if mode == "gemm_only":
    run(mode, lambda: ops.pw_gemm(a, w, bias, 0, gate_rand, HW, res))
elif mode == "se_then_gemm":
    def f():
        g = ops.se_gate(pool, HW, w1, b1, w2t, b2, torch.bfloat16)
        ops.pw_gemm(a, w, bias, 0, g, HW, res)
    run(mode, f)
elif mode == "se_then_gemm_randgate":
    def f():
        ops.se_gate(pool, HW, w1, b1, w2t, b2, torch.bfloat16)
        ops.pw_gemm(a, w, bias, 0, gate_rand, HW, res)
    run(mode, f)
elif mode == "torchop_then_gemm":
    def f():
        pool.mul_(1.0)
        ops.pw_gemm(a, w, bias, 0, gate_rand, HW, res)
    run(mode, f)
