"""Tensor-core depthwise kernel (csrc/dwconv_tc.cu) against the SIMT kernel and torch on the 5x5 stride-1 layers of B4 at
batch 256: element-wise error, pool sums, and time per launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import deepfake_vit_b200 as d
ops, lib = d.ops, d._lib.lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device="cuda").manual_seed(3)
for (C, H, K) in ((336, 48, 5), (960, 24, 5), (1632, 12, 5), (672, 24, 3), (2688, 12, 3)):
    pad = K // 2
    x = torch.randn(B, H, H, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(K * K, C, device="cuda", generator=g) * 0.2).bfloat16().float()
    bias = torch.randn(C, device="cuda", generator=g) * 0.1
    y_simt, pool_simt = ops.dwconv(x, w, bias, K, 1, pad, pad)
    parts = lib.dfv_dwconv_tc_pool_parts(H, H, C, K, pad, pad)
    y = torch.empty_like(y_simt)
    pool = torch.zeros(B, parts, C, device="cuda")
    def run():
        d._lib.check(lib.dfv_dwconv_tc_fwd(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), pool.data_ptr(), B, H, H, C, K, pad, pad, 1, None))
    run(); torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2)
    ref = F.conv2d(F.pad(xr, (pad, pad, pad, pad)), w.t().reshape(C, 1, K, K), bias, groups=C)
    ref = (ref * torch.sigmoid(ref)).permute(0, 2, 3, 1)
    bad = ((y.float() - ref).abs() > 0.03 * (ref.abs() + 1.0)).sum().item()
    bad_s = ((y_simt.float() - ref).abs() > 0.03 * (ref.abs() + 1.0)).sum().item()
    diff = (y.float() - y_simt.float()).abs().max().item()
    pr = (pool.sum(1) - pool_simt.sum(1)).abs().max().item() / (pool_simt.sum(1).abs().max().item() + 1e-9)
    def t(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10 * 1000
    t_tc, t_simt = t(run), t(lambda: ops.dwconv(x, w, bias, K, 1, pad, pad))
    gb = 2 * B * H * H * C * 2 / 1e9
    print(f"C={C} H={H} k={K}: bad vs torch tc {bad} simt {bad_s}; max |tc - simt| {diff:.4f}; pool rel {pr:.2e}; "
          f"tc {t_tc:.0f} us ({gb / t_tc * 1e3:.2f} TB/s)  simt {t_simt:.0f} us ({gb / t_simt * 1e3:.2f} TB/s)  parts {parts}")
