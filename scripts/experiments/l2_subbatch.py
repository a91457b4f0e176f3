"""Experiment: does running a stage's MBConv blocks DEPTH-FIRST over sub-batches keep the 6x-expanded tensors in the 126 MB L2?

For each stride-1 B4 stage shape: a chain of `nb` identical blocks (expand 1x1 + swish -> depthwise + swish + SE pool -> SE gate ->
gated project 1x1 + residual) at batch 256, bf16, run (a) layer by layer over the whole batch (what dfv_infer_fwd does) and (b) sub-batch
by sub-batch, each sub-batch through all nb blocks before the next one starts, the expanded / depthwise scratch reused by every
sub-batch so that it can stay L2-resident.  Both captured into CUDA graphs; prints ms per variant.  Library calls with caller-owned
buffers only (no allocation inside the captured region)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import deepfake_vit_b200 as d
from deepfake_vit_b200 import _lib

lib, check = _lib.lib, _lib.check
DEV, BF = "cuda", torch.bfloat16
DT = d.ops.dtype_code(BF)
B = 256
# (name, c_in, c_mid, H, kernel, pad_lo, pad_hi, squeeze, blocks)
STAGES = [("stage2 95x95 k3", 32, 192, 95, 3, 1, 1, 8, 3), ("stage3 48x48 k5", 56, 336, 48, 5, 2, 2, 14, 3),
          ("stage4 24x24 k3", 112, 672, 24, 3, 1, 1, 28, 5), ("stage5 24x24 k5", 160, 960, 24, 5, 2, 2, 40, 5),
          ("stage6 12x12 k5", 272, 1632, 12, 5, 2, 2, 68, 7)]


def p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


f = p


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


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
    return e0.elapsed_time(e1) / n


g = torch.Generator(device=DEV).manual_seed(0)
for name, cin, cmid, H, k, pl, ph, sq, nb in STAGES:
    hw = H * H
    we = [(torch.randn(cmid, cin, device=DEV, generator=g) * 0.1).to(BF) for _ in range(nb)]
    be = [torch.randn(cmid, device=DEV, generator=g) * 0.1 for _ in range(nb)]
    wd = [torch.randn(k * k, cmid, device=DEV, generator=g) * 0.2 for _ in range(nb)]
    bd = [torch.randn(cmid, device=DEV, generator=g) * 0.1 for _ in range(nb)]
    wp = [(torch.randn(cin, cmid, device=DEV, generator=g) * 0.05).to(BF) for _ in range(nb)]
    bp = [torch.randn(cin, device=DEV, generator=g) * 0.1 for _ in range(nb)]
    w1 = [torch.randn(sq, cmid, device=DEV, generator=g) * 0.05 for _ in range(nb)]
    b1 = [torch.zeros(sq, device=DEV) for _ in range(nb)]
    w2t = [torch.randn(sq, cmid, device=DEV, generator=g) * 0.05 for _ in range(nb)]
    b2 = [torch.zeros(cmid, device=DEV) for _ in range(nb)]
    x0 = torch.randn(B, H, H, cin, device=DEV, generator=g).to(BF)
    acts = [x0] + [torch.empty_like(x0) for _ in range(nb)]        # block outputs: full batch, every layer its own buffer

    def run(sub, results):
        """sub = images per sub-batch; scratch sized for one sub-batch and shared by all of them."""
        exp = torch.empty(sub, H, H, cmid, device=DEV, dtype=BF)
        dwo = torch.empty(sub, H, H, cmid, device=DEV, dtype=BF)
        parts = lib.dfv_dwconv_pool_parts_tuned(DT, sub, H, H, cmid, k, 1, pl, ph, None)
        pool = torch.empty(sub, parts, cmid, device=DEV, dtype=torch.float32)
        gate = torch.empty(sub, cmid, device=DEV, dtype=BF)
        scratch = torch.empty(max(1, lib.dfv_se_scratch_floats(sub, cmid, sq)), device=DEV, dtype=torch.float32)

        def body():
            st = stream()
            for s0 in range(0, B, sub):
                for i in range(nb):
                    xin, xout = acts[i][s0:s0 + sub], acts[i + 1][s0:s0 + sub]
                    check(lib.dfv_pw_gemm_fwd_tuned(p(xin), p(we[i]), f(be[i]), None, 0, None, p(exp), DT, sub * hw, cin, cmid, 1, None, st))
                    check(lib.dfv_dwconv_fwd_tuned(p(exp), f(wd[i]), f(bd[i]), p(dwo), p(pool), DT, sub, H, H, cmid, k, 1, pl, ph, 1, None, st))
                    check(lib.dfv_se_gate_fwd(f(pool), parts, 1.0 / hw, f(w1[i]), f(b1[i]), f(w2t[i]), f(b2[i]), p(gate), DT, f(scratch),
                                              sub, cmid, sq, st))
                    check(lib.dfv_pw_gemm_fwd_tuned(p(dwo), p(wp[i]), f(bp[i]), p(gate), hw, p(xin), p(xout), DT, sub * hw, cmid, cin, 0, None, st))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            body()
        ms = timed(gr.replay)
        results.append((sub, ms, acts[nb].float().abs().mean().item()))

    res = []
    for sub in (256, 128, 64, 32, 16, 8):
        if sub * hw * cmid * 2 * 2 > 400e6 and sub < 16:
            continue
        run(sub, res)
    mb = hw * cmid * 2 / 1e6
    print(f"{name}: C {cin}->{cmid}, {nb} blocks, expanded tensor {mb:.2f} MB/image")
    for sub, ms, chk in res:
        print(f"   sub-batch {sub:4d} (scratch {2 * sub * mb:7.1f} MB): {ms:7.3f} ms   [mean |out| {chk:.4f}]")
