"""Probe for the planned tensor-core depthwise kernel: a tcgen05.mma A operand that starts at an
arbitrary 128-byte pixel row of a TMA-written SWIZZLE_128B tile, against diagonal weight blocks.  Prints, per descriptor
mode (0 = plain, 1 = matrix-base-offset field set), the error against torch for several sets of pixel offsets."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import deepfake_vit_b200 as d
lib = d._lib.lib
torch.manual_seed(0)
P = 200
x = torch.randn(P, 64, device="cuda").bfloat16()
for name, offs in (("aligned (multiples of 8 rows)", [0, 8, 16, 64]), ("row offsets 0..4", [0, 1, 2, 3, 4]),
                   ("5x5 taps of a 16-wide tile", [ky * 16 + kx for ky in range(5) for kx in range(5)])):
    T = len(offs)
    w = torch.randn(T, 64, device="cuda").bfloat16()
    off_t = torch.tensor(offs, device="cuda", dtype=torch.int32)
    ref = torch.zeros(128, 64, device="cuda")
    for t, o in enumerate(offs):
        ref += x[o:o + 128].float() * w[t].float()
    for mode in (0, 1):
        out = torch.full((128, 64), float("nan"), device="cuda")
        rc = lib.dfv_debug_dwconv_tc_probe(x.data_ptr(), w.data_ptr(), off_t.data_ptr(), T, P, mode, out.data_ptr(), None)
        torch.cuda.synchronize()
        err = (out - ref).abs().max().item()
        print(f"{name:32s} mode {mode}: rc {rc} max abs err {err:.3e} (ref max {ref.abs().max().item():.2f})")
