"""B200 only: the building block of the planned tensor-core depthwise kernel (csrc/dw_tc_probe.cu).
A tcgen05.mma A operand may start at ANY 128-byte pixel row of a SWIZZLE_128B tile written by TMA (plain descriptor, no
matrix-base-offset), so a depthwise tap is "the same tile viewed from pixel offset ky * TWI + kx"; against diagonal
16 x 16 weight blocks this gives the exact depthwise sum (bf16 products are exact in the fp32 accumulator)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("offs", [[0, 8, 16, 64], [0, 1, 2, 3, 4], [ky * 16 + kx for ky in range(5) for kx in range(5)],
                                  [ky * 28 + kx for ky in range(3) for kx in range(3)]])
def test_tcgen05_views_a_swizzled_tile_from_any_pixel_row(offs):
    import deepfake_vit_b200 as d
    g = torch.Generator(device="cuda").manual_seed(5)
    P, T = 200, len(offs)
    x = torch.randn(P, 64, device="cuda", generator=g).bfloat16()
    w = torch.randn(T, 64, device="cuda", generator=g).bfloat16()
    off_t = torch.tensor(offs, device="cuda", dtype=torch.int32)
    ref = torch.zeros(128, 64, device="cuda", dtype=torch.float64)
    for t, o in enumerate(offs):
        ref += x[o:o + 128].double() * w[t].double()
    out = torch.full((128, 64), float("nan"), device="cuda")
    d._lib.check(d._lib.lib.dfv_debug_dwconv_tc_probe(x.data_ptr(), w.data_ptr(), off_t.data_ptr(), T, P, 0, out.data_ptr(), None))
    torch.cuda.synchronize()
    assert (out.double() - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())
