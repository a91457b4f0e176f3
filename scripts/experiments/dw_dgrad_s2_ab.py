"""Same-box A/B of the stride-2 depthwise input-gradient kernel (dfv_dwconv_dgrad) between two builds of libdfvit.so.

    python scripts/experiments/dw_dgrad_s2_ab.py <tag> [<other tag to compare the saved outputs with>]

Runs the four stride-2 layers of EfficientNet-B4 at the training batch (64 images, 380 x 380 input) in bf16 and fp32 with
whatever deepfake_vit_b200/libdfvit.so is in place, prints the CUDA-event time per launch (L2 flushed between launches) and
writes the outputs to /tmp/dw_dgrad_<tag>.pt; with a second tag the outputs are compared bit for bit with that run's.
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from deepfake_vit_b200 import ops  # noqa: E402

LAYERS = [(3, 0, 1, 144, 190), (5, 1, 2, 192, 95), (3, 0, 1, 336, 48), (5, 1, 2, 960, 24)]      # kernel, pad_lo, pad_hi, C, H (TF 'same' pads)


def main():
    tag = sys.argv[1]
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(7)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    outs, rows = {}, []
    for dt in (torch.bfloat16, torch.float32):
        for (k, pl, ph, C, H) in LAYERS:
            Ho = (H + pl + ph - k) // 2 + 1
            g = torch.randn(64, Ho, Ho, C, device=dev, generator=gen).to(dt)
            w = torch.randn(k * k, C, device=dev, generator=gen) * 0.3
            for _ in range(3):
                dx = ops.dwconv_dgrad(g, w, H, H, k, 2, pl, ph)
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dx = ops.dwconv_dgrad(g, w, H, H, k, 2, pl, ph)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            nbytes = (g.numel() + dx.numel()) * g.element_size()
            rows.append({"dtype": str(dt), "k": k, "C": C, "H": H, "us": round(ts[len(ts) // 2] * 1e3, 1),
                         "GBps": round(nbytes / ts[len(ts) // 2] / 1e6)})
            outs[f"{dt}_{k}_{C}_{H}"] = dx.cpu()
    for r in rows:
        print(json.dumps(r))
    torch.save(outs, f"/tmp/dw_dgrad_{tag}.pt")
    if len(sys.argv) > 2:
        other = torch.load(f"/tmp/dw_dgrad_{sys.argv[2]}.pt")
        for key, v in outs.items():
            same = torch.equal(v, other[key])
            md = (v.double() - other[key].double()).abs().max().item()
            print(f"{key}: bit-identical to {sys.argv[2]}: {same} (max abs diff {md:.3e})")


if __name__ == "__main__":
    main()
