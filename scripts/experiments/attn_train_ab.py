"""Same-box A/B of the training-path HybridAttention kernels (dfv_hybrid_attention_train_fwd / dfv_hybrid_attention_bwd) between
two builds of libdfvit.so, at the training step's shape (64 images, 12 x 12 x 1792 map, hidden 112, bf16 and fp32).

    python scripts/experiments/attn_train_ab.py <tag> [<other tag to compare the saved outputs with>]
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from deepfake_vit_b200 import ops  # noqa: E402


def timed(fn, flush, n=10):
    for _ in range(3):
        out = fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return out, ts[len(ts) // 2] * 1e3


def main():
    tag = sys.argv[1]
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(11)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B, H, W, C, hid = 64, 12, 12, 1792, 112
    outs = {}
    for dt in (torch.bfloat16, torch.float32):
        x = (torch.randn(B, H, W, C, device=dev, generator=gen) * 0.8).to(dt)
        heat = torch.rand(B, H, W, device=dev, generator=gen).clamp_(0.1, 1.0)
        w1 = torch.randn(hid, C, device=dev, generator=gen) * 0.05
        w2 = torch.randn(C, hid, device=dev, generator=gen) * 0.1
        sa = torch.randn(98, device=dev, generator=gen) * 0.3
        df = torch.randn(B, C, device=dev, generator=gen)
        (f, saved), t_f = timed(lambda: ops.hybrid_attention_train(x, heat, w1, w2, sa, True, True), flush)
        res, t_b = timed(lambda: ops.hybrid_attention_bwd(x, heat, w1, w2, sa, df, saved, True, True), flush)
        print(json.dumps({"dtype": str(dt), "fwd_us": round(t_f, 1), "bwd_us": round(t_b, 1)}))
        outs[str(dt)] = [f.cpu(), saved.cpu()] + [r.float().cpu() for r in res]
    torch.save(outs, f"/tmp/attn_train_{tag}.pt")
    if len(sys.argv) > 2:
        other = torch.load(f"/tmp/attn_train_{sys.argv[2]}.pt")
        names = ["features", "saved", "dfmap", "dheat", "dw1", "dw2", "dsa_w"]
        for key, vs in outs.items():
            for name, v, o in zip(names, vs, other[key]):
                r = ((v.double() - o.double()).norm() / o.double().norm().clamp_min(1e-30)).item()
                print(f"{key} {name}: rel L2 diff vs {sys.argv[2]} {r:.3e}")


if __name__ == "__main__":
    main()
