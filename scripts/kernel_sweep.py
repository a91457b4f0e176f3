"""BASELINE.json configs[4]: isolated kernels at B4 stage shapes across batch 1..1024 (bf16, one B200).

  depthwise 5x5 (+ folded BN + swish + SE pool partials)  at (C, H, stride) of SURVEY 8(d) config 5
  squeeze-excite gate                                      at the same blocks' (C, squeeze, H*W)
  HybridAttention (heat-map + channel + spatial + pool)    at 1792 x 12 x 12
Every iteration is timed alone with CUDA events after an L2 flush (a 512 MB buffer is overwritten between iterations), so small
batches are measured HBM-cold like the layers inside a forward.  Prints one JSON object; --md writes a markdown table.
"""
import argparse, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d

ops, DEV = d.ops, "cuda"
PEAK = 6449.4
if os.path.isfile("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
DW = [(192, 95, 2, 1, 2), (336, 48, 1, 2, 2), (672, 24, 1, 2, 2), (960, 24, 1, 2, 2), (960, 24, 2, 1, 2), (1632, 12, 1, 2, 2)]   # C, H, stride, pad_lo, pad_hi
SE = [(192, 8, 48 * 48), (336, 14, 48 * 48), (672, 28, 24 * 24), (960, 40, 24 * 24), (960, 40, 12 * 12), (1632, 68, 12 * 12)]  # C, squeeze, HW
flush_buf = None


def timed(fn, iters):
    global flush_buf
    if flush_buf is None:
        flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--md", default=None)
    ap.add_argument("--max-batch", type=int, default=1024)
    a = ap.parse_args()
    batches = [b for b in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024) if b <= a.max_batch]
    g = torch.Generator(device=DEV).manual_seed(0)
    rows = []
    for (C, H, s, pl, ph) in DW:
        w = torch.randn(25, C, device=DEV, generator=g) * 0.2
        bias = torch.randn(C, device=DEV, generator=g) * 0.1
        Ho = (H + pl + ph - 5) // s + 1
        for B in batches:
            x = torch.randn(B, H, H, C, device=DEV, generator=g).bfloat16()
            ms = timed(lambda: ops.dwconv(x, w, bias, 5, s, pl, ph), a.iters)
            nbytes = 2.0 * B * C * (H * H + Ho * Ho)
            rows.append(dict(kernel="depthwise5x5", shape=f"C{C} {H}x{H} s{s}", batch=B, us=ms * 1e3, gbs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / PEAK))
            del x
    for (C, sq, hw) in SE:
        w1 = torch.randn(sq, C, device=DEV, generator=g) * 0.05
        b1 = torch.zeros(sq, device=DEV)
        w2t = torch.randn(sq, C, device=DEV, generator=g) * 0.05
        b2 = torch.zeros(C, device=DEV)
        for B in batches:
            pool = torch.randn(B, 2, C, device=DEV, generator=g)
            ms = timed(lambda: ops.se_gate(pool, hw, w1, b1, w2t, b2, torch.bfloat16), a.iters)
            nbytes = 4.0 * (B * 2 * C + 2 * sq * C) + 2.0 * B * C
            rows.append(dict(kernel="se_gate", shape=f"C{C} sq{sq}", batch=B, us=ms * 1e3, gbs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / PEAK))
    C, H, hid = 1792, 12, 112
    ca1 = torch.randn(hid, C, device=DEV, generator=g) * 0.02
    ca2t = torch.randn(hid, C, device=DEV, generator=g) * 0.02
    sa = torch.randn(98, device=DEV, generator=g) * 0.1
    lw = torch.ones(5, device=DEV)
    for B in batches:
        fmap = torch.randn(B, H, H, C, device=DEV, generator=g).bfloat16()
        lm = torch.rand(B, 5, 2, device=DEV, generator=g) * 224

        def run():
            heat = ops.landmark_heatmap(lm, lw, H, H)
            return ops.hybrid_attention(fmap, heat, ca1, ca2t, sa)
        ms = timed(run, a.iters)
        nbytes = 2.0 * 2 * B * H * H * C      # algorithmic: two reads of the map
        rows.append(dict(kernel="hybrid_attention", shape="C1792 12x12", batch=B, us=ms * 1e3, gbs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / PEAK))
    out = dict(what="isolated-kernel sweep, bf16, L2 flushed between iterations, median of %d" % a.iters, hbm_peak_gbs=PEAK, rows=rows)
    print(json.dumps(out))
    if a.md:
        with open(a.md, "w") as f:
            f.write("| kernel | shape | " + " | ".join(f"B={b}" for b in batches) + " |\n|---|---|" + "---|" * len(batches) + "\n")
            keys = []
            for r in rows:
                k = (r["kernel"], r["shape"])
                if k not in keys:
                    keys.append(k)
            for k in keys:
                cells = [next(r for r in rows if (r["kernel"], r["shape"]) == k and r["batch"] == b) for b in batches]
                f.write(f"| {k[0]} | {k[1]} | " + " | ".join(f"{c['us']:.0f} us, {c['gbs']:.0f} GB/s ({c['frac']:.2f})" for c in cells) + " |\n")


if __name__ == "__main__":
    main()
