"""Experiment: one batch-256 bf16 forward replayed (a) as ONE stream of 256 images (GraphedInference), (b) as S concurrent
streams of 256/S images each inside one CUDA graph (fork / join by events).  The forward is ~180 launches whose ramp-up,
tail and the latency-bound SE / attention / head chains leave SMs idle; a second stream's kernels can fill those holes.
Each chunk runs on its own model copy (own workspace + packed blob), so nothing is shared but the read-only input.
Prints ms per 256 images for each S.  argv: steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import deepfake_vit_b200 as d

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = 256
torch.manual_seed(42)
m = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG).cuda().eval().set_compute_dtype(torch.bfloat16)
x = torch.randn(B, 3, 380, 380, device="cuda")
lm = torch.rand(B, 5, 2, device="cuda") * 380


def timed(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


gi = d.GraphedInference(m, x, lm)
ref = gi.replay()[0].clone()
print(f"S=1 (GraphedInference): {timed(gi.replay, steps):.3f} ms")

for S in (2, 3, 4):
    sizes = [B // S + (1 if i < B % S else 0) for i in range(S)]
    sizes = [s + (s & 1) for s in sizes[:-1]]
    sizes.append(B - sum(sizes))
    offs = [sum(sizes[:i]) for i in range(S)]
    models = []
    for _ in range(S):
        mm = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG)
        mm.load_state_dict(m.state_dict())
        models.append(mm.cuda().eval().set_compute_dtype(torch.bfloat16))
    xs = [x[o:o + s].contiguous() for o, s in zip(offs, sizes)]
    ls = [lm[o:o + s].contiguous() for o, s in zip(offs, sizes)]
    streams = [torch.cuda.Stream() for _ in range(S)]
    with torch.no_grad():
        for mm, xx, ll in zip(models, xs, ls):
            for _ in range(2):
                mm(xx, ll)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    outs = [None] * S
    with torch.cuda.graph(g), torch.no_grad():
        cur = torch.cuda.current_stream()
        for i, st in enumerate(streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs[i] = models[i](xs[i], ls[i])[0]
        for st in streams:
            cur.wait_stream(st)
    t = timed(g.replay, steps)
    got = torch.cat(outs)
    err = ((got - ref).norm() / ref.norm()).item()
    print(f"S={S} sizes {sizes}: {t:.3f} ms   (logits vs S=1: rel {err:.2e}; the heat-map maximum is per chunk here)")
    del g, models
