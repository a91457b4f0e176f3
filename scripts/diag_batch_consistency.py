"""Which stage / which images differ between one batch-N forward and the same images run as two half batches?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepfake_vit_b200 as d
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
size = int(sys.argv[2]) if len(sys.argv) > 2 else 380
torch.manual_seed(42)
m = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG).cuda().eval().set_compute_dtype(torch.bfloat16)
# non-degenerate BN statistics so that the signal survives 32 blocks
for mod in m.modules():
    if isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
        mod.running_var.uniform_(0.02, 0.2); mod.running_mean.normal_(0, 0.1)
g = torch.Generator().manual_seed(13)
x = torch.randn(N, 3, size, size, generator=g).cuda()
lo_f, fe_f, _, taps_f = m.forward_with_taps(x, None)
taps_f = [t.float().cpu() for t in taps_f]
h = N // 2
outs = [m.forward_with_taps(x[i * h:(i + 1) * h].contiguous(), None) for i in range(2)]
for s in range(len(taps_f)):
    th = torch.cat([outs[0][3][s].float().cpu(), outs[1][3][s].float().cpu()])
    tf = taps_f[s]
    per = ((tf - th).flatten(1).norm(dim=1) / (th.flatten(1).norm(dim=1) + 1e-20))
    bad = (per > 1e-3).nonzero().flatten().tolist()
    name = "stem" if s == 0 else ("head" if s == len(taps_f) - 1 else f"block{s-1}")
    print(f"{name:8s} max rel {per.max().item():.2e}  differing images: {len(bad)} {bad[:24]}")
    if len(bad) and s > 3: break
