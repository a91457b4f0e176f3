"""Diagnostic: per-parameter gradient agreement of the bf16 training step with the fp32 oracle, next to the
oracle's own autocast-bf16 gradients (the yardstick)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import deepfake_vit_b200 as d
from oracle import calibrate, refmodel
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from test_gpu_train_model import _pair, LOSS_W

size, B = int(sys.argv[1]) if len(sys.argv) > 1 else 96, int(sys.argv[2]) if len(sys.argv) > 2 else 4
om, m, d, refmodel = _pair(size)
x, lm, y = calibrate.synthetic_batch(B, size)
cw = torch.tensor([1.0, 1.5])
sd0 = {k: v.clone() for k, v in om.state_dict().items()}

def oracle_grads(autocast):
    om.load_state_dict(sd0); om.zero_grad(set_to_none=True)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        lo, fe = om(x, lm, return_features=True)
    loss = refmodel.CombinedLoss(LOSS_W, cw)(lo.float(), y, fe.float())["total"]
    loss.backward()
    return {n: p.grad.clone().double() for n, p in om.named_parameters()}, loss.item()

g32, l32 = oracle_grads(False)
gac, lac = oracle_grads(True)

def ours(dtype):
    m.load_state_dict(sd0); m.zero_grad(set_to_none=True); m.set_compute_dtype(dtype)
    lo, fe = m(x.cuda(), lm.cuda(), return_features=True)
    loss = d.CombinedLoss(LOSS_W, cw.cuda())(lo, y.cuda(), fe)["total"]
    loss.backward()
    return {n: p.grad.detach().cpu().double() for n, p in m.named_parameters()}, loss.item()

o32, lo32 = ours(torch.float32)
o16, lo16 = ours(torch.bfloat16)
print("loss fp32-oracle %.6f autocast-oracle %.6f ours-fp32 %.6f ours-bf16 %.6f" % (l32, lac, lo32, lo16))

def cos(a, b):
    return float((a.flatten() @ b.flatten()) / (a.norm() * b.norm() + 1e-30))

def flat(g):
    return torch.cat([g[n].flatten() for n in g32])

print("flat cosine vs fp32 oracle: ours-fp32 %.4f  ours-bf16 %.4f  oracle-autocast %.4f" % (cos(flat(o32), flat(g32)), cos(flat(o16), flat(g32)), cos(flat(gac), flat(g32))))
print("norms: fp32 %.4f ours-bf16 %.4f autocast %.4f" % (flat(g32).norm(), flat(o16).norm(), flat(gac).norm()))
print("%-70s %10s %9s %9s %9s" % ("parameter", "|g32|", "cos ours16", "cos acast", "ratio16"))
for n in g32:
    if n.endswith("weight") and ("conv" in n or "classifier" in n or "fc" in n) or "attention_weights" in n:
        print("%-70s %10.3e %9.4f %9.4f %9.3f" % (n[-70:], g32[n].norm(), cos(o16[n], g32[n]), cos(gac[n], g32[n]), o16[n].norm() / (g32[n].norm() + 1e-30)))
