"""CombinedLoss drop-in (src/training/losses.py:164-247) computed by one libdfvit kernel.

`CombinedLoss(weights, class_weights)(logits, targets, features=None) -> dict` with keys
`ce`, `focal`, `contrastive` (only when features are given and B >= 2) and `total`, exactly as
the reference returns them.  The kernel produces the loss values and d total / d logits,
d total / d features in the same pass; autograd is wired through a custom Function so
`losses['total'].backward()` works as in `Trainer.train_epoch` (trainer.py:144-153).
"""
from typing import Optional

import torch
import torch.nn as nn

from . import ops


class _CombinedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, features, targets, class_weights, w_ce, w_focal, w_con, ce_norm=None):
        lo = logits.detach().float().contiguous()
        fe = features.detach().float().contiguous() if features is not None else None
        losses, has_con, dlogits, dfeat = ops.combined_loss(lo, targets.contiguous(), fe, class_weights, w_ce, w_focal,
                                                            w_con, want_grad=True, ce_norm=ce_norm)
        ctx.save_for_backward(dlogits, dfeat if dfeat is not None else torch.empty(0, device=lo.device))
        ctx.has_feat = features is not None
        ctx.in_dtypes = (logits.dtype, features.dtype if features is not None else None)
        ctx.mark_non_differentiable()
        ctx.has_con = has_con
        return losses

    @staticmethod
    def backward(ctx, g):
        dlogits, dfeat = ctx.saved_tensors
        # only `total` (index 3) carries the kernel's gradient; the components are reported values
        gt = g[3]
        gl = (dlogits * gt).to(ctx.in_dtypes[0])
        gf = (dfeat * gt).to(ctx.in_dtypes[1]) if ctx.has_feat else None
        return gl, gf, None, None, None, None, None, None


class CombinedLoss(nn.Module):
    """ce_scope (not part of the reference API): "rank" -- the weighted cross-entropy is normalised by THIS rank's
    sum of w[y_i], i.e. the reference loss on the shard, averaged over ranks by the gradient all-reduce (DDP semantics);
    "global" (torch.distributed, class weights given) -- by the mean over ranks of that sum (one 4-byte all-reduce), so
    that the mean over ranks of `ce` / `total` and of their gradients is exactly the loss of the global batch on one GPU
    (SURVEY.md 8(e) caveat 3).  Focal and contrastive terms are plain means and need nothing when shards are equal."""

    def __init__(self, weights: dict, class_weights: Optional[torch.Tensor] = None, ce_scope: str = "rank", group=None):
        super().__init__()
        assert ce_scope in ("rank", "global")
        self.weights = weights
        self.class_weights = class_weights
        self.ce_scope, self.group = ce_scope, group

    def forward(self, logits, targets, features=None) -> dict:
        if not logits.is_cuda:
            raise RuntimeError("deepfake_vit_b200.CombinedLoss runs on sm_100 CUDA devices only (no CPU path)")
        w = self.weights
        w_ce = float(w["ce"]) if "ce" in w and w["ce"] > 0 else 0.0
        w_focal = float(w["focal"]) if "focal" in w and w["focal"] > 0 else 0.0
        w_con = float(w["contrastive"]) if "contrastive" in w and w["contrastive"] > 0 else 0.0
        cw = None
        if self.class_weights is not None:
            cw = self.class_weights.detach().to(device=logits.device, dtype=torch.float32).contiguous()
        feats = features if w_con > 0 else None
        ce_norm = None
        if self.ce_scope == "global" and cw is not None and w_ce > 0:
            from . import parallel
            ce_norm = parallel.allreduce_mean_(ops.class_weight_sum(targets.contiguous(), cw, logits.shape[1]), self.group)
        vec = _CombinedLossFn.apply(logits, feats, targets, cw, w_ce, w_focal, w_con, ce_norm)
        out = {}
        if w_ce > 0:
            out["ce"] = vec[0]
        if w_focal > 0:
            out["focal"] = vec[1]
        if feats is not None and features.size(0) >= 2:
            out["contrastive"] = vec[2]
        out["total"] = vec[3]
        return out
