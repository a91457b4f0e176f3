"""Optimizer step either side of the hot path: the reference's
`clip_grad_norm_(model.parameters(), 1.0)` + `AdamW(lr, weight_decay)` (src/training/trainer.py:158-167,
scripts/train.py:96-102) as ONE pass over flat fp32 buffers (libdfvit dfv_clip_adamw_step).

The class is a torch.optim.Optimizer, so `param_groups` (LR schedulers such as
CosineAnnealingWarmRestarts, scripts/train.py:49-55) and `state_dict()` / `load_state_dict()` keep the
stock AdamW format: a checkpoint written by the reference trainer (`optimizer_state_dict`,
trainer.py:299-306) loads here and vice versa.
"""
from typing import Iterable, Optional

import torch

from ._lib import check, lib
from .parallel import flat_offsets


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: Optional[float] = None, grad_source=None):
        """max_grad_norm: global-norm clip applied inside the step (None / 0 = no clipping).
        grad_source: an object with `_last_flat_grad` (DeepfakeDetectionModel): its backward already leaves the
        gradients in one flat buffer in parameter order, which the step then reads without gathering."""
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        assert len(self.param_groups) == 1, "FusedAdamW keeps one parameter group (as the reference trainer does)"
        self.max_grad_norm = float(max_grad_norm or 0.0)
        self._grad_source = grad_source
        self._params = [p for p in self.param_groups[0]["params"]]
        assert all(p.dtype == torch.float32 for p in self._params), "master parameters are fp32"
        dev = self._params[0].device
        self._offsets, n = flat_offsets([p.numel() for p in self._params])   # same layout as the model's flat gradient buffer
        self._n = n
        # parameters become views of one flat buffer (values preserved)
        self._flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self._flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        self._flat_v = torch.zeros(n, dtype=torch.float32, device=dev)
        self._norm_ws = torch.zeros(1, dtype=torch.float64, device=dev)
        self._total_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        with torch.no_grad():
            self._flat_p.zero_()
            for p, off in zip(self._params, self._offsets):
                k = p.numel()
                self._flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self._flat_p[off:off + k].view_as(p)
        self._step = 0
        self._frozen_key, self._frozen_chunks = None, None
        self._bind_state()
        if grad_source is not None and hasattr(grad_source, "invalidate_packed"):
            grad_source.invalidate_packed()       # the parameters moved into the flat buffer

    def _frozen_mask(self, grads_present):
        """One byte per 64-element chunk of the flat buffers: 1 = the optimizer must not touch it.  torch.optim.AdamW
        skips parameters whose .grad is None (requires_grad False, freeze_bn): no update, no weight decay."""
        key = tuple(grads_present)
        if all(key):
            return None
        if self._frozen_key != key:
            mask = torch.zeros(self._n // 64, dtype=torch.uint8)
            for p, off, ok in zip(self._params, self._offsets, key):
                if not ok:
                    mask[off // 64:(off + p.numel() + 63) // 64] = 1
            self._frozen_key, self._frozen_chunks = key, mask.to(self._flat_p.device)
        return self._frozen_chunks

    def _bind_state(self):
        for p, off in zip(self._params, self._offsets):
            k = p.numel()
            self.state[p] = {"step": torch.tensor(float(self._step)),
                             "exp_avg": self._flat_m[off:off + k].view_as(p),
                             "exp_avg_sq": self._flat_v[off:off + k].view_as(p)}

    def _flat_grad(self) -> torch.Tensor:
        src = getattr(self._grad_source, "_last_flat_grad", None) if self._grad_source is not None else None
        if src is not None and src.numel() == self._n:
            g0 = self._params[0].grad
            if g0 is not None and g0.data_ptr() == src.data_ptr():
                return src
        flat = torch.zeros(self._n, dtype=torch.float32, device=self._flat_p.device)
        for p, off in zip(self._params, self._offsets):
            if p.grad is not None:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
        return flat

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """One clip + AdamW step.  Returns the pre-clip global gradient norm (a 1-element device tensor: no
        host synchronisation), the value `clip_grad_norm_` returns in the reference loop."""
        loss = closure() if closure is not None else None
        lo, hi = self._flat_p.data_ptr(), self._flat_p.data_ptr() + 4 * self._n
        for p in self._params:
            assert lo <= p.data_ptr() < hi, ("a parameter is no longer a view of FusedAdamW's flat buffer (model.to(...) / "
                                             ".float() after the optimizer was built): rebuild the optimizer")
        present = [p.grad is not None and p.requires_grad for p in self._params]
        g = self._grad_flat_checked()
        frozen = self._frozen_mask(present)
        grp = self.param_groups[0]
        self._step += 1
        b1, b2 = grp["betas"]
        check(lib.dfv_clip_adamw_step(self._flat_p.data_ptr(), g.data_ptr(), self._flat_m.data_ptr(), self._flat_v.data_ptr(),
                                      self._n, self._norm_ws.data_ptr(), self.max_grad_norm, float(grad_scale), float(grp["lr"]),
                                      float(b1), float(b2), float(grp["eps"]), float(grp["weight_decay"]), self._step,
                                      self._total_norm.data_ptr(), frozen.data_ptr() if frozen is not None else None,
                                      torch.cuda.current_stream().cuda_stream))
        for p in self._params:
            self.state[p]["step"].fill_(float(self._step))
        if self._grad_source is not None and hasattr(self._grad_source, "invalidate_packed"):
            self._grad_source.invalidate_packed()     # weights were rewritten through raw pointers
        return self._total_norm if loss is None else loss

    def _grad_flat_checked(self):
        g = self._flat_grad()
        assert g.is_cuda and g.is_contiguous() and g.dtype == torch.float32
        return g

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # the stock loader replaces the state tensors: copy them back into the flat buffers and re-bind the views
        step = 0
        with torch.no_grad():
            for p, off in zip(self._params, self._offsets):
                st = self.state.get(p, {})
                k = p.numel()
                if "exp_avg" in st:
                    self._flat_m[off:off + k].copy_(st["exp_avg"].reshape(-1))
                    self._flat_v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                    step = max(step, int(float(st["step"])))
        self._step = step
        self._bind_state()
