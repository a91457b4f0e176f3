"""Serving helper: the eval forward of a DeepfakeDetectionModel captured ONCE into a CUDA graph and replayed per batch.

The forward is ~140 kernel launches of fixed shapes; a replay removes their launch gaps (14.10 -> 13.92 ms per batch of
256 on a B200, logits bit-identical) and, more importantly for a serving loop, takes the host out of the loop: one launch
per batch.  Inputs are copied into the graph's static buffers (or written there directly: `gi.images` / `gi.landmarks`
are the H2D targets of a serving loop); the returned tensors are the graph's static outputs and stay valid until the
next call.

The graph holds raw device pointers, so this object keeps alive everything it captured -- the model's inference
workspace for this shape (pinned: eager calls at other shapes no longer evict it) and the packed weight blob -- and
re-captures when the model's weights change (load_state_dict, an optimizer step, a train-mode forward).
"""
import torch


class GraphedInference:
    def __init__(self, model, images, landmarks=None, return_features=False):
        assert not model.training, "capture the eval forward (model.eval())"
        self.model, self.return_features = model, return_features
        self.images = images.detach().clone()
        self.landmarks = None if landmarks is None else landmarks.detach().clone()
        self._capture()

    def _capture(self):
        from . import ops
        model, dev = self.model, self.images.device
        if self.images.dtype == torch.uint8:
            B, H, W = self.images.shape[0], self.images.shape[1], self.images.shape[2]
        else:
            B, H, W = self.images.shape[0], self.images.shape[2], self.images.shape[3]
        self._ws_key = (ops.dtype_code(model.compute_dtype), B, H, W, str(dev))
        model._pinned_ws.add(self._ws_key)                     # _ws() never evicts a pinned shape
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():          # warm-up off the capture stream (allocator, weight packing)
            for _ in range(2):
                model(self.images, self.landmarks, return_features=self.return_features)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.outputs = model(self.images, self.landmarks, return_features=self.return_features)
        # what the captured kernels read and write: owned here for the lifetime of the graph
        self._version = model._version_key()
        self._dtype = model.compute_dtype
        self._keep = (model._workspace[self._ws_key], model._packed[(model.compute_dtype, str(dev))])
        self._blob = self._keep[1].blob

    def replay(self):
        """Replay on the static buffers (fill `self.images` / `self.landmarks` first)."""
        m = self.model
        if m._version_key() != self._version or m.compute_dtype != self._dtype or m.training:
            assert not m.training, "the captured forward is the eval forward"
            self._capture()                                     # weights changed: fold again and capture again
        self.graph.replay()
        return self.outputs

    def __call__(self, images, landmarks=None):
        if images.shape != self.images.shape or images.dtype != self.images.dtype:
            raise ValueError(f"graph captured for images {tuple(self.images.shape)} {self.images.dtype}, got {tuple(images.shape)} {images.dtype}")
        if (landmarks is None) != (self.landmarks is None):
            raise ValueError("graph captured with landmarks" if self.landmarks is not None else "graph captured without landmarks")
        self.images.copy_(images, non_blocking=True)
        if landmarks is not None:
            self.landmarks.copy_(landmarks, non_blocking=True)
        return self.replay()


class GraphedTrainStep:
    """One training step -- train-mode forward + CombinedLoss + backward (+ the bucketed gradient all-reduce under
    torch.distributed) -- captured ONCE into a CUDA graph and replayed per batch.

    The eager step is ~3500 kernel launches for ~30 ms of GPU work: the host needs ~28 ms to enqueue them, so any GPU-side
    saving is invisible until the launches come off the host.  A replay is one launch.  Dropout / drop-connect masks are
    functions of (seed, position); the captured kernels read the seed from a device word that is rewritten before every
    replay, so each step draws fresh masks although the graph's launch arguments are frozen.

        step = GraphedTrainStep(model, criterion, images, landmarks, labels)      # shapes fixed here
        for images, landmarks, labels in loader:
            losses = step(images, landmarks, labels)       # dict of static tensors: ce / focal / contrastive / total
            optimizer.step()                               # p.grad (views of one flat buffer) hold this step's gradients

    Gradients are OVERWRITTEN by every replay (the zeroing of the flat buffer is part of the graph): no accumulation across
    replays, and no zero_grad() needed.
    """

    def __init__(self, model, criterion, images, landmarks, targets, warmup: int = 3):
        assert model.training, "capture the training step (model.train())"
        self.model, self.criterion = model, criterion
        dev = images.device
        self.images = images.detach().clone()
        self.landmarks = None if landmarks is None else landmarks.detach().clone()
        self.targets = targets.detach().clone()
        model._seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._seed_host = torch.zeros(1, dtype=torch.int64).pin_memory()
        self._new_seed()
        self._params = [p for p in model.parameters()]
        # Warm-up and capture run on ONE side stream, and the parameters are not autograd inputs of the captured step: its
        # only differentiable leaf is an anchor created here, so no AccumulateGrad node of an earlier (default-stream)
        # iteration can be pulled into the capture.  The step attaches the gradient views to p.grad itself.
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            anchor = torch.zeros(1, device=dev, requires_grad=True)
            model._detached_anchor = anchor
            try:
                for _ in range(warmup):                         # allocator, arenas, NCCL communicators
                    self._step()
                side.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=side):
                    self.losses = self._step()
            finally:
                model._detached_anchor = None
        torch.cuda.current_stream(dev).wait_stream(side)
        self._grads = list(model._detached_grads)               # views of the captured flat gradient buffer
        self._flat = model._last_flat_grad
        for p, g in zip(self._params, self._grads):
            p.grad = g

    def _step(self):
        logits, feats = self.model(self.images, self.landmarks, return_features=True)
        losses = self.criterion(logits, self.targets, feats)
        losses["total"].backward()
        return {k: v.detach() for k, v in losses.items()}

    def _new_seed(self):
        from .model import _mix_seed
        self._seed_host[0] = _mix_seed(int(torch.randint(0, 2 ** 62, (1,)).item()))
        self.model._seed_dev.copy_(self._seed_host, non_blocking=True)

    def replay(self):
        """Replay on the static buffers (fill `self.images` / `self.landmarks` / `self.targets` first)."""
        m = self.model
        assert m.training, "the captured step is the training step"
        self._new_seed()
        self.graph.replay()
        for p, g in zip(self._params, self._grads):             # survive a zero_grad(set_to_none=True) between steps
            p.grad = g
        m._last_flat_grad = self._flat
        m.invalidate_packed()                                   # running statistics moved
        return self.losses

    def __call__(self, images, landmarks, targets):
        if images.shape != self.images.shape or images.dtype != self.images.dtype:
            raise ValueError(f"step captured for images {tuple(self.images.shape)} {self.images.dtype}, got {tuple(images.shape)} {images.dtype}")
        self.images.copy_(images, non_blocking=True)
        if self.landmarks is not None:
            self.landmarks.copy_(landmarks, non_blocking=True)
        self.targets.copy_(targets, non_blocking=True)
        return self.replay()
