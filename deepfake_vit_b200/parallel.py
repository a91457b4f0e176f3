"""Data-parallel host logic: one process per GPU, the batch sharded over images.

Inference needs no collective.  Training exchanges gradients with ONE all-reduce over the flat fp32
gradient buffer that dfv_train_bwd fills (BatchNorm statistics stay rank-local, as DDP over the
single-GPU reference would leave them).  Nothing here computes on tensors beyond the collective itself.
"""
from typing import List, Tuple

import torch


def shard_bounds(n_items: int, world: int, rank: int, multiple: int = 2) -> Tuple[int, int]:
    """Contiguous shard [start, end) of n_items for `rank`.  Shard sizes are multiples of `multiple`
    (CombinedLoss pairs consecutive samples (2i, 2i+1), losses.py:233-238, so a pair never straddles two
    ranks); the remainder goes to the last rank."""
    assert 0 <= rank < world and n_items >= 0 and multiple >= 1
    units = n_items // multiple
    base, extra = divmod(units, world)
    start = (rank * base + min(rank, extra)) * multiple
    end = start + (base + (1 if rank < extra else 0)) * multiple
    if rank == world - 1:
        end = n_items
    return start, end


def clip_shards(n_clips: int, frames_per_clip: int, world: int, rank: int) -> Tuple[int, int]:
    """Video scoring (BASELINE.json configs[3]): whole clips per rank, so the per-clip heat-map normaliser
    group and the per-clip mean never cross ranks.  Returns the image range [start, end)."""
    c0, c1 = shard_bounds(n_clips, world, rank, 1)
    return c0 * frames_per_clip, c1 * frames_per_clip


def allreduce_gradients(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Average the flat gradient buffer over ranks in place.  NCCL averages inside the collective; gloo
    (the CPU tests) sums, then divides."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return flat
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    return flat


def allreduce_mean_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over ranks of a small tensor (the weighted-CE normaliser of the exact data-parallel loss)."""
    return allreduce_gradients(t, group)


def allreduce_max_key(key: torch.Tensor, group=None) -> torch.Tensor:
    """Element-wise maximum over ranks of order-preserving heat-map keys (ops.landmark_max_key): int32 storage of uint32
    keys whose UNSIGNED order is the float order.  Collectives compare int32 as signed, so the sign bit is flipped for
    the exchange (unsigned order -> signed order) and flipped back.  Returns a new tensor; one 4-byte all-reduce(MAX)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return key
    k = key ^ (-2 ** 31)
    dist.all_reduce(k, op=dist.ReduceOp.MAX, group=group)
    return k ^ (-2 ** 31)


FLAT_ALIGN = 64   # floats: every tensor starts on a 256-byte boundary inside a flat buffer, like a torch allocation
                  # (the kernels read parameters with 16-byte vector loads)


def flat_offsets(numels, align: int = FLAT_ALIGN):
    """Start offset of every tensor inside a flat buffer (given order, each start rounded up to `align` elements)
    and the buffer's total length.  Shared by the model's flat gradient buffer and FusedAdamW's flat parameter /
    moment buffers, so the optimizer reads the gradient buffer the backward pass filled without gathering."""
    offs, total = [], 0
    for n in numels:
        offs.append(total)
        total += (int(n) + align - 1) // align * align
    return offs, total


def flat_layout(named_shapes: List[Tuple[str, torch.Size]]):
    """Offsets of every parameter gradient inside the flat buffer (registration order, FLAT_ALIGN-aligned starts):
    {name: (offset, numel)}, total."""
    numels = []
    for _, shape in named_shapes:
        n = 1
        for s in shape:
            n *= int(s)
        numels.append(n)
    offs, total = flat_offsets(numels)
    return {name: (o, n) for (name, _), o, n in zip(named_shapes, offs, numels)}, total


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None):
    """Make every rank start from rank `src`'s parameters and buffers (what DDP does at construction)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():       # in-place on the tensors themselves: version counters see it
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=group)
    if hasattr(module, "invalidate_packed"):
        module.invalidate_packed()


def merge_units(unit_ranges, bucket_floats: int):
    """Gradient buckets from the backward pass's completion units.  `unit_ranges[u] = (lo, hi)` is the flat-buffer range
    whose gradients are complete after unit u (completion order = reverse registration order, so consecutive units are
    adjacent ranges going DOWN the buffer).  Consecutive units are merged until a bucket holds >= bucket_floats elements.
    Returns [(lo, hi, last_unit)] in completion order; the buckets tile [0, total) exactly."""
    buckets, lo, hi = [], None, None
    for u, (a, b) in enumerate(unit_ranges):
        if lo is None:
            lo, hi = a, b
        else:
            assert b == lo, "units must be adjacent, descending"
            lo = a
        if hi - lo >= bucket_floats or u == len(unit_ranges) - 1:
            buckets.append((lo, hi, u))
            lo = hi = None
    return buckets


class BucketedAllReduce:
    """Gradient all-reduce overlapped with the backward pass: dfv_train_bwd records one CUDA event per
    gradient unit on the compute stream; each bucket's NCCL all-reduce is issued on a side stream behind the event of
    its last unit, so it runs while the remaining (earlier-layer) backward kernels execute.  The compute stream joins
    the side stream once, after the last bucket."""

    def __init__(self, unit_ranges, total: int, bucket_floats: int, device):
        import ctypes as C
        self.total, self.bucket_floats = total, bucket_floats
        self.buckets = merge_units(unit_ranges, bucket_floats)
        self.stream = torch.cuda.Stream(device=device)
        self.events = {}
        table = [None] * len(unit_ranges)
        cur = torch.cuda.current_stream(device)
        for _, _, unit in self.buckets:
            ev = torch.cuda.Event()
            ev.record(cur)                      # torch creates the CUDA event lazily, on its first record
            self.events[unit] = ev
            table[unit] = ev.cuda_event
        self._table = (C.c_void_p * len(table))(*table)

    @staticmethod
    def active(group=None) -> bool:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def event_table(self):
        return self._table

    def reduce(self, flat: torch.Tensor, group=None):
        cur = torch.cuda.current_stream(flat.device)
        with torch.cuda.stream(self.stream):
            for lo, hi, unit in self.buckets:
                self.stream.wait_event(self.events[unit])
                allreduce_gradients(flat[lo:hi], group)
        if not torch.cuda.is_current_stream_capturing():
            flat.record_stream(self.stream)
        cur.wait_stream(self.stream)
