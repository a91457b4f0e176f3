"""N > 1 host logic on CPU: world_size-2 gloo process group.  Covers the batch / clip
sharding rules and the flat-gradient all-reduce the training backward runs (no CUDA involved)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deepfake_vit_b200 import parallel
        # 1. flat gradient all-reduce = mean over ranks
        torch.manual_seed(100 + rank)
        flat = torch.randn(18_939_345 // 64)
        mine = flat.clone()
        parallel.allreduce_gradients(flat)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        want = torch.stack(gathered).mean(0)
        ok_mean = bool(torch.allclose(flat, want, atol=1e-6))
        # 2. sharding: contiguous, even, covers the batch exactly once
        bounds = [parallel.shard_bounds(130, world, r) for r in range(world)]
        ok_shard = bounds[0][0] == 0 and bounds[-1][1] == 130 and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
        ok_even = all((e - s) % 2 == 0 for s, e in bounds[:-1])
        # 3. parameters broadcast from rank 0
        lin = torch.nn.Linear(8, 4)
        bn = torch.nn.BatchNorm1d(4)
        with torch.no_grad():
            bn.running_mean.fill_(float(rank))
        mod = torch.nn.Sequential(lin, bn)
        parallel.broadcast_parameters(mod)
        w = [torch.empty_like(lin.weight) for _ in range(world)]
        dist.all_gather(w, lin.weight.data)
        ok_bcast = all(torch.equal(w[0], t) for t in w) and float(bn.running_mean[0]) == 0.0
        # 4. a data-parallel step of a toy loss: averaged per-rank gradients == gradient of the mean of rank losses
        torch.manual_seed(7)
        wgt = torch.randn(6, requires_grad=True)
        x = torch.arange(24, dtype=torch.float32).view(4, 6) / 10.0
        s, e = parallel.shard_bounds(4, world, rank)
        (x[s:e] @ wgt).pow(2).mean().backward()
        g = wgt.grad.clone()
        parallel.allreduce_gradients(g)
        wg2 = wgt.detach().clone().requires_grad_(True)
        sum((x[a:b] @ wg2).pow(2).mean() for a, b in [parallel.shard_bounds(4, world, r) for r in range(world)]).div(world).backward()
        ok_ddp = bool(torch.allclose(g, wg2.grad, atol=1e-6))
        # 5. the optional `global` modes' exchanges (SURVEY 8(e) caveats 1 and 3): the heat-map maximum travels as an
        #    order-preserving key (uint32 order == float order, int32 storage), the weighted-CE normaliser as a mean
        import struct

        def key_of(v):      # the library's key: non-negative floats get the sign bit set, negative ones are inverted
            u = struct.unpack("<I", struct.pack("<f", v))[0]
            u = (~u & 0xFFFFFFFF) if (u & 0x80000000) else (u | 0x80000000)
            return u - (1 << 32) if u >= (1 << 31) else u
        vals = [[0.75, -0.5], [1.25, -0.25]]        # rank 1 holds the larger maximum in both cases
        ok_key = True
        for j in range(2):
            got = parallel.allreduce_max_key(torch.tensor([key_of(vals[rank][j])], dtype=torch.int32))
            ok_key = ok_key and int(got[0]) == key_of(max(vals[0][j], vals[1][j]))
        norm = parallel.allreduce_mean_(torch.tensor([10.0 + 4.0 * rank]))
        ok_norm = abs(float(norm[0]) - 12.0) < 1e-6
        out[rank] = (ok_mean, ok_shard, ok_even, ok_bcast, ok_ddp, ok_key, ok_norm)
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    for r, flags in res.items():
        assert all(flags), (r, flags)


@pytest.mark.parametrize("n,world", [(256, 8), (64, 8), (130, 4), (6, 4), (2, 2), (0, 2)])
def test_shard_bounds_cover(n, world):
    from deepfake_vit_b200.parallel import shard_bounds
    b = [shard_bounds(n, world, r) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == n
    assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    assert all((e - s) % 2 == 0 for s, e in b[:-1])


def test_clip_shards_whole_clips():
    from deepfake_vit_b200.parallel import clip_shards
    spans = [clip_shards(64, 32, 8, r) for r in range(8)]       # BASELINE.json configs[3]: 64 clips x 32 frames on 8 GPUs
    assert spans[0] == (0, 256) and spans[-1] == (1792, 2048)
    assert all((e - s) % 32 == 0 for s, e in spans)


def test_flat_layout_matches_registration_order():
    from deepfake_vit_b200.parallel import flat_layout
    offs, total = flat_layout([("a", torch.Size([3, 4])), ("b", torch.Size([5])), ("c", torch.Size([]))])
    # registration order, every tensor on a 64-float (256-byte) boundary: the kernels read parameters with vector loads
    assert offs == {"a": (0, 12), "b": (64, 5), "c": (128, 1)} and total == 192


def test_flat_offsets_shared_by_gradient_buffer_and_optimizer():
    from deepfake_vit_b200.parallel import FLAT_ALIGN, flat_offsets
    offs, total = flat_offsets([1296, 48, 7, 64, 65])
    assert offs == [0, 1344, 1408, 1472, 1536] and total == 1664
    assert all(o % FLAT_ALIGN == 0 for o in offs)


def test_gradient_buckets_tile_the_flat_buffer_in_completion_order():
    """The backward's gradient units (head+attention+classifier, block 31..0, stem) merged into all-reduce buckets:
    every bucket is one contiguous flat range, buckets follow completion order and tile [0, total) exactly."""
    import deepfake_vit_b200 as d
    from deepfake_vit_b200.parallel import merge_units
    m = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG)
    params, starts, total = m._flat_layout()
    units = m._unit_ranges(params, starts, total)
    assert len(units) == d._lib.GRAD_UNITS and units[0][1] == total and units[-1][0] == 0
    names = [n for n, _ in m.named_parameters()]
    lo31 = starts[names.index("feature_extractor.backbone.backbone._blocks.31._expand_conv.weight")]
    assert units[1][0] == lo31 and units[0][0] == starts[names.index("feature_extractor.backbone.backbone._conv_head.weight")]
    for bucket_floats in (1 << 18, 4 << 20, 1 << 30):
        buckets = merge_units(units, bucket_floats)
        assert buckets[0][1] == total and buckets[-1][0] == 0 and buckets[-1][2] == len(units) - 1
        assert all(buckets[i][0] == buckets[i + 1][1] for i in range(len(buckets) - 1))
        assert all(hi - lo >= bucket_floats for lo, hi, _ in buckets[:-1])
    assert len(merge_units(units, 1 << 30)) == 1 and 3 <= len(merge_units(units, 4 << 20)) <= 8
