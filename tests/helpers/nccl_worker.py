"""Worker of tests/test_gpu_nccl.py (one process per GPU, launched by torch.distributed.run).

Checks, on real NCCL over NVLink, that the data-parallel training step of DeepfakeDetectionModel leaves on EVERY rank
the mean of the ranks' single-GPU gradients: the flat gradient after the bucketed, backward-overlapped all-reduce equals
(bitwise, fp32) the average of the per-rank flat gradients computed with the collective switched off; BatchNorm buffers
stay rank-local (DDP semantics over the single-GPU reference); ranks draw different dropout masks.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("DFV_WORKER_DEADLINE_S", "240")), exit=True)   # a hang becomes a stack dump + exit
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import deepfake_vit_b200 as d
    from deepfake_vit_b200.parallel import broadcast_parameters, shard_bounds

    torch.manual_seed(100 + rank)                       # ranks start DIFFERENT on purpose: broadcast must fix that
    m = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG).to(dev).train().set_compute_dtype(torch.float32)
    broadcast_parameters(m)
    for mod in m.modules():                             # exact comparison: stochastic parts off
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.feature_extractor.backbone.backbone.drop_connect_rate = 0.0
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    crit = d.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, torch.tensor([1.0, 1.5], device=dev))

    B, size = 8 * world, 96
    g = torch.Generator().manual_seed(7)                # the same global batch on every rank, sharded below
    X = torch.randn(B, 3, size, size, generator=g)
    LM = torch.rand(B, 5, 2, generator=g) * size
    Y = torch.randint(0, 2, (B,), generator=g)
    lo, hi = shard_bounds(B, world, rank, 2)
    x, lm, y = X[lo:hi].to(dev), LM[lo:hi].to(dev), Y[lo:hi].to(dev)

    def step(allreduce, bucket_floats):
        m.load_state_dict(sd0)
        m.zero_grad(set_to_none=True)
        m.ddp_allreduce, m.ddp_bucket_floats = allreduce, bucket_floats
        logits, feats = m(x, lm, return_features=True)
        crit(logits, y, feats)["total"].backward()
        torch.cuda.synchronize()
        return m._last_flat_grad.clone()

    local_flat = step(False, 4 << 20)
    gathered = [torch.empty_like(local_flat) for _ in range(world)]
    dist.all_gather(gathered, local_flat)
    want = gathered[0].clone()
    for t in gathered[1:]:
        want += t
    want /= world
    assert not torch.equal(gathered[0], gathered[-1]), "ranks computed identical gradients: the batch was not sharded"
    for bucket_floats in (4 << 20, 1 << 18, 1 << 30):   # ~4 buckets, one bucket per unit, a single bucket
        # (1) the whole step: backward + bucketed all-reduce behind the events dfv_train_bwd records.  The weight-gradient
        # kernels reduce with fp32 atomics, so two runs of the same step agree to rounding (~1e-6), not bit for bit.
        got = step(True, bucket_floats)
        nb = len(m._reducer.buckets)
        err = ((got - want).double().norm() / want.double().norm()).item()
        assert err < 2e-5, (rank, nb, err)
        # (2) the collective itself on a FIXED buffer: bitwise the mean for 2 ranks ((a + b) / 2 is exact however NCCL averages)
        buf = local_flat.clone()
        for ev in m._reducer.events.values():
            ev.record()
        m._reducer.reduce(buf)
        torch.cuda.synchronize()
        if world == 2:
            assert torch.equal(buf, want), f"rank {rank}: bucketed all-reduce ({nb} buckets) != mean of the ranks' buffers"
        else:
            assert ((buf - want).double().norm() / want.double().norm()).item() < 1e-6
        if rank == 0:
            print(f"nccl_worker: world {world}, {nb} buckets: step gradient rel err {err:.1e}; all-reduce of a fixed buffer == mean, bitwise", flush=True)
    # parameter .grad views are the reduced values
    p = m.feature_extractor.backbone.backbone._conv_head.weight
    names = [n for n, _ in m.named_parameters()]
    params, starts, _ = m._flat_layout()
    off = starts[names.index("feature_extractor.backbone.backbone._conv_head.weight")]
    assert torch.equal(p.grad.flatten(), m._last_flat_grad[off:off + p.numel()])
    assert float((p.grad.flatten() - want[off:off + p.numel()]).norm()) < 2e-5 * float(want[off:off + p.numel()].norm())
    # the captured step (GraphedTrainStep) records the NCCL collectives and the cross-stream event edges into the graph
    m.load_state_dict(sd0)
    m.ddp_allreduce, m.ddp_bucket_floats = True, 4 << 20
    gstep = d.GraphedTrainStep(m, crit, x, lm, y)
    m.load_state_dict(sd0)
    gstep.replay()
    torch.cuda.synchronize()
    gerr = ((m._last_flat_grad - want).double().norm() / want.double().norm()).item()
    assert gerr < 2e-5, (rank, gerr)
    if rank == 0:
        print(f"nccl_worker: graph-replayed step over NCCL: gradient rel err {gerr:.1e}", flush=True)
    m.load_state_dict(sd0)
    m.zero_grad(set_to_none=True)
    step(True, 4 << 20)
    # BatchNorm buffers are rank-local
    rm = m.feature_extractor.backbone.backbone._bn0.running_mean.clone()
    all_rm = [torch.empty_like(rm) for _ in range(world)]
    dist.all_gather(all_rm, rm)
    assert not torch.equal(all_rm[0], all_rm[-1]), "BatchNorm running statistics must stay rank-local"
    # stochastic parts: ranks draw different masks from the same torch seed
    m.load_state_dict(sd0)
    m.feature_extractor.backbone.dropout.p = 0.4
    torch.manual_seed(5)
    with torch.no_grad():
        _, f = m(X[:8].to(dev), LM[:8].to(dev), return_features=True)
    masks = [torch.empty_like(f) for _ in range(world)]
    dist.all_gather(masks, (f == 0).float())
    assert not torch.equal(masks[0], masks[-1]), "every rank drew the same dropout mask"
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("nccl_worker: OK", flush=True)
    # Tear-down: the captured graph holds NCCL kernels of this communicator -- drop it first.  destroy_process_group()
    # blocked forever with the graph alive (seen on the 2-GPU box), so it also gets a bounded wait; everything asserted
    # above has been checked and printed by now.
    del gstep
    import gc
    import threading
    gc.collect()
    torch.cuda.synchronize()
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(30)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
