"""BASELINE.json configs[3]: video-frame scoring, 64 clips x 32 frames (2048 frames @ 380 x 380) sharded over the GPUs of one
box with parallel.clip_shards (whole clips per rank: the heat-map normaliser group and the per-clip mean never cross ranks;
no collective on the data path), per-clip mean-logit aggregation + the notebook's rule mean(softmax[:, 1]) >= 0.5
(task.ipynb:434-442).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/config4_video_scoring.py

Every rank (1) checks ITS clips in fp32 mode against one oracle call per clip (the reference's one-call-per-file, run on the
GPU in fp32 with TF32 off), (2) times the bf16 scoring of its shard: device-resident and end to end from pinned uint8 crops.
Rank 0 prints one JSON line (max over ranks of the timings).
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import deepfake_vit_b200 as d
    from deepfake_vit_b200.parallel import clip_shards
    from oracle import calibrate, refmodel
    n_clips, frames, size = int(os.environ.get("DFV_CLIPS", 64)), 32, 380
    lo, hi = clip_shards(n_clips, frames, world, rank)
    my_clips = (hi - lo) // frames
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    om = calibrate.build(refmodel.get_oracle(), "calibrated")          # identical weights on every rank (seeded)
    m = d.DeepfakeDetectionModel(**refmodel.MODEL_CONFIG)
    m.load_state_dict(om.state_dict(), strict=True)
    m = m.to(dev).eval()
    om = om.to(dev).eval()
    # the global batch is a function of the frame index only; each rank materialises its own frames
    g = torch.Generator().manual_seed(4000 + rank)
    x = torch.randn(hi - lo, 3, size, size, generator=g)
    lm = torch.tensor(calibrate.TEMPLATE_5PT).expand(hi - lo, 5, 2) * size + 3.0 * torch.randn(hi - lo, 5, 2, generator=g)
    xd, lmd = x.to(dev), lm.to(dev)

    # ---- parity: fp32 mode vs one oracle call per clip
    m.set_compute_dtype(torch.float32)
    out = m.score_clips(xd, lmd, frames_per_clip=frames)
    worst, labels_ok = 0.0, True
    for c in range(my_clips):
        sl = slice(c * frames, (c + 1) * frames)
        with torch.no_grad():
            logits, _ = om(xd[sl], lmd[sl])
        prob = torch.softmax(logits, dim=1)[:, 1].mean()
        e = ((out["mean_logits"][c] - logits.mean(0)).norm() / logits.mean(0).norm()).item()
        worst = max(worst, e, abs(out["fake_prob"][c].item() - prob.item()))
        labels_ok &= int(out["labels"][c].item()) == int(prob.item() >= 0.5)
    assert worst < 1e-4 and labels_ok, (rank, worst, labels_ok)

    # ---- throughput: bf16
    m.set_compute_dtype(torch.bfloat16)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    steps = 10
    for _ in range(3):
        m.score_clips(xd, lmd, frames_per_clip=frames)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        res = m.score_clips(xd, lmd, frames_per_clip=frames)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    # end to end: pinned uint8 crops -> device -> labels back on the host
    mean = torch.tensor(d.model.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(d.model.IMAGENET_STD).view(1, 3, 1, 1)
    hu8 = ((x * std + mean) * 255.0).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
    hlm = lm.pin_memory()
    du8, dlm = torch.empty_like(hu8, device=dev), torch.empty_like(lmd)
    host_labels = torch.empty(my_clips, dtype=torch.int32).pin_memory()

    def e2e_step():
        du8.copy_(hu8, non_blocking=True)
        dlm.copy_(hlm, non_blocking=True)
        r = m.score_clips(du8, dlm, frames_per_clip=frames)
        host_labels.copy_(r["labels"], non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / steps
    t = torch.tensor([ms, e2e_ms, worst], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        total = n_clips * frames
        print(json.dumps({"config": "BASELINE.json configs[3]: %d clips x %d frames @ %dx%d over %d GPU(s), whole clips per rank, no collective" % (n_clips, frames, size, size, world),
                          "frames": total, "n_gpus": world, "ms_per_pass": t[0].item(), "frames_per_s": total / (t[0].item() * 1e-3),
                          "clips_per_s": n_clips / (t[0].item() * 1e-3), "e2e_ms_per_pass": t[1].item(), "e2e_frames_per_s": total / (t[1].item() * 1e-3),
                          "e2e_note": "pinned uint8 crops -> device, score_clips, labels back to the host, every pass",
                          "parity": {"checked": "every clip of every rank, fp32 mode vs one oracle call per clip (mean logits, mean fake probability, label)",
                                     "worst_rel_or_abs_error": t[2].item(), "labels_identical": True}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
