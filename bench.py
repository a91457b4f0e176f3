#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract in the task statement, section (4)).

A "step" is one eval-mode forward pass of DeepfakeDetectionModel over one batch of synthetic
380x380 face crops + 5-point landmarks.  At every N the per-GPU workload is BASELINE.json
configs[1] (batch 256, bf16, folded BN) -- weak scaling, no data-path collective (inference needs
none).  `value` is images/s with inputs resident in HBM; `e2e` is the same metric
through the public nn.Module call with HOST (pinned) inputs, host->device copies and the
device->host read of the logits inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec @380x380 fwd (bf16)"
SIZE = 380


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clocks / throttle reasons while the timed region runs.  NVML (nvidia_ml_py) in-process when it is
    importable: spawning nvidia-smi every 200 ms re-initialises the driver interface each time and was seen to stall the
    launch thread of the 3475-launch training step for tens of milliseconds (30.5 -> 38 ms per step in two runs of ten);
    nvidia-smi stays as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.sample_nvml()      # first-call costs of the queries are paid here, outside the timed region
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        act = lambda bit: "Active" if mask & bit else "Not Active"
        return [str(sm), str(self.max_sm), act(0x8), act(0x40), act(0x20), act(0x4)]     # hw, hw thermal, sw thermal, sw power cap

    def run(self):
        while not self.stop.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self.sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.1 if self.nvml is not None else 0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, averaged over this kernel's launches of one
    step) from the committed ncu pass profiles/r01_traffic.json (scripts/gpu_final_profile.sh); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            return json.load(f).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def cpu_reference_run(steps, warmup, batch=8):
    """The reference's own CPU implementation of the path (reference wrapper code over the
    efficientnet-pytorch restatement; the real import when /root/reference is mounted, else its
    restatement), fp32 eval forward, all host threads.  Bounded sample: `batch` images per step."""
    import torch
    from oracle import calibrate, refmodel
    ns = refmodel.get_oracle()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = calibrate.build(ns, "default")
    x, lm, _ = calibrate.synthetic_batch(batch, SIZE)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            model(x, lm)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": batch * steps / total, "unit": "images/s", "cores": torch.get_num_threads(),
            "kind": "reference" if ns.kind == "reference" else "port",
            "sample": f"{steps} eval forwards of batch {batch} @ {SIZE}x{SIZE} fp32 (BASELINE.json configs[0]), "
                      f"{warmup} warm-up, torch CPU {cores} threads",
            "ms_per_step": 1e3 * total / steps}


def train_bench(args, d, _lib, dev, rank, world, local, warmup):
    """BASELINE.json configs[2]: training step = train-mode forward + CombinedLoss + backward + NCCL gradient
    all-reduce (inside backward), batch 64 per GPU, bf16 activations / fp32 master weights and gradients."""
    import torch
    import torch.distributed as dist
    B = args.batch if args.batch != 256 else 64
    torch.manual_seed(42)
    model = d.DeepfakeDetectionModel(**d.DEFAULT_MODEL_CONFIG).to(dev).train().set_compute_dtype(torch.bfloat16)
    if world > 1:
        d.parallel.broadcast_parameters(model)
    crit = d.CombinedLoss({"ce": 1.0, "focal": 0.5, "contrastive": 0.2}, torch.tensor([1.0, 1.5], device=dev))
    g = torch.Generator().manual_seed(1234 + rank)
    host_x = torch.randn(B, 3, SIZE, SIZE, generator=g).pin_memory()
    host_lm = (torch.rand(B, 5, 2, generator=g) * SIZE).pin_memory()
    host_y = torch.randint(0, 2, (B,), generator=g).pin_memory()
    x, lm, y = host_x.to(dev), host_lm.to(dev), host_y.to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step(xx, ll, yy):
        model.zero_grad(set_to_none=True)
        lo, fe = model(xx, ll, return_features=True)
        loss = crit(lo, yy, fe)["total"]
        loss.backward()
        return loss

    for _ in range(warmup):
        step(x, lm, y)
    barrier()
    _lib.lib.dfv_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gc.collect()
    gc.disable()        # a collector pause inside a 3475-launch step drains the launch queue (seen as sporadic 33-38 ms steps)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step(x, lm, y)
        ev1.record()
        barrier()
    gc.enable()
    ms = ev0.elapsed_time(ev1)
    launches = int(_lib.lib.dfv_launch_count(0))

    _lib.lib.dfv_profile_enable(1)
    prof_steps = min(args.steps, 3)
    for _ in range(prof_steps):
        step(x, lm, y)
    torch.cuda.synchronize()
    recs = _lib.profile_records()
    _lib.lib.dfv_profile_enable(0)
    if os.environ.get("DFV_BENCH_DUMP"):          # per-launch list of the last profiled step
        per = len(recs) // prof_steps
        with open(os.environ["DFV_BENCH_DUMP"], "w") as f:
            json.dump([{"kind": k, "bytes": b, "flops": fl, "ms": m} for k, b, fl, m in recs[-per:]], f)
    agg = {}
    for kind, nbytes, flops, kms in recs:
        a = agg.setdefault(kind, [0.0, 0.0, 0.0, 0])
        a[0] += nbytes; a[1] += flops; a[2] += kms; a[3] += 1
    pk = peaks()
    kernels = {k: {"launches_per_step": n // prof_steps, "ms_per_step": kms / prof_steps,
                   "gbs": nb / (kms * 1e-3) / 1e9 if kms > 0 else None, "tflops": fl / (kms * 1e-3) / 1e12 if kms > 0 else None,
                   "hbm_frac": nb / (kms * 1e-3) / 1e9 / pk["hbm_gbs"] if kms > 0 else None}
               for k, (nb, fl, kms, n) in agg.items()}
    top = max(agg, key=lambda k: agg[k][2])
    tb, tf, tms, tn = agg[top]
    roofline = {"kernel": top, "bound": "hbm", "achieved": tb / (tms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": tb / (tms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": measured_traffic(top), "peak_source": pk["source"],
                "launches": tn // prof_steps, "avg_launch_ms": tms / tn, "algorithmic_bytes_per_launch": tb / tn, "kernels": kernels}

    # end to end: pinned host batch -> device every step, loss read back every step
    dx, dl, dy = torch.empty_like(x), torch.empty_like(lm), torch.empty_like(y)

    def e2e_step():
        dx.copy_(host_x, non_blocking=True)
        dl.copy_(host_lm, non_blocking=True)
        dy.copy_(host_y, non_blocking=True)
        return step(dx, dl, dy).item()

    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = t[0].item(), t[1].item()
    n_gpus = max(world, 1)
    if rank == 0:
        line = {"metric": "images/sec @380x380 train-step (bf16)", "value": B * n_gpus * args.steps / (ms * 1e-3), "unit": "images/s",
                "n_gpus": n_gpus, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"training step fwd + CombinedLoss(CE+Focal+Contrastive, class weights) + bwd + gradient all-reduce, "
                                       f"batch {B}/GPU @ {SIZE}x{SIZE} (BASELINE.json configs[2]); optimizer step not included",
                           "per_gpu_batch": B, "global_batch": B * n_gpus, "image_size": SIZE,
                           "parallelism": f"dp{n_gpus} (batch sharded, one NCCL all-reduce of the flat fp32 gradient buffer per step)",
                           "l2_policy": "activations exceed the 126 MB L2; no flush needed"},
                "roofline": roofline, "cpu_baseline": None,
                "e2e": {"value": B * n_gpus * args.steps / (e2e_ms * 1e-3), "unit": "images/s",
                        "h2d_bytes_per_step": (x.numel() * 4 + lm.numel() * 4 + y.numel() * 8) * n_gpus, "d2h_bytes_per_step": 4 * n_gpus,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches, "clocks": clk.summary(),
                "memory_gb": torch.cuda.max_memory_allocated() / 1e9}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer = BASELINE.json configs[1] (the headline line); train = configs[2], fwd + CombinedLoss + bwd + "
                         "gradient all-reduce at batch 64/GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    config = {"workload": f"batch-{args.batch}/GPU bf16 eval forward @ {SIZE}x{SIZE}, folded BN, fused SE epilogues "
                          "(BASELINE.json configs[1])",
              "per_gpu_batch": args.batch, "global_batch": args.batch * max(world, 1), "image_size": SIZE,
              "parallelism": f"dp{max(world, 1)} (batch sharded, no collective)",
              "l2_policy": "activations (>=0.9 GB per layer at batch 256) exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = min(args.steps, 8)
        cb = cpu_reference_run(steps, min(warmup, 2))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "images/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": min(warmup, 2), "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import deepfake_vit_b200 as d
    from deepfake_vit_b200 import _lib
    MODEL_CONFIG = d.DEFAULT_MODEL_CONFIG         # the YAML `model:` mapping

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a B200"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.mode == "train":
        return train_bench(args, d, _lib, dev, rank, world, local, warmup)

    torch.manual_seed(42)
    model = d.DeepfakeDetectionModel(**MODEL_CONFIG).to(dev).eval().set_compute_dtype(torch.bfloat16)
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    host_x = [torch.randn(B, 3, SIZE, SIZE, generator=g).pin_memory() for _ in range(2)]
    host_lm = [(torch.rand(B, 5, 2, generator=g) * SIZE).pin_memory() for _ in range(2)]
    x, lm = host_x[0].to(dev), host_lm[0].to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput
    for _ in range(warmup):
        model(x, lm)
    barrier()
    _lib.lib.dfv_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gc.collect()
    gc.disable()        # no collector pause inside the timed region
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            logits, _ = model(x, lm)
        ev1.record()
        barrier()
    gc.enable()
    ms = ev0.elapsed_time(ev1)
    launches = int(_lib.lib.dfv_launch_count(0))

    # ---- per-kernel CUDA-event profile on the same stream, same buffers (roofline leg)
    _lib.lib.dfv_profile_enable(1)
    prof_steps = min(args.steps, 5)
    for _ in range(prof_steps):
        model(x, lm)
    torch.cuda.synchronize()
    recs = _lib.profile_records()
    _lib.lib.dfv_profile_enable(0)
    if os.environ.get("DFV_BENCH_DUMP"):          # per-launch list of the last profiled step (layer-by-layer analysis)
        per = len(recs) // prof_steps
        with open(os.environ["DFV_BENCH_DUMP"], "w") as f:
            json.dump([{"kind": k, "bytes": b, "flops": fl, "ms": m} for k, b, fl, m in recs[-per:]], f)
    agg = {}
    for kind, nbytes, flops, kms in recs:
        a = agg.setdefault(kind, [0.0, 0.0, 0.0, 0])
        a[0] += nbytes; a[1] += flops; a[2] += kms; a[3] += 1
    pk = peaks()
    kernels = {}
    for kind, (nbytes, flops, kms, n) in agg.items():
        kernels[kind] = {"launches_per_step": n // prof_steps, "ms_per_step": kms / prof_steps,
                         "gbs": nbytes / (kms * 1e-3) / 1e9 if kms > 0 else None,
                         "tflops": flops / (kms * 1e-3) / 1e12 if kms > 0 else None,
                         "hbm_frac": nbytes / (kms * 1e-3) / 1e9 / pk["hbm_gbs"] if kms > 0 else None}
    top = max(agg, key=lambda k: agg[k][2])
    tb, tf, tms, tn = agg[top]
    roofline = {"kernel": top, "bound": "hbm", "achieved": tb / (tms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": tb / (tms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": measured_traffic(top), "peak_source": pk["source"],
                "launches": tn // prof_steps, "avg_launch_ms": tms / tn,
                "algorithmic_bytes_per_launch": tb / tn, "kernels": kernels,
                "note": "achieved = sum of algorithmic bytes of this kernel's launches / sum of their CUDA-event "
                        "durations, measured on the launch stream in a profiled pass right after the timed region"}

    # ---- end to end through the public API: pinned host inputs, double-buffered H2D, D2H of logits
    copy_stream = torch.cuda.Stream(device=dev)
    dev_x = [torch.empty_like(x) for _ in range(2)]
    dev_lm = [torch.empty_like(lm) for _ in range(2)]
    host_out = torch.empty(B, 2).pin_memory()
    ready = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    def stage(i):
        s = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[s])          # buffer free (previous user finished)
            dev_x[s].copy_(host_x[s], non_blocking=True)
            dev_lm[s].copy_(host_lm[s], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n):
        cur = torch.cuda.current_stream()
        stage(0)
        for i in range(n):
            s = i & 1
            if i + 1 < n:
                stage(i + 1)
            cur.wait_event(ready[s])
            lo, _ = model(dev_x[s], dev_lm[s])
            done[s].record(cur)
            host_out.copy_(lo, non_blocking=True)
        cur.synchronize()

    for s in range(2):
        done[s].record(torch.cuda.current_stream())
    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e2e_ms, 0.0) if e2e_ms > 0 else wall_ms

    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = t[0].item(), t[1].item()
    n_gpus = max(world, 1)
    value = B * n_gpus * args.steps / (ms * 1e-3)
    e2e_value = B * n_gpus * args.steps / (e2e_ms * 1e-3)

    if rank == 0:
        cpu = None
        if n_gpus == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_run(5, 2)
            cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": n_gpus, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
                "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "images/s",
                        "h2d_bytes_per_step": (x.numel() * 4 + lm.numel() * 4) * n_gpus,
                        "d2h_bytes_per_step": B * 2 * 4 * n_gpus, "ms_per_step": e2e_ms / args.steps,
                        "note": "model(images, landmarks) with pinned host inputs, double-buffered H2D on a copy stream"},
                "gpu_launches": launches, "clocks": clk.summary(),
                "end_to_end_roofline": {"algorithmic_bytes_per_image": 204.7e6, "bound_images_per_s_per_gpu": pk["hbm_gbs"] * 1e9 / 204.7e6,
                                        "frac": value / n_gpus / (pk["hbm_gbs"] * 1e9 / 204.7e6)}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
