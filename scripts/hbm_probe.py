"""What this B200 sustains for read-only, write-only and copy streams (1 GiB buffers, CUDA events, best of 10) --
context for the write-dominated 1x1-conv GEMMs (an expand GEMM writes 6x what it reads)."""
import json, torch
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
af, bf = a.view(torch.float32), b.view(torch.float32)
def best(fn, nbytes, reps=10):
    for _ in range(2): fn()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return nbytes / min(t) / 1e6
out = {"copy_gbs": best(lambda: b.copy_(a), 2 * n), "write_only_gbs": best(lambda: a.zero_(), n),
       "read_only_gbs": best(lambda: af.sum(), n)}
# 6:1 write:read mix, like an expand GEMM: read 1/6 GiB, write 1 GiB
src = af[: n // 24]
out["write6_read1_gbs"] = best(lambda: torch.cat([src] * 6, out=bf[: n // 4]), n // 4 * 4 + n // 24 * 4 * 1)
print(json.dumps(out))
