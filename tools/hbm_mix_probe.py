#!/usr/bin/env python
"""Calibrates the HBM roofline for write-heavy mixes (K1 writes 3.4x what it reads from DRAM):
copy (1R:1W), fill (0R:1W), and a 1R:4W pattern, all with torch ops on 1 GiB-class buffers."""
import json, os, torch
dev = torch.device("cuda", 0)
n = 256 * 1024 * 1024  # float32 elements = 1 GiB
a = torch.empty(n, dtype=torch.float32, device=dev).normal_()
b = torch.empty_like(a)
outs = [torch.empty(n // 4, dtype=torch.float32, device=dev) for _ in range(4)]


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best

res = {}
t = timeit(lambda: b.copy_(a)); res["copy_1R_1W_GBps"] = 2 * n * 4 / t / 1e9
t = timeit(lambda: b.fill_(1.0)); res["fill_0R_1W_GBps"] = n * 4 / t / 1e9
t = timeit(lambda: torch.sum(a)); res["read_1R_0W_GBps"] = n * 4 / t / 1e9
src = a[: n // 4]
def one_to_four():
    for o in outs:
        o.copy_(src)
t = timeit(one_to_four); res["copy_small_src_4W_GBps_counting_1R_4W"] = (n // 4 * 4 + 4 * (n // 4) * 4) / t / 1e9
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/hbm_mix_probe.json", "w"), indent=1)
print(json.dumps(res))
