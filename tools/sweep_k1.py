#!/usr/bin/env python
"""In-process sweep of the K1 launch geometry (run under gpurun). Prints the best configs."""
import itertools, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as graft
import bench
ob = graft.load_package()
F = int(os.environ.get("SWEEP_FRAMES", "64"))
H, W, R = bench.H, bench.W, bench.R
dev = torch.device("cuda", 0)
rng = torch.from_numpy(bench.synth_pool(F).view(np.int32)).to(dev)
d, o = bench.synth_lut()
lut = ob.XYZLutT.from_arrays(torch.from_numpy(d).to(dev), torch.from_numpy(o).to(dev), H, W)
xyz = torch.empty((F, R, H * W, 3), dtype=torch.float32, device=dev)
rd = torch.empty((F, R, H, W), dtype=torch.int32, device=dev)
st = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
peak, _ = bench.measured_peaks()
rows = []
grid = list(itertools.product([0, 1], [256, 512, 1024], [3, 4, 5, 6], [2, 3, 4, 6, 8], [64, 128, 256]))
for lag, tw, stg, cta, th in grid:
    smem = 256 + stg * (tw * 32)
    if smem * cta > 225 * 1024 or smem > 225 * 1024 or (th + 32) * cta > 2048:
        continue
    for k, v in (("cloud_store_lag", lag), ("cloud_tw", tw), ("cloud_stages", stg), ("cloud_ctas_per_sm", cta), ("cloud_threads", th)):
        ob.set_tunable(k, v)
    try:
        for _ in range(2):
            ob.scan_to_cloud(lut, bench.SHIFTS, rng, xyz=xyz, range_destaggered=rd, stream=st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        n = 5
        for _ in range(n):
            ob.scan_to_cloud(lut, bench.SHIFTS, rng, xyz=xyz, range_destaggered=rd, stream=st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    except Exception as ex:
        print("fail", lag, tw, stg, cta, th, ex)
        continue
    gbps = bench.K1_BYTES_PER_FRAME_F32 * F / (ms * 1e-3) / 1e9
    rows.append({"lag": lag, "tw": tw, "stages": stg, "ctas": cta, "threads": th, "ms": ms, "gbps": gbps, "frac": gbps / peak})
rows.sort(key=lambda r: -r["gbps"])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep_k1.json", "w"), indent=0)
for r in rows[:12]:
    print(r)
print("worst", rows[-1])
