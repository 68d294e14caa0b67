#!/usr/bin/env python
"""K1 on the 512-wide single-return shape: the single-return launch geometry (1024-pixel tiles, 4 stages,
2 CTAs/SM) vs the dual-return one (512, 3, 3) (run under gpurun)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as graft
import bench, bench_sweep, bench_common as bc
ob = graft.load_package()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream()
obs = ob.Stream(0, cuda_stream=stream.cuda_stream)
peak, _ = bc.measured_peaks()
cases = []
for (h, w) in ((32, 512), (32, 1024), (64, 1024)):
    shifts = np.tile(np.array([3 * (w // 128), 2 * (w // 128), w // 128, 0], np.int32), h // 4)
    d, o = bench.synth_lut(seed=43, h=h, w=w)
    lut = ob.XYZLutT.from_arrays(torch.from_numpy(d).to(dev), torch.from_numpy(o).to(dev), h, w, device=0)
    _, c1 = bc.k1_bytes(h, w, 1, 1)
    F = int(max(8, min(2048, bench_sweep.TARGET_BYTES_K1 // c1)))
    pool = bench.synth_pool(16, seed=7, h=h, w=w, returns=1)
    t_rng = torch.from_numpy(np.concatenate([pool] * ((F + 15) // 16))[:F].view(np.int32)).to(dev)
    t_xyz = torch.empty((F, 1, h * w, 3), dtype=torch.float32, device=dev)
    t_rd = torch.empty((F, 1, h, w), dtype=torch.int32, device=dev)
    _, comp = bc.k1_bytes(h, w, 1, F)
    cases.append((h, w, shifts, lut, t_rng, t_xyz, t_rd, comp))
for mode in ("auto", "tw512_s3_c3", "tw512_s4_c4", "tw256_s3_c6"):
    if mode != "auto":
        _, tw, st, ct = mode.replace("tw", "").replace("s", "").replace("c", "").split("_")[0], *[int(x) for x in mode.replace("tw", "").replace("s", "").replace("c", "").split("_")]
        ob.set_tunable("cloud_tw", tw); ob.set_tunable("cloud_stages", st); ob.set_tunable("cloud_ctas_per_sm", ct)
    for (h, w, shifts, lut, t_rng, t_xyz, t_rd, comp) in cases:
        s = bench_sweep._time(torch, stream, lambda: ob.scan_to_cloud(lut, shifts, t_rng, xyz=t_xyz, range_destaggered=t_rd, stream=obs), 10, 3)
        print(f"{h}x{w} single {mode}: {s*1e3:.4f} ms frac {comp / s / 1e9 / peak:.3f}", flush=True)
