#!/usr/bin/env python
"""Launch-size sweep (SURVEY 7 'launch latency vs work size'): K1 (f32 and f64) and K2 at 128x2048
dual return for 1..64 frames per launch.  -> gpurun_out/sweep_batch.{json,md}"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as graft
import bench, bench_k2 as k2
ob = graft.load_package()
dev = torch.device("cuda", 0)
H, W, R = bench.H, bench.W, bench.R
peak, _ = bench.measured_peaks()
st = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
FMAX = 64
rng = torch.from_numpy(bench.synth_pool(FMAX).view(np.int32)).to(dev)
d, o = bench.synth_lut()
rows = []


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


for dtype, tdt, bpf in ((np.float32, torch.float32, 16777216), (np.float64, torch.float64, 29360128)):
    lut = ob.XYZLutT.from_arrays(torch.from_numpy(d.astype(dtype)).to(dev), torch.from_numpy(o.astype(dtype)).to(dev), H, W)
    xyz = torch.empty((FMAX, R, H * W, 3), dtype=tdt, device=dev)
    rd = torch.empty((FMAX, R, H, W), dtype=torch.int32, device=dev)
    for F in (1, 2, 4, 8, 16, 32, 64):
        t = timeit(lambda: ob.scan_to_cloud(lut, bench.SHIFTS, rng[:F], xyz=xyz[:F], range_destaggered=rd[:F], stream=st))
        rows.append({"kernel": "K1", "dtype": np.dtype(dtype).name, "frames": F, "us_per_launch": t * 1e6,
                     "us_per_frame": t * 1e6 / F, "mpoints_s": F * H * W * R / t / 1e6,
                     "gbps": F * bpf / t / 1e9, "frac": F * bpf / t / 1e9 / peak})
    del xyz, rd
si, pk, src = k2.synth_packets(ob, 4)
n_slots, psz = pk.shape[1], pk.shape[2]
t_pk = torch.from_numpy(np.stack([pk[i % 4] for i in range(32)])).to(dev)
lut = ob.XYZLutT.from_arrays(torch.from_numpy(d).to(dev), torch.from_numpy(o).to(dev), H, W)
dec = ob.Decoder.from_sensor(si, src[0])
tdt = {1: torch.uint8, 2: torch.int16, 4: torch.int32}
fields = {f["name"]: torch.empty((32, H, W), dtype=tdt[f["elem_size"]], device=dev) for f in dec.fields}
xyz = [torch.empty((32, H * W, 3), dtype=torch.float32, device=dev) for _ in range(R)]
rd = [torch.empty((32, H, W), dtype=torch.int32, device=dev) for _ in range(R)]
for F in (1, 2, 4, 8, 16, 32):
    t = timeit(lambda: dec.decode_batch(F, t_pk, n_slots, psz, n_slots * psz, fields, lut=lut,
                                        pixel_shift_by_row=k2.SHIFTS, xyz=xyz, range_destaggered=rd, stream=st))
    rows.append({"kernel": "K2", "dtype": "float32", "frames": F, "us_per_launch": t * 1e6,
                 "us_per_frame": t * 1e6 / F, "mpoints_s": F * H * W * R / t / 1e6,
                 "gbps": F * k2.K2_BYTES_PER_FRAME_F32 / t / 1e9,
                 "frac": F * k2.K2_BYTES_PER_FRAME_F32 / t / 1e9 / peak})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep_batch.json", "w"), indent=0)
md = ["| kernel | dtype | frames/launch | us/launch | us/frame | Mpoints/s | GB/s (alg.) | frac |", "|---|---|---|---|---|---|---|---|"]
for r in rows:
    md.append(f"| {r['kernel']} | {r['dtype']} | {r['frames']} | {r['us_per_launch']:.1f} | {r['us_per_frame']:.2f} | "
              f"{r['mpoints_s']:.0f} | {r['gbps']:.0f} | {r['frac']:.2f} |")
open("gpurun_out/sweep_batch.md", "w").write("\n".join(md) + "\n")
print("\n".join(md))
