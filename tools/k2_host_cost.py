#!/usr/bin/env python
"""Host time per prepared K2 launch vs its device time, for many-frame launches of small shapes (is the sweep
host-submission-bound?).  Run under gpurun."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as graft
import bench, bench_k2, bench_sweep, bench_common as bc
ob = graft.load_package()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream()
obs = ob.Stream(0, cuda_stream=stream.cuda_stream)
tdt = {1: torch.uint8, 2: torch.int16, 4: torch.int32}
for (h, w, returns) in ((32, 512, 1), (32, 1024, 2), (64, 1024, 2), (128, 2048, 2)):
    shifts = np.tile(np.array([3 * (w // 128), 2 * (w // 128), w // 128, 0], np.int32), h // 4)
    d, o = bench.synth_lut(seed=43, h=h, w=w)
    lut = ob.XYZLutT.from_arrays(torch.from_numpy(d).to(dev), torch.from_numpy(o).to(dev), h, w, device=0)
    si, pk, src = bench_k2.synth_packets(ob, 2, seed=5, profile=bench_sweep.PROFILES[returns], h=h, w=w, shifts=shifts)
    n_slots, psz = pk.shape[1], pk.shape[2]
    dec = ob.Decoder.from_sensor(si, src[0], device=0)
    fbytes = sum(f["elem_size"] for f in dec.fields)
    _, c2 = bc.k2_bytes(h, w, returns, 1, psz, bench_k2.CPP, fbytes)
    F2 = int(max(8, min(1024, bench_sweep.TARGET_BYTES // c2)))
    t_pk = torch.from_numpy(np.stack([pk[i % 2] for i in range(F2)])).to(dev)
    fields = {f["name"]: torch.empty((F2, h, w), dtype=tdt[f["elem_size"]], device=dev) for f in dec.fields}
    xyz = [torch.empty((F2, h * w, 3), dtype=torch.float32, device=dev) for _ in range(returns)]
    rd = [torch.empty((F2, h, w), dtype=torch.int32, device=dev) for _ in range(returns)]
    plan = dec.prepare_batch(F2, t_pk, n_slots, psz, n_slots * psz, fields, lut=lut, pixel_shift_by_row=shifts, xyz=xyz,
                             range_destaggered=rd, stream=obs)
    for _ in range(5):
        plan()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(50):
        plan()
    e1.record(stream)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{h}x{w} r{returns} F={F2}: host {1e3*(t1-t0)/50:.4f} ms per call, device {e0.elapsed_time(e1)/50:.4f} ms per launch", flush=True)
