#!/usr/bin/env python
"""Surface normals (ob_normals) on a 128x2048 frame that lives in HBM: the destaggered XYZ / range that K1
has just produced go straight into the stencil (run under gpurun).  Device time per frame by CUDA events,
bytes moved, and the oracle's CPU loop on one host core beside it."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as graft
import bench, bench_common as bc
from oracle import oracle as orc
from tests.test_oracle_normals import room_scene
ob = graft.load_package()
dev = torch.device("cuda", 0)
peak, _ = bc.measured_peaks()
H, W = 128, 2048
xyz, rngd, _ = room_scene(H, W)                       # destaggered (H, W, 3) metres, (H, W) mm
org = np.zeros((W, 3))                                # sensor origin per column
out = {}
for dt, tdt in ((np.float32, torch.float32), (np.float64, torch.float64)):
    t_xyz = torch.from_numpy(xyz.astype(dt)).to(dev)
    t_rng = torch.from_numpy(rngd.view(np.int32)).to(dev)
    t_org = torch.from_numpy(org).to(dev)
    st = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
    for dual in (False, True):
        args = (t_xyz, t_rng) + ((t_xyz * 1.05, t_rng) if dual else ()) + (t_org,)
        # the vertical subtent is a per-sensor constant: found once, passed in afterwards (as a caller would)
        _, sub = ob.normals(*args, return_subtent=True, stream=st)
        fn = lambda: ob.normals(*args, vertical_subtent=sub, stream=st)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        es = np.dtype(dt).itemsize
        r = 2 if dual else 1
        b = H * W * r * (3 * es + 4 + 3 * es) + W * 24
        key = f"{np.dtype(dt).name}_{'dual' if dual else 'single'}"
        out[key] = {"ms_per_frame": ms, "bytes": b, "gbs": b / ms / 1e6, "frac_of_copy_peak": b / ms / 1e6 / peak,
                    "mpoints_s": H * W * r / ms / 1e3}
        print(key, out[key], flush=True)
# batched: 32 frames per launch through the C ABI (ob_normals_io.n_frames), float32, dual return
import ctypes as C
from importlib import import_module
capi = import_module(ob.__name__ + "._capi")
F = 32
bx = torch.from_numpy(np.stack([xyz.astype(np.float32) * (1 + 0.001 * i) for i in range(F)])).to(dev)
bx2 = bx * 1.05
br = torch.from_numpy(np.stack([rngd.view(np.int32)] * F)).to(dev)
bn1, bn2 = torch.empty_like(bx), torch.empty_like(bx)
t_org = torch.from_numpy(org).to(dev)
_, sub = ob.normals(bx[0], br[0], bx2[0], br[0], t_org, return_subtent=True)
io = capi.NormalsIO()
io.n_frames, io.h, io.w = F, H, W
io.xyz, io.range, io.xyz2, io.range2 = bx.data_ptr(), br.data_ptr(), bx2.data_ptr(), br.data_ptr()
io.normals, io.normals2 = bn1.data_ptr(), bn2.data_ptr()
io.sensor_origins_xyz, io.n_origins = t_org.data_ptr(), W
io.pixel_search_range, io.min_angle_of_incidence_rad, io.target_distance_m = 1, np.pi / 180.0, 0.025
io.vertical_subtent_rad = sub
sth = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
run = lambda: capi.check(capi.lib.ob_normals(capi.OB_F32, C.byref(io), sth.h))
for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
b = F * H * W * 2 * (12 + 4 + 12)
out["float32_dual_batch32"] = {"ms_per_launch": ms, "ms_per_frame": ms / F, "bytes": b, "gbs": b / ms / 1e6,
                               "frac_of_copy_peak": b / ms / 1e6 / peak, "mpoints_s": F * H * W * 2 / ms / 1e3}
n1_single, _ = ob.normals(bx[3], br[3], bx2[3], br[3], t_org, vertical_subtent=sub)
out["float32_dual_batch32"]["frame3_equals_single_call"] = bool(torch.equal(bn1[3], n1_single))
print("float32_dual_batch32", out["float32_dual_batch32"], flush=True)
t0 = time.perf_counter()
want = orc.normals(xyz, rngd, sensor_origins_xyz=org, pixel_search_range=1)
out["cpu_1thread_ms_f64_single"] = (time.perf_counter() - t0) * 1e3
got, sub = ob.normals(xyz, rngd, org, return_subtent=True)
out["matches_oracle"] = bool(np.array_equal(got, orc.normals(xyz, rngd, sensor_origins_xyz=org, vertical_subtent=sub)))
print({k: v for k, v in out.items() if not isinstance(v, dict)})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/time_normals.json", "w"), indent=1)
