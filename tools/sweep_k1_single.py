#!/usr/bin/env python
"""Launch-geometry sweep of K1 for SINGLE-return frames, LUT path and LUT-free path (run under gpurun)."""
import itertools, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as graft
import bench, bench_common as bc
ob = graft.load_package()
F, H, W = 128, bench.H, bench.W
dev = torch.device("cuda", 0)
meta = json.load(open("tests/golden/OS-1-128_767798045_1024x10_20230712_120049.json"))
args_i = (W, H, 0.001, meta["beam_to_lidar_transform"], meta["lidar_to_sensor_transform"],
          meta["beam_azimuth_angles"], meta["beam_altitude_angles"])
lut = ob.XYZLutT.from_intrinsics(*args_i, dtype=np.float32)
st = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
peak, _ = bc.measured_peaks()
rows = []
for returns in (1, 2):
    rng = torch.from_numpy(bench.synth_pool(F, returns=returns).view(np.int32)).to(dev)
    xyz = torch.empty((F, returns, H * W, 3), dtype=torch.float32, device=dev)
    rd = torch.empty((F, returns, H, W), dtype=torch.int32, device=dev)
    for analytic in (0, 1):
        lut.set_analytic(bool(analytic))
        for tw, stg, cta, th in itertools.product([512, 1024, 2048], [2, 3, 4], [2, 3, 4, 6], [128, 256]):
            smem = 256 + stg * (tw * 24 + returns * 4 * tw)
            if smem * cta > 225 * 1024 or (th + 32) * cta > 2048:
                continue
            for k, v in (("cloud_tw", tw), ("cloud_stages", stg), ("cloud_ctas_per_sm", cta), ("cloud_threads", th)):
                ob.set_tunable(k, v)
            plan = ob.plan_scan_to_cloud(lut, bench.SHIFTS, rng, xyz=xyz, range_destaggered=rd, stream=st)
            for _ in range(2):
                plan()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                plan()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            alg, comp = bc.k1_bytes(H, W, returns, F, n_luts=0 if analytic else 1)
            rows.append({"returns": returns, "analytic": analytic, "tw": tw, "stages": stg, "ctas": cta, "threads": th,
                         "ms": ms, "frac": comp / (ms * 1e-3) / 1e9 / peak, "gpts": F * H * W * returns / (ms * 1e-3) / 1e9})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep_k1_single.json", "w"), indent=0)
for returns in (1, 2):
    for analytic in (0, 1):
        sel = sorted([r for r in rows if r["returns"] == returns and r["analytic"] == analytic], key=lambda r: r["ms"])
        print("returns", returns, "analytic", analytic)
        for r in sel[:4]:
            print("  ", r)
        dflt = [r for r in sel if (r["tw"], r["stages"], r["ctas"], r["threads"]) == (512, 3, 3, 128)]
        print("   default", dflt[0] if dflt else None)
