#!/usr/bin/env python
"""K2 over the sweep shapes with 1 and 2 CTAs per SM (decode_pipe_ctas): ms per step, compulsory fraction of
the copy peak, round trip of the first/last frame (run under gpurun)."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as graft
import bench, bench_k2, bench_sweep, bench_common as bc
ob = graft.load_package()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream()
obs = ob.Stream(0, cuda_stream=stream.cuda_stream)
peak, _ = bc.measured_peaks()
tdt = {1: torch.uint8, 2: torch.int16, 4: torch.int32}
out = []
for (h, w) in [x for x in bench_sweep.SHAPES if x[0] <= int(os.environ.get('MAXH', '128'))]:
    shifts = np.tile(np.array([3 * (w // 128), 2 * (w // 128), w // 128, 0], np.int32), h // 4)
    d, o = bench.synth_lut(seed=43, h=h, w=w)
    lut = ob.XYZLutT.from_arrays(torch.from_numpy(d).to(dev), torch.from_numpy(o).to(dev), h, w, device=0)
    for returns in (1, 2):
        si, pk, src = bench_k2.synth_packets(ob, 2, seed=(0xdeadbeef + h * 7 + w) & 0x7fffffff,
                                             profile=bench_sweep.PROFILES[returns], h=h, w=w, shifts=shifts)
        n_slots, psz = pk.shape[1], pk.shape[2]
        dec = ob.Decoder.from_sensor(si, src[0], device=0)
        fbytes = sum(f["elem_size"] for f in dec.fields)
        _, c2 = bc.k2_bytes(h, w, returns, 1, psz, bench_k2.CPP, fbytes)
        F2 = int(max(8, min(1024, bench_sweep.TARGET_BYTES // c2)))
        t_pk = torch.from_numpy(np.stack([pk[i % 2] for i in range(F2)])).to(dev)
        fields = {f["name"]: torch.empty((F2, h, w), dtype=tdt[f["elem_size"]], device=dev) for f in dec.fields}
        xyz = [torch.empty((F2, h * w, 3), dtype=torch.float32, device=dev) for _ in range(returns)]
        rd = [torch.empty((F2, h, w), dtype=torch.int32, device=dev) for _ in range(returns)]
        t_ts = torch.empty((F2, w), dtype=torch.int64, device=dev)
        t_mid = torch.empty((F2, w), dtype=torch.int16, device=dev)
        t_st = torch.empty((F2, w), dtype=torch.int32, device=dev)
        alg, comp = bc.k2_bytes(h, w, returns, F2, psz, bench_k2.CPP, fbytes)
        for ctas in (1, 2, 3, 4):
            ob.set_tunable("decode_pipe_ctas", ctas)
            for t in list(fields.values()) + xyz + rd:
                t.zero_()
            lp0 = ob.kernel_launch_count("decode_pipe")
            plan = dec.prepare_batch(F2, t_pk, n_slots, psz, n_slots * psz, fields, lut=lut, pixel_shift_by_row=shifts,
                                     xyz=xyz, range_destaggered=rd, timestamp=t_ts, measurement_id=t_mid, status=t_st,
                                     stream=obs)
            s = bench_sweep._time(torch, stream, plan, 10, 3)
            piped = ob.kernel_launch_count("decode_pipe") > lp0
            ok = True
            for i in (0, F2 - 1):
                for f in dec.fields:
                    ok &= bool(np.array_equal(fields[f["name"]][i].cpu().numpy().view(src[i % 2].field(f["name"]).dtype),
                                              src[i % 2].field(f["name"])))
            e = {"shape": f"{h}x{w}", "returns": returns, "ctas": ctas, "F": F2, "ms": s * 1e3,
                 "frac": comp / s / 1e9 / peak, "piped": bool(piped), "roundtrip_ok": ok,
                 "xyz_sum": float(xyz[0][0].double().abs().sum().item())}
            out.append(e)
            print(e, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_k2_ctas.json", "w"), indent=1)
