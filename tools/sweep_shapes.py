#!/usr/bin/env python
"""BASELINE configs[4]: frame-size sweep 32x512..128x2048, single+dual return, K1 and K2:
Mpoints/s and algorithmic HBM GB/s vs the measured roofline, next to the CPU oracle (1 thread).
Run under gpurun (1 GPU):  python tools/sweep_shapes.py  -> gpurun_out/sweep_shapes.{json,md}"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as graft
import bench
from oracle import oracle as orc
from tests.helpers import oracle_pf

ob = graft.load_package()
dev = torch.device("cuda", 0)
peak, _ = bench.measured_peaks()
SHAPES = [(32, 512), (32, 1024), (64, 1024), (64, 2048), (128, 1024), (128, 2048)]
PROFILES = {1: "RNG19_RFL8_SIG16_NIR16", 2: "RNG19_RFL8_SIG16_NIR16_DUAL"}
st = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


rows = []
for (h, w) in SHAPES:
    n = h * w
    shifts = np.tile(np.array([3, 2, 1, 0], np.int32) * (w // 128), h // 4)
    rs = np.random.default_rng(1)
    d = (rs.random((n, 3)) + 0.5).astype(np.float32)
    o = (rs.random((n, 3)) * 0.01).astype(np.float32)
    lut = ob.XYZLutT.from_arrays(torch.from_numpy(d).to(dev), torch.from_numpy(o).to(dev), h, w)
    for R in (1, 2):
        # ---------------- K1 ----------------
        k1_bytes = n * (4 * R + 24 + 12 * R + 4 * R)
        F = max(8, int(1.0e9 // k1_bytes))
        rng_h = rs.integers(1, 1 << 19, size=(F, R, h, w), dtype=np.uint32)
        rng_h[rs.random(rng_h.shape) < 0.5] = 0
        t_rng = torch.from_numpy(rng_h.view(np.int32)).to(dev)
        t_xyz = torch.empty((F, R, n, 3), dtype=torch.float32, device=dev)
        t_rd = torch.empty((F, R, h, w), dtype=torch.int32, device=dev)
        t = timeit(lambda: ob.scan_to_cloud(lut, shifts, t_rng, xyz=t_xyz, range_destaggered=t_rd, stream=st))
        t0 = time.perf_counter()
        for r in range(R):
            orc.destagger(rng_h[0, r], shifts)
            orc.cartesian(rng_h[0, r], d, o)
        cpu1 = time.perf_counter() - t0
        rows.append({"kernel": "K1", "h": h, "w": w, "returns": R, "frames_per_launch": F,
                     "bytes_per_frame": k1_bytes, "ms_per_launch": t * 1e3,
                     "mpoints_s": F * n * R / t / 1e6, "gbps": F * k1_bytes / t / 1e9,
                     "frac": F * k1_bytes / t / 1e9 / peak, "cpu_1thread_mpoints_s": n * R / cpu1 / 1e6})
        del t_rng, t_xyz, t_rd
        # ---------------- K2 ----------------
        prof = PROFILES[R]
        si = ob.SensorInfo(prof, h, w, 16, fw_rev="v3.2.1", pixel_shift_by_row=shifts)
        masks = {f[0]: f[6] for f in si.fields()}
        fr = ob.LidarFrame(si)
        for name in fr.fields:
            a = fr.field(name)
            a[...] = (rs.integers(0, 1 << 32, size=a.shape, dtype=np.uint64) & np.uint64(masks[name])).astype(a.dtype)
        fr.measurement_id[:] = np.arange(w)
        fr.timestamp[:] = 1000 + np.arange(w)
        fr.status[:] = 1
        fr.packet_timestamp[:] = 10 + np.arange(w // 16)
        fr.frame_id = 700
        pk, ts = ob.frame_to_packets(fr, si)
        n_slots, psz = pk.shape
        dec = ob.Decoder.from_sensor(si, fr)
        field_bytes = sum(f["elem_size"] for f in dec.fields)
        k2_bytes = n_slots * psz + 24 * n + field_bytes * n + 12 * R * n + 4 * R * n + 14 * w + 9 * (w // 16)
        F2 = max(8, int(0.7e9 // k2_bytes))
        t_pk = torch.from_numpy(np.broadcast_to(pk, (F2,) + pk.shape).copy()).to(dev)
        tdt = {1: torch.uint8, 2: torch.int16, 4: torch.int32}
        fields = {f["name"]: torch.empty((F2, h, w), dtype=tdt[f["elem_size"]], device=dev) for f in dec.fields}
        xyz = [torch.empty((F2, n, 3), dtype=torch.float32, device=dev) for _ in range(R)]
        rd = [torch.empty((F2, h, w), dtype=torch.int32, device=dev) for _ in range(R)]
        t_ts = torch.empty((F2, w), dtype=torch.int64, device=dev)
        t = timeit(lambda: dec.decode_batch(F2, t_pk, n_slots, psz, n_slots * psz, fields, lut=lut,
                                            pixel_shift_by_row=shifts, xyz=xyz, range_destaggered=rd,
                                            timestamp=t_ts, stream=st))
        opf = oracle_pf(prof, h, w)
        of = orc.Frame(opf, with_window=True)
        b = orc.Batcher(opf)
        t0 = time.perf_counter()
        for k, p in enumerate(pk):
            b.batch(p, 10 + k, of)
        for nm in (["RANGE", "RANGE2"][:R]):
            orc.destagger(of.field(nm), shifts)
            orc.cartesian(of.field(nm), d, o)
        cpu1 = time.perf_counter() - t0
        ok = all(np.array_equal(fields[f["name"]][0].cpu().numpy().view(fr.field(f["name"]).dtype),
                                fr.field(f["name"])) for f in dec.fields)
        rows.append({"kernel": "K2", "h": h, "w": w, "returns": R, "frames_per_launch": F2,
                     "bytes_per_frame": k2_bytes, "ms_per_launch": t * 1e3,
                     "mpoints_s": F2 * n * R / t / 1e6, "gbps": F2 * k2_bytes / t / 1e9,
                     "frac": F2 * k2_bytes / t / 1e9 / peak, "cpu_1thread_mpoints_s": n * R / cpu1 / 1e6,
                     "parity": bool(ok)})
        del t_pk, fields, xyz, rd

os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep_shapes.json", "w"), indent=0)
md = ["| kernel | HxW | returns | frames/launch | B/frame | ms/launch | Mpoints/s | GB/s (alg.) | frac of %.0f GB/s | CPU 1-thread Mpoints/s |" % peak,
      "|---|---|---|---|---|---|---|---|---|---|"]
for r in rows:
    md.append(f"| {r['kernel']} | {r['h']}x{r['w']} | {r['returns']} | {r['frames_per_launch']} | {r['bytes_per_frame']} | "
              f"{r['ms_per_launch']:.3f} | {r['mpoints_s']:.0f} | {r['gbps']:.0f} | {r['frac']:.2f} | {r['cpu_1thread_mpoints_s']:.0f} |")
open("gpurun_out/sweep_shapes.md", "w").write("\n".join(md) + "\n")
print("\n".join(md))
