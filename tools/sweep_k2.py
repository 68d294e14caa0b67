#!/usr/bin/env python
"""In-process sweep of the K2 launch geometry (run under gpurun)."""
import itertools, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as graft
import bench, bench_k2 as k2
ob = graft.load_package()
F = int(os.environ.get("SWEEP_FRAMES", "32"))
H, W, R = k2.H, k2.W, k2.R
dev = torch.device("cuda", 0)
si, pk, src = k2.synth_packets(ob, 4)
n_slots, psz = pk.shape[1], pk.shape[2]
t_pk = torch.from_numpy(np.stack([pk[i % 4] for i in range(F)])).to(dev)
d, o = bench.synth_lut()
lut = ob.XYZLutT.from_arrays(torch.from_numpy(d).to(dev), torch.from_numpy(o).to(dev), H, W)
dec = ob.Decoder.from_sensor(si, src[0])
tdt = {1: torch.uint8, 2: torch.int16, 4: torch.int32}
fields = {f["name"]: torch.empty((F, H, W), dtype=tdt[f["elem_size"]], device=dev) for f in dec.fields}
xyz = [torch.empty((F, H * W, 3), dtype=torch.float32, device=dev) for _ in range(R)]
rd = [torch.empty((F, H, W), dtype=torch.int32, device=dev) for _ in range(R)]
st = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
peak, _ = bench.measured_peaks()
def step():
    dec.decode_batch(F, t_pk, n_slots, psz, n_slots * psz, fields, lut=lut, pixel_shift_by_row=k2.SHIFTS,
                     xyz=xyz, range_destaggered=rd, stream=st)
rows = []
PF = [int(x) for x in os.environ.get("SWEEP_PREFETCH", "2").split(",")]
GEOM = list(itertools.product([1, 2], [1, 2, 3], [192, 256, 320, 384], [1, 2, 3, 4]))
if os.environ.get("SWEEP_QUICK"):
    GEOM = [(2, 1, 384, 3), (2, 1, 320, 3), (1, 2, 384, 3), (2, 1, 256, 3)]
for (tp, stg, th, cta), pf in itertools.product(GEOM, PF):
    smem = 5120 + stg * 33152 * tp
    if smem * cta > 227 * 1024 or th * cta > 2048:
        continue
    for k, v in (("decode_tile_packets", tp), ("decode_stages", stg), ("decode_threads", th),
                 ("decode_ctas_per_sm", cta), ("decode_prefetch", pf)):
        ob.set_tunable(k, v)
    try:
        for _ in range(2):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        n = 5
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    except Exception as ex:
        print("fail", stg, th, cta, ex)
        continue
    gbps = k2.K2_BYTES_PER_FRAME_F32 * F / (ms * 1e-3) / 1e9
    rows.append({"tile_packets": tp, "stages": stg, "threads": th, "ctas": cta, "prefetch": pf, "ms": ms,
                 "gbps": gbps, "frac": gbps / peak})
rows.sort(key=lambda r: -r["gbps"])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep_k2.json", "w"), indent=0)
for r in rows[:10]:
    print(r)
print("worst", rows[-1])
