#!/usr/bin/env python
"""Which part of K2 is how far from the copy peak: the same 32-frame launch with subsets of the outputs
(all / fields only / fields + destaggered range / XYZ only / headers only), event-timed, with the bytes each
variant has to move (run under gpurun)."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as graft
import bench, bench_k2, bench_common as bc
ob = graft.load_package()
args = argparse.Namespace(k2_frames=int(os.environ.get("K2_FRAMES", "32")))
st = bench_k2.K2State(args, ob, torch, None, 0, 0, 1)
H, W, R, F = bench_k2.H, bench_k2.W, bench_k2.R, st.F
peak, _ = bc.measured_peaks()
lut = ob.XYZLutT.from_arrays(st.t_dir, st.t_off, H, W, device=0)
wire = st.n_slots * st.psz
fbytes = H * W * st.field_bytes_px
variants = {
    "all": dict(fields=st.fields, xyz=st.xyz, rd=st.rd, hdr=True),
    "fields_only": dict(fields=st.fields, xyz=None, rd=None, hdr=False),
    "fields_rd": dict(fields=st.fields, xyz=None, rd=st.rd, hdr=True),
    "xyz_only": dict(fields={}, xyz=st.xyz, rd=None, hdr=False),
    "xyz_rd": dict(fields={}, xyz=st.xyz, rd=st.rd, hdr=False),
    "none": dict(fields={}, xyz=None, rd=None, hdr=False),
    "headers_only": dict(fields={}, xyz=None, rd=None, hdr=True),
    "range_field_only": dict(fields={k: v for k, v in st.fields.items() if k in ("RANGE",)}, xyz=None, rd=None, hdr=False),
}
out = {}
configs = [tuple(int(x) for x in c.split(",")) for c in os.environ.get("K2_SPLITS", "1,1").split(";")]
only = os.environ.get("K2_VARIANTS")
dyns = [int(x) for x in os.environ.get("K2_DYN", "3").split(",")]
arrs = [int(x) for x in os.environ.get("K2_LANE_ARRIVE", "1").split(",")]
helps = [int(x) for x in os.environ.get("K2_HELPERS", "0").split(",")]
tmas = [int(x) for x in os.environ.get("K2_TMA_XYZ", "1").split(",")]
configs = [(a, b, d, la, h, tx) for (a, b) in configs for d in dyns for la in arrs for h in helps for tx in tmas]
for pk_split, lut_split, dyn, la, nh, tx in configs:
  ob.set_tunable("decode_pipe_tma_xyz", tx)
  ob.set_tunable("decode_pipe_helpers", nh)
  ob.set_tunable("decode_pipe_lane_arrive", la)
  ob.set_tunable("decode_pipe_pk_split", pk_split)
  ob.set_tunable("decode_pipe_lut_split", lut_split)
  ob.set_tunable("decode_pipe_dyn_rows", dyn)
  for name, v in variants.items():
    if only and name not in only.split(","):
        continue
    plan = st.dec.prepare_batch(F, st.t_pk, st.n_slots, st.psz, wire, v["fields"], lut=lut if v["xyz"] else None,
                                pixel_shift_by_row=bench_k2.SHIFTS, xyz=v["xyz"], range_destaggered=v["rd"],
                                timestamp=st.t_ts if v["hdr"] else None, measurement_id=st.t_mid if v["hdr"] else None,
                                status=st.t_st if v["hdr"] else None, stream=st.obs)
    for _ in range(5):
        plan()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st.stream)
    for _ in range(20):
        plan()
    e1.record(st.stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    fb = sum(int(np.prod(t.shape[1:])) * t.element_size() for t in v["fields"].values())
    b = F * (wire + fb + (R * H * W * 12 if v["xyz"] else 0) + (R * H * W * 4 if v["rd"] else 0)) + (H * W * 24 if v["xyz"] else 0)
    key = name if len(configs) == 1 else f"{name}@pk{pk_split},lut{lut_split},dyn{dyn},la{la},h{nh},tma{tx}"
    out[key] = {"ms": ms, "bytes": b, "gbs": b / ms / 1e6, "frac": b / ms / 1e6 / peak}
    print("%-28s ms %.4f  %.0f GB/s  frac %.3f" % (key, ms, b / ms / 1e6, b / ms / 1e6 / peak), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/k2_parts.json", "w"), indent=1)
