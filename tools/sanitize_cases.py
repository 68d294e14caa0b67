#!/usr/bin/env python
"""Small end-to-end invocations of every kernel, checked against the oracle; meant to run under
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_cases.py
(kept tiny: the sanitizer slows kernels down by orders of magnitude)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
from oracle import oracle as orc
from tests.helpers import (col_map_from_packets, decoder_desc_from_oracle, oracle_pf, random_frame,
                           random_lut, random_range)

ob = graft.load_package()
st = ob.Stream(0)
h, w = 16, 256
for dtype in (np.float32, np.float64):
    for R in (1, 2):
        for aligned in (True, False):
            rs = np.random.default_rng(3)
            rng = np.stack([np.stack([random_range(h, w, 10 * f + r) for r in range(R)]) for f in range(2)])
            d, o = random_lut(h * w, 1, dtype)
            sh = rs.integers(-20, 21, h).astype(np.int32)
            if aligned:
                sh = sh // 4 * 4
            lut = ob.XYZLutT.from_arrays(d, o, h, w)
            xyz = np.zeros((2, R, h * w, 3), dtype)
            rd = np.zeros((2, R, h, w), np.uint32)
            xd = np.zeros((2, R, h, w, 3), dtype)
            ob.scan_to_cloud(lut, sh, rng, xyz=xyz, range_destaggered=rd, xyz_destaggered=xd, stream=st)
            st.sync()
            for f in range(2):
                for r in range(R):
                    want = orc.cartesian(rng[f, r], d, o)
                    assert np.array_equal(xyz[f, r], want)
                    assert np.array_equal(rd[f, r], orc.destagger(rng[f, r], sh))
                    assert np.array_equal(xd[f, r], orc.destagger(want.reshape(h, w, 3), sh))
# K1 with fused per-column poses (streamed rows + resident pose planes) and K3 (count/scan/emit)
for dtype in (np.float32, np.float64):
    hh, ww = 40, 256
    rngp = np.stack([np.stack([random_range(hh, ww, 50 + 10 * f + r) for r in range(2)]) for f in range(2)])
    d, o = random_lut(hh * ww, 4, dtype)
    lut = ob.XYZLutT.from_arrays(d, o, hh, ww)
    poses = np.tile(np.eye(4, dtype=dtype), (ww, 1, 1))
    poses[:, :3, 3] = np.random.default_rng(2).random((ww, 3)).astype(dtype)
    poses[:, 0, 1] = 0.25
    sh = (np.arange(hh, dtype=np.int32) * 5) % 23
    xyz = np.zeros((2, 2, hh * ww, 3), dtype)
    xd = np.zeros((2, 2, hh, ww, 3), dtype)
    ob.scan_to_cloud(lut, sh, rngp, xyz=xyz, xyz_destaggered=xd, stream=st, poses=poses)
    st.sync()
    for f in range(2):
        for r in range(2):
            want = orc.dewarp(orc.cartesian(rngp[f, r], d, o).reshape(hh, ww, 3), poses)
            assert np.array_equal(xyz[f, r].reshape(hh, ww, 3), want)
            assert np.array_equal(xd[f, r], orc.destagger(want, sh))
    status = np.ones(ww, np.uint32)
    status[:5] = 0
    status[100] = 0
    ts = np.arange(ww, dtype=np.uint64)
    got = ob.dewarp_frame(lut, rngp[0, 0], poses.astype(np.float64), status, ts, 1.0, 300.0, provenance=True)
    want = orc.dewarp_frame(rngp[0, 0], d, o, poses.astype(np.float64), status, ts, 1.0, 300.0)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
print("K1 ok (plain, posed), K3 ok")
img = np.random.default_rng(1).integers(0, 255, (h, w), dtype=np.uint8)
sh = np.arange(h, dtype=np.int32) - 5
assert np.array_equal(ob.destagger(img, sh), orc.destagger(img, sh))
img64 = np.random.default_rng(1).random((9, 35))
assert np.array_equal(ob.destagger(img64, np.arange(9, dtype=np.int32)), orc.destagger(img64, np.arange(9)))
print("destagger ok")
ident = np.eye(4)
l = ob.XYZLutT.from_intrinsics(64, 8, 0.001, ident, ident, np.linspace(-3, 3, 8), np.linspace(-10, 10, 8))
assert np.all(np.isfinite(l.direction))
print("lut ok")
for profile, hh, ww in (("RNG19_RFL8_SIG16_NIR16_DUAL", 16, 128), ("LEGACY", 16, 64), ("RNG19_RFL8_SIG16_NIR16_RGB16", 8, 64)):
    pf = oracle_pf(profile, hh, ww)
    src = random_frame(pf, seed=4)
    pk, ts = orc.frame_to_packets(src, pf)
    layout, fields = decoder_desc_from_oracle(pf, src)
    dec = ob.Decoder(layout, fields)
    d, o = random_lut(hh * ww, 2)
    lut = ob.XYZLutT.from_arrays(d, o, hh, ww)
    shifts = (np.arange(hh, dtype=np.int32) * 3) % 17
    for cmap in (None, "faulty"):
        pkk = pk.copy()
        col_src = None
        ref = src
        if cmap:
            pkk = np.delete(pkk, 1, axis=0)
            col_src = col_map_from_packets(pf, pkk)
            ref = orc.Frame(pf, with_window=True)
            b = orc.Batcher(pf)
            for p in pkk:
                b.batch(p, 5, ref)
            b.batch(orc.frame_to_packets(random_frame(pf, 9, frame_id=701), pf)[0][0], 5, ref)  # finalize
        outs = {f["name"]: np.zeros(ref.field(f["name"]).shape, ref.field(f["name"]).dtype) for f in fields}
        io = {"packets": np.ascontiguousarray(pkk), "n_slots": len(pkk), "packet_stride": pkk.shape[1],
              "col_src": col_src, "fields": outs, "timestamp": np.zeros(ww, np.uint64)}
        has_r = ref.has_field("RANGE")
        if has_r:
            io["xyz"] = [np.zeros((hh * ww, 3), np.float32)]
            io["range_destaggered"] = [np.zeros((hh, ww), np.uint32)]
        dec.decode([io], lut=lut if has_r else None, pixel_shift_by_row=shifts if has_r else None, stream=st)
        st.sync()
        for n, a in outs.items():
            assert np.array_equal(a, ref.field(n)), (profile, cmap, n)
        if has_r:
            assert np.array_equal(io["xyz"][0], orc.cartesian(ref.field("RANGE"), d, o))
            assert np.array_equal(io["range_destaggered"][0], orc.destagger(ref.field("RANGE"), shifts))
print("K2 ok")

# K2 through the runtime-plan phase A as well (the loop above took the compile-time layouts where
# one exists), then the pipelined host path: decode jobs, zero-copy bursts, frames in flight
ob.set_tunable("decode_runtime_plans", 1)
pf = oracle_pf("RNG19_RFL8_SIG16_NIR16_DUAL", 16, 128)
src = random_frame(pf, seed=6)
pk, ts = orc.frame_to_packets(src, pf)
layout, fields = decoder_desc_from_oracle(pf, src)
dec = ob.Decoder(layout, fields)
outs = {f["name"]: np.zeros(src.field(f["name"]).shape, src.field(f["name"]).dtype) for f in fields}
dec.decode([{"packets": np.ascontiguousarray(pk), "n_slots": len(pk), "packet_stride": pk.shape[1],
             "col_src": None, "fields": outs}], stream=st)
st.sync()
for n, a in outs.items():
    assert np.array_equal(a, src.field(n)), n
ob.set_tunable("decode_runtime_plans", 0)
print("K2 runtime plans ok")

si = ob.SensorInfo("RNG19_RFL8_SIG16_NIR16_DUAL", 16, 128, fw_rev="v3.2.1")
d, o = random_lut(16 * 128, 5)
lut = ob.XYZLutT.from_arrays(d, o, 16, 128)
shifts = (np.arange(16, dtype=np.int32) * 3) % 17
pipe = ob.FramePipeline(si, depth=2, lut=lut, pixel_shift_by_row=shifts)
buf = ob.pinned_empty(pk.shape, np.uint8)
want = []
got = []
for k in range(4):
    f = random_frame(pf, seed=30 + k, frame_id=900 + k)
    p, t = orc.frame_to_packets(f, pf)
    want.append(f)
    buf[...] = p
    used, slot = pipe.push_burst(buf, t)
    assert used == len(p)
    if slot is not None:
        got.append((slot.frame.field("RANGE").copy(), slot.xyz[0].copy(), slot.range_destaggered[1].copy()))
while (slot := pipe.drain()) is not None:
    got.append((slot.frame.field("RANGE").copy(), slot.xyz[0].copy(), slot.range_destaggered[1].copy()))
assert len(got) == 4
for f, (r, x, rd2) in zip(want, got):
    assert np.array_equal(r, f.field("RANGE"))
    assert np.array_equal(x, orc.cartesian(f.field("RANGE"), d, o))
    assert np.array_equal(rd2, orc.destagger(f.field("RANGE2"), shifts))
print("pipeline ok")
# ---- round 2 kernels: pipelined K2 with both LUT dtypes and the LUT-free mode, K4 encode, normals, batched K3 ----
from tests.helpers import default_os1_64
from tests.test_oracle_normals import room_scene
pf = oracle_pf("RNG19_RFL8_SIG16_NIR16_DUAL", 64, 128)
src = random_frame(pf, seed=8)
pk, ts = orc.frame_to_packets(src, pf)
layout, fields = decoder_desc_from_oracle(pf, src)
dec = ob.Decoder(layout, fields)
shifts = (np.arange(64, dtype=np.int32) * 3) % 17
n0 = ob.kernel_launch_count("decode_pipe")
for dt in (np.float32, np.float64):
    d, o = random_lut(64 * 128, 7, dt)
    lut = ob.XYZLutT.from_arrays(d, o, 64, 128)
    io = {"packets": np.ascontiguousarray(pk), "n_slots": len(pk), "packet_stride": pk.shape[1], "col_src": None,
          "fields": {f["name"]: np.zeros(src.field(f["name"]).shape, src.field(f["name"]).dtype) for f in fields},
          "xyz": [np.zeros((64 * 128, 3), dt) for _ in range(2)],
          "range_destaggered": [np.zeros((64, 128), np.uint32) for _ in range(2)]}
    dec.decode([io], lut=lut, pixel_shift_by_row=shifts, stream=st)
    st.sync()
    for n, a in io["fields"].items():
        assert np.array_equal(a, src.field(n)), n
    for r, nm in enumerate(("RANGE", "RANGE2")):
        assert np.array_equal(io["xyz"][r], orc.cartesian(src.field(nm), d, o))
        assert np.array_equal(io["range_destaggered"][r], orc.destagger(src.field(nm), shifts))
assert ob.kernel_launch_count("decode_pipe") == n0 + 2, "the pipelined K2 did not run"
si64 = default_os1_64(1024)
al = ob.XYZLutT.from_intrinsics(128, 64, 0.001, si64["beam_to_lidar_transform"], si64["lidar_to_sensor_transform"],
                                si64["beam_azimuth_angles"], si64["beam_altitude_angles"], dtype=np.float32).set_analytic(True)
io = {"packets": np.ascontiguousarray(pk), "n_slots": len(pk), "packet_stride": pk.shape[1], "col_src": None,
      "fields": {}, "xyz": [np.zeros((64 * 128, 3), np.float32) for _ in range(2)]}
dec.decode([io], lut=al, pixel_shift_by_row=shifts, stream=st)
rngs = np.stack([np.stack([src.field("RANGE"), src.field("RANGE2")])])
xa = np.zeros((1, 2, 64 * 128, 3), np.float32)
ob.scan_to_cloud(al, shifts, rngs, xyz=xa, stream=st)
st.sync()
assert np.allclose(xa[0, 0], io["xyz"][0], rtol=1e-5, atol=1e-4)
print("K2 pipelined (f32, f64, LUT-free) ok")
# the pipelined kernel's other launch shapes: 3 CTAs per SM (32-row sensor), phase-A helper warps, store-warp tensor
# copies for XYZ (batched launch), each against the oracle
pf32 = oracle_pf("RNG19_RFL8_SIG16_NIR16_DUAL", 32, 128)
src32 = random_frame(pf32, seed=9)
pk32, _ = orc.frame_to_packets(src32, pf32)
layout32, fields32 = decoder_desc_from_oracle(pf32, src32)
dec32 = ob.Decoder(layout32, fields32)
sh32 = (np.arange(32, dtype=np.int32) * 5) % 23
d32, o32 = random_lut(32 * 128, 11, np.float32)
lut32 = ob.XYZLutT.from_arrays(d32, o32, 32, 128)
for tun in ({"decode_pipe_ctas": 3}, {"decode_pipe_ctas": 1, "decode_pipe_helpers": 5}, {"decode_pipe_ctas": 1, "decode_pipe_tma_xyz": 1}):
    for k, v in tun.items():
        ob.set_tunable(k, v)
    n1 = ob.kernel_launch_count("decode_pipe")
    io = {"packets": np.ascontiguousarray(pk32), "n_slots": len(pk32), "packet_stride": pk32.shape[1], "col_src": None,
          "fields": {f["name"]: np.zeros(src32.field(f["name"]).shape, src32.field(f["name"]).dtype) for f in fields32},
          "xyz": [np.zeros((32 * 128, 3), np.float32) for _ in range(2)],
          "range_destaggered": [np.zeros((32, 128), np.uint32) for _ in range(2)]}
    if "decode_pipe_tma_xyz" in tun:   # the store warp needs a uniformly strided batch: the batch entry point
        import torch
        dev0 = torch.device("cuda", 0)
        F = 3
        t_pk = torch.from_numpy(np.stack([pk32] * F)).to(dev0)
        tf = {f["name"]: torch.zeros((F, 32, 128), dtype={1: torch.uint8, 2: torch.int16, 4: torch.int32}[f["elem_size"]], device=dev0)
              for f in dec32.fields}
        tx = [torch.zeros((F, 32 * 128, 3), dtype=torch.float32, device=dev0) for _ in range(2)]
        trd = [torch.zeros((F, 32, 128), dtype=torch.int32, device=dev0) for _ in range(2)]
        dec32.decode_batch(F, t_pk, pk32.shape[0], pk32.shape[1], pk32.shape[0] * pk32.shape[1], tf, lut=lut32,
                           pixel_shift_by_row=sh32, xyz=tx, range_destaggered=trd, stream=st)
        st.sync()
        for r, nm in enumerate(("RANGE", "RANGE2")):
            assert np.array_equal(tx[r][F - 1].cpu().numpy(), orc.cartesian(src32.field(nm), d32, o32))
            assert np.array_equal(trd[r][F - 1].cpu().numpy().view(np.uint32), orc.destagger(src32.field(nm), sh32))
    else:
        dec32.decode([io], lut=lut32, pixel_shift_by_row=sh32, stream=st)
        st.sync()
        for n, a in io["fields"].items():
            assert np.array_equal(a, src32.field(n)), n
        for r, nm in enumerate(("RANGE", "RANGE2")):
            assert np.array_equal(io["xyz"][r], orc.cartesian(src32.field(nm), d32, o32))
            assert np.array_equal(io["range_destaggered"][r], orc.destagger(src32.field(nm), sh32))
    assert ob.kernel_launch_count("decode_pipe") > n1, "the pipelined K2 did not run"
    for k in tun:
        ob.set_tunable(k, 0)
print("K2 pipelined: 3 CTAs/SM, helper warps, store-warp XYZ ok")
sih = ob.SensorInfo("RNG19_RFL8_SIG16_NIR16_DUAL", 16, 128, fw_rev="v3.2.1")
fr = ob.LidarFrame(sih)
rs = np.random.default_rng(12)
for name in fr.fields:
    a = fr.field(name)
    a[...] = rs.integers(0, 200, size=a.shape).astype(a.dtype)
fr.status[:] = 1
fr.status[3] = 0
fr.timestamp[:] = 5 + np.arange(128)
fr.packet_timestamp[:] = 1 + np.arange(8)
fr.frame_id = 77
hp, _ = ob.frame_to_packets(fr, sih, 3, 9)
dp, _ = ob.frame_to_packets(fr, sih, 3, 9, device=True)
assert np.array_equal(hp, dp)
print("K4 encode ok")
xyz, rngd, dirs = room_scene(24, 96)
org = np.zeros((96, 3))
nn, sub = ob.normals(xyz, rngd, org, 2, return_subtent=True)
assert np.array_equal(nn, orc.normals(xyz, rngd, sensor_origins_xyz=org, pixel_search_range=2, vertical_subtent=sub))
n1, n2 = ob.normals(xyz, rngd, xyz * 1.1, rngd, org)
print("normals ok")
frames = []
want = []
for i, (hh, ww) in enumerate(((16, 64), (24, 96))):
    rg = random_range(hh, ww, 80 + i, p_zero=0.3, max_range=50000)
    d, o = random_lut(hh * ww, 9 + i)
    poses = np.tile(np.eye(4), (ww, 1, 1))
    poses[:, :3, 3] = np.random.default_rng(i).random((ww, 3))
    stt = np.ones(ww, np.uint32)
    stt[:3] = 0
    tsn = np.arange(ww, dtype=np.uint64)
    frames.append({"lut": ob.XYZLutT.from_arrays(d, o, hh, ww), "range": rg, "poses": poses, "status": stt, "timestamps": tsn})
    want.append(orc.dewarp_frame(rg, d, o, poses, stt, tsn, 0.5, 40.0)[0])
assert np.array_equal(ob.dewarp_frames(frames, 0.5, 40.0), np.concatenate(want))
print("batched K3 ok")
print("SANITIZE CASES OK")
