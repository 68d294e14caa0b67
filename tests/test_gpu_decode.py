"""GPU parity tests for K2 (fused packet decode -> fields + destagger + XYZ) through the C ABI,
bit-exact against the CPU oracle's FrameBatcher on the reference's pcap fixtures, on randomised
frames of every profile, and under the reference's fault-injection scenarios."""
import numpy as np
import pytest

import __graft_entry__ as graft
from oracle import oracle as orc
from tests.helpers import (PCAP_FIXTURES, col_map_from_packets, decoder_desc_from_oracle,
                           load_fixture, oracle_pf, random_frame, random_lut)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    graft.build()
    m = graft.load_package()
    assert m.device_count() > 0
    return m


def gpu_decode(ob, pf, ref_frame, packets, col_src, lut=None, shifts=None, device_inputs=False):
    layout, fields = decoder_desc_from_oracle(pf, ref_frame)
    dec = ob.Decoder(layout, fields)
    h, w = pf.pixels_per_column, pf.columns_per_frame
    outs = {f["name"]: np.full((h, w, 3) if f["name"] == "RGB" else (h, w), 0xAB,
                               ref_frame.field(f["name"]).dtype) for f in fields}
    io = {"packets": np.ascontiguousarray(packets), "n_slots": len(packets),
          "packet_stride": packets.shape[1], "col_src": col_src, "fields": outs,
          "timestamp": np.full(w, 7, np.uint64), "measurement_id": np.full(w, 7, np.uint16),
          "status": np.full(w, 7, np.uint32)}
    n_ret = 1 + int(any(f["range_return"] == 1 for f in fields))
    if lut is not None:
        io["xyz"] = [np.full((h * w, 3), 9, lut.dtype) for _ in range(n_ret)]
    if shifts is not None:
        io["range_destaggered"] = [np.full((h, w), 9, np.uint32) for _ in range(n_ret)]
    st = ob.Stream(0)
    dec.decode([io], lut=lut, pixel_shift_by_row=shifts, stream=st)
    st.sync()
    return io


def oracle_batch(pf, packets, meta=None, with_window=False, extra_ts=1234):
    frame = orc.Frame(pf, with_window=with_window)
    b = orc.Batcher(pf, init_id=(meta or {}).get("init_id", 0),
                    column_window=(meta or {}).get("column_window"))
    for p in packets:
        b.batch(p, extra_ts, frame)
    return frame


def check_frame(io, ref):
    for name, a in io["fields"].items():
        assert np.array_equal(a, ref.field(name)), name
    assert np.array_equal(io["timestamp"], ref.timestamp)
    assert np.array_equal(io["measurement_id"], ref.measurement_id)
    assert np.array_equal(io["status"], ref.status)


@pytest.mark.parametrize("name", PCAP_FIXTURES)
def test_pcap_fixtures_match_oracle_batcher(ob, name):
    meta, packets = load_fixture(name)
    pf = oracle_pf(meta)
    ref = oracle_batch(pf, packets, meta)
    col_src = col_map_from_packets(pf, packets)
    io = gpu_decode(ob, pf, ref, packets, col_src)
    check_frame(io, ref)
    if np.array_equal(col_src, np.arange(pf.columns_per_frame)):
        io2 = gpu_decode(ob, pf, ref, packets, None)  # identity map: whole-packet TMA path
        check_frame(io2, ref)


@pytest.mark.parametrize("name", ["OS-0-32-U1_v2.2.0_1024x10", "OS-1-128_767798045_1024x10_20230712_120049"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pcap_fused_xyz_and_destagger(ob, name, dtype):
    meta, packets = load_fixture(name)
    pf = oracle_pf(meta)
    ref = oracle_batch(pf, packets, meta)
    h, w = meta["h"], meta["w"]
    d, o = orc.make_xyz_lut(w, h, 0.001, meta["beam_to_lidar_transform"],
                            meta["lidar_to_sensor_transform"], meta["beam_azimuth_angles"],
                            meta["beam_altitude_angles"])
    d, o = d.astype(dtype), o.astype(dtype)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    shifts = np.array(meta["pixel_shift_by_row"], np.int32)
    io = gpu_decode(ob, pf, ref, packets, col_map_from_packets(pf, packets), lut=lut, shifts=shifts)
    check_frame(io, ref)
    for r, fname in enumerate(["RANGE", "RANGE2"]):
        assert np.array_equal(io["xyz"][r], orc.cartesian(ref.field(fname), d, o)), fname
        assert np.array_equal(io["range_destaggered"][r], orc.destagger(ref.field(fname), shifts)), fname


PROFILE_CASES = [
    ("RNG19_RFL8_SIG16_NIR16_DUAL", "STANDARD", 128, 2048), ("RNG19_RFL8_SIG16_NIR16_DUAL", "STANDARD", 32, 512),
    ("RNG19_RFL8_SIG16_NIR16", "STANDARD", 64, 1024), ("RNG19_RFL8_SIG16_NIR16", "STANDARD", 128, 1024),
    ("RNG15_RFL8_NIR8", "STANDARD", 32, 512), ("LEGACY", "STANDARD", 64, 1024),
    ("FIVE_WORD_PIXEL", "STANDARD", 32, 1024), ("FUSA_RNG15_RFL8_NIR8_DUAL", "FUSA", 128, 1024),
    ("RNG15_RFL8_NIR8_DUAL", "STANDARD", 64, 512), ("RNG15_RFL8_WIN8", "STANDARD", 32, 512),
    ("RNG19_RFL8_SIG16_NIR16_ZONE16", "STANDARD", 32, 512), ("RNG15_RFL8_NIR8_ZONE16", "STANDARD", 32, 512),
    ("RNG19_RFL8_SIG16_ZONE16_DUAL", "STANDARD", 32, 512), ("RNG19_RFL8_SIG16_NIR16_RGB16", "STANDARD", 32, 512),
    ("RNG19_RFL8_SIG16_NIR16_RGB16_DUAL", "STANDARD", 32, 512),
]


@pytest.fixture(params=["pipe_static", "pipe_runtime", "pipe_one_cta", "pipe_helpers", "pipe_tma_store", "static_layouts",
                        "runtime_plans", "narrow_tiles"])
def plans(ob, request):
    """K2 has two kernels (the pipelined one, ob_decode_pipe.cu, and decode_kernel for the shapes it
    does not take) and two phase-A code paths in each: compile-time pixel layouts for the standard
    profiles and runtime extraction plans for everything else.  The parity cases run through all of
    them, and through one-packet tiles (16 columns: a warp covers two rows per instruction, 2-stage
    ring) of decode_kernel."""
    pipe = request.param.startswith("pipe_")
    ob.set_tunable("decode_pipe", int(pipe))
    ob.set_tunable("decode_runtime_plans", int(request.param in ("runtime_plans", "pipe_runtime")))
    if request.param == "narrow_tiles":
        ob.set_tunable("decode_tile_packets", 1)
        ob.set_tunable("decode_stages", 2)
    if request.param == "pipe_one_cta":
        for k, v in (("decode_pipe_ctas", 1), ("decode_pipe_dyn_rows", 2), ("decode_pipe_lut_split", 4),
                     ("decode_pipe_pk_split", 8), ("decode_pipe_lane_arrive", 0)):
            ob.set_tunable(k, v)
    if request.param == "pipe_tma_store":   # XYZ through the store warp's tensor copies (batched / job launches)
        ob.set_tunable("decode_pipe_tma_xyz", 1)
    if request.param == "pipe_helpers":
        for k, v in (("decode_pipe_ctas", 1), ("decode_pipe_helpers", 5), ("decode_pipe_dyn_rows", 1)):
            ob.set_tunable(k, v)
    yield request.param
    for k, v in (("decode_pipe_ctas", 0), ("decode_pipe_helpers", 0), ("decode_pipe_dyn_rows", 3),
                 ("decode_pipe_lut_split", 1), ("decode_pipe_pk_split", 1), ("decode_pipe_lane_arrive", 1),
                 ("decode_pipe_tma_xyz", 0)):
        ob.set_tunable(k, v)
    ob.set_tunable("decode_pipe", 1)
    ob.set_tunable("decode_runtime_plans", 0)
    ob.set_tunable("decode_tile_packets", 0)
    ob.set_tunable("decode_stages", 1)


@pytest.mark.parametrize("profile,header,h,w", PROFILE_CASES)
def test_random_frame_roundtrip_all_profiles(ob, plans, profile, header, h, w):
    """frame -> frame_to_packets -> GPU decode == frame (tests/packet_format_test.cpp:218-326)."""
    pf = oracle_pf(profile, h, w, 16, header)
    src = random_frame(pf, seed=0xdeadbeef % (1 << 31))
    packets, ts = orc.frame_to_packets(src, pf, init_id=5, prod_sn=1234)
    assert len(packets) == w // 16
    d, o = random_lut(h * w, 3)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    shifts = np.random.default_rng(1).integers(-20, 21, h).astype(np.int32)
    io = gpu_decode(ob, pf, src, packets, None, lut=lut if src.has_field("RANGE") else None,
                    shifts=shifts if src.has_field("RANGE") else None)
    check_frame(io, src)
    if src.has_field("RANGE"):
        assert np.array_equal(io["xyz"][0], orc.cartesian(src.field("RANGE"), d, o))
        assert np.array_equal(io["range_destaggered"][0], orc.destagger(src.field("RANGE"), shifts))
    if src.has_field("RANGE2"):
        assert np.array_equal(io["xyz"][1], orc.cartesian(src.field("RANGE2"), d, o))


@pytest.mark.parametrize("keep", [("RANGE",), ("RANGE2", "SIGNAL", "WINDOW"), ("FLAGS", "NEAR_IR")])
def test_subset_of_fields_other_images_untouched(ob, plans, keep):
    """a frame that carries only some profile fields: those decode, nothing else is written"""
    pf = oracle_pf("RNG19_RFL8_SIG16_NIR16_DUAL", 64, 512)
    src = random_frame(pf, seed=21)
    packets, _ = orc.frame_to_packets(src, pf)
    layout, fields = decoder_desc_from_oracle(pf, src)
    fields = [f for f in fields if f["name"] in keep]
    dec = ob.Decoder(layout, fields)
    outs = {f["name"]: np.full((64, 512), 0xAB, src.field(f["name"]).dtype) for f in fields}
    d, o = random_lut(64 * 512, 4)
    lut = ob.XYZLutT.from_arrays(d, o, 64, 512)
    n_ret = 1 + int("RANGE2" in keep)
    io = {"packets": np.ascontiguousarray(packets), "n_slots": len(packets), "packet_stride": packets.shape[1],
          "col_src": None, "fields": outs}
    has_r0 = "RANGE" in keep
    if has_r0 or "RANGE2" in keep:
        io["xyz"] = [np.full((64 * 512, 3), 9, np.float32) if (r == 0 and has_r0) or (r == 1 and "RANGE2" in keep)
                     else None for r in range(n_ret)]
    st = ob.Stream(0)
    dec.decode([io], lut=lut, stream=st)
    st.sync()
    for name, a in outs.items():
        assert np.array_equal(a, src.field(name)), name
    for r, nm in enumerate(("RANGE", "RANGE2")[:n_ret]):
        if io.get("xyz") and io["xyz"][r] is not None:
            assert np.array_equal(io["xyz"][r], orc.cartesian(src.field(nm), d, o)), nm


@pytest.mark.parametrize("missing", [("NEAR_IR",), ("RANGE2", "FLAGS2"), ("SIGNAL", "SIGNAL2", "REFLECTIVITY")])
@pytest.mark.parametrize("with_rd", [True, False])
def test_full_decoder_frame_without_some_fields(ob, plans, missing, with_rd):
    """A decoder built for the whole profile, a frame that lacks some of its fields (LidarFrame without an
    optional channel): the no-null-test row loops (all outputs present, with or without the destaggered range)
    must not be chosen -- the fields that are there decode, the fused cloud and the destaggered range too."""
    h, w = 64, 512
    pf = oracle_pf("RNG19_RFL8_SIG16_NIR16_DUAL", h, w)
    src = random_frame(pf, seed=77)
    packets, _ = orc.frame_to_packets(src, pf)
    layout, fields = decoder_desc_from_oracle(pf, src)
    dec = ob.Decoder(layout, fields)
    outs = {f["name"]: np.full((h, w), 0xAB, src.field(f["name"]).dtype) for f in fields if f["name"] not in missing}
    d, o = random_lut(h * w, 6)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    shifts = np.random.default_rng(2).integers(0, 40, h).astype(np.int32)
    io = {"packets": np.ascontiguousarray(packets), "n_slots": len(packets), "packet_stride": packets.shape[1],
          "col_src": None, "fields": outs, "xyz": [np.full((h * w, 3), 9, np.float32) for _ in range(2)]}
    if with_rd:
        io["range_destaggered"] = [np.full((h, w), 9, np.uint32) for _ in range(2)]
    st = ob.Stream(0)
    dec.decode([io], lut=lut, pixel_shift_by_row=shifts if with_rd else None, stream=st)
    st.sync()
    for name, a in outs.items():
        assert np.array_equal(a, src.field(name)), name
    for r, nm in enumerate(("RANGE", "RANGE2")):
        assert np.array_equal(io["xyz"][r], orc.cartesian(src.field(nm), d, o)), nm
        if with_rd:
            assert np.array_equal(io["range_destaggered"][r], orc.destagger(src.field(nm), shifts)), nm


def test_fault_injection_matches_oracle_batcher(ob, plans):
    """dropped packet, invalidated columns, swapped packets, duplicate packet
    (tests/frame_batcher_test.cpp:119-170, 208-259)."""
    pf = oracle_pf("RNG19_RFL8_SIG16_NIR16_DUAL", 128, 1024)
    src = random_frame(pf, seed=77)
    packets, ts = orc.frame_to_packets(src, pf)
    pk = [p.copy() for p in packets]
    # invalidate every 4th column of packets 7 and 20 (clear status bit 0)
    import ctypes as C
    for slot in (7, 20):
        for c in range(0, 16, 4):
            base = pf.packet_header_size + c * pf.col_size
            col = np.concatenate([pk[slot][base:base + pf.col_size], np.zeros(8, np.uint8)])
            orc.lib().orc_field_set(C.byref(pf.c.col_status_info), col.ctypes.data, 0)
            pk[slot][base:base + pf.col_size] = col[:pf.col_size]
    pk[4], pk[5] = pk[5], pk[4]          # swapped
    del pk[14]                            # dropped
    pk.insert(30, pk[29].copy())          # duplicate
    arr = np.stack(pk)
    ref = orc.Frame(pf, with_window=True)
    b = orc.Batcher(pf)
    for p in arr[:-1]:
        assert not b.batch(p, 55, ref)
    b.batch(arr[-1], 55, ref)
    # the frame never completes by count (one packet short): finalize by feeding the next frame id
    nxt = random_frame(pf, seed=78, frame_id=701)
    p2, _ = orc.frame_to_packets(nxt, pf)
    for i in range(5):
        if b.batch(p2[i], 55, ref):
            break
    col_src = col_map_from_packets(pf, arr)
    io = gpu_decode(ob, pf, ref, arr, col_src)
    check_frame(io, ref)
    assert (col_src < 0).sum() == 16 + 8


def test_device_resident_batch_of_frames(ob):
    torch = pytest.importorskip("torch")
    pf = oracle_pf("RNG19_RFL8_SIG16_NIR16_DUAL", 128, 2048)
    F = 3
    srcs = [random_frame(pf, seed=100 + i, frame_id=700 + i) for i in range(F)]
    pk = np.stack([orc.frame_to_packets(s, pf)[0] for s in srcs])   # [F, 128, 33024]
    layout, fields = decoder_desc_from_oracle(pf, srcs[0])
    dec = ob.Decoder(layout, fields)
    h, w = 128, 2048
    d, o = random_lut(h * w, 8)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    shifts = np.tile(np.array([48, 32, 16, 0], np.int32), 32)
    t_pk = torch.from_numpy(pk).cuda()
    tdt = {1: torch.uint8, 2: torch.int16, 4: torch.int32}
    ios, keep = [], []
    for i in range(F):
        outs = {f["name"]: torch.empty((h, w), dtype=tdt[f["elem_size"]], device="cuda") for f in fields}
        xyz = [torch.empty((h * w, 3), dtype=torch.float32, device="cuda") for _ in range(2)]
        rd = [torch.empty((h, w), dtype=torch.int32, device="cuda") for _ in range(2)]
        ios.append({"packets": t_pk[i], "n_slots": 128, "packet_stride": pk.shape[2], "col_src": None,
                    "fields": outs, "xyz": xyz, "range_destaggered": rd})
    st = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
    dec.decode(ios, lut=lut, pixel_shift_by_row=shifts, stream=st)
    torch.cuda.synchronize()
    for i in range(F):
        for f in fields:
            got = ios[i]["fields"][f["name"]].cpu().numpy().view(srcs[i].field(f["name"]).dtype)
            assert np.array_equal(got, srcs[i].field(f["name"])), (i, f["name"])
        for r, nm in enumerate(["RANGE", "RANGE2"]):
            assert np.array_equal(ios[i]["xyz"][r].cpu().numpy(), orc.cartesian(srcs[i].field(nm), d, o))
            assert np.array_equal(ios[i]["range_destaggered"][r].cpu().numpy().view(np.uint32),
                                  orc.destagger(srcs[i].field(nm), shifts))


def test_custom_profile_widening_decode(ob):
    """custom profile entries may widen on the fly, e.g. {UINT32, 0, 0x7fff, -3}
    (tests/frame_batcher_test.cpp:652-692): decode with the alternative table == stock table."""
    name = "OS-0-128-U1_v2.3.0_1024x10"
    meta, packets = load_fixture(name)
    pf = oracle_pf(meta)
    ref = oracle_batch(pf, packets, meta)
    alt = oracle_pf(meta)
    alt.set_fields([("RANGE", orc.UINT32, 0, 0x7fff, -3), ("FLAGS", orc.UINT8, 1, 0b10000000, 7),
                    ("REFLECTIVITY", orc.UINT8, 1, 0xff00, 8), ("NEAR_IR", orc.UINT16, 2, 0xff00, 4)], 4)
    io = gpu_decode(ob, alt, ref, packets, col_map_from_packets(alt, packets))
    check_frame(io, ref)


def test_multi_stream_batch_with_per_frame_luts(ob):
    """Frames of different sensors (own LUT each) decoded by ONE launch (BASELINE configs[3])."""
    torch = pytest.importorskip("torch")
    pf = oracle_pf("RNG19_RFL8_SIG16_NIR16_DUAL", 32, 512)
    h, w, F = 32, 512, 4
    srcs = [random_frame(pf, seed=200 + i, frame_id=700 + i) for i in range(F)]
    pk = np.stack([orc.frame_to_packets(s, pf)[0] for s in srcs])
    layout, fields = decoder_desc_from_oracle(pf, srcs[0])
    dec = ob.Decoder(layout, fields)
    luts_host = [random_lut(h * w, 30 + i) for i in range(F)]
    luts = [ob.XYZLutT.from_arrays(d, o, h, w) for d, o in luts_host]
    shifts = np.arange(h, dtype=np.int32) % 7
    t_pk = torch.from_numpy(pk).cuda()
    tdt = {1: torch.uint8, 2: torch.int16, 4: torch.int32}
    outs = {f["name"]: torch.empty((F, h, w), dtype=tdt[f["elem_size"]], device="cuda") for f in fields}
    xyz = [torch.empty((F, h * w, 3), dtype=torch.float32, device="cuda") for _ in range(2)]
    rd = [torch.empty((F, h, w), dtype=torch.int32, device="cuda") for _ in range(2)]
    st = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
    dec.decode_batch(F, t_pk, pk.shape[1], pk.shape[2], pk.shape[1] * pk.shape[2], outs, lut=None,
                     pixel_shift_by_row=shifts, xyz=xyz, range_destaggered=rd, stream=st, frame_luts=luts)
    torch.cuda.synchronize()
    for i in range(F):
        d, o = luts_host[i]
        for f in fields:
            got = outs[f["name"]][i].cpu().numpy().view(srcs[i].field(f["name"]).dtype)
            assert np.array_equal(got, srcs[i].field(f["name"]))
        for r, nm in enumerate(["RANGE", "RANGE2"]):
            assert np.array_equal(xyz[r][i].cpu().numpy(), orc.cartesian(srcs[i].field(nm), d, o)), (i, nm)
            assert np.array_equal(rd[r][i].cpu().numpy().view(np.uint32), orc.destagger(srcs[i].field(nm), shifts))


def test_pipelined_kernel_takes_the_standard_shapes(ob):
    """The default configuration must run the pipelined K2 (not silently fall back to decode_kernel)
    for a standard profile with a fused cloud, f32 and f64 LUTs."""
    pf = oracle_pf("RNG19_RFL8_SIG16_NIR16_DUAL", 128, 1024, 16, "STANDARD")
    src = random_frame(pf, seed=11)
    packets, ts = orc.frame_to_packets(src, pf, init_id=5, prod_sn=1234)
    shifts = np.random.default_rng(2).integers(-40, 41, 128).astype(np.int32)
    for dt in (np.float32, np.float64):
        d, o = random_lut(128 * 1024, 5, dt)
        lut = ob.XYZLutT.from_arrays(d, o, 128, 1024)
        n0 = ob.kernel_launch_count("decode_pipe")
        io = gpu_decode(ob, pf, src, packets, None, lut=lut, shifts=shifts)
        assert ob.kernel_launch_count("decode_pipe") == n0 + 1
        check_frame(io, src)
        for r, nm in enumerate(("RANGE", "RANGE2")):
            assert np.array_equal(io["xyz"][r], orc.cartesian(src.field(nm), d, o))
            assert np.array_equal(io["range_destaggered"][r], orc.destagger(src.field(nm), shifts))


def test_lut_free_projection_in_fused_decode(ob):
    """K2 with a LUT in LUT-free mode: fields and destaggered ranges bit-exact, XYZ within the
    north_star tolerance (1e-5 norm-wise) of the oracle's float LUT path."""
    from tests.helpers import default_os1_64
    si = default_os1_64(1024)
    h, w = si["h"], si["w"]
    args = (w, h, 0.001, si["beam_to_lidar_transform"], si["lidar_to_sensor_transform"],
            si["beam_azimuth_angles"], si["beam_altitude_angles"])
    d, o = orc.make_xyz_lut(*args)
    d, o = d.astype(np.float32), o.astype(np.float32)
    lut = ob.XYZLutT.from_intrinsics(*args, dtype=np.float32).set_analytic(True)
    pf = oracle_pf("RNG19_RFL8_SIG16_NIR16_DUAL", h, w, 16, "STANDARD")
    src = random_frame(pf, seed=21)
    packets, ts = orc.frame_to_packets(src, pf, init_id=5, prod_sn=1234)
    n0 = ob.kernel_launch_count("decode_pipe")
    io = gpu_decode(ob, pf, src, packets, None, lut=lut, shifts=si["pixel_shift_by_row"])
    assert ob.kernel_launch_count("decode_pipe") == n0 + 1
    check_frame(io, src)
    for r, nm in enumerate(("RANGE", "RANGE2")):
        ref = orc.cartesian(src.field(nm), d, o)
        err = np.linalg.norm(io["xyz"][r].astype(np.float64) - ref, axis=-1)
        assert np.all(err <= 1e-5 * np.linalg.norm(ref.astype(np.float64), axis=-1) + 1e-7)
        assert np.all(io["xyz"][r][src.field(nm).reshape(-1) == 0] == 0.0)
        assert np.array_equal(io["range_destaggered"][r], orc.destagger(src.field(nm), si["pixel_shift_by_row"]))
