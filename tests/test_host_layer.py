"""CPU tests of the product's host-side mirror (PacketFormat tables, frame_to_packets, header
getters, FrameBatcher state machine in header-only mode) against the reference's own tables
(tests/golden/profile_tables.json, parsed from ouster_core/src/parsing.cpp by make_golden.py)
and against the CPU oracle."""
import json
import os

import numpy as np
import pytest

import __graft_entry__ as graft
from oracle import oracle as orc
from tests.helpers import GOLDEN, PCAP_FIXTURES, load_fixture, oracle_pf, random_frame

PROFILES = json.load(open(os.path.join(GOLDEN, "profile_tables.json")))
SLOTS = json.load(open(os.path.join(GOLDEN, "default_field_slots.json")))
TAGS = {"UINT8": 1, "UINT16": 2, "UINT32": 3, "UINT64": 4, "FLOAT16": 12}


@pytest.fixture(scope="module")
def ob():
    graft.build()
    return graft.load_package()


def expected_info(bit, bits, up, nel):
    """field_info() of the reference (parsing.cpp:57-122) evaluated in python."""
    offset, lsb = bit // 8, bit % 8
    mask = sum(1 << i for i in range(lsb, lsb + bits))
    shift = lsb - up
    nbytes = ((bits + up + 7) // 8) // nel
    tag = {1: 1, 2: 2, 3: 3, 4: 3, 5: 4, 6: 4, 7: 4, 8: 4}.get(nbytes, 0)
    return offset, mask, shift, tag, nel


@pytest.mark.parametrize("profile", sorted(p for p in PROFILES if p != "OFF"))
def test_profile_tables_match_reference_source(ob, profile):
    spec = PROFILES[profile]
    si = ob.SensorInfo(profile, 32, 512)
    L = si.layout
    legacy = profile == "LEGACY"
    assert L.channel_data_size == spec["chan_data_size"]
    assert L.col_size == (16 if legacy else 12) + 32 * spec["chan_data_size"] + (4 if legacy else 0)
    assert L.packet_size == (0 if legacy else 64) + 16 * L.col_size
    got = {f[0]: f for f in si.fields()}
    assert sorted(got) == sorted(spec["fields"])
    assert [f[0] for f in si.fields()] == sorted(spec["fields"])      # std::map iteration order
    opf = orc.PacketFormat(profile, 32, 512)
    assert opf.lidar_packet_size == L.packet_size and opf.field_names == sorted(spec["fields"])
    for name, (bit, bits, up, nel) in spec["fields"].items():
        off, mask, shift, tag, n = expected_info(bit, bits, up, nel)
        assert got[name][1:6] == (tag, off, mask, shift, n), name
        oi = opf.field_info(name)
        assert (oi.ty_tag, oi.offset, oi.mask, oi.shift, oi.num_elements) == (tag, off, mask, shift, n), name
        assert got[name][6] == opf.value_mask(name)


@pytest.mark.parametrize("profile", sorted(p for p in SLOTS if p != "OFF"))
def test_default_frame_fields_match_reference_source(ob, profile):
    want = {n: TAGS[t] for n, t in SLOTS[profile]}
    # WINDOW only from firmware 3.2.0 (3.2.1 for zone profiles), lidar_frame.cpp:1097-1110
    for fw, has_window in (("v3.2.1", True), ("v2.5.0", False), ("UNKNOWN", False)):
        si = ob.SensorInfo(profile, 16, 64, fw_rev=fw)
        fr = ob.LidarFrame(si)
        exp = {n: t for n, t in want.items() if has_window or n != "WINDOW"}
        assert sorted(fr.fields) == sorted(exp)
        for n, t in exp.items():
            a = fr.field(n)
            assert a.dtype == ob.host.TAG_NP[t] if hasattr(ob, "host") else True
            assert a.shape[:2] == (16, 64)
        of = orc.Frame(orc.PacketFormat(profile, 16, 64), with_window=has_window)
        assert sorted(of.field_names) == sorted(exp)
    zone = ob.SensorInfo(profile, 16, 64, fw_rev="v3.2.0")
    has = "WINDOW" in ob.LidarFrame(zone).fields
    assert has == ("WINDOW" in want and "ZONE16" not in profile.replace("ZONE16_DUAL", "x"))


def test_packet_geometry_headline_config(ob):
    si = ob.SensorInfo("RNG19_RFL8_SIG16_NIR16_DUAL", 128, 2048)
    L = si.layout
    assert (L.col_size, L.packet_size) == (2060, 33024)
    assert si.block_parsable() == 16
    si = ob.SensorInfo("RNG19_RFL8_SIG16_NIR16", 128, 2048)
    assert si.layout.packet_size == 24832
    assert ob.SensorInfo("LEGACY", 64, 1024).layout.packet_size == 12608


@pytest.mark.parametrize("name", PCAP_FIXTURES)
def test_header_getters_on_pcap_fixtures(ob, name):
    meta, packets = load_fixture(name)
    si = ob.SensorInfo.from_meta(meta)
    assert si.lidar_packet_size == packets.shape[1]
    opf = oracle_pf(meta)
    for p in packets[:4]:
        assert si.frame_id(p) == opf.frame_id(p)
        if meta["profile"] != "LEGACY":
            assert si.packet_init_id(p) == meta["init_id"]
            assert si.packet_prod_sn(p) == meta["prod_sn"]
    if meta["md5_digests"]:
        assert si.frame_id(packets[0]) == int(meta["md5_digests"]["FRAME_ID"])
    for p in packets[:2]:
        assert si.calculate_crc(p) == orc.crc64(p[:-8])


def test_crc64_against_sensor_computed_crc(ob):
    # fw 3.2 packets of the reference's crc_test.pcap end with the sensor's own CRC64
    pk = np.load(os.path.join(GOLDEN, "crc_test.npz"))["packets"]
    si = ob.SensorInfo("RNG15_RFL8_NIR8", 128, 1024)
    assert si.lidar_packet_size == pk.shape[1]
    for p in pk:
        want = int(p[-8:].view(np.uint64)[0])
        assert orc.crc64(p[:-8]) == want
        assert si.calculate_crc(p) == want


def test_frame_id_difference_wraparound(ob):
    # tests/packet_format_test.cpp:776-820
    si = ob.SensorInfo("RNG19_RFL8_SIG16_NIR16", 32, 512)
    assert si.frame_id_difference(65535, 0) == 1
    assert si.frame_id_difference(0, 65535) == -1
    assert si.frame_id_difference(10, 20) == 10
    assert si.frame_id_difference(0, 32768) == -32768
    fu = ob.SensorInfo("FUSA_RNG15_RFL8_NIR8_DUAL", 32, 512, header_type="FUSA")
    assert fu.frame_id_difference(0xffffffff, 0) == 1
    assert fu.frame_id_difference(0, 0xffffffff) == -1
    opf = orc.PacketFormat("RNG19_RFL8_SIG16_NIR16", 32, 512)
    for a, b in [(65535, 0), (0, 65535), (100, 40000), (40000, 100), (7, 7)]:
        assert si.frame_id_difference(a, b) == opf.frame_id_difference(a, b)


def _fill_from_oracle(fr, src):
    for n in src.field_names:
        fr.field(n)[...] = src.field(n)
    fr.timestamp[:] = src.timestamp
    fr.measurement_id[:] = src.measurement_id
    fr.status[:] = src.status
    fr.packet_timestamp[:] = src.packet_timestamp
    fr.alert_flags[:] = src.alert_flags
    fr.frame_id = src.frame_id


CASES = [("RNG19_RFL8_SIG16_NIR16_DUAL", "STANDARD", 32, 512), ("RNG19_RFL8_SIG16_NIR16", "STANDARD", 64, 1024),
         ("RNG15_RFL8_NIR8", "STANDARD", 32, 512), ("LEGACY", "STANDARD", 64, 512),
         ("FIVE_WORD_PIXEL", "STANDARD", 16, 256), ("FUSA_RNG15_RFL8_NIR8_DUAL", "FUSA", 32, 512),
         ("RNG19_RFL8_SIG16_NIR16_RGB16_DUAL", "STANDARD", 16, 256), ("RNG19_RFL8_SIG16_ZONE16_DUAL", "STANDARD", 16, 256)]


@pytest.mark.parametrize("profile,header,h,w", CASES)
def test_frame_to_packets_bytes_match_oracle(ob, profile, header, h, w):
    opf = oracle_pf(profile, h, w, 16, header)
    src = random_frame(opf, seed=11)
    want, want_ts = orc.frame_to_packets(src, opf, init_id=0x123456, prod_sn=0x9876543210)
    si = ob.SensorInfo(profile, h, w, header_type=header, fw_rev="v3.2.1")
    fr = ob.LidarFrame(si)
    assert sorted(fr.fields) == sorted(src.field_names)
    _fill_from_oracle(fr, src)
    got, got_ts = ob.frame_to_packets(fr, si, init_id=0x123456, prod_sn=0x9876543210)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert np.array_equal(got_ts, want_ts)


def _drive(ob, si, opf, meta_init, packets, ts, with_window=True, cache=None):
    """Feed the same packet sequence to the product batcher (header-only) and the oracle."""
    b = ob.FrameBatcher(si)
    b.set_headers_only(True)
    fr = ob.LidarFrame(si)
    oframe = orc.Frame(opf, with_window=with_window)
    obat = orc.Batcher(opf, init_id=meta_init)
    if cache:
        b.set_max_cache_size(cache)
        obat.set_max_cache_size(cache)
    for i, (p, t) in enumerate(zip(packets, ts)):
        r1 = b.batch(p, int(t), fr)
        r2 = obat.batch(p, int(t), oframe)
        assert r1 == r2, i
        assert fr.frame_id == oframe.frame_id, i
        assert np.array_equal(fr.timestamp, oframe.timestamp), i
        assert np.array_equal(fr.measurement_id, oframe.measurement_id), i
        assert np.array_equal(fr.status, oframe.status), i
        assert np.array_equal(fr.packet_timestamp, oframe.packet_timestamp), i
        assert np.array_equal(fr.alert_flags, oframe.alert_flags), i
        assert b.batched_packets == obat.batched_packets, i
        assert b.dropped_packets == obat.dropped_packets, i
        assert fr.status_tuple() == (oframe.c.frame_status, oframe.c.shutdown_countdown,
                                     oframe.c.shot_limiting_countdown)
    return b, fr


def _stream(opf, n_frames, first_id=700, seed=5, init_id=0):
    pk, ts = [], []
    for k in range(n_frames):
        f = random_frame(opf, seed=seed + k, frame_id=(first_id + k) & opf.max_frame_id)
        p, t = orc.frame_to_packets(f, opf, init_id=init_id, prod_sn=77)
        pk += list(p)
        ts += list(t)
    return pk, ts


def test_batcher_state_machine_in_order_and_faults(ob):
    """Return values, headers and counters follow the reference state machine
    (tests/frame_batcher_test.cpp:73-303, 771-1399) -- compared step by step with the oracle."""
    profile, h, w = "RNG19_RFL8_SIG16_NIR16_DUAL", 16, 256
    opf = oracle_pf(profile, h, w)
    si = ob.SensorInfo(profile, h, w, fw_rev="v3.2.1")
    pk, ts = _stream(opf, 4)
    _drive(ob, si, opf, 0, pk, ts)
    # dropped packet + swapped packets + duplicate + a late packet of the previous frame
    pk2, ts2 = list(pk), list(ts)
    pk2[3], pk2[4] = pk2[4], pk2[3]
    ts2[3], ts2[4] = ts2[4], ts2[3]
    del pk2[9], ts2[9]
    pk2.insert(20, pk2[19].copy()); ts2.insert(20, ts2[19])
    late = pk[2].copy()
    pk2.insert(40, late); ts2.insert(40, 999)
    _drive(ob, si, opf, 0, pk2, ts2)
    # host timestamp 0 never completes a frame by count (SURVEY 8a'-1)
    _drive(ob, si, opf, 0, pk, [0] * len(pk))
    # interleaved frames exercise the cache, incl. a small cache limit
    order = []
    n = w // 16
    for k in range(3):
        a = list(range(k * n, (k + 1) * n))
        order += a[: n - 3]
        if k < 3:
            order += list(range((k + 1) * n, (k + 1) * n + 3)) if (k + 1) * n + 3 <= len(pk) else []
        order += a[n - 3:]
    order = [i for i in order if i < len(pk)]
    _drive(ob, si, opf, 0, [pk[i] for i in order], [ts[i] for i in order])
    _drive(ob, si, opf, 0, [pk[i] for i in order], [ts[i] for i in order], cache=1)


def test_batcher_frame_id_wraparound_and_init_id_change(ob):
    profile, h, w = "RNG19_RFL8_SIG16_NIR16", 16, 256
    opf = oracle_pf(profile, h, w)
    si = ob.SensorInfo(profile, h, w, fw_rev="v3.2.1", init_id=5)
    pk, ts = _stream(opf, 4, first_id=65534, init_id=5)
    _drive(ob, si, opf, 5, pk, ts)
    # init id changes mid-stream: current frame is released, new stream continues
    pk_b, ts_b = _stream(opf, 2, first_id=10, seed=50, init_id=9)
    _drive(ob, si, opf, 5, pk[: 16 + 5] + pk_b, ts[: 16 + 5] + ts_b)
    # init id differs from the very first packet
    _drive(ob, si, opf, 5, pk_b, ts_b)


def test_batcher_invalid_columns_and_out_of_range_ids(ob):
    import ctypes as C
    profile, h, w = "RNG15_RFL8_NIR8", 16, 256
    opf = oracle_pf(profile, h, w)
    si = ob.SensorInfo(profile, h, w)
    pk, ts = _stream(opf, 2)
    pk = [p.copy() for p in pk]
    lib = orc.lib()
    def poke(slot, c, info, val):
        base = opf.packet_header_size + c * opf.col_size
        col = np.concatenate([pk[slot][base: base + opf.col_size], np.zeros(8, np.uint8)])
        lib.orc_field_set(C.byref(info), col.ctypes.data, val)
        pk[slot][base: base + opf.col_size] = col[: opf.col_size]
    for c in (0, 5, 15):
        poke(3, c, opf.c.col_status_info, 0)            # invalid columns -> column path
    poke(6, 2, opf.c.col_measurement_id_info, 4000)     # measurement id beyond the frame width
    poke(7, 0, opf.c.col_status_info, 0)
    _drive(ob, si, opf, 0, pk, ts, with_window=False)


def test_batcher_argument_errors(ob):
    si = ob.SensorInfo("RNG19_RFL8_SIG16_NIR16", 16, 256)
    other = ob.SensorInfo("RNG19_RFL8_SIG16_NIR16", 16, 512)
    b = ob.FrameBatcher(si)
    b.set_headers_only(True)
    with pytest.raises(ValueError, match="unexpected frame dimensions"):
        b.batch(np.zeros(si.lidar_packet_size, np.uint8), 1, ob.LidarFrame(other))
    with pytest.raises(ValueError, match="max_cache_size must be > 0"):
        b.set_max_cache_size(0)
    with pytest.raises(ValueError, match="Unknown lidar udp profile"):
        ob.SensorInfo("NOT_A_PROFILE", 16, 256)


def test_cpp_headers_host_only_example():
    """tests/cpp/host_only_example.cpp: the replacement headers' host-only surface (PacketFormat
    geometry / CRC, LidarFrame fields + column poses, frame_to_packets, the FrameBatcher state
    machine in header-only mode) compiled with plain g++ and run on the CPU."""
    import subprocess
    import __graft_entry__ as graft
    graft.build()
    root = graft.ROOT
    lib_dir = os.path.join(root, "ouster-sdk_b200", "lib")
    exe = os.path.join(root, "tests", "cpp", "host_only_example.bin")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "host_only_example.cpp"), "-L", lib_dir,
                           "-louster_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "HOST OK" in out.stdout


def test_python_frame_poses_and_valid_columns(ob):
    """LidarScan.body_to_world / get_first|last_valid_column through the host C ABI (no GPU needed)."""
    si = ob.SensorInfo("RNG19_RFL8_SIG16_NIR16", 16, 64)
    fr = ob.LidarFrame(si)
    assert fr.body_to_world.shape == (64, 4, 4) and np.array_equal(fr.body_to_world[5], np.eye(4))
    fr.body_to_world[5, 0, 3] = 2.5
    fr2 = ob.LidarFrame(si)
    assert fr2.body_to_world[5, 0, 3] == 0.0 and fr.pose[5, 0, 3] == 2.5
    with pytest.raises(RuntimeError, match="No valid columns in LidarFrame"):
        fr.get_first_valid_column()
    fr.status[3] = 1
    fr.status[40] = 3
    fr.status[50] = 2
    assert fr.get_first_valid_column() == 3 and fr.get_last_valid_column() == 40


@pytest.mark.parametrize("burst", [1, 5, 16, 37])
def test_batch_burst_consumption_matches_per_packet_semantics(ob, burst):
    """FrameBatcher.batch_burst (host logic, header-only, no GPU): a burst stops after the packet
    that completes a frame and reports how many packets it took; fed burst by burst, the frames,
    their headers and the drop counter equal the oracle batcher fed packet by packet."""
    profile, h, w = "RNG19_RFL8_SIG16_NIR16_DUAL", 16, 256
    opf = oracle_pf(profile, h, w)
    si = ob.SensorInfo(profile, h, w, fw_rev="v3.2.1")
    pk, ts = _stream(opf, 5, seed=31)
    pk, ts = [p.copy() for p in pk], list(ts)
    pk[3], pk[4] = pk[4], pk[3]
    ts[3], ts[4] = ts[4], ts[3]
    del pk[20], ts[20]
    pk.insert(40, pk[39].copy()); ts.insert(40, ts[39])
    arr, tsa = np.stack(pk), np.asarray(ts, np.uint64)
    # oracle, packet by packet: snapshot the headers of every completed frame
    oframe, obat, want = orc.Frame(opf, with_window=True), orc.Batcher(opf), []
    for p, t in zip(pk, ts):
        if obat.batch(p, int(t), oframe):
            want.append((oframe.frame_id, oframe.timestamp.copy(), oframe.status.copy(),
                         oframe.packet_timestamp.copy()))
    b = ob.FrameBatcher(si)
    b.set_headers_only(True)
    fr, got, pos = ob.LidarFrame(si), [], 0
    while pos < len(pk):
        end = min(pos + burst, len(pk))
        while pos < end:
            used, done = b.batch_burst(arr[pos:end], tsa[pos:end], fr)
            assert 1 <= used <= end - pos
            assert done or used == end - pos      # an unfinished burst is consumed completely
            pos += used
            if done:
                got.append((fr.frame_id, fr.timestamp.copy(), fr.status.copy(), fr.packet_timestamp.copy()))
    assert len(got) == len(want) >= 4
    for g, wv in zip(got, want):
        assert g[0] == wv[0]
        for a, c in zip(g[1:], wv[1:]):
            assert np.array_equal(a, c)
    assert b.dropped_packets == obat.dropped_packets
