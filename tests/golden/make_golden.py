#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ from the reference tree.

Run ONLY in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py

Produces, per reference pcap fixture (tests/pcaps/*.pcap of the reference):
  <name>.npz   lidar UDP payloads (uint8 [n, packet_size]) extracted with a ~30 line pcap parser
  <name>.json  the handful of metadata scalars/vectors the hot path needs + the reference's
               known answers: md5 field digests ("scans"[0] of *_digest.json, semantics in
               python/src/ouster/sdk/core/_digest.py:69-82) and the 64-bit snapshot hashes
               transcribed from tests/frame_batcher_test.cpp:553-592.
and xyz_reference.npz / destagger_reference.npz: outputs of the reference's OWN pure-python
restatements (python/src/ouster/sdk/examples/reference.py:18-76,134-163), imported UNMODIFIED
with a stub `ouster.sdk.core` module (the native bindings cannot be built here).
"""
import importlib.util
import json
import os
import struct
import sys
import types

import numpy as np

REF = "/root/reference"
PCAPS = os.path.join(REF, "tests", "pcaps")
OUT = os.path.dirname(os.path.abspath(__file__))

# tests/frame_batcher_test.cpp:553-592 (snapshot_param table)
SNAPSHOTS = {
    "OS-0-128-U1_v2.3.0_1024x10": {"RANGE": 0xf605c68634d4d496, "REFLECTIVITY": 0x308446ce12113b5c,
                                   "NEAR_IR": 0xacbe4e6963b1d6c7, "FLAGS": 6373750807750774351},
    "OS-0-32-U1_v2.2.0_1024x10": {"RANGE": 0xda815ba0ea0173dd, "RANGE2": 0x9d07c3e610c99239,
                                  "SIGNAL": 0xb2d846ac47621f7b, "SIGNAL2": 0x4553138a62c59e37,
                                  "REFLECTIVITY": 0x63d4c6e69ced4423,
                                  "REFLECTIVITY2": 0x415f5e481688fe5a,
                                  "NEAR_IR": 0x2c32a3e5be6b01d5, "FLAGS": 6902511898004997142,
                                  "FLAGS2": 14986456617710294519},
    "OS-1-128_767798045_1024x10_20230712_120049": {
        "RANGE": 0x8327b9d4c44c45a3, "RANGE2": 0x87288b444ddb9c9e,
        "REFLECTIVITY": 0x6912ca3fa04b0d1f, "REFLECTIVITY2": 0xf58aa5594d9749dc,
        "NEAR_IR": 0xc99384623c5d9feb, "FLAGS": 15585490641324286966,
        "FLAGS2": 3655442015794344596},
    "OS-2-128-U1_v2.3.0_1024x10": {"RANGE": 0x5940899c1190d02d, "SIGNAL": 0x4446bddd21f14dd4,
                                   "REFLECTIVITY": 0xea599b8814d2eac1,
                                   "NEAR_IR": 0x8a5a3df8896e317a, "FLAGS": 3655442015794344596},
    "OS-2-32-U0_v2.0.0_1024x10": {"RANGE": 0x5937f3d8f3762184, "SIGNAL": 0xbb4b7f22d1231e80,
                                  "REFLECTIVITY": 0x3D37AAEB2792F714,
                                  "NEAR_IR": 0xe972940ca8b204f0, "FLAGS": 13284364481018348283},
}

FIXTURES = [
    "OS-0-128-U1_v2.3.0_1024x10",                    # RNG15_RFL8_NIR8 (low data rate)
    "OS-0-32-U1_v2.2.0_1024x10",                     # RNG19_RFL8_SIG16_NIR16_DUAL
    "OS-1-128_767798045_1024x10_20230712_120049",    # FUSA dual low-bandwidth, negative shifts
    "OS-2-128-U1_v2.3.0_1024x10",                    # RNG19_RFL8_SIG16_NIR16 (single)
    "OS-2-32-U0_v2.0.0_1024x10",                     # LEGACY
    "OS-1-32-G_v2.1.1_1024x10",                      # LEGACY (digest only)
]


def read_pcap_udp(path, port=7502):
    """Classic pcap (magic a1b2c3d4, linktype 1), unfragmented IPv4/UDP."""
    data = open(path, "rb").read()
    magic = struct.unpack_from("<I", data, 0)[0]
    assert magic == 0xa1b2c3d4, hex(magic)
    assert struct.unpack_from("<I", data, 20)[0] == 1
    off, out = 24, []
    while off + 16 <= len(data):
        _, _, incl, _ = struct.unpack_from("<IIII", data, off)
        pkt = data[off + 16: off + 16 + incl]
        off += 16 + incl
        if len(pkt) < 42 or pkt[12:14] != b"\x08\x00" or pkt[23] != 17:
            continue
        ihl = (pkt[14] & 0xf) * 4
        frag = struct.unpack_from(">H", pkt, 20)[0] & 0x3fff
        assert frag == 0, "fragmented"
        udp = 14 + ihl
        dport, ulen = struct.unpack_from(">HH", pkt, udp + 2)
        if dport == port:
            out.append(pkt[udp + 8: udp + ulen])
    return out


def mat4(v):
    return [float(x) for x in v]


def load_meta(path):
    d = json.load(open(path))
    if "lidar_data_format" in d:   # nested (fw >= 2.4 style)
        fmt = d["lidar_data_format"]
        bi, li, si = d["beam_intrinsics"], d["lidar_intrinsics"], d["sensor_info"]
        az, alt = bi["beam_azimuth_angles"], bi["beam_altitude_angles"]
        n = bi.get("lidar_origin_to_beam_origin_mm")
        b2l = bi.get("beam_to_lidar_transform")
        l2s = li["lidar_to_sensor_transform"]
        init_id, sn, prod_line = si["initialization_id"], si["prod_sn"], si["prod_line"]
        fw = si.get("image_rev", "")
    else:                          # flat legacy layout
        fmt = d["data_format"]
        az, alt = d["beam_azimuth_angles"], d["beam_altitude_angles"]
        n = d.get("lidar_origin_to_beam_origin_mm")
        b2l = d.get("beam_to_lidar_transform")
        l2s = d["lidar_to_sensor_transform"]
        init_id, sn, prod_line = d.get("initialization_id", 0), d.get("prod_sn", "0"), d["prod_line"]
        fw = d.get("build_rev", "")
    if b2l is None:                # metadata.cpp:751-760: identity with (0,3) = origin offset
        b2l = [1, 0, 0, n, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
    profile = fmt.get("udp_profile_lidar", "LEGACY")
    header = fmt.get("header_type")
    if header is None:             # metadata.cpp:545-555
        header = "FUSA" if profile == "FUSA_RNG15_RFL8_NIR8_DUAL" else "STANDARD"
    return {
        "profile": profile, "header_type": header,
        "h": fmt["pixels_per_column"], "w": fmt["columns_per_frame"],
        "columns_per_packet": fmt["columns_per_packet"],
        "pixel_shift_by_row": fmt["pixel_shift_by_row"],
        "column_window": fmt.get("column_window", [0, fmt["columns_per_frame"] - 1]),
        "beam_azimuth_angles": az, "beam_altitude_angles": alt,
        "beam_to_lidar_transform": mat4(b2l), "lidar_to_sensor_transform": mat4(l2s),
        "init_id": init_id, "prod_sn": int(sn), "prod_line": prod_line, "fw": fw,
    }


def import_reference_py():
    """Import python/src/ouster/sdk/examples/reference.py unmodified, stubbing ouster.sdk.core."""
    core = types.ModuleType("ouster.sdk.core")

    class SensorInfo:  # only for the type annotations in reference.py
        pass

    class LidarFrame:
        pass

    class ChanField:
        RANGE = "RANGE"

    core.SensorInfo, core.LidarFrame, core.ChanField = SensorInfo, LidarFrame, ChanField
    ouster = types.ModuleType("ouster")
    sdk = types.ModuleType("ouster.sdk")
    ouster.sdk, sdk.core = sdk, core
    sys.modules.update({"ouster": ouster, "ouster.sdk": sdk, "ouster.sdk.core": core})
    spec = importlib.util.spec_from_file_location(
        "ouster_reference_examples",
        os.path.join(REF, "python", "src", "ouster", "sdk", "examples", "reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    for name in FIXTURES:
        pk = read_pcap_udp(os.path.join(PCAPS, name + ".pcap"))
        sizes = {len(p) for p in pk}
        assert len(sizes) == 1, sizes
        arr = np.frombuffer(b"".join(pk), np.uint8).reshape(len(pk), -1)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), packets=arr)
        meta = load_meta(os.path.join(PCAPS, name + ".json"))
        dg = os.path.join(PCAPS, name + "_digest.json")
        meta["md5_digests"] = json.load(open(dg))["scans"][0] if os.path.exists(dg) else None
        snap = SNAPSHOTS.get(name)
        meta["snapshot_hashes"] = {k: str(v) for k, v in snap.items()} if snap else None
        json.dump(meta, open(os.path.join(OUT, name + ".json"), "w"), indent=1)
        print(name, arr.shape, meta["profile"])

    # packets carrying a sensor-computed CRC64 (fw 3.2): pins crc64 (parsing.cpp:1183-1234)
    crc_pk = read_pcap_udp(os.path.join(PCAPS, "crc_test.pcap"))[:4]
    np.savez_compressed(os.path.join(OUT, "crc_test.npz"),
                        packets=np.frombuffer(b"".join(crc_pk), np.uint8).reshape(4, -1))

    # ---- outputs of the reference's own python restatements ----
    ref = import_reference_py()
    rng = np.random.default_rng(1234)
    cases = {}
    for fx in ("OS-0-32-U1_v2.2.0_1024x10", "OS-1-128_767798045_1024x10_20230712_120049"):
        m = json.load(open(os.path.join(OUT, fx + ".json")))
        h, w = m["h"], 64   # reference.py is a pure-python double loop: keep w small
        rimg = rng.integers(0, 1 << 19, size=(h, w), dtype=np.uint32)
        rimg[rng.random((h, w)) < 0.3] = 0

        meta = types.SimpleNamespace(
            beam_to_lidar_transform=np.array(m["beam_to_lidar_transform"]).reshape(4, 4),
            beam_azimuth_angles=m["beam_azimuth_angles"],
            beam_altitude_angles=m["beam_altitude_angles"],
            lidar_to_sensor_transform=np.array(m["lidar_to_sensor_transform"]).reshape(4, 4))
        frame = types.SimpleNamespace(w=w, h=h, field=lambda _n, _r=rimg: _r,
                                      measurement_id=np.arange(w))
        xyz = ref.xyz_proj_beam_to_sensor_transform(meta, frame)
        cases[fx + "/range"] = rimg
        cases[fx + "/xyz"] = xyz
    np.savez_compressed(os.path.join(OUT, "xyz_reference.npz"), **cases)

    dcases = {}
    for i, (h, w) in enumerate([(32, 512), (128, 256), (7, 33)]):
        img = rng.integers(0, 4096, size=(h, w)).astype(np.float64)
        shifts = rng.integers(-30, 31, size=h)
        dcases[f"c{i}/img"] = img
        dcases[f"c{i}/shifts"] = shifts.astype(np.int32)
        dcases[f"c{i}/out"] = ref.destagger(list(int(s) for s in shifts), img)
    np.savez_compressed(os.path.join(OUT, "destagger_reference.npz"), **dcases)
    print("reference.py vectors written")


def extract_profile_tables():
    """Parse the profile tables out of the reference SOURCE TEXT (ouster_core/src/parsing.cpp:170-356)
    into tests/golden/profile_tables.json: {profile: {chan_data_size, fields: {name: [bit, bits,
    upshift, num_elements]}}}.  Pins the oracle's and the product's transcriptions of every profile,
    including those without a pcap fixture."""
    import re
    src = open(os.path.join(REF, "ouster_core", "src", "parsing.cpp")).read()
    tables = {}
    for m in re.finditer(r"static const Table<std::string, FieldDecodeInfo, \d+> (\w+)\{\{(.*?)\}\};", src, re.S):
        name, body = m.group(1), m.group(2)
        fields = {}
        for fm in re.finditer(r"\{ChanField::(\w+),\s*field_info\(([^)]*)\)\}", body):
            args = [a.strip() for a in fm.group(2).split(",")]
            vals = [int(eval(re.sub(r"size_t\{(\d+)\}", r"\1", a))) for a in args]
            bit, bits = vals[0], vals[1]
            up = vals[2] if len(vals) > 2 else 0
            nel = vals[4] if len(vals) > 4 else 1
            fields[fm.group(1)] = [bit, bits, up, nel]
        tables[name] = fields
    out = {}
    for m in re.finditer(r"\{UDPProfileLidar::(\w+),\s*\{(\w+)\.data\(\), \w+\.size\(\), (\d+)\}\}", src):
        prof, tab, cds = m.group(1), m.group(2), int(m.group(3))
        out[prof] = {"chan_data_size": cds, "fields": tables.get(tab, {})}
    assert len(out) == 14, sorted(out)
    json.dump(out, open(os.path.join(OUT, "profile_tables.json"), "w"), indent=1, sort_keys=True)
    # default LidarFrame field slots (ouster_core/src/lidar_frame.cpp:73-226)
    src2 = open(os.path.join(REF, "ouster_core", "src", "lidar_frame.cpp")).read()
    slots = {}
    for m in re.finditer(r"static const Table<std::string, ChanFieldType, \d+> (\w+)\{\s*\{(.*?)\}\};", src2, re.S):
        slots[m.group(1)] = re.findall(r"\{ChanField::(\w+),\s*ChanFieldType::(\w+)\}", m.group(2))
    dflt = {}
    for m in re.finditer(r"\{UDPProfileLidar::(\w+),\s*\{(\w+)\.data\(\), \w+\.size\(\)\}\}", src2):
        dflt[m.group(1)] = slots.get(m.group(2), [])
    assert len(dflt) == 14, sorted(dflt)
    json.dump(dflt, open(os.path.join(OUT, "default_field_slots.json"), "w"), indent=1, sort_keys=True)
    print("profile tables:", len(out), "default slots:", len(dflt))


if __name__ == "__main__":
    main()
    extract_profile_tables()
