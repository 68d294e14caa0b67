"""PcapLidarSource (include/ouster/core/pcap_source.h): capture file -> page-locked ring of lidar packets.
CPU: the parser on capture files written here (both byte orders, micro/nanosecond magic, VLAN tag, other
ports and sizes, fragments, truncated tail).  GPU: a capture of a reference fixture's packets read in bursts
and fed to FrameBatcher.batch_burst in place reproduces the reference's md5 digests."""
import hashlib
import struct

import numpy as np
import pytest

import __graft_entry__ as graft
from tests.helpers import PCAP_FIXTURES, load_fixture


def _udp_frame(payload, dport=7502, vlan=False, frag=0, proto=17, ip_options=0):
    udp = struct.pack(">HHHH", 40000, dport, 8 + len(payload), 0) + bytes(payload)
    ihl = 5 + ip_options
    ip = struct.pack(">BBHHHBBH4s4s", 0x40 | ihl, 0, ihl * 4 + len(udp), 1, frag, 64, proto, 0,
                     bytes([10, 0, 0, 1]), bytes([10, 0, 0, 2])) + b"\x00" * (4 * ip_options)
    eth = b"\x02" * 6 + b"\x04" * 6 + (b"\x81\x00\x00\x05" if vlan else b"") + b"\x08\x00"
    return eth + ip + udp


def write_pcap(path, records, swapped=False, nanos=False, linktype=1):
    """records: (sec, frac, frame bytes)."""
    e = ">" if swapped else "<"
    magic = 0xa1b23c4d if nanos else 0xa1b2c3d4
    with open(path, "wb") as f:
        f.write(struct.pack(e + "IHHiIII", magic, 2, 4, 0, 0, 65535, linktype))
        for sec, frac, fr in records:
            f.write(struct.pack(e + "IIII", sec, frac, len(fr), len(fr)) + fr)


@pytest.fixture(scope="module")
def ob():
    graft.build()
    return graft.load_package()


@pytest.mark.parametrize("swapped,nanos", [(False, False), (True, False), (False, True), (True, True)])
def test_parser_filters_and_timestamps(ob, tmp_path, swapped, nanos):
    rng = np.random.default_rng(3)
    size = 1040
    good = [rng.integers(0, 256, size, dtype=np.uint8) for _ in range(7)]
    rec = [
        (10, 5, _udp_frame(good[0])),
        (10, 6, _udp_frame(rng.integers(0, 256, 48, dtype=np.uint8), dport=7503)),   # an IMU-sized datagram
        (11, 7, _udp_frame(good[1], vlan=True)),
        (11, 8, _udp_frame(good[2], frag=0x2000)),                                     # first fragment: skipped
        (12, 9, _udp_frame(good[2], ip_options=2)),
        (12, 10, _udp_frame(good[3], proto=6)),                                         # TCP: skipped
        (13, 11, _udp_frame(good[3], dport=9999)),                                      # other port
        (13, 12, b"\x00" * 10),                                                        # runt
        (14, 13, _udp_frame(good[4])),
    ]
    p = tmp_path / "a.pcap"
    write_pcap(p, rec, swapped, nanos)
    with open(p, "ab") as f:      # truncated record at the tail: end of file, not an error
        f.write(b"\x01\x02\x03")
    src = ob.PcapLidarSource(p, size, dst_port=7502, ring_packets=3)
    got, ts = [], []
    while True:
        pk, t = src.next_burst(8)                 # clipped to the ring (3)
        if len(t) == 0:
            break
        assert pk.shape[0] <= 3 and pk.shape[1] == size and pk.strides[0] % 16 == 0
        got += [np.array(x) for x in pk]
        ts += list(t)
    want = [good[0], good[1], good[2], good[4]]
    assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))
    k = 1 if nanos else 1000
    assert ts == [10 * 10**9 + 5 * k, 11 * 10**9 + 7 * k, 12 * 10**9 + 9 * k, 14 * 10**9 + 13 * k]
    assert src.packets_read == 4 and src.skipped == 5
    # any port
    src = ob.PcapLidarSource(p, size)
    n = sum(len(t) for _, t in src)
    assert n == 5


def test_open_errors(ob, tmp_path):
    with pytest.raises(RuntimeError, match="Failed to open pcap file"):
        ob.PcapLidarSource(tmp_path / "missing.pcap", 100)
    bad = tmp_path / "ng.pcap"
    bad.write_bytes(struct.pack("<I", 0x0a0d0d0a) + b"\x00" * 40)      # pcapng section header
    with pytest.raises(RuntimeError, match="Unsupported pcap format"):
        ob.PcapLidarSource(bad, 100)


@pytest.mark.gpu
@pytest.mark.parametrize("name", PCAP_FIXTURES[:3])
def test_capture_to_batcher_in_place(ob, tmp_path, name):
    meta, packets = load_fixture(name)
    if len(packets) != 64:
        pytest.skip("capture does not hold a whole frame")
    size = len(packets[0])
    rec = [(100 + i, i, _udp_frame(p)) for i, p in enumerate(packets)]
    p = tmp_path / "f.pcap"
    write_pcap(p, rec)
    si = ob.SensorInfo.from_meta(meta)
    b, fr = ob.FrameBatcher(si), ob.LidarFrame(si)
    src = ob.PcapLidarSource(p, size, ring_packets=24)
    done = False
    for pk, ts in src:
        off = 0
        while off < len(ts):
            used, done = b.batch_burst(pk[off:], ts[off:], fr)
            off += max(used, 1)
    assert done and src.packets_read == 64
    dg = meta["md5_digests"]
    for k, v in dg.items():
        if k in ("FRAME_ID", "TIMESTAMP", "STATUS", "MEASUREMENT_ID", "ENCODER_COUNT"):
            continue
        assert hashlib.md5(np.ascontiguousarray(fr.field(k)).tobytes()).hexdigest() == v, k
    assert np.array_equal(fr.packet_timestamp, np.array([(100 + i) * 10**9 + i * 1000 for i in range(64)], np.uint64))
