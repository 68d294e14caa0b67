"""Edge cases through the C ABI: empty batches, tiny and maximal shapes, widths that leave partial
tiles, non-default columns_per_packet, wire layouts that are not 4-byte aligned (generic extraction
path), 64-bit destination fields."""
import numpy as np
import pytest

import __graft_entry__ as graft
from oracle import oracle as orc
from tests.helpers import (col_map_from_packets, decoder_desc_from_oracle, oracle_pf, random_frame,
                           random_lut, random_range)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    graft.build()
    m = graft.load_package()
    assert m.device_count() > 0
    return m


def test_empty_batches_are_no_ops(ob):
    d, o = random_lut(4 * 8, 1)
    lut = ob.XYZLutT.from_arrays(d, o, 4, 8)
    st = ob.Stream(0)
    ob.scan_to_cloud(lut, None, np.zeros((0, 1, 4, 8), np.uint32), xyz=np.zeros((0, 1, 32, 3), np.float32), stream=st)
    st.sync()
    assert ob.destagger(np.zeros((0, 8), np.uint32), np.zeros(0, np.int32)).shape == (0, 8)


@pytest.mark.parametrize("h,w", [(1, 4), (2, 8), (128, 4096), (256, 512), (512, 64)])
def test_extreme_shapes_scan_to_cloud(ob, h, w):
    rng = np.stack([random_range(h, w, 5), random_range(h, w, 6)])[None]
    d, o = random_lut(h * w, 7)
    shifts = (np.arange(h, dtype=np.int32) * 5) % max(w, 1) - (w // 3)
    if w & (w - 1):
        shifts = np.abs(shifts)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    xyz = np.zeros((1, 2, h * w, 3), np.float32)
    rd = np.zeros((1, 2, h, w), np.uint32)
    st = ob.Stream(0)
    ob.scan_to_cloud(lut, shifts, rng, xyz=xyz, range_destaggered=rd, stream=st)
    st.sync()
    for r in range(2):
        assert np.array_equal(xyz[0, r], orc.cartesian(rng[0, r], d, o))
        assert np.array_equal(rd[0, r], orc.destagger(rng[0, r], shifts))


def test_more_rows_than_shift_table_is_rejected(ob):
    h, w = 600, 16
    d, o = random_lut(h * w, 1)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    with pytest.raises(ValueError, match="at most 512 rows"):
        ob.scan_to_cloud(lut, np.zeros(h, np.int32), np.zeros((1, 1, h, w), np.uint32),
                         range_destaggered=np.zeros((1, 1, h, w), np.uint32))
    # the stand-alone destagger has no such limit
    img = np.arange(h * w, dtype=np.uint16).reshape(h, w)
    sh = (np.arange(h) % 16).astype(np.int32)
    assert np.array_equal(ob.destagger(img, sh), orc.destagger(img, sh))


def _decode_vs_oracle(ob, pf, with_window=True, drop=None):
    src = random_frame(pf, seed=21, with_window=with_window)
    pk, ts = orc.frame_to_packets(src, pf)
    if drop is not None:
        pk = np.delete(pk, drop, axis=0)
    ref = orc.Frame(pf, with_window=with_window)
    b = orc.Batcher(pf)
    for p in pk:
        b.batch(p, 3, ref)
    if drop is not None:   # finalize through the next frame's first packets
        nxt = orc.frame_to_packets(random_frame(pf, 22, with_window=with_window, frame_id=701), pf)[0]
        for p in nxt[:6]:
            if b.batch(p, 3, ref):
                break
    layout, fields = decoder_desc_from_oracle(pf, ref)
    dec = ob.Decoder(layout, fields)
    h, w = pf.pixels_per_column, pf.columns_per_frame
    outs = {f["name"]: np.zeros(ref.field(f["name"]).shape, ref.field(f["name"]).dtype) for f in fields}
    io = {"packets": np.ascontiguousarray(pk), "n_slots": len(pk), "packet_stride": pk.shape[1],
          "col_src": None if drop is None else col_map_from_packets(pf, pk), "fields": outs,
          "timestamp": np.zeros(w, np.uint64), "measurement_id": np.zeros(w, np.uint16),
          "status": np.zeros(w, np.uint32)}
    st = ob.Stream(0)
    dec.decode([io], stream=st)
    st.sync()
    for n, a in outs.items():
        assert np.array_equal(a, ref.field(n)), n
    assert np.array_equal(io["timestamp"], ref.timestamp)
    assert np.array_equal(io["status"], ref.status)


@pytest.mark.parametrize("w", [16, 48, 80, 1040])
def test_decode_partial_tiles(ob, w):
    _decode_vs_oracle(ob, oracle_pf("RNG19_RFL8_SIG16_NIR16", 16, w))
    _decode_vs_oracle(ob, oracle_pf("RNG19_RFL8_SIG16_NIR16", 16, w), drop=1 if w > 16 else None)


@pytest.mark.parametrize("cpp,h,w", [(8, 8, 64), (4, 4, 64), (32, 32, 128), (5, 7, 50)])
def test_decode_other_columns_per_packet(ob, cpp, h, w):
    # block_parsable() picks 8 / 4 / 16, or none at all for (5, 7): the column path
    _decode_vs_oracle(ob, oracle_pf("RNG19_RFL8_SIG16_NIR16_DUAL", h, w, cpp))
    _decode_vs_oracle(ob, oracle_pf("RNG15_RFL8_NIR8", h, w, cpp), with_window=False, drop=2)


def test_decode_unaligned_wire_layout_and_wide_fields(ob):
    """chan_data_size = 7 bytes (nothing is word aligned) and a 64-bit destination field."""
    pf = oracle_pf("RNG19_RFL8_SIG16_NIR16", 8, 64)
    pf.set_fields([("RANGE", orc.UINT32, 0, 0x7ffff, 0), ("SIGNAL", orc.UINT16, 3, 0xffff, 0),
                   ("FLAGS", orc.UINT8, 5, 0xf0, 4), ("WIDE", orc.UINT64, 0, 0x00ffffffffffffff, 0)], 7)
    ref0 = orc.Frame(pf, with_window=False)
    for n in list(ref0.field_names):
        pass
    src = orc.Frame(pf, with_window=False, extra_fields=[("WIDE", orc.UINT64)])
    rs = np.random.default_rng(5)
    # WIDE covers the whole 7-byte pixel and is encoded last (name order), so it defines the wire
    # bytes; the narrower fields decode overlapping views of it
    src.field("WIDE")[...] = rs.integers(0, 1 << 56, size=(8, 64), dtype=np.uint64)
    src.measurement_id[:] = np.arange(64)
    src.status[:] = 1
    src.packet_timestamp[:] = 7
    src.frame_id = 9
    # WIDE overlaps the other fields: encode it first as zeros, the others define the bytes
    pk, ts = orc.frame_to_packets(src, pf)
    ref = orc.Frame(pf, with_window=False, extra_fields=[("WIDE", orc.UINT64)])
    b = orc.Batcher(pf)
    for p in pk:
        b.batch(p, 3, ref)
    layout, fields = decoder_desc_from_oracle(pf, ref)
    assert layout["channel_data_size"] == 7
    dec = ob.Decoder(layout, fields)
    outs = {f["name"]: np.zeros(ref.field(f["name"]).shape, ref.field(f["name"]).dtype) for f in fields}
    io = {"packets": np.ascontiguousarray(pk), "n_slots": len(pk), "packet_stride": pk.shape[1],
          "col_src": None, "fields": outs}
    st = ob.Stream(0)
    d, o = random_lut(8 * 64, 2)
    lut = ob.XYZLutT.from_arrays(d, o, 8, 64)
    io["xyz"] = [np.zeros((8 * 64, 3), np.float32)]
    io["range_destaggered"] = [np.zeros((8, 64), np.uint32)]
    sh = np.arange(8, dtype=np.int32)
    dec.decode([io], lut=lut, pixel_shift_by_row=sh, stream=st)
    st.sync()
    for n, a in outs.items():
        assert np.array_equal(a, ref.field(n)), n
    assert np.array_equal(outs["WIDE"], src.field("WIDE")) and np.any(outs["WIDE"] > (1 << 40))
    assert np.array_equal(outs["RANGE"], (src.field("WIDE") & np.uint64(0x7ffff)).astype(np.uint32))
    assert np.array_equal(io["xyz"][0], orc.cartesian(ref.field("RANGE"), d, o))
    assert np.array_equal(io["range_destaggered"][0], orc.destagger(ref.field("RANGE"), sh))
