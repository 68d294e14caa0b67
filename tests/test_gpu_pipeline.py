"""GPU tests of the pipelined decode path: ob_decode_job (persistent device buffers, asynchronous
completion), FrameBatcher.batch_burst (zero-copy uploads from page-locked bursts),
FrameBatcher.set_pipeline_depth / wait, and FramePipeline.  Everything is compared with the CPU
oracle's batcher fed the same packets (bit-exact fields and headers; XYZ equal to the oracle's
float path)."""
import numpy as np
import pytest

import __graft_entry__ as graft
from oracle import oracle as orc
from tests.helpers import oracle_pf, random_frame, random_lut

pytestmark = pytest.mark.gpu

PROFILE, H, W = "RNG19_RFL8_SIG16_NIR16_DUAL", 64, 512
SHIFTS = np.tile(np.array([12, 8, 4, 0], np.int32), H // 4)


@pytest.fixture(scope="module")
def ob():
    graft.build()
    m = graft.load_package()
    assert m.device_count() > 0
    return m


def _frames(opf, n, first_id=900, seed=11):
    out = []
    for k in range(n):
        f = random_frame(opf, seed=seed + k, frame_id=first_id + k)
        p, t = orc.frame_to_packets(f, opf, prod_sn=5)
        out.append((f, np.ascontiguousarray(p), np.asarray(t, np.uint64)))
    return out


def _oracle_frames(opf, packets, ts):
    """the oracle batcher's finished frames for a packet stream (list of deep copies)"""
    b = orc.Batcher(opf)
    of = orc.Frame(opf, with_window=True)
    done = []
    for p, t in zip(packets, ts):
        if b.batch(p, int(t), of):
            snap = {n: of.field(n).copy() for n in of.field_names}
            snap.update(frame_id=of.frame_id, timestamp=of.timestamp.copy(), status=of.status.copy(),
                        measurement_id=of.measurement_id.copy(), packet_timestamp=of.packet_timestamp.copy())
            done.append(snap)
    return done, b.dropped_packets


def _check(fr, snap, names):
    assert fr.frame_id == snap["frame_id"]
    for n in names:
        assert np.array_equal(fr.field(n), snap[n]), n
    for k in ("timestamp", "status", "measurement_id", "packet_timestamp"):
        assert np.array_equal(getattr(fr, k), snap[k]), k


def _check_cloud(slot, snap, d, o):
    for r, nm in enumerate(("RANGE", "RANGE2")):
        assert np.array_equal(slot.xyz[r], orc.cartesian(snap[nm], d, o)), nm
        assert np.array_equal(slot.range_destaggered[r], orc.destagger(snap[nm], SHIFTS)), nm


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("depth", [1, 3])
def test_pipeline_matches_oracle(ob, pinned, depth):
    opf = oracle_pf(PROFILE, H, W)
    si = ob.SensorInfo(PROFILE, H, W, fw_rev="v3.2.1", pixel_shift_by_row=SHIFTS)
    d, o = random_lut(H * W, 3)
    lut = ob.XYZLutT.from_arrays(d, o, H, W)
    frames = _frames(opf, 7)
    n_pk = frames[0][1].shape[0]
    names = [n for n in orc.Frame(opf, with_window=True).field_names]
    allp = [p for _, pk, _ in frames for p in pk]
    allt = [t for _, _, ts in frames for t in ts]
    want, _ = _oracle_frames(opf, allp, allt)
    assert len(want) == 7

    pipe = ob.FramePipeline(si, depth=depth, lut=lut, pixel_shift_by_row=SHIFTS)
    buf = ob.pinned_empty((n_pk, frames[0][1].shape[1]), np.uint8) if pinned else np.empty_like(frames[0][1])
    got = 0
    for _, pk, ts in frames:
        buf[...] = pk
        used, slot = pipe.push_burst(buf, ts)
        assert used == n_pk
        buf[...] = 0xA5            # the burst memory is the caller's again once push_burst returns
        if slot is not None:
            _check(slot.frame, want[got], names)
            _check_cloud(slot, want[got], d, o)
            got += 1
    assert got == 7 - depth and pipe.in_flight == depth
    while True:
        slot = pipe.drain()
        if slot is None:
            break
        _check(slot.frame, want[got], names)
        _check_cloud(slot, want[got], d, o)
        got += 1
    assert got == 7
    assert pipe.gpu_launches == 7


def test_pipeline_with_faults_and_split_bursts(ob):
    """dropped / duplicated / swapped packets and bursts that straddle frame boundaries"""
    opf = oracle_pf(PROFILE, H, W)
    si = ob.SensorInfo(PROFILE, H, W, fw_rev="v3.2.1")
    frames = _frames(opf, 6, seed=40)
    n = frames[0][1].shape[0]
    pk = [p.copy() for _, pp, _ in frames for p in pp]
    ts = [int(t) for _, _, tt in frames for t in tt]
    pk[3], pk[4] = pk[4], pk[3]
    ts[3], ts[4] = ts[4], ts[3]
    del pk[n + 7], ts[n + 7]
    pk.insert(2 * n + 5, pk[2 * n + 4].copy()); ts.insert(2 * n + 5, ts[2 * n + 4])
    # a packet of frame 4 arrives early, inside frame 3 (goes through the batcher's cache)
    early = pk.pop(4 * n + 2); et = ts.pop(4 * n + 2)
    pk.insert(3 * n + 9, early); ts.insert(3 * n + 9, et)
    want, dropped = _oracle_frames(opf, pk, ts)
    names = [x for x in orc.Frame(opf, with_window=True).field_names]

    pipe = ob.FramePipeline(si, depth=2)
    stream = ob.pinned_empty((len(pk), pk[0].size), np.uint8)
    stream[...] = np.stack(pk)
    tsa = np.asarray(ts, np.uint64)
    got, pos, burst = [], 0, 11          # 11 does not divide the 32 packets of a frame
    while pos < len(pk):
        end = min(pos + burst, len(pk))
        while pos < end:
            used, slot = pipe.push_burst(stream[pos:end], tsa[pos:end])
            assert used > 0
            pos += used
            if slot is not None:
                got.append(slot)
                _check(slot.frame, want[len(got) - 1], names)
    while (slot := pipe.drain()) is not None:
        got.append(slot)
        _check(slot.frame, want[len(got) - 1], names)
    assert len(got) == len(want) >= 5
    assert pipe.dropped_packets == dropped


def test_deferred_batcher_api(ob):
    """set_pipeline_depth(2): batch() returns at submission, wait(frame) materialises it"""
    opf = oracle_pf(PROFILE, H, W)
    si = ob.SensorInfo(PROFILE, H, W, fw_rev="v3.2.1")
    frames = _frames(opf, 4, seed=70)
    names = [x for x in orc.Frame(opf, with_window=True).field_names]
    want, _ = _oracle_frames(opf, [p for _, pk, _ in frames for p in pk], [t for _, _, ts in frames for t in ts])
    b = ob.FrameBatcher(si)
    b.set_pipeline_depth(2)
    frs = [ob.LidarFrame(si) for _ in range(4)]
    for k, (_, pk, ts) in enumerate(frames):
        used, done = b.batch_burst(pk, ts, frs[k])         # pageable numpy: bounce-buffer path
        assert done and used == pk.shape[0]
        # host-written headers are valid without waiting
        assert np.array_equal(frs[k].timestamp, want[k]["timestamp"])
    b.wait()
    for k in range(4):
        _check(frs[k], want[k], names)
    # back to synchronous mode: frame is complete when batch returns
    b.set_pipeline_depth(1)
    more = _frames(opf, 1, first_id=2000, seed=99)[0]
    fr = ob.LidarFrame(si)
    used, done = b.batch_burst(more[1], more[2], fr)
    assert done
    w2, _ = _oracle_frames(opf, list(more[1]), list(more[2]))
    _check(fr, w2[0], names)


def test_same_frame_object_reused_in_deferred_mode(ob):
    """reusing one LidarFrame with depth 2 must still yield the last frame's data (implicit wait)"""
    opf = oracle_pf(PROFILE, H, W)
    si = ob.SensorInfo(PROFILE, H, W, fw_rev="v3.2.1")
    frames = _frames(opf, 3, seed=120)
    names = [x for x in orc.Frame(opf, with_window=True).field_names]
    want, _ = _oracle_frames(opf, [p for _, pk, _ in frames for p in pk], [t for _, _, ts in frames for t in ts])
    b = ob.FrameBatcher(si)
    b.set_pipeline_depth(2)
    fr = ob.LidarFrame(si)
    for _, pk, ts in frames:
        _, done = b.batch_burst(pk, ts, fr)
        assert done
    b.wait(fr)
    _check(fr, want[2], names)


def test_decode_job_c_abi_device_outputs(ob):
    """ob_decode_job_* directly: uploads in two pieces, device-resident outputs, resubmission"""
    import ctypes as C
    import torch
    from importlib import import_module
    capi = import_module(ob.__name__ + "._capi")
    lib, check = capi.lib, capi.check
    opf = oracle_pf(PROFILE, H, W)
    src = random_frame(opf, seed=8)
    pk, _ = orc.frame_to_packets(src, opf)
    pk = np.ascontiguousarray(pk)
    from tests.helpers import decoder_desc_from_oracle
    layout, fields = decoder_desc_from_oracle(opf, src)
    dec = ob.Decoder(layout, fields)
    st = ob.Stream()
    job = C.c_void_p()
    lib.ob_decode_job_create.restype = C.c_int
    lib.ob_decode_job_create.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.ob_decode_job_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
    lib.ob_decode_job_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    for f in ("ob_decode_job_uploads_done", "ob_decode_job_wait", "ob_decode_job_destroy", "ob_decode_job_busy"):
        getattr(lib, f).argtypes = [C.c_void_p]
    check(lib.ob_decode_job_create(dec._h, 4, st.h, C.byref(job)))     # reserve < needed: grows
    half = pk.shape[0] // 2
    check(lib.ob_decode_job_upload(job, pk.ctypes.data, pk.strides[0], 0, half))
    check(lib.ob_decode_job_upload(job, pk[half:].ctypes.data, pk.strides[0], half, pk.shape[0] - half))
    check(lib.ob_decode_job_uploads_done(job))
    io = capi.DecodeIO()
    io.n_slots = pk.shape[0]
    outs = {}
    for i, f in enumerate(dec.fields):
        t = torch.zeros(H * W * f["elem_size"], dtype=torch.uint8, device="cuda")
        outs[f["name"]] = t
        io.fields[i] = t.data_ptr()
    check(lib.ob_decode_job_submit(job, C.byref(io), None, None, 0))
    assert lib.ob_decode_job_busy(job) == 1
    check(lib.ob_decode_job_wait(job))
    assert lib.ob_decode_job_busy(job) == 0
    for f in dec.fields:
        want = src.field(f["name"])
        got = outs[f["name"]].cpu().numpy().view(want.dtype).reshape(want.shape)
        assert np.array_equal(got, want), f["name"]
    # host outputs through the job's slab on resubmission of the same slots
    host = {f["name"]: np.zeros_like(src.field(f["name"])) for f in dec.fields}
    for i, f in enumerate(dec.fields):
        io.fields[i] = host[f["name"]].ctypes.data
    check(lib.ob_decode_job_submit(job, C.byref(io), None, None, 0))
    check(lib.ob_decode_job_wait(job))
    for f in dec.fields:
        assert np.array_equal(host[f["name"]], src.field(f["name"])), f["name"]
    # more slots than uploaded -> refused with the reference-style argument error
    io.n_slots = pk.shape[0] + 1
    assert lib.ob_decode_job_submit(job, C.byref(io), None, None, 0) != 0
    check(lib.ob_decode_job_destroy(job))


def test_frame_destroyed_before_wait_does_not_corrupt_the_next_frame(ob):
    """depth 2: a LidarFrame dropped right after its GPU pass was submitted.  The job in flight shares ownership
    of the frame's page-locked block, so the block cannot go back to the pool (and into the next LidarFrame that
    is created) before the device->host copy has landed."""
    opf = oracle_pf(PROFILE, H, W)
    si = ob.SensorInfo(PROFILE, H, W, fw_rev="v3.2.1")
    frames = _frames(opf, 6, seed=300)
    b = ob.FrameBatcher(si)
    b.set_pipeline_depth(2)
    for k, (_, pk, ts) in enumerate(frames):
        fr = ob.LidarFrame(si)
        _, done = b.batch_burst(pk, ts, fr)
        assert done
        del fr                                   # destroyed with its GPU pass in flight
        nxt = ob.LidarFrame(si)                  # would pick the freed block up from the pool
        for name in nxt.fields:
            nxt.field(name)[...] = 0x5A
        b.wait()                                 # the orphaned job lands (somewhere else)
        for name in nxt.fields:
            a = nxt.field(name)
            assert np.all(a == 0x5A), (k, name)
        del nxt
