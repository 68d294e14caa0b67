"""The reference's own Python tests for this path, restated against the drop-in module
(python/tests/test_destagger.py:12-114, python/tests/test_xyzlut.py:119-136) on the 'legacy-2.0'
fixture (OS-2-32-U0_v2.0.0_1024x10)."""
import os

import numpy as np
import pytest

import __graft_entry__ as graft
from tests.helpers import GOLDEN, load_fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def core():
    graft.build()
    ob = graft.load_package()
    assert ob.device_count() > 0
    return ob.pyapi


@pytest.fixture(scope="module")
def meta_frame(core):
    meta, packets = load_fixture("OS-2-32-U0_v2.0.0_1024x10")
    info = core.SensorInfo.from_meta(meta)
    scan = core.LidarScan(info)
    batch = core.ScanBatcher(info)
    done = [batch(p, 77, scan) for p in packets]
    assert done[-1]
    return info, scan, meta


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.uint64, np.int8, np.int16,
                                   np.int32, np.int64, np.float32, np.float64])
def test_destagger_type_good(core, meta_frame, dtype):
    info = meta_frame[0]
    assert core.destagger(info, np.zeros((info.h, info.w), dtype)).dtype == dtype
    assert core.destagger(info, np.zeros((info.h, info.w, 2), dtype)).dtype == dtype


@pytest.mark.parametrize("shape", [(32, 1024), (32, 1024, 1), (32, 1024, 10)])
def test_destagger_shape_good(core, meta_frame, shape):
    info = meta_frame[0]
    assert core.destagger(info, np.zeros(shape)).shape == shape
    assert core.destagger(info, np.zeros(shape), inverse=True).shape == shape


def test_destagger_shape_bad(core, meta_frame):
    info = meta_frame[0]
    h, w = info.h, info.w
    for shape in [(0, w), (h, 0, 2), (h, w + 1), (h - 1, w), (h, w - 1, 1), (h + 1, w, 2)]:
        with pytest.raises(ValueError):
            core.destagger(info, np.zeros(shape))


def test_destagger_inverse(core, meta_frame):
    info = meta_frame[0]
    a = np.arange(info.h * info.w).reshape((info.h, info.w))
    assert np.array_equal(a, core.destagger(info, core.destagger(info, a, inverse=True)))
    d = core.destagger(info, a)
    assert np.array_equal(a, core.destagger(info, d, inverse=True))
    assert np.array_equal(a, core.stagger(info, d))


def test_destagger_xyz_and_correct(core, meta_frame):
    info, scan, meta = meta_frame
    xyz = core.XYZLut(info)(scan)
    assert xyz.shape == (info.h, info.w, 3) and xyz.dtype == np.float64
    assert core.destagger(info, xyz).shape == (info.h, info.w, 3)
    rng = scan.field(core.ChanField.RANGE)
    ref = np.stack([np.roll(rng[u], info.pixel_shift_by_row[u]) for u in range(info.h)])  # reference.py:158-161
    assert np.array_equal(ref, core.destagger(info, rng))
    near_ir = scan.field(core.ChanField.NEAR_IR)
    stacked = np.repeat(near_ir[..., None], 5, axis=2)
    ref_ir = np.stack([np.roll(near_ir[u], info.pixel_shift_by_row[u]) for u in range(info.h)])
    out = core.destagger(info, stacked)
    assert out.dtype == np.uint16 and np.array_equal(out, np.repeat(ref_ir[..., None], 5, axis=2))


def test_xyzlut_vs_doc_formula(core, meta_frame):
    """python/tests/test_xyzlut.py:119-136: XYZLut(info)(scan) allclose the manual's formula."""
    info, scan, meta = meta_frame
    rng = scan.field("RANGE").astype(np.float64)
    h, w = info.h, info.w
    n = np.hypot(info.beam_to_lidar_transform[0, 3], info.beam_to_lidar_transform[2, 3])
    v = np.arange(w)
    te = 2.0 * np.pi * (1.0 - v / w)
    ta = -2.0 * np.pi * np.asarray(info.beam_azimuth_angles) / 360.0
    phi = 2.0 * np.pi * np.asarray(info.beam_altitude_angles) / 360.0
    b03, b23 = info.beam_to_lidar_transform[0, 3], info.beam_to_lidar_transform[2, 3]
    x = (rng - n) * np.cos(te[None] + ta[:, None]) * np.cos(phi[:, None]) + b03 * np.cos(te[None])
    y = (rng - n) * np.sin(te[None] + ta[:, None]) * np.cos(phi[:, None]) + b03 * np.sin(te[None])
    z = (rng - n) * np.sin(phi[:, None]) + b23
    ref = np.stack([x, y, z, np.ones_like(x)], -1) @ info.lidar_to_sensor_transform.T
    ref = ref[..., :3] * 0.001
    ref[rng == 0] = 0
    assert np.allclose(core.XYZLut(info, use_extrinsics=False)(scan), ref)
    assert np.allclose(core.XYZLutFloat(info, use_extrinsics=False)(scan), ref, rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError, match="Image dimensions do not match lut."):
        core.XYZLut(info)(np.zeros((h, w + 1), np.uint32))


def test_pyapi_dewarp_transform_dispatch(core):
    """python/tests/test_pose_util.py:300-360: dtype dispatch and error types of dewarp/transform"""
    poses = np.tile(np.eye(4), (4, 1, 1))
    poses[:, :3, 3] = [1, -2, 3]
    pts = np.array([[i - 3, i + 1, i + 2] for i in range(8)], np.float64).reshape(2, 4, 3)
    out = core.dewarp(pts, poses)
    assert out.dtype == np.float64 and np.array_equal(out, pts + [1, -2, 3])
    out32 = core.dewarp(pts.astype(np.float32), poses)
    assert out32.dtype == np.float32 and np.array_equal(out32, (pts + [1, -2, 3]).astype(np.float32))
    assert np.array_equal(core.transform(pts, poses[0]), pts + [1, -2, 3])
    with pytest.raises(TypeError, match="floating-point"):
        core.dewarp(pts.astype(np.int32), poses)
    with pytest.raises(RuntimeError, match="Number of points per set must match number of poses"):
        core.dewarp(np.zeros((2, 5, 3)), poses)


def test_device_chain_scanbatcher_xyzlut_destagger_normals(core):
    """ScanBatcher -> XYZLut -> destagger -> normals on CUDA tensors / DLPack, nothing
    copied to the host in between; every stage equal to the CPU oracle on the same packets."""
    import torch
    from oracle import oracle as orc
    from tests.helpers import oracle_pf
    meta, packets = load_fixture("OS-1-128_767798045_1024x10_20230712_120049")
    info = core.SensorInfo.from_meta(meta)
    h, w = info.h, info.w
    lut = core.XYZLut(info)
    batcher = core.DeviceScanBatcher(info, lut=lut)
    scan = batcher.new_scan()
    done = [batcher(p, 77, scan) for p in packets]
    if not done[-1]:                  # the fixture ends inside the frame: materialise what arrived
        batcher.flush(scan)
    # oracle decode of the same packets
    pf = oracle_pf(meta)
    oframe = orc.Frame(pf)
    ob_ = orc.Batcher(pf, init_id=meta["init_id"], column_window=meta["column_window"])
    for p in packets:
        ob_.batch(p, 77, oframe)
    for name in scan.fields:
        t = scan.field(name)
        assert t.is_cuda
        ref = oframe.field(name)
        assert np.array_equal(t.cpu().numpy().view(ref.dtype), ref), name
    assert np.array_equal(scan.timestamp, oframe.timestamp)
    shifts = np.asarray(meta["pixel_shift_by_row"], np.int32)
    d, o = lut.direction, lut.offset
    ref_xyz = orc.cartesian(oframe.field("RANGE"), d, o)
    # fused products of the same launch, on the device
    assert scan.xyz[0].is_cuda and np.array_equal(scan.xyz[0].cpu().numpy(), ref_xyz)
    assert np.array_equal(scan.range_destaggered[0].cpu().numpy().view(np.uint32),
                          orc.destagger(oframe.field("RANGE"), shifts))
    # XYZLut on a device range image (through DLPack) -> device points
    class Capsule:   # a foreign DLPack exporter
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, stream=None):
            return self.t.__dlpack__()

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()
    pts = lut(Capsule(scan.field("RANGE")))
    assert pts.is_cuda and pts.shape == (h, w, 3)
    assert np.array_equal(pts.cpu().numpy().reshape(-1, 3), ref_xyz)
    # destagger on the device, any dtype / trailing dims
    xd = core.destagger(info, pts)
    rd = core.destagger(info, scan.field("RANGE"))
    assert xd.is_cuda and rd.is_cuda
    assert np.array_equal(xd.cpu().numpy(), orc.destagger(ref_xyz.reshape(h, w, 3), shifts))
    # normals on the device
    org = np.zeros((w, 3))
    n, sub = core.normals(xd, rd, org, return_subtent=True)
    assert n.is_cuda and n.shape == (h, w, 3)
    ref_n = orc.normals(xd.cpu().numpy(), rd.cpu().numpy().view(np.uint32), sensor_origins_xyz=org,
                        vertical_subtent=sub)
    assert np.array_equal(n.cpu().numpy(), ref_n)
    back = torch.from_dlpack(n)          # results are DLPack exporters themselves
    assert back.data_ptr() == n.data_ptr()
