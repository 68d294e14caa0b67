"""Python drop-in for the hot-path part of `ouster.sdk.core`: same call shapes and
error behaviour as the nanobind binding (python/src/cpp/client/processing.cpp:340-357, 527-700):

    from ouster_sdk_b200 import pyapi as core
    xyz = core.XYZLut(info)(scan)                 # (H, W, 3) float64, staggered
    img = core.destagger(info, scan.field("RANGE"))
"""
import numpy as np

from . import core as _c
from .host import FrameBatcher, LidarFrame, LidarScan, ScanBatcher, SensorInfo  # noqa: F401

# ---- device-resident data -----------------------------------------------------------
# Every function below also takes torch CUDA tensors, or any object exporting __dlpack__ from device
# memory, and then returns a torch CUDA tensor (itself a DLPack exporter): a caller can chain
# DeviceScanBatcher -> XYZLut -> destagger -> normals without the data ever leaving HBM.


def _torch():
    import torch
    return torch


def _dev(x):
    """torch CUDA tensor view of `x` when it lives on a GPU (torch tensor or DLPack exporter), else None."""
    if isinstance(x, np.ndarray) or x is None:
        return None
    if _c._is_torch(x):
        return x if x.is_cuda else None
    if hasattr(x, "__dlpack__") and hasattr(x, "__dlpack_device__"):
        kind = int(x.__dlpack_device__()[0])
        if kind in (2, 13):          # kDLCUDA, kDLCUDAManaged
            return _torch().from_dlpack(x)
    return None


class ChanField:
    RANGE, RANGE2, SIGNAL, SIGNAL2 = "RANGE", "RANGE2", "SIGNAL", "SIGNAL2"
    REFLECTIVITY, REFLECTIVITY2, NEAR_IR = "REFLECTIVITY", "REFLECTIVITY2", "NEAR_IR"
    FLAGS, FLAGS2, WINDOW = "FLAGS", "FLAGS2", "WINDOW"


def _info_dict(info):
    return {"w": info.w, "h": info.h, "beam_to_lidar_transform": info.beam_to_lidar_transform,
            "lidar_to_sensor_transform": info.lidar_to_sensor_transform,
            "sensor_to_body": getattr(info, "sensor_to_body", None),
            "beam_azimuth_angles": info.beam_azimuth_angles,
            "beam_altitude_angles": info.beam_altitude_angles}


class _XYZLutBase:
    _dtype = np.float64

    def __init__(self, info, use_extrinsics=True, device=0):
        self._lut = _c.XYZLutT.from_sensor_info(_info_dict(info), use_extrinsics, self._dtype, device)
        self.h, self.w = self._lut.h, self._lut.w

    @property
    def direction(self):
        return self._lut.direction

    @property
    def offset(self):
        return self._lut.offset

    def __call__(self, scan_or_range):
        """lut(scan) / lut(range image) -> (H, W, 3).  Raises ValueError on a dimension mismatch
        ("Frame dimensions do not match lut." / "Image dimensions do not match lut.")."""
        is_scan = hasattr(scan_or_range, "field")
        rng = scan_or_range.field("RANGE") if is_scan else scan_or_range
        t = _dev(rng)
        if t is not None:   # device range image -> device points
            if tuple(t.shape) != (self.h, self.w):
                raise ValueError("Frame dimensions do not match lut." if is_scan else "Image dimensions do not match lut.")
            torch = _torch()
            if t.dtype not in (torch.int32, torch.uint32):
                raise ValueError("range must be uint32")
            out = torch.empty((self.h * self.w, 3), device=t.device,
                              dtype=torch.float64 if self._dtype == np.float64 else torch.float32)
            st = _c.Stream(t.device.index, cuda_stream=torch.cuda.current_stream(t.device).cuda_stream)
            _c.check(_c.lib.ob_cartesian(self._lut._h, t.contiguous().data_ptr(), self.h * self.w, out.data_ptr(), st.h))
            return out.reshape(self.h, self.w, 3)
        if is_scan:
            if rng.shape != (self.h, self.w):
                raise ValueError("Frame dimensions do not match lut.")
        else:
            rng = np.asarray(scan_or_range)
            if rng.shape != (self.h, self.w):
                raise ValueError("Image dimensions do not match lut.")
        return self._lut(np.ascontiguousarray(rng, np.uint32)).reshape(self.h, self.w, 3)


class XYZLut(_XYZLutBase):
    _dtype = np.float64


class XYZLutFloat(_XYZLutBase):
    _dtype = np.float32


def destagger(info, fields, inverse=False):
    """core.destagger(info, fields, inverse=False): (H, W) or (H, W, k) array of any numeric dtype;
    dtype and shape are preserved; ValueError when the shape does not match the sensor."""
    t = _dev(fields)
    if t is not None:   # device image -> device image (same dtype / shape)
        if t.dim() < 2 or t.dim() > 3:
            raise ValueError("Invalid dimensions for destaggering")
        if t.shape[0] != info.h or t.shape[1] != info.w or t.numel() == 0:
            raise ValueError("Image resolution must match SensorInfo.")
        torch = _torch()
        st = _c.Stream(t.device.index, cuda_stream=torch.cuda.current_stream(t.device).cuda_stream)
        return _c.destagger(t.contiguous(), info.pixel_shift_by_row, inverse, stream=st, device=t.device.index)
    a = np.asarray(fields)
    if a.ndim < 2 or a.ndim > 3:
        raise ValueError("Invalid dimensions for destaggering")
    if a.shape[0] != info.h or a.shape[1] != info.w or a.size == 0:
        raise ValueError("Image resolution must match SensorInfo.")
    if a.dtype == np.bool_:
        return destagger(info, a.view(np.uint8), inverse).view(np.bool_)
    return _c.destagger(np.ascontiguousarray(a), info.pixel_shift_by_row, inverse)


def stagger(info, fields):
    return destagger(info, fields, True)


def _floating(a, what):
    a = np.asarray(a)
    if a.dtype.kind != "f":
        raise TypeError(f"{what} must be floating-point arrays")
    return a


def dewarp(points, poses):
    """core.dewarp(points (H, W, 3), poses (W, 4, 4)) -> (H, W, 3) (processing.cpp:132-161, 300-310):
    float32 points stay float32, anything else is computed in float64; TypeError for non-floating
    input, RuntimeError when W differs."""
    tp = _dev(points)
    if tp is not None:   # device points (+ host or device poses) -> device points
        torch = _torch()
        if not tp.dtype.is_floating_point:
            raise TypeError("points and poses must be floating-point arrays")
        dt = torch.float32 if tp.dtype == torch.float32 else torch.float64
        tq = _dev(poses)
        tq = tq.to(dt) if tq is not None else torch.as_tensor(np.ascontiguousarray(poses), dtype=dt, device=tp.device)
        st = _c.Stream(tp.device.index, cuda_stream=torch.cuda.current_stream(tp.device).cuda_stream)
        return _c.dewarp(tp.to(dt).contiguous(), tq.contiguous(), stream=st, device=tp.device.index)
    p, q = _floating(points, "points and poses"), _floating(poses, "points and poses")
    dt = np.float32 if p.dtype == np.float32 else np.float64
    return _c.dewarp(np.ascontiguousarray(p, dt), np.ascontiguousarray(q, dt))


def transform(points, pose):
    """core.transform(points (..., 3), pose (4, 4)) (processing.cpp:312-329)."""
    p, q = _floating(points, "points and pose"), _floating(pose, "points and pose")
    dt = np.float32 if p.dtype == np.float32 else np.float64
    return _c.transform(np.ascontiguousarray(p, dt), np.ascontiguousarray(q, dt))


def normals(xyz, range, *args, **kwargs):
    """algorithm.normals(xyz, range[, xyz2, range2], sensor_origins_xyz, pixel_search_range=1,
    min_angle_of_incidence_rad=1 deg, target_distance_m=0.025) -- python binding of
    ouster_algorithm/include/ouster/algorithm/normals.h:58-108.  Device inputs give device outputs."""
    t = _dev(xyz)
    if t is None:
        return _c.normals(xyz, range, *args, **kwargs)
    torch = _torch()
    conv = [(_dev(a) if _dev(a) is not None else a) for a in args]
    conv = [torch.as_tensor(np.ascontiguousarray(a, np.float64), device=t.device)
            if isinstance(a, np.ndarray) and a.ndim == 2 and a.shape[-1] == 3 else a for a in conv]
    if "sensor_origins_xyz" in kwargs and isinstance(kwargs["sensor_origins_xyz"], np.ndarray):
        kwargs["sensor_origins_xyz"] = torch.as_tensor(np.ascontiguousarray(kwargs["sensor_origins_xyz"], np.float64),
                                                       device=t.device)
    st = _c.Stream(t.device.index, cuda_stream=torch.cuda.current_stream(t.device).cuda_stream)
    return _c.normals(t.contiguous(), _dev(range), *conv, stream=st, device=t.device.index, **kwargs)


class DeviceLidarScan:
    """The pixel fields of a LidarScan as CUDA tensors (h x w each, torch, DLPack-exportable); the
    per-column / per-packet headers stay in the host LidarScan `self.host` (they are a few KB and the
    host state machine owns them)."""

    def __init__(self, info, device=0, fused_returns=0, xyz_dtype=np.float32):
        torch = _torch()
        self.info, self.host = info, LidarScan(info)
        self.h, self.w = self.host.h, self.host.w
        dev = torch.device("cuda", device)
        tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.int16, np.dtype(np.uint32): torch.int32,
               np.dtype(np.uint64): torch.int64, np.dtype(np.int8): torch.int8, np.dtype(np.int16): torch.int16,
               np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
               np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}
        self._fields = {}
        for name in self.host.fields:
            a = self.host.field(name)
            self._fields[name] = torch.zeros(a.shape, dtype=tdt[a.dtype], device=dev)
        xt = torch.float64 if np.dtype(xyz_dtype) == np.float64 else torch.float32
        self.xyz = [torch.zeros((self.h * self.w, 3), dtype=xt, device=dev) for _ in range(fused_returns)]
        self.range_destaggered = [torch.zeros((self.h, self.w), dtype=torch.int32, device=dev)
                                  for _ in range(fused_returns)]

    @property
    def fields(self):
        return list(self._fields)

    def field(self, name):
        """CUDA tensor of the field (unsigned 16/32-bit fields come as the same-width signed torch dtype:
        identical bits; `.view(torch.uint16)` etc. where torch has the type)."""
        return self._fields[name]

    # headers, as on LidarScan
    timestamp = property(lambda self: self.host.timestamp)
    measurement_id = property(lambda self: self.host.measurement_id)
    status = property(lambda self: self.host.status)
    packet_timestamp = property(lambda self: self.host.packet_timestamp)
    frame_id = property(lambda self: self.host.frame_id)


class DeviceScanBatcher:
    """ScanBatcher whose decoded scan stays on the GPU: the reference's per-packet state machine
    (lidar_frame.cpp:1698-1959, host) + the fused decode launch writing into the DeviceLidarScan's
    CUDA tensors.  With `lut` the same launch also produces scan.xyz[r] / scan.range_destaggered[r].

        batcher = DeviceScanBatcher(info, lut=XYZLutFloat(info))
        scan = batcher.new_scan()
        for packet, ts in stream:
            if batcher(packet, ts, scan):
                n = normals(destagger(info, scan.xyz[0].reshape(h, w, 3)), scan.range_destaggered[0], origins)
    """

    def __init__(self, info, lut=None, device=0):
        self.info, self.device = info, device
        self._b = FrameBatcher(info)
        self._lut = lut._lut if isinstance(lut, _XYZLutBase) else lut
        if self._lut is not None:
            self._b.set_fused_cloud(self._lut, info.pixel_shift_by_row)
        self._bound = None

    def new_scan(self):
        n_ret = 0
        if self._lut is not None:
            names = LidarScan(self.info).fields
            n_ret = 1 + int("RANGE2" in names)
        return DeviceLidarScan(self.info, self.device, n_ret, self._lut.dtype if self._lut is not None else np.float32)

    def _bind(self, scan):
        if self._bound is not scan:
            self._b.set_device_outputs({n: scan.field(n) for n in scan.fields}, scan.xyz or None,
                                       scan.range_destaggered or None)
            self._bound = scan

    def batch(self, packet, host_timestamp, scan):
        self._bind(scan)
        return self._b.batch(packet, host_timestamp, scan.host)

    __call__ = batch

    def batch_burst(self, packets, host_timestamps, scan):
        self._bind(scan)
        return self._b.batch_burst(packets, host_timestamps, scan.host)

    def flush(self, scan):
        self._bind(scan)
        self._b.flush(scan.host)

    def reset(self):
        self._b.reset()

    batched_packets = property(lambda self: self._b.batched_packets)
    dropped_packets = property(lambda self: self._b.dropped_packets)
