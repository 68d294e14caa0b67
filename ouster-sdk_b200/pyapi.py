"""Python drop-in for the hot-path part of `ouster.sdk.core`: same call shapes and
error behaviour as the nanobind binding (python/src/cpp/client/processing.cpp:340-357, 527-700):

    from ouster_sdk_b200 import pyapi as core
    xyz = core.XYZLut(info)(scan)                 # (H, W, 3) float64, staggered
    img = core.destagger(info, scan.field("RANGE"))
"""
import numpy as np

from . import core as _c
from .host import FrameBatcher, LidarFrame, LidarScan, ScanBatcher, SensorInfo  # noqa: F401


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
        if hasattr(scan_or_range, "field"):
            rng = scan_or_range.field("RANGE")
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
    p, q = _floating(points, "points and poses"), _floating(poses, "points and poses")
    dt = np.float32 if p.dtype == np.float32 else np.float64
    return _c.dewarp(np.ascontiguousarray(p, dt), np.ascontiguousarray(q, dt))


def transform(points, pose):
    """core.transform(points (..., 3), pose (4, 4)) (processing.cpp:312-329)."""
    p, q = _floating(points, "points and pose"), _floating(pose, "points and pose")
    dt = np.float32 if p.dtype == np.float32 else np.float64
    return _c.transform(np.ascontiguousarray(p, dt), np.ascontiguousarray(q, dt))
