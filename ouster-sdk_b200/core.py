"""Host-side Python mirror of the reference's hot-path API over the C ABI.

Reference interfaces mirrored (python binding python/src/cpp/client/processing.cpp):
  XYZLut / XYZLutFloat .__call__(range|frame)   :340-357, 640-700
  destagger(info|shifts, field, inverse)         :527-638
Arrays may be numpy arrays (host) or torch tensors (host or CUDA); CUDA tensors are used
in place (zero copy), host arrays are staged by the C library.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import CloudIO, check, lib


def device_count():
    return lib.ob_device_count()


def kernel_launch_count(name=None):
    """Kernels launched since load; `name` restricts the count to one family ("decode_pipe", ...)."""
    if name is not None:
        return lib.ob_kernel_launch_count_of(name.encode())
    return lib.ob_kernel_launch_count()


def set_tunable(name, value, device=0):
    check(lib.ob_set_tunable(device, name.encode(), int(value)))


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        assert x.is_contiguous(), "tensors must be contiguous"
        return x.data_ptr()
    assert x.flags["C_CONTIGUOUS"], "arrays must be C-contiguous"
    return x.ctypes.data


def _itemsize(x):
    return x.element_size() if _is_torch(x) else x.dtype.itemsize


def _np_dtype(x):
    if _is_torch(x):
        import torch
        return {torch.float32: np.dtype(np.float32), torch.float64: np.dtype(np.float64),
                torch.uint8: np.dtype(np.uint8), torch.int32: np.dtype(np.int32),
                torch.uint32: np.dtype(np.uint32), torch.int64: np.dtype(np.int64),
                torch.uint16: np.dtype(np.uint16), torch.int16: np.dtype(np.int16),
                torch.uint64: np.dtype(np.uint64), torch.int8: np.dtype(np.int8)}[x.dtype]
    return x.dtype


class PinnedBuffer:
    """numpy view over cudaHostAlloc'd memory (ob_host_alloc)."""

    def __init__(self, nbytes):
        p = C.c_void_p()
        check(lib.ob_host_alloc(nbytes, C.byref(p)))
        self.ptr, self.nbytes = p.value, nbytes
        self.raw = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p.value))

    def __del__(self):
        if getattr(self, "ptr", None) and lib is not None:
            try:
                lib.ob_host_free(self.ptr)
            except Exception:
                pass
            self.ptr = None


def pinned_empty(shape, dtype):
    """Pinned host array; keeps its allocation alive through the `.base` chain."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    buf = PinnedBuffer(max(n, 1))
    arr = buf.raw[:n].view(dtype).reshape(shape)
    arr_holder = _Holder(arr, buf)
    return arr_holder.arr


class _Holder:
    _keep = []

    def __init__(self, arr, buf):
        self.arr = arr
        _Holder._keep.append(buf)  # pinned buffers live for the process (few, large)


class Stream:
    """ob_stream: CUDA stream + staging; one per caller thread / sensor stream."""

    def __init__(self, device=0, cuda_stream=None):
        h = C.c_void_p()
        if cuda_stream is None:
            check(lib.ob_stream_create(device, C.byref(h)))
        else:
            check(lib.ob_stream_wrap(device, C.c_void_p(cuda_stream), C.byref(h)))
        self.h, self.device = h, device

    @property
    def cuda_stream(self):
        return lib.ob_stream_cuda_handle(self.h)

    def sync(self):
        check(lib.ob_stream_sync(self.h))

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            try:
                lib.ob_stream_destroy(self.h)
            except Exception:
                pass
            self.h = None


_default_streams = {}


def _stream(stream, device=0):
    if stream is not None:
        return stream
    if device not in _default_streams:
        _default_streams[device] = Stream(device)
    return _default_streams[device]


class XYZLutT:
    """XYZLutT<T> (ouster_core/include/ouster/core/xyzlut.h:76-158) with device-resident tables.

    `direction`/`offset` are fetched lazily from the device (the reference keeps host copies)."""

    def __init__(self, handle, h, w, dtype, device):
        self._h, self.h, self.w, self.dtype, self.device = handle, h, w, np.dtype(dtype), device
        self._host = None

    # -- constructors ----------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, direction, offset, h, w, device=0):
        """XYZLutT(direction, offset, h, w)  (xyzlut.h:135)."""
        dt = _np_dtype(direction)
        if dt not in (np.dtype(np.float32), np.dtype(np.float64)) or _np_dtype(offset) != dt:
            raise ValueError("direction/offset must both be float32 or float64")
        n = h * w * 3
        dn = direction.numel() if _is_torch(direction) else direction.size
        on = offset.numel() if _is_torch(offset) else offset.size
        if dn != n or on != n:
            raise ValueError("unexpected image dimensions")
        hd = C.c_void_p()
        check(lib.ob_lut_create(_capi.OB_F64 if dt == np.float64 else _capi.OB_F32, _ptr(direction),
                                _ptr(offset), h, w, device, C.byref(hd)))
        return cls(hd, h, w, dt, device)

    @classmethod
    def from_intrinsics(cls, w, h, range_unit, beam_to_lidar_transform, transform,
                        azimuth_angles_deg, altitude_angles_deg, dtype=np.float64, device=0):
        """impl::make_xyz_lut(w, h, range_unit, ...)  (ouster_core/src/xyzlut.cpp:11-89)."""
        b2l = np.ascontiguousarray(beam_to_lidar_transform, np.float64).reshape(16)
        tr = np.ascontiguousarray(transform, np.float64).reshape(16)
        az = np.ascontiguousarray(azimuth_angles_deg, np.float64)
        alt = np.ascontiguousarray(altitude_angles_deg, np.float64)
        dt = np.dtype(dtype)
        hd = C.c_void_p()
        check(lib.ob_lut_from_intrinsics(_capi.OB_F64 if dt == np.float64 else _capi.OB_F32, w, h,
                                         range_unit, b2l.ctypes.data, tr.ctypes.data,
                                         az.ctypes.data, az.size, alt.ctypes.data, alt.size,
                                         device, C.byref(hd)))
        return cls(hd, h, w, dt, device)

    @classmethod
    def from_sensor_info(cls, info, use_extrinsics=True, dtype=np.float64, device=0):
        """XYZLutT(const SensorInfo&, bool use_extrinsics) (xyzlut.h:111-112, xyzlut.cpp:91-106).
        `info`: mapping/obj with w, h, beam_to_lidar_transform, lidar_to_sensor_transform,
        sensor_to_body (optional), beam_azimuth_angles, beam_altitude_angles."""
        g = (lambda k, d=None: info.get(k, d)) if isinstance(info, dict) else \
            (lambda k, d=None: getattr(info, k, d))
        RANGE_UNIT = 0.001  # types.h:46
        tr = np.array(g("lidar_to_sensor_transform"), np.float64).reshape(4, 4)
        ext = g("sensor_to_body")
        if use_extrinsics and ext is not None:
            ext = np.array(ext, np.float64).reshape(4, 4).copy()
            ext[:3, 3] /= RANGE_UNIT
            tr = ext @ tr
        return cls.from_intrinsics(g("w"), g("h"), RANGE_UNIT, g("beam_to_lidar_transform"), tr,
                                   g("beam_azimuth_angles"), g("beam_altitude_angles"), dtype, device)

    # -- members ---------------------------------------------------------------------------
    def _download(self):
        if self._host is None:
            d = np.empty((self.h * self.w, 3), self.dtype)
            o = np.empty((self.h * self.w, 3), self.dtype)
            check(lib.ob_lut_download(self._h, d.ctypes.data, o.ctypes.data))
            self._host = (d, o)
        return self._host

    @property
    def direction(self):
        return self._download()[0]

    @property
    def offset(self):
        return self._download()[1]

    def set_analytic(self, enable=True):
        """Opt into the LUT-free projection (ob_lut_set_analytic): direction/offset recomputed in the
        kernels from per-row / per-column tables; <= 1e-5 norm-wise vs the LUT path, not bit-exact.
        Only LUTs built from per-beam intrinsics have the tables (ValueError otherwise)."""
        check(lib.ob_lut_set_analytic(self._h, int(bool(enable))))
        return self

    @property
    def analytic(self):
        return bool(lib.ob_lut_is_analytic(self._h))

    def __call__(self, rng, out=None, stream=None):
        """lut(range) -> (h*w, 3) points, staggered order (xyzlut.h:139-150)."""
        return cartesian(self, rng, out=out, stream=stream)

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            try:
                lib.ob_lut_destroy(self._h)
            except Exception:
                pass
            self._h = None


def XYZLut(info, use_extrinsics=True, device=0):
    """Python `XYZLut` (double), python/src/cpp/client/processing.cpp:640-700."""
    return XYZLutT.from_sensor_info(info, use_extrinsics, np.float64, device)


def XYZLutFloat(info, use_extrinsics=True, device=0):
    """Python `XYZLutFloat`."""
    return XYZLutT.from_sensor_info(info, use_extrinsics, np.float32, device)


def _numel(x):
    return x.numel() if _is_torch(x) else x.size


def cartesian(lut, rng, out=None, stream=None):
    """cartesian(range, lut) / lut(range).  Raises ValueError("unexpected image dimensions")
    on size mismatch (ouster_core/src/xyzlut.cpp:117-119)."""
    st = _stream(stream, lut.device)
    if _np_dtype(rng) != np.dtype(np.uint32):
        if _is_torch(rng):
            raise ValueError("range must be uint32")
        rng = np.ascontiguousarray(rng, np.uint32)
    n = _numel(rng)
    own = out is None
    if own:
        if _is_torch(rng) and rng.is_cuda:
            import torch
            out = torch.empty((n, 3), dtype=torch.float64 if lut.dtype == np.float64 else torch.float32,
                              device=rng.device)
        else:
            out = np.empty((n, 3), lut.dtype)
    check(lib.ob_cartesian(lut._h, _ptr(rng), n, _ptr(out), st.h))
    if own and not (_is_torch(out) and out.is_cuda):
        st.sync()
    return out


def destagger(img, pixel_shift_by_row, inverse=False, out=None, stream=None, device=0):
    """destagger<T>(img, pixel_shift_by_row, inverse) for (H,W) or (H,W,...) images
    (ouster_core/include/ouster/core/impl/lidar_frame_impl.h:733-860)."""
    st = _stream(stream, device)
    shape = tuple(img.shape)
    if len(shape) < 2:
        raise ValueError("image must be at least 2-dimensional")
    h, w = shape[0], shape[1]
    k = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    sh = np.ascontiguousarray(pixel_shift_by_row, np.int32)
    own = out is None
    if own:
        if _is_torch(img):
            import torch
            out = torch.empty_like(img)
        else:
            img = np.ascontiguousarray(img)
            out = np.empty_like(img)
    check(lib.ob_destagger(_itemsize(img), k, _ptr(img), sh.ctypes.data, sh.size, h, w,
                           int(bool(inverse)), _ptr(out), st.h))
    if own and not (_is_torch(out) and out.is_cuda):
        st.sync()
    return out


def dewarp(points, poses, out=None, stream=None, device=0):
    """dewarp(points (H, W, 3), poses (W, 4, 4)) -> (H, W, 3): per-column pose application
    (python/src/cpp/client/processing.cpp:132-161; pose_util.h:37-59).  float32 or float64."""
    st = _stream(stream, device)
    dt = _np_dtype(points)
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("points must be float32 or float64")
    if not _is_torch(poses):
        poses = np.ascontiguousarray(poses, dt)
    n_points = _numel(points) // 3
    n_poses = _numel(poses) // 16
    if len(points.shape) == 3 and points.shape[1] != n_poses:
        raise RuntimeError("Number of points per set must match number of poses")
    own = out is None
    if own:
        if _is_torch(points):
            import torch
            out = torch.empty_like(points)
        else:
            points = np.ascontiguousarray(points)
            out = np.empty_like(points)
    status = lib.ob_dewarp(_capi.OB_F64 if dt == np.float64 else _capi.OB_F32, _ptr(points), _ptr(poses),
                           n_points, n_poses, _ptr(out), st.h)
    if status == _capi.OB_RUNTIME_ERROR:
        raise RuntimeError(lib.ob_last_error().decode())
    check(status)
    if own and not (_is_torch(out) and out.is_cuda):
        st.sync()
    return out


def dewarp_frame(lut, rng, poses, status, timestamps=None, min_range=0.0, max_range=float("inf"),
                 provenance=False, stream=None, out=None, out_count=None):
    """dewarp(lidar_frame, xyzlut, min_range, max_range) (pose_util.h:456-485): project the range
    image, apply each column's body_to_world pose, keep the points with min_range <= r <= max_range
    (metres) of the columns between the first and last valid one, in column-major order.
    Returns points [n, 3] (LUT dtype); with provenance=True also (col_idx u32 [n], timestamps u64 [n]).
    One fused GPU launch: nothing but the surviving points is written.
    Asynchronous device form: `out` = CUDA tensor [capacity, 3] of the LUT dtype and `out_count` = CUDA int64
    tensor [1] (inputs in device memory too): nothing waits for the GPU; returns (out, out_count), the count
    being written in stream order (a count above the capacity means the list was cut there)."""
    from ._capi import DewarpFrameIO
    st = _stream(stream, lut.device)
    n_px = lut.h * lut.w
    if _numel(rng) != n_px:
        raise ValueError("unexpected image dimensions")
    if max_range == float("inf"):
        max_range = 4294967.295
    rng = rng if _is_torch(rng) else np.ascontiguousarray(rng, np.uint32)
    poses = poses if _is_torch(poses) else np.ascontiguousarray(poses, np.float64)
    status = status if _is_torch(status) else np.ascontiguousarray(status, np.uint32)
    if _numel(poses) != lut.w * 16 or _numel(status) != lut.w:
        raise ValueError("poses must be [W, 4, 4] and status [W]")
    io = DewarpFrameIO()
    io.range, io.poses, io.status = _ptr(rng), _ptr(poses), _ptr(status)
    io.min_range, io.max_range = float(min_range), float(max_range)
    if out is not None:
        if out_count is None or provenance:
            raise ValueError("the asynchronous form takes out= and out_count= (device tensors) and no provenance")
        io.points, io.capacity = _ptr(out), _numel(out) // 3
        check(lib.ob_dewarp_frame(lut._h, C.byref(io), C.cast(_ptr(out_count), C.POINTER(C.c_size_t)), st.h))
        return out, out_count
    pts = np.empty((n_px, 3), lut.dtype)
    io.points, io.capacity = pts.ctypes.data, n_px
    ci = ts_out = None
    if provenance:
        if timestamps is None:
            raise ValueError("provenance needs the column timestamps")
        timestamps = timestamps if _is_torch(timestamps) else np.ascontiguousarray(timestamps, np.uint64)
        ci, ts_out = np.empty(n_px, np.uint32), np.empty(n_px, np.uint64)
        io.timestamps, io.col_idx, io.timestamps_out = _ptr(timestamps), ci.ctypes.data, ts_out.ctypes.data
    n = C.c_size_t(0)
    check(lib.ob_dewarp_frame(lut._h, C.byref(io), C.byref(n), st.h))
    if provenance:
        return pts[:n.value], ci[:n.value], ts_out[:n.value]
    return pts[:n.value]


DEFAULT_TARGET_DISTANCE_METER = 0.025                    # ouster/algorithm/normals.h:23
DEFAULT_MIN_ANGLE_INCIDENCE_RAD = 1 * np.pi / 180.0     # ouster/algorithm/normals.h:25


def normals(xyz, rng, *args, sensor_origins_xyz=None, pixel_search_range=1,
            min_angle_of_incidence_rad=DEFAULT_MIN_ANGLE_INCIDENCE_RAD,
            target_distance_m=DEFAULT_TARGET_DISTANCE_METER, vertical_subtent=0.0, return_subtent=False,
            stream=None, device=0):
    """algorithm.normals(xyz, range, sensor_origins_xyz, ...) and the dual-return form
    normals(xyz, range, xyz2, range2, sensor_origins_xyz, ...) (python binding of
    ouster_algorithm/include/ouster/algorithm/normals.h:58-108): destaggered (H, W, 3) points and
    (H, W) ranges in, (H, W, 3) unit normals out (a pair for dual returns).  numpy arrays or torch
    tensors (CUDA tensors stay on the device); float32 / float64.  RuntimeError texts as the reference."""
    from ._capi import NormalsIO
    pos = list(args)
    xyz2 = range2 = None
    # dual-return form: the two leading extra positionals are arrays (xyz2, range2); in the single-return
    # form the second one is the scalar pixel_search_range
    if len(pos) >= 2 and hasattr(pos[0], "shape") and hasattr(pos[1], "shape") and len(pos[1].shape) == 2:
        xyz2, range2 = pos[0], pos[1]
        pos = pos[2:]
    names = ["sensor_origins_xyz", "pixel_search_range", "min_angle_of_incidence_rad", "target_distance_m"]
    kw = {"sensor_origins_xyz": sensor_origins_xyz, "pixel_search_range": pixel_search_range,
          "min_angle_of_incidence_rad": min_angle_of_incidence_rad, "target_distance_m": target_distance_m}
    for nm, v in zip(names, pos):
        kw[nm] = v
    st = _stream(stream, device)
    if len(rng.shape) != 2:
        raise RuntimeError("normals: xyz dimensions mismatch")
    h, w = int(rng.shape[0]), int(rng.shape[1])
    dt = _np_dtype(xyz)
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        xyz = np.ascontiguousarray(xyz, np.float64)
        dt = np.dtype(np.float64)
    if _numel(xyz) != h * w * 3:
        raise RuntimeError("normals: xyz dimensions mismatch")
    dual = xyz2 is not None
    if dual:
        if _np_dtype(xyz2) != dt:
            xyz2 = np.ascontiguousarray(xyz2, dt)
        if _numel(xyz2) != h * w * 3:
            raise RuntimeError("normals: xyz dimensions mismatch")
        if tuple(range2.shape) != (h, w):
            raise RuntimeError("normals: range2 dimensions mismatch")
    org = kw["sensor_origins_xyz"]
    if org is None:
        raise TypeError("normals(): incompatible function arguments (sensor_origins_xyz is required)")
    if not _is_torch(org):
        org = np.ascontiguousarray(org, np.float64)
    if len(org.shape) != 2 or org.shape[1] != 3:
        raise TypeError("normals(): incompatible function arguments (sensor_origins_xyz must be (W, 3))")
    if org.shape[0] != w:
        raise RuntimeError("normals: sensor_origins size must match image width")

    def prep_rng(a):
        if _is_torch(a):
            return a
        return np.ascontiguousarray(a, np.uint32)

    def prep_xyz(a):
        return a if _is_torch(a) else np.ascontiguousarray(a)

    xyz, rng = prep_xyz(xyz), prep_rng(rng)
    on_dev = _is_torch(xyz) and xyz.is_cuda

    def new_out():
        if _is_torch(xyz):
            import torch
            return torch.empty((h, w, 3), dtype=xyz.dtype, device=xyz.device)
        return np.empty((h, w, 3), dt)

    n1 = new_out()
    io = NormalsIO()
    io.n_frames, io.h, io.w = 1, h, w
    io.xyz, io.range, io.normals = _ptr(xyz), _ptr(rng), _ptr(n1)
    n2 = None
    if dual:
        xyz2, range2 = prep_xyz(xyz2), prep_rng(range2)
        n2 = new_out()
        io.xyz2, io.range2, io.normals2 = _ptr(xyz2), _ptr(range2), _ptr(n2)
    io.sensor_origins_xyz, io.n_origins = _ptr(org), w
    io.pixel_search_range = int(kw["pixel_search_range"])
    io.min_angle_of_incidence_rad = float(kw["min_angle_of_incidence_rad"])
    io.target_distance_m = float(kw["target_distance_m"])
    io.vertical_subtent_rad = float(vertical_subtent)
    sub = np.zeros(1, np.float64)
    io.vertical_subtent_out = sub.ctypes.data
    status = lib.ob_normals(_capi.OB_F64 if dt == np.float64 else _capi.OB_F32, C.byref(io), st.h)
    if status == _capi.OB_RUNTIME_ERROR:
        raise RuntimeError(lib.ob_last_error().decode())
    check(status)
    if not on_dev or return_subtent:
        st.sync()
    res = (n1, n2) if dual else n1
    return (res, float(sub[0])) if return_subtent else res


def dewarp_frames(frames, min_range=0.0, max_range=float("inf"), provenance=False, stream=None):
    """dewarp(frame_set, xyzluts, min_range, max_range) (pose_util.h:475, impl/dewarp_impl.h:84-117):
    `frames` is a list with one entry per slot of the set -- None for an empty slot, else a dict
    {lut, range, poses, status[, timestamps]} -- and the result is the concatenation of the frames'
    dewarped points in slot order.  provenance=True also returns (frame_idx, col_idx, timestamps).
    One launch for the whole set."""
    from ._capi import DewarpFramesIO
    n = len(frames)
    ios = (DewarpFramesIO * max(n, 1))()
    keep, cap, lut0 = [], 0, None
    for i, fr in enumerate(frames):
        if fr is None:
            continue
        lut = fr["lut"]
        lut0 = lut0 or lut
        rng = fr["range"] if _is_torch(fr["range"]) else np.ascontiguousarray(fr["range"], np.uint32)
        poses = fr["poses"] if _is_torch(fr["poses"]) else np.ascontiguousarray(fr["poses"], np.float64)
        status = fr["status"] if _is_torch(fr["status"]) else np.ascontiguousarray(fr["status"], np.uint32)
        if _numel(rng) != lut.h * lut.w:
            raise ValueError("unexpected image dimensions")
        if _numel(poses) != lut.w * 16 or _numel(status) != lut.w:
            raise ValueError("poses must be [W, 4, 4] and status [W]")
        ts = fr.get("timestamps")
        if provenance:
            if ts is None:
                raise ValueError("provenance needs the column timestamps")
            ts = ts if _is_torch(ts) else np.ascontiguousarray(ts, np.uint64)
        keep += [rng, poses, status, ts]
        ios[i].lut, ios[i].range, ios[i].poses, ios[i].status = lut._h, _ptr(rng), _ptr(poses), _ptr(status)
        ios[i].timestamps = _ptr(ts) if provenance else None
        cap += lut.h * lut.w
    if lut0 is None:
        empty = np.empty((0, 3), np.float64)
        return (empty, np.empty(0, np.uint32), np.empty(0, np.uint32), np.empty(0, np.uint64)) if provenance else empty
    st = _stream(stream, lut0.device)
    if max_range == float("inf"):
        max_range = 4294967.295
    pts = np.empty((cap, 3), lut0.dtype)
    fi = ci = ts_out = None
    if provenance:
        fi, ci, ts_out = np.empty(cap, np.uint32), np.empty(cap, np.uint32), np.empty(cap, np.uint64)
    counts = (C.c_size_t * max(n, 1))()
    total = C.c_size_t(0)
    check(lib.ob_dewarp_frames(ios, n, float(min_range), float(max_range), pts.ctypes.data, cap,
                               fi.ctypes.data if provenance else None, ci.ctypes.data if provenance else None,
                               ts_out.ctypes.data if provenance else None, counts, C.byref(total), st.h))
    k = total.value
    if provenance:
        return pts[:k], fi[:k], ci[:k], ts_out[:k]
    return pts[:k]


def transform(points, pose, out=None, stream=None, device=0):
    """transform(points (..., 3), pose (4, 4)): one pose for every point (pose_util.h:118-131)."""
    pose = pose if _is_torch(pose) else np.ascontiguousarray(pose, _np_dtype(points)).reshape(1, 16)
    shape = tuple(points.shape)
    flat = points.reshape(-1, 3)
    res = dewarp(flat, pose, out=None if out is None else out.reshape(-1, 3), stream=stream, device=device)
    return res.reshape(shape)


def plan_scan_to_cloud(lut, pixel_shift_by_row, rng, xyz=None, range_destaggered=None,
                       xyz_destaggered=None, stream=None, poses=None):
    """Marshal an ob_scan_to_cloud call once and return a callable that launches it (`plan()`): for
    steady-state callers that process into the same buffers every step.  Arguments as scan_to_cloud."""
    st = _stream(stream, lut.device)
    F, R, H, W = tuple(rng.shape)
    io = CloudIO()
    io.n_frames, io.n_returns = F, R
    n = H * W
    io.range, io.range_frame_stride, io.range_return_stride = _ptr(rng), R * n, n
    if xyz is not None:
        io.xyz, io.xyz_frame_stride, io.xyz_return_stride = _ptr(xyz), R * n * 3, n * 3
    if range_destaggered is not None:
        io.range_destaggered, io.rd_frame_stride, io.rd_return_stride = _ptr(range_destaggered), R * n, n
    if xyz_destaggered is not None:
        io.xyz_destaggered, io.xd_frame_stride, io.xd_return_stride = _ptr(xyz_destaggered), R * n * 3, n * 3
    if poses is not None:
        if _np_dtype(poses) != lut.dtype:
            raise ValueError("poses must have the dtype of the lut")
        pn = _numel(poses)
        if pn == W * 16:
            io.poses, io.poses_frame_stride = _ptr(poses), 0
        elif pn == F * W * 16:
            io.poses, io.poses_frame_stride = _ptr(poses), W * 16
        else:
            raise ValueError("poses must be [W, 4, 4] or [F, W, 4, 4]")
    sh, nsh = None, 0
    if pixel_shift_by_row is not None:
        sh = np.ascontiguousarray(pixel_shift_by_row, np.int32)
        nsh = sh.size
    lut_h, sh_p, io_ref, st_h, run = lut._h, (sh.ctypes.data if sh is not None else None), C.byref(io), st.h, \
        lib.ob_scan_to_cloud

    def plan():
        check(run(lut_h, sh_p, nsh, io_ref, st_h))
    plan._keep = (lut, sh, io, st, rng, xyz, range_destaggered, xyz_destaggered, poses)
    return plan


def scan_to_cloud(lut, pixel_shift_by_row, rng, xyz=None, range_destaggered=None,
                  xyz_destaggered=None, stream=None, poses=None):
    """Fused batch: rng is [F, R, H, W] uint32; outputs [F, R, H*W, 3] (xyz), [F, R, H, W]
    (range_destaggered), [F, R, H, W, 3] (xyz_destaggered).  `poses` ([W, 4, 4] shared by all
    frames or [F, W, 4, 4], LUT dtype, e.g. LidarScan.body_to_world) fuses dewarp(xyz, poses) into
    the same pass.  Asynchronous on `stream`."""
    plan_scan_to_cloud(lut, pixel_shift_by_row, rng, xyz, range_destaggered, xyz_destaggered, stream, poses)()


class Decoder:
    """ob_decoder: device-side PacketFormat decode table.

    layout: dict with packet_header_size, col_header_size, channel_data_size, col_size,
            packet_size, columns_per_packet, pixels_per_column, columns_per_frame and the three
            column-header infos col_timestamp / col_measurement_id / col_status as
            (offset, mask, shift) tuples.
    fields: list of dicts {name, offset, mask, shift, elem_size, range_return, zero_pattern}.
    """

    def __init__(self, layout, fields, device=0):
        from ._capi import FieldDesc, PacketLayout
        L = PacketLayout()
        for k in ("packet_header_size", "col_header_size", "channel_data_size", "col_size",
                  "packet_size", "columns_per_packet", "pixels_per_column", "columns_per_frame"):
            setattr(L, k, int(layout[k]))
        for k in ("col_timestamp", "col_measurement_id", "col_status"):
            o, m, sft = layout[k]
            setattr(L, k, FieldDesc(int(o), 8, int(m), int(sft), -1, 0, 0))
        arr = (FieldDesc * max(len(fields), 1))()
        for i, f in enumerate(fields):
            arr[i] = FieldDesc(int(f["offset"]), int(f["elem_size"]), int(f["mask"]), int(f["shift"]),
                               int(f.get("range_return", -1)), int(f.get("zero_pattern", 0)), 0)
        h = C.c_void_p()
        check(lib.ob_decoder_create(C.byref(L), arr, len(fields), device, C.byref(h)))
        self._h, self.device = h, device
        self.layout, self.fields = dict(layout), [dict(f) for f in fields]
        self.h_px, self.w_px = int(layout["pixels_per_column"]), int(layout["columns_per_frame"])

    def decode(self, frames, lut=None, pixel_shift_by_row=None, stream=None):
        """frames: list of dicts {packets, n_slots, packet_stride, col_src (np.int32 [W] or None),
        fields: {name: array}, timestamp, measurement_id, status, xyz: [a0, a1],
        range_destaggered: [a0, a1]}.  Asynchronous on `stream`."""
        from ._capi import DecodeIO
        st = _stream(stream, self.device)
        ios = (DecodeIO * len(frames))()
        keep = []
        for i, fr in enumerate(frames):
            io = ios[i]
            io.packets = _ptr(fr["packets"])
            io.n_slots = int(fr["n_slots"])
            io.packet_stride = int(fr["packet_stride"])
            cs = fr.get("col_src")
            if cs is not None:
                cs = np.ascontiguousarray(cs, np.int32)
                keep.append(cs)
                io.col_src = cs.ctypes.data
            outs = fr.get("fields", {})
            for k, f in enumerate(self.fields):
                a = outs.get(f["name"])
                if a is not None:
                    io.fields[k] = _ptr(a)
            for k in ("timestamp", "measurement_id", "status"):
                if fr.get(k) is not None:
                    setattr(io, k, _ptr(fr[k]))
            for r, a in enumerate(fr.get("xyz", []) or []):
                if a is not None:
                    io.xyz[r] = _ptr(a)
            for r, a in enumerate(fr.get("range_destaggered", []) or []):
                if a is not None:
                    io.range_destaggered[r] = _ptr(a)
            if fr.get("lut") is not None:
                io.lut = fr["lut"]._h
        sh, nsh = None, 0
        if pixel_shift_by_row is not None:
            sh = np.ascontiguousarray(pixel_shift_by_row, np.int32)
            nsh = sh.size
        check(lib.ob_decode_frames(self._h, ios, len(frames), lut._h if lut is not None else None,
                                   sh.ctypes.data if sh is not None else None, nsh, st.h))

    @classmethod
    def from_sensor(cls, info, frame, device=0):
        """Decoder for the fields `frame` (host LidarFrame) shares with the sensor's PacketFormat."""
        L = info.layout
        layout = {k: getattr(L, k) for k in ("packet_header_size", "col_header_size", "channel_data_size",
                                             "col_size", "packet_size", "columns_per_packet",
                                             "pixels_per_column", "columns_per_frame")}
        for k in ("col_timestamp", "col_measurement_id", "col_status"):
            d = getattr(L, k)
            layout[k] = (d.offset, d.mask, d.shift)
        fields = []
        have = set(frame.fields)
        for name, tag, off, mask, shift, nel, vm in info.fields():
            if name not in have:
                continue
            a = frame.field(name)
            es = a.dtype.itemsize * (a.shape[2] if a.ndim == 3 else 1)
            fields.append({"name": name, "offset": off, "mask": mask, "shift": shift, "elem_size": es,
                           "range_return": {"RANGE": 0, "RANGE2": 1}.get(name, -1),
                           "zero_pattern": 0x7e00 if name == "RGB" else 0})
        return cls(layout, fields, device)

    def prepare_batch(self, n_frames, packets, n_slots, packet_stride, packets_frame_stride, fields,
                      lut=None, pixel_shift_by_row=None, xyz=None, range_destaggered=None,
                      timestamp=None, measurement_id=None, status=None, stream=None, frame_luts=None):
        """Build the ob_decode_batch descriptor of a uniformly strided batch ONCE and return a callable
        that launches it (`plan()`): a steady-state caller decodes into the same buffers every step, and
        re-marshalling ~100 pointers through ctypes costs more host time than the launch takes on the GPU.
        Arguments as decode_batch."""
        from ._capi import DecodeBatch
        st = _stream(stream, self.device)
        b = DecodeBatch()
        b.n_frames = n_frames
        b.packets, b.n_slots = _ptr(packets), n_slots
        b.packet_stride, b.packets_frame_stride = packet_stride, packets_frame_stride
        n_px = self.h_px * self.w_px
        for k, f in enumerate(self.fields):
            a = fields.get(f["name"])
            if a is not None:
                b.fields[k] = _ptr(a)
                b.field_frame_stride[k] = n_px * f["elem_size"]
        if timestamp is not None:
            b.timestamp, b.timestamp_frame_stride = _ptr(timestamp), self.w_px * 8
        if measurement_id is not None:
            b.measurement_id, b.measurement_id_frame_stride = _ptr(measurement_id), self.w_px * 2
        if status is not None:
            b.status, b.status_frame_stride = _ptr(status), self.w_px * 4
        any_lut = lut if lut is not None else (frame_luts[0] if frame_luts else None)
        esz = 8 if (any_lut is not None and any_lut.dtype == np.float64) else 4
        keep = [packets, fields, xyz, range_destaggered, timestamp, measurement_id, status, lut, frame_luts]
        if frame_luts:
            arr = (C.c_void_p * n_frames)(*[l._h.value for l in frame_luts])
            keep.append(arr)
            b.frame_luts = C.cast(arr, C.POINTER(C.c_void_p))
        for r, a in enumerate(xyz or []):
            if a is not None:
                b.xyz[r] = _ptr(a)
        b.xyz_frame_stride = n_px * 3 * esz
        for r, a in enumerate(range_destaggered or []):
            if a is not None:
                b.range_destaggered[r] = _ptr(a)
        b.rd_frame_stride = n_px * 4
        sh, nsh = None, 0
        if pixel_shift_by_row is not None:
            sh = np.ascontiguousarray(pixel_shift_by_row, np.int32)
            nsh = sh.size
        keep.append(sh)
        dec_h, b_ref, lut_h = self._h, C.byref(b), (lut._h if lut is not None else None)
        sh_p, st_h, run = (sh.ctypes.data if sh is not None else None), st.h, lib.ob_decode_batch_run

        def plan():
            check(run(dec_h, b_ref, lut_h, sh_p, nsh, st_h))
        plan._keep = (keep, b, st)
        return plan

    def decode_batch(self, n_frames, packets, n_slots, packet_stride, packets_frame_stride, fields,
                     lut=None, pixel_shift_by_row=None, xyz=None, range_destaggered=None,
                     timestamp=None, measurement_id=None, status=None, stream=None, frame_luts=None):
        """Uniformly strided batch of complete frames (ob_decode_batch_run).  `fields` maps a field
        name to an array/tensor shaped [n_frames, H, W(, k)]; xyz / range_destaggered are lists (one
        entry per return) of [n_frames, H*W, 3] / [n_frames, H, W] arrays."""
        self.prepare_batch(n_frames, packets, n_slots, packet_stride, packets_frame_stride, fields, lut,
                           pixel_shift_by_row, xyz, range_destaggered, timestamp, measurement_id, status,
                           stream, frame_luts)()

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            try:
                lib.ob_decoder_destroy(self._h)
            except Exception:
                pass
            self._h = None
