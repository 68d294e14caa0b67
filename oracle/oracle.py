"""ctypes/numpy front-end of the CPU oracle (oracle/ouster_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libouster_oracle.so")

MAX_FIELDS = 24
NAME_LEN = 24

VOID, UINT8, UINT16, UINT32, UINT64, INT8, INT16, INT32, INT64, FLOAT32, FLOAT64, CHAR, FLOAT16 = range(13)
TYPE_NP = {UINT8: np.uint8, UINT16: np.uint16, UINT32: np.uint32, UINT64: np.uint64,
           INT8: np.int8, INT16: np.int16, INT32: np.int32, INT64: np.int64,
           FLOAT32: np.float32, FLOAT64: np.float64, FLOAT16: np.uint16}
NP_TYPE = {np.dtype(v): k for k, v in TYPE_NP.items() if k != FLOAT16}

PROFILES = {
    "LEGACY": 1, "RNG19_RFL8_SIG16_NIR16_DUAL": 2, "RNG19_RFL8_SIG16_NIR16": 3,
    "RNG15_RFL8_NIR8": 4, "FIVE_WORD_PIXEL": 5, "FUSA_RNG15_RFL8_NIR8_DUAL": 6,
    "RNG15_RFL8_NIR8_DUAL": 7, "RNG15_RFL8_NIR8_ZONE16": 8, "RNG19_RFL8_SIG16_NIR16_ZONE16": 9,
    "RNG15_RFL8_WIN8": 10, "RNG19_RFL8_SIG16_ZONE16_DUAL": 11, "RNG19_RFL8_SIG16_NIR16_RGB16": 12,
    "RNG19_RFL8_SIG16_NIR16_RGB16_DUAL": 13,
}
HEADER_STANDARD, HEADER_FUSA = 0, 1


class FieldInfo(C.Structure):
    _fields_ = [("ty_tag", C.c_int), ("offset", C.c_size_t), ("mask", C.c_uint64),
                ("shift", C.c_int), ("num_elements", C.c_int)]


class NamedField(C.Structure):
    _fields_ = [("name", C.c_char * NAME_LEN), ("info", FieldInfo)]


class CPacketFormat(C.Structure):
    _fields_ = [("profile", C.c_int), ("header_type", C.c_int),
                ("pixels_per_column", C.c_uint32), ("columns_per_packet", C.c_uint32),
                ("columns_per_frame", C.c_uint32),
                ("packet_header_size", C.c_size_t), ("col_header_size", C.c_size_t),
                ("channel_data_size", C.c_size_t), ("col_footer_size", C.c_size_t),
                ("packet_footer_size", C.c_size_t), ("col_size", C.c_size_t),
                ("lidar_packet_size", C.c_size_t), ("max_frame_id", C.c_uint32),
                ("n_fields", C.c_int), ("fields", NamedField * MAX_FIELDS),
                ("packet_type_info", FieldInfo), ("frame_id_info", FieldInfo),
                ("init_id_info", FieldInfo), ("prod_sn_info", FieldInfo),
                ("alert_flags_info", FieldInfo), ("countdown_thermal_shutdown_info", FieldInfo),
                ("countdown_shot_limiting_info", FieldInfo), ("thermal_shutdown_info", FieldInfo),
                ("shot_limiting_info", FieldInfo), ("col_status_info", FieldInfo),
                ("col_timestamp_info", FieldInfo), ("col_measurement_id_info", FieldInfo)]


class CFrameField(C.Structure):
    _fields_ = [("name", C.c_char * NAME_LEN), ("ty_tag", C.c_int), ("elem_size", C.c_size_t),
                ("data", C.POINTER(C.c_uint8))]


class CFrame(C.Structure):
    _fields_ = [("w", C.c_size_t), ("h", C.c_size_t), ("n_packets", C.c_size_t),
                ("frame_id", C.c_int64), ("frame_status", C.c_uint64),
                ("shutdown_countdown", C.c_uint8), ("shot_limiting_countdown", C.c_uint8),
                ("n_fields", C.c_int), ("fields", CFrameField * MAX_FIELDS),
                ("timestamp", C.POINTER(C.c_uint64)), ("measurement_id", C.POINTER(C.c_uint16)),
                ("status", C.POINTER(C.c_uint32)), ("packet_timestamp", C.POINTER(C.c_uint64)),
                ("alert_flags", C.POINTER(C.c_uint8))]


def build(force=False):
    """Compile the oracle with its Makefile (gcc); no-op when the .so is up to date."""
    srcs = [os.path.join(_HERE, f) for f in ("ouster_oracle.c", "orc_bench.c", "orc_normals.c", "ouster_oracle.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    vp, sz, i32, u32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint64
    PF = C.POINTER(CPacketFormat)
    FR = C.POINTER(CFrame)
    L.orc_field_info_make.argtypes = [sz, sz, sz, sz, sz, C.POINTER(FieldInfo)]
    L.orc_field_get.argtypes = [C.POINTER(FieldInfo), vp]
    L.orc_field_get.restype = u64
    L.orc_field_set.argtypes = [C.POINTER(FieldInfo), vp, u64]
    L.orc_field_set.restype = None
    L.orc_value_mask.argtypes = [C.POINTER(FieldInfo)]
    L.orc_value_mask.restype = u64
    L.orc_packet_format_init.argtypes = [PF, i32, i32, u32, u32, u32]
    L.orc_packet_format_set_fields.argtypes = [PF, C.POINTER(NamedField), i32, sz]
    L.orc_block_parsable.argtypes = [PF]
    L.orc_frame_id_difference.argtypes = [PF, u32, u32]
    L.orc_crc64.argtypes = [vp, sz]
    L.orc_crc64.restype = u64
    L.orc_default_field_type.argtypes = [i32, C.c_char_p]
    L.orc_block_field.argtypes = [PF, C.c_char_p, sz, vp, i32, vp, i32]
    L.orc_col_field.argtypes = [PF, C.c_char_p, sz, vp, vp, i32]
    L.orc_frame_create.argtypes = [PF, i32]
    L.orc_frame_create.restype = FR
    L.orc_frame_add_field.argtypes = [FR, C.c_char_p, i32]
    L.orc_frame_destroy.argtypes = [FR]
    L.orc_frame_destroy.restype = None
    L.orc_batcher_create.argtypes = [PF, u32, u32, u32]
    L.orc_batcher_create.restype = vp
    L.orc_batcher_destroy.argtypes = [vp]
    L.orc_batcher_destroy.restype = None
    L.orc_batcher_batch.argtypes = [vp, vp, sz, u64, FR]
    L.orc_batcher_reset.argtypes = [vp]
    L.orc_batcher_reset.restype = None
    L.orc_batcher_batched_packets.argtypes = [vp]
    L.orc_batcher_batched_packets.restype = sz
    L.orc_batcher_dropped_packets.argtypes = [vp]
    L.orc_batcher_dropped_packets.restype = sz
    L.orc_batcher_set_max_cache_size.argtypes = [vp, sz]
    L.orc_batcher_force_col_path.argtypes = [vp, i32]
    L.orc_batcher_force_col_path.restype = None
    L.orc_frame_to_packets.argtypes = [FR, PF, u32, u64, vp, vp]
    L.orc_destagger.argtypes = [sz, sz, vp, vp, sz, sz, sz, i32, vp]
    for n in ("orc_cartesian_f64", "orc_cartesian_f32", "orc_cartesian_f64_omp",
              "orc_cartesian_f32_omp"):
        getattr(L, n).argtypes = [vp, vp, vp, vp, sz]
        getattr(L, n).restype = None
    L.orc_make_xyz_lut.argtypes = [sz, sz, C.c_double, vp, vp, vp, sz, vp, sz, vp, vp]
    for n in ("orc_dewarp_f64", "orc_dewarp_f32"):
        getattr(L, n).argtypes = [vp, vp, vp, sz, sz]
        getattr(L, n).restype = None
    for n in ("orc_dewarp_frame_f64", "orc_dewarp_frame_f32"):
        getattr(L, n).argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, sz, C.c_double, C.c_double]
        getattr(L, n).restype = sz
    L.orc_snapshot_hash.argtypes = [vp, sz, sz]
    L.orc_snapshot_hash.restype = u64
    L.orc_normals_vertical_subtent.argtypes = [vp, vp, vp, sz, sz]
    L.orc_normals_vertical_subtent.restype = C.c_double
    L.orc_normals.argtypes = [vp, vp, vp, vp, sz, sz, vp, sz, C.c_double, C.c_double, C.c_double, vp, vp]
    L.orc_normals.restype = i32
    # orc_bench.c: CPU-baseline harness (bench.py only)
    L.orc_bench_max_threads.restype = i32
    L.orc_bench_k1.argtypes = [i32, i32, vp, sz, sz, sz, sz, vp, vp, vp, i32, i32, C.POINTER(C.c_double)]
    L.orc_bench_k1.restype = C.c_double
    L.orc_bench_k2.argtypes = [i32, i32, PF, vp, sz, sz, vp, vp, vp, i32, i32, C.POINTER(C.c_double)]
    L.orc_bench_k2.restype = C.c_double
    L.orc_pool_k1.argtypes = [i32, vp, sz, sz, sz, sz, vp, vp, vp, vp, vp]
    L.orc_pool_k1.restype = None
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PacketFormat:
    """orc_packet_format wrapper (PacketFormat, ouster_core/src/parsing.cpp:386-626)."""

    def __init__(self, profile, h, w, columns_per_packet=16, header_type=HEADER_STANDARD):
        if isinstance(profile, str):
            profile = PROFILES[profile]
        self.c = CPacketFormat()
        rc = lib().orc_packet_format_init(C.byref(self.c), profile, header_type, h,
                                          columns_per_packet, w)
        if rc != 0:
            raise ValueError("Unknown lidar udp profile / invalid packet format")

    def __getattr__(self, k):
        return getattr(self.c, k)

    @property
    def field_names(self):
        return [self.c.fields[i].name.decode() for i in range(self.c.n_fields)]

    def field_info(self, name):
        for i in range(self.c.n_fields):
            if self.c.fields[i].name.decode() == name:
                return self.c.fields[i].info
        raise KeyError(name)

    def set_fields(self, fields, channel_data_size):
        """fields: list of (name, ty_tag, offset, mask, shift[, num_elements])."""
        arr = (NamedField * len(fields))()
        for i, f in enumerate(fields):
            arr[i].name = f[0].encode()
            arr[i].info = FieldInfo(f[1], f[2], f[3], f[4], f[5] if len(f) > 5 else 1)
        lib().orc_packet_format_set_fields(C.byref(self.c), arr, len(fields), channel_data_size)

    def value_mask(self, name):
        return lib().orc_value_mask(C.byref(self.field_info(name)))

    def frame_id(self, buf):
        b = np.concatenate([np.frombuffer(bytes(buf), np.uint8), np.zeros(8, np.uint8)])
        return lib().orc_field_get(C.byref(self.c.frame_id_info), _ptr(b)) & 0xffffffff

    def header(self, info_name, buf):
        b = np.concatenate([np.frombuffer(bytes(buf), np.uint8), np.zeros(8, np.uint8)])
        return lib().orc_field_get(C.byref(getattr(self.c, info_name)), _ptr(b))

    def block_parsable(self):
        return lib().orc_block_parsable(C.byref(self.c))

    def frame_id_difference(self, cur, other):
        return lib().orc_frame_id_difference(C.byref(self.c), cur, other)


class Frame:
    """orc_frame wrapper: numpy views over the C-owned buffers."""

    def __init__(self, pf, with_window=True, extra_fields=()):
        self.pf = pf
        self.p = lib().orc_frame_create(C.byref(pf.c), int(with_window))
        for name, ty in extra_fields:
            lib().orc_frame_add_field(self.p, name.encode(), ty)

    def __del__(self):
        if getattr(self, "p", None) and _lib is not None:
            try:
                _lib.orc_frame_destroy(self.p)
            except Exception:
                pass
            self.p = None

    @property
    def c(self):
        return self.p.contents

    @property
    def w(self):
        return self.c.w

    @property
    def h(self):
        return self.c.h

    @property
    def frame_id(self):
        return self.c.frame_id

    @frame_id.setter
    def frame_id(self, v):
        self.c.frame_id = v

    @property
    def field_names(self):
        return [self.c.fields[i].name.decode() for i in range(self.c.n_fields)]

    def field(self, name):
        c = self.c
        for i in range(c.n_fields):
            f = c.fields[i]
            if f.name.decode() == name:
                n = c.h * c.w * f.elem_size
                raw = np.ctypeslib.as_array(f.data, shape=(n,))
                if name == "RGB":
                    return raw.view(np.uint16).reshape(c.h, c.w, 3)
                return raw.view(TYPE_NP[f.ty_tag]).reshape(c.h, c.w)
        raise KeyError(name)

    def has_field(self, name):
        return name in self.field_names

    def _arr(self, ptr, n):
        return np.ctypeslib.as_array(ptr, shape=(n,))

    @property
    def timestamp(self):
        return self._arr(self.c.timestamp, self.c.w)

    @property
    def measurement_id(self):
        return self._arr(self.c.measurement_id, self.c.w)

    @property
    def status(self):
        return self._arr(self.c.status, self.c.w)

    @property
    def packet_timestamp(self):
        return self._arr(self.c.packet_timestamp, self.c.n_packets)

    @property
    def alert_flags(self):
        return self._arr(self.c.alert_flags, self.c.n_packets)


class Batcher:
    """orc_batcher wrapper (FrameBatcher, ouster_core/src/lidar_frame.cpp:1248-1959)."""

    def __init__(self, pf, init_id=0, column_window=None):
        cw = column_window or (0, pf.columns_per_frame - 1)
        self.pf = pf
        self.p = lib().orc_batcher_create(C.byref(pf.c), init_id, cw[0], cw[1])
        if not self.p:
            raise ValueError("unexpected columns_per_packet/pixels_per_column: 0")

    def __del__(self):
        if getattr(self, "p", None) and _lib is not None:
            try:
                _lib.orc_batcher_destroy(self.p)
            except Exception:
                pass
            self.p = None

    def batch(self, buf, host_timestamp, frame):
        b = np.frombuffer(bytes(buf), np.uint8) if not isinstance(buf, np.ndarray) else buf
        rc = lib().orc_batcher_batch(self.p, _ptr(b), b.size, host_timestamp, frame.p)
        if rc == -1:
            raise ValueError("invalid argument")
        if rc == -2:
            raise RuntimeError("32-bit frame id did not increase since the last frame")
        if rc < 0:
            raise RuntimeError("oracle batch error %d" % rc)
        return bool(rc)

    def reset(self):
        lib().orc_batcher_reset(self.p)

    def force_col_path(self, on=True):
        lib().orc_batcher_force_col_path(self.p, int(on))

    @property
    def batched_packets(self):
        return lib().orc_batcher_batched_packets(self.p)

    @property
    def dropped_packets(self):
        return lib().orc_batcher_dropped_packets(self.p)

    def set_max_cache_size(self, n):
        if lib().orc_batcher_set_max_cache_size(self.p, n) != 0:
            raise ValueError("max_cache_size must be > 0")


def frame_to_packets(frame, pf, init_id=0, prod_sn=0):
    """-> (packets uint8 [n, lidar_packet_size], host_ts uint64 [n])."""
    n = frame.c.n_packets
    out = np.zeros((n, pf.lidar_packet_size), np.uint8)
    ts = np.zeros(n, np.uint64)
    k = lib().orc_frame_to_packets(frame.p, C.byref(pf.c), init_id, prod_sn, _ptr(out), _ptr(ts))
    if k < 0:
        raise ValueError("Mismatch between expected number of packets and PacketFormat.columns_per_packet")
    return out[:k].copy(), ts[:k].copy()


def destagger(img, shifts, inverse=False):
    """destagger<T>(img, pixel_shift_by_row, inverse); img is (H,W) or (H,W,k)."""
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    k = int(np.prod(img.shape[2:])) if img.ndim > 2 else 1
    sh = np.ascontiguousarray(shifts, dtype=np.int32)
    out = np.empty_like(img)
    rc = lib().orc_destagger(img.dtype.itemsize, k, _ptr(img), _ptr(sh), sh.size, h, w,
                             int(inverse), _ptr(out))
    if rc != 0:
        raise ValueError("image height does not match shifts size")
    return out


def cartesian(rng, direction, offset, omp=False):
    """cartesianT<T>: rng uint32 (H,W) or (N,), direction/offset (N,3) float32|float64."""
    rng = np.ascontiguousarray(rng, dtype=np.uint32)
    direction = np.ascontiguousarray(direction)
    offset = np.ascontiguousarray(offset)
    if rng.size != direction.shape[0]:
        raise ValueError("unexpected image dimensions")
    pts = np.empty_like(direction)
    name = {np.dtype(np.float32): "orc_cartesian_f32", np.dtype(np.float64): "orc_cartesian_f64"}[direction.dtype]
    if omp:
        name += "_omp"
    getattr(lib(), name)(_ptr(pts), _ptr(rng), _ptr(direction), _ptr(offset), rng.size)
    return pts


def make_xyz_lut(w, h, range_unit, beam_to_lidar, transform, az_deg, alt_deg):
    b2l = np.ascontiguousarray(beam_to_lidar, np.float64).reshape(16)
    tr = np.ascontiguousarray(transform, np.float64).reshape(16)
    az = np.ascontiguousarray(az_deg, np.float64)
    alt = np.ascontiguousarray(alt_deg, np.float64)
    d = np.empty((w * h, 3), np.float64)
    o = np.empty((w * h, 3), np.float64)
    rc = lib().orc_make_xyz_lut(w, h, range_unit, _ptr(b2l), _ptr(tr), _ptr(az), az.size,
                                _ptr(alt), alt.size, _ptr(d), _ptr(o))
    if rc == -1:
        raise ValueError("lut dimensions must be greater than zero")
    if rc == -2:
        raise ValueError("unexpected frame dimensions")
    return d, o


def dewarp(points, poses):
    """dewarp<T>(points (H,W,3)|(N,3), poses (W,4,4)|(W,16)) -- pose_util.h:37-59."""
    pts = np.ascontiguousarray(points)
    ps = np.ascontiguousarray(poses, dtype=pts.dtype).reshape(-1, 16)
    out = np.empty_like(pts)
    n = pts.size // 3
    name = {np.dtype(np.float32): "orc_dewarp_f32", np.dtype(np.float64): "orc_dewarp_f64"}[pts.dtype]
    getattr(lib(), name)(_ptr(out), _ptr(pts), _ptr(ps), n, ps.shape[0])
    return out


def dewarp_frame(rng, direction, offset, poses, status, timestamps, min_range, max_range):
    """dewarp<T>(LidarFrame, XYZLutT<T>, min_range, max_range) with provenance vectors --
    impl/dewarp_impl.h:22-76.  Returns (points [n,3], col_idx [n] u32, timestamps_ns [n] u64)."""
    rng = np.ascontiguousarray(rng, np.uint32)
    h, w = rng.shape
    d = np.ascontiguousarray(direction)
    o = np.ascontiguousarray(offset, d.dtype)
    ps = np.ascontiguousarray(poses, np.float64).reshape(w, 16)
    st = np.ascontiguousarray(status, np.uint32)
    ts = np.ascontiguousarray(timestamps, np.uint64)
    out = np.empty((h * w, 3), d.dtype)
    ci = np.empty(h * w, np.uint32)
    to = np.empty(h * w, np.uint64)
    name = {np.dtype(np.float32): "orc_dewarp_frame_f32", np.dtype(np.float64): "orc_dewarp_frame_f64"}[d.dtype]
    n = getattr(lib(), name)(_ptr(out), _ptr(ci), _ptr(to), _ptr(rng), _ptr(d), _ptr(o), _ptr(ps), _ptr(st),
                             _ptr(ts), h, w, float(min_range), float(max_range))
    return out[:n].copy(), ci[:n].copy(), to[:n].copy()


def snapshot_hash(a):
    a = np.ascontiguousarray(a)
    return lib().orc_snapshot_hash(_ptr(a), a.size, a.dtype.itemsize)


def crc64(buf):
    b = np.frombuffer(bytes(buf), np.uint8)
    return lib().orc_crc64(_ptr(b), b.size)


# ------------------------------------------------------------------------------------------------
# CPU-baseline harness (orc_bench.c) -- bench.py's cpu_baseline / --impl reference legs only
# ------------------------------------------------------------------------------------------------
BENCH_MODES = {"as_shipped": 0, "ouster_omp": 1, "thread_per_stream": 2}


def bench_max_threads():
    return int(lib().orc_bench_max_threads())


def bench_k1(mode, rng, shifts, direction, offset, threads=0, reps=1):
    """Seconds for `reps` passes of destagger<u32>() + cartesian() over rng [F, R, H, W] (uint32), run
    from C in one of BENCH_MODES; the LUT dtype (float32 / float64) selects cartesianT<float|double>."""
    rng = np.ascontiguousarray(rng, np.uint32)
    F, R, h, w = rng.shape
    f64 = int(direction.dtype == np.float64)
    d = np.ascontiguousarray(direction)
    o = np.ascontiguousarray(offset, d.dtype)
    sh = np.ascontiguousarray(shifts, np.int32)
    sink = C.c_double(0)
    return float(lib().orc_bench_k1(BENCH_MODES[mode], f64, _ptr(rng), F, R, h, w, _ptr(sh), _ptr(d), _ptr(o),
                                    int(threads), int(reps), C.byref(sink)))


def bench_k2(mode, pf, packets, shifts, direction, offset, threads=0, reps=1):
    """Seconds for `reps` passes of FrameBatcher decode + destagger + cartesian over packets
    [F, n_packets, packet_size] (uint8 wire bytes of complete frames)."""
    pk = np.ascontiguousarray(packets, np.uint8)
    F, n_pk, psz = pk.shape
    assert psz == pf.lidar_packet_size
    f64 = int(direction.dtype == np.float64)
    d = np.ascontiguousarray(direction)
    o = np.ascontiguousarray(offset, d.dtype)
    sh = np.ascontiguousarray(shifts, np.int32)
    sink = C.c_double(0)
    return float(lib().orc_bench_k2(BENCH_MODES[mode], f64, C.byref(pf.c), _ptr(pk), F, n_pk, _ptr(sh), _ptr(d),
                                    _ptr(o), int(threads), int(reps), C.byref(sink)))


def pool_k1(rng, shifts, direction, offset):
    """cartesianT + destagger<u32> of a whole pool rng [F, R, H, W] -> (xyz [F, R, H*W, 3], rd [F, R, H, W]);
    every frame computed by the single-thread oracle functions, frames spread over the host cores."""
    rng = np.ascontiguousarray(rng, np.uint32)
    F, R, h, w = rng.shape
    d = np.ascontiguousarray(direction)
    o = np.ascontiguousarray(offset, d.dtype)
    sh = np.ascontiguousarray(shifts, np.int32)
    xyz = np.empty((F, R, h * w, 3), d.dtype)
    rd = np.empty((F, R, h, w), np.uint32)
    lib().orc_pool_k1(int(d.dtype == np.float64), _ptr(rng), F, R, h, w, _ptr(sh), _ptr(d), _ptr(o),
                      _ptr(xyz), _ptr(rd))
    return xyz, rd


# ------------------------------------------------------------------------------------------------
# surface normals (orc_normals.c) -- ouster_algorithm/src/normals.cpp
# ------------------------------------------------------------------------------------------------
DEFAULT_MIN_ANGLE_INCIDENCE_RAD = 1 * np.pi / 180.0   # normals.h:25
DEFAULT_TARGET_DISTANCE_METER = 0.025                  # normals.h:23


def normals_vertical_subtent(xyz, rng, sensor_origins_xyz):
    h, w = rng.shape
    x = np.ascontiguousarray(xyz, np.float64).reshape(h * w, 3)
    r = np.ascontiguousarray(rng, np.uint32)
    o = np.ascontiguousarray(sensor_origins_xyz, np.float64)
    return float(lib().orc_normals_vertical_subtent(_ptr(x), _ptr(r), _ptr(o), h, w))


def normals(xyz, rng, xyz2=None, range2=None, sensor_origins_xyz=None, pixel_search_range=1,
            min_angle_of_incidence_rad=DEFAULT_MIN_ANGLE_INCIDENCE_RAD,
            target_distance_m=DEFAULT_TARGET_DISTANCE_METER, vertical_subtent=0.0):
    """normals(xyz, range[, xyz2, range2], sensor_origins_xyz, ...) with the reference's error texts
    (RuntimeError).  Returns (H, W, 3) or a pair of them."""
    r = np.ascontiguousarray(rng, np.uint32)
    if r.ndim != 2:
        raise RuntimeError("normals: xyz dimensions mismatch")
    h, w = r.shape
    x = np.ascontiguousarray(xyz, np.float64)
    if x.size != h * w * 3:
        raise RuntimeError("normals: xyz dimensions mismatch")
    dual = xyz2 is not None
    x2 = r2 = None
    if dual:
        x2 = np.ascontiguousarray(xyz2, np.float64)
        r2 = np.ascontiguousarray(range2, np.uint32)
        if x2.size != h * w * 3:
            raise RuntimeError("normals: xyz dimensions mismatch")
        if r2.shape != (h, w):
            raise RuntimeError("normals: range2 dimensions mismatch")
    o = np.ascontiguousarray(sensor_origins_xyz, np.float64)
    if o.ndim != 2 or o.shape[0] != w or o.shape[1] != 3:
        raise RuntimeError("normals: sensor_origins size must match image width")
    n1 = np.zeros((h, w, 3), np.float64)
    n2 = np.zeros((h, w, 3), np.float64) if dual else None
    rc = lib().orc_normals(_ptr(x), _ptr(r), _ptr(x2) if dual else None, _ptr(r2) if dual else None, h, w,
                           _ptr(o), int(pixel_search_range), float(min_angle_of_incidence_rad),
                           float(target_distance_m), float(vertical_subtent), _ptr(n1),
                           _ptr(n2) if dual else None)
    if rc == -1:
        raise RuntimeError("normals: target_distance_m must be positive")
    if rc == -2:
        raise RuntimeError("normals: min_angle_of_incidence_rad must be positive")
    return (n1, n2) if dual else n1
