"""ctypes declarations of the C ABI (include/ouster_b200.h).  Loading this module loads
libouster_b200.so; it raises ImportError when the library has not been built -- there is no
Python/CPU fallback for the compute path."""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libouster_b200.so")

OB_MAX_FIELDS = 24
OB_MAX_RETURNS = 2
OB_OK, OB_INVALID_ARGUMENT, OB_RUNTIME_ERROR, OB_CUDA_ERROR, OB_NO_DEVICE = range(5)
OB_F32, OB_F64 = 0, 1

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python ouster-sdk_b200/build.py` "
        "(__graft_entry__.build()); the B200 path has no fallback implementation")

lib = C.CDLL(LIB_PATH)

vp, sz, i32, u32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint64


class CloudIO(C.Structure):
    _fields_ = [("n_frames", u32), ("n_returns", u32),
                ("range", vp), ("range_frame_stride", sz), ("range_return_stride", sz),
                ("xyz", vp), ("xyz_frame_stride", sz), ("xyz_return_stride", sz),
                ("range_destaggered", vp), ("rd_frame_stride", sz), ("rd_return_stride", sz),
                ("xyz_destaggered", vp), ("xd_frame_stride", sz), ("xd_return_stride", sz),
                ("poses", vp), ("poses_frame_stride", sz)]


class DewarpFrameIO(C.Structure):
    _fields_ = [("range", vp), ("poses", vp), ("status", vp), ("timestamps", vp),
                ("min_range", C.c_double), ("max_range", C.c_double),
                ("points", vp), ("col_idx", vp), ("timestamps_out", vp), ("capacity", sz)]


class NormalsIO(C.Structure):
    _fields_ = [("n_frames", sz), ("h", sz), ("w", sz), ("xyz", vp), ("range", vp), ("xyz2", vp), ("range2", vp),
                ("normals", vp), ("normals2", vp), ("xyz_frame_stride", sz), ("range_frame_stride", sz),
                ("normals_frame_stride", sz), ("sensor_origins_xyz", vp), ("n_origins", sz),
                ("origins_frame_stride", sz), ("pixel_search_range", sz),
                ("min_angle_of_incidence_rad", C.c_double), ("target_distance_m", C.c_double),
                ("vertical_subtent_rad", C.c_double), ("vertical_subtent_out", vp)]


class DewarpFramesIO(C.Structure):
    _fields_ = [("lut", vp), ("range", vp), ("poses", vp), ("status", vp), ("timestamps", vp)]


class FieldDesc(C.Structure):
    _fields_ = [("offset", u32), ("elem_size", u32), ("mask", u64), ("shift", C.c_int32),
                ("range_return", C.c_int32), ("zero_pattern", u32), ("reserved", u32)]


class PacketLayout(C.Structure):
    _fields_ = [("packet_header_size", u32), ("col_header_size", u32), ("channel_data_size", u32),
                ("col_size", u32), ("packet_size", u32), ("columns_per_packet", u32),
                ("pixels_per_column", u32), ("columns_per_frame", u32),
                ("col_timestamp", FieldDesc), ("col_measurement_id", FieldDesc),
                ("col_status", FieldDesc)]


class DecodeIO(C.Structure):
    _fields_ = [("packets", vp), ("n_slots", sz), ("packet_stride", sz),
                ("col_src", vp),
                ("fields", vp * OB_MAX_FIELDS),
                ("timestamp", vp), ("measurement_id", vp), ("status", vp),
                ("xyz", vp * OB_MAX_RETURNS), ("range_destaggered", vp * OB_MAX_RETURNS),
                ("lut", vp)]


class DecodeBatch(C.Structure):
    _fields_ = [("n_frames", u32), ("packets", vp), ("n_slots", sz), ("packet_stride", sz),
                ("packets_frame_stride", sz),
                ("fields", vp * OB_MAX_FIELDS), ("field_frame_stride", sz * OB_MAX_FIELDS),
                ("timestamp", vp), ("measurement_id", vp), ("status", vp),
                ("timestamp_frame_stride", sz), ("measurement_id_frame_stride", sz),
                ("status_frame_stride", sz),
                ("xyz", vp * OB_MAX_RETURNS), ("xyz_frame_stride", sz),
                ("range_destaggered", vp * OB_MAX_RETURNS), ("rd_frame_stride", sz),
                ("frame_luts", C.POINTER(vp))]


class EncodeIO(C.Structure):
    _fields_ = [("fields", vp * OB_MAX_FIELDS), ("timestamp", vp), ("status", vp), ("packet_headers", vp),
                ("packet_header_bytes", sz), ("packets", vp), ("packet_stride", sz)]


def _sig(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


_sig("ob_abi_version", i32)
_sig("ob_abi_sizeof", sz, C.c_char_p)
_sig("ob_last_error", C.c_char_p)
_sig("ob_device_count", i32)
_sig("ob_kernel_launch_count", u64)
_sig("ob_kernel_launch_count_of", u64, C.c_char_p)
_sig("ob_set_tunable", i32, i32, C.c_char_p, i32)
_sig("ob_stream_create", i32, i32, C.POINTER(vp))
_sig("ob_stream_wrap", i32, i32, vp, C.POINTER(vp))
_sig("ob_stream_sync", i32, vp)
_sig("ob_stream_cuda_handle", vp, vp)
_sig("ob_stream_destroy", i32, vp)
_sig("ob_host_alloc", i32, sz, C.POINTER(vp))
_sig("ob_host_free", i32, vp)
_sig("ob_lut_create", i32, i32, vp, vp, sz, sz, i32, C.POINTER(vp))
_sig("ob_lut_from_intrinsics", i32, i32, sz, sz, C.c_double, vp, vp, vp, sz, vp, sz, i32,
     C.POINTER(vp))
_sig("ob_lut_download", i32, vp, vp, vp)
_sig("ob_lut_info", i32, vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(i32), C.POINTER(i32))
_sig("ob_lut_device_ptrs", i32, vp, C.POINTER(vp), C.POINTER(vp))
_sig("ob_lut_destroy", i32, vp)
_sig("ob_lut_set_analytic", i32, vp, i32)
_sig("ob_lut_is_analytic", i32, vp)
_sig("ob_cartesian", i32, vp, vp, sz, vp, vp)
_sig("ob_destagger", i32, sz, sz, vp, vp, sz, sz, sz, i32, vp, vp)
_sig("ob_dewarp", i32, i32, vp, vp, sz, sz, vp, vp)
_sig("ob_scan_to_cloud", i32, vp, vp, sz, C.POINTER(CloudIO), vp)
_sig("ob_dewarp_frame", i32, vp, C.POINTER(DewarpFrameIO), C.POINTER(sz), vp)
_sig("ob_normals", i32, i32, C.POINTER(NormalsIO), vp)
_sig("ob_dewarp_frames", i32, C.POINTER(DewarpFramesIO), sz, C.c_double, C.c_double, vp, sz, vp, vp, vp,
     C.POINTER(sz), C.POINTER(sz), vp)
if hasattr(lib, "ob_decoder_create"):
    _sig("ob_decoder_create", i32, C.POINTER(PacketLayout), C.POINTER(FieldDesc), sz, i32,
         C.POINTER(vp))
    _sig("ob_decoder_destroy", i32, vp)
    _sig("ob_decode_frames", i32, vp, C.POINTER(DecodeIO), sz, vp, vp, sz, vp)
    _sig("ob_decode_batch_run", i32, vp, C.POINTER(DecodeBatch), vp, vp, sz, vp)
    _sig("ob_encode_frames", i32, vp, C.POINTER(EncodeIO), sz, i32, vp)


class OusterB200Error(RuntimeError):
    pass


def check(status):
    """Map an ob_status to the exception type the reference raises through its Python binding
    (std::invalid_argument -> ValueError, python/tests/test_xyzlut.py:29-51)."""
    if status == OB_OK:
        return
    msg = lib.ob_last_error().decode()
    if status == OB_INVALID_ARGUMENT:
        raise ValueError(msg)
    raise OusterB200Error(f"[status {status}] {msg}")
