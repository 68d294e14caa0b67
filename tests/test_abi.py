"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ouster_b200.h declares, and fails loudly (no CPU fallback) without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as graft

ROOT = graft.ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ouster_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ob_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    graft.build()
    ob = graft.load_package()
    lib = ctypes.CDLL(ob._capi.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.ob_abi_version() == 1


def test_product_does_not_reference_oracle():
    # the oracle is test infrastructure: nothing under ouster-sdk_b200/ may import/link it
    pkg = os.path.join(ROOT, "ouster-sdk_b200")
    for dp, _, fn in os.walk(pkg):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "ouster_oracle" not in src and "from oracle" not in src, os.path.join(dp, f)


def test_no_cpu_fallback_without_device():
    ob = graft.load_package()
    if ob.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(Exception) as ei:
        ob.Stream(0)
    assert "no CUDA device" in str(ei.value)
    with pytest.raises(Exception):
        ob.XYZLutT.from_arrays(np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32), 2, 2)


def test_argument_validation_messages_without_device():
    # texts follow the reference's exceptions (xyzlut.cpp:15,20)
    ob = graft.load_package()
    ident = np.eye(4)
    with pytest.raises(ValueError, match="lut dimensions must be greater than zero"):
        ob.XYZLutT.from_intrinsics(0, 4, 0.001, ident, ident, [0] * 4, [0] * 4)
    with pytest.raises(ValueError, match="unexpected frame dimensions"):
        ob.XYZLutT.from_intrinsics(8, 4, 0.001, ident, ident, [0] * 3, [0] * 4)


def test_ctypes_struct_layouts_match_the_c_abi():
    """The Python binding's structs must have exactly the C sizes (guards against ABI drift)."""
    ob = graft.load_package()
    capi = ob._capi
    pairs = {"ob_cloud_io": capi.CloudIO, "ob_field_desc": capi.FieldDesc,
             "ob_packet_layout": capi.PacketLayout, "ob_decode_io": capi.DecodeIO,
             "ob_decode_batch": capi.DecodeBatch, "ob_dewarp_frame_io": capi.DewarpFrameIO}
    for name, cls in pairs.items():
        assert capi.lib.ob_abi_sizeof(name.encode()) == ctypes.sizeof(cls), name
    assert capi.lib.ob_abi_sizeof(b"nope") == 0


def test_c_headers_are_plain_c99(tmp_path):
    """include/ouster_b200.h and include/ouster_b200_host.h are a C ABI: a C99 translation unit that
    includes both compiles with -pedantic, links against the library and runs (no CUDA call)."""
    import subprocess
    graft.build()
    src = tmp_path / "cabi.c"
    src.write_text('#include "ouster_b200.h"\n#include "ouster_b200_host.h"\n'
                   "int main(void) { ob_cloud_io io; ob_decode_io d; ob_dewarp_frame_io w; obh_slot s;\n"
                   "  (void)io; (void)d; (void)w; (void)s; return ob_abi_version() == OB_ABI_VERSION ? 0 : 1; }\n")
    lib_dir = os.path.join(graft.ROOT, "ouster-sdk_b200", "lib")
    exe = tmp_path / "cabi"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I",
                           os.path.join(graft.ROOT, "include"), str(src), "-L", lib_dir, "-louster_b200",
                           f"-Wl,-rpath,{lib_dir}", "-o", str(exe)])
    assert subprocess.run([str(exe)]).returncode == 0
