"""Builds libouster_b200.so (CUDA kernels + C ABI + host mirror) in-tree with nvcc for sm_100a.

    python ouster-sdk_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box.
"""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libouster_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "-shared",
    "-DOB_BUILD",
]


def sources():
    cu = sorted(glob.glob(os.path.join(PKG, "csrc", "*.cu")))
    cpp = sorted(glob.glob(os.path.join(PKG, "host", "*.cpp")))
    return cu + cpp


def headers():
    return (glob.glob(os.path.join(PKG, "csrc", "*.h")) + glob.glob(os.path.join(PKG, "csrc", "*.cuh"))
            + glob.glob(os.path.join(PKG, "host", "*.h"))
            + glob.glob(os.path.join(ROOT, "include", "*.h"))
            + glob.glob(os.path.join(ROOT, "include", "ouster", "core", "*.h"))
            + glob.glob(os.path.join(ROOT, "include", "ouster", "core", "impl", "*.h")))


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in sources() + headers() + [__file__])


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = os.environ.get("OB_EXTRA_NVCC_FLAGS", "").split()   # experiments only (e.g. -DOB_K2_STREAM_STORES)
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "csrc"),
                                 "-I", os.path.join(PKG, "host"), "-o", LIB] + sources()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
