"""Shared helpers for the test-suite (fixture loading, synthetic inputs)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

PCAP_FIXTURES = [
    "OS-0-128-U1_v2.3.0_1024x10",
    "OS-0-32-U1_v2.2.0_1024x10",
    "OS-1-128_767798045_1024x10_20230712_120049",
    "OS-2-128-U1_v2.3.0_1024x10",
    "OS-2-32-U0_v2.0.0_1024x10",
    "OS-1-32-G_v2.1.1_1024x10",
]


def load_fixture(name):
    meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
    packets = np.load(os.path.join(GOLDEN, name + ".npz"))["packets"]
    return meta, packets


def fw_has_window(meta):
    """WINDOW is in the default field set only for fw >= 3.2 (lidar_frame.cpp:1097-1110);
    every committed fixture is older."""
    return False


def random_range(h, w, seed, p_zero=0.5, max_range=(1 << 19) - 1):
    """Synthetic range image as tests/benchmarks/benchmark_utils.h:93-110 draws it:
    ~p_zero zeros, valid returns uniform in [1, max_range]."""
    rng = np.random.default_rng(seed)
    r = rng.integers(1, max_range + 1, size=(h, w), dtype=np.uint32)
    r[rng.random((h, w)) < p_zero] = 0
    return r


def random_lut(n, seed, dtype=np.float32):
    """Random LUT as tests/benchmarks/benchmark_utils.h:112-126: dir in U(0.5,1.5), off in U(0,0.01)."""
    rng = np.random.default_rng(seed)
    d = (rng.random((n, 3)) + 0.5).astype(dtype)
    o = (rng.random((n, 3)) * 0.01).astype(dtype)
    return d, o
