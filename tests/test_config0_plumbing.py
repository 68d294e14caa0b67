"""BASELINE configs[0]: OS1-64 1024x64 single return -- the reference's destagger() + cartesian()
on the CPU (oracle), checked against an independent numpy statement of the documented formulas
(python/src/ouster/sdk/examples/reference.py:18-76, 134-163); and the same case on the GPU."""
import numpy as np
import pytest

import __graft_entry__ as graft
from oracle import oracle as orc
from tests.helpers import default_os1_64, random_range


def numpy_doc_formula(info, rng):
    h, w = info["h"], info["w"]
    n = info["beam_to_lidar_transform"][0, 3]
    v = np.arange(w)
    te = 2.0 * np.pi * (1.0 - v / w)
    ta = -2.0 * np.pi * info["beam_azimuth_angles"] / 360.0
    phi = 2.0 * np.pi * info["beam_altitude_angles"] / 360.0
    r = rng.astype(np.float64)
    x = (r - n) * np.cos(te[None, :] + ta[:, None]) * np.cos(phi[:, None]) + n * np.cos(te[None, :])
    y = (r - n) * np.sin(te[None, :] + ta[:, None]) * np.cos(phi[:, None]) + n * np.sin(te[None, :])
    z = (r - n) * np.sin(phi[:, None])
    xyz = np.stack([x, y, z, np.ones_like(x)], -1) @ info["lidar_to_sensor_transform"].T
    xyz = xyz[..., :3] * 0.001
    xyz[rng == 0] = 0
    return xyz


def test_config0_oracle_matches_doc_formula_and_np_roll():
    info = default_os1_64(1024)
    h, w = info["h"], info["w"]
    rng = random_range(h, w, seed=42, p_zero=0.5, max_range=10000)   # benchmark_utils.h:102-107
    d, o = orc.make_xyz_lut(w, h, 0.001, info["beam_to_lidar_transform"], info["lidar_to_sensor_transform"],
                            info["beam_azimuth_angles"], info["beam_altitude_angles"])
    xyz = orc.cartesian(rng, d, o).reshape(h, w, 3)
    assert np.allclose(xyz, numpy_doc_formula(info, rng), rtol=1e-9, atol=1e-9)
    des = orc.destagger(rng, info["pixel_shift_by_row"])
    want = np.stack([np.roll(rng[u], info["pixel_shift_by_row"][u]) for u in range(h)])
    assert np.array_equal(des, want)
    # cartesian() (double) and cartesianT<float> agree to 1e-5 norm-wise (tests/cartesian_test.cpp:75-99)
    xf = orc.cartesian(rng, d.astype(np.float32), o.astype(np.float32)).reshape(h, w, 3)
    err = np.linalg.norm(xf - xyz, axis=-1)
    assert np.all(err <= 1e-5 * np.linalg.norm(xyz, axis=-1) + 1e-7)


@pytest.mark.gpu
def test_config0_gpu_matches_oracle():
    graft.build()
    ob = graft.load_package()
    info = default_os1_64(1024)
    h, w = info["h"], info["w"]
    rng = random_range(h, w, seed=42, p_zero=0.5, max_range=10000)
    for dtype in (np.float64, np.float32):
        lut = ob.XYZLutT.from_sensor_info(info, dtype=dtype)
        d, o = lut.direction, lut.offset
        xyz = lut(rng)
        assert np.array_equal(xyz, orc.cartesian(rng, d, o))
        assert np.allclose(xyz.reshape(h, w, 3), numpy_doc_formula(info, rng), rtol=1e-5, atol=1e-6)
    assert np.array_equal(ob.destagger(rng, info["pixel_shift_by_row"]),
                          orc.destagger(rng, info["pixel_shift_by_row"]))
