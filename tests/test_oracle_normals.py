"""Pins the normals oracle (oracle/orc_normals.c) to the reference's own known answers
(python/tests/test_normals.py:362-442, regression values written into the reference's test) and to
geometric properties (unit length, planar scenes as test_normals_cube_boundaries checks them)."""
import numpy as np
import pytest

from oracle import oracle as orc


def test_reference_known_answers_single_and_dual():
    # python/tests/test_normals.py:362-401 (test_normals)
    xyz = np.array([[[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]], [[0.0, 1.0, 0.0], [1.0, 1.0, 0.0]]])
    rng = np.array([[0.0, 1.0], [1.0, 1.0]], np.uint32)
    org = np.zeros((2, 3))
    want = np.array([[[0.0, 0.0, 0.0], [-1.0, 0.0, 0.0]], [[0.0, -1.0, 0.0], [-0.70710678, -0.70710678, 0.0]]])
    got = orc.normals(xyz, rng, sensor_origins_xyz=org, pixel_search_range=1, min_angle_of_incidence_rad=0.1,
                      target_distance_m=100)
    assert np.allclose(got, want)
    got1, got2 = orc.normals(xyz, rng, xyz, rng, org, 1, 0.1, 100)
    assert np.allclose(got1, want)


def test_reference_known_answers_isolated_pixel():
    # python/tests/test_normals.py:404-442 (test_normals_2)
    xyz = np.array([[[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]], [[0.0, 0.0, 0.0], [0.0, 0.0, 0.0]]])
    rng = np.array([[0.0, 1.0], [0.0, 0.0]], np.uint32)
    org = np.zeros((2, 3))
    want = np.array([[[0.0, 0.0, 0.0], [-1.0, 0.0, 0.0]], [[0.0, 0.0, 0.0], [0, 0.0, 0.0]]])
    assert np.allclose(orc.normals(xyz, rng, sensor_origins_xyz=org, pixel_search_range=1,
                                   min_angle_of_incidence_rad=0.1, target_distance_m=100), want)
    got1, _ = orc.normals(xyz, rng, xyz, rng, org, 1, 0.1, 100)
    assert np.allclose(got1, want)


def test_reference_error_texts():
    # python/tests/test_normals.py:212-359
    xyz = np.zeros((2, 2, 3))
    rng = np.ones((2, 2), np.uint32)
    with pytest.raises(RuntimeError, match="target_distance_m must be positive"):
        orc.normals(xyz, rng, sensor_origins_xyz=np.zeros((2, 3)), target_distance_m=-100)
    with pytest.raises(RuntimeError, match="min_angle_of_incidence_rad must be positive"):
        orc.normals(xyz, rng, sensor_origins_xyz=np.zeros((2, 3)), min_angle_of_incidence_rad=-0.1)
    with pytest.raises(RuntimeError, match="sensor_origins size must match image width"):
        orc.normals(xyz, rng, sensor_origins_xyz=np.zeros((0, 3)))
    with pytest.raises(RuntimeError, match="xyz dimensions mismatch"):
        orc.normals(xyz, np.ones((1, 2), np.uint32), sensor_origins_xyz=np.zeros((2, 3)))


def room_scene(h=64, w=512, half=40000.0, seed=0):
    """Sensor at the origin of a cube room (walls at +-half mm; 40 m keeps the 1 mm range quantisation below 0.2 deg): destaggered XYZ (metres) and range (mm)
    of a spinning lidar with +-22.5 deg vertical field of view."""
    alt = np.deg2rad(np.linspace(22.5, -22.5, h))[:, None]
    az = (2 * np.pi * (1 - np.arange(w) / w))[None, :]
    d = np.stack([np.cos(az) * np.cos(alt), np.sin(az) * np.cos(alt), np.sin(alt) * np.ones_like(az)], -1)
    t = half / np.max(np.abs(d), axis=-1)           # distance to the nearest wall along the beam
    rng = np.round(t).astype(np.uint32)
    xyz = d * rng[..., None] * 0.001
    return xyz, rng, d


def test_planar_room_normals_point_inwards_and_are_unit():
    xyz, rng, d = room_scene()
    n = orc.normals(xyz, rng, sensor_origins_xyz=np.zeros((xyz.shape[1], 3)))
    norms = np.linalg.norm(n, axis=-1)
    assert np.all(norms[rng > 0] > 0)
    assert np.allclose(norms[rng > 0], 1.0, atol=1e-9)
    # away from the wall edges the normal is the inward wall normal (test_normals_cube_boundaries: 0.5 deg)
    axis = np.argmax(np.abs(d), axis=-1)
    expected = np.zeros_like(d)
    np.put_along_axis(expected, axis[..., None], -np.sign(np.take_along_axis(d, axis[..., None], -1)), -1)
    cosang = np.sum(n * expected, -1)
    sorted_abs = np.sort(np.abs(d), axis=-1)
    interior = sorted_abs[..., 2] - sorted_abs[..., 1] > 0.05      # not next to an edge of the cube
    assert np.mean(cosang[interior] > np.cos(np.deg2rad(0.5))) > 0.99


def test_dual_return_shares_the_vertical_subtent():
    xyz, rng, _ = room_scene(32, 128)
    rs = np.random.default_rng(3)
    rng2 = np.where(rs.random(rng.shape) < 0.3, rng + 700, 0).astype(np.uint32)
    xyz2 = xyz * (rng2 / np.maximum(rng, 1))[..., None]
    org = np.zeros((128, 3))
    n1, n2 = orc.normals(xyz, rng, xyz2, rng2, org)
    sub = orc.normals_vertical_subtent(xyz, rng, org)
    a1, a2 = orc.normals(xyz, rng, xyz2, rng2, org, vertical_subtent=sub)
    assert np.array_equal(n1, a1) and np.array_equal(n2, a2)
    assert np.all(n2[rng2 == 0] == 0)
