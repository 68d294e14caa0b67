"""GPU parity tests for K1 (fused range -> XYZ + destagger) and the stand-alone cartesian /
destagger entry points, through the C ABI, against the CPU oracle.  Bit-exact for integers
and -- because the kernel keeps the reference's separate multiply/add roundings -- for the
float and double coordinates as well (tolerance 0; north_star allows 1e-5 relative)."""
import os

import numpy as np
import pytest

import __graft_entry__ as graft
from oracle import oracle as orc
from tests.helpers import GOLDEN, load_fixture, random_lut, random_range

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    graft.build()
    m = graft.load_package()
    assert m.device_count() > 0
    return m


SHAPES = [(32, 512), (64, 1024), (128, 1024), (128, 2048), (16, 64), (8, 36)]


@pytest.mark.parametrize("h,w", SHAPES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cartesian_matches_oracle(ob, h, w, dtype):
    rng = random_range(h, w, seed=h * w)
    d, o = random_lut(h * w, seed=3, dtype=dtype)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    got = lut(rng)
    want = orc.cartesian(rng, d, o)
    assert got.dtype == dtype and got.shape == (h * w, 3)
    assert np.array_equal(got, want)
    # r == 0 gives exactly +0.0 in all components (impl/cartesian.h:58-59)
    z = rng.reshape(-1) == 0
    assert not np.any(got[z]) and not np.any(np.signbit(got[z]))


def test_cartesian_dimension_error(ob):
    d, o = random_lut(32 * 64, 1)
    lut = ob.XYZLutT.from_arrays(d, o, 32, 64)
    with pytest.raises(ValueError, match="unexpected image dimensions"):
        lut(np.zeros((32, 65), np.uint32))


@pytest.mark.parametrize("h,w", [(32, 512), (128, 1024), (128, 2048), (7, 33), (64, 100)])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64])
def test_destagger_matches_oracle(ob, h, w, dtype):
    rs = np.random.default_rng(h + w)
    if np.issubdtype(dtype, np.floating):
        img = rs.random((h, w)).astype(dtype)
    else:
        img = rs.integers(0, np.iinfo(dtype).max, size=(h, w), dtype=dtype)
    pow2 = (w & (w - 1)) == 0
    lo = -30 if pow2 else 0  # negative shifts on non power-of-two widths: reference quirk, see DESIGN.md
    shifts = rs.integers(lo, 31, size=h).astype(np.int32)
    got = ob.destagger(img, shifts)
    assert np.array_equal(got, orc.destagger(img, shifts))
    back = ob.destagger(got, shifts, inverse=True)
    assert np.array_equal(back, img)


def test_destagger_nd_and_np_roll(ob):
    rs = np.random.default_rng(5)
    img = rs.random((64, 512, 3))
    shifts = rs.integers(-64, 65, size=64).astype(np.int32)
    got = ob.destagger(img, shifts)
    want = np.stack([np.roll(img[u], shifts[u], axis=0) for u in range(64)])
    assert np.array_equal(got, want)
    assert np.array_equal(got, orc.destagger(img, shifts))


def test_destagger_reference_py_vectors(ob):
    z = np.load(os.path.join(GOLDEN, "destagger_reference.npz"))
    for i in range(2):
        img, shifts, out = z[f"c{i}/img"], z[f"c{i}/shifts"], z[f"c{i}/out"]
        assert np.array_equal(ob.destagger(img, shifts), out)


def test_destagger_errors(ob):
    with pytest.raises(ValueError, match="image height does not match shifts size"):
        ob.destagger(np.zeros((4, 8), np.uint32), [0, 0, 0])


@pytest.mark.parametrize("h,w", [(32, 512), (64, 1024), (128, 2048), (16, 64)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("aligned_shifts", [True, False])
@pytest.mark.parametrize("n_returns", [1, 2])
def test_scan_to_cloud_matches_oracle(ob, h, w, dtype, aligned_shifts, n_returns):
    F = 3
    rs = np.random.default_rng(h * 7 + w + n_returns)
    rng = np.stack([np.stack([random_range(h, w, seed=100 * f + r, p_zero=0.5 if r == 0 else 0.8)
                              for r in range(n_returns)]) for f in range(F)])
    d, o = random_lut(h * w, seed=11, dtype=dtype)
    shifts = rs.integers(-30, 31, size=h).astype(np.int32)
    if aligned_shifts:
        shifts = (shifts // 4) * 4
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    xyz = np.zeros((F, n_returns, h * w, 3), dtype)
    rd = np.zeros((F, n_returns, h, w), np.uint32)
    xd = np.zeros((F, n_returns, h, w, 3), dtype)
    st = ob.Stream(0)
    ob.scan_to_cloud(lut, shifts, rng, xyz=xyz, range_destaggered=rd, xyz_destaggered=xd, stream=st)
    st.sync()
    for f in range(F):
        for r in range(n_returns):
            want = orc.cartesian(rng[f, r], d, o)
            assert np.array_equal(xyz[f, r], want), (f, r)
            assert np.array_equal(rd[f, r], orc.destagger(rng[f, r], shifts)), (f, r)
            assert np.array_equal(xd[f, r], orc.destagger(want.reshape(h, w, 3), shifts)), (f, r)


def test_scan_to_cloud_generic_fallback_shapes(ob):
    # widths that are not a multiple of 4 take the generic kernel
    h, w = 9, 35
    rng = random_range(h, w, 1)[None, None]
    d, o = random_lut(h * w, 2)
    shifts = np.arange(h, dtype=np.int32)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    xyz = np.zeros((1, 1, h * w, 3), np.float32)
    rd = np.zeros((1, 1, h, w), np.uint32)
    st = ob.Stream(0)
    ob.scan_to_cloud(lut, shifts, rng, xyz=xyz, range_destaggered=rd, stream=st)
    st.sync()
    assert np.array_equal(xyz[0, 0], orc.cartesian(rng[0, 0], d, o))
    assert np.array_equal(rd[0, 0], orc.destagger(rng[0, 0], shifts))


def test_scan_to_cloud_device_tensors(ob):
    torch = pytest.importorskip("torch")
    h, w, F = 128, 2048, 4
    rng = np.stack([np.stack([random_range(h, w, 10 * f + r) for r in range(2)]) for f in range(F)])
    d, o = random_lut(h * w, 4)
    shifts = np.tile(np.array([48, 32, 16, 0], np.int32), h // 4)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    t_rng = torch.from_numpy(rng.view(np.int32)).cuda()
    t_xyz = torch.empty((F, 2, h * w, 3), dtype=torch.float32, device="cuda")
    t_rd = torch.empty((F, 2, h, w), dtype=torch.int32, device="cuda")
    st = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
    ob.scan_to_cloud(lut, shifts, t_rng, xyz=t_xyz, range_destaggered=t_rd, stream=st)
    torch.cuda.synchronize()
    xyz = t_xyz.cpu().numpy()
    rd = t_rd.cpu().numpy().view(np.uint32)
    for f in range(F):
        for r in range(2):
            assert np.array_equal(xyz[f, r], orc.cartesian(rng[f, r], d, o))
            assert np.array_equal(rd[f, r], orc.destagger(rng[f, r], shifts))


def test_lut_from_intrinsics_matches_oracle_and_reference_py(ob):
    z = np.load(os.path.join(GOLDEN, "xyz_reference.npz"))
    for fx in ("OS-0-32-U1_v2.2.0_1024x10", "OS-1-128_767798045_1024x10_20230712_120049"):
        meta, _ = load_fixture(fx)
        rng, ref_xyz = z[fx + "/range"], z[fx + "/xyz"]
        h, w = rng.shape
        lut = ob.XYZLutT.from_intrinsics(w, h, 0.001, meta["beam_to_lidar_transform"],
                                         meta["lidar_to_sensor_transform"],
                                         meta["beam_azimuth_angles"], meta["beam_altitude_angles"])
        d, o = orc.make_xyz_lut(w, h, 0.001, meta["beam_to_lidar_transform"],
                                meta["lidar_to_sensor_transform"],
                                meta["beam_azimuth_angles"], meta["beam_altitude_angles"])
        # device libm vs host libm: a few ulp in double
        assert np.allclose(lut.direction, d, rtol=0, atol=1e-15)
        assert np.allclose(lut.offset, o, rtol=0, atol=1e-13)
        xyz = lut(rng).reshape(h, w, 3)
        assert np.allclose(xyz, ref_xyz, rtol=1e-9, atol=1e-9)      # doc formula, reference.py:18-76
        lutf = ob.XYZLutT.from_intrinsics(w, h, 0.001, meta["beam_to_lidar_transform"],
                                          meta["lidar_to_sensor_transform"],
                                          meta["beam_azimuth_angles"], meta["beam_altitude_angles"],
                                          dtype=np.float32)
        xyzf = lutf(rng).reshape(h, w, 3)
        err = np.linalg.norm(xyzf - ref_xyz, axis=-1)
        assert np.all(err <= 1e-5 * np.linalg.norm(ref_xyz, axis=-1) + 1e-7)  # north_star tolerance


def _normwise_ok(got, ref, tol=1e-5):
    err = np.linalg.norm(got.astype(np.float64) - ref.astype(np.float64), axis=-1)
    return bool(np.all(err <= tol * np.linalg.norm(ref.astype(np.float64), axis=-1) + 1e-7))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lut_free_projection_within_tolerance_of_lut_path(ob, dtype):
    """SURVEY 8d's LUT-free variant (opt-in): direction/offset rebuilt in the kernel from per-row and
    per-column tables (factorisation of xyzlut.cpp:35-86).  north_star tolerance: 1e-5 norm-wise
    relative vs the oracle's LUT path of the same dtype; empty returns stay exactly +0.0; the
    destaggered range is untouched (bit-exact)."""
    meta, _ = load_fixture("OS-1-128_767798045_1024x10_20230712_120049")
    h, w = meta["h"], meta["w"]
    args = (w, h, 0.001, meta["beam_to_lidar_transform"], meta["lidar_to_sensor_transform"],
            meta["beam_azimuth_angles"], meta["beam_altitude_angles"])
    d, o = orc.make_xyz_lut(*args)
    d, o = d.astype(dtype), o.astype(dtype)
    lut = ob.XYZLutT.from_intrinsics(*args, dtype=dtype)
    assert not lut.analytic
    lut.set_analytic(True)
    assert lut.analytic
    F = 3
    rng = np.stack([np.stack([random_range(h, w, 5 + 2 * f), random_range(h, w, 6 + 2 * f, 0.8)]) for f in range(F)])
    rng[0, 0, :, :8] = (1 << 19) - 1          # longest 19-bit range
    shifts = np.asarray(meta["pixel_shift_by_row"], np.int32)
    xyz = np.empty((F, 2, h * w, 3), dtype)
    rd = np.empty((F, 2, h, w), np.uint32)
    st = ob.Stream(0)
    ob.scan_to_cloud(lut, shifts, rng, xyz=xyz, range_destaggered=rd, stream=st)
    st.sync()
    for f in range(F):
        for r in range(2):
            ref = orc.cartesian(rng[f, r], d, o)
            assert _normwise_ok(xyz[f, r], ref)
            assert np.all(xyz[f, r][rng[f, r].reshape(-1) == 0] == 0.0)
            assert np.array_equal(rd[f, r], orc.destagger(rng[f, r], shifts))
    # single-frame entry point takes the same path
    assert _normwise_ok(lut(rng[1, 0]), orc.cartesian(rng[1, 0], d, o))
    # back to the bit-exact LUT path
    lut.set_analytic(False)
    assert np.array_equal(lut(rng[1, 0]), orc.cartesian(rng[1, 0], lut.direction, lut.offset))


def test_lut_free_projection_needs_intrinsics(ob):
    d, o = random_lut(32 * 512, 3)
    lut = ob.XYZLutT.from_arrays(d, o, 32, 512)
    with pytest.raises(ValueError, match="LUT-free projection needs a lut built from per-beam intrinsics"):
        lut.set_analytic(True)
