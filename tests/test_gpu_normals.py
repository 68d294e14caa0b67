"""GPU parity tests for surface normals on destaggered XYZ (ob_normals) against the CPU
oracle (oracle/orc_normals.c, restating ouster_algorithm/src/normals.cpp) and the reference's own
known answers (python/tests/test_normals.py)."""
import numpy as np
import pytest

import __graft_entry__ as graft
from oracle import oracle as orc
from tests.test_oracle_normals import room_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    graft.build()
    m = graft.load_package()
    assert m.device_count() > 0
    return m


def noisy_scene(h, w, seed, p_zero=0.15):
    """Cube-room scene with range noise, dropped returns and a sparse second return."""
    rs = np.random.default_rng(seed)
    xyz, rng, d = room_scene(h, w)
    rng = (rng.astype(np.int64) + rs.integers(-40, 41, rng.shape)).astype(np.uint32)
    rng[rs.random(rng.shape) < p_zero] = 0
    xyz = d * rng[..., None] * 0.001
    rng2 = np.where(rs.random(rng.shape) < 0.25, rng + rs.integers(100, 1500, rng.shape), 0).astype(np.uint32)
    rng2[rng == 0] = 0
    xyz2 = d * rng2[..., None] * 0.001
    return xyz, rng, xyz2, rng2


def ulp_close(a, b, n=8):
    return abs(a - b) <= n * np.spacing(max(abs(a), abs(b)))


@pytest.mark.parametrize("search", [1, 3])
@pytest.mark.parametrize("h,w", [(64, 512), (37, 200)])
def test_single_return_bit_exact_vs_oracle(ob, h, w, search):
    xyz, rng, _, _ = noisy_scene(h, w, 5 + search)
    org = np.random.default_rng(1).normal(0, 0.05, (w, 3))
    got, sub = ob.normals(xyz, rng, org, search, return_subtent=True)
    assert ulp_close(sub, orc.normals_vertical_subtent(xyz, rng, org))   # device acos vs glibc acos
    ref = orc.normals(xyz, rng, sensor_origins_xyz=org, pixel_search_range=search, vertical_subtent=sub)
    assert np.array_equal(got, ref)
    norms = np.linalg.norm(got, axis=-1)
    assert np.allclose(norms[norms > 0], 1.0, atol=1e-9)
    assert np.all(got[rng == 0] == 0)


@pytest.mark.parametrize("search", [1, 2])
def test_dual_return_bit_exact_vs_oracle(ob, search):
    xyz, rng, xyz2, rng2 = noisy_scene(64, 512, 11)
    org = np.zeros((512, 3))
    (g1, g2), sub = ob.normals(xyz, rng, xyz2, rng2, org, search, return_subtent=True)
    r1, r2 = orc.normals(xyz, rng, xyz2, rng2, org, search, vertical_subtent=sub)
    assert np.array_equal(g1, r1) and np.array_equal(g2, r2)
    # without the override the only difference allowed is the acos of the subtent
    o1, o2 = orc.normals(xyz, rng, xyz2, rng2, org, search)
    assert np.allclose(g1, o1, atol=1e-9) and np.allclose(g2, o2, atol=1e-9)


def test_float_inputs_are_widened_and_rounded_once(ob):
    xyz, rng, _, _ = noisy_scene(32, 256, 21)
    org = np.zeros((256, 3))
    xf = xyz.astype(np.float32)
    got, sub = ob.normals(xf, rng, org, return_subtent=True)
    assert got.dtype == np.float32
    ref = orc.normals(xf.astype(np.float64), rng, sensor_origins_xyz=org, vertical_subtent=sub)
    assert np.array_equal(got, ref.astype(np.float32))


def test_reference_known_answers(ob):
    # python/tests/test_normals.py:362-442
    xyz = np.array([[[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]], [[0.0, 1.0, 0.0], [1.0, 1.0, 0.0]]])
    rng = np.array([[0.0, 1.0], [1.0, 1.0]], np.uint32)
    org = np.zeros((2, 3))
    want = np.array([[[0.0, 0.0, 0.0], [-1.0, 0.0, 0.0]], [[0.0, -1.0, 0.0], [-0.70710678, -0.70710678, 0.0]]])
    assert np.allclose(ob.normals(xyz, rng, org, 1, 0.1, 100), want)
    r1, r2 = ob.normals(xyz, rng, xyz, rng, org, 1, 0.1, 100)
    assert np.allclose(r1, want)
    xyz = np.array([[[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]], [[0.0, 0.0, 0.0], [0.0, 0.0, 0.0]]])
    rng = np.array([[0.0, 1.0], [0.0, 0.0]], np.uint32)
    want = np.array([[[0.0, 0.0, 0.0], [-1.0, 0.0, 0.0]], [[0.0, 0.0, 0.0], [0, 0.0, 0.0]]])
    assert np.allclose(ob.normals(xyz, rng, org, 1, 0.1, 100), want)
    r1, _ = ob.normals(xyz, rng, xyz, rng, org, 1, 0.1, 100)
    assert np.allclose(r1, want)


def test_reference_error_texts(ob):
    # python/tests/test_normals.py:212-359
    xyz = np.array([[[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]], [[0.0, 1.0, 0.0], [1.0, 1.0, 0.0]]])
    rng = np.array([[0.0, 1.0], [1.0, 1.0]], np.uint32)
    org = np.zeros((2, 3))
    for extra in ((), (xyz, rng)):
        with pytest.raises(RuntimeError, match=r"target_distance_m must be positive"):
            ob.normals(xyz, rng, *extra, org, 1, 0.017453292519943295, -100)
        with pytest.raises(RuntimeError, match=r"normals: min_angle_of_incidence_rad must be positive"):
            ob.normals(xyz, rng, *extra, org, 1, -0.1, 100)
        with pytest.raises(RuntimeError, match=r"normals: sensor_origins size must match image width"):
            ob.normals(xyz, rng, *extra, np.zeros((0, 3)), 1, 0.017453292519943295, 100)
        with pytest.raises(TypeError, match=r"incompatible function arguments"):
            ob.normals(xyz, rng, *extra, np.zeros((0, 0)), 1, 0.017453292519943295, 100)
    with pytest.raises(RuntimeError, match=r"normals: xyz dimensions mismatch"):
        ob.normals(xyz, np.array([[0.0, 1.0]], np.uint32), org, 1, 0.017453292519943295, 100)


def test_chain_stays_on_the_device(ob):
    """range -> fused destagger + XYZ (K1, xyz_destaggered) -> normals without leaving HBM: torch CUDA
    tensors in, torch CUDA tensors out, results equal to the host path."""
    import torch
    h, w = 64, 512
    _, rng_d, d = room_scene(h, w)
    shifts = np.tile(np.array([12, 8, 4, 0], np.int32), h // 4)
    rng_st = orc.destagger(rng_d, shifts, inverse=True)          # what the sensor delivers (staggered)
    dirs_st = orc.destagger(d.astype(np.float64), shifts, inverse=True).reshape(h * w, 3) * 0.001
    lut = ob.XYZLutT.from_arrays(dirs_st, np.zeros_like(dirs_st), h, w)
    dev = torch.device("cuda", 0)
    t_rng = torch.from_numpy(rng_st.view(np.int32)[None, None]).to(dev)
    t_xd = torch.empty((1, 1, h, w, 3), dtype=torch.float64, device=dev)
    t_rd = torch.empty((1, 1, h, w), dtype=torch.int32, device=dev)
    st = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
    ob.scan_to_cloud(lut, shifts, t_rng, range_destaggered=t_rd, xyz_destaggered=t_xd, stream=st)
    t_org = torch.zeros((w, 3), dtype=torch.float64, device=dev)
    t_n = ob.normals(t_xd[0, 0], t_rd[0, 0], t_org, stream=st)
    assert t_n.is_cuda and t_n.shape == (h, w, 3)
    torch.cuda.synchronize()
    host = ob.normals(t_xd[0, 0].cpu().numpy(), t_rd[0, 0].cpu().numpy().view(np.uint32), np.zeros((w, 3)))
    assert np.array_equal(t_n.cpu().numpy(), host)
    assert np.array_equal(t_rd[0, 0].cpu().numpy().view(np.uint32), rng_d)
