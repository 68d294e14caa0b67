"""Next-row: per-column pose application.  Known answers are the reference's own
(python/tests/test_pose_util.py:300-360, rtol 1e-5); GPU vs oracle is bit-exact."""
import numpy as np
import pytest

import __graft_entry__ as graft
from oracle import oracle as orc


def _ref_cases():
    poses = np.array([[1, 0, 0, 1, 0, 1, 0, -2, 0, 0, 1, 3, 0, 0, 0, 1] for _ in range(4)], np.float64)
    points = np.array([[i - 3, i + 1, i + 2] for i in range(8)], np.float64).reshape(2, 4, 3)
    expected = np.array([[[-2, -1, 5], [-1, 0, 6], [0, 1, 7], [1, 2, 8]],
                         [[2, 3, 9], [3, 4, 10], [4, 5, 11], [5, 6, 12]]], np.float64)
    pts2 = np.arange(1, 25, dtype=np.float64).reshape(2, 4, 3)
    tf = np.array([[0.866, -0.5, 0.0, 1.0], [0.5, 0.866, 0.0, 2.0], [0.0, 0.0, 1.0, -1.0], [0, 0, 0, 1.0]])
    exp2 = np.array([[[0.866, 4.232, 2], [1.964, 8.33, 5], [3.062, 12.428, 8], [4.16, 16.526, 11]],
                     [[5.258, 20.624, 14], [6.356, 24.722, 17], [7.454, 28.82, 20], [8.552, 32.918, 23]]])
    return poses, points, expected, pts2, tf, exp2


def test_oracle_dewarp_known_answers():
    poses, points, expected, pts2, tf, exp2 = _ref_cases()
    np.testing.assert_allclose(orc.dewarp(points, poses), expected, rtol=1e-5, atol=1e-8)
    np.testing.assert_almost_equal(orc.dewarp(pts2.reshape(-1, 3), tf.reshape(1, 16)).reshape(2, 4, 3), exp2, decimal=5)


@pytest.fixture(scope="module")
def ob():
    graft.build()
    m = graft.load_package()
    assert m.device_count() > 0
    return m


@pytest.mark.gpu
def test_gpu_dewarp_known_answers(ob):
    poses, points, expected, pts2, tf, exp2 = _ref_cases()
    got = ob.dewarp(points, poses.reshape(4, 4, 4))
    assert got.shape == (2, 4, 3)
    np.testing.assert_allclose(got, expected, rtol=1e-5, atol=1e-8)
    np.testing.assert_almost_equal(ob.transform(pts2, tf), exp2, decimal=5)
    with pytest.raises(RuntimeError, match="Number of points per set must match number of poses"):
        ob.dewarp(np.zeros((2, 5, 3)), poses.reshape(4, 4, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", [(128, 2048), (64, 1024), (7, 130), (1, 1)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gpu_dewarp_matches_oracle(ob, h, w, dtype):
    rs = np.random.default_rng(h * w)
    pts = (rs.random((h, w, 3)) * 100 - 50).astype(dtype)
    ang = rs.random(w) * 2 * np.pi
    poses = np.zeros((w, 4, 4), dtype)
    poses[:, 0, 0], poses[:, 0, 1] = np.cos(ang), -np.sin(ang)
    poses[:, 1, 0], poses[:, 1, 1] = np.sin(ang), np.cos(ang)
    poses[:, 2, 2] = 1
    poses[:, 3, 3] = 1
    poses[:, :3, 3] = rs.random((w, 3)) * 10
    got = ob.dewarp(pts, poses)
    assert np.array_equal(got, orc.dewarp(pts, poses))
    ref = np.einsum("wij,hwj->hwi", poses[:, :3, :3].astype(np.float64), pts.astype(np.float64)) + poses[None, :, :3, 3]
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-4 if dtype == np.float32 else 1e-9)


def _random_poses(w, dtype, seed, frames=None):
    rs = np.random.default_rng(seed)
    n = w if frames is None else frames * w
    ang = rs.random(n) * 2 * np.pi
    poses = np.zeros((n, 4, 4), dtype)
    poses[:, 0, 0], poses[:, 0, 1] = np.cos(ang), -np.sin(ang)
    poses[:, 1, 0], poses[:, 1, 1] = np.sin(ang), np.cos(ang)
    poses[:, 2, 2] = 1
    poses[:, 3, 3] = 1
    poses[:, :3, 3] = rs.random((n, 3)) * 10
    return poses if frames is None else poses.reshape(frames, w, 4, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,returns,frames", [(128, 2048, 2, 3), (64, 1024, 1, 2), (20, 516, 2, 2), (7, 130, 1, 1)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("per_frame", [False, True])
def test_fused_projection_and_dewarp_matches_two_passes(ob, h, w, returns, frames, dtype, per_frame):
    """scan_to_cloud(poses=...) == dewarp(cartesian(range), poses) of the oracle, bit for bit,
    for the staggered and the destaggered cloud (pose_util.h:37-59 after impl/cartesian.h:36-66).
    (20,516) has rows that do not fill the last 8-row tile; (7,130) takes the generic kernel."""
    from tests.helpers import random_lut, random_range
    rng = np.stack([np.stack([random_range(h, w, 10 * f + r) for r in range(returns)]) for f in range(frames)])
    d, o = random_lut(h * w, 2, dtype)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    poses = _random_poses(w, dtype, 5, frames if per_frame else None)
    # negative shifts only on power-of-two widths (the reference's size_t modulo)
    lo = -30 if (w & (w - 1)) == 0 else 0
    shifts = np.random.default_rng(3).integers(lo, 31, h).astype(np.int32)
    xyz = np.zeros((frames, returns, h * w, 3), dtype)
    xd = np.zeros((frames, returns, h, w, 3), dtype)
    rd = np.zeros((frames, returns, h, w), np.uint32)
    st = ob.Stream(0)
    ob.scan_to_cloud(lut, shifts, rng, xyz=xyz, range_destaggered=rd, xyz_destaggered=xd, stream=st, poses=poses)
    st.sync()
    for f in range(frames):
        pf = poses[f] if per_frame else poses
        for r in range(returns):
            want = orc.dewarp(orc.cartesian(rng[f, r], d, o).reshape(h, w, 3), pf)
            assert np.array_equal(xyz[f, r].reshape(h, w, 3), want), (f, r)
            assert np.array_equal(xd[f, r], orc.destagger(want, shifts)), (f, r)
            assert np.array_equal(rd[f, r], orc.destagger(rng[f, r], shifts)), (f, r)
    # zero-range points land on the column translation, as dewarp() of a zero point does
    zr = rng[0, 0] == 0
    if zr.any():
        p0 = poses[0] if per_frame else poses
        cols = np.nonzero(zr)[1]
        assert np.array_equal(xyz[0, 0].reshape(h, w, 3)[zr], p0[cols, :3, 3])


@pytest.mark.gpu
def test_fused_dewarp_argument_errors(ob):
    from tests.helpers import random_lut, random_range
    h, w = 16, 64
    d, o = random_lut(h * w, 2)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    rng = random_range(h, w, 1).reshape(1, 1, h, w)
    with pytest.raises(ValueError, match="dtype of the lut"):
        ob.scan_to_cloud(lut, None, rng, xyz=np.zeros((1, 1, h * w, 3), np.float32), poses=np.zeros((w, 4, 4)))
    with pytest.raises(ValueError, match="poses must be"):
        ob.scan_to_cloud(lut, None, rng, xyz=np.zeros((1, 1, h * w, 3), np.float32),
                         poses=np.zeros((w - 1, 4, 4), np.float32))
    with pytest.raises(ValueError, match="without an xyz output"):
        ob.scan_to_cloud(lut, np.zeros(h, np.int32), rng, range_destaggered=np.zeros((1, 1, h, w), np.uint32),
                         poses=np.zeros((w, 4, 4), np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", [(128, 2048), (64, 1024), (20, 516), (5, 33)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_dewarp_frame_matches_reference_loop(ob, h, w, dtype):
    """dewarp(LidarFrame, lut, min_range, max_range) (pose_util.h:456-485, impl/dewarp_impl.h:22-76):
    same points, same order, same provenance as the oracle's restatement of the reference loop."""
    from tests.helpers import random_lut, random_range
    rs = np.random.default_rng(h + w)
    rng = random_range(h, w, 7, p_zero=0.3, max_range=60000)
    d, o = random_lut(h * w, 2, dtype)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    poses = _random_poses(w, np.float64, 9)
    status = np.ones(w, np.uint32)
    status[: w // 10] = 0                       # leading invalid columns: before the first valid one
    status[rs.integers(w // 10, w, w // 8)] = 0  # holes: skipped
    status[w // 2] = 2                          # non-zero word without the valid bit: still visited
    status[-3:] = 0                             # trailing invalid columns
    ts = (1000 + 17 * np.arange(w)).astype(np.uint64)
    for lo, hi in ((0.5, 30.0), (0.0, 100.0), (10.0, 10.5), (50.0, 40.0)):
        want_p, want_c, want_t = orc.dewarp_frame(rng, d, o, poses, status, ts, lo, hi)
        got_p, got_c, got_t = ob.dewarp_frame(lut, rng, poses, status, ts, lo, hi, provenance=True)
        assert got_p.shape == want_p.shape, (lo, hi)
        assert np.array_equal(got_p, want_p), (lo, hi)
        assert np.array_equal(got_c, want_c) and np.array_equal(got_t, want_t)
        assert np.array_equal(ob.dewarp_frame(lut, rng, poses, status, None, lo, hi), want_p)
    # no valid column at all -> empty (get_first_valid_column throws in the reference, :44-49)
    none = np.zeros(w, np.uint32)
    none[3] = 2
    assert ob.dewarp_frame(lut, rng, poses, none, ts, 0.5, 30.0).shape == (0, 3)
    assert orc.dewarp_frame(rng, d, o, poses, none, ts, 0.5, 30.0)[0].shape == (0, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_dewarp_frame_set_is_the_concatenation_of_its_frames(ob, dtype):
    """dewarp(FrameSet, xyzluts, min_range, max_range) (pose_util.h:475, impl/dewarp_impl.h:84-117): the
    frames of a set -- different sensors, different shapes, one empty slot -- dewarped in one batched
    pass == the oracle's single-frame loop applied frame by frame and concatenated, with the
    frame / column / timestamp provenance of dewarp_impl."""
    from tests.helpers import random_lut, random_range
    shapes = [(64, 1024), None, (128, 512), (20, 516)]
    frames, want_p, want_f, want_c, want_t = [], [], [], [], []
    for i, hw in enumerate(shapes):
        if hw is None:
            frames.append(None)
            continue
        h, w = hw
        rng = random_range(h, w, 11 + i, p_zero=0.3, max_range=60000)
        d, o = random_lut(h * w, 5 + i, dtype)
        poses = _random_poses(w, np.float64, 3 + i)
        status = np.ones(w, np.uint32)
        status[: w // 9] = 0
        status[np.random.default_rng(i).integers(w // 9, w, w // 7)] = 0
        ts = (500 * i + 3 * np.arange(w)).astype(np.uint64)
        frames.append({"lut": ob.XYZLutT.from_arrays(d, o, h, w), "range": rng, "poses": poses, "status": status,
                       "timestamps": ts})
        p, c, t = orc.dewarp_frame(rng, d, o, poses, status, ts, 0.5, 45.0)
        want_p.append(p)
        want_c.append(c)
        want_t.append(t)
        want_f.append(np.full(len(c), i, np.uint32))
    got_p, got_f, got_c, got_t = ob.dewarp_frames(frames, 0.5, 45.0, provenance=True)
    assert np.array_equal(got_p, np.concatenate(want_p))
    assert np.array_equal(got_f, np.concatenate(want_f))
    assert np.array_equal(got_c, np.concatenate(want_c))
    assert np.array_equal(got_t, np.concatenate(want_t))
    assert np.array_equal(ob.dewarp_frames(frames, 0.5, 45.0), np.concatenate(want_p))
    assert ob.dewarp_frames([None, None]).shape == (0, 3)
    assert ob.dewarp_frames(frames, 50.0, 40.0).shape == (0, 3)     # empty range window


@pytest.mark.gpu
def test_dewarp_frame_long_lookback_chain_and_device_count(ob):
    """The single-launch compaction: 128 CTAs in one frame (four look-back windows of 32) and a set of
    wide frames whose chain continues across the frames; then the asynchronous form (device outputs and a
    device-side count, no host wait inside the call) against the host-output form."""
    import torch
    from tests.helpers import random_lut, random_range
    h, w = 32, 4096
    rng = random_range(h, w, 21, p_zero=0.6, max_range=60000)
    rng[:, 700:2100] = 0                      # long runs of CTAs with nothing to emit
    d, o = random_lut(h * w, 4, np.float32)
    lut = ob.XYZLutT.from_arrays(d, o, h, w)
    poses = _random_poses(w, np.float64, 5)
    status = np.ones(w, np.uint32)
    status[np.random.default_rng(8).integers(0, w, w // 5)] = 0
    ts = np.arange(w, dtype=np.uint64)
    want_p, want_c, _ = orc.dewarp_frame(rng, d, o, poses, status, ts, 0.2, 55.0)
    for _ in range(3):                        # scheduling order of the CTAs differs from run to run
        got_p, got_c, _ = ob.dewarp_frame(lut, rng, poses, status, ts, 0.2, 55.0, provenance=True)
        assert np.array_equal(got_p, want_p) and np.array_equal(got_c, want_c)
    frames = [{"lut": lut, "range": np.roll(rng, 37 * i, axis=1), "poses": poses, "status": status, "timestamps": ts}
              for i in range(5)]
    want = np.concatenate([orc.dewarp_frame(f["range"], d, o, poses, status, ts, 0.2, 55.0)[0] for f in frames])
    assert np.array_equal(ob.dewarp_frames(frames, 0.2, 55.0), want)
    # asynchronous device form
    dev = torch.device("cuda", 0)
    out = torch.full((h * w, 3), -1.0, dtype=torch.float32, device=dev)
    cnt = torch.full((1,), -1, dtype=torch.int64, device=dev)
    t = {k: torch.from_numpy(v).to(dev) for k, v in
         (("rng", rng.view(np.int32)), ("poses", poses), ("status", status.view(np.int32)))}
    ob.dewarp_frame(lut, t["rng"], t["poses"], t["status"], None, 0.2, 55.0, out=out, out_count=cnt)
    torch.cuda.synchronize()
    assert int(cnt.item()) == len(want_p)
    assert np.array_equal(out[:len(want_p)].cpu().numpy(), want_p)
    assert bool((out[len(want_p):] == -1.0).all())
    # capacity smaller than the result: the list is cut, the count still says how many passed
    small = torch.full((100, 3), -1.0, dtype=torch.float32, device=dev)
    ob.dewarp_frame(lut, t["rng"], t["poses"], t["status"], None, 0.2, 55.0, out=small, out_count=cnt)
    torch.cuda.synchronize()
    assert int(cnt.item()) == len(want_p) and np.array_equal(small.cpu().numpy(), want_p[:100])
