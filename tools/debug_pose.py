import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as graft
from oracle import oracle as orc
from tests.helpers import random_lut, random_range
from tests.test_gpu_dewarp import _random_poses
ob = graft.load_package()
h, w, returns, frames, dtype = 20, 516, 2, 2, np.float32
rng = np.stack([np.stack([random_range(h, w, 10 * f + r) for r in range(returns)]) for f in range(frames)])
d, o = random_lut(h * w, 2, dtype)
lut = ob.XYZLutT.from_arrays(d, o, h, w)
poses = _random_poses(w, dtype, 5, None)
shifts = np.random.default_rng(3).integers(-30, 31, h).astype(np.int32)
for use_pose in (False, True):
    xyz = np.zeros((frames, returns, h * w, 3), dtype)
    xd = np.zeros((frames, returns, h, w, 3), dtype)
    rd = np.zeros((frames, returns, h, w), np.uint32)
    st = ob.Stream(0)
    ob.scan_to_cloud(lut, shifts, rng, xyz=xyz, range_destaggered=rd, xyz_destaggered=xd, stream=st,
                     poses=poses if use_pose else None)
    st.sync()
    for f in range(frames):
        for r in range(returns):
            want = orc.cartesian(rng[f, r], d, o).reshape(h, w, 3)
            if use_pose:
                want = orc.dewarp(want, poses)
            got = xyz[f, r].reshape(h, w, 3)
            bad = np.argwhere((got != want).any(axis=2))
            bad_d = np.argwhere((xd[f, r] != orc.destagger(want, shifts)).any(axis=2))
            bad_r = np.argwhere(rd[f, r] != orc.destagger(rng[f, r], shifts))
            print("pose", use_pose, "f", f, "r", r, "xyz bad", len(bad), bad[:5].tolist(), "xd bad", len(bad_d), bad_d[:3].tolist(), "rd bad", len(bad_r))
            if len(bad):
                y, x = bad[0]
                print("   got", got[y, x], "want", want[y, x], "rows bad", sorted(set(bad[:, 0].tolist()))[:10], "cols", sorted(set(bad[:, 1].tolist()))[:10])
