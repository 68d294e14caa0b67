#!/usr/bin/env python
"""One launch of K1 with fused poses for ncu (128x2048 dual, 64 frames)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as graft
import bench
ob = graft.load_package()
F, H, W, R = 64, bench.H, bench.W, bench.R
dev = torch.device("cuda", 0)
rng = torch.from_numpy(bench.synth_pool(F).view(np.int32)).to(dev)
d, o = bench.synth_lut()
lut = ob.XYZLutT.from_arrays(torch.from_numpy(d).to(dev), torch.from_numpy(o).to(dev), H, W)
xyz = torch.empty((F, R, H * W, 3), dtype=torch.float32, device=dev)
rd = torch.empty((F, R, H, W), dtype=torch.int32, device=dev)
poses = np.tile(np.eye(4, dtype=np.float32), (F, W, 1, 1))
poses[..., :3, 3] = np.random.default_rng(1).random((F, W, 3))
t_pose = torch.from_numpy(poses).to(dev)
st = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    ob.scan_to_cloud(lut, bench.SHIFTS, rng, xyz=xyz, range_destaggered=rd, stream=st, poses=t_pose)
torch.cuda.synchronize()
