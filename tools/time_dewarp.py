import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g
ob = g.load_package()
h, w, F = 128, 2048, 64
pts = torch.rand((F * h, w, 3), device='cuda')
poses = torch.eye(4, device='cuda').repeat(w, 1, 1).contiguous()
out = torch.empty_like(pts)
st = ob.Stream(0, cuda_stream=torch.cuda.current_stream().cuda_stream)
for _ in range(3): ob.dewarp(pts, poses, out=out, stream=st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): ob.dewarp(pts, poses, out=out, stream=st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("dewarp f32 %d pts: %.3f ms, %.0f GB/s (24 B/pt)" % (F*h*w, ms, F*h*w*24/ms/1e6))
