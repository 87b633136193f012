"""Time the fused grouping tail (pn2_group_concat) at the BASELINE layer shapes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet2_b200 import workloads as W
from pointnet2_b200.pointnet_util import group_and_concat
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
for (b, n, c, m, s) in [(16, 1024, 64, 256, 32), (16, 256, 128, 64, 32), (32, 512, 320, 128, 32), (32, 512, 320, 128, 128), (16, 64, 256, 16, 32), (8, 2048, 67, 512, 32)]:
    g = torch.Generator(device=dev).manual_seed(1)
    xyz = torch.rand((b, n, 3), device=dev, generator=g); new_xyz = xyz[:, :m].contiguous()
    pts = torch.randn((b, n, c), device=dev, generator=g)
    idx = torch.randint(0, n, (b, m, s), device=dev, generator=g, dtype=torch.int32)
    for _ in range(3): out, gx = group_and_concat(xyz, new_xyz, pts, idx, xyz_first=False)
    ts = []
    for _ in range(15):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); group_and_concat(xyz, new_xyz, pts, idx, xyz_first=False); e.record(); e.synchronize(); ts.append(a.elapsed_time(e))
    ts.sort(); ms = ts[len(ts) // 2]
    by = 4 * b * m * s + 4 * b * min(n, m * s) * (c + 3) + 4 * b * m * s * (c + 3) + 12 * b * m * s
    rows.append(dict(b=b, n=n, c=c, m=m, s=s, ms=round(ms, 4), GBps=round(by / ms / 1e6), checksum=float(out.double().sum())))
    print(rows[-1], flush=True)
