"""Time group_point at the cfg3 layer-2 size (C=320, the HBM-bound tensor) for tuning."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet2_b200 import _lib, workloads as W
lib = _lib.load(); dev = torch.device("cuda:0")
b, n, c, m = 32, 512, 320, 128
pts = torch.randn((b, n, c), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for s in (32, 64, 128):
    idx = torch.randint(0, n, (b, m, s), dtype=torch.int32, device=dev)
    out = torch.empty((b, m, s, c), device=dev)
    ts = []
    for _ in range(3): lib.pn2_group_point(b, n, c, m, s, pts.data_ptr(), idx.data_ptr(), out.data_ptr(), None)
    for _ in range(15):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); lib.pn2_group_point(b, n, c, m, s, pts.data_ptr(), idx.data_ptr(), out.data_ptr(), None); e.record(); e.synchronize()
        ts.append(a.elapsed_time(e))
    ms = statistics.median(ts); by = W.bytes_group(b, n, m, s, c)
    print(f"S={s} ms {ms:.4f} GB/s {by/ms/1e6:.0f} frac {by/ms/1e6/6571.2:.3f}")
# plain copy of the same size for reference (torch clone = read+write)
src = torch.empty((b, m, 128, c), device=dev); dst = torch.empty_like(src)
ts=[]
for _ in range(10):
    flush.zero_(); a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dst.copy_(src); e.record(); e.synchronize(); ts.append(a.elapsed_time(e))
ms = statistics.median(ts); print(f"torch copy 671MB: ms {ms:.4f} -> {2*src.numel()*4/ms/1e6:.0f} GB/s (r+w)")
ts=[]
for _ in range(10):
    flush.zero_(); a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dst.zero_(); e.record(); e.synchronize(); ts.append(a.elapsed_time(e))
ms = statistics.median(ts); print(f"torch memset 671MB: ms {ms:.4f} -> {src.numel()*4/ms/1e6:.0f} GB/s (write only)")
