"""Run three_nn / three_interpolate / fused FP front end at the cfg4 last-FP-layer size (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet2_b200 import _lib, workloads as W
lib = _lib.load()
dev = torch.device("cuda:0")
b, n, m, c = 16, 8192, 1024, 128
x1 = torch.from_numpy(W.cloud_duplicates(b, n, 100)).to(dev)
x2 = x1[:, :m].contiguous()
p2 = torch.from_numpy(W.features(b, m, c, 105)).to(dev)
d = torch.empty((b, n, 3), device=dev); i = torch.empty((b, n, 3), dtype=torch.int32, device=dev)
o = torch.empty((b, n, c), device=dev)
for _ in range(3):
    lib.pn2_three_nn(b, n, m, x1.data_ptr(), x2.data_ptr(), d.data_ptr(), i.data_ptr(), None)
    lib.pn2_three_nn_interpolate(b, n, m, c, x1.data_ptr(), x2.data_ptr(), p2.data_ptr(), o.data_ptr(), None, None, None, None)
torch.cuda.synchronize()
print("ok")
import statistics
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
w = torch.full((b, n, 3), 1 / 3, device=dev)
def t(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); e.synchronize(); ts.append(a.elapsed_time(e))
    return statistics.median(ts)
print("three_nn ms", t(lambda: lib.pn2_three_nn(b, n, m, x1.data_ptr(), x2.data_ptr(), d.data_ptr(), i.data_ptr(), None)))
print("three_interpolate ms", t(lambda: lib.pn2_three_interpolate(b, m, c, n, p2.data_ptr(), i.data_ptr(), w.data_ptr(), o.data_ptr(), None)))
print("fused ms", t(lambda: lib.pn2_three_nn_interpolate(b, n, m, c, x1.data_ptr(), x2.data_ptr(), p2.data_ptr(), o.data_ptr(), None, None, None, None)))
