"""Does the grouping grid keep up with the sampling chain?  cfg3 layer 1 (B=32, N=1024 -> 512) per scale and as the
MSG layer, with 1..3 consumer CTAs per cloud and scale (pn2_set_sa_consumer_ctas)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointnet2_b200 import _lib, workloads as W
from pointnet2_b200.sa_layer import sample_group, sample_group_msg
from pointnet2_b200.tf_sampling import farthest_point_sample_and_gather
dev = torch.device("cuda:0"); lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn):
    ts = []
    for it in range(9):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize()
        if it >= 2: ts.append(a.elapsed_time(e))
    ts.sort(); return round(ts[len(ts) // 2], 4)
for gen in ("U", "S"):
    x = torch.from_numpy(W.DISTRIBUTIONS[gen](32, 1024, 100)).to(dev)
    print(gen, "fps+gather", t(lambda: farthest_point_sample_and_gather(512, x)), flush=True)
    for ctas in (1, 2, 3):
        lib.pn2_set_sa_consumer_ctas(ctas)
        row = {"ctas": ctas}
        for r, s in ((0.1, 16), (0.2, 32), (0.4, 128)):
            row[f"r{r}_S{s}"] = t(lambda: sample_group(512, r, s, x, center=False))
        row["msg"] = t(lambda: sample_group_msg(512, [0.1, 0.2, 0.4], [16, 32, 128], x))
        print(gen, row, flush=True)
    lib.pn2_set_sa_consumer_ctas(0)
