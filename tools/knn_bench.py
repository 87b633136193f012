"""Time pn2_knn_point at cfg2's shape (32 clouds x 4096 points, 1024 queries) for k = 8..128, on a uniform cloud
(tie-free rows: sorted fast path) and a duplicate-heavy one (ties: exact selection-sort replay)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointnet2_b200 import workloads as W
from pointnet2_b200.tf_grouping import knn_point
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
for gen in ("U", "D"):
    x = torch.from_numpy(W.DISTRIBUTIONS[gen](32, 4096, 100)).to(dev)
    q = x[:, :1024].contiguous() if gen == "D" else torch.from_numpy(W.cloud_uniform(32, 1024, 101)).to(dev)
    for k in (8, 32, 64, 128):
        ts = []
        for it in range(7):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); knn_point(k, x, q); e.record(); torch.cuda.synchronize()
            if it >= 2: ts.append(a.elapsed_time(e))
        ts.sort()
        rows.append(dict(cloud=gen, k=k, ms=round(ts[len(ts) // 2], 4), pairs_per_s=round(32 * 1024 * 4096 / (ts[len(ts) // 2] * 1e-3), 0)))
        print(rows[-1], flush=True)
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "knn_bench.json"), "w"), indent=1)
