"""Time query_ball_point (auto = grid where flagged, vs brute force only) per input distribution."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet2_b200 import _lib, workloads as W
from pointnet2_b200.tf_sampling import farthest_point_sample_and_gather
from pointnet2_b200.tf_grouping import query_ball_point
dev = torch.device("cuda:0"); lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=30):
    for _ in range(5): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
rows = []
for dist in "USD":
    for (b, n, m, r, s) in [(32, 4096, 1024, 0.1, 32), (16, 8192, 1024, 0.1, 32), (32, 4096, 1024, 0.2, 32), (32, 4096, 1024, 0.05, 32), (8, 16384, 4096, 0.1, 32)]:
        xyz = torch.from_numpy(W.DISTRIBUTIONS[dist](b, n, 500)).to(dev)
        _, q = farthest_point_sample_and_gather(m, xyz)
        ws_bytes = int(lib.pn2_query_ball_point_workspace_bytes(b, n))
        ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=dev)
        lib.pn2_ball_grid_build(b, n, r, s, xyz.data_ptr(), ws.data_ptr(), ws_bytes, None)
        torch.cuda.synchronize()
        stride = ws_bytes // b // 4
        flags = ws.view(torch.int32)[::stride][:b].cpu().tolist() if ws_bytes else []
        lib.pn2_set_bq_mode(0); ia, ca = query_ball_point(r, s, xyz, q); t_auto = timeit(lambda: query_ball_point(r, s, xyz, q))
        lib.pn2_set_bq_mode(1); ib, cb = query_ball_point(r, s, xyz, q); t_brute = timeit(lambda: query_ball_point(r, s, xyz, q))
        lib.pn2_set_bq_mode(0)
        assert torch.equal(ia, ib) and torch.equal(ca, cb)
        rows.append(dict(dist=dist, b=b, n=n, m=m, r=r, s=s, auto_ms=round(t_auto, 4), brute_ms=round(t_brute, 4), grid_clouds=sum(flags), mean_cnt=float(ca.float().mean())))
        print(rows[-1], flush=True)
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "bq_dist.json"), "w"), indent=1)
