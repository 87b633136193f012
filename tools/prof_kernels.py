"""Run one kernel family a few times at a BASELINE shape, for ncu (python tools/prof_kernels.py CASE [reps]).
CASES: layer_cfg2, layer_cfg4, layer_cfg3, concat64, concat320, interp128, interp_grad, fps65536, fps262144, knn"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointnet2_b200 import _lib, workloads as W
from pointnet2_b200.sa_layer import sample_group
from pointnet2_b200.pointnet_util import group_and_concat
from pointnet2_b200.tf_grouping import query_ball_point, knn_point
from pointnet2_b200.tf_interpolate import three_interpolate, three_nn
from pointnet2_b200.tf_sampling import farthest_point_sample_and_gather

case = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run(fn):
    for _ in range(reps):
        flush.zero_()
        fn()
    torch.cuda.synchronize()

if case == "layer_cfg2":
    x = T(W.cloud_uniform(32, 4096, 100)); run(lambda: sample_group(1024, 0.1, 32, x, center=False))
elif case == "layer_cfg4":
    x = T(W.cloud_duplicates(16, 8192, 100)); run(lambda: sample_group(1024, 0.1, 32, x, center=False))
elif case == "bg_cfg3":
    from pointnet2_b200.sa_layer import ball_group
    x = T(W.cloud_surface(32, 1024, 100)); q = x[:, :512].contiguous(); run(lambda: ball_group(0.4, 128, x, q, center=True))
elif case == "layer_cfg3":
    x = T(W.cloud_surface(32, 1024, 100)); run(lambda: sample_group(512, 0.2, 32, x, center=False))
elif case in ("concat64", "concat320"):
    b, n, m, s, c = (16, 1024, 256, 32, 64) if case == "concat64" else (32, 512, 128, 64, 320)
    x = T(W.cloud_uniform(b, n, 1)); f = T(W.features(b, n, c, 2)); nx = x[:, :m].contiguous()
    idx, _ = query_ball_point(0.3, s, x, nx)
    run(lambda: group_and_concat(x, nx, f, idx, xyz_first=True))
elif case == "interp128":
    b, n, m, c = 16, 8192, 1024, 128
    x1 = T(W.cloud_uniform(b, n, 1)); x2 = x1[:, :m].contiguous(); p = T(W.features(b, m, c, 2))
    d, i = three_nn(x1, x2); w = torch.full((b, n, 3), 1 / 3, device=dev)
    run(lambda: three_interpolate(p, i, w))
elif case == "interp_grad":
    b, n, m, c = 16, 8192, 1024, 128
    x1 = T(W.cloud_uniform(b, n, 1)); x2 = x1[:, :m].contiguous(); p = T(W.features(b, m, c, 2)).requires_grad_(True)
    d, i = three_nn(x1, x2); w = torch.full((b, n, 3), 1 / 3, device=dev); g = torch.randn(b, n, c, device=dev)
    def f():
        p.grad = None
        (three_interpolate(p, i, w) * g).sum().backward()
    run(f)
elif case in ("fps65536", "fps262144"):
    n = int(case[3:]); x = T(W.cloud_uniform(8, n, 116)); run(lambda: farthest_point_sample_and_gather(128, x))
elif case == "knn":
    x = T(W.cloud_uniform(32, 4096, 100)); q = x[:, :1024].contiguous(); run(lambda: knn_point(32, x, q))
print("done", case)
