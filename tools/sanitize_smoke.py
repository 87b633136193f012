"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointnet2_b200 import _lib, workloads as W
from pointnet2_b200.pointnet_util import sample_and_group, pointnet_fp_module, pointnet_sa_module_msg
from pointnet2_b200.tf_sampling import farthest_point_sample
from pointnet2_b200.tf_grouping import query_ball_point, select_top_k
dev = torch.device("cuda:0")
lib = _lib.load()
xyz = torch.from_numpy(W.cloud_duplicates(2, 2500, 3)).to(dev)
f = torch.from_numpy(W.features(2, 2500, 8, 4)).to(dev).requires_grad_(True)
nx, npts, idx, gx = sample_and_group(64, 0.3, 16, xyz, f)
npts.sum().backward()
up = pointnet_fp_module(xyz, nx, None, npts.max(dim=2).values.detach())
_ = pointnet_sa_module_msg(xyz, f.detach(), 32, [0.2, 0.4], [8, 16])
for cfg in [(128, 32, 1), (128, 8, 4), (256, 8, 2), (512, 32, 2), (512, 44, 3), (512, 52, 5)]:
    lib.pn2_set_fps_config(*cfg)
    farthest_point_sample(40, xyz)
lib.pn2_set_fps_config(0, 0, 0)
u = torch.from_numpy(W.cloud_uniform(2, 3000, 5)).to(dev)
q = u[:, :200].contiguous()
query_ball_point(0.08, 16, u, q)          # grid path
lib.pn2_set_bq_mode(1); query_ball_point(0.08, 16, u, q); lib.pn2_set_bq_mode(0)
select_top_k(3, torch.rand(2, 5, 40, device=dev))
from pointnet2_b200.sa_layer import ball_group, sample_group
from pointnet2_b200.tf_grouping import knn_point
sample_group(128, 0.08, 16, u)                      # overlapped layer, grid mode
sample_group(64, 0.3, 16, xyz, center=False)        # overlapped layer, ordered-scan mode (duplicate-heavy cloud)
ball_group(0.08, 16, u, q)
knn_point(8, u, q)
knn_point(40, xyz, nx)
torch.cuda.synchronize()
print("sanitize smoke done")
