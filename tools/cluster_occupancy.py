"""Print cudaOccupancyMaxActiveClusters for every FPS cluster kernel variant (B200: which cluster
sizes keep 8 clouds co-resident?).  python tools/cluster_occupancy.py > profiles/r2_fps_cluster_occupancy.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet2_b200 import _lib
torch.cuda.init(); torch.zeros(1, device="cuda")
lib = _lib.load()
print("threads ppt cluster smem_KB max_active_clusters ctas")
for t, p in [(128, 4), (128, 8), (128, 16), (128, 32), (256, 2), (256, 4), (256, 8), (256, 16), (256, 32), (512, 1), (512, 2), (512, 4), (512, 8), (512, 16), (512, 32), (1024, 2), (1024, 4), (1024, 8)]:
    for c in (2, 4, 8, 16):
        k = lib.pn2_fps_cluster_capacity(t, p, c)
        print(f"{t:5d} {p:3d} {c:4d} {3*p*t*4/1024:7.1f} {k:5d} {k*c:5d}")
print("# register + shared-memory cluster kernel (fps_cluster_big_kernel), any cluster size")
for t, p in [(512, 44), (512, 48), (512, 52)]:
    for c in range(2, 17):
        k = lib.pn2_fps_cluster_capacity(t, p, c)
        print(f"{t:5d} {p:3d} {c:4d} {3*(p-(12 if p == 48 else 16))*t*4/1024:7.1f} {k:5d} {k*c:5d}")
