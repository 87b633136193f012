"""Time FPS single-CTA / cluster configurations (marginal cost per step) via pn2_set_fps_config."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet2_b200 import _lib, workloads as W
from pointnet2_b200.tf_sampling import farthest_point_sample_and_gather
dev = torch.device("cuda:0"); lib = _lib.load()
CASES = [
    (32, 4096, 512, [(0, 0, 0), (128, 8, 4), (256, 8, 2), (128, 4, 8)]),
    (16, 8192, 512, [(0, 0, 0), (128, 16, 4), (256, 16, 2), (128, 8, 8)]),
    (8, 16384, 512, [(0, 0, 0), (128, 8, 16), (256, 8, 8), (128, 16, 8)]),
    (8, 65536, 256, [(0, 0, 0), (128, 32, 16), (256, 16, 16)]),
    (1, 65536, 256, [(0, 0, 0)]),
    (8, 262144, 64, [(0, 0, 0)]),
]
rows = []
for b, n, m, cfgs in CASES:
    xyz = torch.from_numpy(W.cloud_uniform(b, n, 100)).to(dev)
    ref = None
    for cfg in cfgs:
        lib.pn2_set_fps_config(*cfg)
        try:
            idx, _ = farthest_point_sample_and_gather(m, xyz)
        except Exception as e:
            rows.append(dict(b=b, n=n, cfg=cfg, error=str(e)[:80])); print(rows[-1], flush=True); continue
        torch.cuda.synchronize()
        if ref is None: ref = idx.clone()
        same = bool(torch.equal(idx, ref))
        def t(mm):
            for _ in range(2): farthest_point_sample_and_gather(mm, xyz)
            ts = []
            for _ in range(7):
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); farthest_point_sample_and_gather(mm, xyz); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
            ts.sort(); return ts[len(ts) // 2]
        t1, t2 = t(m), t(2 * m)   # marginal cost per step removes the set-up
        rows.append(dict(b=b, n=n, cfg=cfg, us_per_step=round((t2 - t1) / m * 1e3, 4), ms_m=round(t1, 4), same_as_default=same))
        print(rows[-1], flush=True)
    lib.pn2_set_fps_config(0, 0, 0)
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "fps_cluster.json"), "w"), indent=1)
