"""Time FPS single-CTA / cluster configurations (marginal cost per step) via pn2_set_fps_config."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet2_b200 import _lib, workloads as W
from pointnet2_b200.tf_sampling import farthest_point_sample_and_gather
dev = torch.device("cuda:0"); lib = _lib.load()
CASES = [
    (8, 262144, 64, [(0, 0, 0), (512, 32, 16), (512, 52, 10), (512, 44, 12)]),
    (1, 262144, 64, [(0, 0, 0), (512, 44, 12), (512, 52, 10)]),
    (7, 262144, 64, [(0, 0, 0)]),
    (2, 400000, 64, [(0, 0, 0)]),
    (10, 200000, 64, [(0, 0, 0), (512, 32, 16), (512, 44, 9), (512, 44, 10)]),
    (12, 150000, 64, [(0, 0, 0), (512, 32, 16)]),
    (40, 60000, 64, [(0, 0, 0), (512, 44, 3), (128, 32, 16)]),
    (8, 131072, 128, [(0, 0, 0), (512, 44, 6)]),
    (8, 65536, 256, [(0, 0, 0)]),
]
rows = []
for b, n, m, cfgs in CASES:
    xyz = torch.from_numpy(W.cloud_uniform(b, n, 100)).to(dev)
    ref = None
    for cfg in cfgs:
        if not cfg[0]:
            import ctypes
            t_, p_, c_ = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            lib.pn2_set_fps_config(0, 0, 0); lib.pn2_fps_plan(b, n, ctypes.byref(t_), ctypes.byref(p_), ctypes.byref(c_))
            print("  default plan for", (b, n), "->", (t_.value, p_.value, c_.value), flush=True)
        lib.pn2_set_fps_config(*cfg)
        cap = lib.pn2_fps_cluster_capacity(*cfg) if cfg[0] else None
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
        rows.append(dict(b=b, n=n, cfg=cfg, us_per_step=round((t2 - t1) / m * 1e3, 4), ms_m=round(t1, 4), same_as_default=same, co_resident_clusters=cap))
        print(rows[-1], flush=True)
    lib.pn2_set_fps_config(0, 0, 0)
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "fps_cluster.json"), "w"), indent=1)
