"""Packed-chain probe: for every single-CTA plan the planner uses, run farthest point sampling with the plain chain
(override cluster = -1) and with fps_chain_packed (cluster = -2), require identical picks and coordinates, and time
both (CUDA events, median of 9 launches after warm-up).  Writes gpurun_out/fps_packed.json; exit code 0 and the
line `DECISION packed=1` only if every case is bit-identical and the packed chain is faster on the cfg2 shape."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointnet2_b200 import _lib, workloads as W
from pointnet2_b200.tf_sampling import farthest_point_sample_and_gather

dev = torch.device("cuda:0")
lib = _lib.load()
CASES = [  # gen, b, n, m
    ("U", 32, 4096, 1024), ("D", 32, 4096, 1024), ("S", 32, 4096, 1024),   # cfg2 and its two harder clouds
    ("D", 16, 8192, 1024), ("U", 16, 8192, 1024),                          # cfg4 SA1
    ("S", 32, 1024, 512), ("U", 32, 2048, 512), ("U", 8, 1000, 999), ("D", 4, 3000, 3000),
]


def plan(b, n):
    t, p, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib.pn2_set_fps_config(0, 0, 0)
    lib.pn2_fps_plan(b, n, ctypes.byref(t), ctypes.byref(p), ctypes.byref(c))
    return t.value, p.value, c.value


def timed(m, xyz, reps=9):
    for _ in range(3):
        farthest_point_sample_and_gather(m, xyz)
    ts = []
    for _ in range(reps):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        farthest_point_sample_and_gather(m, xyz)
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


rows, ok = [], True
for gen, b, n, m in CASES:
    xyz = torch.from_numpy(W.DISTRIBUTIONS[gen](b, n, 77)).to(dev)
    t, p, c = plan(b, n)
    row = dict(gen=gen, b=b, n=n, m=m, threads=t, ppt=p, cluster=c)
    if c != 1 or p < 8:
        row["skipped"] = "no packed instantiation for this plan"
        rows.append(row)
        print(row, flush=True)
        continue
    res = {}
    for name, code in (("plain", -1), ("packed", -2)):
        lib.pn2_set_fps_config(t, p, code)
        try:
            idx, nx = farthest_point_sample_and_gather(m, xyz)
            torch.cuda.synchronize()
            res[name] = (idx.clone(), nx.clone(), timed(m, xyz))
        finally:
            lib.pn2_set_fps_config(0, 0, 0)
    same = bool(torch.equal(res["plain"][0], res["packed"][0]) and torch.equal(res["plain"][1], res["packed"][1]))
    ok = ok and same
    row.update(identical=same, plain_ms=round(res["plain"][2], 5), packed_ms=round(res["packed"][2], 5),
               ratio=round(res["packed"][2] / res["plain"][2], 4))
    rows.append(row)
    print(row, flush=True)

cfg2 = [r for r in rows if r["n"] == 4096 and r["gen"] == "U"][0]
faster = ok and cfg2.get("ratio", 9.0) < 0.98

# other single-CTA shapes for the cfg2 cloud under the packed chain (information for the planner)
xyz = torch.from_numpy(W.cloud_uniform(32, 4096, 77)).to(dev)
alt = []
for t, p in ((256, 16), (128, 32), (512, 8)):
    lib.pn2_set_fps_config(t, p, -2)
    try:
        alt.append(dict(threads=t, ppt=p, packed_ms=round(timed(1024, xyz), 5)))
    finally:
        lib.pn2_set_fps_config(0, 0, 0)
    print("alt", alt[-1], flush=True)

# cluster kernels: the planner's plan with the plain (+2) and the packed (+1) update
CLUSTER_CASES = [("U", 8, 16384, 256), ("D", 8, 16384, 256), ("U", 8, 65536, 256), ("U", 1, 65536, 256), ("D", 2, 65536, 128),
                 ("U", 8, 262144, 64), ("U", 1, 262144, 128), ("D", 2, 262144, 64), ("U", 8, 131072, 64), ("U", 2, 400000, 48)]
crows, cok = [], True
for gen, b, n, m in CLUSTER_CASES:
    xyz = torch.from_numpy(W.DISTRIBUTIONS[gen](b, n, 78)).to(dev)
    t, p, c = plan(b, n)
    row = dict(gen=gen, b=b, n=n, m=m, threads=t, ppt=p, cluster=c)
    if c < 2 or p % 4:
        row["skipped"] = "no packed instantiation for this plan"
        crows.append(row)
        print(row, flush=True)
        continue
    res = {}
    for name, bit in (("plain", 2), ("packed", 1)):
        lib.pn2_set_fps_config(t + bit, p, c)
        try:
            idx, nx = farthest_point_sample_and_gather(m, xyz)
            torch.cuda.synchronize()
            res[name] = (idx.clone(), nx.clone(), timed(m, xyz, reps=5))
        finally:
            lib.pn2_set_fps_config(0, 0, 0)
    same = bool(torch.equal(res["plain"][0], res["packed"][0]) and torch.equal(res["plain"][1], res["packed"][1]))
    cok = cok and same
    row.update(identical=same, plain_ms=round(res["plain"][2], 5), packed_ms=round(res["packed"][2], 5),
               ratio=round(res["packed"][2] / res["plain"][2], 4))
    crows.append(row)
    print(row, flush=True)
ratios = [r["ratio"] for r in crows if "ratio" in r]
gmean = 1.0
for r in ratios:
    gmean *= r
gmean = gmean ** (1.0 / max(len(ratios), 1))
cfaster = cok and len(ratios) > 0 and gmean < 0.98 and max(ratios) < 1.03
out = dict(rows=rows, all_identical=ok, packed_default=int(faster), alt_plans_cfg2=alt, cluster_rows=crows,
           cluster_all_identical=cok, cluster_ratio_geomean=round(gmean, 4), packed_cluster_default=int(cfaster))
print("DECISION_CLUSTER packed=%d" % int(cfaster))
ok = ok and cok
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fps_packed.json", "w"), indent=1)
print("DECISION packed=%d" % int(faster))
sys.exit(0 if ok else 1)
