"""Generate golden vectors for tests/golden/ from the REFERENCE'S OWN code (TEST INFRASTRUCTURE).

Runs on a GPU box (the reference's sampling/grouping ops are GPU-only, tf_sampling.cpp:123,
tf_grouping.cpp:106): executes the reference CUDA kernels rebuilt unmodified for sm_100a
(oracle/_ref/libref_{sampling,grouping}.so) and the reference CPU functions
(oracle/_ref/libref_cpu.so) on small seeded inputs and stores inputs + outputs as .npz.

    gpurun -- python oracle/gen_golden.py gpurun_out/golden      # on the box
    cp gpurun_out/golden/*.npz tests/golden/                     # here, then commit

The fixtures pin the oracle (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_parity_gpu.py, GPU).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle as O  # noqa: E402
from pointnet2_b200 import workloads as W  # noqa: E402


def prob_cases():
    """ProbSample inputs: (probabilities (b,n), uniform draws (b,m))."""
    rng = np.random.RandomState(150)
    cases = {}
    for name, n, m in (("prob_n5", 5, 64), ("prob_n1000", 1000, 256), ("prob_two_chunks", 9000, 256)):
        p = rng.random_sample((2, n)).astype(np.float32)
        if n >= 1000:
            p[:, rng.randint(0, n, n // 3)] = 0.0  # plateaus in the cumulative sum
        r = rng.random_sample((2, m)).astype(np.float32)
        r[:, 0], r[:, 1] = 0.0, 1.0
        cases[name] = (p, r)
    return cases


def main_prob(outdir: str) -> None:
    """Only the ProbSample fixtures (added after the first fixture set was committed)."""
    import torch
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    os.makedirs(outdir, exist_ok=True)
    for name, (p, r) in prob_cases().items():
        out, cum = O.refcuda_prob_sample(torch.from_numpy(p).to(dev), torch.from_numpy(r).to(dev), return_cumsum=True)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), inp=p, inpr=r, cumsum=cum.cpu().numpy(), out=out.cpu().numpy())
        print("wrote", name)


def main(outdir: str) -> None:
    import torch
    assert torch.cuda.is_available(), "gen_golden.py needs a GPU: the reference's FPS/ball-query kernels are GPU-only"
    dev = torch.device("cuda:0")
    os.makedirs(outdir, exist_ok=True)

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def save(name, **arrs):
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **arrs)
        print("wrote", name, {k: v.shape for k, v in arrs.items()})

    # ---- FPS + gather_point (reference CUDA) --------------------------------------------------
    fps_cases = {
        "fps_uniform": (W.cloud_uniform(4, 1024, 100), 256),      # cfg1-like
        "fps_dups": (W.cloud_duplicates(2, 1500, 101), 700),      # real ties; npoint > #distinct
        "fps_small_n": (W.cloud_uniform(3, 100, 102), 50),        # n < 512: empty slots
        "fps_beyond_smem": (W.cloud_surface(2, 5000, 103), 64),   # n > 3072: reference's global path
        "fps_npoint_gt_n": (W.cloud_uniform(2, 40, 104), 64),     # npoint > n
    }
    for name, (xyz, m) in fps_cases.items():
        idx = O.refcuda_fps(m, t(xyz))
        new_xyz = O.refcuda_gather_point(t(xyz), idx)
        save(name, xyz=xyz, npoint=np.int32(m), idx=idx.cpu().numpy(), new_xyz=new_xyz.cpu().numpy())

    # ---- ball query + group_point (reference CUDA) -------------------------------------------
    bq_cases = {
        "bq_uniform_r0.1": (W.cloud_uniform(4, 1024, 110), 128, 0.1, 32),   # sparse: rows padded
        "bq_uniform_r0.4": (W.cloud_uniform(2, 1024, 111), 64, 0.4, 16),    # dense: early exit
        "bq_dups_r0.2": (W.cloud_duplicates(2, 2500, 112), 96, 0.2, 32),    # coincident points, tile boundary
        "bq_surface_r0.2_s128": (W.cloud_surface(2, 700, 113), 50, 0.2, 128),
    }
    for name, (xyz, m, r, s) in bq_cases.items():
        fidx = O.refcuda_fps(m, t(xyz))
        new_xyz = O.refcuda_gather_point(t(xyz), fidx)
        idx, cnt = O.refcuda_query_ball_point(r, s, t(xyz), new_xyz)
        grouped = O.refcuda_group_point(t(xyz), idx)
        feats = W.features(xyz.shape[0], xyz.shape[1], 16, 7)
        gfeat = O.refcuda_group_point(t(feats), idx)
        save(name, xyz=xyz, new_xyz=new_xyz.cpu().numpy(), radius=np.float32(r), nsample=np.int32(s),
             idx=idx.cpu().numpy(), pts_cnt=cnt.cpu().numpy(), grouped_xyz=grouped.cpu().numpy(),
             feats=feats, grouped_feats=gfeat.cpu().numpy())
    # queries that are NOT data points and may have empty balls (rows left to the pre-zeroed buffer)
    xyz = W.cloud_uniform(2, 300, 114)
    q = (W.cloud_uniform(2, 40, 115) * 1.6 - 0.3).astype(np.float32)
    idx, cnt = O.refcuda_query_ball_point(0.15, 8, t(xyz), t(q))
    save("bq_free_queries", xyz=xyz, new_xyz=q, radius=np.float32(0.15), nsample=np.int32(8),
         idx=idx.cpu().numpy(), pts_cnt=cnt.cpu().numpy())

    # ---- selection sort (reference CUDA) -----------------------------------------------------
    dist = np.random.RandomState(120).random_sample((2, 6, 40)).astype(np.float32)
    dist[0, 0, 5] = dist[0, 0, 17]  # a tie
    outi, out = O.refcuda_selection_sort(7, t(dist))
    save("selection_sort", dist=dist, k=np.int32(7), outi=outi.cpu().numpy(), out=out.cpu().numpy())
    # the reference's only known-answer case (test/selection_sort.cpp:68-92): dist = 10 - i
    kat = np.tile((10.0 - np.arange(4, dtype=np.float32)), (2, 2, 1))
    outi, out = O.refcuda_selection_sort(3, t(kat))
    save("selection_sort_kat", dist=kat, k=np.int32(3), outi=outi.cpu().numpy(), out=out.cpu().numpy())

    # ---- three_nn / three_interpolate (reference CPU) -----------------------------------------
    for name, (n, m, c, seed) in {"interp_1024_256_c64": (1024, 256, 64, 130), "interp_64_16_c5": (64, 16, 5, 131),
                                  "interp_m2": (20, 2, 8, 132)}.items():
        xyz1 = W.cloud_duplicates(2, n, seed, drop=False) if "1024" in name else W.cloud_uniform(2, n, seed)
        xyz2 = xyz1[:, :m].copy() if "1024" in name else W.cloud_uniform(2, m, seed + 1)
        pts = W.features(2, m, c, seed + 2)
        d, i = O.refcpu_three_nn(xyz1, xyz2)
        dd = np.maximum(d, 1e-10)
        w = ((1.0 / dd) / (1.0 / dd).sum(axis=2, keepdims=True)).astype(np.float32)
        w = np.nan_to_num(w, nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)
        out = O.refcpu_three_interpolate(pts, i, w)
        go = W.features(2, n, c, seed + 3)
        gp = O.refcpu_three_interpolate_grad(pts.shape, i, w, go)
        save(name, xyz1=xyz1, xyz2=xyz2, points=pts, dist=d, idx=i, weight=w, out=out, grad_out=go, grad_points=gp)

    # ---- gradients of gather/group (reference CUDA, float atomics: compare with tolerance) ----
    xyz = W.cloud_uniform(2, 256, 140)
    fidx = O.refcuda_fps(32, t(xyz))
    new_xyz = O.refcuda_gather_point(t(xyz), fidx)
    idx, _ = O.refcuda_query_ball_point(0.3, 16, t(xyz), new_xyz)
    feats = W.features(2, 256, 12, 141)
    go = W.features(2, 32 * 16, 12, 142).reshape(2, 32, 16, 12)
    gp = O.refcuda_group_point_grad(feats.shape, idx, t(go))
    og = W.features(2, 32, 3, 143)
    gi = O.refcuda_gather_point_grad(xyz.shape, fidx, t(og))
    save("grads", xyz=xyz, fps_idx=fidx.cpu().numpy(), idx=idx.cpu().numpy(), feats=feats, grad_out=go,
         grad_points=gp.cpu().numpy(), out_g=og, inp_g=gi.cpu().numpy())
    main_prob(outdir)
    print("golden vectors written to", outdir)


if __name__ == "__main__":
    if "--prob-only" in sys.argv:
        sys.argv.remove("--prob-only")
        main_prob(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(HERE), "gpurun_out", "golden"))
        sys.exit(0)
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(HERE), "gpurun_out", "golden"))
