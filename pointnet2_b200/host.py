"""Host-buffer entry to the set-abstraction path.

The reference is driven from numpy through ``sess.run(feed_dict=...)`` (train.py:226-231): inputs
start in host memory and results come back to host memory.  ``SetAbstractionHost`` is that call
for one SSG sampling+grouping layer (farthest_point_sample + gather_point + query_ball_point +
group_point(xyz), utils/pointnet_util.py:40-45): it owns pinned host staging buffers and a device
workspace, and each ``run`` issues H2D copy -> 4 kernels -> D2H copies on one stream through the
C-ABI ``pn2_sa_layer_host``.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib


class SetAbstractionHost:
    def __init__(self, b: int, n: int, npoint: int, radius: float, nsample: int, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("SetAbstractionHost needs a CUDA device: pointnet2_b200 has no CPU path")
        self.b, self.n, self.m, self.radius, self.s = int(b), int(n), int(npoint), float(radius), int(nsample)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = _lib.load()
        ws = int(self.lib.pn2_sa_layer_workspace_bytes(self.b, self.n, self.m, self.s))
        if ws <= 0:
            raise ValueError("SetAbstractionHost expects positive b, n, npoint, nsample")
        self.workspace = torch.empty(ws, dtype=torch.uint8, device=self.device)
        pin = dict(pin_memory=True)
        self.h_xyz = torch.empty((self.b, self.n, 3), dtype=torch.float32, **pin)
        self.h_new_xyz = torch.empty((self.b, self.m, 3), dtype=torch.float32, **pin)
        self.h_idx = torch.empty((self.b, self.m, self.s), dtype=torch.int32, **pin)
        self.h_pts_cnt = torch.empty((self.b, self.m), dtype=torch.int32, **pin)
        self.h_grouped_xyz = torch.empty((self.b, self.m, self.s, 3), dtype=torch.float32, **pin)
        self.h2d_bytes = self.h_xyz.numel() * 4
        self.d2h_bytes = 4 * (self.h_new_xyz.numel() + self.h_idx.numel() + self.h_pts_cnt.numel()
                              + self.h_grouped_xyz.numel())

    def launch(self, stream: torch.cuda.Stream | None = None) -> None:
        """Enqueue copy-in, the four kernels and copy-out for whatever is in ``self.h_xyz``."""
        with torch.cuda.device(self.device):
            st = stream if stream is not None else torch.cuda.current_stream(self.device)
            rc = self.lib.pn2_sa_layer_host(
                self.b, self.n, self.m, self.radius, self.s,
                ctypes.c_void_p(self.h_xyz.data_ptr()), ctypes.c_void_p(self.h_new_xyz.data_ptr()),
                ctypes.c_void_p(self.h_idx.data_ptr()), ctypes.c_void_p(self.h_pts_cnt.data_ptr()),
                ctypes.c_void_p(self.h_grouped_xyz.data_ptr()), ctypes.c_void_p(self.workspace.data_ptr()),
                ctypes.c_size_t(self.workspace.numel()), ctypes.c_void_p(st.cuda_stream))
        _lib.check(rc, "pn2_sa_layer_host")

    def run(self, xyz: np.ndarray):
        """xyz: (b,n,3) float32 numpy array. Returns numpy (new_xyz, idx, pts_cnt, grouped_xyz)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        if xyz.shape != (self.b, self.n, 3):
            raise ValueError(f"expected xyz of shape {(self.b, self.n, 3)}, got {xyz.shape}")
        self.h_xyz.numpy()[...] = xyz
        self.launch()
        torch.cuda.current_stream(self.device).synchronize()
        return (self.h_new_xyz.numpy().copy(), self.h_idx.numpy().copy(), self.h_pts_cnt.numpy().copy(),
                self.h_grouped_xyz.numpy().copy())
