"""Host-buffer entry to the set-abstraction path.

The reference is driven from numpy through ``sess.run(feed_dict=...)`` (train.py:226-231): inputs
start in host memory and results come back to host memory.  ``SetAbstractionHost`` is that call
for one SSG sampling+grouping layer (farthest_point_sample + gather_point + query_ball_point +
group_point(xyz), utils/pointnet_util.py:40-45): it owns pinned host staging buffers and a device
workspace, and each ``run`` issues H2D copy -> the overlapped sampling + grouping kernels (sa_fused.cu) -> D2H copies on one
stream through the C-ABI ``pn2_sa_layer_host``.

``SetAbstractionPipeline`` is the same call for a STREAM of batches (the reference's training loop
feeds one batch per ``sess.run`` while its input queue prepares the next, train.py:207-231): a ring
of ``depth`` such sessions, each on its own CUDA stream, so batch k+1's copy-in and sampling overlap
batch k's grouping and copy-out.  One FPS launch occupies one SM per cloud (32 of 148 at the
benchmark's batch size), so consecutive batches really do run side by side.
"""
from __future__ import annotations

import collections
import ctypes

import numpy as np
import torch

from . import _lib, numa


class SetAbstractionHost:
    def __init__(self, b: int, n: int, npoint: int, radius: float, nsample: int, device=None, want_grouped: bool = True):
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
        # pinned staging buffers are placed on the NUMA node the GPU hangs off (the copies of a rank
        # whose buffers sit on the other socket cross the inter-socket link and do not scale)
        with numa.prefer_node_of(self.device):
            self.h_xyz = torch.empty((self.b, self.n, 3), dtype=torch.float32, **pin)
            self.h_new_xyz = torch.empty((self.b, self.m, 3), dtype=torch.float32, **pin)
            self.h_idx = torch.empty((self.b, self.m, self.s), dtype=torch.int32, **pin)
            self.h_pts_cnt = torch.empty((self.b, self.m), dtype=torch.int32, **pin)
            # want_grouped=False: grouped_xyz is neither computed nor copied back (it is xyz[idx], which a
            # host-side caller can regroup itself) — 3/4 of the device-to-host bytes of the layer
            self.h_grouped_xyz = (torch.empty((self.b, self.m, self.s, 3), dtype=torch.float32, **pin)
                                  if want_grouped else None)
            for t in (self.h_xyz, self.h_new_xyz, self.h_idx, self.h_pts_cnt, self.h_grouped_xyz):
                if t is not None:
                    t.zero_()  # first touch under the NUMA preference
        self.h2d_bytes = self.h_xyz.numel() * 4
        self.d2h_bytes = 4 * (self.h_new_xyz.numel() + self.h_idx.numel() + self.h_pts_cnt.numel()
                              + (self.h_grouped_xyz.numel() if want_grouped else 0))

    def launch(self, stream: torch.cuda.Stream | None = None) -> None:
        """Enqueue copy-in, the four kernels and copy-out for whatever is in ``self.h_xyz``."""
        with torch.cuda.device(self.device):
            st = stream if stream is not None else torch.cuda.current_stream(self.device)
            rc = self.lib.pn2_sa_layer_host(
                self.b, self.n, self.m, self.radius, self.s,
                ctypes.c_void_p(self.h_xyz.data_ptr()), ctypes.c_void_p(self.h_new_xyz.data_ptr()),
                ctypes.c_void_p(self.h_idx.data_ptr()), ctypes.c_void_p(self.h_pts_cnt.data_ptr()),
                ctypes.c_void_p(self.h_grouped_xyz.data_ptr() if self.h_grouped_xyz is not None else 0),
                ctypes.c_void_p(self.workspace.data_ptr()),
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
                self.h_grouped_xyz.numpy().copy() if self.h_grouped_xyz is not None else None)


class SetAbstractionPipeline:
    """A ring of ``depth`` SetAbstractionHost sessions on private streams.

    Usage::
        pipe = SetAbstractionPipeline(b, n, npoint, radius, nsample, depth=2)
        for batch in batches:
            if pipe.full():
                new_xyz, idx, pts_cnt, grouped_xyz = pipe.collect()   # oldest batch, in order
            pipe.input_buffer()[...] = batch                          # fill the pinned slot
            pipe.submit()
        while pipe.pending():
            ... = pipe.collect()

    ``collect`` returns numpy views of the slot's pinned output buffers; they stay valid until the
    next ``submit`` that reuses the slot (``depth`` submits later).
    """

    def __init__(self, b: int, n: int, npoint: int, radius: float, nsample: int, depth: int = 2, device=None,
                 want_grouped: bool = True):
        if depth < 1:
            raise ValueError("SetAbstractionPipeline expects depth >= 1")
        self.slots = [SetAbstractionHost(b, n, npoint, radius, nsample, device=device, want_grouped=want_grouped)
                      for _ in range(int(depth))]
        self.device = self.slots[0].device
        self.streams = [torch.cuda.Stream(self.device) for _ in self.slots]
        self.done = [torch.cuda.Event() for _ in self.slots]
        self._next = 0
        self._inflight: collections.deque[int] = collections.deque()
        self.h2d_bytes, self.d2h_bytes = self.slots[0].h2d_bytes, self.slots[0].d2h_bytes

    @property
    def depth(self) -> int:
        return len(self.slots)

    def pending(self) -> int:
        return len(self._inflight)

    def full(self) -> bool:
        return len(self._inflight) == len(self.slots)

    def input_buffer(self) -> np.ndarray:
        """The pinned (b,n,3) float32 input of the slot the next ``submit`` will use."""
        if self.full():
            raise RuntimeError("SetAbstractionPipeline is full: collect() the oldest batch first")
        return self.slots[self._next].h_xyz.numpy()

    def submit(self, xyz: np.ndarray | None = None, after: torch.cuda.Event | None = None) -> int:
        """Enqueue one batch (``xyz`` is copied into the slot's pinned input when given; otherwise
        whatever ``input_buffer()`` holds is used). Returns the slot index. Never blocks."""
        if self.full():
            raise RuntimeError("SetAbstractionPipeline is full: collect() the oldest batch first")
        i = self._next
        slot = self.slots[i]
        if xyz is not None:
            xyz = np.ascontiguousarray(xyz, dtype=np.float32)
            if xyz.shape != (slot.b, slot.n, 3):
                raise ValueError(f"expected xyz of shape {(slot.b, slot.n, 3)}, got {xyz.shape}")
            slot.h_xyz.numpy()[...] = xyz
        if after is not None:
            self.streams[i].wait_event(after)
        slot.launch(self.streams[i])
        self.done[i].record(self.streams[i])
        self._inflight.append(i)
        self._next = (i + 1) % len(self.slots)
        return i

    def collect(self):
        """Wait for the OLDEST submitted batch; returns (new_xyz, idx, pts_cnt, grouped_xyz) views."""
        if not self._inflight:
            raise RuntimeError("SetAbstractionPipeline.collect() with nothing submitted")
        i = self._inflight.popleft()
        self.done[i].synchronize()
        s = self.slots[i]
        return (s.h_new_xyz.numpy(), s.h_idx.numpy(), s.h_pts_cnt.numpy(),
                s.h_grouped_xyz.numpy() if s.h_grouped_xyz is not None else None)
