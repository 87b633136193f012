"""NUMA placement of the host side of a rank (one process per GPU).

The host-buffer path is PCIe / host-memory bound; on a two-socket HGX box GPUs 0-3 hang off one
socket and 4-7 off the other, and a rank whose pinned staging buffers (or whose submitting thread)
sit on the far socket pays the inter-socket link on every copy.  The reference never has to care
(its CPU ops run where TensorFlow puts them); a one-process-per-GPU runtime does.

``prefer_node_of(device)`` is a context manager: pages first touched inside it are allocated on
the NUMA node of that GPU (``set_mempolicy(MPOL_PREFERRED)``, restored on exit).  ``bind_cpus(device)``
restricts the calling process to that node's CPUs.  Both are best-effort: on a single-node host, in
a container without the sysfs entries, or on a non-Linux kernel they do nothing and say so in
``status()``.
"""
from __future__ import annotations

import ctypes
import os
from contextlib import contextmanager

_SYS_set_mempolicy = 238  # x86_64
_MPOL_DEFAULT, _MPOL_PREFERRED = 0, 1
_status: dict = {}


def _pci_address(device) -> str | None:
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        dom = getattr(p, "pci_domain_id", 0)
        return f"{dom:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        return None


def node_of(device) -> int | None:
    """NUMA node of a CUDA device from sysfs, or None when unknown / single-node."""
    addr = _pci_address(device)
    if addr is None:
        return None
    try:
        node = int(open(f"/sys/bus/pci/devices/{addr}/numa_node").read().strip())
    except (OSError, ValueError):
        return None
    if node < 0 or not os.path.isdir(f"/sys/devices/system/node/node{node}"):
        return None
    return node


def _set_mempolicy(mode: int, node: int | None) -> bool:
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        if node is None:
            rc = libc.syscall(_SYS_set_mempolicy, ctypes.c_int(mode), ctypes.c_void_p(0), ctypes.c_ulong(0))
        else:
            nwords = node // 64 + 1
            mask = (ctypes.c_ulong * nwords)()
            mask[node // 64] = 1 << (node % 64)
            rc = libc.syscall(_SYS_set_mempolicy, ctypes.c_int(mode), mask, ctypes.c_ulong(64 * nwords + 1))
        return rc == 0
    except Exception:
        return False


@contextmanager
def prefer_node_of(device):
    node = node_of(device) if os.environ.get("PN2_NUMA", "1") != "0" else None
    ok = node is not None and _set_mempolicy(_MPOL_PREFERRED, node)
    _status["mempolicy"] = {"node": node, "applied": bool(ok)}
    try:
        yield node if ok else None
    finally:
        if ok:
            _set_mempolicy(_MPOL_DEFAULT, None)


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_cpus(device) -> bool:
    """Restrict this process to the CPUs of the GPU's NUMA node (intersected with its current mask)."""
    node = node_of(device) if os.environ.get("PN2_NUMA", "1") != "0" else None
    applied = False
    if node is not None:
        try:
            cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                applied = True
        except (OSError, ValueError):
            applied = False
    _status["cpus"] = {"node": node, "applied": applied}
    return applied


def status() -> dict:
    return dict(_status)
