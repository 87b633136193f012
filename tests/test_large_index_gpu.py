"""Tensors with more than 2^32 elements (the reference's 32-bit `int` sizes stop at 2^31): the
64-bit indexing paths of the gather / interpolate kernels, checked through size-independent
properties — every sampled output row equals the gathered source row, including the last rows of the
tensor — because a full oracle pass over 17 GB is not a test.  Skipped on GPUs with < 80 GB free."""
import pytest
import torch

from pointnet2_b200 import _lib
from pointnet2_b200.pointnet_util import group_and_concat
from pointnet2_b200.tf_grouping import group_point
from pointnet2_b200.tf_interpolate import three_interpolate

pytestmark = pytest.mark.gpu


def _need(dev, gib):
    free, _ = torch.cuda.mem_get_info(dev)
    if free < gib * (1 << 30):
        pytest.skip(f"needs {gib} GiB of free device memory")


def _sample_rows(total_rows, dev, k=200000):
    g = torch.Generator(device=dev).manual_seed(7)
    return torch.cat([torch.randint(0, total_rows, (k,), device=dev, generator=g),
                      torch.arange(0, 1000, device=dev), torch.arange(total_rows - 1000, total_rows, device=dev)])


@pytest.mark.parametrize("c", [520, 515])  # vector path (c % 4 == 0) and the odd-width row kernel
def test_group_point_beyond_2_to_32_elements(dev, c):
    _need(dev, 80)
    b, n, m, s = 4, 1536, 16384, 128  # 4 * 2^21 rows of c floats: > 4.3e9 elements, 17 GB
    g = torch.Generator(device=dev).manual_seed(1)
    points = torch.randn((b, n, c), device=dev, generator=g)
    idx = torch.randint(0, n, (b, m, s), device=dev, generator=g, dtype=torch.int32)
    out = group_point(points, idx)
    assert out.numel() > (1 << 32)
    rows = _sample_rows(b * m * s, dev)
    want = points[rows // (m * s), idx.view(-1)[rows].long()]
    assert torch.equal(out.view(-1, c)[rows], want)
    if c % 4 == 0:  # the backward kernel at the same size: scatter-add of all 8.4 M rows (float atomics)
        grad_points = torch.zeros_like(points)
        rc = _lib.load().pn2_group_point_grad(b, n, c, m, s, out.data_ptr(), idx.data_ptr(), grad_points.data_ptr(), None)
        assert rc == 0
        flat = (torch.arange(b, device=dev).view(b, 1, 1) * n + idx.long()).view(-1)
        want_g = torch.zeros((b * n, c), device=dev).index_add_(0, flat, out.view(-1, c))
        torch.cuda.synchronize()
        assert torch.allclose(grad_points.view(-1, c), want_g, rtol=2e-3, atol=5e-2)
    del out


def test_group_concat_beyond_2_to_32_elements(dev):
    _need(dev, 80)
    b, n, c, m, s = 4, 1536, 512, 16384, 128
    g = torch.Generator(device=dev).manual_seed(2)
    xyz = torch.rand((b, n, 3), device=dev, generator=g)
    new_xyz = torch.rand((b, m, 3), device=dev, generator=g)
    points = torch.randn((b, n, c), device=dev, generator=g)
    idx = torch.randint(0, n, (b, m, s), device=dev, generator=g, dtype=torch.int32)
    out, gxyz = group_and_concat(xyz, new_xyz, points, idx, xyz_first=False)  # MSG order: [features, xyz]
    assert out.numel() > (1 << 32)
    rows = _sample_rows(b * m * s, dev)
    cloud, src = rows // (m * s), idx.view(-1)[rows].long()
    centre = new_xyz.view(-1, 3)[rows // s]
    want_xyz = xyz[cloud, src] - centre
    got = out.view(-1, c + 3)[rows]
    assert torch.equal(got[:, :c], points[cloud, src])
    assert torch.equal(got[:, c:], want_xyz)
    assert torch.equal(gxyz.view(-1, 3)[rows], want_xyz)


def test_three_interpolate_beyond_2_to_32_elements(dev):
    _need(dev, 80)
    b, n, m, c = 2, 1 << 21, 512, 1032  # 2 * 2^21 rows of 1032 floats: > 4.3e9 elements
    g = torch.Generator(device=dev).manual_seed(3)
    points = torch.randn((b, m, c), device=dev, generator=g)
    idx = torch.randint(0, m, (b, n, 3), device=dev, generator=g, dtype=torch.int32)
    w = torch.rand((b, n, 3), device=dev, generator=g)
    w = w / w.sum(dim=2, keepdim=True)
    out = three_interpolate(points, idx, w)
    assert out.numel() > (1 << 32)
    rows = _sample_rows(b * n, dev)
    cloud = rows // n
    ii, ww = idx.view(-1, 3)[rows].long(), w.view(-1, 3)[rows]
    want = (points[cloud, ii[:, 0]] * ww[:, 0:1] + points[cloud, ii[:, 1]] * ww[:, 1:2]) + points[cloud, ii[:, 2]] * ww[:, 2:3]
    assert torch.allclose(out.view(-1, c)[rows], want, rtol=0, atol=1e-5)
