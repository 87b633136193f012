"""GPU tests of the layer glue (torch twin of utils/pointnet_util.py): the fused path, the
unfused reference op sequence and an oracle-side numpy composite must all agree."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from pointnet2_b200 import workloads as W
from pointnet2_b200.host import SetAbstractionHost, SetAbstractionPipeline
from pointnet2_b200.pointnet_util import (pointnet_fp_module, pointnet_sa_module, pointnet_sa_module_msg,
                                          sample_and_group, sample_and_group_all)

pytestmark = pytest.mark.gpu


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


def oracle_sample_and_group(npoint, radius, nsample, xyz, points, xyz_first=True):
    """numpy composite of utils/pointnet_util.py:40-54 over the oracle ops."""
    new_xyz = O.oracle_gather_point(xyz, O.oracle_fps(npoint, xyz))
    idx, _ = O.oracle_query_ball_point(radius, nsample, xyz, new_xyz)
    gx = O.oracle_group_point(xyz, idx) - new_xyz[:, :, None, :]
    if points is None:
        return new_xyz, gx, idx, gx
    gp = O.oracle_group_point(points, idx)
    new_points = np.concatenate([gx, gp] if xyz_first else [gp, gx], axis=-1)
    return new_xyz, new_points, idx, gx


@pytest.mark.parametrize("c", [0, 6, 64])
@pytest.mark.parametrize("fused", [True, False])
def test_sample_and_group_matches_oracle_composite(dev, c, fused):
    xyz = W.cloud_surface(3, 900, 201)
    pts = W.features(3, 900, c, 202) if c else None
    got = sample_and_group(100, 0.25, 24, T(xyz, dev), T(pts, dev) if c else None, fused=fused)
    want = oracle_sample_and_group(100, 0.25, 24, xyz, pts)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(N(g), w)


def test_sample_and_group_use_xyz_false_and_knn(dev):
    xyz = W.cloud_uniform(2, 400, 203)
    pts = W.features(2, 400, 5, 204)
    nx, npts, idx, gx = sample_and_group(50, 0.3, 8, T(xyz, dev), T(pts, dev), use_xyz=False)
    want = oracle_sample_and_group(50, 0.3, 8, xyz, pts)
    np.testing.assert_array_equal(N(npts), O.oracle_group_point(pts, want[2]))
    np.testing.assert_array_equal(N(gx), want[3])
    nx, npts, idx, gx = sample_and_group(50, 0.3, 8, T(xyz, dev), None, knn=True)
    assert idx.shape == (2, 50, 8) and idx.dtype == torch.int32
    # the seed point is its own nearest neighbour
    assert torch.equal(idx[:, :, 0].long(), T(O.oracle_fps(50, xyz), dev).long())


def test_sample_and_group_all(dev):
    xyz = W.cloud_uniform(2, 16, 205)
    pts = W.features(2, 16, 4, 206)
    nx, npts, idx, gx = sample_and_group_all(T(xyz, dev), T(pts, dev))
    assert nx.shape == (2, 1, 3) and float(nx.abs().max()) == 0.0
    np.testing.assert_array_equal(N(npts), np.concatenate([xyz, pts], 2)[:, None])
    np.testing.assert_array_equal(N(idx), np.tile(np.arange(16, dtype=np.int32), (2, 1, 1)))
    np.testing.assert_array_equal(N(gx), xyz[:, None])


def test_msg_module_channel_order_and_values(dev):
    """MSG concatenates [features, xyz] (utils/pointnet_util.py:184), the opposite of SSG (:50)."""
    xyz = W.cloud_surface(2, 512, 207)
    pts = W.features(2, 512, 9, 208)
    radii, ns = [0.2, 0.4], [16, 32]
    new_xyz, feats = pointnet_sa_module_msg(T(xyz, dev), T(pts, dev), 64, radii, ns)
    nf_unfused = pointnet_sa_module_msg(T(xyz, dev), T(pts, dev), 64, radii, ns, fused=False)[1]
    assert torch.equal(feats, nf_unfused)
    want = []
    for r, s in zip(radii, ns):
        _, npts, _, _ = oracle_sample_and_group(64, r, s, xyz, pts, xyz_first=False)
        want.append(npts.max(axis=2))
    np.testing.assert_array_equal(N(feats), np.concatenate(want, -1))
    assert feats.shape == (2, 64, 2 * 12)


def test_sa_module_ssg_pooling(dev):
    xyz = W.cloud_uniform(2, 300, 209)
    pts = W.features(2, 300, 7, 210)
    new_xyz, feats, idx = pointnet_sa_module(T(xyz, dev), T(pts, dev), 40, 0.3, 16)
    _, npts, oidx, _ = oracle_sample_and_group(40, 0.3, 16, xyz, pts)
    np.testing.assert_array_equal(N(feats), npts.max(axis=2))
    np.testing.assert_array_equal(N(idx), oidx)
    _, fa, _ = pointnet_sa_module(T(xyz, dev), T(pts, dev), 40, 0.3, 16, group_all=True)
    np.testing.assert_array_equal(N(fa), np.concatenate([xyz, pts], 2).max(axis=1, keepdims=True))


@pytest.mark.parametrize("fused", [True, False])
def test_fp_module_matches_oracle_composite(dev, fused):
    xyz1, xyz2 = W.cloud_uniform(2, 256, 211), W.cloud_uniform(2, 64, 212)
    p1, p2 = W.features(2, 256, 6, 213), W.features(2, 64, 32, 214)
    got = pointnet_fp_module(T(xyz1, dev), T(xyz2, dev), T(p1, dev), T(p2, dev), fused=fused)
    d, i = O.oracle_three_nn(xyz1, xyz2)
    dd = np.maximum(d, np.float32(1e-10))
    r = (np.float32(1.0) / dd).astype(np.float32)
    w = (r / r.sum(axis=2, keepdims=True, dtype=np.float32)).astype(np.float32)
    want = np.concatenate([O.oracle_three_interpolate(p2, i, w), p1], axis=2)
    assert got.shape == (2, 256, 38)
    assert np.abs(N(got) - want).max() <= 1e-5


def test_fused_group_concat_backward_matches_unfused(dev):
    xyz = W.cloud_uniform(2, 200, 215)
    pts = W.features(2, 200, 8, 216)
    g = T(W.features(2, 30 * 12, 11, 217).reshape(2, 30, 12, 11), dev)
    grads = []
    for fused in (True, False):
        p = T(pts, dev).requires_grad_(True)
        _, npts, _, _ = sample_and_group(30, 0.3, 12, T(xyz, dev), p, fused=fused)
        (npts * g).sum().backward()
        grads.append(N(p.grad))
    np.testing.assert_allclose(grads[0], grads[1], atol=1e-4, rtol=1e-5)


def test_host_buffer_sa_layer_matches_device_path(dev):
    xyz = W.cloud_uniform(4, 1024, 218)
    sess = SetAbstractionHost(4, 1024, 128, 0.2, 16)
    new_xyz, idx, cnt, grouped = sess.run(xyz)
    wn, _, widx, _ = oracle_sample_and_group(128, 0.2, 16, xyz, None)
    np.testing.assert_array_equal(new_xyz, wn)
    np.testing.assert_array_equal(idx, widx)
    np.testing.assert_array_equal(grouped, O.oracle_group_point(xyz, widx))
    assert (cnt == O.oracle_query_ball_point(0.2, 16, xyz, wn)[1]).all()
    assert sess.h2d_bytes == xyz.nbytes and sess.d2h_bytes == new_xyz.nbytes + idx.nbytes + cnt.nbytes + grouped.nbytes


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_host_pipeline_returns_every_batch_in_order(dev, depth):
    """A stream of different batches through the ring: each comes back in submission order and
    equals the oracle, with up to `depth` batches in flight on separate streams."""
    b, n, m, r, s = 3, 2048, 96, 0.15, 12
    batches = [W.DISTRIBUTIONS["UDS"[k % 3]](b, n, 300 + k) for k in range(7)]
    pipe = SetAbstractionPipeline(b, n, m, r, s, depth=depth)
    got = []
    for x in batches:
        if pipe.full():
            got.append(tuple(a.copy() for a in pipe.collect()))
        pipe.input_buffer()[...] = x
        pipe.submit()
    while pipe.pending():
        got.append(tuple(a.copy() for a in pipe.collect()))
    assert len(got) == len(batches)
    for x, (new_xyz, idx, cnt, grouped) in zip(batches, got):
        wn, _, widx, _ = oracle_sample_and_group(m, r, s, x, None)
        np.testing.assert_array_equal(new_xyz, wn)
        np.testing.assert_array_equal(idx, widx)
        np.testing.assert_array_equal(grouped, O.oracle_group_point(x, widx))
        assert (cnt == O.oracle_query_ball_point(r, s, x, wn)[1]).all()


def test_host_pipeline_refuses_overflow_and_empty_collect(dev):
    pipe = SetAbstractionPipeline(1, 256, 16, 0.2, 4, depth=2)
    with pytest.raises(RuntimeError):
        pipe.collect()
    x = W.cloud_uniform(1, 256, 9)
    pipe.submit(x)
    pipe.submit(x)
    with pytest.raises(RuntimeError):
        pipe.submit(x)
    with pytest.raises(RuntimeError):
        pipe.input_buffer()
    a = pipe.collect()
    b2 = pipe.collect()
    np.testing.assert_array_equal(a[1], b2[1])
    with pytest.raises(ValueError):
        pipe.submit(np.zeros((2, 256, 3), np.float32))
