"""The reference's own call sites run unchanged (VERDICT r1 a7 caveat): models/pointnet2_sem_seg.py:28-37 and
models/pointnet2_cls_msg.py:27-29 are replayed verbatim — positional order, mlp width lists, is_training,
bn_decay, scope — against pointnet2_b200.pointnet_util."""
import inspect

import pytest
import torch

from pointnet2_b200 import layers, workloads as W
from pointnet2_b200.pointnet_util import pointnet_fp_module, pointnet_sa_module, pointnet_sa_module_msg

pytestmark = pytest.mark.gpu


def sem_seg_trunk(point_cloud, is_training, bn_decay=None):
    l0_xyz = point_cloud
    l0_points = None
    # Layer 1 ... (models/pointnet2_sem_seg.py:28-31, verbatim)
    l1_xyz, l1_points, l1_indices = pointnet_sa_module(l0_xyz, l0_points, npoint=1024, radius=0.1, nsample=32, mlp=[32,32,64], mlp2=None, group_all=False, is_training=is_training, bn_decay=bn_decay, scope='layer1')
    l2_xyz, l2_points, l2_indices = pointnet_sa_module(l1_xyz, l1_points, npoint=256, radius=0.2, nsample=32, mlp=[64,64,128], mlp2=None, group_all=False, is_training=is_training, bn_decay=bn_decay, scope='layer2')
    l3_xyz, l3_points, l3_indices = pointnet_sa_module(l2_xyz, l2_points, npoint=64, radius=0.4, nsample=32, mlp=[128,128,256], mlp2=None, group_all=False, is_training=is_training, bn_decay=bn_decay, scope='layer3')
    l4_xyz, l4_points, l4_indices = pointnet_sa_module(l3_xyz, l3_points, npoint=16, radius=0.8, nsample=32, mlp=[256,256,512], mlp2=None, group_all=False, is_training=is_training, bn_decay=bn_decay, scope='layer4')
    # Feature Propagation layers (:34-37, verbatim)
    l3_points = pointnet_fp_module(l3_xyz, l4_xyz, l3_points, l4_points, [256,256], is_training, bn_decay, scope='fa_layer1')
    l2_points = pointnet_fp_module(l2_xyz, l3_xyz, l2_points, l3_points, [256,256], is_training, bn_decay, scope='fa_layer2')
    l1_points = pointnet_fp_module(l1_xyz, l2_xyz, l1_points, l2_points, [256,128], is_training, bn_decay, scope='fa_layer3')
    l0_points = pointnet_fp_module(l0_xyz, l1_xyz, l0_points, l1_points, [128,128,128], is_training, bn_decay, scope='fa_layer4')
    return l0_points, (l1_indices, l2_indices, l3_indices, l4_indices)


def test_sem_seg_call_sites_run_unchanged(dev):
    layers.reset_scopes()
    x = torch.from_numpy(W.cloud_duplicates(2, 8192, 7)).to(dev)
    feats, idxs = sem_seg_trunk(x, is_training=True, bn_decay=0.5)
    assert tuple(feats.shape) == (2, 8192, 128)
    assert [tuple(i.shape) for i in idxs] == [(2, 1024, 32), (2, 256, 32), (2, 64, 32), (2, 16, 32)]
    assert set(layers.scope_modules()) == {f"layer{k}/conv" for k in (1, 2, 3, 4)} | {f"fa_layer{k}/conv" for k in (1, 2, 3, 4)}
    n_params = sum(p.numel() for p in layers.scope_parameters())
    assert n_params > 0
    feats.sum().backward()  # gradients reach the scoped layers
    assert all(p.grad is not None for p in layers.scope_parameters())
    # bn_decay = 0.5 is the weight of the OLD average -> torch momentum 0.5
    bns = [m for mod in layers.scope_modules().values() for m in mod.modules() if isinstance(m, torch.nn.BatchNorm1d)]
    assert bns and all(abs(m.momentum - 0.5) < 1e-12 for m in bns)
    # a second pass reuses the variables (no new scopes), and in inference mode it is deterministic
    a, _ = sem_seg_trunk(x, is_training=False)
    b, _ = sem_seg_trunk(x, is_training=False)
    assert sum(p.numel() for p in layers.scope_parameters()) == n_params
    assert torch.equal(a, b)
    with pytest.raises(ValueError):  # same scope, other widths: TensorFlow raises too
        pointnet_sa_module(x, None, 64, 0.2, 8, [16, 16], None, False, False, None, 'layer1')
    layers.reset_scopes()


def test_cls_msg_call_site_runs_unchanged(dev):
    layers.reset_scopes()
    l0_xyz = torch.from_numpy(W.cloud_surface(4, 1024, 8)).to(dev)
    l0_points = None
    is_training, bn_decay = False, None
    # models/pointnet2_cls_msg.py:27-28, verbatim
    l1_xyz, l1_points = pointnet_sa_module_msg(l0_xyz, l0_points, 512, [0.1,0.2,0.4], [16,32,128], [[32,32,64], [64,64,128], [64,96,128]], is_training, bn_decay, scope='layer1', use_nchw=True)
    l2_xyz, l2_points = pointnet_sa_module_msg(l1_xyz, l1_points, 128, [0.2,0.4,0.8], [32,64,128], [[64,64,128], [128,128,256], [128,128,256]], is_training, bn_decay, scope='layer2')
    l3_xyz, l3_points, _ = pointnet_sa_module(l2_xyz, l2_points, npoint=None, radius=None, nsample=None, mlp=[256,512,1024], mlp2=None, group_all=True, is_training=is_training, bn_decay=bn_decay, scope='layer3')
    assert tuple(l1_points.shape) == (4, 512, 320) and tuple(l2_points.shape) == (4, 128, 640) and tuple(l3_points.shape) == (4, 1, 1024)
    layers.reset_scopes()


def test_signatures_equal_the_references():
    """Positional parameter names, in order, as in utils/pointnet_util.py:87,156,199 (plus our trailing `fused`)."""
    assert list(inspect.signature(pointnet_sa_module).parameters)[:16] == [
        "xyz", "points", "npoint", "radius", "nsample", "mlp", "mlp2", "group_all", "is_training", "bn_decay", "scope", "bn", "pooling",
        "knn", "use_xyz", "use_nchw"]
    assert list(inspect.signature(pointnet_sa_module_msg).parameters)[:12] == [
        "xyz", "points", "npoint", "radius_list", "nsample_list", "mlp_list", "is_training", "bn_decay", "scope", "bn", "use_xyz", "use_nchw"]
    assert list(inspect.signature(pointnet_fp_module).parameters)[:9] == [
        "xyz1", "xyz2", "points1", "points2", "mlp", "is_training", "bn_decay", "scope", "bn"]
