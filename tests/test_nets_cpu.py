"""CPU tests of the learned tails (pointnet2_b200/nets.py): layer algebra and
architecture bookkeeping only — the geometry ops need a GPU (tests/test_nets_gpu.py)."""
import numpy as np
import pytest
import torch
from torch import nn

from pointnet2_b200 import nets


def test_shared_mlp_is_a_1x1_convolution_with_batch_norm():
    """tf_util.conv2d(kernel [1,1]) + batch norm + relu on NHWC == Linear + BatchNorm1d + ReLU on the last axis."""
    torch.manual_seed(0)
    mlp = nets.SharedMLP(7, [5, 9])
    x = torch.randn(2, 6, 4, 7)
    got = mlp(x)  # training mode: batch statistics
    t = x.permute(0, 3, 1, 2)  # NCHW
    lins = [m for m in mlp.body if isinstance(m, nn.Linear)]
    bns = [m for m in mlp.body if isinstance(m, nn.BatchNorm1d)]
    for lin, bn in zip(lins, bns):
        t = nn.functional.conv2d(t, lin.weight[:, :, None, None], lin.bias)
        t = nn.functional.batch_norm(t, None, None, bn.weight, bn.bias, training=True, eps=bn.eps)
        t = torch.relu(t)
    np.testing.assert_allclose(got.detach().numpy(), t.permute(0, 2, 3, 1).detach().numpy(), atol=2e-5, rtol=1e-5)
    assert got.shape == (2, 6, 4, 9)


def test_shared_mlp_logit_layer_has_no_bn_or_relu():
    m = nets.SharedMLP(4, [3], bn=False, last_activation=False)
    assert [type(x) for x in m.body] == [nn.Linear]
    assert float(m.body[0].bias.abs().sum()) == 0.0  # zero-initialised bias, as tf_util does


@pytest.mark.parametrize("ctor,count", [(lambda: nets.PointNet2ClsSSG(40), 1475688), (lambda: nets.PointNet2ClsMSG(40), 1747368),
                                        (lambda: nets.PointNet2SemSeg(13), 967981)])
def test_network_parameter_counts_pin_the_layer_widths(ctor, count):
    """Widths quoted from models/pointnet2_cls_ssg.py:32-43, pointnet2_cls_msg.py:27-38,
    pointnet2_sem_seg.py:28-46 (1.48 M / 1.74 M parameters are the published sizes)."""
    assert sum(p.numel() for p in ctor().parameters()) == count


def test_bn_decay_maps_to_torch_momentum():
    m = nets.PointNet2ClsSSG(4)
    nets.set_bn_momentum(m, 0.9)
    assert all(abs(x.momentum - 0.1) < 1e-12 for x in m.modules() if isinstance(x, nn.BatchNorm1d))


def test_sem_seg_loss_is_weighted_mean_over_nonzero_weights():
    torch.manual_seed(1)
    pred = torch.randn(2, 5, 3)
    lab = torch.randint(0, 3, (2, 5))
    w = torch.tensor([[1.0, 0.0, 2.0, 0.5, 0.0], [0.0, 0.0, 1.0, 1.0, 3.0]])
    per = -torch.log_softmax(pred, -1).gather(-1, lab[..., None]).squeeze(-1)
    want = (per * w).sum() / 6
    assert abs(float(nets.sem_seg_loss(pred, lab, w)) - float(want)) < 1e-6
    assert abs(float(nets.cls_loss(pred[:, 0], lab[:, 0])) - float(per[:, 0].mean())) < 1e-6
