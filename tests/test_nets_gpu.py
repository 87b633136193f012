"""GPU tests of the networks built from the CUDA geometry ops + torch tails."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from pointnet2_b200 import nets, workloads as W

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _demo():
    spec = importlib.util.spec_from_file_location("train_ddp_demo", os.path.join(ROOT, "tools", "train_ddp_demo.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name,b,n,out", [("cls_ssg", 4, 1024, (4, 40)), ("cls_msg", 3, 1024, (3, 40)),
                                          ("sem_seg", 2, 2048, (2, 2048, 21))])
def test_network_forward_backward(dev, name, b, n, out):
    torch.manual_seed(0)
    net = {"cls_ssg": nets.PointNet2ClsSSG, "cls_msg": nets.PointNet2ClsMSG, "sem_seg": nets.PointNet2SemSeg}[name]().to(dev)
    xyz = torch.from_numpy(W.cloud_surface(b, n, 400)).to(dev)
    pred, _ = net(xyz)
    assert tuple(pred.shape) == out and bool(torch.isfinite(pred).all())
    pred.square().mean().backward()
    missing = [k for k, p in net.named_parameters() if p.grad is None or not bool(torch.isfinite(p.grad).all())]
    assert not missing, missing
    # gradients reach the first layer through group_point / three_interpolate backward kernels
    first = next(net.parameters())
    assert float(first.grad.abs().sum()) > 0


def test_eval_mode_is_deterministic(dev):
    torch.manual_seed(0)
    net = nets.PointNet2ClsSSG(8).to(dev).eval()
    xyz = torch.from_numpy(W.cloud_uniform(2, 1024, 401)).to(dev)
    with torch.no_grad():
        a, _ = net(xyz)
        b2, _ = net(xyz)
    assert torch.equal(a, b2)


def test_training_reduces_the_loss_on_synthetic_shapes(dev):
    demo = _demo()
    torch.manual_seed(0)
    net = nets.PointNet2ClsSSG(3).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    rs = np.random.RandomState(7)
    losses = []
    for _ in range(24):
        xyz, lab = demo.synthetic_shapes(16, 512, 3, rs)
        pred, _ = net(torch.from_numpy(xyz).to(dev))
        loss = nets.cls_loss(pred, torch.from_numpy(lab).to(dev))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert np.mean(losses[-6:]) < 0.7 * np.mean(losses[:6]), losses
