"""GPU: the dedicated late-fusion kernels (egaze_lf_fwd / egaze_lf_bwd, csrc/lf.cu) against stock torch.nn autograd over the
same parameter containers (reference models/late_fusion.py:10-23), at sizes that exercise ragged tiles (H, W not multiples
of the 16x16 tile), batch 1, input gradients, and the full 224x224 shape of LF.trainLate (LF.py:79-105)."""
import copy

import pytest
import torch

import torch_ref

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _pair(dev, seed=3):
    from models.late_fusion import late_fusion
    torch.manual_seed(0)
    m = torch_ref.randomize_(late_fusion(), seed).to(dev)
    return m, copy.deepcopy(m).double()


@pytest.mark.parametrize("shape", [(1, 16, 16), (3, 40, 56), (2, 224, 224), (2, 23, 37)])
@pytest.mark.parametrize("train", [False, True])
def test_lf_forward_shapes(cuda_dev, shape, train):
    m, m64 = _pair(cuda_dev)
    m.train(train)
    m64.train(train)
    B, H, W = shape
    g = torch.Generator(device="cpu").manual_seed(11)
    f = torch.rand(B, 1, H, W, generator=g).to(cuda_dev)
    a = torch.rand(B, 1, H, W, generator=g).to(cuda_dev)
    with torch.no_grad():
        got = m(f, a)
        ref = torch_ref.late_fusion_forward(m64, f.double(), a.double())
    assert got.shape == (B, 1, H, W)
    err = (got.double() - ref).abs().max().item()
    print('lf fwd %s train=%s max-abs err vs fp64 %.2e' % (shape, train, err))
    assert err <= 5e-4, err
    if train:
        sd, sr = m.state_dict(), m64.state_dict()
        for k in sd:
            if "running_" in k:
                assert (sd[k].double() - sr[k]).abs().max().item() <= 1e-5, k
            if "num_batches_tracked" in k:
                assert int(sd[k]) == int(sr[k]) == 1


@pytest.mark.parametrize("shape", [(2, 32, 48), (3, 40, 56), (2, 224, 224)])
def test_lf_backward_vs_fp64_autograd(cuda_dev, shape):
    """Every parameter gradient and both input gradients of a train-mode step against stock autograd in fp64."""
    import floss as floss_mod
    m, m64 = _pair(cuda_dev, 5)
    m.train()
    m64.train()
    B, H, W = shape
    g = torch.Generator(device="cpu").manual_seed(12)
    f = torch.rand(B, 1, H, W, generator=g).to(cuda_dev).requires_grad_(True)
    a = torch.rand(B, 1, H, W, generator=g).to(cuda_dev).requires_grad_(True)
    gt = torch.rand(B, 1, H, W, generator=g).to(cuda_dev)
    if H == W:
        loss = floss_mod.floss()(m(f, a), gt)
    else:   # floss assumes square maps (floss.py:35): plain BCE keeps the ragged shapes in the test
        loss = torch.nn.functional.binary_cross_entropy(m(f, a), gt)
    loss.backward()
    f64 = f.detach().double().requires_grad_(True)
    a64 = a.detach().double().requires_grad_(True)
    out64 = torch_ref.late_fusion_forward(m64, f64, a64)
    if H == W:
        loss64 = torch.nn.functional.binary_cross_entropy(out64, gt.double(), weight=torch_ref.floss_weight(gt).double())
    else:
        loss64 = torch.nn.functional.binary_cross_entropy(out64, gt.double())
    loss64.backward()
    assert abs(loss.item() - loss64.item()) <= 1e-4 * abs(loss64.item())
    for (k, p), (_, q) in zip(m.named_parameters(), m64.named_parameters()):
        if k in ("fusion.0.bias", "fusion.3.bias", "fusion.6.bias"):
            # bias in front of a batch-statistics BatchNorm: exactly zero here, rounding noise in stock autograd
            assert p.grad.abs().max().item() == 0.0 and q.grad.abs().max().item() <= 1e-8 * max(1.0, loss64.item()), k
            continue
        print("lf grad %s %s rel-L2 vs fp64 %.2e" % (shape, k, rel_l2(p.grad, q.grad)))
        # ReLU routing flips (a forward value within rounding distance of zero) move single gradients by ~1e-3..1e-2
        assert rel_l2(p.grad, q.grad) <= 2e-2, "%s: %.3e" % (k, rel_l2(p.grad, q.grad))
    print("lf input grads %s rel-L2 vs fp64 %.2e %.2e" % (shape, rel_l2(f.grad, f64.grad), rel_l2(a.grad, a64.grad)))
    assert rel_l2(f.grad, f64.grad) <= 4e-2, rel_l2(f.grad, f64.grad)
    assert rel_l2(a.grad, a64.grad) <= 4e-2, rel_l2(a.grad, a64.grad)


def test_lf_frozen_weights_and_determinism(cuda_dev):
    """requires_grad=False on a conv weight skips its weight-gradient kernel; two identical steps give bit-identical grads
    (per-CTA partials are reduced in a fixed order)."""
    m, _ = _pair(cuda_dev, 7)
    m.train()
    m.fusion[3].weight.requires_grad_(False)
    f = torch.rand(2, 1, 64, 64, device=cuda_dev)
    a = torch.rand(2, 1, 64, 64, device=cuda_dev)
    grads = []
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        m(f, a).sum().backward()
        assert m.fusion[3].weight.grad is None
        grads.append([p.grad.clone() for p in m.parameters() if p.grad is not None])
    for x, y in zip(*grads):
        assert torch.equal(x, y)


def test_lf_eval_mode_backward(cuda_dev):
    """BatchNorm on running statistics (model.eval()) with autograd on: stock nn.BatchNorm2d supports it, so do we."""
    m, m64 = _pair(cuda_dev, 9)
    m.eval()
    m64.eval()
    g = torch.Generator(device="cpu").manual_seed(13)
    f = torch.rand(2, 1, 48, 32, generator=g).to(cuda_dev).requires_grad_(True)
    a = torch.rand(2, 1, 48, 32, generator=g).to(cuda_dev)
    wgt = torch.rand(2, 1, 48, 32, generator=g).to(cuda_dev)
    (m(f, a) * wgt).sum().backward()
    f64 = f.detach().double().requires_grad_(True)
    (torch_ref.late_fusion_forward(m64, f64, a.double()) * wgt.double()).sum().backward()
    for (k, p), (_, q) in zip(m.named_parameters(), m64.named_parameters()):
        assert rel_l2(p.grad, q.grad) <= 5e-3, "%s: %.3e" % (k, rel_l2(p.grad, q.grad))
    assert rel_l2(f.grad, f64.grad) <= 5e-3
