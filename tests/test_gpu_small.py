"""GPU parity of the HBM-/latency-bound kernels (layout, head, floss, AT glue, LSTM) vs plain PyTorch fp32."""
import pytest
import torch
import torch.nn.functional as F

import torch_ref

pytestmark = pytest.mark.gpu


def test_layout_roundtrip(cuda_dev):
    from egaze import ops
    for shape in [(2, 3, 20, 24), (1, 20, 17, 9), (2, 64, 14, 14), (1, 130, 5, 7)]:
        x = torch.randn(*shape, device=cuda_dev)
        act = ops.to_split(x)
        assert act.Cp % 16 == 0 and act.Cp >= shape[1]
        y = ops.from_split(act)
        assert (x - y).abs().max().item() <= 2.0 ** -16 * x.abs().max().item()
        for fmt, bits in ((0, 16), (1, 21)):   # bf16 split: 16 significand bits; fp16 split: 22 (21 gated)
            act_f = ops.to_split(x, fmt=fmt, xb=bool(fmt))
            assert act_f.fmt == fmt
            assert (x - ops.from_split(act_f)).abs().max().item() <= 2.0 ** -bits * x.abs().max().item()
            if fmt:
                assert torch.equal(act_f.xb[..., :shape[1]].float(), x.permute(0, 2, 3, 1).bfloat16().float())
        if act.Cp > shape[1]:
            assert act.hi[..., shape[1]:].float().abs().max().item() == 0.0
        z = ops.nhwc_f32_to_nchw(ops.nchw_to_nhwc_f32(x))
        assert torch.equal(x, z)
        assert torch.equal(ops.nchw_to_nhwc_f32(x), x.permute(0, 2, 3, 1).contiguous())


def test_pack_weights(cuda_dev):
    from egaze import ops
    w = torch.randn(24, 10, 3, 3, device=cuda_dev)
    hi, lo, rows, cp, _ = ops.pack_cache.get(w, 0, fmt=0)
    got = (hi.float() + lo.float())  # [9][24][16]
    ref = w.permute(2, 3, 0, 1).reshape(9, 24, 10)
    assert cp == 16 and rows == 24
    assert (got[:, :, :10] - ref).abs().max().item() <= 2.0 ** -16 * ref.abs().max().item()
    assert got[:, :, 10:].abs().max().item() == 0
    hi, lo, rows, cp, _ = ops.pack_cache.get(w, 1)
    got = (hi.float() + lo.float())  # [9][10][32]
    ref = torch.flip(w, (2, 3)).permute(2, 3, 1, 0).reshape(9, 10, 24)
    assert (got[:, :, :24] - ref).abs().max().item() <= 2.0 ** -16 * ref.abs().max().item()
    # fp16 forward operand: planes hold w * scale, 22 significand bits
    hi, lo, rows, cp, fmt = ops.pack_cache.get(w, 0, fmt=1)
    assert fmt == 1 and hi.dtype == torch.float16
    got = (hi.float() + lo.float()) / ops.f16_weight_scale()
    ref = w.permute(2, 3, 0, 1).reshape(9, 24, 10)
    assert (got[:, :, :10] - ref).abs().max().item() <= 2.0 ** -21 * ref.abs().max().item()


@pytest.mark.parametrize("C,Cs", [(64, 64), (8, 16)])
def test_head_fwd(cuda_dev, C, Cs):
    from egaze import ops
    x = torch.randn(2, C, 40, 24, device=cuda_dev)
    w = torch.randn(1, C, 1, 1, device=cuda_dev) * 0.3
    b = torch.randn(1, device=cuda_dev)
    act = ops.to_split(x, Cs)
    y, logit = ops.head_fwd(act, w, b, want_logit=True)
    ref_logit = F.conv2d(ops.from_split(act), w, b)
    assert (logit - ref_logit).abs().max().item() <= 1e-5 * max(1.0, ref_logit.abs().max().item())
    assert (y - torch.sigmoid(ref_logit)).abs().max().item() <= 1e-6


def _blob_targets(B, S, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.arange(S).float(), torch.arange(S).float(), indexing="ij")
    out = []
    for _ in range(B):
        cy, cx = (torch.rand(2, generator=g) * 0.7 + 0.15) * S
        gmap = torch.exp(-((ys - cy) ** 2 / (2 * (S * 16.3 / 224) ** 2) + (xs - cx) ** 2 / (2 * (S * 12.25 / 224) ** 2)))
        gmap = (gmap - gmap.min()) / (gmap.max() - gmap.min())
        out.append(torch.round(gmap * 255) / 255)
    return torch.stack(out).unsqueeze(1).to(dev)


def test_floss_kats(cuda_dev):
    """SURVEY 4.3 / App. D known answers: single peak -> weight W at the peak; 2-pixel vertical plateau -> 149.33."""
    import floss as floss_mod
    fl = floss_mod.floss()
    t = torch.zeros(1, 1, 224, 224, device=cuda_dev)
    t[0, 0, 100, 50] = 1.0
    w = fl.build_weight_from_target(t)
    assert w.shape == (1, 1, 224, 224) and abs(float(w[0, 0, 100, 50]) - 224.0) < 1e-4
    t[0, 0, 101, 50] = 1.0
    w = fl.build_weight_from_target(t)
    assert abs(float(w[0, 0, 100, 50]) - 149.3333) < 1e-3 and abs(float(w[0, 0, 101, 50]) - 149.3333) < 1e-3
    # saturated predictions hit the -100 clamp: p=0,t=1 -> 100*w
    p = torch.full((1, 1, 224, 224), 0.5, device=cuda_dev)
    tt = torch.zeros(1, 1, 224, 224, device=cuda_dev)
    p[0, 0, 3, 4] = 0.0
    tt[0, 0, 3, 4] = 1.0
    ref = torch_ref.floss_loss(p, tt)
    got = fl(p, tt)
    assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item())


@pytest.mark.parametrize("B,S", [(4, 224), (3, 64)])
def test_floss_fwd_bwd(cuda_dev, B, S):
    import floss as floss_mod
    fl = floss_mod.floss()
    t = _blob_targets(B, S, cuda_dev)
    p = torch.sigmoid(torch.randn(B, 1, S, S, device=cuda_dev) * 3).requires_grad_(True)
    w = torch.from_numpy(fl.build_weight_from_target(t)).to(cuda_dev)
    assert (w - torch_ref.floss_weight(t)).abs().max().item() <= 1e-4
    loss = fl(p, t)
    loss.backward()
    p2 = p.detach().clone().requires_grad_(True)
    ref = torch_ref.floss_loss(p2, t)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert (p.grad - p2.grad).abs().max().item() <= 1e-5 * p2.grad.abs().max().item()


def test_at_glue(cuda_dev):
    from egaze import ops
    B = 5
    feat = F.relu(torch.randn(B, 512, 14, 14, device=cuda_dev))
    gaze = [[0, 0], [223, 223], [100, 37], [15, 208], [120, 120]]
    cm = ops.crop_mean(feat, gaze)
    ref = torch_ref.crop_mean(feat, gaze)
    assert (cm - ref).abs().max().item() <= 1e-6 * max(1.0, ref.abs().max().item())
    wm = ops.weighted_map(cm, feat)
    refm = torch_ref.get_weighted(cm, feat)
    assert (wm - refm).abs().max().item() <= 2e-5
    up = ops.bilinear_up(wm, 16, False)
    refu = F.interpolate(wm.unsqueeze(1), scale_factor=16, mode="bilinear", align_corners=False).squeeze(1)
    assert (up - refu).abs().max().item() <= 1e-5
    up = ops.bilinear_up(wm, 16, True)
    refu = F.interpolate(wm.unsqueeze(1), scale_factor=16, mode="bilinear", align_corners=True).squeeze(1)
    assert (up - refu).abs().max().item() <= 1e-5


@pytest.mark.parametrize("T,B", [(1, 1), (5, 3), (30, 16)])
def test_lstmnet_fwd(cuda_dev, T, B):
    import models.LSTMnet as L
    torch.manual_seed(0)
    net = L.lstmnet().to(cuda_dev).eval()
    x = torch.randn(T, B, 512, device=cuda_dev)
    h0 = torch.randn(2, B, 512, device=cuda_dev) * 0.3
    c0 = torch.randn(2, B, 512, device=cuda_dev) * 0.3
    with torch.no_grad():
        out, (hn, cn) = net(x, (h0, c0))
        ref, (rh, rc) = torch_ref.lstmnet_forward(net, x, h0, c0)
    assert (out - ref).abs().max().item() <= 2e-5
    assert (hn - rh).abs().max().item() <= 2e-5 and (cn - rc).abs().max().item() <= 2e-5
    # stepwise == full sequence
    with torch.no_grad():
        hid = (h0, c0)
        outs = []
        for t in range(T):
            o, hid = net(x[t:t + 1], hid)
            outs.append(o)
    assert (torch.cat(outs) - out).abs().max().item() <= 1e-6


def test_lstmnet_hidden_none(cuda_dev):
    """hidden=None uses the module-global batch_size (=1): works at batch 1, raises at batch 16."""
    import models.LSTMnet as L
    net = L.lstmnet().to(cuda_dev).eval()
    with torch.no_grad():
        out, _ = net(torch.randn(1, 1, 512, device=cuda_dev), None)
        assert out.shape == (1, 1, 512)
        with pytest.raises(RuntimeError):
            net(torch.randn(30, 16, 512, device=cuda_dev), None)
        L.batch_size = 16
        try:
            out, _ = net(torch.randn(30, 16, 512, device=cuda_dev), None)
            assert out.shape == (30, 16, 512)
        finally:
            L.batch_size = 1


@pytest.mark.parametrize("T,B", [(1, 1), (4, 3), (30, 16)])
def test_lstmnet_bwd(cuda_dev, T, B):
    """BPTT through the egaze LSTM kernels vs PyTorch autograd of nn.LSTM/nn.Linear over the same parameters
    (AT.trainLSTM's loss: MSE against tanh(target), AT.py:138)."""
    import copy
    import models.LSTMnet as L
    torch.manual_seed(1)
    net = L.lstmnet().to(cuda_dev).train()
    ref = copy.deepcopy(net)
    x = torch.randn(T, B, 512, device=cuda_dev)
    h0 = (torch.randn(2, B, 512, device=cuda_dev) * 0.3).requires_grad_(True)
    c0 = (torch.randn(2, B, 512, device=cuda_dev) * 0.3).requires_grad_(True)
    tgt = torch.tanh(torch.randn(T, B, 512, device=cuda_dev))
    xa = x.clone().requires_grad_(True)
    out, (hn, cn) = net(xa, (h0, c0))
    loss = torch.nn.functional.mse_loss(out, tgt) + 0.1 * hn.sum() + 0.05 * cn.pow(2).sum()
    loss.backward()
    xb = x.clone().requires_grad_(True)
    h0r, c0r = h0.detach().clone().requires_grad_(True), c0.detach().clone().requires_grad_(True)
    outr, (hnr, cnr) = torch_ref.lstmnet_forward(ref, xb, h0r, c0r)
    lossr = torch.nn.functional.mse_loss(outr, tgt) + 0.1 * hnr.sum() + 0.05 * cnr.pow(2).sum()
    lossr.backward()
    assert abs(loss.item() - lossr.item()) <= 1e-5 * abs(lossr.item())
    rel = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-20)).item()
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert rel(p.grad, q.grad) <= 1e-4, "%s: %.3e" % (k, rel(p.grad, q.grad))
    assert rel(xa.grad, xb.grad) <= 1e-4
    assert rel(h0.grad, h0r.grad) <= 1e-4 and rel(c0.grad, c0r.grad) <= 1e-4
