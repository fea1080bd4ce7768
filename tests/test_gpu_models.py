"""GPU parity of the drop-in modules (through the module API -> C-ABI) against the same parameters executed by stock
PyTorch fp32 ops (cuDNN with TF32 disabled), at sizes up to the BASELINE config's 224x224.

Gates: eval-mode gaze map max-abs <= 1e-3 in `precise` mode (the BASELINE bar); train-mode forward:
BN running stats max-abs <= 1e-4; fast mode is reported, loosely gated.
"""
import os

import pytest
import torch

import torch_ref

pytestmark = pytest.mark.gpu


def _sp_inputs(B, S, dev, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x_s = torch.randn(B, 3, S, S, generator=g)
    x_t = (torch.randint(0, 256, (B, 20, S, S), generator=g).float() / 255 - 0.5) / 0.5
    return x_s.to(dev), x_t.to(dev)


def _make_sp(dev, seed=0):
    from utils import make_layers, cfg
    from models.model_SP import model_SP
    torch.manual_seed(seed)
    m = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20))
    torch_ref.randomize_(m, seed)
    return m.to(dev)


@pytest.mark.parametrize("B,S", [(2, 64), (1, 224), (2, 288)])
def test_trunk_eval(cuda_dev, B, S):
    m = _make_sp(cuda_dev).eval()
    x_s, x_t = _sp_inputs(B, S, cuda_dev)
    with torch.no_grad():
        got = m.features_s(x_s)
        ref = torch_ref.seq_forward(m.features_s, x_s)
        got_t = m.features_t(x_t)
        ref_t = torch_ref.seq_forward(m.features_t, x_t)
    assert got.shape == ref.shape == (B, 512, S // 16, S // 16)
    assert (got - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())
    assert (got_t - ref_t).abs().max().item() <= 1e-3 * max(1.0, ref_t.abs().max().item())


@pytest.mark.parametrize("B,S", [(2, 64), (2, 224)])
def test_model_sp_eval(cuda_dev, B, S):
    m = _make_sp(cuda_dev).eval()
    x_s, x_t = _sp_inputs(B, S, cuda_dev)
    seen = []
    h = m._modules.get('features_s').register_forward_hook(lambda mod, i, o: seen.append(o))  # AT.py:105 idiom
    with torch.no_grad():
        got = m(x_s, x_t)
        ref, f_s, _ = torch_ref.model_sp_forward(m, x_s, x_t, return_feats=True)
    h.remove()
    assert got.shape == (B, 1, S, S)
    err = (got - ref).abs().max().item()
    assert err <= 1e-3, "gaze map max-abs err %.3e" % err
    assert len(seen) == 1 and seen[0].shape == f_s.shape
    assert (seen[0] - f_s).abs().max().item() <= 1e-3 * max(1.0, f_s.abs().max().item())


def test_model_sp_eval_batch_invariance(cuda_dev):
    """Size-independent property at the BASELINE batch: eval-mode output of sample i does not depend on the batch."""
    m = _make_sp(cuda_dev).eval()
    x_s, x_t = _sp_inputs(32, 224, cuda_dev)
    with torch.no_grad():
        full = m(x_s, x_t)
        part = m(x_s[5:7].contiguous(), x_t[5:7].contiguous())
    assert torch.equal(full[5:7], part) or (full[5:7] - part).abs().max().item() <= 1e-6


def test_model_sp_train_forward(cuda_dev):
    """Train-mode BatchNorm (batch statistics + running-stat updates) forward, grads disabled."""
    import copy
    m = _make_sp(cuda_dev).train()
    m_ref = copy.deepcopy(m)
    x_s, x_t = _sp_inputs(4, 64, cuda_dev)
    with torch.no_grad():
        got = m(x_s, x_t)
        ref = torch_ref.model_sp_forward(m_ref, x_s, x_t)
    sd, sr = m.state_dict(), m_ref.state_dict()
    for k in sd:
        if "running_" in k:
            assert (sd[k] - sr[k]).abs().max().item() <= 1e-4, k
        if "num_batches_tracked" in k:
            assert int(sd[k]) == int(sr[k]) == 1, k
    err = (got - ref).abs().max().item()
    print("train-mode gaze map max-abs err vs stock fp32: %.3e" % err)
    assert err <= 1e-3, "train-mode gaze map max-abs err %.3e" % err   # north_star's bar (fp16-split forward: 22 operand bits)


def test_headline_shape_vs_stock_fp32(cuda_dev):
    """The benchmarked shape itself -- batch 32, 224x224, BASELINE configs[1] -- against the same parameters executed by stock
    PyTorch ops on the GPU (cuDNN, TF32 off): eval gaze map, train-mode gaze map + loss + every BatchNorm running statistic
    against stock fp32 (the arithmetic the reference executes; north_star's 1e-3 max-abs bar), and the parameter gradients of
    one training step against stock autograd in fp64, gated per tensor at 3x the distance stock fp32 autograd itself has from
    fp64 on that tensor (ReLU / max-pool routing flips make fp32 that noisy), floored at 1e-2."""
    import copy
    import numpy as np
    import floss as floss_mod
    from oracle import egaze_oracle as orc
    B, S = 32, 224
    m = _make_sp(cuda_dev, 0)
    with torch.no_grad():
        for mod in m.decoder:
            if isinstance(mod, torch.nn.Conv2d):
                mod.weight.mul_(0.8)   # keep the un-normalised decoder's logits O(1) (well-conditioned regime)
    x_s, x_t, gt = [torch.from_numpy(a).to(cuda_dev) for a in orc.synth_sp_inputs(B, S, 1234)]
    # eval
    m.eval()
    with torch.no_grad():
        got = m(x_s, x_t)
        ref = torch_ref.model_sp_forward(m, x_s, x_t)
    e_eval = (got - ref).abs().max().item()
    del got, ref
    # train step: egaze, stock fp32, stock fp64
    m.train()
    m_ref = copy.deepcopy(m)
    m64 = copy.deepcopy(m).double()
    out = m(x_s, x_t)
    loss = floss_mod.floss()(out, gt)
    loss.backward()
    out_r = torch_ref.model_sp_forward(m_ref, x_s, x_t)
    loss_r = torch_ref.floss_loss(out_r, gt)
    loss_r.backward()
    e_train = (out.detach() - out_r.detach()).abs().max().item()
    e_loss = abs(loss.item() - loss_r.item()) / abs(loss_r.item())
    sd, sr = m.state_dict(), m_ref.state_dict()
    e_stats = max((sd[k] - sr[k]).abs().max().item() for k in sd if "running_" in k)
    del out, out_r
    o64 = torch_ref.model_sp_forward(m64, x_s.double(), x_t.double())
    torch.nn.functional.binary_cross_entropy(o64, gt.double(), weight=torch_ref.floss_weight(gt).double()).backward()
    del o64
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
    rows = []
    for (k, p), (_, q), (_, r) in zip(m.named_parameters(), m_ref.named_parameters(), m64.named_parameters()):
        if r.grad.norm().item() < 1e-7:
            assert p.grad.double().norm().item() < 1e-4, k
            continue
        rows.append((k, rel(p.grad, r.grad), rel(q.grad, r.grad)))
    worst = max(rows, key=lambda r: r[1] / max(r[2], 1e-12))
    print("B=32x224: eval max-abs %.2e | train max-abs %.2e | loss rel %.2e | BN stats %.2e | grads vs fp64: egaze median %.2e "
          "max %.2e, stock fp32 median %.2e max %.2e; worst ratio %.2f (%s)"
          % (e_eval, e_train, e_loss, e_stats, np.median([r[1] for r in rows]), max(r[1] for r in rows),
             np.median([r[2] for r in rows]), max(r[2] for r in rows), worst[1] / max(worst[2], 1e-12), worst[0]))
    assert e_eval <= 1e-3 and e_train <= 1e-3, (e_eval, e_train)
    assert e_loss <= 1e-4 and e_stats <= 1e-4, (e_loss, e_stats)
    for k, e, n in rows:
        assert e <= max(1e-2, 3 * n), "%s: rel-L2 to fp64 %.3e, stock fp32 %.3e" % (k, e, n)


@pytest.mark.parametrize("train", [False, True])
def test_late_fusion(cuda_dev, train):
    import copy
    from models.late_fusion import late_fusion
    torch.manual_seed(0)
    m = torch_ref.randomize_(late_fusion(), 3).to(cuda_dev)
    m.train(train)
    m_ref = copy.deepcopy(m)
    f = torch.rand(4, 1, 224, 224, device=cuda_dev)
    g = torch.rand(4, 1, 224, 224, device=cuda_dev)
    with torch.no_grad():
        got = m(f, g)
        ref = torch_ref.late_fusion_forward(m_ref, f, g)
    assert got.shape == (4, 1, 224, 224)
    assert (got - ref).abs().max().item() <= 1e-3
    if train:
        sd, sr = m.state_dict(), m_ref.state_dict()
        for k in sd:
            if "running_" in k:
                assert (sd[k] - sr[k]).abs().max().item() <= 1e-4, k


def test_fast_mode_reported(cuda_dev):
    """`fast` (single bf16 pass) is a reported mode: check it runs and log its error; gate only loosely."""
    m = _make_sp(cuda_dev).eval()
    x_s, x_t = _sp_inputs(2, 224, cuda_dev)
    old = os.environ.get("EGAZE_PRECISION")
    os.environ["EGAZE_PRECISION"] = "fast"
    try:
        with torch.no_grad():
            got = m(x_s, x_t)
    finally:
        if old is None:
            del os.environ["EGAZE_PRECISION"]
        else:
            os.environ["EGAZE_PRECISION"] = old
    with torch.no_grad():
        ref = torch_ref.model_sp_forward(m, x_s, x_t)
    err = (got - ref).abs().max().item()
    print("fast-mode gaze map max-abs err: %.3e" % err)
    assert err <= 0.25


def test_at_sequence_vs_oracle_frame_loop(cuda_dev):
    """BASELINE configs[2] shape of work (AT over feature sequences): the batched path -- one crop-mean over all T*B frames,
    one LSTM call over the T steps, one weighted-map over all frames -- against the NumPy oracle walking the sequences frame
    by frame the way AT.extract_late does (AT.py:224-252)."""
    import numpy as np
    import models.LSTMnet as L
    from egaze import ops
    from oracle import egaze_oracle as orc
    T, NB = 6, 3
    torch.manual_seed(3)
    net = L.lstmnet().to(cuda_dev).eval()
    g = torch.Generator().manual_seed(9)
    feat = torch.relu(torch.randn(T * NB, 512, 14, 14, generator=g))
    gaze = torch.randint(0, 224, (T * NB, 2), generator=g).int()
    with torch.no_grad():
        vec = ops.crop_mean(feat.to(cuda_dev), gaze.to(cuda_dev), 3)
        hid = (torch.zeros(2, NB, 512, device=cuda_dev), torch.zeros(2, NB, 512, device=cuda_dev))
        w, _ = net(vec.view(T, NB, 512), hid)
        got = ops.weighted_map(w.reshape(T * NB, 512), feat.to(cuda_dev)).cpu().numpy()
    sd = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
    h = np.zeros((2, NB, 512), np.float32)
    c = np.zeros((2, NB, 512), np.float32)
    fn, gn = feat.numpy(), gaze.numpy()
    for t in range(T):
        f_t, g_t = fn[t * NB:(t + 1) * NB], gn[t * NB:(t + 1) * NB]
        v = orc.crop_mean(f_t, g_t, 3)
        out, h, c = orc.lstmnet_forward(sd, v[None], h, c)
        for b in range(NB):   # the reference normalises each frame's map on its own (batch 1)
            ref = orc.get_weighted(out[0, b:b + 1], f_t[b:b + 1])
            assert np.abs(got[t * NB + b] - ref[0]).max() <= 1e-4


def test_full_pipeline_train_step(cuda_dev):
    """BASELINE configs[3] per rank at a small size: SP train step, AT step on the hooked map, LF train step; the two
    losses must match stock PyTorch modules fed the same parameters and inputs."""
    import bench
    wl = bench.Workload("full_train", 2, 64, 0, 1, cuda_dev)
    import copy
    sp_ref, lf_ref = copy.deepcopy(wl.model), copy.deepcopy(wl.lf)
    lf_before = [p.detach().clone() for p in wl.lf.parameters()]
    losses = wl.step(*wl.dev).cpu()
    assert torch.isfinite(losses).all()
    x_s, x_t, gt = wl.dev
    ref_out = torch_ref.model_sp_forward(sp_ref, x_s, x_t)
    ref_loss = torch_ref.floss_loss(ref_out, gt)
    assert abs(losses[0].item() - ref_loss.item()) <= 1e-3 * abs(ref_loss.item())
    assert any((p.detach() - q).abs().max().item() > 0 for p, q in zip(wl.lf.parameters(), lf_before))
    losses2 = wl.step(*wl.dev).cpu()
    assert torch.isfinite(losses2).all()


def test_at_step_fixation_saccade_vs_oracle(cuda_dev):
    """egaze.at.at_sequence (B videos advancing together, fixation frames keep their crop weights and LSTM state) against the
    NumPy oracle running every video on its own through the reference's loop body (AT.py:236-248)."""
    import numpy as np
    import models.LSTMnet as L
    from egaze import at
    from oracle import egaze_oracle as orc
    T, B = 7, 3
    torch.manual_seed(4)
    net = L.lstmnet().to(cuda_dev).eval()
    g = torch.Generator().manual_seed(10)
    feats = torch.relu(torch.randn(T, B, 512, 14, 14, generator=g))
    gazes = torch.randint(0, 224, (T, B, 2), generator=g).int()
    fixsac = (torch.rand(T, B, generator=g) < 0.4).int()
    maps, _ = at.at_sequence(feats.to(cuda_dev), gazes.to(cuda_dev), fixsac.to(cuda_dev), net)
    maps = maps.cpu().numpy()
    sd = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
    fn, gn, fs = feats.numpy(), gazes.numpy(), fixsac.numpy()
    for b in range(B):
        h = np.zeros((2, 1, 512), np.float32)
        c = np.zeros((2, 1, 512), np.float32)
        for t in range(T):
            f = fn[t, b:b + 1]
            w = orc.crop_mean(f, gn[t, b:b + 1], 3)
            if fs[t, b] != 1:
                out, h, c = orc.lstmnet_forward(sd, w[None], h, c)
                w = out[0]
            ref = orc.get_weighted(w, f)[0]
            assert np.abs(maps[t, b] - ref).max() <= 1e-4, (t, b)
