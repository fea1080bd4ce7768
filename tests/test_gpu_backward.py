"""GPU parity of the backward path (dgrad / wgrad on tcgen05, BatchNorm / pool / pair-max / head / floss backward)
against PyTorch autograd over the same parameters with stock fp32 ops, and against the reference's golden gradients.

Gates: kernel-level (wgrad / dgrad / BN backward on identical operands) rel-L2 <= 2e-4 (measured ~5e-6, stock fp32 ~3e-6).
Whole-network gradients are compared with an fp64 run of the same stock modules AND with the stock fp32 run: the
network's discrete routing (ReLU masks, max-pool / pair-max arg-max) flips wherever a forward value sits within
rounding distance of a tie, so even stock fp32 is 5e-3..1.3e-2 (rel-L2) from fp64 at the trunk.  The split-bf16
forward carries 16 instead of 24 significand bits (forward rel-err 9e-6 vs 3e-6 per conv), flips ~5x more often and
lands at ~3e-2, worst single tensor 5.04e-2 (measured, tools/precision_probe.py; which tensor is worst moves with the
summation order inside the kernels).  Gate: per-tensor rel-L2 <= 6e-2 and cosine >= 0.998, median
rel-L2 <= max(1e-2, 8 x median stock-fp32 noise), loss rel-err <= 1e-3."""
import copy
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import torch_ref
from oracle import egaze_oracle as orc
from test_oracle_golden import GOLD, sp_shapes, lf_shapes

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return ((a - b).double().norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 64), (2, 28, 28, 128, 256), (1, 14, 14, 512, 512), (1, 56, 56, 64, 128),
                                   (2, 32, 32, 3, 64), (2, 36, 36, 64, 64), (1, 224, 224, 64, 64)])
def test_wgrad(cuda_dev, shape, numeric_mode):
    """The weight-gradient GEMM against conv2d_weight of exactly the planes it multiplies (ops.wgrad_operands: bf16(X) and
    dY_hi in the default mode, the hi+lo planes of both in `precise3`)."""
    from egaze import ops
    N, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, Cin, H, W, generator=g).to(cuda_dev)
    dy = torch.randn(N, Cout, H, W, generator=g).to(cuda_dev)
    xa, dya = ops.to_split(x, xb=True), ops.grad_split(dy)
    gw = ops.wgrad3x3(xa, dya, Cout, Cin)
    x_hi, x_lo, dy_hi, dy_lo, _ = ops.wgrad_operands(xa, dya)
    xr, dyr = ops.from_split(ops.Act(x_hi, x_lo, Cin)), ops.from_split(ops.Act(dy_hi, dy_lo, Cout))
    assert (xr - x).abs().max().item() <= 2.0 ** -8 * x.abs().max().item()
    ref = torch.nn.grad.conv2d_weight(xr.double(), (Cout, Cin, 3, 3), dyr.double(), padding=1).float()
    assert rel_l2(gw, ref) <= 2e-4, rel_l2(gw, ref)
    assert (gw - ref).abs().max().item() <= 3e-4 * ref.abs().max().item()


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 64), (2, 28, 28, 128, 256), (1, 56, 56, 256, 128)])
def test_dgrad(cuda_dev, shape, numeric_mode):
    from egaze import ops
    N, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(4)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.05).to(cuda_dev)
    dy = torch.randn(N, Cout, H, W, generator=g).to(cuda_dev)
    dya = ops.grad_split(dy)
    wp = ops.pack_cache.get(w, 1, cols_p=dya.Cp)
    _, dx, _ = ops.conv3x3(dya, wp, want_f32=True, want_split=False)
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), w.double(), ops.from_split(dya).double(), padding=1).float()
    got = ops.nhwc_f32_to_nchw(dx, Cin)
    assert rel_l2(got, ref) <= 2e-4


@pytest.mark.parametrize("pool", [False, True])
def test_bn_bwd(cuda_dev, pool):
    from egaze import ops
    N, C, H, W = 3, 64, 12, 16
    g = torch.Generator().manual_seed(6)
    raw = torch.randn(N, C, H, W, generator=g).to(cuda_dev).requires_grad_(True)
    bn = torch.nn.BatchNorm2d(C).to(cuda_dev).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.3)
    y = F.relu(bn(raw))
    if pool:
        y = F.max_pool2d(y, 2, 2)
    gy = torch.randn(y.shape, generator=g).to(cuda_dev)
    y.backward(gy)
    raw_nhwc = ops.nchw_to_nhwc_f32(raw.detach())
    st = ops.col_stats(raw_nhwc.view(-1, C))
    mean, invstd, scale, shift = ops.bn_finalize(st, C, bn.eps, 0.1, bn.weight.detach(), bn.bias.detach(), None, None)
    act, f32, dgamma, dbeta = ops.bn_bwd(raw_nhwc, ops.nchw_to_nhwc_f32(gy), scale, shift, mean, invstd, pool, True,
                                         want_f32=True)
    assert rel_l2(ops.nhwc_f32_to_nchw(f32), raw.grad) <= 1e-4
    assert rel_l2(ops.from_split(act), raw.grad) <= (1e-4 if act.lo is not None else 2.0 ** -8)   # hi plane only: bf16 rounding
    assert rel_l2(dgamma, bn.weight.grad) <= 1e-4 and rel_l2(dbeta, bn.bias.grad) <= 1e-4


def _sp_pair(dev, seed=0, decoder_gain=1.0):
    from utils import make_layers, cfg
    from models.model_SP import model_SP
    torch.manual_seed(seed)
    m = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20))
    torch_ref.randomize_(m, seed)
    with torch.no_grad():
        for mod in m.decoder:
            if isinstance(mod, torch.nn.Conv2d):
                mod.weight.mul_(decoder_gain)
    m = m.to(dev).train()
    return m, copy.deepcopy(m)


def _grad_errors(m, m32, m64):
    """per-parameter (name, rel-L2 of egaze vs fp64, rel-L2 of stock fp32 vs fp64)"""
    rows = []
    for (k, p), (_, q), (_, r) in zip(m.named_parameters(), m32.named_parameters(), m64.named_parameters()):
        assert p.grad is not None, "no grad for %s" % k
        if r.grad.double().norm().item() < 1e-7:  # conv biases in front of a BatchNorm: zero up to rounding noise
            assert p.grad.double().norm().item() < 1e-4, k
            continue
        rows.append((k, rel_l2(p.grad, r.grad), rel_l2(q.grad, r.grad)))
    return rows


@pytest.mark.parametrize("B,S,gain", [(2, 64, 1.0), (2, 64, 0.8), (2, 224, 0.8)])
def test_model_sp_train_step_vs_autograd(cuda_dev, B, S, gain):
    """Train step (forward + floss + backward) vs PyTorch autograd over the same parameters.
    Truth = stock ops in fp64.  Gates are relative to what stock fp32 autograd loses against fp64 on the same step (see the
    module docstring for why stock fp32 itself is ~1e-2 from fp64 here)."""
    import floss as floss_mod
    m, m32 = _sp_pair(cuda_dev, 0, gain)
    m64 = copy.deepcopy(m32).double()
    x_s, x_t, gt = [torch.from_numpy(a).to(cuda_dev) for a in orc.synth_sp_inputs(B, S, 5)]
    out = m(x_s, x_t)
    loss = floss_mod.floss()(out, gt)
    loss.backward()
    l32 = torch_ref.floss_loss(torch_ref.model_sp_forward(m32, x_s, x_t), gt)
    l32.backward()
    o64 = torch_ref.model_sp_forward(m64, x_s.double(), x_t.double())
    l64 = F.binary_cross_entropy(o64, gt.double(), weight=torch_ref.floss_weight(gt).double())
    l64.backward()
    assert abs(loss.item() - l64.item()) <= 1e-3 * abs(l64.item())
    rows = _grad_errors(m, m32, m64)
    worst = max(rows, key=lambda r: r[1])
    med_e = float(np.median([r[1] for r in rows])), float(np.median([r[2] for r in rows]))
    print("B=%d S=%d gain %.1f: egaze-vs-fp64 median %.2e worst %.2e (%s) | stock fp32-vs-fp64 median %.2e worst %.2e"
          % (B, S, gain, med_e[0], worst[1], worst[0], med_e[1], max(r[2] for r in rows)))
    # At the benchmark's resolution every tensor must stay within 3x of what stock fp32 autograd itself loses against fp64 on
    # this step (on that tensor or in the median), floored at 1e-2, and the median within 2x.  At 64x64 with batch 2 the deep
    # layers see 32 pixels per channel: the 1-MMA weight gradient sums too few products for its bf16 roundings to average out
    # and the batch statistics amplify every perturbation -- measured 2-3e-2 there (stock fp32: 0.5-1.3e-2), gated at 5e-2.
    if S >= 224:
        for k, e, n in rows:
            assert e <= max(1e-2, 3.0 * max(n, med_e[1])), "%s: egaze-vs-fp64 %.3e, stock fp32-vs-fp64 %.3e" % (k, e, n)
        assert med_e[0] <= max(1e-2, 2 * med_e[1]), "median egaze-vs-fp64 %.3e vs median stock fp32-vs-fp64 %.3e" % med_e
    else:
        for k, e, n in rows:
            assert e <= 5e-2, "%s: egaze-vs-fp64 %.3e, stock fp32-vs-fp64 %.3e" % (k, e, n)
        assert med_e[0] <= 4e-2, "median egaze-vs-fp64 %.3e vs median stock fp32-vs-fp64 %.3e" % med_e
    for (k, p), (_, r) in zip(m.named_parameters(), m64.named_parameters()):
        if r.grad.double().norm().item() >= 1e-7:
            cos = F.cosine_similarity(p.grad.double().reshape(1, -1), r.grad.reshape(1, -1)).item()
            assert cos >= 0.999, "%s: cosine %.5f" % (k, cos)


def test_model_sp_frozen_trunks(cuda_dev):
    """SP.py:99-102,110: trunk parameters frozen -> no trunk grads, fusion/bn/decoder grads unchanged."""
    import floss as floss_mod
    m, m_ref = _sp_pair(cuda_dev, 1, 0.8)
    for mm in (m, m_ref):
        for p in list(mm.features_s.parameters()) + list(mm.features_t.parameters()):
            p.requires_grad = False
    x_s, x_t, gt = [torch.from_numpy(a).to(cuda_dev) for a in orc.synth_sp_inputs(2, 64, 6)]
    floss_mod.floss()(m(x_s, x_t), gt).backward()
    torch_ref.floss_loss(torch_ref.model_sp_forward(m_ref, x_s, x_t), gt).backward()
    for (k, p), (_, q) in zip(m.named_parameters(), m_ref.named_parameters()):
        if k.startswith("features"):
            assert p.grad is None
        elif q.grad.norm().item() < 1e-6:   # fusion.bias feeds a batch-stat BatchNorm: exactly zero up to rounding noise
            assert p.grad.norm().item() < 1e-4, k
        else:
            assert rel_l2(p.grad, q.grad) <= 5e-2, k


@pytest.mark.parametrize("tag,tol", [("sp_train_wellcond_b4_s32", 6e-2), ("sp_train_b4_s32", 6e-2)])
def test_model_sp_train_step_vs_reference_golden(cuda_dev, tag, tol):
    """Gradients of the UNMODIFIED reference (CPU fp32 fixture) vs the CUDA path: both sides carry routing-flip noise
    (module docstring), hence rel-L2 / norm agreement within 6e-2; BN running stats within 1e-4; loss within 1e-3."""
    import floss as floss_mod
    from test_gpu_golden import load_sd
    from utils import make_layers, cfg
    from models.model_SP import model_SP
    g = np.load(os.path.join(GOLD, tag + ".npz"))
    m = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20))
    m = load_sd(m, orc.synth_state_dict(sp_shapes(), int(g["seed_w"]), float(g["decoder_gain"]))).to(cuda_dev).train()
    x_s, x_t, gt = [torch.from_numpy(a).to(cuda_dev) for a in orc.synth_sp_inputs(int(g["B"]), int(g["S"]), int(g["seed_x"]))]
    loss = floss_mod.floss()(m(x_s, x_t), gt)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    for k, v in m.state_dict().items():
        if "running_" in k:
            assert np.abs(v.cpu().numpy() - g["buf/" + k]).max() <= 1e-4, k
    for k, p in m.named_parameters():
        ref_norm = float(g["gnorm/" + k])
        got_norm = p.grad.double().norm().item()
        if ref_norm < 1e-5:   # conv biases in front of a BatchNorm: exactly zero up to rounding noise
            assert got_norm < 1e-4, k
            continue
        assert abs(got_norm - ref_norm) <= tol * ref_norm, "%s: |g| %.4e vs %.4e" % (k, got_norm, ref_norm)
        if "grad/" + k in g:
            assert rel_l2(p.grad, torch.from_numpy(g["grad/" + k]).to(cuda_dev)) <= tol, k


def test_late_fusion_train_step(cuda_dev):
    import floss as floss_mod
    from models.late_fusion import late_fusion
    from test_gpu_golden import load_sd
    g = np.load(os.path.join(GOLD, "lf_b2_s64.npz"))
    m = load_sd(late_fusion(), orc.synth_state_dict(lf_shapes(), int(g["seed_w"]))).to(cuda_dev).train()
    rs = np.random.RandomState(int(g["seed_x"]))
    f = torch.from_numpy(rs.rand(2, 1, 64, 64).astype(np.float32)).to(cuda_dev)
    gg = torch.from_numpy(rs.rand(2, 1, 64, 64).astype(np.float32)).to(cuda_dev)
    _, _, gt = orc.synth_sp_inputs(2, 64, int(g["seed_gt"]))
    loss = floss_mod.floss()(m(f, gg), torch.from_numpy(gt).to(cuda_dev))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    for k, p in m.named_parameters():
        ref = torch.from_numpy(g["grad/" + k]).to(cuda_dev)
        assert (p.grad - ref).norm().item() <= 1e-2 * ref.norm().item() + 1e-6, "%s: %.3e" % (k, rel_l2(p.grad, ref))


def test_vgg_single_stream_train_step(cuda_dev):
    """spatialstream.py inner loop: frozen trunk (still batch-stat BN), decoder trained with floss."""
    import floss as floss_mod
    from utils import make_layers, cfg
    from egaze.vgg import VGG
    torch.manual_seed(0)
    m = torch_ref.randomize_(VGG(make_layers(cfg['D'], 3)), 2)
    with torch.no_grad():
        for mod in m.decoder:
            if isinstance(mod, torch.nn.Conv2d):
                mod.weight.mul_(0.8)  # keep the logits O(1): well-conditioned regime (see test_model_sp_train_step_vs_autograd)
    m = m.to(cuda_dev).train()
    m_ref = copy.deepcopy(m)
    x_s, _, gt = [torch.from_numpy(a).to(cuda_dev) for a in orc.synth_sp_inputs(2, 64, 8)]
    floss_mod.floss()(m(x_s), gt).backward()
    out_r = torch.sigmoid(torch_ref.seq_forward(m_ref.decoder, torch_ref.seq_forward(m_ref.features, x_s)))
    torch_ref.floss_loss(out_r, gt).backward()
    for (k, p), (_, q) in zip(m.named_parameters(), m_ref.named_parameters()):
        if k.startswith("features"):
            assert p.grad is None
        else:
            assert rel_l2(p.grad, q.grad) <= 5e-2, k


@pytest.mark.parametrize("shape", [(32, 224, 224, 64, 64), (32, 56, 56, 256, 256), (32, 14, 14, 512, 512)])
def test_full_size_adjoint_and_linearity(cuda_dev, shape):
    """BASELINE-size properties that need no oracle (the NumPy oracle cannot finish a B=32, 224x224 layer in seconds):
      * linearity:  conv(x1 + 2*x2) == conv(x1) + 2*conv(x2)
      * adjoint identities of the three tcgen05 kernels on the same operands:
            <conv(x; w), dy> == <x, dgrad(dy; w)> == <w, wgrad(x, dy)>
    all at the batch-32 layer sizes the headline benchmark runs (fp64 dot products of the fp32 outputs).  The operands are
    first rounded to what the kernels see (hi+lo split), so the identities hold to accumulation-order rounding."""
    from egaze import ops
    N, H, W, Cin, Cout = shape
    g = torch.Generator().manual_seed(12)
    # bf16-exact operands: every plane a mode drops (activation / gradient lo planes) is then exactly zero, so the identities
    # hold to accumulation-order rounding in every numeric mode
    rnd = lambda *s: torch.randn(*s, generator=g).to(cuda_dev).bfloat16().float()
    x1, x2, dy = rnd(N, Cin, H, W), rnd(N, Cin, H, W), rnd(N, Cout, H, W)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.05).to(cuda_dev)
    w = w.bfloat16().float() + (w - w.bfloat16().float()).bfloat16().float()     # hi + lo exactly

    def conv(x):
        a = ops.to_split(x)
        _, y, _ = ops.conv3x3(a, ops.pack_cache.get(w, 0, cols_p=a.Cp), want_f32=True, want_split=False)
        return ops.nhwc_f32_to_nchw(y, Cout)

    y1, y2 = conv(x1), conv(x2)
    y12 = conv(ops.from_split(ops.to_split(x1 + 2 * x2)))
    x12 = ops.from_split(ops.to_split(x1 + 2 * x2))
    lin_ref = y1 + 2 * y2 + conv(x12 - (x1 + 2 * x2))      # the split of the sum may round: account for that exactly
    assert rel_l2(y12, lin_ref) <= 2e-5, rel_l2(y12, lin_ref)

    dya = ops.grad_split(dy)
    _, dx, _ = ops.conv3x3(dya, ops.pack_cache.get(w, 1, cols_p=dya.Cp), want_f32=True, want_split=False)
    dx = ops.nhwc_f32_to_nchw(dx, Cin)
    gw = ops.wgrad3x3(ops.to_split(x1, xb=True), dya, Cout, Cin)
    a = (y1.double() * dy.double()).sum().item()
    b = (x1.double() * dx.double()).sum().item()
    c = (w.double() * gw.double()).sum().item()
    scale = (y1.double().norm() * dy.double().norm()).item()
    assert abs(a - b) <= 2e-6 * scale and abs(a - c) <= 2e-6 * scale, (a, b, c, scale)


def test_multi_stream_schedule_matches_single_stream(cuda_dev):
    """The default schedule runs the weight-gradient GEMMs and the temporal trunk on their own CUDA streams.  A training
    step must give the same forward (bit for bit: the forward kernels are deterministic) and the same gradients (weight
    gradients are summed with fp32 atomics, so to rounding) as the single-stream schedule; repeated to catch ordering bugs."""
    import floss as floss_mod

    def run(wgrad_stream, trunk_stream):
        os.environ["EGAZE_WGRAD_STREAM"] = "1" if wgrad_stream else "0"
        os.environ["EGAZE_TRUNK_STREAM"] = "1" if trunk_stream else "0"
        try:
            m, _ = _sp_pair(cuda_dev, 0, 0.8)
            x_s, x_t, gt = [torch.from_numpy(a).to(cuda_dev) for a in orc.synth_sp_inputs(4, 64, 5)]
            out = m(x_s, x_t)
            loss = floss_mod.floss()(out, gt)
            loss.backward()
            torch.cuda.synchronize()
            return out.detach().clone(), [p.grad.detach().clone() for p in m.parameters()]
        finally:
            os.environ.pop("EGAZE_WGRAD_STREAM", None)
            os.environ.pop("EGAZE_TRUNK_STREAM", None)

    ref_out, ref_grads = run(False, False)
    # both on (default), and the mixed settings: with the weight-gradient stream off and the trunk stream on, the two trunks
    # (identical layer shapes) run their weight gradients concurrently -- each stream must own its accumulators
    for knobs in ((True, True), (False, True), (True, False)):
        for _ in range(3):
            out, grads = run(*knobs)
            assert torch.equal(out, ref_out), knobs
            for g, r in zip(grads, ref_grads):
                assert rel_l2(g, r) <= 1e-5 or (g - r).abs().max().item() <= 1e-9, knobs


@pytest.mark.parametrize("frozen_stats_only", [False, True])
def test_model_sp_eval_mode_backward(cuda_dev, frozen_stats_only):
    """BatchNorm on running statistics with autograd on (model.eval(), or only the BatchNorm layers in eval mode: the usual
    'frozen statistics' fine-tuning setup): stock nn.BatchNorm2d supports the backward, and the DDP gradient-equality check
     needs it.  Every parameter gradient against stock autograd."""
    import floss as floss_mod
    m, m_ref = _sp_pair(cuda_dev, 2, 0.8)
    for mm in (m, m_ref):
        if frozen_stats_only:
            for mod in mm.modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.eval()
        else:
            mm.eval()
    m64 = copy.deepcopy(m_ref).double()
    if frozen_stats_only:
        for mod in m64.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.eval()
    x_s, x_t, gt = [torch.from_numpy(a).to(cuda_dev) for a in orc.synth_sp_inputs(2, 64, 9)]
    floss_mod.floss()(m(x_s, x_t), gt).backward()
    torch_ref.floss_loss(torch_ref.model_sp_forward(m_ref, x_s, x_t), gt).backward()
    o64 = torch_ref.model_sp_forward(m64, x_s.double(), x_t.double())
    F.binary_cross_entropy(o64, gt.double(), weight=torch_ref.floss_weight(gt).double()).backward()
    rows = []
    for (k, p), (_, q), (_, r) in zip(m.named_parameters(), m_ref.named_parameters(), m64.named_parameters()):
        assert p.grad is not None, k
        rows.append((k, rel_l2(p.grad, r.grad), rel_l2(q.grad, r.grad)))
    med = float(np.median([r[2] for r in rows]))
    print("eval-mode BatchNorm backward: egaze-vs-fp64 median %.2e worst %.2e | stock fp32-vs-fp64 median %.2e worst %.2e"
          % (np.median([r[1] for r in rows]), max(r[1] for r in rows), med, max(r[2] for r in rows)))
    # 64x64, batch 2 (see test_model_sp_train_step_vs_autograd): the bf16 roundings of the 2-MMA data gradient / 1-MMA weight
    # gradient are what is left once BatchNorm no longer couples the samples -- measured median 1.5e-2, worst 3.7e-2
    for k, e, n in rows:
        assert e <= 5e-2, "%s: egaze-vs-fp64 %.3e, stock fp32-vs-fp64 %.3e" % (k, e, n)
    assert float(np.median([r[1] for r in rows])) <= 2.5e-2
    # running statistics untouched
    for (k, v), (_, r) in zip(m.state_dict().items(), m_ref.state_dict().items()):
        if "running_" in k or "num_batches" in k:
            assert torch.equal(v, r), k
