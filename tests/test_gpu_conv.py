"""GPU parity of the tcgen05 3x3 conv (through the C-ABI) against a plain PyTorch fp32 conv2d of the same operands.

Tolerances: the gated modes (`precise`: fp16 hi+lo split, `precise3`: bf16 hi+lo split; 3 MMAs per product) must agree with
fp32 to ~2^-16 relative per product -> we gate at max-abs <= 3e-4 * scale; `fast` (single bf16 pass) is gated loosely at
3e-2 * scale (reported mode).
"""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mk(N, H, W, Cin, Cout, dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5).to(dev)
    b = (torch.randn(Cout, generator=g) * 0.1).to(dev)
    return x, w, b


def _run(x, w, b, precise, grad=False, **kw):
    """precise=False: the `fast` mode.  grad=True: the operand is a gradient (bf16 planes; hi only in the default mode)."""
    from egaze import ops
    old = os.environ.get("EGAZE_PRECISION")
    if not precise:
        os.environ["EGAZE_PRECISION"] = "fast"
    try:
        Cout, Cin = w.shape[0], w.shape[1]
        act = ops.grad_split(x) if grad else ops.to_split(x)
        cout_p = ops.pad_channels(Cout) if Cout % 16 else Cout
        wp = ops.pack_cache.get(w, 0, rows_p=cout_p, cols_p=act.Cp, fmt=act.fmt)
        bias = b
        if bias is not None and cout_p != Cout:
            bias = torch.cat([b, b.new_zeros(cout_p - Cout)])
        return ops.conv3x3(act, wp, bias=bias, **kw)
    finally:
        if not precise:
            if old is None:
                os.environ.pop("EGAZE_PRECISION", None)
            else:
                os.environ["EGAZE_PRECISION"] = old


SHAPES = [
    # N, H, W, Cin, Cout
    (2, 16, 16, 64, 64),
    (1, 14, 14, 512, 512),
    (2, 28, 28, 256, 512),
    (1, 56, 56, 128, 256),
    (1, 112, 112, 64, 128),
    (1, 224, 224, 64, 64),
    (2, 32, 32, 3, 64),      # KC=16 / SWIZZLE_32B path (RGB conv1_1)
    (2, 32, 32, 20, 64),     # KC=32 / SWIZZLE_64B path (flow conv1_1)
    (2, 32, 32, 32, 8),      # LF conv3: Cout padded to 16
    (2, 18, 18, 128, 128),   # 288-input conv5 size (partial tiles)
    (1, 36, 36, 64, 192),    # BN tile = 64, 3 n-tiles
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("precise", [True, False])
def test_conv3x3_plain(cuda_dev, shape, precise, numeric_mode):
    from egaze import ops
    N, H, W, Cin, Cout = shape
    x, w, b = _mk(N, H, W, Cin, Cout, cuda_dev)
    act_out, f32_out, _ = _run(x, w, b, precise, want_f32=True, want_split=True)
    torch.cuda.synchronize()
    if precise:
        # the kernel sees x, w rounded to 16 mantissa bits (hi+lo); compare against fp32 conv of the same rounded operands
        xr = ops.from_split(ops.to_split(x))
        ref = F.conv2d(xr.double(), w.double(), b.double(), padding=1).float()
        tol = 3e-4
    else:
        ref = F.conv2d(x.double(), w.double(), b.double(), padding=1).float()
        tol = 3e-2
    got = ops.nhwc_f32_to_nchw(f32_out, Cout)
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= tol * scale, "f32 out: max-abs err %.3e (scale %.3e)" % (err, scale)
    got2 = ops.from_split(act_out, Cout)
    err2 = (got2 - ref).abs().max().item()
    assert err2 <= (tol + 2e-5) * scale, "split out: max-abs err %.3e (scale %.3e)" % (err2, scale)


@pytest.mark.parametrize("shape", [(2, 32, 32, 64, 128), (1, 56, 56, 64, 64), (2, 28, 28, 128, 256)])
def test_conv3x3_epilogues(cuda_dev, shape, numeric_mode):
    """bias + folded scale/shift + relu + 2x2 max-pool; relu + nearest-2x replicate; 2x2 sum + mask."""
    from egaze import ops
    N, H, W, Cin, Cout = shape
    x, w, b = _mk(N, H, W, Cin, Cout, cuda_dev, seed=1)
    g = torch.Generator().manual_seed(5)
    scale = (torch.rand(Cout, generator=g) + 0.5).to(cuda_dev)
    shift = (torch.randn(Cout, generator=g) * 0.2).to(cuda_dev)
    xr = ops.from_split(ops.to_split(x))
    conv = F.conv2d(xr.double(), w.double(), b.double(), padding=1).float()
    sc = conv.abs().max().item()
    # (a) eval-BN fold + relu + pool
    a, _, _ = _run(x, w, b, True, scale=scale, shift=shift, relu=True, reduce=1)
    ref = F.max_pool2d(F.relu(conv * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)), 2, 2)
    assert (ops.from_split(a) - ref).abs().max().item() <= 4e-4 * sc
    # (b) relu + upsample
    a, _, _ = _run(x, w, b, True, relu=True, ups=True)
    ref = F.interpolate(F.relu(conv), scale_factor=2, mode="nearest")
    assert (ops.from_split(a) - ref).abs().max().item() <= 4e-4 * sc
    # (c) 2x2 sum + mask (dgrad-of-upsample epilogue)
    mask_src = torch.randn(N, Cout, H // 2, W // 2, generator=g).to(cuda_dev)
    mact = ops.to_split(mask_src, Cout)
    a, f, _ = _run(x, w, None, True, reduce=2, mask=mact.hi, want_f32=True)
    conv_nb = F.conv2d(xr.double(), w.double(), None, padding=1).float()
    ref = F.avg_pool2d(conv_nb, 2, 2) * 4 * (ops.from_split(ops.Act(mact.hi, None, Cout)) > 0).float()
    assert (ops.nhwc_f32_to_nchw(f) - ref).abs().max().item() <= 4e-4 * 4 * sc
    assert (ops.from_split(a) - ref).abs().max().item() <= 5e-4 * 4 * sc
    # (c') the same epilogues the way the decoder's data gradient uses them: bf16 gradient operand (hi plane only in the
    #      default mode -> 2 MMAs per product), bf16 output, mask read from a forward activation's hi plane
    xg = ops.grad_split(x)
    xg_r = ops.from_split(xg)
    conv_g = F.conv2d(xg_r.double(), w.double(), None, padding=1).float()
    for red, msrc in ((2, mask_src), (0, torch.randn(N, Cout, H, W, generator=g).to(cuda_dev))):
        m_act = ops.to_split(msrc, Cout)
        a, _, _ = _run(x, w, None, True, grad=True, reduce=red, mask=m_act.hi, want_lo=ops.mode()["dy_lo"])
        assert a.fmt == 0 and (a.lo is None) == (not ops.mode()["dy_lo"])
        ref_g = (F.avg_pool2d(conv_g, 2, 2) * 4 if red else conv_g) * (msrc > 0).float()
        tol_g = (5e-4 if a.lo is not None else 2.0 ** -8) * 4 * conv_g.abs().max().item()
        assert (ops.from_split(a) - ref_g).abs().max().item() <= tol_g
    # (d) column sums of the stored values (the bias gradient the dgrad epilogue hands to the upstream conv), both for the
    #     masked 2x2-sum epilogue and for the plain masked one; accumulated on top of what the buffer already holds
    for red in (2, 0):
        m_src = mask_src if red else torch.randn(N, Cout, H, W, generator=g).to(cuda_dev)
        m_act = ops.to_split(m_src, Cout)
        cs = torch.full((Cout,), 3.0, device=cuda_dev)
        a, _, _ = _run(x, w, None, True, reduce=red, mask=m_act.hi, colsum=cs)
        got = ops.from_split(a)
        want = got.double().sum((0, 2, 3)) + 3.0
        tol = 1e-4 * got.abs().double().sum((0, 2, 3)).max().item() + 1e-3
        assert (cs.double() - want).abs().max().item() <= tol


@pytest.mark.parametrize("shape", [(4, 32, 32, 64, 64), (2, 14, 14, 128, 512), (3, 28, 28, 64, 128), (2, 36, 36, 32, 32)])
def test_conv3x3_bn_stats(cuda_dev, shape, numeric_mode):
    """Per-tile (mean, M2) partials + finalize == torch batch statistics; running stats follow nn.BatchNorm2d."""
    from egaze import ops
    N, H, W, Cin, Cout = shape
    x, w, b = _mk(N, H, W, Cin, Cout, cuda_dev, seed=2)
    _, raw, st = _run(x, w, b, True, want_f32=True, want_split=False, stats=True)
    bn = torch.nn.BatchNorm2d(Cout).to(cuda_dev).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.1)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    mean, invstd, scale, shift = ops.bn_finalize(st, Cout, bn.eps, bn.momentum, bn.weight.detach(), bn.bias.detach(), rm, rv)
    raw_nchw = ops.nhwc_f32_to_nchw(raw)
    ref_y = bn(raw_nchw)
    ref_mean = raw_nchw.double().mean((0, 2, 3))
    ref_var = raw_nchw.double().var((0, 2, 3), unbiased=False)
    assert (mean.double() - ref_mean).abs().max().item() <= 1e-5
    assert (invstd.double() - 1.0 / torch.sqrt(ref_var + bn.eps)).abs().max().item() <= 1e-4
    assert (rm - bn.running_mean).abs().max().item() <= 1e-5
    assert (rv - bn.running_var).abs().max().item() <= 1e-5
    for pool in (False, True):
        a, f = ops.bn_apply(raw, scale, shift, relu=True, pool=pool, want_f32=True)
        ref = F.relu(ref_y)
        if pool:
            ref = F.max_pool2d(ref, 2, 2)
        assert (ops.nhwc_f32_to_nchw(f) - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
        assert (ops.from_split(a) - ref).abs().max().item() <= 5e-5 * max(1.0, ref.abs().max().item())


def test_conv_plans_match_plain_entry(cuda_dev, monkeypatch):
    """egaze_conv3x3_plan_create / _run (frozen TMA descriptors + tile configuration) == egaze_conv3x3_tc bit for bit,
    and a repeated call with the same arguments reuses its plan."""
    from egaze import ops
    x, w, b = _mk(2, 28, 28, 128, 256, cuda_dev, seed=3)
    act = ops.to_split(x)
    wp = ops.pack_cache.get(w, 0, cols_p=act.Cp)
    monkeypatch.setenv("EGAZE_CONV_PLANS", "0")
    ref, _, _ = ops.conv3x3(act, wp, bias=b, relu=True)
    monkeypatch.setenv("EGAZE_CONV_PLANS", "1")
    n0 = len(ops._plans)
    outs = []
    for _ in range(3):
        o, _, _ = ops.conv3x3(act, wp, bias=b, relu=True)
        outs.append((o.hi.clone(), o.lo.clone()))
        del o          # the allocator hands the next call the same output addresses: same argument tuple, same plan
    assert len(ops._plans) - n0 <= 3      # 1 when the allocator recycles the output addresses (it need not, e.g. under compute-sanitizer)
    for hi, lo in outs:
        assert torch.equal(hi, ref.hi) and torch.equal(lo, ref.lo)
