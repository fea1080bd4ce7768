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


# ---- sub-pixel form of nn.Upsample(scale_factor=2) -> conv3x3 (model_SP.py:16-17,20-21,24-25,27-28) ------------------------------
SUB_SHAPES = [
    # N, H (low-res), W, Cin, Cout
    (2, 14, 14, 128, 128),   # partial tiles (14 rows), 2 K chunks
    (1, 28, 28, 256, 128),
    (2, 16, 24, 64, 64),     # one K chunk, Cout = 64 (single-CTA / pair split)
    (1, 56, 56, 128, 64),    # the decoder's last upsample-fed layer shape family
]


def _planar(t):
    """[N, C, 2H, 2W] -> phase-planar [4N, C, H, W]: image (py*2+px)*N + n holds t[n, :, py::2, px::2]."""
    return torch.cat([t[:, :, py::2, px::2] for py in (0, 1) for px in (0, 1)], 0).contiguous()


@pytest.mark.parametrize("shape", SUB_SHAPES)
def test_conv3x3_subpixel_forward(cuda_dev, shape, numeric_mode):
    """sub = 1 on the low-resolution map == conv3x3(upsample_nearest_2x(x)) (bias + ReLU epilogue), and == the kernel's own direct
    path on the materialised upsampled map to rounding."""
    from egaze import ops
    N, H, W, Cin, Cout = shape
    x, w, b = _mk(N, H, W, Cin, Cout, cuda_dev, seed=7)
    act = ops.to_split(x)
    wp = ops.pack_cache.get(w, 2, cols_p=act.Cp, fmt=act.fmt)
    out, f32, _ = ops.conv3x3(act, wp, bias=b, relu=True, sub=1, want_f32=True)
    xr = ops.from_split(act)
    ref = F.relu(F.conv2d(F.interpolate(xr.double(), scale_factor=2, mode="nearest"), w.double(), b.double(), padding=1)).float()
    sc = ref.abs().max().item()
    assert tuple(f32.shape) == (N, 2 * H, 2 * W, Cout)
    err = (ops.nhwc_f32_to_nchw(f32, Cout) - ref).abs().max().item()
    assert err <= 3e-4 * sc, "f32 out: %.3e (scale %.3e)" % (err, sc)
    err2 = (ops.from_split(out, Cout) - ref).abs().max().item()
    assert err2 <= 3.2e-4 * sc, "split out: %.3e (scale %.3e)" % (err2, sc)


@pytest.mark.parametrize("shape", SUB_SHAPES)
def test_conv3x3_subpixel_backward(cuda_dev, shape, numeric_mode):
    """sub = 2 (data gradient from a phase-planar dY, masked, with the bias-gradient column sums), the phase-planar store of an
    ordinary data-gradient launch, and the sub-pixel weight gradient -- against fp64 autograd of conv3x3(upsample(x))."""
    from egaze import ops
    N, H, W, Cin, Cout = shape
    x, w, b = _mk(N, H, W, Cin, Cout, cuda_dev, seed=9)
    g = torch.Generator().manual_seed(11)
    dy = torch.randn(N, Cout, 2 * H, 2 * W, generator=g).to(cuda_dev)
    x_act = ops.to_split(x, xb=True)
    dy_act = ops.grad_split(_planar(dy))                      # [4N, H, W, Cout] bf16 planes
    # reference on the operands the kernels see
    xq = (x_act.xb.float() if x_act.xb is not None else ops.from_split_nhwc(x_act)).permute(0, 3, 1, 2)[:, :Cin].double()
    dyq = ops.from_split_nhwc(dy_act).permute(0, 3, 1, 2).double()     # planar, rounded
    dyq_full = torch.zeros(N, Cout, 2 * H, 2 * W, dtype=torch.float64, device=cuda_dev)
    for ph, (py, px) in enumerate([(0, 0), (0, 1), (1, 0), (1, 1)]):
        dyq_full[:, :, py::2, px::2] = dyq[ph * N:(ph + 1) * N]
    xq.requires_grad_(True)
    wd = w.double().requires_grad_(True)
    y = F.conv2d(F.interpolate(xq, scale_factor=2, mode="nearest"), wd, None, padding=1)
    y.backward(dyq_full)
    # (a) weight gradient
    gw = ops.wgrad3x3(x_act, dy_act, Cout, Cin, sub=True)
    e = ((gw.double() - wd.grad).norm() / wd.grad.norm()).item()
    assert e <= (2e-4 if ops.mode()["wgrad_precise"] or x_act.fmt == 1 else 3e-3), "wgrad rel-L2 %.3e" % e
    # (b) data gradient, fp32 out
    wp = ops.pack_cache.get(w, 3, cols_p=dy_act.Cp)
    _, gx, _ = ops.conv3x3(dy_act, wp, sub=2, want_f32=True, want_split=False, want_lo=ops.mode()["dy_lo"])
    assert tuple(gx.shape) == (N, H, W, Cin)
    ref = xq.grad.float()
    err = (ops.nhwc_f32_to_nchw(gx, Cin) - ref).abs().max().item()
    assert err <= 3e-4 * ref.abs().max().item(), "dgrad max-abs %.3e (scale %.3e)" % (err, ref.abs().max().item())
    # (c) masked, bf16 out, column sums
    mask_src = torch.randn(N, Cin, H, W, generator=g).to(cuda_dev)
    mact = ops.to_split(mask_src, Cin)
    cs = torch.zeros(Cin, device=cuda_dev)
    gm, _, _ = ops.conv3x3(dy_act, wp, sub=2, mask=mact.hi, colsum=cs, want_lo=ops.mode()["dy_lo"])
    refm = ref * (mact.hi.float().permute(0, 3, 1, 2) > 0)
    got = ops.from_split_nhwc(gm).permute(0, 3, 1, 2)
    tol = 3e-4 if ops.mode()["dy_lo"] else 5e-3          # hi-only bf16 store: 2^-9 relative
    assert (got - refm).abs().max().item() <= tol * ref.abs().max().item()
    ecs = (cs - refm.sum((0, 2, 3))).abs().max().item()
    assert ecs <= 3e-4 * refm.abs().sum((0, 2, 3)).max().item() + 1e-3, "colsum %.3e" % ecs


def test_conv3x3_planar_store(cuda_dev, numeric_mode):
    """out_planar: the same values as the ordinary store, laid out [4N][H/2][W/2][C] by output-pixel parity (masked dgrad mode)."""
    from egaze import ops
    N, H, W, Cin, Cout = 2, 28, 24, 64, 128
    x, w, b = _mk(N, H, W, Cout, Cin, cuda_dev, seed=13)     # a data-gradient launch: operand has the conv's Cout channels
    wt = torch.randn(Cout, Cin, 3, 3, device=cuda_dev) * 0.05
    act = ops.grad_split(x)
    wp = ops.pack_cache.get(wt, 1, cols_p=act.Cp)
    g = torch.Generator().manual_seed(3)
    mact = ops.to_split(torch.randn(N, Cin, H, W, generator=g).to(cuda_dev), Cin)
    cs0, cs1 = torch.zeros(Cin, device=cuda_dev), torch.zeros(Cin, device=cuda_dev)
    a, _, _ = ops.conv3x3(act, wp, mask=mact.hi, colsum=cs0, want_lo=ops.mode()["dy_lo"])
    p, _, _ = ops.conv3x3(act, wp, mask=mact.hi, colsum=cs1, want_lo=ops.mode()["dy_lo"], planar=True)
    assert tuple(p.hi.shape) == (4 * N, H // 2, W // 2, Cin)
    ref = torch.cat([a.hi[:, py::2, px::2] for py in (0, 1) for px in (0, 1)], 0)
    assert torch.equal(p.hi, ref)
    if a.lo is not None:
        assert torch.equal(p.lo, torch.cat([a.lo[:, py::2, px::2] for py in (0, 1) for px in (0, 1)], 0))
    assert torch.allclose(cs0, cs1, rtol=1e-5, atol=1e-5)
    _, f, _ = ops.conv3x3(act, wp, want_f32=True, want_split=False, planar=True)
    _, f0, _ = ops.conv3x3(act, wp, want_f32=True, want_split=False)
    assert torch.equal(f, torch.cat([f0[:, py::2, px::2] for py in (0, 1) for px in (0, 1)], 0))


def test_subpixel_weight_pack_roundtrip(cuda_dev):
    """The 16-plane pack (modes 2 / 3) is the fp32 pre-sum of the 3x3 taps, and the sub-pixel unpack is its transpose."""
    from egaze import ops
    Co, Ci = 64, 128
    w = torch.randn(Co, Ci, 3, 3, device=cuda_dev)
    V = {(0, 0): [0], (0, 1): [1, 2], (1, 0): [0, 1], (1, 1): [2]}
    ref = torch.zeros(16, Co, Ci, device=cuda_dev)
    for py in (0, 1):
        for px in (0, 1):
            for a in (0, 1):
                for b in (0, 1):
                    for r in V[(py, a)]:
                        for s in V[(px, b)]:
                            ref[(py * 2 + px) * 4 + a * 2 + b] += w[:, :, r, s]
    hi, lo, _, _, _ = ops.pack_cache.get(w, 2, cols_p=Ci, fmt=0)
    assert (hi.float() + lo.float() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    hi3, lo3, _, _, _ = ops.pack_cache.get(w, 3, cols_p=Co, fmt=0)
    assert (hi3.float() + lo3.float() - ref.transpose(1, 2)).abs().max().item() <= 2e-5 * ref.abs().max().item()
    q = torch.randn(16, Co, Ci, device=cuda_dev)
    gw = torch.empty(Co, Ci, 3, 3, device=cuda_dev)
    ops.call("egaze_unpack_wgrad", q.clone(), Co, Ci, Co, Ci, 0.0, 0, 1, gw, ops.stream_ptr())
    refg = torch.zeros_like(gw)
    for py in (0, 1):
        for px in (0, 1):
            for a in (0, 1):
                for b in (0, 1):
                    for r in V[(py, a)]:
                        for s in V[(px, b)]:
                            refg[:, :, r, s] += q[(py * 2 + px) * 4 + a * 2 + b]
    assert torch.allclose(gw, refg, rtol=1e-6, atol=1e-6)
