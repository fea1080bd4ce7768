"""Forward (and, via egaze.autograd, backward) schedules of the hot-path networks on top of egaze.ops.

The nn.Module containers (utils.make_layers trunk, model_SP, late_fusion) keep the reference's parameter
layout (SURVEY 8b "State layout"); this module reads their parameters and drives the fused sm_100a kernels:

  trunk layer, eval BN : conv3x3_tc [BN folded into scale/shift + ReLU (+2x2 max-pool) in the epilogue] -> split act
  trunk layer, train BN: conv3x3_tc [+bias, raw fp32 out, per-tile (mean, M2)] -> bn_finalize (running stats)
                         -> bn_apply [normalise + ReLU (+pool)] -> split act
  decoder layer        : conv3x3_tc [+bias, ReLU (, 2x nearest replicate for the following nn.Upsample)] -> split act
"""
import os

import torch
import torch.nn as nn

from . import ops


class ConvSpec(object):
    """One fused step of a Sequential: conv [+ bn] [+ relu] [+ pool | + upsample].
    ups_folded: the nn.Upsample after this conv is NOT materialised -- the next conv (`sub`) consumes the low-resolution map in
    sub-pixel form (four 2x2-tap phase convolutions with pre-summed weights, 16 instead of 36 MACs per low-resolution pixel)."""
    __slots__ = ("conv", "bn", "relu", "pool", "ups", "ups_folded", "sub")

    def __init__(self, conv):
        self.conv, self.bn, self.relu, self.pool, self.ups = conv, None, False, False, False
        self.ups_folded, self.sub = False, False


def subpixel_enabled():
    """EGAZE_SUBPIXEL=0 materialises every nn.Upsample and runs the direct 3x3 convolution on it (the round-1 schedule)."""
    return os.environ.get("EGAZE_SUBPIXEL", "1") != "0"


def parse_sequential(seq):
    """Group the children of a reference-style nn.Sequential into fused ConvSpecs (+ trailing 1x1 conv if any)."""
    specs, tail = [], None
    for m in seq.children():
        if isinstance(m, nn.Conv2d):
            if m.kernel_size == (3, 3):
                if m.padding != (1, 1) or m.stride != (1, 1) or m.dilation != (1, 1) or m.groups != 1:
                    raise RuntimeError("egaze: only 3x3/pad1/stride1 convs are on the hot path, got %r" % (m,))
                specs.append(ConvSpec(m))
            elif m.kernel_size == (1, 1) and m.out_channels == 1:
                tail = m
            else:
                raise RuntimeError("egaze: unsupported conv on the hot path: %r" % (m,))
        elif isinstance(m, nn.BatchNorm2d):
            specs[-1].bn = m
        elif isinstance(m, nn.ReLU):
            specs[-1].relu = True
        elif isinstance(m, nn.MaxPool2d):
            if m.kernel_size not in (2, (2, 2)) or m.stride not in (2, (2, 2)):
                raise RuntimeError("egaze: only MaxPool2d(2,2) is on the hot path")
            specs[-1].pool = True
        elif isinstance(m, nn.Upsample):
            if m.scale_factor not in (2, 2.0) or m.mode != "nearest":
                raise RuntimeError("egaze: only nearest Upsample(scale_factor=2) is on the hot path")
            specs[-1].ups = True
        else:
            raise RuntimeError("egaze: unsupported module on the hot path: %r" % (m,))
    if subpixel_enabled():
        # conv+ReLU+Upsample -> conv+ReLU -> conv (the decoder pattern, models/model_SP.py:16-17,20-21,24-25,27-28): the middle
        # conv runs in sub-pixel form; in the backward the following conv's data gradient hands it a phase-planar gradient
        for i in range(1, len(specs) - 1):
            prev, sp, nxt = specs[i - 1], specs[i], specs[i + 1]
            if (prev.ups and prev.bn is None and sp.bn is None and nxt.bn is None and not sp.pool and not nxt.pool
                    and sp.conv.in_channels % 64 == 0 and sp.conv.out_channels % 64 == 0):
                prev.ups_folded, sp.sub = True, True
    return specs, tail


def _nbt(bn):
    """nn.BatchNorm2d.num_batches_tracked (int64 device scalar) when the layer tracks it: bn_finalize bumps it in its own launch."""
    t = bn.num_batches_tracked if bn.track_running_stats else None
    return t if (t is not None and t.is_cuda and t.dtype == torch.int64) else None


def _bn_uses_batch_stats(bn):
    return bn.training or (bn.running_mean is None and bn.running_var is None)


def run_conv_spec(act, spec, saved=None, xb=False):
    """Execute one ConvSpec on a split activation.  `saved` (list) collects what backward needs.  xb: the output feeds a conv
    whose weight gradient will be computed, so (fp16 forward) it also carries its bf16 copy."""
    conv, bn = spec.conv, spec.bn
    xb = ops.want_xb(xb)
    cout_p = ops.pad_channels(conv.out_channels) if conv.out_channels % 16 else conv.out_channels
    wpack = ops.pack_cache.get(conv.weight, 2 if spec.sub else 0, rows_p=cout_p, cols_p=act.Cp, fmt=act.fmt)
    bias = conv.bias.detach() if conv.bias is not None else None
    if bias is not None and cout_p != conv.out_channels:
        bias = torch.cat([bias, bias.new_zeros(cout_p - conv.out_channels)])
    if bn is None:
        out, _, _ = ops.conv3x3(act, wpack, bias=bias, relu=spec.relu, reduce=1 if spec.pool else 0,
                                ups=spec.ups and not spec.ups_folded, xb=xb, sub=1 if spec.sub else 0)
        if saved is not None:
            saved.append({"x": act, "y": out})
        out.C = conv.out_channels
        return out
    C = conv.out_channels
    gamma = bn.weight.detach() if bn.weight is not None else None
    beta = bn.bias.detach() if bn.bias is not None else None
    if cout_p != C:
        gamma = torch.cat([gamma, gamma.new_ones(cout_p - C)]) if gamma is not None else None
        beta = torch.cat([beta, beta.new_zeros(cout_p - C)]) if beta is not None else None
    if not _bn_uses_batch_stats(bn):
        rm, rv = bn.running_mean, bn.running_var
        if cout_p != C:
            rm = torch.cat([rm, rm.new_zeros(cout_p - C)])
            rv = torch.cat([rv, rv.new_ones(cout_p - C)])
        if saved is not None:
            # eval-mode BatchNorm with autograd on (stock nn.BatchNorm2d supports it): keep the raw conv output so that the
            # backward can use the same kernels as the batch-statistics case, with the statistics terms switched off
            if spec.ups:
                raise RuntimeError("egaze: conv+BN followed by Upsample is not on the hot path")
            _, raw, _ = ops.conv3x3(act, wpack, bias=bias, want_f32=True, want_split=False)
            scale, shift = ops.bn_fold(gamma, beta, rm, rv, None, bn.eps)
            out, _ = ops.bn_apply(raw, scale, shift, relu=spec.relu, pool=spec.pool, xb=xb)
            saved.append({"x": act, "raw": raw, "mean": rm.detach(), "invstd": torch.rsqrt(rv.detach() + bn.eps), "scale": scale,
                          "shift": shift, "y": out, "eval": True})
            out.C = C
            return out
        scale, shift = ops.bn_fold(gamma, beta, rm, rv, bias, bn.eps)
        out, _, _ = ops.conv3x3(act, wpack, scale=scale, shift=shift, relu=spec.relu, reduce=1 if spec.pool else 0,
                                ups=spec.ups, xb=xb)
        if saved is not None:
            saved.append({"x": act, "y": out, "scale": scale})
        out.C = C
        return out
    # training-mode BatchNorm: batch statistics are needed before normalising
    if spec.ups:
        raise RuntimeError("egaze: conv+BN followed by Upsample is not on the hot path")
    _, raw, st = ops.conv3x3(act, wpack, bias=bias, want_f32=True, want_split=False, stats=True)
    if bn.momentum is None:
        raise RuntimeError("egaze: BatchNorm2d(momentum=None) (cumulative average) is not supported")
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    if cout_p != C and rm is not None:
        rm_p = torch.cat([rm, rm.new_zeros(cout_p - C)])
        rv_p = torch.cat([rv, rv.new_ones(cout_p - C)])
        mean, invstd, scale, shift = ops.bn_finalize(st, cout_p, bn.eps, bn.momentum, gamma, beta, rm_p, rv_p, _nbt(bn))
        rm.copy_(rm_p[:C])
        rv.copy_(rv_p[:C])
    else:
        mean, invstd, scale, shift = ops.bn_finalize(st, cout_p, bn.eps, bn.momentum, gamma, beta, rm, rv, _nbt(bn))
    out, _ = ops.bn_apply(raw, scale, shift, relu=spec.relu, pool=spec.pool, xb=xb)
    if saved is not None:
        saved.append({"x": act, "raw": raw, "mean": mean, "invstd": invstd, "scale": scale, "shift": shift, "y": out})
    out.C = C
    return out


def xb_flags(specs, after_last=False, recording=True):
    """Per spec: does its OUTPUT need the bf16 copy, i.e. will the NEXT conv's weight gradient be computed?"""
    flags = [recording and specs[i + 1].conv.weight.requires_grad for i in range(len(specs) - 1)]
    return flags + [bool(recording and after_last)]


def run_sequential(seq, act, saved=None, xb_after_last=False):
    """Run a reference-style Sequential of 3x3 convs.  Returns (act, tail_1x1_conv | None)."""
    specs, tail = parse_sequential(seq)
    for spec, xb in zip(specs, xb_flags(specs, xb_after_last, saved is not None)):
        act = run_conv_spec(act, spec, saved, xb)
    return act, tail


def attach_act(t, act):
    """Side channel: keep the internal split-NHWC activation next to the NCHW fp32 tensor the API returns."""
    t._egaze_act = act
    return t


def get_act(t, Cp=None):
    act = getattr(t, "_egaze_act", None)
    if act is not None and act.N == t.shape[0] and act.C == t.shape[1] and (Cp is None or act.Cp == Cp) \
            and act.fmt == ops.mode()["fwd_fmt"]:
        return act
    return ops.to_split(t, Cp)


# ---- model_SP tail: fusion (shared conv on both streams + max) -> bn -> relu -> decoder -> sigmoid ----------------
def run_sp_tail(model, a_s, a_t, saved=None):
    """models/model_SP.py:38-49.  a_s / a_t: split conv5_3 activations of the spatial / temporal trunks."""
    B = a_s.N
    fusion, bn = model.fusion, model.bn
    xb_cat = torch.cat([a_s.xb, a_t.xb], 0) if (a_s.xb is not None and a_t.xb is not None) else None
    cat = ops.Act(torch.cat([a_s.hi, a_t.hi], 0), torch.cat([a_s.lo, a_t.lo], 0), a_s.C, xb_cat)  # depth order (s, t): model_SP.py:40
    wpack = ops.pack_cache.get(fusion.weight, 0, cols_p=cat.Cp, fmt=cat.fmt)
    bias = fusion.bias.detach() if fusion.bias is not None else None
    _, raw2, _ = ops.conv3x3(cat, wpack, bias=bias, want_f32=True, want_split=False)
    mx = ops.pairmax(raw2)  # [B,14,14,512] fp32
    C = mx.shape[-1]
    gamma = bn.weight.detach() if bn.weight is not None else None
    beta = bn.bias.detach() if bn.bias is not None else None
    rec = {"cat": cat, "raw2": raw2, "mx": mx}
    if _bn_uses_batch_stats(bn):
        st = ops.col_stats(mx.view(-1, C))
        rm = bn.running_mean if bn.track_running_stats else None
        rv = bn.running_var if bn.track_running_stats else None
        mean, invstd, scale, shift = ops.bn_finalize(st, C, bn.eps, bn.momentum, gamma, beta, rm, rv, _nbt(bn))
        rec.update(mean=mean, invstd=invstd)
    else:
        scale, shift = ops.bn_fold(gamma, beta, bn.running_mean, bn.running_var, None, bn.eps)
        rec.update(mean=bn.running_mean.detach(), invstd=torch.rsqrt(bn.running_var.detach() + bn.eps), eval=True)
    dec0 = next(m for m in model.decoder.children() if isinstance(m, nn.Conv2d))
    act, _ = ops.bn_apply(mx, scale, shift, relu=True, pool=False, xb=ops.want_xb(saved is not None and dec0.weight.requires_grad))
    rec.update(scale=scale, shift=shift, y=act)
    dec_saved = [] if saved is not None else None
    act, tail = run_sequential(model.decoder, act, dec_saved)
    if tail is None:
        raise RuntimeError("egaze: decoder must end with the 1x1 conv (model_SP.py:30)")
    out = ops.head_fwd(act, tail.weight, tail.bias)
    if saved is not None:
        rec.update(decoder=dec_saved, head_in=act, out=out)
        saved.append(rec)
    return out
