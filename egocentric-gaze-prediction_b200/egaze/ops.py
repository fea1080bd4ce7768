"""Thin tensor-level wrappers over the C-ABI (include/egaze.h).  PyTorch is used for device memory and streams
only; all arithmetic happens in libegaze.so."""
import ctypes
import os
import weakref

import torch

from . import _lib
from ._lib import call, stream_ptr

BF16 = torch.bfloat16
F16 = torch.float16
F32 = torch.float32

# Numeric modes (EGAZE_PRECISION).  Every tensor-core contraction accumulates in fp32 (TMEM); the modes differ in how the
# fp32 operands are presented to the 16-bit tensor-core inputs:
#   precise  (default) forward: fp16 hi+lo split of activations AND weights, 3 MMAs per product (hi*hi + hi*lo + lo*hi): 22
#            significand bits, what train-mode BatchNorm needs for the 1e-3 gate.  Backward: gradients stay bf16
#            (they need fp32's exponent range): data gradient dY_hi x [W_hi | W_lo] (2 MMAs), weight gradient dY_hi x X_hi
#            (1 MMA).  Measured (tools/grad_modes.py): the cheaper backward moves the parameter gradients by a median 5e-3
#            rel-L2, below the 1.4e-2 by which stock fp32 autograd itself differs from fp64 on the same step (ReLU / max-pool
#            routing flips), and the errors do not compound in the weight gradients.
#   precise3 round-1 scheme: bf16 hi+lo split (16 bits), 3 MMAs per product in forward, data and weight gradient.
#   fast     one bf16 MMA per product everywhere -- reported, never gated.
_MODES = {
    "precise": dict(fwd_fmt=1, fwd_lo=True, w_lo=True, dy_lo=False, wgrad_precise=False,
                    dtype="fp16 hi+lo split operands, 3 MMAs per product (forward); bf16 gradients: 2 MMAs per product (dgrad: "
                          "dY_hi x W_hi+lo), 1 MMA (wgrad); fp32 accumulate"),
    "precise3": dict(fwd_fmt=0, fwd_lo=True, w_lo=True, dy_lo=True, wgrad_precise=True,
                     dtype="bf16 hi+lo split operands, 3 MMAs per product (fwd, dgrad, wgrad), fp32 accumulate"),
    "fast": dict(fwd_fmt=0, fwd_lo=False, w_lo=False, dy_lo=False, wgrad_precise=False,
                 dtype="bf16, 1 MMA per product, fp32 accumulate"),
}


def precision():
    p = os.environ.get("EGAZE_PRECISION", "precise")
    if p not in _MODES:
        raise RuntimeError("EGAZE_PRECISION must be one of %s, got %r" % (sorted(_MODES), p))
    return p


def mode():
    return _MODES[precision()]


def is_precise():
    return precision() != "fast"


def dtype_string():
    """What bench.py reports as `dtype`: the arithmetic type of the tensor-core contractions, MMAs per product included."""
    return mode()["dtype"]


_wscale = [None]


def f16_weight_scale():
    if _wscale[0] is None:
        v = ctypes.c_float(0.0)
        call("egaze_f16_weight_scale", ctypes.addressof(v))
        _wscale[0] = float(v.value)
    return _wscale[0]


def pad_channels(c):
    """Channel padding of the K dimension: multiples of 64 when >= 64, else 16/32/48."""
    if c % 64 == 0:
        return c
    if c > 64:
        return (c + 63) // 64 * 64
    return (c + 15) // 16 * 16


class Act(object):
    """NHWC split activation: x ~= hi + lo, both planes bf16 (gradients; round-1 forward) or fp16 (forward operands).
    `xb` (optional) = bf16(x): what the weight-gradient GEMM reads when hi / lo are fp16.  `C` = logical channels,
    hi.shape[-1] = padded channels."""
    __slots__ = ("hi", "lo", "C", "xb")

    def __init__(self, hi, lo, C, xb=None):
        self.hi, self.lo, self.C, self.xb = hi, lo, C, xb

    @property
    def fmt(self):
        """0 = bf16 planes, 1 = fp16 planes (the C-ABI's `fmt`)."""
        return 1 if self.hi.dtype == F16 else 0

    @property
    def N(self):
        return self.hi.shape[0]

    @property
    def H(self):
        return self.hi.shape[1]

    @property
    def W(self):
        return self.hi.shape[2]

    @property
    def Cp(self):
        return self.hi.shape[3]


def empty_act(N, H, W, Cp, C, device, lo=True, fmt=0, xb=False):
    dt = F16 if fmt else BF16
    hi = torch.empty((N, H, W, Cp), dtype=dt, device=device)
    lo_t = torch.empty((N, H, W, Cp), dtype=dt, device=device) if lo else None
    xb_t = torch.empty((N, H, W, Cp), dtype=BF16, device=device) if xb else None
    return Act(hi, lo_t, C, xb_t)


def want_xb(flag):
    """A forward activation needs its bf16 copy iff a weight gradient will read it and the forward planes are fp16."""
    return bool(flag) and mode()["fwd_fmt"] == 1


def to_split(x, Cp=None, fmt=None, xb=False):
    """NCHW fp32 -> Act (channels zero-padded to Cp) in the forward operand format of the current mode."""
    _lib.check_device(x.device)
    x = x.contiguous().float()
    N, C, H, W = x.shape
    Cp = pad_channels(C) if Cp is None else Cp
    fmt = mode()["fwd_fmt"] if fmt is None else fmt
    act = empty_act(N, H, W, Cp, C, x.device, lo=True, fmt=fmt)
    cp_xb = (Cp + 63) // 64 * 64      # the weight-gradient GEMM reads 64-channel rows; the forward conv keeps the narrow K
    if xb and fmt == 1:
        act.xb = torch.empty((N, H, W, cp_xb), dtype=BF16, device=x.device)
    call("egaze_nchw_to_nhwc_split", x, N, C, H, W, Cp, act.hi, act.lo, act.xb, cp_xb, fmt, stream_ptr())
    return act


def grad_split(x, Cp=None):
    """NCHW fp32 gradient -> bf16 Act (gradients keep bf16's exponent range; the lo plane exists only in modes that use it)."""
    act = to_split(x, Cp, fmt=0)
    if not mode()["dy_lo"]:
        act.lo = None
    return act


def from_split(act, C=None):
    """Act -> NCHW fp32 [N, C, H, W]."""
    C = act.C if C is None else C
    out = torch.empty((act.N, C, act.H, act.W), dtype=F32, device=act.hi.device)
    call("egaze_nhwc_to_nchw", act.hi, act.lo, None, act.N, C, act.H, act.W, act.Cp, act.fmt, out, stream_ptr())
    return out


def from_split_nhwc(act):
    """Act -> NHWC fp32 (same channel stride)."""
    x = act.hi.float()
    if act.lo is not None:
        x += act.lo.float()
    return x


def nhwc_f32_to_nchw(x, C=None):
    N, H, W, Cs = x.shape
    C = Cs if C is None else C
    out = torch.empty((N, C, H, W), dtype=F32, device=x.device)
    call("egaze_nhwc_to_nchw", None, None, x, N, C, H, W, Cs, 0, out, stream_ptr())
    return out


def nchw_to_nhwc_f32(x):
    x = x.contiguous().float()
    N, C, H, W = x.shape
    out = torch.empty((N, H, W, C), dtype=F32, device=x.device)
    call("egaze_nchw_to_nhwc_f32", x, N, C, H, W, out, stream_ptr())
    return out


def f32_to_split(x_nhwc, C=None, fmt=0):
    x_nhwc = x_nhwc.contiguous()
    N, H, W, Cs = x_nhwc.shape
    act = empty_act(N, H, W, Cs, Cs if C is None else C, x_nhwc.device, lo=True, fmt=fmt)
    call("egaze_f32_to_split", x_nhwc, x_nhwc.numel(), act.hi, act.lo, fmt, stream_ptr())
    return act


# ---- packed-weight cache: derived copies keyed on (parameter identity, version) ---------------------------------
class _PackCache(object):
    def __init__(self):
        self._d = {}
        self._tables = {}

    @staticmethod
    def _dims(w, mode, rows_p, cols_p):
        Co, Ci = int(w.shape[0]), int(w.shape[1])
        rows = Co if mode in (0, 2) else Ci
        cols = Ci if mode in (0, 2) else Co
        return Co, Ci, rows, (rows if rows_p is None else rows_p), (pad_channels(cols) if cols_p is None else cols_p)

    def get(self, w, mode, rows_p=None, cols_p=None, fmt=None):
        """w: nn.Conv2d weight [Co, Ci, 3, 3] (or Conv3d [Co, Ci, 1, 3, 3]).  mode 0: forward operand, 1: flipped / transposed
        (data gradient); 2 / 3: the 16-plane sub-pixel forms of 0 / 1 for a conv that follows nn.Upsample (egaze.h).  fmt 1: fp16
        planes pre-scaled by f16_weight_scale(); default: the numeric mode's forward format for the forward copies, bf16 for the
        gradient copies.  Returns (hi, lo, rows, cols_p, fmt)."""
        planes = 9 if mode < 2 else 16
        if fmt is None:
            fmt = _MODES[precision()]["fwd_fmt"] if mode in (0, 2) else 0
        Co, Ci, rows, rp, cp = self._dims(w, mode, rows_p, cols_p)
        dt = F16 if fmt else BF16
        key = (id(w), mode, rp, cp, fmt)
        ent = self._d.get(key)
        ver = w._version
        if ent is not None and ent[0]() is w and ent[1] == ver and ent[2] == w.data_ptr():
            return ent[3]
        w4 = w.detach().reshape(Co, Ci, 3, 3).contiguous().float()
        if rp == rows:
            # the pack kernel writes every element (padding columns included): reuse the previous buffers when possible
            if ent is not None and ent[3][0].shape == (planes, rp, cp) and ent[3][0].device == w.device:
                hi, lo = ent[3][0], ent[3][1]
            else:
                hi = torch.empty((planes, rp, cp), dtype=dt, device=w.device)
                lo = torch.empty((planes, rp, cp), dtype=dt, device=w.device)
            call("egaze_pack_w3x3", w4, Co, Ci, cp, mode, fmt, hi, lo, stream_ptr())
        else:
            hi = torch.zeros((planes, rp, cp), dtype=dt, device=w.device)
            lo = torch.zeros((planes, rp, cp), dtype=dt, device=w.device)
            # padded rows (tiny layers only): pack densely then copy into the padded buffer
            thi = torch.empty((planes, rows, cp), dtype=dt, device=w.device)
            tlo = torch.empty((planes, rows, cp), dtype=dt, device=w.device)
            call("egaze_pack_w3x3", w4, Co, Ci, cp, mode, fmt, thi, tlo, stream_ptr())
            hi[:, :rows].copy_(thi)
            lo[:, :rows].copy_(tlo)
        val = (hi, lo, rp, cp, fmt)
        self._d[key] = (weakref.ref(w), ver, w.data_ptr(), val)
        return val

    def entries_for(self, w):
        """Packed copies of weight w the fused optimiser can maintain: [(mode, rows_p, cols_p, fmt, hi, lo)] (one per mode; copies
        with padded rows are excluded -- their padding is written by get() only)."""
        out = []
        for (wid, mode, rp, cp, fmt), ent in self._d.items():
            if wid == id(w) and ent[0]() is w and ent[2] == w.data_ptr():
                rows = int(w.shape[0]) if mode in (0, 2) else int(w.shape[1])
                if rp == rows and not any(o[0] == mode for o in out):
                    out.append((mode, rp, cp, fmt, ent[3][0], ent[3][1]))
        return out

    def mark_current(self, w, entries):
        """The copies in `entries` (from entries_for) were just rewritten from w's current values (egaze.optim.Adam)."""
        for mode, rp, cp, fmt, _, _ in entries:
            key = (id(w), mode, rp, cp, fmt)
            ent = self._d.get(key)
            if ent is not None:
                self._d[key] = (ent[0], w._version, ent[2], ent[3])

    def held_tables(self):
        """Every job table currently alive (egaze/graph.py keeps the list next to a captured graph so that the memory a
        captured re-pack launch reads can never be recycled while the graph exists)."""
        return list(self._tables.values())

    def mark_stale(self):
        """Forget that any packed copy is current: the next get() / refresh() re-packs it.  For weights that change without
        their tensor version changing (an optimiser step replayed inside a CUDA graph, egaze/graph.py)."""
        for key, ent in list(self._d.items()):
            self._d[key] = (ent[0], -1, ent[2], ent[3])

    def refresh(self, min_stale=2):
        """Re-pack, in ONE launch, every cached copy whose weight changed since it was packed (after an optimiser step that is
        all of them: 76 small launches per SP training step otherwise).  Copies of dead or re-allocated weights are dropped.
        With fewer than min_stale stale copies nothing happens (get() re-packs a single one lazily)."""
        stale = []
        for key, ent in list(self._d.items()):
            w = ent[0]()
            if w is None or ent[2] != w.data_ptr():
                del self._d[key]
                continue
            _, mode, rp, cp, _fmt = key
            Co, Ci, rows, _, _ = self._dims(w, mode, rp, cp)
            if ent[1] != w._version and rp == rows and w.is_contiguous() and w.dtype == F32:
                stale.append((key, ent, w, Co, Ci, rows))
        if len(stale) < max(1, min_stale):
            return
        tkey = tuple((k, e[2], e[3][0].data_ptr()) for k, e, _, _, _, _ in stale)
        table = self._tables.get(tkey)
        if table is None:
            import numpy as np
            rec = np.zeros((len(stale), 6), dtype=np.int64)
            for i, (key, ent, w, Co, Ci, rows) in enumerate(stale):
                hi, lo, rp, cp, fmt = ent[3]
                rec[i] = (w.data_ptr(), hi.data_ptr(), lo.data_ptr(), Co | (Ci << 32), rows | (cp << 32), key[1] | (fmt << 32))
            # Tables are never freed behind a CUDA graph's back: a captured egaze_pack_w3x3_multi launch has this table's
            # address baked in.  Old tables age out of a small LRU; graphs keep their own references (held_tables()).
            while len(self._tables) >= 16:
                self._tables.pop(next(iter(self._tables)))
            table = self._tables[tkey] = torch.from_numpy(rec).to(stale[0][2].device)
        else:
            self._tables[tkey] = self._tables.pop(tkey)   # most recently used last
        self._last_table = table
        call("egaze_pack_w3x3_multi", table, len(stale), stream_ptr())
        for key, ent, w, _, _, _ in stale:
            self._d[key] = (ent[0], w._version, ent[2], ent[3])


pack_cache = _PackCache()


# optional per-launch timing of the tensor-core conv (bench.py roofline): CUDA events on the launching stream
_conv_timer = {"on": False, "events": [], "prof": False}


def conv_timer_reset(enable):
    _conv_timer["on"] = bool(enable)
    _conv_timer["events"] = []


def conv_timer_read():
    """-> (total ms, number of launches) since the last reset."""
    torch.cuda.synchronize()
    ev = _conv_timer["events"]
    return sum(e[0].elapsed_time(e[1]) for e in ev), len(ev)


def conv_timer_table():
    """-> [(description, ms)] per timed launch since the last reset."""
    torch.cuda.synchronize()
    return [(e[2], e[0].elapsed_time(e[1])) for e in _conv_timer["events"]]


def conv_tiles(N, H, W, need_even=False):
    nt = ctypes.c_int(0)
    bh = ctypes.c_int(0)
    bw = ctypes.c_int(0)
    call("egaze_conv3x3_tiles", N, H, W, int(need_even), ctypes.addressof(nt), ctypes.addressof(bh), ctypes.addressof(bw))
    return nt.value, bh.value, bw.value


_stats_shape_cache = {}
_stats_ws = {}


def conv_stats_shape(Cout, precise):
    key = (Cout, bool(precise))
    if key not in _stats_shape_cache:
        a, b, c = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        call("egaze_conv3x3_stats_shape", Cout, int(precise), ctypes.addressof(a), ctypes.addressof(b), ctypes.addressof(c))
        _stats_shape_cache[key] = (a.value, b.value, c.value)
    return _stats_shape_cache[key]


# ---- conv plans: one frozen launch (TMA descriptors + tile configuration) per distinct argument tuple -----------------------
# The caching allocator hands a training loop the same addresses step after step, so after the first step nearly every conv
# call finds its plan: the host then passes 2 arguments through ctypes instead of 28 and skips four cuTensorMapEncodeTiled calls
# and the tile search.  A plan is keyed on the FULL argument tuple (every pointer, shape and flag), so a hit is the same call by
# construction, whatever tensor now lives at those addresses.  EGAZE_CONV_PLANS=0: always the plain entry point.
_plans = {}
_PLAN_CAP = 4096


def _use_plans():
    return os.environ.get("EGAZE_CONV_PLANS", "1") != "0" and not _conv_timer["prof"]


def _conv_plan_run(dev, args):
    key = (dev.index,) + tuple((a.data_ptr() if isinstance(a, torch.Tensor) else a) for a in args)
    plan = _plans.get(key)
    if plan is None:
        for a in args:
            if isinstance(a, torch.Tensor) and (not a.is_contiguous() or a.device != dev):
                raise RuntimeError("egaze: conv3x3 needs contiguous tensors on one device")
        h = ctypes.c_longlong(0)
        call("egaze_conv3x3_plan_create", *(args + (ctypes.addressof(h),)))
        if len(_plans) >= _PLAN_CAP:
            for k in list(_plans)[:_PLAN_CAP // 4]:
                call("egaze_plan_destroy", _plans.pop(k))
        plan = _plans[key] = h.value
    _lib.call_on(dev.index, "egaze_conv3x3_plan_run", plan, stream_ptr())


def conv3x3(act, wpack, bias=None, scale=None, shift=None, relu=False, reduce=0, ups=False, mask=None,
            want_f32=False, want_split=True, stats=False, mask_ups=False, colsum=None, want_lo=True, xb=False, sub=0,
            planar=False):
    """3x3/pad-1 conv on the tcgen05 path.  wpack = (w_hi, w_lo, Cout_p, Cin_p, fmt) from pack_cache, in the activation's format.
    The operand mode follows from the planes present (and the numeric mode): act.lo given -> 3 MMAs per product, act.lo None ->
    act.hi x [w_hi | w_lo] (2), `fast` -> 1.  The split output has the input's format (a forward activation stays fp16, a
    gradient stays bf16); want_lo=False writes the hi plane only (bf16), xb=True adds the bf16 copy (fp16 outputs).
    colsum: optional [Cout] fp32 tensor the kernel ADDS the per-channel sums of the stored values to.
    sub (egaze.h): 1 = sub-pixel forward of "nearest-2x upsample -> this conv" (act is the LOW-resolution map, wpack a mode-2 pack, the
    output is 2H x 2W); 2 = its data gradient (act is the phase-planar output gradient [4N, H, W, C], wpack a mode-3 pack, the output
    the low-resolution gradient).  planar: store the output phase-planar ([4N, H/2, W/2, C]).
    Returns (out_act | None, out_f32 | None, (stats_partial, stats_cnt) | None)."""
    w_hi, w_lo, Cout, Cin_p, wfmt = wpack
    if Cin_p != act.Cp:
        raise RuntimeError("egaze: conv3x3 channel mismatch: activation Cp=%d, weight Cin_p=%d" % (act.Cp, Cin_p))
    fmt = act.fmt
    if wfmt != fmt:
        raise RuntimeError("egaze: conv3x3 operand formats differ (activation %s, weights %s)" % (act.hi.dtype, w_hi.dtype))
    md = mode()
    use_wlo = md["w_lo"]
    x_lo = act.lo if (use_wlo and md["fwd_lo"] and act.lo is not None) else None
    N, H, W = act.N, act.H, act.W
    if sub == 2:
        N //= 4
    dev = act.hi.device
    Ho, Wo = (H // 2, W // 2) if reduce else (H, W)
    if ups or sub == 1:
        Ho, Wo = Ho * 2, Wo * 2
    No = N
    if planar:
        No, Ho, Wo = 4 * N, Ho // 2, Wo // 2
    out_act = empty_act(No, Ho, Wo, Cout, Cout, dev, lo=want_lo or fmt == 1, fmt=fmt, xb=xb and fmt == 1) if want_split else None
    out_f32 = torch.empty((No, Ho, Wo, Cout), dtype=F32, device=dev) if want_f32 else None
    st = None
    if stats:
        # Statistics workspace: persistent per (call shape, stream).  bn_finalize consumes it right after the conv on the same
        # stream, and the kernel rewrites the count of every CTA it launches -- the same CTAs for the same shape -- so the slots
        # of CTAs that never run stay at the zero they were allocated with: no fill launch per layer and step.
        parts, cs, cd = conv_stats_shape(Cout, use_wlo)
        key = (N, H, W, Cin_p, Cout, bool(use_wlo), x_lo is not None, dev, _lib.stream_key(dev))
        ws = _stats_ws.get(key)
        if ws is None:
            ws = _stats_ws[key] = (torch.empty((parts, 2, Cout), dtype=F32, device=dev), torch.zeros((parts, cs), dtype=F32, device=dev))
        st = (ws[0], ws[1], cs, cd)
    if _conv_timer["on"]:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    args = (act.hi, x_lo, w_hi, w_lo if use_wlo else None, N, H, W, Cin_p, Cout,
            bias, scale, shift, int(relu), int(reduce), int(ups), mask, int(mask_ups), out_f32,
            out_act.hi if out_act is not None else None, out_act.lo if out_act is not None else None,
            out_act.xb if out_act is not None else None,
            st[0] if st else None, st[1] if st else None, colsum, fmt, fmt, (1.0 / f16_weight_scale()) if fmt else 1.0,
            int(sub), int(planar))
    if _use_plans():
        _conv_plan_run(dev, args)
    else:
        call("egaze_conv3x3_tc", *(args + (stream_ptr(),)))
    if _conv_timer["on"]:
        ev1.record()
        _conv_timer["events"].append((ev0, ev1, ("conv" if not sub else "conv/sub%d" % sub, N, H, W, Cin_p, Cout, int(reduce), int(ups),
                                                 bool(stats))))
    return out_act, out_f32, st


def bn_finalize(st, C, eps, momentum, gamma, beta, running_mean, running_var, num_batches_tracked=None):
    """-> (mean, invstd, scale, shift); updates running stats (and the int64 num_batches_tracked counter) in place when given."""
    partial, cnt, cnt_stride, cnt_div = st
    dev = partial.device
    mean = torch.empty((C,), dtype=F32, device=dev)
    invstd = torch.empty((C,), dtype=F32, device=dev)
    scale = torch.empty((C,), dtype=F32, device=dev)
    shift = torch.empty((C,), dtype=F32, device=dev)
    call("egaze_bn_finalize", partial, cnt, cnt_stride, cnt_div, partial.shape[0], C, float(eps), float(momentum), gamma, beta, running_mean,
         running_var, mean, invstd, scale, shift, num_batches_tracked, stream_ptr())
    return mean, invstd, scale, shift


def bn_fold(gamma, beta, running_mean, running_var, conv_bias, eps):
    C = running_mean.numel()
    scale = torch.empty((C,), dtype=F32, device=running_mean.device)
    shift = torch.empty((C,), dtype=F32, device=running_mean.device)
    call("egaze_bn_fold", gamma, beta, running_mean, running_var, conv_bias, float(eps), C, scale, shift, stream_ptr())
    return scale, shift


def col_stats(x2d):
    rows, C = x2d.shape
    nb = (rows + 127) // 128
    partial = torch.empty((nb, 2, C), dtype=F32, device=x2d.device)
    cnt = torch.empty((nb,), dtype=F32, device=x2d.device)
    call("egaze_col_stats", x2d, rows, C, partial, cnt, stream_ptr())
    return partial, cnt, 1, C


def bn_apply(x_nhwc, scale, shift, relu=True, pool=False, want_f32=False, want_split=True, xb=False):
    """Normalise (+ReLU, +2x2 max-pool) a raw conv output into the next conv's operand (forward format of the current mode)."""
    N, H, W, C = x_nhwc.shape
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    dev = x_nhwc.device
    fmt = mode()["fwd_fmt"]
    out_act = empty_act(N, Ho, Wo, C, C, dev, lo=True, fmt=fmt, xb=xb and fmt == 1) if want_split else None
    out_f32 = torch.empty((N, Ho, Wo, C), dtype=F32, device=dev) if want_f32 else None
    call("egaze_bn_apply", x_nhwc, N, H, W, C, scale, shift, int(relu), int(pool), out_f32,
         out_act.hi if out_act is not None else None, out_act.lo if out_act is not None else None,
         out_act.xb if out_act is not None else None, fmt, stream_ptr())
    return out_act, out_f32


def pairmax(x_2b):
    """x [2B, H, W, C] fp32 -> [B, H, W, C] elementwise max of the two halves."""
    B2 = x_2b.shape[0]
    out = torch.empty((B2 // 2,) + tuple(x_2b.shape[1:]), dtype=F32, device=x_2b.device)
    call("egaze_pairmax", x_2b, out.numel(), out, stream_ptr())
    return out


def head_fwd(act, w, b, want_logit=False):
    """1x1 conv (C -> 1) + sigmoid.  -> [N, 1, H, W] fp32."""
    N, H, W = act.N, act.H, act.W
    out = torch.empty((N, 1, H, W), dtype=F32, device=act.hi.device)
    logit = torch.empty((N, 1, H, W), dtype=F32, device=act.hi.device) if want_logit else None
    wf = w.detach().reshape(-1).contiguous().float()
    call("egaze_head_fwd", act.hi, act.lo, act.fmt, wf, b.detach() if b is not None else None, wf.numel(), act.Cp, N * H * W,
         out, logit, stream_ptr())
    return (out, logit) if want_logit else out


# ---- backward pieces --------------------------------------------------------------------------------------------
def _pad64(t):
    """A plane with its channel stride padded (zeros) to a multiple of 64 (the weight-gradient GEMM's K-chunk rows)."""
    if t is None or t.shape[-1] % 64 == 0:
        return t
    o = torch.zeros(tuple(t.shape[:-1]) + ((t.shape[-1] + 63) // 64 * 64,), dtype=t.dtype, device=t.device)
    o[..., :t.shape[-1]].copy_(t)
    return o


def wgrad_operands(x_act, dy_act, precise=None):
    """The bf16 planes the weight-gradient GEMM multiplies: (x_hi, x_lo | None, dy_hi, dy_lo | None, precise)."""
    precise = mode()["wgrad_precise"] if precise is None else precise
    if dy_act.fmt != 0:
        raise RuntimeError("egaze: gradients are bf16 planes")
    if x_act.fmt == 1:
        if x_act.xb is None:
            raise RuntimeError("egaze: wgrad of an fp16 forward activation needs its bf16 copy (Act.xb)")
        return x_act.xb, None, dy_act.hi, None, False
    if precise and (x_act.lo is None or dy_act.lo is None):
        precise = False
    return x_act.hi, (x_act.lo if precise else None), dy_act.hi, (dy_act.lo if precise else None), precise


def wgrad3x3(x_act, dy_act, Cout, Cin, precise=None, sub=False):
    """dW (OIHW fp32 [Cout, Cin, 3, 3]) of a 3x3/pad-1 conv from its input activation and output gradient (bf16 planes).
    precise (default: the numeric mode's choice): dY_hi*X_hi + dY_hi*X_lo + dY_lo*X_hi, else dY_hi * bf16(X) (1 MMA).
    sub: the conv follows nn.Upsample and ran in sub-pixel form: x_act is the LOW-resolution input, dy_act the phase-planar
    gradient [4N, H, W, C] of the 2H x 2W output."""
    x_hi, x_lo, dy_hi, dy_lo, precise = [_pad64(t) if torch.is_tensor(t) else t for t in wgrad_operands(x_act, dy_act, precise)]
    N, H, W = x_act.N, x_act.H, x_act.W
    dev = x_act.hi.device
    cin_p, cout_p = x_hi.shape[-1], dy_hi.shape[-1]
    # persistent accumulator per shape: allocated zeroed once, left zeroed again by egaze_unpack_wgrad (clear=1)
    # (keyed by stream too: two streams -- the two trunks of model_SP have identical layer shapes -- must never accumulate
    # into the same buffer concurrently)
    key = (cout_p, cin_p, dev, _lib.stream_key(dev), bool(sub))
    dwp = _dwp_cache.get(key)
    if dwp is None:
        dwp = _dwp_cache[key] = torch.zeros((16 if sub else 9, cout_p, cin_p), dtype=F32, device=dev)
    if _conv_timer["on"]:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    call("egaze_wgrad3x3_tc", x_hi, x_lo, dy_hi, dy_lo, N, H, W, cin_p, cout_p, dwp, int(precise), int(bool(sub)), stream_ptr())
    if _conv_timer["on"]:
        ev1.record()
        _conv_timer["events"].append((ev0, ev1, ("wgrad" if not sub else "wgrad/sub", N, H, W, cin_p, cout_p, 0, 0, False)))
    gw = torch.empty((Cout, Cin, 3, 3), dtype=F32, device=dev)
    call("egaze_unpack_wgrad", dwp, Cout, Cin, cout_p, cin_p, 0.0, 1, int(bool(sub)), gw, stream_ptr())
    return gw


_dwp_cache = {}
_bn_bwd_nblk = [None]
_bn_bwd_ws = {}


def bn_bwd(raw, g, scale, shift, mean, invstd, pool, relu, want_f32=False, want_split=True, batch_stats=True):
    """BatchNorm(+ReLU)(+MaxPool) backward.  raw: conv output NHWC fp32; g: grad w.r.t. the layer output (pooled res).
    batch_stats=False: the forward normalised with running statistics (mean / invstd are those).
    -> (draw Act | None, draw f32 | None, dgamma, dbeta)"""
    N, H, W, C = raw.shape
    dev = raw.device
    if _bn_bwd_nblk[0] is None:
        nb = ctypes.c_int(0)
        call("egaze_bn_bwd_blocks", ctypes.addressof(nb))
        _bn_bwd_nblk[0] = nb.value
    # persistent, self-cleaning accumulators (zero at allocation, re-zeroed by every call): one per (C, device, stream)
    key = (C, dev, _lib.stream_key(dev))
    partial = _bn_bwd_ws.get(key)
    if partial is None:
        partial = _bn_bwd_ws[key] = torch.zeros((_bn_bwd_nblk[0], 2, C), dtype=F32, device=dev)
    dgamma = torch.empty((C,), dtype=F32, device=dev)
    dbeta = torch.empty((C,), dtype=F32, device=dev)
    g = g.contiguous()
    call("egaze_bn_bwd_reduce", raw, g, N, H, W, C, scale, shift, mean, invstd, int(pool), int(relu), partial, dgamma,
         dbeta, stream_ptr())
    out_act = empty_act(N, H, W, C, C, dev, lo=mode()["dy_lo"]) if want_split else None
    out_f32 = torch.empty((N, H, W, C), dtype=F32, device=dev) if want_f32 else None
    call("egaze_bn_bwd_apply", raw, g, N, H, W, C, scale, shift, mean, invstd, dgamma, dbeta, int(pool), int(relu),
         int(batch_stats), out_f32,
         out_act.hi if out_act is not None else None, out_act.lo if out_act is not None else None, stream_ptr())
    return out_act, out_f32, dgamma, dbeta


def pairmax_bwd(raw2, dmx):
    B2 = raw2.shape[0]
    act = empty_act(B2, raw2.shape[1], raw2.shape[2], raw2.shape[3], raw2.shape[3], raw2.device, lo=mode()["dy_lo"])
    call("egaze_pairmax_bwd", raw2, dmx.contiguous(), dmx.numel(), act.hi, act.lo, stream_ptr())
    return act


def col_sum(act, C=None):
    """sum over pixels of a split activation -> [C] fp32 (conv bias gradient)."""
    out = torch.zeros((act.Cp,), dtype=F32, device=act.hi.device)
    call("egaze_col_sum_split", act.hi, act.lo, act.N * act.H * act.W, act.Cp, out, stream_ptr())
    return out if C is None or C == act.Cp else out[:C].contiguous()


def head_bwd(act, w, y, gy, relu_mask=True, zeros=None):
    """Backward of sigmoid(conv1x1(act)).  -> (dx Act, dw [C], db [1]).  zeros(n, device): optional source of zero-initialised
    fp32 vectors (the backward's pool, egaze.autograd._GradBag)."""
    dev = act.hi.device
    dx = empty_act(act.N, act.H, act.W, act.Cp, act.C, dev, lo=mode()["dy_lo"])
    wf = w.detach().reshape(-1).contiguous().float()
    dw = zeros(wf.numel(), dev) if zeros is not None else torch.zeros((wf.numel(),), dtype=F32, device=dev)
    db = zeros(1, dev) if zeros is not None else torch.zeros((1,), dtype=F32, device=dev)
    call("egaze_head_bwd", act.hi, act.lo, act.fmt, wf, wf.numel(), act.Cp, act.N * act.H * act.W, y.contiguous(),
         gy.contiguous().float(), int(relu_mask), dx.hi, dx.lo, dw, db, stream_ptr())
    return dx, dw, db


# ---- late fusion (models/late_fusion.py) ----------------------------------------------------------------------------
_lf_scratch = {}


def _lf_scratch_buf(dev, which):
    """Per (device, stream) scratch for the LF kernels' per-CTA partials (sizes from egaze_lf_scratch)."""
    key = (dev, _lib.stream_key(dev), which)
    buf = _lf_scratch.get(key)
    if buf is None:
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        call("egaze_lf_scratch", ctypes.addressof(a), ctypes.addressof(b))
        buf = _lf_scratch[key] = torch.empty((a.value if which == 0 else b.value,), dtype=F32, device=dev)
    return buf


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def lf_parts(seq):
    """The parameter containers of late_fusion.fusion (models/late_fusion.py:10-14): 3 x (Conv2d 3x3, BatchNorm2d) + Conv2d 1x1."""
    mods = list(seq.children())
    convs = [m for m in mods if isinstance(m, torch.nn.Conv2d)]
    bns = [m for m in mods if isinstance(m, torch.nn.BatchNorm2d)]
    shapes = [tuple(c.weight.shape) for c in convs]
    if shapes != [(32, 2, 3, 3), (32, 32, 3, 3), (8, 32, 3, 3), (1, 8, 1, 1)] or len(bns) != 3:
        raise RuntimeError("egaze: late_fusion kernels are specialised for the reference's 2->32->32->8->1 net, got %r" % (shapes,))
    if any(c.padding != ((1, 1) if c.kernel_size == (3, 3) else (0, 0)) or c.stride != (1, 1) for c in convs):
        raise RuntimeError("egaze: late_fusion convs must be stride 1, 'same' padding")
    if len({(bn.eps, bn.momentum) for bn in bns}) != 1 or bns[0].momentum is None or any(not bn.affine for bn in bns):
        raise RuntimeError("egaze: late_fusion BatchNorm layers must share eps / momentum and be affine")
    return convs, bns


def lf_forward(seq, f, g, keep=False, precise=None):
    """late_fusion.forward (models/late_fusion.py:18-23) in one C-ABI call.  -> (out [B,1,H,W], saved | None)."""
    convs, bns = lf_parts(seq)
    precise = is_precise() if precise is None else precise
    f = f.detach().contiguous().float()
    g = g.detach().contiguous().float()
    if f.shape != g.shape or f.dim() != 4 or f.shape[1] != 1:
        raise RuntimeError("egaze: late_fusion expects two (B,1,H,W) maps, got %s and %s" % (tuple(f.shape), tuple(g.shape)))
    B, _, H, W = f.shape
    dev = f.device
    batch_stats = [bn.training or (bn.running_mean is None and bn.running_var is None) for bn in bns]
    if len(set(batch_stats)) != 1:
        raise RuntimeError("egaze: late_fusion BatchNorm layers must all be in the same mode")
    training = batch_stats[0]
    raw1 = torch.empty((B, H, W, 32), dtype=F32, device=dev)
    raw2 = torch.empty((B, H, W, 32), dtype=F32, device=dev)
    raw3 = torch.empty((B, H, W, 8), dtype=F32, device=dev) if (training or keep) else None
    bn_ws = torch.empty((3, 4, 32), dtype=F32, device=dev)
    out = torch.empty((B, 1, H, W), dtype=F32, device=dev)
    ws = [c.weight.detach() for c in convs]
    bs = [c.bias.detach() if c.bias is not None else None for c in convs]
    gam = [bn.weight.detach() for bn in bns]
    bet = [bn.bias.detach() for bn in bns]
    rms = [bn.running_mean if bn.track_running_stats else None for bn in bns]
    rvs = [bn.running_var if bn.track_running_stats else None for bn in bns]
    for t in ws + gam + bet + [x for x in bs + rms + rvs if x is not None]:
        if not (t.is_cuda and t.is_contiguous() and t.dtype == F32):
            raise RuntimeError("egaze: late_fusion parameters must be contiguous fp32 CUDA tensors")
    call("egaze_lf_fwd", f, g, B, H, W, _ptr_array(ws), _ptr_array(bs), _ptr_array(gam), _ptr_array(bet), _ptr_array(rms),
         _ptr_array(rvs), int(training), float(bns[0].eps), float(bns[0].momentum), raw1, raw2, raw3, bn_ws,
         _lf_scratch_buf(dev, 0), out, int(precise), stream_ptr())
    if training:
        for bn in bns:
            if bn.track_running_stats and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
    saved = (f, g, raw1, raw2, raw3, bn_ws, out, training, precise) if keep else None
    return out, saved


def lf_backward(seq, saved, gout, need_w=(True, True, True, True), need_f=False, need_g=False):
    """Backward of lf_forward.  -> dict(dw=[4], dbh, dgamma=[3], dbeta=[3], gf, gg)."""
    convs, bns = lf_parts(seq)
    f, g, raw1, raw2, raw3, bn_ws, out, training, precise = saved
    B, _, H, W = f.shape
    dev = f.device
    new = lambda *shape: torch.empty(shape, dtype=F32, device=dev)
    g1, g2 = new(B, H, W, 32), new(B, H, W, 32)
    dw = [torch.empty_like(c.weight) if need else None for c, need in zip(convs, need_w)]
    dbh = new(1)
    dgamma, dbeta = [new(32), new(32), new(8)], [new(32), new(32), new(8)]
    gf = new(B, 1, H, W) if need_f else None
    gg = new(B, 1, H, W) if need_g else None
    ws = [c.weight.detach() for c in convs]
    call("egaze_lf_bwd", f, g, B, H, W, _ptr_array(ws), raw1, raw2, raw3, bn_ws, int(training), out,
         gout.detach().contiguous().float(), g1, g2, _lf_scratch_buf(dev, 1), _ptr_array(dw), dbh, _ptr_array(dgamma),
         _ptr_array(dbeta), gf, gg, int(precise), stream_ptr())
    return dict(dw=dw, dbh=dbh, dgamma=dgamma, dbeta=dbeta, gf=gf, gg=gg)


# ---- floss -------------------------------------------------------------------------------------------------------
def floss_centroid(target):
    B, H, W = target.shape[0], target.shape[-2], target.shape[-1]
    cen = torch.empty((B, 2), dtype=torch.float64, device=target.device)
    call("egaze_floss_centroid", target, B, H, W, cen, stream_ptr())
    return cen


def floss_weight(target):
    target = target.contiguous().float()
    B, H, W = target.shape[0], target.shape[-2], target.shape[-1]
    cen = floss_centroid(target)
    w = torch.empty_like(target)
    call("egaze_floss_weight", cen, B, H, W, w, stream_ptr())
    return w


def floss_fwd(inp, target, cen):
    B, H, W = target.shape[0], target.shape[-2], target.shape[-1]
    acc = torch.empty((1,), dtype=torch.float64, device=inp.device)
    loss = torch.empty((), dtype=F32, device=inp.device)
    call("egaze_floss_fwd", inp, target, cen, B, H, W, acc, loss, stream_ptr())
    return loss


def floss_bwd(inp, target, cen, grad_loss):
    B, H, W = target.shape[0], target.shape[-2], target.shape[-1]
    g = torch.empty_like(inp)
    call("egaze_floss_bwd", inp, target, cen, B, H, W, grad_loss.contiguous().float(), g, stream_ptr())
    return g


# ---- AT glue -----------------------------------------------------------------------------------------------------
def crop_mean(feat, gaze, size=3, down=16):
    feat = feat.contiguous().float()
    B, C, H, W = feat.shape
    g = torch.as_tensor(gaze, dtype=torch.int32, device=feat.device).reshape(B, 2).contiguous()
    out = torch.empty((B, C), dtype=F32, device=feat.device)
    call("egaze_crop_mean", feat, g, B, C, H, W, int(size), int(down), out, stream_ptr())
    return out


def crop_align_mean(feat, gaze, size=3, up=16):
    """AT.crop_align_feature + mean (align=True branch of AT.extract_late): -> [B, C]."""
    feat = feat.contiguous().float()
    B, C, H, W = feat.shape
    g = torch.as_tensor(gaze, dtype=torch.int32, device=feat.device).reshape(B, 2).contiguous()
    out = torch.empty((B, C), dtype=F32, device=feat.device)
    call("egaze_crop_align_mean", feat, g, B, C, H, W, int(size), int(up), out, stream_ptr())
    return out


def weighted_map(chn_weight, feat):
    feat = feat.contiguous().float()
    B, C, H, W = feat.shape
    w = chn_weight.reshape(B, C).contiguous().float()
    out = torch.empty((B, H, W), dtype=F32, device=feat.device)
    call("egaze_weighted_map", feat, w, B, C, H * W, out, stream_ptr())
    return out


def bilinear_up(x, scale=16, align_corners=False):
    shp = x.shape
    h, w = shp[-2], shp[-1]
    xb = x.contiguous().float().reshape(-1, h, w)
    out = torch.empty((xb.shape[0], h * scale, w * scale), dtype=F32, device=x.device)
    call("egaze_bilinear_up", xb, xb.shape[0], h, w, int(scale), int(align_corners), out, stream_ptr())
    return out.reshape(tuple(shp[:-2]) + (h * scale, w * scale))


# ---- validation metric ---------------------------------------------------------------------------------------------
def aae_auc(output, target):
    """utils.computeAAEAUC on the device: output / target [B,(1,)224,224] CUDA tensors -> [B,4] fp64 tensor of
    (AAE in degrees, AUC, gaze row, gaze column) per sample.  Only this small tensor needs to travel to the host."""
    _lib.check_device(output.device)
    o = output.detach().contiguous().float().reshape(-1, output.shape[-2], output.shape[-1])
    t = target.detach().contiguous().float().reshape(-1, target.shape[-2], target.shape[-1])
    if o.shape != t.shape:
        raise RuntimeError("egaze: aae_auc shape mismatch %s vs %s" % (tuple(o.shape), tuple(t.shape)))
    res = torch.empty((o.shape[0], 4), dtype=torch.float64, device=o.device)
    call("egaze_aae_auc", o, t, o.shape[0], o.shape[1], o.shape[2], _gauss14(o.device), res, stream_ptr())
    return res


_gauss14_cache = {}


def _gauss14(device):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma=14, order=0, radius=int(4*14+0.5)), restated with NumPy so that the
    device multiplies bit-identical weights (utils.py:117 calls gaussian_filter(z, 14))."""
    w = _gauss14_cache.get(device)
    if w is None:
        import numpy as np
        sigma, radius = 14.0, 56
        x = np.arange(-radius, radius + 1)
        phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
        w = _gauss14_cache[device] = torch.from_numpy(phi / phi.sum()).to(device)
    return w


# ---- LSTM ----------------------------------------------------------------------------------------------------------
def lstm_seq_fwd(x, h0, c0, lstm, lin, save_gates=False):
    """x [T,B,512] raw input (tanh applied inside).  lstm: nn.LSTM(512,512,2) parameter container, lin: nn.Linear."""
    x = x.contiguous().float()
    T, B, Hd = x.shape
    dev = x.device
    out = torch.empty((T, B, Hd), dtype=F32, device=dev)
    hn = torch.empty((2, B, Hd), dtype=F32, device=dev)
    cn = torch.empty((2, B, Hd), dtype=F32, device=dev)
    ws_h = torch.empty((T, 2, B, Hd), dtype=F32, device=dev)
    ws_c = torch.empty((T, 2, B, Hd), dtype=F32, device=dev)
    ws_g = torch.empty((T, 2, B, 4, Hd), dtype=F32, device=dev) if save_gates else None

    def ptrs(fmt):
        ts = [getattr(lstm, fmt % l).detach().contiguous() for l in range(2)]
        arr = (ctypes.c_void_p * 2)(*[t.data_ptr() for t in ts])
        return ts, arr

    k1, w_ih = ptrs("weight_ih_l%d")
    k2, w_hh = ptrs("weight_hh_l%d")
    k3, b_ih = ptrs("bias_ih_l%d")
    k4, b_hh = ptrs("bias_hh_l%d")
    call("egaze_lstm_seq_fwd", x, h0.contiguous().float(), c0.contiguous().float(), w_ih, w_hh, b_ih, b_hh,
         lin.weight.detach().contiguous(), lin.bias.detach().contiguous(), T, B, out, hn, cn, ws_h, ws_c, ws_g,
         stream_ptr())
    del k1, k2, k3, k4
    return out, hn, cn, (ws_h, ws_c, ws_g)


def lstm_seq_bwd(x, h0, c0, lstm, lin, out, ws, gout, ghn, gcn, need_input, need_state):
    """BPTT of lstm_seq_fwd.  ws = (ws_h, ws_c, ws_gates) from the forward (save_gates=True).
    -> dict with dw_ih, dw_hh, db (lists of 2), dlin_w, dlin_b, dinput | None, dh0 | None, dc0 | None."""
    x = x.contiguous().float()
    T, B, Hd = x.shape
    dev = x.device
    ws_h, ws_c, ws_g = ws
    new = lambda *shape: torch.empty(shape, dtype=F32, device=dev)
    xt, dz, dh_top = new(T, B, Hd), new(T, B, Hd), new(T, B, Hd)
    dgates = new(2, T, B, 4 * Hd)
    tmp_x, dh_next, dc_next = new(B, Hd), new(2, B, Hd), new(2, B, Hd)
    dx0 = new(T, B, Hd) if need_input else None
    dinput = new(T, B, Hd) if need_input else None
    dh0 = new(2, B, Hd) if need_state else None
    dc0 = new(2, B, Hd) if need_state else None
    dw_ih = [new(4 * Hd, Hd) for _ in range(2)]
    dw_hh = [new(4 * Hd, Hd) for _ in range(2)]
    db = [new(4 * Hd) for _ in range(2)]
    dlin_w, dlin_b = new(Hd, Hd), new(Hd)
    parr = lambda ts: (ctypes.c_void_p * 2)(*[t.data_ptr() for t in ts])
    keep = [getattr(lstm, "weight_ih_l%d" % l).detach().contiguous() for l in range(2)]
    keep2 = [getattr(lstm, "weight_hh_l%d" % l).detach().contiguous() for l in range(2)]
    call("egaze_lstm_seq_bwd", x, h0.contiguous().float(), c0.contiguous().float(), parr(keep), parr(keep2),
         lin.weight.detach().contiguous(), T, B, out, gout.contiguous().float(),
         ghn.contiguous().float() if ghn is not None else None, gcn.contiguous().float() if gcn is not None else None,
         ws_h, ws_c, ws_g, xt, dz, dgates, dh_top, tmp_x, dh_next, dc_next, dx0, parr(dw_ih), parr(dw_hh), parr(db),
         dlin_w, dlin_b, dinput, dh0, dc0, stream_ptr())
    return dict(dw_ih=dw_ih, dw_hh=dw_hh, db=db, dlin_w=dlin_w, dlin_b=dlin_b, dinput=dinput, dh0=dh0, dc0=dc0)
