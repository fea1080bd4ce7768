"""Autograd bridges: torch.autograd.Function wrappers whose forward/backward run entirely on the egaze kernels.

One Function per drop-in module (model_SP, single-stream VGG, TrunkSequential, late_fusion).  Activations needed by the
backward (split-bf16 conv inputs for wgrad, raw fp32 conv outputs + batch statistics for BatchNorm backward) are kept
in a Python-side record on ctx; gradients flow between layers as NHWC tensors and never leave the device.

Backward schedule per layer (reverse order):
  conv+BN+ReLU(+pool): bn_bwd (reduce -> apply; ReLU mask + max-pool routing fused) -> d(raw) split
                       -> wgrad3x3_tc (dW) + col_sum (db) + conv3x3_tc with flipped weights (dgrad, fp32 out)
  conv+ReLU(+ups)    : dgrad of the NEXT layer already applied this layer's ReLU mask (and the 2x2 sum that is the
                       gradient of nn.Upsample) in its epilogue -> wgrad + col_sum + dgrad
"""
import os

import torch

from . import engine, ops


def _req(p):
    return p is not None and p.requires_grad


# ---- weight gradients on a side stream ------------------------------------------------------------------------------
# dW of a layer feeds nothing downstream in the backward, so its tcgen05 GEMM is enqueued on a second stream: it overlaps
# the HBM-bound BatchNorm-backward passes of the next layers (different SM resources: tensor pipe vs DRAM), which the
# main stream would otherwise run with the tensor pipe idle.  EGAZE_WGRAD_STREAM=0 keeps everything on one stream.
_side_streams = {}


def _use_side_stream():
    return os.environ.get("EGAZE_WGRAD_STREAM", "1") != "0"


def _get_side_stream(dev):
    side = _side_streams.get(dev)
    if side is None:
        side = _side_streams[dev] = torch.cuda.Stream(device=dev)
    return side


def _wgrad(x_act, dy_act, cout, cin, sub=False):
    wg = (lambda x, dy, co, ci: ops.wgrad3x3(x, dy, co, ci, sub=True)) if sub else ops.wgrad3x3
    return _wgrad_impl(wg, x_act, dy_act, cout, cin)


def _wgrad_impl(wg, x_act, dy_act, cout, cin):
    if not _use_side_stream():
        return wg(x_act, dy_act, cout, cin)
    dev = x_act.hi.device
    side = _get_side_stream(dev)
    main = torch.cuda.current_stream(dev)
    side.wait_stream(main)                      # the operands were produced on the main stream
    with torch.cuda.stream(side):
        gw = wg(x_act, dy_act, cout, cin)
    for t in (x_act.hi, x_act.lo, x_act.xb, dy_act.hi, dy_act.lo):
        if t is not None:
            t.record_stream(side)               # the caching allocator must not recycle them while the side stream reads
    gw.record_stream(main)
    return gw


def _join_side_stream():
    """Called once at the end of a backward: the returned gradients are consumed on the main stream."""
    for dev, side in _side_streams.items():
        torch.cuda.current_stream(dev).wait_stream(side)


class _GradBag(object):
    """Collects parameter gradients by identity.  Also hands out the zero-initialised fp32 vectors a backward needs (exactly-zero
    bias gradients, the accumulators the dgrad epilogues add column sums to) as 16-byte aligned slices of ONE zero-filled buffer:
    one fill launch per backward instead of one per layer (~45 in a model_SP step)."""
    POOL = 32768

    def __init__(self):
        self.d = {}
        self._pool = None
        self._used = 0

    def reserve(self, device):
        """Zero the pool NOW, on the current stream -- call before the backward forks its side streams."""
        if self._pool is None or self._pool.device != torch.device(device):
            self._pool = torch.zeros((self.POOL,), dtype=torch.float32, device=device)
            self._used = 0
        return self

    def zeros(self, n, device):
        n4 = (int(n) + 3) // 4 * 4
        if self._pool is None or self._pool.device != torch.device(device) or self._used + n4 > self.POOL:
            return torch.zeros((int(n),), dtype=torch.float32, device=device)
        v = self._pool[self._used:self._used + int(n)]
        self._used += n4
        return v

    def zeros_like(self, p):
        if p.dtype != torch.float32:
            return torch.zeros_like(p)
        return self.zeros(p.numel(), p.device).view(p.shape)

    def put(self, p, g):
        if p is not None and p.requires_grad:
            self.d[id(p)] = g.reshape(p.shape)

    def get(self, p):
        return self.d.get(id(p))


def _conv_param_grads(bag, conv, x_act, gpre_act, bias_grad_is_zero=False, bias_grad=None, sub=False):
    """dW via the tcgen05 wgrad kernel, db via a column sum.  gpre_act: gradient w.r.t. the conv output.
    A conv bias that feeds a training-mode BatchNorm has an exactly-zero gradient (the batch mean removes it; stock
    PyTorch returns rounding noise ~1e-9 there), so no reduction pass is spent on it.  bias_grad: column sums already
    accumulated by the dgrad epilogue that produced gpre_act (then no pass over gpre_act is needed either)."""
    if _req(conv.weight):
        gw = _wgrad(x_act, gpre_act, conv.out_channels, conv.in_channels, sub=sub)
        bag.put(conv.weight, gw)
    if _req(conv.bias):
        if bias_grad_is_zero:
            bag.put(conv.bias, bag.zeros_like(conv.bias))
        elif bias_grad is not None:
            bag.put(conv.bias, bias_grad[:conv.out_channels] if bias_grad.numel() != conv.out_channels else bias_grad)
        else:
            bag.put(conv.bias, ops.col_sum(gpre_act, conv.out_channels))


def _dgrad(conv, gpre_act, sub=False, **kw):
    """Data gradient of a 3x3 conv: the same tcgen05 kernel with flipped / transposed weights.  sub: the conv ran in sub-pixel form
    (gpre_act is phase-planar; the result is the gradient w.r.t. the LOW-resolution map in front of the nn.Upsample)."""
    cin = conv.in_channels
    rows_p = ops.pad_channels(cin) if cin % 16 else cin
    wpack = ops.pack_cache.get(conv.weight, 3 if sub else 1, rows_p=rows_p, cols_p=gpre_act.Cp)
    return ops.conv3x3(gpre_act, wpack, want_lo=ops.mode()["dy_lo"], sub=2 if sub else 0, **kw)


def _first_cp(conv):
    """Channel padding of a trunk's input.  When the first conv's weight is trained the input is laid out with the 64-channel
    stride its weight-gradient GEMM needs, instead of being re-padded (two extra copies per plane) in the backward."""
    if _req(conv.weight) and conv.in_channels <= 64 and ops.mode()["fwd_fmt"] == 0:
        return 64       # (fp16 forward: the weight gradient reads the separate bf16 copy, which to_split pads to 64 on its own)
    return ops.pad_channels(conv.in_channels)


def _any_req(mods):
    return any(p.requires_grad for m in mods for p in m.parameters())


from .ddp import DEEP_FROM as _DEEP_FROM   # trunk layer index of conv3_1 (layers [4:] hold 98 % of a trunk's parameters)


def _bn_chain_steps(specs, saved, g, bag, need_input_grad, result, at=None):
    """Generator: one conv+BN+ReLU(+pool) layer of the backward chain per step (so that two independent chains -- the two
    trunks of model_SP -- can be enqueued alternately on two streams).  result[0] receives the gradient w.r.t. the chain
    input (or None)."""
    first_needed = None
    for i, sp in enumerate(specs):
        if _any_req([sp.conv, sp.bn]):
            first_needed = i
            break
    if first_needed is None and not need_input_grad:
        result[0] = None
        return
    stop = 0 if need_input_grad else first_needed
    for i in range(len(specs) - 1, stop - 1, -1):
        sp, rec = specs[i], saved[i]
        evalbn = bool(rec.get("eval"))
        draw, _, dgamma, dbeta = ops.bn_bwd(rec["raw"], g, rec["scale"], rec["shift"], rec["mean"], rec["invstd"],
                                            pool=sp.pool, relu=sp.relu, batch_stats=not evalbn)
        C = sp.conv.out_channels
        bag.put(sp.bn.weight, dgamma[:C])
        bag.put(sp.bn.bias, dbeta[:C])
        # running-statistics BatchNorm: d(raw) = scale * gz, so the conv bias gradient is scale * sum(gz) = scale * dbeta
        _conv_param_grads(bag, sp.conv, rec["x"], draw, bias_grad_is_zero=not evalbn,
                          bias_grad=(rec["scale"] * dbeta) if evalbn else None)
        if i > stop or (i == 0 and need_input_grad):
            _, g, _ = _dgrad(sp.conv, draw, want_f32=True, want_split=False)
        else:
            g = None
        result[0] = g
        yield
        if at is not None and at[1] is not None and i == at[0]:
            # Called after the step that processed layer at[0] has been enqueued.  When two chains alternate (_alternate drives
            # this generator first), the other chain's same layer is enqueued right after this yield returns control, i.e. before
            # this generator resumes here -- so both chains' layers >= at[0] are enqueued when the callback runs.
            at[1]()


def bn_sequential_backward(specs, saved, g, bag, need_input_grad, at=None):
    """Backward through a conv+BN+ReLU(+pool) chain.  g: NHWC fp32 gradient w.r.t. the chain output.
    Returns the NHWC fp32 gradient w.r.t. the chain input (or None)."""
    result = [None]
    for _ in _bn_chain_steps(specs, saved, g, bag, need_input_grad, result, at=at):
        pass
    return result[0]


# ---- the two trunks of model_SP on two streams --------------------------------------------------------------------------
# features_s and features_t are independent until the fusion layer.  Enqueued alternately on two streams, the HBM-bound
# BatchNorm passes of one trunk run under the tensor-bound convolutions of the other.  EGAZE_TRUNK_STREAM=0 disables it.
_trunk_streams = {}


def _trunk_stream(dev):
    if os.environ.get("EGAZE_TRUNK_STREAM", "1") == "0":
        return None
    st = _trunk_streams.get(dev)
    if st is None:
        st = _trunk_streams[dev] = torch.cuda.Stream(device=dev)
    return st


def _alternate(gen_a, gen_b, stream_b):
    """Drive two generators to exhaustion, one step each in turn; gen_b runs with stream_b current."""
    live_a, live_b = True, True
    while live_a or live_b:
        if live_a:
            live_a = next(gen_a, _DONE) is not _DONE
        if live_b:
            with torch.cuda.stream(stream_b):
                live_b = next(gen_b, _DONE) is not _DONE


_DONE = object()


def relu_sequential_backward(specs, saved, gpre, bag, want_input_grad_f32):
    """Backward through a conv+bias+ReLU(+ups) chain.  gpre: split gradient w.r.t. the LAST conv's output
    (its ReLU mask already applied).  Returns the NHWC fp32 gradient w.r.t. the chain input (or None).
    The dgrad of layer i writes the masked gradient w.r.t. layer i-1's output; its epilogue also accumulates that
    tensor's column sums, which are layer i-1's bias gradient.
    A conv in sub-pixel form (engine.ConvSpec.sub) reads its output gradient phase-planar -- the dgrad of the layer after it stores
    it that way -- and its own dgrad lands directly on the low-resolution map (no 2x2 sum pass)."""
    bias_grad = None
    for i in range(len(specs) - 1, -1, -1):
        sp, rec = specs[i], saved[i]
        _conv_param_grads(bag, sp.conv, rec["x"], gpre, bias_grad=bias_grad, sub=sp.sub)
        bias_grad = None
        if i > 0:
            prev = specs[i - 1]
            if _req(prev.conv.bias):
                cin = sp.conv.in_channels   # == prev.conv.out_channels; the dgrad output carries them padded to 16
                bias_grad = bag.zeros(ops.pad_channels(cin) if cin % 16 else cin, gpre.hi.device)
            ups = prev.ups and not prev.ups_folded      # the upsampled map exists: 2x2 sum + upsampled mask in the epilogue
            gpre, _, _ = _dgrad(sp.conv, gpre, sub=sp.sub, reduce=2 if ups else 0, mask=saved[i - 1]["y"].hi, mask_ups=ups,
                                colsum=bias_grad, planar=prev.sub)
        elif want_input_grad_f32:
            _, g, _ = _dgrad(sp.conv, gpre, want_f32=True, want_split=False)
            return g
    return None


def _params(module):
    return list(module.parameters())


def _ret_grads(bag, params):
    """Gradients in parameter order.  Some were produced on the weight-gradient / trunk streams: tell the caching allocator
    that the main stream (optimiser, all-reduce) uses them too."""
    out = []
    for p in params:
        g = bag.get(p)
        if g is not None and g.is_cuda:
            g.record_stream(torch.cuda.current_stream(g.device))
        out.append(g)
    return tuple(out)


# ---------------------------------------------------------------------------------------------------------------------
# model_SP
# ---------------------------------------------------------------------------------------------------------------------
class _ModelSPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x_s, x_t, *params):
        ops.pack_cache.refresh()   # one launch re-packs every weight copy the optimiser step made stale
        saved_s, saved_t, saved_tail = [], [], []
        specs_s, _ = engine.parse_sequential(model.features_s)
        specs_t, _ = engine.parse_sequential(model.features_t)
        ts = _trunk_stream(x_s.device)
        if ts is None:
            fus_w = model.fusion.weight.requires_grad
            a_s = ops.to_split(x_s, _first_cp(specs_s[0].conv), xb=_req(specs_s[0].conv.weight))
            a_t = ops.to_split(x_t, _first_cp(specs_t[0].conv), xb=_req(specs_t[0].conv.weight))
            for sp, xb in zip(specs_s, engine.xb_flags(specs_s, fus_w)):
                a_s = engine.run_conv_spec(a_s, sp, saved_s, xb)
            for sp, xb in zip(specs_t, engine.xb_flags(specs_t, fus_w)):
                a_t = engine.run_conv_spec(a_t, sp, saved_t, xb)
        else:
            main = torch.cuda.current_stream(x_s.device)
            ts.wait_stream(main)                       # inputs and re-packed weights were produced on the main stream
            box_s, box_t = [None], [None]

            def trunk(x, specs, saved, box):
                box[0] = ops.to_split(x, _first_cp(specs[0].conv), xb=_req(specs[0].conv.weight))
                yield
                for sp, xb in zip(specs, engine.xb_flags(specs, model.fusion.weight.requires_grad)):
                    box[0] = engine.run_conv_spec(box[0], sp, saved, xb)
                    yield

            _alternate(trunk(x_s, specs_s, saved_s, box_s), trunk(x_t, specs_t, saved_t, box_t), ts)
            a_s, a_t = box_s[0], box_t[0]
            x_t.record_stream(ts)
            main.wait_stream(ts)                       # the fusion layer reads both trunks
            for t in (a_t.hi, a_t.lo, a_t.xb):
                if t is not None:
                    t.record_stream(main)
        # forward hooks registered on the trunks (AT.py:105 idiom) still fire, with detached NCHW views
        for mod, act in ((model.features_s, a_s), (model.features_t, a_t)):
            if mod._forward_hooks:
                y = ops.from_split(act)
                for hook in mod._forward_hooks.values():
                    hook(mod, (None,), y)
        out = engine.run_sp_tail(model, a_s, a_t, saved_tail)
        ctx.model, ctx.rec = model, (specs_s, specs_t, saved_s, saved_t, saved_tail[0])
        ctx.need_x = (x_s.requires_grad, x_t.requires_grad)
        return out

    @staticmethod
    def backward(ctx, gout):
        model = ctx.model
        specs_s, specs_t, saved_s, saved_t, tail = ctx.rec
        bag = _GradBag().reserve(gout.device)
        dspecs, head = engine.parse_sequential(model.decoder)
        dsaved = tail["decoder"]
        # 1x1 conv + sigmoid backward; its dx already carries the last decoder ReLU's mask
        gpre, dw, db = ops.head_bwd(tail["head_in"], head.weight, tail["out"], gout, relu_mask=True, zeros=bag.zeros)
        bag.put(head.weight, dw)
        bag.put(head.bias, db)
        trunk_need = _any_req([model.features_s, model.features_t]) or any(ctx.need_x)
        upstream_need = trunk_need or _any_req([model.fusion, model.bn])
        g = relu_sequential_backward(dspecs, dsaved, gpre, bag, upstream_need)
        # data-parallel gradient averaging overlapped with the rest of the backward (egaze.ddp.OverlappedGradReducer): each
        # segment is reduced on a communication stream as soon as its weight gradients are enqueued
        reducer = getattr(model, "_egaze_reducer", None)
        side = _get_side_stream(gout.device) if _use_side_stream() else None
        if reducer is not None:
            reducer.begin()
            reducer.reduce(bag, "decoder", (side,))
        if upstream_need:
            bn = model.bn
            evalbn = bool(tail.get("eval"))
            _, dmx, dgamma, dbeta = ops.bn_bwd(tail["mx"], g, tail["scale"], tail["shift"], tail["mean"], tail["invstd"],
                                               pool=False, relu=True, want_f32=True, want_split=False, batch_stats=not evalbn)
            bag.put(bn.weight, dgamma)
            bag.put(bn.bias, dbeta)
            d2 = ops.pairmax_bwd(tail["raw2"], dmx)  # [2B,14,14,512] split: gradient routed to the arg-max stream
            fus = model.fusion
            if _req(fus.weight):
                bag.put(fus.weight, _wgrad(tail["cat"], d2, fus.out_channels, fus.in_channels))
            if _req(fus.bias):
                if evalbn:   # the shared conv's bias reaches the output through whichever stream won the max: sum of d(mx)
                    bag.put(fus.bias, tail["scale"] * dbeta)
                else:
                    bag.put(fus.bias, bag.zeros_like(fus.bias))  # feeds model_SP.bn in batch-stat mode: exactly zero
            if reducer is not None:
                reducer.reduce(bag, "fusion_bn", (side,))
            if trunk_need:
                wpack = ops.pack_cache.get(fus.weight, 1, cols_p=d2.Cp)
                _, gf, _ = ops.conv3x3(d2, wpack, want_f32=True, want_split=False)
                B = gf.shape[0] // 2
                g_s, g_t = gf[:B], gf[B:]
                ts = _trunk_stream(gf.device)
                # the deep trunk layers (conv3_1 and up: 98 % of the trunk parameters) are final long before the backward ends
                def deep_done():
                    if reducer is not None:
                        reducer.reduce(bag, "trunk_deep", (side, ts))
                if ts is None:
                    gx_s = bn_sequential_backward(specs_s, saved_s, g_s, bag, ctx.need_x[0], at=(_DEEP_FROM, None))
                    gx_t = bn_sequential_backward(specs_t, saved_t, g_t, bag, ctx.need_x[1], at=(_DEEP_FROM, deep_done))
                else:
                    main = torch.cuda.current_stream(gf.device)
                    ts.wait_stream(main)               # g_t comes from the fusion dgrad on the main stream
                    gf.record_stream(ts)
                    res_s, res_t = [None], [None]
                    _alternate(_bn_chain_steps(specs_s, saved_s, g_s, bag, ctx.need_x[0], res_s, at=(_DEEP_FROM, deep_done)),
                               _bn_chain_steps(specs_t, saved_t, g_t, bag, ctx.need_x[1], res_t), ts)
                    main.wait_stream(ts)
                    gx_s, gx_t = res_s[0], res_t[0]
                    if gx_t is not None:
                        gx_t.record_stream(main)
            else:
                gx_s = gx_t = None
        else:
            gx_s = gx_t = None
        gx_s = ops.nhwc_f32_to_nchw(gx_s, specs_s[0].conv.in_channels) if (gx_s is not None and ctx.need_x[0]) else None
        gx_t = ops.nhwc_f32_to_nchw(gx_t, specs_t[0].conv.in_channels) if (gx_t is not None and ctx.need_x[1]) else None
        ctx.rec = None
        _join_side_stream()
        if reducer is not None:
            reducer.reduce_rest(bag)     # the shallow trunk layers, and whatever a pruned backward did not reach
            reducer.finish()
        return (None, gx_s, gx_t) + _ret_grads(bag, _params(model))


def model_sp_with_grad(model, x_s, x_t):
    return _ModelSPFn.apply(model, x_s, x_t, *_params(model))


# ---------------------------------------------------------------------------------------------------------------------
# TrunkSequential used on its own (e.g. the trainable flow trunk of temporalstream.py)
# ---------------------------------------------------------------------------------------------------------------------
class _SequentialFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, seq, x, *params):
        ops.pack_cache.refresh()   # one launch re-packs every weight copy the optimiser step made stale
        specs, tail = engine.parse_sequential(seq)
        if tail is not None or any(sp.bn is None for sp in specs):
            raise NotImplementedError("egaze: stand-alone training is implemented for conv+BN+ReLU trunks")
        saved = []
        act = ops.to_split(x, _first_cp(specs[0].conv), xb=_req(specs[0].conv.weight))
        for sp, xb in zip(specs, engine.xb_flags(specs)):
            act = engine.run_conv_spec(act, sp, saved, xb)
        ctx.seq, ctx.rec, ctx.need_x = seq, (specs, saved), x.requires_grad
        y = ops.from_split(act)
        ctx.out_act = act
        return y

    @staticmethod
    def backward(ctx, gy):
        specs, saved = ctx.rec
        bag = _GradBag().reserve(gy.device)
        g = ops.nchw_to_nhwc_f32(gy)
        gx = bn_sequential_backward(specs, saved, g, bag, ctx.need_x)
        gx = ops.nhwc_f32_to_nchw(gx, specs[0].conv.in_channels) if (gx is not None and ctx.need_x) else None
        ctx.rec = None
        _join_side_stream()
        return (None, gx) + _ret_grads(bag, _params(ctx.seq))


def sequential_with_grad(seq, x):
    return _SequentialFn.apply(seq, x, *_params(seq))


# ---------------------------------------------------------------------------------------------------------------------
# single-stream VGG (script-local class of spatialstream.py / temporalstream.py)
# ---------------------------------------------------------------------------------------------------------------------
class _VGGFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, *params):
        ops.pack_cache.refresh()   # one launch re-packs every weight copy the optimiser step made stale
        specs, _ = engine.parse_sequential(model.features)
        saved_f, saved_d = [], []
        dspecs, head = engine.parse_sequential(model.decoder)
        act = ops.to_split(x, _first_cp(specs[0].conv), xb=_req(specs[0].conv.weight))
        for sp, xb in zip(specs, engine.xb_flags(specs, dspecs[0].conv.weight.requires_grad)):
            act = engine.run_conv_spec(act, sp, saved_f, xb)
        feat_act = act
        for sp, xb in zip(dspecs, engine.xb_flags(dspecs)):
            act = engine.run_conv_spec(act, sp, saved_d, xb)
        out = ops.head_fwd(act, head.weight, head.bias)
        ctx.model, ctx.rec, ctx.need_x = model, (specs, saved_f, dspecs, saved_d, head, act, out), x.requires_grad
        if model.return_features:
            feat = ops.from_split(feat_act)
            ctx.mark_non_differentiable(feat)
            return out, feat
        return out

    @staticmethod
    def backward(ctx, gout, *unused):
        model = ctx.model
        specs, saved_f, dspecs, saved_d, head, head_in, out = ctx.rec
        bag = _GradBag().reserve(gout.device)
        gpre, dw, db = ops.head_bwd(head_in, head.weight, out, gout, relu_mask=True, zeros=bag.zeros)
        bag.put(head.weight, dw)
        bag.put(head.bias, db)
        trunk_need = _any_req([model.features]) or ctx.need_x
        g = relu_sequential_backward(dspecs, saved_d, gpre, bag, trunk_need)
        gx = bn_sequential_backward(specs, saved_f, g, bag, ctx.need_x) if trunk_need else None
        gx = ops.nhwc_f32_to_nchw(gx, specs[0].conv.in_channels) if (gx is not None and ctx.need_x) else None
        ctx.rec = None
        _join_side_stream()
        return (None, gx) + _ret_grads(bag, _params(model))


def vgg_with_grad(model, x):
    return _VGGFn.apply(model, x, *_params(model))


# ---------------------------------------------------------------------------------------------------------------------
# late_fusion
# ---------------------------------------------------------------------------------------------------------------------
class _LateFusionFn(torch.autograd.Function):
    """models/late_fusion.py forward + backward through the dedicated LF kernels (csrc/lf.cu): one C-ABI call each way."""

    @staticmethod
    def forward(ctx, model, f, g, *params):
        out, saved = ops.lf_forward(model.fusion, f, g, keep=True)
        ctx.model, ctx.saved = model, saved
        ctx.need_x = (f.requires_grad, g.requires_grad)
        return out

    @staticmethod
    def backward(ctx, gout):
        model = ctx.model
        convs, bns = ops.lf_parts(model.fusion)
        r = ops.lf_backward(model.fusion, ctx.saved, gout, need_w=tuple(_req(c.weight) for c in convs),
                            need_f=ctx.need_x[0], need_g=ctx.need_x[1])
        bag = _GradBag()
        for i, c in enumerate(convs):
            if r["dw"][i] is not None:
                bag.put(c.weight, r["dw"][i])
        for i, bn in enumerate(bns):
            bag.put(bn.weight, r["dgamma"][i][:bn.num_features])
            bag.put(bn.bias, r["dbeta"][i][:bn.num_features])
            if _req(convs[i].bias):
                if ctx.saved[7]:   # feeds a batch-statistics BatchNorm: exactly zero
                    bag.put(convs[i].bias, bag.zeros_like(convs[i].bias))
                else:              # running statistics: d(raw) = scale * gz, so sum d(raw) = scale * dbeta
                    bag.put(convs[i].bias, ctx.saved[5][i, 2, :bn.num_features] * r["dbeta"][i][:bn.num_features])
        bag.put(convs[3].bias, r["dbh"])
        ctx.saved = None
        return (None, r["gf"], r["gg"]) + _ret_grads(bag, _params(model))


def late_fusion_with_grad(model, f, g):
    return _LateFusionFn.apply(model, f, g, *_params(model))


# ---------------------------------------------------------------------------------------------------------------------
# lstmnet (AT.trainLSTM: one-step-ahead MSE on stored 512-vectors, AT.py:118-147)
# ---------------------------------------------------------------------------------------------------------------------
class _LstmNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, inp, h0, c0, *params):
        out, hn, cn, ws = ops.lstm_seq_fwd(inp, h0, c0, model.lstm, model.lin, save_gates=True)
        ctx.model, ctx.rec = model, (inp.detach(), h0.detach(), c0.detach(), out, ws)
        ctx.need = (inp.requires_grad, h0.requires_grad or c0.requires_grad)
        return out, hn, cn

    @staticmethod
    def backward(ctx, gout, ghn, gcn):
        model = ctx.model
        inp, h0, c0, out, ws = ctx.rec
        if gout is None:
            gout = torch.zeros_like(out)
        r = ops.lstm_seq_bwd(inp, h0, c0, model.lstm, model.lin, out, ws, gout, ghn, gcn, ctx.need[0], ctx.need[1])
        bag = _GradBag()
        for l in range(2):
            bag.put(getattr(model.lstm, "weight_ih_l%d" % l), r["dw_ih"][l])
            bag.put(getattr(model.lstm, "weight_hh_l%d" % l), r["dw_hh"][l])
            bag.put(getattr(model.lstm, "bias_ih_l%d" % l), r["db"][l])
            bag.put(getattr(model.lstm, "bias_hh_l%d" % l), r["db"][l].clone())
        bag.put(model.lin.weight, r["dlin_w"])
        bag.put(model.lin.bias, r["dlin_b"])
        ctx.rec = None
        return (None, r["dinput"], r["dh0"], r["dc0"]) + _ret_grads(bag, _params(model))


def lstmnet_with_grad(model, inp, h0, c0):
    out, hn, cn = _LstmNetFn.apply(model, inp, h0, c0, *_params(model))
    return (out, (hn, cn))
