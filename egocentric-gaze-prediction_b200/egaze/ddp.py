"""Data-parallel plumbing: frames shard across ranks, ONE all-reduce per optimiser step.

The reference is single-GPU (gaze_full.py:37); this is the only multi-GPU mechanism the path needs.  After the
backward the gradients are gathered into ONE flat fp32 buffer with a single multi-tensor copy, averaged with one NCCL
all-reduce over NVLink, and every .grad is re-pointed at its slice of the buffer (the optimiser reads the averaged values
in place, no scatter).  Letting autograd ACCUMULATE into pre-set .grad views instead costs one small add kernel per
parameter (215 for model_SP) every step.  BatchNorm statistics stay per replica.
"""
import torch
import torch.distributed as dist


def _pad4(n):
    return (n + 3) // 4 * 4


class FlatGradBucket(object):
    def __init__(self, params, device=None):
        self.params = [p for p in params if p.requires_grad]
        device = device if device is not None else self.params[0].device
        self.numel = sum(_pad4(p.numel()) for p in self.params)     # every view starts 16-byte aligned (vector loads)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.views = []
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            p.grad = v
            self.views.append(v)
            off += _pad4(p.numel())

    def zero(self):
        """Drop the gradients (the next backward assigns fresh tensors instead of adding into the bucket)."""
        for p in self.params:
            p.grad = None

    def allreduce(self, group=None):
        """Gather every parameter's gradient into the flat buffer (one multi-tensor copy; parameters without a gradient
        contribute zeros), average over all ranks (no-op for a single process) and alias .grad to the buffer."""
        src, dst, missing = [], [], []
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None:
                missing.append(v)
            elif g.data_ptr() != v.data_ptr():
                src.append(g.reshape(v.shape))
                dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)
        if missing:
            torch._foreach_zero_(missing)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:  # gloo (CPU tests) has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))
        for p, v in zip(self.params, self.views):
            p.grad = v
        return self.flat


def shard_seed(base_seed, rank):
    """Synthetic-data seed of a rank: identical replicas (weights seeded separately), disjoint data."""
    return int(base_seed) + int(rank)


def broadcast_parameters(module, src=0, group=None):
    """Make every replica start from rank `src`'s parameters and buffers (what DDP does at construction)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


class OverlappedGradReducer(object):
    """Gradient averaging that starts while the backward is still running ("bucket order decoder -> bn/fusion ->
    trunks").

    model_SP's backward is ONE autograd node (egaze.autograd._ModelSPFn), so stock DDP's per-parameter hooks would all fire at
    its end.  Instead the node itself calls `reduce(bag, name)` each time a SEGMENT of parameters has its final gradients --
    the decoder + head, then fusion + bn, then the deep trunk layers (conv3_1 .. conv5_3 of both trunks: 98 % of the trunk
    parameters, ready while the expensive 112^2 / 224^2 layers of the backward still run), then the rest.  The flat fp32 buffer
    is laid out in that order, so a segment is one contiguous slice: it is filled on a communication stream (which first waits
    for the streams that produced the gradients), all-reduced there (NCCL over NVLink; AVG), and the gradients autograd hands to
    the optimiser ARE views of the buffer -- no scatter, no extra pass.  `finish()` makes the main stream wait for the
    communication stream before the node returns.  Works under CUDA-graph capture (every fork is joined inside the node).
    BatchNorm statistics stay per replica.
    """

    def __init__(self, segments, group=None, device=None):
        """segments: ordered [(name, [parameters])]; parameters with requires_grad=False are skipped."""
        self.group = group
        self.segments = []
        seen = set()
        for name, params in segments:
            ps = [p for p in params if p.requires_grad and id(p) not in seen]
            seen.update(id(p) for p in ps)
            if ps:
                self.segments.append((name, ps))
        self.params = [p for _, ps in self.segments for p in ps]
        device = device if device is not None else self.params[0].device
        self.numel = sum(_pad4(p.numel()) for p in self.params)     # every view starts 16-byte aligned (vector loads)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.offset, self.span, off = {}, {}, 0
        for name, ps in self.segments:
            lo = off
            for p in ps:
                self.offset[id(p)] = (off, p.numel())
                off += _pad4(p.numel())
            self.span[name] = (lo, off)
        self.comm = torch.cuda.Stream(device=device) if self.flat.is_cuda else None
        self.done = set()
        self.calls = 0

    def world(self):
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def view(self, p):
        off, n = self.offset[id(p)]
        return self.flat[off:off + n].view_as(p)

    def begin(self):
        self.done.clear()

    def reduce(self, bag, name, producers=()):
        """Average segment `name` over all ranks: the gradients `bag` holds for its parameters are replaced by views of the
        flat buffer.  producers: CUDA streams whose work the gradients depend on (besides the current one)."""
        if name not in self.span or name in self.done:
            return
        self.done.add(name)
        ps = dict(self.segments)[name]
        lo, hi = self.span[name]
        views, grads, missing = [], [], []
        for p in ps:
            g = bag.get(p)
            v = self.view(p)
            if g is None:
                missing.append(v)
            else:
                views.append(v)
                grads.append(g.reshape(v.shape))
        seg = self.flat[lo:hi]
        if self.comm is not None:
            cur = torch.cuda.current_stream(self.flat.device)
            self.comm.wait_stream(cur)
            for s in producers:
                if s is not None:
                    self.comm.wait_stream(s)
            with torch.cuda.stream(self.comm):
                self._fill(views, grads, missing)
                self._allreduce(seg)
            for g in grads:
                g.record_stream(self.comm)
        else:
            self._fill(views, grads, missing)
            self._allreduce(seg)
        for p in ps:
            bag.d[id(p)] = self.view(p)
        self.calls += 1

    def reduce_rest(self, bag, producers=()):
        for name, _ in self.segments:
            self.reduce(bag, name, producers)

    @staticmethod
    def _fill(views, grads, missing):
        if views:
            torch._foreach_copy_(views, grads)
        if missing:
            torch._foreach_zero_(missing)

    def _allreduce(self, seg):
        if self.world() > 1:
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(seg, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.group)
                seg.div_(self.world())

    def finish(self):
        """The gradients are consumed on the current stream from here on."""
        if self.comm is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm)


DEEP_FROM = 4   # trunk layer index (of 13) of conv3_1: layers [4:] hold 98 % of a VGG16 trunk's parameters


def sp_segments(model):
    """Reduction order of model_SP's parameters = the order its backward finishes them (models/model_SP.py:35-50 reversed)."""
    import torch.nn as nn

    def trunk_layers(seq):
        layers, cur = [], None
        for m in seq.children():
            if isinstance(m, nn.Conv2d):
                cur = list(m.parameters())
                layers.append(cur)
            elif isinstance(m, nn.BatchNorm2d) and cur is not None:
                cur.extend(m.parameters())
        return layers
    lt, ls = trunk_layers(model.features_t), trunk_layers(model.features_s)
    flat = lambda layers: [p for layer in layers for p in layer]
    return [("decoder", list(model.decoder.parameters())),
            ("fusion_bn", list(model.fusion.parameters()) + list(model.bn.parameters())),
            ("trunk_deep", flat(lt[DEEP_FROM:]) + flat(ls[DEEP_FROM:])),
            ("trunk_shallow", flat(lt[:DEEP_FROM]) + flat(ls[:DEEP_FROM])),
            ("rest", list(model.parameters()))]


def attach_reducer(model, group=None):
    """Build the overlapped reducer of a model_SP and hang it on the module, where its backward node looks it up."""
    r = OverlappedGradReducer(sp_segments(model), group=group)
    model._egaze_reducer = r
    return r
