"""Data-parallel plumbing: frames shard across ranks, ONE all-reduce per optimiser step.

The reference is single-GPU (gaze_full.py:37); this is the only multi-GPU mechanism the path needs.  After the
backward the gradients are gathered into ONE flat fp32 buffer with a single multi-tensor copy, averaged with one NCCL
all-reduce over NVLink, and every .grad is re-pointed at its slice of the buffer (the optimiser reads the averaged values
in place, no scatter).  Letting autograd ACCUMULATE into pre-set .grad views instead costs one small add kernel per
parameter (215 for model_SP) every step.  BatchNorm statistics stay per replica.
"""
import torch
import torch.distributed as dist


class FlatGradBucket(object):
    def __init__(self, params, device=None):
        self.params = [p for p in params if p.requires_grad]
        device = device if device is not None else self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.views = []
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            p.grad = v
            self.views.append(v)
            off += p.numel()

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
