"""Data-parallel plumbing: frames shard across ranks, ONE all-reduce per optimiser step.

The reference is single-GPU (gaze_full.py:37); this is the only multi-GPU mechanism the path needs.  Every trainable
parameter's .grad is a view into one flat fp32 buffer, so the backward kernels' outputs are accumulated in place and
the NCCL all-reduce (AVG) over NVLink needs no pack / unpack copies.  BatchNorm statistics stay per replica.
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
        """Zero the bucket and make sure every .grad still aliases it (zero_grad(set_to_none=True) would detach them)."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def allreduce(self, group=None):
        """Average the gradients over all ranks (no-op for a single process)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:  # gloo (CPU tests) has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))
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
