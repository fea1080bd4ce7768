"""Drop-in for the reference's floss.py: distance-weighted BCE.  forward(input, target) -> scalar loss;
build_weight_from_target(target) -> np.ndarray like the reference (floss.py:15-41) but computed on the GPU."""
import torch
import torch.nn as nn

from egaze import ops, _lib


class _FlossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, target):
        inp_c = inp.contiguous().float()
        tgt_c = target.detach().contiguous().float()
        cen = ops.floss_centroid(tgt_c)
        ctx.save_for_backward(inp_c, tgt_c, cen)
        return ops.floss_fwd(inp_c, tgt_c, cen)

    @staticmethod
    def backward(ctx, grad_out):
        inp_c, tgt_c, cen = ctx.saved_tensors
        return ops.floss_bwd(inp_c, tgt_c, cen, grad_out), None


class floss(nn.Module):
    def __init__(self):
        super(floss, self).__init__()

    def forward(self, input, target):
        _lib.check_device(input.device)
        if input.shape != target.shape or input.dim() != 4:
            raise ValueError("floss expects input and target of identical (B,1,H,W) shape")
        return _FlossFn.apply(input, target)

    def build_weight_from_target(self, target):
        return ops.floss_weight(target.data).cpu().numpy()
