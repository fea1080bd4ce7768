"""nn.Module containers that keep the reference's parameter/child layout but run through the egaze engine."""
import torch
import torch.nn as nn

from . import engine, ops, _lib


def _needs_grad(module, *tensors):
    if not torch.is_grad_enabled():
        return False
    if any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        return True
    return any(p.requires_grad for p in module.parameters())


class TrunkSequential(nn.Sequential):
    """What utils.make_layers returns: children/indices/state_dict identical to the reference's nn.Sequential
    (reference utils.py:64-76), forward fused on the tcgen05 path.  Input/outputs are NCHW fp32 like the reference;
    the output additionally carries its split-NHWC form (engine.attach_act) so the next stage skips a conversion."""

    def forward(self, x):
        _lib.check_device(x.device)
        if _needs_grad(self, x):
            from .autograd import sequential_with_grad
            return sequential_with_grad(self, x)
        cin = next(m for m in self.children() if isinstance(m, nn.Conv2d)).in_channels
        act = engine.get_act(x, ops.pad_channels(cin))
        act, tail = engine.run_sequential(self, act)
        if tail is not None:
            raise RuntimeError("egaze: unexpected 1x1 conv inside a trunk")
        return engine.attach_act(ops.from_split(act), act)


class DecoderSequential(nn.Sequential):
    """model_SP.decoder container (reference models/model_SP.py:13-31).  Called stand-alone it maps the fused
    512x14x14 map to the pre-sigmoid logit, like the reference's nn.Sequential."""

    def forward(self, x):
        _lib.check_device(x.device)
        if _needs_grad(self, x):
            raise RuntimeError("egaze: call model_SP.forward for training; the stand-alone decoder is inference-only")
        act = engine.get_act(x, ops.pad_channels(x.shape[1]))
        act, tail = engine.run_sequential(self, act)
        if tail is None:
            return engine.attach_act(ops.from_split(act), act)
        _, logit = ops.head_fwd(act, tail.weight, tail.bias, want_logit=True)
        return logit
