"""Fused multi-tensor Adam for the drop-in modules.

`egaze.optim.Adam` takes the arguments of `torch.optim.Adam` (the reference builds one per stage: SP.py:110-113, LF.py:77,
AT.py:84) and keeps its state layout (`step`, `exp_avg`, `exp_avg_sq` per parameter, `state_dict()` interchangeable with
torch's), but a step is TWO kernel launches whatever the number of parameters (egaze_adam_multi), and for every 3x3 conv weight
the same launch rewrites the packed tensor-core operand copies the next forward / backward reads (egaze.ops.pack_cache), so no
separate re-pack pass follows an optimiser step.  Everything is device-side (the step counters too), so the optimiser can be
captured in a CUDA graph (egaze.graph.GraphedStep) without `capturable=` plumbing.

Opt-in: the reference's loops keep working with torch.optim.Adam; bench.py uses this one inside its captured step.
"""
import ctypes

import numpy as np
import torch

from . import ops, _lib
from ._lib import call, stream_ptr

_JOB_DT = np.dtype([("w", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("step", "<u8"), ("p0_hi", "<u8"), ("p0_lo", "<u8"),
                    ("p1_hi", "<u8"), ("p1_lo", "<u8"), ("n", "<i8"), ("Co", "<i4"), ("Ci", "<i4"), ("rows0", "<i4"),
                    ("cols0", "<i4"), ("fmt0", "<i4"), ("rows1", "<i4"), ("cols1", "<i4"), ("sub", "<i4")])


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **unused):
        if amsgrad:
            raise NotImplementedError("egaze.optim.Adam: amsgrad is not implemented (the reference never uses it)")
        for k in unused:
            if k not in ("capturable", "foreach", "fused", "maximize", "differentiable"):
                raise TypeError("egaze.optim.Adam: unexpected argument %r" % k)
        if unused.get("maximize") or unused.get("differentiable"):
            raise NotImplementedError("egaze.optim.Adam: maximize / differentiable are not implemented")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0:
            raise ValueError("egaze.optim.Adam: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))
        nbytes = ctypes.c_int(0)
        call("egaze_adam_job_bytes", ctypes.addressof(nbytes))
        if nbytes.value != _JOB_DT.itemsize:
            raise RuntimeError("egaze.optim.Adam: job record is %d bytes here, %d in libegaze.so" % (_JOB_DT.itemsize, nbytes.value))
        self._tables = {}     # (group index, pointer tuple) -> (pinned host table, device table)

    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif not torch.is_tensor(st["step"]) or st["step"].device != p.device:   # state loaded from a torch.optim.Adam checkpoint
            st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32, device=p.device)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            recs, packs = [], []
            for p in live:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("egaze.optim.Adam: parameters must be contiguous fp32 CUDA tensors")
                g = p.grad
                if g.is_sparse or g.dtype != torch.float32 or g.device != p.device:
                    raise RuntimeError("egaze.optim.Adam: gradients must be dense fp32 tensors on the parameter's device")
                if not g.is_contiguous():
                    g = p.grad = g.contiguous()
                st = self._init_state(p)
                ents = ops.pack_cache.entries_for(p)
                is_conv = ents and (p.dim() == 4 or p.dim() == 5) and tuple(p.shape[-2:]) == (3, 3)
                # a conv that follows nn.Upsample runs in sub-pixel form: its copies are the 16-plane packs (modes 2 / 3)
                sub = bool(is_conv and any(e[0] >= 2 for e in ents))
                e0 = next((e for e in ents if e[0] == (2 if sub else 0)), None) if is_conv else None
                e1 = next((e for e in ents if e[0] == (3 if sub else 1)), None) if is_conv else None
                # copies the kernel cannot maintain (a second forward format of the same weight, padded rows) are left stale
                maintained = [e for e in (e0, e1) if e is not None]
                rec = (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), st["step"].data_ptr(),
                       e0[4].data_ptr() if e0 else 0, e0[5].data_ptr() if e0 else 0,
                       e1[4].data_ptr() if e1 else 0, e1[5].data_ptr() if e1 else 0,
                       p.numel(), int(p.shape[0]) if maintained else 0, int(p.shape[1]) if maintained else 0,
                       e0[1] if e0 else 0, e0[2] if e0 else 0, e0[3] if e0 else 0, e1[1] if e1 else 0, e1[2] if e1 else 0, int(sub))
                recs.append(rec)
                packs.append((p, maintained))
            key = (gi, tuple(r[:9] + r[-1:] for r in recs))
            tab = self._tables.get(key)
            if tab is None:
                host = torch.from_numpy(np.array(recs, dtype=_JOB_DT).view(np.uint8)).pin_memory()
                dev = torch.empty(host.shape, dtype=torch.uint8, device=live[0].device)
                dev.copy_(host, non_blocking=True)     # pinned -> device: legal under CUDA-graph capture
                while len(self._tables) >= 8:
                    self._tables.pop(next(iter(self._tables)))
                tab = self._tables[key] = (host, dev)
            b1, b2 = group["betas"]
            call("egaze_adam_multi", tab[1], len(recs), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                 float(group["weight_decay"]), ops.f16_weight_scale(), stream_ptr())
            for p, maintained in packs:
                torch.autograd.graph.increment_version(p)      # the kernel wrote p in place behind autograd's back
                ops.pack_cache.mark_current(p, maintained)
        return loss
