#!/usr/bin/env python
"""How much gradient accuracy does a cheaper backward cost?  

Runs one model_SP train step (forward + floss + backward) per backward mode on the SAME parameters and inputs and reports, per
parameter tensor, rel-L2 against (a) the 3-pass split-bf16 backward of the same forward (isolates the backward's own rounding:
the forward, hence every ReLU / max-pool routing decision, is identical) and (b) stock torch autograd in fp64.  Stock fp32
autograd vs fp64 is printed beside it as the noise floor.  Modes are emulated numerically with the existing kernels (a zeroed lo
plane == a skipped MMA):
    dgrad_hi      data gradient from dY_hi x [W_hi | W_lo]      (2 MMAs per product)
    wgrad_1pass   weight gradient from dY_hi x X_hi             (1 MMA)
    wgrad_xhi     weight gradient from [dY_hi; dY_lo] x X_hi    (2 MMAs)
"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "egocentric-gaze-prediction_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np
import torch
import torch.nn.functional as F

import torch_ref
from oracle import egaze_oracle as orc


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    from utils import make_layers, cfg
    from models.model_SP import model_SP
    import floss as floss_mod
    B, S = int(os.environ.get("GM_B", "4")), int(os.environ.get("GM_S", "224"))
    torch.manual_seed(0)
    m = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20))
    torch_ref.randomize_(m, 0)
    with torch.no_grad():
        for mod in m.decoder:
            if isinstance(mod, torch.nn.Conv2d):
                mod.weight.mul_(0.8)
    m = m.to(dev).train()
    m32 = copy.deepcopy(m)
    m64 = copy.deepcopy(m).double()
    x_s, x_t, gt = [torch.from_numpy(a).to(dev) for a in orc.synth_sp_inputs(B, S, 5)]
    torch_ref.floss_loss(torch_ref.model_sp_forward(m32, x_s, x_t), gt).backward()
    o64 = torch_ref.model_sp_forward(m64, x_s.double(), x_t.double())
    F.binary_cross_entropy(o64, gt.double(), weight=torch_ref.floss_weight(gt).double()).backward()

    modes = [("3pass", {}),
             ("dgrad_hi", {"EGAZE_EMU_DGRAD_HI": "1"}),
             ("wgrad_1pass", {"EGAZE_EMU_WGRAD_1PASS": "1"}),
             ("wgrad_xhi", {"EGAZE_EMU_WGRAD_XHI": "1"}),
             ("dgrad_hi+wgrad_1pass", {"EGAZE_EMU_DGRAD_HI": "1", "EGAZE_EMU_WGRAD_1PASS": "1"}),
             ("dgrad_hi+wgrad_xhi", {"EGAZE_EMU_DGRAD_HI": "1", "EGAZE_EMU_WGRAD_XHI": "1"})]
    grads = {}
    for name, env in modes:
        for k in ("EGAZE_EMU_DGRAD_HI", "EGAZE_EMU_WGRAD_1PASS", "EGAZE_EMU_WGRAD_XHI"):
            os.environ.pop(k, None)
        os.environ.update(env)
        mm = copy.deepcopy(m32)
        mm.zero_grad(set_to_none=True)
        floss_mod.floss()(mm(x_s, x_t), gt).backward()
        grads[name] = {k: p.grad.detach().clone() for k, p in mm.named_parameters()}
    names = [k for k, p in m64.named_parameters() if p.grad.norm().item() >= 1e-7]
    noise = {k: rel_l2(dict(m32.named_parameters())[k].grad, dict(m64.named_parameters())[k].grad) for k in names}
    print("B=%d S=%d  stock fp32 vs fp64: median %.2e  max %.2e" % (B, S, np.median(list(noise.values())), max(noise.values())))
    for name, _ in modes:
        vs3 = [rel_l2(grads[name][k], grads["3pass"][k]) for k in names]
        vs64 = [rel_l2(grads[name][k], dict(m64.named_parameters())[k].grad) for k in names]
        ratio = [a / max(noise[k], 1e-12) for a, k in zip(vs64, names)]
        iw = int(np.argmax(vs3))
        print("%-22s vs 3pass: median %.2e max %.2e (%s) | vs fp64: median %.2e max %.2e | max ratio to stock-fp32 noise %.1f"
              % (name, np.median(vs3), max(vs3), names[iw], np.median(vs64), max(vs64), max(ratio)))
    # per-group detail for the combined mode
    for name in ("dgrad_hi+wgrad_1pass",):
        print("-- %s, per tensor (vs 3pass | vs fp64 | stock noise)" % name)
        for k in names:
            if k.endswith("weight") and ("features" in k or "decoder" in k or "fusion" in k):
                print("   %-28s %.2e | %.2e | %.2e" % (k, rel_l2(grads[name][k], grads["3pass"][k]),
                                                     rel_l2(grads[name][k], dict(m64.named_parameters())[k].grad), noise[k]))


if __name__ == "__main__":
    main()
