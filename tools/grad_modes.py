#!/usr/bin/env python
"""How much gradient accuracy does a cheaper backward cost?  

Runs one model_SP train step (forward + floss + backward) per numeric mode (EGAZE_PRECISION = precise3 | precise | fast, see
egaze/ops.py) on the SAME parameters and inputs and reports the train-mode gaze map error and, per parameter tensor, the
gradient's rel-L2 against stock torch autograd in fp64.  Stock fp32 autograd vs fp64 is printed beside it as the noise floor
(ReLU / max-pool routing flips make fp32 itself ~1e-2 from fp64 on this step).
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

    modes = [("precise3", {"EGAZE_PRECISION": "precise3"}), ("precise", {"EGAZE_PRECISION": "precise"}),
             ("fast", {"EGAZE_PRECISION": "fast"})]
    grads, outs = {}, {}
    for name, env in modes:
        os.environ.update(env)
        mm = copy.deepcopy(m32)
        mm.zero_grad(set_to_none=True)
        out = mm(x_s, x_t)
        floss_mod.floss()(out, gt).backward()
        outs[name] = out.detach()
        grads[name] = {k: p.grad.detach().clone() for k, p in mm.named_parameters()}
    os.environ.pop("EGAZE_PRECISION", None)
    o32 = torch_ref.model_sp_forward(copy.deepcopy(m32), x_s, x_t).detach()
    print("train-mode gaze map max-abs vs fp64: stock fp32 %.2e | %s" % (
        (o32.double() - o64.detach()).abs().max().item(),
        " | ".join("%s %.2e" % (n, (outs[n].double() - o64.detach()).abs().max().item()) for n, _ in modes)))
    names = [k for k, p in m64.named_parameters() if p.grad.norm().item() >= 1e-7]
    noise = {k: rel_l2(dict(m32.named_parameters())[k].grad, dict(m64.named_parameters())[k].grad) for k in names}
    print("B=%d S=%d  stock fp32 vs fp64: median %.2e  max %.2e" % (B, S, np.median(list(noise.values())), max(noise.values())))
    for name, _ in modes:
        vs3 = [rel_l2(grads[name][k], grads["precise3"][k]) for k in names]
        vs64 = [rel_l2(grads[name][k], dict(m64.named_parameters())[k].grad) for k in names]
        ratio = [a / max(noise[k], 1e-12) for a, k in zip(vs64, names)]
        iw = int(np.argmax(vs3))
        print("%-22s vs precise3: median %.2e max %.2e (%s) | vs fp64: median %.2e max %.2e | max ratio to stock-fp32 noise %.1f"
              % (name, np.median(vs3), max(vs3), names[iw], np.median(vs64), max(vs64), max(ratio)))
    print("-- per tensor: precise vs fp64 | precise3 vs fp64 | stock fp32 vs fp64")
    for k in names:
        if k.endswith("weight") and ("features" in k or "decoder" in k or "fusion" in k):
            g64 = dict(m64.named_parameters())[k].grad
            print("   %-28s %.2e | %.2e | %.2e" % (k, rel_l2(grads["precise"][k], g64), rel_l2(grads["precise3"][k], g64), noise[k]))


if __name__ == "__main__":
    main()
