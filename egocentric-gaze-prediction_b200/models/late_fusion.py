"""Drop-in for the reference's models/late_fusion.py: same attributes (`upsample`, `fusion`, `final`), state_dict
and forward(f, g) signature.  The whole net runs through the dedicated late-fusion kernels (egaze_lf_fwd / egaze_lf_bwd,
csrc/lf.cu): BatchNorm affine + ReLU applied while the next conv stages its window, statistics in the conv epilogues."""
import math

import torch
import torch.nn as nn

from egaze import ops, _lib
from egaze.modules import _needs_grad


class late_fusion(nn.Module):
    def __init__(self):
        super(late_fusion, self).__init__()
        self.upsample = nn.Upsample(scale_factor=16)  # unused by forward, kept for attribute parity (late_fusion.py:9,19)
        self.fusion = nn.Sequential(
            nn.Conv2d(2, 32, kernel_size=3, padding=1), nn.BatchNorm2d(32), nn.ReLU(inplace=True),
            nn.Conv2d(32, 32, kernel_size=3, padding=1), nn.BatchNorm2d(32), nn.ReLU(inplace=True),
            nn.Conv2d(32, 8, kernel_size=3, padding=1), nn.BatchNorm2d(8), nn.ReLU(inplace=True),
            nn.Conv2d(8, 1, kernel_size=1, padding=0))
        self.final = nn.Sigmoid()
        self._initialize_weights()

    def forward(self, f, g):
        _lib.check_device(f.device)
        if _needs_grad(self, f, g):
            from egaze.autograd import late_fusion_with_grad
            return late_fusion_with_grad(self, f, g)
        return ops.lf_forward(self.fusion, f, g)[0]  # channel order (f, g) as in late_fusion.py:20

    def _initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()
