"""Single-stream VGG saliency net: the class the reference defines INSIDE its scripts (spatialstream.py:65-116,
temporalstream.py:64-113, run_spatialstream.py:17-68) -- trunk + 13-conv decoder (three convs at 14x14) + sigmoid.
Same attribute names / state_dict keys ('features.*', 'decoder.*'); `return_features=True` gives the demo variant
that also returns conv5_3 (run_spatialstream.py:49-53)."""
import math

import torch.nn as nn

from . import engine, ops, _lib
from .modules import DecoderSequential, _needs_grad


class VGG(nn.Module):
    def __init__(self, features, return_features=False, freeze_features=True):
        super(VGG, self).__init__()
        self.features = features
        if freeze_features:  # spatialstream.py:70-71 (temporalstream.py leaves them trainable)
            for param in self.features.parameters():
                param.requires_grad = False
        chans = [(512, 512), (512, 512), (512, 512), 'U', (512, 512), (512, 512), (512, 512), 'U', (512, 256),
                 (256, 256), (256, 256), 'U', (256, 128), (128, 128), 'U', (128, 64), (64, 64)]
        mods = []
        for c in chans:
            if c == 'U':
                mods.append(nn.Upsample(scale_factor=2))
            else:
                mods += [nn.Conv2d(c[0], c[1], kernel_size=3, padding=1), nn.ReLU(inplace=True)]
        mods.append(nn.Conv2d(64, 1, kernel_size=1, padding=0))
        self.decoder = DecoderSequential(*mods)
        self.final = nn.Sigmoid()
        self.return_features = return_features
        self._initialize_weights()

    def forward(self, x):
        _lib.check_device(x.device)
        if _needs_grad(self, x):
            from .autograd import vgg_with_grad
            return vgg_with_grad(self, x)
        xo = self.features(x)
        act, tail = engine.run_sequential(self.decoder, engine.get_act(xo))
        y = ops.head_fwd(act, tail.weight, tail.bias)
        return (y, xo) if self.return_features else y

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
