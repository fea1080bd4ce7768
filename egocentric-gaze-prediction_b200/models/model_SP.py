"""Drop-in for the reference's models/model_SP.py: same constructor, attributes, registration order
(features_t before features_s), state_dict keys/shapes and forward signature; forward runs on the egaze engine."""
import math

import torch.nn as nn

from egaze import engine, _lib
from egaze.modules import DecoderSequential, _needs_grad


class model_SP(nn.Module):
    def __init__(self, features_s, features_t):
        super(model_SP, self).__init__()
        self.features_t = features_t
        self.features_s = features_s
        self.relu = nn.ReLU()
        # parameter containers only: Conv3d(1,3,3)+MaxPool3d(2,1,1) == one shared 3x3 conv on both streams + max
        self.fusion = nn.Conv3d(512, 512, kernel_size=(1, 3, 3), padding=(0, 1, 1))
        self.pool3d = nn.MaxPool3d(kernel_size=(2, 1, 1), padding=0)
        self.bn = nn.BatchNorm2d(512)
        chans = [(512, 512), (512, 512), 'U', (512, 512), (512, 512), (512, 512), 'U', (512, 256), (256, 256),
                 (256, 256), 'U', (256, 128), (128, 128), 'U', (128, 64), (64, 64)]
        mods = []
        for c in chans:
            if c == 'U':
                mods.append(nn.Upsample(scale_factor=2))
            else:
                mods += [nn.Conv2d(c[0], c[1], kernel_size=3, padding=1), nn.ReLU(inplace=True)]
        mods.append(nn.Conv2d(64, 1, kernel_size=1, padding=0))
        self.decoder = DecoderSequential(*mods)
        self.final = nn.Sigmoid()
        self._initialize_weights()

    def forward(self, x_s, x_t):
        _lib.check_device(x_s.device)
        if _needs_grad(self, x_s, x_t):
            from egaze.autograd import model_sp_with_grad
            return model_sp_with_grad(self, x_s, x_t)
        y_s = self.features_s(x_s)  # module calls: forward hooks on features_s keep firing (AT.py:105)
        y_t = self.features_t(x_t)
        return engine.run_sp_tail(self, engine.get_act(y_s), engine.get_act(y_t))

    def _initialize_weights(self):
        # Conv2d ~ N(0, sqrt(2/(k*k*Cout))), zero bias; BN gamma=1 beta=0; the Conv3d keeps PyTorch's default init
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
