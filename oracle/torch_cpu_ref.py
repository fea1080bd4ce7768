"""CPU BASELINE PORT -- TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product path).

The reference (hyf015/egocentric-gaze-prediction) is pure Python that calls stock torch.nn modules; on a box
without /root/reference its CPU path is reproduced here by building the SAME stock modules in the same order
(models/model_SP.py:4-50, utils.py:64-76, models/late_fusion.py:6-23, models/LSTMnet.py:15-37, floss.py:5-41) and
running them with PyTorch's CPU kernels (oneDNN/MKL) -- i.e. exactly the arithmetic the reference executes on CPU.
Used by bench.py (`cpu_baseline` leg and `--impl reference`) and validated against the golden fixtures in
tests/test_oracle_golden.py::test_torch_cpu_port_matches_golden.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import egaze_oracle as orc


def make_trunk(cin):
    layers = []
    for v in orc.CFG_D:
        if v == 'M':
            layers.append(nn.MaxPool2d(2, 2))
        else:
            layers += [nn.Conv2d(cin, v, 3, padding=1), nn.BatchNorm2d(v), nn.ReLU()]
            cin = v
    return nn.Sequential(*layers)


def make_decoder(spec):
    layers = []
    for v in spec:
        if v == 'U':
            layers.append(nn.Upsample(scale_factor=2))
        else:
            layers += [nn.Conv2d(v[0], v[1], 3, padding=1), nn.ReLU(inplace=True)]
    layers.append(nn.Conv2d(64, 1, 1))
    return nn.Sequential(*layers)


class ModelSP(nn.Module):
    """Stock-op port of models/model_SP.py:4-50 (registration order t, s as in the reference)."""

    def __init__(self):
        super().__init__()
        self.features_t = make_trunk(20)
        self.features_s = make_trunk(3)
        self.relu = nn.ReLU()
        self.fusion = nn.Conv3d(512, 512, (1, 3, 3), padding=(0, 1, 1))
        self.pool3d = nn.MaxPool3d((2, 1, 1))
        self.bn = nn.BatchNorm2d(512)
        self.decoder = make_decoder(orc.DEC_SP)

    def forward(self, x_s, x_t):
        x_s = self.features_s(x_s).unsqueeze(2)
        x_t = self.features_t(x_t).unsqueeze(2)
        x = self.pool3d(self.fusion(torch.cat((x_s, x_t), 2))).squeeze(2)
        return torch.sigmoid(self.decoder(self.relu(self.bn(x))))


class LateFusion(nn.Module):
    def __init__(self):
        super().__init__()
        self.fusion = nn.Sequential(nn.Conv2d(2, 32, 3, padding=1), nn.BatchNorm2d(32), nn.ReLU(inplace=True),
                                    nn.Conv2d(32, 32, 3, padding=1), nn.BatchNorm2d(32), nn.ReLU(inplace=True),
                                    nn.Conv2d(32, 8, 3, padding=1), nn.BatchNorm2d(8), nn.ReLU(inplace=True),
                                    nn.Conv2d(8, 1, 1))

    def forward(self, f, g):
        return torch.sigmoid(self.fusion(torch.cat((f, g), 1)))


class LSTMNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.lstm = nn.LSTM(512, 512, 2)
        self.lin = nn.Linear(512, 512)

    def forward(self, x, hidden):
        out, hidden = self.lstm(torch.tanh(x), hidden)
        return F.relu(self.lin(out)), hidden


def floss(inp, target):
    """floss.forward (floss.py:9-13) with the NumPy weight builder of the oracle (floss.py:15-41)."""
    w = torch.from_numpy(orc.floss_weight(target.detach().numpy()))
    return F.binary_cross_entropy(inp, target, weight=w)


def load_synth(model, seed):
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = orc.synth_state_dict(shapes, seed)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return model
