"""Drop-in for the reference's utils.py (same public names).  `make_layers`/`cfg` are the hot path (reference
utils.py:57-76) and build an egaze TrunkSequential; the remaining helpers are small host-side utilities the
reference's trainers star-import (`from utils import *`: SP.py:16, AT.py:13, LF.py:12)."""
import collections
import math
import os

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Variable  # noqa: F401  (re-exported: the reference's `from utils import *` relies on it)

from egaze.modules import TrunkSequential

# VGG configurations; the final max-pool of torchvision's VGG is dropped (reference utils.py:57-62)
cfg = {
    'A': [64, 'M', 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512],
    'B': [64, 64, 'M', 128, 128, 'M', 256, 256, 'M', 512, 512, 'M', 512, 512],
    'D': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512],
    'E': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 256, 'M', 512, 512, 512, 512, 'M', 512, 512, 512, 512],
}


def make_layers(cfg, in_channels, batch_norm=True):
    """VGG trunk factory (reference utils.py:64-76): same child modules at the same indices, fused execution."""
    mods = []
    c_in = in_channels
    for v in cfg:
        if v == 'M':
            mods.append(nn.MaxPool2d(kernel_size=2, stride=2))
            continue
        mods.append(nn.Conv2d(c_in, v, kernel_size=3, padding=1))
        if batch_norm:
            mods.append(nn.BatchNorm2d(v))
            mods.append(nn.ReLU(inplace=False))
        else:
            mods.append(nn.ReLU(inplace=True))
        c_in = v
    return TrunkSequential(*mods)


class generalException(Exception):
    pass


class AverageMeter(object):
    """Running average (reference utils.py:32-46)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def repackage_hidden(h):
    """Detach an LSTM state from its history (reference utils.py:49-55)."""
    if h is None:
        return None
    if type(h) == tuple:
        return tuple(repackage_hidden(v) for v in h)
    return h.data


def save_checkpoint(state, filename, save_path):
    torch.save(state, os.path.join(save_path, filename))


def change_key_names(old_params, in_channels):
    """ImageNet VGG16-BN -> flow trunk: first conv weight = RGB mean repeated over in_channels (reference utils.py:78-94)."""
    out = collections.OrderedDict()
    for n, (k, v) in enumerate(old_params.items()):
        if n >= 25:
            break
        out[k] = v.mean(dim=1, keepdim=True).repeat(1, in_channels, 1, 1) if n == 0 else v
    return out


def var_to_image(var):
    ten = var.data.cpu()
    if ten.dim() == 4:
        ten = ten[0].squeeze()
    if ten.dim() == 3:
        std = torch.FloatTensor([0.229, 0.224, 0.225]).view(3, 1, 1)
        mean = torch.FloatTensor([0.485, 0.456, 0.406]).view(3, 1, 1)
        return (ten * std + mean).numpy().transpose((1, 2, 0))
    if ten.dim() == 2:
        return ten.numpy()
    print('warning: input variable is invalid to transfer to image')
    return np.zeros((224, 224))


def _aae_auc_single(out_sq, tar_sq):
    from scipy import ndimage
    predicted = ndimage.center_of_mass(out_sq)
    i, j = np.unravel_index(tar_sq.argmax(), tar_sq.shape)
    d = 112 / math.tan(math.pi / 6)
    r1 = np.array([predicted[0] - 112, predicted[1] - 112, d])
    r2 = np.array([i - 112, j - 112, d])
    angle = math.degrees(math.atan2(np.linalg.norm(np.cross(r1, r2)), np.dot(r1, r2)))
    z = np.zeros((224, 224))
    z[int(predicted[0])][int(predicted[1])] = 1
    z = ndimage.gaussian_filter(z, 14)
    z = z - np.min(z)
    z = z / np.max(z)
    auc = 1 - float((z > z[i][j]).sum()) / (out_sq.shape[0] * out_sq.shape[1])
    return angle, auc, [i, j]


def computeAAEAUC(output, target):
    """Validation metric (reference utils.py:96-140).  NumPy arrays -- what the reference's loops pass after `.cpu()` -- take
    the scipy path like the reference; CUDA tensors are scored on the device (egaze_aae_auc) and only four numbers per
    sample come back, so a validation loop no longer has to ship both maps to the host."""
    try:
        import torch
        if isinstance(output, torch.Tensor) and output.is_cuda:
            from egaze import ops
            res = ops.aae_auc(output, target if isinstance(target, torch.Tensor) else torch.as_tensor(target, device=output.device))
            r = res.cpu().numpy()
            gp = [[int(a), int(b)] for a, b in r[:, 2:4]]
            if r.shape[0] == 1 and output.dim() == 2:
                return float(r[0, 0]), float(r[0, 1]), gp
            return float(np.mean(r[:, 0])), float(np.mean(r[:, 1])), gp
    except ImportError:
        pass
    if output.ndim == 3:
        res = [_aae_auc_single(output[b].squeeze(), target[b].squeeze()) for b in range(output.shape[0])]
        return np.mean([r[0] for r in res]), np.mean([r[1] for r in res]), [r[2] for r in res]
    a, u, g = _aae_auc_single(output, target)
    return a, u, [g]


def plot_loss(train_loss, test_loss, save_path):
    try:
        import matplotlib
        matplotlib.use('agg')
        import matplotlib.pyplot as plt
    except ImportError:
        return
    plt.plot(train_loss)
    plt.plot(test_loss)
    plt.ylabel('loss')
    plt.xlabel('epoch')
    plt.legend(['train', 'test'], loc='upper right')
    plt.savefig(save_path)
    plt.close()
