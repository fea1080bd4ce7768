"""Plain PyTorch fp32 references for the floating-point kernels (test infrastructure only).
They run the STOCK torch ops over the same parameter containers the egaze modules hold."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def seq_forward(seq, x):
    """Stock nn.Sequential semantics over an egaze container (bypasses the fused forward)."""
    for m in seq.children():
        x = m(x)
    return x


def model_sp_forward(model, x_s, x_t, return_feats=False):
    """reference models/model_SP.py:35-50 semantics with stock ops."""
    f_s = seq_forward(model.features_s, x_s)
    f_t = seq_forward(model.features_t, x_t)
    x = torch.cat((f_s.unsqueeze(2), f_t.unsqueeze(2)), 2)
    x = model.pool3d(model.fusion(x)).squeeze(2)
    x = model.relu(model.bn(x))
    x = seq_forward(model.decoder, x)
    y = torch.sigmoid(x)
    return (y, f_s, f_t) if return_feats else y


def late_fusion_forward(model, f, g):
    return torch.sigmoid(seq_forward(model.fusion, torch.cat((f, g), 1)))


def lstmnet_forward(model, x, h0, c0):
    out, hid = model.lstm(torch.tanh(x), (h0, c0))
    return F.relu(model.lin(out)), hid


def floss_weight(target):
    """reference floss.py:15-41 in torch/fp64."""
    B, _, H, W = target.shape
    out = torch.empty_like(target)
    for b in range(B):
        t = target[b, 0]
        idx = (t == t.max()).nonzero().double()
        cx, cy = idx[:, 0].mean(), idx[:, 1].mean()
        a = torch.arange(W, dtype=torch.float64, device=target.device)
        dist = torch.sqrt((a[:, None] - cx) ** 2 + (a[None, :] - cy) ** 2)
        out[b, 0] = (1.0 / ((dist + 1) / W)).float()
    return out


def floss_loss(inp, target):
    return F.binary_cross_entropy(inp, target, weight=floss_weight(target))


def crop_mean(feat, gaze, size=3):
    """reference AT.py:25-39 + mean (AT.py:239-241)."""
    B, C, H, W = feat.shape
    out = []
    for b in range(B):
        fr, fc = int(gaze[b][0]) // 16, int(gaze[b][1]) // 16
        lo, hi = size // 2, H - (size + 1) // 2
        fr, fc = min(max(fr, lo), hi), min(max(fc, lo), hi)
        crop = feat[b, :, fr - size // 2:fr + (size + 1) // 2, fc - size // 2:fc + (size + 1) // 2]
        out.append(crop.reshape(C, -1).mean(1))
    return torch.stack(out)


def get_weighted(w, feat):
    """reference AT.py:58-66 applied per sample."""
    out = []
    for b in range(feat.shape[0]):
        m = (feat[b:b + 1] * w[b].view(1, -1, 1, 1)).sum(1)
        m = m - m.min()
        out.append(m / m.max())
    return torch.cat(out)


def randomize_(model, seed=0):
    """Non-trivial parameters/buffers everywhere (biases, BN affine + running stats) so every path is exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d)):
                fan = m.weight[0].numel()
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan) ** 0.5)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    return model
