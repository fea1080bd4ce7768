"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product path; see DESIGN.md "Oracle").

A NumPy restatement of the reference's hot-path algorithm (hyf015/egocentric-gaze-prediction @ 4c8ba58).  The
reference is pure Python on top of PyTorch (pinned "Pytorch v0.4.0", README.md:10; run here with torch 2.11.0):
the conv / batch-norm / LSTM / BCE arithmetic therefore lives in a third-party dependency that is not vendored.
This file restates that published arithmetic (SURVEY.md App. D lists the semantics relied on) and follows the
reference's own call sites, cited per function as file:line.

Pinning: the reference ships NO tests, golden vectors or fixtures.  The oracle is pinned against
outputs of the UNMODIFIED reference modules executed in the authoring container (oracle/make_golden.py imports them
read-only from /root/reference and writes tests/golden/*.npz); tests/test_oracle_golden.py checks every fixture.

All functions take/return NumPy arrays in the reference's NCHW layout.  `dtype` selects the accumulation precision
(np.float32 mirrors the reference, np.float64 gives a tighter "truth" for tolerance studies).
"""
import math

import numpy as np

CFG_D = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512]  # utils.py:60
# model_SP.decoder (models/model_SP.py:13-30): (Cin, Cout) convs, 'U' = nn.Upsample(scale_factor=2)
DEC_SP = [(512, 512), (512, 512), 'U', (512, 512), (512, 512), (512, 512), 'U', (512, 256), (256, 256), (256, 256), 'U',
          (256, 128), (128, 128), 'U', (128, 64), (64, 64)]
# script-local VGG decoder (spatialstream.py:71-91, temporalstream.py:70-90, run_spatialstream.py:23-43): 3 convs at 14x14
DEC_VGG = [(512, 512), (512, 512), (512, 512), 'U', (512, 512), (512, 512), (512, 512), 'U', (512, 256), (256, 256),
           (256, 256), 'U', (256, 128), (128, 128), 'U', (128, 64), (64, 64)]


# ---- elementary ops (PyTorch semantics) -------------------------------------------------------------
def conv2d(x, w, b=None, pad=1, dtype=np.float32):
    """nn.Conv2d, stride 1, square kernel k, zero padding `pad` (utils.py:70; model_SP.py:13-30; late_fusion.py:10-13)."""
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    N, C, H, W = x.shape
    Co, Ci, k, _ = w.shape
    assert Ci == C
    if k == 1:
        y = np.einsum('nchw,oc->nohw', x, w[:, :, 0, 0], optimize=True)
    else:
        xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
        Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
        wm = w.reshape(Co, Ci * k * k).T.copy()
        y = np.empty((N, Co, Ho, Wo), dtype=dtype)
        for n in range(N):  # per image keeps the im2col buffer small
            win = np.lib.stride_tricks.sliding_window_view(xp[n], (k, k), axis=(1, 2))  # [C,Ho,Wo,k,k]
            cols = win.transpose(1, 2, 0, 3, 4).reshape(Ho * Wo, Ci * k * k)
            y[n] = (cols @ wm).T.reshape(Co, Ho, Wo)
    if b is not None:
        y = y + np.asarray(b, dtype=dtype).reshape(1, -1, 1, 1)
    return y


def batchnorm2d(x, weight, bias, running_mean, running_var, training, momentum=0.1, eps=1e-5, dtype=np.float32):
    """nn.BatchNorm2d (utils.py:72; model_SP.py:12; late_fusion.py:10-12).  Train: biased batch var normalises,
    unbiased var goes into running_var.  Returns (y, new_running_mean, new_running_var)."""
    x = np.asarray(x, dtype=dtype)
    if training:
        n = x.shape[0] * x.shape[2] * x.shape[3]
        mean = x.mean(axis=(0, 2, 3), dtype=np.float64)
        var = x.var(axis=(0, 2, 3), dtype=np.float64)
        new_rm = ((1 - momentum) * running_mean + momentum * mean).astype(np.float32)
        new_rv = ((1 - momentum) * running_var + momentum * var * n / max(n - 1, 1)).astype(np.float32)
    else:
        mean, var = np.asarray(running_mean, np.float64), np.asarray(running_var, np.float64)
        new_rm, new_rv = running_mean, running_var
    inv = 1.0 / np.sqrt(var + eps)
    y = (x - mean.reshape(1, -1, 1, 1).astype(dtype)) * inv.reshape(1, -1, 1, 1).astype(dtype)
    y = y * np.asarray(weight, dtype).reshape(1, -1, 1, 1) + np.asarray(bias, dtype).reshape(1, -1, 1, 1)
    return y, new_rm, new_rv


def relu(x):
    return np.maximum(x, 0)


def maxpool2x2(x):
    """nn.MaxPool2d(2, 2) (utils.py:68)."""
    N, C, H, W = x.shape
    return x.reshape(N, C, H // 2, 2, W // 2, 2).max(axis=(3, 5))


def upsample_nearest2x(x):
    """nn.Upsample(scale_factor=2), default mode 'nearest' (model_SP.py:16,20,24,27)."""
    return x.repeat(2, axis=2).repeat(2, axis=3)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


# ---- networks ------------------------------------------------------------------------------------------------------
def trunk_forward(sd, prefix, x, training, new_buffers=None, cfg=CFG_D, dtype=np.float32):
    """utils.make_layers(cfg['D'], C_in) forward (utils.py:64-76): [conv3x3 -> BN -> ReLU] x13, pools after 2,4,7,10."""
    i = 0
    for v in cfg:
        if v == 'M':
            x = maxpool2x2(x)
            i += 1
            continue
        x = conv2d(x, sd[prefix + '%d.weight' % i], sd[prefix + '%d.bias' % i], dtype=dtype)
        b = prefix + '%d.' % (i + 1)
        x, rm, rv = batchnorm2d(x, sd[b + 'weight'], sd[b + 'bias'], sd[b + 'running_mean'], sd[b + 'running_var'],
                                training, dtype=dtype)
        if new_buffers is not None:
            new_buffers[b + 'running_mean'], new_buffers[b + 'running_var'] = rm, rv
        x = relu(x)
        i += 3
    return x


def decoder_forward(sd, prefix, x, spec, dtype=np.float32):
    """nn.Sequential decoder: conv3x3+ReLU blocks with nearest upsamples, then the 1x1 conv to one channel."""
    i = 0
    for v in spec:
        if v == 'U':
            x = upsample_nearest2x(x)
            i += 1
            continue
        x = relu(conv2d(x, sd[prefix + '%d.weight' % i], sd[prefix + '%d.bias' % i], dtype=dtype))
        i += 2
    return conv2d(x, sd[prefix + '%d.weight' % i], sd[prefix + '%d.bias' % i], pad=0, dtype=dtype)


def model_sp_forward(sd, x_s, x_t, training, dtype=np.float32):
    """models/model_SP.py:35-50.  Returns (gaze map, features_s output, features_t output, updated BN buffers)."""
    nb = {}
    f_s = trunk_forward(sd, 'features_s.', x_s, training, nb, dtype=dtype)
    f_t = trunk_forward(sd, 'features_t.', x_t, training, nb, dtype=dtype)
    # Conv3d(512,512,(1,3,3),pad (0,1,1)) over the stacked (x_s, x_t) + MaxPool3d((2,1,1))  (model_SP.py:38-44)
    # == the same 3x3 conv on each stream followed by an elementwise max
    w = np.asarray(sd['fusion.weight'])[:, :, 0]
    c_s = conv2d(f_s, w, sd['fusion.bias'], dtype=dtype)
    c_t = conv2d(f_t, w, sd['fusion.bias'], dtype=dtype)
    x = np.maximum(c_s, c_t)
    x, rm, rv = batchnorm2d(x, sd['bn.weight'], sd['bn.bias'], sd['bn.running_mean'], sd['bn.running_var'], training,
                            dtype=dtype)
    nb['bn.running_mean'], nb['bn.running_var'] = rm, rv
    x = relu(x)
    x = decoder_forward(sd, 'decoder.', x, DEC_SP, dtype=dtype)
    return sigmoid(x).astype(np.float32), f_s, f_t, nb


def vgg_forward(sd, x, training, dtype=np.float32):
    """script-local single-stream VGG.forward (spatialstream.py:97-101; run_spatialstream.py:49-53 also returns conv5_3)."""
    nb = {}
    f = trunk_forward(sd, 'features.', x, training, nb, dtype=dtype)
    y = sigmoid(decoder_forward(sd, 'decoder.', f, DEC_VGG, dtype=dtype)).astype(np.float32)
    return y, f, nb


def late_fusion_forward(sd, f, g, training, dtype=np.float32):
    """models/late_fusion.py:18-23: cat(f, g) -> [conv3x3+BN+ReLU]x3 -> conv1x1 -> sigmoid."""
    nb = {}
    x = np.concatenate((f, g), axis=1)
    for i in (0, 3, 6):
        x = conv2d(x, sd['fusion.%d.weight' % i], sd['fusion.%d.bias' % i], dtype=dtype)
        b = 'fusion.%d.' % (i + 1)
        x, rm, rv = batchnorm2d(x, sd[b + 'weight'], sd[b + 'bias'], sd[b + 'running_mean'], sd[b + 'running_var'],
                                training, dtype=dtype)
        nb[b + 'running_mean'], nb[b + 'running_var'] = rm, rv
        x = relu(x)
    x = conv2d(x, sd['fusion.9.weight'], sd['fusion.9.bias'], pad=0, dtype=dtype)
    return sigmoid(x).astype(np.float32), nb


def lstmnet_forward(sd, x, h0, c0, dtype=np.float32):
    """models/LSTMnet.py:26-37: tanh -> nn.LSTM(512,512,2) (gates i,f,g,o) -> Linear(512,512) -> ReLU.
    x [T,B,512]; h0,c0 [2,B,512].  Returns (out [T,B,512], h_n, c_n)."""
    x = np.tanh(np.asarray(x, dtype))
    h = [np.asarray(h0[l], dtype).copy() for l in range(2)]
    c = [np.asarray(c0[l], dtype).copy() for l in range(2)]
    outs = []
    for t in range(x.shape[0]):
        inp = x[t]
        for l in range(2):
            gates = (inp @ np.asarray(sd['lstm.weight_ih_l%d' % l], dtype).T + np.asarray(sd['lstm.bias_ih_l%d' % l], dtype) +
                     h[l] @ np.asarray(sd['lstm.weight_hh_l%d' % l], dtype).T + np.asarray(sd['lstm.bias_hh_l%d' % l], dtype))
            i, f, g, o = np.split(gates, 4, axis=1)
            c[l] = sigmoid(f) * c[l] + sigmoid(i) * np.tanh(g)
            h[l] = sigmoid(o) * np.tanh(c[l])
            inp = h[l]
        outs.append(relu(inp @ np.asarray(sd['lin.weight'], dtype).T + np.asarray(sd['lin.bias'], dtype)))
    return np.stack(outs).astype(np.float32), np.stack(h).astype(np.float32), np.stack(c).astype(np.float32)


# ---- floss ---------------------------------------------------------------------------------------------------------
def floss_weight(target):
    """floss.build_weight_from_target (floss.py:15-41): float64 arithmetic, stored into a float32 array."""
    target = np.asarray(target, np.float32)
    B, W = target.shape[0], target.shape[-1]
    out = np.empty_like(target)
    for b in range(B):
        t = target[b].squeeze()
        xs, ys = np.where(t == t.max())
        cx, cy = xs.mean(), ys.mean()
        a = (np.arange(W) - cx)[:, None]
        c = (np.arange(W) - cy)[None, :]
        out[b, 0] = 1.0 / ((np.sqrt(a ** 2 + c ** 2) + 1) / W)
    return out


def floss_loss(inp, target):
    """floss.forward (floss.py:9-13) = F.binary_cross_entropy(input, target, weight): logs clamped at -100, mean."""
    w = floss_weight(target).astype(np.float64)
    p = np.asarray(inp, np.float32)
    t = np.asarray(target, np.float32)
    with np.errstate(divide='ignore'):
        lp = np.maximum(np.log(p), -100.0)
        lq = np.maximum(np.log1p(-p), -100.0)
    return float(np.mean(-w * (t * lp.astype(np.float64) + (1 - t) * lq.astype(np.float64))))


def floss_grad(inp, target, grad_out=1.0):
    """BCE backward: w * (p - t) / max(p(1-p), 1e-12) / numel."""
    w = floss_weight(target)
    p = np.asarray(inp, np.float32)
    t = np.asarray(target, np.float32)
    return (grad_out * w * (p - t) / np.maximum((1 - p) * p, 1e-12) / p.size).astype(np.float32)


# ---- AT glue -------------------------------------------------------------------------------------------------------
def crop_feature(feature, maxind, size=3):
    """AT.crop_feature (AT.py:25-39): per-sample size x size crop around clip(gaze // 16, size//2, H - ceil(size/2))."""
    H = feature.shape[2]
    res = []
    for b in range(feature.shape[0]):
        fmax = np.array(maxind[b]) // 16
        fmax = np.clip(fmax, size // 2, H - int(math.ceil(size / 2.0))).astype(np.int64)  # run_spatialstream.py:93 int()
        lo, hi = size // 2, int(math.ceil(size / 2.0))
        res.append(feature[b, :, fmax[0] - lo:fmax[0] + hi, fmax[1] - lo:fmax[1] + hi])
    return np.stack(res)


def crop_mean(feature, maxind, size=3):
    """AT.py:239-241: mean over the crop -> channel weights [B,512]."""
    c = crop_feature(feature, maxind, size)
    return c.reshape(c.shape[0], c.shape[1], -1).mean(axis=2, dtype=np.float32)


def crop_align_mean(feature, maxind, size=3):
    """AT.crop_align_feature (AT.py:41-56) + mean (AT.py:239-241): bilinear x16 (align_corners=True), 48x48 crop around
    clip(gaze, 24, 224-24), mean -> [B,512]."""
    H = 224
    win = size * 16
    up = bilinear_upsample(feature, 16, align_corners=True)
    res = []
    for b in range(feature.shape[0]):
        fmax = np.clip(np.array(maxind[b]), win // 2, H - win // 2).astype(np.int64)
        c = up[b, :, fmax[0] - win // 2:fmax[0] + win // 2, fmax[1] - win // 2:fmax[1] + win // 2]
        res.append(c.reshape(c.shape[0], -1).mean(axis=1, dtype=np.float32))
    return np.stack(res)


def get_weighted(chn_weight, feature):
    """AT.get_weighted (AT.py:58-66) applied per sample (the reference only ever passes batch 1)."""
    out = []
    for b in range(feature.shape[0]):
        m = (feature[b] * np.asarray(chn_weight[b]).reshape(-1, 1, 1)).sum(axis=0, dtype=np.float32)
        m = m - m.min()
        out.append(m / m.max())
    return np.stack(out).astype(np.float32)


def aae_auc(output, target):
    """utils.computeAAEAUC (utils.py:96-140) for one 224x224 prediction / target pair -> (AAE degrees, AUC, [i, j]).
    scipy does the centre of mass and the sigma-14 Gaussian filter exactly as the reference does."""
    from scipy import ndimage
    predicted = ndimage.center_of_mass(output)
    i, j = np.unravel_index(target.argmax(), target.shape)
    d = 112 / math.tan(math.pi / 6)
    r1 = np.array([predicted[0] - 112, predicted[1] - 112, d])
    r2 = np.array([i - 112, j - 112, d])
    angle = math.degrees(math.atan2(np.linalg.norm(np.cross(r1, r2)), np.dot(r1, r2)))
    z = np.zeros((224, 224))
    z[int(predicted[0])][int(predicted[1])] = 1
    z = ndimage.gaussian_filter(z, 14)
    z = z - np.min(z)
    z = z / np.max(z)
    auc = 1 - float((z > z[i][j]).sum()) / (output.shape[0] * output.shape[1])
    return angle, auc, [int(i), int(j)]


def synth_metric_inputs(B, seed=21):
    """Seeded (prediction, target) pairs for the metric tests: smooth blobs whose centres sweep the image, including the
    borders (the Gaussian filter reflects there), plus quantised targets with arg-max plateaus."""
    rng = np.random.RandomState(seed)
    ys, xs = np.mgrid[0:224, 0:224].astype(np.float64)
    outs, tgts = [], []
    for b in range(B):
        cy, cx = rng.uniform(0, 223, 2) if b % 3 else rng.choice([2.0, 221.0, 30.0, 200.0], 2)
        o = np.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / (2 * rng.uniform(4, 40) ** 2)) + 0.02 * rng.rand(224, 224)
        ty, tx = rng.uniform(0, 223, 2)
        t = np.exp(-((ys - ty) ** 2 / (2 * 16.3 ** 2) + (xs - tx) ** 2 / (2 * 12.25 ** 2)))
        t = np.round(255 * (t - t.min()) / (t.max() - t.min())) / 255
        outs.append(o.astype(np.float32))
        tgts.append(t.astype(np.float32))
    return np.stack(outs), np.stack(tgts)


def bilinear_upsample(x, scale, align_corners=False):
    """F.upsample(mode='bilinear') == align_corners=False (run_spatialstream.py:136); upsample_bilinear == True (AT.py:46)."""
    x = np.asarray(x, np.float32)
    h, w = x.shape[-2:]
    H, W = h * scale, w * scale

    def src(o, n_in, n_out):
        if align_corners:
            return o * (n_in - 1) / (n_out - 1) if n_out > 1 else np.zeros_like(o)
        return np.maximum((o + 0.5) / scale - 0.5, 0.0)

    sy = src(np.arange(H, dtype=np.float32), h, H).astype(np.float32)
    sx = src(np.arange(W, dtype=np.float32), w, W).astype(np.float32)
    y0 = np.minimum(sy.astype(np.int64), h - 1)
    x0 = np.minimum(sx.astype(np.int64), w - 1)
    y1 = np.minimum(y0 + 1, h - 1)
    x1 = np.minimum(x0 + 1, w - 1)
    ly = (sy - y0).astype(np.float32)[:, None]
    lx = (sx - x0).astype(np.float32)[None, :]
    g = lambda yy, xx: x[..., yy[:, None], xx[None, :]]
    return ((1 - ly) * ((1 - lx) * g(y0, x0) + lx * g(y0, x1)) + ly * ((1 - lx) * g(y1, x0) + lx * g(y1, x1))).astype(np.float32)


# ---- deterministic synthetic parameters / inputs (shared by golden generation, tests, bench) -------------------------
def synth_state_dict(shapes, seed=0, decoder_gain=1.0):
    """Deterministic, box-independent parameters for a {key: shape} description (insertion-ordered).
    decoder_gain scales the BN-less decoder conv weights: 1.0 is He init (random-init logits of std ~8, saturated
    sigmoid, ill-conditioned gradients: stock fp32 is itself ~1e-2 away from fp64 there); ~0.8 keeps logits O(1)."""
    import zlib
    sd = {}
    for k, shp in shapes.items():
        rs = np.random.RandomState((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))
        shp = tuple(shp)
        if k.endswith('num_batches_tracked'):
            sd[k] = np.zeros((), np.int64)
        elif k.endswith('running_mean'):
            sd[k] = (rs.randn(*shp) * 0.1).astype(np.float32)
        elif k.endswith('running_var'):
            sd[k] = (rs.rand(*shp) + 0.5).astype(np.float32)
        elif len(shp) >= 3:  # conv weight
            fan = int(np.prod(shp[1:]))
            gain = decoder_gain if k.startswith('decoder.') else 1.0
            sd[k] = (rs.randn(*shp) * (gain * math.sqrt(2.0 / fan))).astype(np.float32)
        elif len(shp) == 2:  # LSTM / Linear weight
            sd[k] = (rs.uniform(-1, 1, shp) / math.sqrt(shp[1])).astype(np.float32)
        elif 'lstm' in k or k.startswith('lin.'):
            sd[k] = (rs.uniform(-1, 1, shp) / math.sqrt(512)).astype(np.float32)
        elif k.endswith('.weight'):  # BN gamma
            sd[k] = (rs.rand(*shp) + 0.5).astype(np.float32)
        else:  # conv bias / BN beta
            sd[k] = (rs.randn(*shp) * 0.05).astype(np.float32)
    return sd


def synth_sp_inputs(B, S, seed=1234):
    """SURVEY 8(d) config-2 inputs: normalised-image-scale RGB, [-1,1] quantised flow stack, quantised Gaussian gaze blob."""
    rs = np.random.RandomState(seed)
    x_s = rs.randn(B, 3, S, S).astype(np.float32)
    x_t = ((rs.randint(0, 256, (B, 20, S, S)).astype(np.float32) / 255 - 0.5) / 0.5).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(S, dtype=np.float64), np.arange(S, dtype=np.float64), indexing='ij')
    gt = np.empty((B, 1, S, S), np.float32)
    for b in range(B):
        cy, cx = (rs.rand(2) * 0.7 + 0.15) * S
        g = np.exp(-((ys - cy) ** 2 / (2 * (S * 16.3 / 224) ** 2) + (xs - cx) ** 2 / (2 * (S * 12.25 / 224) ** 2)))
        g = (g - g.min()) / (g.max() - g.min())
        gt[b, 0] = np.round(g * 255) / 255
    return x_s, x_t, gt
