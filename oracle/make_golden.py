#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (read-only import from /root/reference) on CPU.

Run in the authoring container only (`python oracle/make_golden.py`); /root/reference does not exist on the GPU box,
so the fixtures are committed.  Inputs and parameters are regenerated from seeds by oracle.egaze_oracle.synth_*; the
fixtures hold only the reference's OUTPUTS (plus the seeds / shapes used).

Reference driving follows SURVEY.md 8(c): `skimage` / `matplotlib` are stubbed (neither is touched on the hot path),
modules are constructed directly, script-local classes are extracted with `ast` (the scripts run argparse /
os.listdir / torch.load at import time).
"""
import ast
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("EGAZE_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
from oracle import egaze_oracle as orc  # noqa: E402


def import_reference():
    for name in ("skimage", "skimage.io", "skimage.transform", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.pyplot"].switch_backend = lambda *a, **k: None
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import utils as ref_utils
    import floss as ref_floss
    from models import model_SP as ref_sp, LSTMnet as ref_lstm, late_fusion as ref_lf
    import AT as ref_at
    return ref_utils, ref_floss, ref_sp, ref_lstm, ref_lf, ref_at


def extract_defs(path, names, ns):
    """exec only the named ClassDef/FunctionDef nodes of a reference script inside namespace ns."""
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in names]
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def load_synth(model, seed, decoder_gain=1.0):
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = orc.synth_state_dict(shapes, seed, decoder_gain)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return sd


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    ref_utils, ref_floss, ref_sp, ref_lstm, ref_lf, ref_at = import_reference()
    make_layers, cfg = ref_utils.make_layers, ref_utils.cfg

    # ---- model_SP eval forward (B=2, 64x64) + features_s hook output -------------------------------------------------
    m = ref_sp.model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20))
    load_synth(m, 11)
    m.eval()
    x_s, x_t, gt = orc.synth_sp_inputs(2, 64, 1234)
    blobs = []
    h = m._modules.get('features_s').register_forward_hook(lambda mod, i, o: blobs.append(o))
    with torch.no_grad():
        y = m(t(x_s), t(x_t))
    h.remove()
    np.savez_compressed(os.path.join(OUT, "sp_eval_b2_s64.npz"), seed_w=11, seed_x=1234, B=2, S=64, y=y.numpy(),
                        f_s=blobs[0].numpy())

    # ---- model_SP train step (B=4, 32x32): forward, floss, backward, BN buffers ----------------------------------------
    # two regimes: He-init decoder (the reference's own init: saturated, ill-conditioned) and a damped decoder (logits O(1))
    for tag, seed_w, gain in (("sp_train_b4_s32", 12, 1.0), ("sp_train_wellcond_b4_s32", 17, 0.8)):
        m = ref_sp.model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20))
        load_synth(m, seed_w, gain)
        m.train()
        x_s, x_t, gt = orc.synth_sp_inputs(4, 32, 77)
        crit = ref_floss.floss()
        out = m(t(x_s), t(x_t))
        loss = crit(out, t(gt))
        loss.backward()
        rec = dict(seed_w=seed_w, seed_x=77, B=4, S=32, decoder_gain=gain, y=out.detach().numpy(), loss=float(loss))
        for k, v in m.state_dict().items():
            if "running_" in k:
                rec["buf/" + k] = v.numpy()
        for k, p in m.named_parameters():
            g = p.grad.numpy()
            rec["gnorm/" + k] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
            rec["ghead/" + k] = g.reshape(-1)[:64].copy()
            if g.size <= 4096:
                rec["grad/" + k] = g
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), **rec)

    # ---- config 1: run_spatialstream.py plumbing on one synthetic 224x224 frame ---------------------------
    ns = {"__name__": "ref_run_spatialstream"}
    exec("from utils import *\nimport numpy as np, torch, math\nimport torch.nn as nn", ns)
    extract_defs(os.path.join(REF, "run_spatialstream.py"), {"VGG", "crop_feature1", "get_weighted", "totensor", "toim"}, ns)
    vgg = ns["VGG"](make_layers(cfg['D'], 3))
    load_synth(vgg, 13)
    vgg.eval()
    lf = ref_lf.late_fusion()
    load_synth(lf, 14)
    lf.eval()
    im = (np.random.RandomState(0).rand(224, 224, 3) * 255).astype(np.uint8)
    imt = ns["totensor"](im.copy())
    from scipy import ndimage
    import contextlib, io
    with torch.no_grad():
        out, feat = vgg(imt)
        outim = ns["toim"](out)
        predicted = ndimage.center_of_mass(outim)
        with contextlib.redirect_stdout(io.StringIO()):
            vec = ns["crop_feature1"](feat, predicted, 3)
        vec = vec.contiguous().view(vec.size(0), vec.size(1), -1)
        vec = torch.mean(vec, 2).squeeze()
        weighted = ns["get_weighted"](vec, feat)
        weighted_up = torch.nn.functional.interpolate(weighted, scale_factor=16, mode='bilinear', align_corners=False)
        fin = lf(out, weighted_up)
    np.savez_compressed(os.path.join(OUT, "config1_run_spatialstream.npz"), seed_vgg=13, seed_lf=14, im=im,
                        x=imt.numpy(), out=out.numpy(), feat=feat.numpy(), predicted=np.array(predicted), vec=vec.numpy(),
                        weighted=weighted.numpy(), fin=fin.numpy())

    # ---- late_fusion eval + train forward (B=2, 64x64) -----------------------------------------------------------------
    rs = np.random.RandomState(5)
    f = rs.rand(2, 1, 64, 64).astype(np.float32)
    g = rs.rand(2, 1, 64, 64).astype(np.float32)
    lf = ref_lf.late_fusion()
    load_synth(lf, 15)
    lf.eval()
    with torch.no_grad():
        y_eval = lf(t(f), t(g)).numpy()
    lf.train()
    out = lf(t(f), t(g))
    _, _, gt = orc.synth_sp_inputs(2, 64, 6)
    loss = ref_floss.floss()(out, t(gt))
    loss.backward()
    rec = dict(seed_w=15, seed_x=5, seed_gt=6, y_eval=y_eval, y_train=out.detach().numpy(), loss=float(loss))
    for k, v in lf.state_dict().items():
        if "running_" in k:
            rec["buf/" + k] = v.numpy()
    for k, p in lf.named_parameters():
        rec["grad/" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "lf_b2_s64.npz"), **rec)

    # ---- lstmnet: explicit hidden (T=5,B=3), hidden=None at batch 1 ---------------------------------------------------
    net = ref_lstm.lstmnet()
    load_synth(net, 16)
    net.eval()
    rs = np.random.RandomState(7)
    x = rs.randn(5, 3, 512).astype(np.float32)
    h0 = (rs.randn(2, 3, 512) * 0.3).astype(np.float32)
    c0 = (rs.randn(2, 3, 512) * 0.3).astype(np.float32)
    with torch.no_grad():
        o, (hn, cn) = net(t(x), (t(h0), t(c0)))
        o1, (hn1, cn1) = net(t(x[:1, :1]), None)
    np.savez_compressed(os.path.join(OUT, "lstm_t5_b3.npz"), seed_w=16, seed_x=7, out=o.numpy(), hn=hn.numpy(), cn=cn.numpy(),
                        out_none=o1.numpy(), hn_none=hn1.numpy(), cn_none=cn1.numpy())

    # ---- floss: KATs + blob batch ---------------------------------------------------------------------------------------
    crit = ref_floss.floss()
    tt = np.zeros((1, 1, 224, 224), np.float32)
    tt[0, 0, 100, 50] = 1
    w_single = crit.build_weight_from_target(t(tt))
    tt[0, 0, 101, 50] = 1
    w_plateau = crit.build_weight_from_target(t(tt))
    _, _, gt = orc.synth_sp_inputs(3, 224, 21)
    p = (1 / (1 + np.exp(-np.random.RandomState(22).randn(3, 1, 224, 224) * 3))).astype(np.float32)
    pt = t(p).requires_grad_(True)
    loss = crit(pt, t(gt))
    loss.backward()
    w_blob = crit.build_weight_from_target(t(gt))
    np.savez_compressed(os.path.join(OUT, "floss.npz"), w_single_peak=np.float64(w_single[0, 0, 100, 50]),
                        w_plateau=np.array([w_plateau[0, 0, 100, 50], w_plateau[0, 0, 101, 50]], np.float64),
                        seed_gt=21, seed_p=22, loss=float(loss), grad=pt.grad.numpy()[:, :, ::7, ::7].copy(),
                        w_blob=w_blob[:, :, ::7, ::7].copy())

    # ---- AT glue: crop_feature + mean + get_weighted (batch 1 as in the reference, several gaze points) -----------------
    rs = np.random.RandomState(9)
    feats = np.maximum(rs.randn(4, 512, 14, 14), 0).astype(np.float32)
    gazes = np.array([[0, 0], [223, 223], [100, 37], [15, 208]])
    vecs, maps, avecs = [], [], []
    import warnings
    for b in range(4):
        c = ref_at.crop_feature(t(feats[b:b + 1]), [list(gazes[b])], 3).contiguous()
        v = torch.mean(c.view(1, 512, -1), 2)
        vecs.append(v.numpy())
        maps.append(ref_at.get_weighted(v, t(feats[b:b + 1])).numpy())
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ca = ref_at.crop_align_feature(t(feats[b:b + 1]), [list(gazes[b])], 3).contiguous()
        avecs.append(torch.mean(ca.view(1, 512, -1), 2).numpy())
    # ---- validation metric (utils.computeAAEAUC) on seeded prediction / target pairs
    mo, mt = orc.synth_metric_inputs(12, 21)
    rows = [ref_utils.computeAAEAUC(mo[b], mt[b]) for b in range(mo.shape[0])]
    batch = ref_utils.computeAAEAUC(mo, mt)
    np.savez_compressed(os.path.join(OUT, "metric_aae_auc.npz"), seed=21, B=12, aae=np.array([r[0] for r in rows]),
                        auc=np.array([r[1] for r in rows]), gp=np.array([r[2][0] for r in rows]),
                        batch_aae=np.float64(batch[0]), batch_auc=np.float64(batch[1]))
    np.savez_compressed(os.path.join(OUT, "at_glue.npz"), seed=9, gazes=gazes, vec=np.concatenate(vecs), map=np.concatenate(maps),
                        align_vec=np.concatenate(avecs))
    print("golden fixtures written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print("  %-36s %8d bytes" % (fn, os.path.getsize(os.path.join(OUT, fn))))


if __name__ == "__main__":
    main()
