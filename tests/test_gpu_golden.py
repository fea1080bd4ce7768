"""GPU: the CUDA path (module API -> C-ABI) against (a) the committed golden outputs of the UNMODIFIED reference and
(b) the NumPy oracle on the same seeded inputs.  Gate: gaze maps within 1e-3 max-abs (BASELINE north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import egaze_oracle as orc
from test_oracle_golden import GOLD, sp_shapes, lf_shapes, lstm_shapes, vgg_shapes

pytestmark = pytest.mark.gpu


def load_sd(model, sd):
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return model


def make_sp(seed, dev):
    from utils import make_layers, cfg
    from models.model_SP import model_SP
    m = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20))
    return load_sd(m, orc.synth_state_dict(sp_shapes(), seed)).to(dev)


def test_model_sp_eval_vs_reference_golden(cuda_dev):
    g = np.load(os.path.join(GOLD, "sp_eval_b2_s64.npz"))
    m = make_sp(int(g["seed_w"]), cuda_dev).eval()
    x_s, x_t, _ = orc.synth_sp_inputs(int(g["B"]), int(g["S"]), int(g["seed_x"]))
    seen = []
    m._modules.get('features_s').register_forward_hook(lambda mod, i, o: seen.append(o))
    with torch.no_grad():
        y = m(torch.from_numpy(x_s).to(cuda_dev), torch.from_numpy(x_t).to(cuda_dev))
    assert np.abs(y.cpu().numpy() - g["y"]).max() <= 1e-3
    assert np.abs(seen[0].cpu().numpy() - g["f_s"]).max() <= 1e-3 * max(1.0, np.abs(g["f_s"]).max())


def test_model_sp_train_forward_vs_reference_golden(cuda_dev):
    g = np.load(os.path.join(GOLD, "sp_train_b4_s32.npz"))
    assert float(g["decoder_gain"]) == 1.0
    m = make_sp(int(g["seed_w"]), cuda_dev).train()
    x_s, x_t, gt = orc.synth_sp_inputs(int(g["B"]), int(g["S"]), int(g["seed_x"]))
    with torch.no_grad():
        y = m(torch.from_numpy(x_s).to(cuda_dev), torch.from_numpy(x_t).to(cuda_dev))
        from floss import floss
        loss = floss()(y, torch.from_numpy(gt).to(cuda_dev))
    for k, v in m.state_dict().items():
        if "running_" in k:
            assert np.abs(v.cpu().numpy() - g["buf/" + k]).max() <= 1e-4, k
    err = np.abs(y.cpu().numpy() - g["y"]).max()
    print("train-mode gaze map max-abs err vs the reference's golden output: %.3e" % err)
    assert err <= 1e-3   # north_star's bar, train mode included (fp32 itself is 2e-4 from fp64 here)
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))


def test_model_sp_eval_vs_oracle_224(cuda_dev):
    m = make_sp(3, cuda_dev).eval()
    x_s, x_t, _ = orc.synth_sp_inputs(1, 224, 99)
    with torch.no_grad():
        y = m(torch.from_numpy(x_s).to(cuda_dev), torch.from_numpy(x_t).to(cuda_dev)).cpu().numpy()
    sd = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    ref = orc.model_sp_forward(sd, x_s, x_t, training=False)[0]
    assert np.abs(y - ref).max() <= 1e-3


def test_config1_pipeline_vs_reference_golden(cuda_dev):
    """BASELINE config 1 (run_spatialstream.py plumbing) through the CUDA path."""
    from utils import make_layers, cfg
    from models.late_fusion import late_fusion
    from egaze.vgg import VGG
    from egaze import ops
    from scipy import ndimage
    g = np.load(os.path.join(GOLD, "config1_run_spatialstream.npz"))
    vgg = load_sd(VGG(make_layers(cfg['D'], 3), return_features=True), orc.synth_state_dict(vgg_shapes(), int(g["seed_vgg"])))
    vgg = vgg.to(cuda_dev).eval()
    lf = load_sd(late_fusion(), orc.synth_state_dict(lf_shapes(), int(g["seed_lf"]))).to(cuda_dev).eval()
    with torch.no_grad():
        out, feat = vgg(torch.from_numpy(g["x"]).to(cuda_dev))
        assert np.abs(out.cpu().numpy() - g["out"]).max() <= 1e-3
        assert np.abs(feat.cpu().numpy() - g["feat"]).max() <= 1e-3 * max(1.0, np.abs(g["feat"]).max())
        im = (out.squeeze().cpu().numpy() * 255).astype(np.uint8)
        predicted = ndimage.center_of_mass(im)
        assert np.abs(np.array(predicted) - g["predicted"]).max() <= 0.5
        gaze = [[int(g["predicted"][0]), int(g["predicted"][1])]]
        vec = ops.crop_mean(feat, gaze, 3)
        assert np.abs(vec.cpu().numpy()[0] - g["vec"]).max() <= 1e-3 * max(1.0, np.abs(g["vec"]).max())
        weighted = ops.weighted_map(vec, feat)
        assert np.abs(weighted.cpu().numpy() - g["weighted"][0]).max() <= 2e-3
        up = ops.bilinear_up(weighted.unsqueeze(1), 16, False)
        fin = lf(out, up)
    assert np.abs(fin.cpu().numpy() - g["fin"]).max() <= 1e-3


def test_late_fusion_vs_reference_golden(cuda_dev):
    from models.late_fusion import late_fusion
    g = np.load(os.path.join(GOLD, "lf_b2_s64.npz"))
    rs = np.random.RandomState(int(g["seed_x"]))
    f = torch.from_numpy(rs.rand(2, 1, 64, 64).astype(np.float32)).to(cuda_dev)
    gg = torch.from_numpy(rs.rand(2, 1, 64, 64).astype(np.float32)).to(cuda_dev)
    m = load_sd(late_fusion(), orc.synth_state_dict(lf_shapes(), int(g["seed_w"]))).to(cuda_dev).eval()
    with torch.no_grad():
        assert np.abs(m(f, gg).cpu().numpy() - g["y_eval"]).max() <= 1e-3
        m.train()
        assert np.abs(m(f, gg).cpu().numpy() - g["y_train"]).max() <= 1e-3
    for k, v in m.state_dict().items():
        if "running_" in k:
            assert np.abs(v.cpu().numpy() - g["buf/" + k]).max() <= 1e-4, k


def test_lstm_vs_reference_golden(cuda_dev):
    from models.LSTMnet import lstmnet
    g = np.load(os.path.join(GOLD, "lstm_t5_b3.npz"))
    net = load_sd(lstmnet(), orc.synth_state_dict(lstm_shapes(), int(g["seed_w"]))).to(cuda_dev).eval()
    rs = np.random.RandomState(int(g["seed_x"]))
    x = rs.randn(5, 3, 512).astype(np.float32)
    h0 = (rs.randn(2, 3, 512) * 0.3).astype(np.float32)
    c0 = (rs.randn(2, 3, 512) * 0.3).astype(np.float32)
    d = lambda a: torch.from_numpy(a).to(cuda_dev)
    with torch.no_grad():
        out, (hn, cn) = net(d(x), (d(h0), d(c0)))
        o1, (h1, c1) = net(d(x[:1, :1].copy()), None)
    assert np.abs(out.cpu().numpy() - g["out"]).max() <= 1e-5
    assert np.abs(hn.cpu().numpy() - g["hn"]).max() <= 1e-5 and np.abs(cn.cpu().numpy() - g["cn"]).max() <= 1e-5
    assert np.abs(o1.cpu().numpy() - g["out_none"]).max() <= 1e-5 and np.abs(h1.cpu().numpy() - g["hn_none"]).max() <= 1e-5


def test_floss_vs_reference_golden(cuda_dev):
    from floss import floss
    g = np.load(os.path.join(GOLD, "floss.npz"))
    _, _, gt = orc.synth_sp_inputs(3, 224, int(g["seed_gt"]))
    p = (1 / (1 + np.exp(-np.random.RandomState(int(g["seed_p"])).randn(3, 1, 224, 224) * 3))).astype(np.float32)
    pt = torch.from_numpy(p).to(cuda_dev).requires_grad_(True)
    fl = floss()
    loss = fl(pt, torch.from_numpy(gt).to(cuda_dev))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert np.abs(pt.grad.cpu().numpy()[:, :, ::7, ::7] - g["grad"]).max() <= 1e-5 * np.abs(g["grad"]).max()
    w = fl.build_weight_from_target(torch.from_numpy(gt).to(cuda_dev))
    assert np.abs(w[:, :, ::7, ::7] - g["w_blob"]).max() <= 1e-4


def test_at_glue_vs_reference_golden(cuda_dev):
    from egaze import ops
    g = np.load(os.path.join(GOLD, "at_glue.npz"))
    rs = np.random.RandomState(int(g["seed"]))
    feats = torch.from_numpy(np.maximum(rs.randn(4, 512, 14, 14), 0).astype(np.float32)).to(cuda_dev)
    vec = ops.crop_mean(feats, g["gazes"].tolist(), 3)
    assert np.abs(vec.cpu().numpy() - g["vec"]).max() <= 1e-6
    wm = ops.weighted_map(vec, feats)
    assert np.abs(wm.cpu().numpy() - g["map"]).max() <= 2e-5
    av = ops.crop_align_mean(feats, g["gazes"].tolist(), 3)
    assert np.abs(av.cpu().numpy() - g["align_vec"]).max() <= 1e-5


def test_metric_aae_auc_vs_reference_golden(cuda_dev):
    """Device computeAAEAUC (egaze_aae_auc) against the reference's scipy implementation run on the same seeded maps:
    gaze point exact, AAE within 1e-4 degrees, AUC within 2 pixels of 50176 (float32 vs float64 centre-of-mass sums can move
    int(predicted) across an integer boundary only in contrived cases; none here)."""
    from egaze import ops
    import utils as egaze_utils
    g = np.load(os.path.join(GOLD, "metric_aae_auc.npz"))
    mo, mt = orc.synth_metric_inputs(int(g["B"]), int(g["seed"]))
    o, t = torch.from_numpy(mo).to(cuda_dev), torch.from_numpy(mt).to(cuda_dev)
    res = ops.aae_auc(o, t).cpu().numpy()
    assert np.array_equal(res[:, 2:4].astype(np.int64), g["gp"])
    assert np.abs(res[:, 0] - g["aae"]).max() <= 1e-4
    assert np.abs(res[:, 1] - g["auc"]).max() <= 2.0 / (224 * 224)
    aae, auc, gp = egaze_utils.computeAAEAUC(o, t)          # the drop-in entry point with CUDA tensors
    assert abs(aae - float(g["batch_aae"])) <= 1e-4 and abs(auc - float(g["batch_auc"])) <= 2.0 / (224 * 224)
    assert gp == [list(map(int, r)) for r in g["gp"]]
    aae_h, auc_h, _ = egaze_utils.computeAAEAUC(mo, mt)     # NumPy arrays keep the reference's host path
    assert abs(aae_h - float(g["batch_aae"])) <= 1e-9 and abs(auc_h - float(g["batch_auc"])) <= 1e-12
