"""GPU: the device-side input pipeline (egaze.data) against the reference dataset's CPU arithmetic
(data/STdatas.py:50-73) and the streaming SP -> AT -> LF pipeline object (egaze.pipeline) against the same steps
composed by hand / the reference's uint8 hand-over (AT.py:228-252)."""
import numpy as np
import pytest
import torch

import torch_ref

pytestmark = pytest.mark.gpu


def test_image_norm_bit_identical(cuda_dev):
    from egaze import data
    rs = np.random.RandomState(0)
    im = rs.randint(0, 256, (3, 37, 53, 3)).astype(np.uint8)                      # cv2.imread layout: H, W, BGR
    got = data.normalize_image(torch.from_numpy(im).to(cuda_dev)).cpu()
    for n in range(3):                                                             # data/STdatas.py:51-55, verbatim
        t = torch.from_numpy(im[n].transpose((2, 0, 1)))
        t = t.float().div(255)
        t = t.sub_(torch.FloatTensor([0.485, 0.456, 0.406]).view(3, 1, 1)).div_(torch.FloatTensor([0.229, 0.224, 0.225]).view(3, 1, 1))
        assert torch.equal(got[n], t)


def test_flow_window_matches_reference_stack(cuda_dev):
    """A video walked in order: after each push the window's stack equals the tensor the reference builds from its per-sample list
    [x_n, y_n, x_{n-1}, y_{n-1}, ..., x_{n-9}, y_{n-9}] (data/STdatas.py:18-20,59-68), bit for bit."""
    from egaze import data
    V, H, W, F = 2, 20, 28, 14
    rs = np.random.RandomState(1)
    fx = rs.randint(0, 256, (F, V, H, W)).astype(np.uint8)
    fy = rs.randint(0, 256, (F, V, H, W)).astype(np.uint8)
    win = data.FlowWindow(V, H, W, frames=10, device=cuda_dev)
    for n in range(F):
        win.push(torch.from_numpy(fx[n]).to(cuda_dev), torch.from_numpy(fy[n]).to(cuda_dev))
        got = win.stack().cpu()
        assert got.shape == (V, 20, H, W)
        for v in range(V):
            arr = []
            for k in range(10):
                src = max(n - k, 0)                       # before the tenth frame the oldest available frame repeats
                arr += [torch.from_numpy(fx[src, v]), torch.from_numpy(fy[src, v])]
            ref = torch.stack(arr).float().div_(255).sub_(0.5).div_(0.5)      # STdatas.py:64-68
            assert torch.equal(got[v], ref), (n, v)


def test_jpeg_decode_close_to_cv2(cuda_dev):
    cv2 = pytest.importorskip("cv2")
    from egaze import data
    ys, xs = np.mgrid[0:224, 0:224].astype(np.float32)
    img = np.stack([127 + 100 * np.sin(xs / 23.0), 127 + 100 * np.cos(ys / 31.0), 0.5 * (xs + ys)], -1).clip(0, 255).astype(np.uint8)
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 95])
    assert ok
    raw = enc.tobytes()
    try:
        w, h, c = data.jpeg_info(raw)
        got = data.decode_jpeg(raw, device=cuda_dev).cpu().numpy()
        got_gray = data.decode_jpeg(raw, gray=True, device=cuda_dev).cpu().numpy()
    except RuntimeError as exc:
        if "nvJPEG is not available" in str(exc):
            pytest.skip("no nvJPEG on this box")
        raise
    assert (w, h, c) == (224, 224, 3)
    ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
    ref_gray = cv2.imdecode(enc, cv2.IMREAD_GRAYSCALE)
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    dg = np.abs(got_gray.astype(np.int32) - ref_gray.astype(np.int32))
    print("nvJPEG vs libjpeg: BGR max %d mean %.3f | gray max %d mean %.3f" % (d.max(), d.mean(), dg.max(), dg.mean()))
    assert d.max() <= 6 and d.mean() <= 1.0          # IDCT / chroma-upsampling differences between the two decoders
    assert dg.max() <= 3 and dg.mean() <= 0.6
    x = data.normalize_image(torch.from_numpy(got).to(cuda_dev))
    assert x.shape == (1, 3, 224, 224)


@pytest.mark.parametrize("quantize", [False, True])
def test_gaze_pipeline_matches_hand_composition(cuda_dev, quantize):
    import models.LSTMnet as L
    from models.late_fusion import late_fusion
    from utils import make_layers, cfg
    from models.model_SP import model_SP
    from egaze import ops
    from egaze.pipeline import GazePipeline
    from oracle import egaze_oracle as orc
    torch.manual_seed(0)
    sp = torch_ref.randomize_(model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20)), 0).to(cuda_dev).eval()
    lstm = L.lstmnet().to(cuda_dev).eval()
    lf = torch_ref.randomize_(late_fusion(), 1).to(cuda_dev).eval()
    pipe = GazePipeline(sp, lstm, lf, quantize=quantize)
    B, S, T = 3, 64, 4
    hidden = (torch.zeros(2, B, 512, device=cuda_dev), torch.zeros(2, B, 512, device=cuda_dev))
    g = torch.Generator().manual_seed(3)
    for t in range(T):
        x_s, x_t, gt = [torch.from_numpy(a).to(cuda_dev) for a in orc.synth_sp_inputs(B, S, 40 + t)]
        fixsac = (torch.rand(B, generator=g) < 0.4).int().to(cuda_dev)
        res = pipe.step(x_s, x_t, fixsac=fixsac, target=gt)
        # the same frame by hand, on the module API
        seen = []
        hk = sp.features_s.register_forward_hook(lambda m, i, o: seen.append(o))
        with torch.no_grad():
            out = sp(x_s, x_t)
            hk.remove()
            feat = seen[0]
            out_q = out
            if quantize:      # AT.py:228-230: np.uint8(255 * outim), read back as uint8 / 255 (lateDataset.py:26-31)
                out_q = torch.from_numpy(np.uint8(255 * out.cpu().numpy()).astype(np.float32) / 255).to(cuda_dev)
            gaze = torch.stack([(gt[b, 0] == gt[b, 0].max()).nonzero().float().mean(0).floor() for b in range(B)]).int()
            vec = ops.crop_mean(feat, gaze, 3)
            o, (h, c) = lstm(vec.unsqueeze(0), hidden)
            sac = fixsac != 1
            w = torch.where(sac[:, None], o.squeeze(0), vec)
            hidden = (torch.where(sac[None, :, None], h, hidden[0]), torch.where(sac[None, :, None], c, hidden[1]))
            at = ops.weighted_map(w, feat)
            at_q = at
            if quantize:
                at_q = torch.from_numpy(np.uint8(255 * at.cpu().numpy()).astype(np.float32) / 255).to(cuda_dev)
            fused = lf(ops.bilinear_up(at_q.unsqueeze(1), 16, False), out_q)
        assert torch.equal(res["gaze"].cpu(), gaze.cpu())
        assert torch.equal(res["sp"], out) and torch.equal(res["at"], at)
        assert torch.equal(res["fused"], fused), (t, (res["fused"] - fused).abs().max().item())
    pipe.close()
