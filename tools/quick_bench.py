"""Quick device-timed probe of the SP forward (not the contract bench)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200"))
import torch
from utils import make_layers, cfg
from models.model_SP import model_SP

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B = int(os.environ.get("B", 32)); S = 224
m = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20)).cuda().eval()
x_s = torch.randn(B, 3, S, S, device="cuda"); x_t = torch.randn(B, 20, S, S, device="cuda")
for mode in ("precise", "fast"):
    os.environ["EGAZE_PRECISION"] = mode
    for train in (False, True):
        m.train(train)
        with torch.no_grad():
            ms = timeit(lambda: m(x_s, x_t))
        print("model_SP fwd %s %s B=%d: %.2f ms  %.0f fps  %.1f TFLOP/s algorithmic" % (mode, "train-BN" if train else "eval", B, ms, B / ms * 1e3, 114.167e9 * B / ms / 1e9))
