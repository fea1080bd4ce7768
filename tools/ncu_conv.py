"""Launch a few representative tcgen05 layers of the SP stack (B=32) for `ncu --set full` captures: fprop with the
train-mode (fp32 + BN statistics) and decoder (ReLU, split) epilogues, a masked dgrad, and two weight gradients."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200"))
import torch
from egaze import ops

LAYERS = [  # (name, kind, N, H, W, Cin, Cout)
    ("dec5_512x256_56", "dec", 32, 56, 56, 512, 256),
    ("trunk_64x64_224", "trunk", 32, 224, 224, 64, 64),
    ("trunk_512x512_28", "trunk", 32, 28, 28, 512, 512),
    ("dec10_128x64_224", "dec", 32, 224, 224, 128, 64),
    ("wgrad_256x256_56", "wgrad", 32, 56, 56, 256, 256),
    ("wgrad_64x64_224", "wgrad", 32, 224, 224, 64, 64),
]
reps = int(os.environ.get("REPS", 2))
for name, kind, N, H, W, Ci, Co in LAYERS:
    x = torch.randn(N, Ci, H, W, device="cuda")
    w = torch.randn(Co, Ci, 3, 3, device="cuda") * 0.02
    b = torch.zeros(Co, device="cuda")
    act = ops.to_split(x)
    if kind == "wgrad":
        dy = ops.to_split(torch.randn(N, Co, H, W, device="cuda"))
        run = lambda: ops.wgrad3x3(act, dy, Co, Ci)
    elif kind == "trunk":
        wp = ops.pack_cache.get(w, 0, cols_p=act.Cp)
        run = lambda: ops.conv3x3(act, wp, bias=b, want_f32=True, want_split=False, stats=True)
    else:
        wp = ops.pack_cache.get(w, 0, cols_p=act.Cp)
        run = lambda: ops.conv3x3(act, wp, bias=b, relu=True)
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fl = 2.0 * N * H * W * Co * Ci * 9
    print("%-20s %.3f ms  %.1f TFLOP/s algorithmic (%s)" % (name, ms, fl / ms / 1e9, ops.precision()))
