"""Launch a few representative conv layers of the SP stack (B=32) for `ncu --set full` captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200"))
import torch
from egaze import ops

LAYERS = [  # (name, N, H, W, Cin, Cout)
    ("dec5_512x256_56", 32, 56, 56, 512, 256),
    ("trunk_64x64_224", 32, 224, 224, 64, 64),
    ("trunk_512x512_28", 32, 28, 28, 512, 512),
    ("dec10_128x64_224", 32, 224, 224, 128, 64),
]
reps = int(os.environ.get("REPS", 2))
for name, N, H, W, Ci, Co in LAYERS:
    x = torch.randn(N, Ci, H, W, device="cuda")
    w = torch.randn(Co, Ci, 3, 3, device="cuda") * 0.02
    b = torch.zeros(Co, device="cuda")
    act = ops.to_split(x)
    wp = ops.pack_cache.get(w, 0, cols_p=act.Cp)
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv3x3(act, wp, bias=b, relu=True)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fl = 2.0 * N * H * W * Co * Ci * 9
    print("%-20s %.3f ms  %.1f TFLOP/s (%s)" % (name, ms, fl / ms / 1e9, ops.precision()))
