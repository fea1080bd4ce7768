"""Launch representative wgrad layers (B=32) for ncu captures / quick timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200"))
import torch
from egaze import ops
LAYERS = [("w_256x256_56", 32, 56, 56, 256, 256), ("w_64x64_224", 32, 224, 224, 64, 64), ("w_512x512_28", 32, 28, 28, 512, 512),
          ("w_128x128_112", 32, 112, 112, 128, 128)]
reps = int(os.environ.get("REPS", 2))
for name, N, H, W, Ci, Co in LAYERS:
    xa = ops.to_split(torch.randn(N, Ci, H, W, device="cuda"))
    dya = ops.to_split(torch.randn(N, Co, H, W, device="cuda"))
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dwp = torch.zeros((9, Co, Ci), device="cuda")
        e0.record()
        ops.call("egaze_wgrad3x3_tc", xa.hi, xa.lo, dya.hi, dya.lo, N, H, W, Ci, Co, dwp, 1, 0, ops.stream_ptr())
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%-16s %.3f ms  %.1f TFLOP/s" % (name, ms, 2.0 * N * H * W * Co * Ci * 9 / ms / 1e9))
