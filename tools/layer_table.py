"""Per-launch timing table of the tcgen05 conv kernels for one SP train step (B=32, 224x224)."""
import os, sys
os.environ["EGAZE_WGRAD_STREAM"] = "0"   # per-kernel durations: keep every launch on one stream (no overlap)
os.environ["EGAZE_TRUNK_STREAM"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200")); sys.path.insert(0, ROOT)
import torch
from egaze import ops
from bench import Workload
B = int(os.environ.get("B", 32))
wl = Workload(os.environ.get("WL", "sp_train"), B, 224, 0, 1, torch.device("cuda"))
for _ in range(3): wl.step(*wl.dev)
torch.cuda.synchronize()
ops.conv_timer_reset(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); wl.step(*wl.dev); e1.record(); torch.cuda.synchronize()
tab = ops.conv_timer_table(); ops.conv_timer_reset(False)
tot = 0.0
for d, ms in tab:
    kind, N, H, W, Ci, Co, red, ups, st = d
    # algorithmic FLOPs of the reference's direct convolution: a sub-pixel launch (low-resolution H x W listed) stands for the
    # 3x3 conv on the 2H x 2W upsampled map
    fl = 2.0 * N * H * W * Ci * Co * 9 * (4 if "sub" in kind else 1)
    tot += ms
    print("%-10s N=%2d %3dx%-3d %3d->%-3d red=%d ups=%d stats=%d  %7.3f ms  %6.1f TFLOP/s(padded)" % (kind, N, H, W, Ci, Co, red, ups, st, ms, fl / ms / 1e9))
print("timed kernels %.2f ms of step %.2f ms (%s)" % (tot, e0.elapsed_time(e1), ops.precision()))
