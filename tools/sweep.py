"""BASELINE configs[4]: synthetic-video sweep, batch {8,16,32,64} x size {224,288}, on this rank's GPU.
For every point: device-timed fps of the SP train step and of the eval forward, and the tcgen05 conv kernels' achieved
algorithmic TFLOP/s against the measured sustained bf16 peak (MEASURED_PEAKS.json).  One JSON line per point.

    python tools/sweep.py [--batches 8,16,32,64] [--sizes 224,288] [--steps 5]
(run it under torchrun with --gpus N for the multi-GPU columns; every rank then runs the same per-GPU batch: weak scaling)"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200")); sys.path.insert(0, ROOT)
import torch
import bench
from egaze import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="8,16,32,64")
ap.add_argument("--sizes", default="224,288")
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
peaks, _ = bench.load_peaks()
peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
FLOP = {224: (114.167e9, 341.171e9), 288: (188.725e9, 188.725e9 * 341.171 / 114.167)}   # SURVEY 8d (fwd, fwd+bwd per frame)
for S in [int(v) for v in args.sizes.split(",")]:
    for B in [int(v) for v in args.batches.split(",")]:
        row = {"size": S, "batch": B}
        for name in ("sp_fwd", "sp_train"):
            wl = bench.Workload(name, B, S, 0, 1, dev)
            for _ in range(8):      # new shapes: launch plans, allocator growth and the pack cache settle within a few steps
                wl.step(*wl.dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                wl.step(*wl.dev)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            # kernel durations in a second pass on one stream (see bench.py: overlapping event intervals double-count)
            os.environ["EGAZE_WGRAD_STREAM"] = "0"
            os.environ["EGAZE_TRUNK_STREAM"] = "0"
            wl.step(*wl.dev)
            ops.conv_timer_reset(True)
            for _ in range(args.steps):
                wl.step(*wl.dev)
            conv_ms, _ = ops.conv_timer_read()
            ops.conv_timer_reset(False)
            os.environ.pop("EGAZE_WGRAD_STREAM", None)
            os.environ.pop("EGAZE_TRUNK_STREAM", None)
            fl = FLOP[S][0 if name == "sp_fwd" else 1] * B
            row[name] = {"ms_per_step": round(ms, 3), "fps": round(B / ms * 1e3, 1),
                         "conv_tflops": round(fl * args.steps / (conv_ms * 1e-3) / 1e12, 1),
                         "conv_roofline_frac": round(fl * args.steps / (conv_ms * 1e-3) / 1e12 / peak, 3)}
            del wl
            torch.cuda.empty_cache()
        print(json.dumps(row), flush=True)
