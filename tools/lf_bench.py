#!/usr/bin/env python
"""CUDA-event timing of the late-fusion kernels (egaze_lf_fwd / egaze_lf_bwd) at the LF.trainLate shape (B x 1 x 224 x 224)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "egocentric-gaze-prediction_b200"), ROOT):
    sys.path.insert(0, p)
import torch


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    from models.late_fusion import late_fusion
    from egaze import ops
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for B in (int(os.environ.get("LF_B", "32")), 64):
        f = torch.rand(B, 1, 224, 224, device=dev)
        g = torch.rand(B, 1, 224, 224, device=dev)
        gout = torch.randn(B, 1, 224, 224, device=dev) * 1e-5
        m = late_fusion().to(dev)
        for precise in (True, False):
            m.eval()
            t_eval = timeit(lambda: ops.lf_forward(m.fusion, f, g, precise=precise))
            m.train()
            t_fwd = timeit(lambda: ops.lf_forward(m.fusion, f, g, keep=True, precise=precise))
            _, saved = ops.lf_forward(m.fusion, f, g, keep=True, precise=precise)
            t_bwd = timeit(lambda: ops.lf_backward(m.fusion, saved, gout))
            flop = 1.2147e9 * B
            print("B=%d precise=%d: eval fwd %.3f ms | train fwd %.3f ms (%.1f TFLOP/s alg) | bwd %.3f ms (%.1f TFLOP/s alg) | "
                  "train step %.3f ms" % (B, precise, t_eval, t_fwd, flop / t_fwd / 1e9, t_bwd, 2 * flop / t_bwd / 1e9, t_fwd + t_bwd))


if __name__ == "__main__":
    main()
