"""Launch representative layers of the SP stack (B=32) and the late-fusion kernels for `ncu --set full` captures, in the default
numeric mode: fp16-split forward (train-mode fp32 + BN-statistics epilogue, decoder ReLU / split epilogue), the 2-MMA data
gradient (fp32 out and masked bf16 out), the 1-MMA weight gradient, one LF train step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200"))
import torch
from egaze import ops

LAYERS = [  # (name, kind, N, H, W, Cin, Cout)
    ("fwd_trunk_64x64_224", "trunk", 32, 224, 224, 64, 64),
    ("fwd_trunk_512x512_28", "trunk", 32, 28, 28, 512, 512),
    ("fwd_trunk_512x512_14", "trunk", 32, 14, 14, 512, 512),
    ("fwd_dec_512x256_56", "dec", 32, 56, 56, 512, 256),
    ("fwd_dec_128x64_224", "dec", 32, 224, 224, 128, 64),
    ("dgrad_f32_64x64_224", "dgrad", 32, 224, 224, 64, 64),
    ("dgrad_mask_64x64_224", "dgrad_mask", 32, 224, 224, 64, 64),
    ("dgrad_f32_512x512_28", "dgrad", 32, 28, 28, 512, 512),
    ("wgrad_64x64_224", "wgrad", 32, 224, 224, 64, 64),
    ("wgrad_512x512_28", "wgrad", 32, 28, 28, 512, 512),
]
reps = int(os.environ.get("REPS", 2))
for name, kind, N, H, W, Ci, Co in LAYERS:
    x = torch.randn(N, Ci, H, W, device="cuda")
    w = torch.randn(Co, Ci, 3, 3, device="cuda") * 0.02
    b = torch.zeros(Co, device="cuda")
    if kind == "wgrad":
        act = ops.to_split(x, xb=True)
        dy = ops.grad_split(torch.randn(N, Co, H, W, device="cuda"))
        run = lambda: ops.wgrad3x3(act, dy, Co, Ci)
    elif kind == "trunk":
        act = ops.to_split(x)
        wp = ops.pack_cache.get(w, 0, cols_p=act.Cp)
        run = lambda: ops.conv3x3(act, wp, bias=b, want_f32=True, want_split=False, stats=True)
    elif kind == "dec":
        act = ops.to_split(x)
        wp = ops.pack_cache.get(w, 0, cols_p=act.Cp)
        run = lambda: ops.conv3x3(act, wp, bias=b, relu=True, xb=True)
    else:
        dy = ops.grad_split(torch.randn(N, Co, H, W, device="cuda"))
        wp = ops.pack_cache.get(w, 1, cols_p=dy.Cp)
        if kind == "dgrad":
            run = lambda: ops.conv3x3(dy, wp, want_f32=True, want_split=False)
        else:
            mask = ops.to_split(torch.randn(N, Ci, H, W, device="cuda")).hi
            cs = torch.zeros(Ci, device="cuda")
            run = lambda: ops.conv3x3(dy, wp, mask=mask, colsum=cs, want_lo=ops.mode()["dy_lo"])
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fl = 2.0 * N * H * W * Co * Ci * 9
    print("%-24s %.3f ms  %.1f TFLOP/s algorithmic (%s)" % (name, ms, fl / ms / 1e9, ops.precision()))

from models.late_fusion import late_fusion
m = late_fusion().cuda().train()
f = torch.rand(32, 1, 224, 224, device="cuda")
g = torch.rand(32, 1, 224, 224, device="cuda")
gout = torch.randn(32, 1, 224, 224, device="cuda") * 1e-5
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, saved = ops.lf_forward(m.fusion, f, g, keep=True)
    ops.lf_backward(m.fusion, saved, gout)
    e1.record(); torch.cuda.synchronize()
print("lf_train_step_b32        %.3f ms" % e0.elapsed_time(e1))
