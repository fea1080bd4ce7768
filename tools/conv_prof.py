"""Per-role cycle accounting of the tcgen05 conv kernel (egaze_conv3x3_set_prof) for a few SP layers at B=32.

Prints, per layer, the average over CTAs of: cycles per work item, and the share of its time each role spent
waiting on each barrier (producer: A/B slot free; MMA issuer: accumulator free, A landed, B landed; epilogue:
accumulator ready)."""
import os; os.environ["EGAZE_CONV_PLANS"] = "0"  # the cycle counters are a launch-time setting
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200"))
import torch
from egaze import ops
from egaze._lib import call

LAYERS = [l for l in [  # name, N, H, W, Cin, Cout, kwargs
    ("trunk 64->64 @224 f32+stats", 32, 224, 224, 64, 64, dict(want_f32=True, want_split=False, stats=True)),
    ("dgrad 64->64 @224 f32 (2 MMA)", 32, 224, 224, 64, 64, dict(want_f32=True, want_split=False, grad=True)),
    ("dgrad 64->64 @224 mask bf16hi (2 MMA)", 32, 224, 224, 64, 64, dict(want_lo=False, grad=True, masked=True)),
    ("dgrad 128->128 @112 f32 (2 MMA)", 32, 112, 112, 128, 128, dict(want_f32=True, want_split=False, grad=True)),
    ("trunk 16->64 @224 f32+stats", 32, 224, 224, 3, 64, dict(want_f32=True, want_split=False, stats=True)),
    ("dec 64->64 @224 relu split", 32, 224, 224, 64, 64, dict(relu=True)),
    ("dec 128->64 @224 relu split", 32, 224, 224, 128, 64, dict(relu=True)),
    ("trunk 64->128 @112 f32+stats", 32, 112, 112, 64, 128, dict(want_f32=True, want_split=False, stats=True)),
    ("trunk 128->128 @112 f32+stats", 32, 112, 112, 128, 128, dict(want_f32=True, want_split=False, stats=True)),
    ("trunk 256->256 @56 f32+stats", 32, 56, 56, 256, 256, dict(want_f32=True, want_split=False, stats=True)),
    ("dec 512->256 @56 relu split", 32, 56, 56, 512, 256, dict(relu=True)),
    ("trunk 512->512 @28 f32+stats", 32, 28, 28, 512, 512, dict(want_f32=True, want_split=False, stats=True)),
    ("trunk 512->512 @14 f32+stats", 32, 14, 14, 512, 512, dict(want_f32=True, want_split=False, stats=True)),
] if not os.environ.get("PROF_ONLY") or os.environ["PROF_ONLY"] in l[0]]
prof = torch.zeros(160, 16, dtype=torch.int64, device="cuda")
for name, N, H, W, Ci, Co, kw in LAYERS:
    x = torch.randn(N, Ci, H, W, device="cuda")
    w = torch.randn(Co, Ci, 3, 3, device="cuda") * 0.02
    b = torch.zeros(Co, device="cuda")
    kw = dict(kw)
    grad, masked = kw.pop("grad", False), kw.pop("masked", False)
    if grad:   # the data-gradient operand mode: one bf16 plane of dY x bf16 hi+lo flipped weights
        act = ops.to_split(x, fmt=0)
        act.lo = None
        wp = ops.pack_cache.get(w, 1, cols_p=act.Cp)
        if masked:
            kw["mask"] = ops.to_split(torch.randn(N, Co, H, W, device="cuda"), fmt=0).hi
    else:
        act = ops.to_split(x)
        wp = ops.pack_cache.get(w, 0, cols_p=act.Cp)
    for _ in range(2):
        ops.conv3x3(act, wp, bias=b, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv3x3(act, wp, bias=b, **kw); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    prof.zero_()
    call("egaze_conv3x3_set_prof", prof)
    ops.conv3x3(act, wp, bias=b, **kw)
    torch.cuda.synchronize()
    call("egaze_conv3x3_set_prof", None)
    p = prof[prof[:, 9] > 0].double()
    items = p[:, 9].mean().item()
    pm = p[p[:, 6] > 0]          # CTAs whose MMA warp issued (rank 0 of every pair in CTA-pair mode)
    tot = pm[:, 6].mean().item()

    def f(c, d):
        q = p[p[:, d] > 0]
        return 100.0 * (q[:, c] / q[:, d]).mean().item()
    fl = 2.0 * N * H * W * Co * Ci * 9
    print("%-32s %.3f ms %6.1f TF/s | %5.1f items/CTA %7.0f clk/item | producer waits A-free %4.1f%% B-free %4.1f%% | "
          "MMA waits acc-free %4.1f%% A-landed %4.1f%% B-landed %4.1f%% inside the MMA issue blocks %4.1f%% | epilogue waits acc-ready %4.1f%%"
          % (name, ms, fl / ms / 1e9, items, tot / items, f(0, 2), f(1, 2), f(3, 6), f(4, 6), f(5, 6), f(15, 6), f(7, 8)))
    ep = [(p[:, 10 + i] / p[:, 9]).mean().item() for i in range(5)]
    print("    epilogue (thread 0) cycles per item: TMEM->smem %.0f | barrier %.0f | store loop %.0f | reductions %.0f | barrier %.0f"
          % tuple(ep))
