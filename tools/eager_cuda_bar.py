"""Same-box GPU bar: the reference's module structure executed by stock PyTorch eager CUDA (cuDNN) on the
B200, fp32 with TF32 off (the arithmetic the parity gate is defined on) and with TF32 on (what most users would run; it
does NOT meet the 1e-3 gate).  Measurement only -- nothing here is on the product path.

    python tools/eager_cuda_bar.py [--batch 32] [--steps 5]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import torch_cpu_ref as ref
from oracle import egaze_oracle as orc

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda")
x_s, x_t, gt = [torch.from_numpy(a).to(dev) for a in orc.synth_sp_inputs(args.batch, 224, 1234)]
# floss weight map built ONCE outside the timed loop (the reference rebuilds it on the host every step, floss.py:15-41;
# leaving that out favours this bar)
w_floss = torch.from_numpy(orc.floss_weight(gt.cpu().numpy())).to(dev)
out = {}
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    for mode in ("train", "fwd"):
        torch.manual_seed(0)
        m = ref.ModelSP().to(dev)
        if mode == "train":
            m.train()
            opt = torch.optim.Adam(m.parameters(), lr=1e-7)

            def step():
                opt.zero_grad(set_to_none=True)
                loss = torch.nn.functional.binary_cross_entropy(m(x_s, x_t), gt.view(-1, 1, 224, 224), weight=w_floss.view(-1, 1, 224, 224))
                loss.backward()
                opt.step()
        else:
            m.eval()

            def step():
                with torch.no_grad():
                    m(x_s, x_t)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out["%s_%s" % (mode, "tf32" if tf32 else "fp32")] = {"ms_per_step": ms, "fps": args.batch / ms * 1e3}
        del m
        torch.cuda.empty_cache()
print(json.dumps({"eager_cuda_bar": out, "batch": args.batch, "torch": torch.__version__,
                  "cudnn": torch.backends.cudnn.version()}))
