"""How much of the gradient mismatch is inherent fp32 noise?  Compare stock fp32, stock fp64 and egaze gradients."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
import torch_ref
from oracle import egaze_oracle as orc
from egaze import ops
import floss as floss_mod
from utils import make_layers, cfg
from models.model_SP import model_SP
dev = torch.device("cuda")
def rel(a, b): return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

# per-kernel actual errors
g = torch.Generator().manual_seed(0)
for (N, H, W, Ci, Co) in [(2, 28, 28, 256, 256), (2, 56, 56, 64, 64)]:
    x = torch.randn(N, Ci, H, W, generator=g).to(dev); w = (torch.randn(Co, Ci, 3, 3, generator=g) * 0.03).to(dev)
    dy = torch.randn(N, Co, H, W, generator=g).to(dev)
    xa = ops.to_split(x); wp = ops.pack_cache.get(w, 0, cols_p=xa.Cp)
    _, y, _ = ops.conv3x3(xa, wp, want_f32=True, want_split=False)
    ref = F.conv2d(x.double(), w.double(), padding=1)
    print("fwd   %s rel-L2 vs fp64: egaze %.2e  torch-fp32 %.2e" % ((N, H, W, Ci, Co), rel(ops.nhwc_f32_to_nchw(y), ref), rel(F.conv2d(x, w, padding=1), ref)))
    dya = ops.to_split(dy)
    gw = ops.wgrad3x3(xa, dya, Co, Ci)
    refw = torch.nn.grad.conv2d_weight(x.double(), (Co, Ci, 3, 3), dy.double(), padding=1)
    print("wgrad rel-L2 vs fp64: egaze %.2e  torch-fp32 %.2e" % (rel(gw, refw), rel(torch.nn.grad.conv2d_weight(x, (Co, Ci, 3, 3), dy, padding=1), refw)))

torch.manual_seed(0)
m = torch_ref.randomize_(model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20)), 0)
with torch.no_grad():
    for mod in m.decoder:
        if isinstance(mod, torch.nn.Conv2d): mod.weight.mul_(float(os.environ.get('GAIN', '0.8')))
m = m.to(dev).train()
m32 = copy.deepcopy(m); m64 = copy.deepcopy(m).double()
x_s, x_t, gt = [torch.from_numpy(a).to(dev) for a in orc.synth_sp_inputs(2, 64, 5)]
out = m(x_s, x_t); loss = floss_mod.floss()(out, gt); loss.backward()
o32 = torch_ref.model_sp_forward(m32, x_s, x_t); l32 = torch_ref.floss_loss(o32, gt); l32.backward()
o64 = torch_ref.model_sp_forward(m64, x_s.double(), x_t.double())
w64 = torch_ref.floss_weight(gt).double()
l64 = F.binary_cross_entropy(o64, gt.double(), weight=w64); l64.backward()
print("loss egaze %.6f fp32 %.6f fp64 %.6f" % (loss.item(), l32.item(), l64.item()))
print("out max-abs: egaze-fp64 %.2e  fp32-fp64 %.2e" % ((out.double() - o64).abs().max().item(), (o32.double() - o64).abs().max().item()))
rows = []
for (k, p), (_, q), (_, r) in zip(m.named_parameters(), m32.named_parameters(), m64.named_parameters()):
    if r.grad.norm().item() < 1e-9: continue
    rows.append((k, rel(p.grad, r.grad), rel(q.grad, r.grad)))
for k, a, b in rows:
    if k.endswith('weight') and ('decoder' in k or 'fusion' in k or k.startswith('bn') or k in ('features_s.40.weight','features_s.0.weight','features_t.20.weight')): print("%-28s egaze-vs-fp64 %.2e   fp32-vs-fp64 %.2e" % (k, a, b))
import statistics
print("median egaze %.2e  median fp32 %.2e" % (statistics.median(r[1] for r in rows), statistics.median(r[2] for r in rows)))
