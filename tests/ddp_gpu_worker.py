"""Worker of tests/test_gpu_ddp.py (launched under torchrun, one rank per GPU, NCCL).

SURVEY 4.5: with BatchNorm on running statistics (per-sample independent forward) the gradients two ranks average over their
half-batches must equal the single-process gradients of the concatenated batch.  Checks the overlapped segmented reducer
(egaze.ddp.OverlappedGradReducer, driven from inside model_SP's backward node) and the round-1 flat bucket against the same
single-process reference, in eager mode and with the whole step captured in a CUDA graph."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "egocentric-gaze-prediction_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch
import torch.distributed as dist


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import torch_ref
    import floss as floss_mod
    from oracle import egaze_oracle as orc
    from utils import make_layers, cfg
    from models.model_SP import model_SP
    from egaze.ddp import FlatGradBucket, attach_reducer, broadcast_parameters
    from egaze.graph import GraphedStep
    B, S = 4 * world, 64
    torch.manual_seed(0)
    m = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20))
    torch_ref.randomize_(m, 0)
    m = m.to(dev).eval()                      # BatchNorm on running statistics; autograd stays on
    broadcast_parameters(m, 0)
    x_s, x_t, gt = [torch.from_numpy(a).to(dev) for a in orc.synth_sp_inputs(B, S, 21)]
    crit = floss_mod.floss()
    # single-process reference on the whole batch (every rank computes it: same kernels, same data)
    m.zero_grad(set_to_none=True)
    crit(m(x_s, x_t), gt).backward()
    ref = [p.grad.detach().clone() for p in m.parameters()]
    sl = slice(rank * B // world, (rank + 1) * B // world)
    xs, xt, g = x_s[sl].contiguous(), x_t[sl].contiguous(), gt[sl].contiguous()
    out = {}
    # (a) overlapped segmented reducer, eager
    red = attach_reducer(m)
    m.zero_grad(set_to_none=True)
    crit(m(xs, xt), g).backward()
    torch.cuda.synchronize()
    out["overlap_eager"] = max(rel_l2(p.grad, r) for p, r in zip(m.parameters(), ref))
    assert all(p.grad.data_ptr() == red.view(p).data_ptr() for p in m.parameters())
    out["segments"] = [n for n, _ in red.segments]
    # (b) the same step captured in a CUDA graph (the NCCL calls and the communication-stream fork / join included)
    def step(a, b, c):
        m.zero_grad(set_to_none=True)
        loss = crit(m(a, b), c)
        loss.backward()
        return loss
    gs = GraphedStep(step, [xs, xt, g], modules=[m], warmup=2)
    gs(xs, xt, g)
    torch.cuda.synchronize()
    out["overlap_graph"] = max(rel_l2(p.grad, r) for p, r in zip(m.parameters(), ref))
    gs.release()
    # (c) round-1 flat bucket
    del m._egaze_reducer
    bucket = FlatGradBucket(m.parameters(), dev)
    bucket.zero()
    crit(m(xs, xt), g).backward()
    bucket.allreduce()
    torch.cuda.synchronize()
    out["flat"] = max(rel_l2(p.grad, r) for p, r in zip(m.parameters(), ref))
    worst = torch.tensor([out["overlap_eager"], out["overlap_graph"], out["flat"]], device=dev, dtype=torch.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.update(overlap_eager=worst[0].item(), overlap_graph=worst[1].item(), flat=worst[2].item(), world=world)
        print("DDP_RESULT " + json.dumps(out), flush=True)
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
