"""CPU, world_size 2, gloo: the data-parallel host logic (flat gradient bucket + one all-reduce, rank sharding).
The averaged gradients of two half-batches must equal the single-process gradient of the whole batch (the loss is a
mean over samples; BatchNorm-free model)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 1, 1), torch.nn.Sigmoid())


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200"))
    from egaze.ddp import FlatGradBucket, shard_seed, broadcast_parameters
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _model()
        if rank == 1:  # replicas that drifted apart are re-synchronised from rank 0
            with torch.no_grad():
                for p in m.parameters():
                    p.add_(1.0)
        broadcast_parameters(m, 0)
        bucket = FlatGradBucket(m.parameters())
        g = torch.Generator().manual_seed(7)
        x = torch.randn(8, 3, 16, 16, generator=g)
        t = torch.rand(8, 1, 16, 16, generator=g)
        xs, ts = x[rank * 4:(rank + 1) * 4], t[rank * 4:(rank + 1) * 4]
        for step in range(2):  # second step checks that zero() drops the old gradients (no accumulation)
            bucket.zero()
            loss = torch.nn.functional.binary_cross_entropy(m(xs), ts)
            loss.backward()
            flat = bucket.allreduce().clone()
            for p, v in zip(bucket.params, bucket.views):   # the optimiser reads the averaged values through .grad
                assert p.grad.data_ptr() == v.data_ptr()
        if rank == 0:
            torch.save({"flat": flat, "seeds": [shard_seed(1234, r) for r in range(world)]}, out)
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_full_batch(tmp_path):
    out = str(tmp_path / "res.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    m = _model()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 3, 16, 16, generator=g)
    t = torch.rand(8, 1, 16, 16, generator=g)
    torch.nn.functional.binary_cross_entropy(m(x), t).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    assert torch.allclose(res["flat"], ref, rtol=1e-5, atol=1e-7)
    assert res["seeds"] == [1234, 1235]
