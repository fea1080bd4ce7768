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
    ref = torch.cat([torch.nn.functional.pad(p.grad.reshape(-1), (0, (-p.numel()) % 4)) for p in m.parameters()])
    assert torch.allclose(res["flat"], ref, rtol=1e-5, atol=1e-7)
    assert res["seeds"] == [1234, 1235]


class _Bag(object):
    """The gradient bag of egaze.autograd (gradients by parameter identity)."""

    def __init__(self):
        self.d = {}

    def put(self, p, g):
        self.d[id(p)] = g

    def get(self, p):
        return self.d.get(id(p))


def _reducer_worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200"))
    from egaze.ddp import OverlappedGradReducer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.zeros(s)) for s in ((4, 3), (5,), (2, 2, 2), (7,), (3,))]
        ps[3].requires_grad_(False)                      # frozen: not part of the buffer at all
        # the segments deliberately do NOT follow the parameters' registration order; "rest" repeats them all
        r = OverlappedGradReducer([("late", [ps[2], ps[4]]), ("early", [ps[0], ps[1], ps[3]]), ("rest", ps)])
        assert [n for n, _ in r.segments] == ["late", "early"] and r.numel == 8 + 4 + 12 + 8
        results = []
        for step in range(2):
            bag = _Bag()
            for i, p in enumerate(ps):
                if p.requires_grad and not (step == 1 and i == 4):      # step 1: one parameter gets no gradient
                    bag.put(p, torch.full(p.shape, float((rank + 1) * (i + 1) * (step + 1))))
            r.begin()
            r.reduce(bag, "late")
            assert bag.get(ps[0]).data_ptr() != r.view(ps[0]).data_ptr()      # not reduced yet
            r.reduce(bag, "late")                                               # idempotent within a step
            r.reduce_rest(bag)
            r.finish()
            for p in ps:
                if p.requires_grad:
                    assert bag.get(p).data_ptr() == r.view(p).data_ptr() and bag.get(p).shape == p.shape
            results.append([bag.get(p).clone() if p.requires_grad else None for p in ps])
        if rank == 0:
            torch.save({"results": results, "calls": r.calls}, out)
    finally:
        dist.destroy_process_group()


def test_overlapped_reducer_segments_average_and_alias(tmp_path):
    """egaze.ddp.OverlappedGradReducer on CPU / gloo: segment-ordered flat layout, averaged values, gradients replaced by
    views of the buffer, a parameter without gradient contributes zeros, frozen parameters are left out."""
    out = str(tmp_path / "red.pt")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_reducer_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["calls"] == 4
    for step, grads in enumerate(res["results"]):
        for i, g in enumerate(grads):
            if i == 3:
                assert g is None
            elif step == 1 and i == 4:
                assert float(g.abs().max()) == 0.0
            else:
                assert torch.allclose(g, torch.full_like(g, 1.5 * (i + 1) * (step + 1)))


def test_sp_segments_cover_every_parameter_once():
    sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200"))
    from egaze.ddp import OverlappedGradReducer, sp_segments, DEEP_FROM
    from utils import make_layers, cfg
    from models.model_SP import model_SP
    m = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20))
    r = OverlappedGradReducer(sp_segments(m), device=torch.device("cpu"))
    assert sorted(id(p) for p in r.params) == sorted(id(p) for p in m.parameters())
    names = [n for n, _ in r.segments]
    assert names == ["decoder", "fusion_bn", "trunk_deep", "trunk_shallow"]      # "rest" is empty: nothing was missed
    sizes = {n: sum(p.numel() for p in ps) for n, ps in r.segments}
    assert sizes["trunk_deep"] > 0.95 * (sizes["trunk_deep"] + sizes["trunk_shallow"]) and DEEP_FROM == 4
    assert r.numel == sum((p.numel() + 3) // 4 * 4 for p in m.parameters())
