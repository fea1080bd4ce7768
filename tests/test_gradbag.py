"""Host logic of the backward's zero pool (egaze.autograd._GradBag): slices are disjoint, 16-byte aligned, zero, shaped like the
parameter, and the pool falls back to a plain allocation when it is exhausted or on another device.  CPU only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "egocentric-gaze-prediction_b200"))


def test_zero_pool_slices():
    from egaze.autograd import _GradBag
    bag = _GradBag().reserve("cpu")
    p = torch.nn.Parameter(torch.ones(5, 3))
    a = bag.zeros_like(p)
    b = bag.zeros(7, "cpu")
    c = bag.zeros(1, "cpu")
    assert a.shape == p.shape and b.shape == (7,) and c.shape == (1,)
    assert float(a.abs().sum() + b.abs().sum() + c.abs().sum()) == 0.0
    ptrs = [t.data_ptr() for t in (a, b, c)]
    assert all(ptr % 16 == 0 for ptr in ptrs)
    a.add_(1.0); b.add_(2.0); c.add_(3.0)           # disjoint: writing one never shows in another
    assert float(a.sum()) == 15.0 and float(b.sum()) == 14.0 and float(c.sum()) == 3.0
    bag.put(p, a)
    assert bag.get(p) is not None and bag.get(p).shape == p.shape


def test_zero_pool_exhaustion_and_dtype():
    from egaze.autograd import _GradBag
    bag = _GradBag().reserve("cpu")
    big = bag.zeros(_GradBag.POOL, "cpu")            # takes the whole pool
    more = bag.zeros(8, "cpu")                       # falls back to its own allocation
    assert big.numel() == _GradBag.POOL and more.numel() == 8 and float(more.abs().sum()) == 0.0
    assert not (big.data_ptr() <= more.data_ptr() < big.data_ptr() + 4 * big.numel())
    h = torch.nn.Parameter(torch.ones(4, dtype=torch.float64))
    assert _GradBag().zeros_like(h).dtype == torch.float64      # no pool reserved, other dtype: plain zeros
