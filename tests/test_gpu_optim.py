"""GPU: egaze.optim.Adam (fused multi-tensor Adam that also rewrites the packed conv-weight copies) against
torch.optim.Adam on the same parameters and gradients (the reference's optimiser: SP.py:110-113, LF.py:77)."""
import copy

import pytest
import torch

import torch_ref

pytestmark = pytest.mark.gpu


def _make_sp(dev, seed=0):
    from utils import make_layers, cfg
    from models.model_SP import model_SP
    torch.manual_seed(seed)
    m = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20))
    torch_ref.randomize_(m, seed)
    return m.to(dev).train()


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_fused_adam_matches_torch_adam(cuda_dev, wd):
    """Three steps on identical gradients: parameters and optimiser state agree to fp32 rounding; state_dicts are interchangeable."""
    from egaze.optim import Adam
    m_a = _make_sp(cuda_dev)
    m_b = copy.deepcopy(m_a)
    opt_a = Adam(m_a.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    opt_b = torch.optim.Adam(m_b.parameters(), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    g = torch.Generator(device="cpu").manual_seed(3)
    for _ in range(3):
        for pa, pb in zip(m_a.parameters(), m_b.parameters()):
            gr = (torch.randn(pa.shape, generator=g) * 1e-2).to(cuda_dev)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        opt_a.step()
        opt_b.step()
    for (k, pa), (_, pb) in zip(m_a.named_parameters(), m_b.named_parameters()):
        assert (pa - pb).abs().max().item() <= 2e-6 * max(1.0, pb.abs().max().item()), k
        sa, sb = opt_a.state[pa], opt_b.state[pb]
        assert float(sa["step"]) == float(sb["step"]) == 3.0
        assert (sa["exp_avg"] - sb["exp_avg"]).abs().max().item() <= 4e-6 * max(1e-6, sb["exp_avg"].abs().max().item()), k
        assert (sa["exp_avg_sq"] - sb["exp_avg_sq"]).abs().max().item() <= 4e-6 * max(1e-12, sb["exp_avg_sq"].abs().max().item()), k
    # state_dict round trip in both directions
    opt_c = torch.optim.Adam(m_a.parameters(), lr=1e-3)
    opt_c.load_state_dict(opt_a.state_dict())
    opt_d = Adam(m_b.parameters(), lr=1e-3)
    opt_d.load_state_dict(opt_b.state_dict())
    for pa, pb in zip(m_a.parameters(), m_b.parameters()):
        gr = torch.full_like(pa, 1e-3)
        pa.grad, pb.grad = gr.clone(), gr.clone()
    opt_c.step()
    opt_d.step()
    for (k, pa), (_, pb) in zip(m_a.named_parameters(), m_b.named_parameters()):
        assert (pa - pb).abs().max().item() <= 4e-6 * max(1.0, pb.abs().max().item()), k


def test_fused_adam_maintains_packed_copies(cuda_dev):
    """After a training step with egaze.optim.Adam the cached packed copies equal a fresh pack of the updated weights (both
    the fp16 forward copy and the bf16 data-gradient copy), no re-pack launch is needed, and a training run with it tracks one
    with torch.optim.Adam."""
    import floss as floss_mod
    from egaze import ops
    from egaze.optim import Adam
    from oracle import egaze_oracle as orc
    m_a = _make_sp(cuda_dev, 1)
    m_b = copy.deepcopy(m_a)
    opt_a = Adam(m_a.parameters(), lr=1e-5)
    opt_b = torch.optim.Adam(m_b.parameters(), lr=1e-5)
    crit = floss_mod.floss()
    losses = {"a": [], "b": []}
    for i in range(3):
        x_s, x_t, gt = [torch.from_numpy(a).to(cuda_dev) for a in orc.synth_sp_inputs(2, 64, 50 + i)]
        for m, opt, key in ((m_a, opt_a, "a"), (m_b, opt_b, "b")):
            opt.zero_grad(set_to_none=True)
            loss = crit(m(x_s, x_t), gt)
            loss.backward()
            opt.step()
            losses[key].append(loss.item())
    for la, lb in zip(losses["a"], losses["b"]):
        assert abs(la - lb) <= 2e-3 * abs(lb), losses
    checked = 0
    for mod in m_a.modules():
        if isinstance(mod, (torch.nn.Conv2d, torch.nn.Conv3d)) and tuple(mod.weight.shape[-2:]) == (3, 3):
            w = mod.weight
            for mode, rp, cp, fmt, hi, lo in ops.pack_cache.entries_for(w):
                key = (id(w), mode, rp, cp, fmt)
                assert ops.pack_cache._d[key][1] == w._version, "copy not marked current"
                fresh = ops._PackCache().get(w, mode, rows_p=rp, cols_p=cp, fmt=fmt)
                assert torch.equal(hi, fresh[0]) and torch.equal(lo, fresh[1]), (mode, fmt, tuple(w.shape))
                checked += 1
    assert checked >= 2 * 39 - 2      # 39 conv weights; the two first-layer convs have no data-gradient copy


def test_graphed_step_with_fused_adam(cuda_dev):
    """A step captured with egaze.optim.Adam (no re-pack pass inside the graph) replays to the same losses as eager steps."""
    import floss as floss_mod
    from egaze.graph import GraphedStep
    from egaze.optim import Adam
    from oracle import egaze_oracle as orc
    batches = [[torch.from_numpy(a).to(cuda_dev) for a in orc.synth_sp_inputs(2, 64, 70 + i)] for i in range(3)]
    m_e = _make_sp(cuda_dev, 2)
    m_g = copy.deepcopy(m_e)
    crit = floss_mod.floss()

    def step_fn(model, opt):
        def step(x_s, x_t, gt):
            opt.zero_grad(set_to_none=True)
            loss = crit(model(x_s, x_t), gt)
            loss.backward()
            opt.step()
            return loss
        return step

    opt_e = Adam(m_e.parameters(), lr=1e-5)
    eager = step_fn(m_e, opt_e)
    loss_e = [eager(*b).item() for b in batches]
    opt_g = Adam(m_g.parameters(), lr=1e-5)
    gs = GraphedStep(step_fn(m_g, opt_g), batches[-1], modules=[m_g], optimizers=[opt_g])
    loss_g = [gs(*b).item() for b in batches]
    for a, b in zip(loss_e, loss_g):
        assert abs(a - b) <= 1e-4 * abs(a), (loss_e, loss_g)
    for st in opt_g.state.values():
        assert float(st["step"]) == 3.0
    m_e.eval()
    m_g.eval()
    with torch.no_grad():
        oe, og = m_e(*batches[0][:2]), m_g(*batches[0][:2])
    assert (oe - og).abs().max().item() <= 2e-3
