"""egaze.graph.GraphedStep: a training step replayed as one CUDA graph must do what the eagerly launched step does.

Three SP training steps (two-stream forward, floss, backward on the trunk / weight-gradient streams, optimiser) on three
different batches, once launched eagerly and once replayed from a graph captured on a model in the same initial state.
The warm-up steps the capture needs must leave no trace (parameters, BatchNorm buffers and optimiser state restored),
and eager calls of the module after graph replays must see the replayed optimiser's weights.
"""
import copy

import pytest
import torch

import torch_ref

pytestmark = pytest.mark.gpu

B, S, STEPS = 2, 64, 3


def _make_sp(dev, seed=0):
    from utils import make_layers, cfg
    from models.model_SP import model_SP
    torch.manual_seed(seed)
    m = model_SP(make_layers(cfg['D'], 3), make_layers(cfg['D'], 20))
    torch_ref.randomize_(m, seed)
    return m.to(dev).train()


def _batches(dev):
    from oracle import egaze_oracle as orc
    return [[torch.from_numpy(a).to(dev) for a in orc.synth_sp_inputs(B, S, 100 + i)] for i in range(STEPS)]


def _optimizer(kind, model):
    if kind == "sgd":
        return torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.9)
    return torch.optim.Adam(model.parameters(), lr=1e-5, capturable=True)


def _step_fn(model, opt):
    import floss as floss_mod
    crit = floss_mod.floss()

    def step(x_s, x_t, gt):
        opt.zero_grad(set_to_none=True)
        out = model(x_s, x_t)
        loss = crit(out, gt.view(out.size()))
        loss.backward()
        opt.step()
        return loss
    return step


@pytest.mark.parametrize("kind", ["sgd", "adam"])
def test_graph_replay_matches_eager_steps(cuda_dev, kind):
    from egaze.graph import GraphedStep
    batches = _batches(cuda_dev)
    m_eager = _make_sp(cuda_dev)
    m_graph = copy.deepcopy(m_eager)
    x_eval = [t.clone() for t in batches[0][:2]]

    opt_e = _optimizer(kind, m_eager)
    step_e = _step_fn(m_eager, opt_e)
    loss_e = [step_e(*b).item() for b in batches]

    opt_g = _optimizer(kind, m_graph)
    gs = GraphedStep(_step_fn(m_graph, opt_g), batches[-1], modules=[m_graph], optimizers=[opt_g])
    # the capture's warm-up steps left nothing behind
    for (n, p), q in zip(m_graph.state_dict().items(), _make_sp(cuda_dev).state_dict().values()):
        assert torch.equal(p, q), n
    loss_g = [gs(*b).item() for b in batches]

    for a, b in zip(loss_e, loss_g):
        assert abs(a - b) <= 1e-4 * abs(a), (kind, loss_e, loss_g)
    sd_e, sd_g = m_eager.state_dict(), m_graph.state_dict()
    for n in sd_e:
        if n.endswith("num_batches_tracked"):
            assert int(sd_e[n]) == int(sd_g[n]) == STEPS, n
    if kind == "sgd":
        # lr * gradient updates: the two runs differ only by the summation order of the split-K atomics
        for n in sd_e:
            a, b = sd_e[n].float(), sd_g[n].float()
            assert (a - b).abs().max().item() <= 1e-5 + 1e-4 * a.abs().max().item(), n
    else:
        # Adam's first updates are +-lr whatever the gradient's size, so noise-level gradients may flip: check the state
        for st in opt_g.state.values():
            assert int(st["step"].item()) == STEPS
    # eager use of the module after graph replays sees the weights the replayed optimiser wrote
    m_eager.eval()
    m_graph.eval()
    with torch.no_grad():
        out_e = m_eager(*x_eval)
        out_g = m_graph(*x_eval)
    # sgd: the replayed and the eager run differ by the summation order of the split-K weight-gradient atomics only; after STEPS
    # updates that is ~1e-4 on the gaze map (measured 0.9e-4 .. 1.2e-4 over kernel revisions), gated at 3e-4
    tol = 3e-4 if kind == "sgd" else 2e-3
    assert (out_e - out_g).abs().max().item() <= tol


def test_graphed_step_rejects_non_capturable_adam(cuda_dev):
    from egaze.graph import GraphedStep
    m = _make_sp(cuda_dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-5)
    with pytest.raises(ValueError):
        GraphedStep(_step_fn(m, opt), _batches(cuda_dev)[0], modules=[m], optimizers=[opt])
