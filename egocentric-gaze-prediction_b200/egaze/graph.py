"""Whole-step CUDA graphs.

One SP training step is ~300 kernel launches on three streams (two trunks, weight gradients) plus ~30 optimiser launches;
the host needs 22-30 ms to enqueue them, close to the 33 ms the B200 needs to run them, so any host hiccup shows up as GPU
idle time.  `GraphedStep` captures the whole step -- forward, loss, backward with its stream forks and joins, the NCCL
gradient all-reduce when there is one, the optimiser -- once and replays it with one launch (0.1 ms of host time).

Everything the step does is already device-side (BatchNorm running statistics, Adam's step counter with
`capturable=True`, the weight re-pack that follows an optimiser step), which is what makes the replay equivalent to the
eager step; tests/test_gpu_graph.py checks losses and parameters of replayed steps against eager ones.

The reference has no counterpart (gaze_full.py / SP.py:118-150 launch eagerly); this is an opt-in for the training
loops built on the drop-in modules, not something the modules do behind the caller's back.
"""
import torch

from . import ops


def _state_tensors(optimizers):
    for opt in optimizers:
        for st in opt.state.values():
            for k, v in st.items():
                if torch.is_tensor(v):
                    yield (id(opt), id(st), k), v


class GraphedStep(object):
    """Capture `fn(*inputs)` once, replay it per call.

    fn          -- the step: a callable that takes the input tensors, launches everything on the current stream (side
                   streams it forks must join back before it returns) and returns a tensor or a tuple of tensors.
    inputs      -- example CUDA tensors; they become the graph's static inputs (cloned unless `own_inputs`), later calls
                   copy their arguments into them.
    modules     -- modules whose parameters and buffers the warm-up steps must not advance (restored after capture).
    optimizers  -- optimisers stepped inside `fn` (torch.optim.Adam/AdamW need capturable=True); their state is restored
                   too: entries that existed before come back, entries the warm-up created are zeroed (the initial state
                   of Adam, AdamW and SGD momentum).
    warmup      -- eager runs of `fn` on a side stream before the capture (torch needs >= 1; lazy state, caches and
                   cuBLAS/NCCL workspaces are created there, not under capture).
    """

    def __init__(self, fn, inputs, modules=(), optimizers=(), warmup=3, own_inputs=False, restore=True):
        inputs = list(inputs)
        if not inputs or not all(torch.is_tensor(t) and t.is_cuda for t in inputs):
            raise ValueError("GraphedStep: inputs must be CUDA tensors")
        optimizers = list(optimizers)
        for opt in optimizers:
            if "capturable" in opt.defaults and not all(g.get("capturable", False) for g in opt.param_groups):
                raise ValueError("GraphedStep: %s must be built with capturable=True" % type(opt).__name__)
        self.fn = fn
        self.device = inputs[0].device
        self.static_inputs = inputs if own_inputs else [t.clone() for t in inputs]
        self.graph = None
        self.outputs = None
        tensors = []
        if restore:
            seen = set()
            for m in modules:
                for t in list(m.parameters()) + list(m.buffers()):
                    if id(t) not in seen:
                        seen.add(id(t))
                        tensors.append(t)
        saved = [t.detach().clone() for t in tensors]
        saved_opt = {k: v.detach().clone() for k, v in _state_tensors(optimizers)} if restore else None

        # An optimiser that rewrites the packed weight copies itself (egaze.optim.Adam) leaves them current at the end of every
        # step, so the captured step needs no re-pack pass; with any other optimiser the capture must contain the re-pack an
        # optimiser step makes necessary, whatever the cache holds at capture time.
        from .optim import Adam as _FusedAdam
        optimizers = list(optimizers)
        self._packs_maintained = bool(optimizers) and all(isinstance(o, _FusedAdam) for o in optimizers)
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, int(warmup))):
                if not self._packs_maintained:
                    ops.pack_cache.mark_stale()   # same re-pack set (and launch table) as the capture below
                fn(*self.static_inputs)
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        if not self._packs_maintained:
            ops.pack_cache.mark_stale()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = fn(*self.static_inputs)
        torch.cuda.synchronize(self.device)
        self.graph = graph
        self.outputs = out
        # device tables a captured multi-weight re-pack launch reads (their addresses are baked into the graph)
        self._held = ops.pack_cache.held_tables()

        if restore:
            with torch.no_grad():
                for t, s in zip(tensors, saved):
                    t.copy_(s)
                for k, v in _state_tensors(optimizers):
                    if k in saved_opt:
                        v.copy_(saved_opt[k])
                    else:
                        v.zero_()
            ops.pack_cache.mark_stale()
            if self._packs_maintained:
                ops.pack_cache.refresh(min_stale=1)   # the restored weights' copies: the graph itself does not re-pack
            torch.cuda.synchronize(self.device)

    def __call__(self, *inputs):
        if len(inputs) != len(self.static_inputs):
            raise ValueError("GraphedStep: expected %d inputs, got %d" % (len(self.static_inputs), len(inputs)))
        for s, t in zip(self.static_inputs, inputs):
            if t is s:
                continue
            if t.shape != s.shape or t.dtype != s.dtype:
                raise ValueError("GraphedStep: input %s/%s does not match the captured %s/%s"
                                 % (tuple(t.shape), t.dtype, tuple(s.shape), s.dtype))
            s.copy_(t, non_blocking=True)
        return self.replay()

    def release(self):
        """Destroy the graph and give its memory pool back.  Do this before `destroy_process_group()` when the step contains
        an NCCL collective: a communicator is not torn down while a graph that captured its work exists."""
        if self.graph is not None:
            torch.cuda.synchronize(self.device)
            self.graph.reset()
        self.graph = None
        self.outputs = None
        self.fn = None
        self._held = None

    def replay(self):
        """Run the step on whatever the static inputs hold."""
        if self.graph is None:
            raise RuntimeError("GraphedStep: released")
        self.graph.replay()
        # the replayed optimiser changed the weights without touching their tensor versions: eager calls of the modules
        # after this must re-pack (the graph itself always does)
        ops.pack_cache.mark_stale()
        return self.outputs
