"""Autograd bridges (filled in as the backward kernels land)."""


def _todo(what):
    raise NotImplementedError("egaze: backward for %s is not built yet; run under torch.no_grad() / eval with "
                              "requires_grad=False parameters" % what)


def sequential_with_grad(seq, x):
    _todo("TrunkSequential")


def model_sp_with_grad(model, x_s, x_t):
    _todo("model_SP")


def late_fusion_with_grad(model, f, g):
    _todo("late_fusion")


def lstmnet_with_grad(model, inp, h0, c0):
    _todo("lstmnet")
