"""Drop-in for the reference's models/LSTMnet.py: lstmnet(num_channel=512, num_layer=2).forward(input, hidden) ->
(relu(lin(lstm(tanh(input)))), (h_n, c_n)).  nn.LSTM / nn.Linear are parameter containers (same state_dict keys);
the recurrence runs in egaze's fp32 LSTM kernels.  The module-global `batch_size` (= 1) sizes the zero state used
when hidden is None, exactly as in the reference (LSTMnet.py:13,29-33)."""
import sys

import torch
import torch.nn as nn

from egaze import ops, _lib
from egaze.modules import _needs_grad

batch_size = 1


class lstmnet(nn.Module):
    def __init__(self, num_channel=512, num_layer=2):
        super(lstmnet, self).__init__()
        self.lstm = nn.LSTM(num_channel, num_channel, num_layer)
        self.tanh = nn.Tanh()
        self.num_channel = num_channel
        self.num_layer = num_layer
        self.lin = nn.Linear(512, 512)
        self.relu = nn.ReLU()

    def forward(self, input, hidden):
        _lib.check_device(input.device)
        if self.num_channel != 512 or self.num_layer != 2:
            raise RuntimeError("egaze: lstmnet kernels are specialised for num_channel=512, num_layer=2")
        if hidden is None:
            bs = sys.modules[__name__].batch_size
            h0 = torch.zeros(self.num_layer, bs, self.num_channel, device=input.device)
            c0 = torch.zeros(self.num_layer, bs, self.num_channel, device=input.device)
        else:
            h0, c0 = hidden
        if input.dim() != 3 or input.shape[2] != 512:
            raise RuntimeError("egaze: lstmnet input must be (seq, batch, 512), got %s" % (tuple(input.shape),))
        if tuple(h0.shape) != (2, input.shape[1], 512) or tuple(c0.shape) != (2, input.shape[1], 512):
            # same failure mode as nn.LSTM in the reference (hidden=None only works at batch 1)
            raise RuntimeError("Expected hidden[0] size (2, %d, 512), got %s" % (input.shape[1], list(h0.shape)))
        if _needs_grad(self, input, h0, c0):
            from egaze.autograd import lstmnet_with_grad
            return lstmnet_with_grad(self, input, h0, c0)
        out, hn, cn, _ = ops.lstm_seq_fwd(input, h0, c0, self.lstm, self.lin)
        return (out, (hn, cn))
