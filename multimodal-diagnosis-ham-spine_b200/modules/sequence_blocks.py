"""SequenceEncoder (modules/sequence_blocks.py:6-70) for the multi-slice branch of model.py:316-331: a (bi)directional
nn.LSTM over the T pooled slice features, last time step -> proj.  nn.LSTM is only the parameter container (state_dict keys
`rnn.weight_ih_l0`, `rnn.weight_hh_l0_reverse`, ...); the gate projections run on the fp32 head kernels, the cell on
csrc/elementwise.cu::lstm_cell_*, back-propagation through time through torch.autograd over those nodes.
GRU / Transformer variants of the reference are not built (NotImplementedError)."""
import torch
import torch.nn as nn

from .. import functional as Fm
from ..encoder import MdhsModule
from .fusion_blocks import _next_seed


class _ParamLinear:
    """Adapter: lets Fm.linear_f32 treat an (weight, bias) pair of nn.LSTM like an nn.Linear."""

    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


class SequenceEncoder(MdhsModule):
    def __init__(self, input_dim, hidden_dim=256, encoder_type="lstm", num_layers=1, bidirectional=True, dropout=0.1, num_heads=4):
        super().__init__()
        self.encoder_type = encoder_type.lower()
        self.hidden_dim = hidden_dim
        if self.encoder_type == "lstm":
            self.rnn = nn.LSTM(input_dim, hidden_dim, num_layers=num_layers, batch_first=True, bidirectional=bidirectional,
                               dropout=dropout if num_layers > 1 else 0.0)
            output_dim = hidden_dim * (2 if bidirectional else 1)
            self.proj = nn.Linear(output_dim, hidden_dim) if output_dim != hidden_dim else nn.Identity()
        elif self.encoder_type in ("gru", "transformer"):
            raise NotImplementedError(f"sequence encoder type {encoder_type!r} is not built on the B200 path (lstm only)")
        else:
            raise ValueError(f"Unsupported sequence encoder type: {encoder_type}")

    def _direction(self, st, xs, layer, reverse):
        sfx = f"_l{layer}" + ("_reverse" if reverse else "")
        rnn = self.rnn
        ih = _ParamLinear(getattr(rnn, "weight_ih" + sfx), getattr(rnn, "bias_ih" + sfx))
        hh = _ParamLinear(getattr(rnn, "weight_hh" + sfx), getattr(rnn, "bias_hh" + sfx))
        order = range(len(xs) - 1, -1, -1) if reverse else range(len(xs))
        # h_0 = 0 goes through the hidden projection too, so that bias_hh receives its gradient from the first step; it is a
        # grad-requiring leaf because our Functions receive parameters as plain buffers (autograd would skip the node)
        h = torch.zeros((xs[0].shape[0], rnn.hidden_size), device=xs[0].device, dtype=torch.float32,
                        requires_grad=torch.is_grad_enabled())
        c = None
        outs = [None] * len(xs)
        for t in order:
            gates = Fm.add_f32(Fm.linear_f32(xs[t], st, ih), Fm.linear_f32(h, st, hh))
            h, c = Fm.lstm_cell(gates, c)
            outs[t] = h
        return outs

    def forward(self, x):
        """x: (B, T, D) fp32 -> (B, hidden)."""
        st = self.store(x.device)
        B, T, _ = x.shape
        xs = [x[:, t, :].float().contiguous() for t in range(T)]
        rnn = self.rnn
        for layer in range(rnn.num_layers):
            fwd = self._direction(st, xs, layer, False)
            if rnn.bidirectional:
                bwd = self._direction(st, xs, layer, True)
                xs = [torch.cat([f, b], dim=1) for f, b in zip(fwd, bwd)]
            else:
                xs = fwd
            if layer + 1 < rnn.num_layers and rnn.dropout > 0:
                xs = [Fm.dropout_f32(v, float(rnn.dropout), _next_seed(), self.training) for v in xs]
        last = xs[-1]                                    # out[:, -1, :]
        if isinstance(self.proj, nn.Identity):
            return last
        return Fm.linear_f32(last, st, self.proj)
