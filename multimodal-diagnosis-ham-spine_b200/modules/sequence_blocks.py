"""SequenceEncoder (modules/sequence_blocks.py:6-70) for the multi-slice branch of model.py:316-331: an nn.LSTM / nn.GRU
(any depth, uni- or bidirectional; last time step -> proj) or an nn.TransformerEncoder (sinusoidal positions, post-LN
layers, mean over the slices -> proj) over the T pooled slice features.  The torch modules are only parameter containers
(state_dict keys `rnn.weight_ih_l0`, `rnn.weight_hh_l0_reverse`, `encoder.layers.0.self_attn.in_proj_weight`, ...): gate
projections run on the fp32 head kernels, the cells on csrc/elementwise.cu::{lstm,gru}_cell_*, back-propagation through
time through torch.autograd over those nodes; the Transformer layers reuse the fused attention / GEMM / LayerNorm kernels."""
import math

import torch
import torch.nn as nn

from .. import functional as Fm
from .. import ops
from ..encoder import MdhsModule
from .fusion_blocks import _next_seed


class _ParamLinear:
    """Adapter: lets Fm.linear_f32 treat an (weight, bias) pair of nn.LSTM / nn.GRU like an nn.Linear."""

    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


class SequenceEncoder(MdhsModule):
    def __init__(self, input_dim, hidden_dim=256, encoder_type="lstm", num_layers=1, bidirectional=True, dropout=0.1, num_heads=4):
        super().__init__()
        self.encoder_type = encoder_type.lower()
        self.hidden_dim = hidden_dim
        if self.encoder_type in ("lstm", "gru"):
            rnn_cls = nn.LSTM if self.encoder_type == "lstm" else nn.GRU
            self.rnn = rnn_cls(input_dim, hidden_dim, num_layers=num_layers, batch_first=True, bidirectional=bidirectional,
                               dropout=dropout if num_layers > 1 else 0.0)
            output_dim = hidden_dim * (2 if bidirectional else 1)
            self.proj = nn.Linear(output_dim, hidden_dim) if output_dim != hidden_dim else nn.Identity()
        elif self.encoder_type == "transformer":
            if input_dim % num_heads or (input_dim // num_heads) not in (32, 64):
                raise ValueError("the fused attention kernel needs head_dim 32 or 64 (input_dim / num_heads)")
            layer = nn.TransformerEncoderLayer(d_model=input_dim, nhead=num_heads, dim_feedforward=max(hidden_dim * 4, input_dim * 2),
                                               dropout=dropout, batch_first=True)
            self.encoder = nn.TransformerEncoder(layer, num_layers=num_layers)
            self.proj = nn.Linear(input_dim, hidden_dim) if input_dim != hidden_dim else nn.Identity()
        else:
            raise ValueError(f"Unsupported sequence encoder type: {encoder_type}")

    # ------------------------------------------------------------------ recurrent encoders
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
            gi, gh = Fm.linear_f32(xs[t], st, ih), Fm.linear_f32(h, st, hh)
            if self.encoder_type == "lstm":
                h, c = Fm.lstm_cell(Fm.add_f32(gi, gh), c)
            else:
                h = Fm.gru_cell(gi, gh, h)
            outs[t] = h
        return outs

    def _recurrent(self, st, x):
        T = x.shape[1]
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
        return xs[-1]                                    # out[:, -1, :]

    # ------------------------------------------------------------------ transformer encoder
    @staticmethod
    def _positional_encoding(seq_len, dim, device):
        position = torch.arange(seq_len, device=device).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, dim, 2, device=device, dtype=torch.float32) * (-math.log(10000.0) / dim))
        pe = torch.zeros(seq_len, dim, device=device, dtype=torch.float32)
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        return pe

    def _transformer(self, st, x):
        B, T, E = x.shape
        pos = self._positional_encoding(T, E, x.device)          # constant table (sequence_blocks.py:48-58)
        x2 = Fm.to_bf16(Fm.add_f32(x.float().reshape(B * T, E).contiguous(), pos.repeat(B, 1)))
        for lyr in self.encoder.layers:
            mha = lyr.self_attn
            H = mha.num_heads
            D = E // H
            p = float(lyr.dropout.p) if self.training else 0.0
            pa = float(mha.dropout) if self.training else 0.0
            qkv = Fm.linear(x2, st, mha.in_proj_weight, mha.in_proj_bias)
            a = Fm.attention(qkv, None, None, B, H, T, T, D, 1.0 / math.sqrt(D), drop_p=pa, seed=_next_seed())
            h = Fm.linear(a, st, mha.out_proj.weight, mha.out_proj.bias, residual=x2, drop_p=p, seed=_next_seed())
            x2 = Fm.layernorm(h, st, lyr.norm1)                   # post-LN (norm_first = False)
            f = Fm.linear(x2, st, lyr.linear1.weight, lyr.linear1.bias, act=ops.ACT_RELU, drop_p=p, seed=_next_seed())
            h = Fm.linear(f, st, lyr.linear2.weight, lyr.linear2.bias, residual=x2, drop_p=p, seed=_next_seed())
            x2 = Fm.layernorm(h, st, lyr.norm2)
        if self.encoder.norm is not None:
            x2 = Fm.layernorm(x2, st, self.encoder.norm)
        return Fm.mean_tokens(x2, B, T)                           # (B, E) fp32

    def forward(self, x):
        """x: (B, T, D) fp32 -> (B, hidden)."""
        st = self.store(x.device)
        out = self._recurrent(st, x) if self.encoder_type in ("lstm", "gru") else self._transformer(st, x)
        if isinstance(self.proj, nn.Identity):
            return out
        return Fm.linear_f32(out, st, self.proj)
