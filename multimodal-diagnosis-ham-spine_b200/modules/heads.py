"""Classification heads with the reference's constructors / parameter names (modules/heads.py:28-140 and
the inline MLP head of model.py:195-200).  Heads work on fp32 (B, hidden) features; they are tiny, so every
Linear is one fused fp32 kernel launch (csrc/heads.cu)."""
import torch
import torch.nn as nn

from .. import functional as Fm
from .. import ops
from ..encoder import MdhsModule
from .fusion_blocks import _next_seed


def _f32(x):
    return Fm.to_f32(x) if x.dtype == torch.bfloat16 else x.float()


class MLPHead(nn.Sequential, MdhsModule):
    """nn.Sequential(Linear, ReLU, Dropout, Linear) of model.py:195-200 (state_dict keys `0.*`, `3.*`)."""

    def __init__(self, hidden_dim, num_classes, dropout):
        nn.Sequential.__init__(self, nn.Linear(hidden_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout),
                               nn.Linear(hidden_dim, num_classes))
        object.__setattr__(self, "_store", None)

    def forward(self, x):
        st = self.store(x.device)
        h = Fm.linear_f32(_f32(x), st, self[0], act=ops.ACT_RELU)
        h = Fm.dropout_f32(h, float(self[2].p), _next_seed(), self.training)
        return Fm.linear_f32(h, st, self[3])


class ResidualBlock(MdhsModule):
    def __init__(self, hidden_dim, dropout=0.1):
        super().__init__()
        self.linear1 = nn.Linear(hidden_dim, hidden_dim)
        self.act = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(hidden_dim, hidden_dim)
        self.norm = nn.LayerNorm(hidden_dim)

    def forward(self, x):
        st = self.store(x.device)
        x = _f32(x)
        out = Fm.linear_f32(x, st, self.linear1, act=ops.ACT_RELU)
        out = Fm.dropout_f32(out, float(self.dropout.p), _next_seed(), self.training)
        out = Fm.linear_f32(out, st, self.linear2)
        return Fm.layernorm_f32(x + out, st, self.norm)


class ResidualClassifier(MdhsModule):
    """project -> ReLU -> residual block -> classifier (heads.py:46-58)."""

    def __init__(self, input_dim, hidden_dim, num_classes, dropout=0.1):
        super().__init__()
        self.project = nn.Linear(input_dim, hidden_dim)
        self.res_block = ResidualBlock(hidden_dim, dropout)
        self.classifier = nn.Linear(hidden_dim, num_classes)
        self.act = nn.ReLU()

    def forward(self, x):
        st = self.store(x.device)
        x = Fm.linear_f32(_f32(x), st, self.project, act=ops.ACT_RELU)
        x = self.res_block(x)
        return Fm.linear_f32(x, st, self.classifier)


class AttentionPoolingClassifier(MdhsModule):
    """Learned query attending to a length-1 sequence (heads.py:61-105).  Softmax over a single key is 1, so
    the attention output is out_proj(v_proj(x)) (times the per-(sample, head) dropout keep-scale in training);
    the query / q / k projections are dead compute whose parameters exist only for the state_dict."""

    def __init__(self, input_dim, hidden_dim, num_classes, num_heads=4, dropout=0.1):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.query = nn.Parameter(torch.randn(1, 1, hidden_dim))
        self.attn = nn.MultiheadAttention(hidden_dim, num_heads, dropout=dropout, batch_first=True)
        self.classifier = nn.Linear(hidden_dim, num_classes)

    def forward(self, x):
        st = self.store(x.device)
        x = _f32(x)
        E = self.hidden_dim
        m = self.attn
        w, b = m.in_proj_weight, m.in_proj_bias
        tr = w.requires_grad
        gw, gb = (st.g32(w), st.g32(b)) if tr else (None, None)
        v = Fm.LinearF32Fn.apply(x, w.data[2 * E:], b.data[2 * E:], None if gw is None else gw[2 * E:],
                                 None if gb is None else gb[2 * E:], ops.ACT_NONE)
        if self.training and m.dropout > 0:
            # dropout acts on the (B*heads, 1, 1) attention weights: one keep-scale per (sample, head)
            B, H, D = x.shape[0], m.num_heads, E // m.num_heads
            keep = Fm.dropout_f32(torch.ones(B, H, device=x.device), float(m.dropout), _next_seed(), True)
            v = (v.view(B, H, D) * keep.unsqueeze(-1)).reshape(B, E)
        out = Fm.linear_f32(v, st, m.out_proj)
        return Fm.linear_f32(out, st, self.classifier)


def build_kan_head(hidden_dim, num_classes, dropout=0.1, num_groups=8, act_mode="gelu"):
    """The reference builds this head from `ikan.GroupKAN.GroupKANLinear` (modules/heads.py:108-140), a sibling
    checkout that is neither vendored nor installed: the arithmetic cannot be pinned, so it is not re-implemented
    (the vendored efficient-KAN of ConNexT/models/block/kan1.py is available as mdhs_b200.connext.KANLinear)."""
    raise ImportError("GroupKANLinear not found. Install the ikan package to use classifier_type='kan'.")


__all__ = ["MLPHead", "ResidualBlock", "ResidualClassifier", "AttentionPoolingClassifier", "build_kan_head"]
