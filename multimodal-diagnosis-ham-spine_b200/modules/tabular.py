"""TabularEncoder (modules/tabular.py:4-15): Linear -> ReLU -> Dropout -> Linear on (B, input_dim) fp32 features, and the
`tabular_fusion` block of model.py:163-167.  Same parameter names (`net.0.*`, `net.3.*`); tiny fp32 kernels (csrc/heads.cu)."""
import torch.nn as nn

from .. import functional as Fm
from .. import ops
from ..encoder import MdhsModule
from .fusion_blocks import _next_seed


class TabularEncoder(MdhsModule):
    def __init__(self, input_dim, hidden_dim=128, dropout=0.1):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, hidden_dim))

    def forward(self, x):
        st = self.store(x.device)
        h = Fm.linear_f32(x.float().contiguous(), st, self.net[0], act=ops.ACT_RELU)
        h = Fm.dropout_f32(h, float(self.net[2].p), _next_seed(), self.training)
        return Fm.linear_f32(h, st, self.net[3])


class TabularFusion(nn.Sequential, MdhsModule):
    """nn.Sequential(Linear(hidden + tabular_hidden, hidden), ReLU, Dropout) of model.py:163-167 (keys `0.*`)."""

    def __init__(self, in_dim, hidden_dim, dropout):
        nn.Sequential.__init__(self, nn.Linear(in_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout))
        object.__setattr__(self, "_store", None)

    def forward(self, x):
        st = self.store(x.device)
        h = Fm.linear_f32(x.contiguous(), st, self[0], act=ops.ACT_RELU)
        return Fm.dropout_f32(h, float(self[2].p), _next_seed(), self.training)
