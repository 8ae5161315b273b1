"""Dual-expert gate (modules/gating.py:5-23): sigmoid(MLP([local | context | entropy])) mixing two logit sets."""
import torch
import torch.nn as nn

from .. import functional as Fm
from .. import ops
from ..encoder import MdhsModule


class DualExpertGate(MdhsModule):
    def __init__(self, lesion_dim, context_dim, hidden_dim=128, use_entropy=True):
        super().__init__()
        self.use_entropy = use_entropy
        in_dim = lesion_dim + context_dim + (1 if use_entropy else 0)
        self.fc = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, 1))

    def forward(self, lesion_feat, context_feat, entropy=None):
        st = self.store(lesion_feat.device)
        if self.use_entropy:
            if entropy is None:
                raise ValueError("entropy is required when use_entropy=True")
            gate_in = torch.cat([lesion_feat.float(), context_feat.float(), entropy.float()], dim=-1)
        else:
            gate_in = torch.cat([lesion_feat.float(), context_feat.float()], dim=-1)
        h = Fm.linear_f32(gate_in, st, self.fc[0], act=ops.ACT_RELU)
        return torch.sigmoid(Fm.linear_f32(h, st, self.fc[2]))
