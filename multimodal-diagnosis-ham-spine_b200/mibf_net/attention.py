"""IBFA cross-attention of MIBF-Net (mibf_net/attention.py:31-70) on the B200 kernels.

Q, K, V come from x and K, V from y; keys / values of both are concatenated along the sequence.  The reference
always calls it with one token per modality (model_resnet.py:41-52), i.e. a softmax over two keys per head; that
case runs as two fused projection GEMMs ([K_x|Q_x|V_x], [K_y|V_y]) + one small mixing kernel + the output GEMM.
"""
import torch
import torch.nn as nn

from .. import functional as Fm
from .. import ops
from ..encoder import MdhsModule


class SelfAttention(nn.Module):
    """Parameter container only: constructed by Resnet50WithOurs (model_resnet.py:21) but never called there."""

    def __init__(self, input_dim):
        super().__init__()
        self.query = nn.Linear(input_dim, input_dim)
        self.key = nn.Linear(input_dim, input_dim)
        self.value = nn.Linear(input_dim, input_dim)
        self.softmax = nn.Softmax(dim=-1)

    def forward(self, x):
        raise NotImplementedError("SelfAttention is dead code in the reference (never called); not on the hot path")


def compute_kl_divergence(p, q, eps=1e-8):
    """attention.py:25-28 (used inside the fused MP-loss kernel; kept for API compatibility on fp32 tensors)."""
    p = torch.clamp(p, min=eps, max=1.0)
    q = torch.clamp(q, min=eps, max=1.0)
    return torch.sum(p * (torch.log(p) - torch.log(q)), dim=-1)


class _IbfaMixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kqv_x, kv_y, B, H, D):
        kqv_x, kv_y = kqv_x.contiguous(), kv_y.contiguous()
        out, probs = ops.ibfa_fwd(kqv_x, kv_y, B, H, D)
        ctx.cfg = (B, H, D)
        ctx.save_for_backward(kqv_x, kv_y, probs)
        return out

    @staticmethod
    def backward(ctx, dout):
        kqv_x, kv_y, probs = ctx.saved_tensors
        B, H, D = ctx.cfg
        dx, dy = ops.ibfa_bwd(kqv_x, kv_y, dout.contiguous(), probs, B, H, D)
        return dx, dy, None, None, None


class MultiHeadCrossAttention_v2(MdhsModule):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        if self.head_dim * num_heads != dim:
            raise ValueError("dim must be divisible by num_heads")
        self.toK_x = nn.Linear(dim, dim)
        self.toQ_x = nn.Linear(dim, dim)
        self.toV_x = nn.Linear(dim, dim)
        self.toK_y = nn.Linear(dim, dim)
        self.toV_y = nn.Linear(dim, dim)
        self.to_out = nn.Linear(dim, dim)

    def _mdhs_groups(self):
        return [[self.toK_x.weight, self.toQ_x.weight, self.toV_x.weight], [self.toK_x.bias, self.toQ_x.bias, self.toV_x.bias],
                [self.toK_y.weight, self.toV_y.weight], [self.toK_y.bias, self.toV_y.bias]]

    def forward(self, x, y):
        st = self.store(x.device)
        B, nx, C = x.shape
        if nx != 1 or y.shape[1] != 1:
            raise NotImplementedError("the fused IBFA kernel covers the reference's call pattern (one token per modality)")
        x2 = x.reshape(B, C)
        y2 = y.reshape(B, C)
        x2 = Fm.to_bf16(x2.float()) if x2.dtype != torch.bfloat16 else x2
        y2 = Fm.to_bf16(y2.float()) if y2.dtype != torch.bfloat16 else y2
        tr = self.toK_x.weight.requires_grad
        _, wx16, gwx = st.fused([self.toK_x.weight, self.toQ_x.weight, self.toV_x.weight], (3 * C, C))
        bx32, _, gbx = st.fused([self.toK_x.bias, self.toQ_x.bias, self.toV_x.bias], (3 * C,))
        _, wy16, gwy = st.fused([self.toK_y.weight, self.toV_y.weight], (2 * C, C))
        by32, _, gby = st.fused([self.toK_y.bias, self.toV_y.bias], (2 * C,))
        kqv_x = Fm.linear(x2, st, None, w16=wx16, gw=gwx if tr else None, b32=bx32, gb=gbx if tr else None)
        kv_y = Fm.linear(y2, st, None, w16=wy16, gw=gwy if tr else None, b32=by32, gb=gby if tr else None)
        mixed = _IbfaMixFn.apply(kqv_x, kv_y, B, self.num_heads, self.head_dim)
        out = Fm.linear(mixed, st, self.to_out.weight, self.to_out.bias)
        return out.view(B, 1, C)
