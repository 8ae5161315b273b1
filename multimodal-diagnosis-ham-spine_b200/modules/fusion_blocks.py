"""Fusion modules with the reference's constructors and parameter names (modules/fusion_blocks.py:7-261),
running on the B200 kernels.  nn.MultiheadAttention / nn.LayerNorm / nn.Linear objects are kept as
parameter containers so the state_dict keys (`attn1.in_proj_weight`, `attn2.q_proj_weight`, ...) match.

Token tensors are bf16 (B, N, dim); pooled outputs are fp32 (B, hidden).
`mamba` / `vmamba` fusion need un-vendored third-party CUDA extensions in the reference
(fusion_blocks.py:264-334) and are out of scope.
"""
import math

import torch
import torch.nn as nn

from .. import functional as Fm
from .. import ops
from ..encoder import MdhsModule

_seed_counter = [0xF00D]


def _next_seed():
    _seed_counter[0] += 97
    return _seed_counter[0]


def _tok2d(t):
    B, N, C = t.shape
    return t.reshape(B * N, C), B, N


def _as_bf16_tokens(t):
    if t.dtype != torch.bfloat16:
        B, N, C = t.shape
        return Fm.to_bf16(t.reshape(B * N, C).float()).view(B, N, C)
    return t


def _key_mask(mask):
    if mask is None:
        return None
    return (mask != 0).to(torch.uint8).contiguous()


class _MHA:
    """Kernel-side view of one nn.MultiheadAttention container (batch_first, bias=True)."""

    def __init__(self, mha):
        self.m = mha
        self.E = mha.embed_dim
        self.H = mha.num_heads
        self.D = self.E // self.H
        if self.D not in (32, 64):
            raise ValueError(f"head_dim {self.D} unsupported by the fused attention kernel (32 or 64)")
        self.p = float(mha.dropout)

    def groups(self):
        m = self.m
        if not m._qkv_same_embed_dim:
            return [[m.k_proj_weight, m.v_proj_weight]]
        return []

    def self_attention(self, st, x2d, B, N, residual, training):
        """x2d: normalised input [B*N, E]; returns out_proj(attn) + residual."""
        m, E = self.m, self.E
        qkv = Fm.linear(x2d, st, m.in_proj_weight, m.in_proj_bias)
        a = Fm.attention(qkv, None, None, B, self.H, N, N, self.D, 1.0 / math.sqrt(self.D),
                         drop_p=self.p if training else 0.0, seed=_next_seed())
        return Fm.linear(a, st, m.out_proj.weight, m.out_proj.bias, residual=residual)

    def cross_attention(self, st, q_in, kv_in, B, Nq, Nk, key_mask, residual, training):
        m, E = self.m, self.E
        tr = m.in_proj_bias.requires_grad
        b32 = m.in_proj_bias.data
        gb = st.g32(m.in_proj_bias) if tr else None
        if m._qkv_same_embed_dim:
            w16, gw = st.w16(m.in_proj_weight), (st.g32(m.in_proj_weight) if m.in_proj_weight.requires_grad else None)
            q = Fm.linear(q_in, st, None, w16=w16[:E], gw=None if gw is None else gw[:E], b32=b32[:E],
                          gb=None if gb is None else gb[:E])
            kv = Fm.linear(kv_in, st, None, w16=w16[E:], gw=None if gw is None else gw[E:], b32=b32[E:],
                           gb=None if gb is None else gb[E:])
        else:
            q = Fm.linear(q_in, st, None, w16=st.w16(m.q_proj_weight),
                          gw=st.g32(m.q_proj_weight) if m.q_proj_weight.requires_grad else None, b32=b32[:E],
                          gb=None if gb is None else gb[:E])
            _, w16kv, gwkv = st.fused([m.k_proj_weight, m.v_proj_weight], (2 * E, m.kdim))
            kv = Fm.linear(kv_in, st, None, w16=w16kv, gw=gwkv if m.k_proj_weight.requires_grad else None, b32=b32[E:],
                           gb=None if gb is None else gb[E:])
        a = Fm.attention(q, kv, key_mask, B, self.H, Nq, Nk, self.D, 1.0 / math.sqrt(self.D),
                         drop_p=self.p if training else 0.0, seed=_next_seed())
        return Fm.linear(a, st, m.out_proj.weight, m.out_proj.bias, residual=residual)


class BasicTransformerBlock(MdhsModule):
    """Self-attention -> cross-attention -> feed-forward, pre-LN, residual (fusion_blocks.py:7-71)."""

    def __init__(self, dim, context_dim, num_heads, dropout=0.1):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = nn.MultiheadAttention(embed_dim=dim, num_heads=num_heads, dropout=dropout, batch_first=True)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = nn.MultiheadAttention(embed_dim=dim, num_heads=num_heads, dropout=dropout, batch_first=True,
                                           kdim=context_dim, vdim=context_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = nn.Sequential(nn.Linear(dim, dim * 4), nn.GELU(), nn.Dropout(dropout), nn.Linear(dim * 4, dim))
        self._a1, self._a2 = _MHA(self.attn1), _MHA(self.attn2)

    def _mdhs_groups(self):
        return self._a1.groups() + self._a2.groups()

    def forward(self, x, context, context_mask=None):
        st = self.store(x.device)
        x = _as_bf16_tokens(x)
        context = _as_bf16_tokens(context)
        x2, B, N = _tok2d(x)
        c2, _, Nt = _tok2d(context)
        h = Fm.layernorm(x2, st, self.norm1)
        x2 = self._a1.self_attention(st, h, B, N, x2, self.training)
        h = Fm.layernorm(x2, st, self.norm2)
        x2 = self._a2.cross_attention(st, h, c2, B, N, Nt, _key_mask(context_mask), x2, self.training)
        h = Fm.layernorm(x2, st, self.norm3)
        p = float(self.ff[2].p) if self.training else 0.0
        f = Fm.linear(h, st, self.ff[0].weight, self.ff[0].bias, act=ops.ACT_GELU, drop_p=p, seed=_next_seed())
        x2 = Fm.linear(f, st, self.ff[3].weight, self.ff[3].bias, residual=x2)
        return x2.view(B, N, -1)


class FusionModule(MdhsModule):
    """BasicTransformerBlock + mean over image tokens (fusion_blocks.py:74-100)."""

    def __init__(self, text_dim, hidden_dim, num_heads=4, dropout=0.1):
        super().__init__()
        self.transformer_block = BasicTransformerBlock(hidden_dim, text_dim, num_heads, dropout)
        self.pool = nn.AdaptiveAvgPool1d(1)

    def forward(self, img_tokens, txt_tokens, txt_mask=None):
        self.store(img_tokens.device)
        x = self.transformer_block(img_tokens, txt_tokens, txt_mask)
        x2, B, N = _tok2d(x)
        return Fm.mean_tokens(x2, B, N)


class CrossAttentionBlock(MdhsModule):
    """txt_proj -> MHA(img queries, text keys/values) -> LayerNorm(img + attn) (fusion_blocks.py:103-128)."""

    def __init__(self, text_dim, hidden_dim, num_heads=4, dropout=0.1):
        super().__init__()
        self.txt_proj = nn.Linear(text_dim, hidden_dim)
        self.attn = nn.MultiheadAttention(embed_dim=hidden_dim, num_heads=num_heads, dropout=dropout, batch_first=True)
        self.norm = nn.LayerNorm(hidden_dim)
        self._a = _MHA(self.attn)

    def forward(self, img_tokens, txt_tokens, txt_mask=None):
        st = self.store(img_tokens.device)
        x2, B, N = _tok2d(_as_bf16_tokens(img_tokens))
        t2, _, Nt = _tok2d(_as_bf16_tokens(txt_tokens))
        tp = Fm.linear(t2, st, self.txt_proj.weight, self.txt_proj.bias)
        y = self._a.cross_attention(st, x2, tp, B, N, Nt, _key_mask(txt_mask), x2, self.training)
        return Fm.layernorm(y, st, self.norm).view(B, N, -1)


class MultiScaleFusionModule(MdhsModule):
    """Three cross-attention blocks over layer2/3/4 tokens, pooled and averaged (fusion_blocks.py:131-160)."""

    def __init__(self, text_dim, hidden_dim, num_heads=4, dropout=0.1):
        super().__init__()
        self.cross_l2 = CrossAttentionBlock(text_dim, hidden_dim, num_heads, dropout)
        self.cross_l3 = CrossAttentionBlock(text_dim, hidden_dim, num_heads, dropout)
        self.cross_l4 = CrossAttentionBlock(text_dim, hidden_dim, num_heads, dropout)
        self.pool = nn.AdaptiveAvgPool1d(1)

    def forward(self, img_tokens, txt_tokens, txt_mask=None):
        self.store(txt_tokens.device)
        pooled = None
        for key, blk in (("layer2", self.cross_l2), ("layer3", self.cross_l3), ("layer4", self.cross_l4)):
            t = blk(img_tokens[key], txt_tokens, txt_mask)
            t2, B, N = _tok2d(t)
            p = Fm.mean_tokens(t2, B, N, mult=1.0 / 3.0)
            pooled = p if pooled is None else pooled + p
        return pooled


class HierarchicalFusionModule(MdhsModule):
    """Hierarchical features (README.md:15, "provide-题4": image layer2 / 3 / 4 x text hidden states 4 / 8 / 12, layer-wise
    interaction with adaptive weighting).  The reference describes this variant but ships no code for it, so the definition
    is ours, assembled from the reference's own pieces: one CrossAttentionBlock (fusion_blocks.py:103-128) per level -- image
    tokens of layer{2,3,4} attend to the BERT hidden state of encoder layer {4,8,12} --, token mean, and a learnable
    softmax weighting `level_logits` (3,) of the three pooled vectors (zeros at init = the plain average of the multiscale
    module).  `fusion_type="hierarchical"`; parity is pinned against oracle.port.fusion_hierarchical only."""

    LEVELS = (("layer2", 4), ("layer3", 8), ("layer4", 12))

    def __init__(self, text_dim, hidden_dim, num_heads=4, dropout=0.1):
        super().__init__()
        self.cross_l2 = CrossAttentionBlock(text_dim, hidden_dim, num_heads, dropout)
        self.cross_l3 = CrossAttentionBlock(text_dim, hidden_dim, num_heads, dropout)
        self.cross_l4 = CrossAttentionBlock(text_dim, hidden_dim, num_heads, dropout)
        self.level_logits = nn.Parameter(torch.zeros(3))

    def forward(self, img_tokens, txt_hidden, txt_mask=None):
        """img_tokens: {"layer2|3|4": (B, N_l, hidden)}; txt_hidden: {4|8|12: (B, S, text_dim)}."""
        any_t = next(iter(txt_hidden.values()))
        st = self.store(any_t.device)
        pooled = []
        for (key, lvl), blk in zip(self.LEVELS, (self.cross_l2, self.cross_l3, self.cross_l4)):
            t = blk(img_tokens[key], txt_hidden[lvl], txt_mask)
            t2, B, N = _tok2d(t)
            pooled.append(Fm.mean_tokens(t2, B, N))
        return Fm.level_mix(pooled, st, self.level_logits)


def _pool_image(image_tokens):
    """mean over tokens; a multi-scale dict averages the three pooled vectors (fusion_blocks.py:174-181)."""
    if isinstance(image_tokens, dict):
        out = None
        for key in ("layer2", "layer3", "layer4"):
            t2, B, N = _tok2d(_as_bf16_tokens(image_tokens[key]))
            p = Fm.mean_tokens(t2, B, N, mult=1.0 / 3.0)
            out = p if out is None else out + p
        return out
    t2, B, N = _tok2d(_as_bf16_tokens(image_tokens))
    return Fm.mean_tokens(t2, B, N)


def _pool_text(text_tokens, mode):
    text_tokens = _as_bf16_tokens(text_tokens)
    if mode == "mean":
        t2, B, N = _tok2d(text_tokens)
        return Fm.mean_tokens(t2, B, N)
    return Fm.to_f32(text_tokens[:, 0, :])


class ConcatFusionModule(MdhsModule):
    """Linear([mean(img) | cls(text)]) (fusion_blocks.py:163-187)."""

    def __init__(self, text_dim, hidden_dim, text_pool="cls"):
        super().__init__()
        self.text_pool = text_pool
        self.proj = nn.Linear(hidden_dim + text_dim, hidden_dim)

    def _scales(self):
        return None, None

    def forward(self, image_tokens, text_tokens, txt_mask=None):
        st = self.store(text_tokens.device)
        img = _pool_image(image_tokens)
        txt = _pool_text(text_tokens, self.text_pool)
        w_img, w_txt = self._scales()
        if w_img is not None:
            img, txt = img * w_img, txt * w_txt
        fused = torch.cat([img, txt], dim=1)
        return Fm.linear_f32(fused, st, self.proj)


class WeightedConcatFusionModule(ConcatFusionModule):
    """Concat fusion with sigmoid-gated scalar modality weights (fusion_blocks.py:190-202)."""

    def __init__(self, text_dim, hidden_dim, text_pool="cls"):
        super().__init__(text_dim, hidden_dim, text_pool=text_pool)
        self.w_img = nn.Parameter(torch.zeros(1))
        self.w_txt = nn.Parameter(torch.zeros(1))

    def _scales(self):
        # two scalars: plain autograd on one-element tensors (their gradients reach the flat buffer through .grad)
        return torch.sigmoid(self.w_img), torch.sigmoid(self.w_txt)


class HadamardFusionModule(MdhsModule):
    """LayerNorm(Linear(img) * Linear(text)) (fusion_blocks.py:205-231)."""

    def __init__(self, text_dim, hidden_dim, text_pool="cls"):
        super().__init__()
        self.text_pool = text_pool
        self.img_proj = nn.Linear(hidden_dim, hidden_dim)
        self.txt_proj = nn.Linear(text_dim, hidden_dim)
        self.norm = nn.LayerNorm(hidden_dim)

    def forward(self, image_tokens, text_tokens, txt_mask=None):
        st = self.store(text_tokens.device)
        img = _pool_image(image_tokens)
        txt = _pool_text(text_tokens, self.text_pool)
        fused = Fm.mul_f32(Fm.linear_f32(img, st, self.img_proj), Fm.linear_f32(txt, st, self.txt_proj))
        return Fm.layernorm_f32(fused, st, self.norm)


class BilinearFusionModule(MdhsModule):
    """Low-rank bilinear pooling: LayerNorm(out_proj(img_proj(img) * txt_proj(text))) (fusion_blocks.py:234-261)."""

    def __init__(self, text_dim, hidden_dim, text_pool="cls", rank=128):
        super().__init__()
        self.text_pool = text_pool
        self.img_proj = nn.Linear(hidden_dim, rank)
        self.txt_proj = nn.Linear(text_dim, rank)
        self.out_proj = nn.Linear(rank, hidden_dim)
        self.norm = nn.LayerNorm(hidden_dim)

    def forward(self, image_tokens, text_tokens, txt_mask=None):
        st = self.store(text_tokens.device)
        img = _pool_image(image_tokens)
        txt = _pool_text(text_tokens, self.text_pool)
        fused = Fm.mul_f32(Fm.linear_f32(img, st, self.img_proj), Fm.linear_f32(txt, st, self.txt_proj))
        return Fm.layernorm_f32(Fm.linear_f32(fused, st, self.out_proj), st, self.norm)


class SSMFusionModule(nn.Module):
    def __init__(self, *a, **k):
        raise ImportError("fusion_type='mamba' needs the un-vendored mamba_ssm CUDA extension in the reference "
                          "(modules/fusion_blocks.py:264-292); it is outside the B200 hot-path scope")


class VMambaFusionModule(nn.Module):
    def __init__(self, *a, **k):
        raise ImportError("fusion_type='vmamba' needs the un-vendored EnergeSnake VMAMBA2Block in the reference "
                          "(modules/fusion_blocks.py:295-334); it is outside the B200 hot-path scope")
