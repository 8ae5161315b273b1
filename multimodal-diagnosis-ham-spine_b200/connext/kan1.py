"""KANLinear / KAN1 (ConNexT/models/block/kan1.py:5-289) on the B200 kernels.

A KAN layer is y = SiLU(x) W_base^T + vec(B(x)) (W_spline * scaler)^T (kan1.py:152-165).  Here it is ONE tcgen05 GEMM
with K = 9 * in: `kan_basis_fwd` expands x into the bf16 operand [SiLU(x) | 8 cubic B-spline bases per input] and
`kan_weight_pack` keeps the matching [W_base | W_spline * scaler] operand (re-packed after every optimizer step /
load_state_dict through the ParamStore packer hook).  Backward = dgrad GEMM + wgrad GEMM on the same operands,
`kan_basis_bwd` (analytic SiLU' and B-spline derivative) and `kan_wgrad_unpack` (base / spline / scaler gradients).
Constructor signatures, parameter / buffer names and default initialisation are the reference's.
"""
import math

import torch
import torch.nn as nn

from .. import ops
from ..encoder import MdhsModule


def _pad8(n):
    return (n + 7) // 8 * 8


class KANLinear(MdhsModule):
    """Parameter container + packed operand of one KAN layer (kan1.py:5-75).  Calling it runs the layer."""

    def __init__(self, in_features, out_features, grid_size=5, spline_order=3, scale_noise=0.1, scale_base=1.0,
                 scale_spline=1.0, enable_standalone_scale_spline=True, base_activation=torch.nn.SiLU, grid_eps=0.02,
                 grid_range=[-1, 1]):
        super().__init__()
        if grid_size != 5 or spline_order != 3:
            raise ValueError("the B200 KAN kernels are built for grid_size=5, spline_order=3 (every reference config)")
        if base_activation is not torch.nn.SiLU:
            raise ValueError("the B200 KAN kernels fuse SiLU as the base activation (the reference default)")
        if in_features % 8:
            raise ValueError("in_features must be a multiple of 8 (16-byte rows of the basis operand)")
        self.in_features, self.out_features = in_features, out_features
        self.grid_size, self.spline_order = grid_size, spline_order
        h = (grid_range[1] - grid_range[0]) / grid_size
        grid = (torch.arange(-spline_order, grid_size + spline_order + 1) * h + grid_range[0]).expand(in_features, -1).contiguous()
        self.register_buffer("grid", grid)
        self.base_weight = nn.Parameter(torch.empty(out_features, in_features))
        self.spline_weight = nn.Parameter(torch.empty(out_features, in_features, grid_size + spline_order))
        if enable_standalone_scale_spline:
            self.spline_scaler = nn.Parameter(torch.empty(out_features, in_features))
        self.scale_noise, self.scale_base, self.scale_spline = scale_noise, scale_base, scale_spline
        self.enable_standalone_scale_spline = enable_standalone_scale_spline
        self.grid_eps = grid_eps
        self.reset_parameters()
        object.__setattr__(self, "_solo", None)

    # -- initialisation (host side, once): same distributions as kan1.py:54-75
    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.base_weight, a=math.sqrt(5) * self.scale_base)
        with torch.no_grad():
            noise = (torch.rand(self.grid_size + 1, self.in_features, self.out_features) - 0.5) * self.scale_noise / self.grid_size
            coeff = self._curve2coeff(self.grid.T[self.spline_order:-self.spline_order], noise)
            self.spline_weight.data.copy_((self.scale_spline if not self.enable_standalone_scale_spline else 1.0) * coeff)
            if self.enable_standalone_scale_spline:
                nn.init.kaiming_uniform_(self.spline_scaler, a=math.sqrt(5) * self.scale_spline)

    def _curve2coeff(self, x, y):
        """Least-squares spline coefficients interpolating (x, y) (kan1.py:112-142); init-time only, host tensors."""
        g = self.grid.unsqueeze(0)
        xx = x.unsqueeze(-1)
        bases = ((xx >= g[..., :-1]) & (xx < g[..., 1:])).to(x.dtype)
        for k in range(1, self.spline_order + 1):
            bases = ((xx - g[..., :-(k + 1)]) / (g[..., k:-1] - g[..., :-(k + 1)]) * bases[..., :-1]
                     + (g[..., k + 1:] - xx) / (g[..., k + 1:] - g[..., 1:(-k)]) * bases[..., 1:])
        sol = torch.linalg.lstsq(bases.transpose(0, 1), y.transpose(0, 1)).solution
        return sol.permute(2, 0, 1).contiguous()

    @property
    def scaled_spline_weight(self):
        return self.spline_weight * (self.spline_scaler.unsqueeze(-1) if self.enable_standalone_scale_spline else 1.0)

    def update_grid(self, x, margin=0.01):
        raise NotImplementedError("update_grid (kan1.py:167-214) is never called on the reference's hot path")

    # -- packed GEMM operand
    @property
    def out_pad(self):
        return _pad8(self.out_features)

    @property
    def ld(self):
        return self.in_features * (1 + ops.KAN_NB)

    def _scaler(self):
        return self.spline_scaler.data if self.enable_standalone_scale_spline else None

    def pack_into(self, wcat):
        ops.kan_weight_pack(self.base_weight.data, self.spline_weight.data, self._scaler(), wcat)

    def forward(self, x):
        """Stand-alone use (kan1.py:152-165).  Inside KAN1 / MoE the owning module drives the layer through _KanChain."""
        st = self.store(x.device)
        if self._solo is None or self._solo.store is not st:
            object.__setattr__(self, "_solo", _KanChain([self], st))
        shape = x.shape
        x2 = x.reshape(-1, self.in_features).contiguous()
        need = torch.is_grad_enabled() and (x2.requires_grad or self.base_weight.requires_grad)
        return _KanFn.apply(x2, st.anchor, self._solo, need).reshape(*shape[:-1], self.out_features)


class _KanChain:
    """Forward / backward of a stack of KAN layers on fp32 [rows, in] activations (shared by KAN1 and the MoE experts)."""

    def __init__(self, layers, store, wcat0=None):
        self.layers = list(layers)
        self.store = store
        dev = store.device
        self.wcat = []
        for i, l in enumerate(self.layers):
            if l.grid.device != dev:
                l.grid.data = l.grid.data.to(dev)
            if i == 0 and wcat0 is not None:
                self.wcat.append(wcat0)
            else:
                self.wcat.append(torch.zeros((l.out_pad, l.ld), device=dev, dtype=torch.bfloat16))
        store.add_packer(self.pack)

    def pack(self):
        for l, w in zip(self.layers, self.wcat):
            l.pack_into(w)

    @staticmethod
    def layer_fwd(op, wcat, rows):
        y = torch.zeros((rows, wcat.shape[0]), device=op.device, dtype=torch.float32)
        ops.gemm(op, wcat, out=y, accumulate=True, split_k=-1)
        return y

    def forward(self, x, first_op=None, first_y=None, need=True):
        """x fp32 [rows, in] (row stride free).  Returns (y fp32 [rows, out_pad_last], saved)."""
        saved = []
        cur = x
        for i, (l, w) in enumerate(zip(self.layers, self.wcat)):
            if i == 0 and first_y is not None:
                op, y = first_op, first_y
            else:
                op = ops.kan_basis_fwd(cur, l.grid)
                y = self.layer_fwd(op, w, cur.shape[0])
            if need:
                saved.append((cur, op))
            cur = y[:, :l.out_features] if l.out_pad != l.out_features else y
        return y, saved

    def layer_bwd(self, i, dy, x, op, need_dx, dx=None, accumulate=False, skip_dgrad=False):
        """dy fp32 [rows, out_pad] -> parameter gradients (+=) and dx fp32 [rows, in]."""
        l, w = self.layers[i], self.wcat[i]
        dy16 = ops.cast_f32_bf16(dy)
        tr = l.base_weight.requires_grad
        if tr:
            st = self.store
            gcat = torch.zeros((l.out_pad, l.ld), device=dy.device, dtype=torch.float32)
            ops.gemm(dy16, op, a_mn=True, b_mn=True, out=gcat, accumulate=True, split_k=-1, M=l.out_pad, N=l.ld, K=dy.shape[0])
            ops.kan_wgrad_unpack(gcat, l.spline_weight.data, l._scaler(), st.g32(l.base_weight), st.g32(l.spline_weight),
                                 st.g32(l.spline_scaler) if l.enable_standalone_scale_spline else None)
        if not need_dx or skip_dgrad:
            return None
        dop = ops.gemm(dy16, w, b_mn=True, M=dy.shape[0], N=l.ld, K=l.out_pad)
        return ops.kan_basis_bwd(x, l.grid, dop, dx=dx, accumulate=accumulate)

    def backward(self, dy, saved, need_dx=True, first_external=False):
        """dy fp32 [rows, out_pad_last].  With first_external the layer-0 backward is left to the caller (fused across
        experts); returns the gradient arriving at layer 0's output in that case."""
        for i in range(len(self.layers) - 1, -1, -1):
            if i == 0 and first_external:
                return dy
            x, op = saved[i]
            dx = self.layer_bwd(i, dy, x, op, need_dx or i > 0)
            if i > 0:
                prev = self.layers[i - 1]
                if prev.out_pad != prev.out_features:
                    full = torch.zeros((dx.shape[0], prev.out_pad), device=dx.device, dtype=torch.float32)
                    full[:, :prev.out_features] = dx
                    dx = full
            dy = dx
        return dy


class _KanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, chain, need):
        y, saved = chain.forward(x, need=need)
        ctx.chain, ctx.saved = chain, saved
        last = chain.layers[-1]
        return y[:, :last.out_features].contiguous() if last.out_pad != last.out_features else y

    @staticmethod
    def backward(ctx, dy):
        chain = ctx.chain
        last = chain.layers[-1]
        dy = dy.contiguous()
        if last.out_pad != last.out_features:
            full = torch.zeros((dy.shape[0], last.out_pad), device=dy.device, dtype=torch.float32)
            full[:, :last.out_features] = dy
            dy = full
        dx = chain.backward(dy, ctx.saved, need_dx=ctx.needs_input_grad[0])
        ctx.saved = None
        return (dx if ctx.needs_input_grad[0] else None), None, None, None


class KAN1(MdhsModule):
    """Stack of KANLinear layers (kan1.py:239-289)."""

    def __init__(self, layers_hidden=[768, 512, 256], grid_size=5, spline_order=3, scale_noise=0.1, scale_base=1.0,
                 scale_spline=1.0, base_activation=torch.nn.SiLU, grid_eps=0.02, grid_range=[-1, 1]):
        super().__init__()
        self.grid_size, self.spline_order = grid_size, spline_order
        self.output_dim = layers_hidden[-1]
        self.layers = nn.ModuleList()
        for i, o in zip(layers_hidden, layers_hidden[1:]):
            self.layers.append(KANLinear(i, o, grid_size=grid_size, spline_order=spline_order, scale_noise=scale_noise,
                                         scale_base=scale_base, scale_spline=scale_spline, base_activation=base_activation,
                                         grid_eps=grid_eps, grid_range=grid_range))
        object.__setattr__(self, "_chain", None)

    def forward(self, x, update_grid=False):
        if update_grid:
            raise NotImplementedError("update_grid is not part of the hot path (kan1.py:167-214)")
        if x.numel() == 0:
            return torch.zeros((*x.shape[:-1], self.output_dim), device=x.device)
        st = self.store(x.device)
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        if x2.dtype != torch.float32:
            raise ops._lib.MdhsError("KAN1 expects fp32 features (use functional.to_f32 on bf16 token tensors)")
        x2 = x2.contiguous()
        if self._chain is None or self._chain.store is not st:   # lazily: experts inside a MoE are driven by the MoE
            object.__setattr__(self, "_chain", _KanChain(self.layers, st))
        need = torch.is_grad_enabled() and (x2.requires_grad or self.layers[0].base_weight.requires_grad)
        y = _KanFn.apply(x2, st.anchor, self._chain, need)
        return y.reshape(*shape[:-1], self.output_dim)

    def regularization_loss(self, regularize_activation=1.0, regularize_entropy=1.0):
        raise NotImplementedError("regularization_loss (kan1.py:216-236) is not used by any reference training loop")


class KAN1Head(KAN1):
    """KAN1 used as the classification head of MultimodalBaselineModel (`classifier_type="kan1"`): accepts the fused
    (B, hidden) features in bf16 or fp32.  state_dict keys: classifier.layers.N.{base_weight,spline_weight,spline_scaler,grid}."""

    def forward(self, x, update_grid=False):
        from .. import functional as Fm
        x32 = Fm.to_f32(x) if x.dtype == torch.bfloat16 else x.float()
        return super().forward(x32, update_grid=update_grid)
