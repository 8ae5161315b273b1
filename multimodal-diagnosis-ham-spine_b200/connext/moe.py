"""Sparsely-gated mixture of KAN experts (ConNexT/models/block/moe.py:130-291) on the B200 kernels.

Same constructor, parameters (`w_gate`, `w_noise`, `experts.N.layers.M.*`, buffers `mean`, `std`) and
`moe(x, loss_coef) -> (y, aux_loss)` surface.  What differs from the reference's execution (not its results):
  * routing never leaves the GPU: noisy top-k gating, importance / load statistics and the cv^2 balance loss are one
    warp-per-row kernel + one tiny loss kernel (the reference's SparseDispatcher does nonzero -> sort -> .tolist());
  * the experts are evaluated densely on all B rows and combined with the (mostly zero) gate matrix -- with
    E = 4 experts on B <= 512 rows that is cheaper than ragged per-expert batches, and the first KAN layer of all
    experts shares one basis operand, so it is ONE GEMM with N = E * hidden;
  * the Gaussian gating noise comes from the counter-based generator used for dropout (graph-replay safe); tests can
    inject a fixed noise tensor through `moe._noise_override` to compare against the reference bit for bit.
"""
import torch
import torch.nn as nn

from .. import ops
from ..encoder import MdhsModule
from .kan1 import KAN1, _KanChain


class _MoEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, moe, noisy, noise, coef, need):
        st = moe._store
        E, k = moe.num_experts, moe.k
        wg, wn = moe.w_gate.data, moe.w_noise.data
        gates, clean, raw, probs, topidx, importance, load = ops.moe_gate_fwd(x, wg, wn, noise, k, noisy, moe.mean, moe.std)
        loss, d_imp, d_load = ops.moe_loss(importance, load, coef)
        B = x.shape[0]
        chains = moe._chains
        l0 = chains[0].layers[0]
        # layer 0 of every expert reads the same x: one basis operand, one GEMM over the stacked expert weights
        op0 = ops.kan_basis_fwd(x, l0.grid)
        y0 = _KanChain.layer_fwd(op0, moe._wcat0, B)                       # [B, E * out_pad0]
        last = chains[0].layers[-1]
        Y = torch.empty((E, B, last.out_pad), device=x.device, dtype=torch.float32)
        saved = []
        for e, ch in enumerate(chains):
            y0e = y0[:, e * l0.out_pad:(e + 1) * l0.out_pad]
            if len(ch.layers) == 1:
                Y[e].copy_(y0e)
                saved.append([(x, op0)])
                continue
            ye, sv = ch.forward(x, first_op=op0, first_y=y0e, need=need)
            Y[e].copy_(ye)
            saved.append(sv)
        y = ops.moe_combine_fwd(gates, Y, moe.output_size)
        ctx.moe, ctx.noisy, ctx.noise = moe, noisy, noise
        ctx.saved = (x, gates, clean, raw, probs, topidx, d_imp, d_load, Y, op0, saved)
        return y, loss.view(())

    @staticmethod
    def backward(ctx, dy, dloss):
        moe = ctx.moe
        st = moe._store
        x, gates, clean, raw, probs, topidx, d_imp, d_load, Y, op0, saved = ctx.saved
        ctx.saved = None
        E, k = moe.num_experts, moe.k
        B = x.shape[0]
        chains = moe._chains
        l0 = chains[0].layers[0]
        dgates, dY = ops.moe_combine_bwd(gates, Y, dy.contiguous())
        dy0 = torch.empty((B, E * l0.out_pad), device=x.device, dtype=torch.float32)
        for e, ch in enumerate(chains):
            g = ch.backward(dY[e], saved[e], need_dx=True, first_external=True)
            dy0[:, e * l0.out_pad:(e + 1) * l0.out_pad].copy_(g)
        # fused layer 0: wgrad for all experts in one GEMM, dgrad sums over the experts through K = E * out_pad0
        dy16 = ops.cast_f32_bf16(dy0)
        tr = l0.base_weight.requires_grad
        if tr:
            gcat = torch.zeros((E * l0.out_pad, l0.ld), device=x.device, dtype=torch.float32)
            ops.gemm(dy16, op0, a_mn=True, b_mn=True, out=gcat, accumulate=True, split_k=-1, M=E * l0.out_pad, N=l0.ld, K=B)
            for e, ch in enumerate(chains):
                l = ch.layers[0]
                ops.kan_wgrad_unpack(gcat[e * l0.out_pad:(e + 1) * l0.out_pad], l.spline_weight.data, l._scaler(),
                                     st.g32(l.base_weight), st.g32(l.spline_weight),
                                     st.g32(l.spline_scaler) if l.enable_standalone_scale_spline else None)
        need_dx = ctx.needs_input_grad[0]
        dx = None
        if need_dx:
            dop = ops.gemm(dy16, moe._wcat0, b_mn=True, M=B, N=l0.ld, K=E * l0.out_pad)
            dx = ops.kan_basis_bwd(x, l0.grid, dop)
        trg = moe.w_gate.requires_grad
        if dloss is None:
            dloss = torch.zeros((), device=x.device, dtype=torch.float32)
        ops.moe_gate_bwd(x, moe.w_gate.data, moe.w_noise.data, ctx.noise, dgates, d_imp, d_load, dloss.contiguous().view(1),
                         clean, raw, probs, topidx, dx, st.g32(moe.w_gate) if trg else None,
                         st.g32(moe.w_noise) if (trg and ctx.noisy) else None, k, ctx.noisy, moe.mean, moe.std)
        return dx, None, None, None, None, None, None


class MoE(MdhsModule):
    def __init__(self, input_size, output_size, num_experts, hidden_size, noisy_gating=True, k=4, layers_hidden=None,
                 grid_size=5, spline_order=3, scale_noise=0.1, scale_base=1.0, scale_spline=1.0):
        super().__init__()
        self.noisy_gating = noisy_gating
        self.num_experts = num_experts
        self.output_size = output_size
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.k = k
        expert_layers = [input_size, 512, 128, 32, output_size] if layers_hidden is None else layers_hidden
        self.experts = nn.ModuleList([KAN1(layers_hidden=expert_layers, grid_size=grid_size, spline_order=spline_order,
                                           scale_noise=scale_noise, scale_base=scale_base, scale_spline=scale_spline)
                                      for _ in range(num_experts)])
        self.w_gate = nn.Parameter(torch.zeros(input_size, num_experts), requires_grad=True)
        self.w_noise = nn.Parameter(torch.zeros(input_size, num_experts), requires_grad=True)
        self.softplus = nn.Softplus()
        self.softmax = nn.Softmax(1)
        self.register_buffer("mean", torch.tensor([0.0]))
        self.register_buffer("std", torch.tensor([1.0]))
        assert self.k <= self.num_experts
        if num_experts > 8:
            raise ValueError("the B200 gating kernel keeps one row's experts in registers: num_experts <= 8")
        object.__setattr__(self, "_chains", None)
        object.__setattr__(self, "_wcat0", None)
        object.__setattr__(self, "_noise_override", None)
        object.__setattr__(self, "_calls", 0)

    def _on_bind(self, store):
        l0 = self.experts[0].layers[0]
        E = self.num_experts
        wcat0 = torch.zeros((E * l0.out_pad, l0.ld), device=store.device, dtype=torch.bfloat16)
        chains = [_KanChain(ex.layers, store, wcat0=wcat0[e * l0.out_pad:(e + 1) * l0.out_pad]) for e, ex in enumerate(self.experts)]
        g0 = chains[0].layers[0].grid
        for ch in chains[1:]:
            if not torch.equal(ch.layers[0].grid, g0):
                raise NotImplementedError("experts with different layer-0 knot grids (update_grid) are not supported")
        object.__setattr__(self, "_wcat0", wcat0)
        object.__setattr__(self, "_chains", chains)

    def forward(self, x, loss_coef=1e-2):
        st = self.store(x.device)
        if x.dtype != torch.float32 or x.dim() != 2:
            raise ops._lib.MdhsError("MoE expects fp32 [batch, input_size] features")
        x = x.contiguous()
        noisy = bool(self.noisy_gating and self.training)
        noise = None
        if noisy:
            noise = self._noise_override
            if noise is None:
                object.__setattr__(self, "_calls", self._calls + 1)
                noise = ops.randn_f32((x.shape[0], self.num_experts), x.device, seed=0x6d6f65 + self._calls * 7919)
        y, loss = _MoEFn.apply(x, st.anchor, self, noisy, noise, float(loss_coef), torch.is_grad_enabled())
        return y, loss
