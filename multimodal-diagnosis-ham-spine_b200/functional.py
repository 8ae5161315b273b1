"""torch.autograd glue: each Function runs our kernels in forward and backward.

Activations that flow between Functions are 2-D bf16 CUDA tensors [rows, features] (token-major) or
fp32 [B, features] for the pooled / head part.  Parameter gradients are accumulated by the kernels
directly into the ParamStore's flat fp32 gradient buffer (views passed as `gw` / `gb`); the Functions
therefore return None for them and autograd only carries activation gradients.
"""
import torch

from . import ops


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) (+ residual), bf16 [T,K] -> bf16 [T,N] on the tcgen05 GEMM."""

    @staticmethod
    def forward(ctx, x, residual, w16, gw, bias, gb, act, drop_p, seed, need):
        x = x.contiguous() if x.stride(1) != 1 else x
        T, K = x.shape
        N = w16.shape[0]
        aux = None
        if act == ops.ACT_GELU and need:
            aux = torch.empty((T, N), device=x.device, dtype=torch.bfloat16)
        y = ops.gemm(x, w16, bias=bias, act=act, aux_out=aux, residual=residual, dropout_p=drop_p, dropout_seed=seed)
        ctx.act, ctx.drop_p, ctx.seed = act, drop_p, seed
        ctx.has_res = residual is not None
        ctx.gw, ctx.gb, ctx.w16 = gw, gb, w16
        ctx.save_for_backward(x, aux if aux is not None else (y if act == ops.ACT_RELU else None))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, aux = ctx.saved_tensors
        dy = dy.contiguous()
        T, K = x.shape
        N = ctx.w16.shape[0]
        dres = dy if ctx.has_res else None
        g = dy
        if ctx.act != ops.ACT_NONE or ctx.drop_p > 0:
            # gradient at the pre-activation (both GEMMs below consume it as a TMA operand)
            g = ops.act_dropout_bwd(dy, aux, ctx.act, ctx.drop_p, ctx.seed)
        if ctx.gw is not None:
            ops.gemm(g, x, a_mn=True, b_mn=True, out=ctx.gw, accumulate=True, split_k=-1, M=N, N=K, K=T)
            if ctx.gb is not None:
                ops.col_stats(g, sum32=ctx.gb)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(g, ctx.w16, b_mn=True, M=T, N=K, K=N)
        return dx, dres, None, None, None, None, None, None, None, None


def linear(x, store, weight, bias=None, act=ops.ACT_NONE, residual=None, drop_p=0.0, seed=0, w16=None, gw=None, b32=None, gb=None):
    """weight/bias: nn.Parameters owned by `store`, or explicit (w16, gw, b32, gb) views for sliced parameters."""
    if w16 is None:
        w16 = store.w16(weight)
        gw = store.g32(weight) if weight.requires_grad else None
        if bias is not None:
            b32 = bias.data
            gb = store.g32(bias) if bias.requires_grad else None
    return LinearFn.apply(x, residual, w16, gw, b32, gb, act, float(drop_p), int(seed), torch.is_grad_enabled())


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, gg, gb, eps):
        x = x.contiguous()
        y, _, mean, rstd = ops.layernorm_fwd(x, gamma, beta, eps, save_stats=True)
        ctx.gg, ctx.gb, ctx.gamma = gg, gb, gamma
        ctx.save_for_backward(x, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        dx, _, _ = ops.layernorm_bwd(dy.contiguous(), x, mean, rstd, ctx.gamma, ctx.gg, ctx.gb)
        return dx, None, None, None, None, None


def layernorm(x, store, ln):
    tr = ln.weight.requires_grad
    return LayerNormFn.apply(x, ln.weight.data, ln.bias.data, store.g32(ln.weight) if tr else None,
                             store.g32(ln.bias) if tr else None, ln.eps)


class LayerNormF32Fn(torch.autograd.Function):
    """LayerNorm on fp32 [B, C] head features (fp32 in, fp32 out)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, gg, gb, eps):
        x = x.contiguous()
        _, y32, mean, rstd = ops.layernorm_fwd(x, gamma, beta, eps, out_bf16=False, out_f32=True)
        ctx.gg, ctx.gb, ctx.gamma = gg, gb, gamma
        ctx.save_for_backward(x, mean, rstd)
        return y32

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        _, _, dx32 = ops.layernorm_bwd(dy.contiguous(), x, mean, rstd, ctx.gamma, ctx.gg, ctx.gb, dx_bf16=False, dx_f32=True)
        return dx32, None, None, None, None, None


def layernorm_f32(x, store, ln):
    tr = ln.weight.requires_grad
    return LayerNormF32Fn.apply(x, ln.weight.data, ln.bias.data, store.g32(ln.weight) if tr else None,
                                store.g32(ln.bias) if tr else None, ln.eps)


class AttentionFn(torch.autograd.Function):
    """softmax(scale q k^T + key mask) v.  Operands are token-major bf16 buffers:
    kv is None  -> `q` is a packed [B*S, 3*H*D] q|k|v buffer (self-attention);
    kv given    -> `q` is [B*Sq, H*D] and `kv` a packed [B*Sk, 2*H*D] k|v buffer (cross-attention).
    The backward writes the packed gradients in place, so no slicing goes through autograd."""

    @staticmethod
    def forward(ctx, q, kv, key_mask, B, H, Sq, Sk, D, scale, drop_p, seed):
        HD = H * D
        q = q.contiguous()
        if kv is None:
            qq, kk, vv = q[:, :HD], q[:, HD:2 * HD], q[:, 2 * HD:]
        else:
            kv = kv.contiguous()
            qq, kk, vv = q, kv[:, :HD], kv[:, HD:]
        o, lse = ops.attention_fwd(qq, kk, vv, B, H, Sq, Sk, D, scale, key_mask=key_mask, drop_p=drop_p, seed=seed)
        ctx.cfg = (B, H, Sq, Sk, D, scale, drop_p, seed)
        ctx.key_mask = key_mask
        ctx.packed = kv is None
        ctx.save_for_backward(q, kv, o, lse)
        return o

    @staticmethod
    def backward(ctx, do):
        q, kv, o, lse = ctx.saved_tensors
        B, H, Sq, Sk, D, scale, drop_p, seed = ctx.cfg
        HD = H * D
        if ctx.packed:
            dqkv = torch.empty_like(q)
            ops.attention_bwd(q[:, :HD], q[:, HD:2 * HD], q[:, 2 * HD:], o, do.contiguous(), lse, B, H, Sq, Sk, D, scale,
                              key_mask=ctx.key_mask, drop_p=drop_p, seed=seed, dq=dqkv[:, :HD], dk=dqkv[:, HD:2 * HD],
                              dv=dqkv[:, 2 * HD:])
            return dqkv, None, None, None, None, None, None, None, None, None, None
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        ops.attention_bwd(q, kv[:, :HD], kv[:, HD:], o, do.contiguous(), lse, B, H, Sq, Sk, D, scale, key_mask=ctx.key_mask,
                          drop_p=drop_p, seed=seed, dq=dq, dk=dkv[:, :HD], dv=dkv[:, HD:])
        return dq, dkv, None, None, None, None, None, None, None, None, None


def attention(q, kv, key_mask, B, H, Sq, Sk, D, scale, drop_p=0.0, seed=0):
    return AttentionFn.apply(q, kv, key_mask, B, H, Sq, Sk, D, float(scale), float(drop_p), int(seed))


class DropoutF32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        ctx.p, ctx.seed = p, seed
        return ops.dropout_f32(x.contiguous(), p, seed)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout_f32(dy.contiguous(), ctx.p, ctx.seed), None, None


def dropout_f32(x, p, seed, training):
    if not training or p <= 0.0:
        return x
    return DropoutF32Fn.apply(x, float(p), int(seed))


class MulF32Fn(torch.autograd.Function):
    """Hadamard product of two fp32 [B, C] tensors (HadamardFusionModule / BilinearFusionModule)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        ctx.save_for_backward(a, b)
        return ops.mul_f32(a, b)

    @staticmethod
    def backward(ctx, dc):
        a, b = ctx.saved_tensors
        dc = dc.contiguous()
        return ops.mul_f32(dc, b), ops.mul_f32(dc, a)


def mul_f32(a, b):
    return MulF32Fn.apply(a, b)


class AddF32Fn(torch.autograd.Function):
    """Sum of two fp32 tensors of the same shape (pooled feature sums); the gradient passes to both."""

    @staticmethod
    def forward(ctx, a, b):
        out = torch.empty_like(a)
        ops.axpby(a.contiguous(), out, a=1.0, b=0.0)
        ops.axpby(b.contiguous(), out, a=1.0, b=1.0)
        return out

    @staticmethod
    def backward(ctx, dc):
        return dc, dc


def add_f32(a, b):
    return AddF32Fn.apply(a, b)


class MeanTokensFn(torch.autograd.Function):
    """[B*T, C] bf16 -> [B, C] fp32 mean over the T tokens of each sample (times `mult`)."""

    @staticmethod
    def forward(ctx, x, B, T, mult):
        x = x.contiguous()
        C = x.shape[1]
        y, _ = ops.mean_tokens_fwd(x, B, T, C, scale=mult / T)
        ctx.cfg = (B, T, C, mult)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, T, C, mult = ctx.cfg
        return ops.mean_tokens_bwd(dy.contiguous(), B, T, C, scale=mult / T), None, None, None


def mean_tokens(x, B, T, mult=1.0):
    return MeanTokensFn.apply(x, B, T, float(mult))


class LinearF32Fn(torch.autograd.Function):
    """Small fp32 Linear (few outputs or few rows): y = act(x W^T + b) with the SIMT head kernels."""

    @staticmethod
    def forward(ctx, x, w, b, gw, gb, act):
        x = x.contiguous()
        y = ops.linear_f32_fwd(x, w, b, act)
        ctx.act, ctx.gw, ctx.gb, ctx.w = act, gw, gb, w
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.act == ops.ACT_RELU:
            dy = ops.relu_bwd_f32(dy, y)
        elif ctx.act != ops.ACT_NONE:
            raise NotImplementedError("only ReLU is fused into the small fp32 linear")
        dx = ops.linear_f32_bwd(dy, x, ctx.w, ctx.gw, ctx.gb, need_dx=ctx.needs_input_grad[0])
        return dx, None, None, None, None, None


def linear_f32(x, store, lin, act=ops.ACT_NONE):
    tr = lin.weight.requires_grad
    return LinearF32Fn.apply(x, lin.weight.data, lin.bias.data if lin.bias is not None else None,
                             store.g32(lin.weight) if tr else None,
                             store.g32(lin.bias) if (tr and lin.bias is not None) else None, act)


class CastFn(torch.autograd.Function):
    """dtype bridge between the bf16 token world and the fp32 head world."""

    @staticmethod
    def forward(ctx, x, to_bf16):
        ctx.to_bf16 = to_bf16
        x = x.contiguous()
        return ops.cast_f32_bf16(x) if to_bf16 else ops.cast_bf16_f32(x)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        return (ops.cast_bf16_f32(dy) if ctx.to_bf16 else ops.cast_f32_bf16(dy)), None


def to_bf16(x):
    return CastFn.apply(x, True)


def to_f32(x):
    return CastFn.apply(x, False)


class CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, class_weights, label_smoothing, focal, gamma):
        logits = logits.contiguous()
        loss, dl = ops.ce_loss(logits, labels, class_weights, label_smoothing, focal, gamma, want_grad=True)
        ctx.save_for_backward(dl)
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        (dl,) = ctx.saved_tensors
        out = torch.empty_like(dl)
        ops.axpby(dl, out, a=1.0, b=0.0, a_dev=dloss.contiguous().view(1))
        return out, None, None, None, None, None


def cross_entropy(logits, labels, class_weights=None, label_smoothing=0.0, focal=False, gamma=2.0):
    return CrossEntropyFn.apply(logits, labels, class_weights, float(label_smoothing), bool(focal), float(gamma))


class SupConFn(torch.autograd.Function):
    """Supervised contrastive loss (scripts/train.py:23-44) on fp32 (B, D) features; the gradient is produced in the
    same pass and scaled by the incoming gradient in backward."""

    @staticmethod
    def forward(ctx, features, labels, temperature):
        loss, dx = ops.supcon_loss(features.contiguous(), labels, temperature, want_grad=True)
        ctx.save_for_backward(dx)
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        (dx,) = ctx.saved_tensors
        out = torch.empty_like(dx)
        ops.axpby(dx, out, a=1.0, b=0.0, a_dev=dloss.contiguous().view(1))
        return out, None, None


def supcon_loss(features, labels, temperature=0.07):
    return SupConFn.apply(features, labels, float(temperature))


class LevelMixFn(torch.autograd.Function):
    """sum_l softmax(logits)_l * p_l on fp32 (B, hidden) pooled features with LEARNABLE level logits (hierarchical fusion)."""

    @staticmethod
    def forward(ctx, logits, glogits, *levels):
        levels = [t.contiguous() for t in levels]
        ctx.logits, ctx.glogits = logits, glogits
        ctx.save_for_backward(*levels)
        return ops.level_mix_fwd(levels, logits)

    @staticmethod
    def backward(ctx, dout):
        levels = list(ctx.saved_tensors)
        dps = ops.level_mix_bwd(levels, ctx.logits, dout.contiguous(), ctx.glogits)
        return (None, None) + tuple(dps)


def level_mix(levels, store, logits_param):
    g = store.g32(logits_param) if logits_param.requires_grad else None
    return LevelMixFn.apply(logits_param.data, g, *levels)


class AvgBf16Fn(torch.autograd.Function):
    """0.5 * (a + b) on bf16 token tensors (global / local token average, model.py:303-315)."""

    @staticmethod
    def forward(ctx, a, b):
        return ops.axpby_bf16(a.contiguous(), b.contiguous(), 0.5, 0.5)

    @staticmethod
    def backward(ctx, dy):
        g = ops.axpby_bf16(dy.contiguous(), None, 0.5, 0.0)
        return g, g


def avg_bf16(a, b):
    return AvgBf16Fn.apply(a, b)


class LstmCellFn(torch.autograd.Function):
    """(gates [B,4H], c_prev [B,H] | None) -> (h, c): pointwise LSTM cell; BPTT runs through autograd over these nodes."""

    @staticmethod
    def forward(ctx, gates, c_prev):
        gates = gates.contiguous()
        c_prev = c_prev.contiguous() if c_prev is not None else None
        h, c, act = ops.lstm_cell_fwd(gates, c_prev)
        ctx.has_prev = c_prev is not None
        ctx.save_for_backward(act, c_prev if c_prev is not None else c, c)
        return h, c

    @staticmethod
    def backward(ctx, dh, dc):
        act, c_prev, c = ctx.saved_tensors
        dh = dh.contiguous() if dh is not None else None
        dc = dc.contiguous() if dc is not None else None
        dgates, dc_prev = ops.lstm_cell_bwd(dh, dc, act, c_prev if ctx.has_prev else None, c)
        return dgates, (dc_prev if ctx.has_prev else None)


def lstm_cell(gates, c_prev):
    return LstmCellFn.apply(gates, c_prev)


class GruCellFn(torch.autograd.Function):
    """(gi [B,3H], gh [B,3H], h_prev [B,H]) -> h: pointwise GRU cell (gate order r, z, n)."""

    @staticmethod
    def forward(ctx, gi, gh, h_prev):
        gi, gh, h_prev = gi.contiguous(), gh.contiguous(), h_prev.contiguous()
        h, act = ops.gru_cell_fwd(gi, gh, h_prev)
        ctx.save_for_backward(act, gh, h_prev)
        return h

    @staticmethod
    def backward(ctx, dh):
        act, gh, h_prev = ctx.saved_tensors
        dgi, dgh, dhp = ops.gru_cell_bwd(dh.contiguous(), act, gh, h_prev)
        return dgi, dgh, dhp


def gru_cell(gi, gh, h_prev):
    return GruCellFn.apply(gi, gh, h_prev)
