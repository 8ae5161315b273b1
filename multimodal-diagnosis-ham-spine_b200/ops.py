"""Thin Python wrappers: torch tensors -> raw device pointers -> C ABI (include/mdhs_b200.h).

torch is plumbing here (device memory, streams); every numeric op is one of our sm_100a kernels.
"""
import ctypes

import torch

from . import _lib
from ._lib import GemmArgs, check

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
ACT_GELU_DERIV, ACT_MUL = 3, 4   # forward GELU saving GELU'(pre) in aux_out / backward multiply by that aux_in
DT_BF16, DT_F32 = 0, 1


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MdhsError("mdhs_b200 kernels need CUDA tensors (there is no CPU fallback)")


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=torch.bfloat16, bias=None, act=ACT_NONE,
         aux_out=None, aux_in=None, dact=ACT_NONE, residual=None, accumulate=False, split_k=1,
         bn_hint=0, colsum=None, colsumsq=None, M=None, N=None, K=None, dropout_p=0.0, dropout_seed=0, conv=None,
         stat_x=None, stat_mean=None, stat_scale=None, stat_shift=None, stat_relu=True):
    """D[M,N] (+)= epi(A . B^T).  `a` is [M,K] (K-major) or [K,M] when a_mn; `b` is [N,K] or [K,N] when b_mn.
    2-D bf16 tensors with unit inner stride (row stride may exceed the row length).
    conv = (mode, N, H, W, C, R, S, stride, pad): implicit-GEMM convolution, the NHWC activation [N*H*W, C] is passed as
    `a` (mode 1: fprop / dgrad) or as `b` (mode 2: wgrad, with a_mn = b_mn = True) and M, N, K must be given.
    stat_x (+ stat_mean / stat_scale / stat_shift, colsum / colsumsq): the epilogue also accumulates the BatchNorm-backward
    reductions sum(dy'), sum(dy' * (stat_x - mean)) of the layer whose output gradient this GEMM produces."""
    _require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if conv is not None:
        assert M is not None and N is not None and K is not None
        x = a if conv[0] == 1 else b
        assert x.is_contiguous() and x.shape[0] == conv[1] * conv[2] * conv[3] and x.shape[1] == conv[4]
        m = k_a = k_b = n = None
    elif a_mn:
        k_a, m = a.shape
    else:
        m, k_a = a.shape
    if conv is not None:
        pass
    elif b_mn:
        k_b, n = b.shape
    else:
        n, k_b = b.shape
    M = m if M is None else M
    N = n if N is None else N
    K = k_a if K is None else K
    assert k_a == k_b or K is not None
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    assert out.stride(1) == 1
    args = GemmArgs()
    args.A, args.lda, args.a_mn_major = a.data_ptr(), a.stride(0), int(a_mn)
    args.B, args.ldb, args.b_mn_major = b.data_ptr(), b.stride(0), int(b_mn)
    args.D, args.ldd = out.data_ptr(), out.stride(0)
    args.d_dtype = DT_F32 if out.dtype == torch.float32 else DT_BF16
    args.accumulate = int(accumulate)
    args.M, args.N, args.K = M, N, K
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        args.bias = bias.data_ptr()
    if aux_out is not None:
        assert aux_out.dtype == torch.bfloat16
        args.aux_out, args.ld_aux_out = aux_out.data_ptr(), aux_out.stride(0)
    if aux_in is not None:
        assert aux_in.dtype == torch.bfloat16
        args.aux_in, args.ld_aux_in = aux_in.data_ptr(), aux_in.stride(0)
    args.act, args.dact = act, dact
    if residual is not None:
        args.residual, args.ldr = residual.data_ptr(), residual.stride(0)
        args.r_dtype = DT_F32 if residual.dtype == torch.float32 else DT_BF16
    args.split_k, args.bn_hint = split_k, bn_hint
    if conv is not None:
        (args.conv_mode, args.cN, args.cH, args.cW, args.cC, args.cR, args.cS, args.c_stride, args.c_pad) = [int(v) for v in conv]
    args.dropout_p, args.dropout_seed = float(dropout_p), int(dropout_seed)
    if colsum is not None:
        assert colsum.dtype == torch.float64 and colsumsq.dtype == torch.float64
        args.colsum, args.colsumsq = colsum.data_ptr(), colsumsq.data_ptr()
    if stat_x is not None:
        assert stat_x.dtype == torch.bfloat16 and stat_x.shape == (M, N) and stat_x.stride(1) == 1 and colsum is not None
        args.stat_x, args.ld_stat_x = stat_x.data_ptr(), stat_x.stride(0)
        args.stat_mean, args.stat_scale, args.stat_shift = stat_mean.data_ptr(), stat_scale.data_ptr(), stat_shift.data_ptr()
        args.stat_relu = int(bool(stat_relu))
    check(_lib.lib().mdhs_gemm_bf16(ctypes.byref(args), _stream()), "mdhs_gemm_bf16")
    return out


# --------------------------------------------------------------------------------------------
# thin wrappers for the remaining entry points (argument order = include/mdhs_b200.h)
# --------------------------------------------------------------------------------------------
def _p(t):
    return None if t is None else t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


def layernorm_fwd(x, gamma, beta, eps, *, out_bf16=True, out_f32=False, drop_p=0.0, seed=0, save_stats=True):
    """x: [rows, C] bf16 or fp32 (row stride may exceed C).  Returns (y_bf16|None, y_f32|None, mean, rstd)."""
    _require_cuda(x, gamma, beta)
    rows, C = x.shape
    y = torch.empty((rows, C), device=x.device, dtype=torch.bfloat16) if out_bf16 else None
    y32 = torch.empty((rows, C), device=x.device, dtype=torch.float32) if out_f32 else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    _lib.call("mdhs_layernorm_fwd", _p(x), int(x.dtype == torch.float32), x.stride(0), _p(gamma), _p(beta), _p(y), C,
              _p(y32), _p(mean), _p(rstd), rows, C, float(eps), float(drop_p), int(seed), _s())
    return y, y32, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, *, dx_bf16=True, dx_f32=False, drop_p=0.0, seed=0,
                  drop2_p=0.0, seed2=0, want_dx_drop=False, dbias=None):
    """dbias (optional fp32 [C]): += column sums of the gradient handed on to the dense branch (dx_drop if requested, else
    dx), i.e. the bias gradient of the Linear that fed this LayerNorm's residual sum."""
    rows, C = x.shape
    dx = torch.empty((rows, C), device=x.device, dtype=torch.bfloat16) if dx_bf16 else None
    dxd = torch.empty((rows, C), device=x.device, dtype=torch.bfloat16) if want_dx_drop else None
    dx32 = torch.empty((rows, C), device=x.device, dtype=torch.float32) if dx_f32 else None
    _lib.call("mdhs_layernorm_bwd", _p(dy), int(dy.dtype == torch.float32), dy.stride(0), _p(x),
              int(x.dtype == torch.float32), x.stride(0), _p(mean), _p(rstd), _p(gamma), _p(dx), C, _p(dxd), _p(dx32),
              _p(dgamma), _p(dbeta), _p(dbias), rows, C, float(drop_p), int(seed), float(drop2_p), int(seed2), _s())
    return dx, dxd, dx32


def bn_finalize(colsum, colsumsq, count, gamma, beta, running_mean, running_var, momentum, eps, training=True):
    C = gamma.numel()
    dev = gamma.device
    mean = torch.empty(C, device=dev, dtype=torch.float32)
    invstd = torch.empty(C, device=dev, dtype=torch.float32)
    scale = torch.empty(C, device=dev, dtype=torch.float32)
    shift = torch.empty(C, device=dev, dtype=torch.float32)
    _lib.call("mdhs_bn_finalize", _p(colsum), _p(colsumsq), int(count), _p(gamma), _p(beta), _p(running_mean),
              _p(running_var), float(momentum), float(eps), _p(mean), _p(invstd), _p(scale), _p(shift), C, int(training), _s())
    return mean, invstd, scale, shift


def bn_apply(x, scale, shift, residual=None, relu=True, out=None):
    rows, C = x.shape
    y = torch.empty_like(x) if out is None else out
    _lib.call("mdhs_bn_apply", _p(x), _p(scale), _p(shift), _p(residual), _p(y), rows, C, int(relu), _s())
    return y


def bn_fwd(x, colsum, colsumsq, gamma, beta, running_mean, running_var, momentum, eps, residual=None, relu=True, training=True,
           want_mask=False):
    """Finalize + apply in one C-ABI call.  Returns (y, mean, invstd, scale, shift[, relu bit mask uint8 [rows, C/8]])."""
    rows, C = x.shape
    stats = torch.empty((4, C), device=x.device, dtype=torch.float32)
    y = torch.empty_like(x)
    mask = torch.empty((rows, C // 8), device=x.device, dtype=torch.uint8) if (want_mask and relu) else None
    _lib.call("mdhs_bn_fwd", _p(x), _p(colsum), _p(colsumsq), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
              float(momentum), float(eps), _p(residual), _p(y), stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(),
              stats[3].data_ptr(), _p(mask), rows, C, int(relu), int(training), _s())
    if want_mask:
        return y, stats[0], stats[1], stats[2], stats[3], mask
    return y, stats[0], stats[1], stats[2], stats[3]


def bn_bwd(dy, x, y, mean, invstd, gamma, dgamma, dbeta, relu=True, want_dz=False, scale=None, shift=None, training=True,
           sums=None, mask=None):
    """y may be None when relu and the layer had no residual input: the mask is recomputed from (x, scale, shift).
    sums: optional fp64 [2, C] workspace already holding sum(dy'), sum(dy' * (x - mean)) (GEMM-epilogue fused reduction)."""
    rows, C = x.shape
    ws = torch.empty((2, C), device=x.device, dtype=torch.float64) if sums is None else sums
    coef = torch.empty((5, C), device=x.device, dtype=torch.float32)
    dx = torch.empty_like(x)
    dz = torch.empty_like(x) if want_dz else None
    _lib.call("mdhs_bn_bwd", _p(dy), _p(x), _p(y), _p(mask), _p(mean), _p(invstd), _p(gamma), _p(scale), _p(shift), ws[0].data_ptr(),
              ws[1].data_ptr(), _p(coef), _p(dx), _p(dz), _p(dgamma), _p(dbeta), rows, C, int(relu), int(training),
              int(sums is not None), _s())
    return dx, dz


def col_stats(x, sum64=None, sumsq64=None, sum32=None):
    rows, C = x.shape
    _lib.call("mdhs_col_stats", _p(x), x.stride(0), _p(sum64), _p(sumsq64), _p(sum32), rows, C, _s())


TTA_CODES = {"identity": 0, "hflip": 1, "vflip": 2, "rot90": 3}


def tta_codes(transforms):
    """Packed 4-bit transform ids of [identity] + transforms (scripts/predict.py:33-42)."""
    names = ["identity"] + list(transforms)
    codes = 0
    for v, n in enumerate(names):
        if n not in TTA_CODES:
            raise ValueError(f"unknown TTA transform {n!r}")
        codes |= TTA_CODES[n] << (4 * v)
    return len(names), codes


def im2col_nchw_f32(x, R, S, stride, pad, ldc, tta=None):
    """tta = (V, codes): the patch matrix holds V augmented variants of every image (variant-major), read from x directly."""
    B, C, H, W = x.shape
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    V = 1 if tta is None else tta[0]
    col = torch.empty((V * B * Ho * Wo, ldc), device=x.device, dtype=torch.bfloat16)
    if tta is None:
        _lib.call("mdhs_im2col_nchw_f32", _p(x), _p(col), B, C, H, W, R, S, stride, pad, ldc, _s())
    else:
        _lib.call("mdhs_im2col_nchw_f32_tta", _p(x), _p(col), B, C, H, W, R, S, stride, pad, ldc, V, int(tta[1]), _s())
    return col, Ho, Wo


def im2col_nhwc(x, B, H, W, C, R, S, stride, pad):
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    col = torch.empty((B * Ho * Wo, R * S * C), device=x.device, dtype=torch.bfloat16)
    _lib.call("mdhs_im2col_nhwc", _p(x), _p(col), B, H, W, C, R, S, stride, pad, _s())
    return col, Ho, Wo


def col2im_nhwc(dcol, B, H, W, C, R, S, stride, pad, add=None):
    dx = torch.empty((B * H * W, C), device=dcol.device, dtype=torch.bfloat16)
    _lib.call("mdhs_col2im_nhwc", _p(dcol), _p(add), _p(dx), B, H, W, C, R, S, stride, pad, _s())
    return dx


def maxpool_fwd(x, B, H, W, C):
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((B * Ho * Wo, C), device=x.device, dtype=torch.bfloat16)
    idx = torch.empty((B * Ho * Wo, C), device=x.device, dtype=torch.uint8)
    _lib.call("mdhs_maxpool3x3s2_fwd", _p(x), _p(y), _p(idx), B, H, W, C, _s())
    return y, idx, Ho, Wo


def maxpool_bwd(dy, idx, B, H, W, C):
    dx = torch.empty((B * H * W, C), device=dy.device, dtype=torch.bfloat16)
    _lib.call("mdhs_maxpool3x3s2_bwd", _p(dy), _p(idx), _p(dx), B, H, W, C, _s())
    return dx


def mean_tokens_fwd(x, B, T, C, scale=None, out32=None, accumulate=False, want_bf16=False):
    scale = 1.0 / T if scale is None else scale
    if out32 is None and not want_bf16:
        out32 = torch.empty((B, C), device=x.device, dtype=torch.float32)
    y16 = torch.empty((B, C), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    _lib.call("mdhs_mean_tokens_fwd", _p(x), _p(out32), _p(y16), B, T, C, float(scale), int(accumulate), _s())
    return out32, y16


def mean_tokens_bwd(dy, B, T, C, scale=None):
    scale = 1.0 / T if scale is None else scale
    dx = torch.empty((B * T, C), device=dy.device, dtype=torch.bfloat16)
    if dy.dtype == torch.float32:
        _lib.call("mdhs_mean_tokens_bwd", _p(dy), None, _p(dx), B, T, C, float(scale), _s())
    else:
        _lib.call("mdhs_mean_tokens_bwd", None, _p(dy), _p(dx), B, T, C, float(scale), _s())
    return dx


def conv_weight_pack(w, ldk=None, out=None):
    O, I, R, S = w.shape
    ldk = R * S * I if ldk is None else ldk
    wp = torch.empty((O, ldk), device=w.device, dtype=torch.bfloat16) if out is None else out
    _lib.call("mdhs_conv_weight_pack", _p(w), _p(wp), O, I, R, S, ldk, _s())
    return wp


def conv_weight_pack_dgrad(w, out=None):
    """OIHW fp32 -> bf16 [I, (R-1-r, S-1-s, o)]: the B operand of the stride-1 implicit-GEMM dgrad."""
    O, I, R, S = w.shape
    wt = torch.empty((I, R * S * O), device=w.device, dtype=torch.bfloat16) if out is None else out
    _lib.call("mdhs_conv_weight_pack_dgrad", _p(w), _p(wt), O, I, R, S, _s())
    return wt


def conv_wgrad_unpack(gp, g):
    O, I, R, S = g.shape
    _lib.call("mdhs_conv_wgrad_unpack", _p(gp), _p(g), O, I, R, S, gp.stride(0), _s())


def cast_f32_bf16(x, out=None):
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if out is None else out
    _lib.call("mdhs_cast_f32_bf16", _p(x), _p(y), x.numel(), _s())
    return y


def cast_bf16_f32(x, out=None):
    y = torch.empty(x.shape, device=x.device, dtype=torch.float32) if out is None else out
    _lib.call("mdhs_cast_bf16_f32", _p(x), _p(y), x.numel(), _s())
    return y


def nhwc_bf16_to_nchw_f32(x, B, H, W, C):
    y = torch.empty((B, C, H, W), device=x.device, dtype=torch.float32)
    _lib.call("mdhs_nhwc_bf16_to_nchw_f32", _p(x), _p(y), B, H, W, C, _s())
    return y


def nchw_f32_to_nhwc_bf16(x):
    B, C, H, W = x.shape
    y = torch.empty((B * H * W, C), device=x.device, dtype=torch.bfloat16)
    _lib.call("mdhs_nchw_f32_to_nhwc_bf16", _p(x), _p(y), B, H, W, C, _s())
    return y


def attention_fwd(q, k, v, B, H, Sq, Sk, D, scale, key_mask=None, drop_p=0.0, seed=0, out=None):
    """q/k/v: 2-D token-major bf16 views [B*S, >= H*D] (row stride free)."""
    o = torch.empty((B * Sq, H * D), device=q.device, dtype=torch.bfloat16) if out is None else out
    lse = torch.empty((B * H * Sq,), device=q.device, dtype=torch.float32)
    _lib.call("mdhs_attention_fwd", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(o), o.stride(0),
              _p(key_mask), _p(lse), B, H, Sq, Sk, D, float(scale), float(drop_p), int(seed), _s())
    return o, lse


def attention_bwd(q, k, v, o, d_o, lse, B, H, Sq, Sk, D, scale, key_mask=None, drop_p=0.0, seed=0, dq=None, dk=None, dv=None,
                  delta_scratch=True):
    """dq/dk/dv: optional pre-allocated bf16 views with the same row strides as q/k/v."""
    if dq is None:
        dq = torch.empty((B * Sq, H * D), device=q.device, dtype=torch.bfloat16)
        assert q.stride(0) == H * D
    if dk is None:
        dk = torch.empty((B * Sk, H * D), device=q.device, dtype=torch.bfloat16)
        dv = torch.empty((B * Sk, H * D), device=q.device, dtype=torch.bfloat16)
        assert k.stride(0) == H * D and v.stride(0) == H * D
    assert dq.stride(0) == q.stride(0) and dk.stride(0) == k.stride(0) and dv.stride(0) == v.stride(0)
    assert d_o.stride(0) == o.stride(0)
    dk32 = dv32 = None
    fp32_path = bool(_lib.lib().mdhs_attention_bwd_workspace(Sq, Sk, D))   # a query, not a status
    if fp32_path:
        dk32 = torch.zeros((B * Sk, k.stride(0)), device=q.device, dtype=torch.float32)
        dv32 = torch.zeros((B * Sk, v.stride(0)), device=q.device, dtype=torch.float32)
    elif Sq > 64 and delta_scratch:
        dk32 = torch.empty((B * H * Sq,), device=q.device, dtype=torch.float32)   # delta, dQ pass -> dK/dV pass
    _lib.call("mdhs_attention_bwd", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(o), _p(d_o), o.stride(0),
              _p(key_mask), _p(lse), _p(dq), _p(dk), _p(dv), _p(dk32), _p(dv32), B, H, Sq, Sk, D, float(scale),
              float(drop_p), int(seed), _s())
    if fp32_path:
        # fp32 accumulators cover the full row stride; copy the head columns back as bf16
        dk.copy_(cast_f32_bf16(dk32)[:, :dk.shape[1]])
        dv.copy_(cast_f32_bf16(dv32)[:, :dv.shape[1]])
    return dq, dk, dv


def embed_gather(ids, type_ids, word, pos, type_emb, S):
    rows = ids.numel()
    C = word.shape[1]
    e = torch.empty((rows, C), device=word.device, dtype=torch.float32)
    _lib.call("mdhs_embed_gather", _p(ids), _p(type_ids), _p(word), _p(pos), _p(type_emb), _p(e), rows, S, C, word.shape[0], _s())
    return e


def embed_scatter(de, ids, type_ids, gword, gpos, gtype, S):
    rows, C = de.shape
    vocab = gword.shape[0] if gword is not None else 1 << 30
    _lib.call("mdhs_embed_scatter", _p(de), _p(ids), _p(type_ids), _p(gword), _p(gpos), _p(gtype), rows, S, C, vocab, _s())


def linear_f32_fwd(x, w, bias=None, act=ACT_NONE):
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), device=x.device, dtype=torch.float32)
    _lib.call("mdhs_linear_f32_fwd", _p(x), x.stride(0), _p(w), _p(bias), _p(y), N, M, N, K, act, _s())
    return y


def linear_f32_bwd(dy, x, w, dw=None, db=None, need_dx=True):
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty((M, K), device=dy.device, dtype=torch.float32) if need_dx else None
    _lib.call("mdhs_linear_f32_bwd", _p(dy), dy.stride(0), _p(x), x.stride(0) if x is not None else 0, _p(w), _p(dx), K, 0,
              _p(dw), _p(db), M, N, K, _s())
    return dx


def ce_loss(logits, labels, class_weights=None, label_smoothing=0.0, focal=False, gamma=2.0, want_grad=True):
    B, C = logits.shape
    loss = torch.empty(1, device=logits.device, dtype=torch.float32)
    dl = torch.empty((B, C), device=logits.device, dtype=torch.float32) if want_grad else None
    _lib.call("mdhs_ce_loss", _p(logits), logits.stride(0), _p(labels), _p(class_weights), _p(loss), _p(dl), B, C,
              float(label_smoothing), int(focal), float(gamma), _s())
    return loss, dl


def supcon_loss(x, labels, temperature=0.07, want_grad=True):
    B, D = x.shape
    dev = x.device
    loss = torch.empty(1, device=dev, dtype=torch.float32)
    f = torch.empty((B, D), device=dev, dtype=torch.float32)
    inv = torch.empty(B, device=dev, dtype=torch.float32)
    g = torch.empty((B, B), device=dev, dtype=torch.float32) if want_grad else None
    dx = torch.empty((B, D), device=dev, dtype=torch.float32) if want_grad else None
    _lib.call("mdhs_supcon_loss", _p(x), x.stride(0), _p(labels), _p(loss), _p(dx), D, _p(f), _p(inv), _p(g), B, D,
              float(temperature), _s())
    return loss, dx


def axpby(x, y, a=1.0, b=0.0, a_dev=None):
    _lib.call("mdhs_axpby_f32", _p(x), _p(y), x.numel(), _p(a_dev), float(a), float(b), _s())
    return y


HAS_GEMM_STAT = True   # mdhs_gemm_args.stat_x (BatchNorm-backward reduction in the dgrad epilogue) is available


_NUM_SMS = {}


def num_sms(device=None):
    """SM count of the (current) CUDA device -- 148 on B200 -- for host-side split / grid heuristics."""
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    n = _NUM_SMS.get(dev)
    if n is None:
        n = _NUM_SMS[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    return n


def set_sm_reserve(n):
    """Persistent GEMM grids leave n SMs free (for a concurrently running collective kernel); 0 = all SMs."""
    _lib.call("mdhs_set_sm_reserve", int(n))


def set_gemm_dynamic(on):
    """Dynamic (work-counter) vs static tile scheduling of the persistent GEMM grids."""
    _lib.call("mdhs_set_gemm_dynamic", int(bool(on)))


def adam_flat(params, grads, exp_avg, exp_avg_sq, shadow, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0,
              adamw=True, zero_grad=True, lr_dev=None, step_dev=None, grads_bf16=None, blocks_per_sm=0):
    _lib.call("mdhs_adam_flat", _p(params), _p(grads), _p(grads_bf16), _p(exp_avg), _p(exp_avg_sq), _p(shadow), params.numel(), float(lr),
              float(beta1), float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale), int(adamw),
              int(zero_grad), _p(lr_dev), _p(step_dev), int(blocks_per_sm), _s())


def sgd_flat(params, grads, mom, shadow, lr, momentum, weight_decay, grad_scale=1.0, first_step=False, zero_grad=True,
             lr_dev=None, step_dev=None, grads_bf16=None, blocks_per_sm=0):
    _lib.call("mdhs_sgd_flat", _p(params), _p(grads), _p(grads_bf16), _p(mom), _p(shadow), params.numel(), float(lr), float(momentum),
              float(weight_decay), float(grad_scale), int(first_step), int(zero_grad), _p(lr_dev), _p(step_dev),
              int(blocks_per_sm), _s())


def step_begin(step_dev=None):
    _lib.call("mdhs_step_begin", _p(step_dev), _s())


def act_dropout_bwd(dy, aux, act, drop_p, seed):
    g = torch.empty_like(dy)
    _lib.call("mdhs_act_dropout_bwd", _p(dy), _p(aux), _p(g), dy.numel(), act, float(drop_p), int(seed), _s())
    return g


def relu_bwd_f32(dy, y):
    dx = torch.empty_like(dy)
    _lib.call("mdhs_relu_bwd_f32", _p(dy), _p(y), _p(dx), dy.numel(), _s())
    return dx


def sum64_to_grad(sum64, sumsq64, grad):
    """grad += sum64 (fp64 column sums from a GEMM epilogue); clears both workspaces."""
    assert sum64.dtype == torch.float64 and grad.dtype == torch.float32 and sum64.is_contiguous() and grad.is_contiguous()
    _lib.call("mdhs_sum64_to_grad", _p(sum64), _p(sumsq64), _p(grad), grad.numel(), _s())


def mul_f32(a, b):
    c = torch.empty_like(a)
    _lib.call("mdhs_mul_f32", _p(a), _p(b), _p(c), a.numel(), _s())
    return c


def dropout_f32(x, p, seed):
    y = torch.empty_like(x)
    _lib.call("mdhs_dropout_f32", _p(x), _p(y), x.numel(), float(p), int(seed), _s())
    return y


def ibfa_fwd(kqv_x, kv_y, B, H, D):
    out = torch.empty((B, H * D), device=kqv_x.device, dtype=torch.bfloat16)
    probs = torch.empty((B * H, 2), device=kqv_x.device, dtype=torch.float32)
    _lib.call("mdhs_ibfa_fwd", _p(kqv_x), kqv_x.stride(0), _p(kv_y), kv_y.stride(0), _p(out), _p(probs), B, H, D, _s())
    return out, probs


def ibfa_bwd(kqv_x, kv_y, dout, probs, B, H, D):
    dx = torch.empty_like(kqv_x)
    dy = torch.empty_like(kv_y)
    _lib.call("mdhs_ibfa_bwd", _p(kqv_x), kqv_x.stride(0), _p(kv_y), kv_y.stride(0), _p(dout), _p(probs), _p(dx), _p(dy), B, H, D, _s())
    return dx, dy


def mp_loss(zi, zt, zf, labels, want_grad=True):
    B, C = zi.shape
    loss = torch.empty(1, device=zi.device, dtype=torch.float32)
    g = [torch.empty_like(zi) for _ in range(3)] if want_grad else [None, None, None]
    _lib.call("mdhs_mp_loss", _p(zi), _p(zt), _p(zf), _p(labels), _p(loss), _p(g[0]), _p(g[1]), _p(g[2]), B, C, _s())
    return loss, g


# ---------------------------------------------------------------- KAN / MoE (csrc/kan_moe.cu)
KAN_NB = 8  # grid_size 5 + spline_order 3 bases per input


def kan_basis_fwd(x, grid, ld_op=None):
    """x fp32 [rows, in] (row stride free) -> bf16 [rows, ld_op] = [SiLU(x) | B_0..7(x_0) | B_0..7(x_1) ...]."""
    rows, n_in = x.shape
    ld_op = n_in * (1 + KAN_NB) if ld_op is None else ld_op
    op = torch.empty((rows, ld_op), device=x.device, dtype=torch.bfloat16)
    _lib.call("mdhs_kan_basis_fwd", _p(x), x.stride(0), _p(grid), _p(op), rows, n_in, ld_op, _s())
    return op


def kan_basis_bwd(x, grid, dop, dx=None, accumulate=False):
    rows, n_in = x.shape
    if dx is None:
        dx = torch.empty((rows, n_in), device=x.device, dtype=torch.float32)
        accumulate = False
    _lib.call("mdhs_kan_basis_bwd", _p(x), x.stride(0), _p(grid), _p(dop), _p(dx), rows, n_in, dop.stride(0), int(accumulate), _s())
    return dx


def kan_weight_pack(base_w, spline_w, scaler, wcat):
    out, n_in = base_w.shape
    _lib.call("mdhs_kan_weight_pack", _p(base_w), _p(spline_w), _p(scaler), _p(wcat), out, wcat.shape[0], n_in, wcat.stride(0), _s())
    return wcat


def kan_wgrad_unpack(gcat, spline_w, scaler, g_base, g_spline, g_scaler):
    out, n_in = spline_w.shape[0], spline_w.shape[1]
    _lib.call("mdhs_kan_wgrad_unpack", _p(gcat), _p(spline_w), _p(scaler), _p(g_base), _p(g_spline), _p(g_scaler), out, n_in,
              gcat.stride(0), _s())


def randn_f32(shape, device, seed):
    out = torch.empty(shape, device=device, dtype=torch.float32)
    _lib.call("mdhs_randn_f32", _p(out), out.numel(), int(seed), _s())
    return out


def moe_gate_fwd(x, wg, wn, noise, k, noisy, nmean=None, nstd=None):
    B, n_in = x.shape
    E = wg.shape[1]
    dev = x.device
    f = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
    gates, clean, probs = f(B, E), f(B, E), f(B, E)
    raw = f(B, E) if noisy else None
    topidx = torch.empty((B, k + 1), device=dev, dtype=torch.int32)
    importance, load = f(E), f(E)
    _lib.call("mdhs_moe_gate_fwd", _p(x), _p(wg), _p(wn), _p(noise), _p(gates), _p(clean), _p(raw), _p(probs), _p(topidx),
              _p(importance), _p(load), _p(nmean), _p(nstd), B, n_in, E, k, int(noisy), _s())
    return gates, clean, raw, probs, topidx, importance, load


def moe_loss(importance, load, coef):
    E = importance.numel()
    loss = torch.empty(1, device=importance.device, dtype=torch.float32)
    d_imp = torch.empty_like(importance)
    d_load = torch.empty_like(load)
    _lib.call("mdhs_moe_loss", _p(importance), _p(load), _p(loss), _p(d_imp), _p(d_load), E, float(coef), _s())
    return loss, d_imp, d_load


def moe_gate_bwd(x, wg, wn, noise, dgates, d_imp, d_load, dloss_dev, clean, raw, probs, topidx, dx, dwg, dwn, k, noisy,
                 nmean=None, nstd=None):
    B, n_in = x.shape
    E = wg.shape[1]
    _lib.call("mdhs_moe_gate_bwd", _p(x), _p(wg), _p(wn), _p(noise), _p(dgates), _p(d_imp), _p(d_load), 1.0, _p(dloss_dev),
              _p(clean), _p(raw), _p(probs), _p(topidx), _p(dx), _p(dwg), _p(dwn), _p(nmean), _p(nstd), B, n_in, E, k, int(noisy), _s())


def moe_combine_fwd(gates, Y, C):
    """Y fp32 [E, B, ldy] -> y fp32 [B, C] = sum_e gates[b, e] * Y[e, b, :C]."""
    E, B, ldy = Y.shape
    y = torch.empty((B, C), device=Y.device, dtype=torch.float32)
    _lib.call("mdhs_moe_combine_fwd", _p(gates), _p(Y), _p(y), B, E, C, ldy, _s())
    return y


def moe_combine_bwd(gates, Y, dy):
    E, B, ldy = Y.shape
    C = dy.shape[1]
    dgates = torch.empty_like(gates)
    dY = torch.empty_like(Y)
    _lib.call("mdhs_moe_combine_bwd", _p(gates), _p(Y), _p(dy), _p(dgates), _p(dY), B, E, C, ldy, _s())
    return dgates, dY


# --------------------------------------------------------------------------------------------
# ConvNeXt pieces (csrc/convnext.cu)
# --------------------------------------------------------------------------------------------
def dwconv7(x, w, bias, B, H, W, flip=False):
    """Depthwise 7x7 (pad 3) on NHWC bf16 [B*H*W, C]; w is the Conv2d weight [C,1,7,7] fp32.  flip=True: input gradient."""
    C = x.shape[1]
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.shape[0] == B * H * W and w.is_contiguous()
    y = torch.empty_like(x)
    _lib.call("mdhs_dwconv7_fwd", _p(x), _p(w), _p(bias), _p(y), B, H, W, C, int(flip), _s())
    return y


def dwconv7_wgrad(x, dy, dw, db, B, H, W):
    C = x.shape[1]
    assert x.is_contiguous() and dy.is_contiguous() and dw.is_contiguous()
    _lib.call("mdhs_dwconv7_wgrad", _p(x), _p(dy), _p(dw), _p(db), B, H, W, C, _s())


def layer_scale_fwd(x, z, ls, rows_per_sample, p=0.0, seed=0):
    rows, C = x.shape
    out = torch.empty_like(x)
    _lib.call("mdhs_layer_scale_fwd", _p(x), _p(z), _p(ls), _p(out), rows, C, int(rows_per_sample), float(p), int(seed), _s())
    return out


def layer_scale_bwd(dy, z, ls, dls, rows_per_sample, p=0.0, seed=0, dbias=None):
    rows, C = dy.shape
    dz = torch.empty_like(dy)
    _lib.call("mdhs_layer_scale_bwd", _p(dy), _p(z), _p(ls), _p(dz), _p(dls), _p(dbias), rows, C, int(rows_per_sample), float(p), int(seed), _s())
    return dz


def sq_attn_fwd(q, k, v, B, T, scale=1.0):
    D = q.shape[1]
    out = torch.empty((B, D), device=q.device, dtype=torch.float32)
    probs = torch.empty((B, T), device=q.device, dtype=torch.float32)
    _lib.call("mdhs_sq_attn_fwd", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), _p(probs), B, T, D,
              float(scale), _s())
    return out, probs


def sq_attn_bwd(q, k, v, dout, probs, B, T, scale=1.0):
    D = q.shape[1]
    dq = torch.empty((B, D), device=q.device, dtype=torch.bfloat16)
    dk = torch.empty((B * T, D), device=q.device, dtype=torch.bfloat16)
    dv = torch.empty((B * T, D), device=q.device, dtype=torch.bfloat16)
    _lib.call("mdhs_sq_attn_bwd", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(dout), _p(probs), _p(dq),
              dq.stride(0), _p(dk), dk.stride(0), _p(dv), dv.stride(0), B, T, D, float(scale), _s())
    return dq, dk, dv


def tta_expand(images, transforms):
    """[B,3,H,W] fp32 -> [V*B,3,H,W]: identity followed by the named transforms (scripts/predict.py:33-42)."""
    B, C, H, W = images.shape
    names = ["identity"] + list(transforms)
    codes = 0
    for v, n in enumerate(names):
        if n not in TTA_CODES:
            raise ValueError(f"unknown TTA transform {n!r}")
        codes |= TTA_CODES[n] << (4 * v)
    x = images.contiguous().float()
    y = torch.empty((len(names) * B, C, H, W), device=x.device, dtype=torch.float32)
    _lib.call("mdhs_tta_expand", _p(x), _p(y), B, C, H, W, len(names), codes, _s())
    return y


def _ptr_array(tensors):
    import ctypes as _ct
    arr = (_ct.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


def level_mix_fwd(levels, logits):
    """out = sum_l softmax(logits)_l * levels[l]; levels: list of same-shaped contiguous fp32 tensors (L <= 4)."""
    import ctypes as _ct
    out = torch.empty_like(levels[0])
    arr = _ptr_array(levels)
    _lib.call("mdhs_level_mix_fwd", _ct.cast(arr, _ct.c_void_p), _p(logits), _p(out), out.numel(), len(levels), _s())
    return out


def level_mix_bwd(levels, logits, dout, dlogits):
    import ctypes as _ct
    dps = [torch.empty_like(t) for t in levels]
    ws = torch.empty(4, device=dout.device, dtype=torch.float32)
    pa, da = _ptr_array(levels), _ptr_array(dps)
    _lib.call("mdhs_level_mix_bwd", _ct.cast(pa, _ct.c_void_p), _ct.cast(da, _ct.c_void_p), _p(logits), _p(dout), _p(ws),
              _p(dlogits), dout.numel(), len(levels), _s())
    return dps


def axpby_bf16(x, y, a, b):
    """out = a * x + b * y on bf16 tensors (y may be None)."""
    out = torch.empty_like(x)
    _lib.call("mdhs_axpby_bf16", _p(x), _p(y), _p(out), x.numel(), float(a), float(b), _s())
    return out


def global_local(images, crop_ratio):
    """[B,3,H,W] fp32 -> [2B,3,H,W]: the images followed by their centre crop resized back (model.py:292-301)."""
    B, C, H, W = images.shape
    x = images.contiguous().float()
    y = torch.empty((2 * B, C, H, W), device=x.device, dtype=torch.float32)
    _lib.call("mdhs_global_local", _p(x), _p(y), B, C, H, W, float(crop_ratio), _s())
    return y


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)   # data_loader.py:358


def preprocess_u8(images_u8, out_hw=(224, 224), boxes=None, flips=None, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """GPU-side replacement of the per-sample torchvision transforms after the decode (data_loader.py:343-372):
    uint8 [B, Hs, Ws, 3] CUDA batch -> crop boxes [B, 4] (y0, x0, h, w) -> bilinear resize -> flips [B] uint8 (bit 0 h, bit 1 v)
    -> ToTensor -> Normalize -> fp32 [B, 3, H, W].  The host only decodes and draws the random crop / flip parameters."""
    import ctypes as _ct
    B, Hs, Ws, C = images_u8.shape
    assert C == 3 and images_u8.dtype == torch.uint8 and images_u8.is_cuda and images_u8.is_contiguous()
    out = torch.empty((B, 3, out_hw[0], out_hw[1]), device=images_u8.device, dtype=torch.float32)
    m3 = (_ct.c_float * 3)(*[float(v) for v in mean])
    s3 = (_ct.c_float * 3)(*[float(v) for v in std])
    if boxes is not None:
        boxes = boxes.to(device=images_u8.device, dtype=torch.float32).contiguous()
    if flips is not None:
        flips = flips.to(device=images_u8.device, dtype=torch.uint8).contiguous()
    _lib.call("mdhs_preprocess_u8", _p(images_u8), _p(out), _p(boxes), _p(flips), B, Hs, Ws, out_hw[0], out_hw[1],
              _ct.cast(m3, _ct.c_void_p), _ct.cast(s3, _ct.c_void_p), _s())
    return out


def lstm_cell_fwd(gates, c_prev):
    B, H4 = gates.shape
    H = H4 // 4
    h = torch.empty((B, H), device=gates.device, dtype=torch.float32)
    c = torch.empty_like(h)
    act = torch.empty_like(gates)
    _lib.call("mdhs_lstm_cell_fwd", _p(gates), _p(c_prev), _p(h), _p(c), _p(act), B, H, _s())
    return h, c, act


def lstm_cell_bwd(dh, dc, act, c_prev, c):
    B, H = c.shape
    dgates = torch.empty_like(act)
    dc_prev = torch.empty_like(c)
    _lib.call("mdhs_lstm_cell_bwd", _p(dh), _p(dc), _p(act), _p(c_prev), _p(c), _p(dgates), _p(dc_prev), B, H, _s())
    return dgates, dc_prev


def gru_cell_fwd(gi, gh, h_prev):
    B, H3 = gi.shape
    h = torch.empty((B, H3 // 3), device=gi.device, dtype=torch.float32)
    act = torch.empty_like(gi)
    _lib.call("mdhs_gru_cell_fwd", _p(gi), _p(gh), _p(h_prev), _p(h), _p(act), B, H3 // 3, _s())
    return h, act


def gru_cell_bwd(dh, act, gh, h_prev):
    B, H = h_prev.shape
    dgi, dgh, dhp = torch.empty_like(act), torch.empty_like(act), torch.empty_like(h_prev)
    _lib.call("mdhs_gru_cell_bwd", _p(dh), _p(act), _p(gh), _p(h_prev), _p(dgi), _p(dgh), _p(dhp), B, H, _s())
    return dgi, dgh, dhp
