"""Thin Python wrappers: torch tensors -> raw device pointers -> C ABI (include/mdhs_b200.h).

torch is plumbing here (device memory, streams); every numeric op is one of our sm_100a kernels.
"""
import ctypes

import torch

from . import _lib
from ._lib import GemmArgs, check

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
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
         bn_hint=0, colsum=None, colsumsq=None, M=None, N=None, K=None):
    """D[M,N] (+)= epi(A . B^T).  `a` is [M,K] (K-major) or [K,M] when a_mn; `b` is [N,K] or [K,N] when b_mn.
    2-D bf16 tensors with unit inner stride (row stride may exceed the row length)."""
    _require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    if a_mn:
        k_a, m = a.shape
    else:
        m, k_a = a.shape
    if b_mn:
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
    if colsum is not None:
        assert colsum.dtype == torch.float64 and colsumsq.dtype == torch.float64
        args.colsum, args.colsumsq = colsum.data_ptr(), colsumsq.data_ptr()
    check(_lib.lib().mdhs_gemm_bf16(ctypes.byref(args), _stream()), "mdhs_gemm_bf16")
    return out
