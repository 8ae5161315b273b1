"""Parity of the tcgen05 GEMM (include/mdhs_b200.h: mdhs_gemm_bf16) against fp32 matmul on the same
bf16-rounded inputs.  Tolerance: fp32 accumulation of bf16 products -> max abs err <= 2e-2 * scale for
bf16 outputs (one bf16 rounding of the result), <= 2e-3 relative for fp32 outputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float().t() if b_mn else b.float()
    return A @ B.t()


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K,bn", [
    (128, 64, 64, 64), (256, 128, 128, 128), (384, 256, 192, 256), (200, 136, 72, 0),
    (1000, 264, 520, 0), (8192, 768, 768, 0), (136, 2304, 768, 256),
])
def test_gemm_layouts(mdhs, M, N, K, bn, a_mn, b_mn):
    from mdhs_b200 import ops
    torch.manual_seed(M * 7 + N * 3 + K)
    a = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
    ref = _ref(a, b, a_mn, b_mn)
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, bn_hint=bn)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, (err, scale)
    out16 = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bn_hint=bn)
    err16 = (out16.float() - ref).abs().max().item()
    assert err16 <= 1e-2 * scale, (err16, scale)


def test_gemm_epilogues(mdhs):
    from mdhs_b200 import ops
    torch.manual_seed(0)
    M, N, K = 520, 328, 264
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16()
    pre = a.float() @ w.float().t() + bias
    # bias + gelu + residual, with the pre-activation saved
    aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, residual=res, aux_out=aux, out_dtype=torch.float32)
    ref = torch.nn.functional.gelu(pre) + res.float()
    assert (out - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    assert (aux.float() - pre).abs().max().item() <= 1e-2 * pre.abs().max().item()
    # relu + colsum statistics on bf16 output
    cs = torch.zeros(N, device="cuda", dtype=torch.float64)
    cq = torch.zeros(N, device="cuda", dtype=torch.float64)
    out = ops.gemm(a, w, bias=bias, act=ops.ACT_RELU, colsum=cs, colsumsq=cq)
    ref = torch.relu(pre)
    assert (out.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert torch.allclose(cs, out.double().sum(0), rtol=1e-5, atol=1e-3)
    assert torch.allclose(cq, (out.double() ** 2).sum(0), rtol=1e-5, atol=1e-3)
    # dgrad-style: multiply by gelu'(aux_in)
    g = ops.gemm(a, w, aux_in=aux, dact=ops.ACT_GELU, out_dtype=torch.float32)
    x = aux.float()
    gp = 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * torch.pi) ** 0.5
    ref = (a.float() @ w.float().t()) * gp
    assert (g - ref).abs().max().item() <= 3e-3 * ref.abs().max().item()
    # forward GELU that saves GELU'(pre) + backward multiply by the saved derivative
    daux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU_DERIV, aux_out=daux, out_dtype=torch.float32)
    assert (out - torch.nn.functional.gelu(pre)).abs().max().item() <= 2e-3 * pre.abs().max().item()
    xp = pre
    gpp = 0.5 * (1 + torch.erf(xp / 2 ** 0.5)) + xp * torch.exp(-0.5 * xp * xp) / (2 * torch.pi) ** 0.5
    assert (daux.float() - gpp).abs().max().item() <= 1e-2
    g = ops.gemm(a, w, aux_in=daux, dact=ops.ACT_MUL, out_dtype=torch.float32)
    ref = (a.float() @ w.float().t()) * daux.float()
    assert (g - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    # relu' mask
    g = ops.gemm(a, w, aux_in=aux, dact=ops.ACT_RELU, out_dtype=torch.float32)
    ref = (a.float() @ w.float().t()) * (aux.float() > 0)
    assert (g - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K,bn", [(50000, 64, 72, 0), (50000, 256, 64, 0), (50000, 256, 64, 128), (30000, 512, 128, 64),
                                      (9000, 2048, 64, 128), (20000, 1024, 64, 256)])
def test_gemm_column_statistics_persistent(mdhs, M, N, K, bn):
    """Train-mode BN statistics fused into the conv GEMM: many tiles per persistent CTA, every tile width."""
    from mdhs_b200 import ops
    torch.manual_seed(2)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.2).bfloat16()
    cs = torch.zeros(N, device="cuda", dtype=torch.float64)
    cq = torch.zeros(N, device="cuda", dtype=torch.float64)
    out = ops.gemm(a, w, colsum=cs, colsumsq=cq, bn_hint=bn)
    ref = a.float() @ w.float().t()
    assert (out.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert torch.allclose(cs, out.double().sum(0), rtol=1e-5, atol=2e-2)
    assert torch.allclose(cq, (out.double() ** 2).sum(0), rtol=1e-5, atol=2e-2)


def test_gemm_splitk_accumulate(mdhs):
    from mdhs_b200 import ops
    torch.manual_seed(1)
    # wgrad shape: dW[N,K] = dY^T[N,M] . X[M,K]  (both operands MN-major, reduction over M rows)
    rows, n_out, k_in = 4096 + 40, 264, 200
    dy = torch.randn(rows, n_out, device="cuda").bfloat16()
    x = torch.randn(rows, k_in, device="cuda").bfloat16()
    acc = torch.ones(n_out, k_in, device="cuda")
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=acc, accumulate=True, split_k=8)
    ref = dy.float().t() @ x.float() + 1.0
    assert (acc - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


def test_gemm_strided_views(mdhs):
    from mdhs_b200 import ops
    torch.manual_seed(2)
    # operands that are column slices of a wider buffer (fused QKV style)
    buf = torch.randn(300, 768, device="cuda").bfloat16()
    a = buf[:, 256:512]
    w = torch.randn(128, 256, device="cuda").bfloat16()
    out = torch.zeros(300, 512, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, out=out[:, 128:256])
    ref = a.float() @ w.float().t()
    assert (out[:, 128:256].float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert out[:, :128].abs().max().item() == 0 and out[:, 256:].abs().max().item() == 0


def test_gemm_rejects_bad_args(mdhs):
    from mdhs_b200 import ops, _lib
    a = torch.randn(64, 60, device="cuda").bfloat16()  # K % 8 != 0
    b = torch.randn(64, 60, device="cuda").bfloat16()
    with pytest.raises(_lib.MdhsError):
        ops.gemm(a, b)


def test_gemm_dropout_epilogue(mdhs):
    """Epilogue dropout: stateless mask hash(seed, m*N+n); forward and act'-backward use the same mask."""
    from mdhs_b200 import ops
    torch.manual_seed(3)
    M, N, K, p = 256, 256, 128, 0.1
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, K, device="cuda").bfloat16()
    y0 = ops.gemm(a, w, out_dtype=torch.float32)
    y1 = ops.gemm(a, w, out_dtype=torch.float32, dropout_p=p, dropout_seed=99)
    y2 = ops.gemm(a, w, out_dtype=torch.float32, dropout_p=p, dropout_seed=99)
    assert torch.equal(y1, y2)
    kept = y1 != 0
    assert abs((~kept).float().mean().item() - p) < 0.01
    assert torch.allclose(y1[kept], y0[kept] / (1 - p), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("bn", [64, 128, 256])
@pytest.mark.parametrize("M,N,K,split", [(512, 768, 128, 1), (64, 576, 3000, 5), (2304, 768, 1024, 2), (136, 72, 264, 1)])
def test_gemm_wgrad_accumulate_all_tiles(mdhs, bn, M, N, K, split):
    """fp32 TMA reduce-add path (gradient accumulation) for every tile width."""
    from mdhs_b200 import ops
    torch.manual_seed(5)
    dy = torch.randn(K, M, device="cuda").bfloat16()
    x = torch.randn(K, N, device="cuda").bfloat16()
    acc = torch.full((M, N), 0.5, device="cuda")
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=acc, accumulate=True, split_k=split, bn_hint=bn)
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=acc, accumulate=True, split_k=split, bn_hint=bn)
    ref = 2 * (dy.float().t() @ x.float()) + 0.5
    assert (acc - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    out = ops.gemm(dy, x, a_mn=True, b_mn=True, out_dtype=torch.float32, bn_hint=bn)
    assert (out - (ref - 0.5) / 2).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,bn", [(8192, 768, 768, False, False, 0), (8192 - 128, 768, 320, False, True, 128),
                                                 (4096 + 40, 1024, 512, False, False, 256), (768, 3072, 4096, True, True, 0),
                                                 (1024, 512, 2048 + 64, True, False, 128), (640, 256, 192, False, False, 256)])
def test_gemm_cta_pair_tiles(mdhs, M, N, K, a_mn, b_mn, bn):
    """Shapes large enough for the CTA-pair (tcgen05.mma.cta_group::2, 256-row) path, incl. an odd number of row blocks
    (the second CTA of the last pair runs a phantom tile) and every operand-major combination; fp32 and bf16 outputs,
    plus the fused epilogues (bias + GELU + saved pre-activation, bf16 residual prefetch + dropout-free, column statistics)."""
    from mdhs_b200 import ops
    torch.manual_seed(5)
    a = (torch.randn((K, M) if a_mn else (M, K), device="cuda") * 0.5).bfloat16()
    b = (torch.randn((K, N) if b_mn else (N, K), device="cuda") * 0.1).bfloat16()
    A = a.float().t() if a_mn else a.float()
    Bm = b.float().t() if b_mn else b.float()
    ref = A @ Bm.t()
    scale = ref.abs().max().item()
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bn_hint=bn, out_dtype=torch.float32)
    assert (out - ref).abs().max().item() <= 2e-3 * scale + 1e-3
    acc = torch.ones(M, N, device="cuda")
    ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out=acc, accumulate=True, split_k=-1)
    assert (acc - 1 - ref).abs().max().item() <= 2e-3 * scale + 1e-3
    if not a_mn:
        bias = torch.randn(N, device="cuda")
        res = torch.randn(M, N, device="cuda").bfloat16()
        aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        y = ops.gemm(a, b, b_mn=b_mn, bias=bias, act=ops.ACT_GELU, aux_out=aux, bn_hint=bn)
        want = torch.nn.functional.gelu(ref + bias)
        assert (y.float() - want).abs().max().item() <= 1e-2 * want.abs().max().item()
        assert (aux.float() - (ref + bias)).abs().max().item() <= 1e-2 * (ref + bias).abs().max().item()
        y = ops.gemm(a, b, b_mn=b_mn, bias=bias, residual=res, bn_hint=min(bn, 128))
        want = ref + bias + res.float()
        assert (y.float() - want).abs().max().item() <= 1e-2 * want.abs().max().item()
        g = ops.gemm(a, b, b_mn=b_mn, aux_in=aux, dact=ops.ACT_RELU, bn_hint=min(bn, 128))
        want = ref * (aux.float() > 0)
        assert (g.float() - want).abs().max().item() <= 1e-2 * want.abs().max().item()
        cs = torch.zeros(N, device="cuda", dtype=torch.float64)
        cq = torch.zeros(N, device="cuda", dtype=torch.float64)
        y = ops.gemm(a, b, b_mn=b_mn, colsum=cs, colsumsq=cq, bn_hint=bn)
        assert torch.allclose(cs, y.double().sum(0), rtol=1e-5, atol=2e-2)
        assert torch.allclose(cq, (y.double() ** 2).sum(0), rtol=1e-5, atol=2e-2)


@pytest.mark.parametrize("M,N,K", [(16384, 512, 256), (40000 + 72, 256, 64), (20000, 1024, 128), (3000, 4096, 1024)])
def test_gemm_dynamic_tile_scheduler_matches_static(mdhs, M, N, K):
    """Several rounds of tiles per CTA: the work-counter schedule (default) must produce exactly what the static round-robin
    produces -- plain, bias + residual through the prefetched operand box (short K: two box slots), GELU' second output,
    act' multiply, fp32 output -- and launches must leave their counter slot clean (the same shapes run repeatedly)."""
    from mdhs_b200 import ops
    torch.manual_seed(11)
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16()

    def run():
        out = {}
        out["plain"] = ops.gemm(a, w)
        out["f32"] = ops.gemm(a, w, out_dtype=torch.float32)
        out["res"] = ops.gemm(a, w, bias=bias, residual=res)
        aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        out["gelu"] = ops.gemm(a, w, bias=bias, act=ops.ACT_GELU_DERIV, aux_out=aux)
        out["aux"] = aux
        out["mul"] = ops.gemm(a, w, aux_in=aux, dact=ops.ACT_MUL)
        out["dgrad"] = ops.gemm(res, w, b_mn=True, M=M, N=K, K=N)          # [M, K] = res . W
        return out
    try:
        ops.set_gemm_dynamic(False)
        want = run()
        ops.set_gemm_dynamic(True)
        for _ in range(3):
            got = run()
            for k in want:
                assert torch.equal(got[k], want[k]), k
    finally:
        ops.set_gemm_dynamic(True)
    ref = a.float() @ w.float().t()
    assert (want["f32"] - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3
