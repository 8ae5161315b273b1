"""Per-kernel parity (GPU): every C-ABI entry point against the same op in plain PyTorch fp32 on
identical (bf16-rounded where the kernel takes bf16) inputs.  Tolerances: bf16 outputs <= 1e-2 of the
tensor's max-norm (one bf16 rounding, 2^-8), fp32 outputs <= 1e-4 relative unless noted."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-12)


@pytest.fixture(scope="module")
def ops(mdhs):
    from mdhs_b200 import ops as o
    return o


@pytest.mark.parametrize("rows,C,xf32", [(100, 256, False), (257, 768, False), (64, 768, True), (33, 1024, False), (16, 2048, False), (50, 136, False)])
def test_layernorm(ops, rows, C, xf32):
    torch.manual_seed(0)
    x = torch.randn(rows, C, device="cuda") * 2 + 0.5
    x = x if xf32 else x.bfloat16()
    g = torch.randn(C, device="cuda")
    b = torch.randn(C, device="cuda")
    y, y32, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12, out_bf16=True, out_f32=True)
    xr = x.float().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (C,), gr, br, 1e-12)
    assert relerr(y32, ref) < 1e-5
    assert relerr(y, ref) < 1e-2
    dy = torch.randn(rows, C, device="cuda").bfloat16()
    ref.backward(dy.float())
    dg = torch.zeros(C, device="cuda")
    db = torch.zeros(C, device="cuda")
    dx, _, dx32 = ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, dx_bf16=True, dx_f32=True)
    assert relerr(dx32, xr.grad) < 1e-4
    assert relerr(dx, xr.grad) < 1e-2
    assert relerr(dg, gr.grad) < 1e-4
    assert relerr(db, br.grad) < 1e-4


@pytest.mark.parametrize("rows,C", [(401, 96), (64, 128), (7, 64), (1, 96), (130, 104)])
def test_layernorm_narrow_rows(ops, rows, C):
    """C <= 128, bf16 in / out: the half-warp-per-row kernels (ConvNeXt stage-1 LayerNorms), forward, dx, dgamma / dbeta
    and the fused dense-bias gradient; odd row counts exercise the unpaired last row."""
    torch.manual_seed(1)
    x = (torch.randn(rows, C, device="cuda") * 2 + 0.5).bfloat16()
    g = torch.randn(C, device="cuda")
    b = torch.randn(C, device="cuda")
    y, _, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6)
    xr = x.float().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (C,), gr, br, 1e-6)
    assert relerr(y, ref) < 1e-2
    assert relerr(mean, xr.detach().mean(1)) < 1e-5
    dy = torch.randn(rows, C, device="cuda").bfloat16()
    ref.backward(dy.float())
    dg, db, dbias = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx, _, _ = ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, dbias=dbias)
    assert relerr(dx, xr.grad) < 1e-2
    assert relerr(dg, gr.grad) < 1e-4
    assert relerr(db, br.grad) < 1e-4
    assert relerr(dbias, dx.float().sum(0)) < 1e-4


def test_layernorm_dropout_consistency(ops):
    """Forward output dropout and the backward mask come from the same stateless hash."""
    torch.manual_seed(1)
    rows, C, p = 64, 768, 0.1
    x = torch.randn(rows, C, device="cuda")
    g = torch.ones(C, device="cuda")
    b = torch.zeros(C, device="cuda")
    y0, _, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12)
    y1, _, _, _ = ops.layernorm_fwd(x, g, b, 1e-12, drop_p=p, seed=1234)
    keep = (y1.float() != 0) | (y0.float() == 0)
    frac = 1.0 - keep.float().mean().item()
    assert abs(frac - p) < 0.02
    assert relerr(y1.float()[keep], (y0.float() / (1 - p))[keep]) < 1e-2
    dy = torch.ones(rows, C, device="cuda").bfloat16()
    dg = torch.zeros(C, device="cuda")
    db = torch.zeros(C, device="cuda")
    ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, drop_p=p, seed=1234)
    # dbeta = sum over rows of the masked dy = keep-count / (1-p)
    assert relerr(db, keep.float().sum(0) / (1 - p)) < 1e-3


@pytest.mark.parametrize("rows,C", [(4 * 56 * 56, 64), (2 * 14 * 14, 1024), (1000, 256)])
def test_batchnorm_train(ops, rows, C):
    torch.manual_seed(0)
    x = (torch.randn(rows, C, device="cuda") * 1.5 + 0.3).bfloat16()
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda")
    res = torch.randn(rows, C, device="cuda").bfloat16()
    rm = torch.zeros(C, device="cuda")
    rv = torch.ones(C, device="cuda")
    cs = torch.zeros(C, device="cuda", dtype=torch.float64)
    cq = torch.zeros(C, device="cuda", dtype=torch.float64)
    ops.col_stats(x, cs, cq)
    mean, invstd, scale, shift = ops.bn_finalize(cs, cq, rows, gamma, beta, rm, rv, 0.1, 1e-5)
    y = ops.bn_apply(x, scale, shift, residual=res, relu=True)
    # reference: NCHW-free formulation on [rows, C]
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_ref, rv_ref = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    bn = F.batch_norm(xr, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5)
    resr = res.float().requires_grad_(True)
    ref = torch.relu(bn + resr)
    assert relerr(y, ref) < 1e-2
    assert relerr(rm, rm_ref) < 1e-4 and relerr(rv, rv_ref) < 1e-4
    dy = torch.randn(rows, C, device="cuda").bfloat16()
    # reference backward with the mask of OUR (bf16-rounded) output to avoid boundary flips
    (bn + resr).backward(dy.float() * (y.float() > 0))
    dg = torch.zeros(C, device="cuda")
    db = torch.zeros(C, device="cuda")
    dx, dz = ops.bn_bwd(dy, x, y, mean, invstd, gamma, dg, db, relu=True, want_dz=True)
    assert relerr(dx, xr.grad) < 1.5e-2
    assert relerr(dz, resr.grad) < 1e-2
    assert relerr(dg, gr.grad) < 2e-3
    assert relerr(db, br.grad) < 2e-3


@pytest.mark.parametrize("rows,C", [(4 * 56 * 56, 64), (3 * 28 * 28, 128), (2 * 14 * 14, 1024), (1001, 256), (1004, 64)])
def test_batchnorm_bwd_mask_recomputed(ops, rows, C):
    """No residual input: the backward gets y=None and rebuilds the ReLU mask from (x, scale, shift); must equal the
    y-based path bit for bit (narrow C exercises the folded 256-wide view of the reductions)."""
    torch.manual_seed(1)
    x = (torch.randn(rows, C, device="cuda") * 1.5 + 0.3).bfloat16()
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda") * 0.3
    cs = torch.zeros(C, device="cuda", dtype=torch.float64)
    cq = torch.zeros(C, device="cuda", dtype=torch.float64)
    ops.col_stats(x, cs, cq)
    assert relerr(cs.float(), x.float().sum(0)) < 1e-5 and relerr(cq.float(), (x.float() ** 2).sum(0)) < 1e-5
    mean, invstd, scale, shift = ops.bn_finalize(cs, cq, rows, gamma, beta, None, None, 0.1, 1e-5)
    y = ops.bn_apply(x, scale, shift, residual=None, relu=True)
    dy = torch.randn(rows, C, device="cuda").bfloat16()
    dg0, db0 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dg1, db1 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx0, _ = ops.bn_bwd(dy, x, y, mean, invstd, gamma, dg0, db0, relu=True)
    dx1, _ = ops.bn_bwd(dy, x, None, mean, invstd, gamma, dg1, db1, relu=True, scale=scale, shift=shift)
    assert torch.equal(dx0, dx1)
    assert relerr(dg1, dg0) < 1e-5 and relerr(db1, db0) < 1e-5
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    bn = F.batch_norm(xr, None, None, gr, br, True, 0.1, 1e-5)
    bn.backward(dy.float() * (y.float() > 0))
    assert relerr(dx1, xr.grad) < 1.5e-2
    assert relerr(dg1, gr.grad) < 2e-3 and relerr(db1, br.grad) < 2e-3


@pytest.mark.parametrize("Cin,Cout,R,stride,pad,H", [(64, 64, 3, 1, 1, 14), (128, 128, 3, 2, 1, 14), (256, 512, 1, 2, 0, 14), (64, 256, 1, 1, 0, 8)])
def test_conv_via_im2col_gemm(ops, Cin, Cout, R, stride, pad, H):
    """conv fprop / dgrad / wgrad = im2col + tcgen05 GEMM + col2im, against F.conv2d autograd."""
    torch.manual_seed(0)
    B = 3
    x = torch.randn(B, Cin, H, H, device="cuda").bfloat16().float()
    w = (torch.randn(Cout, Cin, R, R, device="cuda") / math.sqrt(Cin * R * R)).bfloat16().float()
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, stride=stride, padding=pad)
    x_nhwc = ops.nchw_f32_to_nhwc_bf16(x)
    col, Ho, Wo = ops.im2col_nhwc(x_nhwc, B, H, H, Cin, R, R, stride, pad)
    wp = ops.conv_weight_pack(w)
    y = ops.gemm(col, wp)  # [B*Ho*Wo, Cout]
    y_nchw = ops.nhwc_bf16_to_nchw_f32(y, B, Ho, Wo, Cout)
    assert relerr(y_nchw, ref) < 1e-2
    dy = torch.randn_like(ref).bfloat16().float()
    ref.backward(dy)
    dy_nhwc = ops.nchw_f32_to_nhwc_bf16(dy)
    dcol = ops.gemm(dy_nhwc, wp, b_mn=True)  # [M, R*S*Cin] = dY . Wp
    dx = ops.col2im_nhwc(dcol, B, H, H, Cin, R, R, stride, pad)
    assert relerr(ops.nhwc_bf16_to_nchw_f32(dx, B, H, H, Cin), xr.grad) < 1.5e-2
    gp = torch.zeros(Cout, R * R * Cin, device="cuda")
    ops.gemm(dy_nhwc, col, a_mn=True, b_mn=True, out=gp, accumulate=True, split_k=4)
    gw = torch.zeros_like(w)
    ops.conv_wgrad_unpack(gp, gw)
    assert relerr(gw, wr.grad) < 5e-3


def test_stem_im2col(ops):
    torch.manual_seed(0)
    B, H = 2, 64
    x = torch.randn(B, 3, H, H, device="cuda")
    w = torch.randn(64, 3, 7, 7, device="cuda") * 0.05
    ldk = 152
    col, Ho, Wo = ops.im2col_nchw_f32(x, 7, 7, 2, 3, ldk)
    wp = ops.conv_weight_pack(w, ldk=ldk)
    y = ops.gemm(col, wp)
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), stride=2, padding=3)
    assert relerr(ops.nhwc_bf16_to_nchw_f32(y, B, Ho, Wo, 64), ref) < 1e-2


@pytest.mark.parametrize("R,stride,pad,H,W", [(7, 2, 3, 224, 224), (4, 4, 0, 64, 96), (3, 1, 1, 17, 23)])
def test_stem_im2col_matches_unfold(ops, R, stride, pad, H, W):
    """Patch matrix of an NCHW fp32 image (ResNet stem 7x7/2, ConvNeXt patchify 4x4/4, an odd 3x3) against F.unfold, including
    the zero padding columns up to ldk."""
    torch.manual_seed(0)
    B, C = 3, 3
    x = torch.randn(B, C, H, W, device="cuda")
    K = R * R * C
    ldk = (K + 7) // 8 * 8
    col, Ho, Wo = ops.im2col_nchw_f32(x, R, R, stride, pad, ldk)
    ref = F.unfold(x.bfloat16().float(), R, padding=pad, stride=stride)            # [B, C*R*R, L], rows ordered (c, r, s)
    ref = ref.view(B, C, R * R, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, K)  # -> (r, s, c)
    assert torch.equal(col[:, :K].float(), ref)
    assert not col[:, K:].any()


def test_stem_im2col_tta_variants(ops):
    """TTA variants through the patch-matrix addressing equal the patch matrix of the explicitly augmented batch."""
    torch.manual_seed(0)
    B, H = 2, 32
    x = torch.randn(B, 3, H, H, device="cuda")
    names = ["hflip", "vflip", "rot90"]
    col, Ho, Wo = ops.im2col_nchw_f32(x, 7, 7, 2, 3, 152, tta=ops.tta_codes(names))
    aug = torch.cat([x, x.flip(3), x.flip(2), torch.rot90(x, 1, (2, 3))])
    ref, _, _ = ops.im2col_nchw_f32(aug.contiguous(), 7, 7, 2, 3, 152)
    assert torch.equal(col, ref)


@pytest.mark.parametrize("H", [30, 31, 112])
def test_maxpool(ops, H):
    """even sizes: the 2 x 2-block backward; odd: the per-pixel gather"""
    torch.manual_seed(0)
    B, C = 3, 64
    x = torch.randn(B, C, H, H, device="cuda").bfloat16().float()
    xr = x.clone().requires_grad_(True)
    ref = F.max_pool2d(xr, 3, 2, 1)
    xn = ops.nchw_f32_to_nhwc_bf16(x)
    y, idx, Ho, Wo = ops.maxpool_fwd(xn, B, H, H, C)
    assert relerr(ops.nhwc_bf16_to_nchw_f32(y, B, Ho, Wo, C), ref) == 0
    dy = torch.randn_like(ref).bfloat16().float()
    ref.backward(dy)
    dx = ops.maxpool_bwd(ops.nchw_f32_to_nhwc_bf16(dy), idx, B, H, H, C)
    assert relerr(ops.nhwc_bf16_to_nchw_f32(dx, B, H, H, C), xr.grad) < 1e-2


def test_mean_tokens(ops):
    torch.manual_seed(0)
    B, T, C = 5, 49, 256
    x = torch.randn(B * T, C, device="cuda").bfloat16()
    y, _ = ops.mean_tokens_fwd(x, B, T, C)
    assert relerr(y, x.float().view(B, T, C).mean(1)) < 1e-5
    dy = torch.randn(B, C, device="cuda")
    dx = ops.mean_tokens_bwd(dy, B, T, C)
    assert relerr(dx.view(B, T, C), (dy / T).unsqueeze(1).expand(B, T, C)) < 1e-2


@pytest.mark.parametrize("B,H,W,C,O,R,stride,pad", [(4, 56, 56, 64, 64, 3, 1, 1), (2, 28, 28, 128, 128, 3, 2, 1),
                                                    (2, 14, 14, 256, 512, 1, 2, 0), (2, 7, 7, 512, 512, 3, 1, 1),
                                                    (3, 9, 9, 64, 64, 3, 2, 1), (16, 14, 14, 256, 256, 3, 1, 1)])
def test_implicit_gemm_conv(ops, B, H, W, C, O, R, stride, pad):
    """TMA-im2col implicit GEMM (fprop, wgrad, stride-1 dgrad) vs torch conv2d autograd on bf16-rounded operands.
    Tolerance 1e-2 of the max-norm (bf16 outputs) / 2e-3 (fp32 wgrad accumulators)."""
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, device="cuda").bfloat16()
    w = (torch.randn(O, C, R, R, device="cuda") / math.sqrt(C * R * R)).bfloat16().float()
    x_rows = x.permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    rows, K = B * Ho * Wo, R * R * C
    wp = ops.conv_weight_pack(w)
    y = ops.gemm(x_rows, wp, conv=(1, B, H, W, C, R, R, stride, pad), M=rows, N=O, K=K)
    xr, wr = x.float().requires_grad_(True), w.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, stride=stride, padding=pad)
    assert relerr(y, ref.permute(0, 2, 3, 1).reshape(rows, O)) < 1e-2
    dy = torch.randn(rows, O, device="cuda").bfloat16()
    ref.backward(dy.float().view(B, Ho, Wo, O).permute(0, 3, 1, 2))
    gp = torch.zeros(O, K, device="cuda")
    ops.gemm(dy, x_rows, a_mn=True, b_mn=True, out=gp, accumulate=True, split_k=-1, M=O, N=K, K=rows,
             conv=(2, B, H, W, C, R, R, stride, pad))
    assert relerr(gp, wr.grad.permute(0, 2, 3, 1).reshape(O, K)) < 2e-3
    if stride == 1:
        wt = ops.conv_weight_pack_dgrad(w)
        dx = ops.gemm(dy, wt, conv=(1, B, Ho, Wo, O, R, R, 1, R - 1 - pad), M=B * H * W, N=C, K=R * R * O)
        assert relerr(dx, xr.grad.permute(0, 2, 3, 1).reshape(B * H * W, C)) < 1e-2


def _attn_ref(q, k, v, mask, scale):
    s = torch.einsum("bhqd,bhkd->bhqk", q, k) * scale
    if mask is not None:
        s = s.masked_fill(~mask[:, None, None, :].bool(), float("-inf"))
    p = torch.softmax(s, -1)
    return torch.einsum("bhqk,bhkd->bhqd", p, v)


@pytest.mark.parametrize("B,H,Sq,Sk,D,masked", [(3, 12, 64, 64, 64, True), (2, 8, 49, 49, 32, False), (2, 8, 49, 64, 32, True),
                                                 (2, 8, 196, 40, 32, True), (1, 12, 130, 130, 64, True), (1, 12, 256, 256, 64, False),
                                                 (1, 2, 512, 512, 64, True), (2, 8, 784, 64, 32, True)])
def test_attention(ops, B, H, Sq, Sk, D, masked):
    torch.manual_seed(0)
    q = torch.randn(B * Sq, H * D, device="cuda").bfloat16()
    kv = torch.randn(B * Sk, 2 * H * D, device="cuda").bfloat16()  # K and V as column slices of one buffer
    k, v = kv[:, :H * D], kv[:, H * D:]
    mask = None
    if masked:
        lens = torch.randint(Sk // 2, Sk + 1, (B,), device="cuda")
        mask = (torch.arange(Sk, device="cuda")[None, :] < lens[:, None]).to(torch.uint8)
    scale = 1.0 / math.sqrt(D)
    o, lse = ops.attention_fwd(q, k, v, B, H, Sq, Sk, D, scale, key_mask=mask)

    def heads(t, S):
        return t.float().reshape(B, S, H, D).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    qr, kr, vr = heads(q, Sq), heads(k, Sk), heads(v, Sk)
    ref = _attn_ref(qr, kr, vr, mask, scale)
    ref_tok = ref.permute(0, 2, 1, 3).reshape(B * Sq, H * D)
    assert relerr(o, ref_tok) < 1e-2
    do = torch.randn(B * Sq, H * D, device="cuda").bfloat16()
    ref_tok.backward(do.float())
    dkv = torch.empty_like(kv)
    dq, dk, dv = ops.attention_bwd(q, k, v, o, do, lse, B, H, Sq, Sk, D, scale, key_mask=mask,
                                   dk=dkv[:, :H * D], dv=dkv[:, H * D:])

    def tok(t, S):
        return t.permute(0, 2, 1, 3).reshape(B * S, H * D)
    assert relerr(dq, tok(qr.grad, Sq)) < 2e-2
    assert relerr(dk, tok(kr.grad, Sk)) < 2e-2
    assert relerr(dv, tok(vr.grad, Sk)) < 2e-2
    if Sq > 64:   # long path without the delta scratch: the dK/dV pass recomputes rowsum(dO * O) itself
        dkv2 = torch.empty_like(kv)
        dq2, dk2, dv2 = ops.attention_bwd(q, k, v, o, do, lse, B, H, Sq, Sk, D, scale, key_mask=mask, delta_scratch=False,
                                          dk=dkv2[:, :H * D], dv=dkv2[:, H * D:])
        assert torch.equal(dq2, dq) and relerr(dk2, dk) < 1e-3 and relerr(dv2, dv) < 1e-3


def test_attention_dropout_statistics(ops):
    torch.manual_seed(0)
    B, H, S, D, p = 2, 4, 64, 32, 0.25
    q = torch.randn(B * S, H * D, device="cuda").bfloat16()
    k = torch.randn(B * S, H * D, device="cuda").bfloat16()
    v = torch.ones(B * S, H * D, device="cuda").bfloat16()
    # with V == 1 the output equals sum_j P_ij m_ij, whose expectation is 1
    o, _ = ops.attention_fwd(q, k, v, B, H, S, S, D, 0.1, drop_p=p, seed=7)
    assert abs(o.float().mean().item() - 1.0) < 0.05
    o2, _ = ops.attention_fwd(q, k, v, B, H, S, S, D, 0.1, drop_p=p, seed=7)
    assert torch.equal(o, o2)


@pytest.mark.parametrize("S,D", [(64, 64), (256, 64), (200, 32)])
def test_attention_dropout_backward_matches_autograd(ops, S, D):
    """The dropout mask is a pure function of (seed, element index): recover it from forward passes whose V holds an
    identity block (one pass per D-wide block of keys), then check forward and backward of a random-V problem against
    autograd with that explicit mask.  S > 64 runs the long-sequence kernels (dQ pass + transposed dK/dV pass), whose
    dropout indexing must agree with the forward's."""
    torch.manual_seed(0)
    B, H, p, seed = 2, 3, 0.2, 99
    q = torch.randn(B * S, H * D, device="cuda").bfloat16()
    k = torch.randn(B * S, H * D, device="cuda").bfloat16()
    scale = 1.0 / math.sqrt(D)

    def heads(t):
        return t.float().reshape(B, S, H, D).permute(0, 2, 1, 3).contiguous()
    qh, kh = heads(q).requires_grad_(True), heads(k).requires_grad_(True)
    probs = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    pm = torch.zeros(B, H, S, S, device="cuda")
    for c0 in range(0, S, D):
        n = min(D, S - c0)
        blk = torch.zeros(S, D, device="cuda")
        blk[c0:c0 + n, :n] = torch.eye(n, device="cuda")
        vsel = blk.repeat(B, H).bfloat16()                               # [B*S, H*D] with V_bh = the identity block
        o_sel, _ = ops.attention_fwd(q, k, vsel, B, H, S, S, D, scale, drop_p=p, seed=seed)
        pm[..., c0:c0 + n] = heads(o_sel)[..., :n]                        # = probs * mask (bf16-rounded)
    mask = torch.where(pm > 0.5 * probs.detach() / (1 - p), torch.full_like(pm, 1.0 / (1 - p)), torch.zeros_like(pm))
    assert abs((mask == 0).float().mean().item() - p) < 0.02
    v = torch.randn(B * S, H * D, device="cuda").bfloat16()
    vh = heads(v).requires_grad_(True)
    ref = ((probs * mask) @ vh).permute(0, 2, 1, 3).reshape(B * S, H * D)
    o, lse = ops.attention_fwd(q, k, v, B, H, S, S, D, scale, drop_p=p, seed=seed)
    assert relerr(o, ref) < 1e-2
    do = torch.randn(B * S, H * D, device="cuda").bfloat16()
    ref.backward(do.float())
    dq, dk, dv = ops.attention_bwd(q, k, v, o, do, lse, B, H, S, S, D, scale, drop_p=p, seed=seed)

    def tok(t):
        return t.permute(0, 2, 1, 3).reshape(B * S, H * D)
    assert relerr(dq, tok(qh.grad)) < 2e-2
    assert relerr(dk, tok(kh.grad)) < 2e-2
    assert relerr(dv, tok(vh.grad)) < 2e-2


def test_embedding(ops):
    torch.manual_seed(0)
    B, S, C, V = 4, 16, 768, 1000
    word = torch.randn(V, C, device="cuda")
    pos = torch.randn(512, C, device="cuda")
    typ = torch.randn(2, C, device="cuda")
    ids = torch.randint(0, V, (B, S), device="cuda")
    e = ops.embed_gather(ids, None, word, pos, typ, S)
    ref = word[ids] + pos[:S][None] + typ[0]
    assert relerr(e, ref.reshape(B * S, C)) < 1e-6
    de = torch.randn(B * S, C, device="cuda")
    gw, gp, gt = torch.zeros_like(word), torch.zeros_like(pos), torch.zeros_like(typ)
    ops.embed_scatter(de, ids, None, gw, gp, gt, S)
    gw_ref = torch.zeros_like(word).index_add_(0, ids.reshape(-1), de)
    assert relerr(gw, gw_ref) < 1e-5
    assert relerr(gp[:S], de.view(B, S, C).sum(0)) < 1e-5
    assert relerr(gt[0], de.sum(0)) < 1e-5


def test_linear_f32_and_ce(ops):
    torch.manual_seed(0)
    M, K, N = 37, 256, 7
    x = torch.randn(M, K, device="cuda", requires_grad=True)
    w = (torch.randn(N, K, device="cuda") * 0.1).requires_grad_(True)
    b = torch.randn(N, device="cuda", requires_grad=True)
    labels = torch.randint(0, N, (M,), device="cuda")
    cw = torch.rand(N, device="cuda") + 0.5
    for kwargs, ref_fn in [
        (dict(label_smoothing=0.02), lambda z: F.cross_entropy(z, labels, label_smoothing=0.02)),
        (dict(label_smoothing=0.1, class_weights=cw), lambda z: F.cross_entropy(z, labels, weight=cw, label_smoothing=0.1)),
        (dict(), lambda z: F.cross_entropy(z, labels)),
    ]:
        for t in (x, w, b):
            t.grad = None
        z_ref = F.linear(x, w, b)
        loss_ref = ref_fn(z_ref)
        loss_ref.backward()
        z = ops.linear_f32_fwd(x.detach(), w.detach(), b.detach())
        assert relerr(z, z_ref) < 1e-5
        loss, dl = ops.ce_loss(z, labels, **kwargs)
        assert abs(loss.item() - loss_ref.item()) < 1e-5 * max(1.0, abs(loss_ref.item()))
        dw, db = torch.zeros_like(w), torch.zeros_like(b)
        dx = ops.linear_f32_bwd(dl, x.detach(), w.detach(), dw, db)
        assert relerr(dx, x.grad) < 1e-4
        assert relerr(dw, w.grad) < 1e-4
        assert relerr(db, b.grad) < 1e-4


def test_focal_loss(ops):
    torch.manual_seed(0)
    B, C, gamma = 64, 7, 2.0
    z = torch.randn(B, C, device="cuda", requires_grad=True)
    y = torch.randint(0, C, (B,), device="cuda")
    ce = F.cross_entropy(z, y, reduction="none")
    ref = ((1 - torch.exp(-ce)) ** gamma * ce).mean()
    ref.backward()
    loss, dl = ops.ce_loss(z.detach(), y, focal=True, gamma=gamma)
    assert abs(loss.item() - ref.item()) < 1e-5
    assert relerr(dl, z.grad) < 1e-4


def test_adamw_and_sgd_flat(ops):
    torch.manual_seed(0)
    n = 4096 + 8
    p = torch.randn(n, device="cuda")
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref], lr=2e-4, weight_decay=1e-2)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    shadow = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for step in range(1, 4):
        g = torch.randn(n, device="cuda")
        ref.grad = g.clone()
        opt.step()
        gg = g.clone()
        ops.adam_flat(p, gg, m, v, shadow, 2e-4, 0.9, 0.999, 1e-8, 1e-2, step)
        assert gg.abs().max().item() == 0
    assert relerr(p, ref.detach()) < 1e-6
    assert relerr(shadow, p) < 1e-2
    p2 = torch.randn(n, device="cuda")
    ref2 = p2.clone().requires_grad_(True)
    opt2 = torch.optim.SGD([ref2], lr=0.01, momentum=0.9)
    mom = torch.zeros(n, device="cuda")
    for step in range(3):
        g = torch.randn(n, device="cuda")
        ref2.grad = g.clone()
        opt2.step()
        ops.sgd_flat(p2, g.clone(), mom, None, 0.01, 0.9, 0.0, first_step=(step == 0))
    assert relerr(p2, ref2.detach()) < 1e-6


def test_mp_loss_kernel(ops):
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import port
    torch.manual_seed(0)
    B, C = 33, 6
    zs = [(torch.randn(B, C, device="cuda") * s).requires_grad_(True) for s in (1.0, 2.0, 0.5)]
    y = torch.randint(0, C, (B,), device="cuda")
    cpu = [z.detach().cpu().requires_grad_(True) for z in zs]
    ref = port.mp_loss(cpu[0], cpu[1], cpu[2], y.cpu())
    ref.backward()
    loss, g = ops.mp_loss(zs[0].detach(), zs[1].detach(), zs[2].detach(), y)
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    for a, b in zip(g, cpu):
        assert relerr(a.cpu(), b.grad) < 1e-3


@pytest.mark.parametrize("B,D", [(32, 256), (128, 256), (7, 40)])
def test_supcon_loss_fwd_bwd(ops, B, D):
    """SupConLoss (scripts/train.py:23-44) vs the oracle restatement (pinned to the reference): loss 1e-4, gradient 1e-3."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import port
    torch.manual_seed(3)
    x = torch.randn(B, D)
    labels = torch.randint(0, 4, (B,))
    xr = x.clone().requires_grad_(True)
    want = port.supcon_loss(xr, labels, 0.07)
    want.backward()
    loss, dx = ops.supcon_loss(x.cuda(), labels.cuda(), 0.07)
    assert abs(loss.item() - want.item()) <= 1e-4 * max(1.0, abs(want.item()))
    assert relerr(dx, xr.grad.cuda()) < 1e-3


@pytest.mark.parametrize("rows,C,res", [(4 * 56 * 56, 64, False), (2 * 14 * 14, 1024, True), (1001, 256, True), (77, 2048, False)])
def test_batchnorm_fwd_fused_equals_finalize_plus_apply(ops, rows, C, res):
    """mdhs_bn_fwd (finalize + apply in one launch) is bit-identical to mdhs_bn_finalize + mdhs_bn_apply, in train mode
    (incl. the running-statistics update) and in eval mode."""
    torch.manual_seed(2)
    x = (torch.randn(rows, C, device="cuda") * 1.5 + 0.3).bfloat16()
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda")
    r = torch.randn(rows, C, device="cuda").bfloat16() if res else None
    cs = torch.zeros(C, device="cuda", dtype=torch.float64)
    cq = torch.zeros(C, device="cuda", dtype=torch.float64)
    ops.col_stats(x, cs, cq)
    for training in (True, False):
        rm0, rv0 = torch.randn(C, device="cuda") * 0.1, torch.rand(C, device="cuda") + 0.5
        rm1, rv1 = rm0.clone(), rv0.clone()
        mean, invstd, scale, shift = ops.bn_finalize(cs if training else None, cq if training else None, rows, gamma, beta,
                                                     rm0, rv0, 0.1, 1e-5, training=training)
        y0 = ops.bn_apply(x, scale, shift, residual=r, relu=True)
        y1, m1, i1, s1, h1 = ops.bn_fwd(x, cs if training else None, cq if training else None, gamma, beta, rm1, rv1, 0.1, 1e-5,
                                        residual=r, relu=True, training=training)
        assert torch.equal(y0, y1)
        assert torch.equal(mean, m1) and torch.equal(invstd, i1) and torch.equal(scale, s1) and torch.equal(shift, h1)
        assert torch.equal(rm0, rm1) and torch.equal(rv0, rv1)


def test_batchnorm_bwd_eval_mode(ops):
    """training=0: dx = gamma * rsqrt(running_var + eps) * dy' (what F.batch_norm(training=False) back-propagates);
    dgamma / dbeta still are sum(dy' * xhat) / sum(dy')."""
    torch.manual_seed(3)
    rows, C = 1000, 256
    x = (torch.randn(rows, C, device="cuda") * 1.5 + 0.3).bfloat16()
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda") * 0.3
    rm, rv = torch.randn(C, device="cuda") * 0.2, torch.rand(C, device="cuda") + 0.5
    y, mean, invstd, scale, shift = ops.bn_fwd(x, None, None, gamma, beta, rm, rv, 0.0, 1e-5, relu=True, training=False)
    dy = torch.randn(rows, C, device="cuda").bfloat16()
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx, _ = ops.bn_bwd(dy, x, None, mean, invstd, gamma, dg, db, relu=True, scale=scale, shift=shift, training=False)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    out = F.batch_norm(xr, rm, rv, gr, br, False, 0.0, 1e-5)
    out.backward(dy.float() * (y.float() > 0))
    assert relerr(dx, xr.grad) < 1e-2
    assert relerr(dg, gr.grad) < 2e-3 and relerr(db, br.grad) < 2e-3


@pytest.mark.parametrize("M,N,K,conv", [(4 * 28 * 28, 128, 512, None), (1000, 64, 256, None), (2 * 14 * 14, 256, 9 * 256, (2, 14, 14)),
                                        (8 * 56 * 56, 64, 9 * 64, (8, 56, 56))])
def test_gemm_epilogue_bn_backward_reduction(ops, M, N, K, conv):
    """mdhs_gemm_args.stat_x: the dgrad GEMM that writes dy also accumulates sum(dy') and sum(dy' * (x - mean)) of the
    producer BatchNorm (mask recomputed from the raw activation); feeding those sums to mdhs_bn_bwd(sums_ready) equals the
    separate reduce pass."""
    torch.manual_seed(4)
    if conv is None:
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        kw = {}
    else:
        Bc, H, W = conv
        Co = K // 9
        a = (torch.randn(M, Co, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        kw = dict(conv=(1, Bc, H, W, Co, 3, 3, 1, 1), M=M, N=N, K=K)
    x = (torch.randn(M, N, device="cuda") * 1.5 + 0.3).bfloat16()      # raw activation of the producer layer
    gamma = torch.rand(N, device="cuda") + 0.5
    beta = torch.randn(N, device="cuda") * 0.3
    cs = torch.zeros(N, device="cuda", dtype=torch.float64)
    cq = torch.zeros(N, device="cuda", dtype=torch.float64)
    ops.col_stats(x, cs, cq)
    y, mean, invstd, scale, shift = ops.bn_fwd(x, cs, cq, gamma, beta, None, None, 0.1, 1e-5, relu=True, training=True)
    dy_plain = ops.gemm(a, w, **kw)
    sums = torch.zeros(2, N, device="cuda", dtype=torch.float64)
    dy = ops.gemm(a, w, stat_x=x, stat_mean=mean, stat_scale=scale, stat_shift=shift, stat_relu=True, colsum=sums[0],
                  colsumsq=sums[1], **kw)
    assert torch.equal(dy, dy_plain)
    mask = (torch.addcmul(shift, x.float(), scale) > 0)      # fmaf(x, scale, shift): same rounding as the kernels up to ties
    dm = dy.float() * (y.float() > 0)
    want0 = dm.double().sum(0)
    want1 = (dm.double() * (x.double() - mean.double())).sum(0)
    assert relerr(sums[0].float(), want0.float()) < 2e-4, relerr(sums[0].float(), want0.float())
    assert relerr(sums[1].float(), want1.float()) < 2e-4
    del mask
    dg0, db0 = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
    dg1, db1 = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
    dx0, _ = ops.bn_bwd(dy, x, None, mean, invstd, gamma, dg0, db0, relu=True, scale=scale, shift=shift)
    dx1, _ = ops.bn_bwd(dy, x, None, mean, invstd, gamma, dg1, db1, relu=True, scale=scale, shift=shift, sums=sums)
    assert relerr(dx1, dx0) < 4e-3
    assert relerr(dg1, dg0) < 1e-4 and relerr(db1, db0) < 1e-4


def test_preprocess_u8_matches_torch(ops):
    """GPU input pipeline: crop box + bilinear resize + flips + ToTensor + Normalize of a uint8 HWC batch vs the
    same chain in torch (crop -> F.interpolate(bilinear, align_corners=False) -> flip -> normalise)."""
    torch.manual_seed(0)
    B, Hs, Ws = 5, 300, 260
    src = torch.randint(0, 256, (B, Hs, Ws, 3), dtype=torch.uint8, device="cuda")
    boxes = torch.tensor([[0, 0, 300, 260], [10, 20, 200, 180], [38, 18, 224, 224], [100, 60, 90, 150], [5, 7, 64, 64]], dtype=torch.float32)
    flips = torch.tensor([0, 1, 2, 3, 0], dtype=torch.uint8)
    got = ops.preprocess_u8(src, (224, 224), boxes=boxes, flips=flips)
    mean = torch.tensor(ops.IMAGENET_MEAN, device="cuda").view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD, device="cuda").view(1, 3, 1, 1)
    for b in range(B):
        y0, x0, h, w = [int(v) for v in boxes[b]]
        crop = src[b, y0:y0 + h, x0:x0 + w].permute(2, 0, 1).float().unsqueeze(0)
        ref = F.interpolate(crop, size=(224, 224), mode="bilinear", align_corners=False)
        if int(flips[b]) & 1:
            ref = ref.flip(-1)
        if int(flips[b]) & 2:
            ref = ref.flip(-2)
        ref = (ref / 255.0 - mean) / std
        assert (got[b:b + 1] - ref).abs().max().item() < 2e-4, b
    # whole image, no boxes / flips: identity geometry when the sizes match
    same = ops.preprocess_u8(src[:, :224, :224].contiguous(), (224, 224))
    ref = (src[:, :224, :224].permute(0, 3, 1, 2).float() / 255.0 - mean) / std
    assert (same - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("rows,C", [(4 * 56 * 56, 256), (1001, 64), (392, 2048)])
def test_batchnorm_relu_bitmask_equals_y_mask(ops, rows, C):
    """Residual layers: the forward's 1-bit/element ReLU mask (uint8 [rows, C/8]) gives the backward exactly what re-reading
    the bf16 output gave (bit-identical dx / dz, same reductions)."""
    torch.manual_seed(5)
    x = (torch.randn(rows, C, device="cuda") * 1.5 + 0.3).bfloat16()
    res = torch.randn(rows, C, device="cuda").bfloat16()
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda") * 0.3
    cs = torch.zeros(C, device="cuda", dtype=torch.float64)
    cq = torch.zeros(C, device="cuda", dtype=torch.float64)
    ops.col_stats(x, cs, cq)
    y, mean, invstd, scale, shift, mask = ops.bn_fwd(x, cs, cq, gamma, beta, None, None, 0.1, 1e-5, residual=res, relu=True,
                                                     want_mask=True)
    assert mask.shape == (rows, C // 8) and mask.dtype == torch.uint8
    bits = ((mask.unsqueeze(-1) >> torch.arange(8, device="cuda", dtype=torch.uint8)) & 1).reshape(rows, C).bool()
    assert torch.equal(bits, y > 0)
    dy = torch.randn(rows, C, device="cuda").bfloat16()
    dg0, db0 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dg1, db1 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dx0, dz0 = ops.bn_bwd(dy, x, y, mean, invstd, gamma, dg0, db0, relu=True, want_dz=True)
    dx1, dz1 = ops.bn_bwd(dy, x, None, mean, invstd, gamma, dg1, db1, relu=True, want_dz=True, mask=mask)
    assert torch.equal(dz0, dz1)
    assert relerr(dx1, dx0) < 1e-3 and relerr(dg1, dg0) < 1e-5 and relerr(db1, db0) < 1e-5


def test_layernorm_bwd_fused_dense_bias_gradient(ops):
    """mdhs_layernorm_bwd(dbias=...): the bias gradient of the dense layer behind the LayerNorm = column sums of the
    (dropout-masked) input gradient the same call writes -- equal to the separate col_stats pass it replaces."""
    torch.manual_seed(6)
    rows, C = 1000, 768
    x = torch.randn(rows, C, device="cuda").bfloat16()
    g = torch.rand(C, device="cuda") + 0.5
    b = torch.randn(C, device="cuda")
    _, _, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12)
    dy = torch.randn(rows, C, device="cuda").bfloat16()
    for p2 in (0.0, 0.1):
        dg, db, dbias = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        dx, dxd, _ = ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, drop2_p=p2, seed2=77, want_dx_drop=p2 > 0, dbias=dbias)
        src = dxd if p2 > 0 else dx
        want = torch.zeros(C, device="cuda")
        ops.col_stats(src, sum32=want)
        assert relerr(dbias, want) < 1e-5
        assert relerr(dbias, src.float().sum(0)) < 1e-4
