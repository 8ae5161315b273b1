"""ConvNeXt feature extractor (torchvision `ConvNeXt.features`: tiny / small / base / large) on the B200 kernels.

Reference call sites: ConNexT/models/ourmodel.py:57-62,76 (`models.convnext_base(...).features`),
ConNexT/models/pl_model_MOE2.py:36-53 (`ConvNeXtEncoder`: features -> flatten(2)).  The torchvision Sequential is only
the parameter container (state_dict keys `0.0.weight` ... `7.N.block.5.bias`, `N.M.layer_scale`); the arithmetic is

  stem / down-sampling : non-overlapping k x k stride-k convolution = patch gather + tcgen05 GEMM (+ bias), LayerNorm
  CNBlock              : depthwise 7x7 (csrc/convnext.cu) -> LayerNorm(1e-6) -> GEMM + bias + GELU -> GEMM + bias
                         -> x + layer_scale * stochastic_depth(row) * z  (one fused kernel)

on NHWC bf16 token matrices [B*H*W, C].  Backward runs through torch.autograd Functions whose bodies are our kernels;
parameter gradients are written straight into the ParamStore's flat fp32 gradient buffer.
"""
import torch

from .. import functional as Fm
from .. import ops


class _PatchConv:
    """Conv2d(k x k, stride k, pad 0, bias) lowered to patch-gather + GEMM.  Keeps the packed bf16 weight [O, (r,s,c)]."""

    def __init__(self, store, conv, stem):
        self.store, self.conv, self.stem = store, conv, stem
        self.O, self.I, self.R, self.S = conv.weight.shape
        assert conv.stride[0] == self.R and conv.padding[0] == 0 and conv.groups == 1
        self.K = self.R * self.S * self.I
        self.ldk = (self.K + 7) // 8 * 8
        dev = store.device
        self.wp = torch.zeros((self.O, self.ldk), device=dev, dtype=torch.bfloat16)
        self.gp = torch.zeros((self.O, self.ldk), device=dev, dtype=torch.float32)
        store.add_packer(self.repack)

    def repack(self):
        ops.conv_weight_pack(self.conv.weight.data, ldk=self.ldk, out=self.wp)


class _PatchConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, rec, B, H, W):
        if rec.stem:   # x: NCHW fp32 image
            col, Ho, Wo = ops.im2col_nchw_f32(x.contiguous(), rec.R, rec.S, rec.R, 0, rec.ldk)
        else:          # x: NHWC bf16 [B*H*W, I]
            col, Ho, Wo = ops.im2col_nhwc(x.contiguous(), B, H, W, rec.I, rec.R, rec.S, rec.R, 0)
        bias = rec.conv.bias.data if rec.conv.bias is not None else None
        y = ops.gemm(col, rec.wp, bias=bias, N=rec.O)
        ctx.rec, ctx.geom = rec, (B, H, W)
        ctx.save_for_backward(col)
        return y

    @staticmethod
    def backward(ctx, dy):
        (col,) = ctx.saved_tensors
        rec = ctx.rec
        B, H, W = ctx.geom
        st = rec.store
        dy = dy.contiguous()
        rows = dy.shape[0]
        if rec.conv.weight.requires_grad:
            rec.gp.zero_()
            ops.gemm(dy, col, a_mn=True, b_mn=True, out=rec.gp, accumulate=True, split_k=-1, M=rec.O, N=rec.ldk, K=rows)
            ops.conv_wgrad_unpack(rec.gp, st.g32(rec.conv.weight))
            if rec.conv.bias is not None:
                ops.col_stats(dy, sum32=st.g32(rec.conv.bias))
        dx = None
        if not rec.stem and ctx.needs_input_grad[0]:
            dcol = ops.gemm(dy, rec.wp[:, :rec.K], b_mn=True, M=rows, N=rec.K, K=rec.O)
            dx = ops.col2im_nhwc(dcol, B, H, W, rec.I, rec.R, rec.S, rec.R, 0)
        return dx, None, None, None, None, None


class _DwConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, gw, gb, B, H, W):
        x = x.contiguous()
        y = ops.dwconv7(x, w, bias, B, H, W)
        ctx.w, ctx.gw, ctx.gb, ctx.geom = w, gw, gb, (B, H, W)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        B, H, W = ctx.geom
        dy = dy.contiguous()
        if ctx.gw is not None:
            ops.dwconv7_wgrad(x, dy, ctx.gw, ctx.gb, B, H, W)
        dx = ops.dwconv7(dy, ctx.w, None, B, H, W, flip=True) if ctx.needs_input_grad[0] else None
        return dx, None, None, None, None, None, None, None


class _MlpScaleFn(torch.autograd.Function):
    """CNBlock after its LayerNorm (torchvision CNBlock.block[3:], layer_scale, StochasticDepth("row"), residual):
        out = x + ls * keep(sample) * (GELU(y W1^T + b1) W2^T + b2)
    Forward: the first GEMM's epilogue also stores GELU'(pre-activation); backward: the layer-scale kernel emits dz, dls and
    the second Linear's bias gradient, the dgrad GEMM of the second Linear multiplies by the saved GELU' and column-sums its
    output (= the first Linear's bias gradient), so no separate activation-backward or bias-reduction pass runs."""

    @staticmethod
    def forward(ctx, x, y, p1, p2, lsp, scratch, rows_per_sample, p, seed, need):
        x, y = x.contiguous(), y.contiguous()
        w1, gw1, b1, gb1 = p1
        w2, gw2, b2, gb2 = p2
        ls, gls = lsp
        T, N1 = y.shape[0], w1.shape[0]
        deriv = torch.empty((T, N1), device=y.device, dtype=torch.bfloat16) if need else None
        h = ops.gemm(y, w1, bias=b1, act=ops.ACT_GELU_DERIV if need else ops.ACT_GELU, aux_out=deriv)
        z = ops.gemm(h, w2, bias=b2)
        out = ops.layer_scale_fwd(x, z, ls, rows_per_sample, p, seed)
        ctx.p1, ctx.p2, ctx.lsp, ctx.scratch, ctx.cfg = p1, p2, lsp, scratch, (rows_per_sample, p, seed)
        ctx.save_for_backward(y, h, deriv, z)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, h, deriv, z = ctx.saved_tensors
        w1, gw1, b1, gb1 = ctx.p1
        w2, gw2, b2, gb2 = ctx.p2
        ls, gls = ctx.lsp
        rps, p, seed = ctx.cfg
        dout = dout.contiguous()
        T, C = dout.shape
        N1 = w1.shape[0]
        dz = ops.layer_scale_bwd(dout, z, ls, gls, rps, p, seed, dbias=gb2)
        if gw2 is not None:
            ops.gemm(dz, h, a_mn=True, b_mn=True, out=gw2, accumulate=True, split_k=-1, M=C, N=N1, K=T)
        s64 = ctx.scratch[:, :N1] if gb1 is not None else None
        g = ops.gemm(dz, w2, b_mn=True, aux_in=deriv, dact=ops.ACT_MUL, M=T, N=N1, K=C,
                     colsum=None if s64 is None else s64[0], colsumsq=None if s64 is None else s64[1])
        if gb1 is not None:
            ops.sum64_to_grad(s64[0], s64[1], gb1)
        if gw1 is not None:
            ops.gemm(g, y, a_mn=True, b_mn=True, out=gw1, accumulate=True, split_k=-1, M=N1, N=C, K=T)
        dy = ops.gemm(g, w1, b_mn=True, M=T, N=C, K=N1) if ctx.needs_input_grad[1] else None
        return dout, dy, None, None, None, None, None, None, None, None


class SqAttnFn(torch.autograd.Function):
    """Single-query attention: q [B, D] bf16, k / v [B*T, D] bf16 -> out [B, D] fp32 (softmax over the T positions)."""

    @staticmethod
    def forward(ctx, q, k, v, B, T, scale):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        out, probs = ops.sq_attn_fwd(q, k, v, B, T, scale)
        ctx.cfg = (B, T, scale)
        ctx.save_for_backward(q, k, v, probs)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, probs = ctx.saved_tensors
        B, T, scale = ctx.cfg
        dq, dk, dv = ops.sq_attn_bwd(q, k, v, dout.contiguous().float(), probs, B, T, scale)
        return dq, dk, dv, None, None, None


class ConvNeXtEngine:
    """Runs `features` (a torchvision ConvNeXt feature Sequential) whose parameters live in `store`."""

    def __init__(self, store, features):
        self.store, self.features = store, features
        self.step_seed = 0x5D000000
        self._s64 = None
        self.stages = []   # ("stem" | "down" | "blocks", payload)
        for i, stage in enumerate(features):
            first = stage[0]
            if i == 0:
                self.stages.append(("stem", (_PatchConv(store, stage[0], stem=True), stage[1])))
            elif isinstance(first, torch.nn.Conv2d) or type(first).__name__ == "LayerNorm2d":
                self.stages.append(("down", (stage[0], _PatchConv(store, stage[1], stem=False))))
            else:
                self.stages.append(("blocks", list(stage)))

    @property
    def out_channels(self):
        kind, payload = self.stages[-1]
        return payload[-1].block[0].weight.shape[0] if kind == "blocks" else payload[1].O

    def _block(self, blk, x, B, H, W, training, seed):
        st = self.store
        dw, ln, fc1, fc2 = blk.block[0], blk.block[2], blk.block[3], blk.block[5]
        tr = dw.weight.requires_grad
        y = _DwConvFn.apply(x, dw.weight.data, dw.bias.data if dw.bias is not None else None,
                            st.g32(dw.weight) if tr else None, st.g32(dw.bias) if (tr and dw.bias is not None) else None,
                            B, H, W)
        y = Fm.layernorm(y, st, ln)
        ls = blk.layer_scale
        p = float(blk.stochastic_depth.p) if training else 0.0
        return _MlpScaleFn.apply(x, y, self._lin(fc1), self._lin(fc2),
                                 (ls.data.view(-1), st.g32(ls).view(-1) if ls.requires_grad else None),
                                 self._scratch64(fc1.weight.shape[0]), H * W, p, seed, torch.is_grad_enabled())

    def _lin(self, lin):
        """(bf16 shadow weight, fp32 weight-gradient view, fp32 bias, fp32 bias-gradient view) of an nn.Linear."""
        st = self.store
        tr = lin.weight.requires_grad
        has_b = lin.bias is not None
        return (st.w16(lin.weight), st.g32(lin.weight) if tr else None, lin.bias.data if has_b else None,
                st.g32(lin.bias) if (has_b and lin.bias.requires_grad) else None)

    def _scratch64(self, n):
        """fp64 [2, n] column-sum workspace of the bias-gradient epilogue; zero between uses (mdhs_sum64_to_grad clears it)."""
        if self._s64 is None or self._s64.shape[1] < n:
            self._s64 = torch.zeros((2, max(n, 4096)), device=self.store.device, dtype=torch.float64)
        return self._s64

    def forward(self, images, training, taps=None):
        """images: [B,3,H,W] fp32 CUDA -> (tokens [B*h*w, C] bf16, h, w).  taps: optional list that receives
        (tokens, h, w, C) after every residual stage (torchvision features[1], [3], [5], [7])."""
        st = self.store
        B, _, H, W = images.shape
        if training:
            self.step_seed += 1000
        x = None
        n = 0
        for kind, payload in self.stages:
            if kind == "stem":
                conv, ln = payload
                x = _PatchConvFn.apply(images, st.anchor, conv, B, H, W)
                H, W = H // conv.R, W // conv.S
                x = Fm.layernorm(x, st, ln)
            elif kind == "down":
                ln, conv = payload
                x = Fm.layernorm(x, st, ln)
                x = _PatchConvFn.apply(x, st.anchor, conv, B, H, W)
                H, W = H // conv.R, W // conv.S
            else:
                for blk in payload:
                    n += 1
                    x = self._block(blk, x, B, H, W, training, self.step_seed + n)
                if taps is not None:
                    taps.append((x, H, W, x.shape[1]))
        return x, H, W
