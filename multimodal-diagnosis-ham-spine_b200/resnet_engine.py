"""ResNet trunk (torchvision topology: BasicBlock for resnet18/34, Bottleneck for resnet50) on the B200
kernels: NHWC bf16 activations, every convolution an (implicit) tcgen05 GEMM whose epilogue also produces the
train-mode BatchNorm statistics, BN finalize + apply + residual + ReLU fused in one pass, and a hand-scheduled
backward (BN reduce -- fused into the epilogue of the dgrad GEMM that produces dy where possible --, BN apply with
in-kernel coefficients, dgrad GEMM, split-K wgrad GEMM accumulating in fp32).

Reference path: encoder.py:61-72,88-99 (ImageEncoder stem/layer1-4) and mibf_net/model_resnet.py:15-16,40
(torchvision resnet50 incl. avg-pool + fc).  The nn.Module objects only hold parameters / buffers with
the reference's state_dict keys; all arithmetic happens here.

The trunk can be driven as ONE unit (`forward` / `backward`, the fast path) or stage by stage (`forward_stem`,
`forward_layer`, `backward_layer`, `backward_stem`) -- the latter is what encoder.py uses when analysis hooks are
registered on `image_encoder.stem | layerN[-1]` (analysis_tools.py:29-31), so that activations / gradients can be
exposed at the stage boundaries.
"""
import contextlib
import os as _os

import torch

from . import ops
from . import runtime


# Train-mode BN statistics: fused into the conv GEMM epilogue (column sums of the staged bf16 output box, accumulated in
# registers across the persistent CTA's tiles) or, when False, one extra column pass (col_stats) over the raw conv output.
FUSE_BN_STATS_IN_GEMM = _os.environ.get("MDHS_FUSE_BN_STATS", "1") != "0"
# convolutions with a shorter reduction than this take the separate col_stats pass (their GEMM is epilogue / HBM paced)
# (round 2, tools/bench_gemm_step.py: even the most output-bound shape, 401408 x 256 x 64, pays 32 us for the fused column pass
#  against 60 us for the separate one, so everything is fused by default now)
FUSE_BN_STATS_MIN_K = int(_os.environ.get("MDHS_FUSE_BN_MIN_K", "0"))
# backward: sum(dy'), sum(dy' (x - mean)) of the PRODUCER layer's BatchNorm taken in the epilogue of the dgrad GEMM that
# writes dy (the raw activation arrives through the prefetched operand box), so the separate reduce pass disappears
FUSE_BN_BWD_REDUCE = _os.environ.get("MDHS_FUSE_BN_BWD", "1") != "0"
# 1x1 dgrad GEMMs carry the fused reduction only when their reduction (= the conv's output channels) is at least this long
FUSE_BN_BWD_MIN_K = int(_os.environ.get("MDHS_FUSE_BN_BWD_MIN_K", "1024"))


_nullctx = contextlib.nullcontext


def _out_hw(h, k, s, p):
    return (h + 2 * p - k) // s + 1


class _Conv:
    """One Conv2d(bias=False) + its BatchNorm2d."""

    def __init__(self, store, conv, bn, stem=False):
        self.store, self.conv, self.bn, self.stem = store, conv, bn, stem
        self.O, self.I, self.R, self.S = conv.weight.shape
        self.stride, self.pad = conv.stride[0], conv.padding[0]
        self.K = self.R * self.S * self.I
        self.ldk = (self.K + 7) // 8 * 8
        self.direct = (self.R == 1 and self.stride == 1)  # NHWC activation matrix is already the GEMM operand
        self.plain = (self.R == 1)                        # packed layout == OIHW layout
        # implicit GEMM: the activation is read through a TMA im2col descriptor (no patch matrix in HBM)
        self.implicit = (not stem) and (not self.direct) and self.I % 64 == 0
        self.implicit_dgrad = self.implicit and self.stride == 1 and self.O % 64 == 0
        dev = store.device
        self.wt = None
        if self.plain:
            self.wp = store.w16(conv.weight).view(self.O, self.I)
            self.gp = None
        else:
            self.wp = torch.zeros((self.O, self.ldk), device=dev, dtype=torch.bfloat16)
            self.gp = torch.zeros((self.O, self.ldk), device=dev, dtype=torch.float32)
            self.wt = (torch.zeros((self.I, self.R * self.S * self.O), device=dev, dtype=torch.bfloat16)
                       if self.implicit_dgrad else None)
            store.add_packer(self.repack)

    def repack(self):
        ops.conv_weight_pack(self.conv.weight.data, ldk=self.ldk, out=self.wp)
        if self.wt is not None:
            ops.conv_weight_pack_dgrad(self.conv.weight.data, out=self.wt)

    def conv_desc(self, mode, B, H, W):
        return (mode, B, H, W, self.I, self.R, self.S, self.stride, self.pad)

    @property
    def dgrad_fusable(self):
        """The dgrad of this convolution is ONE GEMM writing dx directly (no col2im), so its epilogue can carry the
        producer BatchNorm's backward reduction."""
        return self.direct or self.implicit_dgrad


class ResNetEngine:
    def __init__(self, store, net):
        """net: a torchvision.models.ResNet instance used as the parameter container."""
        self.store = store
        self.net = net
        self.stem = _Conv(store, net.conv1, net.bn1, stem=True)
        self.layers = []
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
            blocks = []
            for blk in layer:
                convs = [_Conv(store, blk.conv1, blk.bn1), _Conv(store, blk.conv2, blk.bn2)]
                if hasattr(blk, "conv3"):
                    convs.append(_Conv(store, blk.conv3, blk.bn3))
                down = None
                if blk.downsample is not None:
                    down = _Conv(store, blk.downsample[0], blk.downsample[1])
                blocks.append((convs, down))
            self.layers.append(blocks)
        # one fp64 workspace for the batch statistics of every conv (zeroed with ONE fill per forward instead of one per
        # layer) and one multi-tensor add for the BatchNorm step counters: ~100 fewer tiny launches per step.  The backward
        # reductions sum(dy'), sum(dy' (x - mean)) get a same-shaped workspace (zeroed once per backward).
        self._all_convs = [self.stem] + [c for blocks in self.layers for convs, down in blocks
                                         for c in (convs + ([down] if down is not None else []))]
        off = 0
        for c in self._all_convs:
            c.stats_off = off
            off += 2 * c.O
        self._stats_total = off
        self._stats_ws = None
        self._bwd_ws = None
        self._nbt = []
        # called with the stage index (3 = layer4 ... 0 = layer1) right after that stage's backward kernels have been
        # enqueued: the data-parallel trainer uses it to start the stage's gradient all-reduce early
        self.on_stage_backward_done = None

    # ------------------------------------------------------------------ forward pieces
    def _conv_bn(self, c, x, B, H, W, relu, residual, training, need_grad, tta=None):
        """x: [B*H*W, Cin] bf16 (or the NCHW fp32 image for the stem).  Returns y, Ho, Wo, record-for-backward."""
        if c.stem:
            A, Ho, Wo = ops.im2col_nchw_f32(x, c.R, c.S, c.stride, c.pad, c.ldk, tta=tta)
        elif c.direct:
            A, Ho, Wo = x, H, W
        elif c.implicit:
            A, Ho, Wo = x, _out_hw(H, c.R, c.stride, c.pad), _out_hw(W, c.S, c.stride, c.pad)
        else:
            A, Ho, Wo = ops.im2col_nhwc(x, B, H, W, c.I, c.R, c.S, c.stride, c.pad)
        rows = B * Ho * Wo
        bn = c.bn
        # layers with a residual input cannot rebuild their ReLU mask from (raw, scale, shift): the forward leaves a 1-bit
        # mask for the backward (instead of the backward re-reading the whole bf16 output twice)
        want_mask = need_grad and relu and residual is not None
        gkw = dict(conv=c.conv_desc(1, B, H, W), M=rows, K=c.K) if c.implicit else {}
        track = bn.track_running_stats and bn.running_mean is not None
        if training:
            st = self._stats_ws[c.stats_off:c.stats_off + 2 * c.O].view(2, c.O)
            if FUSE_BN_STATS_IN_GEMM and c.K >= FUSE_BN_STATS_MIN_K:
                raw = ops.gemm(A, c.wp, colsum=st[0], colsumsq=st[1], N=c.O, **gkw)
            else:
                raw = ops.gemm(A, c.wp, N=c.O, **gkw)
                ops.col_stats(raw, st[0], st[1])
            mom = bn.momentum if bn.momentum is not None else 0.1
            out = ops.bn_fwd(raw, st[0], st[1], bn.weight.data, bn.bias.data, bn.running_mean if track else None,
                             bn.running_var if track else None, mom, bn.eps, residual=residual, relu=relu, training=True,
                             want_mask=want_mask)
            if track and bn.num_batches_tracked is not None:
                self._nbt.append(bn.num_batches_tracked)
        else:
            raw = ops.gemm(A, c.wp, N=c.O, **gkw)
            out = ops.bn_fwd(raw, None, None, bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, 0.0, bn.eps,
                             residual=residual, relu=relu, training=False, want_mask=want_mask)
        y, mean, invstd, scale, shift = out[:5]
        mask = out[5] if want_mask else None
        rec = None
        if need_grad:
            # y is only kept for the backward ReLU mask when a residual entered the activation; otherwise the mask is
            # recomputed from (raw, scale, shift)
            rec = dict(c=c, A=A, raw=raw, y=None, mask=mask, mean=mean, invstd=invstd,
                       scale=scale, shift=shift, relu=relu, B=B, H=H, W=W, Ho=Ho, Wo=Wo, has_res=residual is not None,
                       training=training, sums=None)
        return y, Ho, Wo, rec

    def begin_forward(self, images, training):
        self._nbt = []
        if training:
            self._stats_ws = torch.zeros(self._stats_total, device=images.device, dtype=torch.float64)

    def end_forward(self):
        if self._nbt:
            torch._foreach_add_(self._nbt, 1)
            self._nbt = []

    def forward_stem(self, images, training, need_grad, tta=None):
        """images [B,3,H,W] fp32 -> (x [B*H2*W2, 64] bf16 after the 3x3/2 max-pool, H2, W2, ctx).  tta = (V, codes): the
        trunk runs on V test-time-augmentation variants of every image (batch V*B), generated by the stem's im2col."""
        B, _, H, W = images.shape
        B = B * (tta[0] if tta else 1)
        x, H1, W1, rec = self._conv_bn(self.stem, images.contiguous(), B, H, W, True, None, training, need_grad, tta=tta)
        y, idx, H2, W2 = ops.maxpool_fwd(x, B, H1, W1, self.stem.O)
        ctx = dict(rec=rec, pool=(idx, H1, W1, self.stem.O), B=B) if need_grad else None
        return y, H2, W2, ctx

    def forward_layer(self, li, x, B, H, W, training, need_grad):
        """One ResNet stage (li = 0..3).  Returns (y, H, W, C, ctx)."""
        blocks = [] if need_grad else None
        for convs, down in self.layers[li]:
            inp, Hi, Wi = x, H, W
            t = inp
            main = []
            for c in convs[:-1]:
                t, H, W, r = self._conv_bn(c, t, B, H, W, True, None, training, need_grad)
                main.append(r)
            rdown = None
            if down is not None:
                idn, _, _, rdown = self._conv_bn(down, inp, B, Hi, Wi, False, None, training, need_grad)
            else:
                idn = inp
            x, H, W, r = self._conv_bn(convs[-1], t, B, H, W, True, idn, training, need_grad)
            main.append(r)
            if need_grad:
                blocks.append(dict(main=main, down=rdown))
        return x, H, W, self.layers[li][-1][0][-1].O, (dict(blocks=blocks, B=B) if need_grad else None)

    def forward(self, images, training, need_grad, tta=None):
        """images: [B,3,H,W] fp32 CUDA.  Returns ({name: (feat2d bf16, h, w, C)}, ctx).  tta: see forward_stem (inference)."""
        if tta is not None and need_grad:
            raise ValueError("test-time augmentation inside the stem is an inference path (no backward)")
        B = images.shape[0] * (tta[0] if tta else 1)
        self.begin_forward(images, training)
        x, H, W, sctx = self.forward_stem(images, training, need_grad, tta=tta)
        feats, lctx = {}, []
        for li in range(4):
            x, H, W, C, c = self.forward_layer(li, x, B, H, W, training, need_grad)
            feats[f"layer{li + 1}"] = (x, H, W, C)
            lctx.append(c)
        self.end_forward()
        return feats, (dict(stem=sctx, layers=lctx, B=B) if need_grad else None)

    # ------------------------------------------------------------------ backward pieces
    def begin_backward(self, device):
        """One zero-fill for every fused backward reduction of this pass."""
        self._bwd_ws = (torch.zeros(self._stats_total, device=device, dtype=torch.float64)
                        if (FUSE_BN_BWD_REDUCE and ops.HAS_GEMM_STAT) else None)

    def _wgrad(self, c, rec, draw, A, rows, st):
        """Weight gradient of one convolution (accumulates into the flat fp32 gradient buffer)."""
        if c.implicit:
            # wgrad as an implicit GEMM: B operand = im2col(x) through TMA; plain 1x1 (strided) convs accumulate
            # straight into the flat gradient buffer, k x k ones into the packed buffer
            cd = c.conv_desc(2, rec["B"], rec["H"], rec["W"])
            if c.plain:
                gw = st.g32(c.conv.weight).view(c.O, c.I)
                ops.gemm(draw, A, a_mn=True, b_mn=True, out=gw, accumulate=True, split_k=-1, M=c.O, N=c.K, K=rows, conv=cd)
            else:
                c.gp.zero_()
                ops.gemm(draw, A, a_mn=True, b_mn=True, out=c.gp, accumulate=True, split_k=-1, M=c.O, N=c.K, K=rows, conv=cd)
                ops.conv_wgrad_unpack(c.gp, st.g32(c.conv.weight))
        elif c.plain:
            gw = st.g32(c.conv.weight).view(c.O, c.I)
            ops.gemm(draw, A, a_mn=True, b_mn=True, out=gw, accumulate=True, split_k=-1,
                     M=c.O, N=c.I, K=rows)
        else:
            c.gp.zero_()
            ops.gemm(draw, A, a_mn=True, b_mn=True, out=c.gp, accumulate=True, split_k=-1,
                     M=c.O, N=c.ldk, K=rows)
            ops.conv_wgrad_unpack(c.gp, st.g32(c.conv.weight))

    def _conv_bn_bwd(self, rec, dy, add_to_dx, need_dx=True, producer=None):
        """dy: grad wrt the BN(+res)(+relu) output.  Returns (dx, dz) where dz is the masked dy (identity branch).
        producer: record of the conv+BN whose OUTPUT is this convolution's input (same block, no residual in between):
        when given (and fusable) the dgrad GEMM's epilogue also accumulates that BatchNorm's backward reductions."""
        c = rec["c"]
        st = self.store
        train_w = c.conv.weight.requires_grad
        bnw = c.bn.weight
        dgamma = st.g32(bnw) if bnw.requires_grad else None
        dbeta = st.g32(c.bn.bias) if bnw.requires_grad else None
        draw, dz = ops.bn_bwd(dy, rec["raw"], rec["y"], rec["mean"], rec["invstd"], bnw.data, dgamma, dbeta,
                              relu=rec["relu"], want_dz=rec["has_res"], scale=rec["scale"], shift=rec["shift"],
                              training=rec["training"], sums=rec["sums"], mask=rec["mask"])
        A = rec["A"]
        rows = rec["B"] * rec["Ho"] * rec["Wo"]
        side = None
        if train_w:
            if runtime.OVERLAP_WGRAD and need_dx:
                side = runtime.fork_side()
            with torch.cuda.stream(side) if side is not None else _nullctx():
                self._wgrad(c, rec, draw, A, rows, st)
        dx = None
        if need_dx:
            skw = {}
            # (measured, tools/bench_gemm_step.py: the column pass is free under a main-loop-bound GEMM -- 3x3 dgrad, K >= 576 --
            #  but an epilogue-bound 1x1 dgrad with a short reduction pays more for it than the separate reduce kernel costs:
            #  401408x64x256: 43 -> 86 us against a 28 us reduce; 25088x256x1024: 19 -> 27 us against 13 us)
            if (producer is not None and self._bwd_ws is not None and c.dgrad_fusable and add_to_dx is None
                    and not producer["has_res"] and (c.implicit_dgrad or c.O >= FUSE_BN_BWD_MIN_K)):
                pc = producer["c"]
                sums = self._bwd_ws[pc.stats_off:pc.stats_off + 2 * pc.O].view(2, pc.O)
                skw = dict(stat_x=producer["raw"], stat_mean=producer["mean"], stat_scale=producer["scale"],
                           stat_shift=producer["shift"], stat_relu=producer["relu"], colsum=sums[0], colsumsq=sums[1])
                producer["sums"] = sums
            if c.direct:
                dx = ops.gemm(draw, c.wp, b_mn=True, residual=add_to_dx, M=rows, N=c.I, K=c.O, **skw)
            elif c.implicit_dgrad:
                # stride-1 k x k dgrad = the same implicit GEMM over dY with the flipped / transposed filter
                dx = ops.gemm(draw, c.wt, residual=add_to_dx, M=rec["B"] * rec["H"] * rec["W"], N=c.I, K=c.R * c.S * c.O,
                              conv=(1, rec["B"], rec["Ho"], rec["Wo"], c.O, c.R, c.S, 1, c.R - 1 - c.pad), **skw)
            else:
                dcol = ops.gemm(draw, c.wp[:, :c.K], b_mn=True, M=rows, N=c.K, K=c.O)
                dx = ops.col2im_nhwc(dcol, rec["B"], rec["H"], rec["W"], c.I, c.R, c.S, c.stride, c.pad, add=add_to_dx)
        if side is not None:
            runtime.join_side(side)
        return dx, dz

    def backward_layer(self, li, ctx, dx):
        """dx: grad of the stage output [rows, C] bf16.  Returns the grad of the stage input."""
        for blk in reversed(ctx["blocks"]):
            main, down = blk["main"], blk["down"]
            d, dz = self._conv_bn_bwd(main[-1], dx, None, producer=main[-2] if len(main) > 1 else None)
            for j in range(len(main) - 2, 0, -1):
                d, _ = self._conv_bn_bwd(main[j], d, None, producer=main[j - 1])
            if down is not None:
                d_idn, _ = self._conv_bn_bwd(down, dz, None)
            else:
                d_idn = dz
            dx, _ = self._conv_bn_bwd(main[0], d, d_idn)
        if self.on_stage_backward_done is not None:
            self.on_stage_backward_done(li)
        return dx

    def backward_stem(self, ctx, dx):
        idx, H1, W1, C = ctx["pool"]
        d = ops.maxpool_bwd(dx, idx, ctx["B"], H1, W1, C)
        self._conv_bn_bwd(ctx["rec"], d, None, need_dx=False)

    def backward(self, ctx, dfeats):
        """dfeats: {name: grad 2-D bf16 or None}.  Accumulates parameter gradients; images get no gradient."""
        dx = None
        for li in range(3, -1, -1):
            g = dfeats.get(f"layer{li + 1}")
            if g is not None:
                if dx is None:
                    self.begin_backward(g.device)
                dx = g if dx is None else dx + g
            if dx is None:
                continue  # nothing downstream needs this stage
            dx = self.backward_layer(li, ctx["layers"][li], dx)
        if dx is None:
            return
        self.backward_stem(ctx["stem"], dx)
