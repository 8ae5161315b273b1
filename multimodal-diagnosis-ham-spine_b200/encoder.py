"""ImageEncoder / TextEncoder with the reference's constructors and state_dict keys
(reference: encoder.py:13-134), computing on the B200 kernels.

* `ImageEncoder.model` is a torchvision ResNet used purely as the parameter container, so the key set
  (`model.*` plus the aliased `stem.*`, `layer1..4.*`) and the default initialisation are the
  reference's.  `backbone` additionally accepts "resnet50" (channels 512/1024/2048), which the
  north-star configuration needs and encoder.py:31-33 rejects.
* `TextEncoder.model` is a transformers BertModel container loaded with `from_pretrained(model_path)`.
* Outputs are bf16 CUDA tensors `(B, tokens, feature_dim)` taking part in autograd; inputs are fp32
  NCHW images / int64 ids + mask exactly as the reference receives them.
"""
import os

import torch
import torch.nn as nn
from torchvision.models import convnext_base, convnext_small, convnext_tiny, resnet18, resnet34, resnet50

from . import functional as Fm
from . import ops
from .bert_engine import BertEngine, qkv_groups
from .resnet_engine import ResNetEngine
from .runtime import ParamStore


# --------------------------------------------------------------------------------------------
# binding a module tree to one ParamStore
# --------------------------------------------------------------------------------------------
def collect_groups(root):
    groups = []
    for m in root.modules():
        fn = getattr(m, "_mdhs_groups", None)
        if fn is not None:
            groups.extend(fn())
    return groups


def bind(root, device):
    """Create the flat ParamStore for `root` and hand it to every mdhs module underneath."""
    store = ParamStore(root, device, groups=collect_groups(root))
    for m in root.modules():
        if hasattr(m, "_mdhs_bind"):
            m._mdhs_bind(store)
    return store


class MdhsModule(nn.Module):
    """Base class: lazily binds to a ParamStore on the first CUDA forward (or re-binds after .to())."""

    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_store", None)

    def _mdhs_bind(self, store):
        object.__setattr__(self, "_store", store)
        self._on_bind(store)

    def _on_bind(self, store):
        pass

    def _probe_param(self):
        for p in self.parameters():
            return p
        return None

    def store(self, device):
        st = self._store
        p = self._probe_param()
        if st is None or (p is not None and not st.owns(p)):
            dev = torch.device(device)
            if dev.type != "cuda":
                raise ops._lib.MdhsError("mdhs_b200 modules run on CUDA only: move the inputs/model to a B200 "
                                         "(there is no CPU fallback)")
            # non-parameter buffers (BN running stats) follow the module to the device
            for b in self.buffers():
                if b.device != dev:
                    b.data = b.data.to(dev)
            st = bind(self, dev)
        st.refresh()
        if torch.is_grad_enabled():
            st.attach_grads()
        return st


class _TrunkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, images, engine, training, names, need, tta=None):
        feats, c = engine.forward(images, training, need, tta=tta)
        # frozen trunk (need == False): nothing was saved, the backward below must be a no-op
        ctx.engine, ctx.c, ctx.names = engine, (c if need else None), names
        return tuple(feats[n][0] for n in names)

    @staticmethod
    def backward(ctx, *grads):
        if ctx.c is not None:  # frozen trunk: outputs still carry grad so that downstream Functions run backward
            ctx.engine.backward(ctx.c, {n: (g.contiguous() if g is not None else None) for n, g in zip(ctx.names, grads)})
        ctx.c = None
        return None, None, None, None, None, None, None


# ---- stage-wise trunk (only used while analysis hooks are registered on image_encoder.stem / layerN / layerN[-1]) ----------
class _StemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, images, engine, training, need):
        engine.begin_forward(images, training)
        y, H, W, c = engine.forward_stem(images, training, need)
        ctx.engine, ctx.c = engine, c
        return y

    @staticmethod
    def backward(ctx, dy):
        if ctx.c is not None:
            ctx.engine.backward_stem(ctx.c, dy.contiguous())
        ctx.c = None
        return None, None, None, None, None


class _StageFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, engine, li, B, H, W, training, need, first_backward):
        y, Ho, Wo, C, c = engine.forward_layer(li, x, B, H, W, training, need)
        ctx.engine, ctx.c, ctx.li, ctx.first_backward = engine, c, li, first_backward
        ctx.out_hw = (Ho, Wo, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        if ctx.c is None:
            return (None,) * 9
        if ctx.first_backward:
            ctx.engine.begin_backward(dy.device)
        dx = ctx.engine.backward_layer(ctx.li, ctx.c, dy.contiguous())
        ctx.c = None
        return (dx,) + (None,) * 8


class _ToNchwFn(torch.autograd.Function):
    """[B*H*W, C] bf16 (NHWC) -> [B, C, H, W] fp32 -- the layout / dtype the reference's hooks see."""

    @staticmethod
    def forward(ctx, x2d, B, H, W, C):
        ctx.dims = (B, H, W, C)
        return ops.nhwc_bf16_to_nchw_f32(x2d.contiguous(), B, H, W, C)

    @staticmethod
    def backward(ctx, g):
        B, H, W, C = ctx.dims
        return ops.nchw_f32_to_nhwc_bf16(g.contiguous().float()), None, None, None, None


class _ToNhwcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x4d):
        B, C, H, W = x4d.shape
        ctx.dims = (B, H, W, C)
        return ops.nchw_f32_to_nhwc_bf16(x4d.contiguous().float())

    @staticmethod
    def backward(ctx, g2d):
        B, H, W, C = ctx.dims
        return ops.nhwc_bf16_to_nchw_f32(g2d.contiguous(), B, H, W, C)


def _has_hooks(m):
    return bool(m._forward_hooks or m._forward_pre_hooks or m._backward_hooks or getattr(m, "_backward_pre_hooks", None))


def _fire_hooks(module, inp, out):
    """Run `module`'s registered hooks around a precomputed output: module.__call__ is invoked with its `forward`
    temporarily replaced by a function that returns `out`, so forward (pre-)hooks see (input, output) and full-backward
    hooks receive grad_output exactly as they would around the reference's torchvision block.  `inp` carries no grad
    (the real data path runs through our kernels), which makes torch deliver the backward hook when grad_output arrives."""
    object.__setattr__(module, "forward", lambda *a, **k: out)
    try:
        return module(inp.detach())
    finally:
        object.__delattr__(module, "forward")


class _BertFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, input_ids, attention_mask, engine, training, need, taps):
        if taps:
            h, c, tapped = engine.forward(input_ids, attention_mask, training, need, taps=taps)
            ctx.engine, ctx.c, ctx.taps = engine, c, tuple(taps)
            return (h,) + tuple(tapped[n] for n in taps)
        h, c = engine.forward(input_ids, attention_mask, training, need)
        ctx.engine, ctx.c, ctx.taps = engine, c, ()
        return h

    @staticmethod
    def backward(ctx, dh, *dtaps):
        if ctx.c is not None:
            dt = {n: g for n, g in zip(ctx.taps, dtaps) if g is not None}
            ctx.engine.backward(ctx.c, dh.contiguous() if dh is not None else None, dt or None)
        ctx.c = None
        return None, None, None, None, None, None, None


_CHANNELS = {
    "resnet18": {"layer2": 128, "layer3": 256, "layer4": 512},
    "resnet34": {"layer2": 128, "layer3": 256, "layer4": 512},
    "resnet50": {"layer2": 512, "layer3": 1024, "layer4": 2048},
}
_BUILDERS = {"resnet18": resnet18, "resnet34": resnet34, "resnet50": resnet50}
# ConvNeXt backbones (the reference only reaches ConvNeXt through ConNexT/models/ourmodel.py): the residual
# stages 2 / 3 / 4 of torchvision `features` play the role of layer2 / layer3 / layer4
_CONVNEXT = {"convnext_tiny": (convnext_tiny, {"layer2": 192, "layer3": 384, "layer4": 768}),
             "convnext_small": (convnext_small, {"layer2": 192, "layer3": 384, "layer4": 768}),
             "convnext_base": (convnext_base, {"layer2": 256, "layer3": 512, "layer4": 1024})}


class ImageEncoder(MdhsModule):
    """ResNet backbone returning patch tokens projected to `feature_dim` (encoder.py:13-109)."""

    def __init__(self, feature_dim=512, pretrained=True, weights_path=None, backbone="resnet18", multi_scale=False):
        super().__init__()
        self.multi_scale = multi_scale
        backbone = backbone.lower()
        self.is_convnext = backbone in _CONVNEXT
        if backbone not in _BUILDERS and not self.is_convnext:
            raise ValueError(f"Unsupported backbone: {backbone}. Use resnet18, resnet34, resnet50, convnext_tiny, "
                             "convnext_small or convnext_base.")
        build_model = _CONVNEXT[backbone][0] if self.is_convnext else _BUILDERS[backbone]
        channels = _CONVNEXT[backbone][1] if self.is_convnext else _CHANNELS[backbone]
        if weights_path:
            self.model = build_model(weights=None)
            if not os.path.exists(weights_path):
                raise FileNotFoundError(f"weights file not found: {weights_path}")
            state_dict = torch.load(weights_path, map_location="cpu")
            self.model.load_state_dict(state_dict, strict=False)
        elif pretrained:
            self.model = build_model(weights="DEFAULT")
        else:
            self.model = build_model(weights=None)
        if self.is_convnext:
            # state_dict keys: image_encoder.model.features.* (torchvision ConvNeXt; avg-pool / classifier are bypassed)
            self.model.classifier = nn.Identity()
            self.model.avgpool = nn.Identity()
        else:
            self.model.fc = nn.Identity()
            self.stem = nn.Sequential(self.model.conv1, self.model.bn1, self.model.relu, self.model.maxpool)
            self.layer1 = self.model.layer1
            self.layer2 = self.model.layer2
            self.layer3 = self.model.layer3
            self.layer4 = self.model.layer4
        if self.multi_scale:
            self.proj2 = nn.Linear(channels["layer2"], feature_dim)
            self.proj3 = nn.Linear(channels["layer3"], feature_dim)
        self.proj4 = nn.Linear(channels["layer4"], feature_dim)
        object.__setattr__(self, "_engine", None)

    def _on_bind(self, store):
        if self.is_convnext:
            from .connext.convnext import ConvNeXtEngine
            object.__setattr__(self, "_engine", ConvNeXtEngine(store, self.model.features))
        else:
            object.__setattr__(self, "_engine", ResNetEngine(store, self.model))

    def _trainable(self):
        return any(p.requires_grad for p in self.model.parameters())

    def _forward_convnext(self, st, x):
        """ConvNeXt trunk: autograd runs through the engine's own Functions; stage outputs are already NHWC token matrices."""
        B = x.shape[0]
        taps = []
        self._engine.forward(x.float(), self.training, taps=taps)
        stages = {"layer2": taps[1], "layer3": taps[2], "layer4": taps[3]}
        names = ("layer2", "layer3", "layer4") if self.multi_scale else ("layer4",)
        out = {}
        for name in names:
            proj = getattr(self, "proj" + name[-1])
            t = Fm.linear(stages[name][0], st, proj.weight, proj.bias)
            out[name] = t.view(B, -1, t.shape[1])
        return out if self.multi_scale else out["layer4"]

    def _hooked_stages(self):
        """Analysis hooks (Grad-CAM: analysis_tools.py:29-31, scripts/run_analysis.py:126-132) on the stem or on a stage /
        its last block: {stage index (-1 = stem): [modules whose hooks fire at that boundary]}."""
        out = {}
        if _has_hooks(self.stem):
            out[-1] = [self.stem]
        for li, layer in enumerate((self.layer1, self.layer2, self.layer3, self.layer4)):
            mods = [m for m in (layer[-1], layer) if _has_hooks(m)]
            if mods:
                out[li] = mods
        return out

    def _forward_staged(self, st, x, hooked, names, need):
        """Stage-by-stage trunk with NCHW fp32 taps at the hooked boundaries (slower than the fused path: extra layout
        conversions; only taken while hooks are registered)."""
        eng = self._engine
        B = x.shape[0]
        need = need or torch.is_grad_enabled()   # Grad-CAM back-propagates through a (possibly frozen / eval) trunk
        cur = _StemFn.apply(st.anchor, x, eng, self.training, need)
        H, W = _out_hw_stem(x.shape[2]), _out_hw_stem(x.shape[3])
        C = eng.stem.O
        prev4d = x
        feats = {}
        first_backward_at = max(i for i, n in enumerate(("layer1", "layer2", "layer3", "layer4")) if n in names)
        for li in range(-1, 4):
            if li >= 0:
                cur = _StageFn.apply(cur, eng, li, B, H, W, self.training, need, li == first_backward_at)
                blocks = eng.layers[li]
                C = blocks[-1][0][-1].O
                s = blocks[0][0][1].stride if len(blocks[0][0]) > 2 else blocks[0][0][0].stride
                H, W = (H + s - 1) // s, (W + s - 1) // s
            if li in hooked:
                t4 = _ToNchwFn.apply(cur, B, H, W, C)
                for m in hooked[li]:
                    t4 = _fire_hooks(m, prev4d, t4)
                prev4d = t4
                cur = _ToNhwcFn.apply(t4)
            if li >= 0 and f"layer{li + 1}" in names:
                feats[f"layer{li + 1}"] = cur
        eng.end_forward()
        return tuple(feats[n] for n in names)

    def forward(self, x, tta=None):
        """tta: optional tuple of transform names (scripts/predict.py:33-42: "hflip", "vflip", "rot90"): inference only; the
        returned tokens then hold [identity] + transforms variants stacked on the batch axis ((1 + len(tta)) * B samples),
        produced by the stem's im2col addressing instead of a materialised augmented batch."""
        st = self.store(x.device)
        if tta:
            if self.is_convnext or torch.is_grad_enabled() or self._hooked_stages():
                x, tta = ops.tta_expand(x, tta), None       # generic path: materialise the variants
            else:
                tta = ops.tta_codes(tta)
        if self.is_convnext:
            return self._forward_convnext(st, x)
        B = x.shape[0] * (tta[0] if tta else 1)
        names = ("layer2", "layer3", "layer4") if self.multi_scale else ("layer4",)
        need = self._trainable() and torch.is_grad_enabled()
        hooked = self._hooked_stages()
        if hooked:
            feats = self._forward_staged(st, x.float(), hooked, names, need)
        else:
            feats = _TrunkFn.apply(st.anchor, x.float(), self._engine, self.training, names, need, tta or None)
        if self.multi_scale:
            out = {}
            for name, f, proj in zip(names, feats, (self.proj2, self.proj3, self.proj4)):
                t = Fm.linear(f, st, proj.weight, proj.bias)
                out[name] = t.view(B, -1, t.shape[1])
            return out
        t = Fm.linear(feats[0], st, self.proj4.weight, self.proj4.bias)
        return t.view(B, -1, t.shape[1])


def _out_hw_stem(h):
    h = (h + 2 * 3 - 7) // 2 + 1        # 7x7 / 2, pad 3
    return (h + 2 * 1 - 3) // 2 + 1     # 3x3 / 2 max-pool, pad 1


class TextEncoder(MdhsModule):
    """BERT wrapper returning last_hidden_state (encoder.py:112-134)."""

    def __init__(self, model_path="bert-base-uncased", feature_dim=768):
        super().__init__()
        from transformers import BertModel
        self.model = BertModel.from_pretrained(model_path)
        object.__setattr__(self, "_engine", None)

    def _mdhs_groups(self):
        return qkv_groups(self.model)

    def _on_bind(self, store):
        object.__setattr__(self, "_engine", BertEngine(store, self.model))

    def forward(self, input_ids, attention_mask, hidden_states=None):
        """hidden_states: optional tuple of encoder-layer numbers (1-based, HF `output_hidden_states` indexing): returns
        {n: (B, S, 768)} of those layers' outputs instead of the last hidden state alone (README.md:15: text hidden states
        4 / 8 / 12 paired with image layer2 / 3 / 4)."""
        st = self.store(input_ids.device)
        B, S = input_ids.shape
        trainable = any(p.requires_grad for p in self.model.encoder.parameters())
        need = trainable and torch.is_grad_enabled()
        if hidden_states:
            n_layers = len(self.model.encoder.layer)
            taps = tuple(int(n) for n in hidden_states if int(n) != n_layers)
            for n in taps:
                if not 1 <= n <= n_layers:
                    raise ValueError(f"hidden state {n} out of range 1..{n_layers}")
            outs = _BertFn.apply(st.anchor, input_ids, attention_mask, self._engine, self.training, need, taps)
            outs = outs if isinstance(outs, tuple) else (outs,)
            res = {n: t.view(B, S, t.shape[1]) for n, t in zip(taps, outs[1:])}
            if n_layers in [int(n) for n in hidden_states]:
                res[n_layers] = outs[0].view(B, S, outs[0].shape[1])
            return res
        h = _BertFn.apply(st.anchor, input_ids, attention_mask, self._engine, self.training, need, ())
        return h.view(B, S, h.shape[1])
