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
from torchvision.models import resnet18, resnet34, resnet50

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
        return st


class _TrunkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, images, engine, training, names, need):
        feats, c = engine.forward(images, training, need)
        # frozen trunk (need == False): nothing was saved, the backward below must be a no-op
        ctx.engine, ctx.c, ctx.names = engine, (c if need else None), names
        return tuple(feats[n][0] for n in names)

    @staticmethod
    def backward(ctx, *grads):
        if ctx.c is not None:  # frozen trunk: outputs still carry grad so that downstream Functions run backward
            ctx.engine.backward(ctx.c, {n: (g.contiguous() if g is not None else None) for n, g in zip(ctx.names, grads)})
        ctx.c = None
        return None, None, None, None, None, None


class _BertFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, input_ids, attention_mask, engine, training, need):
        h, c = engine.forward(input_ids, attention_mask, training, need)
        ctx.engine, ctx.c = engine, c
        return h

    @staticmethod
    def backward(ctx, dh):
        if ctx.c is not None:
            ctx.engine.backward(ctx.c, dh.contiguous())
        ctx.c = None
        return None, None, None, None, None, None


_CHANNELS = {
    "resnet18": {"layer2": 128, "layer3": 256, "layer4": 512},
    "resnet34": {"layer2": 128, "layer3": 256, "layer4": 512},
    "resnet50": {"layer2": 512, "layer3": 1024, "layer4": 2048},
}
_BUILDERS = {"resnet18": resnet18, "resnet34": resnet34, "resnet50": resnet50}


class ImageEncoder(MdhsModule):
    """ResNet backbone returning patch tokens projected to `feature_dim` (encoder.py:13-109)."""

    def __init__(self, feature_dim=512, pretrained=True, weights_path=None, backbone="resnet18", multi_scale=False):
        super().__init__()
        self.multi_scale = multi_scale
        backbone = backbone.lower()
        if backbone not in _BUILDERS:
            raise ValueError(f"Unsupported backbone: {backbone}. Use resnet18, resnet34 or resnet50.")
        build_model = _BUILDERS[backbone]
        channels = _CHANNELS[backbone]
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
        object.__setattr__(self, "_engine", ResNetEngine(store, self.model))

    def _trainable(self):
        return any(p.requires_grad for p in self.model.parameters())

    def forward(self, x):
        st = self.store(x.device)
        B = x.shape[0]
        names = ("layer2", "layer3", "layer4") if self.multi_scale else ("layer4",)
        need = self._trainable() and torch.is_grad_enabled()
        feats = _TrunkFn.apply(st.anchor, x.float(), self._engine, self.training, names, need)
        if self.multi_scale:
            out = {}
            for name, f, proj in zip(names, feats, (self.proj2, self.proj3, self.proj4)):
                t = Fm.linear(f, st, proj.weight, proj.bias)
                out[name] = t.view(B, -1, t.shape[1])
            return out
        t = Fm.linear(feats[0], st, self.proj4.weight, self.proj4.bias)
        return t.view(B, -1, t.shape[1])


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

    def forward(self, input_ids, attention_mask):
        st = self.store(input_ids.device)
        B, S = input_ids.shape
        trainable = any(p.requires_grad for p in self.model.encoder.parameters())
        need = trainable and torch.is_grad_enabled()
        h = _BertFn.apply(st.anchor, input_ids, attention_mask, self._engine, self.training, need)
        return h.view(B, S, h.shape[1])
