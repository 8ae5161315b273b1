"""ConNexT image+text classifier on the B200 kernels (ConNexT/models/ourmodel.py:9-95).

Same constructors, forward(batch dict) -> logits and state_dict keys as the reference:
`text_encoder.bert.*` (BertModel container), `image_encoder.*` (torchvision `convnext_*.features` container),
`conv`, `{textbased,imagbased}_cross_attention.{query,key,value}_conv`, `fc`.

The two CrossAttention blocks degenerate on a 1-token text (ourmodel.py:17-31):
  * image -> text (keys = the single text token): softmax over one key is 1, so every image position receives
    value_conv(text) and the average pool returns exactly that vector; query_conv / key_conv are dead compute and
    get zero gradient (as they do in the reference).
  * text -> image: one query over the 49 image positions, softmax(q.k_t) WITHOUT 1/sqrt(d), out = sum_t p_t v_t.
"""
import torch
import torch.nn as nn
from torchvision import models

from .. import functional as Fm
from .. import runtime
from ..encoder import MdhsModule
from ..mibf_net.bert import BertEncoder as _BertEncoder
from .convnext import ConvNeXtEngine, SqAttnFn

_VARIANTS = {"tiny": models.convnext_tiny, "small": models.convnext_small, "base": models.convnext_base,
             "large": models.convnext_large}
_WEIGHTS = {"tiny": "ConvNeXt_Tiny_Weights", "small": "ConvNeXt_Small_Weights", "base": "ConvNeXt_Base_Weights",
            "large": "ConvNeXt_Large_Weights"}


class BertEncoder(_BertEncoder):
    """ConNexT/models/BERT.py:7-21 (CLS vector); `model_path` replaces the path hard-coded at BERT.py:11."""


class CrossAttention(nn.Module):
    """Parameter container with the reference's layout (three 1x1 Conv2d); evaluated by the owning model."""

    def __init__(self, dim):
        super().__init__()
        self.query_conv = nn.Conv2d(dim, dim, kernel_size=1)
        self.key_conv = nn.Conv2d(dim, dim, kernel_size=1)
        self.value_conv = nn.Conv2d(dim, dim, kernel_size=1)
        self.softmax = nn.Softmax(dim=-1)


def _conv1x1(x, st, conv):
    """1x1 Conv2d on a token matrix [rows, C_in] bf16 = Linear with the [O, I, 1, 1] weight viewed as [O, I]."""
    O, I = conv.weight.shape[:2]
    tr = conv.weight.requires_grad
    return Fm.linear(x, st, None, w16=st.w16(conv.weight).view(O, I), gw=st.g32(conv.weight).view(O, I) if tr else None,
                     b32=conv.bias.data if conv.bias is not None else None,
                     gb=st.g32(conv.bias) if (tr and conv.bias is not None) else None)


class ConvNeXtEncoder(MdhsModule):
    """ConNexT/models/pl_model_MOE2.py:29-53: torchvision ConvNeXt features flattened to (B, C, H*W).
    The reference hard-codes convnext_large; `variant` also admits tiny / small / base (config 4 uses Tiny)."""

    def __init__(self, pretrained=True, variant="large"):
        super().__init__()
        weights = getattr(models, _WEIGHTS[variant]).DEFAULT if pretrained else None
        self.features = _VARIANTS[variant](weights=weights).features
        self.output_dim = self.features[-1][-1].block[0].weight.shape[0]
        object.__setattr__(self, "_engine", None)

    def _on_bind(self, store):
        object.__setattr__(self, "_engine", ConvNeXtEngine(store, self.features))

    def forward_tokens(self, x):
        """(B*h*w, C) bf16 token matrix (NHWC) -- what the downstream B200 modules consume -- plus (h, w)."""
        self.store(x.device)
        return self._engine.forward(x.float(), self.training)

    def forward(self, x):
        t, h, w = self.forward_tokens(x)
        B = x.shape[0]
        return t.view(B, h * w, -1).transpose(1, 2)      # (B, C, H*W) view, like feature_map.flatten(2)


class OurClassfierConvnextV2(MdhsModule):
    def __init__(self, num_labels=2, pretrained=True, pretrained_path="/data/QLI/ConNexT/convnext-base-224",
                 bert_path="/data/QLI/BERT_pretain", variant="base"):
        super().__init__()
        self.text_encoder = BertEncoder(model_path=bert_path)
        # the HuggingFace checkpoint branch of the reference (ourmodel.py:41-48) needs files that are not part of it;
        # like the reference when that load fails, fall through to the torchvision model
        weights = None
        if pretrained:
            try:
                weights = getattr(models, _WEIGHTS[variant]).DEFAULT
            except Exception:
                weights = None
        try:
            convnext_model = _VARIANTS[variant](weights=weights)
        except Exception:
            convnext_model = _VARIANTS[variant](weights=None)
        self.image_encoder = convnext_model.features
        c_last = self.image_encoder[-1][-1].block[0].weight.shape[0]
        self.conv = nn.Conv2d(in_channels=c_last, out_channels=768, kernel_size=1)
        self.textbased_cross_attention = CrossAttention(dim=768)
        self.imagbased_cross_attention = CrossAttention(dim=768)
        self.avg_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(768, num_labels)
        object.__setattr__(self, "_trunk", None)

    def _on_bind(self, store):
        object.__setattr__(self, "_trunk", ConvNeXtEngine(store, self.image_encoder))

    def forward(self, batch_data):
        images = batch_data["transformed_image"]
        st = self.store(images.device)
        B = images.shape[0]
        text = self.text_encoder(batch_data["input_ids"], batch_data["attention_mask"]).contiguous()   # (B, 768) bf16
        tokens, h, w = self._trunk.forward(images.float(), self.training)                              # (B*49, C) bf16
        T = h * w
        img = _conv1x1(tokens, st, self.conv)                                                          # (B*49, 768)
        # image -> text: one key, attention == 1 -> pooled output == value_conv(text)
        pooled_1 = Fm.to_f32(_conv1x1(text, st, self.textbased_cross_attention.value_conv))
        # text -> image: single query over the T image positions (no 1/sqrt(d), ourmodel.py:21-27)
        ca = self.imagbased_cross_attention
        q = _conv1x1(text, st, ca.query_conv)
        k = _conv1x1(img, st, ca.key_conv)
        v = _conv1x1(img, st, ca.value_conv)
        pooled_2 = SqAttnFn.apply(q, k, v, B, T, 1.0)                                                   # (B, 768) fp32
        return Fm.linear_f32(Fm.add_f32(pooled_1, pooled_2), st, self.fc)


class ConvNeXtMoEClassifier(MdhsModule):
    """BASELINE config 4: ConvNeXt features -> mean over the 7x7 positions [-> concatenated with the BERT CLS vector]
    -> sparsely-gated MoE of KAN experts (ConNexT/config.yaml:60-74: `num_experts`, `k`, `layers_hidden` =
    [input, 512, 128, 32, classes]; pl_model_MOE2.py:36-53 for the encoder; moe.py:142,267-291 for the head).
    forward(batch dict) -> (logits, balance_loss); `use_text=False` is the image-only variant."""

    def __init__(self, num_labels=7, variant="tiny", use_text=True, num_experts=4, k=2, layers_hidden=None, pretrained=False,
                 bert_path="/data/QLI/BERT_pretain"):
        super().__init__()
        from .moe import MoE
        self.image_encoder = ConvNeXtEncoder(pretrained=pretrained, variant=variant)
        self.text_encoder = BertEncoder(model_path=bert_path) if use_text else None
        width = self.image_encoder.output_dim + (768 if use_text else 0)
        hidden = list(layers_hidden) if layers_hidden is not None else [width, 512, 128, 32, num_labels]
        self.moe = MoE(input_size=width, output_size=num_labels, num_experts=num_experts, hidden_size=hidden[1], k=k,
                       layers_hidden=hidden)

    def forward(self, batch_data, loss_coef=1e-2):
        images = batch_data["transformed_image"]
        self.store(images.device)
        B = images.shape[0]
        branch = None
        if self.text_encoder is not None and runtime.DUAL_STREAM and images.is_cuda:
            main = torch.cuda.current_stream()
            branch = runtime.fork_branch()     # BERT on a second stream, overlapping the ConvNeXt trunk
        tokens, h, w = self.image_encoder.forward_tokens(images)
        feat = Fm.mean_tokens(tokens, B, h * w)                                   # (B, C) fp32
        if self.text_encoder is not None:
            if branch is not None:
                with torch.cuda.stream(branch):
                    cls = self.text_encoder(batch_data["input_ids"], batch_data["attention_mask"]).contiguous()
                    cls = runtime.gate_branch_outputs(cls, main, branch)
                runtime.join_side(branch)
                runtime.record_on_current(cls)
            else:
                cls = self.text_encoder(batch_data["input_ids"], batch_data["attention_mask"]).contiguous()
            feat = torch.cat([Fm.to_f32(cls), feat], dim=1)                      # text first, as moe.py:297 orders it
        return self.moe(feat, loss_coef)
