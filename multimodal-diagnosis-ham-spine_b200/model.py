"""MultimodalBaselineModel with the reference's constructor (same 38 keyword arguments read from
config.yml by scripts/train.py:179-211) and call surface (model.py:21-345): forward(), forward_features(),
classifier, freeze_encoders(), ablation modes, optional dual-expert gate.

Everything numeric runs on the sm_100a kernels; this file only wires modules together the way the
reference does.  The side branches of SURVEY.md section 8f-4 (tabular MLP + tabular fusion, LSTM / GRU /
Transformer multi-slice sequence encoder, global / local centre-crop branch) are built as well; the only
constructor values that raise are the ones whose arithmetic lives in un-vendored third-party packages
(`classifier_type="kan"` -> ikan GroupKAN, `fusion_type in {"mamba", "vmamba"}`): ImportError, exactly like the
reference without those packages.
"""
import torch
import torch.nn as nn

from . import functional as Fm
from . import runtime
from .encoder import ImageEncoder, MdhsModule, TextEncoder
from .modules.fusion_blocks import (BilinearFusionModule, ConcatFusionModule, FusionModule, HadamardFusionModule,
                                    HierarchicalFusionModule, MultiScaleFusionModule, SSMFusionModule, VMambaFusionModule,
                                    WeightedConcatFusionModule, _pool_image)
from .modules.gating import DualExpertGate
from .modules.heads import AttentionPoolingClassifier, MLPHead, ResidualClassifier, build_kan_head
from .modules.sequence_blocks import SequenceEncoder
from .modules.tabular import TabularEncoder, TabularFusion


class MultimodalBaselineModel(MdhsModule):
    def __init__(
        self,
        num_classes,
        image_feature_dim=512,
        text_feature_dim=768,
        hidden_dim=256,
        dropout=0.2,
        pretrained_image=True,
        image_weights_path="/home/medteam/.cache/torch/hub/checkpoints/resnet18-f37072fd.pth",
        text_model_name="bert-base-uncased",
        num_heads=8,
        image_backbone="resnet18",
        classifier_type="mlp",
        fusion_type="basic",
        text_pool="cls",
        kan_num_groups=8,
        kan_act_mode="gelu",
        tabular_enabled=False,
        tabular_input_dim=0,
        tabular_hidden_dim=128,
        tabular_dropout=0.1,
        gate_enabled=False,
        gate_hidden_dim=128,
        gate_use_entropy=True,
        gate_local_mode="image_only",
        gate_context_mode="full",
        sequence_enabled=False,
        sequence_type="lstm",
        sequence_hidden_dim=256,
        sequence_num_layers=1,
        sequence_bidirectional=True,
        sequence_dropout=0.1,
        sequence_num_heads=4,
        global_local_enabled=False,
        global_local_crop_ratio=0.6,
        global_local_combine="avg",
    ):
        super().__init__()
        fusion_dropout = min(dropout, 0.1)   # model.py:62-63
        head_dropout = min(dropout, 0.1)
        self.fusion_type = fusion_type
        self.tabular_enabled = tabular_enabled
        self.sequence_enabled = sequence_enabled
        self.global_local_enabled = global_local_enabled
        self.global_local_crop_ratio = global_local_crop_ratio
        self.global_local_combine = global_local_combine

        self.image_encoder = ImageEncoder(feature_dim=hidden_dim, pretrained=pretrained_image,
                                          weights_path=image_weights_path, backbone=image_backbone,
                                          multi_scale=(fusion_type in ("multiscale", "hierarchical")))
        if self.sequence_enabled:   # model.py:81-95
            self.sequence_encoder = SequenceEncoder(input_dim=hidden_dim, hidden_dim=sequence_hidden_dim, encoder_type=sequence_type,
                                                    num_layers=sequence_num_layers, bidirectional=sequence_bidirectional,
                                                    dropout=sequence_dropout, num_heads=sequence_num_heads)
            self.sequence_proj = nn.Linear(sequence_hidden_dim, hidden_dim) if sequence_hidden_dim != hidden_dim else nn.Identity()
        self.global_local_proj = None
        if global_local_enabled and global_local_combine == "concat":
            self.global_local_proj = nn.Linear(hidden_dim * 2, hidden_dim)   # model.py:97-99 (unused by the multiscale dict path)
        self.text_encoder = TextEncoder(model_path=text_model_name, feature_dim=text_feature_dim)

        if fusion_type == "hierarchical":
            # README.md:15 (image layer2/3/4 x BERT hidden 4/8/12 with adaptive weighting); opt-in, no reference code
            self.fusion = HierarchicalFusionModule(text_dim=text_feature_dim, hidden_dim=hidden_dim, num_heads=num_heads,
                                                   dropout=fusion_dropout)
        elif fusion_type == "multiscale":
            self.fusion = MultiScaleFusionModule(text_dim=text_feature_dim, hidden_dim=hidden_dim, num_heads=num_heads,
                                                 dropout=fusion_dropout)
        elif fusion_type == "hadamard":
            self.fusion = HadamardFusionModule(text_dim=text_feature_dim, hidden_dim=hidden_dim, text_pool=text_pool)
        elif fusion_type == "bilinear":
            self.fusion = BilinearFusionModule(text_dim=text_feature_dim, hidden_dim=hidden_dim, text_pool=text_pool)
        elif fusion_type == "mamba":
            self.fusion = SSMFusionModule(text_dim=text_feature_dim, hidden_dim=hidden_dim, text_pool=text_pool)
        elif fusion_type == "vmamba":
            self.fusion = VMambaFusionModule(text_dim=text_feature_dim, hidden_dim=hidden_dim, text_pool=text_pool)
        elif fusion_type == "weighted_concat":
            self.fusion = WeightedConcatFusionModule(text_dim=text_feature_dim, hidden_dim=hidden_dim, text_pool=text_pool)
        elif fusion_type == "concat":
            self.fusion = ConcatFusionModule(text_dim=text_feature_dim, hidden_dim=hidden_dim, text_pool=text_pool)
        else:
            self.fusion = FusionModule(text_dim=text_feature_dim, hidden_dim=hidden_dim, num_heads=num_heads,
                                       dropout=fusion_dropout)

        if self.tabular_enabled:   # model.py:155-167
            if tabular_input_dim <= 0:
                raise ValueError("tabular_input_dim must be > 0 when tabular is enabled.")
            self.tabular_encoder = TabularEncoder(tabular_input_dim, hidden_dim=tabular_hidden_dim, dropout=tabular_dropout)
            self.tabular_fusion = TabularFusion(hidden_dim + tabular_hidden_dim, hidden_dim, head_dropout)

        self.gate_enabled = gate_enabled
        self.gate_local_mode = gate_local_mode
        self.gate_context_mode = gate_context_mode
        if self.gate_enabled:
            self.gate = DualExpertGate(lesion_dim=hidden_dim, context_dim=hidden_dim, hidden_dim=gate_hidden_dim,
                                       use_entropy=gate_use_entropy)

        self.classifier_type = classifier_type
        if classifier_type == "kan1":
            # KAN head variant of BASELINE config 5: the vendored efficient-KAN of ConNexT/models/block/kan1.py
            # (KAN1([hidden, 256, C])) in place of the un-vendored GroupKAN of classifier_type="kan"
            from .connext.kan1 import KAN1Head
            self.classifier = KAN1Head([hidden_dim, 256, num_classes])
        elif classifier_type == "kan":
            self.classifier = build_kan_head(hidden_dim=hidden_dim, num_classes=num_classes, dropout=head_dropout,
                                             num_groups=kan_num_groups, act_mode=kan_act_mode)
        elif classifier_type == "residual":
            self.classifier = ResidualClassifier(hidden_dim, hidden_dim, num_classes, head_dropout)
        elif classifier_type == "attention_pooling":
            self.classifier = AttentionPoolingClassifier(hidden_dim, hidden_dim, num_classes, num_heads, head_dropout)
        else:
            self.classifier = MLPHead(hidden_dim, num_classes, head_dropout)

    # ------------------------------------------------------------------ reference call surface
    def forward_features(self, image_input, text_input_ids, text_attention_mask, tabular_input=None, ablation_mode=None):
        self.store(image_input.device)
        branch = None
        if runtime.DUAL_STREAM and ablation_mode != "image_only" and image_input.is_cuda:
            main = torch.cuda.current_stream()
            branch = runtime.fork_branch()     # forked BEFORE the image encoder is enqueued: the two encoders overlap
        image_tokens, pooled_image = self._encode_image_tokens(image_input, want_pooled=(ablation_mode == "image_only"))
        text_tokens = None
        if branch is not None:
            # enqueued second so that autograd (latest node first) back-propagates the text encoder FIRST, on its stream:
            # its gradient bucket can leave while the trunk's backward is still being enqueued (train.py)
            with torch.cuda.stream(branch):
                text_tokens = self._encode_text(text_input_ids, text_attention_mask)
                text_tokens = runtime.gate_branch_outputs(text_tokens, main, branch)
            runtime.join_side(branch)
            runtime.record_on_current(text_tokens)
        return self._features_from_tokens(image_tokens, pooled_image, text_input_ids, text_attention_mask, tabular_input,
                                          ablation_mode, text_tokens)

    def _features_from_tokens(self, image_tokens, pooled_image, text_input_ids, text_attention_mask, tabular_input,
                              ablation_mode, text_tokens=None):
        """model.py:212-237 after the image encoder: ablation switch, text encoder, fusion, tabular fusion."""
        if ablation_mode == "image_only":
            return pooled_image
        if text_tokens is None:
            text_tokens = self._encode_text(text_input_ids, text_attention_mask)
        if ablation_mode == "text_off":
            text_tokens = ({k: torch.zeros_like(v) for k, v in text_tokens.items()} if isinstance(text_tokens, dict)
                           else torch.zeros_like(text_tokens))
        if self.sequence_enabled and self.fusion_type in ("multiscale", "hierarchical") and not isinstance(image_tokens, dict):
            image_tokens = {"layer2": image_tokens, "layer3": image_tokens, "layer4": image_tokens}   # model.py:220-225
        fused = self.fusion(image_tokens, text_tokens, text_attention_mask)
        if self.tabular_enabled:   # model.py:229-235
            if tabular_input is None:
                raise ValueError("tabular_input is required when tabular is enabled.")
            tab = self.tabular_encoder(tabular_input)
            fused32 = Fm.to_f32(fused) if fused.dtype == torch.bfloat16 else fused.float()
            fused = self.tabular_fusion(torch.cat([fused32, tab], dim=1))
        return fused

    def forward(self, image_input, text_input_ids, text_attention_mask, tabular_input=None, ablation_mode=None):
        if ablation_mode is not None or not self.gate_enabled:
            fused = self.forward_features(image_input, text_input_ids, text_attention_mask, tabular_input=tabular_input,
                                          ablation_mode=ablation_mode)
            return self.classifier(fused)
        context_mode = None if self.gate_context_mode == "full" else self.gate_context_mode
        if self.training:
            # the reference runs the whole feature path twice (model.py:257-271); in train mode that is observable
            # (BatchNorm running statistics see two momentum updates), so it is kept
            context_feat = self.forward_features(image_input, text_input_ids, text_attention_mask, tabular_input=tabular_input,
                                                 ablation_mode=context_mode)
            local_feat = self.forward_features(image_input, text_input_ids, text_attention_mask, tabular_input=tabular_input,
                                               ablation_mode=self.gate_local_mode)
        else:
            # inference: ONE image-encoder pass and at most ONE text-encoder pass feed both experts (same numbers: eval-mode
            # encoders are deterministic functions of the input)
            self.store(image_input.device)
            modes = (context_mode, self.gate_local_mode)
            image_tokens, pooled = self._encode_image_tokens(image_input, want_pooled=("image_only" in modes))
            text_tokens = None
            if any(m != "image_only" for m in modes):
                text_tokens = self._encode_text(text_input_ids, text_attention_mask)
            context_feat = self._features_from_tokens(image_tokens, pooled, text_input_ids, text_attention_mask, tabular_input,
                                                      context_mode, text_tokens)
            local_feat = self._features_from_tokens(image_tokens, pooled, text_input_ids, text_attention_mask, tabular_input,
                                                    self.gate_local_mode, text_tokens)
        logits_context = self.classifier(context_feat)
        logits_local = self.classifier(local_feat)
        entropy = None
        if self.gate.use_entropy:
            probs = torch.softmax(logits_local, dim=1)
            entropy = -(probs * (probs + 1e-8).log()).sum(dim=1, keepdim=True)
        alpha = self.gate(local_feat, context_feat, entropy)
        return alpha * logits_local + (1 - alpha) * logits_context

    def _encode_text(self, text_input_ids, text_attention_mask):
        if self.fusion_type == "hierarchical":
            return self.text_encoder(text_input_ids, text_attention_mask,
                                     hidden_states=tuple(l for _, l in HierarchicalFusionModule.LEVELS))
        return self.text_encoder(text_input_ids, text_attention_mask)

    def _pool_image_tokens(self, image_tokens):
        return _pool_image(image_tokens)

    def _encode_image_tokens(self, image_input, want_pooled=True):
        if image_input.dim() == 5:   # model.py:317-331: (B, T, 3, H, W) slices -> pooled per slice -> sequence encoder
            if not self.sequence_enabled:
                raise ValueError("Sequence input provided but sequence encoder is disabled.")
            Bs, Ts = image_input.shape[0], image_input.shape[1]
            flat = image_input.reshape(Bs * Ts, *image_input.shape[2:])
            tokens = self._global_local_tokens(flat) if self.global_local_enabled else self.image_encoder(flat)
            pooled = self._pool_image_tokens(tokens)
            pooled = Fm.to_f32(pooled) if pooled.dtype == torch.bfloat16 else pooled.float()
            seq = self.sequence_encoder(pooled.view(Bs, Ts, -1))
            if not isinstance(self.sequence_proj, nn.Identity):
                seq = Fm.linear_f32(seq, self.store(image_input.device), self.sequence_proj)
            seq_tokens = Fm.to_bf16(seq).view(Bs, 1, -1)
            return seq_tokens, seq
        if self.global_local_enabled:
            tokens = self._global_local_tokens(image_input)
        else:
            tokens = self.image_encoder(image_input)
        pooled = self._pool_image_tokens(tokens) if want_pooled else None
        return tokens, pooled

    def _global_local_tokens(self, image_input):
        """model.py:292-315,337-341: encode the image and its centre crop (resized back, bilinear) and combine the tokens.
        Eval mode runs ONE encoder pass over the 2B stacked views; train mode keeps the reference's two passes because
        BatchNorm batch statistics are per pass."""
        from . import ops
        B = image_input.shape[0]
        both = ops.global_local(image_input, self.global_local_crop_ratio)
        if self.training:
            g, l = self.image_encoder(both[:B]), self.image_encoder(both[B:])
        else:
            t = self.image_encoder(both)
            if isinstance(t, dict):
                g, l = {k: v[:B] for k, v in t.items()}, {k: v[B:] for k, v in t.items()}
            else:
                g, l = t[:B], t[B:]
        if isinstance(g, dict):
            return {k: self._avg_tokens(g[k], l[k]) for k in g}
        if self.global_local_combine == "concat":
            Bt, T, Hd = g.shape
            cat = torch.cat([g, l], dim=-1).reshape(Bt * T, 2 * Hd)
            st = self.store(image_input.device)
            return Fm.linear(cat, st, self.global_local_proj.weight, self.global_local_proj.bias).view(Bt, T, Hd)
        return self._avg_tokens(g, l)

    @staticmethod
    def _avg_tokens(a, b):
        shp = a.shape
        return Fm.avg_bf16(a.reshape(-1, shp[-1]), b.reshape(-1, shp[-1])).view(shp)

    def freeze_encoders(self):
        for param in self.image_encoder.parameters():
            param.requires_grad = False
        for param in self.text_encoder.parameters():
            param.requires_grad = False
