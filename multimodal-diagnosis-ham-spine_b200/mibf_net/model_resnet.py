"""MIBF-Net (mibf_net/model_resnet.py:10-94): ResNet-50 (fc -> 768) + BERT CLS -> two IBFA cross-attentions ->
three classification heads, MP-Loss.  Same constructor / forward(batch dict) / cal_loss surface and state_dict
keys (`image_encoder.*` is a torchvision resnet50 container, `text_encoder.bert.*` a BertModel container)."""
import torch
import torch.nn as nn
from torchvision import models

from .. import functional as Fm
from .. import ops
from .. import runtime
from ..encoder import MdhsModule, _TrunkFn
from ..resnet_engine import ResNetEngine
from .attention import MultiHeadCrossAttention_v2, SelfAttention
from .bert import BertEncoder


class _MpLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, zi, zt, zf, labels):
        loss, grads = ops.mp_loss(zi.contiguous(), zt.contiguous(), zf.contiguous(), labels, want_grad=True)
        ctx.save_for_backward(*grads)
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        gi, gt, gf = ctx.saved_tensors
        s = dloss.contiguous().view(1)
        outs = []
        for g in (gi, gt, gf):
            o = torch.empty_like(g)
            ops.axpby(g, o, a=1.0, b=0.0, a_dev=s)
            outs.append(o)
        return outs[0], outs[1], outs[2], None


class Resnet50WithOurs(MdhsModule):
    def __init__(self, num_labels=6, loss_class="KL_loss", bert_path="/data/QLI/BERT_pretain", pretrained=True):
        super().__init__()
        self.text_encoder = BertEncoder(model_path=bert_path)
        backbone = models.resnet50(weights="DEFAULT" if pretrained else None)   # reference: models.resnet50(pretrained=True)
        backbone.fc = nn.Linear(backbone.fc.in_features, 768)
        self.image_encoder = backbone
        self.textbased_cross_attention = MultiHeadCrossAttention_v2(dim=768, num_heads=1)
        self.imagbased_cross_attention = MultiHeadCrossAttention_v2(dim=768, num_heads=1)
        self.I2Iattention = SelfAttention(input_dim=768)
        self.fc = nn.Linear(768 * 2, num_labels)
        self.fc_image = self._build_mlp(768, num_labels)
        self.fc_text = self._build_mlp(768, num_labels)
        self.loss_class = loss_class
        self.loss = nn.CrossEntropyLoss()
        object.__setattr__(self, "_trunk", None)

    def _build_mlp(self, input_dim, num_labels):
        return nn.Sequential(nn.Flatten(start_dim=1), nn.Linear(input_dim, 512), nn.ReLU(), nn.Linear(512, num_labels))

    def _on_bind(self, store):
        object.__setattr__(self, "_trunk", ResNetEngine(store, self.image_encoder))

    def _mlp(self, st, seq, x32):
        return Fm.linear_f32(Fm.linear_f32(x32, st, seq[1], act=ops.ACT_RELU), st, seq[3])

    def forward(self, batch_data):
        images = batch_data["transformed_image"]
        st = self.store(images.device)
        B = images.shape[0]
        branch = None
        if runtime.DUAL_STREAM and images.is_cuda:     # BERT on a second stream, overlapping the ResNet trunk
            main = torch.cuda.current_stream()
            branch = runtime.fork_branch()
        else:
            text = self.text_encoder(batch_data["input_ids"], batch_data["attention_mask"])      # (B, 768) bf16 view
        trainable = any(p.requires_grad for p in self.image_encoder.parameters())
        need = trainable and torch.is_grad_enabled()
        (f4,) = _TrunkFn.apply(st.anchor, images.float(), self._trunk, self.training, ("layer4",), need)
        if branch is not None:
            with torch.cuda.stream(branch):
                text = self.text_encoder(batch_data["input_ids"], batch_data["attention_mask"])
                text = runtime.gate_branch_outputs(text, main, branch)
            runtime.join_side(branch)
            runtime.record_on_current(text)
        pooled = Fm.mean_tokens(f4, B, f4.shape[0] // B)                                          # avg-pool -> (B, 2048) fp32
        fc = self.image_encoder.fc
        image = Fm.linear(Fm.to_bf16(pooled), st, fc.weight, fc.bias)                             # (B, 768) bf16
        t3 = text.reshape(B, 1, -1)
        i3 = image.reshape(B, 1, -1)
        text_fused = self.textbased_cross_attention(i3, t3)
        imag_fused = self.imagbased_cross_attention(t3, i3)
        tf32 = Fm.to_f32(text_fused.reshape(B, -1))
        if32 = Fm.to_f32(imag_fused.reshape(B, -1))
        return {
            "image_text": Fm.linear_f32(torch.cat([tf32, if32], dim=1), st, self.fc),
            "text": self._mlp(st, self.fc_text, tf32),
            "image": self._mlp(st, self.fc_image, if32),
        }

    def cal_loss(self, output, labels):
        if self.loss_class == "textimage_loss":
            return Fm.cross_entropy(output["image_text"], labels)
        if self.loss_class == "text_image_textimage_loss":
            return (Fm.cross_entropy(output["image"], labels) + Fm.cross_entropy(output["text"], labels)
                    + Fm.cross_entropy(output["image_text"], labels))
        return self.compute_kl_loss(output, labels)

    def compute_kl_loss(self, output, labels):
        return _MpLossFn.apply(output["image"].float(), output["text"].float(), output["image_text"].float(), labels)
