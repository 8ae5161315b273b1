"""Batched test-time augmentation (scripts/predict.py:33-81, scripts/ablation_eval.py:35-74).

The reference calls the whole model once per variant (identity, hflip, vflip, rot90) and averages the logits.  Here the V
variants are produced by ONE kernel stacked on the batch axis, the image encoder runs once on V*B images, the text encoder
once on the B texts (it does not depend on the image variant), and fusion + head run per variant on contiguous slices.
Results equal the reference's `torch.stack(logits_list).mean(0)`.
"""
import torch

from . import functional as Fm
from . import ops


def _slice_tokens(tokens, v, B):
    if isinstance(tokens, dict):
        return {k: t[v * B:(v + 1) * B] for k, t in tokens.items()}
    return tokens[v * B:(v + 1) * B]


@torch.no_grad()
def predict_tta(model, images, input_ids, attention_mask, transforms=("hflip",), tabular_input=None):
    """Mean logits over [identity] + transforms for a MultimodalBaselineModel in eval mode.  Returns (B, C) fp32.
    `tabular_input` is forwarded like scripts/predict.py:62-69 does (`model(aug, ids, mask, tabular_input=tabular)`)."""
    model.eval()
    B = images.shape[0]
    V = 1 + len(transforms)
    side_branches = any(getattr(model, f, False) for f in ("gate_enabled", "tabular_enabled", "global_local_enabled",
                                                           "sequence_enabled"))
    if side_branches or images.dim() != 4:
        # gate (two feature sets per variant, model.py:257-281), tabular fusion, global/local crops and 5-D slice inputs
        # all live in the model's own forward: run it once per variant exactly like the reference loop does
        if images.dim() == 5:
            Bs, Ts = images.shape[:2]
            big = ops.tta_expand(images.reshape(Bs * Ts, *images.shape[2:]), transforms).view(V, Bs, Ts, *images.shape[2:])
            variants = [big[v] for v in range(V)]
        else:
            big = ops.tta_expand(images, transforms)
            variants = [big[v * B:(v + 1) * B] for v in range(V)]
        outs = [model(x, input_ids, attention_mask, tabular_input=tabular_input).float() for x in variants]
    else:
        big = ops.tta_expand(images, transforms)
        model.store(images.device)
        tokens = model.image_encoder(big)
        text = model.text_encoder(input_ids, attention_mask)
        outs = [model.classifier(model.fusion(_slice_tokens(tokens, v, B), text, attention_mask)).float() for v in range(V)]
    acc = torch.empty_like(outs[0])
    ops.axpby(outs[0].contiguous(), acc, a=1.0 / V, b=0.0)
    for o in outs[1:]:
        ops.axpby(o.contiguous(), acc, a=1.0 / V, b=1.0)
    return acc


@torch.no_grad()
def predict_tta_batchdict(model, images, input_ids, attention_mask, transforms=("hflip", "vflip", "rot90"), key=None):
    """Batched TTA for the batch-dict models (MIBF-Net `Resnet50WithOurs`, ConNexT `OurClassfierConvnextV2`; BASELINE config 5:
    "batched TTA inference").  The V variants go through the model as ONE batch of V*B images (text inputs tiled), logits are
    averaged per sample.  `key` selects the logit set of a dict-valued output (MIBF: "image_text")."""
    model.eval()
    B = images.shape[0]
    V = 1 + len(transforms)
    big = ops.tta_expand(images, transforms)
    out = model({"transformed_image": big, "input_ids": input_ids.repeat(V, 1), "attention_mask": attention_mask.repeat(V, 1)})
    if isinstance(out, dict):
        out = out[key or "image_text"]
    out = out.float().contiguous()
    acc = torch.empty((B, out.shape[1]), device=out.device, dtype=torch.float32)
    ops.axpby(out[:B].contiguous(), acc, a=1.0 / V, b=0.0)
    for v in range(1, V):
        ops.axpby(out[v * B:(v + 1) * B].contiguous(), acc, a=1.0 / V, b=1.0)
    return acc
