"""Generates tests/golden/reference_connext.pt: outputs of the REAL reference's ConNexT classifier
(ConNexT/models/ourmodel.py, torchvision ConvNeXt branch; Tiny variant to keep the fixture small) on the deterministic
synthetic weights / inputs of oracle/weights.py.  Dev container only; the fixture travels.

    python tests/golden/make_golden_connext.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import weights  # noqa: E402
from refutil import build_reference_connext  # noqa: E402

KEEP = ["fc.weight", "fc.bias", "conv.weight", "conv.bias", "imagbased_cross_attention.query_conv.weight",
        "imagbased_cross_attention.key_conv.bias", "imagbased_cross_attention.key_conv.weight", "imagbased_cross_attention.value_conv.weight",
        "textbased_cross_attention.value_conv.weight", "textbased_cross_attention.query_conv.weight",
        "image_encoder.7.2.layer_scale", "image_encoder.7.0.block.0.weight", "image_encoder.7.0.block.0.bias",
        "image_encoder.7.1.block.3.weight", "image_encoder.6.1.weight", "image_encoder.6.0.weight",
        "image_encoder.5.4.block.5.bias", "image_encoder.5.8.block.2.weight", "image_encoder.3.1.block.0.weight",
        "image_encoder.1.0.layer_scale", "image_encoder.1.2.block.0.weight", "image_encoder.0.0.weight", "image_encoder.0.1.bias",
        "text_encoder.bert.encoder.layer.11.output.dense.bias", "text_encoder.bert.embeddings.LayerNorm.weight"]


def sample(t, n=4096):
    """Deterministic sub-sample of a large tensor (flattened, fixed stride): keeps the fixture small."""
    f = t.detach().flatten()
    return f.clone() if f.numel() <= n else f[::f.numel() // n].clone()


def main():
    torch.set_num_threads(8)
    out = {"meta": {"torch": torch.__version__, "variant": "tiny", "weights": "oracle.weights.synth_state_dict(seed=5)"}}
    ref = build_reference_connext("tiny", num_labels=7).eval()
    sd = weights.synth_state_dict(ref.state_dict(), seed=5)
    ref.load_state_dict(sd)
    im, ii, mm, ll = weights.synthetic_batch(4, 16, 7, image_hw=64, unit_range=True)
    with torch.no_grad():
        out["eval"] = {"batch": 4, "seq": 16, "hw": 64, "logits": ref({"transformed_image": im, "input_ids": ii, "attention_mask": mm}).clone(),
                       "features": ref.image_encoder(im).clone()}
    # one training step: stochastic depth off (its mask comes from torch's RNG), BERT dropout off
    ref.train()
    ref.text_encoder.bert.eval()
    for m in ref.modules():
        if type(m).__name__ == "StochasticDepth":
            m.p = 0.0
    im, ii, mm, ll = weights.synthetic_batch(4, 16, 7, image_hw=96, unit_range=True)
    logits = ref({"transformed_image": im, "input_ids": ii, "attention_mask": mm})
    loss = torch.nn.functional.cross_entropy(logits, ll)
    loss.backward()
    named = dict(ref.named_parameters())
    out["train"] = {"batch": 4, "seq": 16, "hw": 96, "logits": logits.detach().clone(), "loss": loss.detach().clone(),
                    "grads": {k: sample(named[k].grad) for k in KEEP}}
    path = os.path.join(HERE, "reference_connext.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
