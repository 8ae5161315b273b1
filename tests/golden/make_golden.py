"""Generates tests/golden/reference_outputs.pt by running the REAL reference (imported from /root/reference) on
the deterministic synthetic weights / inputs of oracle/weights.py.  Dev container only; the fixture travels.

    python tests/golden/make_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import weights  # noqa: E402
from refutil import REF_ROOT, bert_dir, build_reference_model, quiet  # noqa: E402

CASES = [  # (name, fusion, head, gate)
    ("basic_mlp", "basic", "mlp", False),
    ("multiscale_residual", "multiscale", "residual", False),
    ("concat_attnpool", "concat", "attention_pooling", False),
    ("weighted_gate", "weighted_concat", "mlp", True),
    ("hadamard_mlp", "hadamard", "mlp", False),
    ("bilinear_mlp", "bilinear", "mlp", False),
]


def main():
    torch.set_num_threads(8)
    out = {"meta": {"torch": torch.__version__, "batch": 4, "seq": 16, "hw": 64, "classes": 7,
                    "weights": "oracle.weights.synth_state_dict(seed=1)", "inputs": "oracle.weights.synthetic_batch(4,16,7,image_hw=64)"}}
    images, ids, mask, labels = weights.synthetic_batch(4, 16, 7, image_hw=64)
    for name, fusion, head, gate in CASES:
        ref = build_reference_model(fusion=fusion, head=head, gate=gate).eval()
        sd = weights.synth_state_dict(ref.state_dict(), seed=1)
        ref.load_state_dict(sd)
        with torch.no_grad():
            logits = ref(images, ids, mask)
        out[name] = {"fusion": fusion, "head": head, "gate": gate, "eval_logits": logits.clone()}
        print(name, logits[0, :3])
    # one training step (train-mode BN, dropout off) of the headline configuration: loss + a few gradients
    ref = build_reference_model(fusion="basic", head="mlp").train()
    for m in ref.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    ref.text_encoder.model.eval()
    sd = weights.synth_state_dict(ref.state_dict(), seed=1)
    ref.load_state_dict(sd)
    im, ii, mm, ll = weights.synthetic_batch(8, 16, 7, image_hw=128)
    feats = ref.forward_features(im, ii, mm)
    logits = ref.classifier(feats)
    loss = torch.nn.functional.cross_entropy(logits, ll, label_smoothing=0.02)
    loss.backward()
    named = dict(ref.named_parameters())
    keep = ["classifier.3.weight", "classifier.3.bias", "classifier.0.bias", "fusion.transformer_block.norm3.weight",
            "fusion.transformer_block.ff.3.bias", "image_encoder.proj4.bias", "text_encoder.model.encoder.layer.11.output.dense.bias",
            "text_encoder.model.encoder.layer.0.attention.output.LayerNorm.weight", "image_encoder.model.layer4.2.bn3.weight"]
    out["train_basic_mlp"] = {"batch": 8, "seq": 16, "hw": 128, "logits": logits.detach().clone(), "loss": loss.detach().clone(),
                              "grads": {k: named[k].grad.clone() for k in keep}}
    # MIBF-Net (eval) + MP-loss, KAN, MoE
    import torchvision
    import tempfile
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "hub", "checkpoints"))
    torch.save(torchvision.models.resnet50(weights=None).state_dict(), os.path.join(tmp, "hub", "checkpoints", "resnet50-0676ba61.pth"))
    torch.hub.set_dir(os.path.join(tmp, "hub"))
    import warnings
    with quiet(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from mibf_net.model_resnet import Resnet50WithOurs
        mibf = Resnet50WithOurs(num_labels=6, bert_path=bert_dir()).eval()
    sd = weights.synth_state_dict(mibf.state_dict(), seed=3)
    mibf.load_state_dict(sd)
    im, ii, mm, ll = weights.synthetic_batch(4, 16, 6, image_hw=64, unit_range=True)
    with torch.no_grad():
        o = mibf({"transformed_image": im, "input_ids": ii, "attention_mask": mm})
        out["mibf"] = {k: v.clone() for k, v in o.items()}
        out["mibf"]["mp_loss"] = mibf.cal_loss(o, ll).clone()
    sys.path.insert(0, os.path.join(REF_ROOT, "ConNexT"))
    from models.block.kan1 import KAN1
    from models.block.moe import MoE
    kan = KAN1([64, 32, 7]).eval()
    sdk = weights.synth_state_dict(kan.state_dict(), seed=4)
    kan.load_state_dict(sdk)
    g = torch.Generator().manual_seed(77)
    x = torch.randn(9, 64, generator=g) * 1.5
    x[0, :4] = torch.tensor([-2.2, 2.2, 3.0, -5.0])
    moe = MoE(input_size=64, output_size=7, num_experts=4, hidden_size=32, k=2, layers_hidden=[64, 32, 7]).eval()
    sdm = weights.synth_state_dict(moe.state_dict(), seed=5)
    moe.load_state_dict(sdm)
    with torch.no_grad():
        y, l = moe(x)
        out["kan_moe"] = {"x": x, "kan_out": kan(x).clone(), "moe_out": y.clone(), "moe_loss": l.clone(),
                          "kan_grid": sdk["layers.0.grid"].clone()}
    # MoE in training mode: noisy gating with the Gaussian noise injected (torch.randn_like patched), loss + gradients
    moe_t = MoE(input_size=64, output_size=7, num_experts=4, hidden_size=32, k=2, layers_hidden=[64, 32, 7]).train()
    sdt = weights.synth_state_dict(moe_t.state_dict(), seed=6)
    moe_t.load_state_dict(sdt)
    g2 = torch.Generator().manual_seed(5)
    xt = torch.randn(12, 64, generator=g2)
    noise = torch.randn(12, 4, generator=g2)
    orig_randn_like = torch.randn_like
    torch.randn_like = lambda t: noise.to(t.dtype)
    try:
        xr = xt.clone().requires_grad_(True)
        yt, lt = moe_t(xr)
        (yt.square().sum() + lt.sum()).backward()
    finally:
        torch.randn_like = orig_randn_like
    named = dict(moe_t.named_parameters())
    keep = ["w_gate", "w_noise", "experts.1.layers.0.spline_weight", "experts.2.layers.1.base_weight",
            "experts.0.layers.0.spline_scaler", "experts.3.layers.0.base_weight"]
    out["moe_train"] = {"x": xt, "noise": noise, "y": yt.detach().clone(), "loss": lt.detach().clone(), "dx": xr.grad.clone(),
                        "grads": {k: named[k].grad.clone() for k in keep}, "mean": sdt["mean"].clone(), "std": sdt["std"].clone()}
    path = os.path.join(HERE, "reference_outputs.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
