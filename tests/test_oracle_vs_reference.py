"""Pins the oracle (oracle/port.py) against the REAL reference modules imported from /root/reference.
Runs only where the reference checkout exists (the dev container); the GPU box relies on tests/golden/.
Tolerance: fp32 CPU, identical weights and inputs -> max-norm relative error <= 1e-5 (observed ~1e-6)."""
import os
import sys

import pytest
import torch

from refutil import REF_ROOT, build_reference_connext, build_reference_model, have_reference, quiet

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port, weights  # noqa: E402

pytestmark = pytest.mark.skipif(not have_reference(), reason="reference checkout not present (GPU box)")


def rel(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


@pytest.mark.parametrize("fusion,head,gate", [("basic", "mlp", False), ("multiscale", "residual", False),
                                              ("concat", "attention_pooling", False), ("weighted_concat", "mlp", True),
                                              ("hadamard", "mlp", False), ("bilinear", "mlp", False)])
def test_model_forward_matches_reference(fusion, head, gate):
    ref = build_reference_model(fusion=fusion, head=head, gate=gate).eval()
    sd = weights.synth_state_dict(ref.state_dict(), seed=1)
    ref.load_state_dict(sd)
    images, ids, mask, _ = weights.synthetic_batch(2, 16, 7, image_hw=64)
    with torch.no_grad():
        want = ref(images, ids, mask)
        got = port.model_forward(sd, images, ids, mask, fusion=fusion, head=head, gate=gate)
    assert rel(got, want) < 1e-5


def test_train_mode_bn_and_grads_match_reference():
    ref = build_reference_model(fusion="concat", head="mlp").train()
    for m in ref.modules():  # dropout off, BatchNorm in train mode
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    ref.text_encoder.model.eval()
    sd = weights.synth_state_dict(ref.state_dict(), seed=2)
    ref.load_state_dict(sd)
    images, ids, mask, labels = weights.synthetic_batch(4, 8, 7, image_hw=64)
    loss_ref = torch.nn.functional.cross_entropy(ref(images, ids, mask), labels, label_smoothing=0.02)
    loss_ref.backward()
    sd_g = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    loss = port.ce_label_smoothing(port.model_forward(sd_g, images, ids, mask, fusion="concat", training_bn=True), labels)
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-5
    named = dict(ref.named_parameters())
    for key in ("classifier.3.weight", "fusion.proj.weight", "text_encoder.model.encoder.layer.5.intermediate.dense.weight",
                "image_encoder.model.layer3.2.conv2.weight", "image_encoder.model.bn1.weight"):
        g_ref, g = named[key].grad, sd_g[key].grad
        assert rel(g, g_ref) < 1e-3, key  # conv grads are ill-conditioned even between fp32 runs


def test_losses_match_reference():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    torch.manual_seed(0)
    logits = torch.randn(16, 7)
    labels = torch.randint(0, 7, (16,))
    feats = torch.randn(16, 32)
    import importlib.util
    # scripts/train.py pulls in tensorboard / data loaders; load only the two loss classes from its source text
    src = open(os.path.join(REF_ROOT, "scripts", "train.py")).read()
    start, end = src.index("class SupConLoss"), src.index("def _compute_class_weights")
    ns = {}
    exec("import torch\nimport torch.nn as nn\nimport torch.nn.functional as F\n" + src[start:end], ns)  # reference code, run as-is
    assert abs(ns["SupConLoss"]()(feats, labels).item() - port.supcon_loss(feats, labels).item()) < 1e-5
    assert abs(ns["FocalLoss"]()(logits, labels).item() - port.focal_loss(logits, labels).item()) < 1e-6


def test_mibf_matches_reference(tmp_path, monkeypatch):
    import torchvision
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from refutil import bert_dir
    # models.resnet50(pretrained=True) is satisfied offline by a seeded checkpoint in TORCH_HOME
    hub = tmp_path / "hub" / "checkpoints"
    hub.mkdir(parents=True)
    torch.manual_seed(0)
    torch.save(torchvision.models.resnet50(weights=None).state_dict(), hub / "resnet50-0676ba61.pth")
    monkeypatch.setenv("TORCH_HOME", str(tmp_path))
    torch.hub.set_dir(str(tmp_path / "hub"))
    with quiet():
        from mibf_net.model_resnet import Resnet50WithOurs
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = Resnet50WithOurs(num_labels=6, bert_path=bert_dir()).eval()
    sd = weights.synth_state_dict(ref.state_dict(), seed=3)
    ref.load_state_dict(sd)
    images, ids, mask, labels = weights.synthetic_batch(2, 16, 6, image_hw=64, unit_range=True)
    batch = {"transformed_image": images, "input_ids": ids, "attention_mask": mask}
    with torch.no_grad():
        want = ref(batch)
        got = port.mibf_forward(sd, images, ids, mask)
    for k in ("image_text", "text", "image"):
        assert rel(got[k], want[k]) < 1e-5, k
    l_ref = ref.cal_loss(want, labels).item()
    assert abs(l_ref - port.mp_loss(got["image"], got["text"], got["image_text"], labels).item()) < 1e-5 * max(1.0, abs(l_ref))


def test_kan_and_moe_match_reference():
    sys.path.insert(0, os.path.join(REF_ROOT, "ConNexT"))
    from models.block.kan1 import KAN1
    from models.block.moe import MoE
    torch.manual_seed(0)
    kan = KAN1([64, 32, 7]).eval()
    sd = weights.synth_state_dict(kan.state_dict(), seed=4)
    kan.load_state_dict(sd)
    x = torch.randn(9, 64) * 1.5  # includes values outside the [-2.2, 2.2) knot range
    x[0, :4] = torch.tensor([-2.2, 2.2, 3.0, -5.0])
    with torch.no_grad():
        assert rel(port.kan_net(sd, "", x), kan(x)) < 1e-5
    moe = MoE(input_size=64, output_size=7, num_experts=4, hidden_size=32, k=2, layers_hidden=[64, 32, 7]).eval()
    sd = weights.synth_state_dict(moe.state_dict(), seed=5)
    moe.load_state_dict(sd)
    with torch.no_grad():
        y_ref, l_ref = moe(x)
        y, l = port.moe_forward_eval(sd, "", x, 4, 2)
    assert rel(y, y_ref) < 1e-5
    assert abs(l.item() - l_ref.item()) < 1e-6


def test_moe_training_mode_matches_reference(monkeypatch):
    """Noisy top-k gating + smooth load estimator + gradients, with the Gaussian noise injected on both sides."""
    sys.path.insert(0, os.path.join(REF_ROOT, "ConNexT"))
    from models.block import moe as ref_moe
    torch.manual_seed(0)
    moe = ref_moe.MoE(input_size=64, output_size=7, num_experts=4, hidden_size=32, k=2, layers_hidden=[64, 32, 7]).train()
    sd = weights.synth_state_dict(moe.state_dict(), seed=6)
    moe.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(12, 64, generator=g)
    noise = torch.randn(12, 4, generator=g)
    monkeypatch.setattr(torch, "randn_like", lambda t: noise.to(t.dtype))
    xr = x.clone().requires_grad_(True)
    y_ref, l_ref = moe(xr)
    (y_ref.square().sum() + l_ref.sum()).backward()
    sd_g = {k: v.clone().requires_grad_(v.is_floating_point() and k not in ("mean", "std")) for k, v in sd.items()}
    xo = x.clone().requires_grad_(True)
    y, l = port.moe_forward_train(sd_g, "", xo, noise, 4, 2)
    (y.square().sum() + l).backward()
    assert rel(y, y_ref) < 1e-5
    assert abs(l.item() - l_ref.item()) < 1e-6
    assert rel(xo.grad, xr.grad) < 1e-4
    named = dict(moe.named_parameters())
    for key in ("w_gate", "w_noise", "experts.1.layers.0.spline_weight", "experts.2.layers.1.base_weight",
                "experts.0.layers.0.spline_scaler"):
        assert rel(sd_g[key].grad, named[key].grad) < 1e-4, key


def test_connext_classifier_matches_reference():
    ref = build_reference_connext("tiny").eval()
    sd = weights.synth_state_dict(ref.state_dict(), seed=5)
    ref.load_state_dict(sd)
    images, ids, mask, _ = weights.synthetic_batch(2, 16, 7, image_hw=64, unit_range=True)
    with torch.no_grad():
        want = ref({"transformed_image": images, "input_ids": ids, "attention_mask": mask})
        got = port.connext_forward(sd, images, ids, mask)
        feat_ref = ref.image_encoder(images)
        feat = port.convnext_features(sd, "image_encoder.", images)
    assert rel(feat, feat_ref) < 1e-5
    assert rel(got, want) < 1e-5


def test_tabular_branch_matches_reference():
    ref = build_reference_model(fusion="concat", head="mlp", tabular_enabled=True, tabular_input_dim=12).eval()
    sd = weights.synth_state_dict(ref.state_dict(), seed=7)
    ref.load_state_dict(sd)
    images, ids, mask, _ = weights.synthetic_batch(3, 16, 7, image_hw=64)
    tab = torch.randn(3, 12, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        want = ref(images, ids, mask, tabular_input=tab)
        got = port.model_forward(sd, images, ids, mask, fusion="concat", head="mlp", tabular=tab)
    assert rel(got, want) < 1e-5


@pytest.mark.parametrize("fusion,combine", [("basic", "avg"), ("concat", "concat"), ("multiscale", "avg")])
def test_global_local_branch_matches_reference(fusion, combine):
    ref = build_reference_model(fusion=fusion, head="mlp", global_local_enabled=True, global_local_crop_ratio=0.6,
                                global_local_combine=combine).eval()
    sd = weights.synth_state_dict(ref.state_dict(), seed=8)
    ref.load_state_dict(sd)
    images, ids, mask, _ = weights.synthetic_batch(2, 16, 7, image_hw=64)
    with torch.no_grad():
        want = ref(images, ids, mask)
        got = port.model_forward(sd, images, ids, mask, fusion=fusion, head="mlp", global_local=0.6, global_local_combine=combine)
    assert rel(got, want) < 1e-5


@pytest.mark.parametrize("fusion,layers,bidir,hid,kind", [("concat", 1, True, 256, "lstm"), ("basic", 2, False, 128, "lstm"),
                                                          ("multiscale", 1, True, 256, "lstm"), ("concat", 2, True, 128, "gru"),
                                                          ("concat", 2, True, 128, "transformer")])
def test_sequence_lstm_branch_matches_reference(fusion, layers, bidir, hid, kind):
    ref = build_reference_model(fusion=fusion, head="mlp", sequence_enabled=True, sequence_type=kind, sequence_hidden_dim=hid,
                                sequence_num_layers=layers, sequence_bidirectional=bidir, sequence_dropout=0.0).eval()
    sd = weights.synth_state_dict(ref.state_dict(), seed=11)
    ref.load_state_dict(sd)
    images, ids, mask, _ = weights.synthetic_batch(2 * 3, 16, 7, image_hw=64)
    images = images.view(2, 3, 3, 64, 64)
    with torch.no_grad():
        want = ref(images, ids[:2], mask[:2])
        got = port.model_forward(sd, images, ids[:2], mask[:2], fusion=fusion, head="mlp")
    assert rel(got, want) < 1e-5


def test_bert_hidden_states_4_8_12_match_hf():
    """The hierarchical variant (README.md:15) consumes BERT hidden states 4 / 8 / 12: the oracle's truncated forward
    (num_layers = n) equals transformers' `output_hidden_states=True` tuple at index n."""
    from transformers import BertModel
    from refutil import bert_dir
    with quiet():
        bert = BertModel.from_pretrained(bert_dir()).eval()
    sd = {"text_encoder.model." + k: v for k, v in weights.synth_state_dict(bert.state_dict(), seed=3).items()}
    bert.load_state_dict({k[len("text_encoder.model."):]: v for k, v in sd.items()})
    _, ids, mask, _ = weights.synthetic_batch(3, 12, 7, image_hw=32)
    with torch.no_grad():
        hs = bert(input_ids=ids, attention_mask=mask, output_hidden_states=True).hidden_states
        for n in (4, 8, 12):
            got = port.bert_last_hidden(sd, "text_encoder.model.", ids, mask, num_layers=n)
            valid = mask.bool()
            assert rel(got[valid], hs[n][valid]) < 1e-5, n


def test_convnext_stage_taps_match_torchvision():
    """ImageEncoder(backbone="convnext_*") takes the outputs of the residual stages 2 / 3 / 4 as layer2 / 3 / 4:
    the oracle's stage taps equal torchvision's `features[:4]`, `[:6]`, `[:8]`."""
    import torchvision
    net = torchvision.models.convnext_tiny(weights=None).eval()
    sd = {"features." + k: v for k, v in weights.synth_state_dict(net.features.state_dict(), seed=2).items()}
    net.features.load_state_dict({k[len("features."):]: v for k, v in sd.items()})
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    taps = []
    with torch.no_grad():
        out = port.convnext_features(sd, "features.", x, taps=taps)
        assert rel(out, net.features(x)) < 1e-5
        for tap, upto in zip(taps[1:], (4, 6, 8)):
            assert rel(tap, net.features[:upto](x)) < 1e-5
