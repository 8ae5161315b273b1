"""Golden vectors: tests/golden/reference_outputs.pt holds outputs of the REAL reference (generated in the dev
container by tests/golden/make_golden.py).  CPU tests pin the oracle to them anywhere; GPU tests compare the CUDA
implementation with them directly (same tolerances as tests/test_model_gpu.py)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port, weights  # noqa: E402
from refutil import build_ours  # noqa: E402

GOLD = torch.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.pt"), weights_only=False)
CASES = [k for k in GOLD if k not in ("meta", "train_basic_mlp", "mibf", "kan_moe", "moe_train")]


def rel(a, b):
    return (a.float().cpu() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-12)


def _template(case):
    return build_ours(fusion=case["fusion"], head=case["head"], gate=case["gate"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_logits(name):
    case = GOLD[name]
    sd = weights.synth_state_dict(_template(case).state_dict(), seed=1)
    images, ids, mask, _ = weights.synthetic_batch(4, 16, 7, image_hw=64)
    with torch.no_grad():
        got = port.model_forward(sd, images, ids, mask, fusion=case["fusion"], head=case["head"], gate=case["gate"])
    assert rel(got, case["eval_logits"]) < 1e-5


def test_oracle_reproduces_reference_train_step():
    case = GOLD["train_basic_mlp"]
    sd = weights.synth_state_dict(build_ours(fusion="basic", head="mlp").state_dict(), seed=1)
    sd_g = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    images, ids, mask, labels = weights.synthetic_batch(8, 16, 7, image_hw=128)
    logits = port.model_forward(sd_g, images, ids, mask, fusion="basic", head="mlp", training_bn=True)
    loss = port.ce_label_smoothing(logits, labels, label_smoothing=0.02)
    loss.backward()
    assert rel(logits.detach(), case["logits"]) < 1e-4
    assert abs(loss.item() - case["loss"].item()) < 1e-5
    for k, g in case["grads"].items():
        assert rel(sd_g[k].grad, g) < 2e-3, k


def test_oracle_reproduces_reference_kan_moe():
    case = GOLD["kan_moe"]
    x = case["x"]
    # state_dict templates come from shapes only: rebuild them from the reference layer sizes
    def kan_template(prefix, sizes, sd):
        for i, (a, b) in enumerate(zip(sizes, sizes[1:])):
            sd[f"{prefix}layers.{i}.grid"] = case["kan_grid"][:1].expand(a, -1).contiguous()
            sd[f"{prefix}layers.{i}.base_weight"] = torch.empty(b, a)
            sd[f"{prefix}layers.{i}.spline_weight"] = torch.empty(b, a, 8)
            sd[f"{prefix}layers.{i}.spline_scaler"] = torch.empty(b, a)
        return sd
    sdk = weights.synth_state_dict(kan_template("", [64, 32, 7], {}), seed=4)
    assert rel(port.kan_net(sdk, "", x), case["kan_out"]) < 1e-5
    tm = {}
    for e in range(4):
        kan_template(f"experts.{e}.", [64, 32, 7], tm)
    tm["w_gate"] = torch.empty(64, 4)
    tm["w_noise"] = torch.empty(64, 4)
    tm["mean"] = torch.tensor([0.0])
    tm["std"] = torch.tensor([1.0])
    # keep the reference's parameter order so per-key seeds line up (keys are what matter)
    sdm = weights.synth_state_dict(tm, seed=5)
    sdm["mean"], sdm["std"] = torch.tensor([0.0]), torch.tensor([1.0])
    y, l = port.moe_forward_eval(sdm, "", x, 4, 2)
    assert rel(y, case["moe_out"]) < 1e-5
    assert abs(l.item() - case["moe_loss"].item()) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_logits(name):
    case = GOLD[name]
    model = _template(case)
    model.load_state_dict(weights.synth_state_dict(model.state_dict(), seed=1))
    model = model.cuda().eval()
    images, ids, mask, _ = weights.synthetic_batch(4, 16, 7, image_hw=64)
    with torch.no_grad():
        got = model(images.cuda(), ids.cuda(), mask.cuda())
    want = case["eval_logits"]
    err = rel(got, want)
    assert err < 2e-2, err
    top2 = want.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 2 * (got.float().cpu() - want).abs().max()
    assert torch.equal(got.float().cpu().argmax(1)[safe], want.argmax(1)[safe])


@pytest.mark.gpu
def test_cuda_mibf_matches_reference_and_oracle():
    """MIBF-Net (ResNet-50 + BERT CLS + IBFA x2 + 3 heads + MP-Loss): eval logits vs the real reference's golden
    outputs (<= 3e-2 max-norm: 768-wide single-head attention over two keys in bf16), MP-loss within 2e-2, and a
    train-mode backward whose head gradients match the fp32 oracle (cosine >= 0.98)."""
    import mdhs_b200
    from mdhs_b200.mibf_net import Resnet50WithOurs
    from refutil import bert_dir, quiet
    with quiet():
        model = Resnet50WithOurs(num_labels=6, bert_path=bert_dir(), pretrained=False)
    sd = weights.synth_state_dict(model.state_dict(), seed=3)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    im, ii, mm, ll = weights.synthetic_batch(4, 16, 6, image_hw=64, unit_range=True)
    batch = {"transformed_image": im.cuda(), "input_ids": ii.cuda(), "attention_mask": mm.cuda()}
    with torch.no_grad():
        out = model(batch)
        loss = model.cal_loss(out, ll.cuda())
    gold = GOLD["mibf"]
    for k in ("image_text", "text", "image"):
        assert rel(out[k], gold[k]) < 3e-2, (k, rel(out[k], gold[k]))
    assert abs(loss.item() - gold["mp_loss"].item()) < 2e-2 * max(1.0, abs(gold["mp_loss"].item()))
    # backward (train-mode BN, BERT dropout off)
    model.text_encoder.bert.config.hidden_dropout_prob = 0.0
    model.text_encoder._engine.p_hidden = model.text_encoder._engine.p_attn = 0.0
    model.train()
    out = model(batch)
    loss = model.cal_loss(out, ll.cuda())
    loss.backward()
    sd_g = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    o = port.mibf_forward(sd_g, im, ii, mm, training_bn=True)
    lref = port.mp_loss(o["image"], o["text"], o["image_text"], ll)
    lref.backward()
    assert abs(loss.item() - lref.item()) < 3e-2 * max(1.0, abs(lref.item()))
    named = dict(model.named_parameters())
    for key in ("fc.weight", "fc_text.3.weight", "fc_image.1.weight", "textbased_cross_attention.to_out.weight",
                "imagbased_cross_attention.toV_y.weight", "image_encoder.fc.weight"):
        g, r = named[key].grad.float().cpu().flatten(), sd_g[key].grad.flatten()
        c = (g @ r / (g.norm() * r.norm() + 1e-30)).item()
        assert c > 0.98, (key, c)
    # parameters that never receive a gradient in the reference (BERT pooler, I2Iattention) stay untouched here too
    assert named["I2Iattention.query.weight"].grad.abs().sum().item() == 0


# ----------------------------------------------------------------------------------------------------------------
# KAN / MoE (ConNexT/models/block/kan1.py, moe.py)
# ----------------------------------------------------------------------------------------------------------------
def _moe_template(prefix_sizes=(64, 32, 7), E=4):
    case = GOLD["kan_moe"]
    tm = {}
    for e in range(E):
        for i, (a, b) in enumerate(zip(prefix_sizes, prefix_sizes[1:])):
            p = f"experts.{e}.layers.{i}."
            tm[p + "grid"] = case["kan_grid"][:1].expand(a, -1).contiguous()
            tm[p + "base_weight"] = torch.empty(b, a)
            tm[p + "spline_weight"] = torch.empty(b, a, 8)
            tm[p + "spline_scaler"] = torch.empty(b, a)
    tm["w_gate"] = torch.empty(prefix_sizes[0], E)
    tm["w_noise"] = torch.empty(prefix_sizes[0], E)
    tm["mean"] = torch.tensor([0.0])
    tm["std"] = torch.tensor([1.0])
    return tm


def test_oracle_reproduces_reference_moe_training():
    """oracle/port.py::moe_forward_train against the real reference's training-mode MoE (noise injected): output,
    balance loss and gradients."""
    case = GOLD["moe_train"]
    sd = weights.synth_state_dict(_moe_template(), seed=6)
    assert torch.equal(sd["mean"], case["mean"]) and torch.equal(sd["std"], case["std"])
    sd_g = {k: v.clone().requires_grad_(v.is_floating_point() and k not in ("mean", "std") and not k.endswith("grid"))
            for k, v in sd.items()}
    x = case["x"].clone().requires_grad_(True)
    y, l = port.moe_forward_train(sd_g, "", x, case["noise"], 4, 2)
    (y.square().sum() + l).backward()
    assert rel(y, case["y"]) < 1e-5
    assert abs(l.item() - case["loss"].item()) < 1e-6
    assert rel(x.grad, case["dx"]) < 1e-4
    for k, g in case["grads"].items():
        assert rel(sd_g[k].grad, g) < 1e-4, k


@pytest.mark.gpu
def test_cuda_kan_matches_reference():
    """KAN1 [64, 32, 7] on the tcgen05 GEMM (bf16 basis operand, fp32 accumulate) vs the real reference's output on
    inputs that include the knot-range edges (-2.2, 2.2) and values outside it.  Tolerance 1e-2 max-norm."""
    import mdhs_b200  # noqa: F401
    from mdhs_b200.connext import KAN1
    case = GOLD["kan_moe"]
    kan = KAN1([64, 32, 7])
    sdk = weights.synth_state_dict(kan.state_dict(), seed=4)
    kan.load_state_dict(sdk)
    kan = kan.cuda().eval()
    with torch.no_grad():
        y = kan(case["x"].cuda())
    assert y.shape == case["kan_out"].shape
    assert rel(y, case["kan_out"]) < 1e-2, rel(y, case["kan_out"])


@pytest.mark.gpu
def test_cuda_moe_eval_matches_reference():
    import mdhs_b200  # noqa: F401
    from mdhs_b200.connext import MoE
    case = GOLD["kan_moe"]
    moe = MoE(input_size=64, output_size=7, num_experts=4, hidden_size=32, k=2, layers_hidden=[64, 32, 7])
    sdm = weights.synth_state_dict(moe.state_dict(), seed=5)
    sdm["mean"], sdm["std"] = torch.tensor([0.0]), torch.tensor([1.0])
    moe.load_state_dict(sdm)
    moe = moe.cuda().eval()
    with torch.no_grad():
        y, l = moe(case["x"].cuda())
    assert rel(y, case["moe_out"]) < 1e-2, rel(y, case["moe_out"])
    assert abs(l.item() - case["moe_loss"].item()) < 1e-5


@pytest.mark.gpu
def test_cuda_moe_training_matches_reference():
    """Noisy top-k gating, smooth load estimator, balance loss and every gradient (experts through the bf16 GEMMs:
    and the gate weights that see them through the combine step: <= 2e-2 max-norm) vs the golden outputs of the real reference."""
    import mdhs_b200  # noqa: F401
    from mdhs_b200.connext import MoE
    case = GOLD["moe_train"]
    moe = MoE(input_size=64, output_size=7, num_experts=4, hidden_size=32, k=2, layers_hidden=[64, 32, 7])
    sd = weights.synth_state_dict(moe.state_dict(), seed=6)
    moe.load_state_dict(sd)
    moe = moe.cuda().train()
    object.__setattr__(moe, "_noise_override", case["noise"].cuda())
    x = case["x"].cuda().requires_grad_(True)
    y, l = moe(x)
    (y.square().sum() + l).backward()
    assert rel(y, case["y"]) < 1e-2, rel(y, case["y"])
    assert abs(l.item() - case["loss"].item()) < 1e-5 * max(1.0, abs(case["loss"].item()))
    assert rel(x.grad, case["dx"]) < 2e-2, rel(x.grad, case["dx"])
    named = dict(moe.named_parameters())
    for k, g in case["grads"].items():
        assert rel(named[k].grad, g) < 2e-2, (k, rel(named[k].grad, g))
