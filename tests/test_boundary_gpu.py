"""Boundary tests (GPU): the ways the reference's own scripts reach into the modules.

  * the stock training loop of scripts/train.py:364-387 -- nn.CrossEntropyLoss(label_smoothing) + torch.optim.AdamW +
    optimizer.zero_grad() / loss.backward() / optimizer.step() -- driven through OUR modules, loss trajectory vs the oracle;
  * freeze flags (scripts/train.py:214-219, model.freeze_encoders()): a frozen trunk / text encoder trains without error,
    frozen weights stay bit-identical, the fused optimizer leaves them (and statically unused parameters) untouched;
  * forward / full-backward hooks on image_encoder.stem | layerN[-1] and model.fusion (analysis_tools.py:29-31,154,
    scripts/run_analysis.py:126-132): Grad-CAM's activations and gradients arrive as NCHW fp32 tensors.
"""
import copy
import os
import sys

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from refutil import build_ours  # noqa: E402
from oracle import port, weights  # noqa: E402
from test_model_gpu import _zero_dropout  # noqa: E402


def _model(fusion="concat", head="mlp", seed=1, **kw):
    model = build_ours(fusion=fusion, head=head, **kw)
    sd = weights.synth_state_dict(model.state_dict(), seed=seed)
    model.load_state_dict(sd)
    return model.cuda(), sd


def test_stock_training_loop_matches_oracle():
    """scripts/train.py:364-387 verbatim on our modules (3 steps, dropout off so that both sides are deterministic)."""
    model, sd = _model("basic")
    images, ids, mask, labels = weights.synthetic_batch(8, 16, 7, image_hw=96)
    ci, cd, cm, cl = images.cuda(), ids.cuda(), mask.cuda(), labels.cuda()
    model.train()
    _zero_dropout(model)
    with torch.no_grad():
        model(ci, cd, cm)          # binds the engines
    _zero_dropout(model)
    model.load_state_dict(sd)      # undo the BN running-stat update of the binding pass
    criterion = nn.CrossEntropyLoss(label_smoothing=0.02)
    # Adam moves every weight by ~lr per step whatever the gradient scale: at the config's 2e-4 an 8-sample batch collapses
    # the loss from 2.8 to 0.8 in ONE step and the third step is chaotic in the last bf16 bits; 1e-5 keeps the trajectory smooth
    LR = 1e-5
    optimizer = torch.optim.AdamW(filter(lambda p: p.requires_grad, model.parameters()), lr=LR)
    ours = []
    for _ in range(3):
        optimizer.zero_grad()
        fused = model.forward_features(ci, cd, cm)
        logits = model.classifier(fused)
        loss = criterion(logits.float(), cl)
        loss.backward()
        optimizer.step()
        ours.append(loss.item())
    # oracle: the same loop on the functional restatement (fp32, CPU)
    state = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    seen, plist = set(), []
    canon = {}
    for k, v in state.items():      # aliased keys (image_encoder.stem.* == image_encoder.model.*) share ONE leaf
        c = weights._canonical(k)
        if c in canon:
            state[k] = canon[c]
        else:
            canon[c] = v
            if v.requires_grad and id(v) not in seen:
                seen.add(id(v))
                plist.append(v)
    opt = torch.optim.AdamW(plist, lr=LR)
    ref = []
    for _ in range(3):
        opt.zero_grad()
        lg = port.model_forward(state, images, ids, mask, fusion="basic", head="mlp", training_bn=True)
        ls = port.ce_label_smoothing(lg, labels, label_smoothing=0.02)
        ls.backward()
        opt.step()
        ref.append(ls.item())
    print(f"[stock loop] ours {ours} oracle {ref}")
    assert ours[-1] < ours[0] and ref[-1] < ref[0], "the loss must go down on both sides"
    assert abs(ours[0] - ref[0]) < 2e-2 * max(1.0, abs(ref[0])), (ours, ref)     # same weights: forward parity only
    for a, b in zip(ours[1:], ref[1:]):
        # after AdamW steps (sign-like updates of size lr amplify bf16 gradient noise) the trajectories stay within 5 %
        assert abs(a - b) < 5e-2 * max(1.0, abs(b)), (ours, ref)
    # the drop in loss over the three steps agrees as well (the optimizer really stepped through .grad views)
    assert abs((ours[0] - ours[-1]) - (ref[0] - ref[-1])) < 0.35 * abs(ref[0] - ref[-1]) + 2e-2


@pytest.mark.parametrize("freeze", ["both", "image", "text"])
def test_frozen_encoders_train_step(freeze):
    """`image_encoder.freeze: true` / `text_encoder.freeze: true` (scripts/train.py:214-219) and model.freeze_encoders():
    the step runs, frozen weights stay bit-identical (AdamW's decoupled decay must not touch them), the rest moves."""
    from mdhs_b200.train import Trainer
    model, sd = _model("basic")
    if freeze == "both":
        model.freeze_encoders()
    elif freeze == "image":
        for p in model.image_encoder.parameters():
            p.requires_grad = False
    else:
        for p in model.text_encoder.parameters():
            p.requires_grad = False
    images, ids, mask, labels = weights.synthetic_batch(8, 16, 7, image_hw=64)
    batch = [t.cuda() for t in (images, ids, mask, labels)]
    tr = Trainer(model, optimizer="adamw", lr=1e-3)
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    l0, _ = tr.step(*batch)
    l1, _ = tr.step(*batch)
    torch.cuda.synchronize()
    assert torch.isfinite(l0).all() and torch.isfinite(l1).all()
    after = dict(model.named_parameters())
    frozen_img = freeze in ("both", "image")
    frozen_txt = freeze in ("both", "text")
    for k, v in before.items():
        changed = not torch.equal(v, after[k].detach())
        if k.startswith("image_encoder.") and frozen_img:
            assert not changed, k
        elif k.startswith("text_encoder.") and frozen_txt:
            assert not changed, k
        elif k.startswith("text_encoder.model.pooler."):
            assert not changed, k          # never part of the graph: torch.optim would skip it (grad is None)
        elif k.startswith(("classifier.", "fusion.transformer_block.ff", "image_encoder.proj4")):
            assert changed, k
    # also through CUDA-graph capture
    tr2 = Trainer(model, optimizer="adamw", lr=1e-3)
    tr2.capture(*batch, warmup=1)
    l2, _ = tr2.replay()
    assert torch.isfinite(l2).all()


def test_fused_adamw_matches_torch_on_trainable_subset():
    """The fused optimizer equals torch.optim.AdamW(filter(requires_grad)) after one step from identical gradients."""
    from mdhs_b200.train import Trainer
    model, sd = _model("concat")
    for p in model.text_encoder.parameters():
        p.requires_grad = False
    images, ids, mask, labels = weights.synthetic_batch(4, 16, 7, image_hw=64)
    batch = [t.cuda() for t in (images, ids, mask, labels)]
    model.train()
    _zero_dropout(model)
    with torch.no_grad():
        model(*batch[:3])
    _zero_dropout(model)
    model.load_state_dict(sd)
    # reference update: stock loop with torch.optim.AdamW on a deep copy of the parameters + our gradients
    criterion = nn.CrossEntropyLoss(label_smoothing=0.02)
    opt = torch.optim.AdamW(filter(lambda p: p.requires_grad, model.parameters()), lr=1e-3)
    opt.zero_grad()
    loss = criterion(model.classifier(model.forward_features(*batch[:3])).float(), batch[3])
    loss.backward()
    opt.step()
    want = {k: v.detach().clone() for k, v in model.named_parameters()}
    model.load_state_dict(sd)
    tr = Trainer(model, optimizer="adamw", lr=1e-3, label_smoothing=0.02)
    tr.step(*batch)
    torch.cuda.synchronize()
    for k, v in model.named_parameters():
        w = want[k]
        if not v.requires_grad:
            assert torch.equal(v.detach().cpu(), sd[k]), k
            continue
        # same gradients up to run-to-run atomics order; Adam's first step moves every weight by ~lr * sign(g)
        assert (v.detach() - w).abs().max().item() <= 2.5e-3, k


def test_hooks_fire_with_nchw_activations_and_gradients():
    """Grad-CAM's access pattern (analysis_tools.py:9-76, scripts/run_analysis.py:126-132): forward + full-backward hooks on
    image_encoder.stem / layerN[-1], a forward hook on model.fusion; eval mode, logits.backward(one_hot)."""
    model, sd = _model("basic")
    model.eval()
    images, ids, mask, _ = weights.synthetic_batch(2, 16, 7, image_hw=64)
    ci, cd, cm = images.cuda(), ids.cuda(), mask.cuda()
    acts, grads, fused = {}, {}, []
    enc = model.image_encoder
    targets = {"stem": enc.stem, "layer1": enc.layer1[-1], "layer2": enc.layer2[-1], "layer3": enc.layer3[-1], "layer4": enc.layer4[-1]}
    handles = []
    for name, layer in targets.items():
        handles.append(layer.register_forward_hook(lambda m, i, o, name=name: acts.__setitem__(name, o)))
        handles.append(layer.register_full_backward_hook(lambda m, gi, go, name=name: grads.__setitem__(name, go[0])))
    handles.append(model.fusion.register_forward_hook(lambda m, i, o: fused.append(o.detach().float().cpu())))
    model.zero_grad()
    logits = model(ci, cd, cm)
    one_hot = torch.zeros_like(logits)
    one_hot[torch.arange(2), logits.argmax(1)] = 1
    logits.backward(gradient=one_hot, retain_graph=True)
    torch.cuda.synchronize()
    assert len(fused) == 1 and fused[0].shape == (2, 256)
    shapes = {"stem": (2, 64, 16, 16), "layer1": (2, 256, 16, 16), "layer2": (2, 512, 8, 8), "layer3": (2, 1024, 4, 4),
              "layer4": (2, 2048, 2, 2)}
    # oracle: activations and their gradients from the fp32 restatement
    state = {k: v.clone() for k, v in sd.items()}
    x = images.clone()
    import torch.nn.functional as F
    feats = port.resnet_features(state, "image_encoder.model.", x, "resnet50", False)
    for name, shp in shapes.items():
        assert name in acts and name in grads, name
        assert tuple(acts[name].shape) == shp and acts[name].dtype == torch.float32, (name, acts[name].shape, acts[name].dtype)
        assert tuple(grads[name].shape) == shp, (name, grads[name].shape)
        assert torch.isfinite(grads[name]).all() and grads[name].abs().max().item() > 0
        if name != "stem":
            w = feats[name]
            assert (acts[name].cpu() - w).abs().max().item() <= 2e-2 * w.abs().max().item(), name
    # gradient wrt layer4 output vs the oracle (eval-mode BN everywhere)
    f4 = feats["layer4"].detach().clone().requires_grad_(True)
    tok = F.linear(f4.flatten(2).transpose(1, 2), sd["image_encoder.proj4.weight"], sd["image_encoder.proj4.bias"])
    txt = port.bert_last_hidden(sd, "text_encoder.model.", ids, mask)
    lg = port.head_mlp(sd, "classifier.", port.fusion_basic(sd, "fusion.", tok, txt, mask, 8))
    lg.backward(gradient=one_hot.cpu())
    g, gw = grads["layer4"].cpu(), f4.grad
    c = torch.nn.functional.cosine_similarity(g.flatten(), gw.flatten(), dim=0).item()
    assert c > 0.98, c
    for h in handles:
        h.remove()
    # without hooks the fast path is back and gives the same logits
    with torch.no_grad():
        again = model(ci, cd, cm)
    assert (again - logits.detach()).abs().max().item() <= 1e-3 * logits.abs().max().item() + 1e-6


def test_eval_mode_backward_uses_running_statistics():
    """Eval-mode BatchNorm backward is dx = dy * gamma * rsqrt(running_var + eps) (no batch-statistics terms): gradient of the
    stem convolution against the oracle in eval mode."""
    import mdhs_b200.functional as Fm
    model, sd = _model("concat")
    model.eval()
    images, ids, mask, labels = weights.synthetic_batch(4, 16, 7, image_hw=64)
    model.store("cuda").zero_grad()
    loss = Fm.cross_entropy(model(images.cuda(), ids.cuda(), mask.cuda()).float(), labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    state = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    ls = port.ce_label_smoothing(port.model_forward(state, images, ids, mask, fusion="concat", head="mlp"), labels, label_smoothing=0.0)
    ls.backward()
    named = dict(model.named_parameters())
    for key, thr in (("image_encoder.model.layer4.2.conv3.weight", 0.98), ("image_encoder.model.layer2.0.conv2.weight", 0.95),
                     ("image_encoder.model.layer1.0.conv1.weight", 0.9), ("image_encoder.model.layer3.1.bn2.weight", 0.95)):
        g, gw = named[key].grad.float().cpu(), state[key].grad
        c = torch.nn.functional.cosine_similarity(g.flatten(), gw.flatten(), dim=0).item()
        assert c > thr, (key, c)


def test_gate_inference_single_encoder_pass():
    """Eval-mode dual-expert gate (model.py:257-281): one image-encoder and one text-encoder pass feed both experts, and the
    logits equal the oracle's two-pass evaluation."""
    model, sd = _model("weighted_concat", gate=True)
    model.eval()
    images, ids, mask, _ = weights.synthetic_batch(4, 16, 7, image_hw=64)
    calls = {"img": 0, "txt": 0}
    h1 = model.image_encoder.register_forward_pre_hook(lambda m, a: calls.__setitem__("img", calls["img"] + 1))
    h2 = model.text_encoder.register_forward_pre_hook(lambda m, a: calls.__setitem__("txt", calls["txt"] + 1))
    with torch.no_grad():
        got = model(images.cuda(), ids.cuda(), mask.cuda()).float().cpu()
    h1.remove()
    h2.remove()
    assert calls == {"img": 1, "txt": 1}, calls
    want = port.model_forward(sd, images, ids, mask, fusion="weighted_concat", head="mlp", gate=True)
    assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item()
    model.train()     # train mode keeps the reference's two passes (BN running statistics see two updates)
    calls.update(img=0, txt=0)
    h1 = model.image_encoder.register_forward_pre_hook(lambda m, a: calls.__setitem__("img", calls["img"] + 1))
    out = model(images.cuda(), ids.cuda(), mask.cuda())
    h1.remove()
    assert calls["img"] == 2 and torch.isfinite(out).all()


def test_kan1_head_matches_oracle():
    """classifier_type="kan1": KAN1([hidden, 256, C]) (ConNexT/models/block/kan1.py:239-289) as the head -- the KAN-head
    variant of BASELINE config 5.  Eval logits and head gradients vs the oracle."""
    import mdhs_b200.functional as Fm
    model, sd = _model("concat", head="kan1")
    model.eval()
    images, ids, mask, labels = weights.synthetic_batch(8, 16, 7, image_hw=64)
    model.store("cuda").zero_grad()
    logits = model(images.cuda(), ids.cuda(), mask.cuda())
    loss = Fm.cross_entropy(logits.float(), labels.cuda(), label_smoothing=0.02)
    loss.backward()
    torch.cuda.synchronize()
    state = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and not k.endswith("grid") else v.clone())
             for k, v in sd.items()}
    want = port.model_forward(state, images, ids, mask, fusion="concat", head="kan1")
    ls = port.ce_label_smoothing(want, labels, label_smoothing=0.02)
    ls.backward()
    assert (logits.detach().float().cpu() - want.detach()).abs().max().item() <= 2e-2 * want.abs().max().item()
    named = dict(model.named_parameters())
    for key in ("classifier.layers.0.base_weight", "classifier.layers.0.spline_weight", "classifier.layers.1.spline_scaler",
                "fusion.proj.weight"):
        g, gw = named[key].grad.float().cpu(), state[key].grad
        c = torch.nn.functional.cosine_similarity(g.flatten(), gw.flatten(), dim=0).item()
        assert c > 0.97, (key, c)


def test_hierarchical_fusion_matches_oracle():
    """fusion_type="hierarchical" (README.md:15: image layer2/3/4 x BERT hidden states 4/8/12, adaptive weighting): the
    tapped hidden states equal the oracle's truncated-BERT outputs; eval logits, and the gradients of the level logits, the
    per-level blocks and an EARLY BERT layer (which receives gradient through all three taps) match the oracle."""
    import mdhs_b200.functional as Fm
    model, sd = _model("hierarchical")
    model.eval()
    images, ids, mask, labels = weights.synthetic_batch(6, 24, 7, image_hw=64)
    ci, cd, cm = images.cuda(), ids.cuda(), mask.cuda()
    with torch.no_grad():
        hs = model.text_encoder(cd, cm, hidden_states=(4, 8, 12))
    assert sorted(hs.keys()) == [4, 8, 12]
    for n in (4, 8, 12):
        want_h = port.bert_last_hidden(sd, "text_encoder.model.", ids, mask, num_layers=n)
        valid = mask.bool()
        err = (hs[n].float().cpu() - want_h)[valid].abs().max().item() / want_h[valid].abs().max().item()
        assert err < 2e-2, (n, err)
    model.store("cuda").zero_grad()
    logits = model(ci, cd, cm)
    loss = Fm.cross_entropy(logits.float(), labels.cuda(), label_smoothing=0.02)
    loss.backward()
    torch.cuda.synchronize()
    state = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    want = port.model_forward(state, images, ids, mask, fusion="hierarchical", head="mlp")
    ls = port.ce_label_smoothing(want, labels, label_smoothing=0.02)
    ls.backward()
    assert (logits.detach().float().cpu() - want.detach()).abs().max().item() <= 2e-2 * want.abs().max().item()
    named = dict(model.named_parameters())
    for key in ("fusion.level_logits", "fusion.cross_l2.txt_proj.weight", "fusion.cross_l3.attn.in_proj_weight",
                "fusion.cross_l4.norm.weight", "text_encoder.model.encoder.layer.2.output.dense.weight",
                "text_encoder.model.encoder.layer.6.intermediate.dense.weight",
                "text_encoder.model.encoder.layer.11.output.dense.weight", "image_encoder.proj2.weight"):
        g, gw = named[key].grad.float().cpu(), state[key].grad
        c = torch.nn.functional.cosine_similarity(g.flatten(), gw.flatten(), dim=0).item()
        assert c > 0.97, (key, c)


@pytest.mark.parametrize("fusion", ["basic", "multiscale"])
def test_convnext_backbone_in_image_encoder(fusion):
    """ImageEncoder(backbone="convnext_tiny"): ConvNeXt stages 2/3/4 as layer2/3/4 tokens behind the same
    MultimodalBaselineModel constructor; eval logits and gradients vs the oracle (torchvision ConvNeXt restated)."""
    import mdhs_b200.functional as Fm
    model, sd = _model(fusion, backbone="convnext_tiny")
    model.eval()
    images, ids, mask, labels = weights.synthetic_batch(4, 16, 7, image_hw=64)
    model.store("cuda").zero_grad()
    logits = model(images.cuda(), ids.cuda(), mask.cuda())
    loss = Fm.cross_entropy(logits.float(), labels.cuda(), label_smoothing=0.02)
    loss.backward()
    torch.cuda.synchronize()
    state = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    want = port.model_forward(state, images, ids, mask, arch="convnext_tiny", fusion=fusion, head="mlp")
    port.ce_label_smoothing(want, labels, label_smoothing=0.02).backward()
    assert (logits.detach().float().cpu() - want.detach()).abs().max().item() <= 2e-2 * want.abs().max().item()
    named = dict(model.named_parameters())
    keys = ["image_encoder.proj4.weight", "image_encoder.model.features.7.2.block.3.weight",
            "image_encoder.model.features.5.4.block.0.weight", "image_encoder.model.features.1.0.block.5.weight",
            "image_encoder.model.features.0.0.weight"]
    if fusion == "multiscale":
        keys.append("image_encoder.proj2.weight")
    for key in keys:
        g, gw = named[key].grad.float().cpu(), state[key].grad
        c = torch.nn.functional.cosine_similarity(g.flatten(), gw.flatten(), dim=0).item()
        assert c > 0.95, (key, c)
