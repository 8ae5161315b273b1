"""End-to-end parity (GPU): MultimodalBaselineModel on the B200 kernels vs the CPU oracle (oracle/port.py, which
is pinned to the real reference) on identical synthetic weights and inputs.

Tolerances (bf16 activations / bf16 tensor-core operands, fp32 accumulation and statistics):
  * eval-mode logits: max-norm relative error <= 2e-2 (SURVEY 8c measured 1e-2 for torch's own bf16 autocast),
    top-1 identical on every sample whose oracle top-1/top-2 margin exceeds the measured logit error;
  * training step (train-mode BN, dropout off): train-mode BatchNorm over 50 layers amplifies bf16 rounding --
    torch's own bf16 autocast of the oracle trunk deviates 4-5e-2 (max-norm) from fp32 at layer4 (measured in
    test_trunk_error_vs_torch_autocast, which requires ours <= 1.5x that) -- so logits must agree within 6e-2 and
    the loss within 2e-2; gradients of head / fusion / BERT parameters cosine >= 0.98-0.99 against the fp32 oracle;
    early ResNet conv gradients cosine >= 0.8 (measured 0.89-0.97; ill-conditioned: torch's own bf16 autocast reaches 0.11-0.57 there,
    SURVEY 8c).
"""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from refutil import build_ours  # noqa: E402
from oracle import port, weights  # noqa: E402


def rel(a, b):
    return (a.float().cpu() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-12)


def cos(a, b):
    a, b = a.float().cpu().flatten(), b.float().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


def _zero_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    for m in model.modules():
        cfg = getattr(m, "config", None)
        if cfg is not None and hasattr(cfg, "hidden_dropout_prob"):  # read by BertEngine when it binds
            cfg.hidden_dropout_prob = 0.0
            cfg.attention_probs_dropout_prob = 0.0
        eng = getattr(m, "_engine", None)
        if eng is not None and hasattr(eng, "p_hidden"):
            eng.p_hidden = eng.p_attn = 0.0


def _setup(fusion, head, gate=False, seed=1, hw=64, B=4, S=16):
    model = build_ours(fusion=fusion, head=head, gate=gate)
    sd = weights.synth_state_dict(model.state_dict(), seed=seed)
    model.load_state_dict(sd)
    model = model.cuda()
    images, ids, mask, labels = weights.synthetic_batch(B, S, 7, image_hw=hw)
    return model, sd, images, ids, mask, labels


@pytest.mark.parametrize("fusion,head,gate", [("basic", "mlp", False), ("multiscale", "residual", False),
                                              ("concat", "attention_pooling", False), ("weighted_concat", "mlp", True),
                                              ("hadamard", "mlp", False), ("bilinear", "mlp", False)])
def test_eval_logits_match_oracle(fusion, head, gate):
    model, sd, images, ids, mask, _ = _setup(fusion, head, gate)
    model.eval()
    with torch.no_grad():
        got = model(images.cuda(), ids.cuda(), mask.cuda())
        want = port.model_forward(sd, images, ids, mask, fusion=fusion, head=head, gate=gate)
    err = rel(got, want)
    assert err < 2e-2, err
    top2 = want.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 2 * (got.float().cpu() - want).abs().max()
    assert torch.equal(got.float().cpu().argmax(1)[safe], want.argmax(1)[safe])


def test_state_dict_round_trip_and_aliases():
    model, sd, *_ = _setup("basic", "mlp")
    out = model.state_dict()
    assert set(out.keys()) == set(sd.keys())
    for k in ("image_encoder.model.conv1.weight", "image_encoder.stem.0.weight", "fusion.transformer_block.attn2.k_proj_weight",
              "text_encoder.model.encoder.layer.3.attention.self.key.weight", "classifier.3.bias"):
        assert torch.equal(out[k].cpu(), sd[k]), k
    # after the first CUDA forward the parameters are views of the flat store; keys and values must not change
    images, ids, mask, _ = weights.synthetic_batch(2, 8, 7, image_hw=64)
    model.eval()
    with torch.no_grad():
        model(images.cuda(), ids.cuda(), mask.cuda())
    out2 = model.state_dict()
    assert set(out2.keys()) == set(sd.keys())
    for k, v in sd.items():
        if v.is_floating_point():
            assert torch.equal(out2[k].cpu(), v), k


@pytest.mark.parametrize("fusion", ["basic", "concat"])
def test_train_step_matches_oracle(fusion):
    model, sd, images, ids, mask, labels = _setup(fusion, "mlp", B=8, S=16, hw=128)
    model.train()
    _zero_dropout(model)
    import mdhs_b200.functional as Fm
    feats = model.forward_features(images.cuda(), ids.cuda(), mask.cuda())
    logits = model.classifier(feats)
    loss = Fm.cross_entropy(logits, labels.cuda(), label_smoothing=0.02)
    loss.backward()
    torch.cuda.synchronize()
    sd_g = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    want = port.model_forward(sd_g, images, ids, mask, fusion=fusion, head="mlp", training_bn=True)
    loss_ref = port.ce_label_smoothing(want, labels, label_smoothing=0.02)
    loss_ref.backward()
    assert rel(logits, want.detach()) < 6e-2
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * max(1.0, abs(loss_ref.item()))
    named = dict(model.named_parameters())
    checks = [("classifier.3.weight", 0.99), ("classifier.0.weight", 0.99),
              ("text_encoder.model.encoder.layer.11.output.dense.weight", 0.99),
              ("text_encoder.model.encoder.layer.0.attention.self.query.weight", 0.98),
              ("text_encoder.model.embeddings.word_embeddings.weight", 0.98),
              ("image_encoder.proj4.weight", 0.99),
              ("image_encoder.model.layer4.2.conv3.weight", 0.95),
              ("image_encoder.model.layer1.0.conv1.weight", 0.8),
              ("image_encoder.model.conv1.weight", 0.8)]
    if fusion == "basic":
        # (attn2 q/k projections and norm2 are not checked: with random-init weights the cross-attention softmax is
        #  almost uniform, their fp32 gradients are ~1e-4 of the value-path gradients, i.e. below bf16 resolution)
        checks += [("fusion.transformer_block.attn2.v_proj_weight", 0.99), ("fusion.transformer_block.ff.0.weight", 0.99),
                   ("fusion.transformer_block.attn1.in_proj_weight", 0.99), ("fusion.transformer_block.norm3.weight", 0.99),
                   ("fusion.transformer_block.attn2.out_proj.weight", 0.99)]
    else:
        checks += [("fusion.proj.weight", 0.99)]
    for key, thr in checks:
        g = named[key].grad
        assert g is not None, key
        c = cos(g, sd_g[key].grad)
        assert c >= thr, (key, c)
    # BatchNorm running statistics were updated like F.batch_norm does
    rm = model.image_encoder.model.bn1.running_mean
    assert (rm.cpu() - sd["image_encoder.model.bn1.running_mean"]).abs().max().item() > 0


def test_trunk_error_vs_torch_autocast():
    """Train-mode ResNet-50 trunk: our bf16 error against the fp32 oracle is no worse than 1.5x the error of
    torch's own bf16 autocast running the same oracle code on the same GPU."""
    model, sd, images, ids, mask, _ = _setup("basic", "mlp", B=8, hw=128)
    model.train()
    with torch.no_grad():
        model.forward_features(images.cuda(), ids.cuda(), mask.cuda())  # binds the engines
        want = port.resnet_features(sd, "image_encoder.model.", images, "resnet50", True)
        feats, _ = model.image_encoder._engine.forward(images.cuda(), True, False)
        sdc = {k: v.cuda() for k, v in sd.items() if k.startswith("image_encoder.model.")}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            auto = port.resnet_features(sdc, "image_encoder.model.", images.cuda(), "resnet50", True)
    from mdhs_b200 import ops
    for name in ("layer2", "layer3", "layer4"):
        x2d, H, W, C = feats[name]
        ours = rel(ops.nhwc_bf16_to_nchw_f32(x2d, images.shape[0], H, W, C), want[name])
        theirs = rel(auto[name].float(), want[name])
        assert ours <= 1.5 * theirs + 5e-3, (name, ours, theirs)


@pytest.mark.gpu
def test_batched_tta_equals_per_variant_mean():
    """inference.predict_tta (one expand kernel, image encoder on V*B, BERT once) == mean over per-variant model() calls
    on torch-made variants (scripts/predict.py:33-81)."""
    import mdhs_b200  # noqa: F401
    from mdhs_b200 import ops
    from mdhs_b200.inference import predict_tta
    model = build_ours(fusion="basic", head="mlp")
    sd = weights.synth_state_dict(model.state_dict(), seed=1)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    images, ids, mask, _ = weights.synthetic_batch(3, 16, 7, image_hw=64)
    images, ids, mask = images.cuda(), ids.cuda(), mask.cuda()
    tr = ("hflip", "vflip", "rot90")
    big = ops.tta_expand(images, tr)
    want_big = torch.cat([images, images.flip(-1), images.flip(-2), torch.rot90(images, k=1, dims=(-2, -1))], dim=0)
    assert torch.equal(big, want_big)
    with torch.no_grad():
        want = torch.stack([model(v, ids, mask).float() for v in want_big.split(3)], dim=0).mean(dim=0)
    got = predict_tta(model, images, ids, mask, tr)
    assert (got - want).abs().max().item() <= 1e-3 * want.abs().max().item() + 1e-5


@pytest.mark.gpu
def test_trainer_supcon_finetune_step_runs_and_changes_loss():
    import mdhs_b200  # noqa: F401
    from mdhs_b200.train import Trainer
    torch.manual_seed(0)
    model = build_ours(fusion="concat", head="mlp").cuda()
    images, ids, mask, labels = weights.synthetic_batch(8, 16, 7, image_hw=64)
    batch = [t.cuda() for t in (images, ids, mask, labels)]
    l0, _ = Trainer(model, lr=0.0, supcon_weight=0.0).step(*batch)
    l1, _ = Trainer(model, lr=0.0, supcon_weight=0.5, supcon_stage="finetune").step(*batch)
    l2, _ = Trainer(model, lr=0.0, supcon_weight=1.0, supcon_stage="pretrain").step(*batch)
    for l in (l0, l1, l2):
        assert torch.isfinite(l).all()
    assert l1.item() > l0.item() - 1.0 and abs(l1.item() - l0.item()) > 1e-4   # the contrastive term is present


@pytest.mark.gpu
def test_tabular_branch_matches_oracle():
    """TabularEncoder + tabular_fusion (modules/tabular.py, model.py:155-167,229-235) on the fp32 head kernels."""
    model = build_ours(fusion="concat", head="mlp", tabular_enabled=True, tabular_input_dim=12)
    sd = weights.synth_state_dict(model.state_dict(), seed=7)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    images, ids, mask, _ = weights.synthetic_batch(3, 16, 7, image_hw=64)
    tab = torch.randn(3, 12, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        got = model(images.cuda(), ids.cuda(), mask.cuda(), tabular_input=tab.cuda()).float().cpu()
        want = port.model_forward(sd, images, ids, mask, fusion="concat", head="mlp", tabular=tab)
    assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item()
    with pytest.raises(ValueError):
        model(images.cuda(), ids.cuda(), mask.cuda())


@pytest.mark.gpu
@pytest.mark.parametrize("fusion,combine", [("basic", "avg"), ("concat", "concat"), ("multiscale", "avg")])
def test_global_local_branch_matches_oracle(fusion, combine):
    """Global + centre-crop views (model.py:292-315): crop/resize kernel vs F.interpolate, then eval logits vs the oracle."""
    from mdhs_b200 import ops
    images, ids, mask, _ = weights.synthetic_batch(2, 16, 7, image_hw=64)
    both = ops.global_local(images.cuda(), 0.6).cpu()
    assert torch.equal(both[:2], images)
    assert (both[2:] - port.center_crop_resize(images, 0.6)).abs().max().item() < 1e-5
    model = build_ours(fusion=fusion, head="mlp", global_local_enabled=True, global_local_crop_ratio=0.6,
                       global_local_combine=combine)
    sd = weights.synth_state_dict(model.state_dict(), seed=8)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    with torch.no_grad():
        got = model(images.cuda(), ids.cuda(), mask.cuda()).float().cpu()
        want = port.model_forward(sd, images, ids, mask, fusion=fusion, head="mlp", global_local=0.6, global_local_combine=combine)
    assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item()


@pytest.mark.gpu
def test_full_size_properties_config2():
    """BASELINE config-2 sizes (B = 128, 3x224x224, S = 64) are too big for the CPU oracle, so they are checked through
    size-independent properties: (i) eval-mode logits of a sample do not depend on the batch it is evaluated in
    (128 at once == 4 chunks of 32, up to bf16 GEMM tiling effects: none expected, rows are independent);
    (ii) repeating the training step from the same weights reproduces loss and logits up to the dropout noise floor;
    (iii) logits are finite and the loss is near ln(7) at random init."""
    import copy
    import math
    from mdhs_b200.train import Trainer
    torch.manual_seed(0)
    model = build_ours(fusion="basic", head="mlp").cuda()
    images, ids, mask, labels = weights.synthetic_batch(128, 64, 7, image_hw=224)
    images, ids, mask, labels = images.cuda(), ids.cuda(), mask.cuda(), labels.cuda()
    model.eval()
    with torch.no_grad():
        full = model(images, ids, mask).float()
        parts = torch.cat([model(images[i:i + 32], ids[i:i + 32], mask[i:i + 32]).float() for i in range(0, 128, 32)])
    assert torch.isfinite(full).all()
    assert (full - parts).abs().max().item() <= 1e-3 * full.abs().max().item() + 1e-5
    assert torch.equal(full.argmax(1), parts.argmax(1))
    # (ii) switch off the dropouts that are reachable as attributes and repeat the step
    model.store("cuda")
    eng = model.text_encoder._engine
    eng.p_hidden = eng.p_attn = 0.0
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float):
            m.dropout = 0.0
    sd0 = copy.deepcopy(model.state_dict())
    losses = []
    for _ in range(2):
        model.load_state_dict(sd0)
        tr = Trainer(model, lr=1e-4)
        loss, logits = tr.step(images, ids, mask, labels)
        losses.append((loss.item(), logits.float().clone()))
    assert abs(losses[0][0] - math.log(7)) < 1.0
    # repeatability: the two steps start from identical weights; what may still differ is the stateless-dropout tick of the
    # fused attention blocks (their p is not an nn.Dropout attribute), so the bound is the dropout noise floor, not 0
    assert abs(losses[0][0] - losses[1][0]) < 2e-2
    assert (losses[0][1] - losses[1][1]).abs().max().item() <= 5e-2 * losses[0][1].abs().max().item() + 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("fusion,layers,bidir,hid,kind", [("concat", 1, True, 256, "lstm"), ("basic", 2, False, 128, "lstm"),
                                                          ("concat", 2, True, 128, "gru"), ("concat", 2, True, 128, "transformer")])
def test_sequence_lstm_branch_matches_oracle(fusion, layers, bidir, hid, kind):
    """Multi-slice input (B, T, 3, H, W) -> SequenceEncoder (LSTM) (model.py:316-331, modules/sequence_blocks.py): eval logits
    and the LSTM parameter gradients of a train step against the oracle (pinned to the reference's nn.LSTM)."""
    from mdhs_b200 import functional as Fm
    kw = dict(sequence_enabled=True, sequence_type=kind, sequence_hidden_dim=hid, sequence_num_layers=layers,
              sequence_bidirectional=bidir, sequence_dropout=0.0)
    model = build_ours(fusion=fusion, head="mlp", **kw)
    sd = weights.synth_state_dict(model.state_dict(), seed=11)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    images, ids, mask, labels = weights.synthetic_batch(2 * 3, 16, 7, image_hw=64)
    images = images.view(2, 3, 3, 64, 64)
    ids, mask, labels = ids[:2], mask[:2], labels[:2]
    with torch.no_grad():
        got = model(images.cuda(), ids.cuda(), mask.cuda()).float().cpu()
        want = port.model_forward(sd, images, ids, mask, fusion=fusion, head="mlp")
    # the whole image side collapses into ONE bf16 token here, so its rounding is not averaged over 49 tokens as in the
    # other model tests (measured 1.0e-2 for concat, 3.2e-2 for the basic block): 4e-2
    assert (got - want).abs().max().item() <= 4e-2 * want.abs().max().item()
    # gradients of the recurrent weights (eval-mode BN / no dropout so that both sides are deterministic)
    sd_g = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    loss_ref = port.ce_label_smoothing(port.model_forward(sd_g, images, ids, mask, fusion=fusion, head="mlp"), labels, label_smoothing=0.0)
    loss_ref.backward()
    st = model.store("cuda")
    st.zero_grad()
    feats = model.forward_features(images.cuda(), ids.cuda(), mask.cuda())
    loss = Fm.cross_entropy(model.classifier(feats), labels.cuda())
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 2e-2
    named = dict(model.named_parameters())
    if kind == "transformer":
        keys = ("sequence_encoder.encoder.layers.0.self_attn.in_proj_weight", "sequence_encoder.encoder.layers.1.linear2.weight",
                "sequence_encoder.encoder.layers.0.norm1.weight")
        for key in keys:
            g, g_ref = st.g32(named[key]).float().cpu(), sd_g[key].grad
            cos = torch.nn.functional.cosine_similarity(g.flatten(), g_ref.flatten(), dim=0).item()
            assert cos > 0.95, (key, cos)
        return
    for key in ("sequence_encoder.rnn.weight_ih_l0", "sequence_encoder.rnn.weight_hh_l0", "sequence_encoder.rnn.bias_hh_l0"):
        g, g_ref = st.g32(named[key]).float().cpu(), sd_g[key].grad
        cos = torch.nn.functional.cosine_similarity(g.flatten(), g_ref.flatten(), dim=0).item()
        # B = 2 samples through a bf16 trunk: the bias gradient (a sum of only 2 x T gate gradients) is the noisiest
        assert cos > (0.95 if "bias" in key else 0.98), (key, cos)
    if kind == "lstm":   # both LSTM biases see the same gate gradients: must agree to fp32 rounding
        gi, gh = st.g32(named["sequence_encoder.rnn.bias_ih_l0"]), st.g32(named["sequence_encoder.rnn.bias_hh_l0"])
        assert (gi - gh).abs().max().item() <= 1e-5 * gi.abs().max().item() + 1e-9


@pytest.mark.gpu
def test_trainer_custom_forward_loss_mibf_and_batched_tta():
    """Trainer(forward_loss=...) drives MIBF-Net (batch dict + MP-Loss) through the shared CUDA-graph / fused-optimizer step:
    the captured replay reproduces the eager loss trajectory's first value and the loss goes down; batched TTA over the
    batch-dict model equals the per-variant mean."""
    import mdhs_b200  # noqa: F401
    from mdhs_b200 import ops
    from mdhs_b200.inference import predict_tta_batchdict
    from mdhs_b200.mibf_net.model_resnet import Resnet50WithOurs
    from mdhs_b200.train import Trainer, mibf_forward_loss
    from refutil import bert_dir, quiet
    torch.manual_seed(0)
    with quiet():
        model = Resnet50WithOurs(num_labels=6, bert_path=bert_dir(), pretrained=False).cuda()
    images, ids, mask, labels = weights.synthetic_batch(8, 16, 6, image_hw=64, unit_range=True)
    batch = [t.cuda() for t in (images, ids, mask, labels)]
    tr = Trainer(model, optimizer="sgd", lr=1e-3, forward_loss=mibf_forward_loss)
    l0, _ = tr.step(*batch)
    tr.capture(*batch, warmup=1)
    ls = [tr.replay()[0].item() for _ in range(6)]
    assert all(map(lambda v: v == v and abs(v) < 1e4, ls))
    assert ls[-1] < l0.item()
    model.eval()
    tr_names = ("hflip", "rot90")
    big = ops.tta_expand(batch[0], tr_names)
    with torch.no_grad():
        want = torch.stack([model({"transformed_image": v, "input_ids": batch[1], "attention_mask": batch[2]})["image_text"].float()
                            for v in big.split(8)], dim=0).mean(dim=0)
    got = predict_tta_batchdict(model, batch[0], batch[1], batch[2], tr_names)
    assert (got - want).abs().max().item() <= 2e-3 * want.abs().max().item() + 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("B,S,hw,fusion", [(1, 8, 32, "basic"), (5, 33, 96, "basic"), (3, 17, 64, "multiscale"), (1, 64, 224, "concat")])
def test_edge_shapes_match_oracle(B, S, hw, fusion):
    """Ragged / minimal inputs: one sample, a single 1x1 layer-4 token (32x32 image), sequence lengths that are not multiples
    of 8, a fully padded tail, the full-resolution single image.  Eval logits against the oracle (<= 2e-2 max-norm)."""
    model = build_ours(fusion=fusion, head="mlp")
    sd = weights.synth_state_dict(model.state_dict(), seed=1)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    images, ids, mask, _ = weights.synthetic_batch(B, S, 7, image_hw=hw, seed=5)
    mask[0, 1:] = 0          # first sample: only the CLS token is valid
    with torch.no_grad():
        got = model(images.cuda(), ids.cuda(), mask.cuda()).float().cpu()
        want = port.model_forward(sd, images, ids, mask, fusion=fusion, head="mlp")
    assert got.shape == want.shape and torch.isfinite(got).all()
    assert (got - want).abs().max().item() <= 2e-2 * want.abs().max().item()
