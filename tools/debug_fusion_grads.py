import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_model_gpu import _setup, _zero_dropout, cos, rel
from oracle import port
import mdhs_b200.functional as Fm
model, sd, images, ids, mask, labels = _setup("basic", "mlp", B=8, S=16, hw=128)
model.train(); _zero_dropout(model)
feats = model.forward_features(images.cuda(), ids.cuda(), mask.cuda())
logits = model.classifier(feats)
loss = Fm.cross_entropy(logits, labels.cuda(), label_smoothing=0.02)
loss.backward(); torch.cuda.synchronize()
sd_g = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
want = port.model_forward(sd_g, images, ids, mask, fusion="basic", head="mlp", training_bn=True)
port.ce_label_smoothing(want, labels, label_smoothing=0.02).backward()
named = dict(model.named_parameters())
for k, p in named.items():
    if k.startswith("fusion.") or k.startswith("classifier.") or "proj4" in k:
        g, r = p.grad, sd_g[k].grad
        print(f"{k:60s} cos {cos(g, r):.4f}  |g| {g.norm().item():.3e} |ref| {r.norm().item():.3e}")
print("mask", mask.sum(1))
