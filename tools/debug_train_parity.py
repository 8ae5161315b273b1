import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from refutil import build_ours
from oracle import port, weights
from test_model_gpu import _zero_dropout, rel

for hw, B in ((64, 8), (128, 8), (224, 4)):
    model = build_ours(fusion="basic", head="mlp")
    sd = weights.synth_state_dict(model.state_dict(), seed=1)
    model.load_state_dict(sd)
    model = model.cuda().train()
    _zero_dropout(model)
    images, ids, mask, labels = weights.synthetic_batch(B, 16, 7, image_hw=hw)
    with torch.no_grad():
        tok = model.image_encoder(images.cuda())
        txt = model.text_encoder(ids.cuda(), mask.cuda())
        feats = model.forward_features(images.cuda(), ids.cuda(), mask.cuda())
        logits = model.classifier(feats)
        otok = port.image_encoder(sd, "image_encoder.", images, "resnet50", False, training=True)
        otxt = port.bert_last_hidden(sd, "text_encoder.model.", ids, mask)
        ofeat = port.fusion_basic(sd, "fusion.", otok, otxt, mask)
        ologit = port.head_mlp(sd, "classifier.", ofeat)
        # layer-wise trunk check
        f = port.resnet_features(sd, "image_encoder.model.", images, "resnet50", True)
        eng = model.image_encoder._engine
        feats_k, _ = eng.forward(images.cuda(), True, False)
        from mdhs_b200 import ops
        for name in ("layer1", "layer2", "layer3", "layer4"):
            x2d, H, W, C = feats_k[name]
            print(hw, name, "rel", rel(ops.nhwc_bf16_to_nchw_f32(x2d, B, H, W, C), f[name]), "rows/channel", B * H * W)
        # what torch's own bf16 autocast does on the same oracle (context for the tolerance)
        sdc = {k: v.cuda() for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            import oracle.port as P
            # port builds a CPU mask tensor; run only the trunk under autocast
            fa = P.resnet_features(sdc, "image_encoder.model.", images.cuda(), "resnet50", True)
        print(hw, "torch-autocast layer4 rel", rel(fa["layer4"].float(), f["layer4"]))
    print(hw, "tokens", rel(tok, otok), "text", rel(txt, otxt), "feats", rel(feats, ofeat), "logits", rel(logits, ologit))
