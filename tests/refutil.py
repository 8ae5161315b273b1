"""Helpers shared by the oracle tests and the golden-fixture generator (dev container only parts are guarded)."""
import contextlib
import io
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the reference itself (dev container) or its verbatim vendored copy (oracle/make_ref.py -> oracle/_ref, travels to the GPU box)
_CANDIDATES = ([os.environ["MDHS_REF_ROOT"]] if os.environ.get("MDHS_REF_ROOT") else []) + [
    "/root/reference", os.path.join(_ROOT, "oracle", "_ref")]
REF_ROOT = next((c for c in _CANDIDATES if os.path.exists(os.path.join(c, "model.py"))), _CANDIDATES[0])
BERT_DIR = os.environ.get("MDHS_BERT_DIR", "/tmp/mdhs_bert_base")


def have_reference():
    return os.path.isdir(REF_ROOT) and os.path.exists(os.path.join(REF_ROOT, "model.py"))


def bert_dir():
    """A local HF directory holding a randomly initialised bert-base-uncased (config defaults), built once."""
    if not os.path.exists(os.path.join(BERT_DIR, "config.json")):
        from transformers import BertConfig, BertModel
        torch.manual_seed(0)
        os.makedirs(BERT_DIR, exist_ok=True)
        BertModel(BertConfig()).save_pretrained(BERT_DIR)
    return BERT_DIR


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def build_reference_model(fusion="basic", head="mlp", num_classes=7, backbone="resnet50", gate=False, **kw):
    """The reference's MultimodalBaselineModel, with the ResNet-50 shim of SURVEY.md section 8c step 3
    (encoder.py:31-33 only admits resnet18/34)."""
    import torch.nn as nn
    import torchvision
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with quiet():
        from model import MultimodalBaselineModel as RefModel
        ref_backbone = "resnet18" if backbone == "resnet50" else backbone
        m = RefModel(num_classes=num_classes, hidden_dim=256, dropout=0.2, pretrained_image=False, image_weights_path=None,
                     text_model_name=bert_dir(), num_heads=8, image_backbone=ref_backbone, classifier_type=head,
                     fusion_type=fusion, gate_enabled=gate, **kw)
    if backbone == "resnet50":
        enc = m.image_encoder
        net = torchvision.models.resnet50(weights=None)
        net.fc = nn.Identity()
        enc.model = net
        enc.stem = nn.Sequential(net.conv1, net.bn1, net.relu, net.maxpool)
        enc.layer1, enc.layer2, enc.layer3, enc.layer4 = net.layer1, net.layer2, net.layer3, net.layer4
        if enc.multi_scale:
            enc.proj2 = nn.Linear(512, 256)
            enc.proj3 = nn.Linear(1024, 256)
        enc.proj4 = nn.Linear(2048, 256)
    return m


def build_ours(fusion="basic", head="mlp", num_classes=7, backbone="resnet50", gate=False, **kw):
    import mdhs_b200
    with quiet():
        return mdhs_b200.MultimodalBaselineModel(
            num_classes=num_classes, hidden_dim=256, dropout=0.2, pretrained_image=False, image_weights_path=None,
            text_model_name=bert_dir(), num_heads=8, image_backbone=backbone, classifier_type=head, fusion_type=fusion,
            gate_enabled=gate, **kw)


def build_reference_connext(variant="tiny", num_labels=7):
    """The reference's OurClassfierConvnextV2 (torchvision branch) with the recipe of SURVEY.md section 8c step 5:
    BertModel.from_pretrained is pointed at the local random-init bert-base; convnext_base is swapped for `variant`."""
    import torchvision
    import transformers
    cn = os.path.join(REF_ROOT, "ConNexT")
    if cn not in sys.path:
        sys.path.insert(0, cn)
    orig_fp = transformers.BertModel.from_pretrained
    orig_base = torchvision.models.convnext_base
    try:
        transformers.BertModel.from_pretrained = classmethod(lambda cls, *a, **k: orig_fp.__func__(cls, bert_dir()))
        torchvision.models.convnext_base = lambda weights=None: getattr(torchvision.models, f"convnext_{variant}")(weights=None)
        with quiet():
            import importlib
            om = importlib.import_module("models.ourmodel")
            m = om.OurClassfierConvnextV2(num_labels=num_labels, pretrained=False, pretrained_path=None)
            c_last = m.image_encoder[-1][-1].block[0].weight.shape[0]
            if c_last != 1024:   # the reference hard-codes the Base width for its 1x1 reduction (ourmodel.py:62)
                m.conv = torch.nn.Conv2d(c_last, 768, kernel_size=1)
    finally:
        transformers.BertModel.from_pretrained = orig_fp
        torchvision.models.convnext_base = orig_base
    return m


def build_ours_connext(variant="tiny", num_labels=7):
    import mdhs_b200  # noqa: F401
    from mdhs_b200.connext.ourmodel import OurClassfierConvnextV2
    with quiet():
        return OurClassfierConvnextV2(num_labels=num_labels, pretrained=False, pretrained_path=None, bert_path=bert_dir(),
                                      variant=variant)
