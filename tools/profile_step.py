"""One eager training step between cudaProfilerStart/Stop (target for `ncu --profile-from-start off`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mdhs_b200  # noqa
from mdhs_b200.train import Trainer
from refutil import bert_dir, quiet
from bench import synthetic_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
fusion = sys.argv[2] if len(sys.argv) > 2 else "basic"
with quiet():
    model = mdhs_b200.MultimodalBaselineModel(num_classes=7, hidden_dim=256, dropout=0.2, pretrained_image=False,
                                              image_weights_path=None, text_model_name=bert_dir(), num_heads=8,
                                              image_backbone="resnet50", classifier_type="mlp", fusion_type=fusion).cuda()
tr = Trainer(model)
batch = [t.cuda() for t in synthetic_batch(B, 64, 7)]
for _ in range(2):
    tr.step(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
