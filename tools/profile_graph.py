"""Kernel timeline of the CAPTURED training step (CUDA-graph replays) through torch.profiler / CUPTI: per-kernel warm
durations and how much of the step is inter-kernel gap.  Writes gpurun_out/graph_timeline.json (summary)."""
import collections
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mdhs_b200  # noqa
from mdhs_b200.train import Trainer
from refutil import bert_dir, quiet
from bench import synthetic_batch

with quiet():
    model = mdhs_b200.MultimodalBaselineModel(num_classes=7, hidden_dim=256, dropout=0.2, pretrained_image=False,
                                              image_weights_path=None, text_model_name=bert_dir(), num_heads=8,
                                              image_backbone="resnet50", classifier_type="mlp", fusion_type="basic").cuda()
tr = Trainer(model)
batch = [t.cuda() for t in synthetic_batch(128, 64, 7)]
tr.capture(*batch, warmup=3)
for _ in range(5):
    tr.replay()
torch.cuda.synchronize()
REPS = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(REPS):
        tr.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name and "Memcpy" not in e.name]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.defaultdict(lambda: [0, 0.0])
busy = 0.0
for e in evs:
    d = e.time_range.end - e.time_range.start
    k = e.name.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    k = k.split("(")[0]
    agg[k][0] += 1
    agg[k][1] += d
    busy += d
span = evs[-1].time_range.end - evs[0].time_range.start
out = {"reps": REPS, "span_us_per_step": span / REPS, "busy_us_per_step": busy / REPS, "kernels_per_step": len(evs) / REPS,
       "top": sorted(([k, v[0] / REPS, v[1] / REPS] for k, v in agg.items()), key=lambda r: -r[2])[:40]}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "graph_timeline.json"), "w"), indent=1)
print(json.dumps({k: out[k] for k in ("span_us_per_step", "busy_us_per_step", "kernels_per_step")}))
for k, n, us in out["top"][:30]:
    print(f"{us:9.1f} us  {n:6.1f}x  {k[:90]}")
