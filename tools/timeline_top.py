"""Per-kernel totals of a bench.py --timeline file: launches, total us, share."""
import json
import sys
from collections import defaultdict

d = json.load(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
agg = defaultdict(lambda: [0, 0.0])
for e in d["kernels"]:
    agg[e["name"]][0] += 1
    agg[e["name"]][1] += e["dur_us"]
tot = sum(v[1] for v in agg.values())
print(f"{sys.argv[1]}: step {d['ms_per_step']} ms, kernel time {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {n[:72]:72s} {c:5d} {t:9.1f} {100 * t / tot:5.1f}%")
