"""Where the warp-stall samples of an `ncu --page source --csv` export fall: by opcode, and the hottest instructions."""
import csv
import sys
from collections import Counter

f = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(f)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        data.append((int(r[col["# Samples"]]), int(r[col["Instructions Executed"]]), r[col["Source"]], r))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
inst = sum(d[1] for d in data)
print(f"{len(data)} SASS lines, {tot} samples, {inst} warp instructions executed")
by_op, by_op_inst = Counter(), Counter()
for n, ie, src, _ in data:
    t = src.split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "")
    by_op[op] += n
    by_op_inst[op] += ie
print("-- by opcode: samples %, executed %")
for op, n in by_op.most_common(top):
    print(f"{op:28s} {100 * n / tot:5.1f}%  {100 * by_op_inst[op] / inst:5.1f}%")
print("-- hottest instructions")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for i, (n, ie, src, r) in enumerate(data):
    pass
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
for i in sorted(order):
    n, ie, src, r = data[i]
    why = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {100 * n / tot:5.1f}% x{ie:<9d} {src[:70]:70s} {why}")
