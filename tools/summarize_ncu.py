"""Summaries of ncu output for profiles/: (1) a `--metrics gpu__time_duration.sum --csv` launch list -> per-kernel table,
(2) `--page raw --csv` exports of `--set full` captures -> the handful of counters DESIGN.md argues from.
  python tools/summarize_ncu.py launches <launches.csv>
  python tools/summarize_ncu.py raw <a.raw.csv> [b.raw.csv ...]
"""
import collections
import csv
import re
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subunit_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total kernel time {tot / 1000:.2f} ms over {sum(v[0] for v in agg.values())} launches\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.0005:
            continue
        print(f"| `{k[:80]}` | {v[0]} | {v[1]:.0f} | {100 * v[1] / tot:.1f}% |")


def raw(paths):
    for f in paths:
        rows = list(csv.reader(open(f)))
        d = dict(zip(rows[0], zip(rows[2], rows[1])))
        print(f"\n### {f.split('/')[-1]}: `{d.get('Kernel Name', ('', ''))[0][:90]}`\n")
        for k in KEYS + [k for k in d if "tensor" in k and k not in KEYS][:12]:
            if k in d:
                print(f"- {k}: {d[k][0]} {d[k][1]}")


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else raw)(sys.argv[2] if sys.argv[1] == "launches" else sys.argv[2:])
