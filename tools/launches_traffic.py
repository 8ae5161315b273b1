"""ncu launch list with per-launch DRAM bytes -> per-kernel table (markdown) + the GEMM traffic record bench.py reads.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
  python tools/launches_traffic.py gpurun_out/launches.csv profiles/r02_launches_step_b128_summary.md profiles/r02_gemm_dram_traffic.json
"""
import collections
import csv
import json
import re
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, md_out, json_out):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = row["ID"]
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        d = per.setdefault(key, {"name": name, "us": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        m = row["Metric Name"]
        if m.startswith("gpu__time_duration"):
            d["us"] = v
        elif m.startswith("dram__bytes_read"):
            d["rd"] = v
        elif m.startswith("dram__bytes_write"):
            d["wr"] = v
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in per.values():
        a = agg[d["name"]]
        a[0] += 1
        a[1] += d["us"]
        a[2] += d["rd"]
        a[3] += d["wr"]
    tot = sum(a[1] for a in agg.values())
    out = [f"total kernel time {tot / 1000:.2f} ms over {len(per)} launches (ncu: cold-cache, serialised -- compare SHARES)\n",
           "| kernel | launches | total us | share | DRAM read MB | DRAM write MB | GB/s |", "|---|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if a[1] / tot < 0.0005:
            continue
        out.append(f"| `{k[:84]}` | {a[0]} | {a[1]:.0f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e6:.0f} | {a[3] / 1e6:.0f} | "
                   f"{(a[2] + a[3]) / max(a[1], 1e-9) / 1e3:.0f} |")
    open(md_out, "w").write("\n".join(out) + "\n")
    g = [d for d in per.values() if d["name"].startswith("gemm_tc_kernel")]
    rec = {"source": path.split("/")[-1], "launches_per_step": len(g), "dram_bytes_per_step": sum(d["rd"] + d["wr"] for d in g),
           "dram_read_bytes_per_step": sum(d["rd"] for d in g), "dram_write_bytes_per_step": sum(d["wr"] for d in g),
           "gemm_us_per_step_ncu": sum(d["us"] for d in g), "step_us_ncu": tot}
    json.dump(rec, open(json_out, "w"), indent=1)
    print("\n".join(out[:14]))
    print(json.dumps(rec))


if __name__ == "__main__":
    main(*sys.argv[1:4])
