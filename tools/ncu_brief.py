"""Short digest of an `ncu --page raw --csv` export: duration, occupancy, issue rate, pipe utilisation, top stalls."""
import csv
import sys

for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    d = dict(zip(rows[0], zip(rows[2], rows[1])))
    print("==", f, d["Kernel Name"][0][:70])
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
    for k in keys:
        if k in d:
            print("  %-82s %s %s" % (k, d[k][0], d[k][1]))
    for k, v in d.items():
        if "tensor" in k and "pct" in k and k not in keys:
            try:
                if float(v[0]) > 1:
                    print("  %-82s %s" % (k, v[0]))
            except ValueError:
                pass
    st = []
    for k, v in d.items():
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
            try:
                st.append((float(v[0]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("  stalls/issue:", ", ".join(f"{n} {x:.2f}" for x, n in sorted(st, reverse=True)[:7]))
