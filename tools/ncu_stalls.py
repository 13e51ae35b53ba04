"""Summarise an .ncu-rep: duration, pipe utilisation, stall reasons.  Usage: python tools/ncu_stalls.py report.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(h, r))
    print("kernel:", d.get("Kernel Name", "")[:90])
    for k in ("gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
              "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"):
        if k in d:
            print(f"  {k} = {d[k]} {u[h.index(k)]}")
    st = []
    for k, v in d.items():
        if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                st.append((float(v.replace(",", "")), k.split("stalled_")[1]))
            except ValueError:
                pass
    tot = sum(x for x, _ in st) or 1.0
    print("  stalls:", ", ".join(f"{k} {100 * x / tot:.1f}%" for x, k in sorted(st, reverse=True)[:7]))
