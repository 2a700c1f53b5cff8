"""Selected metrics of an ncu report (one kernel launch) as 'metric<TAB>value<TAB>unit' lines, plus the stall-reason shares.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/rNN_<kernel>_ncu_summary.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("Kernel Name", "gpu__time_duration", "dram__bytes", "dram__cycles_active", "gpu__dram_throughput", "launch__", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__inst_executed.avg.per_cycle", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct",
        "sm__throughput.avg.pct", "sm__warps_active.avg", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled")
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k) for k in keep) and not h.endswith(".max") and not h.endswith(".min") and "peak_sustained" not in h.split(".")[-1] or h == "Kernel Name":
        print(f"{h}\t{v}\t{u}")
stall = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v.replace(",", "") or 0) for h, v in zip(hdr, vals) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
tot = sum(stall.values()) or 1
print("# stall samples (share of all warp samples):")
for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:14]:
    print(f"#   {k:28s} {v:12.0f}  {v / tot:.3f}")
