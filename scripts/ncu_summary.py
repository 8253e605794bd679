"""Condense an .ncu-rep (read here with `ncu -i`) into the handful of metrics DESIGN.md / the judge cite."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg.per_second"]
STALL = "smsp__average_warps_issue_stalled_"

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"== {d.get('Kernel Name', '?')}  (ID {d.get('ID')})")
    for h, u in zip(hdr, units):
        if h in KEYS:
            print(f"  {h:72s} {d[h]:>18s} {u}")
    stalls = sorted(((float(d[h]), h[len(STALL):-len('_per_issue_active.ratio')]) for h in hdr
                     if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and d[h]), reverse=True)
    print("  top stall reasons (warps per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:6]))
