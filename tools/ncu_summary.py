"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_c2.csv profiles/r01_launches_c2.txt
  python tools/ncu_summary.py full gpurun_out/raster_c2.ncu-rep profiles/r01_k_raster_c2.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"]
        a = agg.setdefault(k, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (source: {src})\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':70s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'avg_us':>9s}  last grid/block\n")
        for k, a in agg.items():
            f.write(f"{k[:70]:70s} {a[0]:8d} {a[1] / 1e3:10.1f} {a[1] / tot * 100:6.1f}% {a[1] / a[0] / 1e3:9.1f}  {a[2]} {a[3]}\n")
    print(open(dst).read())


def full(src, dst):
    # src: an .ncu-rep, or the CSV of `ncu -i <rep> --page raw --csv` written on the GPU box (the reports are too large to bring back)
    out = open(src).read() if src.endswith(".csv") else subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  (source: {src})\n")
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")]
            f.write(f"\n## {name}\n")
            for k in KEEP:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:85s} {vals[i]:>18s} {units[i]}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
