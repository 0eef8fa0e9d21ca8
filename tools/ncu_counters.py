"""Per-launch ncu counters of the hot kernels -> profiles/ncu_counters.json, stamped with the hash of the kernel sources.

    python tools/ncu_counters.py <config> <report.ncu-rep> [<config> <report.ncu-rep> ...]

bench.py quotes these figures (DRAM traffic, warp instructions) next to its live timings, and only while the sources that
are compiled into libslcu.so are the ones the capture was made on (bench.kernel_source_sha)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

WANT = {"smsp__inst_executed.sum": "inst_executed", "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
        "gpu__time_duration.sum": "duration_ns", "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "lts__t_bytes.sum": "l2_bytes", "l1tex__t_bytes.sum": "l1_bytes"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "msecond": 1e6, "usecond": 1e3, "nsecond": 1.0, "second": 1e9}


def read(rep):
    """rep: an .ncu-rep, or the CSV of `ncu -i <rep> --page raw --csv` (written on the GPU box when the reports themselves are too
    large to bring back)"""
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = {}
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        short = name.split("(")[0].replace("void ", "").replace("sld::", "")
        ent = {}
        for k, dst in WANT.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    ent[dst] = float(vals[i].replace(",", "")) * SCALE.get(units[i], 1.0)
                except ValueError:
                    pass
        res[short] = ent
    return res


def main():
    path = os.path.join(ROOT, "profiles", "ncu_counters.json")
    sha = bench.kernel_source_sha()
    db = {"source_sha": sha, "configs": {}}
    if os.path.exists(path):
        old = json.load(open(path))
        if old.get("source_sha") == sha:
            db = old
    args = sys.argv[1:]
    for cfg, rep in zip(args[0::2], args[1::2]):
        db["configs"].setdefault(cfg, {}).update(read(rep))
    db["how"] = "ncu --set full --clock-control none -k <kernel> -c 1 on `python bench.py --config <config>` (one launch of the steady state)"
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(db, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
