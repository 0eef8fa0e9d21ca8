"""Top source lines of a kernel from an ncu report captured with --import-source on (needs -lineinfo):
   python tools/ncu_hot_lines.py gpurun_out/x.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = ""
recs = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        cur_file = r[1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or r[0] in ("", "Kernel Name"):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr[4:], r[4:]))
    def g(k):
        try: return float(d.get(k, "0").replace(",", ""))
        except ValueError: return 0.0
    recs.append((cur_file.split("/")[-1], line, r[1].strip()[:90], g("# Samples"), g("Instructions Executed"), g("stall_no_inst"), g("stall_long_sb"), g("stall_wait"), g("stall_short_sb"), g("stall_branch_resolving")))
tot_s = sum(x[3] for x in recs) or 1; tot_i = sum(x[4] for x in recs) or 1
print(f"total samples {tot_s:.0f}  total warp-instructions {tot_i:.0f}")
print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s} {'no_inst':>7s} {'long_sb':>7s} {'wait':>6s} {'short':>6s} {'branch':>6s}  source")
key = 4 if (len(sys.argv) > 3 and sys.argv[3] == "inst") else 3
for x in sorted(recs, key=lambda x: -x[key])[:top]:
    print(f"{x[0][:20]+':'+str(x[1]):28s} {x[3]/tot_s*100:6.2f} {x[4]/tot_i*100:6.2f} {x[5]:7.0f} {x[6]:7.0f} {x[7]:6.0f} {x[8]:6.0f} {x[9]:6.0f}  {x[2]}")
