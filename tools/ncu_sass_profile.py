"""Where a kernel's issue slots and stall samples go, by SASS address range (ncu report with --import-source on):
   python tools/ncu_sass_profile.py gpurun_out/x.ncu-rep [buckets]
Prints, per bucket of consecutive SASS instructions: static size, share of executed warp instructions, share of
stall samples, average active threads, and the first source-less mnemonic as a landmark."""
import csv, subprocess, sys
rep = sys.argv[1]; nb = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; recs = []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < 10 or not r[0].startswith("0x"): continue
    d = dict(zip(hdr, r))
    f = lambda k: float(d.get(k, "0").replace(",", "") or 0)
    recs.append((int(r[0], 16), r[1].strip(), f("# Samples"), f("Instructions Executed"), f("Thread Instructions Executed")))
n = len(recs); ts = sum(x[2] for x in recs) or 1; ti = sum(x[3] for x in recs) or 1
print(f"{n} SASS instructions ({n*16/1024:.0f} KiB), {ti:.3g} warp instructions, {ts:.0f} samples")
per = (n + nb - 1) // nb
for b in range(0, n, per):
    ch = recs[b:b+per]
    s = sum(x[2] for x in ch); i = sum(x[3] for x in ch); t = sum(x[4] for x in ch)
    calls = [x[1] for x in ch if x[1].startswith(("CALL", "RET", "EXIT", "BAR", "MATCH", "ATOM", "WARPSYNC"))][:3]
    print(f"[{b:5d}..{b+len(ch)-1:5d}] inst {i/ti*100:5.1f}%  samples {s/ts*100:5.1f}%  thr/inst {t/max(i,1):5.1f}  {'; '.join(c[:40] for c in calls)}")
