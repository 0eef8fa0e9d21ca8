"""GPU-box diagnostic: every parity case, CUDA vs the CPU oracle (pinned rcp table), compact diff report.
python tools/gpu_diag.py [case ...]   (test infrastructure: uses oracle/ as the checker)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases

def main():
    frames = np.load(os.path.join(ROOT, "tests", "golden", "golden_frames.npz"))
    rcp = frames["rcp_table"]
    names = sys.argv[1:] or sorted(cases.CASES)
    bad = 0
    for n in names:
        sc = (cases.CASES.get(n) or cases.BIG_CASES[n])()
        cg, dg, sg = cases.render("cuda", sc, rcp_table=rcp)
        cg2, dg2, sg2 = cases.render("cuda", sc, rcp_table=rcp)
        co, do, so = cases.render("port", sc, rcp_table=rcp)
        det = np.array_equal(cases.bits(cg), cases.bits(cg2)) and np.array_equal(cases.bits(dg), cases.bits(dg2))
        dc = (cases.bits(cg) != cases.bits(co)).any(axis=-1)
        dd = cases.bits(dg) != cases.bits(do)
        ok = not dc.any() and not dd.any() and all(sg[k] == so[k] for k in ("fragsShaded", "trisBinned", "spans"))
        print(f"{'ok ' if ok else 'BAD'} {n}: colour px {int(dc.sum())} depth {int(dd.sum())} frags {sg['fragsShaded']}/{so['fragsShaded']} "
              f"spans {sg['spans']}/{so['spans']} binned {sg['trisBinned']}/{so['trisBinned']} deterministic={det}", flush=True)
        if not ok:
            bad += 1
            ys, xs = np.nonzero(dc | dd)
            for y, x in list(zip(ys, xs))[:12]:
                print(f"    ({x},{y}) colour {cg[y, x].tolist()} vs {co[y, x].tolist()}  depth {dg[y, x]!r} vs {do[y, x]!r}")
            rows = sorted(set(ys.tolist()))
            print("    rows:", rows[:40])
    print("BAD CASES:", bad)
    return 0

if __name__ == "__main__":
    sys.exit(main())
