"""Compare the CUDA path with the checkers on a GPU box: python tools/gpu_compare.py [scene ...]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from softlight_b200 import capi as A, scenes as S

def bits(a):
    return a.view({2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize])

def run(impl, sc, rcp=None, reps=1):
    ctx = A.Context(impl)
    if rcp is not None and impl != "ref_strict":
        ctx.set_rcp_table(rcp)
    ctx.reset_stats()
    inst = S.SceneInstance(sc, ctx)
    ts = []
    for _ in range(reps):
        t = time.time(); inst.render(); ctx.sync(); ts.append(time.time() - t)
    st = ctx.get_stats()
    return inst.read_color(), inst.read_depth(), st, min(ts), ctx

def main():
    names = sys.argv[1:] or ["draw_test", "c1_small", "c2_small", "c1_sphere100k"]
    ok = True
    for n in names:
        sc = S.SCENES[n]()
        cg, dg, sg, tg, cu = run("cuda", sc, reps=3)
        rcp = cu.get_rcp_table()
        others = ["port"] + (["ref_strict"] if os.path.exists(A.LIB_PATHS["ref_strict"]) else [])
        for o in others:
            co, do, so, to, _ = run(o, sc, rcp)
            if cg.dtype.kind == "f":
                dc = (bits(cg) != bits(co)).astype(np.int64)
            else:
                dc = np.abs(cg.astype(np.int64) - co.astype(np.int64))
            ncol = int((dc.max(axis=-1) > 0).sum()); ndep = int((bits(dg) != bits(do)).sum())
            print(f"{n}: cuda vs {o}: colour px diff={ncol} (max {int(dc.max())}) depth diff={ndep} | frags {sg['fragsShaded']/3:.0f} vs {so['fragsShaded']} | spans {sg['spans']/3:.0f} vs {so['spans']} | binned {sg['trisBinned']/3:.0f} vs {so['trisBinned']} | t {tg*1e3:.2f} ms vs {to*1e3:.1f} ms", flush=True)
            if ncol or ndep:
                ok = False
                ys, xs = np.nonzero(dc.max(axis=-1) > 0)
                print("   colour diff bbox x", xs.min(), xs.max(), "y", ys.min(), ys.max())
                for y, x in list(zip(ys, xs))[:12]:
                    print("   colour", x, y, cg[y, x], co[y, x])
                ys, xs = np.nonzero(bits(dg) != bits(do))
                for y, x in list(zip(ys, xs))[:5]:
                    print("   depth", x, y, dg[y, x], do[y, x])
    print("ALL MATCH" if ok else "MISMATCH")
    return 0 if ok else 1

if __name__ == "__main__":
    sys.exit(main())
