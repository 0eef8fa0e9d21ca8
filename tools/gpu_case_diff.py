"""Diagnose one parity case on a GPU box: python tools/gpu_case_diff.py <case name from tests/cases.py> [...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
from softlight_b200 import capi as A, scenes as S

frames = np.load(os.path.join(ROOT, "tests", "golden", "golden_frames.npz"))
for name in sys.argv[1:]:
    scene = (cases.CASES.get(name) or cases.BIG_CASES[name])()
    ctx = A.Context("cuda"); ctx.set_rcp_table(frames["rcp_table"]); ctx.reset_stats()
    inst = S.SceneInstance(scene, ctx); inst.render(); ctx.sync()
    cg, dg, sg = inst.read_color(), inst.read_depth(), ctx.get_stats()
    err = ctx.f["last_error"](ctx.h)
    co, do, so = cases.render("port", scene, rcp_table=frames["rcp_table"])
    dc = (cases.bits(cg) != cases.bits(co)) if cg.dtype.kind == "f" else (cg != co)
    dpx = dc.any(axis=-1)
    dd = cases.bits(dg) != cases.bits(do)
    print(name, "cuda", sg, "| port", so, "| last_error:", err.decode() if err else "")
    print("  colour px differing:", int(dpx.sum()), " depth texels differing:", int(dd.sum()))
    ys, xs = np.nonzero(dpx | dd)
    if len(xs):
        print("  bbox x", xs.min(), xs.max(), "y", ys.min(), ys.max())
        for y, x in list(zip(ys, xs))[:10]:
            print("   ", x, y, cg[y, x], co[y, x], dg[y, x], do[y, x])
        print("  x histogram (per 500):", np.bincount(xs // 500).tolist())
