"""GPU-box probe: per-frame device times of a config the way bench.py's other_config_leg measures them.  Not a bench value."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B

for cfg in sys.argv[1:] or ["c4_skinned100"]:
    h = B.Harness(cfg, 0, 1, 0, None, 8)
    for rep in range(2):
        for i in range(40):
            h.render_frame(i)
        h.ctx.sync()
        ms = []
        for i in range(12):
            h.flush_l2()
            h.ctx.timer_begin()
            h.render_frame(i)
            ms.append(round(h.ctx.timer_end(), 3))
        print(cfg, rep, ms, flush=True)
    h.close()
