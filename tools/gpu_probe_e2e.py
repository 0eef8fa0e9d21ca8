"""GPU-box probe: where the end-to-end frame time of a config goes (submission, device, D2H).  Not a bench value.
python tools/gpu_probe_e2e.py [config ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench as B
from softlight_b200 import capi as A


def main():
    for cfg in sys.argv[1:] or ["c3_indoor2m"]:
        h = B.Harness(cfg, 0, 1, 0, None, 8)
        ctx, inst, scene = h.ctx, h.inst, h.scene
        for i in range(30):
            h.render_frame(i)
        ctx.sync()
        cbytes = scene.width * scene.height * A.bytes_per_texel(scene.color_type)
        out_pin = ctx.host_array((cbytes,), np.uint8)
        sub, dev, d2h = [], [], []
        for i in range(20):
            ctx.sync()
            t0 = time.perf_counter()
            h.render_frame(i)
            t1 = time.perf_counter()
            ctx.sync()
            t2 = time.perf_counter()
            ctx.texture_read_into(inst.state["color"], out_pin.ctypes.data, cbytes)
            t3 = time.perf_counter()
            sub.append((t1 - t0) * 1e3); dev.append((t2 - t0) * 1e3); d2h.append((t3 - t2) * 1e3)
        print(f"{cfg}: submit {np.median(sub):.3f} ms, submit+device {np.median(dev):.3f} ms, D2H {cbytes / 1e6:.1f} MB {np.median(d2h):.3f} ms "
              f"({cbytes / np.median(d2h) / 1e6:.1f} GB/s)", flush=True)
        h.close()


if __name__ == "__main__":
    main()
