#!/usr/bin/env python
"""bench.py -- the headline measurement of softlight-b200 (contract: task prompt section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--config NAME]

A "step" is one frame of the workload: clear_framebuffer + every draw of the frame
(+ the sort-first composite when N > 1).  At N=1 the workload is BASELINE.json
configs[1] (10k instanced meshes, 1920x1080), SURVEY.md section 8d "C2".

One JSON line is printed by rank 0:
  value     Mtris/s, whole job, inputs resident in HBM, device time (CUDA events on the library's stream)
  e2e       same metric through the C ABI with host buffers: per-frame H2D of the frame's inputs from pinned
            memory and D2H of the colour attachment inside the timed region (host wall clock)
  roofline  the dominant kernel's algorithmic bytes / its event-measured duration vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the compiled reference (oracle/_ref/libslref_fast.so, shipped flags, all host threads) on the same frame
--impl reference times only that CPU path (the driver computes the ratio itself).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from softlight_b200 import capi as A  # noqa: E402
from softlight_b200 import scenes as S  # noqa: E402

METRIC = "Mtris/s"
DEFAULT_CONFIG = "c2_instancing10k"


# ----------------------------------------------------------------------------- helpers
def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(scene_name: str, scene: S.Scene, frags: int) -> dict:
    """bytes the frame must move, per kernel class.  The fragment stage is two
    kernels: "raster" = k_resolve (depth traffic), "shade" = k_shade (colour + texture traffic); the fragment lists that
    travel between them are this design's own intermediate and are NOT counted as algorithmic bytes."""
    cbpt = A.bytes_per_texel(scene.color_type)
    dbpt = A.bytes_per_texel(scene.depth_type)
    px = scene.width * scene.height
    out = {"clear": px * (cbpt + dbpt)}
    if scene_name.startswith("c2"):
        # one 16x32 sphere reused by every instance: u16 indices, 32 B vertices, one mat4 per instance
        out["vertex"] = 3072 * 2 + 561 * 32 + (scene.tris_per_frame // 1024) * 64
        # per shaded fragment: depth read+write, RGBA8 read+write (blend), 4 bilinear taps of RGB8 (reuse = 0: upper figure)
        out["raster"] = frags * (2 * dbpt)
        out["shade"] = frags * (2 * cbpt + 4 * 3)
    elif scene_name.startswith("c3"):
        # per draw: u32 indices + the room's 24-byte vertices + the 1 KiB uniform block; per shaded fragment: depth read +
        # write, RGBA8 write, 4 bilinear taps of RGBA8 (reuse = 0: upper figure)
        tris_per_draw = scene.tris_per_frame // scene.draws_per_frame
        out["vertex"] = scene.draws_per_frame * (tris_per_draw * 3 * 4 + (tris_per_draw // 2 + 2 * 6 * 27) * 24 + 1024)
        out["raster"] = frags * (2 * dbpt)
        out["shade"] = frags * (cbpt + 4 * 4)
    elif scene_name.startswith("c1"):
        # u32 indices + every vertex once (32 B) ; per shaded fragment: depth read + write, RGBA8 write
        nverts = (int(round((scene.tris_per_frame / 2) ** 0.5)) + 1) ** 2
        out["vertex"] = scene.tris_per_frame * 3 * 4 + nverts * 32
        out["raster"] = frags * (2 * dbpt)
        out["shade"] = frags * cbpt
    elif scene_name.startswith("c4"):
        # per draw: u32 indices + the character's 36-byte vertices + 64 bone matrices + the uniform block; per shaded fragment:
        # depth read + write, RGBA8 read + write (SL_BLEND_ALPHA), 4 bilinear taps of RGBA8 (reuse = 0: upper figure)
        tris_per_draw = scene.tris_per_frame // scene.draws_per_frame
        out["vertex"] = scene.draws_per_frame * (tris_per_draw * 3 * 4 + (tris_per_draw // 2 + 160) * 36 + 64 * 64 + 1024)
        out["raster"] = frags * (2 * dbpt)
        out["shade"] = frags * (2 * cbpt + 4 * 4)
    elif scene_name.startswith("c5"):
        # 36 non-indexed 12-byte vertices; per fragment that reaches the blend: RGBA_FLOAT read + write (premultiplied alpha), no
        # depth traffic (test and mask off).  The ray march's volume taps (<= 32 nearest + 256 x 8 + 3 x 8 per lit step, R8) hit
        # a 16.8 MB L2-resident volume: the shader is issue-bound, the HBM figure is the framebuffer traffic only
        out["vertex"] = 36 * 12 + 1024
        out["raster"] = 0
        out["shade"] = frags * (2 * cbpt)
    else:
        out["vertex"] = 0
        out["raster"] = frags * (2 * dbpt)
        out["shade"] = frags * cbpt
    out["frame"] = sum(out.values())
    return out


def ncu_traffic(dom: str, config: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the newest committed `ncu --set full` summary
    (profiles/rNN?_k_<kernel>_<cfg>.txt, written by tools/ncu_summary.py), per launch; None when there is no capture."""
    kern = {"raster": "resolve", "shade": "shade", "vertex": "vertex", "span_count": "span", "span_fill": "span"}.get(dom)
    cfg = config.split("_")[0]
    if kern is None:
        return None, None
    pdir = os.path.join(ROOT, "profiles")
    try:
        cands = sorted(f for f in os.listdir(pdir) if f.endswith(f"_k_{kern}_{cfg}.txt"))
    except OSError:
        return None, None
    if not cands:
        return None, None
    total = 0.0
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for ln in open(os.path.join(pdir, cands[-1])):
        f = ln.split()
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(f[1]) * unit.get(f[2], 1.0)
    return (total or None), "profiles/" + cands[-1]


def pinned_array(shape, dtype):
    """Host staging in pinned memory when torch+CUDA are present (plumbing only)."""
    try:
        import torch

        t = torch.empty(int(np.prod(shape)) * np.dtype(dtype).itemsize, dtype=torch.uint8, pin_memory=True)
        arr = t.numpy().view(dtype).reshape(shape)
        arr._keep = t if False else None
        return arr, t
    except Exception:
        return np.zeros(shape, dtype=dtype), None


# ----------------------------------------------------------------------------- arms
def run_reference(args, rank, world):
    """The reference's own CPU implementation on this box's host cores (shipped flags, all threads)."""
    if rank != 0:
        return None
    scene = S.SCENES[args.config]()
    cores = os.cpu_count() or 1
    kind = "reference" if os.path.exists(A.LIB_PATHS["ref_fast"]) else "port"
    impl = "ref_fast" if kind == "reference" else "port"
    threads = cores if kind == "reference" else 1
    A.load_library(impl)["reset_stats"](None)  # frag counting off
    ctx = A.Context(impl, threads=threads)
    inst = S.SceneInstance(scene, ctx)
    for _ in range(args.warmup):
        inst.render()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        inst.render()
    dt = time.perf_counter() - t0
    ms = dt * 1e3 / args.steps
    value = scene.tris_per_frame / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {scene.name}, {scene.tris_per_frame} tris/frame, {scene.draws_per_frame} draw(s), "
                               f"{scene.width}x{scene.height}", "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": threads, "kind": kind,
                         "sample": f"{args.steps} full frames of {args.config} ({ctx.impl_name()})"},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fps": 1e3 / ms,
    }
    return line


def cpu_baseline_leg(config: str, budget_s: float = 12.0) -> dict:
    scene = S.SCENES[config]()
    cores = os.cpu_count() or 1
    if os.path.exists(A.LIB_PATHS["ref_fast"]):
        impl, kind, threads = "ref_fast", "reference", cores
    else:
        impl, kind, threads = "port", "port", 1
    A.load_library(impl)["reset_stats"](None)
    ctx = A.Context(impl, threads=threads)
    inst = S.SceneInstance(scene, ctx)
    inst.render()
    n, t0 = 0, time.perf_counter()
    while True:
        inst.render()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 50:
            break
    ms = (time.perf_counter() - t0) * 1e3 / n
    return {"value": scene.tris_per_frame / (ms * 1e-3) / 1e6, "unit": METRIC, "cores": threads, "kind": kind,
            "ms_per_frame": ms, "sample": f"{n} full frames of {config} after 1 warm-up ({ctx.impl_name()}, {threads} threads)"}


def run_cuda(args, rank, world, local_rank, dist):
    scene = S.SCENES[args.config]()
    ctx = A.Context("cuda", device=local_rank)
    inst = S.SceneInstance(scene, ctx)
    sf, final = None, []
    if world > 1:
        # sort-first: this rank rasterises its interleaved row strips; composite = stores into rank 0's texture over NVLink
        from softlight_b200.sortfirst import SortFirst

        sf = SortFirst(ctx, dist, rank, world, strip_rows=args.strip_rows)
        final = [sf.make_target(scene.color_type, scene.width, scene.height) for _ in range(2)]  # double-buffered

    def render_frame(i):
        inst.render()
        if sf is not None:
            sf.composite(inst.state["color"], final[i % 2])

    # L2 flush buffer (written between timed frames, outside the timed region)
    import ctypes as C
    flush_tex = ctx.create_texture(A.SL_COLOR_RGBA_32U, 4096, 4096, 1)  # 256 MiB > 126 MB L2
    flush_fbo_depth = ctx.create_texture(A.SL_COLOR_R_FLOAT, 4, 4, 1)
    flush_fbo = ctx.create_framebuffer([flush_tex], flush_fbo_depth)

    def flush_l2():
        ctx.clear_color_buffer(flush_fbo, 0, (0.0, 0.0, 0.0, 0.0))

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    for i in range(max(args.warmup, 3)):
        render_frame(i)
    barrier()

    # ---- device-timed steps ------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    ctx.reset_stats()
    barrier()
    sampler.start()
    step_ms = []
    for i in range(args.steps):
        flush_l2()
        ctx.timer_begin()
        render_frame(i)
        step_ms.append(ctx.timer_end())
    barrier()
    clocks = sampler.stop()
    st = ctx.get_stats()
    frags_total = st["fragsShaded"]
    if dist is not None:
        # a frame is complete when the slowest rank's rows have landed: per-step MAX over ranks, then the sum
        import torch
        t = torch.tensor(step_ms, dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = [float(v) for v in t.tolist()]
        f = torch.tensor([frags_total], dtype=torch.int64, device=f"cuda:{local_rank}")
        dist.all_reduce(f, op=dist.ReduceOp.SUM)
        frags_total = int(f.item())
    total_ms = float(sum(step_ms))
    ms = total_ms / args.steps
    frags = frags_total // args.steps
    launches = st["kernelLaunches"] - args.steps  # minus the L2-flush clears
    value = scene.tris_per_frame / (ms * 1e-3) / 1e6

    # ---- per-kernel times for the roofline (separate pass; events around every launch) ----
    prof_frames = max(3, min(args.steps, 10))
    ctx.profile_begin()
    for i in range(prof_frames):
        render_frame(i)
    kt = ctx.profile_end()
    alg = algorithmic_bytes(args.config, scene, frags)
    peak, peak_src = measured_peaks()
    classes = {k: v for k, v in kt.items() if v["launches"]}
    dom = max(classes, key=lambda k: classes[k]["ms"])
    dom_ms = classes[dom]["ms"] / classes[dom]["launches"]
    launches_per_frame = classes[dom]["launches"] / prof_frames
    dom_bytes = alg.get(dom, 0) / max(launches_per_frame, 1)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    kernel_ms_per_frame = {k: v["ms"] / prof_frames for k, v in classes.items()}
    frame_kernel_ms = sum(kernel_ms_per_frame.values())
    traffic, traffic_src = ncu_traffic(dom, args.config)
    roofline = {
        "bound": "hbm", "kernel": {"raster": "k_resolve", "shade": "k_shade"}.get(dom, "k_" + dom), "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
        "kernel_ms_per_launch": dom_ms, "kernel_share_of_frame": kernel_ms_per_frame[dom] / frame_kernel_ms,
        "frame_algorithmic_bytes": alg["frame"], "frame_achieved_gbs": alg["frame"] / (ms * 1e-3) / 1e9,
        "frame_frac": alg["frame"] / (ms * 1e-3) / 1e9 / peak, "kernel_ms_per_frame": kernel_ms_per_frame,
    }

    # ---- end to end: host buffers, H2D of the frame inputs, D2H of the colour attachment ------
    per_frame = inst.state.get("per_frame", {})
    staged = {}
    h2d = 0
    for key, arr in per_frame.items():
        pin, keep = pinned_array(arr.shape, arr.dtype)
        pin[...] = arr
        staged[key] = (pin, keep)
        h2d += arr.nbytes
    cbytes = scene.width * scene.height * A.bytes_per_texel(scene.color_type)
    out_pin, out_keep = pinned_array((cbytes,), np.uint8)
    e2e_ms = []
    for i in range(args.steps + 2):
        flush_l2()
        ctx.sync()
        t0 = time.perf_counter()
        for key, (pin, _) in staged.items():
            ctx.vbo_assign_raw(inst.state[key], pin.ctypes.data, pin.nbytes)
        render_frame(i)
        if sf is not None:
            sf.frame_barrier()
            if rank == 0:
                ctx.texture_read_into(final[i % 2], out_pin.ctypes.data, cbytes)
        else:
            ctx.texture_read_into(inst.state["color"], out_pin.ctypes.data, cbytes)
        dt = (time.perf_counter() - t0) * 1e3
        if i >= 2:
            e2e_ms.append(dt)
    if dist is not None:
        import torch
        t = torch.tensor(e2e_ms, dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = [float(v) for v in t.tolist()]
    e2e_ms_avg = float(np.mean(e2e_ms))
    e2e = {"value": scene.tris_per_frame / (e2e_ms_avg * 1e-3) / 1e6, "unit": METRIC, "h2d_bytes_per_step": int((h2d + 1024) * world),
           "d2h_bytes_per_step": int(cbytes), "ms_per_step": e2e_ms_avg, "checksum": int(out_pin[::4097].astype(np.uint64).sum())}

    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {scene.name}, {scene.tris_per_frame} tris/frame, {scene.draws_per_frame} draw(s), "
                               f"{scene.width}x{scene.height}", "l2": "flushed between timed frames (256 MiB clear, untimed)",
                   "parallelism": f"sort-first x{world}" + (f", {args.strip_rows}-row strips interleaved over ranks, geometry replicated, "
                                                              "composite = peer stores into rank 0 over NVLink" if world > 1 else "")},
        "gfrags_per_s": frags / (ms * 1e-3) / 1e9, "fps": 1e3 / ms, "frags_per_frame": int(frags),
        "tris_binned_per_frame": int(st["trisBinned"] // args.steps), "spans_per_frame": int(st["spans"] // args.steps),
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roofline,
    }
    if rank == 0 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(args.config)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strip-rows", type=int, default=8, help="rows per ownership strip of the sort-first partition")
    args = ap.parse_args()

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)

    # stdout carries exactly ONE line (the JSON): anything native libraries print meanwhile (NCCL's version banner, ...)
    # goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        line = run_reference(args, rank, world)
        if line is not None:
            emit(line)
        return 0

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group(backend="nccl")
        dist = dist_mod
    line = run_cuda(args, rank, world, local_rank, dist)
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
