#!/usr/bin/env python
"""bench.py -- the headline measurement of softlight-b200 (contract: task prompt section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--config NAME]

A "step" is one frame of the workload: clear_framebuffer + every draw of the frame
(+ the sort-first composite when N > 1).  The workload is the configuration BASELINE.json's
metric is quoted on ("... shaded at 4K"): configs[2], the 2M-triangle indoor scene with 256 textured
draws at 3840x2160 (SURVEY.md section 8d "C3"), at every N.

One JSON line is printed by rank 0:
  value     Mtris/s, whole job, inputs resident in HBM, device time (CUDA events on the library's stream)
  e2e       same metric through the C ABI with host buffers: per-frame H2D of the frame's inputs from pinned
            memory and D2H of the colour attachment inside the timed region (host wall clock)
  roofline  the dominant kernel against BOTH bounds of SURVEY.md section 8d: algorithmic bytes / event-measured
            duration vs MEASURED_PEAKS.json hbm_gbs, and warp instructions / duration vs 148 SMs x 4 issue slots x f_SM
  cpu_baseline  the compiled reference (oracle/_ref/libslref_fast.so, shipped flags, all host threads) on the same frame
  other_configs (N=1)  device ms + e2e ms of the other BASELINE configs (C1, C2, C4, C5-8K), C5-8K with its CPU baseline
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
DEFAULT_CONFIG = "c3_indoor2m"
OTHER_CONFIGS = ["c1_sphere100k", "c2_instancing10k", "c4_skinned100", "c5_volume8k"]


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
        self.times = []

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
            self.times.append(time.perf_counter())

    def stop(self, since=None) -> dict:
        """since: only samples that arrived after this perf_counter() time (the GPU has been under the frame load since then)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln, at in zip(self.lines, self.times):
            if since is not None and at < since:
                continue
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


def kernel_source_sha() -> str:
    """Hash of everything that is compiled into libslcu.so: ncu figures are only quoted for the sources they were captured on."""
    import hashlib

    from softlight_b200 import build as B

    h = hashlib.sha256()
    for f in B.cuda_sources():
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def ncu_counters(dom: str, config: str):
    """Per-launch ncu counters of the dominant kernel (dram bytes, warp instructions) from profiles/ncu_counters.json, written
    by tools/ncu_counters.py from an `ncu --set full` capture of this same command.  Entries are stamped with the hash of the
    kernel sources they were measured on and ignored when the sources have changed since."""
    kern = {"raster": "k_resolve", "shade": "k_shade", "vertex": "k_vertex", "span_count": "k_span<0>", "span_fill": "k_span<1>"}.get(dom)
    path = os.path.join(ROOT, "profiles", "ncu_counters.json")
    if kern is None or not os.path.exists(path):
        return None
    try:
        db = json.load(open(path))
    except Exception:
        return None
    if db.get("source_sha") != kernel_source_sha():
        return None
    ent = db.get("configs", {}).get(config, {})
    for name, v in ent.items():
        if name.startswith(kern):
            return dict(v, kernel=name, source="profiles/ncu_counters.json@" + db["source_sha"])
    return None


# ----------------------------------------------------------------------------- arms
def workload_config(config: str, scene: S.Scene) -> dict:
    """The `config` object of the JSON line: identical in both arms."""
    return {"workload": f"{config}: {scene.name}, {scene.tris_per_frame} tris/frame, {scene.draws_per_frame} draw(s), "
                        f"{scene.width}x{scene.height}"}


def cpu_context():
    cores = os.cpu_count() or 1
    if os.path.exists(A.LIB_PATHS["ref_fast"]):
        impl, kind, threads = "ref_fast", "reference", cores
    else:
        impl, kind, threads = "port", "port", 1
    A.load_library(impl)["reset_stats"](None)  # frag counting off
    return A.Context(impl, threads=threads), kind, threads


CPU_BUILD_NOTE = ("the unmodified reference compiled by oracle/Makefile with g++ -O2 and the project's shipped flags "
                  "(-mavx2 -mfma -ffast-math; its README names clang as the fastest compiler, which this image does not have)")


def run_reference(args, rank, world):
    """The reference's own CPU implementation on this box's host cores (shipped flags, all threads)."""
    if rank != 0:
        return None
    scene = S.SCENES[args.config]()
    ctx, kind, threads = cpu_context()
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
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.config, scene),
        "cpu_baseline": {"value": value, "unit": METRIC, "cores": threads, "kind": kind,
                         "sample": f"{args.steps} full frames of {args.config} ({ctx.impl_name()})", "build": CPU_BUILD_NOTE},
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fps": 1e3 / ms,
    }
    return line


def cpu_baseline_leg(config: str, budget_s: float = 12.0, max_frames: int = 50) -> dict:
    scene = S.SCENES[config]()
    ctx, kind, threads = cpu_context()
    inst = S.SceneInstance(scene, ctx)
    inst.render()
    n, t0 = 0, time.perf_counter()
    while True:
        inst.render()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= max_frames:
            break
    ms = (time.perf_counter() - t0) * 1e3 / n
    return {"value": scene.tris_per_frame / (ms * 1e-3) / 1e6, "unit": METRIC, "cores": threads, "kind": kind,
            "ms_per_frame": ms, "sample": f"{n} full frames of {config} after 1 warm-up ({ctx.impl_name()}, {threads} threads)",
            "build": CPU_BUILD_NOTE}


class Harness:
    """One scene on one context (+ the sort-first plumbing when world > 1) with the timing loops of the contract."""

    def __init__(self, config, rank, world, local_rank, dist, strip_rows):
        self.config, self.rank, self.world, self.local_rank, self.dist = config, rank, world, local_rank, dist
        self.scene = S.SCENES[config]()
        self.ctx = A.Context("cuda", device=local_rank)
        self.inst = S.SceneInstance(self.scene, self.ctx)
        self.sf, self.final = None, []
        if world > 1:
            # sort-first: this rank rasterises its interleaved row strips; composite = stores into rank 0's texture over NVLink
            from softlight_b200.sortfirst import SortFirst

            self.sf = SortFirst(self.ctx, dist, rank, world, strip_rows=strip_rows)
            if os.environ.get("SLCU_GEOM_SHARE", "1") != "0":
                # the ranks split the vertex stage and route the triangle records (two arenas, fixed record slots, <= 4 varyings)
                self.sf.share_geometry(int(self.scene.tris_per_frame * 1.3) * 246 + (64 << 20))
            self.final = [self.sf.make_target(self.scene.color_type, self.scene.width, self.scene.height) for _ in range(2)]
        # L2 flush buffer (written between timed frames, outside the timed region)
        ctx = self.ctx
        self.flush_tex = ctx.create_texture(A.SL_COLOR_RGBA_32U, 4096, 4096, 1)  # 256 MiB > 126 MB L2
        self.flush_fbo = ctx.create_framebuffer([self.flush_tex], ctx.create_texture(A.SL_COLOR_R_FLOAT, 4, 4, 1))
        self._torch_flush = None

    def render_frame(self, i):
        self.inst.render()
        if self.sf is not None:
            self.sf.composite(self.inst.state["color"], self.final[i % 2])

    def flush_l2(self):
        self.ctx.clear_color_buffer(self.flush_fbo, 0, (0.0, 0.0, 0.0, 0.0))

    def flush_l2_outside_library(self):
        """L2 flush that does not show up in the library's per-kernel timers (used by the per-kernel pass): a 256 MiB fill on
        torch's stream, bracketed by synchronisation."""
        try:
            import torch

            if self._torch_flush is None:
                self._torch_flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{self.local_rank}")
            self.ctx.sync()
            self._torch_flush.zero_()
            torch.cuda.synchronize(self.local_rank)
            return True
        except Exception:
            return False

    def barrier(self):
        self.ctx.sync()
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, values):
        if self.dist is None:
            return list(values)
        import torch

        t = torch.tensor(list(values), dtype=torch.float64, device=f"cuda:{self.local_rank}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def sum_over_ranks(self, value: int) -> int:
        if self.dist is None:
            return int(value)
        import torch

        t = torch.tensor([value], dtype=torch.int64, device=f"cuda:{self.local_rank}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())

    def device_timed(self, steps, warmup):
        """K frames, each bracketed by CUDA events on the library's stream, L2 flushed before each; max over ranks per frame."""
        # W untimed frames, and at least 0.4 s of them: the SM clock of an idle GPU takes that long to come up
        # (with the shared vertex stage a frame is a collective: every rank renders the same number of them, so the ranks
        # agree on going on after every eighth frame)
        # the clock sampler starts with the warm-up: launching nvidia-smi next to an 8-frame timed region stalled single frames
        # by milliseconds (C4 read 1.7 ms instead of 1.02); its samples count from 0.2 s into the warm-up, when the frame
        # load has the clocks up, to the end of the timed region
        sampler = ClockSampler(self.local_rank)
        sampler.start()
        t0, i = time.perf_counter(), 0
        while True:
            self.render_frame(i)
            i += 1
            if i % 8 == 0 or i == max(warmup, 3):
                self.ctx.sync()
                if i >= max(warmup, 3) and (self.max_over_ranks([time.perf_counter() - t0])[0] >= 0.6 or i >= 4000):
                    break
        self.barrier()
        self.ctx.reset_stats()
        self.barrier()
        step_ms = []
        for i in range(steps):
            if self.dist is not None:
                self.barrier()  # untimed: the ranks start the frame together (they meet on the device inside it when the vertex stage is shared)
            self.flush_l2()     # enqueued, not waited for: like at N = 1 the frame is submitted while the flush still runs
            self.ctx.timer_begin()
            self.render_frame(i)
            step_ms.append(self.ctx.timer_end())
        self.barrier()
        clocks = sampler.stop(since=t0 + 0.25)
        st = self.ctx.get_stats()
        step_ms = self.max_over_ranks(step_ms)  # a frame is complete when the slowest rank's rows have landed
        frags = self.sum_over_ranks(st["fragsShaded"]) // steps
        return {"ms": float(sum(step_ms)) / steps, "frags": frags, "stats": st, "clocks": clocks,
                "launches": int(st["kernelLaunches"] - steps)}  # minus the L2-flush clears

    def per_kernel(self, frames):
        """Event-bracketed launches (slcu_profile_*), L2 flushed before every frame like the timed loop."""
        self.ctx.profile_begin()
        flushed = True
        for i in range(frames):
            flushed = self.flush_l2_outside_library() and flushed
            if self.dist is not None:
                self.barrier()
            self.render_frame(i)
        kt = self.ctx.profile_end()
        classes = {k: v for k, v in kt.items() if v["launches"]}
        return classes, flushed

    def end_to_end(self, steps):
        """Host buffers: per-frame H2D of the frame's inputs from pinned memory, the frame, D2H of the colour attachment."""
        scene, ctx, inst = self.scene, self.ctx, self.inst
        per_frame = inst.state.get("per_frame", {})
        staged, h2d = {}, 0
        # host buffers are page-locked (slcu_host_alloc): pageable ones go through the driver's bounce buffer at a third of the link rate
        for key, arr in per_frame.items():
            pin = ctx.host_array(arr.shape, arr.dtype)
            pin[...] = arr
            staged[key] = (pin, None)
            h2d += arr.nbytes
        cbytes = scene.width * scene.height * A.bytes_per_texel(scene.color_type)
        out_pin = ctx.host_array((cbytes,), np.uint8)
        e2e_ms = []
        for i in range(steps + 2):
            self.flush_l2()
            ctx.sync()
            t0 = time.perf_counter()
            for key, (pin, _) in staged.items():
                ctx.vbo_assign_raw(inst.state[key], pin.ctypes.data, pin.nbytes)
            self.render_frame(i)
            if self.sf is not None:
                self.sf.frame_barrier()
                if self.rank == 0:
                    ctx.texture_read_into(self.final[i % 2], out_pin.ctypes.data, cbytes)
            else:
                ctx.texture_read_into(inst.state["color"], out_pin.ctypes.data, cbytes)
            dt = (time.perf_counter() - t0) * 1e3
            if i >= 2:
                e2e_ms.append(dt)
        e2e_ms = self.max_over_ranks(e2e_ms)
        ms = float(np.mean(e2e_ms))
        pipelined = None
        if self.sf is None:
            # the same frames through the present path as a two-deep swap chain (slcu_present): frame k's colour attachment
            # crosses the link while frame k+1 renders.  Every frame's H2D and D2H are inside the timed region; the L2 is not
            # flushed here (the frames overlap; C3's 256 MiB of textures do not fit it anyway)
            views = [ctx.host_array((cbytes,), np.uint8) for _ in range(2)]
            n = steps + 2
            for i in range(n + 2):
                if i == 2:
                    ctx.present_wait(0); ctx.present_wait(1); ctx.sync()
                    t0 = time.perf_counter()
                slot = i & 1
                ctx.present_wait(slot)
                for key, (pin, _) in staged.items():
                    ctx.vbo_assign_raw(inst.state[key], pin.ctypes.data, pin.nbytes)
                self.render_frame(i)
                ctx.present(slot, views[slot].ctypes.data, scene.color_type, scene.width, scene.height, inst.state["color"])
            ctx.present_wait(0); ctx.present_wait(1)
            pms = (time.perf_counter() - t0) * 1e3 / n
            pipelined = {"value": scene.tris_per_frame / (pms * 1e-3) / 1e6, "unit": METRIC, "ms_per_step": pms, "frames": n,
                         "how": "slcu_present, two host views: D2H of frame k overlaps the rendering of frame k+1; wall clock over all frames / frames",
                         "checksum": int(views[(n + 1) & 1][::4097].astype(np.uint64).sum())}
        return {"value": scene.tris_per_frame / (ms * 1e-3) / 1e6, "unit": METRIC, "pipelined": pipelined,
                "h2d_bytes_per_step": int((h2d + scene.upload_bytes_per_frame) * self.world), "d2h_bytes_per_step": int(cbytes), "ms_per_step": ms,
                "host_buffers": "page-locked (slcu_host_alloc)", "checksum": int(out_pin[::4097].astype(np.uint64).sum())}

    def close(self):
        self.ctx.close()


def rooflines(config, scene, frags, frame_ms, classes, prof_frames, clocks):
    """The dominant kernel against both bounds of SURVEY.md section 8d."""
    alg = algorithmic_bytes(config, scene, frags)
    peak, peak_src = measured_peaks()
    dom = max(classes, key=lambda k: classes[k]["ms"])
    dom_ms = classes[dom]["ms"] / classes[dom]["launches"]
    launches_per_frame = classes[dom]["launches"] / prof_frames
    dom_bytes = alg.get(dom, 0) / max(launches_per_frame, 1)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    kernel_ms_per_frame = {k: v["ms"] / prof_frames for k, v in classes.items()}
    frame_kernel_ms = sum(kernel_ms_per_frame.values())
    ncu = ncu_counters(dom, config)
    kname = (ncu or {}).get("kernel") or {"raster": "k_resolve2 (k_resolve for long-span batches)", "shade": "k_shade"}.get(dom, "k_" + dom)
    hbm = {
        "bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": (ncu["dram_bytes_read"] + ncu["dram_bytes_write"]) if ncu else None, "traffic_source": ncu["source"] if ncu else None,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms_per_launch": dom_ms,
        "kernel_share_of_frame": kernel_ms_per_frame[dom] / frame_kernel_ms, "frame_algorithmic_bytes": alg["frame"],
        "frame_achieved_gbs": alg["frame"] / (frame_ms * 1e-3) / 1e9, "frame_frac": alg["frame"] / (frame_ms * 1e-3) / 1e9 / peak,
        "kernel_ms_per_frame": kernel_ms_per_frame,
    }
    # bound 2: instruction issue.  148 SMs x 4 schedulers x f_SM warp instructions per second; f_SM = the SM clock
    # sampled under load in this run.  The kernel's warp-instruction count per launch comes from the ncu capture of the same
    # sources (profiles/ncu_counters.json); its duration is measured live.
    f_sm = (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
    issue_peak = 148 * 4 * f_sm / 1e9
    issue = {"bound": "issue", "kernel": kname, "peak": issue_peak, "unit": "Gwarp-inst/s", "achieved": None, "frac": None,
             "sm_mhz": f_sm / 1e6, "frags_per_launch": frags / max(launches_per_frame, 1)}
    if ncu and dom_ms > 0:
        issue["warp_instructions_per_launch"] = ncu["inst_executed"]
        issue["achieved"] = ncu["inst_executed"] / (dom_ms * 1e-3) / 1e9
        issue["frac"] = issue["achieved"] / issue_peak
        issue["counter_source"] = ncu["source"]
        if frags:
            per_frag = ncu["inst_executed"] * launches_per_frame / frags
            issue["warp_instructions_per_fragment"] = per_frag
            issue["fragment_rate_ceiling_gfrag_s"] = issue_peak / per_frag
            issue["fragment_rate_achieved_gfrag_s"] = frags / launches_per_frame / (dom_ms * 1e-3) / 1e9
    hbm["issue"] = issue
    return hbm


def other_config_leg(config, local_rank, with_cpu):
    """Device ms + e2e ms of another BASELINE config in the same run (N = 1)."""
    h = Harness(config, 0, 1, local_rank, None, 8)
    steps = 3 if config.startswith("c5") else 8
    d = h.device_timed(steps, 3)
    e = h.end_to_end(3 if config.startswith("c5") else 5)
    classes, _ = h.per_kernel(3)
    out = {"workload": workload_config(config, h.scene)["workload"], "ms_per_step": d["ms"], "fps": 1e3 / d["ms"],
           "mtris_per_s": h.scene.tris_per_frame / (d["ms"] * 1e-3) / 1e6, "gfrags_per_s": d["frags"] / (d["ms"] * 1e-3) / 1e9,
           "frags_per_frame": d["frags"], "e2e_ms_per_step": e["ms_per_step"], "h2d_bytes_per_step": e["h2d_bytes_per_step"],
           "d2h_bytes_per_step": e["d2h_bytes_per_step"], "steps": steps,
           "kernel_ms_per_frame": {k: v["ms"] / 3 for k, v in classes.items()}}
    h.close()
    if with_cpu:
        out["cpu_baseline"] = cpu_baseline_leg(config, budget_s=8.0, max_frames=3)
    return out


def run_cuda(args, rank, world, local_rank, dist):
    h = Harness(args.config, rank, world, local_rank, dist, args.strip_rows)
    scene = h.scene
    d = h.device_timed(args.steps, args.warmup)
    ms, frags, st = d["ms"], d["frags"], d["stats"]
    value = scene.tris_per_frame / (ms * 1e-3) / 1e6
    prof_frames = max(3, min(args.steps, 10))
    classes, flushed = h.per_kernel(prof_frames)
    roofline = rooflines(args.config, scene, frags, ms, classes, prof_frames, d["clocks"])
    roofline["per_kernel_pass"] = "L2 flushed before every frame (256 MiB fill outside the library's timers)" if flushed else "L2 not flushed (torch unavailable)"
    e2e = h.end_to_end(args.steps)
    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, scene),
        "l2": "flushed between timed frames (256 MiB clear, untimed)",
        "parallelism": f"sort-first x{world}" + (f", {args.strip_rows}-row strips interleaved over ranks, "
                                                 + ("vertex stage split by vertex block, records routed to the owners of their rows by peer stores, "
                                                    if os.environ.get("SLCU_GEOM_SHARE", "1") != "0" else "geometry replicated, ")
                                                 + "composite = peer stores into rank 0 over NVLink" if world > 1 else ""),
        "gfrags_per_s": frags / (ms * 1e-3) / 1e9, "fps": 1e3 / ms, "frags_per_frame": int(frags),
        "tris_binned_per_frame": int(st["trisBinned"] // args.steps), "spans_per_frame": int(st["spans"] // args.steps),
        "gpu_launches": d["launches"], "clocks": d["clocks"], "e2e": e2e, "roofline": roofline,
    }
    h.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # the CPU leg is reported at N = 1 only
        line["cpu_baseline"] = cpu_baseline_leg(args.config)
    if rank == 0 and world == 1 and not args.no_other_configs and args.config == DEFAULT_CONFIG:
        others = {}
        for cfg in OTHER_CONFIGS:
            try:
                others[cfg] = other_config_leg(cfg, local_rank, with_cpu=cfg.startswith("c5") and not args.no_cpu_baseline)
            except Exception as ex:  # a failing extra must not take the headline down
                others[cfg] = {"error": repr(ex)}
        line["other_configs"] = others
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the extra per-config measurements of the default run")
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
