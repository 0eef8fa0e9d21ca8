"""Regenerates tests/golden/* by running the UNMODIFIED reference (oracle/_ref/libslref_strict.so,
compiled from /root/reference by oracle/Makefile) with num_threads(1) on the synthetic cases of
tests/cases.py.  Run in the dev container (needs /root/reference):  python tests/golden/make_golden.py
The rcp table of the generating CPU is stored next to the frames: parity tests pin it, so the
goldens hold on any host CPU."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from softlight_b200 import capi as A  # noqa: E402
import cases  # noqa: E402
import blit_cases  # noqa: E402


def main():
    """`--only a,b,c` regenerates just those draw cases and merges them into the committed files."""
    only = None
    if "--only" in sys.argv:
        only = set(sys.argv[sys.argv.index("--only") + 1].split(","))
    if "--blits" in sys.argv:  # only the blit cases: hashes of the reference's output, merged into the committed file
        old = json.load(open(os.path.join(HERE, "golden_hashes.json")))
        for name in blit_cases.BLIT_CASES:
            out = blit_cases.run_blit("ref_strict", name)
            old["blits"][name] = hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest()
            print("blit", name, old["blits"][name][:12], flush=True)
        json.dump(old, open(os.path.join(HERE, "golden_hashes.json"), "w"), indent=1)
        return
    ctx = A.Context("ref_strict")
    rcp = ctx.get_rcp_table()
    frames = {"rcp_table": rcp, "rsqrt_table": ctx.get_rsqrt_table()}
    hashes = {}
    allcases = dict(cases.CASES)
    allcases.update(cases.BIG_CASES)
    if only is not None:
        old = json.load(open(os.path.join(HERE, "golden_hashes.json")))
        oldFrames = dict(np.load(os.path.join(HERE, "golden_frames.npz")))
        assert np.array_equal(oldFrames["rcp_table"], rcp), "the committed goldens were generated with another rcp table"
        if "rsqrt_table" in oldFrames:
            assert np.array_equal(oldFrames["rsqrt_table"], ctx.get_rsqrt_table()), "the committed goldens were generated with another rsqrt table"
        oldFrames["rsqrt_table"] = ctx.get_rsqrt_table()
        assert only <= set(allcases), only - set(allcases)
        for name in sorted(only):
            scene = allcases[name]()
            color, depth, stats = cases.render("ref_strict", scene)
            old["cases"][name] = {
                "color_sha256": hashlib.sha256(np.ascontiguousarray(color).tobytes()).hexdigest(),
                "depth_sha256": hashlib.sha256(np.ascontiguousarray(depth).tobytes()).hexdigest(),
                "frags": int(stats["fragsShaded"]), "width": scene.width, "height": scene.height,
            }
            if name in cases.FULL_IMAGE_CASES:
                oldFrames[name + ".color"] = color
                oldFrames[name + ".depth"] = cases.bits(depth)
            print(name, old["cases"][name]["frags"], flush=True)
        np.savez_compressed(os.path.join(HERE, "golden_frames.npz"), **oldFrames)
        json.dump(old, open(os.path.join(HERE, "golden_hashes.json"), "w"), indent=1)
        return
    for name, factory in allcases.items():
        scene = factory()
        color, depth, stats = cases.render("ref_strict", scene)
        hashes[name] = {
            "color_sha256": hashlib.sha256(np.ascontiguousarray(color).tobytes()).hexdigest(),
            "depth_sha256": hashlib.sha256(np.ascontiguousarray(depth).tobytes()).hexdigest(),
            "frags": int(stats["fragsShaded"]), "width": scene.width, "height": scene.height,
        }
        if name in cases.FULL_IMAGE_CASES:
            frames[name + ".color"] = color
            frames[name + ".depth"] = cases.bits(depth)
        print(name, hashes[name]["frags"], flush=True)
    mrt = {}
    for name in sorted(cases.MRT_CASES):
        colors, depth, stats = cases.render_mrt("ref_strict", cases.mrt_scene(name))
        mrt[name] = {"colors_sha256": [hashlib.sha256(np.ascontiguousarray(c).tobytes()).hexdigest() for c in colors],
                     "depth_sha256": hashlib.sha256(np.ascontiguousarray(depth).tobytes()).hexdigest(), "frags": int(stats["fragsShaded"])}
        print("mrt", name, mrt[name]["frags"], flush=True)
    blits = {}
    for name in blit_cases.BLIT_CASES:
        out = blit_cases.run_blit("ref_strict", name)
        blits[name] = hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest()
        print("blit", name, blits[name][:12], flush=True)
    np.savez_compressed(os.path.join(HERE, "golden_frames.npz"), **frames)
    json.dump({"generator": ctx.impl_name(), "cases": hashes, "blits": blits, "mrt": mrt}, open(os.path.join(HERE, "golden_hashes.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
