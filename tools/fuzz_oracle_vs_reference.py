"""Randomised oracle-vs-reference comparison on the CPU (development aid; needs oracle/_ref, i.e. /root/reference).

    python tools/fuzz_oracle_vs_reference.py <seed> <seconds> [all|tris|lines]

Every case is a random_tris_scene with random size, colour type (all 33), depth type, depth function, blend mode, cull mode,
clip fraction, viewport / scissor, draw split and primitive family.  The unmodified reference (ref_strict) and the oracle
port render it in a forked child, because the reference itself dies on some inputs (a triangle that
clips into nine vertices overruns its varyings array, SL_TriProcessor.cpp:399/471; fragments left of / below the
framebuffer are written outside the textures).  Line draws get an even element count per draw: with an odd one the
reference reads one element past the buffer and its picture depends on what the heap holds there.
Prints every mismatch / crash with the keyword arguments that reproduce it, and the totals."""
import os
import pickle
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases as Cs  # noqa: E402
from softlight_b200 import capi as A, scenes as S  # noqa: E402


def run_case(kw, rcp):
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        os.close(r)
        try:
            sc = S.random_tris_scene(**kw)
            c0, d0, s0 = Cs.render("ref_strict", sc)
            c1, d1, s1 = Cs.render("port", sc, rcp_table=rcp)
            b0 = Cs.bits(np.ascontiguousarray(c0)) if c0.dtype.kind == "f" else c0
            b1 = Cs.bits(np.ascontiguousarray(c1)) if c1.dtype.kind == "f" else c1
            res = (int((b0 != b1).any(axis=-1).sum()), int((Cs.bits(d0) != Cs.bits(d1)).sum()), int(s0.get("fragsShaded", 0)), int(s1.get("fragsShaded", 0)))
        except Exception as e:  # noqa: BLE001
            res = ("EXC", repr(e)[:200])
        os.write(w, pickle.dumps(res))
        os._exit(0)
    os.close(w)
    data = b""
    while True:
        b = os.read(r, 65536)
        if not b:
            break
        data += b
    os.close(r)
    _, status = os.waitpid(pid, 0)
    return pickle.loads(data) if data else ("CRASH", status)


def random_case(rng, family="all", max_clip=0.9):
    """Keyword arguments of one random_tris_scene."""
    kw = dict(seed=int(rng.integers(1, 1 << 30)), n=int(rng.integers(20, 1500)), width=int(rng.integers(1, 700)), height=int(rng.integers(1, 400)),
              blend=int(rng.integers(0, 5)), depth_test=int(rng.integers(0, 7)), depth_mask=int(rng.integers(0, 2)),
              depth_type=[A.SL_COLOR_R_HALF, A.SL_COLOR_R_FLOAT, A.SL_COLOR_R_DOUBLE][int(rng.integers(0, 3))],
              color_type=int(rng.integers(0, 33)), cull=int(rng.integers(0, 3)), clip_fraction=float(rng.uniform(0, max_clip)),
              clear_depth=float(rng.choice([1.0, 0.5, 0.0, -1.0])), split_draws=int(rng.integers(1, 4)))
    mode = int(rng.integers(0, 10))
    if family == "lines":
        mode = 2
    elif family == "tris" and mode in (0, 1, 2):
        mode = 9
    if mode == 0:
        kw["points"] = True
    elif mode == 1:
        kw["wire"] = True
    elif mode == 2:
        kw["lines"] = True
        kw["n"] = max(1, kw["n"] // (2 * kw["split_draws"])) * 2 * kw["split_draws"]
    elif mode == 3:
        kw["instances"] = int(rng.integers(2, 5))
    for key in ("viewport", "scissor"):
        if rng.random() < 0.3:
            w, h = kw["width"], kw["height"]
            kw[key] = (int(rng.integers(0, w)), int(rng.integers(0, h)), int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1)))
    return kw


def classify(res):
    return "crash" if res[0] == "CRASH" else "exc" if res[0] == "EXC" else "ok" if (res[0] == 0 and res[1] == 0) else "mismatch"


def main():
    seed, seconds = int(sys.argv[1]), float(sys.argv[2])
    family = sys.argv[3] if len(sys.argv) > 3 else "all"
    ctx = A.Context("ref_strict")
    rcp = ctx.get_rcp_table()
    ctx.close()
    rng = np.random.default_rng(seed)
    t0, n, stats = time.time(), 0, {}
    while time.time() - t0 < seconds:
        kw = random_case(rng, family)
        res = run_case(kw, rcp)
        n += 1
        kind = classify(res)
        stats[kind] = stats.get(kind, 0) + 1
        if kind != "ok":
            print(kind.upper(), res, kw, flush=True)
    print("ran", n, stats)


if __name__ == "__main__":
    main()
