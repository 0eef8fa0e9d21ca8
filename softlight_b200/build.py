"""Build recipes (called by __graft_entry__.build()).

* build_cuda():   nvcc -> softlight_b200/libslcu.so (sm_100a only, -lineinfo, -fmad=false)
* build_oracle(): g++  -> oracle/libslor.so (CPU restatement) and, when /root/reference is
                  present, oracle/_ref/libslref_{strict,fast}.so (the compiled reference).
Everything is built in-tree so the .so files travel to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "softlight_b200", "csrc")
LIB = os.path.join(ROOT, "softlight_b200", "libslcu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # nothing is contracted behind our back; FMAs are explicit (__fmaf_rn)
    "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def cuda_sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    srcs += [os.path.join(ROOT, "shaders", f) for f in sorted(os.listdir(os.path.join(ROOT, "shaders")))]
    srcs.append(os.path.join(ROOT, "include", "slcu.h"))
    return srcs


def build_cuda(force: bool = False, verbose: bool = False, out: str = LIB) -> str:
    """`out` other than LIB: an A/B variant (SLCU_NVCC_EXTRA flags), selected at run time with SLCU_LIB=<path>."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not force and _newer(out, cuda_sources()):
        return out
    cmd = [nvcc] + NVCC_FLAGS + ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "shaders"), "-I" + CSRC,
                                 os.path.join(CSRC, "slcu.cu"), "-o", out]
    cmd[1:1] = os.environ.get("SLCU_NVCC_EXTRA", "").split()  # tuning experiments only (e.g. -DSL_RASTER_MIN_BLOCKS=4)
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(r.stderr)
    return out


HARNESS_LIB = os.path.join(ROOT, "softlight_b200", "libslharness.so")


def build_harness() -> str:
    """The measurement harness' call replayer (plain C): softlight_b200/harness/slharness.c -> libslharness.so."""
    src = os.path.join(ROOT, "softlight_b200", "harness", "slharness.c")
    if _newer(HARNESS_LIB, [src]):
        return HARNESS_LIB
    r = subprocess.run(["gcc", "-O2", "-shared", "-fPIC", src, "-o", HARNESS_LIB], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("harness build failed")
    return HARNESS_LIB


def build_oracle() -> None:
    odir = os.path.join(ROOT, "oracle")
    r = subprocess.run(["make", "-s", "-C", odir, "port"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("oracle port build failed")
    if os.path.isdir("/root/reference/softlight/src"):
        r = subprocess.run(["make", "-s", "-j8", "-C", odir, "ref"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("oracle/_ref build failed")


if __name__ == "__main__":
    _out = LIB
    for _a in sys.argv[1:]:
        if _a.startswith("--out="):
            _out = os.path.abspath(_a[6:])
            os.makedirs(os.path.dirname(_out), exist_ok=True)
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out)
    build_harness()
    build_oracle()
    print("built", _out)
