"""The C++ host side above the C ABI (include/softlight_b200/SL_Context.hpp): a port of the reference's own
sl_draw_test.cpp to the mirror header must compile as plain C++11 with g++ (no CUDA headers on the host side), link
against libslcu.so only, fail loudly without a GPU, and on a GPU reproduce the reference's picture (known answer)."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "cpp", "sl_draw_test_b200.cpp")
OUT_DIR = os.path.join(HERE, "cpp", "_build")
BIN = os.path.join(OUT_DIR, "sl_draw_test_b200")
LIBDIR = os.path.join(ROOT, "softlight_b200")


def build_binary():
    from softlight_b200 import build as B

    B.build_cuda()
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-L", LIBDIR, "-lslcu",
           "-Wl,-rpath," + LIBDIR, "-o", BIN]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return BIN


def test_cpp_host_mirror_compiles_links_and_refuses_to_run_without_a_gpu():
    exe = build_binary()
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present: covered by the gpu test")
    except ImportError:
        pass
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)       # no CPU fallback: the context is invalid
    assert "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_draw_test_known_answer():
    """softlight/tests/sl_draw_test.cpp through the C++ mirror: 38,278 covered pixels, centre pixel 0xff40803f."""
    exe = build_binary()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "covered=38278 centre=3f8040ff" in r.stdout and "MATCH" in r.stdout
