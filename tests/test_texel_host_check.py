"""The product's texel-conversion header (softlight_b200/csrc/sl_texel.cuh: every blit pair of the 33 colour types, the packed
render-target casts) compiled for the HOST (tests/host_check/sl_texel_host.cpp gives the CUDA intrinsics their IEEE meaning)
and compared with the oracle texel by texel.  The GPU tests check the same code where it runs; this one finds a logic slip
in the header without a GPU.  The oracle is the checker here, as everywhere in tests/."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from softlight_b200 import capi as A
import test_blit_matrix as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("texel_host") / "libtexelhost.so")
    src = os.path.join(ROOT, "tests", "host_check", "sl_texel_host.cpp")
    subprocess.run(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.texel_blit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_size_t]
    lib.texel_packed_from_float.argtypes = [C.c_int, C.c_void_p]
    lib.texel_packed_from_float.restype = C.c_uint32
    lib.texel_float_from_packed.argtypes = [C.c_int, C.c_uint32, C.c_void_p]
    return lib


def host_blit(lib, src, st, dt):
    """The header's conversion on every texel of `src`, vertically flipped like the blit (same size: x maps to x)."""
    h, w = src.shape[:2]
    flipped = np.ascontiguousarray(src[::-1])
    out = np.zeros((h, w, A.texel_channels(dt)), dtype=A.texel_dtype(dt))
    lib.texel_blit(flipped.ctypes.data, st, out.ctypes.data, dt, h * w)
    return out.tobytes()


@pytest.mark.parametrize("mode", [0, 1])
def test_header_on_host_equals_oracle_on_the_matrix(host_lib, mode):
    want = T.run_matrix("port", mode)
    bad = [T.label(st, dt) for (st, dt) in want if host_blit(host_lib, T.source_texels(st, 8, 4, mode), st, dt) != want[(st, dt)]]
    assert not bad, bad[:10]


def test_header_on_host_equals_oracle_on_the_packed_pairs(host_lib):
    want = T.run_packed("port")
    bad = [T.label(st, dt) for (st, dt) in want if host_blit(host_lib, T.packed_source(st), st, dt) != want[(st, dt)]]
    assert not bad, bad[:10]


def test_header_on_host_render_target_casts_equal_oracle(host_lib):
    port = C.CDLL(A.LIB_PATHS["port"])
    port.slor_test_packed_from_float.argtypes = [C.c_int, C.c_void_p]
    port.slor_test_packed_from_float.restype = C.c_uint32
    port.slor_test_float_from_packed.argtypes = [C.c_int, C.c_uint32, C.c_void_p]
    rng = np.random.default_rng(5)
    colors = rng.uniform(-0.2, 1.2, size=(4000, 4)).astype(np.float32)
    colors[:2000] = rng.uniform(0, 1, size=(2000, 4)).astype(np.float32)
    colors[0] = [np.nan, np.inf, -np.inf, 1e20]
    colors[1] = [0.5, 1.0, 0.0, -0.0]
    colors[2] = [0.4995, 0.5005, 1.5 / 1023, 2.5 / 3]  # ties of the cvtps2dq in the 1010102 cast
    for t in range(28, 33):
        for c in colors:
            c = np.ascontiguousarray(c)
            assert host_lib.texel_packed_from_float(t, c.ctypes.data) == port.slor_test_packed_from_float(t, c.ctypes.data), (t, c)
        nbits = 8 * A.bytes_per_texel(t)
        values = np.arange(2 ** nbits, dtype=np.uint64) if nbits <= 16 else rng.integers(0, 2 ** 32, size=20000, dtype=np.uint64)
        a, b = np.zeros(4, np.float32), np.zeros(4, np.float32)
        for v in values[:: max(1, len(values) // 20000)]:
            host_lib.texel_float_from_packed(t, int(v), a.ctypes.data)
            port.slor_test_float_from_packed(t, int(v), b.ctypes.data)
            assert a.tobytes() == b.tobytes(), (t, int(v))
