"""CPU tests of the boundary: the product library loads without a GPU, exports every symbol
include/slcu.h declares, refuses to run without a device (no CPU fallback), and the host-side
validation mirrors the reference's return codes."""
import ctypes as C
import os
import re

import pytest

from softlight_b200 import capi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "slcu.h")).read()
    return sorted(set(re.findall(r"SLCU_FN\((\w+)\)\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == A.EXPORTED_SYMBOLS


@pytest.mark.parametrize("impl", ["cuda", "port"])
def test_library_exports_every_declared_symbol(impl):
    if impl == "cuda" and not os.path.exists(A.LIB_PATHS["cuda"]):
        from softlight_b200 import build as B

        B.build_cuda()
    lib = C.CDLL(A.LIB_PATHS[impl])
    for name in header_symbols():
        if impl != "cuda" and name in A._SIGS_PRODUCT:
            continue  # the multi-GPU entry points exist in the product only; the checker is one process
        assert hasattr(lib, A.PREFIX[impl] + name), name


def test_no_cpu_fallback_without_device():
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(A.SlError):
        A.Context("cuda")


def test_product_does_not_link_the_oracle():
    """The product .so must not depend on anything under oracle/."""
    import subprocess

    out = subprocess.run(["ldd", A.LIB_PATHS["cuda"]], capture_output=True, text=True).stdout
    assert "slor" not in out and "slref" not in out
    src_dir = os.path.join(ROOT, "softlight_b200")
    for dirpath, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "#include \"slor" not in text and "libslor" not in text and "libslref" not in text
                assert "dlopen" not in text  # the product loads nothing at run time, least of all a checker


def test_reference_error_codes_port():
    ctx = A.Context("port")
    f = ctx.f
    fbo = f["create_framebuffer"](ctx.h)
    assert f["fbo_reserve_color_buffers"](ctx.h, fbo, 5) == -2   # src/SL_Framebuffer.cpp:823-848
    assert f["fbo_reserve_color_buffers"](ctx.h, fbo, 2) == 0
    assert f["fbo_reserve_color_buffers"](ctx.h, fbo, 0) == -1
    tex = ctx.create_texture(A.SL_COLOR_RGBA_8U, 8, 8)
    assert f["fbo_attach_color_buffer"](ctx.h, fbo, 2, tex) == -1
    vao = f["create_vao"](ctx.h)
    assert f["vao_set_num_bindings"](ctx.h, vao, 8) == -1        # src/SL_VertexArray.cpp:162-167
    assert f["vao_set_num_bindings"](ctx.h, vao, 7) == 7
    ubo = ctx.create_ubo(b"\0" * 16)
    assert ctx.create_shader("no_such_shader", "vcolor", ubo, vs_varyings=1, fs_varyings=1) == A.INVALID_ID
    assert ctx.create_shader("vcolor", "vcolor", ubo, vs_varyings=1, fs_varyings=2) == A.INVALID_ID  # fs > vs varyings
    assert ctx.num_threads(0) == 1
