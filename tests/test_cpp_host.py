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


def build_binary(src=SRC, out=BIN):
    from softlight_b200 import build as B

    B.build_cuda()
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-L", LIBDIR, "-lslcu",
           "-Wl,-rpath," + LIBDIR, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


INST_SRC = os.path.join(HERE, "cpp", "sl_instancing_test_b200.cpp")
INST_BIN = os.path.join(OUT_DIR, "sl_instancing_test_b200")


def test_cpp_instancing_demo_compiles():
    build_binary(INST_SRC, INST_BIN)


@pytest.mark.gpu
def test_cpp_instancing_demo_equals_the_oracle(tmp_path):
    """softlight/tests/sl_instancing_test.cpp's pattern through the C++ mirror: draw_instanced, a texture and the instance
    matrices reached through pointers in the uniform block, and uniforms / instance data changed between two draws through
    pointers taken BEFORE the first one (the reference hands out live pointers).  The program dumps its inputs; the same
    inputs go through the CPU oracle and the pictures must be equal."""
    import numpy as np

    from softlight_b200 import capi as A

    exe = build_binary(INST_SRC, INST_BIN)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)

    def load(name, dtype=np.uint8):
        return np.fromfile(os.path.join(str(tmp_path), name), dtype=dtype)

    W, H, n_inst = 320, 200, 27
    ctx = A.Context("port")
    ctx.set_rcp_table(load("rcp.bin", np.uint32))
    color = ctx.create_texture(A.SL_COLOR_RGBA_8U, W, H)
    depth = ctx.create_texture(A.SL_COLOR_R_HALF, W, H)
    fbo = ctx.create_framebuffer([color], depth)
    vbo = ctx.create_vbo(load("verts.bin"))
    idx = load("idx.bin", np.uint16)
    ibo = ctx.create_ibo(idx)
    vao = ctx.create_vao(vbo, ibo, [(0, 32, A.VERTEX_DIMENSION_3, A.VERTEX_DATA_FLOAT), (12, 32, A.VERTEX_DIMENSION_2, A.VERTEX_DATA_FLOAT),
                                    (20, 32, A.VERTEX_DIMENSION_3, A.VERTEX_DATA_FLOAT)])
    tex = ctx.create_texture(A.SL_COLOR_RGB_8U, 16, 16, 1, load("tex.bin").reshape(16, 16, 3))
    inst = ctx.create_vbo(load("mats1.bin"))
    ubo = ctx.create_ubo(load("blob1.bin").tobytes())
    ctx.ubo_bind_texture(ubo, 0, tex)
    ctx.ubo_bind_buffer(ubo, 48, inst, 0)
    sh = ctx.create_shader("inst_tex", "inst_tex", ubo, vs_varyings=3, fs_varyings=3, cull=A.SL_CULL_BACK_FACE,
                           blend=A.SL_BLEND_PREMULTIPLED_ALPHA, depth_test=A.SL_DEPTH_TEST_GREATER_EQUAL, depth_mask=A.SL_DEPTH_MASK_ON)
    mesh = ctx.mesh(vao, 0, idx.size)
    ctx.clear_framebuffer(fbo, 0, (0.6, 0.6, 0.6, 1.0), 0.0)
    ctx.draw_instanced(mesh, n_inst, sh, fbo)
    ctx.vbo_assign(inst, load("mats2.bin"))
    ctx.ubo_assign(ubo, load("blob2.bin").tobytes())
    ctx.draw_instanced(mesh, n_inst, sh, fbo)
    want = ctx.texture_read(color, A.SL_COLOR_RGBA_8U, W, H)
    got = load("color.bin").reshape(H, W, 4)
    assert (want != np.array([153, 153, 153, 255], dtype=np.uint8)).any(axis=-1).sum() > 2000   # something was drawn
    assert np.array_equal(got, want)


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
