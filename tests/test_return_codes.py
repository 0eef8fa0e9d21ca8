"""Return-code parity of the host-side validation: the same call sequences go into the CUDA product, the CPU restatement
and the compiled, unmodified reference, and every integer that comes back must be equal.

  SL_Framebuffer::reserve_color_buffers / attach_color_buffer / valid  (softlight/src/SL_Framebuffer.cpp:823-990)
  SL_VertexArray::set_num_bindings                                     (softlight/src/SL_VertexArray.cpp:162-167)
  SL_Context::create_shader                                            (softlight/src/SL_Context.cpp:488-540)
  SL_Texture::init                                                     (softlight/src/SL_Texture.cpp:276-294)

The CPU leg (oracle vs compiled reference) runs wherever oracle/_ref was built; the GPU leg adds the product."""
import os

import pytest

from softlight_b200 import capi as A

HAVE_REF = os.path.exists(A.LIB_PATHS["ref_strict"])


def drive(impl: str):
    """One fixed script of valid and invalid calls; returns [(label, code), ...]."""
    ctx = A.Context(impl)
    f, h = ctx.f, ctx.h
    out = []

    def tex(color_type, w, h_, d=1):
        t = f["create_texture"](h)
        out.append((f"texture_init({color_type},{w},{h_},{d})", f["texture_init"](h, t, color_type, w, h_, d)))
        return t

    c8 = tex(A.SL_COLOR_RGBA_8U, 8, 8)
    c8b = tex(A.SL_COLOR_RGBA_FLOAT, 8, 8)
    c16x8 = tex(A.SL_COLOR_RGBA_8U, 16, 8)
    c8x16 = tex(A.SL_COLOR_RGBA_8U, 8, 16)
    c3d = tex(A.SL_COLOR_RGBA_8U, 8, 8, 2)
    dh = tex(A.SL_COLOR_R_HALF, 8, 8)
    df = tex(A.SL_COLOR_R_FLOAT, 8, 8)
    dd = tex(A.SL_COLOR_R_DOUBLE, 8, 8)
    d16x8 = tex(A.SL_COLOR_R_FLOAT, 16, 8)
    d8x16 = tex(A.SL_COLOR_R_FLOAT, 8, 16)
    d3d = tex(A.SL_COLOR_R_FLOAT, 8, 8, 2)
    dbad = tex(A.SL_COLOR_RGBA_8U, 8, 8)
    dbad2 = tex(A.SL_COLOR_R_16U, 8, 8)
    never = f["create_texture"](h)  # created, never initialised: a view without texels

    def fbo_case(label, ncolors, colors, depth):
        fb = f["create_framebuffer"](h)
        out.append((label + ":valid-before-reserve", f["fbo_valid"](h, fb)))
        out.append((label + ":reserve", f["fbo_reserve_color_buffers"](h, fb, ncolors)))
        for i, t in colors:
            out.append((label + f":attach{i}", f["fbo_attach_color_buffer"](h, fb, i, t)))
        if depth is not None:
            out.append((label + ":depth", f["fbo_attach_depth_buffer"](h, fb, depth)))
        out.append((label + ":valid", f["fbo_valid"](h, fb)))
        return fb

    fb = f["create_framebuffer"](h)
    out.append(("reserve(0) on a fresh fbo", f["fbo_reserve_color_buffers"](h, fb, 0)))  # same count: no-op, 0
    out.append(("reserve(5)", f["fbo_reserve_color_buffers"](h, fb, 5)))
    out.append(("reserve(2)", f["fbo_reserve_color_buffers"](h, fb, 2)))
    out.append(("reserve(0)", f["fbo_reserve_color_buffers"](h, fb, 0)))
    out.append(("attach(index 2 of 2)", f["fbo_attach_color_buffer"](h, fb, 2, c8)))
    out.append(("attach(index 4 of 2)", f["fbo_attach_color_buffer"](h, fb, 4, c8)))

    fbo_case("ok_half", 1, [(0, c8)], dh)
    fbo_case("ok_float", 1, [(0, c8)], df)
    fbo_case("ok_double", 1, [(0, c8)], dd)
    fbo_case("ok_mrt", 2, [(0, c8), (1, c8b)], df)
    fbo_case("no_attachments", 1, [], None)                       # -2
    fbo_case("second_colour_missing", 2, [(0, c8)], df)           # -2
    fbo_case("colour_without_texels", 1, [(0, never)], df)        # -2
    fbo_case("colour_width_differs", 2, [(0, c8), (1, c16x8)], df)   # -3
    fbo_case("colour_height_differs", 2, [(0, c8), (1, c8x16)], df)  # -4
    fbo_case("colour_depth_differs", 2, [(0, c8), (1, c3d)], df)     # -5
    fbo_case("no_depth", 1, [(0, c8)], None)                      # -6
    fbo_case("depth_without_texels", 1, [(0, c8)], never)         # -6
    fbo_case("depth_width_differs", 1, [(0, c8)], d16x8)          # -7
    fbo_case("depth_height_differs", 1, [(0, c8)], d8x16)         # -8
    fbo_case("depth_is_3d", 1, [(0, c8)], d3d)                    # -9
    fbo_case("depth_type_rgba8", 1, [(0, c8)], dbad)              # -10
    fbo_case("depth_type_r16u", 1, [(0, c8)], dbad2)              # -10
    # precedence: a width mismatch of the depth wins over its type
    fbo_case("depth_width_and_type", 1, [(0, c16x8)], dbad)       # -7
    # reserve resets every view
    fb2 = fbo_case("reset_by_reserve", 1, [(0, c8)], df)
    out.append(("reset_by_reserve:reserve-same", f["fbo_reserve_color_buffers"](h, fb2, 1)))
    out.append(("reset_by_reserve:valid-after-same", f["fbo_valid"](h, fb2)))       # same count: nothing is reset
    out.append(("reset_by_reserve:reserve-2", f["fbo_reserve_color_buffers"](h, fb2, 2)))
    out.append(("reset_by_reserve:valid-after-2", f["fbo_valid"](h, fb2)))          # colour views gone: -2
    out.append(("reset_by_reserve:attach0", f["fbo_attach_color_buffer"](h, fb2, 0, c8)))
    out.append(("reset_by_reserve:attach1", f["fbo_attach_color_buffer"](h, fb2, 1, c8b)))
    out.append(("reset_by_reserve:valid-depth-kept", f["fbo_valid"](h, fb2)))       # the depth view survived the reserve: 0

    vao = f["create_vao"](h)
    for n in (0, 1, 7, 8, 9):
        out.append((f"set_num_bindings({n})", f["vao_set_num_bindings"](h, vao, n)))

    ubo = ctx.create_ubo(b"\0" * 16)
    for label, kw in [
        ("shader_ok", dict(vs_varyings=1, fs_varyings=1)),
        ("shader_fs_more_varyings", dict(vs_varyings=1, fs_varyings=2)),
        ("shader_vs_5_varyings", dict(vs_varyings=5, fs_varyings=1)),
        ("shader_fs_5_varyings", dict(vs_varyings=5, fs_varyings=5)),
        ("shader_ok_0_varyings", dict(vs_varyings=1, fs_varyings=0)),
    ]:
        sid = ctx.create_shader("vcolor", "vcolor", ubo, **kw)
        out.append((label, -1 if sid == A.INVALID_ID else 0))
    out.append(("num_threads(0)", ctx.num_threads(0)))
    out.append(("num_threads(3)", ctx.num_threads(3)))
    ctx.num_threads(1)
    return out


def _compare(a_name, b_name):
    a, b = drive(a_name), drive(b_name)
    assert [x[0] for x in a] == [x[0] for x in b]
    diff = [(la, ca, cb) for (la, ca), (_, cb) in zip(a, b) if ca != cb]
    assert not diff, f"{a_name} vs {b_name}: " + ", ".join(f"{l}: {x} != {y}" for l, x, y in diff)


def test_oracle_codes_follow_the_reference_table():
    """The published codes of SL_Framebuffer::valid, spelled out (src/SL_Framebuffer.cpp:928-990)."""
    got = dict(drive("port"))
    want = {
        "ok_half:valid": 0, "ok_float:valid": 0, "ok_double:valid": 0, "ok_mrt:valid": 0,
        "ok_half:valid-before-reserve": -1, "no_attachments:valid": -2, "second_colour_missing:valid": -2,
        "colour_without_texels:valid": -2, "colour_width_differs:valid": -3, "colour_height_differs:valid": -4,
        "colour_depth_differs:valid": -5, "no_depth:valid": -6, "depth_without_texels:valid": -6,
        "depth_width_differs:valid": -7, "depth_height_differs:valid": -8, "depth_is_3d:valid": -9,
        "depth_type_rgba8:valid": -10, "depth_type_r16u:valid": -10, "depth_width_and_type:valid": -7,
        "reset_by_reserve:valid": 0, "reset_by_reserve:valid-after-same": 0, "reset_by_reserve:valid-after-2": -2,
        "reset_by_reserve:valid-depth-kept": 0, "reserve(0) on a fresh fbo": 0,
        "reserve(0)": -1, "reserve(5)": -2, "reserve(2)": 0, "attach(index 2 of 2)": -1, "attach(index 4 of 2)": -1,
        "set_num_bindings(7)": 7, "set_num_bindings(8)": -1, "set_num_bindings(9)": -1,
        "shader_ok": 0, "shader_fs_more_varyings": -1, "shader_vs_5_varyings": -1, "shader_fs_5_varyings": -1,
        "num_threads(0)": 1, "num_threads(3)": 3,
    }
    for k, v in want.items():
        assert got[k] == v, (k, got[k], v)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_codes_equal_the_compiled_reference():
    _compare("port", "ref_strict")


@pytest.mark.gpu
def test_cuda_codes_equal_the_oracle():
    _compare("cuda", "port")


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref did not travel to this box")
def test_cuda_codes_equal_the_compiled_reference():
    _compare("cuda", "ref_strict")
