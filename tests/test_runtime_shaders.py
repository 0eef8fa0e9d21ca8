"""Run-time shader registration (slcu_register_shader_source): the stand-in for the reference taking any host function
pointer as a shader (softlight/include/softlight/SL_Shader.hpp:63, 97).  The source is compiled with NVRTC against the
headers the product library was built from, so a new shader needs no rebuild of libslcu.so.

CPU: the source compiles (NVRTC needs no device), a broken source is rejected with the compiler's log.
GPU: a shader registered at run time under a new name renders bit-identically to the same body compiled into the
library, on triangle, point and wireframe paths; a shader that exists ONLY at run time matches its known answer."""
import ctypes as C

import numpy as np
import pytest

import cases
from softlight_b200 import capi as A
from softlight_b200 import scenes as S

# the draw-test shader (softlight/tests/sl_draw_test.cpp:27-83) under new names + a shader the library does not have
SOURCE = r"""
struct RtVertex { math::vec4 pos; math::vec4 color; };

SL_SHADER_FN math::vec4 rt_vcolor_vs(SL_VertexParam& param)
{
    const RtVertex& vert = *(param.pVbo->template element<const RtVertex>(param.pVao->offset(0, param.vertId)));
    param.pVaryings[0] = vert.color;
    return vert.pos;
}
SL_SHADER_FN bool rt_vcolor_fs(SL_FragmentParam& fragParam)
{
    fragParam.pOutputs[0] = fragParam.pVaryings[0];
    return true;
}
/* not in the library: inverts the interpolated colour and discards every other column */
SL_SHADER_FN bool rt_invert_fs(SL_FragmentParam& fragParam)
{
    const math::vec4 c = fragParam.pVaryings[0];
    fragParam.pOutputs[0] = math::vec4{1.f - c[0], 1.f - c[1], 1.f - c[2], 1.f};
    return (fragParam.coord.x & 1u) == 0u;
}
SL_REGISTER_VERTEX_SHADER("rt_vcolor", rt_vcolor_vs)
SL_REGISTER_FRAGMENT_SHADER("rt_vcolor", rt_vcolor_fs)
SL_REGISTER_FRAGMENT_SHADER("rt_invert", rt_invert_fs)
"""


def check(source):
    fns = A.load_library("cuda")
    log = C.create_string_buffer(1 << 16)
    rc = fns["shader_source_check"](source.encode(), log, len(log))
    return rc, log.value.decode(errors="replace")


def test_source_compiles_without_a_device():
    rc, log = check(SOURCE)
    assert rc == 3, log


def test_broken_source_is_rejected_with_the_compiler_log():
    rc, log = check(SOURCE.replace("return vert.pos;", "return vert.position_that_does_not_exist;"))
    assert rc < 0 and "position_that_does_not_exist" in log
    rc, log = check("/* no registration */ struct X { int a; };")
    assert rc < 0 and "registers no shader" in log


def _render(scene, shader_names=None, register=None):
    ctx = A.Context("cuda")
    if register:
        assert ctx.register_shader_source(register) == 3
    if shader_names:
        real = ctx.create_shader

        def patched(vs_name, fs_name, ubo, **kw):
            return real(shader_names[0] if vs_name == "vcolor" else vs_name, shader_names[1] if fs_name == "vcolor" else fs_name, ubo, **kw)

        ctx.create_shader = patched
    inst = S.SceneInstance(scene, ctx)
    inst.render()
    ctx.sync()
    color, depth, stats = inst.read_color(), inst.read_depth(), ctx.get_stats()
    ctx.close()
    return color, depth, stats


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rand_blend1", "rand_clip_heavy", "points_dense", "wire_alpha"])
def test_runtime_copy_of_a_builtin_shader_is_bit_identical(name):
    scene_fn = cases.CASES[name]
    a = _render(scene_fn())
    b = _render(scene_fn(), shader_names=("rt_vcolor", "rt_vcolor"), register=SOURCE)
    assert np.array_equal(a[0], b[0]) and np.array_equal(cases.bits(a[1]), cases.bits(b[1]))
    assert a[2]["fragsShaded"] == b[2]["fragsShaded"] > 0


@pytest.mark.gpu
def test_shader_that_exists_only_at_run_time():
    scene = S.draw_test_scene()
    base, _, _ = _render(scene)
    inv, _, st = _render(S.draw_test_scene(), shader_names=("rt_vcolor", "rt_invert"), register=SOURCE)
    inside = (base != np.array([255, 0, 255, 255], dtype=np.uint8)).any(axis=-1)   # the triangle (clear colour is magenta)
    even = (np.arange(base.shape[1]) % 2 == 0)[None, :]
    # odd columns were discarded: still the clear colour; even columns hold the inverted colour
    assert (inv[inside & ~even] == np.array([255, 0, 255, 255], dtype=np.uint8)).all()
    got = inv[inside & even].astype(np.int32)
    want = 255 - base[inside & even].astype(np.int32)
    want[:, 3] = 255
    assert np.abs(got - want).max() <= 1   # 1 - c in float, then * 255 and rounded: within one LSB of 255 - round(c * 255)
    assert st["fragsShaded"] > 0


@pytest.mark.gpu
def test_unknown_shader_name_is_still_refused():
    ctx = A.Context("cuda")
    ubo = ctx.create_ubo(b"\0" * 16)
    assert ctx.create_shader("rt_vcolor", "rt_vcolor", ubo, vs_varyings=1, fs_varyings=1) == A.INVALID_ID  # not registered in THIS context
    assert ctx.register_shader_source(SOURCE) == 3
    assert ctx.create_shader("rt_vcolor", "rt_invert", ubo, vs_varyings=1, fs_varyings=1) != A.INVALID_ID
