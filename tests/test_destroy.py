"""SL_Context::destroy_* (softlight/src/SL_Context.cpp:192-195, 243-247, 295-298, 346-349, 397-400, 448-451, 612-615): the
object is erased from its vector and every later id of that kind moves down by one.  The same script runs in the CPU
restatement, the compiled reference and (GPU) the CUDA product: dummies are created in front of a scene's objects and
destroyed afterwards, the scene is then drawn through the shifted ids and must come out unchanged."""
import os

import numpy as np
import pytest

import cases
from softlight_b200 import capi as A
from softlight_b200 import scenes as S

HAVE_REF = os.path.exists(A.LIB_PATHS["ref_strict"])


def shifted_frame(impl: str, scene_fn, rcp_table=None, tex_uniform=False):
    scene = scene_fn()
    ctx = A.Context(impl)
    if rcp_table is not None and not impl.startswith("ref"):
        ctx.set_rcp_table(rcp_table)
    # dummies in front of everything the scene creates
    d_tex = ctx.create_texture(A.SL_COLOR_RGBA_8U, 4, 4)
    d_vbo = ctx.create_vbo(np.zeros(16, dtype=np.float32))
    d_ibo = ctx.create_ibo(np.zeros(3, dtype=np.uint8))
    d_vao = ctx.create_vao(d_vbo, d_ibo, [(0, 16, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_FLOAT)])
    d_fbo = ctx.create_framebuffer([d_tex], ctx.create_texture(A.SL_COLOR_R_FLOAT, 4, 4))  # second dummy texture: id 1
    assert (d_tex, d_vbo, d_ibo, d_vao, d_fbo) == (0, 0, 0, 0, 0)
    inst = S.SceneInstance(scene, ctx)
    st = inst.state
    inst.render()                      # one frame before anything is destroyed
    ctx.sync()
    before = inst.read_color().copy()
    # a shader and a uniform block behind the scene's own (destroying them moves nothing the scene uses)
    d_ubo = ctx.create_ubo(b"\0" * 16)
    d_sh = ctx.create_shader("vcolor", "vcolor", d_ubo, vs_varyings=1, fs_varyings=1)
    assert d_sh != A.INVALID_ID and d_sh > st["shader"]
    assert ctx.destroy("shader", d_sh) == 0 and ctx.destroy("ubo", d_ubo) == 0
    assert ctx.destroy("framebuffer", d_fbo) == 0
    assert ctx.destroy("texture", 1) == 0 and ctx.destroy("texture", 0) == 0   # two textures in front of the scene's
    assert ctx.destroy("vao", d_vao) == 0 and ctx.destroy("ibo", d_ibo) == 0 and ctx.destroy("vbo", d_vbo) == 0
    # the caller's ids follow the erase, exactly as with the reference
    st["color"] -= 2
    st["depth"] -= 2
    st["fbo"] -= 1
    meshes = st["meshes"] if "meshes" in st else [st["mesh"]]
    moved = [ctx.mesh(m.vaoId - 1, m.elementBegin, m.elementEnd, m.mode) for m in meshes]
    if "meshes" in st:
        st["meshes"] = moved
    else:
        st["mesh"] = moved[0]
    # a vertex array names its buffers by INDEX (SL_VertexArray::set_vertex_buffer): the caller points it at the moved ids
    for vao in sorted({m.vaoId for m in moved}):
        ctx.f["vao_set_vertex_buffer"](ctx.h, vao, 0)
        ctx.f["vao_set_index_buffer"](ctx.h, vao, 0)
    assert ctx.fbo_valid(st["fbo"]) == 0   # the attachments followed their textures
    st.pop("calls", None)
    inst.render()
    ctx.sync()
    after = inst.read_color()
    depth = inst.read_depth()
    if not impl.startswith("ref"):
        assert ctx.destroy("texture", 99) != 0   # an id that does not exist: refused (undefined behaviour in the reference)
    ctx.close()
    return before, after, depth


@pytest.mark.parametrize("impl", ["port"] + (["ref_strict"] if HAVE_REF else []))
def test_destroy_shifts_ids_cpu(impl, golden):
    frames, _ = golden
    before, after, _ = shifted_frame(impl, S.draw_test_scene, frames["rcp_table"])
    assert np.array_equal(before, after)
    assert after[240, 320].tolist() != [255, 0, 255, 255]   # the triangle is there (clear colour is magenta)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
def test_destroy_oracle_equals_reference(golden):
    frames, _ = golden
    a = shifted_frame("port", cases.CASES["rand_blend1"], frames["rcp_table"])
    b = shifted_frame("ref_strict", cases.CASES["rand_blend1"])
    assert np.array_equal(a[1], b[1]) and np.array_equal(cases.bits(a[2]), cases.bits(b[2]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["draw_test", "rand_blend1", "rand_index_u16"])
def test_destroy_shifts_ids_cuda(name, golden):
    frames, _ = golden
    fn = S.draw_test_scene if name == "draw_test" else cases.CASES[name]
    bg, ag, dg = shifted_frame("cuda", fn, frames["rcp_table"])
    bo, ao, do = shifted_frame("port", fn, frames["rcp_table"])
    assert np.array_equal(bg, ag) and np.array_equal(ag, ao)
    assert np.array_equal(cases.bits(dg), cases.bits(do))
