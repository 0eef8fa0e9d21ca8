"""GPU tests (-m gpu): the CUDA path, called through the C ABI, against
  * the committed golden frames (rendered by the unmodified reference; rcp table pinned),
  * the CPU oracle restatement on the same seeded inputs,
  * the compiled reference itself when oracle/_ref travelled to the box (host-measured rcp table),
  * size-independent properties at BASELINE.json's full sizes (idempotence, order independence
    of disjoint draws, clear-only frames).
Bar (north_star): coverage, depth bits, integer outputs bit-exact; colours <= 1 LSB.  We assert
bit-exact everywhere; the tolerance is only the documented fallback."""
import hashlib
import os

import numpy as np
import pytest

import cases
from softlight_b200 import capi as A
from softlight_b200 import scenes as S

pytestmark = pytest.mark.gpu

COLOR_TOL_LSB = 0      # north_star allows 1; we hold 0
DEPTH_REL_TOL = 0.0    # north_star allows 1e-6 relative; we hold bit-exact


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def assert_same(color, depth, ref_color, ref_depth):
    assert color.shape == ref_color.shape
    if color.dtype.kind == "f":
        assert np.array_equal(cases.bits(color), cases.bits(ref_color))
    else:
        d = np.abs(color.astype(np.int64) - ref_color.astype(np.int64))
        assert d.max(initial=0) <= COLOR_TOL_LSB, f"{int((d > 0).sum())} colour values differ, max {int(d.max())}"
    nd = int((cases.bits(depth) != cases.bits(ref_depth)).sum())
    assert nd == 0, f"{nd} depth texels differ"


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_cuda_matches_golden_reference_frames(name, golden):
    frames, hashes = golden
    scene = cases.CASES[name]()
    color, depth, stats = cases.render("cuda", scene, rcp_table=frames["rcp_table"])
    h = hashes[name]
    assert stats["kernelLaunches"] > 0
    assert stats["fragsShaded"] == h["frags"]
    assert sha(color) == h["color_sha256"]
    assert sha(depth) == h["depth_sha256"]
    if name in cases.FULL_IMAGE_CASES:
        assert_same(color, depth, frames[name + ".color"], frames[name + ".depth"].view(depth.dtype))


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_cuda_matches_oracle(name, golden):
    frames, _ = golden
    scene = cases.CASES[name]()
    cg, dg, sg = cases.render("cuda", scene, rcp_table=frames["rcp_table"])
    co, do, so = cases.render("port", scene, rcp_table=frames["rcp_table"])
    assert sg["fragsShaded"] == so["fragsShaded"]
    assert sg["trisBinned"] == so["trisBinned"]
    assert sg["spans"] == so["spans"]
    assert_same(cg, dg, co, do)


@pytest.mark.parametrize("name", sorted(cases.MRT_CASES))
def test_cuda_mrt_matches_golden_and_oracle(name, golden):
    """Multiple render targets: every attachment bit-exact against the golden hashes (unmodified reference) and the oracle."""
    import json
    frames, _ = golden
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_hashes.json")))["mrt"][name]
    scene = cases.mrt_scene(name)
    cg, dg, sg = cases.render_mrt("cuda", scene, rcp_table=frames["rcp_table"])
    co, do, so = cases.render_mrt("port", scene, rcp_table=frames["rcp_table"])
    assert sg["fragsShaded"] == so["fragsShaded"] == want["frags"]
    assert len(cg) == len(co) == len(want["colors_sha256"])
    for a, b, h in zip(cg, co, want["colors_sha256"]):
        assert np.array_equal(cases.bits(a) if a.dtype.kind == "f" else a, cases.bits(b) if b.dtype.kind == "f" else b)
        assert sha(a) == h
    assert sha(dg) == want["depth_sha256"]


@pytest.mark.parametrize("name", sorted(cases.BIG_CASES))
def test_cuda_full_size_configs_match_golden(name, golden):
    frames, hashes = golden
    scene = cases.BIG_CASES[name]()
    color, depth, stats = cases.render("cuda", scene, rcp_table=frames["rcp_table"])
    h = hashes[name]
    assert stats["fragsShaded"] == h["frags"]
    assert sha(color) == h["color_sha256"]
    assert sha(depth) == h["depth_sha256"]


@pytest.mark.skipif(not os.path.exists(A.LIB_PATHS["ref_strict"]), reason="oracle/_ref did not travel")
@pytest.mark.parametrize("name", ["draw_test", "c1_small", "c2_small", "rand_clip_heavy", "rand_blend2", "c1_sphere100k"])
def test_cuda_matches_live_reference_on_this_host(name):
    """Default mode: the rcp table is measured from this box's CPU, so the CUDA frame equals what the
    reference renders on the same box."""
    scene = (cases.CASES.get(name) or cases.BIG_CASES[name])()
    cg, dg, sg = cases.render("cuda", scene)
    cr, dr, sr = cases.render("ref_strict", scene)
    assert sg["fragsShaded"] == sr["fragsShaded"]
    assert_same(cg, dg, cr, dr)


def test_idempotent_and_deterministic_full_c2(golden):
    """Rendering the same frame twice (and on a second context) gives identical bits: no order races."""
    frames, _ = golden
    scene = S.SCENES["c2_instancing10k"]()
    a = cases.render("cuda", scene, rcp_table=frames["rcp_table"])
    b = cases.render("cuda", scene, rcp_table=frames["rcp_table"])
    assert np.array_equal(a[0], b[0]) and np.array_equal(cases.bits(a[1]), cases.bits(b[1]))
    assert a[2]["fragsShaded"] == b[2]["fragsShaded"]


@pytest.mark.parametrize("name", ["c3_small", "rand_split3", "rand_multi3", "rand_overflow_blend", "c2_small"])
def test_batched_draws_equal_draw_by_draw(name, golden, monkeypatch):
    """Consecutive draws with one shader + framebuffer are executed as ONE pipeline pass; with
    SLCU_SYNC_DRAWS=1 every draw call is its own pass, like the reference.  Both give the same bits, and the batched
    context really did batch."""
    frames, hashes = golden
    scene = cases.CASES[name]()
    cb, db, sb = cases.render("cuda", scene, rcp_table=frames["rcp_table"])
    monkeypatch.setenv("SLCU_SYNC_DRAWS", "1")
    cs, ds, ss = cases.render("cuda", scene, rcp_table=frames["rcp_table"])
    assert ss["batches"] == ss["drawCalls"] or scene.draws_per_frame == 1
    assert sb["batches"] <= ss["batches"]
    if name == "c3_small":
        assert sb["batches"] == 1 and sb["drawCalls"] == 16
    assert sb["fragsShaded"] == ss["fragsShaded"] == hashes[name]["frags"]
    assert sha(cb) == sha(cs) == hashes[name]["color_sha256"]
    assert sha(db) == sha(ds) == hashes[name]["depth_sha256"]


def test_uniform_updates_between_batched_draws_are_snapshotted():
    """ubo_assign between two draws of a batch must not leak into the first draw (the reference's draw is synchronous)."""
    ctx = A.Context("cuda")
    c = ctx.create_texture(A.SL_COLOR_RGBA_8U, 64, 64)
    d = ctx.create_texture(A.SL_COLOR_R_FLOAT, 64, 64)
    fbo = ctx.create_framebuffer([c], d)
    tri = np.array([[-1, -1, 0, 1], [1, -1, 0, 1], [0, 1, 0, 1]], dtype=np.float32)
    left = tri.copy(); left[:, 0] = left[:, 0] * 0.5 - 0.5
    right = tri.copy(); right[:, 0] = right[:, 0] * 0.5 + 0.5
    verts = np.zeros((6, 8), dtype=np.float32)
    verts[:3, :4], verts[3:, :4] = left, right
    vbo = ctx.create_vbo(verts)
    vao = ctx.create_vao(vbo, None, [(0, 32, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_FLOAT), (16, 32, A.VERTEX_DIMENSION_4, A.VERTEX_DATA_FLOAT)])
    ident = np.identity(4, dtype=np.float32)

    def blob(col):
        return ident.tobytes() + np.array(col, dtype=np.float32).tobytes()

    ubo = ctx.create_ubo(blob((1, 0, 0, 1)))
    sh = ctx.create_shader("flat", "flat", ubo, vs_varyings=0, fs_varyings=0, cull=A.SL_CULL_OFF, blend=A.SL_BLEND_OFF,
                           depth_test=A.SL_DEPTH_TEST_OFF, depth_mask=A.SL_DEPTH_MASK_OFF)
    assert sh != A.INVALID_ID
    ctx.clear_framebuffer(fbo, 0, (0, 0, 0, 1), 1.0)
    ctx.draw(ctx.mesh(vao, 0, 3, A.RENDER_MODE_TRIANGLES), sh, fbo)
    ctx.ubo_assign(ubo, blob((0, 0, 1, 1)))
    ctx.draw(ctx.mesh(vao, 3, 6, A.RENDER_MODE_TRIANGLES), sh, fbo)
    img = ctx.texture_read(c, A.SL_COLOR_RGBA_8U, 64, 64)
    st = ctx.get_stats()
    assert st["batches"] == 1 and st["drawCalls"] == 2
    assert tuple(img[20, 16]) == (255, 0, 0, 255)
    assert tuple(img[20, 48]) == (0, 0, 255, 255)


def test_clear_only_frame():
    ctx = A.Context("cuda")
    c = ctx.create_texture(A.SL_COLOR_RGBA_8U, 1920, 1080)
    d = ctx.create_texture(A.SL_COLOR_R_HALF, 1920, 1080)
    fbo = ctx.create_framebuffer([c], d)
    ctx.clear_framebuffer(fbo, 0, (0.6, 0.6, 0.6, 1.0), 0.25)
    col = ctx.texture_read(c, A.SL_COLOR_RGBA_8U, 1920, 1080)
    dep = ctx.texture_read(d, A.SL_COLOR_R_HALF, 1920, 1080)
    assert np.all(col == np.array([153, 153, 153, 255], dtype=np.uint8))
    assert np.all(dep == np.float16(0.25))


def test_unregistered_shader_is_refused():
    ctx = A.Context("cuda")
    ubo = ctx.create_ubo(b"\0" * 16)
    assert ctx.create_shader("nope", "vcolor", ubo, vs_varyings=1, fs_varyings=1) == A.INVALID_ID


def test_empty_point_draw_is_a_noop():
    """The unmodified reference dies with SIGFPE on a point draw without elements, so there is no golden; here it leaves the
    cleared frame untouched."""
    scene = S.random_tris_scene(seed=5, n=30, points=True, empty_draw=True)
    color, depth, stats = cases.render("cuda", scene)
    assert stats["fragsShaded"] == 0
    assert np.all(color == color[0, 0]) and np.all(depth == 1.0)


@pytest.mark.parametrize("mode", [0x1234, 0x02])
def test_unimplemented_render_modes_are_refused_loudly(mode):
    ctx = A.Context("cuda")
    st = S.SceneInstance(S.random_tris_scene(seed=2, n=10), ctx)
    m = st.state["meshes"][0]
    with pytest.raises(A.SlError):
        ctx.draw(ctx.mesh(m.vaoId, m.elementBegin, m.elementEnd, mode), st.state["shader"], st.state["fbo"])


def test_lines_between_triangle_draws_keep_submission_order(golden):
    """triangles, lines, triangles on one framebuffer: the line pass executes between the two triangle batches, exactly
    like the oracle's draw-by-draw execution."""
    frames, _ = golden

    def run(impl):
        ctx = A.Context(impl)
        ctx.set_rcp_table(frames["rcp_table"])
        ctx.set_rsqrt_table(frames["rsqrt_table"])
        tri = S.SceneInstance(S.random_tris_scene(seed=41, n=300, blend=1), ctx)
        st = tri.state
        ctx.clear_framebuffer(st["fbo"], 0, (0.1, 0.2, 0.3, 1.0), 1.0)
        m = st["meshes"][0]
        third = (m.elementEnd - m.elementBegin) // 9 * 3
        ctx.draw(ctx.mesh(m.vaoId, 0, third, A.RENDER_MODE_INDEXED_TRIANGLES), st["shader"], st["fbo"])
        ctx.draw(ctx.mesh(m.vaoId, 0, m.elementEnd, A.RENDER_MODE_INDEXED_LINES), st["shader"], st["fbo"])
        ctx.draw(ctx.mesh(m.vaoId, third, 2 * third, A.RENDER_MODE_INDEXED_TRIANGLES), st["shader"], st["fbo"])
        ctx.sync()
        return tri.read_color(), tri.read_depth()

    cg, dg = run("cuda")
    co, do = run("port")
    assert_same(cg, dg, co, do)


def test_odd_line_element_count_past_the_index_buffer_is_refused():
    """for (i = begin; i < end; i += 2) fetches element `end` for the last segment (src/SL_LineProcessor.cpp:236-251).  Inside
    the index buffer that is the next element (reproduced: lines_multi4_u32_double_ge has odd sub-draws); past its end the
    reference draws whatever its heap holds there -- the product refuses the draw instead of fetching through it."""
    ctx = A.Context("cuda")
    st = S.SceneInstance(S.random_tris_scene(seed=3, n=11), ctx).state  # 33 elements
    m = st["meshes"][0]
    assert (m.elementEnd - m.elementBegin) == 33
    ctx.draw(ctx.mesh(m.vaoId, 0, 31, A.RENDER_MODE_INDEXED_LINES), st["shader"], st["fbo"])  # element 31 exists
    ctx.sync()
    with pytest.raises(A.SlError):
        ctx.draw(ctx.mesh(m.vaoId, 0, 33, A.RENDER_MODE_INDEXED_LINES), st["shader"], st["fbo"])
    ctx.close()


def test_points_between_triangle_draws_keep_submission_order(golden):
    """triangles, points, triangles on one framebuffer: the point pass executes between the two triangle batches, exactly
    like the oracle's draw-by-draw execution."""
    frames, _ = golden

    def run(impl):
        ctx = A.Context(impl)
        ctx.set_rcp_table(frames["rcp_table"])
        tri = S.SceneInstance(S.random_tris_scene(seed=31, n=300, blend=1), ctx)
        st = tri.state
        ctx.clear_framebuffer(st["fbo"], 0, (0.1, 0.2, 0.3, 1.0), 1.0)
        m = st["meshes"][0]
        third = (m.elementEnd - m.elementBegin) // 9 * 3
        ctx.draw(ctx.mesh(m.vaoId, 0, third, A.RENDER_MODE_INDEXED_TRIANGLES), st["shader"], st["fbo"])
        ctx.draw(ctx.mesh(m.vaoId, 0, m.elementEnd, A.RENDER_MODE_INDEXED_POINTS), st["shader"], st["fbo"])
        ctx.draw(ctx.mesh(m.vaoId, third, m.elementEnd, A.RENDER_MODE_INDEXED_TRIANGLES), st["shader"], st["fbo"])
        c = ctx.texture_read(st["color"], A.SL_COLOR_RGBA_8U, 200, 150)
        d = ctx.texture_read(st["depth"], A.SL_COLOR_R_FLOAT, 200, 150)
        n = ctx.get_stats()["fragsShaded"]
        ctx.close()
        return c, d, n

    a, b = run("cuda"), run("port")
    assert np.array_equal(a[0], b[0]) and np.array_equal(cases.bits(a[1]), cases.bits(b[1])) and a[2] == b[2]


@pytest.mark.parametrize("name,world,strip", [("c1_small", 2, 8), ("rand_clip_heavy", 2, 1), ("rand_overflow_wide", 3, 2),
                                               ("rand_overflow_blend", 4, 1), ("c2_small", 8, 8), ("rand_overflow_half_ge", 2, 1),
                                               ("points_dense", 2, 1), ("points_sparse_1080p", 4, 8),
                                               ("wire_alpha", 2, 1), ("wire_overflow_premult", 3, 2), ("wire_overflow_half_ge", 2, 1)])
def test_sort_first_partition_is_bit_identical(name, world, strip, golden):
    """Every rank's context (all on this GPU) rasterises only its row strips (+ the rows its hand-overs come from);
    the composite of the owned rows equals the unpartitioned frame bit for bit, spans that wrap across an ownership
    boundary included."""
    from softlight_b200 import sortfirst as SF

    frames, hashes = golden
    scene = cases.CASES[name]()
    ctxs, insts, finals_c, finals_d = [], [], [], []
    for r in range(world):
        ctx = A.Context("cuda")
        ctx.set_rcp_table(frames["rcp_table"])
        ctx.set_partition(r, world, strip)
        inst = S.SceneInstance(scene, ctx)
        inst.render()
        # local composite path (root == rank): owned rows -> a zero-initialised texture
        fc = ctx.create_texture(scene.color_type, scene.width, scene.height)
        fd = ctx.create_texture(scene.depth_type, scene.width, scene.height)
        ctx.composite(inst.state["color"], fc, r)
        ctx.composite(inst.state["depth"], fd, r)
        ctx.sync()
        finals_c.append(ctx.texture_read(fc, scene.color_type, scene.width, scene.height))
        finals_d.append(ctx.texture_read(fd, scene.depth_type, scene.width, scene.height)[..., 0])
        mine = SF.owned_rows(scene.height, r, world, strip)
        assert not finals_c[-1][~mine].any(), "composite wrote a row this rank does not own"
        ctxs.append(ctx)
    color = SF.merge_owned_rows(finals_c, strip)
    depth = SF.merge_owned_rows(finals_d, strip)
    assert sha(color) == hashes[name]["color_sha256"]
    assert sha(depth) == hashes[name]["depth_sha256"]


@pytest.mark.parametrize("name,world,strip", [("c1_small", 2, 8), ("c2_small", 4, 8), ("rand_clip_heavy", 2, 1), ("rand_overflow_wide", 3, 2),
                                               ("rand_overflow_blend", 4, 1), ("rand_multi3", 2, 4), ("rand_split3", 3, 8),
                                               ("rand_index_u8", 2, 2), ("rand_depthfunc3", 4, 1), ("c4_small", 2, 8)])
def test_shared_vertex_stage_is_bit_identical(name, world, strip, golden, monkeypatch):
    """The ranks split the vertex stage (slcu_geometry_share_*): each shades every world-th vertex block and stores the
    records into the arenas of the ranks whose rows they touch; the device-side rendezvous replaces the host.  All ranks
    are contexts on this one GPU, one host thread each (a rank waits on the device for its peers' blocks), their arenas
    attached by address.  The composite of the owned rows equals the unpartitioned frame bit for bit, and every rank
    reports the whole frame's record count."""
    import threading

    from softlight_b200 import sortfirst as SF

    if name not in cases.CASES:
        pytest.skip("case not defined")
    monkeypatch.setenv("SLCU_GEOM_SHARE_MIN_PRIMS", "0")  # by default only batches of millions of primitives are shared
    frames, hashes = golden
    scene = cases.CASES[name]()
    ctxs = []
    for r in range(world):
        ctx = A.Context("cuda")
        ctx.set_rcp_table(frames["rcp_table"])
        ctx.set_partition(r, world, strip)
        ctx.geometry_share_create(96 << 20)
        ctxs.append(ctx)
    for r in range(world):
        for t in range(world):
            if t != r:
                ctxs[r].geometry_share_attach(t, ctxs[t].geometry_share_address())
    finals_c, finals_d, stats, errors = [None] * world, [None] * world, [None] * world, []
    ready = threading.Barrier(world)

    def run(r):
        try:
            ctx = ctxs[r]
            inst = S.SceneInstance(scene, ctx)
            ctx.sync()
            ready.wait(timeout=60)  # every rank's resources exist before any rank waits on the device for its peers
            for _ in range(2):  # the second frame runs with the sizes the first one learned: no refusal, no local re-run
                ctx.reset_stats()
                inst.render()
                ctx.sync()
            stats[r] = ctx.get_stats()
            fc = ctx.create_texture(scene.color_type, scene.width, scene.height)
            fd = ctx.create_texture(scene.depth_type, scene.width, scene.height)
            ctx.composite(inst.state["color"], fc, r)
            ctx.composite(inst.state["depth"], fd, r)
            ctx.sync()
            finals_c[r] = ctx.texture_read(fc, scene.color_type, scene.width, scene.height)
            finals_d[r] = ctx.texture_read(fd, scene.depth_type, scene.width, scene.height)[..., 0]
        except Exception as ex:  # noqa: BLE001
            errors.append((r, repr(ex)))

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    assert all(f is not None for f in finals_c), "a rank did not finish"
    color = SF.merge_owned_rows(finals_c, strip)
    depth = SF.merge_owned_rows(finals_d, strip)
    assert sha(color) == hashes[name]["color_sha256"]
    assert sha(depth) == hashes[name]["depth_sha256"]
    binned = {st["trisBinned"] for st in stats}
    assert len(binned) == 1, f"ranks disagree on the frame's record count: {binned}"
