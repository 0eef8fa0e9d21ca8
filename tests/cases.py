"""Named parity cases shared by the CPU (oracle pinning) and GPU (CUDA parity) tests."""
import numpy as np

from softlight_b200 import capi as A
from softlight_b200 import scenes as S


def _variants():
    v = []
    for dt in (A.SL_DEPTH_TEST_OFF, A.SL_DEPTH_TEST_LESS_THAN, A.SL_DEPTH_TEST_LESS_EQUAL, A.SL_DEPTH_TEST_GREATER_THAN,
               A.SL_DEPTH_TEST_GREATER_EQUAL, A.SL_DEPTH_TEST_EQUAL, A.SL_DEPTH_TEST_NOT_EQUAL):
        v.append((f"depthfunc{dt}", dict(depth_test=dt, clear_depth=0.5)))
    for b in range(5):
        v.append((f"blend{b}", dict(blend=b)))
    for d, n in ((A.SL_COLOR_R_HALF, "half"), (A.SL_COLOR_R_FLOAT, "float"), (A.SL_COLOR_R_DOUBLE, "double")):
        v.append((f"depth_{n}_premult", dict(depth_type=d, blend=2)))
    for ct in (A.SL_COLOR_RGBA_8U, A.SL_COLOR_RGBA_16U, A.SL_COLOR_RGBA_HALF, A.SL_COLOR_RGBA_FLOAT, A.SL_COLOR_RGB_8U,
               A.SL_COLOR_RG_8U, A.SL_COLOR_R_8U, A.SL_COLOR_RGB_FLOAT, A.SL_COLOR_R_FLOAT, A.SL_COLOR_RGBA_DOUBLE):
        v.append((f"color{ct}", dict(color_type=ct)))
        v.append((f"color{ct}_alpha", dict(color_type=ct, blend=1)))
    for c in range(3):
        v.append((f"cull{c}", dict(cull=c)))
    for idt, n in ((None, "none"), (np.uint8, "u8"), (np.uint16, "u16"), (np.uint32, "u32")):
        v.append((f"index_{n}", dict(index_dtype=idt, n=80)))
    v.append(("viewport", dict(viewport=(20, 10, 120, 90))))
    v.append(("scissor", dict(scissor=(30, 20, 100, 80))))
    v.append(("viewport_scissor", dict(viewport=(20, 10, 120, 90), scissor=(30, 20, 100, 80))))
    v.append(("split3", dict(split_draws=3)))
    v.append(("multi3", dict(split_draws=3, multiple=True)))
    v.append(("clip_heavy", dict(n=3000, width=320, height=200, clip_fraction=0.5)))
    v.append(("nomask_additive", dict(depth_mask=A.SL_DEPTH_MASK_OFF, blend=3)))
    v.append(("empty", dict(n=30, empty_draw=True)))
    v.append(("ragged_odd_size", dict(n=300, width=333, height=77)))
    v.append(("tiny_1x1", dict(n=50, width=1, height=1)))
    # clip-generated triangles that end right of / above the viewport: fragments wrap into the next row / the padding rows
    v.append(("overflow_wide", dict(n=900, width=640, height=48, clip_fraction=0.8)))
    v.append(("overflow_blend", dict(n=700, width=257, height=131, clip_fraction=0.9, blend=1)))
    v.append(("overflow_half_ge", dict(n=700, width=512, height=64, clip_fraction=0.9, depth_type=A.SL_COLOR_R_HALF,
                                       depth_test=A.SL_DEPTH_TEST_GREATER_EQUAL, clear_depth=-1.0)))
    # the remaining colour formats the path supports (every 8/16-bit unorm, half, float and double layout; 32/64-bit
    # integer targets are refused): appended last so that the seeds of the cases above do not move
    for ct in (A.SL_COLOR_R_16U, A.SL_COLOR_R_HALF, A.SL_COLOR_R_DOUBLE, A.SL_COLOR_RG_16U, A.SL_COLOR_RG_HALF, A.SL_COLOR_RG_FLOAT,
               A.SL_COLOR_RG_DOUBLE, A.SL_COLOR_RGB_16U, A.SL_COLOR_RGB_HALF, A.SL_COLOR_RGB_DOUBLE):
        v.append((f"color{ct}", dict(color_type=ct)))
        v.append((f"color{ct}_premult", dict(color_type=ct, blend=2)))
    # extreme framebuffer shape: the row binner's height limit (SL_MAX_ROWS = 8192).  (A 16000-pixel-wide stress case is NOT
    # here: with clip-generated spans thousands of pixels long, the reference enters the behaviours DESIGN.md section 2 lists
    # as not reproduced -- its oracle restatement already differs there; widths up to 7680 (config C5) are pinned.)
    v.append(("tall_8192", dict(n=600, width=48, height=8192, clip_fraction=0.4)))
    # 32- and 64-bit integer colour targets: the reference converts with cvtps2dq (0.5 and above become 0x80000000) and with
    # g++'s float -> uint64 sequence (1.0 becomes 0); reproduced as they are
    for ct in (A.SL_COLOR_R_32U, A.SL_COLOR_R_64U, A.SL_COLOR_RG_32U, A.SL_COLOR_RG_64U, A.SL_COLOR_RGB_32U, A.SL_COLOR_RGB_64U,
               A.SL_COLOR_RGBA_32U, A.SL_COLOR_RGBA_64U):
        v.append((f"color{ct}", dict(color_type=ct)))
        v.append((f"color{ct}_alpha", dict(color_type=ct, blend=1)))
    # the packed colour types as render targets (src/SL_Framebuffer.cpp:85-145, 371-590): rgba_cast both ways, incl. the
    # sign-extending AVX2 forms of RGBA1010102
    for ct in (A.SL_COLOR_RGB_332, A.SL_COLOR_RGB_565, A.SL_COLOR_RGBA_5551, A.SL_COLOR_RGBA_4444, A.SL_COLOR_RGBA_1010102):
        v.append((f"color{ct}", dict(color_type=ct)))
        v.append((f"color{ct}_alpha", dict(color_type=ct, blend=1)))
        v.append((f"color{ct}_premult", dict(color_type=ct, blend=2)))
    return v


# multiple render targets (sl_mrt_test.cpp attaches RGB8, RGB8, RG8, RGB8 + half depth): compared attachment by attachment
MRT_CASES = {
    "mrt_rgb8x3_rg8_half": dict(n=500, width=211, height=160, color_type=A.SL_COLOR_RGB_8U,
                                mrt_types=[A.SL_COLOR_RGB_8U, A.SL_COLOR_RG_8U, A.SL_COLOR_RGB_8U], depth_type=A.SL_COLOR_R_HALF),
    "mrt_mixed_alpha": dict(n=400, width=200, height=150, color_type=A.SL_COLOR_RGBA_8U, blend=1,
                            mrt_types=[A.SL_COLOR_RGBA_FLOAT, A.SL_COLOR_RGBA_16U, A.SL_COLOR_RGBA_HALF]),
    "mrt_two_noblend_4seg": dict(n=1500, width=1100, height=64, color_type=A.SL_COLOR_RGBA_8U, mrt_types=[A.SL_COLOR_R_FLOAT]),
}


def mrt_scene(name):
    return S.random_tris_scene(seed=700 + sorted(MRT_CASES).index(name), name=name, **MRT_CASES[name])


def render_mrt(impl: str, scene, rcp_table=None):
    ctx = A.Context(impl)
    if rcp_table is not None and not impl.startswith("ref"):
        ctx.set_rcp_table(rcp_table)
    ctx.reset_stats()
    inst = S.SceneInstance(scene, ctx)
    inst.render()
    ctx.sync()
    colors, depth, stats = inst.read_colors(), inst.read_depth(), ctx.get_stats()
    ctx.close()
    return colors, depth, stats


CASES = {}
for _i, (_name, _kw) in enumerate(_variants()):
    CASES["rand_" + _name] = (lambda i=_i, kw=_kw, name=_name: S.random_tris_scene(seed=100 + i, name="rand_" + name, **kw))

# RENDER_MODE_POINTS / INDEXED_POINTS (SL_PointProcessor + SL_PointRasterizer): many points per pixel, so the per-pixel
# submission order decides every depth test and blend
POINT_CASES = {
    "points_dense": dict(n=4000, width=64, height=48),
    "points_alpha_nodepth": dict(n=4000, width=64, height=48, blend=1, depth_test=A.SL_DEPTH_TEST_OFF, depth_mask=A.SL_DEPTH_MASK_OFF),
    "points_premult_half_le": dict(n=3000, width=100, height=50, blend=2, depth_type=A.SL_COLOR_R_HALF, depth_test=A.SL_DEPTH_TEST_LESS_EQUAL),
    "points_noindex_split3": dict(n=3000, width=200, height=150, index_dtype=None, split_draws=3),
    "points_multi4_u32": dict(n=3000, width=100, height=50, multiple=True, split_draws=4, index_dtype=np.uint32),
    "points_instanced3_additive": dict(n=1000, width=100, height=50, instances=3, blend=3),
    "points_viewport_scissor": dict(n=3000, width=100, height=50, viewport=(10, 5, 80, 40), scissor=(20, 10, 50, 30)),
    "points_rgbf_double_ge": dict(n=2000, width=77, height=33, color_type=A.SL_COLOR_RGB_FLOAT, depth_type=A.SL_COLOR_R_DOUBLE,
                                  depth_test=A.SL_DEPTH_TEST_GREATER_EQUAL, clear_depth=-1.0),
    "points_sparse_1080p": dict(n=20000, width=1920, height=1080),
    "points_one_pixel": dict(n=700, width=1, height=1, blend=1, depth_test=A.SL_DEPTH_TEST_OFF, depth_mask=A.SL_DEPTH_MASK_OFF),
    # (no empty point draw here: the unmodified reference dies with SIGFPE on one -- a modulo by the element count in its
    # partitioning -- so there is nothing to pin; the CUDA path treats it as a no-op, tests/test_gpu_parity.py)
}
for _i, _name in enumerate(POINT_CASES):
    CASES[_name] = (lambda i=_i, name=_name: S.random_tris_scene(seed=900 + i, name=name, points=True, **POINT_CASES[name]))

# RENDER_MODE_TRI_WIRE / INDEXED_TRI_WIRE (SL_TriRasterizer::render_wireframe): the triangle scenes, edges only.  Blending
# makes the rasteriser's pixel revisits visible (a revisited pixel is blended twice), the overflow cases wrap edge runs
# across the row end.
WIRE_CASES = {
    "wire_basic": dict(n=400),
    "wire_alpha": dict(n=400, blend=1),
    "wire_noindex_split3_additive": dict(n=300, index_dtype=None, split_draws=3, blend=3, depth_mask=A.SL_DEPTH_MASK_OFF),
    "wire_multi3_u32": dict(n=300, index_dtype=np.uint32, split_draws=3, multiple=True),
    "wire_clip_heavy": dict(n=3000, width=320, height=200, clip_fraction=0.5),
    "wire_overflow_wide": dict(n=900, width=640, height=48, clip_fraction=0.8),
    "wire_overflow_premult": dict(n=700, width=257, height=131, clip_fraction=0.9, blend=2),
    "wire_overflow_half_ge": dict(n=700, width=512, height=64, clip_fraction=0.9, depth_type=A.SL_COLOR_R_HALF,
                                  depth_test=A.SL_DEPTH_TEST_GREATER_EQUAL, clear_depth=-1.0),
    "wire_nodepth_rgbf": dict(n=500, depth_test=A.SL_DEPTH_TEST_OFF, color_type=A.SL_COLOR_RGB_FLOAT, cull=1),
    "wire_viewport_scissor": dict(n=400, viewport=(20, 10, 120, 90), scissor=(30, 20, 100, 80)),
    "wire_1080p": dict(n=200, width=1920, height=1080, clip_fraction=0.2),
    "wire_tiny_1x1": dict(n=50, width=1, height=1),
    "wire_empty": dict(n=30, empty_draw=True),
    "wire_instanced3_alpha": dict(n=200, instances=3, blend=1, depth_test=A.SL_DEPTH_TEST_LESS_EQUAL),
}
for _i, _name in enumerate(WIRE_CASES):
    CASES[_name] = (lambda i=_i, name=_name: S.random_tris_scene(seed=950 + i, name=name, wire=True, **WIRE_CASES[name]))
# RENDER_MODE_LINES / INDEXED_LINES (SL_LineProcessor + SL_LineRasterizer): consecutive element pairs of the same random
# vertices are segments; clip_fraction puts vertices behind the eye (segment dropped) or outside the frustum (Liang-Barsky
# against the viewport rectangle), blending makes the per-texel submission order visible.
LINE_CASES = {
    "lines_basic": dict(n=300),
    "lines_alpha_clip": dict(n=300, width=64, height=48, blend=1, clip_fraction=0.6),
    "lines_premult_half_le": dict(n=2000, width=320, height=200, clip_fraction=0.5, blend=2, depth_type=A.SL_COLOR_R_HALF,
                                  depth_test=A.SL_DEPTH_TEST_LESS_EQUAL),
    "lines_noindex_split3_additive": dict(n=500, index_dtype=None, split_draws=3, blend=3, depth_mask=A.SL_DEPTH_MASK_OFF),
    "lines_viewport_scissor": dict(n=500, viewport=(20, 10, 120, 90), scissor=(30, 20, 100, 80)),
    "lines_instanced3_alpha": dict(n=300, instances=3, blend=1, depth_test=A.SL_DEPTH_TEST_LESS_EQUAL),
    "lines_1080p": dict(n=200, width=1920, height=1080, clip_fraction=0.2),
    "lines_tiny_1x1": dict(n=50, width=1, height=1),
    "lines_nodepth_small": dict(n=700, width=37, height=21, clip_fraction=0.9, blend=1, depth_test=A.SL_DEPTH_TEST_OFF),
    "lines_multi4_u32_double_ge": dict(n=333, width=257, height=131, multiple=True, split_draws=4, index_dtype=np.uint32,
                                       depth_type=A.SL_COLOR_R_DOUBLE, depth_test=A.SL_DEPTH_TEST_GREATER_EQUAL, clear_depth=-1.0),
    "lines_u8_rgbf": dict(n=80, index_dtype=np.uint8, color_type=A.SL_COLOR_RGB_FLOAT),
    "lines_empty": dict(n=30, empty_draw=True),
    # NOT_EQUAL through the functors' scalar overload (`a != b`, SL_ShaderUtil.hpp:471-476): TRUE for the NaN depth of a
    # clipped zero-length line, where the SIMD overload the filled triangles use (_CMP_NEQ_OQ) says false
    "lines_ne_nan": dict(n=1400, width=364, height=362, clip_fraction=0.55, depth_test=A.SL_DEPTH_TEST_NOT_EQUAL, clear_depth=0.0),
    "lines_ne_nan_premult_nomask": dict(n=1200, width=228, height=155, clip_fraction=0.6, blend=2, depth_mask=A.SL_DEPTH_MASK_OFF,
                                        depth_test=A.SL_DEPTH_TEST_NOT_EQUAL, depth_type=A.SL_COLOR_R_DOUBLE, viewport=(13, 20, 191, 120)),
    "lines_ne_nan_half_1010102": dict(n=1000, width=110, height=56, clip_fraction=0.4, blend=4, depth_test=A.SL_DEPTH_TEST_NOT_EQUAL,
                                      depth_type=A.SL_COLOR_R_HALF, color_type=A.SL_COLOR_RGBA_1010102, split_draws=2),
}
for _i, _name in enumerate(LINE_CASES):
    CASES[_name] = (lambda i=_i, name=_name: S.random_tris_scene(seed=1000 + i, name=name, lines=True, **LINE_CASES[name]))
CASES["draw_test"] = S.SCENES["draw_test"]
CASES["c1_small"] = S.SCENES["c1_small"]
CASES["c2_small"] = S.SCENES["c2_small"]
CASES["c3_small"] = S.SCENES["c3_small"]
CASES["c4_small"] = S.SCENES["c4_small"]
CASES["c5_small"] = S.SCENES["c5_small"]
CASES["sampler_matrix"] = S.SCENES["sampler_matrix"]

# full-size configs: compared through hashes / live checkers only
BIG_CASES = {
    "c1_sphere100k": S.SCENES["c1_sphere100k"],
    "c2_instancing10k": S.SCENES["c2_instancing10k"],
    "c3_indoor2m": S.SCENES["c3_indoor2m"],
    "c4_skinned100": S.SCENES["c4_skinned100"],
    "c5_volume4k": S.SCENES["c5_volume4k"],
    "c5_volume8k": S.SCENES["c5_volume8k"],
    "c1_sphere100k_4k": S.SCENES["c1_sphere100k_4k"],
    "c1_sphere1m_4k": S.SCENES["c1_sphere1m_4k"],
}

# cases whose images are committed in full (the rest of CASES is pinned by sha256)
FULL_IMAGE_CASES = ["draw_test", "c1_small", "c2_small", "c3_small", "c4_small", "c5_small", "sampler_matrix", "rand_blend1", "rand_clip_heavy", "rand_depth_half_premult", "points_dense", "wire_alpha", "lines_alpha_clip"]


def bits(a: np.ndarray) -> np.ndarray:
    return a.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize])


_GOLDEN_RSQRT = None


def golden_rsqrt_table():
    """The rsqrtss table of the CPU the goldens were generated on (stored next to the rcp table)."""
    global _GOLDEN_RSQRT
    if _GOLDEN_RSQRT is None:
        import os

        f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_frames.npz"))
        _GOLDEN_RSQRT = f["rsqrt_table"]
    return _GOLDEN_RSQRT


def render(impl: str, scene, rcp_table=None, device=0, threads=1):
    ctx = A.Context(impl, device=device, threads=threads)
    if rcp_table is not None and not impl.startswith("ref"):
        ctx.set_rcp_table(rcp_table)
        ctx.set_rsqrt_table(golden_rsqrt_table())  # pinned together: both tables belong to the CPU that made the goldens
    ctx.reset_stats()
    inst = S.SceneInstance(scene, ctx)
    inst.render()
    ctx.sync()
    color, depth, stats = inst.read_color(), inst.read_depth(), ctx.get_stats()
    ctx.close()
    return color, depth, stats
