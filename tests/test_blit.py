"""Blit parity: SL_Context::blit -> SL_BlitProcessor::blit_nearest (src/SL_BlitProcessor.cpp:620-715).
CPU: the oracle restatement against golden outputs of the unmodified reference (and the live reference when present).
GPU: k_blit through the C ABI against the same goldens.  Bit-exact (integer index math + defined casts)."""
import hashlib
import os

import numpy as np
import pytest

import blit_cases
from softlight_b200 import capi as A


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(blit_cases.BLIT_CASES))
def test_port_blit_matches_golden(name, golden_blits):
    assert sha(blit_cases.run_blit("port", name)) == golden_blits[name]


@pytest.mark.skipif(not os.path.exists(A.LIB_PATHS["ref_strict"]), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["rgba8_rect", "rgbaf_to_rgba8_scaled", "rgba8_to_rgbaf"])
def test_port_blit_matches_live_reference(name):
    a, b = blit_cases.run_blit("port", name), blit_cases.run_blit("ref_strict", name)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_blit_flips_vertically_known_answer():
    """srcY = srcY1 - (srcY0 + yf) - 1: row 0 of the destination is the LAST row of the source."""
    ctx = A.Context("port")
    src_px = np.zeros((4, 2, 4), dtype=np.uint8)
    src_px[:, :, 0] = np.arange(4)[:, None]
    src = ctx.create_texture(A.SL_COLOR_RGBA_8U, 2, 4, 1, src_px)
    dst = ctx.create_texture(A.SL_COLOR_RGBA_8U, 2, 4, 1)
    ctx.blit(dst, src)
    out = ctx.texture_read(dst, A.SL_COLOR_RGBA_8U, 2, 4)
    assert out[:, 0, 0].tolist() == [3, 2, 1, 0]


def test_cross_layout_conversion_is_converted():
    """RGB8 -> RGBA16 was refused in round 1; the conversion matrix (tests/test_blit_matrix.py) now covers it."""
    ctx = A.Context("port")
    src = ctx.create_texture(A.SL_COLOR_RGB_8U, 1, 1, 1, np.array([[[255, 0, 128]]], dtype=np.uint8))
    dst = ctx.create_texture(A.SL_COLOR_RGBA_16U, 1, 1, 1)
    ctx.blit(dst, src)
    assert ctx.texture_read(dst, A.SL_COLOR_RGBA_16U, 1, 1).reshape(-1).tolist() == [65535, 0, 128 * 257, 65535]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(blit_cases.BLIT_CASES))
def test_cuda_blit_matches_golden(name, golden_blits):
    assert sha(blit_cases.run_blit("cuda", name)) == golden_blits[name]


@pytest.mark.gpu
@pytest.mark.parametrize("src_type,dst_type", [(A.SL_COLOR_RGBA_8U, A.SL_COLOR_RGBA_8U), (A.SL_COLOR_RGBA_FLOAT, A.SL_COLOR_RGBA_8U),
                                               (A.SL_COLOR_RGB_8U, A.SL_COLOR_RGBA_HALF), (A.SL_COLOR_RGBA_8U, A.SL_COLOR_RGB_16U),
                                               (A.SL_COLOR_RGBA_8U, A.SL_COLOR_RGB_565), (A.SL_COLOR_RGBA_FLOAT, A.SL_COLOR_RGBA_1010102)])
def test_blit_to_host_view_equals_blit_into_a_texture(src_type, dst_type):
    """SL_Context::blit(SL_TextureView& buffer, textureId[, rects]) (src/SL_Context.cpp:816-874), the present path: the host
    view receives exactly what the same blit writes into a texture (pinned to the reference by the blit matrix), whole
    view or rectangle, and every texel outside the destination rectangle keeps its value."""
    rng = np.random.default_rng(5)
    ctx = A.Context("cuda")
    sw, sh, dw, dh = 97, 61, 64, 40
    dt = A.texel_dtype(src_type)
    nch = src_type // 7 + 1
    src = (rng.random((sh, sw, nch)) * (255 if dt == np.uint8 else 1)).astype(dt)
    tin = ctx.create_texture(src_type, sw, sh, 1, src)
    for rects in (None, ((5, 3, 80, 50), (10, 7, 50, 33))):
        tout = ctx.create_texture(dst_type, dw, dh, 1)
        fill = ctx.host_array((dh, dw, A.texel_channels(dst_type)), A.texel_dtype(dst_type))
        fill.view(np.uint8)[...] = 0xAB
        before = fill.copy()
        if rects is None:
            ctx.blit(tout, tin)
            ctx.blit_to_host(fill, dst_type, tin)
            want = ctx.texture_read(tout, dst_type, dw, dh)
            assert np.array_equal(fill.view(np.uint8), want.view(np.uint8))
        else:
            ctx.blit(tout, tin, rects[0] + rects[1])
            ctx.blit_to_host(fill, dst_type, tin, rects[0], rects[1])
            want = ctx.texture_read(tout, dst_type, dw, dh)
            x0, y0, x1, y1 = rects[1]
            assert np.array_equal(fill[y0:y1, x0:x1].view(np.uint8), want[y0:y1, x0:x1].view(np.uint8))
            mask = np.ones((dh, dw), dtype=bool)
            mask[y0:y1, x0:x1] = False
            assert np.array_equal(fill[mask].view(np.uint8), before[mask].view(np.uint8)), "texels outside the rectangle were touched"
    ctx.close()


@pytest.mark.gpu
def test_two_deep_present_delivers_each_frame_to_its_own_view():
    """slcu_present: frame k's blit is enqueued, frame k+1 renders while it crosses the link; after present_wait every host
    view holds its own frame."""
    ctx = A.Context("cuda")
    w, h = 320, 200
    color = ctx.create_texture(A.SL_COLOR_RGBA_8U, w, h, 1)
    depth = ctx.create_texture(A.SL_COLOR_R_FLOAT, w, h, 1)
    fbo = ctx.create_framebuffer([color], depth)
    views = [ctx.host_array((h, w, 4), np.uint8) for _ in range(2)]
    got = []
    for k in range(6):
        slot = k & 1
        ctx.present_wait(slot)
        if k >= 2:
            got.append(views[slot].copy())
        ctx.clear_color_buffer(fbo, 0, (k / 8.0, 0.0, 1.0 - k / 8.0, 1.0))
        ctx.present(slot, views[slot].ctypes.data, A.SL_COLOR_RGBA_8U, w, h, color)
    for slot in ((6 & 1), (7 & 1)):
        ctx.present_wait(slot)
        got.append(views[slot].copy())
    assert len(got) == 6
    for k, frame in enumerate(got):
        want = np.array([int(k / 8.0 * 255.0), 0, int((1.0 - k / 8.0) * 255.0), 255], dtype=np.uint8)
        assert (frame == want).all(), (k, frame[0, 0], want)
    ctx.close()
