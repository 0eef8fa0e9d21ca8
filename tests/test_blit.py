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
