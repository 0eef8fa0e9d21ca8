"""The blit conversion matrix over all 33 colour types, every (source, destination) pair:
* SL_BlitProcessor (softlight/src/SL_BlitProcessor.cpp:20-715 + color_cast, SL_Color.hpp:236-700): the 28 x 28 pairs of
  the plain layouts (R/RG/RGB/RGBA x u8, u16, u32, u64, half, float, double);
* SL_BlitCompressedProcessor (softlight/src/SL_BlitCompressedProcessor.cpp + rgb_cast / rgba_cast, SL_ColorCompressed.hpp):
  the 305 pairs with a packed type (RGB332, RGB565, RGBA5551, RGBA4444, RGBA1010102) on either side.

CPU: the oracle restatement against the compiled reference on all 1089 pairs (in-range and out-of-range component values)
plus an exhaustive run of the packed pairs (every 8- and 16-bit packed value, 65536 RGBA1010102 values, edge-laden plain
sources), and against committed hashes of the reference's output where oracle/_ref is absent.
GPU: k_blit through the C ABI against the oracle on the same inputs."""
import hashlib
import json
import os

import numpy as np
import pytest

from softlight_b200 import capi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KINDS = [0, 1, 2, 3, 4, 5, 6]  # u8, u16, u32, u64, half, float, double
NAMES = {0: "u8", 1: "u16", 2: "u32", 3: "u64", 4: "half", 5: "float", 6: "double"}
PACKED = {28: "rgb332", 29: "rgb565", 30: "rgba5551", 31: "rgba4444", 32: "rgba1010102"}
HAVE_REF = os.path.exists(A.LIB_PATHS["ref_strict"])
GOLDEN = os.path.join(ROOT, "tests", "golden", "blit_matrix_hashes.json")


def source_texels(ctype, w, h, mode):
    rng = np.random.default_rng(ctype * 7 + mode)
    n = A.texel_channels(ctype)
    dt = A.texel_dtype(ctype)
    if np.dtype(dt).kind == "f":
        a = rng.uniform(0.0, 1.0, size=(h, w, n)) if mode == 0 else rng.uniform(-0.5, 1.5, size=(h, w, n))
        a = a.astype(dt)
        a[0, 0, :] = 0
        a[0, 1, :] = 1
        return a
    a = rng.integers(0, np.iinfo(dt).max, size=(h, w, n), endpoint=True, dtype=np.uint64).astype(dt)
    if np.dtype(dt).itemsize >= 4:  # the wide integers: values around the powers of two the x86 conversions branch on
        top = np.iinfo(dt).max
        edges = [0, 1, top, top - 1, top // 2, top // 2 + 1, 2 ** 24 + 1, 2 ** 31, 2 ** 31 - 1, 2 ** 31 + 129, 255, 256, 65535, 65536,
                 (2 ** 53 + 1) & top, (2 ** 63 + 2 ** 39 + 1) & top]
        a.reshape(-1)[: len(edges)] = np.array(edges, dtype=np.uint64).astype(dt)
    return a


def all_pairs():
    """Every (source, destination) pair of the 33 colour types: 28 x 28 through SL_BlitProcessor, the 305 pairs with a
    packed type on either side through SL_BlitCompressedProcessor."""
    types = list(range(28)) + sorted(PACKED)
    for st in types:
        for dt in types:
            yield st, dt


def run_matrix(impl, mode):
    """One context, all 400 pairs; returns {(src, dst): bytes}."""
    ctx = A.Context(impl)
    w, h = 8, 4
    srcs = {}
    out = {}
    for st, dt in all_pairs():
        if st not in srcs:
            srcs[st] = ctx.create_texture(st, w, h, 1, source_texels(st, w, h, mode))
        d = ctx.create_texture(dt, w, h, 1)
        ctx.blit(d, srcs[st])
        ctx.sync()
        out[(st, dt)] = ctx.texture_read(d, dt, w, h).tobytes()
    ctx.close()
    return out


def digest(m):
    h = hashlib.sha256()
    for k in sorted(m):
        h.update(m[k])
    return h.hexdigest()


def type_name(t):
    return PACKED[t] if t in PACKED else f"{NAMES[t % 7]}x{t // 7 + 1}"


def label(st, dt):
    return f"{type_name(st)}->{type_name(dt)}"


@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_matrix_matches_committed_reference_hashes(mode):
    want = json.load(open(GOLDEN))
    assert digest(run_matrix("port", mode)) == want[f"mode{mode}"]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_matrix_equals_compiled_reference(mode):
    a, b = run_matrix("port", mode), run_matrix("ref_strict", mode)
    bad = [label(*k) for k in a if a[k] != b[k]]
    assert not bad, bad[:10]


def test_two_component_integer_quirk_known_answer():
    """SL_Color.hpp:373-380: the RG integer overload divides by max(T) itself -- an RG8 -> RG8 blit leaves 0 or 1."""
    ctx = A.Context("port")
    px = np.array([[[255, 254], [128, 0]]], dtype=np.uint8)
    s = ctx.create_texture(A.SL_COLOR_RG_8U, 2, 1, 1, px)
    d = ctx.create_texture(A.SL_COLOR_RG_8U, 2, 1, 1)
    ctx.blit(d, s)
    assert ctx.texture_read(d, A.SL_COLOR_RG_8U, 2, 1).reshape(-1).tolist() == [1, 0, 0, 0]
    # R lands in the LAST colour channel of RGB (src/SL_BlitProcessor.cpp:196-204)
    s1 = ctx.create_texture(A.SL_COLOR_R_8U, 1, 1, 1, np.array([[[200]]], dtype=np.uint8))
    d3 = ctx.create_texture(A.SL_COLOR_RGB_8U, 1, 1, 1)
    ctx.blit(d3, s1)
    assert ctx.texture_read(d3, A.SL_COLOR_RGB_8U, 1, 1).reshape(-1).tolist() == [0, 0, 200]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_cuda_matrix_equals_oracle(mode):
    a, b = run_matrix("cuda", mode), run_matrix("port", mode)
    bad = [label(*k) for k in a if a[k] != b[k]]
    assert not bad, bad[:10]


def packed_source(st):
    """Sources for the exhaustive packed run: every value of the 8/16-bit packed types, 65536 RGBA1010102 values (random plus
    a sweep that walks all four fields), and 64 x 256 plain textures with the values the x86 conversions branch on."""
    rng = np.random.default_rng(st)
    dt = A.texel_dtype(st)
    n = A.texel_channels(st)
    if A.is_packed_color(st):
        nb = A.bytes_per_texel(st)
        if nb == 1:
            return np.arange(256, dtype=np.uint8).reshape(1, 256, 1)
        if nb == 2:
            return np.arange(65536, dtype=np.uint16).reshape(256, 256, 1)
        sweep = np.arange(4096, dtype=np.uint32)
        a = np.concatenate([rng.integers(0, 2 ** 32, size=65536 - 4096, dtype=np.uint64).astype(np.uint32),
                            sweep * np.uint32(0x00100401) + ((sweep % 4) << 30).astype(np.uint32)])
        return a.reshape(256, 256, 1)
    h, w = 64, 256
    if np.dtype(dt).kind == "f":
        a = rng.uniform(-0.6, 1.6, size=(h, w, n))
        a[:8] = rng.uniform(0, 1, size=(8, w, n))
        with np.errstate(over="ignore"):
            a = a.astype(dt)
            a[8, :16, :] = np.array([0, 1, 0.5, 1e9, -1e9, np.inf, -np.inf, np.nan, 1e-8, 0.999999, 1.0000001, 2, 3, 4, 255.5, -0.0]).astype(dt)[:, None]
        return a
    a = rng.integers(0, np.iinfo(dt).max, size=(h, w, n), endpoint=True, dtype=np.uint64).astype(dt)
    top = np.iinfo(dt).max
    edges = [0, 1, top, top - 1, top // 2, top // 2 + 1, (2 ** 24 + 1) & top, (2 ** 31) & top, 255, 256 & top, 65535 & top, 65536 & top,
             (2 ** 53 + 1) & top, (2 ** 63 + 2 ** 39 + 1) & top, top // 1023, top // 7]
    a[0, :16, :] = np.array(edges, dtype=np.uint64).astype(dt)[:, None]
    return a


def run_packed(impl):
    """All 305 pairs with a packed type on either side, on packed_source(); returns {(src, dst): bytes}."""
    ctx = A.Context(impl)
    srcs = {}
    for st in range(33):
        a = packed_source(st)
        srcs[st] = (ctx.create_texture(st, a.shape[1], a.shape[0], 1, a), a.shape[1], a.shape[0])
    out = {}
    for st in range(33):
        for dt in range(33):
            if not (A.is_packed_color(st) or A.is_packed_color(dt)):
                continue
            tid, w, h = srcs[st]
            d = ctx.create_texture(dt, w, h, 1)
            ctx.blit(d, tid)
            ctx.sync()
            out[(st, dt)] = ctx.texture_read(d, dt, w, h).tobytes()
    ctx.close()
    return out


def test_oracle_packed_pairs_match_committed_reference_hash():
    assert digest(run_packed("port")) == json.load(open(GOLDEN))["packed_exhaustive"]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
def test_oracle_packed_pairs_equal_compiled_reference():
    a, b = run_packed("port"), run_packed("ref_strict")
    bad = [label(*k) for k in a if a[k] != b[k]]
    assert not bad, bad[:10]


def test_packed_known_answers():
    """RGBA1010102's AVX2 casts sign-extend their fields (SL_ColorCompressed.hpp:830-866): 1023 reads back as 1/1023 in
    float, as 255 (the top eight bits) in uint8 RGBA, and as the LOW eight bits through the integer rgb_cast."""
    ctx = A.Context("port")
    px = np.array([[[1023 | (512 << 10) | (511 << 20) | (3 << 30)]]], dtype=np.uint32)
    s = ctx.create_texture(A.SL_COLOR_RGBA_1010102, 1, 1, 1, px)
    f4 = ctx.create_texture(A.SL_COLOR_RGBA_FLOAT, 1, 1, 1)
    ctx.blit(f4, s)
    got = ctx.texture_read(f4, A.SL_COLOR_RGBA_FLOAT, 1, 1).reshape(-1)
    want = np.array([1, 512, 511, 1], dtype=np.float32) * np.array([1 / 1023, 1 / 1023, 1 / 1023, 1 / 3], dtype=np.float32)
    assert got.tolist() == want.tolist()
    u4 = ctx.create_texture(A.SL_COLOR_RGBA_8U, 1, 1, 1)
    ctx.blit(u4, s)
    assert ctx.texture_read(u4, A.SL_COLOR_RGBA_8U, 1, 1).reshape(-1).tolist() == [255, 128, 127, 255]
    u3 = ctx.create_texture(A.SL_COLOR_RGB_8U, 1, 1, 1)
    ctx.blit(u3, s)
    assert ctx.texture_read(u3, A.SL_COLOR_RGB_8U, 1, 1).reshape(-1).tolist() == [1023 & 255, 512 & 255, 511 & 255]
    # a source alpha never reaches a packed destination: RGBA8 -> RGBA4444 stores the maximum (src/SL_BlitCompressedProcessor.cpp:213-233)
    s8 = ctx.create_texture(A.SL_COLOR_RGBA_8U, 1, 1, 1, np.array([[[255, 128, 17, 0]]], dtype=np.uint8))
    p4 = ctx.create_texture(A.SL_COLOR_RGBA_4444, 1, 1, 1)
    ctx.blit(p4, s8)
    assert int(ctx.texture_read(p4, A.SL_COLOR_RGBA_4444, 1, 1).reshape(-1)[0]) == 15 | (7 << 4) | (1 << 8) | (15 << 12)


@pytest.mark.gpu
def test_cuda_packed_pairs_equal_oracle():
    a, b = run_packed("cuda"), run_packed("port")
    bad = [label(*k) for k in a if a[k] != b[k]]
    assert not bad, bad[:10]


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
def test_oracle_equals_reference_on_random_blits():
    """Random type pairs, sizes and NON-EMPTY rectangles (resampling, vertical flip, partial destinations).  An empty source
    rectangle makes the reference read row or column -1 and is left out."""
    rng = np.random.default_rng(77)
    bad = []
    for _ in range(400):
        st, dt = int(rng.integers(0, 33)), int(rng.integers(0, 33))
        sw, sh, dw, dh = [int(rng.integers(1, 90)) for _ in range(4)]
        pool = packed_source(st).reshape(-1, A.texel_channels(st))
        a = np.ascontiguousarray(pool[rng.integers(0, pool.shape[0], size=sw * sh)].reshape(sh, sw, -1))
        rect = None
        if rng.random() < 0.6:
            x0 = int(rng.integers(0, sw)); x1 = int(rng.integers(x0 + 1, sw + 1))
            y0 = int(rng.integers(0, sh)); y1 = int(rng.integers(y0 + 1, sh + 1))
            u0, u1 = sorted(int(v) for v in rng.integers(0, dw + 1, size=2))
            v0, v1 = sorted(int(v) for v in rng.integers(0, dh + 1, size=2))
            rect = (x0, y0, x1, y1, u0, v0, u1, v1)
        outs = []
        for impl in ("port", "ref_strict"):
            ctx = A.Context(impl)
            s = ctx.create_texture(st, sw, sh, 1, a)
            d = ctx.create_texture(dt, dw, dh, 1)
            ctx.blit(d, s, rect)
            ctx.sync()
            outs.append(ctx.texture_read(d, dt, dw, dh).tobytes())
            ctx.close()
        if outs[0] != outs[1]:
            bad.append((label(st, dt), (sw, sh, dw, dh), rect))
    assert not bad, bad[:5]
