"""Blit parity cases (SL_Context::blit -> SL_BlitProcessor::blit_nearest), replayed through any of the three libraries."""
import numpy as np

from softlight_b200 import capi as A

# name: (src type, src w, h, dst type, dst w, h, rect or None)   rect = (sx0, sy0, sx1, sy1, dx0, dy0, dx1, dy1)
BLIT_CASES = {
    "rgba8_same_size": (A.SL_COLOR_RGBA_8U, 64, 48, A.SL_COLOR_RGBA_8U, 64, 48, None),
    "rgba8_upscale": (A.SL_COLOR_RGBA_8U, 64, 48, A.SL_COLOR_RGBA_8U, 200, 150, None),
    "rgba8_downscale_odd": (A.SL_COLOR_RGBA_8U, 257, 131, A.SL_COLOR_RGBA_8U, 31, 17, None),
    "rgba8_rect": (A.SL_COLOR_RGBA_8U, 128, 96, A.SL_COLOR_RGBA_8U, 160, 120, (16, 8, 100, 70, 20, 30, 140, 100)),
    "rgbaf_to_rgba8": (A.SL_COLOR_RGBA_FLOAT, 96, 64, A.SL_COLOR_RGBA_8U, 96, 64, None),
    "rgbaf_to_rgba8_scaled": (A.SL_COLOR_RGBA_FLOAT, 96, 64, A.SL_COLOR_RGBA_8U, 333, 77, None),
    "rgba8_to_rgbaf": (A.SL_COLOR_RGBA_8U, 80, 60, A.SL_COLOR_RGBA_FLOAT, 120, 90, None),
    "rfloat_same": (A.SL_COLOR_R_FLOAT, 50, 40, A.SL_COLOR_R_FLOAT, 75, 60, None),
    "rgb8_same": (A.SL_COLOR_RGB_8U, 33, 21, A.SL_COLOR_RGB_8U, 66, 42, None),
    "rgbaf_same_4k_to_1080": (A.SL_COLOR_RGBA_FLOAT, 384, 216, A.SL_COLOR_RGBA_FLOAT, 192, 108, None),
    # resampling + rectangles through the general conversion: the 32/64-bit integer layouts and the packed types
    # (SL_BlitCompressedProcessor::blit_nearest has the same loop as SL_BlitProcessor's)
    "rgba8_to_rgb565_upscale": (A.SL_COLOR_RGBA_8U, 64, 48, A.SL_COLOR_RGB_565, 200, 150, None),
    "rgb565_to_rgba8_rect": (A.SL_COLOR_RGB_565, 128, 96, A.SL_COLOR_RGBA_8U, 160, 120, (16, 8, 100, 70, 20, 30, 140, 100)),
    "rgbaf_to_rgba1010102_downscale": (A.SL_COLOR_RGBA_FLOAT, 257, 131, A.SL_COLOR_RGBA_1010102, 31, 17, None),
    "rgba1010102_to_rgba4444_rect": (A.SL_COLOR_RGBA_1010102, 96, 64, A.SL_COLOR_RGBA_4444, 120, 90, (3, 5, 90, 60, 7, 11, 111, 80)),
    "rgb332_to_rgbf_scaled": (A.SL_COLOR_RGB_332, 50, 40, A.SL_COLOR_RGB_FLOAT, 75, 60, None),
    "rg16_to_rgba64_rect": (A.SL_COLOR_RG_16U, 80, 60, A.SL_COLOR_RGBA_64U, 100, 70, (10, 10, 70, 50, 5, 5, 95, 65)),
    "rgbad_to_rgb32_scaled": (A.SL_COLOR_RGBA_DOUBLE, 40, 30, A.SL_COLOR_RGB_32U, 90, 70, None),
}


def source_texels(name, ctype, w, h):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    n = A.texel_channels(ctype)
    dt = A.texel_dtype(ctype)
    if np.dtype(dt).kind == "f":
        a = rng.uniform(-0.25, 1.25, size=(h, w, n)).astype(dt)  # out-of-range values exercise the saturating casts
        a[0, 0, :] = np.nan if n == 4 else 0.5
        a[1 % h, 1 % w, :] = 1e20
        a[2 % h, 2 % w, :] = -1e20
        return a
    if np.dtype(dt).itemsize == 8:
        return rng.integers(0, np.iinfo(dt).max, size=(h, w, n), endpoint=True, dtype=np.uint64)
    return rng.integers(0, np.iinfo(dt).max, size=(h, w, n), endpoint=True).astype(dt)


def run_blit(impl, name, device=0):
    st, sw, sh, dt, dw, dh, rect = BLIT_CASES[name]
    ctx = A.Context(impl, device=device)
    src = ctx.create_texture(st, sw, sh, 1, source_texels(name, st, sw, sh))
    dst = ctx.create_texture(dt, dw, dh, 1)
    ctx.blit(dst, src, rect)
    ctx.sync()
    out = ctx.texture_read(dst, dt, dw, dh)
    ctx.close()
    return out
