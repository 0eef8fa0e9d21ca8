/*
 * sl_texel.cuh -- texel conversions between the 33 colour types (SL_ColorDataType).
 *
 *   sl_blit_texel        one texel of SL_BlitProcessor (src/SL_BlitProcessor.cpp:20-426 + color_cast, SL_Color.hpp:236-700):
 *                        all 28 x 28 pairs of the plain layouts (R/RG/RGB/RGBA x u8, u16, u32, u64, half, float, double)
 *                        and of SL_BlitCompressedProcessor (src/SL_BlitCompressedProcessor.cpp:20-258 + rgb_cast / rgba_cast,
 *                        SL_ColorCompressed.hpp): every pair with a packed type (RGB332, RGB565, RGBA5551, RGBA4444,
 *                        RGBA1010102) on either side.
 *   sl_packed_from_float / sl_float_from_packed
 *                        rgba_cast<packed, float> / rgba_cast<float, packed>: the packed types as render targets
 *                        (assign_pixel / assign_alpha_pixel specialisations, src/SL_Framebuffer.cpp:85-145, 371-590).
 *
 * Plain layouts.  A blit functor builds a colour of the DESTINATION's channel count in the SOURCE's component type (missing
 * channels are the type's minimum, alpha its maximum; R lands in the LAST colour channel of RGB/RGBA, RG in the middle of
 * RGBA), converts it with color_cast and stores it; "to R" functors convert the whole source colour and keep component 0.
 * Four-component conversions go through vec4_t's SSE cast operators where one exists (x86/vec4f_impl.h:190-355: float ->
 * u8/u16/u32 = cvtps2dq + saturating packs, float -> half = cvtps2ph toward zero, u32 -> float = the halves trick), the
 * others are element-wise C casts (cvttss2si / cvttsd2si -- the 64-bit form towards u32 and u64 -- and the low bits;
 * g++'s compare / subtract / flip sequence towards u64; scalar half conversions round to nearest).  The two-component
 * integer overload scales by max(T) instead of by the ratio of the maxima (SL_Color.hpp:373-380) -- reproduced,
 * including for equal types.
 *
 * Packed types are bit-fields with r in the lowest bits.  Which cast a blit uses is fixed by the functor:
 *   packed -> R, RGB: rgb_cast (component 0 for R);   packed -> RG, RGBA: rgba_cast (leading two for RG);
 *   R/RG/RGB/RGBA -> packed: ALWAYS rgb_cast of (r, min, min) / (r, g, min) / (r, g, b) / (r, g, b): a source alpha is
 *   dropped, the packed alpha becomes its maximum;
 *   packed -> another packed: through uint8 (uint16 towards 1010102), rgb_cast both ways towards the 3-component types,
 *   rgba_cast both ways towards the 4-component ones.
 * 1010102 has its own arithmetic (SL_ColorCompressed.hpp:745-775, 830-870): the integer rgb_cast multiplies by max(T) in
 * T's own (promoted) width and divides by (T)1023 -- for uint8 that is the LOW 8 bits, for the 32/64-bit types the product
 * wraps; the integer rgba_cast multiplies by (uint32_t)max(T) in 32 bits; the AVX2 specialisations for float and uint8
 * RGBA sign-extend the fields (abs() of it for float: 1023 reads as 1/1023; the top eight bits for uint8).
 */
#pragma once

#ifndef SL_TEXEL_HOST_CHECK /* tests/host_check/ compiles this header for the host with its own stand-ins */
#include "sl_shader_api.cuh"
#endif

namespace sld
{

struct SLTexelValue /* one component in its own type: kinds 0..3 u8/u16/u32/u64 and 4 (half bits) in u, 5 float, 6 double */
{
    unsigned long long u;
    float f;
    double d;
};

SLD_FN unsigned long long sl_int_max(int kind) { return kind == 0 ? 0xffull : (kind == 1 ? 0xffffull : (kind == 2 ? 0xffffffffull : 0xffffffffffffffffull)); }
SLD_FN float sl_int_max_f(int kind) { return kind == 0 ? 255.f : (kind == 1 ? 65535.f : (kind == 2 ? 4294967296.f : 18446744073709551616.f)); }
SLD_FN double sl_int_max_d(int kind) { return kind == 0 ? 255.0 : (kind == 1 ? 65535.0 : (kind == 2 ? 4294967295.0 : 18446744073709551616.0)); }
SLD_FN int sl_kind_bytes(int kind) { return kind == 0 ? 1 : ((kind == 1 || kind == 4) ? 2 : ((kind == 2 || kind == 5) ? 4 : 8)); }

SLD_FN SLTexelValue sl_texel_load(const unsigned char* p, int kind)
{
    SLTexelValue t{0ull, 0.f, 0.0};
    if (kind == 0) t.u = *p;
    else if (kind == 1 || kind == 4) t.u = *reinterpret_cast<const uint16_t*>(p);
    else if (kind == 2) t.u = *reinterpret_cast<const uint32_t*>(p);
    else if (kind == 3) t.u = *reinterpret_cast<const unsigned long long*>(p);
    else if (kind == 5) t.f = *reinterpret_cast<const float*>(p);
    else t.d = *reinterpret_cast<const double*>(p);
    return t;
}
SLD_FN void sl_texel_store(unsigned char* p, int kind, const SLTexelValue& t)
{
    if (kind == 0) *p = (unsigned char)t.u;
    else if (kind == 1 || kind == 4) *reinterpret_cast<uint16_t*>(p) = (uint16_t)t.u;
    else if (kind == 2) *reinterpret_cast<uint32_t*>(p) = (uint32_t)t.u;
    else if (kind == 3) *reinterpret_cast<unsigned long long*>(p) = t.u;
    else if (kind == 5) *reinterpret_cast<float*>(p) = t.f;
    else *reinterpret_cast<double*>(p) = t.d;
}
SLD_FN SLTexelValue sl_texel_limit(int kind, bool maxValue) /* SL_ColorLimits (SL_Color.hpp:90-175): min() is 0 for every type */
{
    SLTexelValue t{0ull, 0.f, 0.0};
    if (!maxValue) return t;
    if (kind <= 3) t.u = sl_int_max(kind);
    else if (kind == 4) t.u = 0x3c00ull;
    else if (kind == 5) t.f = 1.f;
    else t.d = 1.0;
    return t;
}

/* cvttsd2si, 32- and 64-bit: "integer indefinite" on NaN / overflow */
SLD_FN int32_t sl_cvt_tr_d(double x) { return (x > -2147483649.0 && x < 2147483648.0) ? __double2int_rz(x) : (int32_t)0x80000000; }
SLD_FN long long sl_cvt_tr64_d(double x)
{
    return (x >= -9223372036854775808.0 && x < 9223372036854775808.0) ? __double2ll_rz(x) : (long long)0x8000000000000000ull;
}
/* (T)x for an unsigned integer T the way g++ compiles it on x86-64: cvttss2si / cvttsd2si and the low bits for u8 / u16,
 * the 64-bit conversion and the low 32 bits for u32, and for u64: below 2^63 (NaN and negatives included) the 64-bit
 * conversion, otherwise the conversion of x - 2^63 with the top bit flipped */
SLD_FN unsigned long long sl_trunc_f(float x, int kind)
{
    if (kind <= 1) return (unsigned long long)(uint32_t)sl_cvt_tr(x) & sl_int_max(kind);
    if (kind == 2) return (unsigned long long)sl_cvt_tr64(x) & 0xffffffffull;
    if (x >= 9223372036854775808.f) return (unsigned long long)sl_cvt_tr64(__fsub_rn(x, 9223372036854775808.f)) ^ 0x8000000000000000ull;
    return (unsigned long long)sl_cvt_tr64(x);
}
SLD_FN unsigned long long sl_trunc_d(double x, int kind)
{
    if (kind <= 1) return (unsigned long long)(uint32_t)sl_cvt_tr_d(x) & sl_int_max(kind);
    if (kind == 2) return (unsigned long long)sl_cvt_tr64_d(x) & 0xffffffffull;
    if (x >= 9223372036854775808.0) return (unsigned long long)sl_cvt_tr64_d(__dsub_rn(x, 9223372036854775808.0)) ^ 0x8000000000000000ull;
    return (unsigned long long)sl_cvt_tr64_d(x);
}
SLD_FN float sl_int_to_float(unsigned long long v, int kind, bool simd4)
{
    if (kind == 2 && simd4) return sl_u32_to_float((uint32_t)v); /* vec4_t<uint32_t> -> vec4_t<float> (vec4f_impl.h:236-245) */
    if (kind == 3) return __ull2float_rn(v);                      /* g++'s uint64 -> float sequence is correctly rounded */
    return __uint2float_rn((uint32_t)v);
}
SLD_FN unsigned long long sl_float_to_int(float x, int kind, bool simd4)
{
    if (simd4 && kind <= 2)
    {
        const int32_t r = sl_cvt_rn(x); /* cvtps2dq */
        if (kind == 0) { const int32_t s = r < -32768 ? -32768 : (r > 32767 ? 32767 : r); return (unsigned long long)(s < 0 ? 0 : (s > 255 ? 255 : s)); } /* packs_epi32, packus_epi16 */
        if (kind == 1) return (unsigned long long)(r < 0 ? 0 : (r > 65535 ? 65535 : r));                                                                /* packus_epi32 */
        return (unsigned long long)(uint32_t)r;
    }
    return sl_trunc_f(x, kind);
}

/* color_cast<T, U> of one component; `width` = the channel count the cast runs on */
SLD_FN SLTexelValue sl_texel_convert(const SLTexelValue& in, int ik, int ok, int width)
{
    const bool simd4 = width == 4;
    SLTexelValue o{0ull, 0.f, 0.0};
    if (ik <= 3 && ok <= 3) /* integer -> integer (SL_Color.hpp:236-246) */
    {
        const unsigned long long tMax = sl_int_max(ok), uMax = sl_int_max(ik);
        if (width == 2) /* scales by max(T) itself (:373-380): widening multiplies and wraps, narrowing (and same-type!) divides */
            o.u = tMax > uMax ? (((in.u & tMax) * tMax) & tMax) : ((in.u / tMax) & tMax);
        else
            o.u = tMax > uMax ? (((in.u & tMax) * (tMax / (uMax & tMax))) & tMax) : ((in.u / (uMax / (tMax & uMax))) & tMax);
        return o;
    }
    if (ok <= 3) /* float kinds -> integer (:250-266): p * (U)max(T), then the cast */
    {
        if (ik == 6) { o.u = sl_trunc_d(__dmul_rn(in.d, sl_int_max_d(ok)), ok); return o; }
        const float x = __fmul_rn(ik == 4 ? sl_h2f((uint16_t)in.u) : in.f, sl_int_max_f(ok));
        o.u = sl_float_to_int(x, ok, simd4);
        return o;
    }
    if (ik <= 3) /* integer -> float kinds (:270-300) */
    {
        if (ok == 6)
        {
            const double scale = ik == 0 ? (1.0 / 255.0) : (ik == 1 ? (1.0 / 65535.0) : (ik == 2 ? (1.0 / 4294967295.0) : (1.0 / 18446744073709551616.0)));
            o.d = __dmul_rn(__ull2double_rn(in.u), scale);
            return o;
        }
        const float scale = ik == 0 ? (ok == 5 ? 0.00392156862745f : (1.f / 255.f))
                                    : (ik == 1 ? (1.f / 65535.f) : (ik == 2 ? (1.f / 4294967296.f) : (1.f / 18446744073709551616.f)));
        const float f = __fmul_rn(sl_int_to_float(in.u, ik, simd4), scale);
        if (ok == 5) o.f = f;
        else o.u = simd4 ? sl_f2h_rz(f) : sl_f2h_rn(f);
        return o;
    }
    /* float kinds -> float kinds (:304-310): plain casts; half only talks to float */
    if (ik == ok) return in;
    if (ik == 4)
    {
        const float f = sl_h2f((uint16_t)in.u);
        if (ok == 5) o.f = f; else o.d = (double)f;
        return o;
    }
    if (ik == 5)
    {
        if (ok == 6) o.d = (double)in.f;
        else o.u = simd4 ? sl_f2h_rz(in.f) : sl_f2h_rn(in.f);
        return o;
    }
    if (ok == 5) o.f = __double2float_rn(in.d);
    else o.u = sl_f2h_rn(__double2float_rn(in.d)); /* vec4_t<double> has no SSE cast: (half)(float)(double) */
    return o;
}
/* which source component (or limit) feeds destination component j: 0..3 = source component, -1 = min, -2 = max */
SLD_FN int sl_blit_channel_source(int srcN, int dstN, int j)
{
    if (dstN <= srcN) return j;
    if (srcN == 1) return dstN == 2 ? (j == 0 ? 0 : -1) : (j == 2 ? 0 : (j == 3 ? -2 : -1));
    if (srcN == 2) return dstN == 3 ? (j < 2 ? j : -1) : (j == 0 ? -1 : (j == 3 ? -2 : j - 1));
    return j < 3 ? j : -2;
}

/* ---- packed types: type 28 + ci, ci = 0 RGB332, 1 RGB565, 2 RGBA5551, 3 RGBA4444, 4 RGBA1010102 ---- */
SLD_FN bool sl_is_packed(int type) { return type >= 28; }
SLD_FN int sl_pack_bits(int ci, int j)
{
    /* {3,3,2,0}, {5,6,5,0}, {5,5,5,1}, {4,4,4,4}, {10,10,10,2} */
    const uint32_t row = ci == 0 ? 0x0233u : (ci == 1 ? 0x0565u : (ci == 2 ? 0x1555u : (ci == 3 ? 0x4444u : 0x2aaau)));
    return (int)((row >> (4 * j)) & 0xfu);
}
SLD_FN int sl_pack_bytes(int ci) { return ci == 0 ? 1 : (ci == 4 ? 4 : 2); }
SLD_FN uint32_t sl_pack_max(int ci, int j) { return (1u << sl_pack_bits(ci, j)) - 1u; }
SLD_FN void sl_unpack(int ci, uint32_t v, uint32_t c[4])
{
    int sh = 0;
    for (int j = 0; j < 4; ++j) { c[j] = (v >> sh) & sl_pack_max(ci, j); sh += sl_pack_bits(ci, j); }
}
SLD_FN uint32_t sl_pack(int ci, const uint32_t c[4])
{
    uint32_t v = 0;
    int sh = 0;
    for (int j = 0; j < 4; ++j) { v |= (c[j] & sl_pack_max(ci, j)) << sh; sh += sl_pack_bits(ci, j); }
    return v;
}
SLD_FN uint32_t sl_packed_load(const unsigned char* p, int ci)
{
    return ci == 0 ? (uint32_t)*p : (ci == 4 ? *reinterpret_cast<const uint32_t*>(p) : (uint32_t)*reinterpret_cast<const uint16_t*>(p));
}
SLD_FN void sl_packed_store(unsigned char* p, int ci, uint32_t v)
{
    if (ci == 0) *p = (unsigned char)v;
    else if (ci == 4) *reinterpret_cast<uint32_t*>(p) = v;
    else *reinterpret_cast<uint16_t*>(p) = (uint16_t)v;
}

/* rgb_cast / rgba_cast<packed, U>: integers c / (max(U) / (U)max(field)), float kinds (value_type)(c * (U)max(field)) */
SLD_FN uint32_t sl_packed_from_plain(const SLTexelValue* c, int ik, int ci, bool withAlpha)
{
    uint32_t out[4] = {0u, 0u, 0u, sl_pack_max(ci, 3)};
    const int elemKind = ci == 4 ? 1 : 0; /* the casts go to the packed colour's value_type: uint8_t, uint16_t for 1010102 */
    for (int j = 0; j < (withAlpha ? 4 : 3); ++j)
    {
        const uint32_t cm = sl_pack_max(ci, j);
        if (!cm) continue;
        if (ik <= 3)
        {
            const unsigned long long uMax = sl_int_max(ik);
            out[j] = (uint32_t)((c[j].u / (uMax / ((unsigned long long)cm & uMax))) & sl_int_max(elemKind));
        }
        else if (ik == 6) out[j] = (uint32_t)sl_trunc_d(__dmul_rn(c[j].d, (double)cm), elemKind);
        else if (ik == 5) out[j] = (uint32_t)sl_trunc_f(__fmul_rn(c[j].f, (float)cm), elemKind);
        else
        {
            const uint16_t h = sl_f2h_rn(__fmul_rn(sl_h2f((uint16_t)c[j].u), sl_h2f(sl_f2h_rn((float)cm)))); /* half * half rounds to half */
            out[j] = (uint32_t)sl_trunc_f(sl_h2f(h), elemKind);
        }
    }
    return sl_pack(ci, out);
}
/* rgb_cast (rgba == false) / rgba_cast (rgba == true) <T, packed>; out[3] is only written by rgba_cast */
SLD_FN void sl_plain_from_packed(uint32_t v, int ci, int ok, bool rgba, SLTexelValue out[4])
{
    uint32_t c[4];
    sl_unpack(ci, v, c);
    for (int j = 0; j < (rgba ? 4 : 3); ++j)
    {
        const uint32_t cm = sl_pack_max(ci, j);
        SLTexelValue o{0ull, 0.f, 0.0};
        if (!cm) { out[j] = sl_texel_limit(ok, true); continue; } /* 332 / 565 have no alpha: max(T) */
        if (ok <= 3)
        {
            const unsigned long long tMax = sl_int_max(ok);
            if (ci != 4) o.u = ((unsigned long long)c[j] * (tMax / ((unsigned long long)cm & tMax))) & tMax;
            else if (!rgba)
            {
                unsigned long long prod = (unsigned long long)c[j] * tMax; /* int for uint8/uint16 (no overflow), else T's own width */
                if (ok == 2) prod &= 0xffffffffull;
                o.u = (prod / (1023ull & tMax)) & tMax;
            }
            else if (ok == 0) o.u = j < 3 ? (unsigned long long)((c[j] >> 2) & 0xffu) : (unsigned long long)(c[j] * 85u); /* AVX2 specialisation (:846-866) */
            else o.u = (unsigned long long)((uint32_t)(c[j] * (uint32_t)tMax) / cm) & tMax;
        }
        else if (ci == 4 && rgba && ok == 5) /* AVX2 specialisation (:830-843): sign-extended field, abs, * (1 / max) */
        {
            const int bits = sl_pack_bits(ci, j);
            const int32_t ext = (int32_t)(c[j] << (32 - bits)) >> (32 - bits);
            o.f = __fmul_rn(__int2float_rn(ext < 0 ? -ext : ext), __fdiv_rn(1.f, (float)cm));
        }
        else if (ok == 5) o.f = __fdiv_rn((float)c[j], (float)cm);
        else if (ok == 6) o.d = __ddiv_rn((double)c[j], (double)cm);
        else o.u = sl_f2h_rn(__fdiv_rn((float)c[j], (float)cm)); /* half / half: float division, rounded to half */
        out[j] = o;
    }
}
SLD_FN uint32_t sl_packed_from_packed(uint32_t v, int si, int di)
{
    if (si == di) return v;
    const bool viaRgba = di >= 2;
    const int midKind = di == 4 ? 1 : 0;
    SLTexelValue mid[4];
    sl_plain_from_packed(v, si, midKind, viaRgba, mid);
    return sl_packed_from_plain(mid, midKind, di, viaRgba);
}
/* rgba_cast<packed, float>: generic for the 8-bit-valued types, the AVX2 specialisation (cvtps2dq, shift, mask) for 1010102 */
SLD_FN uint32_t sl_packed_from_float(int ci, const math::vec4& c)
{
    if (ci == 4)
    {
        uint32_t v = 0;
        int sh = 0;
        for (int j = 0; j < 4; ++j)
        {
            const int bits = sl_pack_bits(4, j);
            v |= ((uint32_t)sl_cvt_rn(__fmul_rn(c.v[j], (float)sl_pack_max(4, j))) << sh) & (sl_pack_max(4, j) << sh);
            sh += bits;
        }
        return v;
    }
    SLTexelValue t[4];
    for (int j = 0; j < 4; ++j) t[j] = SLTexelValue{0ull, c.v[j], 0.0};
    return sl_packed_from_plain(t, 5, ci, true);
}
SLD_FN math::vec4 sl_float_from_packed(int ci, uint32_t v)
{
    SLTexelValue t[4];
    sl_plain_from_packed(v, ci, 5, true, t);
    return math::vec4{t[0].f, t[1].f, t[2].f, t[3].f};
}

SLD_FN int sl_texel_bytes(int type)
{
    return sl_is_packed(type) ? sl_pack_bytes(type - 28) : (type / 7 + 1) * sl_kind_bytes(type % 7);
}

/* one texel of a converting blit, any (source type, destination type) pair */
SLD_FN void sl_blit_texel(const unsigned char* src, int srcType, unsigned char* dst, int dstType)
{
    if (sl_is_packed(srcType) && sl_is_packed(dstType))
    {
        sl_packed_store(dst, dstType - 28, sl_packed_from_packed(sl_packed_load(src, srcType - 28), srcType - 28, dstType - 28));
        return;
    }
    if (sl_is_packed(dstType))
    {
        const int ik = srcType % 7, srcN = srcType / 7 + 1, ib = sl_kind_bytes(ik);
        SLTexelValue c[3];
        for (int j = 0; j < 3; ++j) c[j] = j < srcN ? sl_texel_load(src + j * ib, ik) : sl_texel_limit(ik, false);
        sl_packed_store(dst, dstType - 28, sl_packed_from_plain(c, ik, dstType - 28, false));
        return;
    }
    const int ok = dstType % 7, dstN = dstType / 7 + 1, ob = sl_kind_bytes(ok);
    if (sl_is_packed(srcType))
    {
        SLTexelValue c[4];
        sl_plain_from_packed(sl_packed_load(src, srcType - 28), srcType - 28, ok, dstN == 2 || dstN == 4, c);
        for (int j = 0; j < dstN; ++j) sl_texel_store(dst + j * ob, ok, c[j]);
        return;
    }
    const int ik = srcType % 7, srcN = srcType / 7 + 1, ib = sl_kind_bytes(ik);
    const int width = dstN == 1 ? srcN : dstN; /* the channel count color_cast runs on */
    for (int j = 0; j < dstN; ++j)
    {
        const int from = sl_blit_channel_source(srcN, dstN, j);
        const SLTexelValue in = from >= 0 ? sl_texel_load(src + from * ib, ik) : sl_texel_limit(ik, from == -2);
        SLTexelValue out = sl_texel_convert(in, ik, ok, width);
        /* the constant alpha of a double colour, 1.0 * (double)UINT64_MAX = 2^64, is converted by the compiler, which
         * saturates; the same product at run time (a float colour's alpha) goes through the cvttss2si sequence and gives 0 */
        if (from == -2 && ik == 6 && ok == 3) out.u = 0xffffffffffffffffull;
        sl_texel_store(dst + j * ob, ok, out);
    }
}

} // namespace sld
