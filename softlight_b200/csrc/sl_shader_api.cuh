/*
 * sl_shader_api.cuh -- the device mirrors of the objects a SoftLight shader
 * sees, with the reference's member names so shader bodies port textually:
 * SL_UniformBuffer::as<T>(), SL_VertexBuffer::element<T>(), SL_VertexArray::
 * offset(), SL_Texture (+ ORDERED / SWIZZLED index maps, bit-identical to
 * SL_Texture.hpp:189-231 / SL_Swizzle.hpp:43-88), SL_VertexParam,
 * SL_FragmentParam (SL_Shader.hpp:41-97), color_cast (SL_Color.hpp:480-690) and
 * the samplers (SL_Sampler.hpp:51-372).
 */
#pragma once

#include <stddef.h>
#include <string.h>

#include "sl_math.cuh"

namespace sld
{

/* SL_UniformBuffer.hpp:23: a 1 KiB POD blob.  On the device the blob of the
 * draw call lives in kernel-parameter (constant) space. */
struct alignas(16) SL_UniformBuffer
{
    unsigned char mBytes[1024];
    template <typename T> SLD_FN const T* as() const { return reinterpret_cast<const T*>(mBytes); }
};

/* SL_VertexBuffer.hpp:71-88 */
struct SL_VertexBuffer
{
    const unsigned char* mBuffer;
    template <typename T> SLD_FN const T* element(ptrdiff_t offset) const { return reinterpret_cast<const T*>(mBuffer + offset); }
};

/* SL_VertexArray.hpp:18-37,107-126 */
struct SL_VertexArray
{
    struct Binding { int dimens; int type; ptrdiff_t offset; ptrdiff_t stride; };
    Binding mBindings[8];
    SLD_FN ptrdiff_t offset(size_t bindId, size_t vertId) const { return mBindings[bindId].offset + mBindings[bindId].stride * (ptrdiff_t)vertId; }
};

enum class SL_TexelOrder { ORDERED, SWIZZLED };

/* SL_TextureView (SL_Texture.hpp:29-44) + the index maps.  Storage is padded
 * and zero-filled like _sl_allocate_texture (src/SL_Texture.cpp:28-54) so the
 * one-past-the-edge taps the samplers make read the same zeros. */
struct SL_Texture
{
    uint16_t mWidth, mHeight, mDepth;
    uint8_t mBpt, mChannels;
    int mType;
    unsigned char* mTexels;

    SLD_FN uint16_t width() const { return mWidth; }
    SLD_FN uint16_t height() const { return mHeight; }
    SLD_FN uint16_t depth() const { return mDepth; }
    SLD_FN uint32_t channels() const { return (uint32_t)mChannels; } /* SL_Texture.hpp:349-352 */
    SLD_FN const void* data() const { return mTexels; }
    /* raw texel by linear index (SL_Texture.hpp:451-455) */
    template <typename color_type> SLD_FN const color_type texel(ptrdiff_t index) const { return reinterpret_cast<const color_type*>(mTexels)[index]; }

    template <SL_TexelOrder order> SLD_FN ptrdiff_t map_coordinate(uint32_t x, uint32_t y) const
    {
        if (order == SL_TexelOrder::ORDERED) return (ptrdiff_t)(x + (uint32_t)mWidth * y);
        const uint32_t tileId = (x >> 2) + ((uint32_t)mWidth >> 2) * (y >> 2);
        const uint32_t innerId = (x & 3u) + ((y & 3u) << 2);
        return (ptrdiff_t)(innerId + tileId * 16u);
    }
    template <SL_TexelOrder order> SLD_FN ptrdiff_t map_coordinate(uint32_t x, uint32_t y, uint32_t z) const
    {
        if (order == SL_TexelOrder::ORDERED) return (ptrdiff_t)(x + mWidth * (y + mHeight * z));
        const uint32_t tileId = (x >> 2) + (((uint32_t)mWidth >> 2) * ((y >> 2) + (((uint32_t)mHeight >> 2) * (z >> 2))));
        const uint32_t innerId = (x & 3u) + (((y & 3u) << 2) + (4u * ((z & 3u) << 2)));
        return (ptrdiff_t)(innerId + tileId * 64u);
    }
    template <typename color_type, SL_TexelOrder order = SL_TexelOrder::ORDERED>
    SLD_FN const color_type texel(uint16_t x, uint16_t y) const { return reinterpret_cast<const color_type*>(mTexels)[map_coordinate<order>(x, y)]; }
    template <typename color_type, SL_TexelOrder order = SL_TexelOrder::ORDERED>
    SLD_FN const color_type texel(uint16_t x, uint16_t y, uint16_t z) const { return reinterpret_cast<const color_type*>(mTexels)[map_coordinate<order>(x, y, z)]; }
};

struct SL_FragCoordXYZ { uint16_t x, y; float depth; };

struct SL_VertexParam
{
    const SL_UniformBuffer* pUniforms;
    size_t vertId;
    size_t instanceId;
    const SL_VertexArray* pVao;
    const SL_VertexBuffer* pVbo;
    math::vec4* pVaryings;
};

struct SL_FragmentParam
{
    SL_FragCoordXYZ coord;
    const SL_UniformBuffer* pUniforms;
    math::vec4 pVaryings[4];
    math::vec4 pOutputs[4];
};


/* constants.h:13-21 (double literals, like the reference: expressions that mix them with floats evaluate in double) */
#ifndef LS_PI
#define LS_PI 3.14159265358979323846
#define LS_PI_INVERSE 0.31830988618379067153
#define LS_EPSILON 1.0e-9
#endif

/* SL_PackedVertex.hpp:20-100,185-188: three sign-extended 10-bit fields * (1/511); the w field is always 0 */
SLD_FN math::vec4 sl_unpack_vec4_10_10_10_2I(int32_t norm)
{
    const int32_t v = norm & 0x3FFFFFFF;
    const int32_t x = (int32_t)((uint32_t)v << 22) >> 22;
    const int32_t y = (int32_t)((uint32_t)v << 12) >> 22;
    const int32_t z = (int32_t)((uint32_t)v << 2) >> 22;
    return math::vec4{(float)x * (1.f / 511.f), (float)y * (1.f / 511.f), (float)z * (1.f / 511.f), 0.f * (1.f / 511.f)};
}

/* SL_ColorRType (SL_Color.hpp:184-241): one channel with element-wise operators */
template <typename T> struct alignas(sizeof(T)) SL_ColorRType
{
    typedef T value_type;
    T r;
    SL_ColorRType() = default;
    SLD_FN explicit constexpr SL_ColorRType(T n) : r{n} {}
    SLD_FN constexpr SL_ColorRType operator+(const SL_ColorRType& n) const { return SL_ColorRType{(T)(r + n.r)}; }
    SLD_FN constexpr SL_ColorRType operator-(const SL_ColorRType& n) const { return SL_ColorRType{(T)(r - n.r)}; }
    SLD_FN constexpr SL_ColorRType operator*(const SL_ColorRType& n) const { return SL_ColorRType{(T)(r * n.r)}; }
    SLD_FN constexpr SL_ColorRType operator*(const T n) const { return SL_ColorRType{(T)(r * n)}; }
};
typedef SL_ColorRType<uint8_t> SL_ColorR8;
typedef SL_ColorRType<float> SL_ColorRf;
typedef math::vec3_t<float> SL_ColorRGBf; /* SL_Color.hpp:466,477 */

/* ------------------------------------------------------------ color_cast */
template <typename T, typename U> struct ColorCast;
template <> struct ColorCast<float, uint8_t> /* SL_Color.hpp:549,666: (float)p * 0.00392156862745f */
{
    static SLD_FN math::vec4 run(const math::vec4_t<uint8_t>& p) { return math::vec4{(float)p.v[0], (float)p.v[1], (float)p.v[2], (float)p.v[3]} * 0.00392156862745f; }
    static SLD_FN math::vec3 run(const math::vec3_t<uint8_t>& p) { return math::vec3{(float)p.v[0], (float)p.v[1], (float)p.v[2]} * 0.00392156862745f; }
    static SLD_FN float run(uint8_t p) { return (float)p * 0.00392156862745f; }
    static SLD_FN SL_ColorRType<float> run(const SL_ColorRType<uint8_t>& p) { return SL_ColorRType<float>{(float)p.r * 0.00392156862745f}; } /* :314-318 */
};
template <> struct ColorCast<float, uint16_t>
{
    static SLD_FN math::vec4 run(const math::vec4_t<uint16_t>& p) { return math::vec4{(float)p.v[0], (float)p.v[1], (float)p.v[2], (float)p.v[3]} * (1.f / 65535.f); }
};
template <> struct ColorCast<float, float>
{
    static SLD_FN math::vec4 run(const math::vec4& p) { return p; }
    static SLD_FN math::vec3 run(const math::vec3& p) { return p; }
    static SLD_FN float run(float p) { return p; }
};
template <> struct ColorCast<uint8_t, float> /* :506,:623: (vecN<u8>)(p*255): vec4 = cvtps2dq+packs, vec3/scalar = C truncation */
{
    static SLD_FN math::vec4_t<uint8_t> run(const math::vec4& p) { return math::vec4_to_u8(p * 255.f); }
    static SLD_FN math::vec3_t<uint8_t> run(const math::vec3& p)
    {
        const math::vec3 s = p * 255.f;
        return math::vec3_t<uint8_t>{(uint8_t)sl_cvt_tr(s.v[0]), (uint8_t)sl_cvt_tr(s.v[1]), (uint8_t)sl_cvt_tr(s.v[2])};
    }
    static SLD_FN uint8_t run(float p) { return (uint8_t)sl_cvt_tr(p * 255.f); }
    /* :270-275: (SL_ColorRType<T>)(p.r * max) = explicit ctor from float = C truncation */
    static SLD_FN SL_ColorRType<uint8_t> run(const SL_ColorRType<float>& p) { return SL_ColorRType<uint8_t>{(uint8_t)sl_cvt_tr(p.r * 255.f)}; }
};
template <typename T, typename U, typename V> SLD_FN auto color_cast(const V& p) -> decltype(ColorCast<T, U>::run(p)) { return ColorCast<T, U>::run(p); }

/* -------------------------------------------------------------- samplers */
struct SL_WrapModeClampEdge
{
    SLD_FN float operator()(float uvw) const { return math::clamp(uvw, 0.f, 1.f); }
    static constexpr bool border = false;
};
struct SL_WrapModeClampBorder
{
    SLD_FN float operator()(float uvw) const { return (uvw >= 0.f && uvw < 1.f) ? uvw : -1.f; }
    static constexpr bool border = true;
};
struct SL_WrapModeRepeat
{
    SLD_FN float operator()(float uvw) const { return (uvw < 0.f ? 1.f : 0.f) + math::fmod_1(uvw); }
    static constexpr bool border = false;
};
namespace SL_WrapMode
{
typedef SL_WrapModeClampEdge EDGE;
typedef SL_WrapModeClampBorder BORDER;
typedef SL_WrapModeRepeat REPEAT;
}

template <typename C> SLD_FN C sl_zero_color() { C c; memset(&c, 0, sizeof(c)); return c; }
SLD_FN uint16_t sl_f2u16(float f) { return (uint16_t)sl_cvt_tr(f); } /* (uint16_t)float: cvttss2si, low 16 bits */
SLD_FN uint32_t sl_f2u32(float f)                                    /* (uint_fast32_t)float: 64-bit cvttss2si, low 32 bits */
{
    if (!(f >= -9223372036854775808.0f && f < 9223372036854775808.0f)) return 0u;
    return (uint32_t)(long long)f;
}

/* SL_Sampler.hpp:122-143 */
template <typename color_type, class WrapMode, SL_TexelOrder order = SL_TexelOrder::ORDERED>
SLD_FN color_type sl_sample_nearest(const SL_Texture& tex, float x, float y)
{
    if (WrapMode::border && (x < 0.f || x >= 1.f || y < 0.f || y >= 1.f)) return sl_zero_color<color_type>();
    const WrapMode wrapMode{};
    const float wx = wrapMode(x);
    const float wy = wrapMode(y);
    const uint32_t xi = sl_f2u32((float)tex.width() * wx);
    const uint32_t yi = sl_f2u32((float)tex.height() * wy);
    return reinterpret_cast<const color_type*>(tex.data())[tex.map_coordinate<order>(xi, yi)];
}

/* SL_Sampler.hpp:145-179 */
template <typename color_type, class WrapMode, SL_TexelOrder order = SL_TexelOrder::ORDERED>
SLD_FN color_type sl_sample_nearest(const SL_Texture& tex, float x, float y, float z)
{
    if (WrapMode::border && (x < 0.f || x >= 1.f || y < 0.f || y >= 1.f || z < 0.f || z >= 1.f)) return sl_zero_color<color_type>();
    const WrapMode wrapMode{};
    const float wx = wrapMode(x);
    const float wy = wrapMode(y);
    const float wz = wrapMode(z);
    const uint32_t xi = sl_f2u32((float)tex.width() * wx);
    const uint32_t yi = sl_f2u32((float)tex.height() * wy);
    const uint32_t zi = sl_f2u32(math::round((float)tex.depth() * wz));
    return reinterpret_cast<const color_type*>(tex.data())[tex.map_coordinate<order>(xi, yi, zi)];
}

/* SL_Sampler.hpp:186-218 */
template <typename color_type, class WrapMode, SL_TexelOrder order = SL_TexelOrder::ORDERED>
SLD_FN color_type sl_sample_bilinear(const SL_Texture& tex, float x, float y)
{
    typedef typename color_type::value_type T;
    if (WrapMode::border && (x < 0.f || x >= 1.f || y < 0.f || y >= 1.f)) return sl_zero_color<color_type>();
    const WrapMode wrapMode{};
    const float xf = wrapMode(x) * (float)tex.width();
    const float yf = wrapMode(y) * (float)tex.height();
    const uint16_t xi0 = sl_f2u16(xf);
    const uint16_t yi0 = sl_f2u16(yf);
    const uint16_t xi1 = (uint16_t)math::clamp<unsigned>(xi0 + 1u, 0u, tex.width());
    const uint16_t yi1 = (uint16_t)math::clamp<unsigned>(yi0 + 1u, 0u, tex.height());
    const float dx = xf - (float)xi0;
    const float dy = yf - (float)yi0;
    const float omdx = 1.f - dx;
    const float omdy = 1.f - dy;
    const auto pixel0 = color_cast<float, T>(tex.texel<color_type, order>(xi0, yi0));
    const auto pixel1 = color_cast<float, T>(tex.texel<color_type, order>(xi0, yi1));
    const auto pixel2 = color_cast<float, T>(tex.texel<color_type, order>(xi1, yi0));
    const auto pixel3 = color_cast<float, T>(tex.texel<color_type, order>(xi1, yi1));
    const auto weight0 = pixel0 * omdx * omdy;
    const auto weight1 = pixel1 * omdx * dy;
    const auto weight2 = pixel2 * dx * omdy;
    const auto weight3 = pixel3 * dx * dy;
    const auto ret = math::sum(weight0, weight1, weight2, weight3);
    return color_cast<T, float>(ret);
}

/* SL_Sampler.hpp:315-372: "trilinear" = 8 taps at (x-1|x, y-1|y, z-1|z) weighted by the fractions of x, y, z; the
 * implicit float -> uint16_t conversions of si/ti/ri are cvttss2si */
template <typename color_type, class WrapMode, SL_TexelOrder order = SL_TexelOrder::ORDERED>
SLD_FN color_type sl_sample_trilinear(const SL_Texture& tex, float x, float y, float z)
{
    typedef typename color_type::value_type T;
    if (WrapMode::border && (x < 0.f || x >= 1.f || y < 0.f || y >= 1.f || z < 0.f || z >= 1.f)) return sl_zero_color<color_type>();
    const WrapMode wrapMode{};

    x = wrapMode(x) * (float)tex.width();
    const uint16_t xi = sl_f2u16(x);
    const uint16_t si = sl_f2u16(math::max(x - 1.f, 0.f));

    y = wrapMode(y) * (float)tex.height();
    const uint16_t yi = sl_f2u16(y);
    const uint16_t ti = sl_f2u16(math::max(y - 1.f, 0.f));

    z = wrapMode(z) * (float)tex.depth();
    const uint16_t zi = sl_f2u16(z);
    const uint16_t ri = sl_f2u16(math::max(z - 1.f, 0.f));

    /* the eight taps.  Texel addresses are integer arithmetic (SL_Texture::map_coordinate, SL_Texture.hpp:189-231), so
     * they may be formed any way that gives the same index: the texture's shape and base pointer are read once and
     * the (y, z) row bases are shared by the two x taps (ORDERED layout). */
    const color_type* texels = reinterpret_cast<const color_type*>(tex.data());
    ptrdiff_t i000, i100, i010, i001, i101, i011, i110, i111;
    if (order == SL_TexelOrder::ORDERED)
    {
        const uint32_t W = tex.width(), H = tex.height();
        const uint32_t rowTR = W * ((uint32_t)ti + H * (uint32_t)ri), rowYR = W * ((uint32_t)yi + H * (uint32_t)ri);
        const uint32_t rowTZ = W * ((uint32_t)ti + H * (uint32_t)zi), rowYZ = W * ((uint32_t)yi + H * (uint32_t)zi);
        i000 = (ptrdiff_t)((uint32_t)si + rowTR); i100 = (ptrdiff_t)((uint32_t)xi + rowTR);
        i010 = (ptrdiff_t)((uint32_t)si + rowYR); i110 = (ptrdiff_t)((uint32_t)xi + rowYR);
        i001 = (ptrdiff_t)((uint32_t)si + rowTZ); i101 = (ptrdiff_t)((uint32_t)xi + rowTZ);
        i011 = (ptrdiff_t)((uint32_t)si + rowYZ); i111 = (ptrdiff_t)((uint32_t)xi + rowYZ);
    }
    else
    {
        i000 = tex.map_coordinate<order>(si, ti, ri); i100 = tex.map_coordinate<order>(xi, ti, ri);
        i010 = tex.map_coordinate<order>(si, yi, ri); i110 = tex.map_coordinate<order>(xi, yi, ri);
        i001 = tex.map_coordinate<order>(si, ti, zi); i101 = tex.map_coordinate<order>(xi, ti, zi);
        i011 = tex.map_coordinate<order>(si, yi, zi); i111 = tex.map_coordinate<order>(xi, yi, zi);
    }
    const auto c000 = color_cast<float, T>(texels[i000]);
    const auto c100 = color_cast<float, T>(texels[i100]);
    const auto c010 = color_cast<float, T>(texels[i010]);
    const auto c001 = color_cast<float, T>(texels[i001]);
    const auto c101 = color_cast<float, T>(texels[i101]);
    const auto c011 = color_cast<float, T>(texels[i011]);
    const auto c110 = color_cast<float, T>(texels[i110]);
    const auto c111 = color_cast<float, T>(texels[i111]);

    const float xf = x - math::floor(x);
    const float yf = y - math::floor(y);
    const float zf = z - math::floor(z);
    const float xd = 1.f - xf;
    const float yd = 1.f - yf;
    const float zd = 1.f - zf;

    const auto weight000 = c000 * xd * yd * zd;
    const auto weight100 = c100 * xf * yd * zd;
    const auto weight010 = c010 * xd * yf * zd;
    const auto weight001 = c001 * xd * yd * zf;
    const auto weight101 = c101 * xf * yd * zf;
    const auto weight011 = c011 * xd * yf * zf;
    const auto weight110 = c110 * xf * yf * zd;
    const auto weight111 = c111 * xf * yf * zf;

    const auto ret = math::sum(weight000, weight100, weight010, weight001, weight101, weight011, weight110, weight111);
    return color_cast<T, float>(ret);
}

} // namespace sld
