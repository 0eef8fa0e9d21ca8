/*
 * slor_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the SoftLight draw path: a plain, sequential,
 * scalar C++ re-telling of what the reference's x86/AVX2 source path computes
 * with num_threads(1), function by function, each citing the reference
 * file:line it follows.  It exists so that the CUDA product can be checked on
 * a box where /root/reference is absent, and so that the algorithm is written
 * down once without SIMD.  It is pinned against the compiled reference itself
 * (oracle/_ref/libslref_strict.so; tests/test_oracle_cpu.py) and against the
 * committed golden frames (tests/golden/).
 *
 * Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may load
 * this library.  The product (softlight_b200/libslcu.so) never does.
 *
 * Exports the C ABI of include/slcu.h with the prefix slor_.
 */
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "slor_math.h"
#include "slor_shader_api.h"

#define SLCU_PREFIX slor_
#include "slcu.h"

namespace slo
{

uint32_t g_rcp_table[2048];
uint32_t g_rsqrt_table[4096];

/* rsqrtss on this CPU: entries for inputs in [1, 2) and [2, 4), indexed by the top 11 mantissa bits */
static void measure_host_rsqrt_table(uint32_t* t)
{
    for (uint32_t p = 0; p < 2; ++p)
    {
        for (uint32_t i = 0; i < 2048; ++i)
        {
            float in = u2f(((127u + p) << 23) | (i << 12));
            float out;
            __asm__ volatile("rsqrtss %1, %0" : "=x"(out) : "x"(in));
            t[p * 2048u + i] = f2u(out);
        }
    }
}

/* Measure the rcpss table from the CPU this runs on (the reference executes
 * the real instruction, so "same box" means "same table"). */
static void measure_host_rcp_table(uint32_t* t)
{
    for (uint32_t i = 0; i < 2048; ++i)
    {
        float in = u2f(0x3f800000u | (i << 12));
        float out;
        __asm__ volatile("rcpss %1, %0" : "=x"(out) : "x"(in));
        t[i] = f2u(out);
    }
}

/* ---- shader registry ---------------------------------------------------- */
typedef math::vec4 (*VsFn)(SL_VertexParam&);
typedef bool (*FsFn)(SL_FragmentParam&);
static std::map<std::string, VsFn>& vs_registry() { static std::map<std::string, VsFn> r; return r; }
static std::map<std::string, FsFn>& fs_registry() { static std::map<std::string, FsFn> r; return r; }
struct VsReg { VsReg(const char* n, VsFn f) { vs_registry()[n] = f; } };
struct FsReg { FsReg(const char* n, FsFn f) { fs_registry()[n] = f; } };

#define SL_SHADER_FN static inline
#define SL_CAT2(a, b) a##b
#define SL_CAT(a, b) SL_CAT2(a, b)
#define SL_REGISTER_VERTEX_SHADER(name, fn) static VsReg SL_CAT(vsreg_, fn){name, fn};
#define SL_REGISTER_FRAGMENT_SHADER(name, fn) static FsReg SL_CAT(fsreg_, fn){name, fn};

#include "sl_shaders_basic.inl"
#include "sl_shaders_scene.inl"
#include "sl_shaders_anim_volume.inl"

/* ---- objects ------------------------------------------------------------- */
struct IndexBuffer
{
    int type = 0;
    uint32_t count = 0;
    unsigned bytesPerId = 0;
    std::vector<unsigned char> data;
};

/* the SL_TextureView snapshot SL_Framebuffer keeps per attachment (src/SL_Framebuffer.cpp:855-920) */
struct FboView { bool texels = false; uint16_t w = 0, h = 0, d = 0; int type = 0; };
struct Framebuffer
{
    unsigned numColors = 0;
    size_t colors[4] = {(size_t)-1, (size_t)-1, (size_t)-1, (size_t)-1};
    size_t depth = (size_t)-1;
    FboView colorViews[4];
    FboView depthView;
};

struct Shader
{
    int cullMode, depthTest, depthMask, blend, numVaryings, numOutputs;
    VsFn vs;
    FsFn fs;
    size_t ubo;
};

struct UboFixup { size_t ubo, offset; int kind; size_t id, byteOffset; };

/* SL_FragmentBin (SL_ShaderUtil.hpp:627-647) */
struct Bin
{
    math::vec4 mScreenCoords[3];
    math::vec4 mBarycentricCoords[3];
    math::vec4 mVaryings[12];
    uint64_t primIndex;
};

struct TransformedVert
{
    math::vec4 vert;
    math::vec4 varyings[4];
};

} // namespace slo

using namespace slo;

struct slcu_context
{
    std::vector<SL_VertexBuffer> vbos;
    std::vector<IndexBuffer> ibos;
    std::vector<SL_VertexArray> vaos;
    std::vector<SL_UniformBuffer> ubos;
    std::vector<SL_Texture*> textures;
    std::vector<Framebuffer> fbos;
    std::vector<Shader> shaders;
    std::vector<UboFixup> fixups;
    int32_t viewport[4] = {0, 0, 65535, 65535}; /* SL_ViewportState defaults */
    int32_t scissor[4] = {0, 0, 65535, 65535};
    uint32_t rcpTable[2048];
    uint32_t rsqrtTable[4096];
    std::string err;
    slcu_stats stats{};

    /* per-draw state */
    const Shader* shader = nullptr;
    bool wire = false; /* RENDER_MODE_TRI_WIRE / INDEXED_TRI_WIRE: the rasteriser draws edges only */
    SL_Texture* colorBufs[4] = {};
    SL_Texture* depthBuf = nullptr;
    std::vector<Bin> bins;

    ~slcu_context()
    {
        for (SL_Texture* t : textures) { if (t) { free(t->mTexels); delete t; } }
    }
};

namespace slo
{

/* ===========================================================================
 * Fragment stage
 * =========================================================================*/

/* ---- colour placement: assign_pixel<T> (src/SL_Framebuffer.cpp:44-217) ---- */
static inline unsigned chan_bytes(int type) { static const unsigned b[7] = {1, 2, 4, 8, 2, 4, 8}; return b[type % 7]; }

/* the packed colour types (RGB332 ... RGBA1010102, see the conversions further down): field widths, r lowest */
static const int kPackBits[5][4] = {{3, 3, 2, 0}, {5, 6, 5, 0}, {5, 5, 5, 1}, {4, 4, 4, 4}, {10, 10, 10, 2}};
static const unsigned kPackBytes[5] = {1, 2, 2, 2, 4};
static inline bool is_packed(int type) { return type >= 28; }
static inline uint32_t pack_max(int ci, int j) { return (1u << kPackBits[ci][j]) - 1u; }
static inline void unpack(int ci, uint32_t v, uint32_t c[4])
{
    for (int j = 0, sh = 0; j < 4; sh += kPackBits[ci][j], ++j) c[j] = (v >> sh) & pack_max(ci, j);
}
static inline uint32_t pack(int ci, const uint32_t c[4])
{
    uint32_t v = 0;
    for (int j = 0, sh = 0; j < 4; sh += kPackBits[ci][j], ++j) v |= (c[j] & pack_max(ci, j)) << sh;
    return v;
}
static inline uint32_t load_packed(const unsigned char* p, int ci) { uint32_t v = 0; memcpy(&v, p, kPackBytes[ci]); return v; }
static inline void store_packed(unsigned char* p, int ci, uint32_t v) { memcpy(p, &v, kPackBytes[ci]); }

static uint32_t packed_from_float(int ci, const math::vec4& c); /* rgba_cast<packed, float> */
static math::vec4 float_from_packed(int ci, uint32_t v);        /* rgba_cast<float, packed> */

static void store_color(SL_Texture& t, uint16_t x, uint16_t y, const math::vec4& c)
{
    const size_t idx = (size_t)x + (size_t)t.mWidth * y; /* linear, never swizzled */
    unsigned char* p = t.mTexels + idx * t.mBpt;
    if (is_packed(t.mType)) { store_packed(p, t.mType - 28, packed_from_float(t.mType - 28, c)); return; } /* src/SL_Framebuffer.cpp:85-145 */
    const int kind = t.mType % 7;
    const unsigned n = t.mChannels;
    switch (kind)
    {
        case 0: /* 8U: color_cast<uint8_t,float> = (vec4<u8>)(c*255) */
        {
            const math::vec4_t<uint8_t> r = (math::vec4_t<uint8_t>)(c * 255.f);
            for (unsigned i = 0; i < n; ++i) p[i] = r.v[i];
            break;
        }
        case 1: /* 16U: cvtps2dq + packus_epi32 (vec4f_impl.h) */
        {
            const math::vec4 s = c * 65535.f;
            for (unsigned i = 0; i < n; ++i)
            {
                const int32_t v = cvt_rn(s.v[i]);
                const uint16_t u = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
                memcpy(p + 2 * i, &u, 2);
            }
            break;
        }
        case 4: /* HALF: vec4<float> -> vec4<half> is cvtps2ph round-toward-zero */
        {
            for (unsigned i = 0; i < n; ++i) { const uint16_t h = float_to_half_rz(c.v[i]); memcpy(p + 2 * i, &h, 2); }
            break;
        }
        case 5: /* FLOAT */
            memcpy(p, c.v, 4 * n);
            break;
        case 6: /* DOUBLE */
        {
            for (unsigned i = 0; i < n; ++i) { const double d = (double)c.v[i]; memcpy(p + 8 * i, &d, 8); }
            break;
        }
        case 2: /* 32U: color_cast<uint32_t,float> = (vec4<u32>)(c * (float)UINT32_MAX); the vec4 conversion is cvtps2dq
                 * (x86/vec4f_impl.h:333-342): signed, RNE, 0x80000000 from 0.5 upwards -- reproduced, not repaired */
        {
            const math::vec4 s = c * 4294967296.f;
            for (unsigned i = 0; i < n; ++i) { const uint32_t u = (uint32_t)cvt_rn(s.v[i]); memcpy(p + 4 * i, &u, 4); }
            break;
        }
        case 3: /* 64U: the generic element-wise (uint64_t)float of vec4_impl.h:184-187 as g++ compiles it: cvttss2si below
                 * 2^63 (NaN and negatives included: unordered compares take this branch), otherwise cvttss2si(f - 2^63) ^ 2^63 */
        {
            const math::vec4 s = c * 18446744073709551616.f;
            for (unsigned i = 0; i < n; ++i)
            {
                const float f = s.v[i];
                auto cvtt64 = [](float x) -> int64_t { return (x >= -9223372036854775808.f && x < 9223372036854775808.f) ? (int64_t)x : INT64_MIN; };
                const uint64_t u = !(f >= 9223372036854775808.f) ? (uint64_t)cvtt64(f) : ((uint64_t)cvtt64(f - 9223372036854775808.f) ^ 0x8000000000000000ull);
                memcpy(p + 8 * i, &u, 8);
            }
            break;
        }
        default:
            break;
    }
}

/* vec4<uint32_t> -> vec4<float> (x86/vec4f_impl.h:236-245): (float)(v >> 1) twice plus (float)(v & 1), every step rounded */
static inline float u32_to_float(uint32_t v)
{
    const float v2f = (float)(int32_t)(v >> 1);
    const float v1f = (float)(int32_t)(v & 1u);
    return (v2f + v2f) + v1f;
}

static math::vec4 load_color_as_float(const SL_Texture& t, uint16_t x, uint16_t y)
{
    /* assign_alpha_pixel's destination read (src/SL_Framebuffer.cpp:284-312),
     * including the 3-channel quirk that replicates channel 0 (:308). */
    const size_t idx = (size_t)x + (size_t)t.mWidth * y;
    const unsigned char* p = t.mTexels + idx * t.mBpt;
    if (is_packed(t.mType)) return float_from_packed(t.mType - 28, load_packed(p, t.mType - 28)); /* :371-590 */
    const int kind = t.mType % 7;
    const unsigned n = t.mChannels;
    float ch[4];
    float lo, hi, scale;
    auto rd = [&](unsigned i) -> float {
        switch (kind)
        {
            case 0: return (float)p[i];
            case 1: { uint16_t u; memcpy(&u, p + 2 * i, 2); return (float)u; }
            case 2: { uint32_t u; memcpy(&u, p + 4 * i, 4); return u32_to_float(u); }
            case 3: { uint64_t u; memcpy(&u, p + 8 * i, 8); return (float)u; } /* cvtsi2ss, or the halve / round-to-odd / double sequence: correctly rounded */
            case 4: { uint16_t u; memcpy(&u, p + 2 * i, 2); return half_to_float(u); }
            case 5: { float f; memcpy(&f, p + 4 * i, 4); return f; }
            case 6: { double d; memcpy(&d, p + 8 * i, 8); return (float)d; }
            default: return 0.f;
        }
    };
    switch (kind)
    {
        case 0: lo = 0.f; hi = 255.f; scale = 0.00392156862745f; break;
        case 1: lo = 0.f; hi = 65535.f; scale = 1.f / 65535.f; break;
        case 2: lo = 0.f; hi = u32_to_float(0xffffffffu); scale = 1.f / 4294967296.f; break;
        case 3: lo = 0.f; hi = (float)0xffffffffffffffffull; scale = 1.f / 18446744073709551616.f; break;
        default: lo = 0.f; hi = 1.f; scale = 1.f; break;
    }
    switch (n)
    {
        case 1: ch[0] = rd(0); ch[1] = lo; ch[2] = lo; ch[3] = hi; break;
        case 2: ch[0] = rd(0); ch[1] = rd(1); ch[2] = lo; ch[3] = hi; break;
        case 3: ch[0] = rd(0); ch[1] = rd(0); ch[2] = rd(0); ch[3] = hi; break;
        default: ch[0] = rd(0); ch[1] = rd(1); ch[2] = rd(2); ch[3] = rd(3); break;
    }
    math::vec4 d{ch[0], ch[1], ch[2], ch[3]};
    if (kind <= 3) d = d * scale;
    return d;
}

/* assign_alpha_pixel<T> (src/SL_Framebuffer.cpp:270-368) */
static void blend_color(SL_Texture& t, uint16_t x, uint16_t y, const math::vec4& rgba, int blendMode)
{
    math::vec4 d = load_color_as_float(t, x, y);
    const math::vec4 srcAlpha{rgba.v[3]};
    const math::vec4 modulation = math::vec4{1.f} - srcAlpha;
    if (blendMode == 1) /* SL_BLEND_ALPHA */
    {
        const math::vec4 dstMod = modulation * d.v[3];
        const math::vec4 dstAlpha = dstMod + srcAlpha;
        const math::vec4 dstRgb = math::fmadd(rgba, srcAlpha, (d * dstMod)) * math::rcp(dstAlpha);
        d = math::vec4{dstRgb.v[0], dstRgb.v[1], dstRgb.v[2], dstAlpha.v[0]};
    }
    else if (blendMode == 2) /* SL_BLEND_PREMULTIPLED_ALPHA */
    {
        d = math::fmadd(d, modulation, rgba);
    }
    else if (blendMode == 3) /* SL_BLEND_ADDITIVE */
    {
        d = math::fmadd(rgba, srcAlpha, d);
    }
    else if (blendMode == 4) /* SL_BLEND_SCREEN */
    {
        d = (rgba * srcAlpha) + (d * modulation);
    }
    d = math::clamp(d, math::vec4{0.f, 0.f, 0.f, 0.f}, math::vec4{1.f, 1.f, 1.f, 1.f});
    store_color(t, x, y, d);
}

/* depth texel read (src/SL_TriRasterizer.cpp:40-102) / write (src/SL_FragmentProcessor.cpp:526-529) */
static inline float load_depth(const SL_Texture& t, size_t idx)
{
    switch (t.mBpt)
    {
        case 2: { uint16_t h; memcpy(&h, t.mTexels + idx * 2, 2); return half_to_float(h); }
        case 4: { float f; memcpy(&f, t.mTexels + idx * 4, 4); return f; }
        default: { double d; memcpy(&d, t.mTexels + idx * 8, 8); return (float)d; }
    }
}
static inline void store_depth(SL_Texture& t, size_t idx, float z)
{
    switch (t.mBpt)
    {
        case 2: { const uint16_t h = float_to_half_rn(z); memcpy(t.mTexels + idx * 2, &h, 2); break; }
        case 4: memcpy(t.mTexels + idx * 4, &z, 4); break;
        default: { const double d = (double)z; memcpy(t.mTexels + idx * 8, &d, 8); break; }
    }
}

/* SL_DepthFunc* (SL_ShaderUtil.hpp:240-507): _CMP_xx_OQ, false on NaN */
static inline bool depth_cmp(int func, float z, float d)
{
    switch (func)
    {
        case 0: return true;
        case 1: return z < d;
        case 2: return z <= d;
        case 3: return z > d;
        case 4: return z >= d;
        case 5: return z == d;
        case 6: return (z < d) || (z > d); /* NEQ_OQ */
        default: return false;
    }
}
/* the functors' SCALAR overloads (SL_ShaderUtil.hpp:240-507, `constexpr int operator()(float a, float b)`), which the point,
 * line and wireframe rasterisers call (src/SL_PointRasterizer.cpp:58, src/SL_LineRasterizer.cpp:64, src/SL_TriRasterizer.cpp:
 * 177): the same as the SIMD ones except NOT_EQUAL, which is the C++ `a != b` -- TRUE when the fragment's depth is NaN (a
 * zero-length line that is clipped interpolates with 0 * inf) */
static inline bool depth_cmp_scalar(int func, float z, float d)
{
    return func == 6 ? (z != d) : depth_cmp(func, z, d);
}

/* SL_ScanlineBounds (SL_ScanlineBounds.hpp:125-347) */
struct ScanlineBounds
{
    float v0y, v1y, v0x, v1x, p20y, p20x, p21xy, p10xy;

    void init(math::vec4 p0, math::vec4 p1, math::vec4 p2)
    {
        /* sl_sort_minmax x3 (:60-68,139-141): swap when a.y < b.y (ordered) */
        if (p0.v[1] < p1.v[1]) std::swap(p0, p1);
        if (p0.v[1] < p2.v[1]) std::swap(p0, p2);
        if (p1.v[1] < p2.v[1]) std::swap(p1, p2);
        /* :164-181 */
        const math::vec4 p2p0 = p2 - p0;
        const math::vec4 p2p1 = p2 - p1;
        const math::vec4 p1p0 = p1 - p0;
        v0y = p0.v[1];
        v1y = p1.v[1];
        v0x = p0.v[0];
        v1x = p1.v[0];
        p20y = x86_rcp(p2p0.v[1]);
        p20x = p2p0.v[0];
        p21xy = p2p1.v[0] * x86_rcp(p2p1.v[1]);
        p10xy = p1p0.v[0] * x86_rcp(p1p0.v[1]);
    }

    /* :326-347 */
    void step(float yf, int32_t& xMin, int32_t& xMax) const
    {
        const float d0 = yf - v0y;
        const float d1 = yf - v1y;
        const float alpha = d0 * p20y;
        const float lo = sfma(p20x, alpha, v0x);
        const float pdv0 = sfma(p10xy, d0, v0x);
        const float pdv1 = sfma(p21xy, d1, v1x);
        const float hi = sign_bit(d1) ? pdv1 : pdv0; /* blendv on the sign bit of d1 */
        xMin = cvt_rn(smin(lo, hi));
        xMax = cvt_rn(smax(lo, hi));
    }
};

/* One queued fragment (SL_FragCoord, SL_ShaderUtil.hpp:668-689) */
struct QueuedFrag { math::vec4 bc; uint16_t x, y; float z; };

/* debug aid: SL_FRAG_TRACE=<file> records the (x,y) of every fragment-shader call */
static std::vector<uint32_t> gFragTrace;
static const char* gFragTracePath = getenv("SL_FRAG_TRACE");

/* SL_FragmentProcessor::flush_tri_fragments (src/SL_FragmentProcessor.cpp:416-531) for ONE fragment */
static void shade_fragment(slcu_context* c, const Bin& bin, const QueuedFrag& q)
{
    const Shader& sh = *c->shader;
    /* :438-460 perspective correction */
    const math::vec4 homogenous{bin.mScreenCoords[0].v[3], bin.mScreenCoords[1].v[3], bin.mScreenCoords[2].v[3], 0.f};
    math::vec4 bc = q.bc * homogenous;
    const float d = (bc.v[0] + bc.v[1]) + (bc.v[2] + bc.v[3]);
    const float persp = x86_rcp(d);
    bc = bc * persp;

    /* interpolate_tri_varyings, AVX2 branch (:163-193): all four slots */
    SL_FragmentParam fp;
    fp.pUniforms = &c->ubos[sh.ubo];
    for (int s = 0; s < 4; ++s)
    {
        for (int k = 0; k < 4; ++k)
        {
            float v = bc.v[0] * bin.mVaryings[s + 0].v[k];
            v = sfma(bc.v[1], bin.mVaryings[s + 4].v[k], v);
            v = sfma(bc.v[2], bin.mVaryings[s + 8].v[k], v);
            fp.pVaryings[s].v[k] = v;
        }
    }
    fp.coord.x = q.x;
    fp.coord.y = q.y;
    fp.coord.depth = q.z;
    for (int i = 0; i < 4; ++i) fp.pOutputs[i] = math::vec4{0.f};

    /* the metric counts fragments that land inside the w*h texels (see oracle/ref/slref_shim.cpp counting_fs) */
    if ((size_t)q.x + (size_t)c->depthBuf->mWidth * q.y < (size_t)c->depthBuf->mWidth * c->depthBuf->mHeight) c->stats.fragsShaded++;
    if (gFragTracePath) gFragTrace.push_back((uint32_t)q.x | ((uint32_t)q.y << 16));
    const bool haveOutputs = sh.fs(fp); /* :501 */
    if (haveOutputs)
    {
        for (int i = sh.numOutputs - 1; i >= 0; --i) /* :507-523 */
        {
            if (sh.blend != 0) blend_color(*c->colorBufs[i], fp.coord.x, fp.coord.y, fp.pOutputs[i], sh.blend);
            else store_color(*c->colorBufs[i], fp.coord.x, fp.coord.y, fp.pOutputs[i]);
        }
    }
    if (sh.depthMask) /* :526-529, regardless of the shader's result */
    {
        store_depth(*c->depthBuf, (size_t)fp.coord.x + (size_t)c->depthBuf->mWidth * fp.coord.y, fp.coord.depth);
    }
}

/* SL_TriRasterizer::render_triangle_simd (src/SL_TriRasterizer.cpp:413-559), one thread */
static void render_triangle(slcu_context* c, const Bin& bin)
{
    const Shader& sh = *c->shader;
    SL_Texture& depthBuffer = *c->depthBuf;
    const math::vec4& p0 = bin.mScreenCoords[0];
    const math::vec4& p1 = bin.mScreenCoords[1];
    const math::vec4& p2 = bin.mScreenCoords[2];

    /* :431-433 */
    const int32_t bboxMinY = cvt_rn(smin(smin(p0.v[1], p1.v[1]), p2.v[1]));
    const int32_t bboxMaxY = cvt_rn(smax(smax(p0.v[1], p1.v[1]), p2.v[1]));
    int32_t y = bboxMinY; /* sl_scanline_offset(1, 0, y) == 0 */
    if (y >= bboxMaxY) return;

    const float depth[4] = {p0.v[2], p1.v[2], p2.v[2], 0.f}; /* :444-445 */
    ScanlineBounds scanline;
    scanline.init(p0, p1, p2);
    const math::vec4 bc0 = bin.mBarycentricCoords[0];
    const math::vec4 bc1 = bin.mBarycentricCoords[1];
    const math::vec4 bc2 = bin.mBarycentricCoords[2];

    static thread_local std::vector<QueuedFrag> queue;
    queue.clear();

    do
    {
        int32_t xMin, xMax;
        const float yf = (float)y;
        scanline.step(yf, xMin, xMax);
        if (!(xMin < xMax)) { y += 1; continue; }
        if (y < (int32_t)depthBuffer.mHeight) c->stats.spans++; /* rows a read-back can see */

        /* :471-481 */
        math::vec4 bcY;
        for (int k = 0; k < 4; ++k) bcY.v[k] = sfma(bc1.v[k], yf, bc2.v[k]);
        int32_t x4[4] = {xMin, xMin + 1, xMin + 2, xMin + 3};
        math::vec4 bc[4];
        for (int j = 0; j < 4; ++j)
        {
            const float xf = (float)x4[j];
            for (int k = 0; k < 4; ++k) bc[j].v[k] = (xf * bc0.v[k]) + bcY.v[k];
        }
        math::vec4 bcX;
        for (int k = 0; k < 4; ++k) bcX.v[k] = bc0.v[k] * 4.f;
        int64_t pd = (int64_t)xMin + (int64_t)depthBuffer.mWidth * y;

        do
        {
            /* :484-490 */
            bool m[4];
            float z[4];
            int any = 0;
            for (int j = 0; j < 4; ++j)
            {
                /* _sl_mul_vec4_mat4_ps (:376-393): (w+z) + (y+x) */
                const float rx = depth[0] * bc[j].v[0];
                const float ry = depth[1] * bc[j].v[1];
                const float rz = depth[2] * bc[j].v[2];
                const float rw = depth[3] * bc[j].v[3];
                z[j] = (rw + rz) + (ry + rx);
                const bool xBound = x4[j] < xMax;
                /* the reference loads 4 texels unconditionally; lanes past xMax are masked off */
                const float dv = xBound ? load_depth(depthBuffer, (size_t)(pd + j)) : 0.f;
                m[j] = xBound && depth_cmp(sh.depthTest, z[j], dv);
                any |= m[j];
            }
            if (any)
            {
                /* :496-498: in-place compaction of the barycentric registers (kept for the next groups!) */
                bc[2] = m[2] ? bc[2] : bc[3];
                bc[1] = m[1] ? bc[1] : bc[2];
                bc[0] = m[0] ? bc[0] : bc[1];
                /* :500-526: survivor s gets bc[s] (new) and the coords of its own lane */
                int s = 0;
                for (int j = 0; j < 4; ++j)
                {
                    if (m[j])
                    {
                        QueuedFrag q;
                        q.bc = bc[s];
                        q.x = (uint16_t)x4[j];
                        q.y = (uint16_t)y;
                        q.z = z[j];
                        queue.push_back(q);
                        ++s;
                    }
                }
                /* :529-533: flush when the queue is nearly full (SL_SHADER_MAX_QUEUED_FRAGS - 4).  This is observable:
                 * a span that runs past the row end wraps onto the start of the next row, which the same triangle's
                 * next span may cover too; whether that span's depth test sees the wrapped fragments depends on
                 * whether a flush happened in between. */
                if (queue.size() > 600 - 4)
                {
                    for (const QueuedFrag& q : queue) shade_fragment(c, bin, q);
                    queue.clear();
                }
            }
            for (int j = 0; j < 4; ++j)
            {
                for (int k = 0; k < 4; ++k) bc[j].v[k] = bc[j].v[k] + bcX.v[k]; /* :539-542 */
                x4[j] += 4;
            }
            pd += 4;
        } while (x4[0] < xMax); /* :548 */

        y += 1;
    } while (y < bboxMaxY);

    for (const QueuedFrag& q : queue) shade_fragment(c, bin, q);
}

static inline int32_t wire_edge_width(int32_t a, int32_t b)
{
    const int32_t d = (int32_t)((uint32_t)a - (uint32_t)b);
    const int32_t ad = d < 0 ? (int32_t)(0u - (uint32_t)d) : d;
    return ad > 1 ? ad : 1;
}

/* SL_TriRasterizer::render_wireframe (src/SL_TriRasterizer.cpp:116-202), one thread: per scan line only the pixels
 * next to the two edges -- d0 pixels from xMin (the distance the left bound moved since the previous line, at least 1)
 * and the last d1 pixels before xMax -- scalar barycentrics (fmadd), depth = dot(depth, bc). */
static void render_wireframe(slcu_context* c, const Bin& bin)
{
    const Shader& sh = *c->shader;
    SL_Texture& depthBuffer = *c->depthBuf;
    const math::vec4& p0 = bin.mScreenCoords[0];
    const math::vec4& p1 = bin.mScreenCoords[1];
    const math::vec4& p2 = bin.mScreenCoords[2];
    const int32_t bboxMinY = cvt_tr(smin(p2.v[1], smin(p0.v[1], p1.v[1]))); /* :136-137, C casts */
    const int32_t bboxMaxY = cvt_tr(smax(p2.v[1], smax(p0.v[1], p1.v[1])));
    ScanlineBounds scanline;
    scanline.init(p0, p1, p2);
    const math::vec4 bc0 = bin.mBarycentricCoords[0];
    const math::vec4 bc1 = bin.mBarycentricCoords[1];
    const math::vec4 bc2 = bin.mBarycentricCoords[2];
    const math::vec4 depth{p0.v[2], p1.v[2], p2.v[2], 0.f};

    static thread_local std::vector<QueuedFrag> queue;
    queue.clear();
    for (int32_t y = bboxMinY; y < bboxMaxY; ++y)
    {
        const float yf = (float)y;
        math::vec4 bcY;
        for (int k = 0; k < 4; ++k) bcY.v[k] = sfma(bc1.v[k], k < 3 ? yf : 0.f, bc2.v[k]); /* :149 */
        int32_t xMinMax0[2], xMinMax1[2];
        scanline.step(yf, xMinMax0[0], xMinMax0[1]);
        scanline.step(yf - 1.f, xMinMax1[0], xMinMax1[1]);
        /* :157-158 max(abs(a - b), 1) in wrapping 32-bit arithmetic: on the line above the top vertex the bounds are the x86
         * "integer indefinite" 0x80000000 (cvtss2si of NaN/inf), a - b and abs() wrap back to INT_MIN and the max picks 1 */
        const int32_t d0 = wire_edge_width(xMinMax0[0], xMinMax1[0]);
        const int32_t d1 = wire_edge_width(xMinMax0[1], xMinMax1[1]);
        const int64_t rowBase = (int64_t)depthBuffer.mWidth * y;
        bool counted = false;
        for (int32_t ix = 0, x = xMinMax0[0]; (uint32_t)x < (uint32_t)xMinMax0[1]; ++ix, ++x)
        {
            if (ix == d0) /* :165-169 skip to the start of the right edge */
            {
                x = std::max(xMinMax0[0], (int32_t)((uint32_t)xMinMax0[1] - (uint32_t)d1 - 1u));
                continue;
            }
            if (!counted && y < (int32_t)depthBuffer.mHeight) { c->stats.spans++; counted = true; }
            const float xf = (float)x;
            math::vec4 bc;
            for (int k = 0; k < 4; ++k) bc.v[k] = sfma(bc0.v[k], k < 3 ? xf : 0.f, bcY.v[k]); /* :173 */
            const float z = math::dot(depth, bc);
            const float d = load_depth(depthBuffer, (size_t)(rowBase + x));
            if (!depth_cmp_scalar(sh.depthTest, z, d)) continue;
            QueuedFrag q;
            q.bc = bc;
            q.x = (uint16_t)x;
            q.y = (uint16_t)y;
            q.z = z;
            queue.push_back(q);
            if (queue.size() == 600) /* :188-192 SL_SHADER_MAX_QUEUED_FRAGS */
            {
                for (const QueuedFrag& f : queue) shade_fragment(c, bin, f);
                queue.clear();
            }
        }
    }
    for (const QueuedFrag& f : queue) shade_fragment(c, bin, f);
}

/* SL_VertexProcessor::flush_rasterizer<SL_TriRasterizer> (src/SL_VertexProcessor.cpp:84-181), one thread */
static void flush_bins(slcu_context* c)
{
    std::vector<uint32_t> ids(c->bins.size());
    for (size_t i = 0; i < ids.size(); ++i) ids[i] = (uint32_t)i;
    if (c->shader->blend != 0)
    {
        /* :110-119 LSD radix sort by primIndex == a stable sort */
        std::stable_sort(ids.begin(), ids.end(), [&](uint32_t a, uint32_t b) { return c->bins[a].primIndex < c->bins[b].primIndex; });
    }
    for (uint32_t id : ids) { if (c->wire) render_wireframe(c, c->bins[id]); else render_triangle(c, c->bins[id]); } /* SL_TriRasterizer::dispatch_bins :970-1005 */
    c->bins.clear();
}

/* ===========================================================================
 * Vertex stage
 * =========================================================================*/

/* SL_TriProcessor::push_bin (src/SL_TriProcessor.cpp:248-342) */
static void push_bin(slcu_context* c, uint64_t primIndex, const TransformedVert& a, const TransformedVert& b, const TransformedVert& cc)
{
    const math::vec4& p0 = a.vert;
    const math::vec4& p1 = b.vert;
    const math::vec4& p2 = cc.vert;
    const math::vec4 bboxMin = math::min(math::min(p0, p1), p2);
    const math::vec4 bboxMax = math::max(math::max(p0, p1), p2);
    const int isPrimHidden = (bboxMax.v[0] - bboxMin.v[0] < 1.f) || (bboxMax.v[1] - bboxMin.v[1] < 1.f);
    if (isPrimHidden) return;

    /* :271-287 */
    const float xy0 = p0.v[0] - p1.v[0], xy1 = p0.v[1] - p1.v[1];
    const float zy0 = p2.v[0] - p1.v[0], zy1 = p2.v[1] - p1.v[1];
    /* math::cross(vec2 zy, vec2 xy) = zy.x*xy.y - zy.y*xy.x (generic/vec_utils_impl.h) */
    const float denom = x86_rcp((zy0 * xy1) - (zy1 * xy0));
    const math::vec4 ptX{p0.v[0], p1.v[0], p2.v[0], 0.f}; /* transpose(mat4{p0,p1,p2,0}).m[0] */
    const math::vec4 ptY{p0.v[1], p1.v[1], p2.v[1], 0.f};
    const math::vec4 one{1.f};
    const math::vec4 ddx = math::cross(ptY, one);
    const math::vec4 ddy = math::cross(one, ptX);
    const math::vec4 ddz = math::cross(ptX, ptY);

    if (c->bins.size() >= 1024) flush_bins(c); /* :290-303 SL_SHADER_MAX_BINNED_PRIMS */

    Bin bin;
    bin.mScreenCoords[0] = p0;
    bin.mScreenCoords[1] = p1;
    bin.mScreenCoords[2] = p2;
    bin.mBarycentricCoords[0] = ddx * denom;
    bin.mBarycentricCoords[1] = ddy * denom;
    bin.mBarycentricCoords[2] = ddz * denom;
    for (int i = 0; i < 12; ++i) bin.mVaryings[i] = math::vec4{0.f};
    for (int s = 0; s < c->shader->numVaryings; ++s)
    {
        bin.mVaryings[s + 0] = a.varyings[s];
        bin.mVaryings[s + 4] = b.varyings[s];
        bin.mVaryings[s + 8] = cc.varyings[s];
    }
    bin.primIndex = primIndex;
    c->bins.push_back(bin);
    c->stats.trisBinned++;
}

/* sl_perspective_divide (SL_PostVertexTransform.hpp:160-178) */
static inline void perspective_divide(math::vec4& v)
{
    const float wInv = x86_rcp(v.v[3]);
    v = math::vec4{v.v[0] * wInv, v.v[1] * wInv, v.v[2] * wInv, wInv};
}

/* sl_ndc_to_screen_coords (SL_PostVertexTransform.hpp:412-449, FMA branch) */
static inline void ndc_to_screen(math::vec4& p, const math::vec4& dims)
{
    const float whx = dims.v[2] * 0.5f; /* wh = (halfDims.z, halfDims.w, 1, 1) */
    const float why = dims.v[3] * 0.5f;
    float vx = sfma(1.f + p.v[0], whx, dims.v[0]);
    float vy = sfma(1.f + p.v[1], why, dims.v[1]);
    vx = smax(math::floor(vx), 0.f);
    vy = smax(math::floor(vy), 0.f);
    p.v[0] = vx;
    p.v[1] = vy;
}

/* sl_ndc_clip_status (SL_PostVertexTransform.hpp:590-616) */
static inline int clip_status(const math::vec4& c0, const math::vec4& c1, const math::vec4& c2)
{
    const math::vec4* cs[3] = {&c0, &c1, &c2};
    bool vis = true, part = false;
    for (int i = 0; i < 3; ++i)
    {
        const float w = cs[i]->v[3];
        const float nw = u2f(f2u(w) | 0x80000000u);
        for (int k = 0; k < 4; ++k) vis = vis && (nw <= cs[i]->v[k]) && (w >= cs[i]->v[k]);
        part = part || (w > 0.f);
    }
    return (vis ? 3 : 0) | (part ? 1 : 0);
}

/* SL_TriProcessor::clip_and_process_tris (src/SL_TriProcessor.cpp:370-519) */
static void clip_and_process_tris(slcu_context* c, uint64_t primIndex, const math::vec4& viewportDims,
                                  const TransformedVert& a, const TransformedVert& b, const TransformedVert& cc)
{
    const unsigned numVarys = (unsigned)c->shader->numVaryings;
    unsigned numTotalVerts = 3;
    math::vec4 tempVerts[9], newVerts[9];
    math::vec4 tempVarys[9 * 4], newVarys[9 * 4];
    const math::vec4 clipEdges[6] = {
        {1.f, 0.f, 0.f, 1.f}, {-1.f, 0.f, 0.f, 1.f}, {0.f, 1.f, 0.f, 1.f},
        {0.f, -1.f, 0.f, 1.f}, {0.f, 0.f, 1.f, 1.f}, {0.f, 0.f, -1.f, 1.f}};

    newVerts[0] = a.vert;
    newVerts[1] = b.vert;
    newVerts[2] = cc.vert;
    for (unsigned i = 0; i < numVarys; ++i)
    {
        newVarys[0 * 4 + i] = a.varyings[i];
        newVarys[1 * 4 + i] = b.varyings[i];
        newVarys[2 * 4 + i] = cc.varyings[i];
    }

    for (const math::vec4& edge : clipEdges)
    {
        unsigned numNewVerts = 0;
        unsigned j = numTotalVerts - 1;
        math::vec4 p0 = newVerts[numTotalVerts - 1];
        float t0 = math::dot(p0, edge);
        int visible0 = !sign_bit(t0);
        for (unsigned k = 0; k < numTotalVerts; ++k)
        {
            const math::vec4& p1 = newVerts[k];
            const float t1 = math::dot(p1, edge);
            const int visible1 = !sign_bit(t1);
            if (visible0 ^ visible1)
            {
                const float t = t0 * x86_rcp(t0 - t1);
                tempVerts[numNewVerts] = math::mix(p0, p1, t);
                for (unsigned i = 0; i < numVarys; ++i)
                {
                    tempVarys[numNewVerts * 4 + i] = math::mix(newVarys[j * 4 + i], newVarys[k * 4 + i], t);
                }
                ++numNewVerts;
            }
            if (visible1)
            {
                tempVerts[numNewVerts] = p1;
                for (unsigned i = 0; i < numVarys; ++i) tempVarys[numNewVerts * 4 + i] = newVarys[k * 4 + i];
                ++numNewVerts;
            }
            j = k;
            p0 = p1;
            t0 = t1;
            visible0 = visible1;
        }
        if (!numNewVerts) return;
        numTotalVerts = numNewVerts;
        for (unsigned i = 0; i < numNewVerts; ++i)
        {
            newVerts[i] = tempVerts[i];
            for (unsigned v = 0; v < 4; ++v) newVarys[i * 4 + v] = tempVarys[i * 4 + v];
        }
    }
    if (numTotalVerts < 3) return;

    for (unsigned i = 0; i < numTotalVerts; ++i)
    {
        perspective_divide(newVerts[i]);
        ndc_to_screen(newVerts[i], viewportDims);
    }
    /* NB (:482-499): the reference transforms vertices in groups of three, so
     * for 4,5,7,8 vertices it also transforms stale entries past the end; those
     * are never read. */

    TransformedVert p0, p1, p2;
    p0.vert = newVerts[0];
    for (unsigned i = 0; i < numVarys; ++i) p0.varyings[i] = newVarys[i];
    for (unsigned i = numTotalVerts - 2; i--;) /* :505-518, reverse fan order */
    {
        const unsigned j = i + 1;
        const unsigned k = i + 2;
        p1.vert = newVerts[j];
        p2.vert = newVerts[k];
        for (unsigned v = 0; v < numVarys; ++v)
        {
            p1.varyings[v] = newVarys[j * 4 + v];
            p2.varyings[v] = newVarys[k * 4 + v];
        }
        push_bin(c, primIndex, p0, p1, p2);
    }
}

/* face_determinant, SSE4.1 branch (src/SL_TriProcessor.cpp:40-66) */
static inline float face_determinant(const math::vec4& p0, const math::vec4& p1, const math::vec4& p2)
{
    /* col4 = (p0.x, p0.y, p0.w, 0); p1_201 = (p1.w,p1.x,p1.y,.), p2_120 = (p2.y,p2.w,p2.x,.)
     * p1_120 = (p1.y,p1.w,p1.x,.), p2_201 = (p2.w,p2.x,p2.y,.) */
    const float col3[3] = {p1.v[3] * p2.v[1], p1.v[0] * p2.v[3], p1.v[1] * p2.v[0]};
    const float col1[3] = {p1.v[1] * p2.v[3], p1.v[3] * p2.v[0], p1.v[0] * p2.v[1]};
    const float col4[3] = {p0.v[0], p0.v[1], p0.v[3]};
    float mul2[3];
    for (int i = 0; i < 3; ++i) mul2[i] = (col1[i] - col3[i]) * col4[i];
    /* hi = movehdup(mul2) -> lane0 = mul2[1]; lo = hi + mul2 (ss) = mul2[1]+mul2[0]; w = mul2[2]; w + lo */
    return mul2[2] + (mul2[1] + mul2[0]);
}

/* sl_scissor + SL_ViewportState::scissor_matrix / viewport_rect (src/SL_ViewportState.cpp:17-111) */
static math::mat4 scissor_matrix(const int32_t sc[4], float fboW, float fboH)
{
    const math::vec4 fboInv = math::rcp(math::vec4{fboW, fboH, fboW, fboH});
    const math::vec4 s = math::vec4{(float)sc[0], (float)sc[1], (float)sc[2], (float)sc[3]} * fboInv;
    float x = s.v[0], y = s.v[1], w = s.v[2], h = s.v[3];
    if (x < 0.f) { w += x; x = 0.f; }
    if (y < 0.f) { h += y; y = 0.f; }
    w = math::min(1.f - x, w);
    h = math::min(1.f - y, h);
    const float nm00 = x86_rcp(w);
    const float nm11 = x86_rcp(h);
    const float nm30 = nm00 - 1.f;
    const float nm31 = nm11 - 1.f;
    const float m30 = sfma(x, (-2.f * nm00), nm30);
    const float m31 = sfma(y, (-2.f * nm11), nm31);
    math::mat4 r;
    r.m[0] = math::vec4{nm00, 0.f, 0.f, 0.f};
    r.m[1] = math::vec4{0.f, nm11, 0.f, 0.f};
    r.m[2] = math::vec4{0.f, 0.f, 1.f, 0.f};
    r.m[3] = math::vec4{m30, m31, 0.f, 1.f};
    return r;
}

static inline float min3(float a, float b, float c) { return smin(c, smin(a, b)); } /* scalarf_utils_impl.h:32-38 */

static math::vec4 viewport_rect(const int32_t vp[4], const int32_t sc[4], float fboW, float fboH)
{
    const float s0 = (float)sc[0], s1 = (float)sc[1], s2 = (float)sc[2], s3 = (float)sc[3];
    math::vec4 d{(float)vp[0], (float)vp[1], (float)vp[2], (float)vp[3]};
    d.v[0] = math::clamp(d.v[0] + s0, 0.f, fboW);
    d.v[1] = math::clamp(d.v[1] + s1, 0.f, fboH);
    d.v[2] = min3(d.v[2], (s0 + s2) - d.v[0], fboW - d.v[0]);
    d.v[3] = min3(d.v[3], (s1 + s3) - d.v[1], fboH - d.v[1]);
    return d;
}

/* SL_TriProcessor::process_verts (src/SL_TriProcessor.cpp:527-640), one thread */
static void process_verts(slcu_context* c, const slcu_mesh& m, size_t instanceId, bool usingIndices,
                          const math::mat4& scissorMat, const math::vec4& viewportDims)
{
    const Shader& sh = *c->shader;
    const SL_VertexArray& vao = c->vaos[m.vaoId];
    const IndexBuffer* pIbo = vao.has_index_buffer() ? &c->ibos[vao.mIboId] : nullptr;

    SL_VertexParam params;
    params.pUniforms = &c->ubos[sh.ubo];
    params.instanceId = instanceId;
    params.pVao = &vao;
    params.pVbo = &c->vbos[vao.mVboId];

    const size_t numElements = m.elementEnd - m.elementBegin;
    const size_t primOffset = numElements * instanceId;
    TransformedVert v[3];

    for (size_t i = m.elementBegin; i < m.elementEnd; i += 3)
    {
        c->stats.trisSubmitted++;
        size_t vertId[3];
        if (usingIndices) /* get_next_vertex3 (:125-234) */
        {
            for (int k = 0; k < 3; ++k)
            {
                const unsigned char* p = pIbo->data.data() + (i + k) * pIbo->bytesPerId;
                switch (pIbo->type)
                {
                    case 0: vertId[k] = *p; break;
                    case 1: { uint16_t u; memcpy(&u, p, 2); vertId[k] = u; break; }
                    default: { uint32_t u; memcpy(&u, p, 4); vertId[k] = u; break; }
                }
            }
        }
        else
        {
            vertId[0] = i; vertId[1] = i + 1; vertId[2] = i + 2;
        }
        for (int k = 0; k < 3; ++k) /* :584-594 */
        {
            params.vertId = vertId[k];
            params.pVaryings = v[k].varyings;
            v[k].vert = sh.vs(params);
        }
        if (sh.cullMode != 2) /* :597-612 */
        {
            const float det = face_determinant(v[0].vert, v[1].vert, v[2].vert);
            const bool culled = (sh.cullMode == 1) ^ sign_bit(det);
            if (culled) continue;
        }
        for (int k = 0; k < 3; ++k) v[k].vert = scissorMat * v[k].vert; /* :615-617 */

        const int visStatus = clip_status(v[0].vert, v[1].vert, v[2].vert);
        if (visStatus == 3)
        {
            for (int k = 0; k < 3; ++k) { perspective_divide(v[k].vert); ndc_to_screen(v[k].vert, viewportDims); }
            push_bin(c, primOffset + i, v[0], v[1], v[2]);
        }
        else if (visStatus == 1)
        {
            clip_and_process_tris(c, primOffset + i, viewportDims, v[0], v[1], v[2]);
        }
    }
}

/* SL_PointProcessor::process_verts + SL_PointRasterizer::render_point (src/SL_PointProcessor.cpp:70-113,
 * src/SL_PointRasterizer.cpp:27-101), one thread: every element is one vertex; a vertex inside the clip volume
 * (sl_ndc_clip_status, one-vertex form, SL_PostVertexTransform.hpp:536-553) becomes one fragment at its screen
 * position: depth test against the buffer, varyings = the vertex's own, shader, colour write / blend, depth write. */
static void process_points(slcu_context* c, const slcu_mesh& m, size_t instanceId, bool usingIndices, const math::mat4& scissorMat,
                           const math::vec4& viewportDims)
{
    const Shader& sh = *c->shader;
    const SL_VertexArray& vao = c->vaos[m.vaoId];
    const IndexBuffer* pIbo = vao.has_index_buffer() ? &c->ibos[vao.mIboId] : nullptr;
    SL_Texture& depthBuffer = *c->depthBuf;

    SL_VertexParam params;
    params.pUniforms = &c->ubos[sh.ubo];
    params.instanceId = instanceId;
    params.pVao = &vao;
    params.pVbo = &c->vbos[vao.mVboId];

    c->stats.trisSubmitted += (m.elementEnd - m.elementBegin) / 3; /* the harness counts elements / 3 for every mode */
    for (size_t i = m.elementBegin; i < m.elementEnd; ++i)
    {
        size_t vertId = i;
        if (usingIndices)
        {
            const unsigned char* p = pIbo->data.data() + i * pIbo->bytesPerId;
            switch (pIbo->type)
            {
                case 0: vertId = *p; break;
                case 1: { uint16_t u; memcpy(&u, p, 2); vertId = u; break; }
                default: { uint32_t u; memcpy(&u, p, 4); vertId = u; break; }
            }
        }
        TransformedVert tv;
        params.vertId = vertId;
        params.pVaryings = tv.varyings;
        tv.vert = scissorMat * sh.vs(params);

        const float w0p = tv.vert.v[3];
        const float w0n = -w0p;
        bool visible = w0p > 0.f;
        for (int k = 0; k < 4; ++k) visible = visible && (tv.vert.v[k] <= w0p) && (tv.vert.v[k] >= w0n);
        if (!visible) continue;
        perspective_divide(tv.vert);
        ndc_to_screen(tv.vert, viewportDims);
        c->stats.trisBinned++;

        SL_FragmentParam fp;
        fp.pUniforms = &c->ubos[sh.ubo];
        fp.coord.x = (uint16_t)cvt_tr(tv.vert.v[0]);
        fp.coord.y = (uint16_t)cvt_tr(tv.vert.v[1]);
        fp.coord.depth = tv.vert.v[2];
        const size_t idx = (size_t)fp.coord.x + (size_t)depthBuffer.mWidth * fp.coord.y;
        const float d = load_depth(depthBuffer, idx);
        if (!depth_cmp_scalar(sh.depthTest, fp.coord.depth, d)) continue;
        for (int s2 = 0; s2 < sh.numVaryings; ++s2) fp.pVaryings[s2] = tv.varyings[s2];
        for (int k = 0; k < 4; ++k) fp.pOutputs[k] = math::vec4{0.f};
        if (idx < (size_t)depthBuffer.mWidth * depthBuffer.mHeight) c->stats.fragsShaded++;
        const bool haveOutputs = sh.fs(fp);
        if (haveOutputs)
        {
            for (int k = sh.numOutputs - 1; k >= 0; --k)
            {
                if (sh.blend != 0) blend_color(*c->colorBufs[k], fp.coord.x, fp.coord.y, fp.pOutputs[k], sh.blend);
                else store_color(*c->colorBufs[k], fp.coord.x, fp.coord.y, fp.pOutputs[k]);
            }
        }
        if (sh.depthMask) store_depth(depthBuffer, idx, fp.coord.depth);
    }
}

/* ---- lines: SL_LineProcessor + SL_LineRasterizer ---------------------------------------------------------------------- */

/* sl_draw_line_bresenham (SL_Geometry.hpp:314-366): sl_lowp_t is int16_t (SL_Setup.hpp:12), arithmetic in int, results
 * truncated to 16 bits at every assignment; the callback takes uint16_t coordinates. */
template <typename Callback>
static void draw_line_bresenham(int16_t x1, int16_t y1, int16_t x2, int16_t y2, Callback&& cb)
{
    const bool steep = std::abs((int)x1 - (int)x2) < std::abs((int)y1 - (int)y2);
    if (steep) { std::swap(x1, y1); std::swap(x2, y2); }
    if (x1 > x2) { std::swap(x1, x2); std::swap(y1, y2); }
    const int16_t dx = (int16_t)(x2 - x1);
    const int16_t dy = (int16_t)(y2 - y1);
    const int16_t dErr = (int16_t)(std::abs((int)dy) * 2);
    const int16_t yErr = (y2 > y1) ? 1 : -1;
    int16_t currentErr = 0;
    int16_t y = y1;
    for (int16_t x = x1; x <= x2; ++x)
    {
        if (steep) cb((uint16_t)y, (uint16_t)x); else cb((uint16_t)x, (uint16_t)y);
        currentErr = (int16_t)(currentErr + dErr);
        if (currentErr > dx)
        {
            y = (int16_t)(y + yErr);
            currentErr = (int16_t)(currentErr - 2 * dx);
        }
        if (x == 32767) break; /* ++x would overflow (undefined in the reference; lines never get here) */
    }
}

/* SL_LineRasterizer::render_line (src/SL_LineRasterizer.cpp:28-87) + flush_line_fragments
 * (src/SL_FragmentProcessor.cpp:346-403), one thread.  The depth test reads the buffer while up to 600 of the line's own
 * fragments are still queued; the queue is restated because a line that touches column `width` and column 0 of the next
 * row visits one texel twice (linear addressing). */
static void render_line(slcu_context* c, const TransformedVert& a, const TransformedVert& b)
{
    const Shader& sh = *c->shader;
    SL_Texture& depthBuffer = *c->depthBuf;
    const float z0 = a.vert.v[2], z1 = b.vert.v[2];
    const float sc0x = a.vert.v[0], sc0y = a.vert.v[1], sc1x = b.vert.v[0], sc1y = b.vert.v[1];
    const float ddx = sc1x - sc0x, ddy = sc1y - sc0y;
    const float dist = math::inversesqrt(ddx * ddx + ddy * ddy); /* length_squared(vec2): sum(v * v), generic */

    struct Queued { float interp; uint16_t x, y; float depth; };
    std::vector<Queued> queue;
    auto flush = [&]()
    {
        for (const Queued& q : queue)
        {
            SL_FragmentParam fp;
            fp.pUniforms = &c->ubos[sh.ubo];
            /* interpolate_line_varyings (src/SL_FragmentProcessor.cpp:27-52, AVX2 branch): all four slots, fmadd(v1 - v0, p, v0) */
            for (int s2 = 0; s2 < 4; ++s2)
            {
                for (int k = 0; k < 4; ++k) fp.pVaryings[s2].v[k] = sfma(b.varyings[s2].v[k] - a.varyings[s2].v[k], q.interp, a.varyings[s2].v[k]);
            }
            fp.coord.x = q.x;
            fp.coord.y = q.y;
            fp.coord.depth = q.depth;
            for (int k = 0; k < 4; ++k) fp.pOutputs[k] = math::vec4{0.f};
            const size_t idx = (size_t)q.x + (size_t)depthBuffer.mWidth * q.y;
            if (idx < (size_t)depthBuffer.mWidth * depthBuffer.mHeight) c->stats.fragsShaded++;
            if (sh.fs(fp))
            {
                for (int k = sh.numOutputs - 1; k >= 0; --k)
                {
                    if (sh.blend != 0) blend_color(*c->colorBufs[k], q.x, q.y, fp.pOutputs[k], sh.blend);
                    else store_color(*c->colorBufs[k], q.x, q.y, fp.pOutputs[k]);
                }
            }
            if (sh.depthMask) store_depth(depthBuffer, idx, q.depth);
        }
        queue.clear();
    };
    draw_line_bresenham((int16_t)(uint16_t)cvt_tr(sc0x), (int16_t)(uint16_t)cvt_tr(sc0y), (int16_t)(uint16_t)cvt_tr(sc1x), (int16_t)(uint16_t)cvt_tr(sc1y),
                        [&](uint16_t x, uint16_t y)
                        {
                            const math::vec4 d{(float)x - sc0x, (float)y - sc0y, 0.f - 0.f, 0.f - 0.f};
                            const float currLen = math::length(d);
                            const float interp = currLen * dist;
                            const float z = z0 + ((z1 - z0) * interp); /* math::mix, scalar (generic/scalar_utils_impl.h:52) */
                            const size_t idx = (size_t)x + (size_t)depthBuffer.mWidth * y;
                            const float dd = load_depth(depthBuffer, idx);
                            if (!depth_cmp_scalar(sh.depthTest, z, dd)) return;
                            queue.push_back(Queued{interp, x, y, z});
                            if (queue.size() == 600) flush(); /* SL_SHADER_MAX_QUEUED_FRAGS */
                        });
    if (!queue.empty()) flush();
}

/* SL_LineProcessor::clip_and_process_lines (src/SL_LineProcessor.cpp:68-177): Liang-Barsky against
 * (viewportDims[0], [1]) .. ([2], [3]) -- the rectangle's width and height are used as its far corner, like the source */
static void clip_and_process_lines(slcu_context* c, const math::vec4& viewportDims, const TransformedVert& a, const TransformedVert& b)
{
    const int numVarys = c->shader->numVaryings;
    const float dx = b.vert.v[0] - a.vert.v[0];
    const float dy = b.vert.v[1] - a.vert.v[1];
    const float xMin = viewportDims.v[0], yMin = viewportDims.v[1], xMax = viewportDims.v[2], yMax = viewportDims.v[3];
    const double eps = 1.0e-9; /* LS_EPSILON is a double literal: the comparisons are done in double */
    if ((double)math::abs(dx) < eps && (double)math::abs(dy) < eps)
    {
        if (a.vert.v[0] >= xMin && a.vert.v[0] <= xMax && a.vert.v[1] >= yMin && a.vert.v[1] <= yMax)
        {
            c->stats.trisBinned++;
            render_line(c, a, b);
            return;
        }
    }
    auto clip_segment = [eps](float num, float denom, float& tE, float& tL) -> bool
    {
        if ((double)math::abs(denom) < eps) return num < 0.f;
        const float t = num / denom;
        if (denom > 0.f)
        {
            if (t > tL) return false;
            if (t > tE) tE = t;
        }
        else
        {
            if (t < tE) return false;
            if (t < tL) tL = t;
        }
        return true;
    };
    TransformedVert p0 = a, p1 = b;
    float& x1 = p0.vert.v[0];
    float& y1 = p0.vert.v[1];
    float& x2 = p1.vert.v[0];
    float& y2 = p1.vert.v[1];
    const float ex = p0.vert.v[0] - p1.vert.v[0], ey = p0.vert.v[1] - p1.vert.v[1];
    const float dist = math::inversesqrt(ex * ex + ey * ey);
    float tE = 0.f, tL = 1.f;
    if (clip_segment(xMin - x1, dx, tE, tL) && clip_segment(x1 - xMax, -dx, tE, tL) && clip_segment(yMin - y1, dy, tE, tL) && clip_segment(y1 - yMax, -dy, tE, tL))
    {
        if (tL < 1.f)
        {
            x2 = x1 + tL * dx;
            y2 = y1 + tL * dy;
        }
        if (tE > 0.f)
        {
            x1 += tE * dx;
            y1 += tE * dy;
        }
        const math::vec4 d0{p0.vert.v[0] - a.vert.v[0], p0.vert.v[1] - a.vert.v[1], p0.vert.v[2] - a.vert.v[2], p0.vert.v[3] - a.vert.v[3]};
        const math::vec4 d1{p1.vert.v[0] - b.vert.v[0], p1.vert.v[1] - b.vert.v[1], p1.vert.v[2] - b.vert.v[2], p1.vert.v[3] - b.vert.v[3]};
        const float len0 = math::length(d0);
        const float len1 = math::length(d1);
        const float interp0 = len0 * dist;
        const float interp1 = len1 * dist;
        p0.vert.v[2] = a.vert.v[2] + ((b.vert.v[2] - a.vert.v[2]) * interp0);
        p0.vert.v[3] = a.vert.v[3] + ((b.vert.v[3] - a.vert.v[3]) * interp0);
        p1.vert.v[2] = b.vert.v[2] + ((a.vert.v[2] - b.vert.v[2]) * interp1);
        p1.vert.v[3] = b.vert.v[3] + ((a.vert.v[3] - b.vert.v[3]) * interp1);
        for (int i = 0; i < numVarys; ++i)
        {
            p0.varyings[i] = math::mix(a.varyings[i], b.varyings[i], interp0);
            p1.varyings[i] = math::mix(b.varyings[i], a.varyings[i], interp1);
        }
        c->stats.trisBinned++;
        render_line(c, p0, p1);
    }
}

/* SL_LineProcessor::process_verts (src/SL_LineProcessor.cpp:188-279), one thread, no vertex cache: elements are taken two
 * at a time; a segment with a vertex behind the eye (w < 0) is dropped, the others are tested with the two-vertex
 * sl_ndc_clip_status (SL_PostVertexTransform.hpp:560-583), divided, mapped to the screen and drawn or clipped. */
static void process_lines(slcu_context* c, const slcu_mesh& m, size_t instanceId, bool usingIndices, const math::mat4& scissorMat,
                          const math::vec4& viewportDims)
{
    const Shader& sh = *c->shader;
    const SL_VertexArray& vao = c->vaos[m.vaoId];
    const IndexBuffer* pIbo = vao.has_index_buffer() ? &c->ibos[vao.mIboId] : nullptr;
    SL_VertexParam params;
    params.pUniforms = &c->ubos[sh.ubo];
    params.instanceId = instanceId;
    params.pVao = &vao;
    params.pVbo = &c->vbos[vao.mVboId];
    auto fetch = [&](size_t i) -> size_t
    {
        if (!usingIndices) return i;
        const unsigned char* p = pIbo->data.data() + i * pIbo->bytesPerId;
        switch (pIbo->type)
        {
            case 0: return *p;
            case 1: { uint16_t u; memcpy(&u, p, 2); return u; }
            default: { uint32_t u; memcpy(&u, p, 4); return u; }
        }
    };
    c->stats.trisSubmitted += (m.elementEnd - m.elementBegin) / 3; /* the harness counts elements / 3 for every mode */
    for (size_t i = m.elementBegin; i < m.elementEnd; i += 2)
    {
        TransformedVert v0, v1;
        for (int k = 0; k < 4; ++k) { v0.varyings[k] = math::vec4{0.f}; v1.varyings[k] = math::vec4{0.f}; }
        params.vertId = fetch(i);
        params.pVaryings = v0.varyings;
        v0.vert = scissorMat * sh.vs(params);
        params.vertId = fetch(i + 1);
        params.pVaryings = v1.varyings;
        v1.vert = scissorMat * sh.vs(params);
        if (v0.vert.v[3] < 0.f || v1.vert.v[3] < 0.f) continue;
        const float w0p = v0.vert.v[3], w1p = v1.vert.v[3];
        bool vis = true;
        for (int k = 0; k < 4; ++k) vis = vis && (v0.vert.v[k] <= w0p) && (v1.vert.v[k] <= w1p) && (v0.vert.v[k] >= -w0p) && (v1.vert.v[k] >= -w1p);
        const bool part = (w0p > 0.f) || (w1p > 0.f);
        if (!vis && !part) continue;
        perspective_divide(v0.vert);
        perspective_divide(v1.vert);
        ndc_to_screen(v0.vert, viewportDims);
        ndc_to_screen(v1.vert, viewportDims);
        if (vis)
        {
            c->stats.trisBinned++;
            render_line(c, v0, v1);
        }
        else clip_and_process_lines(c, viewportDims, v0, v1);
    }
}

static void apply_fixups(slcu_context* c)
{
    for (const UboFixup& f : c->fixups)
    {
        const void* ptr;
        if (f.kind == 0) ptr = c->textures[f.id];
        else ptr = c->vbos[f.id].mBuffer.data() + f.byteOffset;
        memcpy(c->ubos[f.ubo].mBytes + f.offset, &ptr, sizeof(ptr));
    }
}

/* SL_ProcessorPool::run_shader_processors + SL_TriProcessor::execute
 * (src/SL_ProcessorPool.cpp:280-380, src/SL_TriProcessor.cpp:665-716) */
static int run_draw(slcu_context* c, const slcu_mesh* meshes, size_t numMeshes, size_t numInstances, size_t shaderId, size_t fboId)
{
    if (shaderId >= c->shaders.size() || fboId >= c->fbos.size()) return SLCU_ERR_ARG;
    memcpy(g_rcp_table, c->rcpTable, sizeof(g_rcp_table));
    memcpy(g_rsqrt_table, c->rsqrtTable, sizeof(g_rsqrt_table));
    apply_fixups(c);
    const Framebuffer& fbo = c->fbos[fboId];
    c->shader = &c->shaders[shaderId];
    for (unsigned i = 0; i < 4; ++i) c->colorBufs[i] = (i < fbo.numColors && fbo.colors[i] != (size_t)-1) ? c->textures[fbo.colors[i]] : nullptr;
    if (fbo.depth == (size_t)-1) return SLCU_ERR_ARG;
    c->depthBuf = c->textures[fbo.depth];
    c->stats.drawCalls++;

    const float fboW = (float)c->depthBuf->mWidth, fboH = (float)c->depthBuf->mHeight;
    const math::mat4 scissorMat = scissor_matrix(c->scissor, fboW, fboH);
    const math::vec4 viewportDims = viewport_rect(c->viewport, c->scissor, fboW, fboH);

    c->wire = (meshes[0].mode == 0x50) || (meshes[0].mode == 0x70); /* mRenderMode = meshes->mode (src/SL_ProcessorPool.cpp:301,354) */
    const bool points = (meshes[0].mode == 0x01) || (meshes[0].mode == 0x03); /* RENDER_MODE_POINTS / INDEXED_POINTS -> SL_PointProcessor */
    const bool lines = (meshes[0].mode == 0x04) || (meshes[0].mode == 0x0c); /* RENDER_MODE_LINES / INDEXED_LINES -> SL_LineProcessor */
    if (lines)
    {
        if (numInstances == 1)
        {
            for (size_t i = 0; i < numMeshes; ++i) process_lines(c, meshes[i], 0, meshes[i].mode == 0x0c, scissorMat, viewportDims);
        }
        else
        {
            for (size_t i = 0; i < numInstances; ++i) process_lines(c, meshes[0], i, meshes[0].mode == 0x0c, scissorMat, viewportDims);
        }
        return 0;
    }
    if (points)
    {
        if (numInstances == 1)
        {
            for (size_t i = 0; i < numMeshes; ++i) process_points(c, meshes[i], 0, meshes[i].mode == 0x03, scissorMat, viewportDims);
        }
        else
        {
            for (size_t i = 0; i < numInstances; ++i) process_points(c, meshes[0], i, meshes[0].mode == 0x03, scissorMat, viewportDims);
        }
        return 0;
    }
    if (numInstances == 1)
    {
        for (size_t i = 0; i < numMeshes; ++i)
        {
            const bool usingIndices = (meshes[i].mode == 0x30) || (meshes[i].mode == 0x70);
            process_verts(c, meshes[i], 0, usingIndices, scissorMat, viewportDims);
        }
    }
    else
    {
        const bool usingIndices = (meshes[0].mode == 0x30) || (meshes[0].mode == 0x70);
        for (size_t i = 0; i < numInstances; ++i) process_verts(c, meshes[0], i, usingIndices, scissorMat, viewportDims);
    }
    if (!c->bins.empty()) flush_bins(c); /* cleanup<SL_TriRasterizer>() */
    return 0;
}

/* SL_ClearProcessor::clear_texture (src/SL_ClearProcessor.cpp:23-63): w*h texels (depth dim ignored) */
static void clear_texture(SL_Texture& t, const void* texel)
{
    /* sl_calc_indexed_parition2<1>(count, 1, 0) (SL_ShaderUtil.hpp:222-231): the single thread's range is
     * count + (1 - count % 3) clipped to count, i.e. the LAST texel stays untouched when count % 3 == 2 */
    const size_t count = (size_t)t.mWidth * t.mHeight;
    const size_t range = count + (1 - (count % 3));
    const size_t n = range < count ? range : count;
    for (size_t i = 0; i < n; ++i) memcpy(t.mTexels + i * t.mBpt, texel, t.mBpt);
}

/* sl_match_color_for_type (src/SL_Color.cpp:135-190): vecN<double> -> typed texel */
static void match_color(int type, const double rgba[4], unsigned char out[32])
{
    if (is_packed(type)) /* rgba_cast<packed, double>: (value_type)(c * (double)max(field)), truncated into the field */
    {
        uint32_t f[4];
        for (int j = 0; j < 4; ++j) f[j] = (uint32_t)(type == 32 ? (uint16_t)(rgba[j] * (double)pack_max(4, j)) : (uint8_t)(rgba[j] * (double)pack_max(type - 28, j)));
        store_packed(out, type - 28, pack(type - 28, f));
        return;
    }
    const int kind = type % 7;
    const unsigned n = (unsigned)(type / 7 + 1);
    for (unsigned i = 0; i < n; ++i)
    {
        switch (kind)
        {
            case 0: { const uint8_t v = (uint8_t)(rgba[i] * 255.0); out[i] = v; break; }          /* generic (T)(double) truncation */
            case 1: { const uint16_t v = (uint16_t)(rgba[i] * 65535.0); memcpy(out + 2 * i, &v, 2); break; }
            case 2: { const uint32_t v = (uint32_t)(rgba[i] * 4294967295.0); memcpy(out + 4 * i, &v, 4); break; }
            case 3: { const uint64_t v = (uint64_t)(rgba[i] * 18446744073709551615.0); memcpy(out + 8 * i, &v, 8); break; }
            case 4: { const uint16_t v = float_to_half_rn((float)rgba[i]); memcpy(out + 2 * i, &v, 2); break; }
            case 5: { const float v = (float)rgba[i]; memcpy(out + 4 * i, &v, 4); break; }
            default: { memcpy(out + 8 * i, &rgba[i], 8); break; }
        }
    }
}

static bool depth_texel(const SL_Texture& t, double depth, unsigned char out[8])
{
    /* SL_Context::clear_depth_buffer (src/SL_Context.cpp:896-925) */
    switch (t.mBpt)
    {
        case 2: { const uint16_t h = float_to_half_rn((float)depth); memcpy(out, &h, 2); return true; }
        case 4: { const float f = (float)depth; memcpy(out, &f, 4); return true; }
        case 8: memcpy(out, &depth, 8); return true;
        default: return false;
    }
}

} // namespace slo

/* ===========================================================================
 * C ABI
 * =========================================================================*/
extern "C" {

int slor_context_create(slcu_context** out, int)
{
    if (!out) return SLCU_ERR_ARG;
    slcu_context* c = new slcu_context{};
    measure_host_rcp_table(c->rcpTable);
    measure_host_rsqrt_table(c->rsqrtTable);
    *out = c;
    return 0;
}
void slor_context_destroy(slcu_context* c) { delete c; }
unsigned slor_num_threads(slcu_context*, unsigned n) { return n < 1 ? 1 : n; }
const char* slor_last_error(slcu_context* c) { return c->err.c_str(); }
const char* slor_impl_name(void) { return "oracle-port(scalar C++ restatement)"; }
int slor_set_rcp_table(slcu_context* c, const uint32_t* t) { memcpy(c->rcpTable, t, sizeof(c->rcpTable)); return 0; }
int slor_get_rcp_table(slcu_context* c, uint32_t* t) { memcpy(t, c->rcpTable, sizeof(c->rcpTable)); return 0; }
int slor_set_rsqrt_table(slcu_context* c, const uint32_t* t) { memcpy(c->rsqrtTable, t, sizeof(c->rsqrtTable)); return 0; }
int slor_get_rsqrt_table(slcu_context* c, uint32_t* t) { memcpy(t, c->rsqrtTable, sizeof(c->rsqrtTable)); return 0; }

size_t slor_create_vbo(slcu_context* c) { c->vbos.emplace_back(); return c->vbos.size() - 1; }
size_t slor_create_ibo(slcu_context* c) { c->ibos.emplace_back(); return c->ibos.size() - 1; }
size_t slor_create_vao(slcu_context* c) { c->vaos.emplace_back(); return c->vaos.size() - 1; }
size_t slor_create_ubo(slcu_context* c) { c->ubos.emplace_back(); memset(c->ubos.back().mBytes, 0, 1024); return c->ubos.size() - 1; }
size_t slor_create_texture(slcu_context* c) { c->textures.push_back(new SL_Texture{}); return c->textures.size() - 1; }
size_t slor_create_framebuffer(slcu_context* c) { c->fbos.emplace_back(); return c->fbos.size() - 1; }

/* SL_Context::destroy_* (src/SL_Context.cpp:192-195, 243-247, 295-298, 346-349, 397-400, 448-451, 612-615): vector erase.
 * Framebuffer attachments and the pointers declared with ubo_bind_* are pointers in the reference and keep following
 * their object; ids held by vertex arrays and shaders are indices there and are left alone. */
static void remap_fixups(slcu_context* c, int kind, size_t gone)
{
    for (size_t i = 0; i < c->fixups.size();)
    {
        UboFixup& f = c->fixups[i];
        if (f.kind == kind && f.id == gone) { c->fixups.erase(c->fixups.begin() + i); continue; }
        if (f.kind == kind && f.id > gone) --f.id;
        ++i;
    }
}
int slor_destroy_vbo(slcu_context* c, size_t id)
{
    if (id >= c->vbos.size()) return SLCU_ERR_ARG;
    c->vbos.erase(c->vbos.begin() + id);
    remap_fixups(c, 1, id);
    return 0;
}
int slor_destroy_ibo(slcu_context* c, size_t id) { if (id >= c->ibos.size()) return SLCU_ERR_ARG; c->ibos.erase(c->ibos.begin() + id); return 0; }
int slor_destroy_vao(slcu_context* c, size_t id) { if (id >= c->vaos.size()) return SLCU_ERR_ARG; c->vaos.erase(c->vaos.begin() + id); return 0; }
int slor_destroy_ubo(slcu_context* c, size_t id)
{
    if (id >= c->ubos.size()) return SLCU_ERR_ARG;
    c->ubos.erase(c->ubos.begin() + id);
    for (size_t i = 0; i < c->fixups.size();)
    {
        UboFixup& f = c->fixups[i];
        if (f.ubo == id) { c->fixups.erase(c->fixups.begin() + i); continue; }
        if (f.ubo > id) --f.ubo;
        ++i;
    }
    return 0;
}
int slor_destroy_texture(slcu_context* c, size_t id)
{
    if (id >= c->textures.size()) return SLCU_ERR_ARG;
    if (c->textures[id]) { free(c->textures[id]->mTexels); delete c->textures[id]; }
    c->textures.erase(c->textures.begin() + id);
    remap_fixups(c, 0, id);
    for (Framebuffer& f : c->fbos)
    {
        for (int i = 0; i < 4; ++i)
        {
            if (f.colors[i] == (size_t)-1) continue;
            if (f.colors[i] == id) { f.colors[i] = (size_t)-1; f.colorViews[i] = FboView{}; }
            else if (f.colors[i] > id) --f.colors[i];
        }
        if (f.depth != (size_t)-1)
        {
            if (f.depth == id) { f.depth = (size_t)-1; f.depthView = FboView{}; }
            else if (f.depth > id) --f.depth;
        }
    }
    return 0;
}
int slor_destroy_framebuffer(slcu_context* c, size_t id) { if (id >= c->fbos.size()) return SLCU_ERR_ARG; c->fbos.erase(c->fbos.begin() + id); return 0; }
int slor_destroy_shader(slcu_context* c, size_t id) { if (id >= c->shaders.size()) return SLCU_ERR_ARG; c->shaders.erase(c->shaders.begin() + id); return 0; }

int slor_vbo_init(slcu_context* c, size_t vbo, size_t n)
{
    if (vbo >= c->vbos.size()) return SLCU_ERR_ARG;
    c->vbos[vbo].mBuffer.assign(n + 64, 0);
    return 0;
}
int slor_vbo_assign(slcu_context* c, size_t vbo, const void* d, size_t off, size_t n)
{
    if (vbo >= c->vbos.size() || off + n > c->vbos[vbo].mBuffer.size()) return SLCU_ERR_ARG;
    memcpy(c->vbos[vbo].mBuffer.data() + off, d, n);
    return 0;
}
int slor_ibo_init(slcu_context* c, size_t ibo, uint32_t n, int type, const void* d)
{
    if (ibo >= c->ibos.size() || type < 0 || type > 2) return SLCU_ERR_ARG;
    IndexBuffer& b = c->ibos[ibo];
    b.type = type;
    b.count = n;
    b.bytesPerId = 1u << type;
    b.data.assign((size_t)n * b.bytesPerId + 16, 0);
    if (d) memcpy(b.data.data(), d, (size_t)n * b.bytesPerId);
    return 0;
}
int slor_vao_set_vertex_buffer(slcu_context* c, size_t vao, size_t vbo) { c->vaos[vao].mVboId = vbo; return 0; }
int slor_vao_set_index_buffer(slcu_context* c, size_t vao, size_t ibo) { c->vaos[vao].mIboId = ibo; return 0; }
int slor_vao_set_num_bindings(slcu_context* c, size_t vao, int n)
{
    if (n >= 8 || n < 0) return -1; /* src/SL_VertexArray.cpp:162-167 */
    c->vaos[vao].mNumBindings = n;
    return n;
}
int slor_vao_set_binding(slcu_context* c, size_t vao, size_t bind, ptrdiff_t off, ptrdiff_t stride, int dim, int type)
{
    if (bind >= 8) return SLCU_ERR_ARG;
    c->vaos[vao].mBindings[bind] = SL_VertexArray::Binding{dim, type, off, stride};
    return 0;
}
int slor_ubo_assign(slcu_context* c, size_t ubo, const void* d, size_t off, size_t n)
{
    if (ubo >= c->ubos.size() || off + n > 1024) return SLCU_ERR_ARG;
    memcpy(c->ubos[ubo].mBytes + off, d, n);
    return 0;
}
static void set_fixup(slcu_context* c, const UboFixup& f)
{
    for (UboFixup& g : c->fixups) { if (g.ubo == f.ubo && g.offset == f.offset) { g = f; return; } }
    c->fixups.push_back(f);
}
int slor_ubo_bind_texture(slcu_context* c, size_t ubo, size_t off, size_t tex) { set_fixup(c, UboFixup{ubo, off, 0, tex, 0}); return 0; }
int slor_ubo_bind_buffer(slcu_context* c, size_t ubo, size_t off, size_t vbo, size_t bo) { set_fixup(c, UboFixup{ubo, off, 1, vbo, bo}); return 0; }

int slor_texture_init(slcu_context* c, size_t tex, int type, uint16_t w, uint16_t h, uint16_t d)
{
    if (tex >= c->textures.size() || type < 0 || type >= 33) return SLCU_ERR_ARG;
    SL_Texture& t = *c->textures[tex];
    free(t.mTexels);
    t.mWidth = w; t.mHeight = h; t.mDepth = d;
    if (type >= 28) /* the packed types (src/SL_Color.cpp:10-110) */
    {
        t.mChannels = (uint8_t)(type <= 29 ? 3 : 4);
        t.mBpt = (uint8_t)slo::kPackBytes[type - 28];
    }
    else
    {
        t.mChannels = (uint8_t)(type / 7 + 1);
        t.mBpt = (uint8_t)(chan_bytes(type) * t.mChannels);
    }
    t.mType = type;
    /* _sl_allocate_texture (src/SL_Texture.cpp:28-54): every dimension padded up by 4-(dim%4) */
    const size_t pw = w + (4 - (w % 4)), ph = h + (4 - (h % 4)), pd = d + (4 - (d % 4));
    t.mAllocBytes = pw * ph * pd * t.mBpt + 4 * t.mBpt;
    t.mTexels = (unsigned char*)calloc(t.mAllocBytes, 1);
    return t.mTexels ? 0 : -1;
}
int slor_texture_set_texels(slcu_context* c, size_t tex, int order, uint16_t x, uint16_t y, uint16_t z, uint16_t w, uint16_t h, uint16_t d, const void* data)
{
    if (tex >= c->textures.size()) return SLCU_ERR_ARG;
    SL_Texture& t = *c->textures[tex];
    const unsigned char* src = (const unsigned char*)data;
    for (uint16_t k = 0, z0 = z; k < d; ++k, ++z0)
        for (uint16_t j = 0, y0 = y; j < h; ++j, ++y0)
            for (uint16_t i = 0, x0 = x; i < w; ++i, ++x0)
            {
                const ptrdiff_t index = i + w * (j + (h * k));
                /* set_texel<order> (SL_Texture.hpp:388-410): 2-D map when depth <= 1, else the 3-D map */
                ptrdiff_t dst;
                if (1 >= t.mDepth) dst = order == 0 ? t.map_coordinate<SL_TexelOrder::ORDERED>(x0, y0) : t.map_coordinate<SL_TexelOrder::SWIZZLED>(x0, y0);
                else dst = order == 0 ? t.map_coordinate<SL_TexelOrder::ORDERED>(x0, y0, z0) : t.map_coordinate<SL_TexelOrder::SWIZZLED>(x0, y0, z0);
                memcpy(t.mTexels + dst * t.mBpt, src + index * t.mBpt, t.mBpt);
            }
    return 0;
}
int slor_texture_read(slcu_context* c, size_t tex, void* out, size_t n)
{
    if (tex >= c->textures.size() || n > c->textures[tex]->mAllocBytes) return SLCU_ERR_ARG;
    memcpy(out, c->textures[tex]->mTexels, n);
    return 0;
}

int slor_fbo_reserve_color_buffers(slcu_context* c, size_t fbo, unsigned n)
{
    /* src/SL_Framebuffer.cpp:823-848: same count = no-op (checked first), the colour views are reset, the depth view stays */
    Framebuffer& f = c->fbos[fbo];
    if (n == f.numColors) return 0;
    if (n < 1) return -1;
    if (n > 4) return -2;
    for (int i = 0; i < 4; ++i) { f.colors[i] = (size_t)-1; f.colorViews[i] = FboView{}; }
    f.numColors = n;
    return 0;
}
static FboView make_view(const SL_Texture& t) { return FboView{t.mTexels != nullptr, t.mWidth, t.mHeight, t.mDepth, t.mType}; }
int slor_fbo_attach_color_buffer(slcu_context* c, size_t fbo, unsigned i, size_t tex)
{
    if (i >= c->fbos[fbo].numColors) return -1;
    c->fbos[fbo].colors[i] = tex;
    c->fbos[fbo].colorViews[i] = make_view(*c->textures[tex]);
    return 0;
}
int slor_fbo_attach_depth_buffer(slcu_context* c, size_t fbo, size_t tex)
{
    c->fbos[fbo].depth = tex;
    c->fbos[fbo].depthView = make_view(*c->textures[tex]);
    return 0;
}
/* SL_Framebuffer::valid (src/SL_Framebuffer.cpp:928-990): same checks, same order, same codes */
int slor_fbo_valid(slcu_context* c, size_t fbo)
{
    const Framebuffer& f = c->fbos[fbo];
    if (!f.numColors) return -1;
    const FboView& c0 = f.colorViews[0];
    for (unsigned i = 0; i < f.numColors; ++i)
    {
        const FboView& v = f.colorViews[i];
        if (!v.texels) return -2;
        if (v.w != c0.w) return -3;
        if (v.h != c0.h) return -4;
        if (v.d != c0.d) return -5;
    }
    const FboView& d = f.depthView;
    if (!d.texels) return -6;
    if (d.w != c0.w) return -7;
    if (d.h != c0.h) return -8;
    if (d.d > 1) return -9;
    if (d.type != 4 && d.type != 5 && d.type != 6) return -10;
    return 0;
}

size_t slor_create_shader(slcu_context* c, const slcu_vertex_shader* vs, const slcu_fragment_shader* fs, size_t ubo)
{
    /* SL_Context::create_shader validation (src/SL_Context.cpp:488-540) */
    if (ubo >= c->ubos.size() || vs->numVaryings > 4 || fs->numVaryings > vs->numVaryings || fs->numOutputs > 4 || !fs->numOutputs) return SLCU_INVALID_ID;
    auto v = vs_registry().find(vs->shader);
    auto f = fs_registry().find(fs->shader);
    if (v == vs_registry().end() || f == fs_registry().end()) { c->err = "unknown shader"; return SLCU_INVALID_ID; }
    Shader s;
    s.cullMode = vs->cullMode; s.depthTest = fs->depthTest; s.depthMask = fs->depthMask; s.blend = fs->blend;
    s.numVaryings = fs->numVaryings; s.numOutputs = fs->numOutputs; s.vs = v->second; s.fs = f->second; s.ubo = ubo;
    c->shaders.push_back(s);
    return c->shaders.size() - 1;
}

int slor_viewport(slcu_context* c, int32_t x, int32_t y, uint16_t w, uint16_t h) { c->viewport[0] = x; c->viewport[1] = y; c->viewport[2] = w; c->viewport[3] = h; return 0; }
int slor_scissor(slcu_context* c, int32_t x, int32_t y, uint16_t w, uint16_t h) { c->scissor[0] = x; c->scissor[1] = y; c->scissor[2] = w; c->scissor[3] = h; return 0; }

int slor_clear_color_buffer(slcu_context* c, size_t fbo, unsigned att, const double rgba[4])
{
    SL_Texture& t = *c->textures[c->fbos[fbo].colors[att]];
    unsigned char texel[32];
    match_color(t.mType, rgba, texel);
    clear_texture(t, texel);
    return 0;
}
int slor_clear_depth_buffer(slcu_context* c, size_t fbo, double depth)
{
    SL_Texture& t = *c->textures[c->fbos[fbo].depth];
    unsigned char texel[8];
    if (depth_texel(t, depth, texel)) clear_texture(t, texel);
    return 0;
}
int slor_clear_framebuffer(slcu_context* c, size_t fbo, unsigned att, const double rgba[4], double depth)
{
    SL_Texture& d = *c->textures[c->fbos[fbo].depth];
    unsigned char dt[8];
    if (!depth_texel(d, depth, dt)) return 0; /* reference returns before clearing anything */
    slor_clear_color_buffer(c, fbo, att, rgba);
    clear_texture(d, dt);
    return 0;
}

int slor_draw(slcu_context* c, const slcu_mesh* m, size_t shader, size_t fbo) { return run_draw(c, m, 1, 1, shader, fbo); }
int slor_draw_multiple(slcu_context* c, const slcu_mesh* m, size_t n, size_t shader, size_t fbo)
{
    if (!m || !n) return 0;
    return run_draw(c, m, n, 1, shader, fbo);
}
int slor_draw_instanced(slcu_context* c, const slcu_mesh* m, size_t inst, size_t shader, size_t fbo) { return run_draw(c, m, 1, inst, shader, fbo); }

/* ---- SL_BlitProcessor's texel conversions (src/SL_BlitProcessor.cpp:20-426 + color_cast, SL_Color.hpp:236-700) ----------
 * A blit functor builds a colour of the DESTINATION's channel count in the SOURCE's component type (missing channels are
 * the type's minimum, alpha its maximum; R lands in the LAST colour channel of RGB/RGBA, RG in the middle of RGBA --
 * :196-204, 304-326), converts it with color_cast and stores it.  "to R" functors convert the whole source colour and keep
 * component 0, so their conversion runs on the source's channel count.  Conversions of 4-component colours go through
 * vec4_t's SSE cast operators where one exists (x86/vec4f_impl.h:190-355: float -> u8/u16/u32 = cvtps2dq + saturating packs,
 * float -> half = cvtps2ph toward zero, u32 -> float = the halves trick); everything else is the element-wise C cast. */
namespace slo
{
struct TexelValue /* one component, in its own type */
{
    uint64_t u; /* kinds 0..3 */
    uint16_t h; /* kind 4: half bits */
    float f;    /* kind 5 */
    double d;   /* kind 6 */
};
static const uint64_t kIntMax[4] = {0xffull, 0xffffull, 0xffffffffull, 0xffffffffffffffffull};

static inline TexelValue load_component(const unsigned char* p, int kind)
{
    TexelValue t{};
    switch (kind)
    {
        case 0: t.u = *p; break;
        case 1: { uint16_t v; memcpy(&v, p, 2); t.u = v; break; }
        case 2: { uint32_t v; memcpy(&v, p, 4); t.u = v; break; }
        case 3: memcpy(&t.u, p, 8); break;
        case 4: memcpy(&t.h, p, 2); break;
        case 5: memcpy(&t.f, p, 4); break;
        default: memcpy(&t.d, p, 8); break;
    }
    return t;
}
static inline TexelValue limit_component(int kind, bool maxValue)
{
    TexelValue t{};
    if (!maxValue) return t; /* min(): 0 for every type (SL_ColorLimits, SL_Color.hpp:90-175) */
    if (kind <= 3) t.u = kIntMax[kind];
    else if (kind == 4) t.h = 0x3c00u;
    else if (kind == 5) t.f = 1.f;
    else t.d = 1.0;
    return t;
}
/* (T)x for an integer T, as the x86 code does it: cvttss2si / cvttsd2si (the 64-bit form for 32- and 64-bit unsigned
 * targets), then the low bits */
static inline uint64_t trunc_to_int(double x, bool isFloat, int kind)
{
    if (kind <= 1)
    {
        const bool in = isFloat ? ((float)x >= -2147483648.0f && (float)x < 2147483648.0f) : (x > -2147483649.0 && x < 2147483648.0);
        const int32_t v = in ? (int32_t)x : INT32_MIN;
        return (uint64_t)(uint32_t)v & kIntMax[kind];
    }
    if (kind == 2)
    {
        const bool in = x >= -9223372036854775808.0 && x < 9223372036854775808.0;
        const int64_t v = in ? (int64_t)x : INT64_MIN;
        return (uint64_t)v & 0xffffffffull;
    }
    /* uint64_t: gcc's sequence (compare with 2^63, subtract, convert, flip the top bit) */
    if (x >= 9223372036854775808.0)
    {
        const double y = x - 9223372036854775808.0;
        const bool in = y < 9223372036854775808.0;
        return (uint64_t)(in ? (int64_t)y : INT64_MIN) ^ 0x8000000000000000ull;
    }
    const bool in = x >= -9223372036854775808.0;
    return (uint64_t)(in ? (int64_t)x : INT64_MIN);
}
static inline float u32_to_float_halves(uint32_t v) /* vec4_t<uint32_t> -> vec4_t<float> (vec4f_impl.h:236-245) */
{
    const float hi = (float)(int32_t)(v >> 1), lo = (float)(int32_t)(v & 1u);
    return (hi + hi) + lo;
}
static inline float int_to_float(uint64_t v, int kind, bool simd4)
{
    if (kind == 2 && simd4) return u32_to_float_halves((uint32_t)v);
    if (kind == 3) return (float)v;
    return (float)(uint32_t)v;
}
static inline uint64_t float_to_int(float x, int kind, bool simd4)
{
    if (simd4 && kind <= 2)
    {
        const int32_t r = cvt_rn(x); /* cvtps2dq */
        if (kind == 0) { const int32_t s = r < -32768 ? -32768 : (r > 32767 ? 32767 : r); return (uint64_t)(s < 0 ? 0 : (s > 255 ? 255 : s)); } /* packs_epi32, packus_epi16 */
        if (kind == 1) return (uint64_t)(r < 0 ? 0 : (r > 65535 ? 65535 : r));                                                                /* packus_epi32 */
        return (uint64_t)(uint32_t)r;
    }
    return trunc_to_int((double)x, true, kind);
}

static inline TexelValue convert_component(const TexelValue& in, int ik, int ok, int width)
{
    const bool simd4 = width == 4;
    TexelValue o{};
    if (ik <= 3 && ok <= 3) /* integer -> integer (SL_Color.hpp:236-246) */
    {
        const uint64_t tMax = kIntMax[ok], uMax = kIntMax[ik];
        if (width == 2)
        {
            /* the two-component overload scales by max(T) itself instead of by the ratio of the two maxima (:373-380):
             * widening multiplies by max(T) and wraps, narrowing (and same-type!) divides by max(T) */
            if (tMax > uMax) o.u = ((in.u & tMax) * tMax) & tMax;
            else o.u = (in.u / tMax) & tMax;
            return o;
        }
        if (tMax > uMax) o.u = ((in.u & tMax) * (tMax / (uMax & tMax))) & tMax;
        else o.u = (in.u / (uMax / (tMax & uMax))) & tMax;
        return o;
    }
    if (ok <= 3) /* float kinds -> integer (:250-266): p * (U)max(T), then the cast */
    {
        if (ik == 6) { o.u = trunc_to_int(in.d * (double)kIntMax[ok], false, ok); return o; }
        const float x = (ik == 4 ? half_to_float(in.h) : in.f) * (float)kIntMax[ok];
        o.u = float_to_int(x, ok, simd4);
        return o;
    }
    if (ik <= 3) /* integer -> float kinds (:270-300) */
    {
        if (ok == 6) { o.d = (double)in.u * (1.0 / (double)kIntMax[ik]); return o; }
        const float scale = (ik == 0 && ok == 5) ? 0.00392156862745f : (1.f / (float)kIntMax[ik]);
        const float f = int_to_float(in.u, ik, simd4) * scale;
        if (ok == 5) o.f = f;
        else o.h = simd4 ? float_to_half_rz(f) : float_to_half_rn(f);
        return o;
    }
    /* float kinds -> float kinds (:304-310): plain casts; half only talks to float */
    if (ik == ok) return in;
    if (ik == 4)
    {
        const float f = half_to_float(in.h);
        if (ok == 5) o.f = f; else o.d = (double)f;
        return o;
    }
    if (ik == 5)
    {
        if (ok == 6) o.d = (double)in.f;
        else o.h = simd4 ? float_to_half_rz(in.f) : float_to_half_rn(in.f);
        return o;
    }
    if (ok == 5) o.f = (float)in.d;
    else o.h = float_to_half_rn((float)in.d); /* vec4_t<double> has no SSE cast: (half)(float)(double) */
    return o;
}
static inline void store_component(unsigned char* p, int kind, const TexelValue& t)
{
    switch (kind)
    {
        case 0: *p = (unsigned char)t.u; break;
        case 1: { const uint16_t v = (uint16_t)t.u; memcpy(p, &v, 2); break; }
        case 2: { const uint32_t v = (uint32_t)t.u; memcpy(p, &v, 4); break; }
        case 3: memcpy(p, &t.u, 8); break;
        case 4: memcpy(p, &t.h, 2); break;
        case 5: memcpy(p, &t.f, 4); break;
        default: memcpy(p, &t.d, 8); break;
    }
}
/* which source component (or limit) feeds destination component j: 0..3 = source component, -1 = min, -2 = max */
static inline int blit_channel_source(int srcN, int dstN, int j)
{
    if (dstN <= srcN) return j;                                 /* leading components (vec2_cast / vec3_cast, [0]) */
    if (srcN == 1) return dstN == 2 ? (j == 0 ? 0 : -1) : (j == 2 ? 0 : (j == 3 ? -2 : -1)); /* R -> RG (r, min); -> RGB (min, min, r); -> RGBA (min, min, r, max) */
    if (srcN == 2) return dstN == 3 ? (j < 2 ? j : -1) : (j == 0 ? -1 : (j == 3 ? -2 : j - 1)); /* RG -> RGB (r, g, min); -> RGBA (min, r, g, max) */
    return j < 3 ? j : -2;                                      /* RGB -> RGBA (r, g, b, max) */
}
static void blit_texel(const unsigned char* src, int srcType, unsigned char* dst, int dstType)
{
    static const int kBytes[7] = {1, 2, 4, 8, 2, 4, 8};
    const int ik = srcType % 7, srcN = srcType / 7 + 1, ok = dstType % 7, dstN = dstType / 7 + 1;
    const int width = dstN == 1 ? srcN : dstN; /* the channel count color_cast runs on */
    for (int j = 0; j < dstN; ++j)
    {
        const int from = blit_channel_source(srcN, dstN, j);
        const TexelValue in = from >= 0 ? load_component(src + from * kBytes[ik], ik) : limit_component(ik, from == -2);
        TexelValue out = convert_component(in, ik, ok, width);
        /* the constant alpha of a double colour, 1.0 * (double)UINT64_MAX = 2^64, is converted by the compiler, which
         * saturates; the same product at run time (a float colour's alpha) goes through the cvttss2si sequence and gives 0 */
        if (from == -2 && ik == 6 && ok == 3) out.u = kIntMax[3];
        store_component(dst + j * kBytes[ok], ok, out);
    }
}

/* ---- the packed colour types ("compressed", include/softlight/SL_ColorCompressed.hpp): RGB332, RGB565, RGBA5551,
 * RGBA4444, RGBA1010102 -- bit-fields with r in the lowest bits.  Conversions are rgb_cast / rgba_cast; which of the two a
 * blit uses is fixed by the functor (src/SL_BlitCompressedProcessor.cpp:20-258):
 *   packed -> R, RGB: rgb_cast (component 0 for R);   packed -> RG, RGBA: rgba_cast (leading two for RG);
 *   R/RG/RGB/RGBA -> packed: ALWAYS rgb_cast of (r, min, min) / (r, g, min) / (r, g, b) / (r, g, b): a source alpha is
 *   dropped, the packed alpha becomes its maximum;
 *   packed -> another packed: through uint8 (uint16 towards 1010102), rgb_cast both ways towards the 3-component
 *   types, rgba_cast both ways towards the 4-component ones.
 * 1010102 has its own arithmetic (:745-775, 830-870): the integer rgb_cast multiplies by max(T) in T's own (promoted)
 * width and divides by (T)1023 -- for uint8 that is the LOW 8 bits, for the 32/64-bit types the product wraps; the integer
 * rgba_cast multiplies by (uint32_t)max(T) in 32 bits; and the AVX2 specialisations for float and uint8 RGBA sign-extend
 * the fields (abs() of it for float: 1023 reads as 1/1023; the top eight bits for uint8). */
/* rgb_cast / rgba_cast<packed, U>: integers c / (max(U) / (U)max(field)), float kinds (value_type)(c * (U)max(field)) */
static uint32_t packed_from_plain(const TexelValue* c, int ik, int ci, bool withAlpha)
{
    uint32_t out[4] = {0, 0, 0, pack_max(ci, 3)};
    const int elemKind = ci == 4 ? 1 : 0; /* the casts go to the packed colour's value_type: uint8_t, uint16_t for 1010102 */
    for (int j = 0; j < (withAlpha ? 4 : 3); ++j)
    {
        const uint32_t cm = pack_max(ci, j);
        if (!cm) continue;
        if (ik <= 3)
        {
            const uint64_t uMax = kIntMax[ik];
            out[j] = (uint32_t)((c[j].u / (uMax / ((uint64_t)cm & uMax))) & kIntMax[elemKind]);
        }
        else if (ik == 6) out[j] = (uint32_t)trunc_to_int(c[j].d * (double)cm, false, elemKind);
        else if (ik == 5) out[j] = (uint32_t)trunc_to_int((double)(c[j].f * (float)cm), true, elemKind);
        else
        {
            const uint16_t h = float_to_half_rn(half_to_float(c[j].h) * half_to_float(float_to_half_rn((float)cm))); /* half * half rounds to half */
            out[j] = (uint32_t)trunc_to_int((double)half_to_float(h), true, elemKind);
        }
    }
    return pack(ci, out);
}
/* rgba_cast<SL_ColorRGB1010102, float>, the AVX2 specialisation (:745-762): cvtps2dq, shift, mask */
static uint32_t packed1010102_from_float(const math::vec4& c)
{
    static const float kMax[4] = {1023.f, 1023.f, 1023.f, 3.f};
    static const int kShift[4] = {0, 10, 20, 30};
    static const uint32_t kMask[4] = {0x000003ffu, 0x000ffc00u, 0x3ff00000u, 0xc0000000u};
    uint32_t v = 0;
    for (int j = 0; j < 4; ++j) v |= ((uint32_t)cvt_rn(c.v[j] * kMax[j]) << kShift[j]) & kMask[j];
    return v;
}
/* rgb_cast (rgba == false) / rgba_cast (rgba == true) <T, packed>; out[3] is only written by rgba_cast */
static void plain_from_packed(uint32_t v, int ci, int ok, bool rgba, TexelValue out[4])
{
    uint32_t c[4];
    unpack(ci, v, c);
    for (int j = 0; j < (rgba ? 4 : 3); ++j)
    {
        const uint32_t cm = pack_max(ci, j);
        TexelValue o{};
        if (!cm) { out[j] = limit_component(ok, true); continue; } /* 332 / 565 have no alpha: max(T) */
        if (ok <= 3)
        {
            const uint64_t tMax = kIntMax[ok];
            if (ci != 4) o.u = ((uint64_t)c[j] * (tMax / ((uint64_t)cm & tMax))) & tMax;
            else if (!rgba)
            {
                uint64_t prod = (uint64_t)c[j] * tMax; /* int for uint8/uint16 (no overflow), else T's own width */
                if (ok == 2) prod &= 0xffffffffull;
                o.u = (prod / (1023ull & tMax)) & tMax;
            }
            else if (ok == 0) o.u = j < 3 ? (uint64_t)((c[j] >> 2) & 0xffu) : (uint64_t)(c[j] * 85u); /* AVX2 specialisation (:846-866) */
            else o.u = (uint64_t)((uint32_t)(c[j] * (uint32_t)tMax) / cm) & tMax;
        }
        else if (ci == 4 && rgba && ok == 5) /* AVX2 specialisation (:830-843): sign-extended field, abs, * (1 / max) */
        {
            const int bits = kPackBits[ci][j];
            const int32_t ext = (int32_t)(c[j] << (32 - bits)) >> (32 - bits);
            o.f = (float)(ext < 0 ? -ext : ext) * (1.f / (float)cm);
        }
        else if (ok == 5) o.f = (float)c[j] / (float)cm;
        else if (ok == 6) o.d = (double)c[j] / (double)cm;
        else o.h = float_to_half_rn((float)c[j] / (float)cm); /* half / half: float division, rounded to half */
        out[j] = o;
    }
}
static uint32_t packed_from_packed(uint32_t v, int si, int di)
{
    if (si == di) return v;
    const bool viaRgba = di >= 2;
    const int midKind = di == 4 ? 1 : 0;
    TexelValue mid[4];
    plain_from_packed(v, si, midKind, viaRgba, mid);
    return packed_from_plain(mid, midKind, di, viaRgba);
}
extern "C++" { /* declared next to store_color, outside the C ABI block */
static uint32_t packed_from_float(int ci, const math::vec4& c)
{
    if (ci == 4) return packed1010102_from_float(c);
    TexelValue t[4];
    for (int j = 0; j < 4; ++j) { t[j] = TexelValue{}; t[j].f = c.v[j]; }
    return packed_from_plain(t, 5, ci, true);
}
static math::vec4 float_from_packed(int ci, uint32_t v)
{
    TexelValue t[4];
    plain_from_packed(v, ci, 5, true, t);
    return math::vec4{t[0].f, t[1].f, t[2].f, t[3].f};
}
} // extern "C++"
static void blit_texel_packed(const unsigned char* src, int srcType, unsigned char* dst, int dstType)
{
    static const int kBytes[7] = {1, 2, 4, 8, 2, 4, 8};
    if (is_packed(srcType) && is_packed(dstType))
    {
        store_packed(dst, dstType - 28, packed_from_packed(load_packed(src, srcType - 28), srcType - 28, dstType - 28));
    }
    else if (is_packed(dstType))
    {
        const int ik = srcType % 7, srcN = srcType / 7 + 1;
        TexelValue c[3];
        for (int j = 0; j < 3; ++j) c[j] = j < srcN ? load_component(src + j * kBytes[ik], ik) : limit_component(ik, false);
        store_packed(dst, dstType - 28, packed_from_plain(c, ik, dstType - 28, false));
    }
    else
    {
        const int ok = dstType % 7, dstN = dstType / 7 + 1;
        TexelValue c[4];
        plain_from_packed(load_packed(src, srcType - 28), srcType - 28, ok, dstN == 2 || dstN == 4, c);
        for (int j = 0; j < dstN; ++j) store_component(dst + j * kBytes[ok], ok, c[j]);
    }
}
} // namespace slo

/* SL_BlitProcessor::blit_nearest (src/SL_BlitProcessor.cpp:620-667), one thread */
int slor_blit_rect(slcu_context* c, size_t outTex, size_t inTex, uint16_t sx0, uint16_t sy0, uint16_t sx1, uint16_t sy1,
                   uint16_t dx0, uint16_t dy0, uint16_t dx1, uint16_t dy1)
{
    if (outTex >= c->textures.size() || inTex >= c->textures.size()) return SLCU_ERR_ARG;
    const SL_Texture& s = *c->textures[inTex];
    SL_Texture& d = *c->textures[outTex];
    if (sx1 < sx0 || sy1 < sy0 || dx1 <= dx0 || dy1 <= dy0) return 0;
    const uint64_t inW = (uint64_t)sx1 - sx0, inH = (uint64_t)sy1 - sy0, outW = (uint64_t)dx1 - dx0;
    const uint64_t totalOutW = d.mWidth, totalOutH = d.mHeight;
    const uint64_t x0 = dx0, x1 = std::min<uint64_t>(totalOutW, x0 + outW);
    const uint64_t foutW = ((inW << 16) / totalOutW) + 1u, foutH = ((inH << 16) / totalOutH) + 1u;
    for (uint64_t y = dy0; y < dy1 && y < totalOutH; ++y)
    {
        const uint64_t yf = (y * foutH) >> 16;
        const uint64_t srcY = (uint64_t)sy1 - ((uint64_t)sy0 + yf) - 1u;
        for (uint64_t x = x0; x < x1; ++x)
        {
            const uint64_t srcX = (x * foutW) >> 16;
            if (srcX >= s.mWidth || srcY >= s.mHeight) continue;
            const size_t si = (size_t)(srcX + (uint64_t)s.mWidth * srcY), di = (size_t)(x + totalOutW * y);
            const bool rgIntQuirk = s.mType == d.mType && s.mType / 7 == 1 && s.mType % 7 <= 3; /* RG integer -> same type: p / max(T), see convert_component */
            if (slo::is_packed(s.mType) || slo::is_packed(d.mType)) slo::blit_texel_packed(s.mTexels + si * s.mBpt, s.mType, d.mTexels + di * d.mBpt, d.mType);
            else if (s.mType == d.mType && !rgIntQuirk) memcpy(d.mTexels + di * d.mBpt, s.mTexels + si * s.mBpt, d.mBpt);
            else slo::blit_texel(s.mTexels + si * s.mBpt, s.mType, d.mTexels + di * d.mBpt, d.mType);
        }
    }
    return 0;
}
/* test hooks: the packed render-target casts on their own (tests/test_texel_host_check.py compares the product's device
 * header, compiled for the host, with them) */
SLCU_EXPORT uint32_t slor_test_packed_from_float(int type, const float rgba[4])
{
    return slo::packed_from_float(type - 28, slo::math::vec4{rgba[0], rgba[1], rgba[2], rgba[3]});
}
SLCU_EXPORT void slor_test_float_from_packed(int type, uint32_t v, float out[4])
{
    const slo::math::vec4 c = slo::float_from_packed(type - 28, v);
    memcpy(out, c.v, 16);
}
int slor_blit(slcu_context* c, size_t outTex, size_t inTex)
{
    if (outTex >= c->textures.size() || inTex >= c->textures.size()) return SLCU_ERR_ARG;
    return slor_blit_rect(c, outTex, inTex, 0, 0, c->textures[inTex]->mWidth, c->textures[inTex]->mHeight, 0, 0,
                          c->textures[outTex]->mWidth, c->textures[outTex]->mHeight);
}
int slor_sync(slcu_context*) { return 0; }
int slor_reset_stats(slcu_context* c) { gFragTrace.clear(); if (c) memset(&c->stats, 0, sizeof(c->stats)); return 0; }
int slor_get_stats(slcu_context* c, slcu_stats* out)
{
    if (gFragTracePath)
    {
        FILE* f = fopen(gFragTracePath, "wb");
        if (f) { fwrite(gFragTrace.data(), 4, gFragTrace.size(), f); fclose(f); }
    }
    c->stats.batches = c->stats.drawCalls; /* every draw call is executed on its own */
    *out = c->stats;
    return 0;
}

static double g_t0;
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int slor_timer_begin(slcu_context*) { g_t0 = now_ms(); return 0; }
int slor_timer_end(slcu_context*, double* ms) { *ms = now_ms() - g_t0; return 0; }
int slor_profile_begin(slcu_context*) { return 0; }
int slor_profile_end(slcu_context*, slcu_kernel_times* out) { memset(out, 0, sizeof(*out)); return 0; }

} // extern "C"
