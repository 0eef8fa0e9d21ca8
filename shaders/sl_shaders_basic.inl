/*
 * Built-in shader pairs, part 1: the draw-test shader and config C1 (Lambert).
 *
 * This file is *shader source*, i.e. user input to the pipeline, not part of
 * it.  It is written once against the reference's shader-facing API
 * (`math::` = ls::math, SL_VertexParam / SL_FragmentParam field names,
 * SL_Shader.hpp:41-97) and compiled three times:
 *   - oracle/ref/slref_shim.cpp : math = ls::math (the real x86 intrinsics)
 *   - oracle/port/slor_oracle.cpp : math = the oracle's plain-C restatement
 *   - softlight_b200/csrc/*.cu  : math = sld::math (__device__ mirrors)
 * The including context defines SL_SHADER_FN and the `math` namespace alias.
 */

/* ---- "vcolor": the shader of softlight/tests/sl_draw_test.cpp:27-83 ------ */
struct VColorVertex
{
    math::vec4 pos;
    math::vec4 color;
};

SL_SHADER_FN math::vec4 vcolor_vs(SL_VertexParam& param)
{
    const VColorVertex& vert = *(param.pVbo->template element<const VColorVertex>(param.pVao->offset(0, param.vertId)));
    param.pVaryings[0] = vert.color;
    return vert.pos;
}

SL_SHADER_FN bool vcolor_fs(SL_FragmentParam& fragParam)
{
    fragParam.pOutputs[0] = fragParam.pVaryings[0];
    return true;
}

/* ---- "lambert": config C1 ------------------------- */
struct LambertUniforms
{
    math::mat4 mvp;
    math::vec4 lightDir;
};

struct LambertVertex
{
    math::vec4 pos;
    math::vec4 nrm;
};

SL_SHADER_FN math::vec4 lambert_vs(SL_VertexParam& param)
{
    const LambertUniforms* pUniforms = param.pUniforms->template as<LambertUniforms>();
    const LambertVertex&   vert      = *(param.pVbo->template element<const LambertVertex>(param.pVao->offset(0, param.vertId)));

    param.pVaryings[0] = vert.nrm;
    return pUniforms->mvp * vert.pos;
}

SL_SHADER_FN bool lambert_fs(SL_FragmentParam& fragParam)
{
    const LambertUniforms* pUniforms = fragParam.pUniforms->template as<LambertUniforms>();
    const math::vec4&      n         = fragParam.pVaryings[0];
    const float            d         = math::max(math::dot(n, pUniforms->lightDir), 0.f);

    fragParam.pOutputs[0] = math::vec4{d, d, d, 1.f};
    return true;
}

/* ---- "flat": constant colour from the uniform block; discards odd primitives
 * never -- used by tests that need a shader with zero varyings. ------------ */
struct FlatUniforms
{
    math::mat4 mvp;
    math::vec4 color;
};

SL_SHADER_FN math::vec4 flat_vs(SL_VertexParam& param)
{
    const FlatUniforms* pUniforms = param.pUniforms->template as<FlatUniforms>();
    const math::vec4&   pos       = *(param.pVbo->template element<const math::vec4>(param.pVao->offset(0, param.vertId)));
    return pUniforms->mvp * pos;
}

SL_SHADER_FN bool flat_fs(SL_FragmentParam& fragParam)
{
    const FlatUniforms* pUniforms = fragParam.pUniforms->template as<FlatUniforms>();
    fragParam.pOutputs[0] = pUniforms->color;
    return true;
}

/* four render targets, like softlight/tests/sl_mrt_test.cpp:105-123 (clamped colour / position / uv / normal): here four
 * clamped functions of the interpolated colour and of the fragment coordinate */
SL_SHADER_FN bool vcolor_mrt_fs(SL_FragmentParam& fragParam)
{
    const math::vec4& c = fragParam.pVaryings[0];
    const math::vec4 coord{(float)(fragParam.coord.x & 255) * (1.f / 255.f), (float)(fragParam.coord.y & 255) * (1.f / 255.f), fragParam.coord.depth, 1.f};
    fragParam.pOutputs[0] = math::clamp(c, math::vec4{0.f}, math::vec4{1.f});
    fragParam.pOutputs[1] = math::clamp(c * c, math::vec4{0.f}, math::vec4{1.f});
    fragParam.pOutputs[2] = math::clamp(coord, math::vec4{0.f}, math::vec4{1.f});
    fragParam.pOutputs[3] = math::clamp(math::vec4{1.f} - c, math::vec4{0.f}, math::vec4{1.f});
    return true;
}

/* ---- "sampler_matrix": parity of SL_Sampler.hpp / SL_Texture.hpp addressing beyond what the configs use: every filter
 * (nearest, bilinear, trilinear) x wrap mode (REPEAT, EDGE, BORDER) x texel order (ORDERED, SWIZZLED) over 8-bit RGBA / RGB /
 * R textures in 2-D and 3-D; uniforms select the combination, the interpolated colour carries (u, v, w) beyond [0, 1). ---- */
struct SamplerMatrixUniforms
{
    const SL_Texture* pRgba2dOrdered;
    const SL_Texture* pRgba2dSwizzled;
    const SL_Texture* pRgb2dOrdered;
    const SL_Texture* pR3dOrdered;
    const SL_Texture* pR3dSwizzled;
    const SL_Texture* pRgba3dOrdered;
    int               mode;
    int               pad;
};

SL_SHADER_FN bool sampler_matrix_fs(SL_FragmentParam& fragParam)
{
    const SamplerMatrixUniforms* pUniforms = fragParam.pUniforms->template as<SamplerMatrixUniforms>();
    const float u = fragParam.pVaryings[0][0], v = fragParam.pVaryings[0][1], w = fragParam.pVaryings[0][2];
    math::vec4 out{0.f, 0.f, 0.f, 1.f};
    switch (pUniforms->mode)
    {
        case 0: out = color_cast<float, uint8_t>(sl_sample_nearest<math::vec4_t<uint8_t>, SL_WrapMode::REPEAT, SL_TexelOrder::ORDERED>(*pUniforms->pRgba2dOrdered, u, v)); break;
        case 1: out = color_cast<float, uint8_t>(sl_sample_nearest<math::vec4_t<uint8_t>, SL_WrapMode::EDGE, SL_TexelOrder::SWIZZLED>(*pUniforms->pRgba2dSwizzled, u, v)); break;
        case 2: out = math::vec4_cast(color_cast<float, uint8_t>(sl_sample_nearest<math::vec3_t<uint8_t>, SL_WrapMode::BORDER, SL_TexelOrder::ORDERED>(*pUniforms->pRgb2dOrdered, u, v)), 1.f); break;
        case 3: out = color_cast<float, uint8_t>(sl_sample_bilinear<math::vec4_t<uint8_t>, SL_WrapMode::REPEAT, SL_TexelOrder::SWIZZLED>(*pUniforms->pRgba2dSwizzled, u, v)); break;
        case 4: out = math::vec4_cast(color_cast<float, uint8_t>(sl_sample_bilinear<math::vec3_t<uint8_t>, SL_WrapMode::EDGE, SL_TexelOrder::ORDERED>(*pUniforms->pRgb2dOrdered, u, v)), 1.f); break;
        case 5: out = color_cast<float, uint8_t>(sl_sample_bilinear<math::vec4_t<uint8_t>, SL_WrapMode::BORDER, SL_TexelOrder::ORDERED>(*pUniforms->pRgba2dOrdered, u, v)); break;
        case 6: out = math::vec4{color_cast<float, uint8_t>(sl_sample_nearest<SL_ColorR8, SL_WrapMode::REPEAT, SL_TexelOrder::ORDERED>(*pUniforms->pR3dOrdered, u, v, w)).r, 0.f, 0.f, 1.f}; break;
        case 7: out = math::vec4{color_cast<float, uint8_t>(sl_sample_nearest<SL_ColorR8, SL_WrapMode::BORDER, SL_TexelOrder::SWIZZLED>(*pUniforms->pR3dSwizzled, u, v, w)).r, 0.f, 0.f, 1.f}; break;
        case 8: out = math::vec4{color_cast<float, uint8_t>(sl_sample_trilinear<SL_ColorR8, SL_WrapMode::REPEAT, SL_TexelOrder::ORDERED>(*pUniforms->pR3dOrdered, u, v, w)).r, 0.f, 0.f, 1.f}; break;
        case 9: out = math::vec4{color_cast<float, uint8_t>(sl_sample_trilinear<SL_ColorR8, SL_WrapMode::EDGE, SL_TexelOrder::SWIZZLED>(*pUniforms->pR3dSwizzled, u, v, w)).r, 0.f, 0.f, 1.f}; break;
        case 10: out = color_cast<float, uint8_t>(sl_sample_trilinear<math::vec4_t<uint8_t>, SL_WrapMode::BORDER, SL_TexelOrder::ORDERED>(*pUniforms->pRgba3dOrdered, u, v, w)); break;
        default: out = color_cast<float, uint8_t>(sl_sample_bilinear<math::vec4_t<uint8_t>, SL_WrapMode::EDGE, SL_TexelOrder::SWIZZLED>(*pUniforms->pRgba2dSwizzled, u, v)); break;
    }
    fragParam.pOutputs[0] = out;
    return true;
}

SL_REGISTER_VERTEX_SHADER("vcolor", vcolor_vs)
SL_REGISTER_FRAGMENT_SHADER("vcolor", vcolor_fs)
SL_REGISTER_FRAGMENT_SHADER("sampler_matrix", sampler_matrix_fs)
SL_REGISTER_FRAGMENT_SHADER("vcolor_mrt", vcolor_mrt_fs)
SL_REGISTER_VERTEX_SHADER("lambert", lambert_vs)
SL_REGISTER_FRAGMENT_SHADER("lambert", lambert_fs)
SL_REGISTER_VERTEX_SHADER("flat", flat_vs)
SL_REGISTER_FRAGMENT_SHADER("flat", flat_fs)
