/*
 * Built-in shader pairs, part 3: configs C4 (skinned characters) and C5 (volume ray march) of BASELINE.json.
 * Shader source = user input to the pipeline (see sl_shaders_basic.inl for how this file is compiled three times
 * against three `math::` back-ends: the real ls::math in oracle/ref, the CPU restatement in oracle/port and the
 * __device__ mirror in softlight_b200/csrc).
 */

/* ---- "skin_tex": config C4, the shader pair of softlight/tests/sl_animation_test.cpp:199-280
 * (_textured_skin_vert_shader_impl: vertex format Tuple<vec3 pos, vec2h uv, int32 10-10-10-2 normal, vec4<u16> bone
 * ids, vec4h bone weights> = 36 bytes, four-bone linear blend skinning; _textured_frag_shader: bilinear REPEAT albedo,
 * half-Lambert towards the camera).  AnimUniforms keeps the demo's members; the two host pointers it chases
 * (texture, bone matrices) are declared with ubo_bind_texture / ubo_bind_buffer. ------------------------------------ */
struct AnimUniforms
{
    const SL_Texture* pTexture;
    const math::mat4* pBones;
    math::mat4        modelMatrix;
    math::mat4        vpMatrix;
    math::vec4        camPos;
};

struct SkinVertex
{
    math::vec3              pos;
    math::vec2_t<math::half> uv;
    int32_t                 norm;
    math::vec4_t<uint16_t>  boneIds;
    math::vec4_t<math::half> boneWeights;
};

SL_SHADER_FN math::vec4 skin_tex_vs(SL_VertexParam& param)
{
    const AnimUniforms* pUniforms   = param.pUniforms->template as<AnimUniforms>();
    const math::mat4*   pBones      = pUniforms->pBones;

    const SkinVertex* const v       = param.pVbo->template element<const SkinVertex>(param.pVao->offset(0, param.vertId));
    const math::vec4&&  vert        = math::vec4_cast(v->pos, 1.f);
    const math::vec2_t<math::half> uv = v->uv;
    const math::vec4&&  norm        = sl_unpack_vec4_10_10_10_2I(v->norm);
    const math::vec4_t<uint16_t> boneIds = v->boneIds;
    const math::vec4&&  boneWeights = (math::vec4)v->boneWeights;

    math::mat4&& boneTrans =
        (pBones[boneIds[0]] * boneWeights[0]) +
        (pBones[boneIds[1]] * boneWeights[1]) +
        (pBones[boneIds[2]] * boneWeights[2]) +
        (pBones[boneIds[3]] * boneWeights[3]);

    const math::mat4&& modelPos = pUniforms->modelMatrix * boneTrans;
    const math::vec4&& pos = modelPos * vert;

    param.pVaryings[0] = pos;
    param.pVaryings[1] = math::vec4_cast((math::vec2)uv, 0.f, 0.f);
    param.pVaryings[2] = modelPos * norm;

    return pUniforms->vpMatrix * pos;
}

SL_SHADER_FN bool skin_tex_fs(SL_FragmentParam& fragParam)
{
    const AnimUniforms*  pUniforms = fragParam.pUniforms->template as<AnimUniforms>();
    const math::vec4&    pos       = fragParam.pVaryings[0];
    const math::vec4&    uv        = fragParam.pVaryings[1];
    const math::vec4&&   norm      = math::normalize(fragParam.pVaryings[2]);
    const SL_Texture*    pTexture  = pUniforms->pTexture;
    const math::vec4     ambient   = {0.5f, 0.5f, 0.5f, 1.f};
    math::vec4           albedo;

    // normalize the texture colors to within (0.f, 1.f)
    if (!pTexture)
    {
        albedo = math::vec4{1.f, 1.f, 1.f, 1.f};
    }
    else if (pTexture->channels() == 3)
    {
        const math::vec3_t<uint8_t>&& pixel8 = sl_sample_bilinear<math::vec3_t<uint8_t>, SL_WrapMode::REPEAT>(*pTexture, uv[0], uv[1]);
        const math::vec4_t<uint8_t>&& pixelF = math::vec4_cast<uint8_t>(pixel8, 255);
        albedo = color_cast<float, uint8_t>(pixelF);
    }
    else
    {
        const math::vec4_t<uint8_t>&& pixelF = sl_sample_bilinear<math::vec4_t<uint8_t>, SL_WrapMode::REPEAT>(*pTexture, uv[0], uv[1]);
        albedo = color_cast<float, uint8_t>(pixelF);
    }

    // Light direction calculation
    math::vec4          lightDir   = math::normalize(pUniforms->camPos - pos);
    const float         lightAngle = 0.5f * math::dot(-lightDir, norm) + 0.5f;
    const math::vec4&&  diffuse    = math::vec4{1.f} * lightAngle;

    const math::vec4 rgba = albedo * (ambient+diffuse);
    fragParam.pOutputs[0] = math::min(rgba, math::vec4{1.f});

    return true;
}

SL_REGISTER_VERTEX_SHADER("skin_tex", skin_tex_vs)
SL_REGISTER_FRAGMENT_SHADER("skin_tex", skin_tex_fs)

static_assert(sizeof(AnimUniforms) == 160, "AnimUniforms layout must be identical in every back-end");
static_assert(sizeof(SkinVertex) == 36, "SkinVertex is Tuple<vec3, vec2h, int32, vec4<u16>, vec4h>");


/* ---- "volume": config C5, the shader pair of softlight/tests/sl_volume_rendering_test.cpp:72-254: a unit cube scaled
 * by the volume spacing; the fragment shader intersects the view ray with the box, rejects rays that only see empty
 * space with a 32-step nearest-neighbour probe, then marches 256 trilinear steps back to front through an R8 3-D
 * texture, looking up a float opacity table and an RGB-float colour table and lighting with a central-difference
 * style normal.  Output is premultiplied RGBA. ------------------------------------------------------------------- */
struct VolumeUniforms
{
    const SL_Texture* pCubeMap;
    const SL_Texture* pOpacityMap;
    const SL_Texture* pColorMap;
    math::vec4        spacing;
    math::vec4        camPos;
    math::mat4        mvpMatrix;
};

SL_SHADER_FN math::vec4 volume_vs(SL_VertexParam& param)
{
    const VolumeUniforms* pUniforms = param.pUniforms->template as<VolumeUniforms>();
    const math::vec4      spacing   = pUniforms->spacing;
    const math::vec3&     vert      = *(param.pVbo->template element<const math::vec3>(param.pVao->offset(0, param.vertId)));
    const math::mat4&&    modelMat  = math::scale(math::mat4{1.f}, math::vec3_cast(spacing));
    const math::vec4&&    pos       = modelMat * math::vec4_cast(vert, 1.f);

    param.pVaryings[0] = pos;
    param.pVaryings[1] = pos - pUniforms->camPos;

    return pUniforms->mvpMatrix * pos;
}

SL_SHADER_FN bool volume_intersect_ray_box(
    const math::vec4& rayPos,
    const math::vec4& rayDir,
    const math::vec4& spacing,
    float& texNear,
    float& texFar)
{
    const math::vec4&& invR    = math::rcp(rayDir);
    const math::vec4&& tbot    = invR * (-spacing-rayPos);
    const math::vec4&& ttop    = invR * (spacing-rayPos);

    const math::vec4&& tmin    = math::min(ttop, tbot);
    const math::vec2   minXX   = {tmin[0], tmin[0]};
    const math::vec2   minYZ   = {tmin[1], tmin[2]};
    const math::vec2&& nearVal = math::max(minXX, minYZ);
    texNear = math::max(0.f, nearVal[0], nearVal[1]);

    const math::vec4&& tmax    = math::max(ttop, tbot);
    const math::vec2   maxXX   = {tmax[0], tmax[0]};
    const math::vec2   maxYZ   = {tmax[1], tmax[2]};
    const math::vec2&& farVal  = math::min(maxXX, maxYZ);
    texFar = math::min(farVal[0], farVal[1]);

    return texNear <= texFar;
}

template <unsigned step>
SL_SHADER_FN math::vec4 volume_calc_normal(const SL_Texture& tex, const math::vec4& p)
{
    constexpr float stepLen = 1.f / (float)step;
    const math::vec4&& a = p - math::vec4{stepLen, 0.f, 0.f, 0.f};
    const math::vec4&& b = p - math::vec4{0.f, stepLen, 0.f, 0.f};
    const math::vec4&& c = p - math::vec4{0.f, 0.f, stepLen, 0.f};

    return (math::normalize((math::vec4)math::vec4_t<unsigned int>{
        sl_sample_trilinear<SL_ColorRType<unsigned char>, SL_WrapMode::EDGE, SL_TexelOrder::ORDERED>(tex, a[0], a[1], a[2]).r,
        sl_sample_trilinear<SL_ColorRType<unsigned char>, SL_WrapMode::EDGE, SL_TexelOrder::ORDERED>(tex, b[0], b[1], b[2]).r,
        sl_sample_trilinear<SL_ColorRType<unsigned char>, SL_WrapMode::EDGE, SL_TexelOrder::ORDERED>(tex, c[0], c[1], c[2]).r,
        0u
    }) * 2.f) - math::vec4{1.f, 1.f, 1.f, 0.f};
}

SL_SHADER_FN bool volume_can_skip_render(const SL_Texture& volumeTex, const math::vec4& ray, math::vec4 rayPos)
{
    constexpr unsigned numTestSteps = 32;
    const math::vec4&& rayStep = ray * (1.f/(float)numTestSteps);

    for (unsigned i = 0; i < numTestSteps; ++i)
    {
        const SL_ColorR8&& intensity = sl_sample_nearest<SL_ColorR8, SL_WrapMode::EDGE>(volumeTex, rayPos[0], rayPos[1], rayPos[2]);

        if (intensity.r > 16)
        {
            return false;
        }

        rayPos -= rayStep;
    }

    return true;
}

SL_SHADER_FN bool volume_fs(SL_FragmentParam& fragParam)
{
    constexpr unsigned numSteps = 256;
    constexpr float step = 1.f / (float)numSteps;

    const VolumeUniforms* pUniforms = fragParam.pUniforms->template as<VolumeUniforms>();
    const math::vec4&     spacing   = pUniforms->spacing;
    const math::vec4&&    scaling   = math::rcp(spacing);
    const SL_Texture&     volumeTex = *pUniforms->pCubeMap;
    const SL_Texture&     alphaTex  = *pUniforms->pOpacityMap;
    const SL_Texture&     colorTex  = *pUniforms->pColorMap;
    const math::vec4&&    pos       = fragParam.pVaryings[0] * scaling;
    const math::vec4&&    rayDir    = math::normalize(fragParam.pVaryings[1]);
    float                 nearPos;
    float                 farPos;

    const bool intersectInternalBox = volume_intersect_ray_box(pos, rayDir, spacing, nearPos, farPos);
    if (!intersectInternalBox)
    {
        return false;
    }

    const math::vec4&& rayFar    = (pos + rayDir * farPos + 1.f) * 0.5f;
    const math::vec4&& rayNear   = (pos + rayDir * nearPos + 1.f) * 0.5f;
    const math::vec4&& ray       = rayFar - rayNear;
    const math::vec4&& rayStep   = ray * step;
    math::vec4         rayPos    = rayFar;
    math::vec4         dstTexel  = math::vec4{0.f};
    unsigned           intensity;
    float              srcAlpha;

    // Test pixels with minimal filtering before attempting to do anything
    // more expensive
    if (volume_can_skip_render(volumeTex, ray, rayFar))
    {
        return false;
    }

    for (unsigned i = 0; (i < numSteps) && (dstTexel[3] < 1.f); ++i)
    {
        intensity = sl_sample_trilinear<SL_ColorR8, SL_WrapMode::EDGE>(volumeTex, rayPos[0], rayPos[1], rayPos[2]).r;

        if (intensity > 16)
        {
            // regular opacity (doesn't take ray steps into account).
            srcAlpha = alphaTex.template texel<float>(intensity) * step * 100.f;
            if (srcAlpha > 0.f)
            {
                const math::vec4&& norm      = volume_calc_normal<numSteps>(volumeTex, rayPos);
                const float        luminance = 2.f * math::clamp(math::dot(norm, math::vec4{1.f, 0.f, 1.f, 0.f}), 0.f, 1.f);
                const math::vec3&& volColor  = colorTex.template texel<SL_ColorRGBf>(intensity) * luminance;
                const math::vec4&& srcRGBA   = math::vec4_cast(volColor, 1.f) * srcAlpha;

                dstTexel = math::fmadd(dstTexel, math::vec4{1.f}-srcAlpha, srcRGBA);
            }
        }

        rayPos -= rayStep;
    }

    // output composition
    fragParam.pOutputs[0] = math::clamp(dstTexel, math::vec4{0.f}, math::vec4{1.f});

    return dstTexel[3] > 0.f;
}

SL_REGISTER_VERTEX_SHADER("volume", volume_vs)
SL_REGISTER_FRAGMENT_SHADER("volume", volume_fs)

/* Diagnostic variants of "volume" (parity bisection on the GPU box, tools/gpu_compare.py): same vertex shader, the
 * fragment shader writes intermediate values of volume_fs instead of the composited colour. */
SL_SHADER_FN bool volume_dbg0_fs(SL_FragmentParam& fragParam)
{
    fragParam.pOutputs[0] = fragParam.pVaryings[0] + fragParam.pVaryings[1];
    return true;
}

SL_SHADER_FN bool volume_dbg1_fs(SL_FragmentParam& fragParam)
{
    const VolumeUniforms* pUniforms = fragParam.pUniforms->template as<VolumeUniforms>();
    const math::vec4&     spacing   = pUniforms->spacing;
    const math::vec4&&    scaling   = math::rcp(spacing);
    const math::vec4&&    pos       = fragParam.pVaryings[0] * scaling;
    const math::vec4&&    rayDir    = math::normalize(fragParam.pVaryings[1]);
    float                 nearPos   = 0.f;
    float                 farPos    = 0.f;
    volume_intersect_ray_box(pos, rayDir, spacing, nearPos, farPos);
    fragParam.pOutputs[0] = math::vec4{nearPos, farPos, rayDir[0] + rayDir[1], rayDir[2]};
    return true;
}

SL_SHADER_FN bool volume_dbg2_fs(SL_FragmentParam& fragParam)
{
    constexpr unsigned numSteps = 256;
    constexpr float step = 1.f / (float)numSteps;
    const VolumeUniforms* pUniforms = fragParam.pUniforms->template as<VolumeUniforms>();
    const math::vec4&     spacing   = pUniforms->spacing;
    const math::vec4&&    scaling   = math::rcp(spacing);
    const SL_Texture&     volumeTex = *pUniforms->pCubeMap;
    const SL_Texture&     alphaTex  = *pUniforms->pOpacityMap;
    const math::vec4&&    pos       = fragParam.pVaryings[0] * scaling;
    const math::vec4&&    rayDir    = math::normalize(fragParam.pVaryings[1]);
    float                 nearPos;
    float                 farPos;
    if (!volume_intersect_ray_box(pos, rayDir, spacing, nearPos, farPos))
    {
        return false;
    }
    const math::vec4&& rayFar    = (pos + rayDir * farPos + 1.f) * 0.5f;
    const math::vec4&& rayNear   = (pos + rayDir * nearPos + 1.f) * 0.5f;
    const math::vec4&& ray       = rayFar - rayNear;
    const math::vec4&& rayStep   = ray * step;
    math::vec4         rayPos    = rayFar;
    float sumI = 0.f, sumA = 0.f, sumL = 0.f, sumN = 0.f;
    for (unsigned i = 0; i < numSteps; ++i)
    {
        const unsigned intensity = sl_sample_trilinear<SL_ColorR8, SL_WrapMode::EDGE>(volumeTex, rayPos[0], rayPos[1], rayPos[2]).r;
        sumI += (float)(int)intensity;
        if (intensity > 16)
        {
            const float srcAlpha = alphaTex.template texel<float>(intensity) * step * 100.f;
            sumA += srcAlpha;
            const math::vec4&& norm = volume_calc_normal<numSteps>(volumeTex, rayPos);
            sumN += (norm[0] + norm[1]) + norm[2];
            sumL += 2.f * math::clamp(math::dot(norm, math::vec4{1.f, 0.f, 1.f, 0.f}), 0.f, 1.f);
        }
        rayPos -= rayStep;
    }
    fragParam.pOutputs[0] = math::vec4{sumI, sumA, sumN, sumL};
    return true;
}

SL_REGISTER_FRAGMENT_SHADER("volume_dbg0", volume_dbg0_fs)
SL_REGISTER_FRAGMENT_SHADER("volume_dbg1", volume_dbg1_fs)
SL_REGISTER_FRAGMENT_SHADER("volume_dbg2", volume_dbg2_fs)

static_assert(sizeof(VolumeUniforms) == 128, "VolumeUniforms layout must be identical in every back-end");
