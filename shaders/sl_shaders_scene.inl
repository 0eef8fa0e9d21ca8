/*
 * Built-in shader pairs, part 2: the scene configs of BASELINE.json.
 * Shader source = user input to the pipeline (see sl_shaders_basic.inl for how
 * this file is compiled three times against three `math::` back-ends).
 */

/* ---- "inst_tex": config C2, the shader pair of
 * softlight/tests/sl_instancing_test.cpp:80-171 (vertex format
 * Tuple<vec3 pos, vec2 uv, vec3 nrm>, bilinear REPEAT RGB8 albedo, half-Lambert).
 * The uniform block keeps the demo's members; the two host pointers it chases
 * (texture, per-instance matrices) are declared with ubo_bind_texture /
 * ubo_bind_buffer so they are valid where the shader runs. ------------------ */
struct InstTexUniforms
{
    const SL_Texture* pTexture;
    math::vec4        ambient;
    math::vec4        diffuse;
    const math::mat4* instanceMatrix;
    math::mat4        modelMatrix;
    math::mat4        vpMatrix;
    math::vec4        camPos;
};

struct InstTexVertex
{
    math::vec3 pos;
    math::vec2 uv;
    math::vec3 norm;
};

SL_SHADER_FN math::vec4 inst_tex_vs(SL_VertexParam& param)
{
    const InstTexUniforms* pUniforms = param.pUniforms->template as<InstTexUniforms>();
    const InstTexVertex* const v     = param.pVbo->template element<const InstTexVertex>(param.pVao->offset(0, param.vertId));
    const math::vec4&&  vert         = math::vec4_cast(v->pos, 1.f);
    const math::vec4&&  uv           = math::vec4_cast(v->uv, 0.f, 0.f);
    const math::vec4&&  norm         = math::vec4_cast(v->norm, 0.f);

    const math::mat4&  instanceMat = pUniforms->instanceMatrix[param.instanceId];
    const math::mat4&& modelMat    = instanceMat * pUniforms->modelMatrix;
    const math::vec4&& pos         = modelMat * vert;
    const math::vec4&& n           = modelMat * norm;

    param.pVaryings[0] = pos;
    param.pVaryings[1] = uv;
    param.pVaryings[2] = n;

    return pUniforms->vpMatrix * pos;
}

SL_SHADER_FN bool inst_tex_fs(SL_FragmentParam& fragParam)
{
    const InstTexUniforms* pUniforms = fragParam.pUniforms->template as<InstTexUniforms>();
    const math::vec4&    pos       = fragParam.pVaryings[0];
    const math::vec4&    uv        = fragParam.pVaryings[1];
    const math::vec4&&   norm      = math::normalize(fragParam.pVaryings[2]);
    const SL_Texture*    pTexture  = pUniforms->pTexture;
    const math::vec4&    ambient   = pUniforms->ambient;
    math::vec4           albedo;

    // normalize the texture colors to within (0.f, 1.f)
    {
        const math::vec3_t<uint8_t>&& pixel8 = sl_sample_bilinear<math::vec3_t<uint8_t>, SL_WrapMode::REPEAT>(*pTexture, uv[0], uv[1]);
        const math::vec4_t<uint8_t>&& pixelF = math::vec4_cast<uint8_t>(pixel8, 255);
        albedo = color_cast<float, uint8_t>(pixelF);
    }

    // Light direction calculation
    math::vec4          lightDir   = math::normalize(pUniforms->camPos - pos);
    const float         lightAngle = 0.5f * math::dot(-lightDir, norm) + 0.5f;
    const math::vec4&&  diffuse    = pUniforms->diffuse * lightAngle;
    const math::vec4&&  rgba       = albedo * (ambient+diffuse);

    fragParam.pOutputs[0] = math::min(rgba, math::vec4{1.f});

    return true;
}

SL_REGISTER_VERTEX_SHADER("inst_tex", inst_tex_vs)
SL_REGISTER_FRAGMENT_SHADER("inst_tex", inst_tex_fs)

static_assert(sizeof(InstTexUniforms) == 208, "InstTexUniforms layout must be identical in every back-end");
