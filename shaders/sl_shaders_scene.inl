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


/* ---- "mesh_pbr": config C3, the shader pair of softlight/tests/sl_large_scene_test.cpp:384-408 (texture_vert_shader:
 * vertex format Tuple<vec3 pos, vec2 uv, int32 10-10-10-2 normal>) and :517-632 (texture_frag_shader_pbr<4>) with the
 * PBR helpers of :110-160.  The albedo is sampled bilinear / REPEAT from a SWIZZLED RGBA8 texture.
 * MeshUniforms keeps the demo's members (no bump map); pTexture is declared with ubo_bind_texture. ------------------- */
struct PbrLight
{
    math::vec4 pos;
    math::vec4 ambient;
    math::vec4 diffuse;
};

struct PbrPointLight
{
    float constant;
    float linear;
    float quadratic;
    float padding;
};

struct MeshUniforms
{
    const SL_Texture* pTexture;
    math::vec4 camPos;
    PbrLight light;
    PbrPointLight point;
    math::mat4 modelMatrix;
    math::mat4 mvpMatrix;
};

struct MeshVertex
{
    math::vec3 pos;
    math::vec2 uv;
    int32_t    norm;
};

SL_SHADER_FN math::vec4 mesh_pbr_vs(SL_VertexParam& param)
{
    const MeshUniforms* pUniforms = param.pUniforms->template as<MeshUniforms>();
    const MeshVertex*   v         = param.pVbo->template element<const MeshVertex>(param.pVao->offset(0, param.vertId));
    const math::vec4&&  vert      = math::vec4_cast(v->pos, 1.f);
    const math::vec4&&  uv        = math::vec4_cast(v->uv, 0.f, 0.f);
    const math::vec4&&  norm      = sl_unpack_vec4_10_10_10_2I(v->norm);

    param.pVaryings[0] = pUniforms->modelMatrix * vert;
    param.pVaryings[1] = uv;
    param.pVaryings[2] = pUniforms->modelMatrix * norm;

    return pUniforms->mvpMatrix * vert;
}

// Calculate the metallic component of a surface
SL_SHADER_FN math::vec4 pbr_fresnel_schlick(float cosTheta, const math::vec4& surfaceReflection)
{
    return math::fmadd(math::vec4{1.f} - surfaceReflection, math::vec4{math::pow(1.f - cosTheta, 5.f)}, surfaceReflection);
}

// normal distribution function within a hemisphere
SL_SHADER_FN float pbr_distribution_ggx(const math::vec4& norm, const math::vec4& hemisphere, float roughness)
{
    float roughSquared = roughness * roughness;
    float roughQuad    = roughSquared * roughSquared;
    float nDotH        = math::max(math::dot(norm, hemisphere), 0.f);
    float nDotH2       = nDotH * nDotH;
    float distribution = nDotH2 * (roughQuad - 1.f) + 1.f;

    return nDotH2 / (LS_PI * distribution  * distribution);
}

// Determine how a surface's roughness affects how it reflects light
SL_SHADER_FN float pbr_geometry_schlick_ggx(float normDotView, float roughness)
{
    roughness += 1.f;
    roughness = (roughness*roughness) * 0.125f; // 1/8

    float geometry = normDotView * (1.f - roughness) + roughness;
    return normDotView / geometry;
}

// PBR Geometry function for determining how light bounces off a surface
SL_SHADER_FN float pbr_geometry_smith(const math::vec4& norm, const math::vec4& viewDir, const math::vec4& radiance, float roughness)
{
    float normDotView = math::max(math::dot(norm, viewDir), 0.f);
    float normDotRad = math::max(math::dot(norm, radiance), 0.f);

    return pbr_geometry_schlick_ggx(normDotView, roughness) * pbr_geometry_schlick_ggx(normDotRad, roughness);
}

SL_SHADER_FN bool mesh_pbr_fs(SL_FragmentParam& fragParams)
{
    const MeshUniforms* pUniforms  = fragParams.pUniforms->template as<MeshUniforms>();
    const math::vec4     pos       = fragParams.pVaryings[0];
    const math::vec4     uv        = fragParams.pVaryings[1];
    math::vec4&&         norm      = math::normalize(fragParams.pVaryings[2]);
    math::vec4&          output    = fragParams.pOutputs[0];
    const SL_Texture*    pTexture  = pUniforms->pTexture;
    math::vec4           pixel;

    // normalize the texture colors to within (0.f, 1.f)
    {
        const math::vec4_t<uint8_t>&& pixelF = sl_sample_bilinear<math::vec4_t<uint8_t>, SL_WrapMode::REPEAT, SL_TexelOrder::SWIZZLED>(*pTexture, uv[0], uv[1]);
        pixel = color_cast<float, uint8_t>(pixelF);
    }

    // gamma correction
    pixel = math::pow(pixel, math::vec4{2.2f});

    // surface model
    const math::vec4     camPos           = pUniforms->camPos;
    const math::vec4     viewDir          = math::normalize(camPos - pos);
    const math::vec4     lightPos         = pUniforms->light.pos;
    const math::vec4     albedo           = pixel;
    constexpr float      metallic         = 0.4f;
    constexpr float      roughness        = 0.35f;
    constexpr float      ambientIntensity = 0.5f;
    constexpr float      diffuseIntensity = 50.f;

    // Metallic reflectance at a normal incidence
    // 0.04f should be close to plastic.
    const math::vec4   surfaceConstant   = {0.4f, 0.4f, 0.4f, 1.f};
    const math::vec4&& surfaceReflection = math::mix(surfaceConstant, albedo, metallic);

    math::vec4         lightDir0         = math::vec4{0.f};

    math::vec4         lightDirN         = lightPos - pos;
    const float        distance          = math::length(lightDirN);
    lightDirN                            = lightDirN * math::rcp(distance); // normalize
    math::vec4         hemisphere        = math::normalize(viewDir + lightDirN);
    const float        attenuation       = math::rcp(distance);
    math::vec4         radianceObj       = pUniforms->light.diffuse * attenuation * diffuseIntensity;

    const float        ndf               = pbr_distribution_ggx(norm, hemisphere, roughness);
    const float        geom              = pbr_geometry_smith(norm, viewDir, lightDirN, roughness);
    const math::vec4   fresnel           = pbr_fresnel_schlick(math::max(math::dot(hemisphere, viewDir), 0.f), surfaceReflection);

    const math::vec4   brdf              = fresnel * ndf * geom;
    const float        cookTorrance      = 4.f * math::max(math::dot(norm, viewDir), 0.f) * math::max(math::dot(norm, lightDirN), 0.f) + LS_EPSILON;  // avoid divide-by-0
    const math::vec4   specular          = brdf * math::rcp(cookTorrance);

    const math::vec4&  specContrib       = fresnel;
    const math::vec4   refractRatio      = (math::vec4{1.f} - specContrib) * (math::vec4{1.f} - metallic);

    const float normDotLight             = math::max(math::dot(lightDirN, norm), 0.f);
    lightDir0                            += (refractRatio * albedo * LS_PI_INVERSE + specular) * radianceObj * normDotLight;

    const math::vec4   ambient           = pUniforms->light.ambient * ambientIntensity;

    // Color normalization and light contribution
    math::vec4 outRGB = albedo * (ambient + lightDir0);

    // HDR Tone mapping
    const float exposure = 4.f;
    outRGB = math::vec4{1.f} - math::exp(-outRGB * exposure);
    outRGB[3] = 1.f;

    output = outRGB;

    return true;
}

SL_REGISTER_VERTEX_SHADER("mesh_pbr", mesh_pbr_vs)
SL_REGISTER_FRAGMENT_SHADER("mesh_pbr", mesh_pbr_fs)

static_assert(sizeof(MeshUniforms) == 224, "MeshUniforms layout must be identical in every back-end");
static_assert(sizeof(MeshVertex) == 24, "MeshVertex is Tuple<vec3, vec2, int32>");
