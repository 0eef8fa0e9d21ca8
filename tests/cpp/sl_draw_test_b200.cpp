/*
 * sl_draw_test_b200.cpp -- softlight/tests/sl_draw_test.cpp:85-197 written against the C++ host mirror
 * (include/softlight_b200/SL_Context.hpp): same calls in the same order, the shader named instead of pointed to.
 * Prints the number of covered pixels and the centre pixel and checks them against what the unmodified reference
 * renders (38,278 covered pixels, centre 0xff40803f: SURVEY.md appendix B, tests/test_oracle_cpu.py).
 *
 *   g++ -std=c++11 -I include tests/cpp/sl_draw_test_b200.cpp -L softlight_b200 -lslcu -o sl_draw_test_b200
 *
 * Exit codes: 0 match, 1 mismatch, 77 no CUDA device (there is no CPU fallback).
 */
#include <cstdio>
#include <cstdlib>

#include "softlight_b200/SL_Context.hpp"

using namespace sl_b200;

struct ColoredVertex
{
    float pos[4];
    float color[4];
};

static SL_VertexShader line_vert_shader()
{
    SL_VertexShader shader;
    shader.numVaryings = 1;
    shader.cullMode = SL_CULL_BACK_FACE;
    shader.shader = "vcolor"; /* _line_vert_shader_impl: position + colour pass-through */
    return shader;
}

static SL_FragmentShader line_frag_shader()
{
    SL_FragmentShader shader;
    shader.numVaryings = 1;
    shader.numOutputs = 1;
    shader.blend = SL_BLEND_OFF;
    shader.depthTest = SL_DEPTH_TEST_OFF;
    shader.depthMask = SL_DEPTH_MASK_OFF;
    shader.shader = "vcolor"; /* _line_frag_shader_impl: output = interpolated colour */
    return shader;
}

int main()
{
    int retCode = 0;
    SL_Context context;
    if (!context.valid())
    {
        std::fprintf(stderr, "sl_draw_test_b200: %s (status %d)\n", context.last_error(), context.status());
        return 77;
    }

    retCode = (int)context.num_threads(1);
    if (retCode != 1) return 1;

    size_t fboId = context.create_framebuffer();
    size_t texId = context.create_texture();
    size_t depthId = context.create_texture();
    size_t vaoId = context.create_vao();
    size_t vboId = context.create_vbo();
    size_t iboId = context.create_ibo();
    size_t uboId = context.create_ubo();

    const SL_VertexShader&& vertShader = line_vert_shader();
    const SL_FragmentShader&& fragShader = line_frag_shader();
    size_t shaderId = context.create_shader(vertShader, fragShader, uboId);
    if (shaderId == (size_t)-1) { std::fprintf(stderr, "create_shader failed: %s\n", context.last_error()); return 1; }

    /* an unregistered callable is refused like the reference refuses an invalid shader */
    SL_FragmentShader bogus = fragShader;
    bogus.shader = "no_such_shader";
    if (context.create_shader(vertShader, bogus, uboId) != (size_t)-1) { std::fprintf(stderr, "unregistered shader was accepted\n"); return 1; }

    SL_VertexBuffer& vbo = context.vbo(vboId);
    ColoredVertex tri[3] = {
        {{-0.5f, -0.5f, 0.f, 1.f}, {1.f, 0.f, 0.f, 1.f}},
        {{ 0.f,   0.5f, 0.f, 1.f}, {0.f, 1.f, 0.f, 1.f}},
        {{ 0.5f, -0.5f, 0.f, 1.f}, {0.f, 0.f, 1.f, 1.f}}
    };
    const size_t numVboBytes = sizeof(tri);
    retCode = vbo.init(numVboBytes);
    if (retCode != 0) { std::fprintf(stderr, "Error while creating a VBO: %d\n", retCode); return 1; }
    vbo.assign(tri, 0, numVboBytes);

    SL_IndexBuffer& ibo = context.ibo(iboId);
    unsigned char indices[3] = {0, 2, 1};
    retCode = ibo.init(3, VERTEX_DATA_BYTE, indices);
    if (retCode != 0) { std::fprintf(stderr, "Error while creating an IBO: %d\n", retCode); return 1; }

    SL_VertexArray& vao = context.vao(vaoId);
    vao.set_vertex_buffer(vboId);
    vao.set_index_buffer(iboId);
    retCode = vao.set_num_bindings(2);
    if (retCode != 2) { std::fprintf(stderr, "Error while setting the number of VAO bindings: %d\n", retCode); return 1; }
    vao.set_binding(0, 0, sizeof(ColoredVertex), VERTEX_DIMENSION_4, VERTEX_DATA_FLOAT);
    vao.set_binding(1, 4 * sizeof(float), sizeof(ColoredVertex), VERTEX_DIMENSION_4, VERTEX_DATA_FLOAT);

    SL_Texture& tex = context.texture(texId);
    retCode = tex.init(SL_COLOR_RGBA_8U, 640, 480, 1);
    if (retCode != 0) { std::fprintf(stderr, "Error while creating a color texture: %d\n", retCode); return 1; }
    SL_Texture& depth = context.texture(depthId);
    retCode = depth.init(SL_COLOR_R_FLOAT, 640, 480, 1);
    if (retCode != 0) { std::fprintf(stderr, "Error while creating a depth texture: %d\n", retCode); return 1; }

    SL_Framebuffer& fbo = context.framebuffer(fboId);
    retCode = fbo.reserve_color_buffers(1);
    if (retCode != 0) { std::fprintf(stderr, "Error while reserving FBO color buffers: %d\n", retCode); return 1; }
    retCode = fbo.attach_color_buffer(0, tex.view());
    if (retCode != 0) { std::fprintf(stderr, "Error while attaching a color buffer to an FBO: %d\n", retCode); return 1; }
    retCode = fbo.attach_depth_buffer(depth.view());
    if (retCode != 0) { std::fprintf(stderr, "Error while attaching a depth buffer to an FBO: %d\n", retCode); return 1; }
    if (fbo.valid() != 0) { std::fprintf(stderr, "framebuffer not valid: %d\n", fbo.valid()); return 1; }
    context.clear_framebuffer(fboId, 0, SL_ColorRGBAd{{1.0, 0.0, 1.0, 1.0}}, 0.0);

    SL_Mesh m;
    m.elementBegin = 0;
    m.elementEnd = ibo.count();
    m.vaoId = vaoId;
    m.mode = RENDER_MODE_INDEXED_TRIANGLES;
    m.materialId = 0;

    context.num_threads(8);
    context.draw(m, shaderId, fboId);
    if (context.status() != 0) { std::fprintf(stderr, "draw failed: %s\n", context.last_error()); return 1; }

    const unsigned char* px = static_cast<const unsigned char*>(tex.data());
    if (!px) { std::fprintf(stderr, "read-back failed: %s\n", context.last_error()); return 1; }
    unsigned covered = 0;
    for (size_t i = 0; i < 640u * 480u; ++i)
    {
        if (!(px[4 * i] == 255 && px[4 * i + 1] == 0 && px[4 * i + 2] == 255 && px[4 * i + 3] == 255)) ++covered;
    }
    const unsigned char* c = px + 4 * (320 + 640 * 240);
    std::printf("covered=%u centre=%02x%02x%02x%02x\n", covered, c[0], c[1], c[2], c[3]);
    const bool ok = covered == 38278u && c[0] == 0x3f && c[1] == 0x80 && c[2] == 0x40 && c[3] == 0xff;
    std::printf("%s\n", ok ? "MATCH" : "MISMATCH");
    return ok ? 0 : 1;
}
