/*
 * sl_instancing_test_b200.cpp -- the pattern of softlight/tests/sl_instancing_test.cpp:173-290 written against the C++
 * host mirror (include/softlight_b200/SL_Context.hpp): an indexed mesh drawn with draw_instanced, per-instance matrices
 * reached through a pointer inside the uniform block, a sampled texture reached through another, and -- like the
 * reference's program (:236, 251-274) -- ONE pointer into the uniform block taken up front and written through between
 * draws.  The instance matrices are changed through a cached data() pointer as well.
 *
 * The program dumps every input it used and the picture it got into the directory given as argv[1]; the test
 * (tests/test_cpp_host.py) replays the same inputs through the CPU oracle and compares the bytes.
 *
 *   g++ -std=c++11 -I include tests/cpp/sl_instancing_test_b200.cpp -L softlight_b200 -lslcu -o sl_instancing_test_b200
 *
 * Exit codes: 0 rendered and dumped, 1 error, 77 no CUDA device (there is no CPU fallback).
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "softlight_b200/SL_Context.hpp"

using namespace sl_b200;

struct Vertex { float pos[3]; float uv[2]; float nrm[3]; };        /* Tuple<vec3, vec2, vec3>, 32 bytes */
struct Mat4 { float m[16]; };                                        /* column-major like ls::math::mat4 */
struct InstTexUniforms                                               /* shaders/sl_shaders_scene.inl */
{
    const void* pTexture;     /* declared with bind_texture */
    const void* pad0_;
    float ambient[4];
    float diffuse[4];
    const void* instanceMatrix; /* declared with bind_buffer */
    const void* pad1_;
    Mat4 modelMatrix;
    Mat4 vpMatrix;
    float camPos[4];
};
static_assert(sizeof(InstTexUniforms) == 208, "layout of the shader's uniform block");

static Mat4 identity() { Mat4 r; std::memset(&r, 0, sizeof(r)); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.f; return r; }
static Mat4 translate_scale(float x, float y, float z, float s) { Mat4 r = identity(); r.m[0] = r.m[5] = r.m[10] = s; r.m[12] = x; r.m[13] = y; r.m[14] = z; return r; }
/* infinite reversed-Z perspective looking down -z, as the demos use */
static Mat4 projection(float fovy, float aspect, float znear)
{
    Mat4 r; std::memset(&r, 0, sizeof(r));
    const float f = 1.f / std::tan(fovy * 0.5f);
    r.m[0] = f / aspect; r.m[5] = f; r.m[11] = -1.f; r.m[14] = znear;
    return r;
}

static void dump(const std::string& dir, const char* name, const void* p, size_t n)
{
    FILE* f = std::fopen((dir + "/" + name).c_str(), "wb");
    if (!f) { std::perror(name); std::exit(1); }
    std::fwrite(p, 1, n, f);
    std::fclose(f);
}

int main(int argc, char** argv)
{
    const std::string dir = argc > 1 ? argv[1] : ".";
    SL_Context context;
    if (!context.valid())
    {
        std::fprintf(stderr, "sl_instancing_test_b200: %s (status %d)\n", context.last_error(), context.status());
        return 77;
    }
    const uint16_t W = 320, H = 200;

    /* a cube: 24 vertices (4 per face, own normal), 12 triangles, u16 indices */
    std::vector<Vertex> verts;
    std::vector<uint16_t> idx;
    const float n[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    for (int f = 0; f < 6; ++f)
    {
        const float* nn = n[f];
        const float u[3] = {nn[1], nn[2], nn[0]}, v[3] = {nn[1] * u[2] - nn[2] * u[1], nn[2] * u[0] - nn[0] * u[2], nn[0] * u[1] - nn[1] * u[0]};
        const float c[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
        const uint16_t base = (uint16_t)verts.size();
        for (int k = 0; k < 4; ++k)
        {
            Vertex vx;
            for (int a = 0; a < 3; ++a) { vx.pos[a] = nn[a] + c[k][0] * u[a] + c[k][1] * v[a]; vx.nrm[a] = nn[a]; }
            vx.uv[0] = 0.5f * (c[k][0] + 1.f) * 2.f;
            vx.uv[1] = 0.5f * (c[k][1] + 1.f) * 2.f;
            verts.push_back(vx);
        }
        const uint16_t q[6] = {0, 1, 2, 0, 2, 3};
        for (int k = 0; k < 6; ++k) idx.push_back((uint16_t)(base + q[k]));
    }

    const size_t numInstances = 27;
    std::vector<Mat4> mats(numInstances), mats2(numInstances);
    for (size_t i = 0; i < numInstances; ++i)
    {
        const float x = (float)(i % 3) - 1.f, y = (float)((i / 3) % 3) - 1.f, z = (float)(i / 9);
        mats[i] = translate_scale(4.f * x, 4.f * y, -12.f - 5.f * z, 1.f + 0.1f * (float)(i % 4));
        mats2[i] = translate_scale(4.f * x + 2.f, 4.f * y + 2.f, -10.f - 5.f * z, 0.6f);
    }

    std::vector<unsigned char> texels(16 * 16 * 3);
    for (int y = 0; y < 16; ++y)
        for (int x = 0; x < 16; ++x)
        {
            unsigned char* t = &texels[3 * (x + 16 * y)];
            t[0] = (unsigned char)(16 * x + 8); t[1] = (unsigned char)(16 * y + 8); t[2] = (unsigned char)(((x ^ y) & 1) ? 230 : 40);
        }

    const size_t fboId = context.create_framebuffer(), colorId = context.create_texture(), depthId = context.create_texture(), albedoId = context.create_texture();
    const size_t vaoId = context.create_vao(), vboId = context.create_vbo(), iboId = context.create_ibo(), instId = context.create_vbo(), uboId = context.create_ubo();

    SL_VertexBuffer& vbo = context.vbo(vboId);
    if (vbo.init(verts.size() * sizeof(Vertex)) != 0) return 1;
    vbo.assign(verts.data(), 0, verts.size() * sizeof(Vertex));
    SL_IndexBuffer& ibo = context.ibo(iboId);
    if (ibo.init((uint32_t)idx.size(), VERTEX_DATA_SHORT, idx.data()) != 0) return 1;
    SL_VertexArray& vao = context.vao(vaoId);
    vao.set_vertex_buffer(vboId);
    vao.set_index_buffer(iboId);
    if (vao.set_num_bindings(3) != 3) return 1;
    vao.set_binding(0, 0, sizeof(Vertex), VERTEX_DIMENSION_3, VERTEX_DATA_FLOAT);
    vao.set_binding(1, 12, sizeof(Vertex), VERTEX_DIMENSION_2, VERTEX_DATA_FLOAT);
    vao.set_binding(2, 20, sizeof(Vertex), VERTEX_DIMENSION_3, VERTEX_DATA_FLOAT);

    /* the per-instance matrices: the reference keeps them in host memory and points at them from the uniform block */
    SL_VertexBuffer& inst = context.vbo(instId);
    if (inst.init(numInstances * sizeof(Mat4)) != 0) return 1;
    Mat4* const pInstances = static_cast<Mat4*>(inst.data()); /* taken ONCE, written through below */
    std::memcpy(pInstances, mats.data(), numInstances * sizeof(Mat4));

    if (context.texture(colorId).init(SL_COLOR_RGBA_8U, W, H, 1) != 0) return 1;
    if (context.texture(depthId).init(SL_COLOR_R_HALF, W, H, 1) != 0) return 1;
    SL_Texture& albedo = context.texture(albedoId);
    if (albedo.init(SL_COLOR_RGB_8U, 16, 16, 1) != 0) return 1;
    if (albedo.set_texels(0, 0, 0, 16, 16, 1, texels.data()) != 0) return 1;

    SL_Framebuffer& fbo = context.framebuffer(fboId);
    if (fbo.reserve_color_buffers(1) != 0 || fbo.attach_color_buffer(0, context.texture(colorId).view()) != 0 || fbo.attach_depth_buffer(context.texture(depthId).view()) != 0) return 1;
    if (fbo.valid() != 0) { std::fprintf(stderr, "framebuffer not valid: %d\n", fbo.valid()); return 1; }

    SL_UniformBuffer& ubo = context.ubo(uboId);
    InstTexUniforms* const pUniforms = ubo.as<InstTexUniforms>(); /* taken ONCE (sl_instancing_test.cpp:236) */
    std::memset(pUniforms, 0, sizeof(*pUniforms));
    const float ambient[4] = {0.2f, 0.2f, 0.2f, 1.f}, diffuse[4] = {0.8f, 0.8f, 0.8f, 0.f}, cam[4] = {0.f, 0.f, 6.f, 1.f};
    std::memcpy(pUniforms->ambient, ambient, 16);
    std::memcpy(pUniforms->diffuse, diffuse, 16);
    std::memcpy(pUniforms->camPos, cam, 16);
    pUniforms->modelMatrix = identity();
    Mat4 view = translate_scale(-cam[0], -cam[1], -cam[2], 1.f);
    Mat4 proj = projection(1.0471976f, (float)W / (float)H, 0.01f);
    Mat4 vp;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r)
        {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += proj.m[k * 4 + r] * view.m[c * 4 + k];
            vp.m[c * 4 + r] = s;
        }
    pUniforms->vpMatrix = vp;
    if (ubo.bind_texture(offsetof(InstTexUniforms, pTexture), albedoId) != 0) return 1;
    if (ubo.bind_buffer(offsetof(InstTexUniforms, instanceMatrix), instId, 0) != 0) return 1;

    SL_VertexShader vs; vs.numVaryings = 3; vs.cullMode = SL_CULL_BACK_FACE; vs.shader = "inst_tex";
    SL_FragmentShader fs; fs.numVaryings = 3; fs.numOutputs = 1; fs.blend = SL_BLEND_PREMULTIPLED_ALPHA; fs.depthTest = SL_DEPTH_TEST_GREATER_EQUAL; fs.depthMask = SL_DEPTH_MASK_ON; fs.shader = "inst_tex";
    const size_t shaderId = context.create_shader(vs, fs, uboId);
    if (shaderId == (size_t)-1) { std::fprintf(stderr, "create_shader failed: %s\n", context.last_error()); return 1; }

    SL_Mesh m;
    m.vaoId = vaoId; m.elementBegin = 0; m.elementEnd = ibo.count(); m.mode = RENDER_MODE_INDEXED_TRIANGLES; m.materialId = 0;

    context.clear_framebuffer(fboId, 0, SL_ColorRGBAd{{0.6, 0.6, 0.6, 1.0}}, 0.0);
    InstTexUniforms blob1 = *pUniforms;
    context.draw_instanced(m, numInstances, shaderId, fboId);
    if (context.status() != 0) { std::fprintf(stderr, "draw failed: %s\n", context.last_error()); return 1; }

    /* second draw: everything changes through the pointers taken before the first one */
    const float diffuse2[4] = {0.3f, 0.9f, 0.4f, 0.f};
    std::memcpy(pUniforms->diffuse, diffuse2, 16);
    pUniforms->modelMatrix = translate_scale(0.f, 0.f, 0.f, 0.75f);
    std::memcpy(pInstances, mats2.data(), numInstances * sizeof(Mat4));
    InstTexUniforms blob2 = *pUniforms;
    context.draw_instanced(m, numInstances, shaderId, fboId);
    if (context.status() != 0) { std::fprintf(stderr, "second draw failed: %s\n", context.last_error()); return 1; }

    const void* px = context.texture(colorId).data();
    if (!px) { std::fprintf(stderr, "read-back failed: %s\n", context.last_error()); return 1; }
    uint32_t rcp[2048];
    if (slcu_get_rcp_table(context.native_handle(), rcp) != 0) return 1;

    dump(dir, "verts.bin", verts.data(), verts.size() * sizeof(Vertex));
    dump(dir, "idx.bin", idx.data(), idx.size() * 2);
    dump(dir, "mats1.bin", mats.data(), mats.size() * sizeof(Mat4));
    dump(dir, "mats2.bin", mats2.data(), mats2.size() * sizeof(Mat4));
    dump(dir, "tex.bin", texels.data(), texels.size());
    dump(dir, "blob1.bin", &blob1, sizeof(blob1));
    dump(dir, "blob2.bin", &blob2, sizeof(blob2));
    dump(dir, "rcp.bin", rcp, sizeof(rcp));
    dump(dir, "color.bin", px, (size_t)W * H * 4);
    std::printf("rendered %ux%u, %zu instances x %zu triangles, two draws\n", (unsigned)W, (unsigned)H, numInstances, idx.size() / 3);
    return 0;
}
