/*
 * slref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A C-ABI veneer (same shape as include/slcu.h, prefix slref_) over the
 * UNMODIFIED reference renderer compiled from the sources where they lie under
 * /root/reference (see oracle/Makefile).  No reference source is copied here:
 * this file only *calls* the reference's public API (SL_Context & friends,
 * softlight/include/softlight/SL_Context.hpp) and instantiates the shared
 * shader sources under shaders/ against the real ls::math.
 *
 * Built twice: libslref_strict.so (pinned strict flags, parity oracle, use
 * with num_threads(1)) and libslref_fast.so (the project's shipped flags,
 * timing baseline).  See SURVEY.md section 8c / appendix B.
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <atomic>

#include "lightsky/math/vec_utils.h"
#include "lightsky/math/mat_utils.h"
#include "lightsky/math/half.h"

#include "softlight/SL_Color.hpp"
#include "softlight/SL_Context.hpp"
#include "softlight/SL_Framebuffer.hpp"
#include "softlight/SL_IndexBuffer.hpp"
#include "softlight/SL_Mesh.hpp"
#include "softlight/SL_PackedVertex.hpp"
#include "softlight/SL_Sampler.hpp"
#include "softlight/SL_Shader.hpp"
#include "softlight/SL_Texture.hpp"
#include "softlight/SL_UniformBuffer.hpp"
#include "softlight/SL_VertexArray.hpp"
#include "softlight/SL_VertexBuffer.hpp"
#include "softlight/SL_ViewportState.hpp"
#include "softlight/SL_ImgFile.hpp"

#define SLCU_PREFIX slref_
#include <immintrin.h>
#include "slcu.h"

/* The only undefined symbol of the hot-path source subset. */
const void* SL_ImgFile::data() const { return nullptr; }

namespace math = ls::math;

/* --------------------------------------------------------------------------
 * Shader registry: name -> host function pointer, filled by the shared shader
 * sources.
 * ------------------------------------------------------------------------*/
namespace
{
typedef math::vec4 (*VsFn)(SL_VertexParam&);
typedef bool (*FsFn)(SL_FragmentParam&);

std::map<std::string, VsFn>& vs_registry() { static std::map<std::string, VsFn> r; return r; }
std::map<std::string, FsFn>& fs_registry() { static std::map<std::string, FsFn> r; return r; }

struct VsReg { VsReg(const char* n, VsFn f) { vs_registry()[n] = f; } };
struct FsReg { FsReg(const char* n, FsFn f) { fs_registry()[n] = f; } };

std::atomic<uint64_t> gFragsShaded{0};
bool gCountFrags = false;

/* debug aid: SL_FRAG_TRACE=<file> records the (x,y) of every fragment-shader call (1 thread only) */
std::vector<uint32_t> gFragTrace;
const char* gFragTracePath = getenv("SL_FRAG_TRACE");

/* Only fragments that land inside the w*h texels of the framebuffer are counted.  The reference addresses pixels as
 * x + w*y inside an allocation padded to ~4x (src/SL_Texture.cpp:28-54), so clip-generated triangles that end above
 * the viewport shade fragments nobody can read back; they are not part of the metric. */
size_t gFboW = 0, gFboH = 0;

template <FsFn fn>
bool counting_fs(SL_FragmentParam& p)
{
    if ((size_t)p.coord.x + gFboW * (size_t)p.coord.y < gFboW * gFboH) gFragsShaded.fetch_add(1, std::memory_order_relaxed);
    if (gFragTracePath) gFragTrace.push_back((uint32_t)p.coord.x | ((uint32_t)p.coord.y << 16));
    return fn(p);
}
struct FsCountReg { FsCountReg(const char* n, FsFn f) { fs_registry()[std::string("#count#") + n] = f; } };
} // namespace

#define SL_SHADER_FN static inline
#define SL_CAT2(a, b) a##b
#define SL_CAT(a, b) SL_CAT2(a, b)
#define SL_REGISTER_VERTEX_SHADER(name, fn) static VsReg SL_CAT(vsreg_, fn){name, fn};
#define SL_REGISTER_FRAGMENT_SHADER(name, fn) \
    static FsReg SL_CAT(fsreg_, fn){name, fn}; \
    static FsCountReg SL_CAT(fscreg_, fn){name, counting_fs<fn>};

#include "sl_shaders_basic.inl"
#include "sl_shaders_scene.inl"
#include "sl_shaders_anim_volume.inl"

/* --------------------------------------------------------------------------
 * Context wrapper
 * ------------------------------------------------------------------------*/
struct UboFixup
{
    size_t ubo;
    size_t offset;
    int kind; /* 0 texture, 1 buffer */
    size_t id;
    size_t byteOffset;
};

struct slcu_context
{
    SL_Context ctx;
    std::vector<UboFixup> fixups;
    std::string err;
    slcu_stats stats;
};

static void apply_fixups(slcu_context* c)
{
    for (const UboFixup& f : c->fixups)
    {
        unsigned char* p = c->ctx.ubo(f.ubo).buffer() + f.offset;
        const void* ptr;
        if (f.kind == 0)
        {
            ptr = &c->ctx.texture(f.id);
        }
        else
        {
            ptr = reinterpret_cast<const unsigned char*>(c->ctx.vbo(f.id).data()) + f.byteOffset;
        }
        memcpy(p, &ptr, sizeof(ptr));
    }
}

static size_t count_prims(const slcu_mesh* m, size_t n, size_t inst)
{
    size_t t = 0;
    for (size_t i = 0; i < n; ++i)
    {
        t += (m[i].elementEnd - m[i].elementBegin) / 3;
    }
    return t * inst;
}

static SL_Mesh to_mesh(const slcu_mesh& m)
{
    SL_Mesh r;
    r.vaoId = m.vaoId;
    r.elementBegin = m.elementBegin;
    r.elementEnd = m.elementEnd;
    r.mode = (SL_RenderMode)m.mode;
    r.materialId = m.materialId;
    return r;
}

extern "C" {

int slref_context_create(slcu_context** out, int)
{
    if (!out) return SLCU_ERR_ARG;
    slcu_context* c = new slcu_context{};
    c->ctx.num_threads(1);
    memset(&c->stats, 0, sizeof(c->stats));
    *out = c;
    return 0;
}

void slref_context_destroy(slcu_context* c) { delete c; }

unsigned slref_num_threads(slcu_context* c, unsigned n) { return c->ctx.num_threads(n); }

const char* slref_last_error(slcu_context* c) { return c->err.c_str(); }

const char* slref_impl_name(void)
{
#ifdef SLREF_FAST
    return "reference(SoftLight, shipped flags: -O2 -mavx2 -mfma -ffast-math)";
#else
    return "reference(SoftLight, strict flags: -O2 -mavx2 -mfma -ffp-contract=off)";
#endif
}

int slref_set_rcp_table(slcu_context*, const uint32_t*) { return SLCU_ERR_UNSUPPORTED; /* hardware rcpps */ }

int slref_get_rcp_table(slcu_context*, uint32_t* t)
{
    for (uint32_t i = 0; i < 2048; ++i)
    {
        union { uint32_t u; float f; } in, o;
        in.u = 0x3f800000u | (i << 12);
        o.f = math::rcp(in.f);
        t[i] = o.u;
    }
    return 0;
}

int slref_set_rsqrt_table(slcu_context*, const uint32_t*) { return SLCU_ERR_UNSUPPORTED; /* hardware rsqrtss */ }

/* the instruction math::inversesqrt is built on (light_math x86/scalarf_utils_impl.h:183-197), inputs in [1, 2) and [2, 4) */
int slref_get_rsqrt_table(slcu_context*, uint32_t* t)
{
    for (uint32_t p = 0; p < 2; ++p)
    {
        for (uint32_t i = 0; i < 2048; ++i)
        {
            union { uint32_t u; float f; } in, o;
            in.u = ((127u + p) << 23) | (i << 12);
            o.f = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(in.f)));
            t[p * 2048u + i] = o.u;
        }
    }
    return 0;
}

size_t slref_create_vbo(slcu_context* c) { return c->ctx.create_vbo(); }
size_t slref_create_ibo(slcu_context* c) { return c->ctx.create_ibo(); }
size_t slref_create_vao(slcu_context* c) { return c->ctx.create_vao(); }
size_t slref_create_ubo(slcu_context* c) { return c->ctx.create_ubo(); }
size_t slref_create_texture(slcu_context* c) { return c->ctx.create_texture(); }
size_t slref_create_framebuffer(slcu_context* c) { return c->ctx.create_framebuffer(); }

/* SL_Context::destroy_* of the unmodified reference.  A host program keeps real pointers in its uniform blocks; this shim
 * re-derives them from ids at every draw, so the ids of its own book-keeping follow the erase like the pointers would. */
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
int slref_destroy_vbo(slcu_context* c, size_t id) { c->ctx.destroy_vbo(id); remap_fixups(c, 1, id); return 0; }
int slref_destroy_ibo(slcu_context* c, size_t id) { c->ctx.destroy_ibo(id); return 0; }
int slref_destroy_vao(slcu_context* c, size_t id) { c->ctx.destroy_vao(id); return 0; }
int slref_destroy_ubo(slcu_context* c, size_t id)
{
    c->ctx.destroy_ubo(id);
    for (size_t i = 0; i < c->fixups.size();)
    {
        UboFixup& f = c->fixups[i];
        if (f.ubo == id) { c->fixups.erase(c->fixups.begin() + i); continue; }
        if (f.ubo > id) --f.ubo;
        ++i;
    }
    return 0;
}
int slref_destroy_texture(slcu_context* c, size_t id) { c->ctx.destroy_texture(id); remap_fixups(c, 0, id); return 0; }
int slref_destroy_framebuffer(slcu_context* c, size_t id) { c->ctx.destroy_framebuffer(id); return 0; }
int slref_destroy_shader(slcu_context* c, size_t id) { c->ctx.destroy_shader(id); return 0; }

int slref_vbo_init(slcu_context* c, size_t vbo, size_t n) { return c->ctx.vbo(vbo).init(n); }

int slref_vbo_assign(slcu_context* c, size_t vbo, const void* d, size_t off, size_t n)
{
    c->ctx.vbo(vbo).assign(d, (ptrdiff_t)off, n);
    return 0;
}

int slref_ibo_init(slcu_context* c, size_t ibo, uint32_t n, int type, const void* d)
{
    return c->ctx.ibo(ibo).init(n, (SL_DataType)type, d);
}

int slref_vao_set_vertex_buffer(slcu_context* c, size_t vao, size_t vbo) { c->ctx.vao(vao).set_vertex_buffer(vbo); return 0; }
int slref_vao_set_index_buffer(slcu_context* c, size_t vao, size_t ibo) { c->ctx.vao(vao).set_index_buffer(ibo); return 0; }
int slref_vao_set_num_bindings(slcu_context* c, size_t vao, int n) { return c->ctx.vao(vao).set_num_bindings((size_t)n); }

int slref_vao_set_binding(slcu_context* c, size_t vao, size_t bind, ptrdiff_t off, ptrdiff_t stride, int dim, int type)
{
    c->ctx.vao(vao).set_binding(bind, off, stride, (SL_Dimension)dim, (SL_DataType)type);
    return 0;
}

int slref_ubo_assign(slcu_context* c, size_t ubo, const void* d, size_t off, size_t n)
{
    if (off + n > 1024) return SLCU_ERR_ARG;
    memcpy(c->ctx.ubo(ubo).buffer() + off, d, n);
    return 0;
}

static void set_fixup(slcu_context* c, const UboFixup& f)
{
    for (UboFixup& g : c->fixups)
    {
        if (g.ubo == f.ubo && g.offset == f.offset) { g = f; return; }
    }
    c->fixups.push_back(f);
}

int slref_ubo_bind_texture(slcu_context* c, size_t ubo, size_t off, size_t tex)
{
    set_fixup(c, UboFixup{ubo, off, 0, tex, 0});
    return 0;
}

int slref_ubo_bind_buffer(slcu_context* c, size_t ubo, size_t off, size_t vbo, size_t byteOff)
{
    set_fixup(c, UboFixup{ubo, off, 1, vbo, byteOff});
    return 0;
}

int slref_texture_init(slcu_context* c, size_t tex, int type, uint16_t w, uint16_t h, uint16_t d)
{
    return c->ctx.texture(tex).init((SL_ColorDataType)type, w, h, d);
}

int slref_texture_set_texels(slcu_context* c, size_t tex, int order, uint16_t x, uint16_t y, uint16_t z, uint16_t w, uint16_t h, uint16_t d, const void* data)
{
    SL_Texture& t = c->ctx.texture(tex);
    if (order == 0) t.set_texels<SL_TexelOrder::ORDERED>(x, y, z, w, h, d, data);
    else t.set_texels<SL_TexelOrder::SWIZZLED>(x, y, z, w, h, d, data);
    return 0;
}

int slref_texture_read(slcu_context* c, size_t tex, void* out, size_t n)
{
    memcpy(out, c->ctx.texture(tex).data(), n);
    return 0;
}

int slref_fbo_reserve_color_buffers(slcu_context* c, size_t fbo, unsigned n) { return c->ctx.framebuffer(fbo).reserve_color_buffers(n); }
int slref_fbo_attach_color_buffer(slcu_context* c, size_t fbo, unsigned i, size_t tex) { return c->ctx.framebuffer(fbo).attach_color_buffer(i, c->ctx.texture(tex).view()); }
int slref_fbo_attach_depth_buffer(slcu_context* c, size_t fbo, size_t tex) { return c->ctx.framebuffer(fbo).attach_depth_buffer(c->ctx.texture(tex).view()); }
int slref_fbo_valid(slcu_context* c, size_t fbo) { return c->ctx.framebuffer(fbo).valid(); }

size_t slref_create_shader(slcu_context* c, const slcu_vertex_shader* vs, const slcu_fragment_shader* fs, size_t ubo)
{
    auto v = vs_registry().find(vs->shader);
    auto f = fs_registry().find(gCountFrags ? std::string("#count#") + fs->shader : std::string(fs->shader));
    if (v == vs_registry().end() || f == fs_registry().end())
    {
        c->err = "unknown shader name";
        return SLCU_INVALID_ID;
    }
    SL_VertexShader sv;
    sv.numVaryings = vs->numVaryings;
    sv.cullMode = (SL_CullMode)vs->cullMode;
    sv.shader = v->second;
    SL_FragmentShader sf;
    sf.numVaryings = fs->numVaryings;
    sf.numOutputs = fs->numOutputs;
    sf.blend = (SL_BlendMode)fs->blend;
    sf.depthTest = (SL_DepthTest)fs->depthTest;
    sf.depthMask = (SL_DepthMask)fs->depthMask;
    sf.shader = f->second;
    return c->ctx.create_shader(sv, sf, ubo);
}

int slref_viewport(slcu_context* c, int32_t x, int32_t y, uint16_t w, uint16_t h) { c->ctx.viewport_state().viewport(x, y, w, h); return 0; }
int slref_scissor(slcu_context* c, int32_t x, int32_t y, uint16_t w, uint16_t h) { c->ctx.viewport_state().scissor(x, y, w, h); return 0; }

int slref_clear_color_buffer(slcu_context* c, size_t fbo, unsigned att, const double rgba[4])
{
    c->ctx.clear_color_buffer(fbo, att, math::vec4_t<double>{rgba[0], rgba[1], rgba[2], rgba[3]});
    return 0;
}

int slref_clear_depth_buffer(slcu_context* c, size_t fbo, double depth)
{
    c->ctx.clear_depth_buffer(fbo, depth);
    return 0;
}

int slref_clear_framebuffer(slcu_context* c, size_t fbo, unsigned att, const double rgba[4], double depth)
{
    c->ctx.clear_framebuffer(fbo, att, math::vec4_t<double>{rgba[0], rgba[1], rgba[2], rgba[3]}, depth);
    return 0;
}

static void set_fbo_dims(slcu_context* c, size_t fbo)
{
    const SL_TextureView& d = c->ctx.framebuffer(fbo).get_depth_buffer();
    gFboW = d.width;
    gFboH = d.height;
}

int slref_draw(slcu_context* c, const slcu_mesh* m, size_t shader, size_t fbo)
{
    apply_fixups(c);
    c->stats.trisSubmitted += count_prims(m, 1, 1);
    c->stats.drawCalls++;
    set_fbo_dims(c, fbo);
    c->ctx.draw(to_mesh(*m), shader, fbo);
    return 0;
}

int slref_draw_multiple(slcu_context* c, const slcu_mesh* m, size_t n, size_t shader, size_t fbo)
{
    apply_fixups(c);
    std::vector<SL_Mesh> meshes;
    for (size_t i = 0; i < n; ++i) meshes.push_back(to_mesh(m[i]));
    c->stats.trisSubmitted += count_prims(m, n, 1);
    c->stats.drawCalls++;
    set_fbo_dims(c, fbo);
    c->ctx.draw_multiple(meshes.data(), n, shader, fbo);
    return 0;
}

int slref_draw_instanced(slcu_context* c, const slcu_mesh* m, size_t inst, size_t shader, size_t fbo)
{
    apply_fixups(c);
    c->stats.trisSubmitted += count_prims(m, 1, inst);
    c->stats.drawCalls++;
    set_fbo_dims(c, fbo);
    c->ctx.draw_instanced(to_mesh(*m), inst, shader, fbo);
    return 0;
}

int slref_blit(slcu_context* c, size_t out, size_t in) { c->ctx.blit(out, in); return 0; }

int slref_blit_rect(slcu_context* c, size_t out, size_t in, uint16_t sx0, uint16_t sy0, uint16_t sx1, uint16_t sy1, uint16_t dx0, uint16_t dy0, uint16_t dx1, uint16_t dy1)
{
    c->ctx.blit(out, in, sx0, sy0, sx1, sy1, dx0, dy0, dx1, dy1);
    return 0;
}

int slref_sync(slcu_context*) { return 0; }

/* Fragment counting costs an atomic per fragment: it is switched on per
 * context *before* create_shader (reset_stats(ctx) with ctx == NULL toggles
 * the global flag off; any non-null call switches it on). */
int slref_reset_stats(slcu_context* c)
{
    if (!c) { gCountFrags = false; return 0; }
    gCountFrags = true;
    gFragsShaded.store(0);
    gFragTrace.clear();
    memset(&c->stats, 0, sizeof(c->stats));
    return 0;
}

int slref_get_stats(slcu_context* c, slcu_stats* out)
{
    c->stats.fragsShaded = gFragsShaded.load();
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
int slref_timer_begin(slcu_context*) { g_t0 = now_ms(); return 0; }
int slref_timer_end(slcu_context*, double* ms) { *ms = now_ms() - g_t0; return 0; }
int slref_profile_begin(slcu_context*) { return 0; }
int slref_profile_end(slcu_context*, slcu_kernel_times* out) { memset(out, 0, sizeof(*out)); return 0; }

} // extern "C"
