/*
 * slcu.h -- the C-ABI drop-in boundary of softlight-b200.
 *
 * SoftLight (the reference) has no C API of its own (its README lists "Wrap the
 * rendering module in a C API" as a TODO, /root/reference/README.md:164); its
 * only entry points into parallel work are SL_ProcessorPool::run_shader_processors,
 * run_clear_processors and run_blit_processors, reached from SL_Context
 * (softlight/src/SL_ProcessorPool.cpp:280-753).  This header is the thin C ABI
 * those three bodies are replaced by, plus the resource calls SL_Context needs
 * to mirror its objects on the device.  Every entry point cites the reference
 * interface it stands in for.  Plain pointers and sizes only; every function
 * returns an int status (0 ok, negative = failure) or an id, and never throws.
 *
 * Numeric enum values are the reference's own (SL_ColorDataType, SL_DataType,
 * SL_Dimension, SL_RenderMode, SL_CullMode, SL_DepthTest, SL_DepthMask,
 * SL_BlendMode), so a binding can pass them through unchanged.
 *
 * The same declarations, with SLCU_PREFIX set to slref_ / slor_, describe the
 * two test-only checker libraries under oracle/ (the compiled reference and
 * the CPU restatement); the product library is libslcu.so with the slcu_ prefix.
 */
#ifndef SLCU_H
#define SLCU_H

#include <stddef.h>
#include <stdint.h>

#ifndef SLCU_PREFIX
#define SLCU_PREFIX slcu_
#endif
#define SLCU_CAT2(a, b) a##b
#define SLCU_CAT(a, b) SLCU_CAT2(a, b)
#define SLCU_FN(name) SLCU_CAT(SLCU_PREFIX, name)

#if defined(__GNUC__)
#define SLCU_EXPORT __attribute__((visibility("default")))
#else
#define SLCU_EXPORT
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct slcu_context slcu_context;

#define SLCU_INVALID_ID ((size_t)-1)

/* Error codes (all negative). */
enum
{
    SLCU_OK              = 0,
    SLCU_ERR_ARG         = -1,   /* bad id / null pointer / out of range        */
    SLCU_ERR_CUDA        = -100, /* a CUDA runtime call failed                  */
    SLCU_ERR_NO_DEVICE   = -101, /* no CUDA device: there is NO CPU fallback    */
    SLCU_ERR_NO_SHADER   = -102, /* shader name not in the device registry      */
    SLCU_ERR_UNSUPPORTED = -103, /* state combination not implemented           */
    SLCU_ERR_OOM         = -104
};

/* SL_Mesh (softlight/include/softlight/SL_Mesh.hpp:30-39), same field order. */
typedef struct slcu_mesh
{
    uint64_t vaoId;
    uint64_t elementBegin;
    uint64_t elementEnd;
    uint32_t mode;       /* SL_RenderMode: 0x10/0x30 triangles, 0x50/0x70 wireframe triangles, 0x01/0x03 points, 0x04/0x0C lines */
    uint32_t materialId;
} slcu_mesh;

/* SL_VertexShader / SL_FragmentShader (SL_Shader.hpp:58-63, 88-97).  The
 * reference stores a raw C function pointer; the device cannot call it, so the
 * `shader` member is the registry name of the __device__ callable that was
 * registered for that function (SL_REGISTER_VERTEX_SHADER / _FRAGMENT_SHADER,
 * softlight_b200/csrc/slcu.cu).  An unknown name makes
 * create_shader return SLCU_INVALID_ID, like the reference's validation
 * failures (softlight/src/SL_Context.cpp:488-540). */
typedef struct slcu_vertex_shader
{
    uint8_t numVaryings;
    uint8_t cullMode; /* SL_CullMode */
    const char* shader;
} slcu_vertex_shader;

typedef struct slcu_fragment_shader
{
    uint8_t numVaryings;
    uint8_t numOutputs;
    uint8_t blend;     /* SL_BlendMode */
    uint8_t depthTest; /* SL_DepthTest */
    uint8_t depthMask; /* SL_DepthMask */
    const char* shader;
} slcu_fragment_shader;

/* Counters of the last draw call(s) since reset_stats (ours; the reference
 * keeps none).  Used by bench.py for Mtris/s and Gfragments/s. */
typedef struct slcu_stats
{
    uint64_t trisSubmitted;   /* primitives fed to the vertex stage             */
    uint64_t trisBinned;      /* triangle records that reached the rasteriser   */
    uint64_t spans;           /* non-empty (triangle,row) spans in rows < height */
    uint64_t fragsShaded;     /* fragment-shader invocations that land inside
                                 the width*height texels of the framebuffer    */
    uint64_t kernelLaunches;  /* CUDA kernels launched by this context          */
    uint64_t drawCalls;
    uint64_t batches;         /* pipeline passes the draw calls were executed in (CUDA: consecutive draws with one
                                 shader + framebuffer share a pass; CPU checkers: == drawCalls)                 */
} slcu_stats;

/* ---- context: SL_Context ctor/dtor, num_threads (SL_Context.cpp:1135-1148) */
SLCU_EXPORT int  SLCU_FN(context_create)(slcu_context** out, int device);
SLCU_EXPORT void SLCU_FN(context_destroy)(slcu_context* ctx);
SLCU_EXPORT unsigned SLCU_FN(num_threads)(slcu_context* ctx, unsigned n);
SLCU_EXPORT const char* SLCU_FN(last_error)(slcu_context* ctx);
SLCU_EXPORT const char* SLCU_FN(impl_name)(void);

/* rcpss/rcpps table.  2048 entries: the 11 low-zero-padded
 * 12-bit result mantissas (bits 22..11 of the result, as result_bits >> 11 of
 * the [1,2) input binade... see softlight_b200/csrc/sl_x86emu.cuh).  By default
 * the table is measured from the host CPU at context_create so results match
 * the reference running on the same box; tests pin it to the committed one. */
SLCU_EXPORT int SLCU_FN(set_rcp_table)(slcu_context* ctx, const uint32_t* table2048);
SLCU_EXPORT int SLCU_FN(get_rcp_table)(slcu_context* ctx, uint32_t* table2048);
/* rsqrtss table (line primitives: math::inversesqrt is rsqrtss plus one Newton step, light_math
 * x86/scalarf_utils_impl.h:183-197).  4096 entries: the instruction's results for the inputs
 * 2^p * (1 + i / 2048), p = 0, 1 -- exponent parity and the top 11 mantissa bits decide the result,
 * the exponent is handled exactly.  Measured from the host CPU at context_create like the rcp table. */
SLCU_EXPORT int SLCU_FN(set_rsqrt_table)(slcu_context* ctx, const uint32_t* table4096);
SLCU_EXPORT int SLCU_FN(get_rsqrt_table)(slcu_context* ctx, uint32_t* table4096);

/* ---- resources: SL_Context::create_* (SL_Context.cpp:160-460) */
SLCU_EXPORT size_t SLCU_FN(create_vbo)(slcu_context* ctx);
SLCU_EXPORT size_t SLCU_FN(create_ibo)(slcu_context* ctx);
SLCU_EXPORT size_t SLCU_FN(create_vao)(slcu_context* ctx);
SLCU_EXPORT size_t SLCU_FN(create_ubo)(slcu_context* ctx);
SLCU_EXPORT size_t SLCU_FN(create_texture)(slcu_context* ctx);
SLCU_EXPORT size_t SLCU_FN(create_framebuffer)(slcu_context* ctx);

/* SL_Context::destroy_* (softlight/src/SL_Context.cpp:192-195, 243-247, 295-298, 346-349, 397-400, 448-451, 612-615): the
 * object is erased from its vector, so every later id of that kind moves down by one -- exactly the reference's
 * behaviour, which callers written against it already account for.  What the reference holds by POINTER keeps pointing
 * at the same object here too: framebuffer attachments and declared uniform pointers (ubo_bind_*) follow a texture /
 * vertex buffer whose id moved, and lose it when it is the one destroyed.  What the reference holds by INDEX (a vertex
 * array's buffer ids, a shader's uniform block) is left alone, like there.  Returns 0, or SLCU_ERR_ARG for an id that
 * does not exist (undefined behaviour in the reference). */
SLCU_EXPORT int SLCU_FN(destroy_vbo)(slcu_context* ctx, size_t id);
SLCU_EXPORT int SLCU_FN(destroy_ibo)(slcu_context* ctx, size_t id);
SLCU_EXPORT int SLCU_FN(destroy_vao)(slcu_context* ctx, size_t id);
SLCU_EXPORT int SLCU_FN(destroy_ubo)(slcu_context* ctx, size_t id);
SLCU_EXPORT int SLCU_FN(destroy_texture)(slcu_context* ctx, size_t id);
SLCU_EXPORT int SLCU_FN(destroy_framebuffer)(slcu_context* ctx, size_t id);
SLCU_EXPORT int SLCU_FN(destroy_shader)(slcu_context* ctx, size_t id);

/* SL_VertexBuffer::init / assign (SL_VertexBuffer.hpp:45-49, src/SL_VertexBuffer.cpp) */
SLCU_EXPORT int SLCU_FN(vbo_init)(slcu_context* ctx, size_t vbo, size_t numBytes);
SLCU_EXPORT int SLCU_FN(vbo_assign)(slcu_context* ctx, size_t vbo, const void* data, size_t offset, size_t numBytes);

/* SL_IndexBuffer::init(count, type, data) (src/SL_IndexBuffer.cpp:117-170) */
SLCU_EXPORT int SLCU_FN(ibo_init)(slcu_context* ctx, size_t ibo, uint32_t numElements, int dataType, const void* data);

/* SL_VertexArray (SL_VertexArray.hpp:40-100; set_num_bindings returns n or -1) */
SLCU_EXPORT int SLCU_FN(vao_set_vertex_buffer)(slcu_context* ctx, size_t vao, size_t vbo);
SLCU_EXPORT int SLCU_FN(vao_set_index_buffer)(slcu_context* ctx, size_t vao, size_t ibo);
SLCU_EXPORT int SLCU_FN(vao_set_num_bindings)(slcu_context* ctx, size_t vao, int numBindings);
SLCU_EXPORT int SLCU_FN(vao_set_binding)(slcu_context* ctx, size_t vao, size_t bindId, ptrdiff_t offset, ptrdiff_t stride, int dimension, int dataType);

/* SL_UniformBuffer::assign (SL_UniformBuffer.hpp:96-110).  The reference lets
 * shaders chase host pointers stored inside the blob (textures, instance
 * matrices, bones).  bind_texture / bind_buffer declare such an 8-byte slot:
 * at draw time it holds a pointer to the texture object / into the buffer that
 * is valid where the shader runs. */
SLCU_EXPORT int SLCU_FN(ubo_assign)(slcu_context* ctx, size_t ubo, const void* data, size_t offset, size_t numBytes);
SLCU_EXPORT int SLCU_FN(ubo_bind_texture)(slcu_context* ctx, size_t ubo, size_t offset, size_t texture);
SLCU_EXPORT int SLCU_FN(ubo_bind_buffer)(slcu_context* ctx, size_t ubo, size_t offset, size_t vbo, size_t byteOffset);

/* SL_Texture::init(type,w,h,d) (src/SL_Texture.cpp:276-294), set_texels
 * <ORDERED|SWIZZLED> (SL_Texture.hpp:418-444; texelOrder 0 = ORDERED,
 * 1 = SWIZZLED), data() read-back of the w*h*d*bpt leading bytes.
 * colorType is any of the 33 SL_ColorDataType values (SL_Color.hpp:29-70):
 * 0..27 the plain layouts, 28..32 the packed RGB332 / RGB565 / RGBA5551 /
 * RGBA4444 / RGBA1010102 (SL_ColorCompressed.hpp), as textures, blit sources
 * and destinations, and colour attachments. */
SLCU_EXPORT int SLCU_FN(texture_init)(slcu_context* ctx, size_t tex, int colorType, uint16_t w, uint16_t h, uint16_t d);
SLCU_EXPORT int SLCU_FN(texture_set_texels)(slcu_context* ctx, size_t tex, int texelOrder, uint16_t x, uint16_t y, uint16_t z, uint16_t w, uint16_t h, uint16_t d, const void* data);
SLCU_EXPORT int SLCU_FN(texture_read)(slcu_context* ctx, size_t tex, void* out, size_t numBytes);

/* ours: page-locked host memory.  The reference's textures and window back buffers are plain host allocations
 * (src/SL_Texture.cpp:276-294, SL_WindowBuffer); a host that wants texture_read / vbo_assign to run at the link rate hands
 * the library buffers from here instead (any host pointer is accepted everywhere; pageable ones are just slower). */
SLCU_EXPORT int SLCU_FN(host_alloc)(slcu_context* ctx, void** out, size_t numBytes);
SLCU_EXPORT int SLCU_FN(host_free)(slcu_context* ctx, void* ptr);

/* SL_Framebuffer (src/SL_Framebuffer.cpp:823-990) */
SLCU_EXPORT int SLCU_FN(fbo_reserve_color_buffers)(slcu_context* ctx, size_t fbo, unsigned n);
SLCU_EXPORT int SLCU_FN(fbo_attach_color_buffer)(slcu_context* ctx, size_t fbo, unsigned index, size_t texture);
SLCU_EXPORT int SLCU_FN(fbo_attach_depth_buffer)(slcu_context* ctx, size_t fbo, size_t texture);
SLCU_EXPORT int SLCU_FN(fbo_valid)(slcu_context* ctx, size_t fbo);

/* Run-time shader registration (ours; stands where the reference takes a host function pointer, SL_Shader.hpp:63, 97).
 * `source` is CUDA C++ in the dialect of shaders/(*).inl: SL_SHADER_FN functions over the device mirrors of ls::math and
 * the SL_* shader API, and SL_REGISTER_VERTEX_SHADER("name", fn) / SL_REGISTER_FRAGMENT_SHADER("name", fn) lines.
 * NVRTC instantiates the pipeline's kernel templates for every registered function against the headers the library was
 * built from; the names can then be used in create_shader (a run-time name shadows a built-in one).
 * register_shader_source returns the number of shaders registered or a negative error (compiler log in last_error);
 * shader_source_check only compiles (no device needed) and copies the log out. */
SLCU_EXPORT int SLCU_FN(register_shader_source)(slcu_context* ctx, const char* source);
SLCU_EXPORT int SLCU_FN(shader_source_check)(const char* source, char* log, size_t logBytes);

/* SL_Context::create_shader (src/SL_Context.cpp:488-605) */
SLCU_EXPORT size_t SLCU_FN(create_shader)(slcu_context* ctx, const slcu_vertex_shader* vs, const slcu_fragment_shader* fs, size_t ubo);

/* SL_ViewportState::viewport / scissor (SL_ViewportState.hpp:50-62) */
SLCU_EXPORT int SLCU_FN(viewport)(slcu_context* ctx, int32_t x, int32_t y, uint16_t w, uint16_t h);
SLCU_EXPORT int SLCU_FN(scissor)(slcu_context* ctx, int32_t x, int32_t y, uint16_t w, uint16_t h);

/* SL_Context::clear_* -> SL_ProcessorPool::run_clear_processors
 * (src/SL_Context.cpp:881-963, src/SL_ProcessorPool.cpp:452-573) */
SLCU_EXPORT int SLCU_FN(clear_color_buffer)(slcu_context* ctx, size_t fbo, unsigned attachment, const double rgba[4]);
SLCU_EXPORT int SLCU_FN(clear_depth_buffer)(slcu_context* ctx, size_t fbo, double depth);
SLCU_EXPORT int SLCU_FN(clear_framebuffer)(slcu_context* ctx, size_t fbo, unsigned attachment, const double rgba[4], double depth);

/* SL_Context::draw / draw_multiple / draw_instanced -> run_shader_processors
 * (src/SL_Context.cpp:714-740, src/SL_ProcessorPool.cpp:280-380) */
SLCU_EXPORT int SLCU_FN(draw)(slcu_context* ctx, const slcu_mesh* mesh, size_t shader, size_t fbo);
SLCU_EXPORT int SLCU_FN(draw_multiple)(slcu_context* ctx, const slcu_mesh* meshes, size_t numMeshes, size_t shader, size_t fbo);
SLCU_EXPORT int SLCU_FN(draw_instanced)(slcu_context* ctx, const slcu_mesh* mesh, size_t numInstances, size_t shader, size_t fbo);

/* SL_Context::blit(out, in[, rects]) -> run_blit_processors /
 * run_blit_compressed_processors (src/SL_Context.cpp:747-813,
 * src/SL_BlitProcessor.cpp:620-715, src/SL_BlitCompressedProcessor.cpp):
 * nearest neighbour, vertically flipped, converting between any two of the
 * 33 colour types the way color_cast / rgb_cast / rgba_cast do. */
SLCU_EXPORT int SLCU_FN(blit)(slcu_context* ctx, size_t outTexture, size_t inTexture);
SLCU_EXPORT int SLCU_FN(blit_rect)(slcu_context* ctx, size_t outTexture, size_t inTexture,
                                   uint16_t srcX0, uint16_t srcY0, uint16_t srcX1, uint16_t srcY1,
                                   uint16_t dstX0, uint16_t dstY0, uint16_t dstX1, uint16_t dstY1);

/* SL_Context::blit(SL_TextureView& buffer, size_t textureId[, rectangles]) (src/SL_Context.cpp:816-874): the blit into a
 * host-side view -- the window back buffer SL_Swapchain / SL_WindowBuffer hand out.  `texels` points at w * h texels of
 * colorType (page-locked memory from host_alloc crosses the link at full rate).  Synchronous like the reference's. */
SLCU_EXPORT int SLCU_FN(blit_to_host)(slcu_context* ctx, void* texels, int colorType, uint16_t w, uint16_t h, size_t inTexture);
SLCU_EXPORT int SLCU_FN(blit_to_host_rect)(slcu_context* ctx, void* texels, int colorType, uint16_t w, uint16_t h, size_t inTexture,
                                           uint16_t srcX0, uint16_t srcY0, uint16_t srcX1, uint16_t srcY1,
                                           uint16_t dstX0, uint16_t dstY0, uint16_t dstX1, uint16_t dstY1);
/* ours: the same blit as a two-deep swap chain.  present(slot) returns when the blit is enqueued -- the next frame renders
 * while this one crosses the link -- present_wait(slot) returns when the host view holds the frame. */
SLCU_EXPORT int SLCU_FN(present)(slcu_context* ctx, unsigned slot, void* texels, int colorType, uint16_t w, uint16_t h, size_t inTexture);
SLCU_EXPORT int SLCU_FN(present_wait)(slcu_context* ctx, unsigned slot);

/* ---- sort-first multi-GPU: ours (the reference is one process; its analogue is the per-thread scan-line interleave of
 * SL_ScanlineBounds.hpp:31-50 / src/SL_TriRasterizer.cpp:436-438).  One process and one context per GPU.
 * set_partition: this context rasterises only rows y with (y / stripRows) % world == rank (geometry and uniforms are
 *   replicated by issuing the same calls on every rank).
 * texture_ipc_export/import: a 64-byte CUDA IPC handle of a texture's texels, to be carried to the peer ranks by any
 *   transport, and mapped there.
 * composite: stores the rows this rank owns of its srcTexture into rank `root`'s dstTexture (a peer mapping over
 *   NVLink when root != rank).  The caller synchronises the ranks (sync + barrier) before the root reads. */
SLCU_EXPORT int SLCU_FN(set_partition)(slcu_context* ctx, unsigned rank, unsigned world, unsigned stripRows);
SLCU_EXPORT int SLCU_FN(texture_ipc_export)(slcu_context* ctx, size_t texture, void* handle64);
SLCU_EXPORT int SLCU_FN(texture_ipc_import)(slcu_context* ctx, size_t texture, unsigned peerRank, const void* handle64);
SLCU_EXPORT int SLCU_FN(composite)(slcu_context* ctx, size_t srcTexture, size_t dstTexture, unsigned root);

/* Shared vertex stage (ours; SURVEY.md section 8e "shard K1 by primitive range").  Without it every rank runs the vertex
 * stage on all primitives.  With it rank r shades vertex blocks r, r + world, ... and stores each triangle record straight
 * into the arenas of the ranks whose rows the triangle touches (peer stores over NVLink; a device-side rendezvous, no host
 * round trip).  Every rank creates an arena of the same size after set_partition, exports its handle, and imports the
 * handles of all the others; from then on every triangle batch that fits the arena is shared (the decision is identical
 * on all ranks because they issue identical calls).  A batch needs about numPrimitives * (3 + 3 * varyings) * 16 * 1.15 bytes.
 * address / attach replace export / import for contexts that live in one process. */
SLCU_EXPORT int SLCU_FN(geometry_share_create)(slcu_context* ctx, size_t arenaBytes);
SLCU_EXPORT int SLCU_FN(geometry_share_export)(slcu_context* ctx, void* handle64);
SLCU_EXPORT int SLCU_FN(geometry_share_import)(slcu_context* ctx, unsigned peerRank, const void* handle64);
SLCU_EXPORT int SLCU_FN(geometry_share_address)(slcu_context* ctx, void** address);
SLCU_EXPORT int SLCU_FN(geometry_share_attach)(slcu_context* ctx, unsigned peerRank, void* address);

/* The reference's draw/clear/blit return after all workers joined.  Ours
 * enqueue on one stream; sync() is the join.  texture_read syncs by itself. */
SLCU_EXPORT int SLCU_FN(sync)(slcu_context* ctx);

/* ours: counters of the work enqueued since reset_stats. */
SLCU_EXPORT int SLCU_FN(reset_stats)(slcu_context* ctx);
SLCU_EXPORT int SLCU_FN(get_stats)(slcu_context* ctx, slcu_stats* out);

/* ours: per-kernel device time.  Between profile_begin and profile_end every
 * kernel launch is bracketed by CUDA events on the context's own stream;
 * profile_end synchronises and returns the summed milliseconds and launch
 * counts per kernel class.  Used by bench.py for the roofline of the dominant
 * kernel; off by default (the events serialise nothing but cost a few us). */
enum { SLCU_K_CLEAR = 0, SLCU_K_VERTEX, SLCU_K_SPAN_COUNT, SLCU_K_SCAN, SLCU_K_SPAN_FILL, SLCU_K_RASTER, SLCU_K_BLIT, SLCU_K_COMPOSITE, SLCU_K_SHADE,
       SLCU_K_ROUTE /* shared vertex stage: records to their ranks */, SLCU_K_RENDEZVOUS /* ... and the wait for the other ranks' */, SLCU_K_NUM };
typedef struct slcu_kernel_times
{
    double ms[SLCU_K_NUM];
    uint64_t launches[SLCU_K_NUM];
} slcu_kernel_times;
/* ours: device time of everything enqueued between the two calls, measured
 * with a CUDA event pair on the context's own stream (timer_end synchronises). */
SLCU_EXPORT int SLCU_FN(timer_begin)(slcu_context* ctx);
SLCU_EXPORT int SLCU_FN(timer_end)(slcu_context* ctx, double* ms);
SLCU_EXPORT int SLCU_FN(profile_begin)(slcu_context* ctx);
SLCU_EXPORT int SLCU_FN(profile_end)(slcu_context* ctx, slcu_kernel_times* out);

#ifdef __cplusplus
}
#endif

#endif /* SLCU_H */
