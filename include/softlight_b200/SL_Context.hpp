/*
 * SL_Context.hpp -- the C++ host side above the C ABI (include/slcu.h): a header-only mirror of the slice of
 * SoftLight's public surface that the draw path needs, with the reference's class and method names, argument
 * meaning and return codes, so that host code written against softlight/SL_Context.hpp compiles against this header
 * with two changes only:
 *   - a shader's `shader` member names a registered __device__ callable (const char*) instead of holding a host
 *     function pointer (a GPU cannot call into host memory);
 *   - host pointers stored inside a uniform block are declared (SL_UniformBuffer::bind_texture / bind_buffer).
 *
 * What it mirrors (reference file:line under softlight/):
 *   SL_Context            include/softlight/SL_Context.hpp:66-310, src/SL_Context.cpp:160-1148
 *   SL_VertexBuffer       include/softlight/SL_VertexBuffer.hpp:20-118
 *   SL_IndexBuffer        include/softlight/SL_IndexBuffer.hpp, src/SL_IndexBuffer.cpp:117-170
 *   SL_VertexArray        include/softlight/SL_VertexArray.hpp:40-126
 *   SL_UniformBuffer      include/softlight/SL_UniformBuffer.hpp:23-110
 *   SL_Texture / View     include/softlight/SL_Texture.hpp:29-455, src/SL_Texture.cpp:276-294
 *   SL_Framebuffer        include/softlight/SL_Framebuffer.hpp, src/SL_Framebuffer.cpp:823-990
 *   SL_Shader structs     include/softlight/SL_Shader.hpp:41-115
 *   SL_Mesh               include/softlight/SL_Mesh.hpp:30-39
 *   SL_ViewportState      include/softlight/SL_ViewportState.hpp:34-69
 *
 * Objects are owned by the context and addressed by vector position, like the reference.  Host-visible memory
 * (vertex data, uniform blocks) lives in a CPU shadow the user writes through data()/as<T>(); the context uploads
 * what changed at draw / clear / blit (the three entry points into SL_ProcessorPool that the library replaces).
 * There is no CPU fallback: without a CUDA device the constructor leaves the context invalid and every call fails.
 *
 * Everything lives in namespace sl_b200 so that the header can sit next to the reference's own headers in one
 * program (tests do exactly that); `using namespace sl_b200;` gives the reference's spelling.
 */
#ifndef SOFTLIGHT_B200_SL_CONTEXT_HPP
#define SOFTLIGHT_B200_SL_CONTEXT_HPP

#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include "slcu.h"

namespace sl_b200
{

/* enum values are the reference's (SL_Color.hpp:29-70, SL_Geometry.hpp, SL_Mesh.hpp:17-28, SL_PipelineState.hpp:22-100) */
enum SL_ColorDataType : int
{
    SL_COLOR_R_8U = 0, SL_COLOR_R_16U, SL_COLOR_R_32U, SL_COLOR_R_64U, SL_COLOR_R_HALF, SL_COLOR_R_FLOAT, SL_COLOR_R_DOUBLE,
    SL_COLOR_RG_8U, SL_COLOR_RG_16U, SL_COLOR_RG_32U, SL_COLOR_RG_64U, SL_COLOR_RG_HALF, SL_COLOR_RG_FLOAT, SL_COLOR_RG_DOUBLE,
    SL_COLOR_RGB_8U, SL_COLOR_RGB_16U, SL_COLOR_RGB_32U, SL_COLOR_RGB_64U, SL_COLOR_RGB_HALF, SL_COLOR_RGB_FLOAT, SL_COLOR_RGB_DOUBLE,
    SL_COLOR_RGBA_8U, SL_COLOR_RGBA_16U, SL_COLOR_RGBA_32U, SL_COLOR_RGBA_64U, SL_COLOR_RGBA_HALF, SL_COLOR_RGBA_FLOAT, SL_COLOR_RGBA_DOUBLE,
    /* the packed types of SL_ColorCompressed.hpp: one unsigned integer per texel, r in the lowest bits */
    SL_COLOR_RGB_332, SL_COLOR_RGB_565, SL_COLOR_RGBA_5551, SL_COLOR_RGBA_4444, SL_COLOR_RGBA_1010102
};
/* sl_is_compressed_color (SL_Color.hpp:94-102) */
constexpr bool sl_is_compressed_color(SL_ColorDataType p) noexcept { return (int)p >= (int)SL_COLOR_RGB_332 && (int)p <= (int)SL_COLOR_RGBA_1010102; }
enum SL_DataType : int { VERTEX_DATA_BYTE = 0, VERTEX_DATA_SHORT, VERTEX_DATA_INT, VERTEX_DATA_LONG, VERTEX_DATA_FLOAT, VERTEX_DATA_DOUBLE };
enum SL_Dimension : int { VERTEX_DIMENSION_1 = 0, VERTEX_DIMENSION_2, VERTEX_DIMENSION_3, VERTEX_DIMENSION_4 };
enum SL_RenderMode : uint32_t
{
    RENDER_MODE_POINTS = 0x01, RENDER_MODE_INDEXED_POINTS = 0x03, RENDER_MODE_LINES = 0x04, RENDER_MODE_INDEXED_LINES = 0x0C,
    RENDER_MODE_TRIANGLES = 0x10, RENDER_MODE_INDEXED_TRIANGLES = 0x30, RENDER_MODE_TRI_WIRE = 0x50, RENDER_MODE_INDEXED_TRI_WIRE = 0x70
};
enum SL_CullMode : uint8_t { SL_CULL_BACK_FACE = 0, SL_CULL_FRONT_FACE, SL_CULL_OFF };
enum SL_DepthTest : uint8_t
{
    SL_DEPTH_TEST_OFF = 0, SL_DEPTH_TEST_LESS_THAN, SL_DEPTH_TEST_LESS_EQUAL, SL_DEPTH_TEST_GREATER_THAN, SL_DEPTH_TEST_GREATER_EQUAL,
    SL_DEPTH_TEST_EQUAL, SL_DEPTH_TEST_NOT_EQUAL
};
enum SL_DepthMask : uint8_t { SL_DEPTH_MASK_OFF = 0, SL_DEPTH_MASK_ON };
enum SL_BlendMode : uint8_t { SL_BLEND_OFF = 0, SL_BLEND_ALPHA, SL_BLEND_PREMULTIPLED_ALPHA, SL_BLEND_ADDITIVE, SL_BLEND_SCREEN };
enum class SL_TexelOrder : int { ORDERED = 0, SWIZZLED = 1 };

/* SL_Mesh.hpp:30-39 */
struct SL_Mesh
{
    size_t vaoId;
    size_t elementBegin;
    size_t elementEnd;
    SL_RenderMode mode;
    uint32_t materialId;
};

/* SL_Shader.hpp:58-63, 88-97: `shader` is the registry name of the device callable (see file header) */
struct SL_VertexShader
{
    uint8_t numVaryings;
    SL_CullMode cullMode;
    const char* shader;
};
struct SL_FragmentShader
{
    uint8_t numVaryings;
    uint8_t numOutputs;
    SL_BlendMode blend;
    SL_DepthTest depthTest;
    SL_DepthMask depthMask;
    const char* shader;
};

struct SL_ColorRGBAd { double v[4]; }; /* ls::math::vec4_t<double> of clear_* */

class SL_Context;

/* every resource object knows its context and id; none owns device memory itself */
class SL_Resource
{
  protected:
    friend class SL_Context;
    slcu_context* mDev = nullptr;
    size_t mId = SLCU_INVALID_ID;
};

class SL_VertexBuffer : public SL_Resource
{
    std::vector<unsigned char> mShadow;
    std::vector<unsigned char> mSent; /* what the device holds, kept once a writable pointer has left the object */
    bool mDirty = false;
    bool mHandedOut = false;
    friend class SL_Context;

  public:
    /* SL_VertexBuffer::init(numBytes): 0 on success */
    int init(size_t numBytes) noexcept
    {
        mShadow.assign(numBytes, 0);
        mDirty = true;
        return slcu_vbo_init(mDev, mId, numBytes);
    }
    size_t num_bytes() const noexcept { return mShadow.size(); }
    void* data() noexcept { mDirty = true; mHandedOut = true; return mShadow.data(); } /* written in place by the user, uploaded at the next draw */
    const void* data() const noexcept { return mShadow.data(); }
    template <typename T> T* element(ptrdiff_t offset) noexcept { mDirty = true; mHandedOut = true; return reinterpret_cast<T*>(mShadow.data() + offset); }
    void assign(const void* pInputData, ptrdiff_t offset, size_t numBytes) noexcept
    {
        std::memcpy(mShadow.data() + offset, pInputData, numBytes);
        slcu_vbo_assign(mDev, mId, pInputData, (size_t)offset, numBytes);
    }
};

class SL_IndexBuffer : public SL_Resource
{
    uint32_t mCount = 0;
    SL_DataType mType = VERTEX_DATA_INT;

  public:
    /* SL_IndexBuffer::init(numElements, type, pData): 0 on success, -1 bad type, -2 allocation */
    int init(uint32_t numElements, SL_DataType type, const void* pData) noexcept
    {
        mCount = numElements;
        mType = type;
        return slcu_ibo_init(mDev, mId, numElements, (int)type, pData);
    }
    uint32_t count() const noexcept { return mCount; }
    SL_DataType type() const noexcept { return mType; }
};

class SL_VertexArray : public SL_Resource
{
  public:
    void set_vertex_buffer(size_t vboId) noexcept { slcu_vao_set_vertex_buffer(mDev, mId, vboId); }
    void set_index_buffer(size_t iboId) noexcept { slcu_vao_set_index_buffer(mDev, mId, iboId); }
    int set_num_bindings(size_t numBindings) noexcept { return slcu_vao_set_num_bindings(mDev, mId, (int)numBindings); }
    void set_binding(size_t bindId, ptrdiff_t offset, ptrdiff_t stride, SL_Dimension numDimens, SL_DataType vertType) noexcept
    {
        slcu_vao_set_binding(mDev, mId, bindId, offset, stride, (int)numDimens, (int)vertType);
    }
};

/* SL_UniformBuffer.hpp:23: a 1 KiB POD blob the user reinterprets */
class SL_UniformBuffer : public SL_Resource
{
    alignas(16) unsigned char mBytes[1024];
    bool mDirty = true;
    friend class SL_Context;

  public:
    SL_UniformBuffer() { std::memset(mBytes, 0, sizeof(mBytes)); }
    template <typename T> T* as() noexcept { mDirty = true; return reinterpret_cast<T*>(mBytes); }
    template <typename T> const T* as() const noexcept { return reinterpret_cast<const T*>(mBytes); }
    void assign(const void* pData, size_t offset, size_t numBytes) noexcept { std::memcpy(mBytes + offset, pData, numBytes); mDirty = true; }
    /* ours: the 8-byte slot at `offset` holds a host pointer the shader chases; on the device it must point to the
     * texture object / into the vertex buffer.  Declared once, patched at every draw. */
    int bind_texture(size_t offset, size_t textureId) noexcept { return slcu_ubo_bind_texture(mDev, mId, offset, textureId); }
    int bind_buffer(size_t offset, size_t vboId, size_t byteOffset) noexcept { return slcu_ubo_bind_buffer(mDev, mId, offset, vboId, byteOffset); }
};

/* SL_Texture.hpp:29-44 reduced to what the path needs: which texture, its shape and type */
struct SL_TextureView
{
    size_t textureId;
    uint16_t width, height, depth;
    SL_ColorDataType type;
};

/* The reference's SL_TextureView over HOST texels (SL_Texture.hpp:29-44) as SL_WindowBuffer / SL_Swapchain hand it out for
 * the window's back buffer: the destination of SL_Context::blit(SL_TextureView&, textureId) */
struct SL_HostTextureView
{
    void* pTexels;
    uint16_t width, height;
    SL_ColorDataType type;
};

class SL_Texture : public SL_Resource
{
    SL_TextureView mView{SLCU_INVALID_ID, 0, 0, 0, SL_COLOR_RGBA_8U};
    mutable std::vector<unsigned char> mReadback;

    static size_t bytes_per_texel(SL_ColorDataType t) noexcept
    {
        static const size_t chan[7] = {1, 2, 4, 8, 2, 4, 8}, packed[5] = {1, 2, 2, 2, 4};
        if (sl_is_compressed_color(t)) return packed[(int)t - (int)SL_COLOR_RGB_332];
        return chan[(int)t % 7] * (size_t)((int)t / 7 + 1);
    }

  public:
    /* SL_Texture::init(type, w, h, d): 0 on success, -1 on failure (src/SL_Texture.cpp:276-294) */
    int init(SL_ColorDataType type, uint16_t w, uint16_t h, uint16_t d = 1) noexcept
    {
        const int rc = slcu_texture_init(mDev, mId, (int)type, w, h, d);
        if (rc == 0) mView = SL_TextureView{mId, w, h, d, type};
        return rc;
    }
    uint16_t width() const noexcept { return mView.width; }
    uint16_t height() const noexcept { return mView.height; }
    uint16_t depth() const noexcept { return mView.depth; }
    SL_ColorDataType type() const noexcept { return mView.type; }
    size_t bpp() const noexcept { return bytes_per_texel(mView.type); }
    const SL_TextureView& view() const noexcept { return mView; }
    /* SL_Texture::set_texels<order>(x, y, z, w, h, d, pData) (SL_Texture.hpp:418-444) */
    template <SL_TexelOrder order = SL_TexelOrder::ORDERED>
    int set_texels(uint16_t x, uint16_t y, uint16_t z, uint16_t w, uint16_t h, uint16_t d, const void* pData) noexcept
    {
        return slcu_texture_set_texels(mDev, mId, (int)order, x, y, z, w, h, d, pData);
    }
    /* SL_Texture::data(): the texels in linear order; joins the device work first (the reference's draw is synchronous) */
    const void* data() const noexcept
    {
        mReadback.resize((size_t)mView.width * mView.height * mView.depth * bytes_per_texel(mView.type));
        if (slcu_texture_read(mDev, mId, mReadback.data(), mReadback.size()) != 0) return nullptr;
        return mReadback.data();
    }
};

class SL_Framebuffer : public SL_Resource
{
  public:
    /* return codes as in src/SL_Framebuffer.cpp:823-990 (reserve: 0 / -1 / -2; attach: 0 / -1; valid: 0 ... -10) */
    int reserve_color_buffers(uint64_t numColorBuffers) noexcept { return slcu_fbo_reserve_color_buffers(mDev, mId, (unsigned)numColorBuffers); }
    int attach_color_buffer(uint64_t index, const SL_TextureView& view) noexcept { return slcu_fbo_attach_color_buffer(mDev, mId, (unsigned)index, view.textureId); }
    int attach_depth_buffer(const SL_TextureView& view) noexcept { return slcu_fbo_attach_depth_buffer(mDev, mId, view.textureId); }
    int valid() const noexcept { return slcu_fbo_valid(mDev, mId); }
};

/* SL_ViewportState.hpp:34-69 */
class SL_ViewportState
{
    slcu_context* mDev = nullptr;
    friend class SL_Context;

  public:
    void viewport(int32_t x, int32_t y, uint16_t w, uint16_t h) noexcept { slcu_viewport(mDev, x, y, w, h); }
    void scissor(int32_t x, int32_t y, uint16_t w, uint16_t h) noexcept { slcu_scissor(mDev, x, y, w, h); }
};

class SL_Context
{
    slcu_context* mDev = nullptr;
    int mStatus = SLCU_ERR_NO_DEVICE;
    /* stable addresses: users hold references across create_* calls, like with the reference's aligned vectors of
     * heap objects (textures) -- here every object is heap-allocated */
    std::vector<SL_VertexBuffer*> mVbos;
    std::vector<SL_IndexBuffer*> mIbos;
    std::vector<SL_VertexArray*> mVaos;
    std::vector<SL_UniformBuffer*> mUbos;
    std::vector<SL_Texture*> mTextures;
    std::vector<SL_Framebuffer*> mFbos;
    SL_ViewportState mViewport;
    unsigned mNumThreads = 1;

    template <typename T> size_t add(std::vector<T*>& v, size_t id)
    {
        if (id == SLCU_INVALID_ID) return id;
        T* o = new T();
        o->mDev = mDev;
        o->mId = id;
        if (v.size() <= id) v.resize(id + 1, nullptr);
        v[id] = o;
        return id;
    }
    template <typename T> static void drop(std::vector<T*>& v)
    {
        for (T* o : v) delete o;
        v.clear();
    }

    /* what the user wrote into host-visible objects since the last entry into the pipeline */
    void upload_dirty() noexcept
    {
        /* The reference hands out live pointers, and its own programs cache them and keep writing through them between
         * draws (sl_instancing_test.cpp:236, 251-274).  Uniform blocks (1 KiB, snapshotted at submission anyway) are
         * therefore sent with every draw; a vertex buffer whose non-const pointer was ever handed out is compared with
         * what was last sent and re-sent when it differs. */
        for (SL_VertexBuffer* b : mVbos)
        {
            if (!b || b->mShadow.empty()) continue;
            if (b->mDirty || (b->mHandedOut && (b->mSent.size() != b->mShadow.size() || std::memcmp(b->mSent.data(), b->mShadow.data(), b->mShadow.size()) != 0)))
            {
                slcu_vbo_assign(mDev, b->mId, b->mShadow.data(), 0, b->mShadow.size());
                if (b->mHandedOut) b->mSent = b->mShadow;
                b->mDirty = false;
            }
        }
        for (SL_UniformBuffer* u : mUbos)
        {
            if (u) slcu_ubo_assign(mDev, u->mId, u->mBytes, 0, sizeof(u->mBytes));
        }
    }

  public:
    explicit SL_Context(int device = 0) noexcept
    {
        mStatus = slcu_context_create(&mDev, device);
        if (mStatus != 0) mDev = nullptr;
        mViewport.mDev = mDev;
    }
    ~SL_Context() noexcept { terminate(); }
    SL_Context(const SL_Context&) = delete;
    SL_Context& operator=(const SL_Context&) = delete;

    bool valid() const noexcept { return mDev != nullptr; }              /* ours: false when there is no CUDA device */
    int status() const noexcept { return mStatus; }
    const char* last_error() const noexcept { return mDev ? slcu_last_error(mDev) : "no CUDA device (there is no CPU fallback)"; }
    slcu_context* native_handle() noexcept { return mDev; }

    void terminate() noexcept
    {
        drop(mVbos); drop(mIbos); drop(mVaos); drop(mUbos); drop(mTextures); drop(mFbos);
        if (mDev) slcu_context_destroy(mDev);
        mDev = nullptr;
    }

    /* create_* return the index of the new object, (size_t)-1 on failure (src/SL_Context.cpp:160-460) */
    size_t create_vbo() noexcept { return mDev ? add(mVbos, slcu_create_vbo(mDev)) : SLCU_INVALID_ID; }
    size_t create_ibo() noexcept { return mDev ? add(mIbos, slcu_create_ibo(mDev)) : SLCU_INVALID_ID; }
    size_t create_vao() noexcept { return mDev ? add(mVaos, slcu_create_vao(mDev)) : SLCU_INVALID_ID; }
    size_t create_ubo() noexcept { return mDev ? add(mUbos, slcu_create_ubo(mDev)) : SLCU_INVALID_ID; }
    size_t create_texture() noexcept { return mDev ? add(mTextures, slcu_create_texture(mDev)) : SLCU_INVALID_ID; }
    size_t create_framebuffer() noexcept { return mDev ? add(mFbos, slcu_create_framebuffer(mDev)) : SLCU_INVALID_ID; }

    /* destroy_* (src/SL_Context.cpp:192-195, 243-247, 295-298, 346-349, 397-400, 448-451, 612-615): the object is erased,
     * every later index of that kind moves down by one */
    template <typename T> void erase(std::vector<T*>& v, size_t index) noexcept
    {
        if (index >= v.size()) return;
        delete v[index];
        v.erase(v.begin() + (ptrdiff_t)index);
        for (size_t i = index; i < v.size(); ++i) { if (v[i]) v[i]->mId = i; }
    }
    void destroy_vbo(size_t index) noexcept { if (mDev && slcu_destroy_vbo(mDev, index) == 0) erase(mVbos, index); }
    void destroy_ibo(size_t index) noexcept { if (mDev && slcu_destroy_ibo(mDev, index) == 0) erase(mIbos, index); }
    void destroy_vao(size_t index) noexcept { if (mDev && slcu_destroy_vao(mDev, index) == 0) erase(mVaos, index); }
    void destroy_ubo(size_t index) noexcept { if (mDev && slcu_destroy_ubo(mDev, index) == 0) erase(mUbos, index); }
    void destroy_texture(size_t index) noexcept { if (mDev && slcu_destroy_texture(mDev, index) == 0) erase(mTextures, index); }
    void destroy_framebuffer(size_t index) noexcept { if (mDev && slcu_destroy_framebuffer(mDev, index) == 0) erase(mFbos, index); }
    void destroy_shader(size_t index) noexcept { if (mDev) slcu_destroy_shader(mDev, index); }

    SL_VertexBuffer& vbo(size_t i) noexcept { return *mVbos[i]; }
    SL_IndexBuffer& ibo(size_t i) noexcept { return *mIbos[i]; }
    SL_VertexArray& vao(size_t i) noexcept { return *mVaos[i]; }
    SL_UniformBuffer& ubo(size_t i) noexcept { return *mUbos[i]; }
    SL_Texture& texture(size_t i) noexcept { return *mTextures[i]; }
    SL_Framebuffer& framebuffer(size_t i) noexcept { return *mFbos[i]; }
    const SL_VertexBuffer& vbo(size_t i) const noexcept { return *mVbos[i]; }
    const SL_IndexBuffer& ibo(size_t i) const noexcept { return *mIbos[i]; }
    const SL_VertexArray& vao(size_t i) const noexcept { return *mVaos[i]; }
    const SL_UniformBuffer& ubo(size_t i) const noexcept { return *mUbos[i]; }
    const SL_Texture& texture(size_t i) const noexcept { return *mTextures[i]; }
    const SL_Framebuffer& framebuffer(size_t i) const noexcept { return *mFbos[i]; }
    /* the object lists (SL_Context.hpp:100-180; scene-level callers size their id ranges with them, e.g.
     * SL_SceneGraph.cpp:1059 `mContext.vaos().size()`): pointers to the live mirrors here, objects by value there */
    const std::vector<SL_VertexBuffer*>& vbos() const noexcept { return mVbos; }
    const std::vector<SL_IndexBuffer*>& ibos() const noexcept { return mIbos; }
    const std::vector<SL_VertexArray*>& vaos() const noexcept { return mVaos; }
    const std::vector<SL_UniformBuffer*>& ubos() const noexcept { return mUbos; }
    const std::vector<SL_Texture*>& textures() const noexcept { return mTextures; }
    const std::vector<SL_Framebuffer*>& framebuffers() const noexcept { return mFbos; }

    /* SL_Context::create_shader(vertShader, fragShader, uniformIndex): index or (size_t)-1 (src/SL_Context.cpp:488-605);
     * an unregistered shader name is refused the same way */
    size_t create_shader(const SL_VertexShader& vertShader, const SL_FragmentShader& fragShader, size_t uniformIndex) noexcept
    {
        if (!mDev) return SLCU_INVALID_ID;
        const slcu_vertex_shader vs{vertShader.numVaryings, (uint8_t)vertShader.cullMode, vertShader.shader};
        const slcu_fragment_shader fs{fragShader.numVaryings, fragShader.numOutputs, (uint8_t)fragShader.blend, (uint8_t)fragShader.depthTest,
                                      (uint8_t)fragShader.depthMask, fragShader.shader};
        return slcu_create_shader(mDev, &vs, &fs, uniformIndex);
    }

    SL_ViewportState& viewport_state() noexcept { return mViewport; }

    /* the three entry points SL_ProcessorPool::run_shader_processors stands behind (src/SL_Context.cpp:714-740) */
    void draw(const SL_Mesh& m, size_t shaderId, size_t fboId) noexcept
    {
        upload_dirty();
        const slcu_mesh cm{m.vaoId, m.elementBegin, m.elementEnd, (uint32_t)m.mode, m.materialId};
        mStatus = slcu_draw(mDev, &cm, shaderId, fboId);
    }
    void draw_multiple(const SL_Mesh* meshes, size_t numMeshes, size_t shaderId, size_t fboId) noexcept
    {
        upload_dirty();
        std::vector<slcu_mesh> cm(numMeshes);
        for (size_t i = 0; i < numMeshes; ++i) cm[i] = slcu_mesh{meshes[i].vaoId, meshes[i].elementBegin, meshes[i].elementEnd, (uint32_t)meshes[i].mode, meshes[i].materialId};
        mStatus = slcu_draw_multiple(mDev, cm.data(), numMeshes, shaderId, fboId);
    }
    void draw_instanced(const SL_Mesh& m, size_t numInstances, size_t shaderId, size_t fboId) noexcept
    {
        upload_dirty();
        const slcu_mesh cm{m.vaoId, m.elementBegin, m.elementEnd, (uint32_t)m.mode, m.materialId};
        mStatus = slcu_draw_instanced(mDev, &cm, numInstances, shaderId, fboId);
    }

    /* run_blit_processors (src/SL_Context.cpp:747-813) */
    void blit(size_t outTextureId, size_t inTextureId) noexcept { mStatus = slcu_blit(mDev, outTextureId, inTextureId); }
    void blit(size_t outTextureId, size_t inTextureId, uint16_t srcX0, uint16_t srcY0, uint16_t srcX1, uint16_t srcY1, uint16_t dstX0, uint16_t dstY0,
              uint16_t dstX1, uint16_t dstY1) noexcept
    {
        mStatus = slcu_blit_rect(mDev, outTextureId, inTextureId, srcX0, srcY0, srcX1, srcY1, dstX0, dstY0, dstX1, dstY1);
    }

    /* blit to a window (src/SL_Context.cpp:816-874): into a host-side view, synchronous like the reference's */
    void blit(SL_HostTextureView& buffer, size_t textureId) noexcept
    {
        mStatus = slcu_blit_to_host(mDev, buffer.pTexels, (int)buffer.type, buffer.width, buffer.height, textureId);
    }
    void blit(SL_HostTextureView& buffer, size_t textureId, uint16_t srcX0, uint16_t srcY0, uint16_t srcX1, uint16_t srcY1, uint16_t dstX0,
              uint16_t dstY0, uint16_t dstX1, uint16_t dstY1) noexcept
    {
        mStatus = slcu_blit_to_host_rect(mDev, buffer.pTexels, (int)buffer.type, buffer.width, buffer.height, textureId, srcX0, srcY0, srcX1, srcY1,
                                         dstX0, dstY0, dstX1, dstY1);
    }

    /* run_clear_processors (src/SL_Context.cpp:881-963) */
    void clear_color_buffer(size_t fboId, unsigned attachmentId, const SL_ColorRGBAd& color) noexcept { mStatus = slcu_clear_color_buffer(mDev, fboId, attachmentId, color.v); }
    void clear_depth_buffer(size_t fboId, double depth) noexcept { mStatus = slcu_clear_depth_buffer(mDev, fboId, depth); }
    void clear_framebuffer(size_t fboId, unsigned attachmentId, const SL_ColorRGBAd& color, double depth) noexcept
    {
        mStatus = slcu_clear_framebuffer(mDev, fboId, attachmentId, color.v, depth);
    }
    template <size_t N>
    void clear_framebuffer(size_t fboId, const std::array<unsigned, N>& bufferIndices, const std::array<SL_ColorRGBAd, N>& colors, double depth) noexcept
    {
        for (size_t i = 0; i < N; ++i) slcu_clear_color_buffer(mDev, fboId, bufferIndices[i], colors[i].v);
        mStatus = slcu_clear_depth_buffer(mDev, fboId, depth);
    }

    /* SL_Context::num_threads (src/SL_Context.cpp:1135-1148): kept for source compatibility, the GPU ignores it */
    unsigned num_threads() const noexcept { return mNumThreads; }
    unsigned num_threads(unsigned inNumThreads) noexcept { return mNumThreads = slcu_num_threads(mDev, inNumThreads); }

    /* ours: the reference's draw returns when the frame is done; here work is enqueued on a stream and joined by
     * texture(i).data(), or explicitly */
    int sync() noexcept { return mDev ? slcu_sync(mDev) : SLCU_ERR_NO_DEVICE; }
};

} // namespace sl_b200

#endif /* SOFTLIGHT_B200_SL_CONTEXT_HPP */
