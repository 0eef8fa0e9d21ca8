/*
 * slcu.cu -- host side of the C ABI (include/slcu.h): the device mirrors of
 * SL_Context's resources and the replacements of SL_ProcessorPool::run_*.
 *
 * Everything here only manages memory and launches; the arithmetic of the path
 * runs in the kernels of sl_pipeline.cuh.  The few scalars the reference
 * computes once per draw on the host (scissor matrix, viewport rectangle,
 * clear colours; src/SL_ViewportState.cpp:17-111, src/SL_Context.cpp:881-963)
 * are computed here with the same rcp table the device uses.
 *
 * There is no CPU fallback: without a CUDA device context_create fails.
 */
#include <cuda_runtime.h>
#include <nvrtc.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "slcu.h"
#include "sl_pipeline.cuh"
#include "sl_resolve2.cuh"
#include "sl_points.cuh"
#include "sl_wire.cuh"
#include "sl_lines.cuh"
#include "sl_sort.cuh"

using namespace sld;

/* ===========================================================================
 * Shader registry: SL_REGISTER_VERTEX_SHADER / SL_REGISTER_FRAGMENT_SHADER
 * pair a name with a __device__ callable; the kernels are instantiated with it
 * as a template argument (direct, inlinable calls -- never interpreted).
 * =========================================================================*/
typedef cudaError_t (*VertexLauncher)(unsigned grid, cudaStream_t s, const SLVertexParams& p, bool cache);
static const unsigned kVertexCacheMaxBytes = 3u * SL_VB_PRIMS * (unsigned)sizeof(uint16_t) + 6u * SL_VCACHE_VERTS * (unsigned)sizeof(float4) + SL_VCACHE_VERTS / 8u + SL_VCACHE_VERTS; /* indices, clip + screen position, 4 varyings, used bits, flags */
typedef cudaError_t (*RasterLauncher)(unsigned grid, cudaStream_t s, const SLShadeParams& p);

typedef cudaError_t (*PointVertexLauncher)(unsigned grid, cudaStream_t s, const SLPointParams& p);
typedef cudaError_t (*PointShadeLauncher)(unsigned grid, cudaStream_t s, const SLPointShadeParams& p);
typedef cudaError_t (*WireShadeLauncher)(unsigned grid, cudaStream_t s, const SLWireShadeParams& p);
typedef cudaError_t (*LineVertexLauncher)(unsigned grid, cudaStream_t s, const SLLineParams& p);
typedef cudaError_t (*LineShadeLauncher)(unsigned grid, cudaStream_t s, const SLLineShadeParams& p);

struct VsEntry { const char* name; VertexLauncher launch; PointVertexLauncher launchPoints; LineVertexLauncher launchLines; };
struct FsEntry { const char* name; RasterLauncher launch; PointShadeLauncher launchPoints; WireShadeLauncher launchWire; LineShadeLauncher launchLines; };
static std::vector<VsEntry>& vs_registry() { static std::vector<VsEntry> r; return r; }
static std::vector<FsEntry>& fs_registry() { static std::vector<FsEntry> r; return r; }

template <VsFn VS> static cudaError_t launch_vertex(unsigned grid, cudaStream_t s, const SLVertexParams& p, bool cache)
{
    if (!cache)
    {
        k_vertex<VS, false><<<grid, SL_VB_THREADS, 0, s>>>(p);
        return cudaGetLastError();
    }
    static bool attrSet[64] = {}; /* one opt-in per instantiation and device: the cache may exceed the 48 KB default */
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attrSet[dev & 63])
    {
        cudaError_t e = cudaFuncSetAttribute(k_vertex<VS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVertexCacheMaxBytes);
        if (e != cudaSuccess) return e;
        attrSet[dev & 63] = true;
    }
    k_vertex<VS, true><<<grid, SL_VB_THREADS, p.cacheBytes, s>>>(p);
    return cudaGetLastError();
}
template <FsFn FS> static cudaError_t launch_raster(unsigned grid, cudaStream_t s, const SLShadeParams& p)
{
    k_shade<FS><<<grid, SL_SHADE_WARPS * 32, 0, s>>>(p);
    return cudaGetLastError();
}
template <VsFn VS> static cudaError_t launch_point_vertex(unsigned grid, cudaStream_t s, const SLPointParams& p)
{
    k_point_vertex<VS><<<grid, 256, 0, s>>>(p);
    return cudaGetLastError();
}
template <FsFn FS> static cudaError_t launch_point_shade(unsigned grid, cudaStream_t s, const SLPointShadeParams& p)
{
    k_point_shade<FS><<<grid, 128, 0, s>>>(p);
    return cudaGetLastError();
}
template <FsFn FS> static cudaError_t launch_wire_shade(unsigned grid, cudaStream_t s, const SLWireShadeParams& p)
{
    k_wire_shade<FS><<<grid, 128, 0, s>>>(p);
    return cudaGetLastError();
}
template <VsFn VS> static cudaError_t launch_line_vertex(unsigned grid, cudaStream_t s, const SLLineParams& p)
{
    k_line_vertex<VS><<<grid, SL_LINE_BLOCK, 0, s>>>(p);
    return cudaGetLastError();
}
template <FsFn FS> static cudaError_t launch_line_shade(unsigned grid, cudaStream_t s, const SLLineShadeParams& p)
{
    k_line_shade<FS><<<grid, 128, 0, s>>>(p);
    return cudaGetLastError();
}
struct VsReg { VsReg(const char* n, VertexLauncher l, PointVertexLauncher pl, LineVertexLauncher ll) { vs_registry().push_back(VsEntry{n, l, pl, ll}); } };
struct FsReg { FsReg(const char* n, RasterLauncher l, PointShadeLauncher pl, WireShadeLauncher wl, LineShadeLauncher ll) { fs_registry().push_back(FsEntry{n, l, pl, wl, ll}); } };

#define SL_SHADER_FN static __device__ __forceinline__
#define SL_CAT2(a, b) a##b
#define SL_CAT(a, b) SL_CAT2(a, b)
#define SL_REGISTER_VERTEX_SHADER(name, fn) static VsReg SL_CAT(vsreg_, fn){name, launch_vertex<fn>, launch_point_vertex<fn>, launch_line_vertex<fn>};
#define SL_REGISTER_FRAGMENT_SHADER(name, fn) static FsReg SL_CAT(fsreg_, fn){name, launch_raster<fn>, launch_point_shade<fn>, launch_wire_shade<fn>, launch_line_shade<fn>};

namespace sld
{
#include "sl_shaders_basic.inl"
#include "sl_shaders_scene.inl"
#include "sl_shaders_anim_volume.inl"
}

/* ===========================================================================
 * Host copies of the two x86 behaviours the per-draw host scalars need
 * =========================================================================*/
static inline uint32_t h_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float h_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static float h_rcp(const uint32_t* table, float x)
{
    const uint32_t xb = h_f2u(x);
    const uint32_t s = xb & 0x80000000u, e = (xb >> 23) & 0xffu, m = xb & 0x7fffffu;
    if (e == 0xffu) return h_u2f(m ? (xb | 0x00400000u) : s);
    if (e == 0u) return h_u2f(s | 0x7f800000u);
    const uint32_t r = table[m >> 12];
    const int re = (int)((r >> 23) & 0xffu) - ((int)e - 127);
    if (re <= 0) return h_u2f(s);
    return h_u2f(s | ((uint32_t)re << 23) | (r & 0x7fffffu));
}
static inline float h_min(float a, float b) { return a < b ? a : b; }
static inline float h_max(float a, float b) { return a > b ? a : b; }

#if defined(__x86_64__)
static inline float host_rcpss(float x)
{
    float out;
    __asm__ volatile("rcpss %1, %0" : "=x"(out) : "x"(x));
    return out;
}
#endif

/* The reference executes rcpss on the host CPU; to match it on the same box the
 * table is measured from that CPU, and the "11 mantissa bits in, exponent
 * exact" structure the device emulation relies on is verified by sampling. */
static bool measure_host_rcp_table(uint32_t* t, std::string& err)
{
#if defined(__x86_64__)
    for (uint32_t i = 0; i < 2048; ++i) t[i] = h_f2u(host_rcpss(h_u2f(0x3f800000u | (i << 12))));
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 2000000; ++i)
    {
        s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
        const uint32_t xb = (uint32_t)((s * 0x2545F4914F6CDD1Dull) >> 32);
        const float x = h_u2f(xb);
        if (h_f2u(host_rcpss(x)) != h_f2u(h_rcp(t, x)))
        {
            err = "host rcpss is not an 11-bit table with exact exponent handling; this CPU is not supported yet";
            return false;
        }
    }
    return true;
#else
    (void)t;
    err = "no x86 host: pin the rcp table with set_rcp_table";
    return false;
#endif
}

/* rsqrtss, the same way: exponent parity + 11 mantissa bits in, exact exponent (line primitives: math::inversesqrt) */
static float h_rsqrt(const uint32_t* table, float x)
{
    const uint32_t xb = h_f2u(x);
    const uint32_t e = (xb >> 23) & 0xffu, m = xb & 0x7fffffu;
    if (e == 0xffu) return m ? h_u2f(xb | 0x00400000u) : ((xb >> 31) ? h_u2f(0xffc00000u) : 0.f);
    if (e == 0u) return h_u2f((xb & 0x80000000u) | 0x7f800000u);
    if (xb >> 31) return h_u2f(0xffc00000u);
    const int p = ((int)e - 127) & 1;
    const uint32_t r = table[(uint32_t)p * 2048u + (m >> 12)];
    const int re = (int)((r >> 23) & 0xffu) - ((int)e - (127 + p)) / 2;
    return h_u2f(((uint32_t)re << 23) | (r & 0x7fffffu));
}
static bool measure_host_rsqrt_table(uint32_t* t)
{
#if defined(__x86_64__)
    auto rsq = [](float x) { float out; __asm__ volatile("rsqrtss %1, %0" : "=x"(out) : "x"(x)); return out; };
    for (uint32_t p = 0; p < 2; ++p)
        for (uint32_t i = 0; i < 2048; ++i) t[p * 2048u + i] = h_f2u(rsq(h_u2f(((127u + p) << 23) | (i << 12))));
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 1000000; ++i)
    {
        s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
        const float x = h_u2f((uint32_t)((s * 0x2545F4914F6CDD1Dull) >> 32));
        if (h_f2u(rsq(x)) != h_f2u(h_rsqrt(t, x))) return false; /* lines are then refused until a table is pinned */
    }
    return true;
#else
    (void)t;
    return false;
#endif
}

/* ===========================================================================
 * Context
 * =========================================================================*/
struct DevBuffer
{
    unsigned char* ptr = nullptr;
    size_t bytes = 0;
};

struct HostTexture
{
    uint16_t w = 0, h = 0, d = 0;
    uint8_t bpt = 0, channels = 0;
    int type = 0;
    unsigned char* dev = nullptr;
    size_t allocBytes = 0;
    unsigned depthWritesSinceClear = 0; /* as a depth attachment: depth-writing passes since its last clear (sort-first partition, run_batch) */
    std::vector<unsigned char> shadow; /* only for textures written with set_texels */
};

struct HostIbo { int type = 0; uint32_t count = 0; unsigned long long maxIndex = 0; /* largest index in the buffer (host scan at init): decides the post-transform cache */ DevBuffer buf; };
struct HostVao { SL_VertexArray vao; int numBindings = 0; size_t vbo = (size_t)-1, ibo = (size_t)-1; };
/* SL_TextureView as SL_Framebuffer keeps it: a snapshot taken when the attachment is made (src/SL_Framebuffer.cpp:855-920);
 * valid() judges these snapshots, the draw path addresses the texture itself */
struct FboView { bool texels = false; uint16_t w = 0, h = 0, d = 0; int type = 0; };
struct HostFbo { unsigned numColors = 0; size_t colors[4] = {(size_t)-1, (size_t)-1, (size_t)-1, (size_t)-1}; size_t depth = (size_t)-1; FboView colorViews[4]; FboView depthView; };
/* kernels of a shader that was registered at run time (slcu_register_shader_source): the same templates, instantiated by
 * NVRTC for the user's __device__ callable and loaded with cudaLibraryLoadData */
struct RtVertexKernels { cudaKernel_t vertex[2] = {nullptr, nullptr}; cudaKernel_t pointVertex = nullptr, lineVertex = nullptr; };
struct RtFragmentKernels { cudaKernel_t shade = nullptr, pointShade = nullptr, wireShade = nullptr, lineShade = nullptr; };
struct HostShader { int cullMode, depthTest, depthMask, blend, numVaryings, numOutputs; VertexLauncher vs; RasterLauncher fs; PointVertexLauncher pointVs; PointShadeLauncher pointFs; WireShadeLauncher wireFs; LineVertexLauncher lineVs; LineShadeLauncher lineFs; size_t ubo; RtVertexKernels rtV; RtFragmentKernels rtF; };
struct RtModule { cudaLibrary_t lib = nullptr; void* rcpTable = nullptr; void* rsqrtTable = nullptr; };
struct RtVsEntry { std::string name; RtVertexKernels k; };
struct RtFsEntry { std::string name; RtFragmentKernels k; };
struct UboFixup { size_t ubo, offset; int kind; size_t id, byteOffset; };
struct HostUbo { unsigned char bytes[1024]; };

/* A grow-only device buffer of the pipeline.  Stream-ordered allocation on the context's stream: growing one in the middle
 * of a frame neither waits for the device to drain nor stalls other contexts on the same GPU (cudaFree does both -- with
 * several contexts sharing the vertex stage that is a deadlock: one waits on the device for a peer whose host sits in cudaFree). */
struct Scratch
{
    DevBuffer buf;
    cudaStream_t stream = nullptr;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= buf.bytes) return cudaSuccess;
        if (buf.ptr) cudaFreeAsync(buf.ptr, stream);
        buf.ptr = nullptr;
        buf.bytes = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, want, stream);
        if (e == cudaSuccess) { buf.ptr = (unsigned char*)p; buf.bytes = want; }
        return e;
    }
    void release() { if (buf.ptr) cudaFreeAsync(buf.ptr, stream); buf.ptr = nullptr; buf.bytes = 0; }
};

struct Batch
{
    size_t shaderId = 0, fboId = 0;
    std::vector<SLDrawDesc> descs;
    std::vector<uint32_t> blockDraw;
    unsigned numBlocks = 0;
    unsigned long long totalPrims = 0;
    unsigned long long maxVerts = 0; /* vertices of the largest vertex buffer of the batch (binding 0) */
    bool wire = false;               /* RENDER_MODE_TRI_WIRE / INDEXED_TRI_WIRE: edges only (sl_wire.cuh) */
    bool partFull = false;           /* sort-first: the fast path was refused, count every row (enqueue_batch) */
    unsigned long long maxIndex = 0; /* largest vertex index any draw of the batch can fetch (~0: a draw without index buffer) */
    bool geomLocal = false;          /* shared vertex stage: this attempt runs every block here (after a refusal) */
    void clear() { descs.clear(); blockDraw.clear(); numBlocks = 0; totalPrims = 0; maxVerts = 0; wire = false; partFull = false; maxIndex = 0; geomLocal = false; }
};

struct slcu_context
{
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<DevBuffer> vbos;
    std::vector<HostIbo> ibos;
    std::vector<HostVao> vaos;
    std::vector<HostUbo> ubos;
    std::vector<HostTexture> textures;
    std::vector<HostFbo> fbos;
    std::vector<HostShader> shaders;
    std::vector<UboFixup> fixups;
    std::vector<RtModule> rtModules; /* shader sources compiled at run time */
    std::vector<RtVsEntry> rtVs;
    std::vector<RtFsEntry> rtFs;
    SL_Texture* devTextures = nullptr; /* device array of texture descriptors (fix-up targets) */
    size_t devTexturesCap = 0;
    int32_t viewport[4] = {0, 0, 65535, 65535};
    int32_t scissor[4] = {0, 0, 65535, 65535};
    uint32_t rcpTable[2048];
    bool rcpValid = false;
    uint32_t rsqrtTable[4096];
    bool rsqrtValid = false;
    std::string err;
    slcu_stats stats{};
    SLCounters* counters = nullptr;     /* device */
    SLCounters* hostCounters = nullptr; /* pinned */
    Scratch records, blockInfo, cnt, off, segSums, rowTotals, rowStart, entries, handoff, rowFlags, rowNeeded, drawTable, blockDraw, recDraw;
    Scratch fragPool, chunkNext, listInfo; /* fragment lists, k_resolve -> k_shade */
    Scratch handoffRegs;
    Scratch wireLeaders; /* wireframe: positions where a pixel's sorted candidates start */
    Scratch wireSlots, wireRecInfo; /* wireframe: sort values (double buffer); first slot + candidate count per record */
    Scratch pointKeys, pointRecs, pointSortTemp; /* point primitives: sort keys (double buffer), records, radix-sort scratch */
    Scratch spanItems;   /* (vertex block, row band) work items of k_span */
    Scratch blockRows;   /* rows each vertex block's records touch (k_block_rows) */
    Scratch rowPerm;     /* rows by weight, heaviest first (k_row_order) */
    Scratch ticketOrder; /* skewed work order of k_resolve, cached per (rows, segments, skew) */
    unsigned orderRows = 0, orderSegs = 0, orderSkew = 0;
    Batch batch;            /* submitted, not yet executed draws */
    /* A batch is enqueued with the buffer sizes its predecessors needed and checked at its end (run_batch): no host round
     * trip in the middle of a frame.  When the device found the buffers too small it has skipped every kernel that writes
     * something visible, and the batch is run again with sizes taken from its own counters. */
    cudaEvent_t batchDone = nullptr;
    bool refusedPartition = false; /* the last refusal was the sort-first fast path's (a span crosses into another rank's row) */
    struct Caps { unsigned long long recOverflow = 0, entries = 0, chunks = 0, handoffSlots = 0, avgSpan = 0; } caps;
    bool syncDraws = false; /* SLCU_SYNC_DRAWS=1: execute every draw call on its own (debug / parity mode) */
    /* sort-first partition (slcu_set_partition) */
    unsigned partRank = 0, partWorld = 1, partStripRows = 8;
    struct PeerMap { size_t tex; unsigned rank; void* ptr; };
    std::vector<PeerMap> peerMaps;
    /* shared vertex stage (slcu_geometry_share_*): one allocation per rank = head (rendezvous words) + two arenas that
     * alternate from batch to batch; peer[t] = rank t's allocation as mapped here (own pointer for t == rank) */
    struct GeomShare
    {
        unsigned char* base = nullptr;
        size_t arenaBytes = 0;
        unsigned char* peer[SL_MAX_RANKS] = {};
        bool ipcMapped[SL_MAX_RANKS] = {};
        uint32_t epoch = 0;
        bool ready(unsigned world) const
        {
            if (!base || world < 2 || world > (unsigned)SL_MAX_RANKS) return false;
            for (unsigned t = 0; t < world; ++t) if (!peer[t]) return false;
            return true;
        }
    } geom;
    /* A texture_read that finds the batch rendering into that texture still pending: the batch's last stage (k_shade) then
     * runs band by band and every band's rows cross the link while the next band is shaded (enqueue_batch). */
    struct PendingRead { bool active = false, issued = false; size_t tex = 0; unsigned char* out = nullptr; size_t bytes = 0; } pendingRead;
    cudaEvent_t bandEvent[8] = {};
    cudaEvent_t readDone = nullptr;
    /* present path (slcu_blit_to_host / slcu_present): a second stream for the copy into the host view, two staging slots */
    cudaStream_t copyStream = nullptr;
    struct PresentSlot { Scratch staging; cudaEvent_t blitDone = nullptr, copyDone = nullptr; bool busy = false; } present[2];
    /* optional per-kernel event timing (profile_begin / profile_end) */
    cudaEvent_t timerA = nullptr, timerB = nullptr;
    bool profiling = false;
    struct ProfEvent { int cls; cudaEvent_t a, b; };
    std::vector<ProfEvent> profEvents;
    std::vector<cudaEvent_t> eventPool;
};

/* RAII bracket of one kernel launch: counts it and, when profiling, records events around it */
struct KernelScope
{
    slcu_context* c;
    cudaEvent_t b = nullptr;
    KernelScope(slcu_context* ctx, int cls) : c(ctx)
    {
        c->stats.kernelLaunches++;
        if (!c->profiling) return;
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; ++i)
        {
            if (!c->eventPool.empty()) { ev[i] = c->eventPool.back(); c->eventPool.pop_back(); }
            else cudaEventCreate(&ev[i]);
        }
        cudaEventRecord(ev[0], c->stream);
        b = ev[1];
        c->profEvents.push_back(slcu_context::ProfEvent{cls, ev[0], ev[1]});
    }
    ~KernelScope() { if (b) cudaEventRecord(b, c->stream); }
};

/* what g_rcp_table holds on each device (function attributes and __device__ symbols are per device) */
static bool g_activeRcpValid[64];
static uint32_t g_activeRcpCopy[64][2048];
static bool g_activeRsqrtValid[64];
static uint32_t g_activeRsqrtCopy[64][4096];

/* execute the pending draw batch before anything that reads or changes what it uses */
#define SLCU_FLUSH()                                                                     \
    do {                                                                                 \
        if (c && !c->batch.descs.empty()) { const int frc_ = flush_draws(c); if (frc_) return frc_; } \
    } while (0)

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
            return SLCU_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

static int flush_draws(slcu_context* c);
static void geom_release(slcu_context* c)
{
    for (int t = 0; t < SL_MAX_RANKS; ++t)
    {
        if (c->geom.peer[t] && c->geom.ipcMapped[t]) cudaIpcCloseMemHandle(c->geom.peer[t]);
        c->geom.peer[t] = nullptr;
        c->geom.ipcMapped[t] = false;
    }
    if (c->geom.base) cudaFree(c->geom.base);
    c->geom.base = nullptr;
    c->geom.arenaBytes = 0;
    c->geom.epoch = 0;
}

static int ensure_rcp_table(slcu_context* c)
{
    if (!c->rcpValid) { c->err = "no rcp table"; return SLCU_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    const int d = c->device & 63;
    if (g_activeRcpValid[d] && memcmp(g_activeRcpCopy[d], c->rcpTable, sizeof(c->rcpTable)) == 0) return 0;
    CK(cudaDeviceSynchronize()); /* table switch: rare (tests pinning a table) */
    CK(cudaMemcpyToSymbol(g_rcp_table, c->rcpTable, sizeof(c->rcpTable)));
    memcpy(g_activeRcpCopy[d], c->rcpTable, sizeof(c->rcpTable));
    g_activeRcpValid[d] = true;
    for (const RtModule& m : c->rtModules) CK(cudaMemcpy(m.rcpTable, c->rcpTable, sizeof(c->rcpTable), cudaMemcpyHostToDevice)); /* every module has its own g_rcp_table */
    return 0;
}

static int ensure_rsqrt_table(slcu_context* c)
{
    if (!c->rsqrtValid) { c->err = "no rsqrt table (line primitives need one: set_rsqrt_table)"; return SLCU_ERR_UNSUPPORTED; }
    CK(cudaSetDevice(c->device));
    const int d = c->device & 63;
    if (g_activeRsqrtValid[d] && memcmp(g_activeRsqrtCopy[d], c->rsqrtTable, sizeof(c->rsqrtTable)) == 0) return 0;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyToSymbol(g_rsqrt_table, c->rsqrtTable, sizeof(c->rsqrtTable)));
    memcpy(g_activeRsqrtCopy[d], c->rsqrtTable, sizeof(c->rsqrtTable));
    g_activeRsqrtValid[d] = true;
    for (const RtModule& m : c->rtModules) CK(cudaMemcpy(m.rsqrtTable, c->rsqrtTable, sizeof(c->rsqrtTable), cudaMemcpyHostToDevice));
    return 0;
}

static inline unsigned chan_bytes(int type) { static const unsigned b[7] = {1, 2, 4, 8, 2, 4, 8}; return b[type % 7]; }
/* the packed types SL_COLOR_RGB_332 ... SL_COLOR_RGBA_1010102 (28..32, SL_ColorCompressed.hpp): one integer per texel */
static inline bool type_is_packed(int type) { return type >= 28; }
static inline unsigned type_channels(int type) { return type_is_packed(type) ? (type <= 29 ? 3u : 4u) : (unsigned)(type / 7 + 1); }
static inline unsigned type_bpt(int type)
{
    static const unsigned packed[5] = {1, 2, 2, 2, 4};
    return type_is_packed(type) ? packed[type - 28] : chan_bytes(type) * (unsigned)(type / 7 + 1);
}

static int upload_texture_descs(slcu_context* c)
{
    const size_t n = c->textures.size();
    if (n > c->devTexturesCap)
    {
        /* descriptors must not move once a draw may be in flight: sync, then grow */
        CK(cudaStreamSynchronize(c->stream));
        if (c->devTextures) cudaFree(c->devTextures);
        c->devTexturesCap = n + 64;
        CK(cudaMalloc(&c->devTextures, c->devTexturesCap * sizeof(SL_Texture)));
    }
    std::vector<SL_Texture> h(n);
    for (size_t i = 0; i < n; ++i)
    {
        const HostTexture& t = c->textures[i];
        h[i].mWidth = t.w; h[i].mHeight = t.h; h[i].mDepth = t.d; h[i].mBpt = t.bpt; h[i].mChannels = t.channels;
        h[i].mType = t.type; h[i].mTexels = t.dev;
    }
    if (n) CK(cudaMemcpyAsync(c->devTextures, h.data(), n * sizeof(SL_Texture), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream)); /* h goes out of scope */
    return 0;
}

/* ===========================================================================
 * Per-draw host scalars (src/SL_ViewportState.cpp:17-111)
 * =========================================================================*/
static void host_scissor_matrix(const slcu_context* c, float fboW, float fboH, float m[16])
{
    const uint32_t* T = c->rcpTable;
    const float inv[4] = {h_rcp(T, fboW), h_rcp(T, fboH), h_rcp(T, fboW), h_rcp(T, fboH)};
    float x = (float)c->scissor[0] * inv[0], y = (float)c->scissor[1] * inv[1];
    float w = (float)c->scissor[2] * inv[2], h = (float)c->scissor[3] * inv[3];
    if (x < 0.f) { w += x; x = 0.f; }
    if (y < 0.f) { h += y; y = 0.f; }
    w = h_min(1.f - x, w);
    h = h_min(1.f - y, h);
    const float nm00 = h_rcp(T, w), nm11 = h_rcp(T, h);
    const float nm30 = nm00 - 1.f, nm31 = nm11 - 1.f;
    const float m30 = fmaf(x, (-2.f * nm00), nm30);
    const float m31 = fmaf(y, (-2.f * nm11), nm31);
    const float r[16] = {nm00, 0.f, 0.f, 0.f, 0.f, nm11, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, m30, m31, 0.f, 1.f};
    memcpy(m, r, sizeof(r));
}

static void host_viewport_rect(const slcu_context* c, float fboW, float fboH, float d[4])
{
    const float s0 = (float)c->scissor[0], s1 = (float)c->scissor[1], s2 = (float)c->scissor[2], s3 = (float)c->scissor[3];
    d[0] = (float)c->viewport[0]; d[1] = (float)c->viewport[1]; d[2] = (float)c->viewport[2]; d[3] = (float)c->viewport[3];
    /* clamp(n,lo,hi) = max(min(n,hi),lo); min(a,b,c) = min(c, min(a,b)) (x86/scalarf_utils_impl.h:32-38,88) */
    d[0] = h_max(h_min(d[0] + s0, fboW), 0.f);
    d[1] = h_max(h_min(d[1] + s1, fboH), 0.f);
    d[2] = h_min(fboW - d[0], h_min(d[2], (s0 + s2) - d[0]));
    d[3] = h_min(fboH - d[1], h_min(d[3], (s1 + s3) - d[1]));
}

/* ===========================================================================
 * run_shader_processors replacement
 * =========================================================================*/
static void snapshot_uniforms(slcu_context* c, size_t ubo, SL_UniformBuffer& out)
{
    memcpy(out.mBytes, c->ubos[ubo].bytes, 1024);
    for (const UboFixup& f : c->fixups)
    {
        if (f.ubo != ubo) continue;
        const void* ptr;
        if (f.kind == 0) ptr = c->devTextures + f.id;
        else ptr = c->vbos[f.id].ptr + f.byteOffset;
        memcpy(out.mBytes + f.offset, &ptr, sizeof(ptr));
    }
}

/* one launch of a pipeline kernel: the build-time instantiation, or the run-time one of a registered source */
template <typename Params> static cudaError_t launch_rt(cudaKernel_t k, unsigned grid, unsigned block, size_t smem, cudaStream_t s, const Params& p)
{
    void* args[] = {const_cast<Params*>(&p)};
    return cudaLaunchKernel(reinterpret_cast<const void*>(k), dim3(grid), dim3(block), args, smem, s);
}
static cudaError_t launch_vertex_stage(const HostShader& sh, unsigned grid, cudaStream_t s, const SLVertexParams& p, bool cache)
{
    if (sh.rtV.vertex[0]) return launch_rt(sh.rtV.vertex[cache ? 1 : 0], grid, SL_VB_THREADS, cache ? p.cacheBytes : 0, s, p);
    return sh.vs(grid, s, p, cache);
}
static cudaError_t launch_shade_stage(const HostShader& sh, unsigned grid, cudaStream_t s, const SLShadeParams& p)
{
    if (sh.rtF.shade) return launch_rt(sh.rtF.shade, grid, SL_SHADE_WARPS * 32, 0, s, p);
    return sh.fs(grid, s, p);
}
static cudaError_t launch_point_vertex_stage(const HostShader& sh, unsigned grid, cudaStream_t s, const SLPointParams& p)
{
    if (sh.rtV.pointVertex) return launch_rt(sh.rtV.pointVertex, grid, 256, 0, s, p);
    return sh.pointVs(grid, s, p);
}
static cudaError_t launch_line_vertex_stage(const HostShader& sh, unsigned grid, cudaStream_t s, const SLLineParams& p)
{
    if (sh.rtV.lineVertex) return launch_rt(sh.rtV.lineVertex, grid, SL_LINE_BLOCK, 0, s, p);
    return sh.lineVs(grid, s, p);
}
static cudaError_t launch_line_shade_stage(const HostShader& sh, unsigned grid, cudaStream_t s, const SLLineShadeParams& p)
{
    if (sh.rtF.lineShade) return launch_rt(sh.rtF.lineShade, grid, 128, 0, s, p);
    return sh.lineFs(grid, s, p);
}
static cudaError_t launch_point_shade_stage(const HostShader& sh, unsigned grid, cudaStream_t s, const SLPointShadeParams& p)
{
    if (sh.rtF.pointShade) return launch_rt(sh.rtF.pointShade, grid, 128, 0, s, p);
    return sh.pointFs(grid, s, p);
}
static cudaError_t launch_wire_shade_stage(const HostShader& sh, unsigned grid, cudaStream_t s, const SLWireShadeParams& p)
{
    if (sh.rtF.wireShade) return launch_rt(sh.rtF.wireShade, grid, 128, 0, s, p);
    return sh.wireFs(grid, s, p);
}

/* The ordering step of the point / line / wireframe paths (sl_sort.cuh): stable LSD radix sort of `n` 64-bit keys (and
 * values) on bits [bitBegin, bitEnd).  keys / vals are double buffers (A at [0, n), B at [n, 2n)); returns 0 or 1 = the half
 * that holds the result, or a negative error. */
static int sort_u64(slcu_context* c, unsigned long long* keys, unsigned long long* vals, unsigned long long n, int bitBegin, int bitEnd)
{
    if (n > 0xfffffff0ull) { c->err = "too many keys for one sort"; return SLCU_ERR_UNSUPPORTED; }
    const int bitsTotal = bitEnd - bitBegin;
    if (bitsTotal <= 0) return 0;
    const int passes = (bitsTotal + SL_SORT_MAX_BITS - 1) / SL_SORT_MAX_BITS;
    const int perPass = (bitsTotal + passes - 1) / passes;
    const unsigned numTiles = (unsigned)((n + SL_SORT_TILE - 1) / SL_SORT_TILE);
    const size_t histBytes = ((size_t)(1u << perPass) * numTiles * sizeof(uint32_t) + 255) & ~(size_t)255;
    CK(c->pointSortTemp.ensure(histBytes + (size_t)(1u << perPass) * sizeof(uint32_t)));
    const unsigned smem = (unsigned)((SL_SORT_WARPS + 1) * (1u << perPass) * sizeof(uint32_t));
    static bool attrSet[64] = {};
    if (!attrSet[c->device & 63])
    {
        CK(cudaFuncSetAttribute(k_sort_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((SL_SORT_WARPS + 1) * (1u << SL_SORT_MAX_BITS) * sizeof(uint32_t))));
        attrSet[c->device & 63] = true;
    }
    int cur = 0;
    for (int shift = bitBegin; shift < bitEnd; shift += perPass)
    {
        const int bits = bitEnd - shift < perPass ? bitEnd - shift : perPass;
        SLSortPass sp;
        sp.keysIn = keys + (size_t)cur * n;
        sp.keysOut = keys + (size_t)(cur ^ 1) * n;
        sp.valsIn = vals ? vals + (size_t)cur * n : nullptr;
        sp.valsOut = vals ? vals + (size_t)(cur ^ 1) * n : nullptr;
        sp.hist = reinterpret_cast<uint32_t*>(c->pointSortTemp.buf.ptr);
        sp.totals = reinterpret_cast<uint32_t*>(c->pointSortTemp.buf.ptr + histBytes);
        sp.n = n;
        sp.numTiles = numTiles;
        sp.numDigits = 1u << bits;
        sp.shift = shift;
        KernelScope ks(c, SLCU_K_SCAN);
        k_sort_hist<<<numTiles, SL_SORT_THREADS, 0, c->stream>>>(sp);
        k_sort_rows<<<sp.numDigits, SL_SORT_THREADS, 0, c->stream>>>(sp.hist, sp.totals, numTiles);
        k_sort_scatter<<<numTiles, SL_SORT_THREADS, smem, c->stream>>>(sp);
        CK(cudaGetLastError());
        c->stats.kernelLaunches += 2;
        cur ^= 1;
    }
    return cur;
}

/* Draws are not executed one by one.  Consecutive draws that use the same shader object and framebuffer are collected
 * into a batch (their uniforms, viewport and scissor are snapshotted at submission, like the reference's synchronous draw
 * would have read them) and the whole batch goes through the pipeline at once when anything else needs the result: a draw
 * with another shader or framebuffer, a clear, a blit, a read-back, an upload into a resource, slcu_sync.  Primitive order
 * across the draws of a batch is submission order, so the frame is bit-identical to draw-by-draw execution; what changes
 * is that a 256-draw frame fills the GPU with 8 launches instead of 2,000 under-filled ones. */
static const unsigned kMaxBatchBlocks = 24576; /* 24 M primitives */
static const size_t kMaxBatchDraws = 65535;    /* recDraw is 16 bits */

static int flush_draws(slcu_context* c);
static int run_points(slcu_context* c, const slcu_mesh* meshes, size_t numMeshes, size_t numInstances, size_t shaderId, size_t fboId);
static int run_lines(slcu_context* c, const slcu_mesh* meshes, size_t numMeshes, size_t numInstances, size_t shaderId, size_t fboId);

static int run_draw(slcu_context* c, const slcu_mesh* meshes, size_t numMeshes, size_t numInstances, size_t shaderId, size_t fboId)
{
    if (!meshes || !numMeshes) return 0;
    if (shaderId >= c->shaders.size() || fboId >= c->fbos.size()) { c->err = "bad shader/fbo id"; return SLCU_ERR_ARG; }
    /* the processor is chosen from the first mesh's mode (src/SL_ProcessorPool.cpp:286-293, 339-346) */
    switch (meshes[0].mode)
    {
        case 0x10u: case 0x30u: break;                                                                /* triangles */
        case 0x50u: case 0x70u: break;                                                                /* wireframe triangles */
        case 0x01u: case 0x03u: return run_points(c, meshes, numMeshes, numInstances, shaderId, fboId); /* points */
        case 0x04u: case 0x0cu: return run_lines(c, meshes, numMeshes, numInstances, shaderId, fboId);                 /* lines */
        default: c->err = "unknown render mode"; return SLCU_ERR_ARG;
    }
    const HostShader& sh = c->shaders[shaderId];
    const HostFbo& fbo = c->fbos[fboId];
    if (fbo.depth == (size_t)-1 || fbo.depth >= c->textures.size()) { c->err = "fbo has no depth attachment"; return SLCU_ERR_ARG; }
    const HostTexture& depthTex = c->textures[fbo.depth];
    /* The reference addresses pixels linearly inside an allocation padded to ~4x the texels (src/SL_Texture.cpp:28-54), so
     * clip-generated triangles that end above the viewport shade fragments in rows >= h.  Nothing can read those back and
     * they cannot influence a texel below h*w, so rows >= h are not rasterised (and not counted, see slcu_stats). */
    const unsigned H = depthTex.h;
    if (H > (unsigned)SL_MAX_ROWS) { c->err = "framebuffer taller than SL_MAX_ROWS"; return SLCU_ERR_UNSUPPORTED; }
    if (depthTex.w > 32767) { c->err = "framebuffer wider than 32767"; return SLCU_ERR_UNSUPPORTED; }
    if (depthTex.bpt != 2 && depthTex.bpt != 4 && depthTex.bpt != 8) { c->err = "depth attachment must be R_HALF/R_FLOAT/R_DOUBLE"; return SLCU_ERR_ARG; }
    for (int i = 0; i < sh.numOutputs; ++i)
    {
        if ((unsigned)i >= fbo.numColors || fbo.colors[i] >= c->textures.size()) { c->err = "shader writes an unattached colour buffer"; return SLCU_ERR_ARG; }
        const HostTexture& t = c->textures[fbo.colors[i]];
        if (t.w != depthTex.w || t.h != depthTex.h) { c->err = "colour and depth attachments differ in size (SL_Framebuffer::valid -7 / -8)"; return SLCU_ERR_ARG; }
    }
    const float fboW = (float)depthTex.w, fboH = (float)depthTex.h;
    const bool wire = (meshes[0].mode == 0x50u) || (meshes[0].mode == 0x70u); /* mRenderMode = meshes->mode (src/SL_ProcessorPool.cpp:301,354) */

    /* one table entry per mesh (draw / draw_multiple) or one for all instances (draw_instanced) */
    const size_t launches = (numInstances == 1) ? numMeshes : 1;
    for (size_t i = 0; i < launches; ++i)
    {
        const slcu_mesh& m = meshes[i];
        if (m.vaoId >= c->vaos.size()) { c->err = "bad vao id"; return SLCU_ERR_ARG; }
        const HostVao& vao = c->vaos[m.vaoId];
        if (vao.vbo >= c->vbos.size()) { c->err = "vao has no vertex buffer"; return SLCU_ERR_ARG; }
        const bool usingIndices = (m.mode == 0x30u) || (m.mode == 0x70u);
        if (usingIndices && vao.ibo >= c->ibos.size()) { c->err = "indexed draw without index buffer"; return SLCU_ERR_ARG; }
    }
    c->stats.drawCalls++;
    for (size_t i = 0; i < launches; ++i)
    {
        const slcu_mesh& m = meshes[i];
        /* the reference loop is `for (i = begin; i < end; i += 3)` (src/SL_TriProcessor.cpp:574) */
        const unsigned long long primsPerInstance = (m.elementEnd - m.elementBegin + 2) / 3;
        const unsigned long long prims = primsPerInstance * (numInstances == 1 ? 1 : numInstances);
        /* a vertex block holds primitives of one instance only (k_vertex) */
        const unsigned long long blocksPerInstance = (primsPerInstance + SL_VB_PRIMS - 1) / SL_VB_PRIMS;
        const unsigned long long blocks64 = blocksPerInstance * (numInstances == 1 ? 1 : numInstances);
        c->stats.trisSubmitted += prims;
        if (!blocks64) continue;
        if (blocks64 > 0x3fffffull) { c->err = "draw too large for one batch"; return SLCU_ERR_UNSUPPORTED; }
        const unsigned blocks = (unsigned)blocks64;
        Batch& B = c->batch;
        if (!B.descs.empty() && (B.shaderId != shaderId || B.fboId != fboId || B.wire != wire || B.numBlocks + blocks > kMaxBatchBlocks || B.descs.size() >= kMaxBatchDraws))
        {
            const int rc = flush_draws(c);
            if (rc) return rc;
        }
        B.shaderId = shaderId;
        B.fboId = fboId;
        B.wire = wire;
        const HostVao& vao = c->vaos[m.vaoId];
        const bool usingIndices = (m.mode == 0x30u) || (m.mode == 0x70u);
        B.descs.emplace_back();
        SLDrawDesc& d = B.descs.back();
        snapshot_uniforms(c, sh.ubo, d.uniforms);
        d.vao = vao.vao;
        d.vbo.mBuffer = c->vbos[vao.vbo].ptr;
        d.usingIndices = usingIndices ? 1 : 0;
        d.ibo = usingIndices ? c->ibos[vao.ibo].buf.ptr : nullptr;
        d.iboType = usingIndices ? c->ibos[vao.ibo].type : 2;
        host_scissor_matrix(c, fboW, fboH, d.scissor);
        host_viewport_rect(c, fboW, fboH, d.viewport);
        d.elementBegin = m.elementBegin;
        d.primsPerInstance = primsPerInstance;
        d.numPrims = prims;
        d.firstBlock = B.numBlocks;
        d.blocksPerInstance = (unsigned)blocksPerInstance;
        {
            const unsigned long long mi = usingIndices ? c->ibos[vao.ibo].maxIndex : ~0ull;
            if (mi > B.maxIndex) B.maxIndex = mi;
        }
        {
            const ptrdiff_t stride = vao.vao.mBindings[0].stride;
            const unsigned long long nv = stride > 0 ? (unsigned long long)(c->vbos[vao.vbo].bytes / (size_t)stride) : ~0ull;
            if (nv > B.maxVerts) B.maxVerts = nv;
        }
        B.blockDraw.insert(B.blockDraw.end(), blocks, (uint32_t)(B.descs.size() - 1));
        B.numBlocks += blocks;
        B.totalPrims += prims;
    }
    if (c->syncDraws) return flush_draws(c);
    return 0;
}

/* SL_PointProcessor + SL_PointRasterizer (src/SL_PointProcessor.cpp, src/SL_PointRasterizer.cpp): executed at submission,
 * after the triangle batch that precedes it; one (vertex, sort, shade) pass per mesh / per instanced draw. */
static int run_points(slcu_context* c, const slcu_mesh* meshes, size_t numMeshes, size_t numInstances, size_t shaderId, size_t fboId)
{
    int rc = flush_draws(c);
    if (rc) return rc;
    const HostShader& sh = c->shaders[shaderId];
    const HostFbo& fbo = c->fbos[fboId];
    if (fbo.depth == (size_t)-1 || fbo.depth >= c->textures.size()) { c->err = "fbo has no depth attachment"; return SLCU_ERR_ARG; }
    const HostTexture& depthTex = c->textures[fbo.depth];
    if (depthTex.w > 32767 || depthTex.h > 65535) { c->err = "framebuffer too large"; return SLCU_ERR_UNSUPPORTED; }
    if (depthTex.bpt != 2 && depthTex.bpt != 4 && depthTex.bpt != 8) { c->err = "depth attachment must be R_HALF/R_FLOAT/R_DOUBLE"; return SLCU_ERR_ARG; }
    for (int i = 0; i < sh.numOutputs; ++i)
    {
        if ((unsigned)i >= fbo.numColors || fbo.colors[i] >= c->textures.size()) { c->err = "shader writes an unattached colour buffer"; return SLCU_ERR_ARG; }
        const HostTexture& t = c->textures[fbo.colors[i]];
        if (t.w != depthTex.w || t.h != depthTex.h) { c->err = "colour and depth attachments differ in size (SL_Framebuffer::valid -7 / -8)"; return SLCU_ERR_ARG; }
    }
    const size_t launches = (numInstances == 1) ? numMeshes : 1;
    for (size_t i = 0; i < launches; ++i)
    {
        const slcu_mesh& m = meshes[i];
        if (m.vaoId >= c->vaos.size()) { c->err = "bad vao id"; return SLCU_ERR_ARG; }
        const HostVao& vao = c->vaos[m.vaoId];
        if (vao.vbo >= c->vbos.size()) { c->err = "vao has no vertex buffer"; return SLCU_ERR_ARG; }
        if (m.mode == 0x03u && vao.ibo >= c->ibos.size()) { c->err = "indexed draw without index buffer"; return SLCU_ERR_ARG; }
    }
    rc = ensure_rcp_table(c);
    if (rc) return rc;
    cudaSetDevice(c->device);
    c->stats.drawCalls++;
    CK(cudaMemsetAsync(c->counters, 0, SL_COUNTERS_BATCH_BYTES, c->stream)); /* the per-batch part; k_commit adds it to the totals */
    const float fboW = (float)depthTex.w, fboH = (float)depthTex.h;
    const int NV = sh.numVaryings;
    unsigned pixBits = 1;
    while ((1ull << pixBits) <= (unsigned long long)depthTex.w * depthTex.h) ++pixBits; /* 2^pixBits - 1 >= w*h: the invisible key sorts last */
    for (size_t i = 0; i < launches; ++i)
    {
        const slcu_mesh& m = meshes[i];
        const unsigned long long per = m.elementEnd > m.elementBegin ? (unsigned long long)(m.elementEnd - m.elementBegin) : 0ull;
        const unsigned long long numPoints = per * (numInstances == 1 ? 1 : numInstances);
        c->stats.trisSubmitted += (per / 3) * (numInstances == 1 ? 1 : numInstances);
        if (!numPoints) continue;
        if (numPoints > 0x7ffffff0ull) { c->err = "draw too large for one pass"; return SLCU_ERR_UNSUPPORTED; }
        const HostVao& vao = c->vaos[m.vaoId];
        const bool usingIndices = m.mode == 0x03u; /* per mesh (src/SL_PointProcessor.cpp:86) */
        SLDrawDesc d;
        memset(&d, 0, sizeof(d));
        snapshot_uniforms(c, sh.ubo, d.uniforms);
        d.vao = vao.vao;
        d.vbo.mBuffer = c->vbos[vao.vbo].ptr;
        d.usingIndices = usingIndices ? 1 : 0;
        d.ibo = usingIndices ? c->ibos[vao.ibo].buf.ptr : nullptr;
        d.iboType = usingIndices ? c->ibos[vao.ibo].type : 2;
        host_scissor_matrix(c, fboW, fboH, d.scissor);
        host_viewport_rect(c, fboW, fboH, d.viewport);
        d.elementBegin = m.elementBegin;
        d.primsPerInstance = per;
        d.numPrims = numPoints;
        CK(c->drawTable.ensure(sizeof(SLDrawDesc)));
        CK(cudaMemcpyAsync(c->drawTable.buf.ptr, &d, sizeof(SLDrawDesc), cudaMemcpyHostToDevice, c->stream)); /* pageable source: staged before the call returns */
        CK(c->pointKeys.ensure((size_t)numPoints * 2 * sizeof(unsigned long long)));
        CK(c->pointRecs.ensure((size_t)numPoints * (1 + NV) * sizeof(float4)));
        unsigned long long* keysA = reinterpret_cast<unsigned long long*>(c->pointKeys.buf.ptr);
        SLPointParams vp;
        vp.draw = reinterpret_cast<const SLDrawDesc*>(c->drawTable.buf.ptr);
        vp.numPoints = numPoints;
        vp.elementsPerInstance = per;
        vp.numVaryings = NV;
        vp.width = depthTex.w;
        vp.height = depthTex.h;
        vp.rank = c->partRank;
        vp.world = c->partWorld;
        vp.stripRows = c->partStripRows;
        vp.keys = keysA;
        vp.points = reinterpret_cast<float4*>(c->pointRecs.buf.ptr);
        vp.counters = c->counters;
        { KernelScope ks(c, SLCU_K_VERTEX); CK(launch_point_vertex_stage(sh, (unsigned)((numPoints + 255ull) / 256ull), c->stream, vp)); }
        /* stable LSD sort on the pixel bits only: inside a pixel the keys stay in submission order */
        const int sortedHalf = sort_u64(c, keysA, nullptr, numPoints, 32, 32 + (int)pixBits);
        if (sortedHalf < 0) return sortedHalf;

        SLPointShadeParams sp;
        sp.draw = vp.draw;
        sp.keys = keysA + (size_t)sortedHalf * numPoints;
        sp.points = vp.points;
        sp.numPoints = numPoints;
        sp.numVaryings = NV;
        sp.numOutputs = sh.numOutputs;
        sp.blend = sh.blend;
        sp.depthTest = sh.depthTest;
        sp.depthMask = sh.depthMask;
        sp.depth = depthTex.dev;
        sp.depthBpt = depthTex.bpt;
        for (int k = 0; k < 4; ++k)
        {
            sp.color[k] = nullptr;
            sp.colorType[k] = 0;
            if (k < sh.numOutputs) { const HostTexture& t = c->textures[fbo.colors[k]]; sp.color[k] = t.dev; sp.colorType[k] = t.type; }
        }
        sp.counters = c->counters;
        { KernelScope ks(c, SLCU_K_SHADE); CK(launch_point_shade_stage(sh, (unsigned)((numPoints + 127ull) / 128ull), c->stream, sp)); }
    }
    k_commit<<<1, 1, 0, c->stream>>>(c->counters);
    CK(cudaGetLastError());
    if (sh.depthMask) c->textures[fbo.depth].depthWritesSinceClear++;
    return 0;
}


/* SL_LineProcessor + SL_LineRasterizer (src/SL_LineProcessor.cpp, src/SL_LineRasterizer.cpp): executed at submission, after
 * the triangle batch that precedes it; one (vertex, emit, sort, shade) pass per mesh / per instanced draw (sl_lines.cuh). */
static int run_lines(slcu_context* c, const slcu_mesh* meshes, size_t numMeshes, size_t numInstances, size_t shaderId, size_t fboId)
{
    int rc = flush_draws(c);
    if (rc) return rc;
    const HostShader& sh = c->shaders[shaderId];
    const HostFbo& fbo = c->fbos[fboId];
    if (fbo.depth == (size_t)-1 || fbo.depth >= c->textures.size()) { c->err = "fbo has no depth attachment"; return SLCU_ERR_ARG; }
    const HostTexture& depthTex = c->textures[fbo.depth];
    if (depthTex.w > 32767 || depthTex.h > 32767) { c->err = "framebuffer too large"; return SLCU_ERR_UNSUPPORTED; }
    if (depthTex.bpt != 2 && depthTex.bpt != 4 && depthTex.bpt != 8) { c->err = "depth attachment must be R_HALF/R_FLOAT/R_DOUBLE"; return SLCU_ERR_ARG; }
    for (int i = 0; i < sh.numOutputs; ++i)
    {
        if ((unsigned)i >= fbo.numColors || fbo.colors[i] >= c->textures.size()) { c->err = "shader writes an unattached colour buffer"; return SLCU_ERR_ARG; }
        const HostTexture& t = c->textures[fbo.colors[i]];
        if (t.w != depthTex.w || t.h != depthTex.h) { c->err = "colour and depth attachments differ in size (SL_Framebuffer::valid -7 / -8)"; return SLCU_ERR_ARG; }
    }
    const size_t launches = (numInstances == 1) ? numMeshes : 1;
    for (size_t i = 0; i < launches; ++i)
    {
        const slcu_mesh& m = meshes[i];
        if (m.vaoId >= c->vaos.size()) { c->err = "bad vao id"; return SLCU_ERR_ARG; }
        const HostVao& vao = c->vaos[m.vaoId];
        if (vao.vbo >= c->vbos.size()) { c->err = "vao has no vertex buffer"; return SLCU_ERR_ARG; }
        if (m.mode == 0x0cu && vao.ibo >= c->ibos.size()) { c->err = "indexed draw without index buffer"; return SLCU_ERR_ARG; }
    }
    rc = ensure_rcp_table(c);
    if (rc) return rc;
    rc = ensure_rsqrt_table(c);
    if (rc) return rc;
    cudaSetDevice(c->device);
    c->stats.drawCalls++;
    CK(cudaMemsetAsync(c->counters, 0, SL_COUNTERS_BATCH_BYTES, c->stream)); /* the per-batch part; k_commit adds it to the totals */
    const float fboW = (float)depthTex.w, fboH = (float)depthTex.h;
    const int NV = sh.numVaryings;
    unsigned pixBits = 1;
    while ((1ull << pixBits) <= (unsigned long long)depthTex.w * depthTex.h) ++pixBits; /* 2^pixBits - 1 >= w*h: the invisible key sorts last */
    for (size_t i = 0; i < launches; ++i)
    {
        const slcu_mesh& m = meshes[i];
        const unsigned long long elems = m.elementEnd > m.elementBegin ? (unsigned long long)(m.elementEnd - m.elementBegin) : 0ull;
        const unsigned long long per = (elems + 1ull) / 2ull; /* for (i = begin; i < end; i += 2) */
        const unsigned long long numLines = per * (numInstances == 1 ? 1 : numInstances);
        c->stats.trisSubmitted += (elems / 3) * (numInstances == 1 ? 1 : numInstances);
        if (!numLines) continue;
        if (numLines > 0x3ffffff0ull) { c->err = "draw too large for one pass"; return SLCU_ERR_UNSUPPORTED; }
        const HostVao& vao = c->vaos[m.vaoId];
        const bool usingIndices = m.mode == 0x0cu; /* per mesh (src/SL_LineProcessor.cpp:205) */
        if ((elems & 1ull) && usingIndices && m.elementEnd >= c->ibos[vao.ibo].count)
        {
            /* for (i = begin; i < end; i += 2) fetches element `end` for the last segment (src/SL_LineProcessor.cpp:236-251):
             * past the index buffer the reference draws whatever its heap holds there; here the index would address
             * arbitrary device memory */
            c->err = "line draw with an odd element count: the last segment's second element lies past the index buffer";
            return SLCU_ERR_ARG;
        }
        SLDrawDesc d;
        memset(&d, 0, sizeof(d));
        snapshot_uniforms(c, sh.ubo, d.uniforms);
        d.vao = vao.vao;
        d.vbo.mBuffer = c->vbos[vao.vbo].ptr;
        d.usingIndices = usingIndices ? 1 : 0;
        d.ibo = usingIndices ? c->ibos[vao.ibo].buf.ptr : nullptr;
        d.iboType = usingIndices ? c->ibos[vao.ibo].type : 2;
        host_scissor_matrix(c, fboW, fboH, d.scissor);
        host_viewport_rect(c, fboW, fboH, d.viewport);
        d.elementBegin = m.elementBegin;
        d.primsPerInstance = per;
        d.numPrims = numLines;
        CK(c->drawTable.ensure(sizeof(SLDrawDesc)));
        CK(cudaMemcpyAsync(c->drawTable.buf.ptr, &d, sizeof(SLDrawDesc), cudaMemcpyHostToDevice, c->stream)); /* pageable source: staged before the call returns */
        const unsigned numBlocks = (unsigned)((numLines + SL_LINE_BLOCK - 1) / SL_LINE_BLOCK);
        CK(c->records.ensure((size_t)numLines * (2 + 2 * NV) * sizeof(float4)));
        CK(c->cnt.ensure((size_t)numLines * sizeof(uint32_t)));
        CK(c->segSums.ensure((size_t)numBlocks * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(&c->counters->pixelsCovered, 0, sizeof(unsigned long long), c->stream));

        SLLineParams lp;
        memset(&lp, 0, sizeof(lp));
        lp.draw = reinterpret_cast<const SLDrawDesc*>(c->drawTable.buf.ptr);
        lp.numLines = numLines;
        lp.linesPerInstance = per;
        lp.numVaryings = NV;
        lp.width = depthTex.w;
        lp.height = depthTex.h;
        lp.rank = c->partRank;
        lp.world = c->partWorld;
        lp.stripRows = c->partStripRows;
        lp.lines = reinterpret_cast<float4*>(c->records.buf.ptr);
        lp.lineCount = reinterpret_cast<uint32_t*>(c->cnt.buf.ptr);
        lp.blockSum = reinterpret_cast<unsigned long long*>(c->segSums.buf.ptr);
        lp.counters = c->counters;
        { KernelScope ks(c, SLCU_K_VERTEX); CK(launch_line_vertex_stage(sh, numBlocks, c->stream, lp)); }
        { KernelScope ks(c, SLCU_K_SCAN); k_line_offsets<<<1, 1024, 0, c->stream>>>(lp.blockSum, numBlocks); }
        CK(cudaGetLastError());
        /* the candidates' number sizes the next buffers: one host round trip per line draw */
        CK(cudaMemcpyAsync(c->hostCounters, c->counters, sizeof(SLCounters), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        const unsigned long long numCands = c->hostCounters->pixelsCovered;
        if (!numCands) continue;
        if (numCands > 0xfffffff0ull) { c->err = "too many line fragments for one pass"; return SLCU_ERR_UNSUPPORTED; }
        CK(c->pointKeys.ensure((size_t)numCands * 2 * sizeof(unsigned long long)));
        CK(c->pointRecs.ensure((size_t)numCands * sizeof(float4)));
        unsigned long long* keysA = reinterpret_cast<unsigned long long*>(c->pointKeys.buf.ptr);
        lp.keys = keysA;
        lp.cands = reinterpret_cast<float4*>(c->pointRecs.buf.ptr);
        { KernelScope ks(c, SLCU_K_SPAN_FILL); k_line_emit<<<numBlocks, SL_LINE_BLOCK, 0, c->stream>>>(lp); }
        CK(cudaGetLastError());
        /* stable LSD sort on the texel bits only: inside a texel the keys stay in submission order */
        const int sortedHalf = sort_u64(c, keysA, nullptr, numCands, 32, 32 + (int)pixBits);
        if (sortedHalf < 0) return sortedHalf;

        SLLineShadeParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.draw = lp.draw;
        sp.keys = keysA + (size_t)sortedHalf * numCands;
        sp.cands = lp.cands;
        sp.lines = lp.lines;
        sp.numCands = numCands;
        sp.numVaryings = NV;
        sp.numOutputs = sh.numOutputs;
        sp.blend = sh.blend;
        sp.depthTest = sh.depthTest;
        sp.depthMask = sh.depthMask;
        sp.depth = depthTex.dev;
        sp.depthBpt = depthTex.bpt;
        for (int k = 0; k < 4; ++k)
        {
            sp.color[k] = nullptr;
            sp.colorType[k] = 0;
            if (k < sh.numOutputs) { const HostTexture& t = c->textures[fbo.colors[k]]; sp.color[k] = t.dev; sp.colorType[k] = t.type; }
        }
        sp.counters = c->counters;
        { KernelScope ks(c, SLCU_K_SHADE); CK(launch_line_shade_stage(sh, (unsigned)((numCands + 127ull) / 128ull), c->stream, sp)); }
    }
    k_commit<<<1, 1, 0, c->stream>>>(c->counters);
    CK(cudaGetLastError());
    if (sh.depthMask) c->textures[fbo.depth].depthWritesSinceClear++;
    return 0;
}

static int enqueue_batch(slcu_context* c, Batch& B, bool& verify);

/* The counters of the batch that just ran are in hostCounters[0].  Returns true when the device refused it (abortBatch):
 * the sizes it needs are then known exactly and become the new minimum for every later batch. */
static bool batch_was_refused(slcu_context* c)
{
    const SLCounters& h = c->hostCounters[0];
    if (h.totalEntries) c->caps.avgSpan = h.pixelsCovered / h.totalEntries;
    if (getenv("SLCU_DEBUG"))
        fprintf(stderr, "[slcu] batch: records(overflow area)=%llu entries=%u handoffs=%llu pixels=%llu rowEndCuts=%llu abort=0x%x flags=0x%x\n",
                (unsigned long long)h.recAlloc, h.totalEntries, (unsigned long long)h.handoffs, (unsigned long long)h.pixelsCovered,
                (unsigned long long)h.rowEndCuts, h.abortBatch, h.errorFlags);
    if (!h.abortBatch) return false;
    if (h.abortBatch & SL_ABORT_PARTITION) c->refusedPartition = true;
    auto grow = [](unsigned long long& cap, unsigned long long need) { const unsigned long long want = need + need / 4 + 4096; if (want > cap) cap = want; };
    grow(c->caps.recOverflow, h.recAlloc);
    grow(c->caps.entries, h.totalEntries);
    grow(c->caps.chunks, h.pixelsCovered / SL_CHUNK_FRAGS);
    grow(c->caps.handoffSlots, 2ull * h.handoffs);
    return true;
}

/* Run one batch.  Its kernels are enqueued back to back with the buffer sizes earlier batches needed -- no host round trip
 * in the middle -- and its counters come back with one async copy behind the last kernel.  The host waits for that copy
 * (the end of the batch, where the caller's own join is in the usual clear / draw ... / read frame) and, when the device
 * refused the batch, runs it again with the sizes its own counters ask for.  Everything that depends on the batch
 * (another shader's draws, clears, blits, read-backs) is enqueued after this returns. */
static int run_batch(slcu_context* c, Batch& B)
{
    for (int attempt = 0; attempt < 6; ++attempt)
    {
        bool verify = false;
        const int rc = enqueue_batch(c, B, verify);
        if (rc) return rc;
        if (!verify) /* executed with its own host round trips (wireframe) or empty */
        {
            if (c->shaders[B.shaderId].depthMask) c->textures[c->fbos[B.fboId].depth].depthWritesSinceClear++;
            return 0;
        }
        CK(cudaMemcpyAsync(c->hostCounters, c->counters, sizeof(SLCounters), cudaMemcpyDeviceToHost, c->stream));
        if (!c->batchDone) CK(cudaEventCreateWithFlags(&c->batchDone, cudaEventDisableTiming));
        CK(cudaEventRecord(c->batchDone, c->stream));
        CK(cudaEventSynchronize(c->batchDone));
        HostTexture& depthTex = c->textures[c->fbos[B.fboId].depth];
        const HostShader& bsh = c->shaders[B.shaderId];
        if (!batch_was_refused(c))
        {
            if (bsh.depthMask) depthTex.depthWritesSinceClear++;
            return 0;
        }
        if (c->refusedPartition)
        {
            /* Sort-first: a span runs past the row end into a row of another rank, so this rank also rasterises the rows such
             * hand-overs come from.  Their depth texels were left alone by every earlier pass of the frame (another rank owns
             * them), so with an earlier depth-writing pass behind this one the tests in those rows would see stale depth: the
             * frame is refused loudly instead of diverging from the single-GPU result. */
            if (bsh.depthTest != 0 && depthTex.depthWritesSinceClear != 0)
            {
                c->refusedPartition = false;
                c->err = "sort-first partition: a span crosses into another rank's row in a pass that follows an earlier depth-writing pass on the same "
                         "attachment; that row's depth is not current on this rank (render this frame without the partition)";
                return SLCU_ERR_UNSUPPORTED;
            }
            B.partFull = true;
            c->refusedPartition = false;
        }
        B.geomLocal = true; /* shared vertex stage: the other ranks have moved on; this one repeats the batch by itself */
    }
    c->err = "batch does not fit its buffers after repeated growth";
    return SLCU_ERR_OOM;
}

static int flush_draws(slcu_context* c)
{
    if (c->batch.descs.empty()) return 0;
    const int rc = run_batch(c, c->batch);
    c->batch.clear();
    return rc;
}

static int enqueue_batch(slcu_context* c, Batch& B, bool& verify)
{
    verify = false;
    const size_t shaderId = B.shaderId, fboId = B.fboId;
    const HostShader& sh = c->shaders[shaderId];
    const HostFbo& fbo = c->fbos[fboId];
    const HostTexture& depthTex = c->textures[fbo.depth];
    const unsigned H = depthTex.h;
    int rc = ensure_rcp_table(c);
    if (rc) return rc;
    cudaSetDevice(c->device);
    c->stats.batches++;

    const unsigned rowPitch = (H + 31u) & ~31u;
    /* x-segments of the rasteriser: one warp per (row, segment), about 128 pixels each, at most 32 per row */
    unsigned segTarget = 512u;
    if (const char* e = getenv("SLCU_SEG_WIDTH")) { const int v = atoi(e); if (v >= 16) segTarget = (unsigned)v; }
    unsigned rasterSegs = (depthTex.w + segTarget - 1u) / segTarget;
    if (rasterSegs < 1u) rasterSegs = 1u;
    if (rasterSegs > 32u) rasterSegs = 32u;
    const unsigned segWidth = (depthTex.w + rasterSegs - 1u) / rasterSegs;
    rasterSegs = (depthTex.w + segWidth - 1u) / segWidth;
    const int NV = sh.numVaryings;
    const int recStride = 3 + 3 * NV;
    const unsigned numBlocks = B.numBlocks;
    const unsigned long long totalPrims = B.totalPrims;

    /* the draw table and the block -> draw map of the batch */
    CK(c->drawTable.ensure(B.descs.size() * sizeof(SLDrawDesc)));
    CK(c->blockDraw.ensure((size_t)numBlocks * sizeof(uint32_t)));
    CK(cudaMemcpyAsync(c->drawTable.buf.ptr, B.descs.data(), B.descs.size() * sizeof(SLDrawDesc), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->blockDraw.buf.ptr, B.blockDraw.data(), (size_t)numBlocks * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));

    /* record slots: SL_VB_THREADS per (block, round) at fixed places + an overflow area for rounds that clipping fans out */
    const unsigned long long recFixed = (unsigned long long)numBlocks * SL_VB_PRIMS;
    unsigned long long recOverflow = totalPrims / 8 + 4096;
    if (c->caps.recOverflow > recOverflow) recOverflow = c->caps.recOverflow;
    unsigned long long recCapacity = recFixed + recOverflow;
    /* post-transform cache: batches whose index buffers address few enough vertices for one block's shared memory (small
     * instanced meshes).  SLCU_VERTEX_CACHE=0 turns it off. */
    static const bool allowVertexCache = !(getenv("SLCU_VERTEX_CACHE") && atoi(getenv("SLCU_VERTEX_CACHE")) == 0);
    const uint8_t* rowNeeded = nullptr;
    unsigned numSubLists = (segWidth + (1u << SL_SHADE_SHIFT) - 1u) >> SL_SHADE_SHIFT;
    if (numSubLists > (unsigned)SL_MAX_SUBLISTS) { c->err = "segment too wide for the sub-list table"; return SLCU_ERR_ARG; }
    const unsigned long long numLists = (unsigned long long)H * rasterSegs * numSubLists;
    unsigned long long chunkCap = 0, handoffSlots = 0;
    /* Shared vertex stage (k_vertex): every rank runs 1/world of the blocks and routes the records.  The decision is the same
     * on every rank (same calls, same arena size): the first attempt of a triangle batch that fits the arena.  A refusal
     * of any kind makes the rank run the batch again by itself, all blocks, into its own buffers. */
    static const bool allowGeomShare = !(getenv("SLCU_GEOM_SHARE") && atoi(getenv("SLCU_GEOM_SHARE")) == 0);
    const unsigned world = c->partWorld;
    bool sharded = false;
    size_t arenaRecDraw = 0, arenaBlockInfo = 0, arenaRecMask = 0;
    /* sharing pays when the vertex stage is long against the ranks' skew at the rendezvous (they start a frame ~0.1 ms
     * apart: measured at 8 GPUs, profiles/r02y_share_ab_x8.txt: C2, 10.2 M primitives, 0.94 -> 0.68 ms; C3, 2.1 M, 0.86 -> 0.91)
     * and there are enough ranks to spread the routed records over (from 4 on) */
    const unsigned long long geomShareMinPrims = (getenv("SLCU_GEOM_SHARE_MIN_PRIMS") && *getenv("SLCU_GEOM_SHARE_MIN_PRIMS")) ? strtoull(getenv("SLCU_GEOM_SHARE_MIN_PRIMS"), nullptr, 10) : 4000000ull;
    const unsigned geomShareMinWorld = (getenv("SLCU_GEOM_SHARE_MIN_PRIMS") && *getenv("SLCU_GEOM_SHARE_MIN_PRIMS")) ? 2u : 4u; /* two ranks route half the records over one link pair: C2 1.17 ms shared, 1.11 replicated */
    if (allowGeomShare && !B.wire && c->geom.ready(world) && totalPrims >= geomShareMinPrims && world >= geomShareMinWorld)
    {
        const unsigned long long slice = (totalPrims / 8 + 4096) / world + 1024;
        const unsigned long long cap = recFixed + slice * world;
        auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
        arenaRecDraw = up((size_t)cap * recStride * sizeof(float4));
        arenaBlockInfo = arenaRecDraw + up((size_t)cap * sizeof(uint16_t));
        arenaRecMask = arenaBlockInfo + up((size_t)numBlocks * sizeof(SLBlockInfo));
        const size_t need = arenaRecMask + up((size_t)cap);
        if (cap <= 0x7ffffff0ull && need <= c->geom.arenaBytes)
        {
            if (!B.geomLocal)
            {
                sharded = true;
                recOverflow = slice;
                recCapacity = cap;
            }
        }
    }
    float4* recordsPtr = nullptr;
    uint16_t* recDrawPtr = nullptr;
    SLBlockInfo* blockInfoPtr = nullptr;
    for (int attempt = 0;; ++attempt)
    {
        if (recCapacity > 0x7ffffff0ull) { c->err = "batch too large"; return SLCU_ERR_UNSUPPORTED; }
        CK(cudaMemsetAsync(c->counters, 0, SL_COUNTERS_BATCH_BYTES, c->stream)); /* the per-batch part */
        {
            SLVertexParams vp;
            memset(&vp, 0, sizeof(vp));
            vp.draws = reinterpret_cast<const SLDrawDesc*>(c->drawTable.buf.ptr);
            vp.blockDraw = reinterpret_cast<const uint32_t*>(c->blockDraw.buf.ptr);
            vp.cullMode = sh.cullMode;
            vp.numVaryings = NV;
            /* the cache holds every vertex the batch's index buffers can name */
            const bool cache = allowVertexCache && B.maxIndex < (unsigned long long)SL_VCACHE_VERTS;
            unsigned long long cv = cache ? ((B.maxIndex + 32ull) & ~31ull) : 32ull;
            vp.cacheVerts = (unsigned)cv;
            vp.cacheBytes = cache ? (unsigned)(3u * SL_VB_PRIMS * sizeof(uint16_t) + (2 + NV) * cv * sizeof(float4) + cv / 8 + cv) : 0u;
            vp.recOverflowCapacity = (unsigned)recOverflow;
            vp.counters = c->counters;
            vp.partStrip = (int)c->partStripRows;
            unsigned grid = numBlocks;
            if (sharded)
            {
                const uint32_t epoch = ++c->geom.epoch;
                const size_t arenaOff = SL_GEOM_HEAD_BYTES + (size_t)(epoch & 1u) * c->geom.arenaBytes;
                vp.shardRank = (int)c->partRank;
                vp.shardWorld = (int)world;
                vp.recOverflowBase = (unsigned)(recFixed + (unsigned long long)c->partRank * recOverflow);
                SLGeomRoute gr;
                memset(&gr, 0, sizeof(gr));
                for (unsigned t = 0; t < world; ++t)
                {
                    unsigned char* a = c->geom.peer[t] + arenaOff;
                    gr.out[t].records = reinterpret_cast<float4*>(a);
                    gr.out[t].recDraw = reinterpret_cast<uint16_t*>(a + arenaRecDraw);
                    gr.out[t].blockInfo = reinterpret_cast<SLBlockInfo*>(a + arenaBlockInfo);
                }
                vp.out = gr.out[c->partRank];
                vp.recMask = c->geom.peer[c->partRank] + arenaOff + arenaRecMask;
                recordsPtr = vp.out.records;
                recDrawPtr = vp.out.recDraw;
                blockInfoPtr = vp.out.blockInfo;
                grid = numBlocks > c->partRank ? (numBlocks - c->partRank + world - 1) / world : 0u;
                if (grid)
                {
                    { KernelScope ks(c, SLCU_K_VERTEX); CK(launch_vertex_stage(sh, grid, c->stream, vp, cache)); }
                    gr.blockDraw = vp.blockDraw;
                    gr.recMask = vp.recMask;
                    gr.rank = (int)c->partRank;
                    gr.world = (int)world;
                    gr.recStride = recStride;
                    KernelScope ks(c, SLCU_K_ROUTE);
                    k_geom_route<<<grid, 32 * SL_VB_SEGS, 0, c->stream>>>(gr);
                    CK(cudaGetLastError());
                }
                SLGeomSync gs;
                memset(&gs, 0, sizeof(gs));
                for (unsigned t = 0; t < world; ++t) gs.head[t] = c->geom.peer[t];
                gs.rank = (int)c->partRank;
                gs.world = (int)world;
                gs.epoch = epoch;
                gs.counters = c->counters;
                gs.timeoutClocks = 20000000000ll; /* ~10 s: a rank that never arrives is an error, not a hang */
                { KernelScope ks(c, SLCU_K_RENDEZVOUS); k_geom_barrier<<<1, 32, 0, c->stream>>>(gs); }
                CK(cudaGetLastError());
            }
            else
            {
                CK(c->records.ensure((size_t)recCapacity * recStride * sizeof(float4)));
                CK(c->recDraw.ensure((size_t)recCapacity * sizeof(uint16_t)));
                CK(c->blockInfo.ensure((size_t)numBlocks * sizeof(SLBlockInfo)));
                recordsPtr = reinterpret_cast<float4*>(c->records.buf.ptr);
                recDrawPtr = reinterpret_cast<uint16_t*>(c->recDraw.buf.ptr);
                blockInfoPtr = reinterpret_cast<SLBlockInfo*>(c->blockInfo.buf.ptr);
                vp.shardRank = 0;
                vp.shardWorld = 1;
                vp.recOverflowBase = (unsigned)recFixed;
                vp.out.records = recordsPtr;
                vp.out.recDraw = recDrawPtr;
                vp.out.blockInfo = blockInfoPtr;
                vp.recMask = nullptr;
                { KernelScope ks(c, SLCU_K_VERTEX); CK(launch_vertex_stage(sh, grid, c->stream, vp, cache)); }
            }
        }

        if (B.wire)
        {
            /* ---- wireframe: edge fragments are counted, emitted with an order key, sorted and shaded (sl_wire.cuh) ---- */
            SLWireParams wp;
            memset(&wp, 0, sizeof(wp));
            wp.records = recordsPtr;
            wp.blockInfo = blockInfoPtr;
            wp.numBlocks = numBlocks;
            wp.recStride = recStride;
            wp.width = depthTex.w;
            wp.height = H;
            wp.rank = c->partRank;
            wp.world = c->partWorld;
            wp.stripRows = c->partStripRows;
            wp.counters = c->counters;
            const unsigned wireGrid = numBlocks * SL_WIRE_BLOCKS_PER_VB; /* one warp per record slot */
            CK(c->wireRecInfo.ensure((size_t)recCapacity * 2 * sizeof(uint32_t)));
            wp.recBase = reinterpret_cast<uint32_t*>(c->wireRecInfo.buf.ptr);
            wp.recCount = wp.recBase + recCapacity;
            { KernelScope ks(c, SLCU_K_SPAN_COUNT); k_wire<0><<<wireGrid, 128, 0, c->stream>>>(wp); }
            { KernelScope ks(c, SLCU_K_SCAN); k_wire_offsets<<<1, 1024, 0, c->stream>>>(wp.blockInfo, numBlocks, wp.recCount, wp.recBase); }
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(c->hostCounters, c->counters, sizeof(SLCounters), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            if (c->hostCounters->abortBatch & SL_ABORT_RECORDS)
            {
                if (attempt == 0) { recOverflow = totalPrims * 7 + 4096; recCapacity = recFixed + recOverflow; c->caps.recOverflow = recOverflow; continue; }
                c->err = "triangle record buffer overflow";
                return SLCU_ERR_OOM;
            }
            const unsigned long long numFrags = c->hostCounters->pixelsCovered;
            if (!numFrags) { k_commit<<<1, 1, 0, c->stream>>>(c->counters); return 0; }
            if (numFrags > 0x7ffffff0ull) { c->err = "too many wireframe fragments for one batch"; return SLCU_ERR_UNSUPPORTED; }
            CK(c->pointKeys.ensure((size_t)numFrags * 2 * sizeof(unsigned long long)));
            CK(c->wireSlots.ensure((size_t)numFrags * 2 * sizeof(unsigned long long)));
            const unsigned long long bigRecords = c->hostCounters->rowEndCuts; /* records with 600+ candidates: their queue flushes are observable */
            CK(c->pointRecs.ensure((size_t)numFrags * (sizeof(float4) + 3 * sizeof(uint32_t) + sizeof(uint16_t) + 2)));
            unsigned long long* keysA = reinterpret_cast<unsigned long long*>(c->pointKeys.buf.ptr);
            unsigned long long* valsA = reinterpret_cast<unsigned long long*>(c->wireSlots.buf.ptr);
            int ordBits = 13; /* segment (5) + index (8) + the bits of the vertex block */
            while ((1u << (ordBits - 13)) < numBlocks) ++ordBits;
            wp.lowBits = ordBits + SL_WIRE_SUB_BITS;
            wp.capacity = numFrags;
            wp.keys = keysA;
            wp.vals = valsA;
            wp.fragBc = reinterpret_cast<float4*>(c->pointRecs.buf.ptr);
            wp.fragXY = reinterpret_cast<uint32_t*>(wp.fragBc + numFrags);
            wp.fragRec = wp.fragXY + numFrags;
            uint16_t* fragEpoch = reinterpret_cast<uint16_t*>(wp.fragRec + numFrags);
            uint8_t* fragPass = reinterpret_cast<uint8_t*>(fragEpoch + numFrags);
            CK(cudaMemsetAsync(&c->counters->pixelsCovered, 0, sizeof(unsigned long long), c->stream)); /* now the slot allocator */
            { KernelScope ks(c, SLCU_K_SPAN_FILL); k_wire<1><<<wireGrid, 128, 0, c->stream>>>(wp); }
            CK(cudaGetLastError());

            unsigned pixBits = 1;
            while ((1ull << pixBits) < (unsigned long long)depthTex.w * H) ++pixBits;
            /* slots are in submission order already: only the pixel bits take part */
            const int sortedHalf = sort_u64(c, keysA, valsA, numFrags, wp.lowBits, wp.lowBits + (int)pixBits);
            if (sortedHalf < 0) return sortedHalf;

            SLWireShadeParams ws;
            ws.draws = reinterpret_cast<const SLDrawDesc*>(c->drawTable.buf.ptr);
            ws.recDraw = recDrawPtr;
            ws.records = wp.records;
            ws.recStride = recStride;
            ws.lowBits = wp.lowBits;
            ws.keys = keysA + (size_t)sortedHalf * numFrags;
            ws.vals = valsA + (size_t)sortedHalf * numFrags;
            ws.fragBc = wp.fragBc;
            ws.fragXY = wp.fragXY;
            ws.fragRec = wp.fragRec;
            ws.fragEpoch = nullptr;
            ws.fragPass = fragPass;
            CK(c->wireLeaders.ensure((size_t)numFrags * sizeof(uint32_t)));
            ws.leaders = reinterpret_cast<uint32_t*>(c->wireLeaders.buf.ptr);
            ws.numLeaders = &c->counters->chunkAlloc; /* a free per-batch word */
            ws.numFrags = numFrags;
            ws.numVaryings = NV;
            ws.numOutputs = sh.numOutputs;
            ws.blend = sh.blend;
            ws.depthTest = sh.depthTest;
            ws.depthMask = sh.depthMask;
            ws.depth = depthTex.dev;
            ws.depthBpt = depthTex.bpt;
            for (int k = 0; k < 4; ++k)
            {
                ws.color[k] = nullptr;
                ws.colorType[k] = 0;
                if (k < sh.numOutputs) { const HostTexture& t = c->textures[fbo.colors[k]]; ws.color[k] = t.dev; ws.colorType[k] = t.type; }
            }
            ws.counters = c->counters;
            CK(cudaMemsetAsync(ws.numLeaders, 0, sizeof(unsigned), c->stream));
            {
                SLLeaderParams lp;
                lp.keys = ws.keys;
                lp.n = numFrags;
                lp.shift = ws.lowBits;
                lp.leaders = ws.leaders;
                lp.numLeaders = ws.numLeaders;
                KernelScope ks(c, SLCU_K_SCAN);
                k_pixel_leaders<<<(unsigned)((numFrags + 255ull) / 256ull), 256, 0, c->stream>>>(lp);
            }
            CK(cudaGetLastError());
            /* pixels touched <= min(candidates, texels): the walkers exit past the leader count */
            const unsigned long long maxLeaders = numFrags < (unsigned long long)depthTex.w * H ? numFrags : (unsigned long long)depthTex.w * H;
            const unsigned shadeGrid = (unsigned)((maxLeaders + 127ull) / 128ull);
            if (bigRecords && sh.depthTest != 0 && sh.depthMask)
            {
                /* flush epochs of the big records: iterate depth outcomes <-> epochs to the fixed point (sl_wire.cuh) */
                ws.fragEpoch = fragEpoch;
                CK(cudaMemsetAsync(fragEpoch, 0, (size_t)numFrags * sizeof(uint16_t), c->stream));
                SLWireEpochParams ep;
                ep.blockInfo = wp.blockInfo;
                ep.numBlocks = numBlocks;
                ep.recBase = wp.recBase;
                ep.recCount = wp.recCount;
                ep.fragPass = fragPass;
                ep.fragEpoch = fragEpoch;
                ep.changed = &c->counters->rasterTicket; /* a free per-batch word */
                bool settled = false;
                for (int it = 0; it < 16 && !settled; ++it)
                {
                    CK(cudaMemsetAsync(ep.changed, 0, sizeof(unsigned), c->stream));
                    { KernelScope ks(c, SLCU_K_RASTER); k_wire_depth<<<shadeGrid, 128, 0, c->stream>>>(ws); }
                    { KernelScope ks(c, SLCU_K_SCAN); k_wire_epoch<<<wireGrid, 128, 0, c->stream>>>(ep); }
                    CK(cudaGetLastError());
                    unsigned changed = 0;
                    CK(cudaMemcpyAsync(&c->hostCounters[1], ep.changed, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
                    CK(cudaStreamSynchronize(c->stream));
                    memcpy(&changed, &c->hostCounters[1], sizeof(unsigned));
                    settled = changed == 0;
                }
                if (!settled) { CK(cudaMemsetAsync(ep.changed, 0, sizeof(unsigned), c->stream)); c->err = "wireframe flush epochs did not settle"; return SLCU_ERR_UNSUPPORTED; }
            }
            { KernelScope ks(c, SLCU_K_SHADE); CK(launch_wire_shade_stage(sh, shadeGrid, c->stream, ws)); }
            k_commit<<<1, 1, 0, c->stream>>>(c->counters);
            CK(cudaGetLastError());
            return 0;
        }

        /* ---- span count + scans ---- */
        CK(c->cnt.ensure((size_t)numBlocks * rowPitch * sizeof(uint16_t)));
        CK(c->off.ensure((size_t)numBlocks * rowPitch * sizeof(uint32_t)));
        const unsigned numSegs = numBlocks < (unsigned)SL_SCAN_SEGS ? numBlocks : (unsigned)SL_SCAN_SEGS;
        const unsigned blocksPerSeg = (numBlocks + numSegs - 1) / numSegs;
        const unsigned usedSegs = (numBlocks + blocksPerSeg - 1) / blocksPerSeg;
        CK(c->segSums.ensure((size_t)usedSegs * rowPitch * sizeof(uint32_t)));
        CK(c->rowTotals.ensure((size_t)rowPitch * sizeof(uint32_t)));
        CK(c->rowStart.ensure((size_t)(H + 1) * sizeof(uint32_t)));

        SLSpanParams sp;
        sp.records = recordsPtr;
        sp.recPresent = sharded ? recDrawPtr : nullptr;
        sp.blockInfo = blockInfoPtr;
        sp.firstBlock = 0;
        sp.numRows = H;
        sp.rowPitch = rowPitch;
        sp.width = depthTex.w;
        sp.segWidth = segWidth;
        sp.numSegs = rasterSegs;
        sp.chained = sh.depthTest != 0 ? 1 : 0;
        sp.recStride = recStride;
        sp.cnt = reinterpret_cast<uint16_t*>(c->cnt.buf.ptr);
        sp.off = reinterpret_cast<const uint32_t*>(c->off.buf.ptr);
        sp.rowStart = reinterpret_cast<const uint32_t*>(c->rowStart.buf.ptr);
        sp.entries = nullptr;
        sp.counters = c->counters;
        sp.rowFlags = nullptr;
        sp.rowNeeded = nullptr;
        sp.partFast = 0;
        sp.partRank = (int)c->partRank;
        sp.partWorld = (int)c->partWorld;
        sp.partStrip = (int)c->partStripRows;
        const bool partitioned = c->partWorld > 1;
        const bool partFast = partitioned && !B.partFull;
        if (partitioned)
        {
            CK(c->rowFlags.ensure(H + 1));
            CK(c->rowNeeded.ensure(H + 1));
            CK(cudaMemsetAsync(c->rowFlags.buf.ptr, 0, H + 1, c->stream));
            sp.rowFlags = c->rowFlags.buf.ptr;
            if (partFast)
            {
                /* fast path: the rows this rank owns and nothing else, in the count pass already (see SLSpanParams::partFast) */
                KernelScope ks(c, SLCU_K_SCAN);
                k_row_closure<<<(H + 255) / 256, 256, 0, c->stream>>>(c->rowFlags.buf.ptr, c->rowNeeded.buf.ptr, H, c->partRank, c->partWorld, c->partStripRows);
                rowNeeded = c->rowNeeded.buf.ptr;
                sp.rowNeeded = rowNeeded;
                sp.rowFlags = nullptr;
                sp.partFast = 1;
            }
        }
        /* work items of the span stage: (vertex block, 256-row band) pairs, taken through a ticket by persistent CTAs */
        const unsigned numBands = (H + SL_SPAN_WIN - 1) / SL_SPAN_WIN;
        CK(c->spanItems.ensure((size_t)numBlocks * numBands * sizeof(uint32_t)));
        CK(cudaMemsetAsync(c->cnt.buf.ptr, 0, (size_t)numBlocks * rowPitch * sizeof(uint16_t), c->stream)); /* k_span<0> writes the rows of its items only */
        sp.items = reinterpret_cast<const uint32_t*>(c->spanItems.buf.ptr);
        sp.numItems = &c->counters->spanItems;
        CK(c->blockRows.ensure((size_t)numBlocks * sizeof(uint32_t)));
        { KernelScope ks(c, SLCU_K_SCAN); k_block_rows<<<(numBlocks + 7u) / 8u, 256, 0, c->stream>>>(sp.blockInfo, numBlocks, reinterpret_cast<uint32_t*>(c->blockRows.buf.ptr)); }
        { KernelScope ks(c, SLCU_K_SCAN); k_span_items<<<1, 1024, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(c->blockRows.buf.ptr), numBlocks, H, reinterpret_cast<uint32_t*>(c->spanItems.buf.ptr), &c->counters->spanItems); }
        CK(cudaGetLastError());
        int spanSMs = 148;
        cudaDeviceGetAttribute(&spanSMs, cudaDevAttrMultiProcessorCount, c->device);
        unsigned spanGrid = (unsigned)spanSMs * 6u;
        if (spanGrid > numBlocks * numBands) spanGrid = numBlocks * numBands;
        sp.ticket = &c->counters->spanTicket[0];
        { KernelScope ks(c, SLCU_K_SPAN_COUNT); k_span<0><<<spanGrid, 256, 0, c->stream>>>(sp); }
        CK(cudaGetLastError());
        if (partitioned && !partFast)
        {
            KernelScope ks(c, SLCU_K_SCAN);
            k_row_closure<<<(H + 255) / 256, 256, 0, c->stream>>>(c->rowFlags.buf.ptr, c->rowNeeded.buf.ptr, H, c->partRank, c->partWorld, c->partStripRows);
            rowNeeded = c->rowNeeded.buf.ptr;
            sp.rowNeeded = rowNeeded;
        }
        const dim3 sgrid((rowPitch + 127) / 128, usedSegs);
        { KernelScope ks(c, SLCU_K_SCAN); k_scan_seg_sums<<<sgrid, 128, 0, c->stream>>>(sp.cnt, reinterpret_cast<uint32_t*>(c->segSums.buf.ptr), numBlocks, rowPitch, blocksPerSeg, rowNeeded, H); }
        CK(cudaGetLastError());
        { KernelScope ks(c, SLCU_K_SCAN); k_scan_rows<<<(rowPitch + 127) / 128, 128, 0, c->stream>>>(reinterpret_cast<uint32_t*>(c->segSums.buf.ptr), reinterpret_cast<uint32_t*>(c->rowTotals.buf.ptr), rowPitch, usedSegs); }
        CK(cudaGetLastError());
        /* what this batch gets: the sizes its predecessors needed (first batch: a guess; the device refuses the batch when
         * it does not fit and the host then knows the exact figures) */
        unsigned long long capEntries = c->caps.entries ? c->caps.entries : totalPrims / 2 + 65536ull;
        CK(c->entries.ensure((size_t)capEntries * sizeof(SLSpanEntry)));
        capEntries = c->entries.buf.bytes / sizeof(SLSpanEntry);
        if (capEntries > 0xfffffff0ull) capEntries = 0xfffffff0ull;
        const unsigned long long chunkSlack = 64ull + 160ull * SL_R2_MIN_BLOCKS * SL_R2_WARPS * (SL_R2_FIFO_BLOCK * 3ull); /* what the resolving warps hold in their FIFOs */
        chunkCap = (c->caps.chunks ? c->caps.chunks : (unsigned long long)depthTex.w * H / 8ull) + numLists + chunkSlack;
        if (chunkCap > 0xfffffff0ull || chunkCap * SL_CHUNK_BYTES > (96ull << 30)) { c->err = "fragment list pool too large"; return SLCU_ERR_OOM; }
        CK(c->fragPool.ensure((size_t)chunkCap * SL_CHUNK_BYTES));
        chunkCap = c->fragPool.buf.bytes / SL_CHUNK_BYTES;
        if (chunkCap > 0xfffffff0ull) chunkCap = 0xfffffff0ull;
        CK(c->chunkNext.ensure((size_t)chunkCap * 2 * sizeof(uint32_t))); /* next and previous chunk */
        if (sh.depthTest != 0)
        {
            handoffSlots = 4096;
            while (handoffSlots < c->caps.handoffSlots) handoffSlots <<= 1;
            CK(c->handoff.ensure((size_t)handoffSlots * sizeof(unsigned long long)));
            CK(c->handoffRegs.ensure((size_t)handoffSlots * 4 * sizeof(float4)));
            CK(cudaMemsetAsync(c->handoff.buf.ptr, 0, (size_t)handoffSlots * sizeof(unsigned long long), c->stream));
        }
        SLBatchCaps caps;
        caps.entries = capEntries;
        caps.chunks = chunkCap;
        caps.handoffSlots = handoffSlots;
        caps.numLists = numLists;
        caps.chunkSlack = chunkSlack;
        caps.chained = sh.depthTest != 0 ? 1 : 0;
        { KernelScope ks(c, SLCU_K_SCAN); k_scan_row_starts<<<1, 1024, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(c->rowTotals.buf.ptr), reinterpret_cast<uint32_t*>(c->rowStart.buf.ptr), H, c->counters, caps); }
        CK(cudaGetLastError());
        { KernelScope ks(c, SLCU_K_SCAN); k_scan_offsets<<<sgrid, 128, 0, c->stream>>>(sp.cnt, reinterpret_cast<const uint32_t*>(c->segSums.buf.ptr), reinterpret_cast<uint32_t*>(c->off.buf.ptr), numBlocks, rowPitch, blocksPerSeg, rowNeeded, H); }
        CK(cudaGetLastError());

        sp.entries = reinterpret_cast<SLSpanEntry*>(c->entries.buf.ptr);
        sp.ticket = &c->counters->spanTicket[1];
        { KernelScope ks(c, SLCU_K_SPAN_FILL); k_span<1><<<spanGrid, 256, 0, c->stream>>>(sp); }
        CK(cudaGetLastError());
        break;
    }

    /* ---- fragment stage: k_resolve (depth, coverage, order) -> fragment lists -> k_shade<FS> ---- */
    CK(c->listInfo.ensure((size_t)numLists * 3 * sizeof(uint32_t))); /* head, count, tail */
    CK(cudaMemsetAsync(c->listInfo.buf.ptr, 0, (size_t)numLists * 2 * sizeof(uint32_t), c->stream));
    SLFragLists lists;
    lists.pool = c->fragPool.buf.ptr;
    lists.chunkNext = reinterpret_cast<uint32_t*>(c->chunkNext.buf.ptr);
    lists.chunkPrev = lists.chunkNext + chunkCap;
    lists.listHead = reinterpret_cast<uint32_t*>(c->listInfo.buf.ptr);
    lists.listCount = lists.listHead + numLists;
    lists.listTail = lists.listCount + numLists;
    lists.chunkCapacity = (unsigned)chunkCap;

    SLRasterParams rp;
    rp.records = recordsPtr;
    rp.entries = reinterpret_cast<const SLSpanEntry*>(c->entries.buf.ptr);
    rp.rowStart = reinterpret_cast<const uint32_t*>(c->rowStart.buf.ptr);
    rp.recStride = recStride;
    rp.depthTest = sh.depthTest;
    rp.depthMask = sh.depthMask;
    rp.numRows = H;
    rp.width = depthTex.w;
    rp.depth = depthTex.dev;
    rp.depthBpt = depthTex.bpt;
    rp.handoff = nullptr;
    rp.handoffRegs = nullptr;
    rp.handoffMask = 0;
    rp.rowNeeded = rowNeeded;
    rp.segWidth = segWidth;
    rp.numSegs = rasterSegs;
    rp.numSubLists = numSubLists;
    rp.lists = lists;
    if (handoffSlots)
    {
        rp.handoff = reinterpret_cast<unsigned long long*>(c->handoff.buf.ptr);
        rp.handoffRegs = reinterpret_cast<float4*>(c->handoffRegs.buf.ptr);
        rp.handoffMask = (unsigned)(handoffSlots - 1);
    }
    /* each resolving warp keeps its segment's depth texels in shared memory (at most 32 x ... pixels x 8 bytes) */
    rp.tileBytes = (segWidth * depthTex.bpt + 15u) & ~15u;
    rp.counters = c->counters;
    rp.ticketOrder = nullptr;
    rp.rowPerm = nullptr;
    {
        /* fine segments shorten the longest ticket and multiply the per-layer overhead of the rows they cut: a loss on one
         * GPU (k_resolve2 is throughput-bound there: C3 0.84 -> 1.04 ms), a gain under the sort-first partition, where a
         * rank's few rows leave the kernel as long as its longest ticket (profiles/r02n_*, r02o_*) */
        static const int fineEnv = getenv("SLCU_FINE_ROWS") ? atoi(getenv("SLCU_FINE_ROWS")) : -1;
        rp.fineRows = fineEnv >= 0 ? (unsigned)fineEnv : (c->partWorld > 1 ? 3u : 0u);
    }
    {
        /* skewed work order (see k_resolve): only when every hand-over goes from a segment to its right neighbour */
        static const int skewEnv = getenv("SLCU_TICKET_SKEW") ? atoi(getenv("SLCU_TICKET_SKEW")) : -1;
        int numSMs = 148;
        cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, c->device);
        const unsigned resident = (unsigned)numSMs * (unsigned)SL_RASTER_MIN_BLOCKS * SL_RASTER_WARPS; /* warps of k_resolve that fit the machine */
        /* sort-first partition: only 1/world of the rows does any work on this rank, so the skew is world times longer */
        const unsigned world = c->partWorld > 1 ? c->partWorld : 1u;
        const unsigned skew = skewEnv >= 0 ? (unsigned)skewEnv : (resident * world + rasterSegs - 1) / rasterSegs + 32u;
        if (sh.depthTest != 0 && rasterSegs > 1 && rasterSegs <= 64 && skew > 0) /* used by the kernel only when no span runs past a row end */
        {
            if (c->orderRows != H || c->orderSegs != rasterSegs || c->orderSkew != skew)
            {
                std::vector<uint32_t> order;
                order.reserve((size_t)H * rasterSegs);
                /* pairs (y, s) sorted by key y + s*skew, then by s: walk the keys */
                const unsigned long long maxKey = (unsigned long long)H - 1ull + (unsigned long long)(rasterSegs - 1) * skew;
                for (unsigned long long key = 0; key <= maxKey; ++key)
                {
                    for (unsigned s2 = 0; s2 < rasterSegs; ++s2)
                    {
                        const unsigned long long off = (unsigned long long)s2 * skew;
                        if (key < off) break;
                        const unsigned long long yy = key - off;
                        if (yy < H) order.push_back((uint32_t)(yy << 6) | s2);
                    }
                }
                CK(c->ticketOrder.ensure(order.size() * sizeof(uint32_t)));
                CK(cudaMemcpyAsync(c->ticketOrder.buf.ptr, order.data(), order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
                CK(cudaStreamSynchronize(c->stream)); /* `order` is a temporary */
                c->orderRows = H; c->orderSegs = rasterSegs; c->orderSkew = skew;
            }
            rp.ticketOrder = reinterpret_cast<const uint32_t*>(c->ticketOrder.buf.ptr);
            static const bool heavyFirst = !(getenv("SLCU_HEAVY_FIRST") && atoi(getenv("SLCU_HEAVY_FIRST")) == 0);
            if (heavyFirst)
            {
                CK(c->rowPerm.ensure((size_t)(H + 1) * sizeof(uint32_t)));
                KernelScope ks(c, SLCU_K_SCAN);
                k_row_order<<<1, 1024, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(c->rowStart.buf.ptr), H, reinterpret_cast<uint32_t*>(c->rowPerm.buf.ptr), rp.fineRows ? rp.fineRows : 2u);
                rp.rowPerm = reinterpret_cast<const uint32_t*>(c->rowPerm.buf.ptr);
            }
        }
    }
    rp.diagnostics = getenv("SLCU_DEBUG") ? 1 : 0;
    /* Two implementations of the stage (same inputs, same outputs): k_resolve2 walks every span piece with one lane
     * (sl_resolve2.cuh: short spans, i.e. every triangle-dense scene), k_resolve spreads the groups of a piece over the
     * lanes of a window (long spans).  SLCU_RESOLVE=1|2 forces one of them. */
    static const int resolveEnv = getenv("SLCU_RESOLVE") ? atoi(getenv("SLCU_RESOLVE")) : 0;
    const bool useResolve2 = resolveEnv == 2 || (resolveEnv != 1 && c->caps.avgSpan < 48ull); /* average span of the last verified batch */
    if (useResolve2)
    {
        const unsigned smemBytes = SL_R2_RCP_BYTES + SL_R2_WARPS * sl_r2_warp_bytes(segWidth, rp.tileBytes);
        /* one instantiation per depth texel size and depth function */
        typedef void (*ResolveKernel)(const SLRasterParams);
        static const ResolveKernel resolveKernels[3][7] = {
            {k_resolve2<2, 0>, k_resolve2<2, 1>, k_resolve2<2, 2>, k_resolve2<2, 3>, k_resolve2<2, 4>, k_resolve2<2, 5>, k_resolve2<2, 6>},
            {k_resolve2<4, 0>, k_resolve2<4, 1>, k_resolve2<4, 2>, k_resolve2<4, 3>, k_resolve2<4, 4>, k_resolve2<4, 5>, k_resolve2<4, 6>},
            {k_resolve2<8, 0>, k_resolve2<8, 1>, k_resolve2<8, 2>, k_resolve2<8, 3>, k_resolve2<8, 4>, k_resolve2<8, 5>, k_resolve2<8, 6>}};
        const ResolveKernel kernel = resolveKernels[depthTex.bpt == 2 ? 0 : (depthTex.bpt == 4 ? 1 : 2)][sh.depthTest];
        static bool attrSet2[64][3][7] = {};
        bool& attrDone = attrSet2[c->device & 63][depthTex.bpt == 2 ? 0 : (depthTex.bpt == 4 ? 1 : 2)][sh.depthTest];
        if (!attrDone)
        {
            CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            attrDone = true;
        }
        if (smemBytes > 160u * 1024u) { c->err = "framebuffer too wide for the resolve tile"; return SLCU_ERR_ARG; }
        int numSMs = 148;
        cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, c->device);
        const unsigned long long numWarps = (unsigned long long)H * rasterSegs;
        unsigned long long grid = (unsigned long long)numSMs * SL_R2_MIN_BLOCKS; /* persistent warps */
        const unsigned long long needBlocks = (numWarps + SL_R2_WARPS - 1) / SL_R2_WARPS;
        if (grid > needBlocks) grid = needBlocks;
        KernelScope ks(c, SLCU_K_RASTER);
        kernel<<<(unsigned)grid, SL_R2_WARPS * 32, smemBytes, c->stream>>>(rp);
        CK(cudaGetLastError());
    }
    else
    {
    const unsigned warpsPerBlock = SL_RASTER_WARPS;
    const unsigned smemBytes = warpsPerBlock * (SL_PIECEQ_BYTES + SL_SUBLIST_BYTES + SL_PIECEREG_BYTES + SL_WINSCRATCH_BYTES + ((4u * segWidth + 15u) & ~15u) + rp.tileBytes); /* piece queue, sub-list state, piece registers, window scratch, toucher masks, depth tile: per warp */
    {
        static bool attrSet[64] = {};
        if (!attrSet[c->device & 63])
        {
            CK(cudaFuncSetAttribute(k_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attrSet[c->device & 63] = true;
        }
        if (smemBytes > 96u * 1024u) { c->err = "framebuffer too wide for the resolve tile"; return SLCU_ERR_ARG; }
        KernelScope ks(c, SLCU_K_RASTER);
        const unsigned long long numWarps = (unsigned long long)H * rasterSegs;
        k_resolve<<<(unsigned)((numWarps + warpsPerBlock - 1) / warpsPerBlock), warpsPerBlock * 32, smemBytes, c->stream>>>(rp);
        CK(cudaGetLastError());
    }
    }

    SLShadeParams shp;
    shp.draws = reinterpret_cast<const SLDrawDesc*>(c->drawTable.buf.ptr);
    shp.recDraw = recDrawPtr;
    shp.records = recordsPtr;
    shp.recStride = recStride;
    shp.numVaryings = NV;
    shp.numOutputs = sh.numOutputs;
    shp.blend = sh.blend;
    shp.width = depthTex.w;
    shp.numLists = (unsigned)numLists;
    shp.segWidth = segWidth;
    shp.numSegs = rasterSegs;
    shp.numSubLists = numSubLists;
    {
        static const bool allowOpaqueSkip = !(getenv("SLCU_OPAQUE_SKIP") && atoi(getenv("SLCU_OPAQUE_SKIP")) == 0);
        shp.opaqueSkip = (allowOpaqueSkip && sh.blend == 0) ? 1 : 0;
    }
    shp.lists = lists;
    for (int i = 0; i < 4; ++i)
    {
        shp.color[i] = nullptr;
        shp.colorType[i] = 0;
        if (i < sh.numOutputs)
        {
            const HostTexture& t = c->textures[fbo.colors[i]];
            shp.color[i] = t.dev;
            shp.colorType[i] = t.type;
        }
    }
    shp.shadeTicket = &c->counters->shadeTicket;
    shp.listBegin = 0u;
    shp.abortBatch = &c->counters->abortBatch;
    {
        int numSMs = 148;
        cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, c->device);
        /* a read-back of one of the colour attachments waits for this batch: shade the lists in bands of rows (a list's
         * fragments only touch texels of its own row) and send every band's rows over the link while the next band is shaded */
        int readAtt = -1;
        if (c->pendingRead.active)
        {
            for (int i = 0; i < sh.numOutputs; ++i) if (fbo.colors[i] == c->pendingRead.tex) readAtt = i;
        }
        static const int bandsEnv = getenv("SLCU_READ_BANDS") ? atoi(getenv("SLCU_READ_BANDS")) : 4;
        const unsigned bands = (readAtt >= 0 && bandsEnv > 1 && H >= 64u) ? (unsigned)(bandsEnv > 8 ? 8 : bandsEnv) : 1u;
        const unsigned listsPerRow = rasterSegs * numSubLists;
        for (unsigned b = 0; b < bands; ++b)
        {
            const unsigned row0 = (unsigned)((unsigned long long)H * b / bands), row1 = (unsigned)((unsigned long long)H * (b + 1) / bands);
            shp.listBegin = row0 * listsPerRow;
            shp.numLists = row1 * listsPerRow;
            shp.shadeTicket = b == 0 ? &c->counters->shadeTicket : &c->counters->shadeBandTicket[b - 1];
            const unsigned long long bandLists = (unsigned long long)(row1 - row0) * listsPerRow;
            unsigned long long grid = (unsigned long long)numSMs * SL_SHADE_MIN_BLOCKS; /* persistent warps */
            const unsigned long long needBlocks = (bandLists + SL_SHADE_WARPS - 1) / SL_SHADE_WARPS;
            if (grid > needBlocks) grid = needBlocks;
            if (grid) { KernelScope ks(c, SLCU_K_SHADE); CK(launch_shade_stage(sh, (unsigned)grid, c->stream, shp)); }
            if (bands > 1)
            {
                const HostTexture& t = c->textures[c->pendingRead.tex];
                const size_t rowBytes = (size_t)t.w * t.bpt;
                size_t a = (size_t)row0 * rowBytes, e = (size_t)row1 * rowBytes;
                if (b + 1 == bands) e = c->pendingRead.bytes; /* whatever the caller asked for beyond the last row */
                if (e > c->pendingRead.bytes) e = c->pendingRead.bytes;
                if (!c->copyStream) CK(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
                if (!c->bandEvent[b]) CK(cudaEventCreateWithFlags(&c->bandEvent[b], cudaEventDisableTiming));
                if (!c->readDone) CK(cudaEventCreateWithFlags(&c->readDone, cudaEventDisableTiming));
                CK(cudaEventRecord(c->bandEvent[b], c->stream));
                CK(cudaStreamWaitEvent(c->copyStream, c->bandEvent[b], 0));
                if (a < e) CK(cudaMemcpyAsync(c->pendingRead.out + a, t.dev + a, e - a, cudaMemcpyDeviceToHost, c->copyStream));
                if (b + 1 == bands) { CK(cudaEventRecord(c->readDone, c->copyStream)); c->pendingRead.issued = true; }
            }
        }
    }
    k_commit<<<1, 1, 0, c->stream>>>(c->counters);
    CK(cudaGetLastError());
    verify = true;
    return 0;
}

/* ===========================================================================
 * run_clear_processors replacement (src/SL_Context.cpp:881-963)
 * =========================================================================*/
static void host_match_color(int type, const double rgba[4], unsigned char out[32])
{
    /* sl_match_color_for_type (src/SL_Color.cpp:135-190): color_cast<T,double> = (T)(c * max) element-wise */
    if (type_is_packed(type)) /* rgba_cast<packed, double>: (value_type)(c * (double)max(field)), truncated into the field */
    {
        static const int bits[5][4] = {{3, 3, 2, 0}, {5, 6, 5, 0}, {5, 5, 5, 1}, {4, 4, 4, 4}, {10, 10, 10, 2}};
        static const unsigned bytes[5] = {1, 2, 2, 2, 4};
        const int ci = type - 28;
        uint32_t v = 0;
        for (int j = 0, sh = 0; j < 4; sh += bits[ci][j], ++j)
        {
            const uint32_t fmax = (1u << bits[ci][j]) - 1u;
            const uint32_t f = ci == 4 ? (uint32_t)(uint16_t)(rgba[j] * (double)fmax) : (uint32_t)(uint8_t)(rgba[j] * (double)fmax);
            v |= (f & fmax) << sh;
        }
        memcpy(out, &v, bytes[ci]);
        return;
    }
    const int kind = type % 7;
    const unsigned n = (unsigned)(type / 7 + 1);
    for (unsigned i = 0; i < n; ++i)
    {
        switch (kind)
        {
            case 0: { const uint8_t v = (uint8_t)(rgba[i] * 255.0); out[i] = v; break; }
            case 1: { const uint16_t v = (uint16_t)(rgba[i] * 65535.0); memcpy(out + 2 * i, &v, 2); break; }
            case 2: { const uint32_t v = (uint32_t)(rgba[i] * 4294967295.0); memcpy(out + 4 * i, &v, 4); break; }
            case 3: { const uint64_t v = (uint64_t)(rgba[i] * 18446744073709551615.0); memcpy(out + 8 * i, &v, 8); break; }
            case 4: { const __half hv = __float2half_rn((float)rgba[i]); memcpy(out + 2 * i, &hv, 2); break; }
            case 5: { const float v = (float)rgba[i]; memcpy(out + 4 * i, &v, 4); break; }
            default: memcpy(out + 8 * i, &rgba[i], 8); break;
        }
    }
}

static bool host_depth_texel(const HostTexture& t, double depth, unsigned char out[8])
{
    switch (t.bpt)
    {
        case 2: { const __half hv = __float2half_rn((float)depth); memcpy(out, &hv, 2); return true; }
        case 4: { const float f = (float)depth; memcpy(out, &f, 4); return true; }
        case 8: memcpy(out, &depth, 8); return true;
        default: return false;
    }
}

static int clear_texture(slcu_context* c, const HostTexture& t, const unsigned char* texel)
{
    SLClearParams p;
    p.dst = t.dev;
    /* The reference ignores the depth dimension, and its one-thread slice is count + (1 - count % 3) clipped to count
     * (sl_calc_indexed_parition2<1>, SL_ShaderUtil.hpp:222-231; src/SL_ClearProcessor.cpp:23-33): the last texel is left
     * untouched when count % 3 == 2.  Parity is against the 1-thread reference, so the same texels are written here. */
    p.numTexels = (unsigned long long)t.w * t.h;
    if (p.numTexels % 3ull == 2ull) p.numTexels -= 1ull;
    p.bpt = t.bpt;
    memset(p.texel, 0, sizeof(p.texel));
    memcpy(p.texel, texel, t.bpt);
    const unsigned long long bytes = p.numTexels * t.bpt;
    unsigned grid = (unsigned)((bytes / 16 + 255) / 256);
    if (grid < 1) grid = 1;
    if (grid > 148 * 16) grid = 148 * 16;
    { KernelScope ks(c, SLCU_K_CLEAR); k_clear<<<grid, 256, 0, c->stream>>>(p); }
    CK(cudaGetLastError());
    return 0;
}

/* Device error flags (SL_ERR_*) seen by a read-back of the counters: turned into a return code and a decoded message,
 * then cleared on the device so that each occurrence is reported once. */
static std::vector<Scratch*> all_scratch(slcu_context* c)
{
    return {&c->records, &c->blockInfo, &c->cnt, &c->off, &c->segSums, &c->rowTotals, &c->rowStart, &c->entries, &c->handoff, &c->rowFlags, &c->rowNeeded,
            &c->drawTable, &c->blockDraw, &c->recDraw, &c->fragPool, &c->chunkNext, &c->listInfo, &c->handoffRegs, &c->ticketOrder, &c->spanItems,
            &c->blockRows, &c->rowPerm, &c->pointKeys, &c->pointRecs, &c->pointSortTemp, &c->wireSlots, &c->wireRecInfo, &c->wireLeaders};
}

static int report_device_flags(slcu_context* c, unsigned flags)
{
    if (!flags) return 0;
    std::string m = "device flags:";
    if (flags & SL_ERR_SPAN_TWO_ROWS) m += " [span longer than two rows: reference behaviour not reproduced]";
    if (flags & SL_ERR_SPAN_NEGATIVE_X) m += " [span starts left of x=0: reference behaviour not reproduced]";
    if (flags & SL_ERR_UNEXPECTED_OVERFLOW) m += " [row overflow of a triangle that cannot overflow]";
    if (flags & SL_ERR_HANDOFF_PUBLISH) m += " [hand-over published without a table]";
    if (flags & SL_ERR_HANDOFF_POLL) m += " [hand-over awaited without a table]";
    if (flags & SL_ERR_POOL) m += " [fragment pool exhausted: fragments dropped]";
    if (flags & SL_ERR_WIRE_SEGMENT) m += " [wireframe: record segment over 256]";
    if (flags & SL_ERR_WIRE_SLOTS) m += " [wireframe: candidate slots exhausted]";
    if (flags & SL_ERR_LINE_ALIAS) m += " [line: a segment visits texel (width, y) = (0, y + 1) twice: reference behaviour not reproduced]";
    if (flags & SL_ERR_GEOM_TIMEOUT) m += " [shared vertex stage: a rank never delivered its vertex blocks]";
    c->err = m;
    cudaMemsetAsync(&c->counters->errorFlags, 0, sizeof(unsigned), c->stream);
    if (flags & (SL_ERR_POOL | SL_ERR_WIRE_SLOTS)) return SLCU_ERR_OOM;
    return SLCU_ERR_UNSUPPORTED;
}

/* ===========================================================================
 * C ABI
 * =========================================================================*/
extern "C" {

int slcu_context_create(slcu_context** out, int device)
{
    if (!out) return SLCU_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return SLCU_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return SLCU_ERR_CUDA;
    slcu_context* c = new slcu_context{};
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return SLCU_ERR_CUDA; }
    if (cudaMalloc(&c->counters, sizeof(SLCounters)) != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return SLCU_ERR_CUDA; }
    cudaMemset(c->counters, 0, sizeof(SLCounters));
    if (cudaMallocHost(&c->hostCounters, 2 * sizeof(SLCounters)) != cudaSuccess) { cudaFree(c->counters); cudaStreamDestroy(c->stream); delete c; return SLCU_ERR_CUDA; }
    for (Scratch* x : all_scratch(c)) x->stream = c->stream;
    c->rcpValid = measure_host_rcp_table(c->rcpTable, c->err);
    c->rsqrtValid = measure_host_rsqrt_table(c->rsqrtTable);
    c->syncDraws = getenv("SLCU_SYNC_DRAWS") != nullptr && atoi(getenv("SLCU_SYNC_DRAWS")) != 0;
    *out = c;
    return 0;
}

void slcu_context_destroy(slcu_context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    flush_draws(c);
    cudaStreamSynchronize(c->stream);
    for (RtModule& m : c->rtModules) if (m.lib) cudaLibraryUnload(m.lib);
    if (c->batchDone) cudaEventDestroy(c->batchDone);
    if (c->timerA) cudaEventDestroy(c->timerA);
    if (c->timerB) cudaEventDestroy(c->timerB);
    for (auto& e : c->profEvents) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (cudaEvent_t e : c->eventPool) cudaEventDestroy(e);
    for (DevBuffer& b : c->vbos) if (b.ptr) cudaFree(b.ptr);
    for (HostIbo& b : c->ibos) if (b.buf.ptr) cudaFree(b.buf.ptr);
    for (HostTexture& t : c->textures) if (t.dev) cudaFree(t.dev);
    if (c->devTextures) cudaFree(c->devTextures);
    if (c->counters) cudaFree(c->counters);
    if (c->hostCounters) cudaFreeHost(c->hostCounters);
    for (auto& pm : c->peerMaps) if (pm.ptr) cudaIpcCloseMemHandle(pm.ptr);
    geom_release(c);
    if (c->copyStream) cudaStreamSynchronize(c->copyStream);
    for (cudaEvent_t& e : c->bandEvent) if (e) cudaEventDestroy(e);
    if (c->readDone) cudaEventDestroy(c->readDone);
    for (auto& ps : c->present)
    {
        ps.staging.release();
        if (ps.blitDone) cudaEventDestroy(ps.blitDone);
        if (ps.copyDone) cudaEventDestroy(ps.copyDone);
    }
    if (c->copyStream) cudaStreamDestroy(c->copyStream);
    for (Scratch* x : all_scratch(c)) x->release();
    cudaStreamSynchronize(c->stream);
    cudaStreamDestroy(c->stream);
    delete c;
}

unsigned slcu_num_threads(slcu_context*, unsigned n) { return n < 1 ? 1 : n; } /* src/SL_ProcessorPool.cpp:231-274: returns the value set */
const char* slcu_last_error(slcu_context* c) { return c ? c->err.c_str() : ""; }
const char* slcu_impl_name(void) { return "softlight-b200 (sm_100a CUDA)"; }

int slcu_set_rcp_table(slcu_context* c, const uint32_t* t)
{
    SLCU_FLUSH();
    if (!c || !t) return SLCU_ERR_ARG;
    memcpy(c->rcpTable, t, sizeof(c->rcpTable));
    c->rcpValid = true;
    return 0;
}
int slcu_set_rsqrt_table(slcu_context* c, const uint32_t* t)
{
    SLCU_FLUSH();
    if (!c || !t) return SLCU_ERR_ARG;
    memcpy(c->rsqrtTable, t, sizeof(c->rsqrtTable));
    c->rsqrtValid = true;
    return 0;
}
int slcu_get_rsqrt_table(slcu_context* c, uint32_t* t)
{
    if (!c || !t || !c->rsqrtValid) return SLCU_ERR_ARG;
    memcpy(t, c->rsqrtTable, sizeof(c->rsqrtTable));
    return 0;
}
int slcu_get_rcp_table(slcu_context* c, uint32_t* t)
{
    if (!c || !t || !c->rcpValid) return SLCU_ERR_ARG;
    memcpy(t, c->rcpTable, sizeof(c->rcpTable));
    return 0;
}

size_t slcu_create_vbo(slcu_context* c) { c->vbos.emplace_back(); return c->vbos.size() - 1; }
size_t slcu_create_ibo(slcu_context* c) { c->ibos.emplace_back(); return c->ibos.size() - 1; }
size_t slcu_create_vao(slcu_context* c) { c->vaos.emplace_back(); memset(&c->vaos.back().vao, 0, sizeof(SL_VertexArray)); return c->vaos.size() - 1; }
size_t slcu_create_ubo(slcu_context* c) { c->ubos.emplace_back(); memset(c->ubos.back().bytes, 0, 1024); return c->ubos.size() - 1; }
size_t slcu_create_texture(slcu_context* c) { c->textures.emplace_back(); return c->textures.size() - 1; }
size_t slcu_create_framebuffer(slcu_context* c) { c->fbos.emplace_back(); return c->fbos.size() - 1; }

/* ---- SL_Context::destroy_* : vector erase, later ids shift (see include/slcu.h) ---- */
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
static void remap_fixups_ubo(slcu_context* c, size_t gone)
{
    /* fix-ups are declared per uniform block (an index, like the block itself) */
    for (size_t i = 0; i < c->fixups.size();)
    {
        UboFixup& f = c->fixups[i];
        if (f.ubo == gone) { c->fixups.erase(c->fixups.begin() + i); continue; }
        if (f.ubo > gone) --f.ubo;
        ++i;
    }
}
int slcu_destroy_vbo(slcu_context* c, size_t id)
{
    SLCU_FLUSH();
    if (!c || id >= c->vbos.size()) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    if (c->vbos[id].ptr) cudaFree(c->vbos[id].ptr);
    c->vbos.erase(c->vbos.begin() + id);
    remap_fixups(c, 1, id);
    return 0;
}
int slcu_destroy_ibo(slcu_context* c, size_t id)
{
    SLCU_FLUSH();
    if (!c || id >= c->ibos.size()) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    if (c->ibos[id].buf.ptr) cudaFree(c->ibos[id].buf.ptr);
    c->ibos.erase(c->ibos.begin() + id);
    return 0;
}
int slcu_destroy_vao(slcu_context* c, size_t id)
{
    SLCU_FLUSH();
    if (!c || id >= c->vaos.size()) return SLCU_ERR_ARG;
    c->vaos.erase(c->vaos.begin() + id);
    return 0;
}
int slcu_destroy_ubo(slcu_context* c, size_t id)
{
    SLCU_FLUSH();
    if (!c || id >= c->ubos.size()) return SLCU_ERR_ARG;
    c->ubos.erase(c->ubos.begin() + id);
    remap_fixups_ubo(c, id);
    return 0;
}
int slcu_destroy_texture(slcu_context* c, size_t id)
{
    SLCU_FLUSH();
    if (!c || id >= c->textures.size()) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    if (c->textures[id].dev) cudaFree(c->textures[id].dev);
    c->textures.erase(c->textures.begin() + id);
    remap_fixups(c, 0, id);
    for (HostFbo& f : c->fbos)
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
    for (size_t i = 0; i < c->peerMaps.size();)
    {
        auto& pm = c->peerMaps[i];
        if (pm.tex == id) { if (pm.ptr) cudaIpcCloseMemHandle(pm.ptr); c->peerMaps.erase(c->peerMaps.begin() + i); continue; }
        if (pm.tex > id) --pm.tex;
        ++i;
    }
    return upload_texture_descs(c); /* the device-side descriptor array is indexed by id */
}
int slcu_destroy_framebuffer(slcu_context* c, size_t id)
{
    SLCU_FLUSH();
    if (!c || id >= c->fbos.size()) return SLCU_ERR_ARG;
    c->fbos.erase(c->fbos.begin() + id);
    return 0;
}
int slcu_destroy_shader(slcu_context* c, size_t id)
{
    SLCU_FLUSH();
    if (!c || id >= c->shaders.size()) return SLCU_ERR_ARG;
    c->shaders.erase(c->shaders.begin() + id);
    return 0;
}

int slcu_vbo_init(slcu_context* c, size_t vbo, size_t numBytes)
{
    SLCU_FLUSH();
    if (vbo >= c->vbos.size()) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    DevBuffer& b = c->vbos[vbo];
    CK(cudaStreamSynchronize(c->stream));
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.bytes = 0;
    CK(cudaMalloc(&b.ptr, numBytes + 64));
    CK(cudaMemsetAsync(b.ptr, 0, numBytes + 64, c->stream));
    b.bytes = numBytes;
    return 0;
}

int slcu_vbo_assign(slcu_context* c, size_t vbo, const void* data, size_t offset, size_t numBytes)
{
    SLCU_FLUSH();
    if (vbo >= c->vbos.size() || !data || offset + numBytes > c->vbos[vbo].bytes) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaMemcpyAsync(c->vbos[vbo].ptr + offset, data, numBytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream)); /* the caller may reuse `data` on return */
    return 0;
}

int slcu_ibo_init(slcu_context* c, size_t ibo, uint32_t numElements, int dataType, const void* data)
{
    SLCU_FLUSH();
    if (ibo >= c->ibos.size() || dataType < 0 || dataType > 2) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    HostIbo& b = c->ibos[ibo];
    CK(cudaStreamSynchronize(c->stream));
    if (b.buf.ptr) cudaFree(b.buf.ptr);
    const size_t bytes = (size_t)numElements << dataType;
    CK(cudaMalloc(&b.buf.ptr, bytes + 16)); /* padded like src/SL_IndexBuffer.cpp:138-161 */
    CK(cudaMemsetAsync(b.buf.ptr, 0, bytes + 16, c->stream));
    b.buf.bytes = bytes;
    b.type = dataType;
    b.count = numElements;
    b.maxIndex = 0;
    if (data)
    {
        CK(cudaMemcpyAsync(b.buf.ptr, data, bytes, cudaMemcpyHostToDevice, c->stream));
        for (uint32_t i = 0; i < numElements; ++i)
        {
            const unsigned long long v = dataType == 0 ? ((const uint8_t*)data)[i] : (dataType == 1 ? ((const uint16_t*)data)[i] : ((const uint32_t*)data)[i]);
            if (v > b.maxIndex) b.maxIndex = v;
        }
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int slcu_vao_set_vertex_buffer(slcu_context* c, size_t vao, size_t vbo) { if (vao >= c->vaos.size()) return SLCU_ERR_ARG; c->vaos[vao].vbo = vbo; return 0; }
int slcu_vao_set_index_buffer(slcu_context* c, size_t vao, size_t ibo) { if (vao >= c->vaos.size()) return SLCU_ERR_ARG; c->vaos[vao].ibo = ibo; return 0; }
int slcu_vao_set_num_bindings(slcu_context* c, size_t vao, int n)
{
    if (vao >= c->vaos.size() || n < 0 || n >= 8) return -1; /* src/SL_VertexArray.cpp:162-167 */
    c->vaos[vao].numBindings = n;
    return n;
}
int slcu_vao_set_binding(slcu_context* c, size_t vao, size_t bind, ptrdiff_t offset, ptrdiff_t stride, int dim, int type)
{
    if (vao >= c->vaos.size() || bind >= 8) return SLCU_ERR_ARG;
    c->vaos[vao].vao.mBindings[bind] = SL_VertexArray::Binding{dim, type, offset, stride};
    return 0;
}

int slcu_ubo_assign(slcu_context* c, size_t ubo, const void* data, size_t offset, size_t numBytes)
{
    if (ubo >= c->ubos.size() || !data || offset + numBytes > 1024) return SLCU_ERR_ARG;
    memcpy(c->ubos[ubo].bytes + offset, data, numBytes);
    return 0;
}
static void set_fixup(slcu_context* c, const UboFixup& f)
{
    for (UboFixup& g : c->fixups) { if (g.ubo == f.ubo && g.offset == f.offset) { g = f; return; } }
    c->fixups.push_back(f);
}
int slcu_ubo_bind_texture(slcu_context* c, size_t ubo, size_t offset, size_t tex)
{
    if (ubo >= c->ubos.size() || offset + 8 > 1024 || tex >= c->textures.size()) return SLCU_ERR_ARG;
    set_fixup(c, UboFixup{ubo, offset, 0, tex, 0});
    return 0;
}
int slcu_ubo_bind_buffer(slcu_context* c, size_t ubo, size_t offset, size_t vbo, size_t byteOffset)
{
    if (ubo >= c->ubos.size() || offset + 8 > 1024 || vbo >= c->vbos.size()) return SLCU_ERR_ARG;
    set_fixup(c, UboFixup{ubo, offset, 1, vbo, byteOffset});
    return 0;
}

int slcu_texture_init(slcu_context* c, size_t tex, int type, uint16_t w, uint16_t h, uint16_t d)
{
    SLCU_FLUSH();
    if (tex >= c->textures.size() || type < 0 || type >= 33) return -1;
    cudaSetDevice(c->device);
    HostTexture& t = c->textures[tex];
    CK(cudaStreamSynchronize(c->stream));
    if (t.dev) cudaFree(t.dev);
    t.dev = nullptr;
    t.shadow.clear();
    t.w = w; t.h = h; t.d = d; t.type = type;
    t.channels = (uint8_t)type_channels(type);
    t.bpt = (uint8_t)type_bpt(type);
    /* _sl_allocate_texture (src/SL_Texture.cpp:28-54): w,h padded up by 4-(dim%4), zero-filled.  The reference also
     * pads the depth to 4+ slices even for 2-D textures; the device copy keeps only the slices that can be addressed. */
    const size_t pw = w + (4 - (w % 4)), ph = h + (4 - (h % 4)), pd = (d <= 1) ? 1 : (size_t)d + (4 - (d % 4));
    t.allocBytes = pw * ph * pd * t.bpt + 4 * t.bpt;
    CK(cudaMalloc(&t.dev, t.allocBytes));
    CK(cudaMemsetAsync(t.dev, 0, t.allocBytes, c->stream));
    return upload_texture_descs(c);
}

int slcu_texture_set_texels(slcu_context* c, size_t tex, int order, uint16_t x, uint16_t y, uint16_t z, uint16_t w, uint16_t h, uint16_t d, const void* data)
{
    SLCU_FLUSH();
    if (tex >= c->textures.size() || !data) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    HostTexture& t = c->textures[tex];
    if (!t.dev) return SLCU_ERR_ARG;
    if (t.shadow.empty())
    {
        t.shadow.assign(t.allocBytes, 0);
        CK(cudaMemcpyAsync(t.shadow.data(), t.dev, t.allocBytes, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    /* SL_Texture::set_texels<order> (SL_Texture.hpp:418-444) on the shadow, then one upload */
    const unsigned char* src = (const unsigned char*)data;
    const uint32_t W = t.w, Hh = t.h;
    for (uint32_t k = 0, z0 = z; k < d; ++k, ++z0)
        for (uint32_t j = 0, y0 = y; j < h; ++j, ++y0)
            for (uint32_t i = 0, x0 = x; i < w; ++i, ++x0)
            {
                const size_t index = i + (size_t)w * (j + ((size_t)h * k));
                size_t dst;
                if (order == 0) dst = (t.d <= 1) ? (x0 + (size_t)W * y0) : (x0 + (size_t)W * (y0 + (size_t)Hh * z0));
                else if (t.d <= 1) dst = ((x0 & 3u) + ((y0 & 3u) << 2)) + (size_t)((x0 >> 2) + (W >> 2) * (y0 >> 2)) * 16u;
                else dst = ((x0 & 3u) + (((y0 & 3u) << 2) + (4u * ((z0 & 3u) << 2)))) + (size_t)((x0 >> 2) + ((W >> 2) * ((y0 >> 2) + ((Hh >> 2) * (z0 >> 2))))) * 64u;
                if ((dst + 1) * t.bpt > t.allocBytes) return SLCU_ERR_ARG;
                memcpy(t.shadow.data() + dst * t.bpt, src + index * t.bpt, t.bpt);
            }
    CK(cudaMemcpyAsync(t.dev, t.shadow.data(), t.allocBytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int slcu_texture_read(slcu_context* c, size_t tex, void* out, size_t numBytes)
{
    if (!c || tex >= c->textures.size() || !out || numBytes > c->textures[tex].allocBytes) return SLCU_ERR_ARG;
    /* the pending batch may be the one that renders into this texture: tell it, so that its rows start crossing the link
     * while the rest of the frame is still being shaded (enqueue_batch) */
    const HostTexture& t = c->textures[tex];
    c->pendingRead = slcu_context::PendingRead{};
    if (!c->batch.descs.empty() && !c->batch.wire && numBytes >= (size_t)t.w * t.h * t.bpt)
    {
        c->pendingRead.active = true;
        c->pendingRead.tex = tex;
        c->pendingRead.out = (unsigned char*)out;
        c->pendingRead.bytes = numBytes;
    }
    const int frc = c->batch.descs.empty() ? 0 : flush_draws(c);
    const bool issued = c->pendingRead.issued;
    c->pendingRead = slcu_context::PendingRead{};
    if (frc) return frc;
    cudaSetDevice(c->device);
    if (issued)
    {
        CK(cudaEventSynchronize(c->readDone)); /* the last band's copy; the batch itself was joined by run_batch */
        return 0;
    }
    CK(cudaMemcpyAsync(out, c->textures[tex].dev, numBytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

/* page-locked host memory for read-back / present targets and per-frame inputs: the copy engine reaches the link rate
 * (~56 GB/s on this box) only from / into pinned pages; pageable memory goes through the driver's bounce buffer (~19 GB/s) */
int slcu_host_alloc(slcu_context* c, void** out, size_t numBytes)
{
    if (!c || !out) return SLCU_ERR_ARG;
    *out = nullptr;
    cudaSetDevice(c->device);
    CK(cudaHostAlloc(out, numBytes ? numBytes : 1, cudaHostAllocPortable));
    return 0;
}
int slcu_host_free(slcu_context* c, void* p)
{
    if (!c) return SLCU_ERR_ARG;
    if (!p) return 0;
    SLCU_FLUSH();
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream)); /* a read-back into it may be in flight */
    CK(cudaFreeHost(p));
    return 0;
}

int slcu_fbo_reserve_color_buffers(slcu_context* c, size_t fbo, unsigned n)
{
    SLCU_FLUSH();
    if (fbo >= c->fbos.size()) return SLCU_ERR_ARG;
    /* src/SL_Framebuffer.cpp:823-848: same count = no-op (checked first), the colour views are reset, the depth view stays */
    HostFbo& f = c->fbos[fbo];
    if (n == f.numColors) return 0;
    if (n < 1) return -1;
    if (n > 4) return -2;
    for (int i = 0; i < 4; ++i) { f.colors[i] = (size_t)-1; f.colorViews[i] = FboView{}; }
    f.numColors = n;
    return 0;
}
int slcu_fbo_attach_color_buffer(slcu_context* c, size_t fbo, unsigned index, size_t tex)
{
    SLCU_FLUSH();
    if (fbo >= c->fbos.size() || tex >= c->textures.size()) return SLCU_ERR_ARG;
    if (index >= c->fbos[fbo].numColors) return -1;
    c->fbos[fbo].colors[index] = tex;
    const HostTexture& t = c->textures[tex];
    c->fbos[fbo].colorViews[index] = FboView{t.dev != nullptr, t.w, t.h, t.d, t.type};
    return 0;
}
int slcu_fbo_attach_depth_buffer(slcu_context* c, size_t fbo, size_t tex)
{
    SLCU_FLUSH();
    if (fbo >= c->fbos.size() || tex >= c->textures.size()) return SLCU_ERR_ARG;
    c->fbos[fbo].depth = tex;
    const HostTexture& t = c->textures[tex];
    c->fbos[fbo].depthView = FboView{t.dev != nullptr, t.w, t.h, t.d, t.type};
    return 0;
}
/* SL_Framebuffer::valid (src/SL_Framebuffer.cpp:928-990): same checks, same order, same codes.  An attachment that was
 * never made (or whose texture was never initialised) is the reference's reset view: no texels, zero dimensions. */
static int fbo_valid_code(const HostFbo& f)
{
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
    if (d.type != 4 && d.type != 5 && d.type != 6) return -10; /* SL_COLOR_R_HALF / R_FLOAT / R_DOUBLE */
    return 0;
}

int slcu_fbo_valid(slcu_context* c, size_t fbo)
{
    if (!c || fbo >= c->fbos.size()) return SLCU_ERR_ARG;
    return fbo_valid_code(c->fbos[fbo]);
}

/* ---- run-time shader registration -----------------------------------------------------------------------------------
 * The reference accepts any host function pointer as a shader (SL_Shader.hpp:63, 97).  A GPU cannot call into host code,
 * so the counterpart is SOURCE: the body of the user's vertex / fragment function written against the device mirrors of
 * ls::math and the SL_* shader API (the dialect of shaders/(*).inl), plus SL_REGISTER_VERTEX_SHADER("name", fn) /
 * SL_REGISTER_FRAGMENT_SHADER("name", fn) lines.  NVRTC instantiates the pipeline's kernel templates for each registered
 * function -- direct, inlinable calls, never interpreted -- against the very headers this library was built from (embedded
 * at build time), and the result is loaded into the context's device.  No rebuild of libslcu.so. */
static const struct { const char* name; const char* text; } kEmbeddedHeaders[] = {
#include "_embedded_headers.inc"
};

struct ParsedRegistration { bool vertex; std::string name, fn; };
static std::vector<ParsedRegistration> parse_registrations(const std::string& src)
{
    std::vector<ParsedRegistration> out;
    for (int kind = 0; kind < 2; ++kind)
    {
        const std::string macro = kind == 0 ? "SL_REGISTER_VERTEX_SHADER" : "SL_REGISTER_FRAGMENT_SHADER";
        size_t pos = 0;
        while ((pos = src.find(macro, pos)) != std::string::npos)
        {
            pos += macro.size();
            const size_t open = src.find('(', pos);
            const size_t q0 = src.find('"', pos);
            if (open == std::string::npos || q0 == std::string::npos) break;
            const size_t q1 = src.find('"', q0 + 1);
            const size_t comma = src.find(',', q1);
            const size_t close = src.find(')', comma);
            if (q1 == std::string::npos || comma == std::string::npos || close == std::string::npos) break;
            std::string fn = src.substr(comma + 1, close - comma - 1);
            fn.erase(0, fn.find_first_not_of(" \t\r\n"));
            fn.erase(fn.find_last_not_of(" \t\r\n") + 1);
            out.push_back(ParsedRegistration{kind == 0, src.substr(q0 + 1, q1 - q0 - 1), fn});
            pos = close;
        }
    }
    return out;
}

struct CompiledShaders
{
    std::vector<char> cubin;
    std::vector<ParsedRegistration> regs;
    std::vector<std::string> lowered; /* per registration: 4 kernel names (vertex: k_vertex<,false>, k_vertex<,true>, k_point_vertex, k_line_vertex; fragment: k_shade, k_point_shade, k_wire_shade, k_line_shade) */
    std::string rcpName, rsqrtName;
};

static int compile_shader_source(const char* source, CompiledShaders& out, std::string& log)
{
    if (!source) { log = "no source"; return SLCU_ERR_ARG; }
    out.regs = parse_registrations(source);
    if (out.regs.empty()) { log = "the source registers no shader (SL_REGISTER_VERTEX_SHADER / SL_REGISTER_FRAGMENT_SHADER)"; return SLCU_ERR_ARG; }
    std::string tu = "#include \"sl_pipeline.cuh\"\n#include \"sl_points.cuh\"\n#include \"sl_wire.cuh\"\n#include \"sl_lines.cuh\"\n"
                     "#define SL_SHADER_FN static __device__ __forceinline__\n"
                     "#define SL_REGISTER_VERTEX_SHADER(name, fn)\n#define SL_REGISTER_FRAGMENT_SHADER(name, fn)\n"
                     "namespace sld {\n#line 1 \"user_shader\"\n";
    tu += source;
    tu += "\n}\n";
    /* the pipeline's headers, and just enough of the C headers they name (NVRTC has no host headers) */
    std::vector<const char*> names, texts;
    for (const auto& h : kEmbeddedHeaders) { names.push_back(h.name); texts.push_back(h.text); }
    names.push_back("stdint.h");
    texts.push_back("#pragma once\ntypedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t; typedef int int32_t; "
                    "typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t; typedef unsigned long uintptr_t; typedef long intptr_t;\n");
    names.push_back("stddef.h");
    texts.push_back("#pragma once\ntypedef unsigned long size_t; typedef long ptrdiff_t;\n#ifndef offsetof\n#define offsetof(t, m) __builtin_offsetof(t, m)\n#endif\n");
    names.push_back("string.h");
    texts.push_back("#pragma once\n");
    nvrtcProgram prog = nullptr;
    if (nvrtcCreateProgram(&prog, tu.c_str(), "slcu_runtime_shaders.cu", (int)names.size(), texts.data(), names.data()) != NVRTC_SUCCESS) { log = "nvrtcCreateProgram failed"; return SLCU_ERR_CUDA; }
    std::vector<std::string> exprs;
    for (const ParsedRegistration& r : out.regs)
    {
        if (r.vertex)
        {
            exprs.push_back("&sld::k_vertex<sld::" + r.fn + ", false>");
            exprs.push_back("&sld::k_vertex<sld::" + r.fn + ", true>");
            exprs.push_back("&sld::k_point_vertex<sld::" + r.fn + ">");
            exprs.push_back("&sld::k_line_vertex<sld::" + r.fn + ">");
        }
        else
        {
            exprs.push_back("&sld::k_shade<sld::" + r.fn + ">");
            exprs.push_back("&sld::k_point_shade<sld::" + r.fn + ">");
            exprs.push_back("&sld::k_wire_shade<sld::" + r.fn + ">");
            exprs.push_back("&sld::k_line_shade<sld::" + r.fn + ">");
        }
    }
    exprs.push_back("&sld::g_rsqrt_table");
    exprs.push_back("&sld::g_rcp_table");
    for (const std::string& e : exprs) nvrtcAddNameExpression(prog, e.c_str());
    const char* cudaHome = getenv("CUDA_HOME") ? getenv("CUDA_HOME") : "/usr/local/cuda";
    const std::string inc = std::string("-I") + cudaHome + "/include";
    /* the flags libslcu.so itself is built with: nothing is contracted behind the arithmetic's back (softlight_b200/build.py) */
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-fmad=false", "-lineinfo", "-diag-suppress=177", "-diag-suppress=9", inc.c_str()};
    const nvrtcResult rc = nvrtcCompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts);
    size_t logSize = 0;
    nvrtcGetProgramLogSize(prog, &logSize);
    log.assign(logSize, '\0');
    if (logSize) nvrtcGetProgramLog(prog, &log[0]);
    if (rc != NVRTC_SUCCESS) { nvrtcDestroyProgram(&prog); return SLCU_ERR_ARG; }
    for (size_t i = 0; i < exprs.size(); ++i)
    {
        const char* low = nullptr;
        if (nvrtcGetLoweredName(prog, exprs[i].c_str(), &low) != NVRTC_SUCCESS || !low) { log = "no lowered name for " + exprs[i]; nvrtcDestroyProgram(&prog); return SLCU_ERR_CUDA; }
        if (i + 1 == exprs.size()) out.rcpName = low;
        else if (i + 2 == exprs.size()) out.rsqrtName = low;
        else out.lowered.push_back(low);
    }
    size_t bytes = 0;
    nvrtcGetCUBINSize(prog, &bytes);
    out.cubin.resize(bytes);
    nvrtcGetCUBIN(prog, out.cubin.data());
    nvrtcDestroyProgram(&prog);
    return 0;
}

/* compile only (no device needed): 0 when the source would be accepted, the compiler's log in `log` either way */
int slcu_shader_source_check(const char* source, char* logOut, size_t logBytes)
{
    CompiledShaders cs;
    std::string log;
    const int rc = compile_shader_source(source, cs, log);
    if (logOut && logBytes) { strncpy(logOut, log.c_str(), logBytes - 1); logOut[logBytes - 1] = '\0'; }
    return rc ? rc : (int)cs.regs.size();
}

/* compile, load, register: returns the number of shaders the source registered, or a negative error (log in last_error) */
int slcu_register_shader_source(slcu_context* c, const char* source)
{
    SLCU_FLUSH();
    if (!c) return SLCU_ERR_ARG;
    CompiledShaders cs;
    std::string log;
    int rc = compile_shader_source(source, cs, log);
    if (rc) { c->err = "shader source rejected: " + log; return rc; }
    cudaSetDevice(c->device);
    RtModule m;
    CK(cudaLibraryLoadData(&m.lib, cs.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    size_t rcpBytes = 0;
    CK(cudaLibraryGetGlobal(&m.rcpTable, &rcpBytes, m.lib, cs.rcpName.c_str()));
    if (rcpBytes != sizeof(c->rcpTable)) { c->err = "run-time module: unexpected rcp table size"; return SLCU_ERR_CUDA; }
    if (c->rcpValid) CK(cudaMemcpy(m.rcpTable, c->rcpTable, sizeof(c->rcpTable), cudaMemcpyHostToDevice));
    size_t rsqrtBytes = 0;
    CK(cudaLibraryGetGlobal(&m.rsqrtTable, &rsqrtBytes, m.lib, cs.rsqrtName.c_str()));
    if (rsqrtBytes != sizeof(c->rsqrtTable)) { c->err = "run-time module: unexpected rsqrt table size"; return SLCU_ERR_CUDA; }
    if (c->rsqrtValid) CK(cudaMemcpy(m.rsqrtTable, c->rsqrtTable, sizeof(c->rsqrtTable), cudaMemcpyHostToDevice));
    for (size_t i = 0; i < cs.regs.size(); ++i)
    {
        cudaKernel_t k[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int j = 0; j < 4; ++j) CK(cudaLibraryGetKernel(&k[j], m.lib, cs.lowered[4 * i + j].c_str()));
        if (cs.regs[i].vertex)
        {
            CK(cudaFuncSetAttribute(reinterpret_cast<const void*>(k[1]), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVertexCacheMaxBytes));
            RtVsEntry e;
            e.name = cs.regs[i].name;
            e.k.vertex[0] = k[0]; e.k.vertex[1] = k[1]; e.k.pointVertex = k[2]; e.k.lineVertex = k[3];
            bool replaced = false;
            for (RtVsEntry& old : c->rtVs) if (old.name == e.name) { old = e; replaced = true; }
            if (!replaced) c->rtVs.push_back(e);
        }
        else
        {
            RtFsEntry e;
            e.name = cs.regs[i].name;
            e.k.shade = k[0]; e.k.pointShade = k[1]; e.k.wireShade = k[2]; e.k.lineShade = k[3];
            bool replaced = false;
            for (RtFsEntry& old : c->rtFs) if (old.name == e.name) { old = e; replaced = true; }
            if (!replaced) c->rtFs.push_back(e);
        }
    }
    c->rtModules.push_back(m);
    return (int)cs.regs.size();
}

size_t slcu_create_shader(slcu_context* c, const slcu_vertex_shader* vs, const slcu_fragment_shader* fs, size_t ubo)
{
    /* validation of SL_Context::create_shader (src/SL_Context.cpp:488-540) */
    if (!vs || !fs || !vs->shader || !fs->shader || ubo >= c->ubos.size()) return SLCU_INVALID_ID;
    if (vs->numVaryings > 4 || fs->numVaryings > vs->numVaryings || fs->numOutputs > 4 || fs->numOutputs < 1) return SLCU_INVALID_ID;
    if (vs->cullMode > 2 || fs->blend > 4 || fs->depthTest > 6 || fs->depthMask > 1) return SLCU_INVALID_ID;
    VertexLauncher vl = nullptr;
    RasterLauncher fl = nullptr;
    PointVertexLauncher pvl = nullptr;
    PointShadeLauncher pfl = nullptr;
    LineVertexLauncher lvl = nullptr;
    LineShadeLauncher lfl = nullptr;
    WireShadeLauncher wfl = nullptr;
    for (const VsEntry& e : vs_registry()) if (!strcmp(e.name, vs->shader)) { vl = e.launch; pvl = e.launchPoints; lvl = e.launchLines; }
    for (const FsEntry& e : fs_registry()) if (!strcmp(e.name, fs->shader)) { fl = e.launch; pfl = e.launchPoints; wfl = e.launchWire; lfl = e.launchLines; }
    RtVertexKernels rtV;
    RtFragmentKernels rtF;
    for (const RtVsEntry& e : c->rtVs) if (e.name == vs->shader) rtV = e.k; /* a run-time registration shadows a built-in of the same name */
    for (const RtFsEntry& e : c->rtFs) if (e.name == fs->shader) rtF = e.k;
    if ((!vl && !rtV.vertex[0]) || (!fl && !rtF.shade))
    {
        c->err = "shader not registered as a __device__ callable (no CPU fallback)";
        return SLCU_INVALID_ID;
    }
    HostShader s;
    s.cullMode = vs->cullMode; s.depthTest = fs->depthTest; s.depthMask = fs->depthMask; s.blend = fs->blend;
    s.numVaryings = fs->numVaryings; s.numOutputs = fs->numOutputs; s.vs = vl; s.fs = fl; s.pointVs = pvl; s.pointFs = pfl; s.wireFs = wfl; s.lineVs = lvl; s.lineFs = lfl; s.ubo = ubo; s.rtV = rtV; s.rtF = rtF;
    c->shaders.push_back(s);
    return c->shaders.size() - 1;
}

int slcu_viewport(slcu_context* c, int32_t x, int32_t y, uint16_t w, uint16_t h) { c->viewport[0] = x; c->viewport[1] = y; c->viewport[2] = w; c->viewport[3] = h; return 0; }
int slcu_scissor(slcu_context* c, int32_t x, int32_t y, uint16_t w, uint16_t h) { c->scissor[0] = x; c->scissor[1] = y; c->scissor[2] = w; c->scissor[3] = h; return 0; }

int slcu_clear_color_buffer(slcu_context* c, size_t fbo, unsigned att, const double rgba[4])
{
    SLCU_FLUSH();
    if (fbo >= c->fbos.size() || att >= c->fbos[fbo].numColors || c->fbos[fbo].colors[att] >= c->textures.size()) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    const HostTexture& t = c->textures[c->fbos[fbo].colors[att]];
    unsigned char texel[32];
    host_match_color(t.type, rgba, texel);
    return clear_texture(c, t, texel);
}
int slcu_clear_depth_buffer(slcu_context* c, size_t fbo, double depth)
{
    SLCU_FLUSH();
    if (fbo >= c->fbos.size() || c->fbos[fbo].depth >= c->textures.size()) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    const HostTexture& t = c->textures[c->fbos[fbo].depth];
    unsigned char texel[8];
    if (!host_depth_texel(t, depth, texel)) return 0; /* silently ignored like src/SL_Context.cpp:904-920 */
    c->textures[c->fbos[fbo].depth].depthWritesSinceClear = 0;
    return clear_texture(c, t, texel);
}
int slcu_clear_framebuffer(slcu_context* c, size_t fbo, unsigned att, const double rgba[4], double depth)
{
    SLCU_FLUSH();
    if (fbo >= c->fbos.size() || c->fbos[fbo].depth >= c->textures.size()) return SLCU_ERR_ARG;
    const HostTexture& d = c->textures[c->fbos[fbo].depth];
    unsigned char texel[8];
    if (!host_depth_texel(d, depth, texel)) return 0; /* src/SL_Context.cpp:944-960 returns before clearing anything */
    int rc = slcu_clear_color_buffer(c, fbo, att, rgba);
    if (rc) return rc;
    c->textures[c->fbos[fbo].depth].depthWritesSinceClear = 0;
    return clear_texture(c, d, texel);
}

int slcu_draw(slcu_context* c, const slcu_mesh* m, size_t shader, size_t fbo) { return run_draw(c, m, 1, 1, shader, fbo); }
int slcu_draw_multiple(slcu_context* c, const slcu_mesh* m, size_t n, size_t shader, size_t fbo) { return run_draw(c, m, n, 1, shader, fbo); }
int slcu_draw_instanced(slcu_context* c, const slcu_mesh* m, size_t inst, size_t shader, size_t fbo) { return run_draw(c, m, 1, inst, shader, fbo); }

/* run_blit_processors replacement (src/SL_ProcessorPool.cpp:400-445) */
/* SL_BlitProcessor::blit_nearest for one (source texture, destination texels) pair: the destination is a texture of this
 * context or a device staging copy of a host view */
static int blit_core(slcu_context* c, const HostTexture& s, unsigned char* dstTexels, int dstType, unsigned dstW, unsigned dstH,
                     uint16_t sx0, uint16_t sy0, uint16_t sx1, uint16_t sy1, uint16_t dx0, uint16_t dy0, uint16_t dx1, uint16_t dy1)
{
    SLBlitParams p;
    const bool rgIntSame = s.type == dstType && !type_is_packed(s.type) && s.type / 7 == 1 && s.type % 7 <= 3; /* RG integer -> same type is p / max(T) in the reference (SL_Color.hpp:373-380) */
    if (s.type == dstType && !rgIntSame) p.mode = 0;
    else if (s.type == 26 && dstType == 21) p.mode = 1; /* RGBA_FLOAT -> RGBA_8U */
    else if (s.type == 21 && dstType == 26) p.mode = 2; /* RGBA_8U -> RGBA_FLOAT */
    else p.mode = 3; /* the general conversion (sl_texel.cuh): any pair of the 33 colour types, SL_BlitProcessor's and SL_BlitCompressedProcessor's */
    p.srcType = s.type;
    p.dstType = dstType;
    if (sx1 < sx0 || sy1 < sy0 || dx1 <= dx0 || dy1 <= dy0) return 0;
    p.src = s.dev; p.dst = dstTexels;
    p.srcW = s.w; p.srcH = s.h; p.dstW = (uint16_t)dstW; p.dstH = (uint16_t)dstH;
    p.srcX0 = sx0; p.srcY0 = sy0; p.srcX1 = sx1; p.srcY1 = sy1;
    p.dstX0 = dx0; p.dstY0 = dy0; p.dstX1 = dx1; p.dstY1 = dy1;
    p.bpt = (int)type_bpt(dstType);
    const unsigned outW = (unsigned)dx1 - dx0, outH = (unsigned)dy1 - dy0;
    const dim3 grid((outW + 255u) / 256u, outH < 148u * 8u ? outH : 148u * 8u);
    { KernelScope ks(c, SLCU_K_BLIT); k_blit<<<grid, 256, 0, c->stream>>>(p); }
    CK(cudaGetLastError());
    return 0;
}

int slcu_blit_rect(slcu_context* c, size_t outTex, size_t inTex, uint16_t sx0, uint16_t sy0, uint16_t sx1, uint16_t sy1,
                   uint16_t dx0, uint16_t dy0, uint16_t dx1, uint16_t dy1)
{
    SLCU_FLUSH();
    if (!c || outTex >= c->textures.size() || inTex >= c->textures.size()) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    const HostTexture& s = c->textures[inTex];
    const HostTexture& d = c->textures[outTex];
    if (!s.dev || !d.dev || !d.w || !d.h) return SLCU_ERR_ARG;
    return blit_core(c, s, d.dev, d.type, d.w, d.h, sx0, sy0, sx1, sy1, dx0, dy0, dx1, dy1);
}

/* ---- the present path: SL_Context::blit(SL_TextureView& buffer, size_t textureId[, rectangles]) (src/SL_Context.cpp:816-874),
 * the blit into a host-side view that SL_Swapchain / SL_WindowBuffer hand out as the window's back buffer.  The blit runs on
 * the device into a staging copy of the view's geometry; the rectangle it wrote then crosses the link into the host view
 * (every other texel of the view is left alone, like the reference's in-place blit). */
static int present_enqueue(slcu_context* c, unsigned slot, void* texels, int colorType, uint16_t w, uint16_t h, size_t inTex,
                           uint16_t sx0, uint16_t sy0, uint16_t sx1, uint16_t sy1, uint16_t dx0, uint16_t dy0, uint16_t dx1, uint16_t dy1)
{
    if (!c || slot > 1 || !texels || inTex >= c->textures.size() || colorType < 0 || colorType >= 33 || !w || !h) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    const HostTexture& s = c->textures[inTex];
    if (!s.dev) return SLCU_ERR_ARG;
    if (dx1 > w) dx1 = w;
    if (dy1 > h) dy1 = h;
    slcu_context::PresentSlot& ps = c->present[slot];
    if (!c->copyStream) CK(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
    if (!ps.blitDone)
    {
        CK(cudaEventCreateWithFlags(&ps.blitDone, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ps.copyDone, cudaEventDisableTiming));
    }
    const unsigned bpt = type_bpt(colorType);
    if (ps.busy) CK(cudaStreamWaitEvent(c->stream, ps.copyDone, 0)); /* the staging copy's previous contents have left */
    ps.staging.stream = c->stream;
    CK(ps.staging.ensure((size_t)w * h * bpt));
    const int rc = blit_core(c, s, ps.staging.buf.ptr, colorType, w, h, sx0, sy0, sx1, sy1, dx0, dy0, dx1, dy1);
    if (rc) return rc;
    CK(cudaEventRecord(ps.blitDone, c->stream));
    CK(cudaStreamWaitEvent(c->copyStream, ps.blitDone, 0));
    if (!(sx1 < sx0 || sy1 < sy0 || dx1 <= dx0 || dy1 <= dy0))
    {
        const size_t pitch = (size_t)w * bpt, off = (size_t)dy0 * pitch + (size_t)dx0 * bpt;
        if (dx0 == 0 && dx1 == w) CK(cudaMemcpyAsync((unsigned char*)texels + off, ps.staging.buf.ptr + off, (size_t)(dy1 - dy0) * pitch, cudaMemcpyDeviceToHost, c->copyStream));
        else CK(cudaMemcpy2DAsync((unsigned char*)texels + off, pitch, ps.staging.buf.ptr + off, pitch, (size_t)(dx1 - dx0) * bpt, (size_t)(dy1 - dy0), cudaMemcpyDeviceToHost, c->copyStream));
    }
    CK(cudaEventRecord(ps.copyDone, c->copyStream));
    ps.busy = true;
    return 0;
}

int slcu_blit_to_host_rect(slcu_context* c, void* texels, int colorType, uint16_t w, uint16_t h, size_t inTex,
                           uint16_t sx0, uint16_t sy0, uint16_t sx1, uint16_t sy1, uint16_t dx0, uint16_t dy0, uint16_t dx1, uint16_t dy1)
{
    SLCU_FLUSH();
    const int rc = present_enqueue(c, 0, texels, colorType, w, h, inTex, sx0, sy0, sx1, sy1, dx0, dy0, dx1, dy1);
    if (rc) return rc;
    CK(cudaEventSynchronize(c->present[0].copyDone)); /* synchronous, like the reference's blit */
    return 0;
}

int slcu_blit_to_host(slcu_context* c, void* texels, int colorType, uint16_t w, uint16_t h, size_t inTex)
{
    if (!c || inTex >= c->textures.size()) return SLCU_ERR_ARG;
    const HostTexture& s = c->textures[inTex];
    return slcu_blit_to_host_rect(c, texels, colorType, w, h, inTex, 0, 0, s.w, s.h, 0, 0, w, h);
}

/* ours: the same blit without the join -- the call returns when the work is enqueued, the next frame renders while this
 * one crosses the link (a two-deep swap chain: slot 0 / 1, one staging copy each).  present_wait(slot) is the join. */
int slcu_present(slcu_context* c, unsigned slot, void* texels, int colorType, uint16_t w, uint16_t h, size_t inTex)
{
    SLCU_FLUSH();
    if (!c || inTex >= c->textures.size()) return SLCU_ERR_ARG;
    const HostTexture& s = c->textures[inTex];
    return present_enqueue(c, slot, texels, colorType, w, h, inTex, 0, 0, s.w, s.h, 0, 0, w, h);
}

int slcu_present_wait(slcu_context* c, unsigned slot)
{
    if (!c || slot > 1) return SLCU_ERR_ARG;
    if (!c->present[slot].busy) return 0;
    cudaSetDevice(c->device);
    CK(cudaEventSynchronize(c->present[slot].copyDone));
    return 0;
}

int slcu_blit(slcu_context* c, size_t outTex, size_t inTex)
{
    /* SL_Context::blit(out, in) (src/SL_Context.cpp:747-766): whole source onto whole destination */
    if (!c || outTex >= c->textures.size() || inTex >= c->textures.size()) return SLCU_ERR_ARG;
    const HostTexture& s = c->textures[inTex];
    const HostTexture& d = c->textures[outTex];
    return slcu_blit_rect(c, outTex, inTex, 0, 0, s.w, s.h, 0, 0, d.w, d.h);
}

/* ---- sort-first multi-GPU ---- */
int slcu_set_partition(slcu_context* c, unsigned rank, unsigned world, unsigned stripRows)
{
    SLCU_FLUSH();
    if (!c || world < 1 || rank >= world || stripRows < 1) return SLCU_ERR_ARG;
    if (c->geom.base && (rank != c->partRank || world != c->partWorld || stripRows != c->partStripRows))
    {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        geom_release(c); /* the shared vertex stage was set up for another partition */
    }
    c->partRank = rank;
    c->partWorld = world;
    c->partStripRows = stripRows;
    return 0;
}

int slcu_texture_ipc_export(slcu_context* c, size_t tex, void* handle64)
{
    SLCU_FLUSH();
    if (!c || tex >= c->textures.size() || !handle64 || !c->textures[tex].dev) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->textures[tex].dev));
    memcpy(handle64, &h, 64);
    return 0;
}

int slcu_texture_ipc_import(slcu_context* c, size_t tex, unsigned peerRank, const void* handle64)
{
    SLCU_FLUSH();
    if (!c || tex >= c->textures.size() || !handle64) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    for (auto& pm : c->peerMaps)
    {
        if (pm.tex == tex && pm.rank == peerRank) { cudaIpcCloseMemHandle(pm.ptr); pm.ptr = ptr; return 0; }
    }
    c->peerMaps.push_back(slcu_context::PeerMap{tex, peerRank, ptr});
    return 0;
}

int slcu_composite(slcu_context* c, size_t srcTex, size_t dstTex, unsigned root)
{
    SLCU_FLUSH();
    if (!c || srcTex >= c->textures.size() || dstTex >= c->textures.size()) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    const HostTexture& s = c->textures[srcTex];
    const HostTexture& d = c->textures[dstTex];
    if (s.w != d.w || s.h != d.h || s.bpt != d.bpt) { c->err = "composite: source and destination differ in size or format"; return SLCU_ERR_ARG; }
    SLCompositeParams p;
    p.src = s.dev;
    p.dst = nullptr;
    if (root == c->partRank) p.dst = d.dev;
    else
    {
        for (const auto& pm : c->peerMaps) if (pm.tex == dstTex && pm.rank == root) p.dst = (unsigned char*)pm.ptr;
    }
    if (!p.dst) { c->err = "composite: the root's texture was not imported (slcu_texture_ipc_import)"; return SLCU_ERR_ARG; }
    p.rowBytes = (unsigned long long)s.w * s.bpt;
    p.numRows = s.h;
    p.rank = c->partRank;
    p.world = c->partWorld;
    p.stripRows = c->partStripRows;
    unsigned gx = (unsigned)((p.rowBytes / 16 + 255) / 256);
    if (gx < 1) gx = 1;
    if (gx > 8) gx = 8;
    unsigned gy = s.h < 148u * 4u ? s.h : 148u * 4u;
    { KernelScope ks(c, SLCU_K_COMPOSITE); k_composite<<<dim3(gx, gy), 256, 0, c->stream>>>(p); }
    CK(cudaGetLastError());
    return 0;
}

/* ---- shared vertex stage ------------------------------------------------------------------------------------------ */
int slcu_geometry_share_create(slcu_context* c, size_t arenaBytes)
{
    SLCU_FLUSH();
    if (!c) return SLCU_ERR_ARG;
    if (c->partWorld < 2 || c->partWorld > (unsigned)SL_MAX_RANKS) { c->err = "geometry_share_create: set_partition with 2..8 ranks first"; return SLCU_ERR_ARG; }
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    geom_release(c);
    arenaBytes = (arenaBytes + 255) & ~(size_t)255;
    if (!arenaBytes) return SLCU_ERR_ARG;
    void* p = nullptr;
    if (cudaMalloc(&p, SL_GEOM_HEAD_BYTES + 2 * arenaBytes) != cudaSuccess) { cudaGetLastError(); c->err = "geometry_share_create: out of device memory"; return SLCU_ERR_OOM; }
    c->geom.base = (unsigned char*)p;
    c->geom.arenaBytes = arenaBytes;
    c->geom.peer[c->partRank] = c->geom.base;
    CK(cudaMemset(c->geom.base, 0, SL_GEOM_HEAD_BYTES));
    return 0;
}

int slcu_geometry_share_export(slcu_context* c, void* handle64)
{
    if (!c || !handle64 || !c->geom.base) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->geom.base));
    memcpy(handle64, &h, 64);
    return 0;
}

int slcu_geometry_share_import(slcu_context* c, unsigned peerRank, const void* handle64)
{
    SLCU_FLUSH();
    if (!c || !handle64 || !c->geom.base || peerRank >= c->partWorld || peerRank == c->partRank) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    if (c->geom.peer[peerRank] && c->geom.ipcMapped[peerRank]) cudaIpcCloseMemHandle(c->geom.peer[peerRank]);
    c->geom.peer[peerRank] = (unsigned char*)ptr;
    c->geom.ipcMapped[peerRank] = true;
    return 0;
}

/* peers inside one process (several contexts driven by threads of one host program): the allocation's device address */
int slcu_geometry_share_address(slcu_context* c, void** out)
{
    if (!c || !out || !c->geom.base) return SLCU_ERR_ARG;
    *out = c->geom.base;
    return 0;
}
int slcu_geometry_share_attach(slcu_context* c, unsigned peerRank, void* address)
{
    SLCU_FLUSH();
    if (!c || !address || !c->geom.base || peerRank >= c->partWorld || peerRank == c->partRank) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    cudaPointerAttributes at;
    CK(cudaPointerGetAttributes(&at, address));
    if (at.type != cudaMemoryTypeDevice) { c->err = "geometry_share_attach: not a device address"; return SLCU_ERR_ARG; }
    if (at.device != c->device)
    {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, c->device, at.device));
        if (!can) { c->err = "geometry_share_attach: no peer access between the two devices"; return SLCU_ERR_UNSUPPORTED; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { c->err = cudaGetErrorString(e); return SLCU_ERR_CUDA; }
        cudaGetLastError();
    }
    if (c->geom.peer[peerRank] && c->geom.ipcMapped[peerRank]) cudaIpcCloseMemHandle(c->geom.peer[peerRank]);
    c->geom.peer[peerRank] = (unsigned char*)address;
    c->geom.ipcMapped[peerRank] = false;
    return 0;
}

int slcu_sync(slcu_context* c)
{
    SLCU_FLUSH();
    cudaSetDevice(c->device);
    /* the join of the reference's workers; also where frames that could not be rendered like the reference are reported */
    CK(cudaMemcpyAsync(&c->hostCounters[1].errorFlags, &c->counters->errorFlags, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return report_device_flags(c, c->hostCounters[1].errorFlags);
}

int slcu_reset_stats(slcu_context* c)
{
    SLCU_FLUSH();
    if (!c) return 0;
    cudaSetDevice(c->device);
    memset(&c->stats, 0, sizeof(c->stats));
    CK(cudaMemsetAsync(c->counters, 0, sizeof(SLCounters), c->stream));
    return 0;
}

int slcu_get_stats(slcu_context* c, slcu_stats* out)
{
    SLCU_FLUSH();
    if (!c || !out) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaMemcpyAsync(c->hostCounters, c->counters, sizeof(SLCounters), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->stats.trisBinned = c->hostCounters->totalTrisBinned;
    c->stats.spans = c->hostCounters->totalSpans;
    c->stats.fragsShaded = c->hostCounters->totalFragsShaded;
    if (getenv("SLCU_DEBUG"))
        fprintf(stderr, "[slcu] resolve: windows=%llu groups=%llu waves=%llu idle polls=%llu | warps=%llu clocks sum=%llu max=%llu maxWindows=%llu\n",
                (unsigned long long)c->hostCounters->resolveWindows, (unsigned long long)c->hostCounters->resolveGroups,
                (unsigned long long)c->hostCounters->resolveWaves, (unsigned long long)c->hostCounters->resolveIdle,
                (unsigned long long)c->hostCounters->resolveWarps, (unsigned long long)c->hostCounters->resolveCycles,
                (unsigned long long)c->hostCounters->resolveMaxCycles, (unsigned long long)c->hostCounters->resolveMaxWindows);
    if (getenv("SLCU_DEBUG") && (c->hostCounters->resolveMaxWindows >> 32))
    {
        const unsigned long long w = c->hostCounters->resolveMaxWindows; /* k_resolve2: its longest ticket */
        fprintf(stderr, "[slcu] resolve2 longest ticket: %llu clocks, %llu pieces, %llu rounds\n", w >> 32, (w >> 18) & 16383ull, w & 262143ull);
    }
    *out = c->stats;
    return report_device_flags(c, c->hostCounters->errorFlags);
}

int slcu_timer_begin(slcu_context* c)
{
    SLCU_FLUSH();
    if (!c) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    if (!c->timerA) { CK(cudaEventCreate(&c->timerA)); CK(cudaEventCreate(&c->timerB)); }
    CK(cudaEventRecord(c->timerA, c->stream));
    return 0;
}

int slcu_timer_end(slcu_context* c, double* ms)
{
    SLCU_FLUSH(); /* a batch that has to be run again is part of the timed region */
    if (!c || !ms || !c->timerA) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaEventRecord(c->timerB, c->stream));
    CK(cudaEventSynchronize(c->timerB));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, c->timerA, c->timerB));
    *ms = (double)f;
    return 0;
}

int slcu_profile_begin(slcu_context* c)
{
    SLCU_FLUSH();
    if (!c) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    for (auto& e : c->profEvents) { c->eventPool.push_back(e.a); c->eventPool.push_back(e.b); }
    c->profEvents.clear();
    c->profiling = true;
    return 0;
}

int slcu_profile_end(slcu_context* c, slcu_kernel_times* out)
{
    SLCU_FLUSH();
    if (!c || !out) return SLCU_ERR_ARG;
    cudaSetDevice(c->device);
    memset(out, 0, sizeof(*out));
    c->profiling = false;
    CK(cudaStreamSynchronize(c->stream));
    for (auto& e : c->profEvents)
    {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e.a, e.b));
        out->ms[e.cls] += ms;
        out->launches[e.cls]++;
        c->eventPool.push_back(e.a);
        c->eventPool.push_back(e.b);
    }
    c->profEvents.clear();
    return 0;
}

} // extern "C"
