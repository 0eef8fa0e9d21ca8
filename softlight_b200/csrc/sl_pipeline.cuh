/*
 * sl_pipeline.cuh -- the draw path as sm_100a kernels.
 *
 *   k_vertex<VS>      SL_TriProcessor::process_verts / clip_and_process_tris /
 *                     push_bin's visibility test (src/SL_TriProcessor.cpp:248-640):
 *                     one thread per primitive; index fetch, 3x vertex shader,
 *                     cull, scissor matrix, clip status, Sutherland-Hodgman,
 *                     perspective divide, viewport; writes *triangle records*
 *                     (screen coords + varyings) in submission order per block.
 *   k_span<MODE>      SL_ScanlineBounds::init/step (SL_ScanlineBounds.hpp:137-347)
 *                     for every (record,row): counts / fills the per-row span
 *                     lists in exact submission order.  Rows are the bins: a
 *                     full-width row is the only tile shape for which the
 *                     reference's left-to-right compaction state machine
 *                     (SURVEY.md section 0 finding 6) stays tile-local.
 *   k_resolve         SL_TriRasterizer::render_triangle_simd (src/SL_TriRasterizer.cpp:
 *                     413-559) + the depth write of flush_tri_fragments: one warp
 *                     per (row, x-segment) walks the row's spans in order, lanes
 *                     take spans whose x-ranges do not overlap; output = ordered
 *                     fragment lists in HBM.
 *   k_shade<FS>       SL_FragmentProcessor::flush_tri_fragments + assign_pixel /
 *                     assign_alpha_pixel (src/SL_FragmentProcessor.cpp:416-531,
 *                     src/SL_Framebuffer.cpp:44-368): one warp per fragment list,
 *                     32 fragments per step, converged shader calls.
 *
 * Data layout in HBM:
 *   record r      : float4[3 + 3*NV]  = p0,p1,p2 (x,y,z,1/w) then varyings [slot][vert].  (A plane-major layout was
 *                   measured too: it makes the vertex stage's writes and the span stage's reads coalesced, but the
 *                   rasteriser's per-piece gathers then touch twice the sectors and the frame got slower.)
 *   span entry    : {u32 record, u16 xMin, u16 xMax}, grouped by row, ordered
 *   cnt[vb][y]    : u16 spans of vertex-block vb on row y;  off[vb][y] : u32 exclusive prefix over vb
 */
#pragma once

#include "sl_shader_api.cuh"
#include "sl_texel.cuh"

namespace sld
{

typedef math::vec4 (*VsFn)(SL_VertexParam&);
typedef bool (*FsFn)(SL_FragmentParam&);

constexpr int SL_VB_THREADS = 256;        /* threads per vertex block                       */
constexpr int SL_VB_ROUNDS = 4;           /* primitives per thread                          */
constexpr int SL_VB_PRIMS = SL_VB_THREADS * SL_VB_ROUNDS;
constexpr int SL_MAX_ROWS = 8192;         /* framebuffer height limit of the row binner     */
constexpr int SL_SPAN_WIN = 256;          /* rows per mask window in k_span                 */
constexpr int SL_SHADE_SHIFT_C = 7;       /* log2 of the pixels per fragment list (= SL_SHADE_SHIFT, declared with the lists) */

/* rows: first | one-past-last << 16 of the rows the segment's records can have span entries in (0xffff = none).  Every segment
 * is written exactly once per batch by the warp that owns it -- empty ones too -- so the table needs no clearing and can be
 * filled by several ranks at once (k_vertex) */
struct SLSegment { uint32_t base, count, rows; };
constexpr int SL_VB_SEGS = SL_VB_ROUNDS * (SL_VB_THREADS / 32); /* one record segment per (round, warp) of a vertex block */
struct SLBlockInfo { SLSegment seg[SL_VB_SEGS]; };
constexpr int SL_MAX_RANKS = 8;       /* ranks that can share the vertex stage (one NVSwitch domain) */
constexpr int SL_SPAN_BAND_SHIFT = 8; /* log2(SL_SPAN_WIN): rows per work item of k_span */

struct SLSpanEntry { uint32_t rec; uint16_t xMin, xMax; };

/* device-side counters of one context */
struct SLCounters
{
    /* ---- of the running batch: zeroed by one memset when the batch is enqueued (SL_COUNTERS_BATCH_BYTES) ---- */
    unsigned long long recAlloc;      /* triangle records taken from the overflow area                                  */
    unsigned long long handoffs;      /* span cuts (segment boundaries + row ends) that need a state hand-over          */
    unsigned long long pixelsCovered; /* sum of span lengths = upper bound of the batch's fragments                     */
    unsigned long long rowEndCuts;    /* spans that run past the row end into an existing row                           */
    unsigned long long trisBinned, spans, fragsShaded; /* statistics of the batch; k_commit adds them to the totals     */
    unsigned int abortBatch;          /* SL_ABORT_*: the batch does not fit the buffers it was given.  Set before anything
                                         visible is written; k_span<1>, k_resolve*, k_shade then return at once and the host
                                         runs the batch again with larger buffers (slcu.cu: verify_inflight)            */
    unsigned int rasterTicket;        /* (row, segment) ticket of k_resolve (reset by k_scan_row_starts)                */
    unsigned int chunkAlloc;          /* fragment chunks handed out by k_resolve (reset by k_scan_row_starts)           */
    unsigned int shadeTicket;         /* list ticket of k_shade (reset by k_scan_row_starts)                            */
    unsigned int shadeBandTicket[8];  /* ... of its later launches when the lists are shaded band by band (read-back overlap) */
    unsigned int totalEntries;        /* entries of the row lists (= rowStart[numRows]), for the host's sizing           */
    unsigned int spanTicket[2];       /* item tickets of k_span<0> / k_span<1>                                          */
    unsigned int spanItems;           /* work items of k_span (k_span_items)                                            */
    /* ---- since reset_stats ---- */
    unsigned long long totalTrisBinned, totalSpans, totalFragsShaded;
    unsigned int errorFlags;          /* SL_ERR_*: sticky until the host reports them                                   */
    unsigned int pad_;
    /* k_resolve diagnostics (SLCU_DEBUG=1): windows, their groups and waves, polls that found nothing, per-warp clocks */
    unsigned long long resolveWindows, resolveWaves, resolveIdle, resolveGroups;
    unsigned long long resolveCycles, resolveMaxCycles, resolveMaxWindows, resolveWarps;
};
#define SL_COUNTERS_BATCH_BYTES offsetof(SLCounters, totalTrisBinned)

/* abortBatch bits */
constexpr unsigned SL_ABORT_RECORDS = 1u;  /* more triangle records than the overflow area holds (clipping fan-out)      */
constexpr unsigned SL_ABORT_ENTRIES = 2u;  /* more span entries than the row lists hold                                  */
constexpr unsigned SL_ABORT_CHUNKS = 4u;   /* the fragment pool could run out                                            */
constexpr unsigned SL_ABORT_HANDOFFS = 8u; /* more hand-overs than the table holds                                       */
constexpr unsigned SL_ABORT_PARTITION = 16u; /* sort-first fast path: a span runs past the row end into another rank's row */
constexpr unsigned SL_ABORT_GEOMETRY = 32u;  /* shared vertex stage: some rank's record area overflowed; every rank runs the batch again on its own */
/* errorFlags bits: one per condition */
constexpr unsigned SL_ERR_SPAN_TWO_ROWS = 0x2u;     /* a span longer than two rows: not reproduced                        */
constexpr unsigned SL_ERR_SPAN_NEGATIVE_X = 0x4u;   /* a span that starts left of x = 0: not reproduced                   */
constexpr unsigned SL_ERR_UNEXPECTED_OVERFLOW = 0x8u; /* row overflow of a triangle classified as "cannot overflow"       */
constexpr unsigned SL_ERR_HANDOFF_PUBLISH = 0x10u;  /* a cut the count pass did not see (no hand-over table)              */
constexpr unsigned SL_ERR_POOL = 0x20u;             /* fragment chunk pool exhausted: fragments dropped                   */
constexpr unsigned SL_ERR_HANDOFF_POLL = 0x40u;     /* a piece waits for a hand-over but there is no table                */
constexpr unsigned SL_ERR_WIRE_SEGMENT = 0x80u;     /* wireframe: a record segment with more than 256 records             */
constexpr unsigned SL_ERR_WIRE_SLOTS = 0x100u;      /* wireframe: candidate slots exhausted                               */
constexpr unsigned SL_ERR_GEOM_TIMEOUT = 0x200u;    /* shared vertex stage: a rank never signalled its blocks             */

/* A span whose xMax exceeds the framebuffer width keeps going in linear memory:
 * the reference addresses pixels as x + width*y (src/SL_TriRasterizer.cpp:472,
 * src/SL_FragmentProcessor.cpp:507-529), so those fragments land at the start of
 * row y+1 (clip-generated vertices do end up right of the viewport because the
 * clipper's t comes from rcpps).  Row y+1 is owned by another warp, so such a
 * span is cut at the row end: the in-row part stays in row y's list, the rest
 * becomes an "overflow entry" in row y+1's list (same position in submission
 * order), and the walker state at the cut (lane map of the compaction state
 * machine + depth outcomes of the straddling group's in-row lanes) is handed
 * over through a small hash table in global memory. */
constexpr uint32_t SL_ENTRY_OVERFLOW = 0x80000000u;

/* One submitted draw of a batch (one mesh of draw / draw_multiple, or one draw_instanced): everything that may differ
 * between the draws of a batch.  The table lives in HBM; a batch holds consecutive draws that share the shader object
 * and the framebuffer (SL_Shader state and attachments are batch constants).  Uniforms are snapshotted at submission,
 * exactly what the reference's synchronous draw would have read. */
struct alignas(16) SLDrawDesc
{
    SL_UniformBuffer uniforms;
    alignas(16) float scissor[16]; /* column-major scissor matrix (host, SL_ViewportState::scissor_matrix); read as a mat4 */
    alignas(16) float viewport[4]; /* viewport_rect */
    SL_VertexArray vao;
    SL_VertexBuffer vbo;
    const unsigned char* ibo;
    int iboType;          /* SL_DataType: 0 u8, 1 u16, 2 u32 */
    int usingIndices;
    unsigned long long elementBegin;
    unsigned long long primsPerInstance;
    unsigned long long numPrims;      /* primsPerInstance * instances */
    unsigned int firstBlock;          /* index of this draw's first vertex block in the batch */
    unsigned int blocksPerInstance;   /* vertex blocks never straddle instances: ceil(primsPerInstance / SL_VB_PRIMS) */
};

/* where the vertex stage's output goes: this context's buffers, or -- when ranks share the stage -- rank t's arena */
struct SLGeomOut
{
    float4* records;
    uint16_t* recDraw;                /* record slot -> index into draws (read by the rasteriser); SL_REC_ABSENT: not routed here */
    SLBlockInfo* blockInfo;
};

struct SLVertexParams
{
    const SLDrawDesc* draws;          /* draw table of the batch (HBM)                                  */
    const uint32_t* blockDraw;        /* vertex block -> index into draws                               */
    int cullMode;                     /* SL_CullMode */
    int numVaryings;
    unsigned int cacheBytes;          /* dynamic shared memory of the post-transform cache (0: off) */
    unsigned int cacheVerts;          /* vertices the cache holds (<= SL_VCACHE_VERTS, multiple of 32)  */
    unsigned int recOverflowBase;     /* first slot of this rank's overflow area                        */
    unsigned int recOverflowCapacity; /* its size                                                       */
    SLCounters* counters;
    int shardRank, shardWorld;        /* ranks sharing the stage: this one runs blocks rank, rank + world, ... (world 1: all) */
    int partStrip;                    /* rows per ownership strip of the sort-first partition           */
    SLGeomOut out;                    /* this rank's buffers (its arena when the stage is shared)       */
    uint8_t* recMask;                 /* shared stage: record slot -> ranks whose rows the record touches (k_geom_route sends it there) */
};

/* ------------------------------------------------------------------------
 * Vertex stage helpers
 * ----------------------------------------------------------------------*/
struct SLVert
{
    math::vec4 vert;
    math::vec4 varyings[4];
};

/* face_determinant, SSE4.1 branch (src/SL_TriProcessor.cpp:40-66) */
SLD_FN float sl_face_determinant(const math::vec4& p0, const math::vec4& p1, const math::vec4& p2)
{
    const float m0 = ((p1.v[1] * p2.v[3]) - (p1.v[3] * p2.v[1])) * p0.v[0];
    const float m1 = ((p1.v[3] * p2.v[0]) - (p1.v[0] * p2.v[3])) * p0.v[1];
    const float m2 = ((p1.v[0] * p2.v[1]) - (p1.v[1] * p2.v[0])) * p0.v[3];
    return m2 + (m1 + m0);
}

/* sl_ndc_clip_status (SL_PostVertexTransform.hpp:590-616): 3 fully, 1 partially, 0 not visible */
SLD_FN int sl_clip_status(const math::vec4& c0, const math::vec4& c1, const math::vec4& c2)
{
    bool vis = true, part = false;
    const math::vec4* cs[3] = {&c0, &c1, &c2};
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        const float w = cs[i]->v[3];
        const float nw = __uint_as_float(__float_as_uint(w) | 0x80000000u);
#pragma unroll
        for (int k = 0; k < 4; ++k) vis = vis && (nw <= cs[i]->v[k]) && (w >= cs[i]->v[k]);
        part = part || (w > 0.f);
    }
    return (vis ? 3 : 0) | (part ? 1 : 0);
}

/* sl_perspective_divide + sl_ndc_to_screen_coords (SL_PostVertexTransform.hpp:160-178, 412-449) */
SLD_FN void sl_to_screen(math::vec4& v, const float* dims)
{
    const float wInv = sl_rcp(v.v[3]);
    const float nx = v.v[0] * wInv, ny = v.v[1] * wInv, nz = v.v[2] * wInv;
    float sx = sl_fma(1.f + nx, dims[2] * 0.5f, dims[0]);
    float sy = sl_fma(1.f + ny, dims[3] * 0.5f, dims[1]);
    sx = sl_max(floorf(sx), 0.f);
    sy = sl_max(floorf(sy), 0.f);
    v = math::vec4{sx, sy, nz, wInv};
}

/* push_bin's visibility test (src/SL_TriProcessor.cpp:257-264) */
SLD_FN bool sl_prim_hidden(const math::vec4& p0, const math::vec4& p1, const math::vec4& p2)
{
    const float minX = sl_min(sl_min(p0.v[0], p1.v[0]), p2.v[0]);
    const float maxX = sl_max(sl_max(p0.v[0], p1.v[0]), p2.v[0]);
    const float minY = sl_min(sl_min(p0.v[1], p1.v[1]), p2.v[1]);
    const float maxY = sl_max(sl_max(p0.v[1], p1.v[1]), p2.v[1]);
    return (maxX - minX < 1.f) || (maxY - minY < 1.f);
}

SLD_FN void sl_write_record(float4* rec, int numVaryings, const math::vec4& p0, const math::vec4& p1, const math::vec4& p2,
                            const math::vec4* va, const math::vec4* vb, const math::vec4* vc)
{
    rec[0] = *reinterpret_cast<const float4*>(&p0);
    rec[1] = *reinterpret_cast<const float4*>(&p1);
    rec[2] = *reinterpret_cast<const float4*>(&p2);
#pragma unroll
    for (int s = 0; s < 4; ++s) /* static indices: the shader outputs stay in registers */
    {
        if (s < numVaryings)
        {
            rec[3 + 3 * s + 0] = *reinterpret_cast<const float4*>(&va[s]);
            rec[3 + 3 * s + 1] = *reinterpret_cast<const float4*>(&vb[s]);
            rec[3 + 3 * s + 2] = *reinterpret_cast<const float4*>(&vc[s]);
        }
    }
}

/* SL_TriProcessor::clip_and_process_tris (src/SL_TriProcessor.cpp:370-519).  Rare
 * path: the polygon lives in local memory.  Returns the number of vertices (0
 * when nothing is left); verts are already in screen space. */
struct SLClipPoly
{
    math::vec4 verts[9];
    math::vec4 varys[9 * 4];
};

/* bufs[0] holds the triangle; the polygon alternates between the two buffers, one clip plane at a time, and `which` says
 * where it ended up.  A plane that has every vertex on its visible side leaves the polygon as it is (the reference copies
 * it vertex by vertex: same values). */
__device__ __noinline__ int sl_clip_polygon(SLClipPoly* bufs, int numVarys, const float* dims, int& which)
{
    int src = 0;
    int numTotalVerts = 3;
    const float edges[6][4] = {{1.f, 0.f, 0.f, 1.f}, {-1.f, 0.f, 0.f, 1.f}, {0.f, 1.f, 0.f, 1.f},
                               {0.f, -1.f, 0.f, 1.f}, {0.f, 0.f, 1.f, 1.f}, {0.f, 0.f, -1.f, 1.f}};
    for (int e = 0; e < 6; ++e)
    {
        const math::vec4 edge{edges[e][0], edges[e][1], edges[e][2], edges[e][3]};
        const SLClipPoly& poly = bufs[src];
        SLClipPoly& tmp = bufs[src ^ 1];
        bool allVisible = true;
        for (int k = 0; k < numTotalVerts; ++k) allVisible = allVisible && !sl_signbit(math::dot(poly.verts[k], edge));
        if (allVisible) continue;
        int numNewVerts = 0;
        int j = numTotalVerts - 1;
        math::vec4 p0 = poly.verts[numTotalVerts - 1];
        float t0 = math::dot(p0, edge);
        int visible0 = !sl_signbit(t0);
        for (int k = 0; k < numTotalVerts; ++k)
        {
            const math::vec4 p1 = poly.verts[k];
            const float t1 = math::dot(p1, edge);
            const int visible1 = !sl_signbit(t1);
            if (visible0 ^ visible1)
            {
                const float t = t0 * sl_rcp(t0 - t1);
                tmp.verts[numNewVerts] = math::mix(p0, p1, t);
                for (int i = 0; i < numVarys; ++i) tmp.varys[numNewVerts * 4 + i] = math::mix(poly.varys[j * 4 + i], poly.varys[k * 4 + i], t);
                ++numNewVerts;
            }
            if (visible1)
            {
                tmp.verts[numNewVerts] = p1;
                for (int i = 0; i < numVarys; ++i) tmp.varys[numNewVerts * 4 + i] = poly.varys[k * 4 + i];
                ++numNewVerts;
            }
            j = k;
            p0 = p1;
            t0 = t1;
            visible0 = visible1;
        }
        if (!numNewVerts) return 0;
        numTotalVerts = numNewVerts;
        src ^= 1;
    }
    which = src;
    if (numTotalVerts < 3) return 0;
    for (int i = 0; i < numTotalVerts; ++i) sl_to_screen(bufs[src].verts[i], dims);
    return numTotalVerts;
}

/* Exact shortcut for the clipper above.  Sutherland-Hodgman visits the planes in a fixed order; as long as every
 * vertex was on the visible side of the planes visited so far, the polygon is still the input triangle.  If at that
 * point all three vertices are on the invisible side of the next plane, the clipper would end with zero vertices
 * (src/SL_TriProcessor.cpp:455-459) -- same dot products, same sign tests, no interpolation involved.  The reference
 * sends every triangle that is outside the frustum but in front of the eye through the clipper (status "partially
 * visible" = any w > 0, SL_PostVertexTransform.hpp:598-616); most of them end here. */
SLD_FN bool sl_clip_rejects(const math::vec4& p0, const math::vec4& p1, const math::vec4& p2)
{
    const float edges[6][4] = {{1.f, 0.f, 0.f, 1.f}, {-1.f, 0.f, 0.f, 1.f}, {0.f, 1.f, 0.f, 1.f},
                               {0.f, -1.f, 0.f, 1.f}, {0.f, 0.f, 1.f, 1.f}, {0.f, 0.f, -1.f, 1.f}};
#pragma unroll
    for (int e = 0; e < 6; ++e)
    {
        const math::vec4 edge{edges[e][0], edges[e][1], edges[e][2], edges[e][3]};
        const int vis = (sl_signbit(math::dot(p0, edge)) ? 0 : 1) + (sl_signbit(math::dot(p1, edge)) ? 0 : 1) + (sl_signbit(math::dot(p2, edge)) ? 0 : 1);
        if (vis == 0) return true;
        if (vis != 3) return false;
    }
    return false;
}

/* ------------------------------------------------------------------------
 * k_vertex
 * ----------------------------------------------------------------------*/
constexpr int SL_VCACHE_VERTS = 1024; /* post-transform vertex cache of one vertex block (shared memory) */

/* get_next_vertex3 (src/SL_TriProcessor.cpp:125-234) */
SLD_FN void sl_fetch_vertex_ids(const SLDrawDesc& P, unsigned long long i, unsigned long long vertId[3])
{
    if (P.usingIndices)
    {
        if (P.iboType == 0) { const uint8_t* p = P.ibo + i; vertId[0] = p[0]; vertId[1] = p[1]; vertId[2] = p[2]; }
        else if (P.iboType == 1) { const uint16_t* p = reinterpret_cast<const uint16_t*>(P.ibo) + i; vertId[0] = p[0]; vertId[1] = p[1]; vertId[2] = p[2]; }
        else { const uint32_t* p = reinterpret_cast<const uint32_t*>(P.ibo) + i; vertId[0] = p[0]; vertId[1] = p[1]; vertId[2] = p[2]; }
    }
    else { vertId[0] = i; vertId[1] = i + 1; vertId[2] = i + 2; }
}

/* Ranks whose rows a record can have span entries in (sort-first partition, interleaved strips): the rows of its
 * vertices and one more, for a span that runs past the row end (k_span: y0 .. yEnd).  Screen y is a non-negative
 * integer stored as a float (sl_to_screen); anything else (NaN, huge) goes to every rank. */
SLD_FN uint32_t sl_rank_mask(float ya, float yb, float yc, int strip, int world)
{
    const uint32_t all = (1u << world) - 1u;
    const float lo = fminf(ya, fminf(yb, yc)), hi = fmaxf(ya, fmaxf(yb, yc));
    if (!(ya == ya) || !(yb == yb) || !(yc == yc) || !(hi < 65000.f)) return all;
    const int sa = (int)lo / strip, sb = ((int)hi + 1) / strip;
    if (sb - sa + 1 >= world) return all;
    uint32_t m = 0u;
    for (int s = sa; s <= sb; ++s) m |= 1u << (s % world);
    return m;
}

/* The reference runs the vertex shader three times per triangle with no post-transform cache
 * (src/SL_TriProcessor.cpp:584-594; SL_VertexCache is compiled out, SL_Config.hpp:10-12).  A vertex shader is a pure
 * function of (uniforms, vertId, instanceId), so shading each vertex ONCE per (instance, vertex block) gives the same bits.
 *
 * A vertex block is 1024 consecutive primitives of ONE instance of one draw (the last block of an instance is short), 256
 * threads, four rounds of one primitive per thread.  Two instantiations:
 *   CACHE = false  every primitive shades its own three vertices, like the reference.
 *   CACHE = true   for batches whose index buffers address at most SL_VCACHE_VERTS vertices (small instanced meshes: C2's
 *                  sphere is 561 vertices for 1024 triangles).  The block stages its indices in shared memory, shades every
 *                  vertex they use once -- together with everything after the shader that is a function of the vertex
 *                  alone: scissor transform, clip-status terms, perspective divide, viewport -- and assembles its triangles
 *                  from shared memory.  No per-primitive shader state lives in registers, so twice the blocks fit an SM.
 *
 * Records go to fixed places: every (block, round, warp) owns 32 record slots, enough unless clipping fanned primitives
 * out; only then does the warp take its slots from the overflow area (one atomic).  No barrier, no allocation-order
 * dependence, and slot order = submission order.
 *
 * Ranks that share the vertex stage (slcu_geometry_share_*): rank r runs blocks r, r + world, ... into its own arena and
 * notes for every record which ranks' rows it touches; k_geom_route then copies each record into the arenas of those ranks
 * (peer stores over NVLink, whole records per warp instruction), the 2-byte recDraw word -- 0xffff = "not for you" -- and
 * the segment table into every rank's.  Slots are the same everywhere, so nothing has to be merged. */
#ifndef SL_VERTEX_MIN_BLOCKS
#define SL_VERTEX_MIN_BLOCKS 2 /* (256, 1) lets ptxas take > 128 registers: one block per SM, measured 35 % slower */
#endif
#ifndef SL_VERTEX_CACHED_MIN_BLOCKS
#define SL_VERTEX_CACHED_MIN_BLOCKS 3
#endif
constexpr uint16_t SL_REC_ABSENT = 0xffffu; /* recDraw: the record was not routed to this rank */

template <VsFn VS, bool CACHE>
__global__ void __launch_bounds__(SL_VB_THREADS, CACHE ? SL_VERTEX_CACHED_MIN_BLOCKS : SL_VERTEX_MIN_BLOCKS) k_vertex(const __grid_constant__ SLVertexParams P)
{
    /* CACHE: u16 indices [3 * SL_VB_PRIMS] | float4 [2 + NV][cacheVerts] (clip position, screen position, varyings) | u32 used bits | u8 flags */
    extern __shared__ __align__(16) unsigned char sVCache[];
    uint32_t binned = 0u; /* lane 0: records of this warp */

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int world = P.shardWorld;
    const unsigned vb = world > 1 ? blockIdx.x * (unsigned)world + (unsigned)P.shardRank : blockIdx.x;
    const uint32_t drawIdx = P.blockDraw[vb];
    const SLDrawDesc& D = P.draws[drawIdx];
    const unsigned lb = vb - D.firstBlock;
    const unsigned instanceId = lb / D.blocksPerInstance;
    const unsigned primBase = (lb - instanceId * D.blocksPerInstance) * (unsigned)SL_VB_PRIMS; /* first primitive of the block inside its instance */
    const unsigned primsPerInstance = (unsigned)D.primsPerInstance;
    const int NV = P.numVaryings;
    const int recStride = 3 + 3 * NV;

    SL_VertexParam params;
    params.pUniforms = &D.uniforms;
    params.pVao = &D.vao;
    params.pVbo = &D.vbo;
    params.instanceId = (size_t)instanceId;

    const math::mat4 scissorMat = *reinterpret_cast<const math::mat4*>(D.scissor);
    const float viewport[4] = {D.viewport[0], D.viewport[1], D.viewport[2], D.viewport[3]};

    /* ---- post-transform cache ---- */
    const unsigned CV = P.cacheVerts;
    uint16_t* sIdx = reinterpret_cast<uint16_t*>(sVCache);
    float4* cache = reinterpret_cast<float4*>(sVCache + 3 * SL_VB_PRIMS * sizeof(uint16_t));
    uint32_t* sUsed = reinterpret_cast<uint32_t*>(cache + (size_t)(2 + NV) * CV);
    uint8_t* cacheFlags = reinterpret_cast<uint8_t*>(sUsed + CV / 32u);
    if constexpr (CACHE)
    {
        for (unsigned w = tid; w < CV / 32u; w += SL_VB_THREADS) sUsed[w] = 0u;
        __syncthreads();
#pragma unroll
        for (int round = 0; round < SL_VB_ROUNDS; ++round)
        {
            const unsigned p = (unsigned)round * SL_VB_THREADS + (unsigned)tid;
            if (primBase + p < primsPerInstance)
            {
                unsigned long long id[3];
                sl_fetch_vertex_ids(D, D.elementBegin + 3ull * (primBase + p), id);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                {
                    const unsigned v = min((unsigned)id[k], CV - 1u); /* the host checked the index buffer: never clamps */
                    sIdx[3u * p + k] = (uint16_t)v;
                    atomicOr(&sUsed[v >> 5], 1u << (v & 31u));
                }
            }
        }
        __syncthreads();
        for (unsigned v = tid; v < CV; v += SL_VB_THREADS)
        {
            if (!((sUsed[v >> 5] >> (v & 31u)) & 1u)) continue;
            math::vec4 vary[4];
            params.vertId = (size_t)v;
            params.pVaryings = vary;
            const math::vec4 pos = VS(params);
            cache[v] = make_float4(pos.v[0], pos.v[1], pos.v[2], pos.v[3]);
            /* everything after the vertex shader that is a function of the vertex alone: scissor transform (:615-617),
             * its share of sl_ndc_clip_status, perspective divide + viewport (valid when the triangle turns out fully
             * visible) */
            math::vec4 sc = scissorMat * pos;
            const float w = sc.v[3];
            const float nw = __uint_as_float(__float_as_uint(w) | 0x80000000u);
            bool vis = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) vis = vis && (nw <= sc.v[k]) && (w >= sc.v[k]);
            cacheFlags[v] = (uint8_t)((vis ? 1u : 0u) | ((w > 0.f) ? 2u : 0u));
            sl_to_screen(sc, viewport);
            cache[CV + v] = make_float4(sc.v[0], sc.v[1], sc.v[2], sc.v[3]);
#pragma unroll
            for (int s = 0; s < 4; ++s)
            {
                if (s < NV) cache[(2 + s) * CV + v] = make_float4(vary[s].v[0], vary[s].v[1], vary[s].v[2], vary[s].v[3]);
            }
        }
        __syncthreads();
    }

    /* Four passes over the block's primitives, the last three only when some primitive needs the clipper:
     *   0  the rounds: classify every primitive (culled / hidden / one record / needs the clipper).  A (round, warp) without
     *      clipper candidates writes its records at once; one with candidates is deferred as a whole and the candidates go
     *      into a queue.
     *   1  the queue, one candidate per thread: clip and COUNT the triangles of the fan (and the rows they touch).
     *   2  the deferred (round, warp)s: now every count is known; slots are assigned, the unclipped primitives' records
     *      written, the candidates told where theirs go.
     *   3  the queue again: clip and WRITE.
     * The clipper is a long serial routine over a polygon in local memory.  Run where it is met, it keeps one lane of a
     * warp busy and 31 waiting (profiles/r02q_k_vertex_c2.txt: a third of the kernel's issue slots on C2, where a few per
     * cent of the triangles cross the frustum); run from the queue every lane of the warp clips.  The candidates' vertices
     * are shaded again from their primitive id in passes 1-3 -- they are few. */
    __shared__ uint8_t sN[SL_VB_PRIMS];       /* queue candidates: triangles of the clipped fan                          */
    __shared__ uint32_t sSlot[SL_VB_PRIMS];   /* queue candidates: pass 1 -> rows (first | last + 1 << 16), pass 2 -> first slot (~0: none) */
    __shared__ uint16_t sClipQ[SL_VB_PRIMS];
    __shared__ uint8_t sDefer[SL_VB_SEGS];
    __shared__ unsigned sClipCount;
    if (tid == 0) sClipCount = 0u;
    if (tid < SL_VB_SEGS) sDefer[tid] = 0;
    __syncthreads();

    unsigned clipCount = 0u;
    SLClipPoly poly[2];  /* the clipper's two buffers (local memory); a queue of at most one candidate per thread keeps its */
    int keptVerts = -1;  /* polygon from pass 1 to pass 3 instead of clipping twice                                       */
    int keptWhich = 0;
    for (int pass = 0; pass < 4; ++pass)
    {
        const bool queuePass = (pass & 1) != 0;
        const unsigned iters = queuePass ? (clipCount + SL_VB_THREADS - 1u) / SL_VB_THREADS : (unsigned)SL_VB_ROUNDS;
        for (unsigned it = 0; it < iters; ++it)
        {
            if (pass == 2 && !sDefer[it * (SL_VB_THREADS / 32) + warp]) continue; /* warp-uniform */
            unsigned p;
            bool active;
            if (queuePass)
            {
                const unsigned q = it * SL_VB_THREADS + (unsigned)tid;
                active = q < clipCount;
                p = active ? sClipQ[q] : 0u;
            }
            else
            {
                p = it * SL_VB_THREADS + (unsigned)tid;
                active = primBase + p < primsPerInstance;
            }
            SLVert v[3];
            int n = 0;             /* records this thread emits */
            int polyVerts = 0;     /* > 0: clipped path, polygon in poly[which] */
            int which = 0;
            bool needClip = false; /* partially visible and not rejected by the first plane it misses: the clipper decides */
            unsigned ci[3] = {0u, 0u, 0u}; /* cache slots of the three vertices */

            if (active)
            {
                if constexpr (CACHE)
                {
                    ci[0] = sIdx[3u * p + 0];
                    ci[1] = sIdx[3u * p + 1];
                    ci[2] = sIdx[3u * p + 2];
                    math::vec4 cp[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                    {
                        const float4 pos = cache[ci[k]];
                        cp[k] = math::vec4{pos.x, pos.y, pos.z, pos.w};
                    }
                    bool culled = false;
                    if (P.cullMode != 2) /* :597-612 */
                    {
                        const float det = sl_face_determinant(cp[0], cp[1], cp[2]);
                        culled = (P.cullMode == 1) ^ sl_signbit(det);
                    }
                    if (!culled)
                    {
                        const unsigned f0 = cacheFlags[ci[0]], f1 = cacheFlags[ci[1]], f2 = cacheFlags[ci[2]];
                        if (f0 & f1 & f2 & 1u) /* fully visible */
                        {
                            const float4 s0 = cache[CV + ci[0]], s1 = cache[CV + ci[1]], s2 = cache[CV + ci[2]];
                            n = sl_prim_hidden(math::vec4{s0.x, s0.y, s0.z, s0.w}, math::vec4{s1.x, s1.y, s1.z, s1.w}, math::vec4{s2.x, s2.y, s2.z, s2.w}) ? 0 : 1;
                        }
                        else if ((f0 | f1 | f2) & 2u) /* partially visible: the clipper works on the scissored clip-space vertices */
                        {
#pragma unroll
                            for (int k = 0; k < 3; ++k) cp[k] = scissorMat * cp[k];
                            needClip = !sl_clip_rejects(cp[0], cp[1], cp[2]);
                            if (needClip && queuePass && !(pass == 3 && keptVerts >= 0))
                            {
#pragma unroll
                                for (int k = 0; k < 3; ++k)
                                {
                                    poly[0].verts[k] = cp[k];
#pragma unroll
                                    for (int s = 0; s < 4; ++s)
                                    {
                                        if (s < NV)
                                        {
                                            const float4 f = cache[(2 + s) * CV + ci[k]];
                                            poly[0].varys[k * 4 + s] = math::vec4{f.x, f.y, f.z, f.w};
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
                else
                {
                    unsigned long long vertId[3];
                    sl_fetch_vertex_ids(D, D.elementBegin + 3ull * (primBase + p), vertId);
#pragma unroll
                    for (int k = 0; k < 3; ++k) /* src/SL_TriProcessor.cpp:584-594 */
                    {
                        params.vertId = (size_t)vertId[k];
                        params.pVaryings = v[k].varyings;
                        v[k].vert = VS(params);
                    }

                    bool culled = false;
                    if (P.cullMode != 2) /* :597-612 */
                    {
                        const float det = sl_face_determinant(v[0].vert, v[1].vert, v[2].vert);
                        culled = (P.cullMode == 1) ^ sl_signbit(det);
                    }
                    if (!culled)
                    {
#pragma unroll
                        for (int k = 0; k < 3; ++k) v[k].vert = scissorMat * v[k].vert; /* :615-617 */
                        const int vis = sl_clip_status(v[0].vert, v[1].vert, v[2].vert);
                        if (vis == 3)
                        {
#pragma unroll
                            for (int k = 0; k < 3; ++k) sl_to_screen(v[k].vert, viewport);
                            n = sl_prim_hidden(v[0].vert, v[1].vert, v[2].vert) ? 0 : 1;
                        }
                        else if (vis == 1)
                        {
                            needClip = !sl_clip_rejects(v[0].vert, v[1].vert, v[2].vert);
                            if (needClip && queuePass && !(pass == 3 && keptVerts >= 0))
                            {
#pragma unroll
                                for (int k = 0; k < 3; ++k)
                                {
                                    poly[0].verts[k] = v[k].vert;
#pragma unroll
                                    for (int s = 0; s < 4; ++s) { if (s < NV) poly[0].varys[k * 4 + s] = v[k].varyings[s]; }
                                }
                            }
                        }
                    }
                }
            }

            if (queuePass)
            {
                /* every active lane is a candidate (it was one in pass 0: same inputs, same arithmetic) */
                if (needClip)
                {
                    if (pass == 3 && keptVerts >= 0) { polyVerts = keptVerts; which = keptWhich; }
                    else polyVerts = sl_clip_polygon(poly, NV, viewport, which);
                    if (pass == 1 && clipCount <= (unsigned)SL_VB_THREADS) { keptVerts = polyVerts; keptWhich = which; }
                    const SLClipPoly& pg = poly[which];
                    if (pass == 1)
                    {
                        float fy0 = 3.0e38f, fy1 = -1.f;
                        for (int f = polyVerts - 2; f-- > 0;)
                        {
                            if (!sl_prim_hidden(pg.verts[0], pg.verts[f + 1], pg.verts[f + 2])) ++n;
                        }
                        for (int k = 0; k < polyVerts; ++k) { fy0 = fminf(fy0, pg.verts[k].v[1]); fy1 = fmaxf(fy1, pg.verts[k].v[1]); }
                        uint32_t rows = 0xffffu;
                        if (n && fy1 >= 0.f)
                        {
                            const uint32_t r0 = fy0 < 65535.f ? (uint32_t)fy0 : 65535u;
                            const uint32_t r1 = fy1 < 65533.f ? (uint32_t)fy1 + 2u : 65535u;
                            rows = r0 | (r1 << 16);
                        }
                        sN[p] = (uint8_t)n;
                        sSlot[p] = rows;
                    }
                    else
                    {
                        uint32_t slot = sSlot[p];
                        if (slot != 0xffffffffu)
                        {
                            for (int f = polyVerts - 2; f-- > 0;) /* reverse fan order (:505-518) */
                            {
                                const int j = f + 1, k = f + 2;
                                if (sl_prim_hidden(pg.verts[0], pg.verts[j], pg.verts[k])) continue;
                                uint16_t tag = (uint16_t)drawIdx;
                                if (world > 1)
                                {
                                    const uint32_t mask = sl_rank_mask(pg.verts[0].v[1], pg.verts[j].v[1], pg.verts[k].v[1], P.partStrip, world);
                                    P.recMask[slot] = (uint8_t)mask;
                                    if (!((mask >> P.shardRank) & 1u)) tag = SL_REC_ABSENT;
                                }
                                P.out.recDraw[slot] = tag;
                                sl_write_record(P.out.records + (size_t)slot * recStride, NV, pg.verts[0], pg.verts[j], pg.verts[k],
                                                &pg.varys[0], &pg.varys[j * 4], &pg.varys[k * 4]);
                                ++slot;
                            }
                        }
                    }
                }
                else if (active && pass == 1) { sN[p] = 0; sSlot[p] = 0xffffu; } /* cannot happen: pass 0 queued it on the same arithmetic */
                continue;
            }

            if (pass == 0)
            {
                if (__any_sync(0xffffffffu, needClip)) /* the whole (round, warp) waits for its candidates' counts */
                {
                    if (needClip) sClipQ[atomicAdd(&sClipCount, 1u)] = (uint16_t)p;
                    if (lane == 0) sDefer[it * (SL_VB_THREADS / 32) + warp] = 1;
                    continue;
                }
            }
            else if (needClip) n = sN[p];

            /* ordered compaction inside the warp: exclusive scan of n (one ballot unless clipping fanned a primitive out) */
            uint32_t incl, total;
            if (!__any_sync(0xffffffffu, n > 1))
            {
                const uint32_t m = __ballot_sync(0xffffffffu, n != 0);
                incl = (uint32_t)__popc(m & (0xffffffffu >> (31 - lane)));
                total = (uint32_t)__popc(m);
            }
            else
            {
                incl = (uint32_t)n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1)
                {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                total = __shfl_sync(0xffffffffu, incl, 31);
            }
            uint32_t base = vb * (uint32_t)SL_VB_PRIMS + it * SL_VB_THREADS + (uint32_t)warp * 32u;
            uint32_t granted = total;
            if (total > 32u) /* warp-uniform */
            {
                if (lane == 0)
                {
                    const unsigned long long o = atomicAdd(&P.counters->recAlloc, (unsigned long long)total);
                    if (o + total > (unsigned long long)P.recOverflowCapacity)
                    {
                        atomicOr(&P.counters->abortBatch, SL_ABORT_RECORDS); /* the host runs the batch again with a larger overflow area */
                        base = 0u;
                        granted = 0u;
                    }
                    else base = P.recOverflowBase + (uint32_t)o;
                }
                base = __shfl_sync(0xffffffffu, base, 0);
                granted = __shfl_sync(0xffffffffu, granted, 0);
            }
            /* rows the records of this warp touch (conservative: vertex rows, one more for spans that run past the row end) */
            uint32_t r0 = 65535u, r1 = 0u;
            if (n && granted)
            {
                if (needClip) { const uint32_t rows = sSlot[p]; r0 = rows & 0xffffu; r1 = rows >> 16; }
                else
                {
                    float a, b, cc;
                    if constexpr (CACHE) { a = cache[CV + ci[0]].y; b = cache[CV + ci[1]].y; cc = cache[CV + ci[2]].y; }
                    else { a = v[0].vert.v[1]; b = v[1].vert.v[1]; cc = v[2].vert.v[1]; }
                    const float fy0 = fminf(a, fminf(b, cc)), fy1 = fmaxf(a, fmaxf(b, cc));
                    /* screen y is a non-negative integer stored as a float (sl_to_screen); NaN rows produce no spans */
                    if (fy1 >= 0.f)
                    {
                        r0 = fy0 < 65535.f ? (uint32_t)fy0 : 65535u;
                        r1 = fy1 < 65533.f ? (uint32_t)fy1 + 2u : 65535u;
                    }
                }
            }
            if (granted) /* warp-uniform */
            {
                r0 = __reduce_min_sync(0xffffffffu, r0);
                r1 = __reduce_max_sync(0xffffffffu, r1);
            }
            if (lane == 0)
            {
                binned += granted;
                const SLSegment sg{base, granted, r0 | (r1 << 16)};
                const int s = (int)it * (SL_VB_THREADS / 32) + warp;
                P.out.blockInfo[vb].seg[s] = sg;
            }
            if (needClip) sSlot[p] = (n != 0 && granted != 0) ? base + (incl - (uint32_t)n) : 0xffffffffu; /* written in pass 3 */
            /* (Staging a warp's records in shared memory and storing them as one contiguous stream was measured: slower on
             * every config -- the kernel is bound by its index/vertex fetch latency, not by the record stores.) */
            else if (n != 0 && granted != 0)
            {
                const uint32_t slot = base + (incl - (uint32_t)n);
                uint16_t tag = (uint16_t)drawIdx;
                float4* rec = P.out.records + (size_t)slot * recStride;
                if constexpr (CACHE)
                {
                    const float4 s0 = cache[CV + ci[0]], s1 = cache[CV + ci[1]], s2 = cache[CV + ci[2]];
                    if (world > 1)
                    {
                        const uint32_t mask = sl_rank_mask(s0.y, s1.y, s2.y, P.partStrip, world);
                        P.recMask[slot] = (uint8_t)mask;
                        if (!((mask >> P.shardRank) & 1u)) tag = SL_REC_ABSENT;
                    }
                    rec[0] = s0;
                    rec[1] = s1;
                    rec[2] = s2;
#pragma unroll
                    for (int s = 0; s < 4; ++s)
                    {
                        if (s < NV)
                        {
                            rec[3 + 3 * s + 0] = cache[(2 + s) * CV + ci[0]];
                            rec[3 + 3 * s + 1] = cache[(2 + s) * CV + ci[1]];
                            rec[3 + 3 * s + 2] = cache[(2 + s) * CV + ci[2]];
                        }
                    }
                }
                else
                {
                    if (world > 1)
                    {
                        const uint32_t mask = sl_rank_mask(v[0].vert.v[1], v[1].vert.v[1], v[2].vert.v[1], P.partStrip, world);
                        P.recMask[slot] = (uint8_t)mask;
                        if (!((mask >> P.shardRank) & 1u)) tag = SL_REC_ABSENT;
                    }
                    sl_write_record(rec, NV, v[0].vert, v[1].vert, v[2].vert, v[0].varyings, v[1].varyings, v[2].varyings);
                }
                P.out.recDraw[slot] = tag;
            }
        }
        __syncthreads(); /* pass 0 -> 1: the queue is complete; 1 -> 2: the counts; 2 -> 3: the slots */
        if (pass == 0)
        {
            clipCount = sClipCount;
            if (!clipCount) break; /* block-uniform */
        }
    }
    if (lane == 0 && binned) atomicAdd(&P.counters->trisBinned, (unsigned long long)binned);
}

/* ------------------------------------------------------------------------
 * Span set-up: SL_ScanlineBounds (SL_ScanlineBounds.hpp:125-347)
 * ----------------------------------------------------------------------*/
struct SLScanline
{
    float v0y, v1y, v0x, v1x, p20y, p20x, p21xy, p10xy;

    SLD_FN void init(float4 p0, float4 p1, float4 p2)
    {
        /* sl_sort_minmax x3 (:60-68,139-141): swap when a.y < b.y */
        float4 t;
        if (p0.y < p1.y) { t = p0; p0 = p1; p1 = t; }
        if (p0.y < p2.y) { t = p0; p0 = p2; p2 = t; }
        if (p1.y < p2.y) { t = p1; p1 = p2; p2 = t; }
        v0y = p0.y;
        v1y = p1.y;
        v0x = p0.x;
        v1x = p1.x;
        p20y = sl_rcp(p2.y - p0.y);
        p20x = p2.x - p0.x;
        p21xy = (p2.x - p1.x) * sl_rcp(p2.y - p1.y);
        p10xy = (p1.x - p0.x) * sl_rcp(p1.y - p0.y);
    }

    SLD_FN void step(float yf, int32_t& xMin, int32_t& xMax) const /* :326-347 */
    {
        const float d0 = yf - v0y;
        const float d1 = yf - v1y;
        const float alpha = d0 * p20y;
        const float lo = sl_fma(p20x, alpha, v0x);
        const float pdv0 = sl_fma(p10xy, d0, v0x);
        const float pdv1 = sl_fma(p21xy, d1, v1x);
        const float hi = sl_signbit(d1) ? pdv1 : pdv0;
        xMin = sl_cvt_rn(sl_min(lo, hi));
        xMax = sl_cvt_rn(sl_max(lo, hi));
    }
};

struct SLSpanParams
{
    const float4* records;
    const uint16_t* recPresent;  /* shared vertex stage: recDraw, SL_REC_ABSENT = the record was not routed to this rank (null: all present) */
    const SLBlockInfo* blockInfo;
    unsigned int firstBlock;     /* first vertex block of the batch          */
    unsigned int numRows;        /* rows incl. the texture's padding rows    */
    unsigned int rowPitch;       /* padded row count of cnt/off              */
    unsigned int width;          /* framebuffer width (row stride)           */
    unsigned int segWidth;       /* x-segment width of the rasteriser        */
    unsigned int numSegs;        /* segments per row                         */
    int chained;                 /* depth test on: cut spans need hand-overs */
    int recStride;               /* float4 per record                        */
    uint16_t* cnt;               /* [numBlocks][rowPitch]                    */
    const uint32_t* off;         /* [numBlocks][rowPitch] exclusive over vb  */
    const uint32_t* rowStart;    /* [numRows+1]                              */
    SLSpanEntry* entries;
    SLCounters* counters;
    /* sort-first partition (null when this context owns every row) */
    uint8_t* rowFlags;           /* MODE 0 out: 1 = some span of the row below runs over into this row */
    const uint8_t* rowNeeded;    /* MODE 1 in : rows this rank rasterises (owned + the rows their hand-overs come from) */
    const uint32_t* items;       /* work items: vertex block | row band << 16 (k_span_items)               */
    const uint32_t* numItems;
    unsigned int* ticket;        /* item ticket of this pass (zeroed per batch)                              */
    int partRank, partWorld, partStrip; /* sort-first partition: row y belongs to rank (y / partStrip) % partWorld */
    int partFast;                /* sort-first fast path: rowNeeded = the rows this rank owns, and both passes leave every
                                    other row alone.  Valid as long as no span runs past the row end into a row of another
                                    rank; k_span<0> looks for exactly that and refuses the batch (SL_ABORT_PARTITION), the
                                    host then runs it again with every row counted and the rows closed over (k_row_closure) */
};

/* What one (record,row) span contributes to the row lists.  bit0: an entry in
 * its own row (some pixel left of the row end), bit1: an overflow entry in the
 * next row (some pixel at or beyond the row end). */
SLD_FN int sl_classify_span(const SLSpanParams& P, int32_t& xMin, int32_t& xMax)
{
    if (!(xMin < xMax)) return 0;
    const int32_t W = (int32_t)P.width;
    if (xMin < 0) { atomicOr(&P.counters->errorFlags, SL_ERR_SPAN_NEGATIVE_X); xMin = 0; }      /* would address the previous row: not reproduced */
    if (xMax > 2 * W) { atomicOr(&P.counters->errorFlags, SL_ERR_SPAN_TWO_ROWS); xMax = 2 * W; }
    if (!(xMin < xMax)) return 0;
    return (xMin < W ? 1 : 0) | (xMax > W ? 2 : 0);
}

/* Number of cuts of a (clamped) span of row y that need a hand-over: segment boundaries b*segWidth (b = 1..numSegs-1)
 * and the row end W strictly inside (xMin, xMax), and the same boundaries shifted by W for the part that wraps into row
 * y+1 (only when that row exists). */
SLD_FN unsigned sl_count_cuts(const SLSpanParams& P, int32_t xMin, int32_t xMax, bool nextRowExists)
{
    /* an upper bound: k_resolve2 may cut a heavy row every 2^SL_SHADE_SHIFT pixels inside a segment (its fine segments), so
     * those boundaries are counted for every row */
    const int32_t W = (int32_t)P.width, TW = (int32_t)P.segWidth, NS = (int32_t)P.numSegs;
    unsigned n = 0;
    for (int32_t part = 0; part < 2; ++part)
    {
        /* part 0: the span inside its own row; part 1: what wrapped into row y+1 (shifted by W) */
        int32_t a = xMin, b = xMax;
        if (part == 0) b = min(b, W);
        else
        {
            if (!(xMax > W && nextRowExists)) break;
            if (xMin < W) ++n; /* the row end itself */
            a = max(xMin - W, 0);
            b = xMax - W;
        }
        if (!(a < b)) continue;
        for (int32_t sIdx = a / TW; sIdx < NS && sIdx * TW < b; ++sIdx)
        {
            const int32_t s0 = sIdx * TW;
            if (s0 > a && s0 < b && !(part == 1 && s0 <= xMin - W)) ++n; /* the segment boundary */
            /* fine boundaries strictly inside (a, b) and inside this segment */
            const int32_t s1 = (sIdx == NS - 1) ? 0x7fffffff : s0 + TW;
            const int32_t lo = max(a, s0), hi = min(b, s1);
            const int32_t kFirst = ((lo - s0) >> SL_SHADE_SHIFT_C) + 1, kLast = (hi - 1 - s0) >> SL_SHADE_SHIFT_C;
            if (kLast >= kFirst) n += (unsigned)(kLast - kFirst + 1);
        }
    }
    return n;
}

/* Sort-first partition, interleaved strips: the number of rows below n that `rank` owns, and the k-th owned row at or above a */
SLD_FN int sl_owned_below(int n, int strip, int world, int rank)
{
    const int period = strip * world;
    const int q = n / period, r = n - q * period;
    return q * strip + min(max(r - rank * strip, 0), strip);
}
SLD_FN int sl_owned_kth(int a, int k, int strip, int world, int rank)
{
    const int j = sl_owned_below(a, strip, world, rank) + k;
    return (j / strip) * (strip * world) + rank * strip + (j % strip);
}

/* Rows a vertex block's records can have span entries in: the union of its segments' (one warp per block). */
__global__ void __launch_bounds__(256) k_block_rows(const SLBlockInfo* blockInfo, unsigned numBlocks, uint32_t* blockRows)
{
    const unsigned vb = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (vb >= numBlocks) return;
    const uint32_t rows = blockInfo[vb].seg[threadIdx.x & 31].rows;
    uint32_t lo = rows & 0xffffu, hi = rows >> 16;
    for (int d = 16; d > 0; d >>= 1)
    {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0) blockRows[vb] = lo | (hi << 16);
}

/* Work items of k_span: one per (vertex block, band of SL_SPAN_WIN rows the block's records touch).  A block of large
 * triangles covers a thousand rows and would keep one CTA busy long after the rest of the grid has drained; cut into bands
 * it is shared by several.  Bands of one block are independent: an entry's place in its row list depends on the block's
 * earlier records in that ROW only. */
__global__ void __launch_bounds__(1024) k_span_items(const uint32_t* blockRows, unsigned numBlocks, unsigned numRows, uint32_t* items, uint32_t* numItems)
{
    __shared__ uint32_t sWarp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned per = (numBlocks + 1023u) / 1024u;
    const unsigned a = (unsigned)tid * per, b = min(a + per, numBlocks);
    auto bands = [&](unsigned vb, uint32_t& first) -> uint32_t {
        const uint32_t y0 = blockRows[vb] & 0xffffu;
        uint32_t y1 = blockRows[vb] >> 16;
        if (y1 > numRows) y1 = numRows;
        if (!(y0 < y1)) { first = 0u; return 0u; }
        first = y0 >> SL_SPAN_BAND_SHIFT;
        return ((y1 - 1u) >> SL_SPAN_BAND_SHIFT) - first + 1u;
    };
    uint32_t mine = 0u, first = 0u;
    for (unsigned vb = a; vb < b; ++vb) mine += bands(vb, first);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        uint32_t w = sWarp[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        sWarp[lane] = wi - w;
        if (lane == 31) *numItems = wi;
    }
    __syncthreads();
    uint32_t off = sWarp[warp] + incl - mine;
    for (unsigned vb = a; vb < b; ++vb)
    {
        const uint32_t n = bands(vb, first);
        for (uint32_t k = 0; k < n; ++k) items[off + k] = vb | ((first + k) << 16);
        off += n;
    }
}

/* MODE 0: count entries per (vertex block, row).  MODE 1: fill the row lists.
 *
 * Persistent CTAs take (vertex block, row band) items through a ticket.  The block's records are taken 256 at a time (one
 * per thread: bounding rows + SL_ScanlineBounds set-up, kept in shared memory).  The (record, row) pairs of such a batch
 * that fall into the band are then flattened and dealt out evenly to the 256 threads, so that a triangle that covers the
 * whole band costs the CTA as much as 256 one-row triangles.  Order inside the batch comes from a bit mask per (row, kind):
 * bit i = record i of the batch has an entry in that row; an entry's place in its row list is the number of set bits
 * below its own. */
template <int MODE>
__global__ void __launch_bounds__(256) k_span(const __grid_constant__ SLSpanParams P)
{
    __shared__ uint16_t sRowRun[SL_SPAN_WIN];
    __shared__ uint32_t sMask[2][SL_SPAN_WIN][8]; /* [0] overflow entries coming from the row below, [1] own spans */
    __shared__ SLScanline sSl[256];
    __shared__ int sY0[256], sY1[256], sYEnd[256];
    __shared__ uint32_t sRec[256];
    __shared__ uint32_t sSegStart[SL_VB_SEGS + 1], sSegBase[SL_VB_SEGS];
    __shared__ uint32_t sPairOff[257];
    __shared__ uint32_t sWarpTot[8];
    __shared__ uint32_t sItem;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int H = (int)P.numRows;
    if (MODE == 1 && P.counters->abortBatch) return; /* the lists would not fit: the batch is run again */
    const uint32_t numItems = *P.numItems;

    unsigned myCuts = 0;
    unsigned long long myPix = 0;
    for (;;)
    {
        __syncthreads(); /* the previous item's shared state is no longer read */
        if (tid == 0) sItem = atomicAdd(P.ticket, 1u);
        __syncthreads();
        const uint32_t item = sItem;
        if (item >= numItems) break;
        const uint32_t packed = P.items[item];
        const unsigned int vb = packed & 0xffffu;
        const int band0 = (int)(packed >> 16) << SL_SPAN_BAND_SHIFT;
        const int band1 = min(band0 + SL_SPAN_WIN, H);
        sRowRun[tid] = 0;

        /* the block's records live in SL_VB_SEGS regions (one per round and warp of k_vertex); they are taken as ONE stream
         * in (round, warp) order = submission order, 256 at a time */
        static_assert(SL_VB_SEGS == 32, "the segment scan below is one warp wide");
        static_assert(SL_SPAN_WIN == 256, "one row of the band per thread");
        if (tid < 32)
        {
            /* exclusive scan of the segment counts */
            const SLSegment sg = P.blockInfo[P.firstBlock + vb].seg[tid];
            const uint32_t cnt = sg.count;
            uint32_t inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                if (tid >= d) inc += t;
            }
            sSegStart[tid] = inc - cnt;
            sSegBase[tid] = sg.base;
            if (tid == 31) sSegStart[32] = inc;
        }
        __syncthreads();
        const uint32_t numRecs = sSegStart[32];
        for (uint32_t sub = 0; sub < numRecs; sub += 256)
        {
            /* ---- one record per thread: rows it can have entries in, scan-line set-up ---- */
            const bool valid = sub + tid < numRecs;
            uint32_t myRec = 0u;
            if (valid)
            {
                const uint32_t v = sub + tid;
                int k = 0; /* last segment that starts at or before v */
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                {
                    if (sSegStart[k + step] <= v) k += step;
                }
                myRec = sSegBase[k] + (v - sSegStart[k]);
            }
            int y0 = 0, y1 = 0, yEnd = 0;
            bool mayOver = false;
            if (valid && !(P.recPresent && P.recPresent[myRec] == SL_REC_ABSENT)) /* an absent record has no entry in any of this rank's rows */
            {
                const float4* r = P.records + (size_t)myRec * P.recStride;
                const float4 p0 = r[0], p1 = r[1], p2 = r[2];
                /* src/SL_TriRasterizer.cpp:431-433 */
                y0 = sl_cvt_rn(sl_min(sl_min(p0.y, p1.y), p2.y));
                y1 = sl_cvt_rn(sl_max(sl_max(p0.y, p1.y), p2.y));
                y0 = y0 < 0 ? 0 : y0;
                y1 = y1 > H ? H : y1;
                /* span ends interpolate the vertex x's (rcpps error << 8 px), so only these can pass the row end */
                mayOver = sl_max(sl_max(p0.x, p1.x), p2.x) + 8.f > (float)P.width;
                yEnd = (mayOver && y1 < H) ? y1 + 1 : y1;
                if (!(y0 < y1)) { y0 = 0; y1 = 0; yEnd = 0; }
                if (max(y0, band0) < min(yEnd, band1)) /* only records that reach into the band need the set-up */
                {
                    SLScanline sl;
                    sl.init(p0, p1, p2);
                    sSl[tid] = sl;
                }
            }
            __syncthreads(); /* the previous batch's tables are no longer read */
            sRec[tid] = myRec;
            sY0[tid] = y0;
            sY1[tid] = mayOver ? -y1 - 1 : y1; /* sign bit: the record can run over the row end */
            sYEnd[tid] = yEnd;
            const int w0 = band0, w1 = band1;
            for (int i = tid; i < 2 * SL_SPAN_WIN * 8; i += 256) (&sMask[0][0][0])[i] = 0u;
            /* ---- flatten the (record, row) pairs of the band: exclusive scan of the per-record row counts ---- */
            /* sort-first fast path: a record that cannot run past the row end contributes its rows of THIS rank only */
            const bool ownedOnly = P.partFast && !mayOver;
            const int ra = max(y0, w0), rb = min(yEnd, w1);
            const int na = ra < rb ? (ownedOnly ? sl_owned_below(rb, P.partStrip, P.partWorld, P.partRank) - sl_owned_below(ra, P.partStrip, P.partWorld, P.partRank) : rb - ra) : 0;
            uint32_t incl = (uint32_t)na;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (lane == 31) sWarpTot[warp] = incl;
            __syncthreads();
            uint32_t warpOff = 0, total = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w)
            {
                const uint32_t t = sWarpTot[w];
                if (w < warp) warpOff += t;
                total += t;
            }
            sPairOff[tid] = warpOff + incl - (uint32_t)na;
            if (tid == 0) sPairOff[256] = total;
            __syncthreads();
            if (!total) continue; /* block-uniform */

            for (int pass = 0; pass <= MODE; ++pass)
            {
                for (uint32_t p = tid; p < total; p += 256)
                {
                    /* the record of pair p: last i with sPairOff[i] <= p */
                    int lo = 0, hi = 255;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                    {
                        const int mid = (lo + hi + 1) >> 1;
                        if (sPairOff[mid] <= p) lo = mid; else hi = mid - 1;
                    }
                    const int i = lo;
                    const int ry0 = sY0[i];
                    const int ry1raw = sY1[i];
                    const bool rOver = ry1raw < 0;
                    const int ry1 = rOver ? -ry1raw - 1 : ry1raw;
                    const int y = (P.partFast && !rOver) ? sl_owned_kth(max(ry0, w0), (int)(p - sPairOff[i]), P.partStrip, P.partWorld, P.partRank)
                                                         : max(ry0, w0) + (int)(p - sPairOff[i]);
                    if (MODE == 1 && P.rowNeeded && !P.rowNeeded[y]) continue; /* sort-first: another rank's row */
                    if (MODE == 0 && P.partFast)
                    {
                        /* another rank's row: only a span that could run over into one of MY rows matters here */
                        const bool mine = P.rowNeeded[y] != 0;
                        const bool nextMine = y + 1 < H && P.rowNeeded[y + 1] != 0;
                        if (!mine)
                        {
                            if (!(rOver && nextMine && y < ry1)) continue;
                            int32_t tMin = 0, tMax = 0;
                            sSl[i].step((float)y, tMin, tMax);
                            if (sl_classify_span(P, tMin, tMax) & 2) atomicOr(&P.counters->abortBatch, SL_ABORT_PARTITION);
                            continue;
                        }
                        if (rOver && y < ry1 && y + 1 < H && !nextMine)
                        {
                            int32_t tMin = 0, tMax = 0;
                            sSl[i].step((float)y, tMin, tMax);
                            if (sl_classify_span(P, tMin, tMax) & 2) atomicOr(&P.counters->abortBatch, SL_ABORT_PARTITION);
                        }
                    }
                    const uint32_t rec = sRec[i];
                    const uint32_t bit = 1u << (i & 31);
                    const int q = i >> 5;
                    int32_t xMin = 0, xMax = 0, oMin = 0, oMax = 0;
                    int own = 0, over = 0;
                    if (y < ry1)
                    {
                        sSl[i].step((float)y, xMin, xMax);
                        const int c = sl_classify_span(P, xMin, xMax);
                        own = c & 1;
                        if ((c & 2) && !rOver) atomicOr(&P.counters->errorFlags, SL_ERR_UNEXPECTED_OVERFLOW);
                        if (MODE == 0 && c)
                        {
                            myPix += (unsigned long long)(xMax - xMin);
                            if (P.chained) myCuts += sl_count_cuts(P, xMin, xMax, y + 1 < H);
                            /* the span statistic counts spans, the row lists hold entries: a span of the last row that
                             * lies entirely beyond the row end has no entry at all (its pixels are outside the image) */
                            if (c == 2 && y + 1 >= H) atomicAdd(&P.counters->spans, 1ull);
                            if (c == 3 && y + 1 < H)
                            {
                                atomicAdd(&P.counters->spans, ~0ull); /* two entries, one span */
                                atomicAdd(&P.counters->rowEndCuts, 1ull);
                                if (P.rowFlags) P.rowFlags[y + 1] = 1;
                            }
                        }
                    }
                    if (rOver && y > ry0)
                    {
                        sSl[i].step((float)(y - 1), oMin, oMax);
                        over = (sl_classify_span(P, oMin, oMax) >> 1) & 1;
                    }
                    if (pass == 0)
                    {
                        if (over) atomicOr(&sMask[0][y - w0][q], bit);
                        if (own) atomicOr(&sMask[1][y - w0][q], bit);
                    }
                    else if ((own | over) && (!P.rowNeeded || P.rowNeeded[y]))
                    {
                        uint32_t rank = 0;
                        for (int k = 0; k < q; ++k) rank += __popc(sMask[0][y - w0][k]) + __popc(sMask[1][y - w0][k]);
                        rank += __popc(sMask[0][y - w0][q] & (bit - 1u)) + __popc(sMask[1][y - w0][q] & (bit - 1u));
                        size_t slot = (size_t)P.rowStart[y] + P.off[(size_t)vb * P.rowPitch + y] + sRowRun[y - w0] + rank;
                        if (over) P.entries[slot++] = SLSpanEntry{rec | SL_ENTRY_OVERFLOW, (uint16_t)oMin, (uint16_t)oMax};
                        if (own) P.entries[slot] = SLSpanEntry{rec, (uint16_t)xMin, (uint16_t)xMax};
                    }
                }
                __syncthreads();
            }
            if (w0 + tid < w1)
            {
                uint32_t c = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) c += __popc(sMask[0][tid][k]) + __popc(sMask[1][tid][k]);
                sRowRun[tid] = (uint16_t)(sRowRun[tid] + c);
            }
        }
        __syncthreads();
        if (MODE == 0 && band0 + tid < band1) P.cnt[(size_t)vb * P.rowPitch + band0 + tid] = sRowRun[tid]; /* the other rows stay zero (memset per batch) */
    }
    if (MODE == 0)
    {
        for (int d = 16; d > 0; d >>= 1) myCuts += __shfl_down_sync(0xffffffffu, myCuts, d);
        if (lane == 0 && myCuts) atomicAdd(&P.counters->handoffs, (unsigned long long)myCuts);
        for (int d = 16; d > 0; d >>= 1) myPix += __shfl_down_sync(0xffffffffu, myPix, d);
        if (lane == 0 && myPix) atomicAdd(&P.counters->pixelsCovered, myPix);
    }
}

/* column scans of cnt over the vertex blocks, two-level */
constexpr int SL_SCAN_SEGS = 64;

/* Sort-first partition: a rank rasterises the rows it owns plus, transitively, every row whose spans run over into a row
 * it rasterises (that row's walker state is needed for the hand-over; it is recomputed locally instead of exchanged). */
__global__ void k_row_closure(const uint8_t* rowFlags, uint8_t* rowNeeded, unsigned numRows, unsigned rank, unsigned world, unsigned stripRows)
{
    const unsigned y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= numRows) return;
    bool needed = false;
    for (unsigned k = y;; ++k)
    {
        if ((k / stripRows) % world == rank) { needed = true; break; }
        if (k + 1 >= numRows || !rowFlags[k + 1]) break;
    }
    rowNeeded[y] = needed ? 1 : 0;
}

__global__ void k_scan_seg_sums(const uint16_t* cnt, uint32_t* segSums, unsigned numBlocks, unsigned rowPitch, unsigned blocksPerSeg, const uint8_t* rowNeeded, unsigned numRows)
{
    const unsigned y = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned seg = blockIdx.y;
    if (y >= rowPitch) return;
    if (rowNeeded && (y >= numRows || !rowNeeded[y])) { segSums[(size_t)seg * rowPitch + y] = 0; return; }
    const unsigned b0 = seg * blocksPerSeg;
    const unsigned b1 = min(b0 + blocksPerSeg, numBlocks);
    /* independent loads, eight in flight (the loop is latency-bound otherwise) */
    uint32_t s = 0;
    unsigned b = b0;
    for (; b + 8 <= b1; b += 8)
    {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = cnt[(size_t)(b + k) * rowPitch + y];
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[k];
    }
    for (; b < b1; ++b) s += cnt[(size_t)b * rowPitch + y];
    segSums[(size_t)seg * rowPitch + y] = s;
}

__global__ void k_scan_rows(uint32_t* segSums, uint32_t* rowTotals, unsigned rowPitch, unsigned numSegs)
{
    const unsigned y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= rowPitch) return;
    uint32_t run = 0;
    for (unsigned s = 0; s < numSegs; ++s)
    {
        const uint32_t c = segSums[(size_t)s * rowPitch + y];
        segSums[(size_t)s * rowPitch + y] = run;
        run += c;
    }
    rowTotals[y] = run;
}

/* what the batch was given; k_scan_row_starts compares it with what the count pass found */
struct SLBatchCaps
{
    unsigned long long entries;      /* span entries the row lists hold          */
    unsigned long long chunks;       /* chunks of the fragment pool              */
    unsigned long long handoffSlots; /* slots of the hand-over table             */
    unsigned long long numLists;     /* fragment lists (one partial chunk each)  */
    unsigned long long chunkSlack;   /* chunks the resolving warps may hold back */
    int chained;
};

/* single block: exclusive scan of rowTotals -> rowStart[0..numRows]; also accumulates the span counter and decides
 * whether the batch fits its buffers */
__global__ void __launch_bounds__(1024) k_scan_row_starts(const uint32_t* rowTotals, uint32_t* rowStart, unsigned numRows, SLCounters* counters, const SLBatchCaps caps)
{
    __shared__ uint32_t sPart[1024];
    const int tid = threadIdx.x;
    const unsigned per = (numRows + 1023u) / 1024u;
    const unsigned a = tid * per;
    const unsigned b = min(a + per, numRows);
    uint32_t s = 0;
    for (unsigned y = a; y < b; ++y) s += rowTotals[y];
    sPart[tid] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1)
    {
        const uint32_t t = tid >= d ? sPart[tid - d] : 0u;
        __syncthreads();
        sPart[tid] += t;
        __syncthreads();
    }
    uint32_t run = sPart[tid] - s;
    for (unsigned y = a; y < b; ++y)
    {
        rowStart[y] = run;
        run += rowTotals[y];
    }
    if (tid == 1023)
    {
        const uint32_t total = sPart[1023];
        rowStart[numRows] = total;
        atomicAdd(&counters->spans, (unsigned long long)total);
        counters->totalEntries = total;
        counters->rasterTicket = 0u;
        counters->chunkAlloc = 0u;
        counters->shadeTicket = 0u;
        unsigned abortBits = 0u;
        if ((unsigned long long)total > caps.entries) abortBits |= SL_ABORT_ENTRIES;
        /* every fragment is a pixel of a span, and a list ends with at most one partial chunk */
        if (counters->pixelsCovered / 32ull + caps.numLists + caps.chunkSlack > caps.chunks) abortBits |= SL_ABORT_CHUNKS;
        if (caps.chained && 2ull * counters->handoffs > caps.handoffSlots) abortBits |= SL_ABORT_HANDOFFS;
        if (abortBits) atomicOr(&counters->abortBatch, abortBits);
    }
}

/* Heavy rows first: rowPerm lists the rows with more than heavyFactor times the average number of span entries (in ascending row
 * order), then all the others.  k_resolve2 is as long as its longest (row, segment) ticket -- a row near the horizon of a
 * deep scene holds a few thousand overlapping pieces that can only be settled layer after layer -- so those tickets have
 * to be under way from the start instead of being handed out whenever their row number comes up. */
__global__ void __launch_bounds__(1024) k_row_order(const uint32_t* rowStart, unsigned numRows, uint32_t* rowPerm, unsigned heavyFactor)
{
    __shared__ uint32_t sWarp[32];
    __shared__ uint32_t sTotal;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t limit = heavyFactor * (rowStart[numRows] / (numRows ? numRows : 1u)) + 1u;
    const unsigned per = (numRows + 1023u) / 1024u;
    const unsigned a = (unsigned)tid * per, b = min(a + per, numRows);
    uint32_t heavy = 0u;
    for (unsigned y = a; y < b; ++y) heavy += (rowStart[y + 1] - rowStart[y]) > limit ? 1u : 0u;
    uint32_t incl = heavy;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        uint32_t w = sWarp[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        sWarp[lane] = wi - w;
        if (lane == 31) sTotal = wi;
    }
    __syncthreads();
    uint32_t hPos = sWarp[warp] + incl - heavy; /* heavy rows before this thread's chunk */
    const uint32_t numHeavy = sTotal;
    for (unsigned y = a; y < b; ++y)
    {
        if ((rowStart[y + 1] - rowStart[y]) > limit) rowPerm[hPos++] = y;
        else rowPerm[numHeavy + (y - hPos)] = y; /* light rows keep their order behind the heavy ones */
    }
    if (tid == 0) rowPerm[numRows] = numHeavy; /* k_resolve2 gives these rows fine segments */
}

/* Shared vertex stage: what this rank's blocks produced goes to the other ranks.  One CTA per block of this rank, one warp
 * per record segment: the segment descriptor to every rank, every record to the ranks of its mask, the recDraw word to all
 * of them (SL_REC_ABSENT where the record does not go).  A warp instruction carries whole records -- lanes of one record
 * write consecutive float4 -- so the peer stores leave the SM as full 32-byte sectors; k_vertex's own per-thread stores
 * (one float4 of 32 different records per instruction) would cross NVLink as 16-byte packets. */
struct SLGeomRoute
{
    const uint32_t* blockDraw;
    SLGeomOut out[SL_MAX_RANKS];       /* every rank's arena (peer mappings); [rank] = the source */
    const uint8_t* recMask;
    int rank, world, recStride;
};
__global__ void __launch_bounds__(32 * SL_VB_SEGS) k_geom_route(const __grid_constant__ SLGeomRoute P)
{
    const unsigned vb = blockIdx.x * (unsigned)P.world + (unsigned)P.rank;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint16_t drawIdx = (uint16_t)P.blockDraw[vb];
    const int stride = P.recStride;
    const int perIter = 32 / stride;          /* records per warp instruction (stride <= 15) */
    const int sub = lane / stride, f = lane - sub * stride;
    const float4* src = P.out[P.rank].records;
    {
        const int s = warp; /* one warp per segment: 32 independent load chains per block */
        const SLSegment sg = P.out[P.rank].blockInfo[vb].seg[s];
        if (lane < P.world && lane != P.rank) P.out[lane].blockInfo[vb].seg[s] = sg;
        for (uint32_t r0 = 0; r0 < sg.count; r0 += 32u)
        {
            /* the recDraw words of up to 32 records: one lane each */
            const uint32_t r = r0 + (uint32_t)lane;
            const uint32_t mask = r < sg.count ? P.recMask[sg.base + r] : 0u;
            if (r < sg.count)
            {
                for (int t = 0; t < P.world; ++t)
                {
                    if (t != P.rank) P.out[t].recDraw[sg.base + r] = ((mask >> t) & 1u) ? drawIdx : SL_REC_ABSENT;
                }
            }
            const uint32_t n = min(32u, sg.count - r0);
            for (uint32_t q = 0; q < n; q += (uint32_t)perIter)
            {
                const uint32_t mine = q + (uint32_t)sub;
                const uint32_t m = __shfl_sync(0xffffffffu, mask, (int)(mine & 31u));
                if (sub < perIter && mine < n)
                {
                    const size_t at = (size_t)(sg.base + r0 + mine) * stride + f;
                    const uint32_t others = m & ~(1u << P.rank);
                    if (others)
                    {
                        const float4 val = src[at];
                        for (int t = 0; t < P.world; ++t)
                        {
                            if ((others >> t) & 1u) P.out[t].records[at] = val;
                        }
                    }
                }
            }
        }
    }
}

/* Shared vertex stage: the rendezvous between k_vertex and everything that reads the records.  One thread per rank:
 * thread t tells rank t that this rank's blocks of batch `epoch` have landed in its arena (k_vertex's peer stores
 * completed with that kernel; the release store orders the count word before it), then waits for rank t's word to show
 * up here.  Words alternate with the epoch's parity: a rank can be one batch ahead of another, never two (it cannot
 * pass the next rendezvous without the slower rank's word).  Bit 31 of a word: that rank's record area overflowed.
 * The record counts of the other ranks' blocks join this rank's statistic.  `words`: [2][SL_MAX_RANKS] epoch words, then
 * [2][SL_MAX_RANKS] 64-bit counts, at the head of every rank's shared allocation. */
struct SLGeomSync
{
    unsigned char* head[SL_MAX_RANKS]; /* rank t's shared allocation (peer mapping) */
    int rank, world;
    uint32_t epoch;                    /* 1, 2, ...: < 2^31 */
    SLCounters* counters;
    long long timeoutClocks;
};
constexpr size_t SL_GEOM_HEAD_BYTES = 256;
__global__ void k_geom_barrier(const __grid_constant__ SLGeomSync P)
{
    const int t = (int)threadIdx.x;
    if (t >= P.world) return;
    const unsigned par = P.epoch & 1u;
    uint32_t* theirWords = reinterpret_cast<uint32_t*>(P.head[t]);
    unsigned long long* theirCounts = reinterpret_cast<unsigned long long*>(P.head[t] + 2 * SL_MAX_RANKS * sizeof(uint32_t));
    const uint32_t* myWords = reinterpret_cast<const uint32_t*>(P.head[P.rank]);
    const unsigned long long* myCounts = reinterpret_cast<const unsigned long long*>(P.head[P.rank] + 2 * SL_MAX_RANKS * sizeof(uint32_t));
    const uint32_t word = P.epoch | ((P.counters->abortBatch & SL_ABORT_RECORDS) ? 0x80000000u : 0u);
    theirCounts[par * SL_MAX_RANKS + P.rank] = P.counters->trisBinned;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirWords + par * SL_MAX_RANKS + P.rank), "r"(word) : "memory");
    const uint32_t* src = myWords + par * SL_MAX_RANKS + t;
    const long long t0 = clock64();
    uint32_t got = 0u;
    for (;;)
    {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(src) : "memory");
        if ((got & 0x7fffffffu) == P.epoch) break;
        if (clock64() - t0 > P.timeoutClocks)
        {
            atomicOr(&P.counters->errorFlags, SL_ERR_GEOM_TIMEOUT);
            atomicOr(&P.counters->abortBatch, SL_ABORT_GEOMETRY);
            return;
        }
        __nanosleep(200);
    }
    if (got & 0x80000000u) atomicOr(&P.counters->abortBatch, SL_ABORT_GEOMETRY);
    if (t != P.rank) atomicAdd(&P.counters->trisBinned, myCounts[par * SL_MAX_RANKS + t]);
}

/* end of a batch: its statistics join the totals unless it has to be run again */
__global__ void k_commit(SLCounters* counters)
{
    if (counters->abortBatch) return;
    counters->totalTrisBinned += counters->trisBinned;
    counters->totalSpans += counters->spans;
    counters->totalFragsShaded += counters->fragsShaded;
}

__global__ void k_scan_offsets(const uint16_t* cnt, const uint32_t* segSums, uint32_t* off, unsigned numBlocks, unsigned rowPitch, unsigned blocksPerSeg, const uint8_t* rowNeeded, unsigned numRows)
{
    const unsigned y = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned seg = blockIdx.y;
    if (y >= rowPitch) return;
    if (rowNeeded && (y >= numRows || !rowNeeded[y])) return;
    const unsigned b0 = seg * blocksPerSeg;
    const unsigned b1 = min(b0 + blocksPerSeg, numBlocks);
    uint32_t run = segSums[(size_t)seg * rowPitch + y];
    unsigned b = b0;
    for (; b + 8 <= b1; b += 8)
    {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = cnt[(size_t)(b + k) * rowPitch + y];
#pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            off[(size_t)(b + k) * rowPitch + y] = run;
            run += v[k];
        }
    }
    for (; b < b1; ++b)
    {
        off[(size_t)b * rowPitch + y] = run;
        run += cnt[(size_t)b * rowPitch + y];
    }
}

/* ------------------------------------------------------------------------
 * Fragment stage
 * ----------------------------------------------------------------------*/
/* Fragment lists (k_resolve -> k_shade).  One list per (row, x-segment) = per resolving warp, in submission order per
 * pixel; a list is a chain of chunks of 32 fragments, structure of arrays inside the chunk so that both sides move
 * whole 128/512-byte lines:  chunk c = [float4 bc[32]] [float z[32]] [u32 x|y<<16 [32]] [u32 record[32]]. */
constexpr unsigned SL_CHUNK_FRAGS = 32u;
constexpr unsigned SL_CHUNK_BYTES = SL_CHUNK_FRAGS * (16u + 4u + 4u + 4u);
constexpr uint32_t SL_NO_CHUNK = 0xffffffffu;

struct SLFragLists
{
    unsigned char* pool;      /* chunkCapacity * SL_CHUNK_BYTES                          */
    uint32_t* chunkNext;      /* [chunkCapacity]                                         */
    uint32_t* chunkPrev;      /* [chunkCapacity] the chunk before (k_shade walks opaque lists from their end) */
    uint32_t* listHead;       /* [numRows * numSegs] first chunk of the list             */
    uint32_t* listCount;      /* [numRows * numSegs] fragments in the list (zeroed per batch) */
    uint32_t* listTail;       /* [numRows * numSegs] last chunk of the list              */
    unsigned int chunkCapacity;
};

struct SLRasterParams
{
    const float4* records;
    const SLSpanEntry* entries;
    const uint32_t* rowStart;
    int recStride;
    int depthTest;   /* SL_DepthTest */
    int depthMask;
    unsigned int numRows;     /* rows incl. the texture's padding rows */
    unsigned int width;       /* depth attachment width (row stride of every attachment, like the reference) */
    unsigned long long* handoff; /* hash table of hand-overs (null when the batch has none) */
    float4* handoffRegs;         /* [slots][4]: the four barycentric registers at the cut    */
    unsigned int handoffMask;
    unsigned int segWidth;       /* x-segment width */
    unsigned int numSegs;        /* segments per row (<= 32) */
    unsigned int numSubLists;    /* fragment lists per segment: one per 2^SL_SHADE_SHIFT pixels */
    unsigned int tileBytes;      /* shared memory per warp for its segment's depth texels (0: work in global memory) */
    const uint8_t* rowNeeded;    /* sort-first partition: rows this rank rasterises (null = all) */
    const uint32_t* ticketOrder; /* ticket -> row << 6 | segment (null: row-major)               */
    const uint32_t* rowPerm;     /* with ticketOrder: its rows are ranks, rowPerm[rank] = the row (heaviest first; k_row_order) */
    unsigned int fineRows;       /* k_resolve2: rows with more than fineRows x the average number of entries get fine segments (0: never) */
    unsigned char* depth;
    int depthBpt;
    SLFragLists lists;
    SLCounters* counters;
    int diagnostics;             /* SLCU_DEBUG: count windows / waves / idle polls / clocks */
};

struct SLShadeParams
{
    const SLDrawDesc* draws;     /* draw table of the batch: the fragment shader's uniforms are draws[recDraw[record]].uniforms */
    const uint16_t* recDraw;
    const float4* records;
    int recStride;
    int numVaryings;
    int numOutputs;
    int blend;                /* SL_BlendMode */
    unsigned int width;
    unsigned int numLists;
    unsigned int segWidth, numSegs, numSubLists; /* list l covers 2^SL_SHADE_SHIFT texels from width * row + seg * segWidth + (sub << SL_SHADE_SHIFT) */
    int opaqueSkip;           /* blend off: shade only the fragment that ends up in the texel (see k_shade) */
    unsigned char* color[4];
    int colorType[4];         /* SL_ColorDataType */
    SLFragLists lists;
    unsigned int* shadeTicket; /* list ticket of k_shade (zeroed per batch) */
    unsigned int listBegin;    /* this launch shades lists listBegin .. numLists - 1 (a band of rows when a read-back follows) */
    const unsigned int* abortBatch; /* SLCounters::abortBatch */
};

/* depth texel read (src/SL_TriRasterizer.cpp:40-102) */
SLD_FN float sl_load_depth(const unsigned char* base, int bpt, size_t idx)
{
    if (bpt == 2) return sl_h2f(reinterpret_cast<const uint16_t*>(base)[idx]);
    if (bpt == 4) return reinterpret_cast<const float*>(base)[idx];
    return (float)reinterpret_cast<const double*>(base)[idx];
}
/* depth write (src/SL_FragmentProcessor.cpp:526-529) */
SLD_FN void sl_store_depth(unsigned char* base, int bpt, size_t idx, float z)
{
    if (z != z)
    {
        /* cvtps2ph / cvtss2sd keep a NaN's sign and the top of its payload and set the quiet bit; the GPU's conversions
         * return one canonical NaN */
        const uint32_t u = __float_as_uint(z);
        if (bpt == 2) reinterpret_cast<uint16_t*>(base)[idx] = (uint16_t)(((u >> 16) & 0x8000u) | 0x7e00u | ((u >> 13) & 0x3ffu));
        else if (bpt == 4) reinterpret_cast<float*>(base)[idx] = z;
        else reinterpret_cast<unsigned long long*>(base)[idx] = ((unsigned long long)(u >> 31) << 63) | 0x7ff8000000000000ull | ((unsigned long long)(u & 0x3fffffu) << 29);
        return;
    }
    if (bpt == 2) reinterpret_cast<uint16_t*>(base)[idx] = sl_f2h_rn(z);
    else if (bpt == 4) reinterpret_cast<float*>(base)[idx] = z;
    else reinterpret_cast<double*>(base)[idx] = (double)z;
}

/* SL_DepthFunc* (SL_ShaderUtil.hpp:240-507): ordered, quiet */
SLD_FN bool sl_depth_cmp(int func, float z, float d)
{
    switch (func)
    {
        case 0: return true;
        case 1: return z < d;
        case 2: return z <= d;
        case 3: return z > d;
        case 4: return z >= d;
        case 5: return z == d;
        default: return (z < d) || (z > d);
    }
}
/* The functors' SCALAR overloads (SL_ShaderUtil.hpp:240-507), which the point, line and wireframe rasterisers call
 * (src/SL_PointRasterizer.cpp:58, src/SL_LineRasterizer.cpp:64, src/SL_TriRasterizer.cpp:177): the same as the SIMD ones above
 * except NOT_EQUAL, which is the C++ `a != b` -- TRUE when the fragment's depth is NaN (a clipped zero-length line
 * interpolates with 0 * inf), where _CMP_NEQ_OQ says false. */
SLD_FN bool sl_depth_cmp_scalar(int func, float z, float d)
{
    return func == 6 ? !(z == d) : sl_depth_cmp(func, z, d);
}

/* assign_pixel<SL_ColorRGB332 ... SL_ColorRGB1010102> (src/SL_Framebuffer.cpp:85-145) and rgba_cast<float, packed>, the
 * destination read of the packed assign_alpha_pixel (:371-590).  Out of line and apart from the generic pair below, which
 * stays as it was measured. */
__device__ __noinline__ void sl_store_color_packed(unsigned char* base, int type, size_t idx, const math::vec4& c)
{
    const int ci = type - 28;
    sl_packed_store(base + idx * (size_t)sl_pack_bytes(ci), ci, sl_packed_from_float(ci, c));
}
__device__ __noinline__ math::vec4 sl_load_color_packed(const unsigned char* base, int type, size_t idx)
{
    const int ci = type - 28;
    return sl_float_from_packed(ci, sl_packed_load(base + idx * (size_t)sl_pack_bytes(ci), ci));
}

/* assign_pixel<T> (src/SL_Framebuffer.cpp:44-217), every format.  Kept out of line: the two formats the configs use
 * have inline fast paths below, and the rasteriser's instruction footprint matters (profiles/). */
__device__ __noinline__ void sl_store_color_generic(unsigned char* base, int type, size_t idx, const math::vec4& c)
{
    const int kind = type % 7;
    const int n = type / 7 + 1;
    if (kind == 0)
    {
        const math::vec4_t<uint8_t> r = math::vec4_to_u8(c * 255.f);
        if (n == 4) reinterpret_cast<uint32_t*>(base)[idx] = (uint32_t)r.v[0] | ((uint32_t)r.v[1] << 8) | ((uint32_t)r.v[2] << 16) | ((uint32_t)r.v[3] << 24);
        else { unsigned char* p = base + idx * n; for (int i = 0; i < n; ++i) p[i] = r.v[i]; }
    }
    else if (kind == 1)
    {
        uint16_t* p = reinterpret_cast<uint16_t*>(base) + idx * n;
        for (int i = 0; i < n; ++i)
        {
            const int32_t v = sl_cvt_rn(c.v[i] * 65535.f);
            p[i] = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
        }
    }
    else if (kind == 2)
    {
        /* color_cast<uint32_t,float> = (vec4<u32>)(c * (float)UINT32_MAX) with the vec4 conversion cvtps2dq (x86/vec4f_impl.h:
         * 333-342): signed, RNE, 0x80000000 from 0.5 upwards -- reproduced, not repaired */
        uint32_t* p = reinterpret_cast<uint32_t*>(base) + idx * n;
        for (int i = 0; i < n; ++i) p[i] = (uint32_t)sl_cvt_rn(c.v[i] * 4294967296.f);
    }
    else if (kind == 3)
    {
        /* the generic element-wise (uint64_t)float (vec4_impl.h:184-187) as g++ compiles it: cvttss2si below 2^63 (NaN and
         * negatives included: unordered compares take this branch), otherwise cvttss2si(f - 2^63) ^ 2^63 */
        unsigned long long* p = reinterpret_cast<unsigned long long*>(base) + idx * n;
        for (int i = 0; i < n; ++i)
        {
            const float f = c.v[i] * 18446744073709551616.f;
            p[i] = !(f >= 9223372036854775808.f) ? (unsigned long long)sl_cvt_tr64(f)
                                                 : ((unsigned long long)sl_cvt_tr64(f - 9223372036854775808.f) ^ 0x8000000000000000ull);
        }
    }
    else if (kind == 4)
    {
        uint16_t* p = reinterpret_cast<uint16_t*>(base) + idx * n;
        for (int i = 0; i < n; ++i) p[i] = sl_f2h_rz(c.v[i]); /* vec4<float> -> vec4<half>: cvtps2ph toward zero */
    }
    else if (kind == 5)
    {
        if (n == 4) reinterpret_cast<float4*>(base)[idx] = make_float4(c.v[0], c.v[1], c.v[2], c.v[3]);
        else { float* p = reinterpret_cast<float*>(base) + idx * n; for (int i = 0; i < n; ++i) p[i] = c.v[i]; }
    }
    else if (kind == 6)
    {
        double* p = reinterpret_cast<double*>(base) + idx * n;
        for (int i = 0; i < n; ++i) p[i] = (double)c.v[i];
    }
}

SLD_FN void sl_store_color(unsigned char* base, int type, size_t idx, const math::vec4& c)
{
    if (type == 21) /* SL_COLOR_RGBA_8U */
    {
        const math::vec4_t<uint8_t> r = math::vec4_to_u8(c * 255.f);
        reinterpret_cast<uint32_t*>(base)[idx] = (uint32_t)r.v[0] | ((uint32_t)r.v[1] << 8) | ((uint32_t)r.v[2] << 16) | ((uint32_t)r.v[3] << 24);
    }
    else if (type == 26) /* SL_COLOR_RGBA_FLOAT */
        reinterpret_cast<float4*>(base)[idx] = make_float4(c.v[0], c.v[1], c.v[2], c.v[3]);
    else if (sl_is_packed(type)) sl_store_color_packed(base, type, idx, c);
    else sl_store_color_generic(base, type, idx, c);
}

/* destination read of assign_alpha_pixel (src/SL_Framebuffer.cpp:284-312), 3-channel quirk (:308) included */
__device__ __noinline__ math::vec4 sl_load_color_generic(const unsigned char* base, int type, size_t idx)
{
    const int kind = type % 7;
    const int n = type / 7 + 1;
    float ch[4];
    float hi = 1.f, scale = 1.f;
    if (kind == 0)
    {
        const unsigned char* p = base + idx * n;
        for (int i = 0; i < n; ++i) ch[i] = (float)p[i];
        hi = 255.f;
        scale = 0.00392156862745f;
    }
    else if (kind == 1)
    {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(base) + idx * n;
        for (int i = 0; i < n; ++i) ch[i] = (float)p[i];
        hi = 65535.f;
        scale = 1.f / 65535.f;
    }
    else if (kind == 2)
    {
        /* vec4<uint32_t> -> vec4<float> (x86/vec4f_impl.h:236-245): (float)(v >> 1) twice plus (float)(v & 1), every step rounded */
        const uint32_t* p = reinterpret_cast<const uint32_t*>(base) + idx * n;
        for (int i = 0; i < n; ++i) ch[i] = sl_u32_to_float(p[i]);
        hi = sl_u32_to_float(0xffffffffu);
        scale = 1.f / 4294967296.f;
    }
    else if (kind == 3)
    {
        const unsigned long long* p = reinterpret_cast<const unsigned long long*>(base) + idx * n;
        for (int i = 0; i < n; ++i) ch[i] = __ull2float_rn(p[i]); /* g++'s uint64 -> float sequence is correctly rounded */
        hi = __ull2float_rn(0xffffffffffffffffull);
        scale = 1.f / 18446744073709551616.f;
    }
    else if (kind == 4)
    {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(base) + idx * n;
        for (int i = 0; i < n; ++i) ch[i] = sl_h2f(p[i]);
    }
    else if (kind == 5)
    {
        const float* p = reinterpret_cast<const float*>(base) + idx * n;
        for (int i = 0; i < n; ++i) ch[i] = p[i];
    }
    else
    {
        const double* p = reinterpret_cast<const double*>(base) + idx * n;
        for (int i = 0; i < n; ++i) ch[i] = (float)p[i];
    }
    math::vec4 d;
    if (n == 1) d = math::vec4{ch[0], 0.f, 0.f, hi};
    else if (n == 2) d = math::vec4{ch[0], ch[1], 0.f, hi};
    else if (n == 3) d = math::vec4{ch[0], ch[0], ch[0], hi};
    else d = math::vec4{ch[0], ch[1], ch[2], ch[3]};
    if (kind <= 3) d = d * scale;
    return d;
}

SLD_FN math::vec4 sl_load_color(const unsigned char* base, int type, size_t idx)
{
    if (type == 21) /* SL_COLOR_RGBA_8U: color_cast<float,uint8_t> = (float)c * 0.00392156862745f */
    {
        const uint32_t u = reinterpret_cast<const uint32_t*>(base)[idx];
        return math::vec4{(float)(u & 0xffu), (float)((u >> 8) & 0xffu), (float)((u >> 16) & 0xffu), (float)(u >> 24)} * 0.00392156862745f;
    }
    if (type == 26) /* SL_COLOR_RGBA_FLOAT */
    {
        const float4 f = reinterpret_cast<const float4*>(base)[idx];
        return math::vec4{f.x, f.y, f.z, f.w};
    }
    if (sl_is_packed(type)) return sl_load_color_packed(base, type, idx);
    return sl_load_color_generic(base, type, idx);
}

/* assign_alpha_pixel<T> (src/SL_Framebuffer.cpp:270-368) */
SLD_FN void sl_blend_color(unsigned char* base, int type, size_t idx, const math::vec4& rgba, int blendMode)
{
    math::vec4 d = sl_load_color(base, type, idx);
    const math::vec4 srcAlpha{rgba.v[3]};
    const math::vec4 modulation = math::vec4{1.f} - srcAlpha;
    if (blendMode == 1)
    {
        const math::vec4 dstMod = modulation * d.v[3];
        const math::vec4 dstAlpha = dstMod + srcAlpha;
        const math::vec4 dstRgb = math::fmadd(rgba, srcAlpha, (d * dstMod)) * math::rcp(dstAlpha);
        d = math::vec4{dstRgb.v[0], dstRgb.v[1], dstRgb.v[2], dstAlpha.v[0]};
    }
    else if (blendMode == 2) d = math::fmadd(d, modulation, rgba);
    else if (blendMode == 3) d = math::fmadd(rgba, srcAlpha, d);
    else if (blendMode == 4) d = (rgba * srcAlpha) + (d * modulation);
    d = math::clamp(d, math::vec4{0.f, 0.f, 0.f, 0.f}, math::vec4{1.f, 1.f, 1.f, 1.f});
    sl_store_color(base, type, idx, d);
}

struct SLTriSetup
{
    math::vec4 bc0, bc1, bc2; /* barycentric plane equations (push_bin, src/SL_TriProcessor.cpp:271-314) */
    float z0, z1, z2;         /* vertex depths                                                           */
    float w0, w1, w2;         /* 1/w of the three vertices                                               */
};

SLD_FN void sl_tri_setup(SLTriSetup& t, const float4 p0, const float4 p1, const float4 p2)
{
    const float xy0 = p0.x - p1.x, xy1 = p0.y - p1.y;
    const float zy0 = p2.x - p1.x, zy1 = p2.y - p1.y;
    const float denom = sl_rcp((zy0 * xy1) - (zy1 * xy0)); /* math::cross(vec2 zy, vec2 xy), generic */
    const math::vec4 ptX{p0.x, p1.x, p2.x, 0.f};
    const math::vec4 ptY{p0.y, p1.y, p2.y, 0.f};
    const math::vec4 one{1.f};
    t.bc0 = math::cross(ptY, one) * denom;
    t.bc1 = math::cross(one, ptX) * denom;
    t.bc2 = math::cross(ptX, ptY) * denom;
    t.z0 = p0.z; t.z1 = p1.z; t.z2 = p2.z;
    t.w0 = p0.w; t.w1 = p1.w; t.w2 = p2.w;
}

/* Where a resolving warp's depth texels live while it works on them: the depth texels of its (row, x-segment) staged
 * into shared memory (origin = linear index x + width*y of the segment's first pixel), or the attachment itself
 * (origin 0) when a segment does not fit the shared-memory budget. */
struct SLTile
{
    unsigned char* depth;
    long long origin;
};

/* ------------------------------------------------------------------------
 * The fragment stage is two kernels.
 *
 *   k_resolve (one for every shader) runs the 4-pixel group loop of render_triangle_simd
 *   (src/SL_TriRasterizer.cpp:471-548): depth of the four lanes, depth test against the tile, the in-place compaction
 *   of the barycentric registers (:496-498), the depth write (src/SL_FragmentProcessor.cpp:526-529 -- the value is the
 *   rasteriser's z and does not depend on the shader).  Every surviving fragment goes straight into the chunks of the
 *   fragment list of its (row, 128-pixel column) in HBM: {barycentrics the
 *   reference would hand to the shader, z, x, y, record}.  All ordering constraints of the path (submission order per
 *   pixel, the left-to-right compaction state of a span, hand-overs between segments and rows) are settled here, by a
 *   kernel that is shared by all shaders and never waits for shading.
 *
 *   k_shade<FS>: persistent warps take fragment lists through a ticket; 32 fragments per step, one per lane (flush_tri_fragments,
 *   src/SL_FragmentProcessor.cpp:416-531): perspective correction, varyings, the fragment shader, colour write /
 *   blend.  Lists are in submission order per pixel and a list is walked front to back, so colour writes to a pixel
 *   happen in submission order; fragments of one step that hit the same pixel are ordered by lane.  Lists never share
 *   a pixel, so list order is free.  The shader always runs with converged lanes.
 * ----------------------------------------------------------------------*/
constexpr unsigned SL_PIECEQ_BYTES = 32u * (unsigned)sizeof(SLSpanEntry);
constexpr int SL_MAX_SUBLISTS = 8;      /* fragment lists per resolving warp: one per SL_SHADE_WIDTH pixels of its segment */
constexpr int SL_SHADE_SHIFT = SL_SHADE_SHIFT_C; /* log2(SL_SHADE_WIDTH)                                                  */
constexpr unsigned SL_SUBLIST_BYTES = SL_MAX_SUBLISTS * 3u * 4u + 32u; /* per warp: tail chunk, fill, total of every sub-list */

/* Append the survivors of a window to the warp's fragment lists, straight into the chunks in HBM.  Lists are ordered by
 * (lane, pixel) = submission order per pixel; a group's pixels may fall into two neighbouring sub-lists.  Per sub-list:
 * an exclusive scan of the lanes' counts gives every fragment its position behind the list's tail; the chunks that
 * position range needs beyond the tail chunk are taken with ONE atomic and linked by the first lanes. */
struct SLListState
{
    uint32_t* tail;   /* [nsub] current chunk of the sub-list (SL_NO_CHUNK: none yet)  */
    uint32_t* fill;   /* [nsub] fragments in it, 1..32 (32 when there is no chunk)      */
    uint32_t* total;  /* [nsub] fragments of the sub-list so far                        */
    uint32_t firstList;
    int nsub;
};

SLD_FN void sl_push_frags(const SLRasterParams& P, const SLListState& L, int lane, uint32_t mBits, const math::vec4 (&bcJ)[4], const float (&zJ)[4], int x4,
                          int relX, int y, uint32_t recId)
{
    if (!__any_sync(0xffffffffu, mBits != 0u)) return;
    /* sub-list of every pixel of the group (relX = x of the group's first pixel relative to the segment start) */
    const int sFirst = relX >> SL_SHADE_SHIFT, sLast = (relX + 3) >> SL_SHADE_SHIFT;
    const int jSplit = sLast != sFirst ? (((sFirst + 1) << SL_SHADE_SHIFT) - relX) : 4; /* pixels j >= jSplit belong to sLast */
    const uint32_t loBits = mBits & ((1u << jSplit) - 1u), hiBits = mBits & ~((1u << jSplit) - 1u);
    const int sMin = __reduce_min_sync(0xffffffffu, mBits ? (loBits ? sFirst : sLast) : 0x7fffffff);
    const int sMax = __reduce_max_sync(0xffffffffu, mBits ? (hiBits ? sLast : sFirst) : -1);
    for (int sub = sMin; sub <= sMax; ++sub)
    {
        const uint32_t bits = (sub == sFirst ? loBits : 0u) | (sub == sLast ? hiBits : 0u);
        const int n = __popc(bits);
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int T = __shfl_sync(0xffffffffu, incl, 31);
        if (!T) continue;
        const uint32_t tail = L.tail[sub], fill = L.fill[sub];
        const uint32_t need = (fill + (uint32_t)T - 1u) >> 5; /* chunks beyond the tail chunk */
        uint32_t baseChunk = 0u;
        if (need)
        {
            if (lane == 0)
            {
                baseChunk = atomicAdd(&P.counters->chunkAlloc, need);
                if (baseChunk + need > P.lists.chunkCapacity) { atomicOr(&P.counters->errorFlags, SL_ERR_POOL); baseChunk = SL_NO_CHUNK; }
            }
            baseChunk = __shfl_sync(0xffffffffu, baseChunk, 0);
            if (baseChunk == SL_NO_CHUNK) continue; /* pool exhausted: flagged, fragments dropped */
            if ((uint32_t)lane < need)
            {
                const uint32_t prev = lane == 0 ? tail : baseChunk + (uint32_t)lane - 1u;
                if (prev == SL_NO_CHUNK) P.lists.listHead[L.firstList + sub] = baseChunk;
                else P.lists.chunkNext[prev] = baseChunk + (uint32_t)lane;
                P.lists.chunkPrev[baseChunk + (uint32_t)lane] = prev;
            }
        }
        uint32_t q = fill + (uint32_t)(incl - n);
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            if ((bits >> j) & 1u)
            {
                const uint32_t ci = q >> 5, slot = q & 31u;
                unsigned char* c = P.lists.pool + (size_t)(ci ? baseChunk + ci - 1u : tail) * SL_CHUNK_BYTES;
                reinterpret_cast<float4*>(c)[slot] = make_float4(bcJ[j].v[0], bcJ[j].v[1], bcJ[j].v[2], bcJ[j].v[3]);
                reinterpret_cast<float*>(c + 512)[slot] = zJ[j];
                reinterpret_cast<uint32_t*>(c + 640)[slot] = (uint32_t)(x4 + j) | ((uint32_t)y << 16);
                reinterpret_cast<uint32_t*>(c + 768)[slot] = recId;
                ++q;
            }
        }
        __syncwarp();
        if (lane == 0)
        {
            L.fill[sub] = ((fill + (uint32_t)T - 1u) & 31u) + 1u;
            if (need) L.tail[sub] = baseChunk + need - 1u;
            L.total[sub] += (uint32_t)T;
        }
        __syncwarp();
    }
}

/* Hand-over of the walker state at a cut (see SL_ENTRY_OVERFLOW).  A row is resolved by numSegs warps, one per
 * x-segment; a span that crosses a segment boundary (or the row end) is walked piece by piece, and the piece on the
 * right continues from the state the piece on the left ends with: the four barycentric registers as the reference
 * holds them at the start of the group that contains the cut (accumulated rounding and in-place compaction included,
 * so the right piece neither replays the accumulator chains nor needs the lane map) and the depth outcomes of that
 * group's lanes left of the cut.  Hash table in HBM, one slot per cut: a 64-bit word
 * [63:12] key = record<<20 | source row<<7 | boundary index, [11] claimed, [3] ready, [2:0] outcomes, and 4 x float4
 * registers.  Boundary b of a span of row y: x = b*segWidth for b < numSegs, x = W for b = numSegs (row end),
 * x = W + (b-numSegs)*segWidth beyond (the part that wrapped into row y+1).  A consumer only ever waits for a warp with
 * a smaller ticket (left segment of its row, or the row below), and tickets are taken in launch order, so the wait
 * cannot deadlock. */
struct SLHand
{
    math::vec4 r[4];   /* registers at the start of the group that contains the cut */
    uint32_t fM;       /* outcomes of that group's lanes left of the cut            */
};

SLD_FN unsigned long long sl_handoff_key(uint32_t rec, int y, int boundary)
{
    return ((unsigned long long)rec << 20) | ((unsigned long long)(y & 0x1fff) << 7) | (unsigned long long)(boundary & 0x7f);
}
SLD_FN uint32_t sl_handoff_hash(unsigned long long key) { return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 36); }

SLD_FN void sl_handoff_publish(const SLRasterParams& P, uint32_t rec, int y, int boundary, uint32_t outcomes, const math::vec4& r0, const math::vec4& r1,
                               const math::vec4& r2, const math::vec4& r3)
{
    if (!P.handoff) { atomicOr(&P.counters->errorFlags, SL_ERR_HANDOFF_PUBLISH); return; } /* the count pass and the resolve pass disagree */
    const unsigned long long key = sl_handoff_key(rec, y, boundary);
    const unsigned long long word = (key << 12) | 0x800ull | (unsigned long long)(outcomes & 7u);
    uint32_t i = sl_handoff_hash(key) & P.handoffMask;
    while (atomicCAS(&P.handoff[i], 0ull, word) != 0ull) i = (i + 1u) & P.handoffMask;
    float4* dst = P.handoffRegs + (size_t)i * 4;
    __stcg(dst + 0, make_float4(r0.v[0], r0.v[1], r0.v[2], r0.v[3]));
    __stcg(dst + 1, make_float4(r1.v[0], r1.v[1], r1.v[2], r1.v[3]));
    __stcg(dst + 2, make_float4(r2.v[0], r2.v[1], r2.v[2], r2.v[3]));
    __stcg(dst + 3, make_float4(r3.v[0], r3.v[1], r3.v[2], r3.v[3]));
    __threadfence();
    atomicOr(&P.handoff[i], 0x8ull);
}

SLD_FN bool sl_handoff_poll(const SLRasterParams& P, uint32_t rec, int y, int boundary, SLHand& hand)
{
    if (!P.handoff) { atomicOr(&P.counters->errorFlags, SL_ERR_HANDOFF_POLL); return true; }
    const unsigned long long key = sl_handoff_key(rec, y, boundary);
    uint32_t i = sl_handoff_hash(key) & P.handoffMask;
    for (;;)
    {
        const unsigned long long w = *reinterpret_cast<volatile unsigned long long*>(&P.handoff[i]);
        if (w == 0ull) return false;
        if ((w >> 12) == key)
        {
            if (!(w & 0x8ull)) return false; /* claimed, registers not yet visible */
            __threadfence();
            const float4* src = P.handoffRegs + (size_t)i * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                const float4 f = __ldcg(src + k);
                hand.r[k] = math::vec4{f.x, f.y, f.z, f.w};
            }
            hand.fM = (uint32_t)(w & 7ull);
            return true;
        }
        i = (i + 1u) & P.handoffMask;
    }
}

constexpr uint32_t SL_STATE_IDENTITY = 0xE4u; /* lane map (0,1,2,3) */

/* the compaction of src/SL_TriRasterizer.cpp:496-498 on the lane map: t = outcomes of the four lanes (bit i = lane i) */
SLD_FN uint32_t sl_state_step(uint32_t st, uint32_t t)
{
    if (!t) return st;
    const uint32_t s3 = (st >> 6) & 3u;
    const uint32_t s2 = (t & 4u) ? ((st >> 4) & 3u) : s3;
    const uint32_t s1 = (t & 2u) ? ((st >> 2) & 3u) : s2;
    const uint32_t s0 = (t & 1u) ? (st & 3u) : s1;
    return s0 | (s1 << 2) | (s2 << 4) | (s3 << 6);
}

SLD_FN math::vec4 sl_pick4(uint32_t s, const math::vec4& a0, const math::vec4& a1, const math::vec4& a2, const math::vec4& a3)
{
    return s == 0u ? a0 : (s == 1u ? a1 : (s == 2u ? a2 : a3));
}

#ifndef SL_RASTER_WARPS_N
#define SL_RASTER_WARPS_N 4
#endif
constexpr int SL_RASTER_WARPS = SL_RASTER_WARPS_N;
#ifndef SL_RASTER_MIN_BLOCKS
#define SL_RASTER_MIN_BLOCKS 4
#endif
constexpr unsigned SL_PIECEREG_BYTES = 32u * 4u * (unsigned)sizeof(float4); /* per warp: the registers a piece continues from */
constexpr unsigned SL_WINSCRATCH_BYTES = 32u * 4u + 32u * 16u; /* head -> piece, piece book (+ 4 bytes per pixel of the segment: toucher masks) */

SLD_FN unsigned sl_align16(unsigned n) { return (n + 15u) & ~15u; }

/* warp-wide copy between a segment of an attachment and its shared-memory tile (16-byte accesses when both sides allow) */
SLD_FN void sl_copy_warp(unsigned char* dst, const unsigned char* src, unsigned bytes, int lane)
{
    const uintptr_t both = reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src);
    if ((both & 15u) == 0u)
    {
        const unsigned n16 = bytes >> 4;
        for (unsigned i = lane; i < n16; i += 32) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
        for (unsigned i = (n16 << 4) + lane; i < bytes; i += 32) dst[i] = src[i];
    }
    else if ((both & 3u) == 0u)
    {
        const unsigned n4 = bytes >> 2;
        for (unsigned i = lane; i < n4; i += 32) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
        for (unsigned i = (n4 << 2) + lane; i < bytes; i += 32) dst[i] = src[i];
    }
    else
    {
        for (unsigned i = lane; i < bytes; i += 32) dst[i] = src[i];
    }
}

/* k_resolve: one warp per (row, x-segment); work is handed out through a ticket in row-major order.  The warp stages
 * the depth texels of its segment into shared memory once, streams the row's ordered entry list and keeps the pieces
 * (spans clipped to the segment) that fall into its segment, 32 at a time.
 *
 * The pieces are resolved in *windows of 32 four-pixel groups*: lane l of a window works on one group of the 4-pixel
 * loop of render_triangle_simd (src/SL_TriRasterizer.cpp:471-548), groups of consecutive pieces packed side by side in
 * submission order (a long piece simply fills several windows).  Every lane sets up its triangle, brings the four
 * barycentric registers to its group (one rounded addition per group, :539-542) and then the window is settled in
 * waves that honour the two orders the reference defines:
 *   - pixel order: a lane may test and write depth only when no earlier lane of another piece that touches one of its
 *     pixels is still pending;
 *   - register order: the in-place compaction (:496-498) makes the registers of a group depend on the depth outcomes
 *     of the groups to its left.  What happened so far is a lane map (which initial register each register now holds;
 *     14 reachable states, (3,3,3,3) absorbing).  Lanes evaluate under the lane map of their piece's last settled
 *     group; a wave settles, per piece, the leading run of lanes up to and including the first one that changes it.
 * The earliest pending lane always settles, so a window needs at most as many waves as its longest dependency chain,
 * and a wave is a few dozen instructions.  Survivors are appended to the fragment lists in lane order = submission order.
 *
 * A piece that does not start its span continues from the registers handed over by the piece on its left (SLHand);
 * with the depth test off nothing ever waits: such a piece replays the accumulator chain from the span start instead
 * (every lane left of the cut passed, the registers were never compacted). */
__global__ void __launch_bounds__(SL_RASTER_WARPS * 32, SL_RASTER_MIN_BLOCKS) k_resolve(const __grid_constant__ SLRasterParams P)
{
    extern __shared__ __align__(16) unsigned char sDyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ltMask = (1u << lane) - 1u;
    unsigned dbgWindows = 0, dbgWaves = 0, dbgIdle = 0, dbgGroups = 0;
    /* One ticket per warp.  (Persistent warps that loop over tickets were measured: slower on C2/C3/C4 -- more
     * co-resident working warps do not raise this kernel's throughput -- so the simple form stays.) */
    {
    if (P.counters->abortBatch) return; /* the batch did not fit: it is run again, nothing visible may be written */
    unsigned ticket = 0;
    if (lane == 0) ticket = atomicAdd(&P.counters->rasterTicket, 1u);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    const int NS = (int)P.numSegs;
    if (ticket >= P.numRows * P.numSegs) return;
    /* Work order.  Row-major by default.  When no span of the batch runs past a row end, the only hand-overs are from
     * a segment to its right neighbour, and the host hands out a skewed order (row y of segment s together with row
     * y + s*D of segment 0, D*numSegs >= the number of resident warps): a piece's left neighbour then took its
     * ticket a whole machine-load earlier and its hand-over is normally there before it is asked for. */
    int y = (int)(ticket / (unsigned)NS), seg = (int)(ticket % (unsigned)NS);
    if (P.ticketOrder && P.counters->rowEndCuts == 0ull) /* the skew assumes that hand-overs only go to the right neighbour */
    {
        const uint32_t o = P.ticketOrder[ticket];
        y = (int)(o >> 6);
        seg = (int)(o & 63u);
    }
    if (P.rowNeeded && !P.rowNeeded[y]) return;
    const int W = (int)P.width;
    const int sx0 = seg * (int)P.segWidth;
    const int sx1 = (seg == NS - 1) ? W : sx0 + (int)P.segWidth;
    const bool chained = P.depthTest != 0;
    const uint32_t begin = P.rowStart[y], end = P.rowStart[y + 1];
    if (begin == end) return;
    unsigned shaded = 0;
    dbgWindows = 0; dbgWaves = 0; dbgIdle = 0; dbgGroups = 0;
    const long long dbgT0 = P.diagnostics ? clock64() : 0ll;
    const bool useDepth = P.depthTest != 0 || P.depthMask != 0;

    const unsigned pixMaskBytes = sl_align16(4u * P.segWidth);
    unsigned char* warpMem = sDyn + (size_t)warp * (SL_PIECEQ_BYTES + SL_SUBLIST_BYTES + SL_PIECEREG_BYTES + SL_WINSCRATCH_BYTES + pixMaskBytes + P.tileBytes);
    SLSpanEntry* queue = reinterpret_cast<SLSpanEntry*>(warpMem);
    SLListState L;
    L.tail = reinterpret_cast<uint32_t*>(warpMem + SL_PIECEQ_BYTES);
    L.fill = L.tail + SL_MAX_SUBLISTS;
    L.total = L.fill + SL_MAX_SUBLISTS;
    L.nsub = (int)P.numSubLists;
    L.firstList = (uint32_t)(y * NS + seg) * P.numSubLists;
    if (lane < SL_MAX_SUBLISTS)
    {
        L.tail[lane] = SL_NO_CHUNK;
        L.fill[lane] = 32u;
        L.total[lane] = 0u;
    }
    float4* pieceRegs = reinterpret_cast<float4*>(warpMem + SL_PIECEQ_BYTES + SL_SUBLIST_BYTES);
    uint32_t* slotPiece = reinterpret_cast<uint32_t*>(warpMem + SL_PIECEQ_BYTES + SL_SUBLIST_BYTES + SL_PIECEREG_BYTES);
    int4* pieceBook = reinterpret_cast<int4*>(slotPiece + 32);
    uint32_t* pixMask = reinterpret_cast<uint32_t*>(pieceBook + 32); /* [segWidth] lanes of the current window that touch the pixel */
    unsigned char* tileMem = warpMem + SL_PIECEQ_BYTES + SL_SUBLIST_BYTES + SL_PIECEREG_BYTES + SL_WINSCRATCH_BYTES + pixMaskBytes;
    const unsigned segPixels = (unsigned)(sx1 - sx0);
    if (useDepth)
    {
        for (unsigned i = lane; i < P.segWidth; i += 32) pixMask[i] = 0u;
    }
    const long long segOrigin = (long long)W * y + sx0;
    SLTile T;
    bool staged = false;
    if (P.tileBytes && useDepth)
    {
        T.depth = tileMem;
        T.origin = segOrigin;
    }
    else
    {
        T.depth = P.depth;
        T.origin = 0;
    }

    uint32_t base = begin;
    while (base < end)
    {
        /* ---- collect up to 32 pieces of this segment, in list order (four 32-entry batches in flight) ---- */
        int qn = 0;
        bool full = false;
        while (base < end && !full)
        {
            SLSpanEntry c[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                const uint32_t idx = base + 32u * k + lane;
                c[k] = SLSpanEntry{0u, 0, 0};
                if (idx < end) c[k] = P.entries[idx];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                if (full || base >= end) continue;
                const int cb = (c[k].rec & SL_ENTRY_OVERFLOW) ? W : 0;
                const bool cHas = max((int)c[k].xMin, cb + sx0) < min((int)c[k].xMax, cb + sx1); /* empty entries: 0 < 0 */
                const uint32_t hm = __ballot_sync(0xffffffffu, cHas);
                const int cnt = __popc(hm);
                if (qn + cnt > 32) { full = true; continue; } /* process what is queued, then come back to this batch */
                if (cHas)
                {
                    queue[qn + __popc(hm & ltMask)] = c[k];
                    /* the first access of the window that holds the piece: pull the triangle record towards the SM now */
                    const float4* r = P.records + (size_t)(c[k].rec & ~SL_ENTRY_OVERFLOW) * P.recStride;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(r));
                }
                qn += cnt;
                base += 32;
            }
        }
        __syncwarp();
        if (!qn) continue;
        if (P.tileBytes && useDepth && !staged)
        {
            /* stage the depth segment in: one coalesced read */
            sl_copy_warp(T.depth, P.depth + (size_t)segOrigin * P.depthBpt, segPixels * (unsigned)P.depthBpt, lane);
            staged = true;
            __syncwarp();
        }

        /* ---- lane i keeps the book for piece i of the queue ---- */
        SLSpanEntry e = SLSpanEntry{0u, 0, 0};
        if (lane < qn) e = queue[lane];
        __syncwarp();
        const bool isOver = (e.rec & SL_ENTRY_OVERFLOW) != 0u;
        const uint32_t recId = e.rec & ~SL_ENTRY_OVERFLOW;
        const int xBase = isOver ? W : 0; /* the entry's x coordinates are those of its source row */
        const int ySrc = isOver ? y - 1 : y;
        const int xMin = e.xMin, xMax = e.xMax;
        const int p0 = max(xMin, xBase + sx0), p1 = min(xMax, xBase + sx1);
        const bool has = lane < qn && p0 < p1;
        const bool needsIn = has && xMin < p0 && chained;
        const bool publish = has && chained && xMax > p1 && (seg < NS - 1 || (!isOver && y + 1 < (int)P.numRows));
        const int bIn = (isOver ? NS : 0) + seg;
        int gNext = has ? ((p0 - xMin) >> 2) : 0;               /* next group of the piece to resolve            */
        const int gEnd = has ? (((p1 - 1 - xMin) >> 2) + 1) : 0; /* one past its last group                       */
        bool started = false;                                   /* pieceRegs[lane] holds the registers at gBase  */
        int gBase = 0;
        uint32_t pfM = 0u;
        unsigned idle = 0;

        for (;;)
        {
            const int cnt = gEnd - gNext;
            const uint32_t nz = __ballot_sync(0xffffffffu, cnt > 0);
            if (!nz) break;
            /* pieces that continue a span of the segment on the left (or of the row below) need its hand-over first */
            bool ready = true;
            if (cnt > 0 && needsIn && !started)
            {
                SLHand h;
                ready = sl_handoff_poll(P, recId, ySrc, bIn, h);
                if (ready)
                {
#pragma unroll
                    for (int k = 0; k < 4; ++k) pieceRegs[lane * 4 + k] = make_float4(h.r[k].v[0], h.r[k].v[1], h.r[k].v[2], h.r[k].v[3]);
                    pfM = h.fM;
                    gBase = gNext;
                    started = true;
                }
            }
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const int start = incl - cnt; /* lane of the window this piece's next group would get */
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t notReady = __ballot_sync(0xffffffffu, cnt > 0 && !ready);
            int limit = min(32, total);
            if (notReady) limit = min(limit, __shfl_sync(0xffffffffu, start, __ffs(notReady) - 1));
            if (limit == 0)
            {
                /* the first unresolved piece waits for its hand-over: back off so that the polling does not take issue
                 * slots from the warps that produce it */
                idle = min(idle + 1u, 4u);
                __nanosleep(20u << idle);
                ++dbgIdle;
                continue;
            }
            idle = 0;
            ++dbgWindows;
            dbgGroups += (unsigned)limit;
            __syncwarp();

            /* ---- this lane as a worker: group `lane` of the window ---- */
            const bool headHere = cnt > 0 && start < limit;
            const uint32_t heads = __ballot_sync(0xffffffffu, headHere); /* piece lanes whose piece has groups in the window */
            if (headHere)
            {
                slotPiece[start] = (uint32_t)lane;
                pieceBook[lane] = make_int4(start, gNext, gBase, (int)(pfM | (started ? 8u : 0u) | (publish ? 16u : 0u)));
            }
            __syncwarp();
            const bool act = lane < limit;
            /* the piece of this group: the last piece that starts at or before this lane */
            uint32_t startBits = 0u;
            {
                /* bit k set when some piece of the window starts at lane k */
                const uint32_t mine = headHere ? (1u << start) : 0u;
                startBits = __reduce_or_sync(0xffffffffu, mine);
            }
            const int headLane = 31 - __clz((int)(startBits & (0xffffffffu >> (31 - lane))));
            const int pi = act ? (int)slotPiece[headLane] : 0; /* the lane that keeps this group's piece */
            const SLSpanEntry we = queue[pi];
            const int4 wb = pieceBook[pi];
            const bool wOver = (we.rec & SL_ENTRY_OVERFLOW) != 0u;
            const uint32_t wRec = we.rec & ~SL_ENTRY_OVERFLOW;
            const int wXBase = wOver ? W : 0;
            const int wY = wOver ? y - 1 : y;
            const int wXMin = we.xMin;
            const int wP0 = max(wXMin, wXBase + sx0), wP1 = min((int)we.xMax, wXBase + sx1);
            const int wStart = wb.x, wGNext = wb.y, wGBase = wb.z;
            const uint32_t wFM = (uint32_t)wb.w & 7u;
            const bool wStarted = ((uint32_t)wb.w & 8u) != 0u, wPublish = ((uint32_t)wb.w & 16u) != 0u;
            const int wGEnd = ((wP1 - 1 - wXMin) >> 2) + 1;
            const int wOut = (wOver ? NS : 0) + seg + 1;
            const int g = wGNext + (lane - wStart);
            const int x4 = wXMin + 4 * g;
            (void)heads;

            const float4* rec = P.records + (size_t)wRec * P.recStride;
            SLTriSetup t;
            sl_tri_setup(t, rec[0], rec[1], rec[2]);
            const math::vec4 bcX = t.bc0 * 4.f;
            math::vec4 a0, a1, a2, a3;
            int nAdd;
            uint32_t fM = 0u; /* outcomes of this group's lanes that another piece owns (left of the cut) */
            if (wStarted)
            {
                const float4 r0 = pieceRegs[pi * 4 + 0], r1 = pieceRegs[pi * 4 + 1], r2 = pieceRegs[pi * 4 + 2], r3 = pieceRegs[pi * 4 + 3];
                a0 = math::vec4{r0.x, r0.y, r0.z, r0.w};
                a1 = math::vec4{r1.x, r1.y, r1.z, r1.w};
                a2 = math::vec4{r2.x, r2.y, r2.z, r2.w};
                a3 = math::vec4{r3.x, r3.y, r3.z, r3.w};
                nAdd = g - wGBase;
                if (g == wGBase) fM = wFM;
            }
            else
            {
                const float yf = (float)wY;
                const math::vec4 bcY = math::fmadd(t.bc1, math::vec4{yf}, t.bc2); /* :474 */
                a0 = (t.bc0 * (float)(wXMin + 0)) + bcY;                          /* :476-481 */
                a1 = (t.bc0 * (float)(wXMin + 1)) + bcY;
                a2 = (t.bc0 * (float)(wXMin + 2)) + bcY;
                a3 = (t.bc0 * (float)(wXMin + 3)) + bcY;
                nAdd = g;
                /* depth test off and the piece starts inside its span: the lanes left of the cut passed */
                if (wXMin < wP0 && g == ((wP0 - wXMin) >> 2)) fM = ((1u << (wP0 - x4)) - 1u) & 7u;
            }
            if (!act) nAdd = 0;
            for (int i = 0; i < nAdd; ++i) { a0 += bcX; a1 += bcX; a2 += bcX; a3 += bcX; } /* :539-542, one rounding per group */

            /* depth of the four registers' initial contents at this group (_sl_mul_vec4_mat4_ps, :376-393) */
            const float zc0 = ((a0.v[3] * 0.f) + (a0.v[2] * t.z2)) + ((a0.v[1] * t.z1) + (a0.v[0] * t.z0));
            const float zc1 = ((a1.v[3] * 0.f) + (a1.v[2] * t.z2)) + ((a1.v[1] * t.z1) + (a1.v[0] * t.z0));
            const float zc2 = ((a2.v[3] * 0.f) + (a2.v[2] * t.z2)) + ((a2.v[1] * t.z1) + (a2.v[0] * t.z0));
            const float zc3 = ((a3.v[3] * 0.f) + (a3.v[2] * t.z2)) + ((a3.v[1] * t.z1) + (a3.v[0] * t.z0));
            const int rowBase = (int)((long long)W * (long long)wY - T.origin);
            uint32_t ownBits = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                if (act && x4 + j >= wP0 && x4 + j < wP1) ownBits |= 1u << j;
            }
            /* the earlier lanes of other pieces that touch one of this group's pixels: every lane adds itself to the
             * toucher mask of its pixels, reads them back, and clears them again for the next window */
            const uint32_t before = ltMask & ~((1u << wStart) - 1u); /* earlier lanes of my piece in this window */
            uint32_t ov = 0u;
            if (useDepth)
            {
                const int px = x4 - wXBase - sx0; /* segment-relative x of the group's first pixel */
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    if ((ownBits >> j) & 1u) atomicOr(&pixMask[px + j], 1u << lane);
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    if ((ownBits >> j) & 1u) ov |= pixMask[px + j];
                }
                ov &= ltMask & ~before;
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    if ((ownBits >> j) & 1u) pixMask[px + j] = 0u;
                }
            }

            /* ---- waves ---- */
            bool pending = act;
            uint32_t stCur = SL_STATE_IDENTITY; /* lane map after my piece's last settled group (relative to a0..a3) */
            uint32_t stIn = SL_STATE_IDENTITY, mBits = 0u, tBits = 0u;
            for (;;)
            {
                const uint32_t PEND = __ballot_sync(0xffffffffu, pending);
                if (!PEND) break;
                ++dbgWaves;
                const bool dSafe = pending && (ov & PEND) == 0u;
                uint32_t m = 0u;
                if (dSafe)
                {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        if ((ownBits >> j) & 1u)
                        {
                            const uint32_t sj = (stCur >> (2 * j)) & 3u;
                            const float z = sj == 0u ? zc0 : (sj == 1u ? zc1 : (sj == 2u ? zc2 : zc3));
                            const float d = P.depthTest ? sl_load_depth(T.depth, P.depthBpt, (size_t)(rowBase + x4 + j)) : 0.f;
                            if (sl_depth_cmp(P.depthTest, z, d)) m |= 1u << j;
                        }
                    }
                }
                const uint32_t tt = m | fM;
                const uint32_t stOut = sl_state_step(stCur, tt);
                const uint32_t D = __ballot_sync(0xffffffffu, dSafe);
                const uint32_t CH = __ballot_sync(0xffffffffu, dSafe && stOut != stCur);
                const bool commit = dSafe && (((~D) | CH) & PEND & before) == 0u;
                if (commit)
                {
                    stIn = stCur;
                    mBits = m;
                    tBits = tt;
                    pending = false;
                    if (P.depthMask && m) /* src/SL_FragmentProcessor.cpp:526-529 */
                    {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                        {
                            if ((m >> j) & 1u)
                            {
                                const uint32_t sj = (stCur >> (2 * j)) & 3u;
                                const float z = sj == 0u ? zc0 : (sj == 1u ? zc1 : (sj == 2u ? zc2 : zc3));
                                sl_store_depth(T.depth, P.depthBpt, (size_t)(rowBase + x4 + j), z);
                            }
                        }
                    }
                }
                /* the still pending lanes of a piece go on under the lane map its last settled lane left behind */
                const uint32_t C = __ballot_sync(0xffffffffu, commit);
                const uint32_t cand = C & before;
                const int src = cand ? (31 - __clz(cand)) : lane;
                const uint32_t sNew = __shfl_sync(0xffffffffu, stOut, src);
                if (pending && cand) stCur = sNew;
                __syncwarp(); /* depth stores of this wave before the loads of the next */
            }

            /* ---- the piece's last group: hand over to the piece on the right; a piece cut by the window end: carry on ---- */
            const uint32_t stAfter = sl_state_step(stIn, tBits);
            if (act && g == wGEnd - 1)
            {
                if (wPublish)
                {
                    if (((wP1 - wXMin) >> 2) == g)
                    {
                        /* the cut lies inside this group: the registers before its compaction + outcomes left of the cut */
                        const uint32_t below = (1u << (wP1 - x4)) - 1u;
                        sl_handoff_publish(P, wRec, wY, wOut, tBits & below & 7u, sl_pick4(stIn & 3u, a0, a1, a2, a3), sl_pick4((stIn >> 2) & 3u, a0, a1, a2, a3),
                                           sl_pick4((stIn >> 4) & 3u, a0, a1, a2, a3), sl_pick4((stIn >> 6) & 3u, a0, a1, a2, a3));
                    }
                    else
                    {
                        /* the cut falls on the group boundary: the registers after the compaction, advanced to the next group */
                        sl_handoff_publish(P, wRec, wY, wOut, 0u, sl_pick4(stAfter & 3u, a0, a1, a2, a3) + bcX, sl_pick4((stAfter >> 2) & 3u, a0, a1, a2, a3) + bcX,
                                           sl_pick4((stAfter >> 4) & 3u, a0, a1, a2, a3) + bcX, sl_pick4((stAfter >> 6) & 3u, a0, a1, a2, a3) + bcX);
                    }
                }
            }
            else if (act && lane == limit - 1)
            {
                const math::vec4 c0 = sl_pick4(stAfter & 3u, a0, a1, a2, a3) + bcX, c1 = sl_pick4((stAfter >> 2) & 3u, a0, a1, a2, a3) + bcX;
                const math::vec4 c2 = sl_pick4((stAfter >> 4) & 3u, a0, a1, a2, a3) + bcX, c3 = sl_pick4((stAfter >> 6) & 3u, a0, a1, a2, a3) + bcX;
                pieceRegs[pi * 4 + 0] = make_float4(c0.v[0], c0.v[1], c0.v[2], c0.v[3]);
                pieceRegs[pi * 4 + 1] = make_float4(c1.v[0], c1.v[1], c1.v[2], c1.v[3]);
                pieceRegs[pi * 4 + 2] = make_float4(c2.v[0], c2.v[1], c2.v[2], c2.v[3]);
                pieceRegs[pi * 4 + 3] = make_float4(c3.v[0], c3.v[1], c3.v[2], c3.v[3]);
            }

            /* ---- survivors, in lane order: number s is shaded with register s after the compaction = the initial
             * register of the first passing lane >= s (:500-526), and with its own x and z ---- */
            math::vec4 bcJ[4];
            float zJ[4];
            if (mBits)
            {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    const uint32_t sIdx = (uint32_t)__popc(tBits & ((1u << j) - 1u));
                    const uint32_t ge = tBits & ~((1u << sIdx) - 1u);
                    const int pLane = __ffs(ge) - 1;
                    bcJ[j] = sl_pick4((stIn >> (2 * (pLane & 3))) & 3u, a0, a1, a2, a3);
                    const uint32_t zj = (stIn >> (2 * j)) & 3u;
                    zJ[j] = zj == 0u ? zc0 : (zj == 1u ? zc1 : (zj == 2u ? zc2 : zc3));
                }
                shaded += (unsigned)__popc(mBits);
            }
            sl_push_frags(P, L, lane, mBits, bcJ, zJ, x4, x4 - wXBase - sx0, wY, wRec);

            /* ---- book-keeping of the pieces ---- */
            const int consumed = min(cnt, max(0, limit - start));
            if (consumed > 0)
            {
                gNext += consumed;
                if (gNext < gEnd)
                {
                    started = true; /* pieceRegs[lane] was written by the worker of the last consumed group */
                    gBase = gNext;
                    pfM = 0u;
                }
            }
            __syncwarp();
        }
    }
    __syncwarp();
    if (lane < L.nsub && L.total[lane]) { P.lists.listCount[L.firstList + lane] = L.total[lane]; P.lists.listTail[L.firstList + lane] = L.tail[lane]; }
    const bool dirty = __any_sync(0xffffffffu, shaded != 0u);
    if (staged && dirty && P.depthMask)
    {
        /* write the depth segment back: one coalesced store */
        __syncwarp();
        sl_copy_warp(P.depth + (size_t)segOrigin * P.depthBpt, T.depth, segPixels * (unsigned)P.depthBpt, lane);
    }
    for (int d = 16; d > 0; d >>= 1) shaded += __shfl_down_sync(0xffffffffu, shaded, d);
    if (lane == 0 && shaded) atomicAdd(&P.counters->fragsShaded, (unsigned long long)shaded);
    if (lane == 0 && P.diagnostics)
    {
        const unsigned long long cyc = (unsigned long long)(clock64() - dbgT0);
        atomicAdd(&P.counters->resolveWindows, (unsigned long long)dbgWindows);
        atomicAdd(&P.counters->resolveWaves, (unsigned long long)dbgWaves);
        atomicAdd(&P.counters->resolveIdle, (unsigned long long)dbgIdle);
        atomicAdd(&P.counters->resolveGroups, (unsigned long long)dbgGroups);
        atomicAdd(&P.counters->resolveCycles, cyc);
        atomicMax(&P.counters->resolveMaxCycles, cyc);
        atomicMax(&P.counters->resolveMaxWindows, (unsigned long long)dbgWindows);
        atomicAdd(&P.counters->resolveWarps, 1ull);
    }
    }
}

/* ------------------------------------------------------------------------
 * k_shade<FS>: SL_FragmentProcessor::flush_tri_fragments (src/SL_FragmentProcessor.cpp:416-531) + assign_pixel /
 * assign_alpha_pixel (src/SL_Framebuffer.cpp:44-368) over the fragment lists.  One warp per list.
 * ----------------------------------------------------------------------*/
constexpr int SL_SHADE_WARPS = 4;
#ifndef SL_SHADE_MIN_BLOCKS
#define SL_SHADE_MIN_BLOCKS 8 /* 64 registers: measured best on every config (C3 shade 0.63 -> 0.54 ms) */
#endif

/* one fragment of a chunk: perspective correction (:438-460), varyings (interpolate_tri_varyings, AVX2 branch :163-193 --
 * the reference interpolates all four slots; slots past numVaryings hold stale bin data there and are never read by a
 * well-formed shader), the shader call (:501).  Returns the texel index. */
template <FsFn FS>
SLD_FN uint32_t sl_shade_fragment(const SLShadeParams& P, const unsigned char* c, int slot, SL_FragmentParam& fp, bool& haveOutputs)
{
    const float4 b4 = reinterpret_cast<const float4*>(c)[slot];
    const float z = reinterpret_cast<const float*>(c + 512)[slot];
    const uint32_t xy = reinterpret_cast<const uint32_t*>(c + 640)[slot];
    const uint32_t recId = reinterpret_cast<const uint32_t*>(c + 768)[slot];
    const int x = (int)(xy & 0xffffu), y = (int)(xy >> 16);
    const float4* rec = P.records + (size_t)recId * P.recStride;
    math::vec4 bc = math::vec4{b4.x, b4.y, b4.z, b4.w} * math::vec4{rec[0].w, rec[1].w, rec[2].w, 0.f};
    const float persp = sl_rcp((bc.v[0] + bc.v[1]) + (bc.v[2] + bc.v[3]));
    bc = bc * persp;
    fp.pUniforms = &P.draws[P.recDraw[recId]].uniforms;
    for (int s = 0; s < P.numVaryings; ++s)
    {
        const float4 a = rec[3 + 3 * s + 0], b = rec[3 + 3 * s + 1], cc = rec[3 + 3 * s + 2];
        fp.pVaryings[s] = math::vec4{
            sl_fma(bc.v[2], cc.x, sl_fma(bc.v[1], b.x, bc.v[0] * a.x)), sl_fma(bc.v[2], cc.y, sl_fma(bc.v[1], b.y, bc.v[0] * a.y)),
            sl_fma(bc.v[2], cc.z, sl_fma(bc.v[1], b.z, bc.v[0] * a.z)), sl_fma(bc.v[2], cc.w, sl_fma(bc.v[1], b.w, bc.v[0] * a.w))};
    }
    fp.coord.x = (uint16_t)x;
    fp.coord.y = (uint16_t)y;
    fp.coord.depth = z;
#pragma unroll
    for (int i = 0; i < 4; ++i) fp.pOutputs[i] = math::vec4{0.f};
    haveOutputs = FS(fp);
    /* every attachment is addressed as x + width*y with the depth attachment's width (the reference uses each colour view's
     * own width; SL_Framebuffer::valid requires them to be equal, and run_draw refuses anything else); x, y are relative to
     * the span's source row, so fragments that ran past the row end land at the start of the next row exactly like in the
     * reference */
    return (uint32_t)x + P.width * (uint32_t)y;
}

/* Without blending a texel ends up with the colour of the LAST fragment whose shader produced outputs; a fragment shader
 * has no other effect (depth is written by the resolve stage).  So an opaque list is walked from its end: per chunk, one
 * cheap pass over the texel indices finds the fragments that are the last ones of their texel (a bit per texel of the
 * list's 128-texel window remembers what is already settled), they are collected, and shaded 32 at a time with every
 * lane busy -- the fragments under them are never shaded at all (C3: 11.5 M fragments pass the depth test, 8.3 M texels
 * show one).  A shader that returns false for one of them leaves the texel to an earlier fragment after all: rare, so the
 * warp then simply runs the whole list again in submission order (below), which overwrites what the fast pass stored. */
template <FsFn FS>
__global__ void __launch_bounds__(SL_SHADE_WARPS * 32, SL_SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ SLShadeParams P)
{
    __shared__ uint32_t sSettled[SL_SHADE_WARPS][4];
    __shared__ uint32_t sQueue[SL_SHADE_WARPS][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (P.abortBatch && *P.abortBatch) return; /* the batch did not fit: it is run again, nothing visible may be written */
    /* persistent warps: lists are handed out through a ticket (a block would otherwise live as long as its longest list) */
    for (;;)
    {
    unsigned list = 0;
    if (lane == 0) list = P.listBegin + atomicAdd(P.shadeTicket, 1u);
    list = __shfl_sync(0xffffffffu, list, 0);
    if (list >= P.numLists) break;
    const int count = (int)P.lists.listCount[list];
    if (count <= 0) continue;
    bool inOrder = !P.opaqueSkip;
    if (!inOrder)
    {
        const unsigned perRow = P.numSegs * P.numSubLists;
        const unsigned row = list / perRow, rem = list - row * perRow;
        const unsigned seg = rem / P.numSubLists, sub = rem - seg * P.numSubLists;
        const uint32_t listBase = P.width * row + seg * P.segWidth + (sub << SL_SHADE_SHIFT_C);
        if (lane < 4) sSettled[warp][lane] = 0u;
        __syncwarp();
        uint32_t* queue = sQueue[warp];
        unsigned qn = 0u;
        bool redo = false;
        uint32_t chunk = P.lists.listTail[list];
        for (int remaining = count; remaining > 0 || qn > 0u;)
        {
            if (remaining > 0)
            {
                const int n = ((remaining - 1) & 31) + 1; /* the last chunk is the partly filled one */
                const unsigned char* c = P.lists.pool + (size_t)chunk * SL_CHUNK_BYTES;
                uint32_t key = 0xffffff00u + (uint32_t)lane; /* idle lanes: a key no texel has */
                bool alive = false;
                if (lane < n)
                {
                    const uint32_t xy = reinterpret_cast<const uint32_t*>(c + 640)[lane];
                    key = (xy & 0xffffu) + P.width * (xy >> 16);
                    const uint32_t rel = key - listBase;
                    if (rel >= (1u << SL_SHADE_SHIFT_C)) redo = true; /* not in this list's window: cannot happen; the in-order pass is exact regardless */
                    else alive = !((sSettled[warp][rel >> 5] >> (rel & 31u)) & 1u);
                }
                const uint32_t peers = __match_any_sync(0xffffffffu, alive ? key : 0xffffff00u + (uint32_t)lane);
                const bool last = alive && (peers >> lane) == 1u; /* no later fragment of the chunk on the same texel */
                if (last) atomicOr(&sSettled[warp][(key - listBase) >> 5], 1u << ((key - listBase) & 31u));
                const uint32_t m = __ballot_sync(0xffffffffu, last);
                if (last) queue[qn + (unsigned)__popc(m & ((1u << lane) - 1u))] = (chunk << 5) | (uint32_t)lane;
                qn += (unsigned)__popc(m);
                remaining -= n;
                if (remaining > 0) chunk = P.lists.chunkPrev[chunk];
                __syncwarp();
            }
            if (qn >= 32u || (remaining <= 0 && qn > 0u))
            {
                const unsigned take = qn < 32u ? qn : 32u;
                bool haveOutputs = true;
                if ((unsigned)lane < take)
                {
                    const uint32_t e = queue[qn - take + (unsigned)lane]; /* the newest ones: the rest stays where it is */
                    SL_FragmentParam fp;
                    const uint32_t key = sl_shade_fragment<FS>(P, P.lists.pool + (size_t)(e >> 5) * SL_CHUNK_BYTES, (int)(e & 31u), fp, haveOutputs);
                    if (haveOutputs)
                    {
                        for (int i = P.numOutputs - 1; i >= 0; --i) sl_store_color(P.color[i], P.colorType[i], (size_t)key, fp.pOutputs[i]);
                    }
                }
                if (__any_sync(0xffffffffu, !haveOutputs)) redo = true;
                qn -= take;
                __syncwarp();
            }
        }
        inOrder = __any_sync(0xffffffffu, redo);
        if (inOrder) __syncwarp();
    }
    if (!inOrder) continue;
    int remaining = count;
    uint32_t chunk = P.lists.listHead[list];
    for (; remaining > 0; remaining -= 32)
    {
        const int n = min(remaining, 32);
        const unsigned char* c = P.lists.pool + (size_t)chunk * SL_CHUNK_BYTES;
        if (remaining > 32) chunk = P.lists.chunkNext[chunk];
        SL_FragmentParam fp;
        bool haveOutputs = false;
        uint32_t key = 0xffffff00u + (uint32_t)lane; /* idle lanes: a key no pixel has */
        if (lane < n) key = sl_shade_fragment<FS>(P, c, lane, fp, haveOutputs);
        /* :507-523, in submission order per pixel */
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        if (P.blend == 0)
        {
            /* plain stores: the last fragment of a pixel that has outputs wins */
            const uint32_t hv = __ballot_sync(0xffffffffu, haveOutputs) & peers;
            if (haveOutputs && (hv >> lane) == 1u)
            {
                for (int i = P.numOutputs - 1; i >= 0; --i) sl_store_color(P.color[i], P.colorType[i], (size_t)key, fp.pOutputs[i]);
            }
        }
        else
        {
            const int rank = __popc(peers & ((1u << lane) - 1u));
            const int maxRank = (int)__reduce_max_sync(0xffffffffu, (unsigned)rank);
            for (int r = 0; r <= maxRank; ++r)
            {
                if (rank == r && haveOutputs)
                {
                    for (int i = P.numOutputs - 1; i >= 0; --i) sl_blend_color(P.color[i], P.colorType[i], (size_t)key, fp.pOutputs[i], P.blend);
                }
                __syncwarp();
            }
        }
        __syncwarp(); /* the next step's loads of a pixel come after this step's stores to it */
    }
    } /* next list */
}

/* ------------------------------------------------------------------------
 * SL_ClearProcessor::clear_texture (src/SL_ClearProcessor.cpp:23-63): typed fill
 * ----------------------------------------------------------------------*/
struct SLClearParams
{
    unsigned char* dst;
    unsigned long long numTexels;
    int bpt;
    unsigned char texel[32];
};

__global__ void k_clear(const __grid_constant__ SLClearParams P)
{
    /* 16-byte stores: a 16-byte pattern covers bpt in {1,2,4,8,16}; other sizes go texel by texel */
    const unsigned long long totalBytes = P.numTexels * (unsigned long long)P.bpt;
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    if (16 % P.bpt == 0)
    {
        uint4 pat;
        unsigned char* pb = reinterpret_cast<unsigned char*>(&pat);
        for (int i = 0; i < 16; ++i) pb[i] = P.texel[i % P.bpt];
        const unsigned long long n16 = totalBytes / 16ull;
        uint4* d16 = reinterpret_cast<uint4*>(P.dst);
        for (unsigned long long i = tid; i < n16; i += stride) d16[i] = pat;
        for (unsigned long long i = n16 * 16ull + tid; i < totalBytes; i += stride) P.dst[i] = P.texel[i % P.bpt];
    }
    else
    {
        for (unsigned long long i = tid; i < totalBytes; i += stride) P.dst[i] = P.texel[i % P.bpt];
    }
}

/* ------------------------------------------------------------------------
 * SL_BlitProcessor::blit_nearest (src/SL_BlitProcessor.cpp:620-667): nearest
 * neighbour, 16.16 fixed point, vertically flipped, format converting.  One
 * thread per destination texel.
 *
 * The texel conversions (all 33 x 33 type pairs) are in sl_texel.cuh.
 * ----------------------------------------------------------------------*/
struct SLBlitParams
{
    const unsigned char* src;
    unsigned char* dst;
    unsigned int srcW, srcH, dstW, dstH;
    unsigned int srcX0, srcY0, srcX1, srcY1, dstX0, dstY0, dstX1, dstY1;
    int mode, bpt;         /* mode 0: texel copy (bpt bytes), 1: RGBA_FLOAT -> RGBA_8U, 2: RGBA_8U -> RGBA_FLOAT, 3: general */
    int srcType, dstType;  /* SL_ColorDataType */
};

__global__ void __launch_bounds__(256) k_blit(const __grid_constant__ SLBlitParams P)
{
    const unsigned long long inW = (unsigned long long)P.srcX1 - P.srcX0, inH = (unsigned long long)P.srcY1 - P.srcY0;
    const unsigned long long outW = (unsigned long long)P.dstX1 - P.dstX0;
    const unsigned long long x0 = P.dstX0;
    const unsigned long long x1 = min((unsigned long long)P.dstW, x0 + outW);
    const unsigned long long foutW = ((inW << 16) / P.dstW) + 1ull; /* relative to the WHOLE destination, like the reference */
    const unsigned long long foutH = ((inH << 16) / P.dstH) + 1ull;
    const unsigned long long x = x0 + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= x1) return;
    const int srcBpt = sl_texel_bytes(P.srcType);
    for (unsigned long long y = (unsigned long long)P.dstY0 + blockIdx.y; y < P.dstY1 && y < P.dstH; y += gridDim.y)
    {
        const unsigned long long yf = (y * foutH) >> 16;
        const unsigned long long srcY = (unsigned long long)P.srcY1 - ((unsigned long long)P.srcY0 + yf) - 1ull;
        const unsigned long long srcX = (x * foutW) >> 16;
        if (srcX >= P.srcW || srcY >= P.srcH) continue; /* the reference would read outside the source here */
        const size_t si = (size_t)(srcX + (unsigned long long)P.srcW * srcY), di = (size_t)(x + (unsigned long long)P.dstW * y);
        if (P.mode == 1)
        {
            const float4 f = reinterpret_cast<const float4*>(P.src)[si];
            const math::vec4_t<uint8_t> r = math::vec4_to_u8(math::vec4{f.x, f.y, f.z, f.w} * 255.f);
            reinterpret_cast<uint32_t*>(P.dst)[di] = (uint32_t)r.v[0] | ((uint32_t)r.v[1] << 8) | ((uint32_t)r.v[2] << 16) | ((uint32_t)r.v[3] << 24);
        }
        else if (P.mode == 2)
        {
            const uint32_t u = reinterpret_cast<const uint32_t*>(P.src)[si];
            const math::vec4 c = math::vec4{(float)(u & 0xffu), (float)((u >> 8) & 0xffu), (float)((u >> 16) & 0xffu), (float)(u >> 24)} * 0.00392156862745f;
            reinterpret_cast<float4*>(P.dst)[di] = make_float4(c.v[0], c.v[1], c.v[2], c.v[3]);
        }
        else if (P.mode == 3) sl_blit_texel(P.src + si * srcBpt, P.srcType, P.dst + di * P.bpt, P.dstType);
        else if (P.bpt == 4) reinterpret_cast<uint32_t*>(P.dst)[di] = reinterpret_cast<const uint32_t*>(P.src)[si];
        else if (P.bpt == 16) reinterpret_cast<uint4*>(P.dst)[di] = reinterpret_cast<const uint4*>(P.src)[si];
        else
        {
            for (int b = 0; b < P.bpt; ++b) P.dst[di * P.bpt + b] = P.src[si * P.bpt + b];
        }
    }
}

/* ------------------------------------------------------------------------
 * Sort-first composite: every rank stores the rows it owns into the root's
 * copy of the attachment (a peer mapping over NVLink, or local memory on the
 * root itself).  Ownership is exclusive per row, so this is a pure gather.
 * ----------------------------------------------------------------------*/
struct SLCompositeParams
{
    const unsigned char* src;
    unsigned char* dst;
    unsigned long long rowBytes;
    unsigned int numRows, rank, world, stripRows;
};

__global__ void __launch_bounds__(256) k_composite(const __grid_constant__ SLCompositeParams P)
{
    const bool vec = (P.rowBytes % 16ull == 0ull) && ((reinterpret_cast<uintptr_t>(P.src) | reinterpret_cast<uintptr_t>(P.dst)) % 16u == 0u);
    for (unsigned y = blockIdx.y; y < P.numRows; y += gridDim.y)
    {
        if ((y / P.stripRows) % P.world != P.rank) continue;
        const unsigned char* s = P.src + (unsigned long long)y * P.rowBytes;
        unsigned char* d = P.dst + (unsigned long long)y * P.rowBytes;
        const unsigned long long t0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
        const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
        if (vec)
        {
            const uint4* s4 = reinterpret_cast<const uint4*>(s);
            uint4* d4 = reinterpret_cast<uint4*>(d);
            for (unsigned long long i = t0; i < P.rowBytes / 16ull; i += stride) d4[i] = s4[i];
        }
        else
        {
            for (unsigned long long i = t0; i < P.rowBytes; i += stride) d[i] = s[i];
        }
    }
}

} // namespace sld
