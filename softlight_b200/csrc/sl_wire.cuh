/* Wireframe triangles: RENDER_MODE_TRI_WIRE / RENDER_MODE_INDEXED_TRI_WIRE.
 *
 * The vertex stage is the triangle path's (k_vertex: same SL_TriProcessor); only the rasteriser differs:
 * SL_TriRasterizer::render_wireframe (src/SL_TriRasterizer.cpp:116-202) visits, per scan line, d0 pixels from xMin and
 * the last d1 pixels before xMax (d = how far that bound moved since the previous line, at least 1), with scalar
 * barycentrics.  Edge fragments are sparse, so instead of the span / resolve machinery of filled triangles they are
 * handled like points (sl_points.cuh):
 *
 *   k_wire<EMIT=0>     counts the fragment candidates of every triangle record (sizes the buffers)
 *   k_wire<EMIT=1>     writes one (key, payload) per candidate: key = linear pixel << lowBits | submission ordinal of
 *                      the record << 3 | visit order of that record on that pixel; sort value = depth | payload slot
 *   (radix sort of the keys)
 *   k_pixel_leaders    (sl_points.cuh) compacts the positions where a pixel's candidates start
 *   k_wire_shade<FS>   one thread per pixel walks that pixel's candidates in submission order: depth test,
 *                      SL_FragmentProcessor::flush_tri_fragments shading, colour / depth writes.
 *
 * The reference tests a candidate when it is queued and writes depth when the queue is flushed (600 fragments, or at
 * the end of the triangle), and render_wireframe does revisit pixels (the jump to the right edge can land left of where
 * the left edge stopped; a run that crosses the row end wraps onto pixels of the next line).  All candidates of ONE
 * record on a pixel are therefore tested against the depth the pixel had before that record -- unless a queue flush fell
 * between two visits: a visit sees the depth written by an earlier visit of the same record iff that one was flushed
 * first, i.e. iff its flush epoch (passed fragments of the record before it / 600) is smaller.  Epochs depend on the depth
 * outcomes and the outcomes on the epochs; both are independent of the fragment shader (depth is written whatever the
 * shader returns), so for batches with records of 600+ candidates the host iterates k_wire_depth (per-pixel depth
 * simulation -> pass flags) and k_wire_epoch (per-record prefix of the pass flags -> epochs) to the fixed point before
 * k_wire_shade runs with the final epochs.  Candidates that are not stored (allocation padding beyond width*height, rows
 * of another rank under the sort-first partition) are not counted in the epochs. */
#pragma once

#include "sl_pipeline.cuh"
#include "sl_points.cuh" /* k_pixel_leaders */

namespace sld
{

constexpr int SL_WIRE_SUB_BITS = 3;
/* key = pixel << lowBits | ordinal << 3 | visit order, ordinal = (vertex block * 32 + segment) << 8 | index in the segment;
 * lowBits = 3 + the bits the batch's ordinals need (13 + bits(numBlocks - 1)): fewer radix passes for small batches */

struct SLWireParams
{
    const float4* records;
    const SLBlockInfo* blockInfo;
    unsigned int numBlocks;
    int recStride;
    unsigned int width, height;
    unsigned int rank, world, stripRows;  /* sort-first partition */
    unsigned long long capacity;          /* EMIT: fragment slots */
    int lowBits;
    unsigned long long* keys;
    unsigned long long* vals;             /* sort values: depth bits << 32 | payload slot of the key */
    float4* fragBc;
    uint32_t* fragXY;
    uint32_t* fragRec;
    uint32_t* recBase;                    /* [record] first slot of the record's candidates (k_wire_offsets: submission order) */
    uint32_t* recCount;                   /* [record] candidates (COUNT out; EMIT zeroes what could not be stored)         */
    SLCounters* counters;                 /* pixelsCovered: candidates (COUNT) / slot allocator (EMIT); rowEndCuts: records with 600+ candidates (COUNT) */
};

constexpr unsigned SL_WIRE_QUEUE = 600u; /* SL_SHADER_MAX_QUEUED_FRAGS */

/* max(abs(a - b), 1) in wrapping 32-bit arithmetic (src/SL_TriRasterizer.cpp:157-158): on the line above the top vertex
 * the bounds are 0x80000000 (cvtss2si of NaN / inf); a - b and abs() wrap back to INT_MIN and the max picks 1 */
SLD_FN int32_t sl_wire_edge_width(int32_t a, int32_t b)
{
    const int32_t d = (int32_t)((uint32_t)a - (uint32_t)b);
    const int32_t ad = d < 0 ? (int32_t)(0u - (uint32_t)d) : d;
    return ad > 1 ? ad : 1;
}

/* the visit loop of one scan line (:162-169); f(x, run) is called for every visited x, run = 0 before the jump */
template <typename F>
SLD_FN void sl_wire_row(int32_t x0, int32_t x1, int32_t d0, int32_t d1, int32_t xLimit, F f)
{
    int run = 0;
    for (int32_t ix = 0, x = x0; (uint32_t)x < (uint32_t)x1 && x < xLimit; ++ix, ++x)
    {
        if (ix == d0)
        {
            const int32_t j = (int32_t)((uint32_t)x1 - (uint32_t)d1 - 1u);
            x = x0 > j ? x0 : j;
            run = 1;
            continue;
        }
        f(x, run);
    }
}

/* One WARP per (vertex block, segment, index): segments are the (round, warp) record regions of k_vertex, taken in
 * submission order exactly like k_span does; the lanes take the record's scan lines 32 at a time, and a warp scan of the
 * per-line counts keeps the record's candidates contiguous and in visit order (line by line). */
constexpr unsigned SL_WIRE_BLOCKS_PER_VB = SL_VB_SEGS * 256u / 4u; /* 128 threads = 4 warps = 4 record slots */

template <int EMIT>
__global__ void __launch_bounds__(128) k_wire(const __grid_constant__ SLWireParams P)
{
    const unsigned long long w = ((unsigned long long)blockIdx.x * 128ull + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned vb = (unsigned)(w >> 13), seg = (unsigned)(w >> 8) & 31u, idx = (unsigned)w & 255u;
    if (vb >= P.numBlocks) return; /* warp-uniform exits */
    const SLSegment sg = P.blockInfo[vb].seg[seg];
    if (sg.count > 256u && idx == 0u && lane == 0) atomicOr(&P.counters->errorFlags, SL_ERR_WIRE_SEGMENT);
    if (idx >= sg.count) return;
    const uint32_t recId = sg.base + idx;

    const float4* r = P.records + (size_t)recId * P.recStride;
    const float4 p0 = r[0], p1 = r[1], p2 = r[2];
    const int32_t W = (int32_t)P.width, H = (int32_t)P.height;
    int32_t yBegin = sl_cvt_tr(sl_min(p2.y, sl_min(p0.y, p1.y))); /* :136-137 */
    int32_t yEnd = sl_cvt_tr(sl_max(p2.y, sl_max(p0.y, p1.y)));
    if (yBegin < 0) yBegin = 0;
    if (yEnd > H) yEnd = H; /* rows >= H only reach the allocation padding */
    SLScanline sl;
    sl.init(p0, p1, p2);
    const int32_t xLimit = 4 * W; /* further right than three wraps: flagged, not drawn */
    const unsigned long long numPix = (unsigned long long)P.width * P.height;
    const uint32_t ordinal = ((vb * 32u + seg) << 8) | idx;

    auto owned = [&](unsigned long long pix) -> bool {
        if (pix >= numPix) return false;
        if (P.world <= 1u) return true;
        const unsigned row = (unsigned)(pix / P.width);
        return (row / P.stripRows) % P.world == P.rank;
    };
    /* candidates of line y that this context stores; hit = the line has a visit at all */
    auto count_row = [&](int32_t y, int32_t& x0, int32_t& x1, int32_t& d0, int32_t& d1, bool& hit) -> unsigned {
        const float yf = (float)y;
        int32_t q0, q1;
        sl.step(yf, x0, x1);
        sl.step(yf - 1.f, q0, q1);
        d0 = sl_wire_edge_width(x0, q0);
        d1 = sl_wire_edge_width(x1, q1);
        unsigned c = 0;
        bool h = false;
        sl_wire_row(x0, x1, d0, d1, xLimit, [&](int32_t x, int) {
            h = true;
            if (owned((unsigned long long)x + (unsigned long long)P.width * (unsigned)y)) ++c;
        });
        if ((uint32_t)x0 < (uint32_t)x1 && x1 > xLimit) atomicOr(&P.counters->errorFlags, SL_ERR_SPAN_TWO_ROWS);
        hit = h;
        return c;
    };

    unsigned n = 0, rowsHit = 0;
    for (int32_t y = yBegin + lane; y < yEnd; y += 32)
    {
        int32_t x0, x1, d0, d1;
        bool hit;
        n += count_row(y, x0, x1, d0, d1, hit);
        rowsHit += hit ? 1u : 0u;
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if (!EMIT)
    {
        if (lane == 0)
        {
            P.recCount[recId] = n; /* k_wire_offsets turns the counts into slots in submission order */
            if (n) atomicAdd(&P.counters->pixelsCovered, (unsigned long long)n);
            if (n >= SL_WIRE_QUEUE) atomicAdd(&P.counters->rowEndCuts, 1ull);
        }
        return;
    }
    rowsHit = __reduce_add_sync(0xffffffffu, rowsHit);
    /* the record's slots: handed out in submission order (k_wire_offsets), so that a stable sort on the pixel bits alone
     * leaves every pixel's candidates in (record, line, run) order */
    const unsigned long long base = P.recBase[recId];
    if (lane == 0 && rowsHit) atomicAdd(&P.counters->spans, (unsigned long long)rowsHit); /* the statistic: (record, line) pairs with a visit */
    const bool fits = base + n <= P.capacity && n == P.recCount[recId];
    if (lane == 0 && !fits)
    {
        P.recCount[recId] = 0u;
        atomicOr(&P.counters->errorFlags, SL_ERR_WIRE_SLOTS);
    }
    if (!n || !fits) return;

    SLTriSetup ts;
    sl_tri_setup(ts, p0, p1, p2);
    const math::vec4 depth{p0.z, p1.z, p2.z, 0.f};
    unsigned long long running = base;
    for (int32_t yBase = yBegin; yBase < yEnd; yBase += 32)
    {
        const int32_t y = yBase + lane;
        int32_t x0 = 0, x1 = 0, d0 = 1, d1 = 1;
        bool hit = false;
        unsigned c = 0;
        if (y < yEnd) c = count_row(y, x0, x1, d0, d1, hit);
        unsigned inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        unsigned long long slot = running + (inc - c);
        running += __shfl_sync(0xffffffffu, inc, 31);
        if (!c) continue;
        const float yf = (float)y;
        const math::vec4 yv{yf, yf, yf, 0.f};
        math::vec4 bcY;
#pragma unroll
        for (int k = 0; k < 4; ++k) bcY.v[k] = sl_fma(ts.bc1.v[k], yv.v[k], ts.bc2.v[k]); /* :149 */
        sl_wire_row(x0, x1, d0, d1, xLimit, [&](int32_t x, int run) {
            const unsigned long long pix = (unsigned long long)x + (unsigned long long)P.width * (unsigned)y;
            if (!owned(pix)) return;
            const float xf = (float)x;
            const math::vec4 xv{xf, xf, xf, 0.f};
            math::vec4 bc;
#pragma unroll
            for (int k = 0; k < 4; ++k) bc.v[k] = sl_fma(ts.bc0.v[k], xv.v[k], bcY.v[k]); /* :173 */
            const float z = math::dot(depth, bc);
            /* visit order of this record on this pixel: the line the visit came from (a wrapped run belongs to an earlier
             * line than the pixel's own), then the run */
            const unsigned wraps = (unsigned)(x / W); /* 0..3 */
            const unsigned sub = ((3u - wraps) << 1) | (unsigned)run;
            P.keys[slot] = (pix << P.lowBits) | ((unsigned long long)ordinal << SL_WIRE_SUB_BITS) | sub;
            P.vals[slot] = ((unsigned long long)__float_as_uint(z) << 32) | (unsigned long long)(uint32_t)slot;
            P.fragBc[slot] = make_float4(bc.v[0], bc.v[1], bc.v[2], bc.v[3]);
            P.fragXY[slot] = ((uint32_t)x & 0xffffu) | (((uint32_t)y & 0xffffu) << 16);
            P.fragRec[slot] = recId;
            ++slot;
        });
    }
}

/* Slots of every record's candidates in submission order: an exclusive prefix of recCount over the records taken
 * segment by segment ((round, warp) regions of the vertex blocks, the order k_span uses too).  One CTA: every thread sums a
 * contiguous run of segments, the run totals are scanned in shared memory, the run is walked again to hand out the bases. */
__global__ void __launch_bounds__(1024) k_wire_offsets(const SLBlockInfo* blockInfo, unsigned numBlocks, const uint32_t* recCount, uint32_t* recBase)
{
    __shared__ unsigned long long sWarp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned numSegs = numBlocks * (unsigned)SL_VB_SEGS;
    const unsigned per = (numSegs + 1023u) / 1024u;
    const unsigned a = min((unsigned)tid * per, numSegs), b = min(a + per, numSegs);
    unsigned long long sum = 0ull;
    for (unsigned sIdx = a; sIdx < b; ++sIdx)
    {
        const SLSegment sg = blockInfo[sIdx / SL_VB_SEGS].seg[sIdx % SL_VB_SEGS];
        const uint32_t cnt = min(sg.count, 256u);
        for (uint32_t i = 0; i < cnt; ++i) sum += recCount[sg.base + i];
    }
    unsigned long long incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    unsigned long long run = incl - sum;
    for (int w = 0; w < warp; ++w) run += sWarp[w];
    for (unsigned sIdx = a; sIdx < b; ++sIdx)
    {
        const SLSegment sg = blockInfo[sIdx / SL_VB_SEGS].seg[sIdx % SL_VB_SEGS];
        const uint32_t cnt = min(sg.count, 256u);
        for (uint32_t i = 0; i < cnt; ++i)
        {
            recBase[sg.base + i] = (uint32_t)run;
            run += recCount[sg.base + i];
        }
    }
}

struct SLWireShadeParams
{
    const SLDrawDesc* draws;
    const uint16_t* recDraw;
    const float4* records;
    int recStride;
    int lowBits;
    const unsigned long long* keys;   /* sorted */
    const unsigned long long* vals;   /* sorted along with the keys: depth bits << 32 | slot */
    const float4* fragBc;
    const uint32_t* fragXY;
    const uint32_t* fragRec;
    uint32_t* leaders;                /* k_pixel_leaders out: positions in the sorted order where a pixel's candidates start */
    unsigned int* numLeaders;
    const uint16_t* fragEpoch;        /* flush epoch of every candidate (null: all 0) */
    uint8_t* fragPass;                /* k_wire_depth out: depth outcome of every candidate */
    unsigned long long numFrags;
    int numVaryings;
    int numOutputs;
    int blend;
    int depthTest;
    int depthMask;
    unsigned char* depth;
    int depthBpt;
    unsigned char* color[4];
    int colorType[4];
    SLCounters* counters;
};

/* The walk both kernels share: the thread that holds the first key of a pixel takes that pixel's candidates in order,
 * record by record; f(slot, z) is called for every candidate that passes its depth test and must perform the depth write
 * itself if it wants one (k_wire_depth only tracks the value). */
template <bool SHADE, typename F>
SLD_FN void sl_wire_walk_pixel(const SLWireShadeParams& P, unsigned long long t, F f)
{
    const unsigned long long pix = P.keys[t] >> P.lowBits;
    float dCur = sl_load_depth(P.depth, P.depthBpt, (size_t)pix);
    unsigned long long j = t;
    bool more = true;
    while (more)
    {
        const unsigned long long ordMask = (1ull << (P.lowBits - SL_WIRE_SUB_BITS)) - 1ull;
        const unsigned long long ord = (P.keys[j] >> SL_WIRE_SUB_BITS) & ordMask;
        const float dPre = dCur;
        float zs[8];
        uint32_t eps[8];
        int gs = 0;
        for (;;)
        {
            const unsigned long long v = P.vals[j];
            const uint32_t s = (uint32_t)v;
            const float z = __uint_as_float((uint32_t)(v >> 32));
            const uint32_t ep = P.fragEpoch ? (uint32_t)P.fragEpoch[s] : 0u;
            /* what this visit sees: the depth before the record, or the last earlier visit that passed and was flushed */
            float dSeen = dPre;
            if (P.depthMask)
            {
                for (int i = 0; i < gs; ++i) { if (eps[i] < ep) dSeen = zs[i]; }
            }
            const bool pass = sl_depth_cmp_scalar(P.depthTest, z, dSeen);
            if (!SHADE) P.fragPass[s] = pass ? 1 : 0;
            if (pass)
            {
                f(s, z, pix);
                const float stored = P.depthBpt == 2 ? sl_h2f(sl_f2h_rn(z)) : z; /* what a later read of the texel returns */
                if (P.depthMask) dCur = stored;
                if (gs < 8) { zs[gs] = stored; eps[gs] = ep; ++gs; }
            }
            if (++j >= P.numFrags) { more = false; break; }
            const unsigned long long k = P.keys[j];
            if ((k >> P.lowBits) != pix) { more = false; break; }
            if (((k >> SL_WIRE_SUB_BITS) & ordMask) != ord) break;
        }
    }
}

/* depth-only simulation of the batch with the current epochs: pass flag per candidate; the attachment is not written */
__global__ void __launch_bounds__(128) k_wire_depth(const __grid_constant__ SLWireShadeParams P)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * 128ull + threadIdx.x;
    if (i < (unsigned long long)*P.numLeaders) sl_wire_walk_pixel<false>(P, P.leaders[i], [](uint32_t, float, unsigned long long) {});
}

struct SLWireEpochParams
{
    const SLBlockInfo* blockInfo;
    unsigned int numBlocks;
    const uint32_t* recBase;
    const uint32_t* recCount;
    const uint8_t* fragPass;
    uint16_t* fragEpoch;
    unsigned int* changed;
};

/* flush epoch of every candidate of the records with 600+ candidates: passed candidates before it / 600.  One warp per
 * record slot (the indexing of k_wire), 32 candidates per step. */
__global__ void __launch_bounds__(128) k_wire_epoch(const __grid_constant__ SLWireEpochParams P)
{
    const unsigned long long w = ((unsigned long long)blockIdx.x * 128ull + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned vb = (unsigned)(w >> 13), seg = (unsigned)(w >> 8) & 31u, idx = (unsigned)w & 255u;
    if (vb >= P.numBlocks) return;
    const SLSegment sg = P.blockInfo[vb].seg[seg];
    if (idx >= sg.count) return;
    const uint32_t recId = sg.base + idx;
    const uint32_t n = P.recCount[recId];
    if (n < SL_WIRE_QUEUE) return;
    const uint32_t base = P.recBase[recId];
    uint32_t passedBefore = 0u;
    bool ch = false;
    for (uint32_t i0 = 0; i0 < n; i0 += 32u)
    {
        const uint32_t i = i0 + lane;
        const bool p = i < n && P.fragPass[base + i] != 0;
        const uint32_t b = __ballot_sync(0xffffffffu, p);
        const uint32_t e = (passedBefore + __popc(b & ((1u << lane) - 1u))) / SL_WIRE_QUEUE;
        const uint16_t ep = (uint16_t)(e > 0xffffu ? 0xffffu : e);
        if (i < n && P.fragEpoch[base + i] != ep) { P.fragEpoch[base + i] = ep; ch = true; }
        passedBefore += __popc(b);
    }
    if (__any_sync(0xffffffffu, ch) && lane == 0u) atomicAdd(P.changed, 1u);
}

template <FsFn FS>
__global__ void __launch_bounds__(128) k_wire_shade(const __grid_constant__ SLWireShadeParams P)
{
    const unsigned long long li = (unsigned long long)blockIdx.x * 128ull + threadIdx.x;
    unsigned shaded = 0u;
    if (li < (unsigned long long)*P.numLeaders)
    {
        sl_wire_walk_pixel<true>(P, P.leaders[li], [&](uint32_t s, float z, unsigned long long pix) {
            const float4 b4 = P.fragBc[s];
            const uint32_t xy = P.fragXY[s];
            const uint32_t recId = P.fragRec[s];
            const float4* rec = P.records + (size_t)recId * P.recStride;
            SL_FragmentParam fp;
            /* flush_tri_fragments (src/SL_FragmentProcessor.cpp:438-529), as in k_shade */
            math::vec4 bc = math::vec4{b4.x, b4.y, b4.z, b4.w} * math::vec4{rec[0].w, rec[1].w, rec[2].w, 0.f};
            const float persp = sl_rcp((bc.v[0] + bc.v[1]) + (bc.v[2] + bc.v[3]));
            bc = bc * persp;
            fp.pUniforms = &P.draws[P.recDraw[recId]].uniforms;
            for (int v = 0; v < P.numVaryings; ++v)
            {
                const float4 a = rec[3 + 3 * v + 0], b = rec[3 + 3 * v + 1], cc = rec[3 + 3 * v + 2];
                fp.pVaryings[v] = math::vec4{
                    sl_fma(bc.v[2], cc.x, sl_fma(bc.v[1], b.x, bc.v[0] * a.x)), sl_fma(bc.v[2], cc.y, sl_fma(bc.v[1], b.y, bc.v[0] * a.y)),
                    sl_fma(bc.v[2], cc.z, sl_fma(bc.v[1], b.z, bc.v[0] * a.z)), sl_fma(bc.v[2], cc.w, sl_fma(bc.v[1], b.w, bc.v[0] * a.w))};
            }
            fp.coord.x = (uint16_t)(xy & 0xffffu);
            fp.coord.y = (uint16_t)(xy >> 16);
            fp.coord.depth = z;
#pragma unroll
            for (int i = 0; i < 4; ++i) fp.pOutputs[i] = math::vec4{0.f};
            ++shaded;
            if (FS(fp))
            {
                if (P.blend == 0)
                {
                    for (int i = P.numOutputs - 1; i >= 0; --i) sl_store_color(P.color[i], P.colorType[i], (size_t)pix, fp.pOutputs[i]);
                }
                else
                {
                    for (int i = P.numOutputs - 1; i >= 0; --i) sl_blend_color(P.color[i], P.colorType[i], (size_t)pix, fp.pOutputs[i], P.blend);
                }
            }
            if (P.depthMask) sl_store_depth(P.depth, P.depthBpt, (size_t)pix, z);
        });
    }
    for (int o = 16; o > 0; o >>= 1) shaded += __shfl_xor_sync(0xffffffffu, shaded, o);
    if ((threadIdx.x & 31) == 0 && shaded) atomicAdd(&P.counters->fragsShaded, (unsigned long long)shaded);
}

} // namespace sld
