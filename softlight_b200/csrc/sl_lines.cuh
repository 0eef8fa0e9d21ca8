/* Line primitives: RENDER_MODE_LINES / RENDER_MODE_INDEXED_LINES.
 *
 *   k_line_vertex<VS>  SL_LineProcessor::process_verts + clip_and_process_lines (src/SL_LineProcessor.cpp:68-279): one
 *                      thread per segment (two elements): scissorMat * VS twice, the w < 0 rejection, the two-vertex clip
 *                      status (SL_PostVertexTransform.hpp:560-583), perspective divide, viewport transform, Liang-Barsky
 *                      against the viewport rectangle for partially visible segments.  A surviving segment becomes one
 *                      line record and the number of pixels sl_draw_line_bresenham will visit.
 *   k_line_offsets     exclusive prefix of the per-block pixel counts: candidate slots in submission order.
 *   k_line_emit        SL_LineRasterizer::render_line up to the depth test (src/SL_LineRasterizer.cpp:28-87): a warp per
 *                      segment, a lane per pixel -- step k of the Bresenham walk in closed form (the walk's error term after
 *                      k steps has crossed its threshold floor((k * dErr + dx - 1) / (2 dx)) times) -- interpolation
 *                      parameter (sqrtss, rsqrtss + Newton) and depth; one candidate and one sort key
 *                      (linear texel << 32 | slot) per pixel.
 *   (stable radix sort of the keys on the texel bits: per texel, candidates stay in submission order)
 *   k_line_shade<FS>   the depth test and flush_line_fragments (src/SL_FragmentProcessor.cpp:346-403): the thread that holds
 *                      a texel's first key walks that texel's candidates: depth test against the buffer, varyings =
 *                      fmadd(b - a, interp, a), shader, colour write / blend, depth write.
 *
 * A segment visits every pixel once, so per-texel submission order is the whole ordering constraint -- except for a
 * segment that touches column `width` (the clip rectangle's far edge is inclusive) and column 0 of the next row: the
 * reference addresses texels linearly, both are the same texel, and its second depth test runs while the first fragment
 * is still queued.  That case is detected and flagged (SL_ERR_LINE_ALIAS), not reproduced. */
#pragma once

#include "sl_pipeline.cuh"
#include "sl_points.cuh"

namespace sld
{

constexpr unsigned SL_ERR_LINE_ALIAS = 0x400u;
constexpr int SL_LINE_BLOCK = 256; /* segments per block of k_line_vertex / k_line_emit */

struct SLLineParams
{
    const SLDrawDesc* draw;               /* the one draw of this pass (HBM)                       */
    unsigned long long numLines;          /* segments per instance * instances                     */
    unsigned long long linesPerInstance;
    int numVaryings;
    unsigned int width, height;           /* depth attachment                                      */
    unsigned int rank, world, stripRows;  /* sort-first partition: rows this rank owns             */
    float4* lines;                        /* [numLines][2 + 2 * NV]: (x0, y0, z0, dist), (x1, y1, z1, -), varyings of a, of b */
    uint32_t* lineCount;                  /* [numLines] pixels of the walk (0: nothing to draw)    */
    unsigned long long* blockSum;         /* [blocks] pixels per block, then (k_line_offsets) their exclusive prefix */
    unsigned long long* keys;             /* [candidates]                                          */
    float4* cands;                        /* [candidates] (interp, depth, x | y << 16, segment)    */
    SLCounters* counters;
};

/* the walk's set-up (SL_Geometry.hpp:314-334) from the float end points: (uint16_t) casts, then int16_t arithmetic */
struct SLLineWalk
{
    int x1, y1, dx, dErr, yErr;
    bool steep;
    SLD_FN void init(float ax, float ay, float bx, float by)
    {
        int xa = (int16_t)(uint16_t)sl_cvt_tr(ax), ya = (int16_t)(uint16_t)sl_cvt_tr(ay);
        int xb = (int16_t)(uint16_t)sl_cvt_tr(bx), yb = (int16_t)(uint16_t)sl_cvt_tr(by);
        steep = ::abs(xa - xb) < ::abs(ya - yb);
        if (steep) { int t = xa; xa = ya; ya = t; t = xb; xb = yb; yb = t; }
        if (xa > xb) { int t = xa; xa = xb; xb = t; t = ya; ya = yb; yb = t; }
        x1 = xa;
        y1 = ya;
        dx = (int16_t)(xb - xa);
        const int dy = (int16_t)(yb - ya);
        dErr = (int16_t)(::abs(dy) * 2);
        yErr = (yb > ya) ? 1 : -1;
    }
    SLD_FN unsigned count() const { return dx >= 0 ? (unsigned)dx + 1u : 0u; }
    SLD_FN void pixel(unsigned k, uint32_t& px, uint32_t& py) const
    {
        const int m = dx > 0 ? (int)(((long long)k * dErr + dx - 1) / (2 * dx)) : 0;
        const int x = x1 + (int)k, y = y1 + yErr * m;
        px = (uint32_t)(steep ? y : x) & 0xffffu;
        py = (uint32_t)(steep ? x : y) & 0xffffu;
    }
};

template <VsFn VS>
__global__ void __launch_bounds__(SL_LINE_BLOCK) k_line_vertex(const __grid_constant__ SLLineParams P)
{
    __shared__ unsigned long long sSum[SL_LINE_BLOCK / 32];
    const unsigned long long g = (unsigned long long)blockIdx.x * SL_LINE_BLOCK + threadIdx.x;
    const int NV = P.numVaryings;
    unsigned count = 0u;
    bool binned = false;
    if (g < P.numLines)
    {
        const SLDrawDesc& D = *P.draw;
        const unsigned long long instanceId = g / P.linesPerInstance;
        const unsigned long long i = D.elementBegin + 2ull * (g - instanceId * P.linesPerInstance);
        SLVert a, b;
        SL_VertexParam params;
        params.pUniforms = &D.uniforms;
        params.pVao = &D.vao;
        params.pVbo = &D.vbo;
        params.instanceId = (size_t)instanceId;
        const math::mat4 scissorMat = *reinterpret_cast<const math::mat4*>(D.scissor);
#pragma unroll
        for (int k = 0; k < 2; ++k)
        {
            unsigned long long vertId = i + (unsigned long long)k;
            if (D.usingIndices) /* SL_IndexBuffer::index */
            {
                if (D.iboType == 0) vertId = D.ibo[vertId];
                else if (D.iboType == 1) vertId = reinterpret_cast<const uint16_t*>(D.ibo)[vertId];
                else vertId = reinterpret_cast<const uint32_t*>(D.ibo)[vertId];
            }
            SLVert& v = k ? b : a;
#pragma unroll
            for (int s = 0; s < 4; ++s) v.varyings[s] = math::vec4{0.f};
            params.vertId = (size_t)vertId;
            params.pVaryings = v.varyings;
            v.vert = scissorMat * VS(params); /* :246-252 */
        }
        bool draw = false;
        if (!(a.vert.v[3] < 0.f || b.vert.v[3] < 0.f)) /* :255-258 */
        {
            const float w0p = a.vert.v[3], w1p = b.vert.v[3];
            bool vis = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) vis = vis && (a.vert.v[k] <= w0p) && (b.vert.v[k] <= w1p) && (a.vert.v[k] >= -w0p) && (b.vert.v[k] >= -w1p);
            const bool part = (w0p > 0.f) || (w1p > 0.f);
            if (vis || part)
            {
                const float viewport[4] = {D.viewport[0], D.viewport[1], D.viewport[2], D.viewport[3]};
                sl_to_screen(a.vert, viewport);
                sl_to_screen(b.vert, viewport);
                if (vis) draw = true;
                else
                {
                    /* clip_and_process_lines (:68-177) */
                    const float dx = b.vert.v[0] - a.vert.v[0];
                    const float dy = b.vert.v[1] - a.vert.v[1];
                    const float xMin = viewport[0], yMin = viewport[1], xMax = viewport[2], yMax = viewport[3];
                    const double eps = 1.0e-9; /* LS_EPSILON, a double literal */
                    bool done = false;
                    if ((double)math::abs(dx) < eps && (double)math::abs(dy) < eps)
                    {
                        if (a.vert.v[0] >= xMin && a.vert.v[0] <= xMax && a.vert.v[1] >= yMin && a.vert.v[1] <= yMax) { draw = true; done = true; }
                    }
                    if (!done)
                    {
                        float tE = 0.f, tL = 1.f;
                        auto clip = [&](float num, float denom) -> bool
                        {
                            if ((double)math::abs(denom) < eps) return num < 0.f;
                            const float t = num / denom;
                            if (denom > 0.f)
                            {
                                if (t > tL) return false;
                                if (t > tE) tE = t;
                            }
                            else
                            {
                                if (t < tE) return false;
                                if (t < tL) tL = t;
                            }
                            return true;
                        };
                        SLVert p0 = a, p1 = b;
                        const float ex = p0.vert.v[0] - p1.vert.v[0], ey = p0.vert.v[1] - p1.vert.v[1];
                        const float dist = math::inversesqrt(ex * ex + ey * ey);
                        const float x1 = p0.vert.v[0], y1 = p0.vert.v[1];
                        if (clip(xMin - x1, dx) && clip(x1 - xMax, -dx) && clip(yMin - y1, dy) && clip(y1 - yMax, -dy))
                        {
                            if (tL < 1.f)
                            {
                                p1.vert.v[0] = x1 + tL * dx;
                                p1.vert.v[1] = y1 + tL * dy;
                            }
                            if (tE > 0.f)
                            {
                                p0.vert.v[0] = x1 + tE * dx;
                                p0.vert.v[1] = y1 + tE * dy;
                            }
                            const math::vec4 d0{p0.vert.v[0] - a.vert.v[0], p0.vert.v[1] - a.vert.v[1], p0.vert.v[2] - a.vert.v[2], p0.vert.v[3] - a.vert.v[3]};
                            const math::vec4 d1{p1.vert.v[0] - b.vert.v[0], p1.vert.v[1] - b.vert.v[1], p1.vert.v[2] - b.vert.v[2], p1.vert.v[3] - b.vert.v[3]};
                            const float interp0 = math::length(d0) * dist;
                            const float interp1 = math::length(d1) * dist;
                            p0.vert.v[2] = a.vert.v[2] + ((b.vert.v[2] - a.vert.v[2]) * interp0);
                            p0.vert.v[3] = a.vert.v[3] + ((b.vert.v[3] - a.vert.v[3]) * interp0);
                            p1.vert.v[2] = b.vert.v[2] + ((a.vert.v[2] - b.vert.v[2]) * interp1);
                            p1.vert.v[3] = b.vert.v[3] + ((a.vert.v[3] - b.vert.v[3]) * interp1);
#pragma unroll
                            for (int s = 0; s < 4; ++s)
                            {
                                if (s < NV)
                                {
                                    p0.varyings[s] = math::mix(a.varyings[s], b.varyings[s], interp0);
                                    p1.varyings[s] = math::mix(b.varyings[s], a.varyings[s], interp1);
                                }
                            }
                            a = p0;
                            b = p1;
                            draw = true;
                        }
                    }
                }
            }
        }
        if (draw)
        {
            binned = true;
            SLLineWalk walk;
            walk.init(a.vert.v[0], a.vert.v[1], b.vert.v[0], b.vert.v[1]);
            count = walk.count();
            const float ddx = b.vert.v[0] - a.vert.v[0], ddy = b.vert.v[1] - a.vert.v[1];
            const float dist = math::inversesqrt(ddx * ddx + ddy * ddy); /* render_line :38 */
            float4* rec = P.lines + g * (unsigned long long)(2 + 2 * NV);
            rec[0] = make_float4(a.vert.v[0], a.vert.v[1], a.vert.v[2], dist);
            rec[1] = make_float4(b.vert.v[0], b.vert.v[1], b.vert.v[2], 0.f);
#pragma unroll
            for (int s = 0; s < 4; ++s)
            {
                if (s < NV)
                {
                    rec[2 + s] = *reinterpret_cast<const float4*>(&a.varyings[s]);
                    rec[2 + NV + s] = *reinterpret_cast<const float4*>(&b.varyings[s]);
                }
            }
        }
        P.lineCount[g] = count;
    }
    /* pixels of this block */
    unsigned long long sum = count;
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31) == 0) sSum[threadIdx.x >> 5] = sum;
    const unsigned bm = __ballot_sync(0xffffffffu, binned);
    if ((threadIdx.x & 31) == 0 && bm) atomicAdd(&P.counters->trisBinned, (unsigned long long)__popc(bm));
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned long long t = 0;
        for (int w = 0; w < SL_LINE_BLOCK / 32; ++w) t += sSum[w];
        P.blockSum[blockIdx.x] = t;
        if (t) atomicAdd(&P.counters->pixelsCovered, t);
    }
}

/* exclusive prefix of blockSum in place (one CTA; a few thousand blocks) */
__global__ void __launch_bounds__(1024) k_line_offsets(unsigned long long* blockSum, unsigned numBlocks)
{
    __shared__ unsigned long long sWarp[32];
    __shared__ unsigned long long sCarry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sCarry = 0ull;
    __syncthreads();
    for (unsigned base = 0; base < numBlocks; base += 1024u)
    {
        const unsigned i = base + (unsigned)tid;
        const unsigned long long v = i < numBlocks ? blockSum[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) sWarp[warp] = incl;
        __syncthreads();
        unsigned long long before = sCarry;
        for (int w = 0; w < warp; ++w) before += sWarp[w];
        if (i < numBlocks) blockSum[i] = before + incl - v;
        __syncthreads();
        if (tid == 1023) sCarry = before + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SL_LINE_BLOCK) k_line_emit(const __grid_constant__ SLLineParams P)
{
    __shared__ unsigned long long sBase[SL_LINE_BLOCK];
    __shared__ unsigned long long sWarp[SL_LINE_BLOCK / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long g0 = (unsigned long long)blockIdx.x * SL_LINE_BLOCK;
    {
        const unsigned long long g = g0 + tid;
        const unsigned long long v = g < P.numLines ? P.lineCount[g] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) sWarp[warp] = incl;
        __syncthreads();
        unsigned long long before = P.blockSum[blockIdx.x];
        for (int w = 0; w < warp; ++w) before += sWarp[w];
        sBase[tid] = before + incl - v;
        __syncthreads();
    }
    const int NV = P.numVaryings;
    const unsigned long long texels = (unsigned long long)P.width * P.height;
    for (int l = warp; l < SL_LINE_BLOCK; l += SL_LINE_BLOCK / 32)
    {
        const unsigned long long g = g0 + (unsigned long long)l;
        if (g >= P.numLines) break;
        const unsigned count = P.lineCount[g];
        if (!count) continue;
        const float4* rec = P.lines + g * (unsigned long long)(2 + 2 * NV);
        const float4 r0 = rec[0], r1 = rec[1];
        SLLineWalk walk;
        walk.init(r0.x, r0.y, r1.x, r1.y);
        const unsigned long long base = sBase[l];
        int wMin = 0x7fffffff, wMax = -1, zMin = 0x7fffffff, zMax = -1; /* rows of the pixels in column `width` / column 0 */
        for (unsigned k = lane; k < count; k += 32u)
        {
            uint32_t px, py;
            walk.pixel(k, px, py);
            /* render_line's callback (:53-66) */
            const math::vec4 d{(float)px - r0.x, (float)py - r0.y, 0.f, 0.f};
            const float interp = math::length(d) * r0.w;
            float z = r0.z + ((r1.z - r0.z) * interp);
            /* a segment of zero length: rsqrtss(0) = inf, 0 * inf = the x86 default NaN (sign bit set), which every later SSE
             * operation passes on unchanged; the GPU's arithmetic yields its own canonical NaN */
            if (z != z) z = __uint_as_float(0xffc00000u);
            const unsigned long long pix = (unsigned long long)px + (unsigned long long)P.width * py;
            unsigned long long key = SL_POINT_INVISIBLE;
            if (pix < texels)
            {
                const unsigned row = (unsigned)(pix / P.width);
                if (P.world <= 1u || (row / P.stripRows) % P.world == P.rank) key = (pix << 32) | (base + k);
            }
            P.keys[base + k] = key;
            P.cands[base + k] = make_float4(interp, z, __uint_as_float(px | (py << 16)), __uint_as_float((uint32_t)g));
            if (px >= P.width) { wMin = min(wMin, (int)py); wMax = max(wMax, (int)py); }
            if (px == 0u) { zMin = min(zMin, (int)py); zMax = max(zMax, (int)py); }
        }
        wMin = __reduce_min_sync(0xffffffffu, wMin);
        wMax = __reduce_max_sync(0xffffffffu, wMax);
        zMin = __reduce_min_sync(0xffffffffu, zMin);
        zMax = __reduce_max_sync(0xffffffffu, zMax);
        /* (width, y) and (0, y + 1) are one texel: the segment visits it twice */
        if (lane == 0 && wMax >= 0 && zMax >= 0 && wMin + 1 <= zMax && zMin <= wMax + 1) atomicOr(&P.counters->errorFlags, SL_ERR_LINE_ALIAS);
    }
}

struct SLLineShadeParams
{
    const SLDrawDesc* draw;
    const unsigned long long* keys;       /* sorted by texel, submission order inside a texel */
    const float4* cands;
    const float4* lines;
    unsigned long long numCands;
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

template <FsFn FS>
__global__ void __launch_bounds__(128) k_line_shade(const __grid_constant__ SLLineShadeParams P)
{
    const unsigned long long t = (unsigned long long)blockIdx.x * 128ull + threadIdx.x;
    unsigned shaded = 0u;
    if (t < P.numCands)
    {
        unsigned long long key = P.keys[t];
        const unsigned long long pix = key >> 32;
        if (key != SL_POINT_INVISIBLE && (t == 0ull || (P.keys[t - 1ull] >> 32) != pix))
        {
            SL_FragmentParam fp;
            fp.pUniforms = &P.draw->uniforms;
            const int NV = P.numVaryings;
            for (unsigned long long j = t;;)
            {
                const float4 c = P.cands[key & 0xffffffffull];
                const float d = sl_load_depth(P.depth, P.depthBpt, (size_t)pix);
                if (sl_depth_cmp_scalar(P.depthTest, c.y, d)) /* :62-66 */
                {
                    const uint32_t xy = __float_as_uint(c.z);
                    fp.coord.x = (uint16_t)(xy & 0xffffu);
                    fp.coord.y = (uint16_t)(xy >> 16);
                    fp.coord.depth = c.y;
                    const float4* rec = P.lines + (unsigned long long)__float_as_uint(c.w) * (unsigned long long)(2 + 2 * NV);
                    for (int s = 0; s < NV; ++s) /* interpolate_line_varyings: fmadd(v1 - v0, p, v0) */
                    {
                        const float4 a = rec[2 + s], b = rec[2 + NV + s];
                        fp.pVaryings[s] = math::vec4{sl_fma(b.x - a.x, c.x, a.x), sl_fma(b.y - a.y, c.x, a.y), sl_fma(b.z - a.z, c.x, a.z), sl_fma(b.w - a.w, c.x, a.w)};
                    }
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
                    if (P.depthMask) sl_store_depth(P.depth, P.depthBpt, (size_t)pix, c.y);
                }
                if (++j >= P.numCands) break;
                key = P.keys[j];
                if ((key >> 32) != pix) break;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) shaded += __shfl_xor_sync(0xffffffffu, shaded, o);
    if ((threadIdx.x & 31) == 0 && shaded) atomicAdd(&P.counters->fragsShaded, (unsigned long long)shaded);
}

} // namespace sld
