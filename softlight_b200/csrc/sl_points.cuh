/* Point primitives: RENDER_MODE_POINTS / RENDER_MODE_INDEXED_POINTS.
 *
 *   k_point_vertex<VS>  SL_PointProcessor::process_verts (src/SL_PointProcessor.cpp:70-113): one thread per element;
 *                       scissorMat * VS, the one-vertex clip test (SL_PostVertexTransform.hpp:536-553), perspective
 *                       divide, viewport transform.  A visible vertex becomes one point record and one sort key
 *                       (linear pixel index << 32 | submission index).
 *   (stable radix sort of the keys on the pixel bits: per pixel, points stay in submission order)
 *   k_point_shade<FS>   SL_PointRasterizer::render_point (src/SL_PointRasterizer.cpp:27-101): the thread that holds the
 *                       first key of a pixel walks that pixel's points in submission order: depth test against the
 *                       buffer, shader, colour write / blend, depth write.  (Compacting the pixel starts first, as the
 *                       wireframe path does with k_pixel_leaders below, was measured: at 1-2 points per pixel the extra
 *                       pass costs what it saves, 0.57 -> 0.62 ms for 4 M points.)
 *
 * The reference shades points strictly in submission order; a point touches exactly one texel (x + width*y, linear like
 * every other framebuffer access of the reference) and the shader cannot read the framebuffer, so ordering per pixel is
 * the whole ordering constraint.  Points whose linear index falls at or past width*height land in the reference's
 * allocation padding (src/SL_Texture.cpp:28-54) where nothing can read them back; they are dropped here and not counted. */
#pragma once

#include "sl_pipeline.cuh"

namespace sld
{

constexpr unsigned long long SL_POINT_INVISIBLE = ~0ull;

struct SLPointParams
{
    const SLDrawDesc* draw;               /* the one draw of this pass (HBM)                      */
    unsigned long long numPoints;         /* elements * instances                                 */
    unsigned long long elementsPerInstance;
    int numVaryings;
    unsigned int width, height;           /* depth attachment                                     */
    unsigned int rank, world, stripRows;  /* sort-first partition: rows this rank owns            */
    unsigned long long* keys;             /* [numPoints]                                          */
    float4* points;                       /* [numPoints][1 + numVaryings]: (x, y, z, -) + varyings */
    SLCounters* counters;
};

template <VsFn VS>
__global__ void __launch_bounds__(256) k_point_vertex(const __grid_constant__ SLPointParams P)
{
    const unsigned long long g = (unsigned long long)blockIdx.x * 256ull + threadIdx.x;
    bool binned = false;
    if (g < P.numPoints)
    {
        const SLDrawDesc& D = *P.draw;
        const unsigned long long instanceId = g / P.elementsPerInstance;
        const unsigned long long i = D.elementBegin + (g - instanceId * P.elementsPerInstance);
        unsigned long long vertId = i;
        if (D.usingIndices) /* SL_IndexBuffer::index (SL_IndexBuffer.hpp) */
        {
            if (D.iboType == 0) vertId = D.ibo[i];
            else if (D.iboType == 1) vertId = reinterpret_cast<const uint16_t*>(D.ibo)[i];
            else vertId = reinterpret_cast<const uint32_t*>(D.ibo)[i];
        }
        SLVert v;
        SL_VertexParam params;
        params.pUniforms = &D.uniforms;
        params.pVao = &D.vao;
        params.pVbo = &D.vbo;
        params.instanceId = (size_t)instanceId;
        params.vertId = (size_t)vertId;
        params.pVaryings = v.varyings;
        const math::mat4 scissorMat = *reinterpret_cast<const math::mat4*>(D.scissor);
        v.vert = scissorMat * VS(params); /* :94-98 */

        /* sl_ndc_clip_status, one vertex: w > 0 && clip <= w && clip >= -w on all four components */
        const float w0p = v.vert.v[3], w0n = -w0p;
        bool visible = w0p > 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) visible = visible && (v.vert.v[k] <= w0p) && (v.vert.v[k] >= w0n);

        unsigned long long key = SL_POINT_INVISIBLE;
        if (visible)
        {
            const float viewport[4] = {D.viewport[0], D.viewport[1], D.viewport[2], D.viewport[3]};
            sl_to_screen(v.vert, viewport);
            binned = true;
            /* (uint16_t)screenCoord (src/SL_PointRasterizer.cpp:51-52): cvttss2si, low 16 bits */
            const uint32_t x = (uint32_t)(int)v.vert.v[0] & 0xffffu;
            const uint32_t y = (uint32_t)(int)v.vert.v[1] & 0xffffu;
            const unsigned long long pix = (unsigned long long)x + (unsigned long long)P.width * y;
            if (pix < (unsigned long long)P.width * P.height)
            {
                const unsigned row = (unsigned)(pix / P.width);
                if (P.world <= 1u || (row / P.stripRows) % P.world == P.rank)
                {
                    key = (pix << 32) | g;
                    float4* rec = P.points + g * (unsigned long long)(1 + P.numVaryings);
                    rec[0] = make_float4(__uint_as_float(x), __uint_as_float(y), v.vert.v[2], 0.f);
#pragma unroll
                    for (int s = 0; s < 4; ++s)
                    {
                        if (s < P.numVaryings) rec[1 + s] = *reinterpret_cast<const float4*>(&v.varyings[s]);
                    }
                }
            }
        }
        P.keys[g] = key;
    }
    const unsigned b = __ballot_sync(0xffffffffu, binned);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&P.counters->trisBinned, (unsigned long long)__popc(b));
}

/* The pixels of a sorted key array: positions whose predecessor belongs to another pixel, compacted so that the walking
 * kernels run one pixel per lane (walking from every key and letting the non-first ones exit left 1.5 of 32 lanes busy
 * on the wireframe path: profiles/r01h_k_wire_depth_before_compaction.txt).  One atomic per block keeps the list close to sorted. */
struct SLLeaderParams
{
    const unsigned long long* keys;
    unsigned long long n;
    int shift;                 /* pixel = key >> shift */
    uint32_t* leaders;
    unsigned int* numLeaders;
};

__global__ void __launch_bounds__(256) k_pixel_leaders(const __grid_constant__ SLLeaderParams P)
{
    __shared__ unsigned sWarp[8];
    __shared__ unsigned sBase;
    const unsigned long long t = (unsigned long long)blockIdx.x * 256ull + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    bool leader = false;
    if (t < P.n) leader = t == 0ull || (P.keys[t - 1ull] >> P.shift) != (P.keys[t] >> P.shift);
    const unsigned b = __ballot_sync(0xffffffffu, leader);
    if (lane == 0u) sWarp[warp] = (unsigned)__popc(b);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        /* one atomic per block (one per warp was 0.7 ms of serialised atomics for 8.6 M keys) */
        unsigned tot = 0u;
        for (int w = 0; w < 8; ++w) { const unsigned c = sWarp[w]; sWarp[w] = tot; tot += c; }
        sBase = tot ? atomicAdd(P.numLeaders, tot) : 0u;
    }
    __syncthreads();
    if (leader) P.leaders[sBase + sWarp[warp] + __popc(b & ((1u << lane) - 1u))] = (uint32_t)t;
}

struct SLPointShadeParams
{
    const SLDrawDesc* draw;
    const unsigned long long* keys;       /* sorted by pixel, submission order inside a pixel */
    const float4* points;
    unsigned long long numPoints;
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
__global__ void __launch_bounds__(128) k_point_shade(const __grid_constant__ SLPointShadeParams P)
{
    const unsigned long long t = (unsigned long long)blockIdx.x * 128ull + threadIdx.x;
    unsigned shaded = 0u;
    if (t < P.numPoints)
    {
        unsigned long long key = P.keys[t];
        const unsigned long long pix = key >> 32;
        if (key != SL_POINT_INVISIBLE && (t == 0ull || (P.keys[t - 1ull] >> 32) != pix))
        {
            SL_FragmentParam fp;
            fp.pUniforms = &P.draw->uniforms;
            for (unsigned long long j = t;;)
            {
                const float4* rec = P.points + (key & 0xffffffffull) * (unsigned long long)(1 + P.numVaryings);
                const float4 p = rec[0];
                const float d = sl_load_depth(P.depth, P.depthBpt, (size_t)pix);
                if (sl_depth_cmp_scalar(P.depthTest, p.z, d)) /* :58-62 */
                {
                    fp.coord.x = (uint16_t)__float_as_uint(p.x);
                    fp.coord.y = (uint16_t)__float_as_uint(p.y);
                    fp.coord.depth = p.z;
                    for (int s = 0; s < P.numVaryings; ++s)
                    {
                        const float4 a = rec[1 + s];
                        fp.pVaryings[s] = math::vec4{a.x, a.y, a.z, a.w};
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) fp.pOutputs[i] = math::vec4{0.f};
                    ++shaded;
                    if (FS(fp)) /* :69-93 */
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
                    if (P.depthMask) sl_store_depth(P.depth, P.depthBpt, (size_t)pix, p.z); /* :95-98 */
                }
                if (++j >= P.numPoints) break;
                key = P.keys[j];
                if ((key >> 32) != pix) break;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) shaded += __shfl_xor_sync(0xffffffffu, shaded, o);
    if ((threadIdx.x & 31) == 0 && shaded) atomicAdd(&P.counters->fragsShaded, (unsigned long long)shaded);
}

} // namespace sld
