/*
 * sl_resolve2.cuh -- k_resolve2: the depth / coverage / order stage with ONE LANE PER SPAN PIECE.
 *
 * Same contract as k_resolve (sl_pipeline.cuh): SL_TriRasterizer::render_triangle_simd (src/SL_TriRasterizer.cpp:413-559)
 * + the depth write of flush_tri_fragments (src/SL_FragmentProcessor.cpp:526-529); input = the ordered per-row span
 * lists, output = ordered fragment lists in HBM, hand-overs between x-segments and rows through the same table.
 *
 * k_resolve spreads the 4-pixel groups of consecutive pieces over the lanes of a window and settles the window in waves;
 * every lane pays the whole book-keeping for one group.  For scenes made of short spans (a few groups per span: every
 * triangle-dense config) that overhead is ~13x the arithmetic.  Here a lane owns a piece (a span clipped to the
 * warp's x-segment) and walks its groups left to right exactly like the reference's loop does: the barycentric
 * registers are advanced by real additions (:539-542) and compacted in place (:496-498), so there is no lane map, no
 * replay and no per-group set-up.  What remains to be coordinated between lanes is the submission order per pixel:
 *
 *   - pieces are handed to free lanes in list order; every lane knows which lanes hold EARLIER pieces (a bit mask kept
 *     up to date at hand-out time);
 *   - a piece marks all its pixels in a per-pixel lane mask in shared memory when it starts and clears a pixel's mark
 *     when it has resolved it; a lane may work on its next SL_R2_GROUPS groups only when no earlier lane still has a mark
 *     on one of those pixels.  The earliest unfinished piece never waits, lanes that run in the same round touch disjoint
 *     pixels, and a later piece follows an earlier one as closely as one round behind.
 *
 * The depth texels of the segment live in shared memory for the whole ticket: staged in with one bulk copy
 * (cp.async.bulk + mbarrier, the TMA engine's 1-D path) and written back with one bulk store.
 * Fragments go straight into the HBM chunks of their (row, 128-pixel column) list: positions come from shared-memory
 * atomics, chunks are reserved once per round for the pixels about to be resolved.
 */
#pragma once

#include "sl_pipeline.cuh"

namespace sld
{

#ifndef SL_R2_GROUPS
#define SL_R2_GROUPS 2 /* groups a lane resolves per round */
#endif
#ifndef SL_R2_KEEP_LANES
#define SL_R2_KEEP_LANES 20 /* with at least this many lanes holding a piece, a round skips the hand-out */
#endif
#ifndef SL_R2_WARPS_N
#define SL_R2_WARPS_N 4
#endif
#ifndef SL_R2_MIN_BLOCKS
#define SL_R2_MIN_BLOCKS 4
#endif
constexpr int SL_R2_WARPS = SL_R2_WARPS_N;
constexpr int SL_R2_RING = 256;      /* pieces waiting for a lane (per warp) */
constexpr int SL_R2_FIFO = 64;       /* free chunk ids held by the warp */
constexpr unsigned SL_R2_FIFO_BLOCK = 16u; /* chunks taken from the global pool at a time */
constexpr int SL_R2_CHUNK_RING = 32; /* newest chunk ids remembered per sub-list */
constexpr unsigned SL_R2_RCP_BYTES = 2048u * 4u; /* per block: the rcp table */
constexpr unsigned SL_R2_RING_BYTES = SL_R2_RING * (unsigned)sizeof(SLSpanEntry);
constexpr unsigned SL_R2_SUB_BYTES = SL_MAX_SUBLISTS * 4u * 4u + SL_MAX_SUBLISTS * SL_R2_CHUNK_RING * 4u; /* count, chunks, want, last chunk + id ring */
constexpr unsigned SL_R2_FIFO_BYTES = SL_R2_FIFO * 4u;
constexpr unsigned SL_R2_FIXED_BYTES = SL_R2_RING_BYTES + SL_R2_SUB_BYTES + SL_R2_FIFO_BYTES + 16u; /* + the tile's mbarrier */

__host__ __device__ inline unsigned sl_r2_warp_bytes(unsigned segWidth, unsigned tileBytes) { return SL_R2_FIXED_BYTES + ((4u * segWidth + 15u) & ~15u) + tileBytes; }

/* ---- 1-D bulk copies through the TMA engine (cp.async.bulk) ---- */
SLD_FN uint32_t sl_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
SLD_FN void sl_mbar_init(void* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sl_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
SLD_FN void sl_bulk_load(void* dstSmem, const void* srcGlobal, unsigned bytes, void* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sl_smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sl_smem_addr(dstSmem)), "l"(srcGlobal), "r"(bytes),
                 "r"(sl_smem_addr(bar))
                 : "memory");
}
SLD_FN void sl_mbar_wait(void* bar, unsigned phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(sl_smem_addr(bar)),
        "r"(phase)
        : "memory");
}
SLD_FN void sl_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
SLD_FN void sl_bulk_store(void* dstGlobal, const void* srcSmem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstGlobal), "r"(sl_smem_addr(srcSmem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
SLD_FN void sl_bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

/* depth of one barycentric register (_sl_mul_vec4_mat4_ps, src/SL_TriRasterizer.cpp:376-393) */
SLD_FN float sl_zdot(const math::vec4& a, float z0, float z1, float z2) { return ((a.v[3] * 0.f) + (a.v[2] * z2)) + ((a.v[1] * z1) + (a.v[0] * z0)); }

/* depth texel of the staged segment, typed at compile time (src/SL_TriRasterizer.cpp:40-102, src/SL_FragmentProcessor.cpp:526-529) */
template <int BPT> SLD_FN float sl_tile_load(const unsigned char* tile, int idx)
{
    if (BPT == 2) return sl_h2f(reinterpret_cast<const uint16_t*>(tile)[idx]);
    if (BPT == 4) return reinterpret_cast<const float*>(tile)[idx];
    return (float)reinterpret_cast<const double*>(tile)[idx];
}
template <int BPT> SLD_FN void sl_tile_store(unsigned char* tile, int idx, float z)
{
    if (BPT == 2) reinterpret_cast<uint16_t*>(tile)[idx] = sl_f2h_rn(z);
    else if (BPT == 4) reinterpret_cast<float*>(tile)[idx] = z;
    else reinterpret_cast<double*>(tile)[idx] = (double)z;
}
/* SL_DepthFunc* (SL_ShaderUtil.hpp:240-507) as a truth table over (z < d, z > d, z == d): ordered, quiet -- all three false
 * on NaN, which only OFF lets pass.  Bit index = lt | gt << 1 | eq << 2. */
SLD_FN uint32_t sl_depth_table(int func)
{
    return func == 0 ? 0xffu : func == 1 ? 0x02u : func == 2 ? 0x12u : func == 3 ? 0x04u : func == 4 ? 0x14u : func == 5 ? 0x10u : 0x06u;
}
SLD_FN bool sl_depth_pass(uint32_t table, float z, float d)
{
    const uint32_t idx = (z < d ? 1u : 0u) | (z > d ? 2u : 0u) | (z == d ? 4u : 0u);
    return ((table >> idx) & 1u) != 0u;
}

/* first probe of a hand-over slot, issued early in the round (acquire: the registers are read after the flag) */
SLD_FN unsigned long long sl_ld_acquire_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

/* SL_DepthFunc* (SL_ShaderUtil.hpp:240-507) chosen at compile time: ordered, quiet -- false on NaN, which only OFF lets pass */
template <int FUNC> SLD_FN bool sl_depth_func(float z, float d)
{
    if (FUNC == 0) return true;
    if (FUNC == 1) return z < d;
    if (FUNC == 2) return z <= d;
    if (FUNC == 3) return z > d;
    if (FUNC == 4) return z >= d;
    if (FUNC == 5) return z == d;
    return (z < d) || (z > d);
}

template <int BPT, int FUNC>
__global__ void __launch_bounds__(SL_R2_WARPS * 32, SL_R2_MIN_BLOCKS) k_resolve2(const __grid_constant__ SLRasterParams P)
{
    extern __shared__ __align__(16) unsigned char sDyn[];
    uint32_t* sRcp = reinterpret_cast<uint32_t*>(sDyn);
    for (unsigned i = threadIdx.x; i < 2048u; i += blockDim.x) sRcp[i] = g_rcp_table[i];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ltMask = (1u << lane) - 1u, myBit = 1u << lane;
    const int W = (int)P.width, NS = (int)P.numSegs;
    constexpr bool chained = FUNC != 0;
    const bool useDepth = FUNC != 0 || P.depthMask != 0;
    const int nsub = (int)P.numSubLists;

    unsigned char* warpMem = sDyn + SL_R2_RCP_BYTES + (size_t)warp * sl_r2_warp_bytes(P.segWidth, P.tileBytes);
    SLSpanEntry* ring = reinterpret_cast<SLSpanEntry*>(warpMem);
    uint32_t* subCount = reinterpret_cast<uint32_t*>(warpMem + SL_R2_RING_BYTES); /* fragments in the sub-list so far  */
    uint32_t* subChunks = subCount + SL_MAX_SUBLISTS;                              /* chunks it owns                    */
    uint32_t* subWant = subChunks + SL_MAX_SUBLISTS;                               /* pixels about to be resolved       */
    uint32_t* subLast = subWant + SL_MAX_SUBLISTS;                                 /* its newest chunk                  */
    uint32_t* chunkRing = subLast + SL_MAX_SUBLISTS;                               /* [sub][SL_R2_CHUNK_RING] chunk ids */
    uint32_t* chunkFifo = chunkRing + SL_MAX_SUBLISTS * SL_R2_CHUNK_RING;          /* free chunk ids of this warp       */
    unsigned long long* tileBar = reinterpret_cast<unsigned long long*>(warpMem + SL_R2_RING_BYTES + SL_R2_SUB_BYTES + SL_R2_FIFO_BYTES);
    uint32_t* pixMask = reinterpret_cast<uint32_t*>(warpMem + SL_R2_FIXED_BYTES); /* [segWidth] lanes whose piece still has to resolve the pixel */
    unsigned char* tile = warpMem + SL_R2_FIXED_BYTES + ((4u * P.segWidth + 15u) & ~15u);
    if (lane == 0) sl_mbar_init(tileBar, 1u);
    unsigned tilePhase = 0u;
    __syncthreads();
    if (P.counters->abortBatch) return; /* the batch did not fit: it is run again, nothing visible may be written */
    const bool skewed = P.ticketOrder && P.counters->rowEndCuts == 0ull; /* the skew assumes that hand-overs only go to the right neighbour */

    unsigned long long dbgRounds = 0, dbgIdle = 0, dbgGroups = 0;
    const long long dbgT0 = P.diagnostics ? clock64() : 0ll;
    unsigned shaded = 0;
    /* Fine segments.  The kernel lasts as long as its longest (row, segment) ticket, and in a deep scene that is a row near
     * the horizon: thousands of overlapping pieces in one segment, settled layer after layer by ONE warp.  The rows
     * k_row_order found heavy (more than P.fineRows times the average number of entries; they lead rowPerm) are cut along
     * the boundaries of their fragment lists instead: every 2^SL_SHADE_SHIFT pixels get a warp of their own, with
     * hand-overs at the finer boundaries.  Their tickets come first, row by row; the tickets of the other rows follow in
     * the skewed order.  Only when no span runs past a row end: all hand-overs stay inside a row, one row = one numbering. */
    const unsigned fineF = (unsigned)nsub;
    const unsigned numHeavy = (skewed && P.rowPerm && nsub > 1 && NS * nsub <= 64 && P.fineRows != 0u) ? P.rowPerm[P.numRows] : 0u;
    const unsigned fineTickets = numHeavy * (unsigned)NS * fineF;
    const unsigned numTickets = fineTickets + P.numRows * P.numSegs;
    /* free chunks of this warp: a FIFO of ids in shared memory, topped up SL_R2_FIFO_BLOCK at a time; the top-up's atomic
     * is issued a round before its result is needed (lane 0 keeps it in flight) and what is left over goes to the next ticket */
    unsigned fifoHead = 0u, fifoCount = 0u;
    uint32_t pendingBase = 0u; /* lane 0 */
    bool pendingValid = false; /* warp-uniform */

    /* persistent warps: (row, segment) tickets in launch order; a consumer of a hand-over only ever waits for a smaller
     * ticket, which a running warp holds */
    for (;;)
    {
        unsigned ticket = 0;
        if (lane == 0) ticket = atomicAdd(&P.counters->rasterTicket, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket >= numTickets) break;
        unsigned fineK = 0u;
        bool fine = false;
        int y, seg;
        if (ticket < fineTickets)
        {
            const unsigned perRow = (unsigned)NS * fineF;
            const unsigned fs = ticket % perRow;
            y = (int)P.rowPerm[ticket / perRow];
            seg = (int)(fs / fineF);
            fineK = fs % fineF;
            fine = true;
        }
        else
        {
            ticket -= fineTickets;
            y = (int)(ticket / (unsigned)NS);
            seg = (int)(ticket % (unsigned)NS);
            if (skewed)
            {
                const uint32_t o = P.ticketOrder[ticket];
                y = (int)(o >> 6);
                seg = (int)(o & 63u);
                if (P.rowPerm)
                {
                    if ((unsigned)y < numHeavy) continue; /* a heavy row: done by the fine tickets */
                    y = (int)P.rowPerm[y];                /* rows are independent here: the heaviest go first */
                }
            }
        }
        if (P.rowNeeded && !P.rowNeeded[y]) continue;
        const uint32_t begin = P.rowStart[y], end = P.rowStart[y + 1];
        if (begin == end) continue;
        const long long dbgTicket0 = P.diagnostics ? clock64() : 0ll;
        unsigned dbgPieces = 0u;
        const unsigned long long dbgRounds0 = dbgRounds;
        int sx0 = seg * (int)P.segWidth;
        int sx1 = (seg == NS - 1) ? W : sx0 + (int)P.segWidth;
        uint32_t firstList = (uint32_t)(y * NS + seg) * P.numSubLists;
        int segIdx = seg, segCount = NS; /* this warp's place among the row's segments (hand-over boundaries are numbered by it) */
        if (fine)
        {
            sx0 += (int)(fineK << SL_SHADE_SHIFT);
            if (sx0 >= sx1) continue;
            sx1 = min(sx0 + (1 << SL_SHADE_SHIFT), sx1);
            firstList += fineK; /* exactly the pixels of one fragment list */
            segIdx = seg * (int)fineF + (int)fineK;
            segCount = NS * (int)fineF;
        }
        const unsigned segPixels = (unsigned)(sx1 - sx0);
        const long long segOrigin = (long long)W * y + sx0;
        const unsigned tileLoadBytes = segPixels * (unsigned)BPT;
        unsigned char* tileSrc = P.depth + (size_t)segOrigin * BPT;
        const bool bulkOk = (tileLoadBytes & 15u) == 0u && (reinterpret_cast<uintptr_t>(tileSrc) & 15u) == 0u;

        if (lane < SL_MAX_SUBLISTS)
        {
            subCount[lane] = 0u;
            subChunks[lane] = 0u;
            subWant[lane] = 0u;
            subLast[lane] = SL_NO_CHUNK;
        }
        bool staged = false, tileReady = false, dirty = false;

        /* lane state: the piece this lane walks */
        bool busy = false, ready = false, publish = false;
        uint32_t earlier = 0u, recId = 0u, fM = 0u;
        int xMin = 0, p0 = 0, p1 = 0, g = 0, gEnd = 0, xBase = 0, ySrc = 0, bOut = 0;
        math::vec4 a0{0.f}, a1{0.f}, a2{0.f}, a3{0.f}, bcX{0.f};
        float z0 = 0.f, z1 = 0.f, z2 = 0.f;

        /* the row's entry list is read 128 entries ahead: four loads per lane in flight while the warp works */
        uint32_t base = begin;
        SLSpanEntry pre[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const uint32_t idx = base + 32u * k + (uint32_t)lane;
            pre[k] = SLSpanEntry{0u, 0, 0};
            if (idx < end) pre[k] = P.entries[idx];
        }
        unsigned ringHead = 0u, ringCount = 0u, idle = 0u;
        /* A round is: top (probe hand-overs, fill the ring, hand pieces to free lanes) + settle (check, reserve chunks,
         * resolve).  While most lanes still hold a piece the top is skipped: the pieces on board are settled layer after
         * layer, and what the longest tickets spend per layer is what the kernel takes. */
        bool skipTop = false;
        for (;;)
        {
            const uint32_t busyMask = __ballot_sync(0xffffffffu, busy);
            /* hand-over slots are probed early: the load travels while the warp fills its ring and sets pieces up */
            unsigned long long pollW = 0ull;
            uint32_t pollI = 0u;
            if (!skipTop && busy && !ready)
            {
                if (!P.handoff) { atomicOr(&P.counters->errorFlags, SL_ERR_HANDOFF_POLL); pollW = ~0ull; }
                else
                {
                    pollI = sl_handoff_hash(sl_handoff_key(recId, ySrc, bOut - 1)) & P.handoffMask;
                    pollW = sl_ld_acquire_u64(&P.handoff[pollI]);
                }
            }
            /* ---- keep the ring of waiting pieces filled: the row's entries that reach into this segment, in list order ---- */
            while (!skipTop && ringCount <= (unsigned)(SL_R2_RING - 128) && base < end)
            {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const SLSpanEntry c = pre[k];
                    const int cb = (c.rec & SL_ENTRY_OVERFLOW) ? W : 0;
                    const bool cHas = max((int)c.xMin, cb + sx0) < min((int)c.xMax, cb + sx1); /* empty entries: 0 < 0 */
                    const uint32_t hm = __ballot_sync(0xffffffffu, cHas);
                    if (cHas)
                    {
                        ring[(ringHead + ringCount + (unsigned)__popc(hm & ltMask)) & (SL_R2_RING - 1)] = c;
                        const float4* r = P.records + (size_t)(c.rec & ~SL_ENTRY_OVERFLOW) * P.recStride;
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(r));
                    }
                    ringCount += (unsigned)__popc(hm);
                }
                base += 128u;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const uint32_t idx = base + 32u * k + (uint32_t)lane;
                    pre[k] = SLSpanEntry{0u, 0, 0};
                    if (idx < end) pre[k] = P.entries[idx];
                }
            }
            if (!busyMask && !ringCount) break;
            __syncwarp();

            /* ---- hand the waiting pieces to the free lanes, in order ---- */
            const unsigned myk = (unsigned)__popc(~busyMask & ltMask);
            const bool take = !skipTop && !busy && myk < ringCount;
            const uint32_t newMask = __ballot_sync(0xffffffffu, take);
            if (newMask)
            {
                if (useDepth && !staged)
                {
                    /* the first piece of this segment: bring its depth texels in (one bulk copy, overlapped with the set-up) */
                    staged = true;
                    if (bulkOk)
                    {
                        if (lane == 0)
                        {
                            sl_fence_async_smem();
                            sl_bulk_load(tile, tileSrc, tileLoadBytes, tileBar);
                        }
                    }
                    else
                    {
                        sl_copy_warp(tile, tileSrc, tileLoadBytes, lane);
                        tileReady = true;
                    }
                    for (unsigned i = lane; i < segPixels; i += 32) pixMask[i] = 0u;
                    __syncwarp();
                }
                if (take)
                {
                    const SLSpanEntry e = ring[(ringHead + myk) & (SL_R2_RING - 1)];
                    const bool isOver = (e.rec & SL_ENTRY_OVERFLOW) != 0u;
                    recId = e.rec & ~SL_ENTRY_OVERFLOW;
                    xBase = isOver ? W : 0; /* the entry's x coordinates are those of its source row */
                    ySrc = isOver ? y - 1 : y;
                    xMin = e.xMin;
                    const int xMax = e.xMax;
                    p0 = max(xMin, xBase + sx0);
                    p1 = min(xMax, xBase + sx1);
                    const bool needsIn = xMin < p0 && chained;
                    publish = chained && xMax > p1 && (segIdx < segCount - 1 || (!isOver && y + 1 < (int)P.numRows));
                    bOut = (isOver ? segCount : 0) + segIdx + 1;
                    g = (p0 - xMin) >> 2;
                    gEnd = ((p1 - 1 - xMin) >> 2) + 1;
                    busy = true;
                    earlier = busyMask | (newMask & ltMask);

                    /* triangle set-up (push_bin, src/SL_TriProcessor.cpp:271-314), once per piece */
                    const float4* rec = P.records + (size_t)recId * P.recStride;
                    const float4 q0 = rec[0], q1 = rec[1], q2 = rec[2];
                    const float xy0 = q0.x - q1.x, xy1 = q0.y - q1.y;
                    const float zy0 = q2.x - q1.x, zy1 = q2.y - q1.y;
                    const float denom = sl_rcp_tab(sRcp, (zy0 * xy1) - (zy1 * xy0));
                    const math::vec4 ptX{q0.x, q1.x, q2.x, 0.f};
                    const math::vec4 ptY{q0.y, q1.y, q2.y, 0.f};
                    const math::vec4 one{1.f};
                    const math::vec4 bc0 = math::cross(ptY, one) * denom;
                    z0 = q0.z; z1 = q1.z; z2 = q2.z;
                    bcX = bc0 * 4.f;
                    fM = 0u;
                    if (!needsIn)
                    {
                        const math::vec4 bc1 = math::cross(one, ptX) * denom;
                        const math::vec4 bc2 = math::cross(ptX, ptY) * denom;
                        const math::vec4 bcY = math::fmadd(bc1, math::vec4{(float)ySrc}, bc2); /* src/SL_TriRasterizer.cpp:474 */
                        a0 = (bc0 * (float)(xMin + 0)) + bcY;                                  /* :476-481 */
                        a1 = (bc0 * (float)(xMin + 1)) + bcY;
                        a2 = (bc0 * (float)(xMin + 2)) + bcY;
                        a3 = (bc0 * (float)(xMin + 3)) + bcY;
                        /* depth test off and the piece starts inside its span: the lanes left of the cut passed and the
                         * registers were never compacted -- replay the additions from the span start (:539-542) */
                        for (int i = 0; i < g; ++i) { a0 += bcX; a1 += bcX; a2 += bcX; a3 += bcX; }
                        if (xMin < p0) fM = ((1u << (p0 - (xMin + 4 * g))) - 1u) & 7u;
                        ready = true;
                    }
                    else ready = false; /* continues a span of the segment on the left (or of the row below): needs its hand-over */
                    if (useDepth)
                    {
                        const int m0 = p0 - xBase - sx0, m1 = p1 - xBase - sx0;
                        for (int px = m0; px < m1; ++px) atomicOr(&pixMask[px], myBit);
                    }
                }
                else earlier &= ~newMask; /* those lanes now hold later pieces */
                const unsigned nTaken = (unsigned)__popc(newMask);
                dbgPieces += nTaken;
                ringHead += nTaken;
                ringCount -= nTaken;
                __syncwarp();
            }

            if (staged && !tileReady)
            {
                sl_mbar_wait(tileBar, tilePhase);
                tilePhase ^= 1u;
                tileReady = true;
            }
            /* ---- which lanes may resolve their next groups ---- */
            bool go = false;
            bool finishedEarly = false;
            int cx = 0, ce = 0;
            if (busy)
            {
                if (!ready)
                {
                    /* the slot probed at the top of the round; on a key mismatch keep probing (open addressing) */
                    const unsigned long long key = sl_handoff_key(recId, ySrc, bOut - 1);
                    bool found = pollW == ~0ull;
                    while (!found && pollW != 0ull)
                    {
                        if ((pollW >> 12) == key) { found = (pollW & 0x8ull) != 0ull; break; }
                        pollI = (pollI + 1u) & P.handoffMask;
                        pollW = sl_ld_acquire_u64(&P.handoff[pollI]);
                    }
                    if (found)
                    {
                        if (pollW != ~0ull)
                        {
                            const float4* src = P.handoffRegs + (size_t)pollI * 4;
                            const float4 f0 = __ldcg(src + 0), f1 = __ldcg(src + 1), f2 = __ldcg(src + 2), f3 = __ldcg(src + 3);
                            a0 = math::vec4{f0.x, f0.y, f0.z, f0.w};
                            a1 = math::vec4{f1.x, f1.y, f1.z, f1.w};
                            a2 = math::vec4{f2.x, f2.y, f2.z, f2.w};
                            a3 = math::vec4{f3.x, f3.y, f3.z, f3.w};
                            fM = (uint32_t)(pollW & 7ull);
                        }
                        ready = true;
                    }
                }
                if (ready)
                {
                    cx = max(p0, xMin + 4 * g) - xBase - sx0;
                    ce = min(p1, xMin + 4 * (g + SL_R2_GROUPS)) - xBase - sx0;
                    uint32_t ov = 0u;
                    if (useDepth)
                    {
#pragma unroll
                        for (int k = 0; k < 4 * SL_R2_GROUPS; ++k)
                        {
                            if (cx + k < ce) ov |= pixMask[cx + k];
                        }
                    }
                    go = (ov & earlier) == 0u;
                    /* Early rejection.  With a monotone depth function (LESS .. GREATER_EQUAL) the texel a pixel is tested against
                     * only ever moves in the direction that makes the test harder, so a pixel that fails against the depth of NOW
                     * fails against the depth the reference would see once the earlier pieces are settled.  A group whose
                     * pixels all fail changes nothing -- no depth write, no fragment, no compaction (:492-495) -- so a piece that
                     * still waits for earlier pieces may walk over such groups at once: the hidden layers of a deep scene drain
                     * without taking their turn in the per-pixel order.  (Not for a group that carries handed-over outcomes or
                     * has to publish its state.) */
                    if (!go && FUNC >= 1 && FUNC <= 4)
                    {
                        bool advanced = false;
                        while (g < gEnd && fM == 0u && !(publish && g == gEnd - 1))
                        {
                            const int x4 = xMin + 4 * g;
                            const int rel = x4 - xBase - sx0;
                            const int oLo = max(p0 - x4, 0), oHi = min(p1 - x4, 4);
                            const float zq[4] = {sl_zdot(a0, z0, z1, z2), sl_zdot(a1, z0, z1, z2), sl_zdot(a2, z0, z1, z2), sl_zdot(a3, z0, z1, z2)};
                            bool anyPass = false;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                            {
                                if (j >= oLo && j < oHi && sl_depth_func<FUNC>(zq[j], sl_tile_load<BPT>(tile, rel + j))) anyPass = true;
                            }
                            if (anyPass) break;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                            {
                                if (j >= oLo && j < oHi) atomicAnd(&pixMask[rel + j], ~myBit); /* other waiting lanes may clear theirs at the same time */
                            }
                            a0 += bcX; a1 += bcX; a2 += bcX; a3 += bcX; /* :539-542 */
                            ++g;
                            ++dbgGroups;
                            advanced = true;
                        }
                        if (g == gEnd) { busy = false; finishedEarly = true; }
                        else if (advanced)
                        {
                            cx = max(p0, xMin + 4 * g) - xBase - sx0;
                            ce = min(p1, xMin + 4 * (g + SL_R2_GROUPS)) - xBase - sx0;
                            ov = 0u;
#pragma unroll
                            for (int k = 0; k < 4 * SL_R2_GROUPS; ++k)
                            {
                                if (cx + k < ce) ov |= pixMask[cx + k];
                            }
                            go = (ov & earlier) == 0u;
                        }
                    }
                }
            }
            const uint32_t goMask = __ballot_sync(0xffffffffu, go);
            if (!goMask) skipTop = false; /* nothing on board can move: probe the hand-overs, take new pieces */
            if (!goMask && __any_sync(0xffffffffu, finishedEarly)) continue; /* lanes became free: hand out the next pieces */
            if (!goMask)
            {
                /* everything that could run waits for a hand-over: back off so that the polling does not take issue slots
                 * from the warps that produce it */
                idle = min(idle + 1u, 4u);
                __nanosleep(20u << idle);
                ++dbgIdle;
                continue;
            }
            idle = 0u;
            ++dbgRounds;
            if (go)
            {
                const int sA = cx >> SL_SHADE_SHIFT, sB = (ce - 1) >> SL_SHADE_SHIFT;
                if (sA == sB) atomicAdd(&subWant[sA], (uint32_t)(ce - cx));
                else
                {
                    const int split = sB << SL_SHADE_SHIFT;
                    atomicAdd(&subWant[sA], (uint32_t)(split - cx));
                    atomicAdd(&subWant[sB], (uint32_t)(ce - split));
                }
            }
            __syncwarp();
            /* ---- chunks for the pixels of this round: lane s looks after sub-list s, ids come from the warp's FIFO ---- */
            if (pendingValid)
            {
                /* the top-up asked for a round ago has landed */
                const uint32_t b = __shfl_sync(0xffffffffu, pendingBase, 0);
                if (lane < (int)SL_R2_FIFO_BLOCK) chunkFifo[(fifoHead + fifoCount + (unsigned)lane) & (SL_R2_FIFO - 1)] = b == SL_NO_CHUNK ? SL_NO_CHUNK : b + (uint32_t)lane;
                fifoCount += SL_R2_FIFO_BLOCK;
                pendingValid = false;
            }
            uint32_t need = 0u, have = 0u;
            if (lane < nsub)
            {
                const uint32_t want = subWant[lane];
                if (want)
                {
                    subWant[lane] = 0u;
                    have = subChunks[lane];
                    const uint32_t needTotal = (subCount[lane] + want + 31u) >> 5;
                    if (needTotal > have) need = needTotal - have;
                }
            }
            uint32_t incl = need;
#pragma unroll
            for (int d = 1; d < SL_MAX_SUBLISTS; d <<= 1)
            {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const uint32_t totalNeed = __shfl_sync(0xffffffffu, incl, SL_MAX_SUBLISTS - 1);
            if (totalNeed)
            {
                while (fifoCount < totalNeed)
                {
                    /* more than the FIFO holds (the first rounds of a warp, bursts): take a block now */
                    uint32_t b = 0u;
                    if (lane == 0)
                    {
                        b = atomicAdd(&P.counters->chunkAlloc, SL_R2_FIFO_BLOCK);
                        if (b + SL_R2_FIFO_BLOCK > P.lists.chunkCapacity) { atomicOr(&P.counters->errorFlags, SL_ERR_POOL); b = SL_NO_CHUNK; } /* pool exhausted: flagged, fragments dropped */
                    }
                    b = __shfl_sync(0xffffffffu, b, 0);
                    if (lane < (int)SL_R2_FIFO_BLOCK) chunkFifo[(fifoHead + fifoCount + (unsigned)lane) & (SL_R2_FIFO - 1)] = b == SL_NO_CHUNK ? SL_NO_CHUNK : b + (uint32_t)lane;
                    fifoCount += SL_R2_FIFO_BLOCK;
                }
                __syncwarp();
                if (need)
                {
                    uint32_t prev = subLast[lane];
                    const unsigned first = fifoHead + (incl - need);
                    for (uint32_t i = 0; i < need; ++i)
                    {
                        const uint32_t id = chunkFifo[(first + i) & (SL_R2_FIFO - 1)];
                        chunkRing[lane * SL_R2_CHUNK_RING + ((have + i) & (SL_R2_CHUNK_RING - 1))] = id;
                        if (id != SL_NO_CHUNK)
                        {
                            if (prev == SL_NO_CHUNK) P.lists.listHead[firstList + lane] = id;
                            else P.lists.chunkNext[prev] = id;
                            P.lists.chunkPrev[id] = prev;
                            prev = id;
                        }
                    }
                    subLast[lane] = prev;
                    subChunks[lane] = have + need;
                }
                fifoHead += totalNeed;
                fifoCount -= totalNeed;
            }
            if (fifoCount < SL_R2_FIFO_BLOCK * 2u && !pendingValid)
            {
                if (lane == 0)
                {
                    pendingBase = atomicAdd(&P.counters->chunkAlloc, SL_R2_FIFO_BLOCK);
                    if (pendingBase + SL_R2_FIFO_BLOCK > P.lists.chunkCapacity) { atomicOr(&P.counters->errorFlags, SL_ERR_POOL); pendingBase = SL_NO_CHUNK; }
                }
                pendingValid = true;
            }
            __syncwarp();

            /* ---- resolve: up to SL_R2_GROUPS groups per lane, the reference's loop body (src/SL_TriRasterizer.cpp:483-548) ---- */
            const int nIt = (int)__reduce_max_sync(0xffffffffu, go ? (unsigned)min(SL_R2_GROUPS, gEnd - g) : 0u);
            for (int it = 0; it < nIt; ++it)
            {
                if (go && g < gEnd)
                {
                    ++dbgGroups;
                    const int x4 = xMin + 4 * g;
                    const int rel = x4 - xBase - sx0; /* segment-relative x of the group's first pixel */
                    const bool last = g == gEnd - 1;
                    const float zc0 = sl_zdot(a0, z0, z1, z2), zc1 = sl_zdot(a1, z0, z1, z2), zc2 = sl_zdot(a2, z0, z1, z2), zc3 = sl_zdot(a3, z0, z1, z2);
                    const float zc[4] = {zc0, zc1, zc2, zc3};
                    /* the pixels of the group that belong to this piece: [max(p0 - x4, 0), min(p1 - x4, 4)) */
                    const int oLo = max(p0 - x4, 0), oHi = min(p1 - x4, 4);
                    const uint32_t own = ((1u << oHi) - 1u) & ~((1u << oLo) - 1u);
                    uint32_t m = 0u;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        if ((own >> j) & 1u)
                        {
                            const float d = FUNC != 0 ? sl_tile_load<BPT>(tile, rel + j) : 0.f;
                            if (sl_depth_func<FUNC>(zc[j], d)) m |= 1u << j;
                        }
                    }
                    const uint32_t tt = m | fM;
                    fM = 0u;
                    const bool cutInside = last && publish && ((p1 - xMin) >> 2) == g;
                    if (cutInside)
                    {
                        /* the cut lies inside this group: the registers before its compaction + the outcomes left of the cut */
                        const uint32_t below = (1u << (p1 - x4)) - 1u;
                        sl_handoff_publish(P, recId, ySrc, bOut, tt & below & 7u, a0, a1, a2, a3);
                    }
                    if (tt)
                    {
                        /* in-place compaction (:496-498) */
                        const math::vec4 n2 = (tt & 4u) ? a2 : a3;
                        const math::vec4 n1 = (tt & 2u) ? a1 : n2;
                        const math::vec4 n0 = (tt & 1u) ? a0 : n1;
                        if (m)
                        {
                            shaded += (unsigned)__popc(m);
                            /* positions in the fragment lists: a group's pixels fall into at most two sub-lists */
                            const int sFirst = rel >> SL_SHADE_SHIFT, sLast = (rel + 3) >> SL_SHADE_SHIFT;
                            const int jSplit = sLast != sFirst ? (((sFirst + 1) << SL_SHADE_SHIFT) - rel) : 4;
                            const uint32_t loBits = m & ((1u << jSplit) - 1u), hiBits = m & ~((1u << jSplit) - 1u);
                            uint32_t posLo = 0u, posHi = 0u;
                            if (loBits) posLo = atomicAdd(&subCount[sFirst], (uint32_t)__popc(loBits));
                            if (hiBits) posHi = atomicAdd(&subCount[sLast], (uint32_t)__popc(hiBits));
                            /* survivor number s of the group (the s-th passing lane, handed-over ones included) is shaded with
                             * register s after the compaction (:500-526) and with its own x and z; register s is known at
                             * compile time when the survivors are visited by number */
                            uint32_t rest = tt;
#pragma unroll
                            for (int sIdx = 0; sIdx < 4; ++sIdx)
                            {
                                if (rest)
                                {
                                    const int j = __ffs((int)rest) - 1;
                                    rest &= rest - 1u;
                                    if ((m >> j) & 1u)
                                    {
                                        const math::vec4& bc = sIdx == 0 ? n0 : (sIdx == 1 ? n1 : (sIdx == 2 ? n2 : a3));
                                        const float zj = j == 0 ? zc0 : (j == 1 ? zc1 : (j == 2 ? zc2 : zc3));
                                        if (P.depthMask) sl_tile_store<BPT>(tile, rel + j, zj); /* src/SL_FragmentProcessor.cpp:526-529 */
                                        const bool lo = j < jSplit;
                                        const uint32_t pos = lo ? posLo++ : posHi++;
                                        const int sub = lo ? sFirst : sLast;
                                        const uint32_t chunk = chunkRing[sub * SL_R2_CHUNK_RING + ((pos >> 5) & (SL_R2_CHUNK_RING - 1))];
                                        if (chunk != SL_NO_CHUNK)
                                        {
                                            unsigned char* c = P.lists.pool + (size_t)chunk * SL_CHUNK_BYTES;
                                            const uint32_t slot = pos & 31u;
                                            reinterpret_cast<float4*>(c)[slot] = make_float4(bc.v[0], bc.v[1], bc.v[2], bc.v[3]);
                                            reinterpret_cast<float*>(c + 512)[slot] = zj;
                                            reinterpret_cast<uint32_t*>(c + 640)[slot] = (uint32_t)(x4 + j) | ((uint32_t)ySrc << 16);
                                            reinterpret_cast<uint32_t*>(c + 768)[slot] = recId;
                                        }
                                    }
                                }
                            }
                        }
                        a0 = n0; a1 = n1; a2 = n2;
                    }
                    if (useDepth)
                    {
                        /* these pixels are resolved: later pieces may have them (lanes of one round touch disjoint pixels) */
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                        {
                            if ((own >> j) & 1u) pixMask[rel + j] &= ~myBit;
                        }
                    }
                    a0 += bcX; a1 += bcX; a2 += bcX; a3 += bcX; /* :539-542, one rounding per group */
                    if (last)
                    {
                        /* the cut falls on the group boundary: the registers after the compaction, advanced to the next group */
                        if (publish && !cutInside) sl_handoff_publish(P, recId, ySrc, bOut, 0u, a0, a1, a2, a3);
                        busy = false;
                    }
                    ++g;
                }
            }
            __syncwarp();
            /* enough pieces still on board (or nothing left to hand out): settle another layer before the next top */
            {
                const uint32_t stillBusy = __ballot_sync(0xffffffffu, busy);
                skipTop = stillBusy != 0u && (__popc(stillBusy) >= SL_R2_KEEP_LANES || (ringCount == 0u && base >= end));
            }
        }

        /* ---- end of the ticket ---- */
        dirty = __any_sync(0xffffffffu, shaded != 0u) && P.depthMask != 0;
        __syncwarp();
        /* the list's last chunk is the one that holds its last fragment: chunks are reserved for the pixels a round is about
         * to resolve, and the ones the depth test leaves empty stay linked behind it */
        if (lane < nsub && subCount[lane])
        {
            P.lists.listCount[firstList + lane] = subCount[lane];
            P.lists.listTail[firstList + lane] = chunkRing[lane * SL_R2_CHUNK_RING + (((subCount[lane] - 1u) >> 5) & (SL_R2_CHUNK_RING - 1))];
        }
        if (staged)
        {
            if (dirty)
            {
                /* write the depth segment back: one bulk store */
                if (bulkOk)
                {
                    sl_fence_async_smem();
                    __syncwarp();
                    if (lane == 0)
                    {
                        sl_bulk_store(tileSrc, tile, tileLoadBytes);
                        sl_bulk_store_wait_read();
                    }
                }
                else sl_copy_warp(tileSrc, tile, tileLoadBytes, lane);
            }
            __syncwarp();
        }
        for (int d = 16; d > 0; d >>= 1) shaded += __shfl_down_sync(0xffffffffu, shaded, d);
        if (lane == 0 && shaded) atomicAdd(&P.counters->fragsShaded, (unsigned long long)shaded);
        shaded = 0u;
        if (P.diagnostics && lane == 0)
        {
            /* the longest ticket: clocks << 32 | pieces << 18 | rounds */
            const unsigned long long cyc = (unsigned long long)(clock64() - dbgTicket0);
            atomicMax(&P.counters->resolveMaxWindows, (cyc << 32) | ((unsigned long long)min(dbgPieces, 16383u) << 18) | min(dbgRounds - dbgRounds0, 262143ull));
        }
    }

    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); /* the last depth store has landed */

    if (P.diagnostics)
    {
        for (int d = 16; d > 0; d >>= 1) dbgGroups += __shfl_down_sync(0xffffffffu, dbgGroups, d);
        if (lane == 0)
        {
            const unsigned long long cyc = (unsigned long long)(clock64() - dbgT0);
            atomicAdd(&P.counters->resolveWindows, dbgRounds);
            atomicAdd(&P.counters->resolveIdle, dbgIdle);
            atomicAdd(&P.counters->resolveGroups, dbgGroups);
            atomicAdd(&P.counters->resolveCycles, cyc);
            atomicMax(&P.counters->resolveMaxCycles, cyc);
            atomicAdd(&P.counters->resolveWarps, 1ull);
        }
    }
}

} // namespace sld
