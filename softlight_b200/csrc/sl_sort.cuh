/* sl_sort.cuh -- the ordering step of the point / line / wireframe paths: a stable least-significant-digit radix sort of
 * 64-bit keys (optionally with 64-bit values) over a bit range, written for exactly what those paths need -- per texel,
 * candidates must stay in submission order, i.e. the sort has to be STABLE and only the texel bits (plus, for wireframe,
 * the ordinal bits) take part.
 *
 * Digits of up to 11 bits (the 21 texel bits of a 1080p target are two passes), three kernels per pass:
 *   k_sort_hist     a CTA counts the digits of its 4096 consecutive keys                    -> hist[digit][cta]
 *   k_sort_rows     CTA d turns row d of hist into its exclusive prefix over the CTAs and notes the row's total
 *   k_sort_scatter  every CTA scans the row totals (the digits' global bases) for itself; its eight warps own 512
 *                   consecutive keys each: a first sweep gives every (warp, digit) its destination (base of the digit +
 *                   the CTA's prefix + what the earlier warps hold), a second sweep walks the keys in order, 32 at a
 *                   time: lanes with equal digits find each other with match.any, their rank is their position among
 *                   those lanes, the running destination per (warp, digit) lives in shared memory.
 * Equal digits therefore leave a pass in the order they entered it. */
#pragma once

#include <stdint.h>

namespace sld
{

constexpr int SL_SORT_THREADS = 256;
constexpr int SL_SORT_ITEMS = 16;                                /* keys per thread */
constexpr int SL_SORT_TILE = SL_SORT_THREADS * SL_SORT_ITEMS;    /* keys per CTA    */
constexpr int SL_SORT_WARPS = SL_SORT_THREADS / 32;
constexpr int SL_SORT_WARP_KEYS = SL_SORT_TILE / SL_SORT_WARPS;  /* consecutive keys a warp owns */
constexpr int SL_SORT_MAX_BITS = 11;

struct SLSortPass
{
    const unsigned long long* keysIn;
    unsigned long long* keysOut;
    const unsigned long long* valsIn; /* null: keys only */
    unsigned long long* valsOut;
    uint32_t* hist;                   /* [numDigits][numTiles] */
    uint32_t* totals;                 /* [numDigits] */
    unsigned long long n;
    unsigned int numTiles;
    unsigned int numDigits;           /* 1 << bits of this pass */
    int shift;                        /* the digit is (key >> shift) & (numDigits - 1) */
};

__global__ void __launch_bounds__(SL_SORT_THREADS) k_sort_hist(const __grid_constant__ SLSortPass P)
{
    __shared__ uint32_t sHist[1 << SL_SORT_MAX_BITS];
    for (unsigned d = threadIdx.x; d < P.numDigits; d += SL_SORT_THREADS) sHist[d] = 0u;
    __syncthreads();
    const unsigned long long base = (unsigned long long)blockIdx.x * SL_SORT_TILE;
    const unsigned mask = P.numDigits - 1u;
#pragma unroll 4
    for (int i = 0; i < SL_SORT_ITEMS; ++i)
    {
        const unsigned long long idx = base + (unsigned long long)i * SL_SORT_THREADS + threadIdx.x;
        if (idx < P.n) atomicAdd(&sHist[(unsigned)(P.keysIn[idx] >> P.shift) & mask], 1u);
    }
    __syncthreads();
    for (unsigned d = threadIdx.x; d < P.numDigits; d += SL_SORT_THREADS) P.hist[(size_t)d * P.numTiles + blockIdx.x] = sHist[d];
}

__global__ void __launch_bounds__(SL_SORT_THREADS) k_sort_rows(uint32_t* hist, uint32_t* totals, unsigned int numTiles)
{
    __shared__ uint32_t sWarp[SL_SORT_WARPS];
    __shared__ uint32_t sCarry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* row = hist + (size_t)blockIdx.x * numTiles;
    if (tid == 0) sCarry = 0u;
    __syncthreads();
    for (unsigned base = 0; base < numTiles; base += SL_SORT_THREADS)
    {
        const unsigned i = base + (unsigned)tid;
        const uint32_t v = i < numTiles ? row[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) sWarp[warp] = incl;
        __syncthreads();
        uint32_t before = sCarry;
        for (int w = 0; w < warp; ++w) before += sWarp[w];
        if (i < numTiles) row[i] = before + incl - v;
        __syncthreads();
        if (tid == SL_SORT_THREADS - 1) sCarry = before + incl;
        __syncthreads();
    }
    if (tid == 0) totals[blockIdx.x] = sCarry;
}

__global__ void __launch_bounds__(SL_SORT_THREADS) k_sort_scatter(const __grid_constant__ SLSortPass P)
{
    extern __shared__ uint32_t sSortDyn[]; /* [warps][numDigits] counts, then running destinations; [numDigits] digit bases */
    __shared__ uint32_t sWarpSum[SL_SORT_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned ND = P.numDigits, mask = ND - 1u;
    uint32_t* sCount = sSortDyn;
    uint32_t* sBase = sSortDyn + (size_t)SL_SORT_WARPS * ND;
    for (unsigned i = tid; i < SL_SORT_WARPS * ND; i += SL_SORT_THREADS) sCount[i] = 0u;
    /* the digits' global bases: exclusive scan of the row totals (ND <= 2048: 8 per thread) */
    {
        const unsigned per = (ND + SL_SORT_THREADS - 1u) / SL_SORT_THREADS;
        const unsigned a = min((unsigned)tid * per, ND), b = min(a + per, ND);
        uint32_t sum = 0u;
        for (unsigned d = a; d < b; ++d) sum += P.totals[d];
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) sWarpSum[warp] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        for (int w = 0; w < warp; ++w) run += sWarpSum[w];
        for (unsigned d = a; d < b; ++d) { sBase[d] = run; run += P.totals[d]; }
    }
    __syncthreads();
    const unsigned long long warpBase = (unsigned long long)blockIdx.x * SL_SORT_TILE + (unsigned long long)warp * SL_SORT_WARP_KEYS;
    unsigned long long keys[SL_SORT_WARP_KEYS / 32];
#pragma unroll
    for (int s = 0; s < SL_SORT_WARP_KEYS / 32; ++s)
    {
        const unsigned long long idx = warpBase + (unsigned long long)s * 32 + lane;
        keys[s] = idx < P.n ? P.keysIn[idx] : 0ull;
        if (idx < P.n) atomicAdd(&sCount[(size_t)warp * ND + ((unsigned)(keys[s] >> P.shift) & mask)], 1u);
    }
    __syncthreads();
    /* digit d of this CTA starts at base[d] + hist[d][cta]; warp w's share of it follows the earlier warps' */
    for (unsigned d = tid; d < ND; d += SL_SORT_THREADS)
    {
        uint32_t run = sBase[d] + P.hist[(size_t)d * P.numTiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SL_SORT_WARPS; ++w)
        {
            const uint32_t c = sCount[(size_t)w * ND + d];
            sCount[(size_t)w * ND + d] = run;
            run += c;
        }
    }
    __syncthreads();
    uint32_t* mine = sCount + (size_t)warp * ND;
#pragma unroll
    for (int s = 0; s < SL_SORT_WARP_KEYS / 32; ++s)
    {
        const unsigned long long idx = warpBase + (unsigned long long)s * 32 + lane;
        const bool valid = idx < P.n;
        const unsigned d = valid ? ((unsigned)(keys[s] >> P.shift) & mask) : (0x10000u + (unsigned)lane); /* idle lanes: a digit nobody has */
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t dst = 0u;
        if (valid && lane == leader) dst = mine[d];
        dst = __shfl_sync(0xffffffffu, dst, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        __syncwarp();
        if (valid && lane == leader) mine[d] += (uint32_t)__popc(peers);
        __syncwarp();
        if (valid)
        {
            P.keysOut[dst] = keys[s];
            if (P.valsIn) P.valsOut[dst] = P.valsIn[idx];
        }
    }
}

} // namespace sld
