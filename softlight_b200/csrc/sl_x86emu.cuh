/*
 * sl_x86emu.cuh -- device primitives that reproduce the x86 instructions the
 * reference's source path is written in:
 *
 *   rsqrtss      -> sl_rsqrt(): 2 x 2048-entry table (exponent parity, top 11 mantissa bits), measured the same way.
 *   rcpss/rcpps  -> sl_rcp():   2048-entry table indexed by the top 11 mantissa
 *                   bits, exponent handled exactly; the table is measured from
 *                   the host CPU at context creation (or pinned by a test).
 *   minps/maxps  -> sl_min()/sl_max(): (a<b)?a:b, second operand on NaN/equal (FSETP + FSEL, never FMNMX/.SAT).
 *   cvtps2dq     -> sl_cvt_rn(): RNE, 0x80000000 when out of range / NaN.
 *   cvttps2dq    -> sl_cvt_tr().
 *   fmadd/fmsub  -> __fmaf_rn (the TU is compiled with -fmad=false so that
 *                   nothing else is contracted).
 *   cmpps _OQ    -> plain C comparisons (false on NaN).
 *   cvtps2ph     -> __float2half_rn (scalar depth path) / _rz (vec4 colour path).
 */
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

namespace sld
{

/* One table per process: contexts that use a different table re-upload it
 * before their work is enqueued (slcu.cu: ensure_rcp_table). */
__device__ uint32_t g_rcp_table[2048];

__device__ __forceinline__ float sl_rcp(float x)
{
    const uint32_t xb = __float_as_uint(x);
    const uint32_t s = xb & 0x80000000u;
    const uint32_t e = (xb >> 23) & 0xffu;
    const uint32_t m = xb & 0x7fffffu;
    if (e == 0xffu) return __uint_as_float(m ? (xb | 0x00400000u) : s);
    if (e == 0u) return __uint_as_float(s | 0x7f800000u);
    const uint32_t r = __ldg(&g_rcp_table[m >> 12]);
    const int re = (int)((r >> 23) & 0xffu) - ((int)e - 127);
    if (re <= 0) return __uint_as_float(s);
    return __uint_as_float(s | ((uint32_t)re << 23) | (r & 0x7fffffu));
}

/* the same with the table at a caller-supplied address (a copy staged in shared memory) */
__device__ __forceinline__ float sl_rcp_tab(const uint32_t* __restrict__ table, float x)
{
    const uint32_t xb = __float_as_uint(x);
    const uint32_t s = xb & 0x80000000u;
    const uint32_t e = (xb >> 23) & 0xffu;
    const uint32_t m = xb & 0x7fffffu;
    if (e == 0xffu) return __uint_as_float(m ? (xb | 0x00400000u) : s);
    if (e == 0u) return __uint_as_float(s | 0x7f800000u);
    const uint32_t r = table[m >> 12];
    const int re = (int)((r >> 23) & 0xffu) - ((int)e - 127);
    if (re <= 0) return __uint_as_float(s);
    return __uint_as_float(s | ((uint32_t)re << 23) | (r & 0x7fffffu));
}

/* rsqrtss: exponent parity and the top 11 mantissa bits pick the entry, the exponent is exact (slcu.cu verifies this
 * structure against the host instruction when it measures the table) */
__device__ uint32_t g_rsqrt_table[4096];
__device__ __forceinline__ float sl_rsqrt(float x)
{
    const uint32_t xb = __float_as_uint(x);
    const uint32_t e = (xb >> 23) & 0xffu;
    const uint32_t m = xb & 0x7fffffu;
    if (e == 0xffu) return m ? __uint_as_float(xb | 0x00400000u) : ((xb >> 31) ? __uint_as_float(0xffc00000u) : 0.f);
    if (e == 0u) return __uint_as_float((xb & 0x80000000u) | 0x7f800000u);
    if (xb >> 31) return __uint_as_float(0xffc00000u);
    const int p = ((int)e - 127) & 1;
    const uint32_t r = __ldg(&g_rsqrt_table[(uint32_t)p * 2048u + (m >> 12)]);
    const int re = (int)((r >> 23) & 0xffu) - ((int)e - (127 + p)) / 2;
    return __uint_as_float(((uint32_t)re << 23) | (r & 0x7fffffu));
}

__device__ __forceinline__ float sl_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
/* minss/maxss: (a < b) ? a : b -- the SECOND operand when either is NaN or both are equal.  Written with the compare
 * result materialised as an integer: for the plain C form (and for setp/selp pairs) ptxas folds max(min(x, 1), 0) into
 * a .SAT modifier, which turns NaN into 0 where the x86 sequence gives 1 (seen on config C5: clamp() of the lighting
 * term of a zero-length normal).  This form compiles to FSETP + FSEL and is never folded. */
__device__ __forceinline__ float sl_min(float a, float b)
{
    unsigned c;
    asm("{.reg .pred p; setp.lt.f32 p, %1, %2; selp.u32 %0, 1, 0, p;}" : "=r"(c) : "f"(a), "f"(b));
    return c ? a : b;
}
__device__ __forceinline__ float sl_max(float a, float b)
{
    unsigned c;
    asm("{.reg .pred p; setp.gt.f32 p, %1, %2; selp.u32 %0, 1, 0, p;}" : "=r"(c) : "f"(a), "f"(b));
    return c ? a : b;
}
__device__ __forceinline__ bool sl_signbit(float x) { return (__float_as_uint(x) >> 31) != 0u; }

__device__ __forceinline__ int32_t sl_cvt_rn(float x)
{
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return (int32_t)0x80000000;
    return __float2int_rn(x);
}
__device__ __forceinline__ int32_t sl_cvt_tr(float x)
{
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return (int32_t)0x80000000;
    return __float2int_rz(x);
}

/* cvttss2si, 64-bit: "integer indefinite" on NaN / overflow */
__device__ __forceinline__ long long sl_cvt_tr64(float x)
{
    if (!(x >= -9223372036854775808.0f && x < 9223372036854775808.0f)) return (long long)0x8000000000000000ull;
    return __float2ll_rz(x);
}
/* vec4<uint32_t> -> vec4<float> of the reference: (float)(int)(v >> 1) + itself + (float)(int)(v & 1), every step rounded */
__device__ __forceinline__ float sl_u32_to_float(uint32_t v)
{
    const float v2f = __int2float_rn((int)(v >> 1));
    const float v1f = __int2float_rn((int)(v & 1u));
    return __fadd_rn(__fadd_rn(v2f, v2f), v1f);
}

__device__ __forceinline__ uint16_t sl_f2h_rn(float f) { return __half_as_ushort(__float2half_rn(f)); }
__device__ __forceinline__ uint16_t sl_f2h_rz(float f) { return __half_as_ushort(__float2half_rz(f)); }
__device__ __forceinline__ float sl_h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }

} // namespace sld
