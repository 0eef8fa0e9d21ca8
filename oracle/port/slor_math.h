/*
 * slor_math.h -- TEST INFRASTRUCTURE ONLY (part of the CPU oracle, never used
 * by the product).
 *
 * Plain scalar C++ restatement of the slice of ls::math that the draw path and
 * the built-in shaders use, following the *x86 source path* of the reference
 * (the only one that compiles, SURVEY.md section 0 finding 3) operation by
 * operation: every function cites the intrinsic sequence it restates.  Rules:
 *   - compiled with -ffp-contract=off: a*b+c is two roundings unless written
 *     as sfma() (= a real FMA, where the reference uses _mm_fmadd_ps & co);
 *   - rcp() is the rcpss/rcpps table, a context-owned table of
 *     2048 result bit patterns for 1.m with the top 11 mantissa bits;
 *   - smin/smax are minps/maxps: (a<b)?a:b / (a>b)?a:b, i.e. the SECOND
 *     operand on NaN or equal;
 *   - float->int is cvtps2dq (round-to-nearest-even, 0x80000000 on overflow).
 */
#ifndef SLOR_MATH_H
#define SLOR_MATH_H

#include <cmath>
#include <cstdint>
#include <cstring>

namespace slo
{

extern uint32_t g_rcp_table[2048];

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* rcpss / rcpps (scalarf_utils_impl.h:220-223, _mm_rcp_ps everywhere). */
static inline float x86_rcp(float x)
{
    const uint32_t xb = f2u(x);
    const uint32_t s = xb & 0x80000000u;
    const uint32_t e = (xb >> 23) & 0xffu;
    const uint32_t m = xb & 0x7fffffu;
    if (e == 0xffu) return u2f(m ? (xb | 0x00400000u) : s);
    if (e == 0u) return u2f(s | 0x7f800000u);
    const uint32_t r = g_rcp_table[m >> 12];
    const int re = (int)((r >> 23) & 0xffu) - ((int)e - 127);
    if (re <= 0) return u2f(s);
    return u2f(s | ((uint32_t)re << 23) | (r & 0x7fffffu));
}

/* rsqrtss: a table over (exponent parity, top 11 mantissa bits) with exact exponent handling, measured from the host CPU
 * like the rcp table (entries: results for inputs in [1, 2) and [2, 4)) */
extern uint32_t g_rsqrt_table[4096];
static inline float x86_rsqrt(float x)
{
    const uint32_t xb = f2u(x);
    const uint32_t e = (xb >> 23) & 0xffu;
    const uint32_t m = xb & 0x7fffffu;
    if (e == 0xffu) return m ? u2f(xb | 0x00400000u) : ((xb >> 31) ? u2f(0xffc00000u) : 0.f); /* NaN quieted, -inf -> NaN, +inf -> 0 */
    if (e == 0u) return u2f((xb & 0x80000000u) | 0x7f800000u);                                   /* zero and denormals: signed infinity */
    if (xb >> 31) return u2f(0xffc00000u);
    const int p = ((int)e - 127) & 1;
    const uint32_t r = g_rsqrt_table[(uint32_t)p * 2048u + (m >> 12)];
    const int re = (int)((r >> 23) & 0xffu) - ((int)e - (127 + p)) / 2;
    return u2f(((uint32_t)re << 23) | (r & 0x7fffffu));
}

static inline float sfma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
static inline float smin(float a, float b) { return (a < b) ? a : b; }   /* minss/minps */
static inline float smax(float a, float b) { return (a > b) ? a : b; }   /* maxss/maxps */
static inline bool sign_bit(float x) { return (f2u(x) >> 31) != 0; }

/* cvtss2si / cvtps2dq: RNE, "integer indefinite" on NaN / overflow. */
static inline int32_t cvt_rn(float x)
{
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return INT32_MIN;
    return (int32_t)__builtin_nearbyintf(x);
}
/* cvttss2si / cvttps2dq */
static inline int32_t cvt_tr(float x)
{
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return INT32_MIN;
    return (int32_t)x;
}

/* cvtps2ph RNE (x86/half_impl.h:24-31) and cvtph2ps, in software. */
static inline uint16_t float_to_half_rn(float f)
{
    const uint32_t x = f2u(f);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((absx > 0x7f800000u) ? (0x0200u | ((absx >> 13) & 0x3ffu)) : 0u));
    if (absx >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u); /* rounds to >= 65520 -> inf */
    if (absx < 0x33000001u) return (uint16_t)sign;              /* <= 2^-25 -> 0 (ties to even) */
    int e = (int)(absx >> 23) - 127;
    uint32_t m = (absx & 0x7fffffu) | 0x800000u;
    int shift;
    uint32_t he;
    if (e < -14) { shift = 13 + (-14 - e); he = 0; }
    else { shift = 13; he = (uint32_t)(e + 15); }
    uint32_t hm = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u);
    const uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (hm & 1u))) hm++;
    if (he == 0) return (uint16_t)(sign | hm); /* denormal (hm may carry into exponent 1: correct) */
    return (uint16_t)(sign | (((he << 10) + (hm - 0x400u))));
}
static inline uint16_t float_to_half_rz(float f)
{
    const uint32_t x = f2u(f);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((absx > 0x7f800000u) ? (0x0200u | ((absx >> 13) & 0x3ffu)) : 0u));
    if (absx >= 0x47800000u) return (uint16_t)(sign | 0x7bffu); /* overflow truncates to max finite */
    int e = (int)(absx >> 23) - 127;
    if (e < -24) return (uint16_t)sign;
    uint32_t m = (absx & 0x7fffffu) | 0x800000u;
    if (e < -14) return (uint16_t)(sign | (m >> (13 + (-14 - e))));
    return (uint16_t)(sign | ((uint32_t)(e + 15) << 10) | ((m >> 13) & 0x3ffu));
}
static inline float half_to_float(uint16_t h)
{
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    const uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    if (e == 0x1fu) return u2f(sign | 0x7f800000u | (m << 13) | (m ? 0x00400000u : 0u));
    if (e == 0)
    {
        if (!m) return u2f(sign);
        int ee = -1;
        do { m <<= 1; ee++; } while (!(m & 0x400u));
        return u2f(sign | ((uint32_t)(127 - 15 - ee) << 23) | ((m & 0x3ffu) << 13));
    }
    return u2f(sign | ((e + 112u) << 23) | (m << 13));
}

struct half
{
    uint16_t bits;
    half() = default;
    explicit half(float f) : bits(float_to_half_rn(f)) {}
    explicit operator float() const { return half_to_float(bits); }
};

namespace math
{
using slo::half;

template <typename T> struct vec2_t
{
    T v[2];
    typedef T value_type;
    vec2_t() = default;
    constexpr vec2_t(T a, T b) : v{a, b} {}
    explicit constexpr vec2_t(T a) : v{a, a} {}
    T operator[](int i) const { return v[i]; }
    T& operator[](int i) { return v[i]; }
    vec2_t operator+(const vec2_t& o) const { return vec2_t{(T)(v[0] + o.v[0]), (T)(v[1] + o.v[1])}; }
    vec2_t operator-(const vec2_t& o) const { return vec2_t{(T)(v[0] - o.v[0]), (T)(v[1] - o.v[1])}; }
    vec2_t operator*(const vec2_t& o) const { return vec2_t{(T)(v[0] * o.v[0]), (T)(v[1] * o.v[1])}; }
    vec2_t operator*(T s) const { return vec2_t{(T)(v[0] * s), (T)(v[1] * s)}; }
    template <typename U> explicit operator vec2_t<U>() const { return vec2_t<U>{(U)v[0], (U)v[1]}; }
};

template <typename T> struct vec3_t
{
    T v[3];
    typedef T value_type;
    vec3_t() = default;
    constexpr vec3_t(T a, T b, T c) : v{a, b, c} {}
    explicit constexpr vec3_t(T a) : v{a, a, a} {}
    T operator[](int i) const { return v[i]; }
    T& operator[](int i) { return v[i]; }
    /* generic/vec3_impl.h: plain element-wise expressions */
    vec3_t operator+(const vec3_t& o) const { return vec3_t{(T)(v[0] + o.v[0]), (T)(v[1] + o.v[1]), (T)(v[2] + o.v[2])}; }
    vec3_t operator-(const vec3_t& o) const { return vec3_t{(T)(v[0] - o.v[0]), (T)(v[1] - o.v[1]), (T)(v[2] - o.v[2])}; }
    vec3_t operator*(const vec3_t& o) const { return vec3_t{(T)(v[0] * o.v[0]), (T)(v[1] * o.v[1]), (T)(v[2] * o.v[2])}; }
    vec3_t operator*(T s) const { return vec3_t{(T)(v[0] * s), (T)(v[1] * s), (T)(v[2] * s)}; }
    vec3_t operator/(T s) const { return vec3_t{(T)(v[0] / s), (T)(v[1] / s), (T)(v[2] / s)}; }
    vec3_t operator-() const { return vec3_t{(T)-v[0], (T)-v[1], (T)-v[2]}; }
    vec3_t& operator+=(const vec3_t& o) { *this = *this + o; return *this; }
    vec3_t& operator*=(T s) { *this = *this * s; return *this; }
    static constexpr unsigned num_components() { return 3; }
    template <typename U> explicit operator vec3_t<U>() const { return vec3_t<U>{(U)v[0], (U)v[1], (U)v[2]}; }
};

template <typename T> struct vec4_align { static constexpr size_t value = alignof(T); };
template <> struct vec4_align<float> { static constexpr size_t value = 16; }; /* union with __m128 (x86/vec4f_impl.h:23) */

template <typename T> struct alignas(vec4_align<T>::value) vec4_t
{
    T v[4];
    typedef T value_type;
    vec4_t() = default;
    constexpr vec4_t(T a, T b, T c, T d) : v{a, b, c, d} {}
    explicit constexpr vec4_t(T a) : v{a, a, a, a} {}
    T operator[](int i) const { return v[i]; }
    T& operator[](int i) { return v[i]; }
    vec4_t operator+(const vec4_t& o) const { return vec4_t{(T)(v[0] + o.v[0]), (T)(v[1] + o.v[1]), (T)(v[2] + o.v[2]), (T)(v[3] + o.v[3])}; }
    vec4_t operator-(const vec4_t& o) const { return vec4_t{(T)(v[0] - o.v[0]), (T)(v[1] - o.v[1]), (T)(v[2] - o.v[2]), (T)(v[3] - o.v[3])}; }
    vec4_t operator*(const vec4_t& o) const { return vec4_t{(T)(v[0] * o.v[0]), (T)(v[1] * o.v[1]), (T)(v[2] * o.v[2]), (T)(v[3] * o.v[3])}; }
    vec4_t operator/(const vec4_t& o) const { return vec4_t{(T)(v[0] / o.v[0]), (T)(v[1] / o.v[1]), (T)(v[2] / o.v[2]), (T)(v[3] / o.v[3])}; }
    vec4_t operator+(T s) const { return vec4_t{(T)(v[0] + s), (T)(v[1] + s), (T)(v[2] + s), (T)(v[3] + s)}; }
    vec4_t operator-(T s) const { return vec4_t{(T)(v[0] - s), (T)(v[1] - s), (T)(v[2] - s), (T)(v[3] - s)}; }
    vec4_t operator*(T s) const { return vec4_t{(T)(v[0] * s), (T)(v[1] * s), (T)(v[2] * s), (T)(v[3] * s)}; }
    vec4_t operator/(T s) const { return vec4_t{(T)(v[0] / s), (T)(v[1] / s), (T)(v[2] / s), (T)(v[3] / s)}; }
    /* x86/vec4f_impl.h operator-(): _mm_xor_ps with -0.f for float == IEEE negate */
    vec4_t operator-() const { return vec4_t{(T)-v[0], (T)-v[1], (T)-v[2], (T)-v[3]}; }
    vec4_t& operator+=(const vec4_t& o) { *this = *this + o; return *this; }
    vec4_t& operator-=(const vec4_t& o) { *this = *this - o; return *this; }
    vec4_t& operator*=(const vec4_t& o) { *this = *this * o; return *this; }
    vec4_t& operator*=(T s) { *this = *this * s; return *this; }
    static constexpr unsigned num_components() { return 4; }
    template <typename U> explicit operator vec4_t<U>() const;
};

/* generic vec4 cast: element-wise C cast (vec4f_impl.h:184-187) ... */
template <typename T> template <typename U>
inline vec4_t<T>::operator vec4_t<U>() const { return vec4_t<U>{(U)v[0], (U)v[1], (U)v[2], (U)v[3]}; }

/* ... except float->integer, which is cvtps2dq + saturating packs (vec4f_impl.h:258-340) */
static inline int32_t sat_i16(int32_t x) { return x < -32768 ? -32768 : (x > 32767 ? 32767 : x); }
template <> template <>
inline vec4_t<float>::operator vec4_t<uint8_t>() const
{
    uint8_t r[4];
    for (int i = 0; i < 4; ++i)
    {
        const int32_t s = sat_i16(cvt_rn(v[i])); /* packs_epi32 */
        r[i] = (uint8_t)(s < 0 ? 0 : (s > 255 ? 255 : s)); /* packus_epi16 */
    }
    return vec4_t<uint8_t>{r[0], r[1], r[2], r[3]};
}
template <> template <>
inline vec4_t<float>::operator vec4_t<int32_t>() const
{
    return vec4_t<int32_t>{cvt_rn(v[0]), cvt_rn(v[1]), cvt_rn(v[2]), cvt_rn(v[3])};
}
/* vec4<float> -> vec4<half>: _mm_cvtps_ph(..., _MM_FROUND_TO_ZERO) (vec4f_impl.h:348-355) */
template <> template <>
inline vec4_t<float>::operator vec4_t<half>() const
{
    vec4_t<half> r;
    for (int i = 0; i < 4; ++i) r.v[i].bits = float_to_half_rz(v[i]);
    return r;
}
template <> template <>
inline vec4_t<half>::operator vec4_t<float>() const
{
    return vec4_t<float>{half_to_float(v[0].bits), half_to_float(v[1].bits), half_to_float(v[2].bits), half_to_float(v[3].bits)};
}

/* vec4<uint32_t> -> vec4<float> (vec4f_impl.h:236-245): cvtdq2ps has no unsigned form, so the reference converts v>>1
 * and v&1 separately: (float)(v>>1) * 2 + (float)(v&1), two roundings for values >= 2^24 */
template <> template <>
inline vec4_t<uint32_t>::operator vec4_t<float>() const
{
    vec4_t<float> r;
    for (int i = 0; i < 4; ++i)
    {
        const float hi = (float)(int32_t)(v[i] >> 1), lo = (float)(int32_t)(v[i] & 1u);
        r.v[i] = (hi + hi) + lo;
    }
    return r;
}

typedef vec2_t<float> vec2;
typedef vec3_t<float> vec3;
typedef vec4_t<float> vec4;

/* mat4: 4 columns; m[col][row] */
template <typename T> struct mat4_t
{
    vec4_t<T> m[4];
    mat4_t() = default;
    /* generic/mat4_impl.h: mat4_t(n) = n on the diagonal */
    explicit constexpr mat4_t(T n) : m{vec4_t<T>{n, T(0), T(0), T(0)}, vec4_t<T>{T(0), n, T(0), T(0)}, vec4_t<T>{T(0), T(0), n, T(0)}, vec4_t<T>{T(0), T(0), T(0), n}} {}
    constexpr mat4_t(const vec4_t<T>& a, const vec4_t<T>& b, const vec4_t<T>& c, const vec4_t<T>& d) : m{a, b, c, d} {}
    /* generic/mat4_impl.h:205-213,474-482: column-wise, element-wise */
    mat4_t operator+(const mat4_t& n) const { return mat4_t{m[0] + n.m[0], m[1] + n.m[1], m[2] + n.m[2], m[3] + n.m[3]}; }
    mat4_t operator*(T n) const { return mat4_t{m[0] * n, m[1] * n, m[2] * n, m[3] * n}; }
    const vec4_t<T>& operator[](int i) const { return m[i]; }
    vec4_t<T>& operator[](int i) { return m[i]; }
    /* x86/mat4f_impl.h:123-146 (FMA branch): v1 = fma(m1,y, m0*x); v3 = fma(m3,w, m2*z); v1+v3 */
    vec4_t<T> operator*(const vec4_t<T>& s) const
    {
        vec4_t<T> r;
        for (int i = 0; i < 4; ++i)
        {
            const float v0 = m[0].v[i] * s.v[0];
            const float v1 = sfma(m[1].v[i], s.v[1], v0);
            const float v2 = m[2].v[i] * s.v[2];
            const float v3 = sfma(m[3].v[i], s.v[3], v2);
            r.v[i] = v1 + v3;
        }
        return r;
    }
    /* x86/mat4f_impl.h:16-41 (AVX2 branch): r = col0*n0; r = fma(col1,n1,r); fma(col2,n2,r); fma(col3,n3,r) */
    mat4_t operator*(const mat4_t& n) const
    {
        mat4_t r;
        for (int j = 0; j < 4; ++j)
        {
            for (int i = 0; i < 4; ++i)
            {
                float a = m[0].v[i] * n.m[j].v[0];
                a = sfma(m[1].v[i], n.m[j].v[1], a);
                a = sfma(m[2].v[i], n.m[j].v[2], a);
                a = sfma(m[3].v[i], n.m[j].v[3], a);
                r.m[j].v[i] = a;
            }
        }
        return r;
    }
};
typedef mat4_t<float> mat4;

/* ---- scalar (x86/scalarf_utils_impl.h) ---- */
static inline float min(float a, float b) { return smin(a, b); }              /* :24 */
static inline float max(float a, float b) { return smax(a, b); }              /* :56 */
static inline float clamp(float n, float lo, float hi) { return smax(smin(n, hi), lo); } /* :88 */
static inline float saturate(float n) { return smax(smin(n, 1.f), 0.f); }     /* :98 (n^n == +0) */
static inline float floor(float n) { return __builtin_floorf(n); }            /* :109 roundps */
static inline float round(float n) { return __builtin_nearbyintf(n); }        /* :139 roundps nearest-even */
static inline float fmod_1(float n) { return n - (float)cvt_tr(n); }          /* :170 */
static inline float rcp(float x) { return x86_rcp(x); }                       /* :220 */
static inline float fmadd(float x, float m, float a) { return sfma(x, m, a); } /* :599 */
static inline float inversesqrt(float x)                                      /* :183-197: rsqrtss + one Newton step, FMA branch */
{
    const float rsqr = x86_rsqrt(x);
    const float sqrr = x * rsqr;
    const float subr = sfma(-sqrr, rsqr, 3.f);
    const float hlfr = 0.5f * rsqr;
    return subr * hlfr;
}
static inline int sign_mask(float x) { return sign_bit(x) ? 1 : 0; }          /* :250 */
static inline float sqrt(float x) { return __builtin_sqrtf(x); }
static inline float abs(float x) { return u2f(f2u(x) & 0x7fffffffu); }
static inline float min(float a, float b, float c) { return smin(c, smin(a, b)); } /* :32-38 */
static inline float max(float a, float b, float c) { return smax(c, smax(a, b)); } /* :64-70 */
template <typename T> static inline T clamp(T n, T lo, T hi) { return n < lo ? lo : (n > hi ? hi : n); } /* generic ints */
template <typename T> static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T max(T a, T b) { return a > b ? a : b; }

/* ---- vec4 (x86/vecf_utils_impl.h) ---- */
static inline float hsum4(float a0, float a1, float a2, float a3) { return (a0 + a1) + (a2 + a3); }
static inline float dot(const vec4& a, const vec4& b)                         /* :154-168 */
{
    return hsum4(a.v[0] * b.v[0], a.v[1] * b.v[1], a.v[2] * b.v[2], a.v[3] * b.v[3]);
}
static inline float sum(const vec4& a) { return hsum4(a.v[0], a.v[1], a.v[2], a.v[3]); }   /* :100 */
static inline float length(const vec4& a) { return __builtin_sqrtf(dot(a, a)); }       /* :173 */
static inline vec4 normalize(const vec4& v)                                   /* :192-207 */
{
    const float e = hsum4(v.v[0] * v.v[0], v.v[1] * v.v[1], v.v[2] * v.v[2], v.v[3] * v.v[3]);
    const float r = x86_rcp(__builtin_sqrtf(e));
    return vec4{v.v[0] * r, v.v[1] * r, v.v[2] * r, v.v[3] * r};
}
static inline vec4 cross(const vec4& a, const vec4& b)                        /* :139-149 */
{
    /* c = fmsub(a, b.yzx, a.yzx*b), then c.yzx */
    float c[4];
    const int yzx[4] = {1, 2, 0, 3};
    for (int i = 0; i < 4; ++i) c[i] = sfma(a.v[i], b.v[yzx[i]], -(a.v[yzx[i]] * b.v[i]));
    return vec4{c[1], c[2], c[0], c[3]};
}
static inline vec4 mix(const vec4& a, const vec4& b, float t)                 /* :212-222 */
{
    vec4 r;
    for (int i = 0; i < 4; ++i) r.v[i] = sfma(b.v[i] - a.v[i], t, a.v[i]);
    return r;
}
static inline vec4 min(const vec4& a, const vec4& b) { return vec4{smin(a.v[0], b.v[0]), smin(a.v[1], b.v[1]), smin(a.v[2], b.v[2]), smin(a.v[3], b.v[3])}; }
static inline vec4 max(const vec4& a, const vec4& b) { return vec4{smax(a.v[0], b.v[0]), smax(a.v[1], b.v[1]), smax(a.v[2], b.v[2]), smax(a.v[3], b.v[3])}; }
static inline vec4 clamp(const vec4& v, const vec4& lo, const vec4& hi)        /* :268: min(hi, max(v, lo)) */
{
    vec4 r;
    for (int i = 0; i < 4; ++i) r.v[i] = smin(hi.v[i], smax(v.v[i], lo.v[i]));
    return r;
}
static inline vec4 saturate(const vec4& v)                                    /* :276 */
{
    vec4 r;
    for (int i = 0; i < 4; ++i) r.v[i] = smin(1.f, smax(v.v[i], 0.f));
    return r;
}
static inline vec4 rcp(const vec4& v) { return vec4{x86_rcp(v.v[0]), x86_rcp(v.v[1]), x86_rcp(v.v[2]), x86_rcp(v.v[3])}; }
static inline vec4 fmadd(const vec4& x, const vec4& m, const vec4& a)         /* :556 */
{
    return vec4{sfma(x.v[0], m.v[0], a.v[0]), sfma(x.v[1], m.v[1], a.v[1]), sfma(x.v[2], m.v[2], a.v[2]), sfma(x.v[3], m.v[3], a.v[3])};
}
static inline vec4 reflect(const vec4& v, const vec4& n)                      /* :227-246 */
{
    const float dp = dot(v, n);
    vec4 r;
    for (int i = 0; i < 4; ++i) r.v[i] = v.v[i] - ((2.f * dp) * n.v[i]);
    return r;
}
static inline vec4 floor(const vec4& v) { return vec4{floor(v.v[0]), floor(v.v[1]), floor(v.v[2]), floor(v.v[3])}; }

/* ---- generic vec3/vec2 (generic/vec_utils_impl.h: plain left-to-right expressions) ---- */
template <typename T> static inline T dot(const vec3_t<T>& a, const vec3_t<T>& b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
template <typename T> static inline T dot(const vec2_t<T>& a, const vec2_t<T>& b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1]; }

/* vec casts (vec_utils.h:1579-1730) */
template <typename T> static inline vec4_t<T> vec4_cast(const vec3_t<T>& v, const T& s) { return vec4_t<T>{v.v[0], v.v[1], v.v[2], s}; }
template <typename T> static inline vec4_t<T> vec4_cast(const vec2_t<T>& v, const T& s0, const T& s1) { return vec4_t<T>{v.v[0], v.v[1], s0, s1}; }
template <typename T> static inline vec3_t<T> vec3_cast(const vec4_t<T>& v) { return vec3_t<T>{v.v[0], v.v[1], v.v[2]}; }
template <typename T> static inline vec3_t<T> vec3_cast(const vec2_t<T>& v, const T& s) { return vec3_t<T>{v.v[0], v.v[1], s}; }
template <typename T> static inline vec2_t<T> vec2_cast(const vec4_t<T>& v) { return vec2_t<T>{v.v[0], v.v[1]}; }
template <typename T> static inline vec2_t<T> vec2_cast(const vec3_t<T>& v) { return vec2_t<T>{v.v[0], v.v[1]}; }

/* generic vec2 min/max (generic/vec_utils_impl.h:120-147): math::min<num_t>/max<num_t> with an explicit template
 * argument = the generic (a < b) ? a : b / (a > b) ? a : b */
template <typename T> static inline vec2_t<T> min(const vec2_t<T>& a, const vec2_t<T>& b) { return vec2_t<T>{a.v[0] < b.v[0] ? a.v[0] : b.v[0], a.v[1] < b.v[1] ? a.v[1] : b.v[1]}; }
template <typename T> static inline vec2_t<T> max(const vec2_t<T>& a, const vec2_t<T>& b) { return vec2_t<T>{a.v[0] > b.v[0] ? a.v[0] : b.v[0], a.v[1] > b.v[1] ? a.v[1] : b.v[1]}; }

/* math::scale(mat4, vec3) (generic/mat_utils_impl.h:433-440): only the diagonal is multiplied */
template <typename T> static inline mat4_t<T> scale(const mat4_t<T>& m, const vec3_t<T>& s)
{
    mat4_t<T> r = m;
    r.m[0].v[0] = m.m[0].v[0] * s.v[0];
    r.m[1].v[1] = m.m[1].v[1] * s.v[1];
    r.m[2].v[2] = m.m[2].v[2] * s.v[2];
    return r;
}

/* variadic sum (generic/scalar_utils_impl.h:1009-1023): right fold */
template <typename T> static inline T sum_fold(const T& a) { return a; }
template <typename T, typename... R> static inline T sum_fold(const T& a, const R&... rest) { return a + sum_fold(rest...); }
template <typename T, typename... R> static inline T sum(const T& a, const T& b, const R&... rest) { return a + sum_fold(b, rest...); }

/* -------------------- log / exp / pow: the "fast approximate" forms that are live in the reference (LS_MATH_HIGH_PREC
 * is never defined; -mfma selects the fmadd branches).  Scalar (x86/scalarf_utils_impl.h:305-326,334-358,389-423,550-590)
 * and vec4 (x86/vecf_utils_impl.h:382-549) versions are lane-wise identical. */
static inline float log2(float n)
{
    const uint32_t vxi = f2u(n);
    const float mxf = u2f((vxi & 0x007FFFFFu) | 0x3f000000u);
    const float a = rcp(0.3520887068f + mxf);
    const float b = sfma(1.1920928955078125e-7f, (float)(int32_t)vxi, -124.22551499f);
    const float c = sfma(-1.498030302f, mxf, b);
    return sfma(-1.72587999f, a, c);
}
static inline float log(float n) { return 0.69314718056f * log2(n); }
static inline float exp2(float x)
{
    const float clipp = smax(x, -126.f);
    const float clipp2 = clipp + 126.94269504f;
    return u2f((uint32_t)cvt_rn((float)(1 << 23) * clipp2));
}
static inline float exp(float x) { return exp2(x * 1.4426950408889634f); }
static inline float pow(float x, float y) { return exp(log(x) * y); }
static inline vec4 log2(const vec4& n) { return vec4{log2(n.v[0]), log2(n.v[1]), log2(n.v[2]), log2(n.v[3])}; }
static inline vec4 log(const vec4& n) { return vec4{log(n.v[0]), log(n.v[1]), log(n.v[2]), log(n.v[3])}; }
static inline vec4 exp2(const vec4& x) { return vec4{exp2(x.v[0]), exp2(x.v[1]), exp2(x.v[2]), exp2(x.v[3])}; }
static inline vec4 exp(const vec4& x) { return vec4{exp(x.v[0]), exp(x.v[1]), exp(x.v[2]), exp(x.v[3])}; }
static inline vec4 pow(const vec4& x, const vec4& y) { return exp(log(x) * y); }

} // namespace math
} // namespace slo

#endif
