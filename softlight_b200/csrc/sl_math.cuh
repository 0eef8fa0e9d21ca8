/*
 * sl_math.cuh -- sld::math: the __device__ mirror of the slice of ls::math that
 * user shaders and the pipeline use.  Same names and call shapes as the
 * reference (light_math/include/lightsky/math/...), so shader bodies port
 * textually; the arithmetic follows the reference's x86 implementation files
 * operation by operation (cited per function) so results are bit-identical.
 *
 * The translation unit is compiled with -fmad=false: `a*b+c` is two roundings;
 * only sl_fma() fuses (where the reference issues a real fmadd/fmsub).
 */
#pragma once

#include "sl_x86emu.cuh"

#define SLD_FN __device__ __forceinline__

namespace sld
{

struct half
{
    uint16_t bits;
    half() = default;
    SLD_FN explicit half(float f) : bits(sl_f2h_rn(f)) {}            /* x86/half_impl.h:24-31: RNE */
    SLD_FN explicit operator float() const { return sl_h2f(bits); } /* cvtph2ps */
};

namespace math
{
using sld::half;

/* ------------------------------------------------------------------ vec2 */
template <typename T> struct vec2_t
{
    T v[2];
    typedef T value_type;
    vec2_t() = default;
    SLD_FN constexpr vec2_t(T a, T b) : v{a, b} {}
    SLD_FN explicit constexpr vec2_t(T a) : v{a, a} {}
    SLD_FN T operator[](int i) const { return v[i]; }
    SLD_FN T& operator[](int i) { return v[i]; }
    SLD_FN vec2_t operator+(const vec2_t& o) const { return vec2_t{(T)(v[0] + o.v[0]), (T)(v[1] + o.v[1])}; }
    SLD_FN vec2_t operator-(const vec2_t& o) const { return vec2_t{(T)(v[0] - o.v[0]), (T)(v[1] - o.v[1])}; }
    SLD_FN vec2_t operator*(const vec2_t& o) const { return vec2_t{(T)(v[0] * o.v[0]), (T)(v[1] * o.v[1])}; }
    SLD_FN vec2_t operator*(T s) const { return vec2_t{(T)(v[0] * s), (T)(v[1] * s)}; }
    template <typename U> SLD_FN explicit operator vec2_t<U>() const { return vec2_t<U>{(U)v[0], (U)v[1]}; }
};

/* ------------------------------------------------------------------ vec3
 * generic/vec3_impl.h: plain element-wise C++ expressions. */
template <typename T> struct vec3_t
{
    T v[3];
    typedef T value_type;
    vec3_t() = default;
    SLD_FN constexpr vec3_t(T a, T b, T c) : v{a, b, c} {}
    SLD_FN explicit constexpr vec3_t(T a) : v{a, a, a} {}
    SLD_FN T operator[](int i) const { return v[i]; }
    SLD_FN T& operator[](int i) { return v[i]; }
    SLD_FN vec3_t operator+(const vec3_t& o) const { return vec3_t{(T)(v[0] + o.v[0]), (T)(v[1] + o.v[1]), (T)(v[2] + o.v[2])}; }
    SLD_FN vec3_t operator-(const vec3_t& o) const { return vec3_t{(T)(v[0] - o.v[0]), (T)(v[1] - o.v[1]), (T)(v[2] - o.v[2])}; }
    SLD_FN vec3_t operator*(const vec3_t& o) const { return vec3_t{(T)(v[0] * o.v[0]), (T)(v[1] * o.v[1]), (T)(v[2] * o.v[2])}; }
    SLD_FN vec3_t operator*(T s) const { return vec3_t{(T)(v[0] * s), (T)(v[1] * s), (T)(v[2] * s)}; }
    SLD_FN vec3_t operator/(T s) const { return vec3_t{(T)(v[0] / s), (T)(v[1] / s), (T)(v[2] / s)}; }
    SLD_FN vec3_t operator-() const { return vec3_t{(T)-v[0], (T)-v[1], (T)-v[2]}; }
    SLD_FN vec3_t& operator+=(const vec3_t& o) { *this = *this + o; return *this; }
    SLD_FN vec3_t& operator*=(T s) { *this = *this * s; return *this; }
    SLD_FN static constexpr unsigned num_components() { return 3; }
    template <typename U> SLD_FN explicit operator vec3_t<U>() const { return vec3_t<U>{(U)v[0], (U)v[1], (U)v[2]}; }
};

/* ------------------------------------------------------------------ vec4
 * vec4_t<float> is a union with __m128 in the reference (16-byte aligned,
 * x86/vec4f_impl.h:23); every operator is the matching packed instruction,
 * i.e. element-wise IEEE. */
template <typename T> struct vec4_align { static constexpr size_t value = alignof(T); };
template <> struct vec4_align<float> { static constexpr size_t value = 16; };

template <typename T> struct alignas(vec4_align<T>::value) vec4_t
{
    T v[4];
    typedef T value_type;
    vec4_t() = default;
    SLD_FN constexpr vec4_t(T a, T b, T c, T d) : v{a, b, c, d} {}
    SLD_FN explicit constexpr vec4_t(T a) : v{a, a, a, a} {}
    SLD_FN T operator[](int i) const { return v[i]; }
    SLD_FN T& operator[](int i) { return v[i]; }
    SLD_FN vec4_t operator+(const vec4_t& o) const { return vec4_t{(T)(v[0] + o.v[0]), (T)(v[1] + o.v[1]), (T)(v[2] + o.v[2]), (T)(v[3] + o.v[3])}; }
    SLD_FN vec4_t operator-(const vec4_t& o) const { return vec4_t{(T)(v[0] - o.v[0]), (T)(v[1] - o.v[1]), (T)(v[2] - o.v[2]), (T)(v[3] - o.v[3])}; }
    SLD_FN vec4_t operator*(const vec4_t& o) const { return vec4_t{(T)(v[0] * o.v[0]), (T)(v[1] * o.v[1]), (T)(v[2] * o.v[2]), (T)(v[3] * o.v[3])}; }
    SLD_FN vec4_t operator/(const vec4_t& o) const { return vec4_t{(T)(v[0] / o.v[0]), (T)(v[1] / o.v[1]), (T)(v[2] / o.v[2]), (T)(v[3] / o.v[3])}; }
    SLD_FN vec4_t operator+(T s) const { return vec4_t{(T)(v[0] + s), (T)(v[1] + s), (T)(v[2] + s), (T)(v[3] + s)}; }
    SLD_FN vec4_t operator-(T s) const { return vec4_t{(T)(v[0] - s), (T)(v[1] - s), (T)(v[2] - s), (T)(v[3] - s)}; }
    SLD_FN vec4_t operator*(T s) const { return vec4_t{(T)(v[0] * s), (T)(v[1] * s), (T)(v[2] * s), (T)(v[3] * s)}; }
    SLD_FN vec4_t operator/(T s) const { return vec4_t{(T)(v[0] / s), (T)(v[1] / s), (T)(v[2] / s), (T)(v[3] / s)}; }
    SLD_FN vec4_t operator-() const { return vec4_t{(T)-v[0], (T)-v[1], (T)-v[2], (T)-v[3]}; }
    SLD_FN vec4_t& operator+=(const vec4_t& o) { *this = *this + o; return *this; }
    SLD_FN vec4_t& operator-=(const vec4_t& o) { *this = *this - o; return *this; }
    SLD_FN vec4_t& operator*=(const vec4_t& o) { *this = *this * o; return *this; }
    SLD_FN vec4_t& operator*=(T s) { *this = *this * s; return *this; }
    SLD_FN static constexpr unsigned num_components() { return 4; }
    template <typename U> SLD_FN explicit operator vec4_t<U>() const { return vec4_t<U>{(U)v[0], (U)v[1], (U)v[2], (U)v[3]}; }
};

/* vec4<uint32_t> -> vec4<float> (x86/vec4f_impl.h:236-245): cvtdq2ps has no unsigned form, so the reference converts
 * v>>1 and v&1 separately: (float)(v>>1) * 2 + (float)(v&1) */
template <> template <>
SLD_FN vec4_t<uint32_t>::operator vec4_t<float>() const
{
    vec4_t<float> r;
#pragma unroll
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

/* float -> u8: cvtps2dq, packs_epi32, packus_epi16 (x86/vec4f_impl.h:299-309) */
SLD_FN vec4_t<uint8_t> vec4_to_u8(const vec4& f)
{
    uint8_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        int32_t s = sl_cvt_rn(f.v[i]);
        s = s < -32768 ? -32768 : (s > 32767 ? 32767 : s);
        r[i] = (uint8_t)(s < 0 ? 0 : (s > 255 ? 255 : s));
    }
    return vec4_t<uint8_t>{r[0], r[1], r[2], r[3]};
}

/* ------------------------------------------------------------------ mat4 */
template <typename T> struct alignas(16) mat4_t
{
    vec4_t<T> m[4]; /* columns */
    mat4_t() = default;
    /* generic/mat4_impl.h: mat4_t(n) = n on the diagonal */
    SLD_FN explicit constexpr mat4_t(T n) : m{vec4_t<T>{n, T(0), T(0), T(0)}, vec4_t<T>{T(0), n, T(0), T(0)}, vec4_t<T>{T(0), T(0), n, T(0)}, vec4_t<T>{T(0), T(0), T(0), n}} {}
    SLD_FN constexpr mat4_t(const vec4_t<T>& a, const vec4_t<T>& b, const vec4_t<T>& c, const vec4_t<T>& d) : m{a, b, c, d} {}
    /* generic/mat4_impl.h:205-213,474-482: column-wise, element-wise */
    SLD_FN mat4_t operator+(const mat4_t& n) const { return mat4_t{m[0] + n.m[0], m[1] + n.m[1], m[2] + n.m[2], m[3] + n.m[3]}; }
    SLD_FN mat4_t operator*(T n) const { return mat4_t{m[0] * n, m[1] * n, m[2] * n, m[3] * n}; }
    SLD_FN const vec4_t<T>& operator[](int i) const { return m[i]; }
    SLD_FN vec4_t<T>& operator[](int i) { return m[i]; }
    /* x86/mat4f_impl.h:123-146, FMA branch: (m0*x (+)fma m1*y) + (m2*z (+)fma m3*w) */
    SLD_FN vec4_t<T> operator*(const vec4_t<T>& s) const
    {
        vec4_t<T> r;
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const float v1 = sl_fma(m[1].v[i], s.v[1], m[0].v[i] * s.v[0]);
            const float v3 = sl_fma(m[3].v[i], s.v[3], m[2].v[i] * s.v[2]);
            r.v[i] = v1 + v3;
        }
        return r;
    }
    /* x86/mat4f_impl.h:16-41, AVX2 branch: per column a chain of one mul + three FMAs */
    SLD_FN mat4_t operator*(const mat4_t& n) const
    {
        mat4_t r;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                float a = m[0].v[i] * n.m[j].v[0];
                a = sl_fma(m[1].v[i], n.m[j].v[1], a);
                a = sl_fma(m[2].v[i], n.m[j].v[2], a);
                a = sl_fma(m[3].v[i], n.m[j].v[3], a);
                r.m[j].v[i] = a;
            }
        }
        return r;
    }
};
typedef mat4_t<float> mat4;

/* ---------------------------------------- scalars (x86/scalarf_utils_impl.h) */
SLD_FN float min(float a, float b) { return sl_min(a, b); }                           /* :24  minss */
SLD_FN float max(float a, float b) { return sl_max(a, b); }                           /* :56  maxss */
SLD_FN float clamp(float n, float lo, float hi) { return sl_max(sl_min(n, hi), lo); } /* :88 */
SLD_FN float saturate(float n) { return sl_max(sl_min(n, 1.f), 0.f); }                /* :98 */
SLD_FN float floor(float n) { return floorf(n); }                                     /* :109 roundps(floor) */
SLD_FN float round(float n) { return rintf(n); }                                      /* :139 roundps(nearest) */
SLD_FN float fmod_1(float n) { return n - (float)sl_cvt_tr(n); }                      /* :170 */
SLD_FN float rcp(float x) { return sl_rcp(x); }                                       /* :220 rcpss */
SLD_FN float fmadd(float x, float m, float a) { return sl_fma(x, m, a); }             /* :599 */
SLD_FN float inversesqrt(float x)                                                     /* :183-197: rsqrtss + one Newton step (FMA branch) */
{
    const float rsqr = sl_rsqrt(x);
    const float sqrr = x * rsqr;
    const float subr = sl_fma(-sqrr, rsqr, 3.f);
    const float hlfr = 0.5f * rsqr;
    return subr * hlfr;
}
SLD_FN int sign_mask(float x) { return sl_signbit(x) ? 1 : 0; }                       /* :250 */
SLD_FN float sqrt(float x) { return sqrtf(x); }
SLD_FN float abs(float x) { return __uint_as_float(__float_as_uint(x) & 0x7fffffffu); }
SLD_FN float min(float a, float b, float c) { return sl_min(c, sl_min(a, b)); }         /* :32-38 */
SLD_FN float max(float a, float b, float c) { return sl_max(c, sl_max(a, b)); }         /* :64-70 */
template <typename T> SLD_FN T clamp(T n, T lo, T hi) { return n < lo ? lo : (n > hi ? hi : n); }
template <typename T> SLD_FN T min(T a, T b) { return a < b ? a : b; }
template <typename T> SLD_FN T max(T a, T b) { return a > b ? a : b; }

/* ------------------------------------------- vec4 (x86/vecf_utils_impl.h) */
SLD_FN float hsum4(float a0, float a1, float a2, float a3) { return (a0 + a1) + (a2 + a3); }
SLD_FN float dot(const vec4& a, const vec4& b)                                        /* :154-168 */
{
    return hsum4(a.v[0] * b.v[0], a.v[1] * b.v[1], a.v[2] * b.v[2], a.v[3] * b.v[3]);
}
SLD_FN float sum(const vec4& a) { return hsum4(a.v[0], a.v[1], a.v[2], a.v[3]); }     /* :100 */
SLD_FN float length(const vec4& a) { return sqrtf(dot(a, a)); }                       /* :173 sqrtss */
SLD_FN vec4 normalize(const vec4& v)                                                   /* :192-207: v * rcpps(sqrtps(dot)) */
{
    const float r = sl_rcp(sqrtf(dot(v, v)));
    return vec4{v.v[0] * r, v.v[1] * r, v.v[2] * r, v.v[3] * r};
}
SLD_FN vec4 cross(const vec4& a, const vec4& b)                                        /* :139-149: shuffle(fmsub(a, b.yzx, a.yzx*b)) */
{
    const float c0 = sl_fma(a.v[0], b.v[1], -(a.v[1] * b.v[0]));
    const float c1 = sl_fma(a.v[1], b.v[2], -(a.v[2] * b.v[1]));
    const float c2 = sl_fma(a.v[2], b.v[0], -(a.v[0] * b.v[2]));
    const float c3 = sl_fma(a.v[3], b.v[3], -(a.v[3] * b.v[3]));
    return vec4{c1, c2, c0, c3};
}
SLD_FN vec4 mix(const vec4& a, const vec4& b, float t)                                 /* :212-222 */
{
    return vec4{sl_fma(b.v[0] - a.v[0], t, a.v[0]), sl_fma(b.v[1] - a.v[1], t, a.v[1]),
                sl_fma(b.v[2] - a.v[2], t, a.v[2]), sl_fma(b.v[3] - a.v[3], t, a.v[3])};
}
SLD_FN vec4 min(const vec4& a, const vec4& b) { return vec4{sl_min(a.v[0], b.v[0]), sl_min(a.v[1], b.v[1]), sl_min(a.v[2], b.v[2]), sl_min(a.v[3], b.v[3])}; }
SLD_FN vec4 max(const vec4& a, const vec4& b) { return vec4{sl_max(a.v[0], b.v[0]), sl_max(a.v[1], b.v[1]), sl_max(a.v[2], b.v[2]), sl_max(a.v[3], b.v[3])}; }
SLD_FN vec4 clamp(const vec4& v, const vec4& lo, const vec4& hi)                        /* :268: min(hi, max(v, lo)) */
{
    return vec4{sl_min(hi.v[0], sl_max(v.v[0], lo.v[0])), sl_min(hi.v[1], sl_max(v.v[1], lo.v[1])),
                sl_min(hi.v[2], sl_max(v.v[2], lo.v[2])), sl_min(hi.v[3], sl_max(v.v[3], lo.v[3]))};
}
SLD_FN vec4 saturate(const vec4& v)                                                    /* :276 */
{
    return vec4{sl_min(1.f, sl_max(v.v[0], 0.f)), sl_min(1.f, sl_max(v.v[1], 0.f)), sl_min(1.f, sl_max(v.v[2], 0.f)), sl_min(1.f, sl_max(v.v[3], 0.f))};
}
SLD_FN vec4 rcp(const vec4& v) { return vec4{sl_rcp(v.v[0]), sl_rcp(v.v[1]), sl_rcp(v.v[2]), sl_rcp(v.v[3])}; }
SLD_FN vec4 fmadd(const vec4& x, const vec4& m, const vec4& a)                          /* :556 */
{
    return vec4{sl_fma(x.v[0], m.v[0], a.v[0]), sl_fma(x.v[1], m.v[1], a.v[1]), sl_fma(x.v[2], m.v[2], a.v[2]), sl_fma(x.v[3], m.v[3], a.v[3])};
}
SLD_FN vec4 reflect(const vec4& v, const vec4& n)                                       /* :227-246 */
{
    const float dp2 = 2.f * dot(v, n);
    return vec4{v.v[0] - dp2 * n.v[0], v.v[1] - dp2 * n.v[1], v.v[2] - dp2 * n.v[2], v.v[3] - dp2 * n.v[3]};
}
SLD_FN vec4 floor(const vec4& v) { return vec4{floorf(v.v[0]), floorf(v.v[1]), floorf(v.v[2]), floorf(v.v[3])}; }

/* -------------------- generic vec3 / vec2 (generic/vec_utils_impl.h: left to right) */
template <typename T> SLD_FN T dot(const vec3_t<T>& a, const vec3_t<T>& b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
template <typename T> SLD_FN T dot(const vec2_t<T>& a, const vec2_t<T>& b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1]; }

/* vec casts (vec_utils.h:1579-1730) */
template <typename T> SLD_FN vec4_t<T> vec4_cast(const vec3_t<T>& v, const T& s) { return vec4_t<T>{v.v[0], v.v[1], v.v[2], s}; }
template <typename T> SLD_FN vec4_t<T> vec4_cast(const vec2_t<T>& v, const T& s0, const T& s1) { return vec4_t<T>{v.v[0], v.v[1], s0, s1}; }
template <typename T> SLD_FN vec3_t<T> vec3_cast(const vec4_t<T>& v) { return vec3_t<T>{v.v[0], v.v[1], v.v[2]}; }
template <typename T> SLD_FN vec3_t<T> vec3_cast(const vec2_t<T>& v, const T& s) { return vec3_t<T>{v.v[0], v.v[1], s}; }
template <typename T> SLD_FN vec2_t<T> vec2_cast(const vec4_t<T>& v) { return vec2_t<T>{v.v[0], v.v[1]}; }
template <typename T> SLD_FN vec2_t<T> vec2_cast(const vec3_t<T>& v) { return vec2_t<T>{v.v[0], v.v[1]}; }

/* generic vec2 min/max (generic/vec_utils_impl.h:120-147): (a < b) ? a : b / (a > b) ? a : b per element */
template <typename T> SLD_FN vec2_t<T> min(const vec2_t<T>& a, const vec2_t<T>& b) { return vec2_t<T>{a.v[0] < b.v[0] ? a.v[0] : b.v[0], a.v[1] < b.v[1] ? a.v[1] : b.v[1]}; }
template <typename T> SLD_FN vec2_t<T> max(const vec2_t<T>& a, const vec2_t<T>& b) { return vec2_t<T>{a.v[0] > b.v[0] ? a.v[0] : b.v[0], a.v[1] > b.v[1] ? a.v[1] : b.v[1]}; }
SLD_FN vec2_t<float> min(const vec2_t<float>& a, const vec2_t<float>& b) { return vec2_t<float>{sl_min(a.v[0], b.v[0]), sl_min(a.v[1], b.v[1])}; } /* same selection, not foldable into .SAT */
SLD_FN vec2_t<float> max(const vec2_t<float>& a, const vec2_t<float>& b) { return vec2_t<float>{sl_max(a.v[0], b.v[0]), sl_max(a.v[1], b.v[1])}; }

/* math::scale(mat4, vec3) (generic/mat_utils_impl.h:433-440): only the diagonal is multiplied */
template <typename T> SLD_FN mat4_t<T> scale(const mat4_t<T>& m, const vec3_t<T>& s)
{
    mat4_t<T> r = m;
    r.m[0].v[0] = m.m[0].v[0] * s.v[0];
    r.m[1].v[1] = m.m[1].v[1] * s.v[1];
    r.m[2].v[2] = m.m[2].v[2] * s.v[2];
    return r;
}

/* variadic sum (generic/scalar_utils_impl.h:1009-1023): right fold */
template <typename T> SLD_FN T sum_fold(const T& a) { return a; }
template <typename T, typename... R> SLD_FN T sum_fold(const T& a, const R&... rest) { return a + sum_fold(rest...); }
template <typename T, typename... R> SLD_FN T sum(const T& a, const T& b, const R&... rest) { return a + sum_fold(b, rest...); }

/* -------------------- log / exp / pow: the "fast approximate" forms that are live in the reference (LS_MATH_HIGH_PREC
 * is never defined; -mfma selects the fmadd branches).  Scalar (x86/scalarf_utils_impl.h:305-326,334-358,389-423,550-590)
 * and vec4 (x86/vecf_utils_impl.h:382-549) versions are lane-wise identical. */
SLD_FN float log2(float n)
{
    const uint32_t vxi = __float_as_uint(n);
    const float mxf = __uint_as_float((vxi & 0x007FFFFFu) | 0x3f000000u);
    const float a = sl_rcp(0.3520887068f + mxf);
    const float b = sl_fma(1.1920928955078125e-7f, (float)(int32_t)vxi, -124.22551499f);
    const float c = sl_fma(-1.498030302f, mxf, b);
    return sl_fma(-1.72587999f, a, c);
}
SLD_FN float log(float n) { return 0.69314718056f * log2(n); }
SLD_FN float exp2(float x)
{
    const float clipp = sl_max(x, -126.f);
    const float clipp2 = clipp + 126.94269504f;
    return __uint_as_float((uint32_t)sl_cvt_rn((float)(1 << 23) * clipp2));
}
SLD_FN float exp(float x) { return exp2(x * 1.4426950408889634f); }
SLD_FN float pow(float x, float y) { return exp(log(x) * y); }
SLD_FN vec4 log2(const vec4& n) { return vec4{log2(n.v[0]), log2(n.v[1]), log2(n.v[2]), log2(n.v[3])}; }
SLD_FN vec4 log(const vec4& n) { return vec4{log(n.v[0]), log(n.v[1]), log(n.v[2]), log(n.v[3])}; }
SLD_FN vec4 exp2(const vec4& x) { return vec4{exp2(x.v[0]), exp2(x.v[1]), exp2(x.v[2]), exp2(x.v[3])}; }
SLD_FN vec4 exp(const vec4& x) { return vec4{exp(x.v[0]), exp(x.v[1]), exp(x.v[2]), exp(x.v[3])}; }
SLD_FN vec4 pow(const vec4& x, const vec4& y) { return exp(log(x) * y); }

} // namespace math
} // namespace sld
