/* TEST INFRASTRUCTURE.  Compiles the product's device header softlight_b200/csrc/sl_texel.cuh for the HOST, so that the
 * texel conversions (all 33 x 33 blit pairs, the packed render-target casts) can be checked against the oracle without a
 * GPU (tests/test_texel_host_check.py).  The CUDA intrinsics the header uses are given their IEEE meaning here; the x86
 * emulation helpers (sl_cvt_*, half conversions) come from the oracle's math header.  Built with -ffp-contract=off. */
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../oracle/port/slor_math.h"

#define SL_TEXEL_HOST_CHECK 1
#define SLD_FN static inline

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline int __double2int_rz(double x) { return (int)x; }
static inline long long __double2ll_rz(double x) { return (long long)x; }
static inline float __ull2float_rn(unsigned long long v) { return (float)v; }
static inline float __uint2float_rn(unsigned v) { return (float)v; }
static inline float __int2float_rn(int v) { return (float)v; }
static inline double __ull2double_rn(unsigned long long v) { return (double)v; }
static inline float __double2float_rn(double x) { return (float)x; }

namespace sld
{
namespace math
{
struct vec4 { float v[4]; };
} // namespace math
static inline int32_t sl_cvt_rn(float x) { return slo::cvt_rn(x); }
static inline int32_t sl_cvt_tr(float x) { return slo::cvt_tr(x); }
static inline long long sl_cvt_tr64(float x)
{
    if (!(x >= -9223372036854775808.0f && x < 9223372036854775808.0f)) return (long long)0x8000000000000000ull;
    return (long long)x;
}
static inline float sl_u32_to_float(uint32_t v)
{
    const float v2f = (float)(int)(v >> 1), v1f = (float)(int)(v & 1u);
    return (v2f + v2f) + v1f;
}
static inline uint16_t sl_f2h_rn(float f) { return slo::float_to_half_rn(f); }
static inline uint16_t sl_f2h_rz(float f) { return slo::float_to_half_rz(f); }
static inline float sl_h2f(uint16_t h) { return slo::half_to_float(h); }
} // namespace sld

#include "../../softlight_b200/csrc/sl_texel.cuh"

extern "C" {
/* n texels, converted one by one (no resampling, no flip) */
void texel_blit(const unsigned char* src, int srcType, unsigned char* dst, int dstType, size_t n)
{
    const int sb = sld::sl_texel_bytes(srcType), db = sld::sl_texel_bytes(dstType);
    for (size_t i = 0; i < n; ++i) sld::sl_blit_texel(src + i * sb, srcType, dst + i * db, dstType);
}
uint32_t texel_packed_from_float(int type, const float rgba[4])
{
    return sld::sl_packed_from_float(type - 28, sld::math::vec4{{rgba[0], rgba[1], rgba[2], rgba[3]}});
}
void texel_float_from_packed(int type, uint32_t v, float out[4])
{
    const sld::math::vec4 c = sld::sl_float_from_packed(type - 28, v);
    memcpy(out, c.v, 16);
}
}
