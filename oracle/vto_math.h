/*
 * vto_math.h -- TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
 *
 * Scalar + small-vector arithmetic of the GLSL built-ins the reference's
 * shaders call, restated in plain C on IEEE binary32 with a FIXED operation
 * order, so that the CPU oracle and the CUDA kernels (which restate the same
 * formulas independently in voxeltoy_b200/csrc/vt_math.cuh) agree bit for bit.
 *
 * GLSL leaves the precision of sin/cos/acos/atan/pow and the evaluation order
 * of dot/cross/normalize/mat*vec to the driver. The contract here (DESIGN.md
 * "Arithmetic contract"):
 *   - every + - * / sqrt is a single correctly rounded binary32 operation,
 *     never fused (build with -ffp-contract=off);
 *   - pow(x,y) = exp2(y*log2(x)) -- the GLSL specification's own definition;
 *   - sin/cos: 3-constant Cody-Waite reduction by pi/2 + degree-7/8 minimax
 *     polynomials; acos/atan: the classic single-precision Cephes reductions.
 *     All max ~2 ulp versus libm on the ranges the shaders use (tests/
 *     test_oracle_math.py pins this).
 *   - min/max/step/sign/mix/mod/clamp follow the GLSL 4.30 spec text
 *     (section 8.3), including their behaviour on NaN that falls out of the
 *     spec's "y < x ? y : x" wording.
 *   - float->int conversion saturates and maps NaN to 0 (what both x86 SSE
 *     after our explicit guard and PTX cvt.rzi.s32.f32 produce).
 */
#ifndef VTO_MATH_H
#define VTO_MATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;

#define VTO_PI        3.14159265359f   /* shaders/shared/constants.h:1 */
#define VTO_TWO_PI    6.28318530718f   /* constants.h:2 */
#define VTO_INV_TWOPI 0.15915494309f   /* constants.h:3 */

static inline uint32_t vto_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    vto_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* ---- GLSL scalar built-ins (GLSL 4.30 spec 8.3) ---------------------------- */
static inline float g_min(float x, float y) { return (y < x) ? y : x; }
static inline float g_max(float x, float y) { return (x < y) ? y : x; }
static inline float g_step(float edge, float x) { return (x < edge) ? 0.0f : 1.0f; }
static inline float g_sign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
static inline float g_abs(float x) { return vto_u2f(vto_f2u(x) & 0x7fffffffu); }
static inline float g_mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
static inline float g_mod(float x, float y) { return x - y * floorf(x / y); }
static inline float g_clamp(float x, float lo, float hi) { return g_min(g_max(x, lo), hi); }

/* float -> int, round toward zero, saturating, NaN -> 0 */
static inline int32_t g_f2i(float f)
{
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

/* ---- transcendental functions ---------------------------------------------- */

/* multiply by 2^k for k in [-300, 300] using two exact power-of-two factors */
static inline float vto_scale2(float p, int k)
{
    if (k > 300) k = 300;
    if (k < -300) k = -300;
    int k1 = k / 2;            /* C division truncates toward zero */
    int k2 = k - k1;
    float s1 = vto_u2f((uint32_t)(k1 + 127) << 23);   /* |k1|,|k2| <= 150 -> need clamping */
    float s2 = vto_u2f((uint32_t)(k2 + 127) << 23);
    return (p * s1) * s2;
}

/* exp2(t): k = floor(t + 0.5), r = t - k in [-0.5, 0.5], Cephes exp2f polynomial */
static inline float vto_exp2(float t)
{
    if (t != t) return t;
    if (t > 128.0f) return INFINITY;
    if (t < -150.0f) return 0.0f;
    float kf = floorf(t + 0.5f);
    float r = t - kf;
    int k = (int)kf;           /* in [-150, 129] */
    float p = 1.535336188319500e-4f;
    p = p * r + 1.339887440266574e-3f;
    p = p * r + 9.618437357674640e-3f;
    p = p * r + 5.550332471162809e-2f;
    p = p * r + 2.402264791363012e-1f;
    p = p * r + 6.931472028550421e-1f;
    p = p * r + 1.0f;
    /* k1,k2 in [-75, 65]: both scale factors are normal numbers */
    return vto_scale2(p, k);
}

/* log2(x), Cephes log2f: frexp to [sqrt(1/2), sqrt(2)), degree-9 polynomial */
static inline float vto_log2(float x)
{
    if (x != x) return x;
    if (x < 0.0f) return NAN;
    if (x == 0.0f) return -INFINITY;
    if (x == INFINITY) return x;
    uint32_t u = vto_f2u(x);
    int e = 0;
    if ((u & 0x7f800000u) == 0u) {           /* subnormal: scale by 2^24 (exact) */
        x = x * 16777216.0f;
        u = vto_f2u(x);
        e = -24;
    }
    e += (int)((u >> 23) & 0xffu) - 126;     /* x = m * 2^e, m in [0.5, 1) */
    float m = vto_u2f((u & 0x007fffffu) | 0x3f000000u);
    if (m < 0.70710678118654752440f) { e -= 1; m = (m + m) - 1.0f; }
    else { m = m - 1.0f; }
    float z = m * m;
    float y = 7.0376836292e-2f;
    y = y * m - 1.1514610310e-1f;
    y = y * m + 1.1676998740e-1f;
    y = y * m - 1.2420140846e-1f;
    y = y * m + 1.4249322787e-1f;
    y = y * m - 1.6668057665e-1f;
    y = y * m + 2.0000714765e-1f;
    y = y * m - 2.4999993993e-1f;
    y = y * m + 3.3333331174e-1f;
    y = (y * m) * z;
    y = y + (-0.5f * z);
    /* log2(1+m) = (m + y) * (1 + LOG2EA), LOG2EA = log2(e) - 1 */
    float r = y * 0.44269504088896340736f;
    r = r + m * 0.44269504088896340736f;
    r = r + y;
    r = r + m;
    r = r + (float)e;
    return r;
}

/* GLSL pow: "results are undefined if x < 0 or x = 0 and y <= 0"; defined as exp2(y*log2(x)) */
static inline float g_pow(float x, float y) { return vto_exp2(y * vto_log2(x)); }

/* shared Cody-Waite reduction by pi/2; returns quadrant, writes reduced argument */
static inline int vto_reduce_pio2(float x, float* r)
{
    float kf = floorf(x * 0.63661977236758134308f + 0.5f);
    float t = x - kf * 1.5703125f;                    /* exact for |k| < 2^15 */
    t = t - kf * 4.837512969970703125e-4f;
    t = t - kf * 7.54978995489188216e-8f;
    *r = t;
    return g_f2i(kf) & 3;
}
static inline float vto_sin_poly(float r)
{
    float z = r * r;
    float p = -1.9515295891e-4f;
    p = p * z + 8.3321608736e-3f;
    p = p * z - 1.6666654611e-1f;
    return (p * z) * r + r;
}
static inline float vto_cos_poly(float r)
{
    float z = r * r;
    float p = 2.443315711809948e-5f;
    p = p * z - 1.388731625493765e-3f;
    p = p * z + 4.166664568298827e-2f;
    return ((p * z) * z - 0.5f * z) + 1.0f;
}
static inline float g_sin(float x)
{
    if (!(g_abs(x) <= 3.0e4f)) return NAN;            /* out of contract range (and NaN/inf) */
    float r; int q = vto_reduce_pio2(x, &r);
    float s = (q & 1) ? vto_cos_poly(r) : vto_sin_poly(r);
    return (q & 2) ? -s : s;
}
static inline float g_cos(float x)
{
    if (!(g_abs(x) <= 3.0e4f)) return NAN;
    float r; int q = vto_reduce_pio2(x, &r);
    float c = (q & 1) ? vto_sin_poly(r) : vto_cos_poly(r);
    return ((q + 1) & 2) ? -c : c;
}

/* Cephes asinf kernel on [0, 0.5] */
static inline float vto_asin_kernel(float a)
{
    float z = a * a;
    float p = 4.2163199048e-2f;
    p = p * z + 2.4181311049e-2f;
    p = p * z + 4.5470025998e-2f;
    p = p * z + 7.4953002686e-2f;
    p = p * z + 1.6666752422e-1f;
    return (p * z) * a + a;
}
/* GLSL acos: undefined for |x| > 1 -> NaN */
static inline float g_acos(float x)
{
    if (!(g_abs(x) <= 1.0f)) return NAN;
    if (x < -0.5f) return VTO_PI - 2.0f * vto_asin_kernel(sqrtf(0.5f * (1.0f + x)));
    if (x > 0.5f)  return 2.0f * vto_asin_kernel(sqrtf(0.5f * (1.0f - x)));
    return 1.57079632679489661923f - ((x < 0.0f) ? -vto_asin_kernel(-x) : vto_asin_kernel(x));
}

/* Cephes atanf for a >= 0 */
static inline float vto_atan_pos(float a)
{
    float y0;
    if (a > 2.414213562373095f) { y0 = 1.57079632679489661923f; a = -(1.0f / a); }
    else if (a > 0.4142135623730950f) { y0 = 0.78539816339744830962f; a = (a - 1.0f) / (a + 1.0f); }
    else { y0 = 0.0f; }
    float z = a * a;
    float p = 8.05374449538e-2f;
    p = p * z - 1.38776856032e-1f;
    p = p * z + 1.99777106478e-1f;
    p = p * z - 3.33329491539e-1f;
    return y0 + ((p * z) * a + a);
}
/* GLSL atan(y, x) */
static inline float g_atan2(float y, float x)
{
    if (x != x || y != y) return NAN;
    if (x == 0.0f && y == 0.0f) return 0.0f;          /* GLSL: undefined; pinned to 0 */
    float ay = g_abs(y), ax = g_abs(x);
    float a;
    if (ax == INFINITY && ay == INFINITY) a = 0.78539816339744830962f;
    else a = vto_atan_pos(ay / ax);                   /* ax == 0 -> +inf -> pi/2 */
    if (x < 0.0f) a = VTO_PI - a;
    return (y < 0.0f) ? -a : a;
}

/* ---- vectors ---------------------------------------------------------------- */
static inline v3 V3(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 v3s(float s) { return V3(s, s, s); }
static inline v3 v3add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3mul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3div(v3 a, v3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 v3scale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 v3divs(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline v3 v3neg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline v3 v3abs(v3 a) { return V3(g_abs(a.x), g_abs(a.y), g_abs(a.z)); }
static inline v3 v3sign(v3 a) { return V3(g_sign(a.x), g_sign(a.y), g_sign(a.z)); }
static inline v3 v3floor(v3 a) { return V3(floorf(a.x), floorf(a.y), floorf(a.z)); }
/* dot: left-to-right sum of products, no fusion */
static inline float v3dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float v3length(v3 a) { return sqrtf(v3dot(a, a)); }
/* GLSL normalize(x) = x / length(x) */
static inline v3 v3normalize(v3 a) { return v3divs(a, v3length(a)); }
/* GLSL cross (spec 8.5) */
static inline v3 v3cross(v3 a, v3 b)
{
    return V3(a.y * b.z - b.y * a.z,
              a.z * b.x - b.z * a.x,
              a.x * b.y - b.x * a.y);
}
/* row-major 4x4 (the host's pm.x[r][c]) times column vector; rows summed left to right */
static inline v4 m4mulv(const float* m, float x, float y, float z, float w)
{
    v4 r;
    r.x = ((m[0] * x + m[1] * y) + m[2] * z) + m[3] * w;
    r.y = ((m[4] * x + m[5] * y) + m[6] * z) + m[7] * w;
    r.z = ((m[8] * x + m[9] * y) + m[10] * z) + m[11] * w;
    r.w = ((m[12] * x + m[13] * y) + m[14] * z) + m[15] * w;
    return r;
}

#endif /* VTO_MATH_H */
