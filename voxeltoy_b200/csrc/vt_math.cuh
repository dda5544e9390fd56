// vt_math.cuh -- device arithmetic of the voxelToy hot path (sm_100a).
//
// The path tracer's discrete decisions (which voxel a DDA hits, which axis a hit
// normal takes, whether a shadow ray is blocked) depend on the exact binary32 value
// of every intermediate, so this file fixes the arithmetic contract of the product:
//   * every + - * / sqrt is one correctly rounded binary32 operation, never fused
//     (the translation units are compiled with -fmad=false, IEEE div/sqrt, no FTZ);
//   * GLSL built-ins follow the GLSL 4.30 specification text (section 8.3):
//       min(x,y) = y < x ? y : x     max(x,y) = x < y ? y : x
//       step(e,x) = x < e ? 0 : 1    sign(0) = 0     mix(x,y,a) = x*(1-a) + y*a
//       mod(x,y) = x - y*floor(x/y)  pow(x,y) = exp2(y*log2(x))
//   * sin/cos/acos/atan/exp2/log2 are fixed polynomial kernels (Cody-Waite reduction
//     + minimax / Cephes single-precision coefficients), a few ulp from libm, so the
//     result does not depend on a driver's or libdevice's choice of approximation;
//   * float -> int is cvt.rzi.s32.f32 (saturating, NaN -> 0).
// Vector helpers evaluate left to right exactly as the GLSL expressions of the
// reference shaders are written (src/shaders/**, cited where used).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VT_DEV __device__ __forceinline__
// kernels have internal linkage: the headers are compiled into more than one translation unit (vt_api.cu, vt_shade.cu)
#define VT_GLOBAL static __global__
// Division and pow are expanded at ~90 and ~20 sites of wf_shade (1 400 + 1 250 of its 6 700 instructions); that kernel
// stalls on instruction fetch (14 % of its samples), so its translation unit (vt_shade.cu, VT_COMPACT_MATH) calls one
// copy of each instead: shade -8 %. Everywhere else they stay inline (wf_generate is 6 % slower with calls).
#ifdef VT_COMPACT_MATH
#define VT_DEV_HEAVY static __device__ __noinline__
#else
#define VT_DEV_HEAVY __device__ __forceinline__
#endif

namespace vt {

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

VT_DEV f2 mk2(float x, float y) { f2 r; r.x = x; r.y = y; return r; }
VT_DEV f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
VT_DEV f3 mk3(float s) { return mk3(s, s, s); }
VT_DEV f4 mk4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
VT_DEV f4 mk4(f3 v, float w) { return mk4(v.x, v.y, v.z, w); }
VT_DEV f3 xyz(f4 v) { return mk3(v.x, v.y, v.z); }

// IEEE binary32 division with the zero numerator peeled off. div.rn's inline fast path rejects a == 0 (FCHK) and calls a
// ~60-instruction subroutine; axis-aligned normals, black albedo channels and zero throughput make that the common case
// in shading (10 % of wf_shade's instructions before this). 0 / b = (sign a ^ sign b) 0 for every b except 0 and NaN.
VT_DEV_HEAVY float gdiv(float a, float b)
{
    if (a == 0.0f && b == b && b != 0.0f) return __int_as_float((__float_as_int(a) ^ __float_as_int(b)) & (int)0x80000000);
    return a / b;
}

// Division by a COMPILE-TIME constant c (the shaders divide by PI and 2 PI about ten times per shaded hit): with rc = RN(1 / c),
//   q0 = RN(a * rc),  r = a - c * q0 (exact, one FMA),  q = RN(q0 + r * rc)
// is the correctly rounded quotient (Markstein; q0 is within an ulp of a / c and the residual is exact) -- 3 instructions instead of
// the ~15 of the called IEEE sequence. The argument needs normal operands and results: zeros (sign!), infinities, NaN and anything
// below 1e-30 take gdiv. Not taken on trust: vt_debug_div_const compares it with div.rn for ALL 2^32 numerators on the device
// (tests/test_gpu_edges.py::test_division_by_constants_is_exact), for every constant used below.
VT_DEV float gdiv_by(float a, const float c, const float rc)
{
    const float aa = __int_as_float(__float_as_int(a) & 0x7fffffff);
    if (!(aa >= 1.0e-30f && aa < __int_as_float(0x7f800000))) return gdiv(a, c);
    const float q0 = a * rc;
    const float r = fmaf(-c, q0, a);          // an explicit fma: -fmad=false only forbids CONTRACTING a * b + c
    return fmaf(r, rc, q0);
}
#define VT_PI_F        3.14159265359f
#define VT_TWO_PI_F    6.28318530718f
VT_DEV float gdiv_pi(float a) { return gdiv_by(a, VT_PI_F, 1.0f / VT_PI_F); }
VT_DEV float gdiv_two_pi(float a) { return gdiv_by(a, VT_TWO_PI_F, 1.0f / VT_TWO_PI_F); }

VT_DEV f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
VT_DEV f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
VT_DEV f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
VT_DEV f3 operator/(f3 a, f3 b) { return mk3(gdiv(a.x, b.x), gdiv(a.y, b.y), gdiv(a.z, b.z)); }
VT_DEV f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
VT_DEV f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
VT_DEV f3 operator/(f3 a, float s) { return mk3(gdiv(a.x, s), gdiv(a.y, s), gdiv(a.z, s)); }
VT_DEV f3 operator+(f3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
VT_DEV f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }

// ---- GLSL scalar built-ins ----------------------------------------------------------
VT_DEV float gmin(float x, float y) { return (y < x) ? y : x; }
VT_DEV float gmax(float x, float y) { return (x < y) ? y : x; }
VT_DEV float gstep(float edge, float x) { return (x < edge) ? 0.0f : 1.0f; }
VT_DEV float gsign(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
VT_DEV float gabs(float x) { return __int_as_float(__float_as_int(x) & 0x7fffffff); }
VT_DEV float gmix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
VT_DEV float gmod(float x, float y) { return x - y * floorf(x / y); }
VT_DEV float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
VT_DEV int f2i(float f) { return __float2int_rz(f); }
// dda.h:29  mix(d, 1e-5, step(abs(d), 1e-5)) = d * (1 - a) + 1e-5 * a with a = (1e-5 < |d|) ? 0 : 1. For |d| > 1e-5 that is d * 1 + 0 = d
// (d is not zero there, so adding +0 changes nothing), for |d| <= 1e-5 it is +-0 + 1e-5 = 1e-5, and a NaN stays a NaN: one compare
// and one select instead of seven instructions per component. Same bits for every binary32 d (vt_debug_div_const, which = 2).
VT_DEV float gclamp_dir(float d) { return !(gabs(d) <= 1e-5f) ? d : 1e-5f; }

VT_DEV f3 gabs(f3 a) { return mk3(gabs(a.x), gabs(a.y), gabs(a.z)); }
VT_DEV f3 gsign(f3 a) { return mk3(gsign(a.x), gsign(a.y), gsign(a.z)); }
VT_DEV f3 gfloor(f3 a) { return mk3(floorf(a.x), floorf(a.y), floorf(a.z)); }
VT_DEV f3 gmin(f3 a, f3 b) { return mk3(gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)); }
VT_DEV f3 gmax(f3 a, f3 b) { return mk3(gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)); }

VT_DEV float dot(f2 a, f2 b) { return a.x * b.x + a.y * b.y; }
VT_DEV float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
VT_DEV float length(f3 a) { return sqrtf(dot(a, a)); }
VT_DEV f3 normalize(f3 a) { return a / length(a); }
VT_DEV f3 cross(f3 a, f3 b)
{
    return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
// row-major 4x4 (host pm.x[r][c], uploaded with transpose = GL_TRUE) times column vector
VT_DEV f4 mul44(const float* __restrict__ m, float x, float y, float z, float w)
{
    f4 r;
    r.x = ((m[0] * x + m[1] * y) + m[2] * z) + m[3] * w;
    r.y = ((m[4] * x + m[5] * y) + m[6] * z) + m[7] * w;
    r.z = ((m[8] * x + m[9] * y) + m[10] * z) + m[11] * w;
    r.w = ((m[12] * x + m[13] * y) + m[14] * z) + m[15] * w;
    return r;
}

// ---- constants: shaders/shared/constants.h:1-3 ----------------------------------------
#define VT_PI        3.14159265359f
#define VT_TWO_PI    6.28318530718f
#define VT_INV_TWOPI 0.15915494309f

// ---- transcendental kernels -----------------------------------------------------------
VT_DEV float pow2i(int k) { return __int_as_float((k + 127) << 23); }

// exp2: round-to-nearest split + degree-6 polynomial on [-0.5, 0.5]
VT_DEV float gexp2(float t)
{
    if (t != t) return t;
    if (t > 128.0f) return __int_as_float(0x7f800000);
    if (t < -150.0f) return 0.0f;
    const float kf = floorf(t + 0.5f);
    const float r = t - kf;
    int k = (int)kf;
    float p = 1.535336188319500e-4f;
    p = p * r + 1.339887440266574e-3f;
    p = p * r + 9.618437357674640e-3f;
    p = p * r + 5.550332471162809e-2f;
    p = p * r + 2.402264791363012e-1f;
    p = p * r + 6.931472028550421e-1f;
    p = p * r + 1.0f;
    k = k > 300 ? 300 : (k < -300 ? -300 : k);
    const int k1 = k / 2, k2 = k - k1;     // two normal power-of-two factors: result rounds once, in the last multiply
    return (p * pow2i(k1)) * pow2i(k2);
}

// log2: mantissa in [sqrt(1/2), sqrt(2)), degree-9 polynomial (Cephes log2f)
VT_DEV float glog2(float x)
{
    if (x != x) return x;
    if (x < 0.0f) return __int_as_float(0x7fc00000);
    if (x == 0.0f) return __int_as_float(0xff800000);
    if (x == __int_as_float(0x7f800000)) return x;
    uint32_t u = (uint32_t)__float_as_int(x);
    int e = 0;
    if ((u & 0x7f800000u) == 0u) { x = x * 16777216.0f; u = (uint32_t)__float_as_int(x); e = -24; }
    e += (int)((u >> 23) & 0xffu) - 126;
    float m = __int_as_float((int)((u & 0x007fffffu) | 0x3f000000u));
    if (m < 0.70710678118654752440f) { e -= 1; m = (m + m) - 1.0f; }
    else { m = m - 1.0f; }
    const float z = m * m;
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
    float r = y * 0.44269504088896340736f;
    r = r + m * 0.44269504088896340736f;
    r = r + y;
    r = r + m;
    r = r + (float)e;
    return r;
}

VT_DEV_HEAVY float gpow(float x, float y) { return gexp2(y * glog2(x)); }

VT_DEV int reduce_pio2(float x, float& r)
{
    const float kf = floorf(x * 0.63661977236758134308f + 0.5f);
    float t = x - kf * 1.5703125f;
    t = t - kf * 4.837512969970703125e-4f;
    t = t - kf * 7.54978995489188216e-8f;
    r = t;
    return f2i(kf) & 3;
}
VT_DEV float sin_poly(float r)
{
    const float z = r * r;
    float p = -1.9515295891e-4f;
    p = p * z + 8.3321608736e-3f;
    p = p * z - 1.6666654611e-1f;
    return (p * z) * r + r;
}
VT_DEV float cos_poly(float r)
{
    const float z = r * r;
    float p = 2.443315711809948e-5f;
    p = p * z - 1.388731625493765e-3f;
    p = p * z + 4.166664568298827e-2f;
    return ((p * z) * z - 0.5f * z) + 1.0f;
}
VT_DEV float gsin(float x)
{
    if (!(gabs(x) <= 3.0e4f)) return __int_as_float(0x7fc00000);
    float r; const int q = reduce_pio2(x, r);
    const float s = (q & 1) ? cos_poly(r) : sin_poly(r);
    return (q & 2) ? -s : s;
}
VT_DEV float gcos(float x)
{
    if (!(gabs(x) <= 3.0e4f)) return __int_as_float(0x7fc00000);
    float r; const int q = reduce_pio2(x, r);
    const float c = (q & 1) ? sin_poly(r) : cos_poly(r);
    return ((q + 1) & 2) ? -c : c;
}
// sin and cos of the same argument share one reduction
VT_DEV void gsincos(float x, float& s, float& c)
{
    if (!(gabs(x) <= 3.0e4f)) { s = c = __int_as_float(0x7fc00000); return; }
    float r; const int q = reduce_pio2(x, r);
    const float sp = sin_poly(r), cp = cos_poly(r);
    const float sv = (q & 1) ? cp : sp;
    const float cv = (q & 1) ? sp : cp;
    s = (q & 2) ? -sv : sv;
    c = ((q + 1) & 2) ? -cv : cv;
}

VT_DEV float asin_kernel(float a)
{
    const float z = a * a;
    float p = 4.2163199048e-2f;
    p = p * z + 2.4181311049e-2f;
    p = p * z + 4.5470025998e-2f;
    p = p * z + 7.4953002686e-2f;
    p = p * z + 1.6666752422e-1f;
    return (p * z) * a + a;
}
VT_DEV float gacos(float x)
{
    if (!(gabs(x) <= 1.0f)) return __int_as_float(0x7fc00000);
    if (x < -0.5f) return VT_PI - 2.0f * asin_kernel(sqrtf(0.5f * (1.0f + x)));
    if (x > 0.5f) return 2.0f * asin_kernel(sqrtf(0.5f * (1.0f - x)));
    return 1.57079632679489661923f - ((x < 0.0f) ? -asin_kernel(-x) : asin_kernel(x));
}
VT_DEV float atan_pos(float a)
{
    float y0;
    if (a > 2.414213562373095f) { y0 = 1.57079632679489661923f; a = -(1.0f / a); }
    else if (a > 0.4142135623730950f) { y0 = 0.78539816339744830962f; a = (a - 1.0f) / (a + 1.0f); }
    else { y0 = 0.0f; }
    const float z = a * a;
    float p = 8.05374449538e-2f;
    p = p * z - 1.38776856032e-1f;
    p = p * z + 1.99777106478e-1f;
    p = p * z - 3.33329491539e-1f;
    return y0 + ((p * z) * a + a);
}
VT_DEV float gatan2(float y, float x)
{
    if (x != x || y != y) return __int_as_float(0x7fc00000);
    if (x == 0.0f && y == 0.0f) return 0.0f;
    const float ay = gabs(y), ax = gabs(x);
    const float inf = __int_as_float(0x7f800000);
    float a;
    if (ax == inf && ay == inf) a = 0.78539816339744830962f;
    else a = atan_pos(ay / ax);
    if (x < 0.0f) a = VT_PI - a;
    return (y < 0.0f) ? -a : a;
}

} // namespace vt
