// Environment-map ingest on the device: the host half of P9, i.e. everything
// Renderer::loadBackgroundImage (renderer.cpp:947-1055) does between reading the file and uploading the three
// textures: RGB -> RGBA32F, importance function (image.cpp:285-346: downscale to <= 512, luminance, 3x3 gaussian),
// sin(theta) weighting + integral (image.cpp:349-389), conditional / marginal CDFs (image.cpp:68-283), and this
// library's guide tables for the CDF searches. Every float operation keeps the reference's order: the prefix sums
// and the integral are SEQUENTIAL chains of rounded additions, so they stay sequential here (one warp per row, one
// thread for the 2-D sum); what runs in parallel is everything around the chains (rows, divisions, filters).
// Compiled with -fmad=false like the rest of the library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef VT_GLOBAL
#define VT_GLOBAL static __global__
#endif
#define VT_ENV_MAX_CDF_SIZE 512          // image.cpp:14

VT_GLOBAL void vt_env_rgba_kernel(const float* __restrict__ rgb, float4* __restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;        // renderer.cpp:994-1002: GL_RGB -> GL_RGBA32F, alpha 1
    if (i < n) out[i] = make_float4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 1.0f);
}

// image.cpp:309-335: reduction to <= 512 texels per side + Rec.709 luminance. The reduction is this build's area-weighted box
// filter (host/image.cpp generateImageFunction: source rectangle [x sx, (x+1) sx) x [y sy, (y+1) sy), edge pixels weighted by
// their overlap, double accumulator, rows outer); same operations in the same order as the host, so the same bits. With
// integer factors every weight is exactly 1.
VT_GLOBAL void vt_env_luminance_kernel(const float* __restrict__ rgb, int w, int h, int nw, int nh, float* __restrict__ lum)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nw || y >= nh) return;
    float ch[3];
    if (nw == w && nh == h) {
        for (int c = 0; c < 3; ++c) ch[c] = rgb[((size_t)y * w + x) * 3 + c];
    } else {
        const double sx = (double)w / (double)nw, sy = (double)h / (double)nh;
        const double y0 = (double)y * sy, y1 = (double)(y + 1) * sy, x0 = (double)x * sx, x1 = (double)(x + 1) * sx;
        const int j0 = (int)floor(y0), j1 = min(h, (int)ceil(y1)), i0 = (int)floor(x0), i1 = min(w, (int)ceil(x1));
        const double area = sx * sy;
        for (int c = 0; c < 3; ++c) {
            double acc = 0.0;
            for (int j = j0; j < j1; ++j) {
                const double wy = fmin(y1, (double)(j + 1)) - fmax(y0, (double)j);
                for (int i = i0; i < i1; ++i) {
                    const double wx = fmin(x1, (double)(i + 1)) - fmax(x0, (double)i);
                    acc += (wy * wx) * (double)rgb[((size_t)j * w + i) * 3 + c];
                }
            }
            ch[c] = (float)(acc / area);
        }
    }
    lum[(size_t)y * nw + x] = (ch[0] * 0.2126f + ch[1] * 0.7152f) + ch[2] * 0.0722f;
}

// image.cpp:337-342: one axis of the separable 3x3 gaussian (1/4, 1/2, 1/4), edges clamped
VT_GLOBAL void vt_env_blur_kernel(const float* __restrict__ in, float* __restrict__ out, int w, int h, int vertical)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    float a, b, c;
    if (vertical) {
        a = in[(size_t)(y ? y - 1 : 0) * w + x]; b = in[(size_t)y * w + x]; c = in[(size_t)(y + 1 < h ? y + 1 : h - 1) * w + x];
    } else {
        a = in[(size_t)y * w + (x ? x - 1 : 0)]; b = in[(size_t)y * w + x]; c = in[(size_t)y * w + (x + 1 < w ? x + 1 : w - 1)];
    }
    out[(size_t)y * w + x] = (a * 0.25f + b * 0.5f) + c * 0.25f;
}

// image.cpp:361-375: functionU = max(0, value) * sinTheta(row); the sine table comes from the host (double-precision
// libm sin of M_PI * (y + 0.5f) / H, rounded to float: the device's double sin is not bit-compatible with glibc's)
VT_GLOBAL void vt_env_function_kernel(const float* __restrict__ img, const float* __restrict__ sin_row, int w, int h, float* __restrict__ fu)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const float v = img[(size_t)y * w + x];
    fu[(size_t)y * w + x] = ((0.f < v) ? v : 0.f) * sin_row[y];
}

// image.cpp:372: textureTimesSinSum += value * sinTheta, row-major, one chain of rounded float additions.
// One CTA: all threads stage a tile in shared memory, thread 0 folds it in order.
VT_GLOBAL void __launch_bounds__(256) vt_env_sum_kernel(const float* __restrict__ fu, size_t n, float* __restrict__ out)
{
    constexpr int kTile = 4096;
    __shared__ float tile[kTile];
    float sum = 0.0f;
    for (size_t base = 0; base < n; base += kTile) {
        const int m = (int)((n - base < (size_t)kTile) ? n - base : (size_t)kTile);
        for (int i = threadIdx.x; i < m; i += blockDim.x) tile[i] = fu[base + i];
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll 8
            for (int i = 0; i < m; ++i) sum = sum + tile[i];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sum;
}

// One CDF (image.cpp:212-247 for a row of CDF-U, :251-280 for CDF-V) by one warp: out[0] = 0,
// out[x] = out[x-1] + f[x-1] / steps (sequential), then out[x] /= out[n] or, for an all-zero function, x / steps.
// The divisions run lane-parallel; the chain is replayed by every lane from shuffled operands, lane i keeps step i.
__device__ __forceinline__ float env_scan_warp(const float* __restrict__ f, int n, float steps, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    float acc = 0.0f;
    if (lane == 0) out[0] = 0.0f;
    for (int x0 = 0; x0 < n; x0 += 32) {
        const float v = (x0 + lane < n) ? f[x0 + lane] / steps : 0.0f;
        float mine = 0.0f;
        const int m = (n - x0 < 32) ? n - x0 : 32;
        for (int i = 0; i < m; ++i) {
            acc = acc + __shfl_sync(0xffffffffu, v, i);
            if (lane == i) mine = acc;
        }
        if (x0 + lane < n) out[x0 + lane + 1] = mine;
    }
    const float total = acc;                                               // out[n], known to every lane
    for (int x = 1 + lane; x <= n; x += 32)                                 // each lane re-reads what it wrote itself
        out[x] = (total > 0.0f) ? out[x] / total : (float)x / steps;
    return total;
}

VT_GLOBAL void __launch_bounds__(128) vt_env_cdf_rows_kernel(const float* __restrict__ fu, int w, int h, float* __restrict__ cdf_u, float* __restrict__ fv)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= h) return;
    const float total = env_scan_warp(fu + (size_t)row * w, w, (float)(unsigned int)w, cdf_u + (size_t)row * (w + 1));
    if ((threadIdx.x & 31) == 0) fv[row] = total;
}

VT_GLOBAL void __launch_bounds__(32) vt_env_cdf_v_kernel(const float* __restrict__ fv, int h, float* __restrict__ cdf_v)
{
    env_scan_warp(fv, h, (float)(unsigned int)h, cdf_v);
}

// Guide tables (vt_api.cu build_guide / vt_device.cuh cdf_search_guided): for CDF row r (n entries, `stride` apart)
// guide[j] = max({0} U {m in [1, n-2] : cdf[m] <= j/K}), guide[K] = n-2; *sorted is cleared when some row's
// cdf[1..n-2] is not non-decreasing (then the device keeps the literal bisection of envMapSample.h:70-123).
VT_GLOBAL void vt_env_guide_kernel(const float* __restrict__ cdf, int n, int stride, int rows, int K, unsigned short* __restrict__ guide, int* __restrict__ sorted)
{
    const int r = blockIdx.x;
    if (r >= rows) return;
    const float* c = cdf + (size_t)r * stride;
    bool ok = true;
    for (int m = 2 + threadIdx.x; m <= n - 2; m += blockDim.x) ok = ok && (c[m] >= c[m - 1]);
    if (threadIdx.x == 0 && n - 2 >= 1) ok = ok && (c[1] == c[1]);
    if (!ok) *sorted = 0;
    for (int j = threadIdx.x; j <= K; j += blockDim.x) {
        int g;
        if (j == K) g = (n - 2 > 0) ? n - 2 : 0;
        else {
            const float t = (float)j / (float)K;
            int lo = 0, hi = n - 2;                                         // largest m in [0, n-2] with m == 0 or cdf[m] <= t
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (c[mid] <= t) lo = mid; else hi = mid - 1; }
            g = lo;
        }
        guide[(size_t)r * (K + 1) + j] = (unsigned short)g;
    }
}
