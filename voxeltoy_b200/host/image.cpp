// image.cpp -- environment-map ingest and the marginal / conditional CDF construction used to importance-sample
// it (PBRT-2 14.6.5). calculateCDF follows renderer/image.cpp:68-283 and :349-389 of the reference. The reference
// delegates file I/O, down-scaling, luminance and the 3x3 blur to OpenImageIO (un-vendored, un-pinned: SURVEY 8c);
// this build owns that chain: PFM / Radiance-HDR readers here, OpenEXR / PNG in image_formats.cpp, area-weighted down-scale to <= 512, Rec.709 luminance, separable
// [1 2 1]/4 blur with clamped edges.
#include "vt_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <zlib.h>

#define MAX_CDF_SIZE 512          // image.cpp:14

// ---- file I/O --------------------------------------------------------------------------------------------------------
static bool readPFM(FILE* fp, unsigned int& w, unsigned int& h, std::vector<float>& rgb)
{
    char tag[3] = { 0, 0, 0 };
    int iw = 0, ih = 0; float scale = 0;
    if (fscanf(fp, "%2s %d %d %f", tag, &iw, &ih, &scale) != 4 || iw <= 0 || ih <= 0) return false;
    fgetc(fp);                                       // single whitespace after the header
    const int ch = (strcmp(tag, "PF") == 0) ? 3 : ((strcmp(tag, "Pf") == 0) ? 1 : 0);
    if (!ch) return false;
    std::vector<float> raw((size_t)iw * ih * ch);
    if (fread(&raw[0], sizeof(float), raw.size(), fp) != raw.size()) return false;
    if (scale > 0) {                                 // big-endian payload
        for (size_t i = 0; i < raw.size(); ++i) { unsigned char* b = (unsigned char*)&raw[i]; std::swap(b[0], b[3]); std::swap(b[1], b[2]); }
    }
    w = iw; h = ih; rgb.resize((size_t)iw * ih * 3);
    for (int y = 0; y < ih; ++y)                     // PFM rows are stored bottom-to-top; image row 0 is the top row
        for (int x = 0; x < iw; ++x)
            for (int c = 0; c < 3; ++c)
                rgb[((size_t)y * iw + x) * 3 + c] = raw[((size_t)(ih - 1 - y) * iw + x) * ch + (ch == 3 ? c : 0)];
    return true;
}

static bool readHDR(FILE* fp, unsigned int& w, unsigned int& h, std::vector<float>& rgb)
{
    char line[256];
    bool fmt = false;
    if (!fgets(line, sizeof line, fp) || strncmp(line, "#?", 2) != 0) return false;
    while (fgets(line, sizeof line, fp)) {
        if (line[0] == '\n' || line[0] == '\r') break;
        if (strncmp(line, "FORMAT=32-bit_rle_rgbe", 22) == 0) fmt = true;
    }
    if (!fmt || !fgets(line, sizeof line, fp)) return false;
    int ih = 0, iw = 0;
    if (sscanf(line, "-Y %d +X %d", &ih, &iw) != 2 || iw <= 0 || ih <= 0) return false;
    w = iw; h = ih; rgb.resize((size_t)iw * ih * 3);
    std::vector<unsigned char> scan((size_t)iw * 4);
    for (int y = 0; y < ih; ++y) {
        unsigned char hd[4];
        if (fread(hd, 1, 4, fp) != 4) return false;
        if (hd[0] == 2 && hd[1] == 2 && !(hd[2] & 0x80) && ((hd[2] << 8) | hd[3]) == iw) {      // adaptive RLE scanline
            for (int c = 0; c < 4; ++c) {
                int x = 0;
                while (x < iw) {
                    int n = fgetc(fp);
                    if (n == EOF) return false;
                    if (n > 128) { n -= 128; const int v = fgetc(fp); if (x + n > iw) return false; while (n--) scan[(size_t)x++ * 4 + c] = (unsigned char)v; }
                    else { if (x + n > iw) return false; while (n--) scan[(size_t)x++ * 4 + c] = (unsigned char)fgetc(fp); }
                }
            }
        } else {                                                                                 // flat scanline
            memcpy(&scan[0], hd, 4);
            if (iw > 1 && fread(&scan[4], 4, iw - 1, fp) != (size_t)(iw - 1)) return false;
        }
        for (int x = 0; x < iw; ++x) {
            const unsigned char* p = &scan[(size_t)x * 4];
            const float f = p[3] ? std::ldexp(1.0f, (int)p[3] - (128 + 8)) : 0.0f;
            for (int c = 0; c < 3; ++c) rgb[((size_t)y * iw + x) * 3 + c] = p[c] * f;
        }
    }
    return true;
}

// image.cpp:28-59: float RGB, row 0 first
bool loadImage(const std::string& path, unsigned int& outWidth, unsigned int& outHeight, std::vector<float>& outPixelData)
{
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    const int c0 = fgetc(fp), c1 = fgetc(fp);
    rewind(fp);
    bool ok = false;
    if (c0 == 'P' && (c1 == 'F' || c1 == 'f')) ok = readPFM(fp, outWidth, outHeight, outPixelData);
    else if (c0 == '#' && c1 == '?') ok = readHDR(fp, outWidth, outHeight, outPixelData);
    fclose(fp);
    if (c0 == 0x76 && c1 == 0x2f) return readEXR(path, outWidth, outHeight, outPixelData);       // OpenEXR magic 76 2f 31 01
    if (c0 == 0x89 && c1 == 'P') return readPNG(path, outWidth, outHeight, outPixelData);
    return ok;
}

bool writePFM(const std::string& path, const float* rgb, unsigned int w, unsigned int h)
{
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    fprintf(fp, "PF\n%u %u\n-1.0\n", w, h);
    for (unsigned int y = 0; y < h; ++y) fwrite(rgb + (size_t)(h - 1 - y) * w * 3, sizeof(float), (size_t)w * 3, fp);
    fclose(fp);
    return true;
}

// ---- PNG (RGBA8, rows top-down) ------------------------------------------------------------------------------------------
// The reference writes whatever format OpenImageIO derives from the file extension (renderer.cpp:1129-1138); PNG is what its
// UI offers. A PNG is a zlib stream inside IDAT; this writer emits *stored* deflate blocks (no compression, no dependency):
// every decoder accepts them. CRC-32 (ISO 3309) over chunk type + data, Adler-32 over the raw scanlines.
static uint32_t crc32_update(uint32_t crc, const unsigned char* p, size_t n)
{
    static uint32_t table[256]; static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return crc;
}

static void put_be32(std::vector<unsigned char>& v, uint32_t x) { v.push_back(x >> 24); v.push_back((x >> 16) & 0xFF); v.push_back((x >> 8) & 0xFF); v.push_back(x & 0xFF); }

static bool write_chunk(FILE* fp, const char type[4], const std::vector<unsigned char>& data)
{
    std::vector<unsigned char> head; put_be32(head, (uint32_t)data.size());
    uint32_t crc = crc32_update(0xFFFFFFFFu, (const unsigned char*)type, 4);
    if (!data.empty()) crc = crc32_update(crc, &data[0], data.size());
    std::vector<unsigned char> tail; put_be32(tail, crc ^ 0xFFFFFFFFu);
    return fwrite(&head[0], 1, 4, fp) == 4 && fwrite(type, 1, 4, fp) == 4 && (data.empty() || fwrite(&data[0], 1, data.size(), fp) == data.size())
        && fwrite(&tail[0], 1, 4, fp) == 4;
}

// Radiance .hdr (RGBE, flat scanlines -- every reader accepts them), rows top-down; `stride` floats per pixel (3 or 4), RGB first
bool writeHDR(const std::string& path, const float* pixels, unsigned int w, unsigned int h, int stride)
{
    if (!pixels || w == 0 || h == 0 || stride < 3) return false;
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    fprintf(fp, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %u +X %u\n", h, w);
    std::vector<unsigned char> line((size_t)w * 4);
    for (unsigned int y = 0; y < h; ++y) {
        for (unsigned int x = 0; x < w; ++x) {
            const float* p = pixels + ((size_t)y * w + x) * stride;
            float r = p[0], g = p[1], b = p[2];
            if (!(r > 0.0f)) r = 0.0f; if (!(g > 0.0f)) g = 0.0f; if (!(b > 0.0f)) b = 0.0f;         // RGBE holds no negatives and no NaN
            const float m = std::max(r, std::max(g, b));
            unsigned char* o = &line[(size_t)x * 4];
            if (m < 1e-32f || !(m < 3.0e38f)) { o[0] = o[1] = o[2] = o[3] = 0; if (!(m < 3.0e38f)) { o[0] = o[1] = o[2] = 255; o[3] = 255; } continue; }
            int e; const float f = std::frexp(m, &e) * 256.0f / m;
            o[0] = (unsigned char)(r * f); o[1] = (unsigned char)(g * f); o[2] = (unsigned char)(b * f); o[3] = (unsigned char)(e + 128);
        }
        fwrite(&line[0], 1, line.size(), fp);
    }
    const bool ok = ferror(fp) == 0;
    return (fclose(fp) == 0) && ok;
}

bool writePNG(const std::string& path, const unsigned char* rgba8, unsigned int w, unsigned int h)
{
    if (!rgba8 || w == 0 || h == 0) return false;
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    bool ok = fwrite(sig, 1, 8, fp) == 8;
    std::vector<unsigned char> ihdr; put_be32(ihdr, w); put_be32(ihdr, h);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8 bit, RGBA, deflate, filter 0, no interlace
    ok = ok && write_chunk(fp, "IHDR", ihdr);
    const size_t stride = (size_t)w * 4 + 1, rawSize = stride * h;                                   // filter byte 0 (None) per scanline
    std::vector<unsigned char> raw(rawSize);
    for (unsigned int y = 0; y < h; ++y) { raw[y * stride] = 0; memcpy(&raw[y * stride + 1], rgba8 + (size_t)y * w * 4, (size_t)w * 4); }
    uLongf zn = compressBound((uLong)rawSize);                                                       // zlib stream: deflate level 6
    std::vector<unsigned char> z(zn);
    ok = ok && compress2(&z[0], &zn, &raw[0], (uLong)rawSize, 6) == Z_OK;
    z.resize(ok ? zn : 0);
    ok = ok && write_chunk(fp, "IDAT", z) && write_chunk(fp, "IEND", std::vector<unsigned char>());
    return (fclose(fp) == 0) && ok;
}

// ---- image function: image.cpp:285-346 (generateImageFunction) -------------------------------------------------------
static bool generateImageFunction(const float* rgbPixels, unsigned int imageWidth, unsigned int imageHeight,
                                  std::vector<float>& result, unsigned int& outW, unsigned int& outH)
{
    std::vector<float> src(rgbPixels, rgbPixels + (size_t)imageWidth * imageHeight * 3);
    unsigned int w = imageWidth, h = imageHeight;
    const unsigned int m = std::max(w, h);
    if (m > MAX_CDF_SIZE) {                                                    // :309-321
        // The reference resizes with OpenImageIO's default filter (ImageBufAlgo::resize), which no test pins and which is not
        // available here. This build defines the reduction as an AREA-WEIGHTED BOX filter: output pixel (x, y)
        // averages the source rectangle [x sx, (x+1) sx) x [y sy, (y+1) sy), sx = w / nw, sy = h / nh, edge pixels weighted by
        // their overlap; double accumulator, rows outer. Integer factors give weights of exactly 1 (the plain box mean).
        const unsigned int nw = (unsigned int)((float)w / m * MAX_CDF_SIZE);
        const unsigned int nh = (unsigned int)((float)h / m * MAX_CDF_SIZE);
        if (nw == 0 || nh == 0) return false;
        const double sx = (double)w / (double)nw, sy = (double)h / (double)nh;
        std::vector<float> small((size_t)nw * nh * 3);
        for (unsigned int y = 0; y < nh; ++y) {
            const double y0 = (double)y * sy, y1 = (double)(y + 1) * sy;
            const unsigned int j0 = (unsigned int)std::floor(y0), j1 = std::min(h, (unsigned int)std::ceil(y1));
            for (unsigned int x = 0; x < nw; ++x) {
                const double x0 = (double)x * sx, x1 = (double)(x + 1) * sx;
                const unsigned int i0 = (unsigned int)std::floor(x0), i1 = std::min(w, (unsigned int)std::ceil(x1));
                for (int c = 0; c < 3; ++c) {
                    double acc = 0;
                    for (unsigned int j = j0; j < j1; ++j) {
                        const double wy = std::min(y1, (double)(j + 1)) - std::max(y0, (double)j);
                        for (unsigned int i = i0; i < i1; ++i) {
                            const double wx = std::min(x1, (double)(i + 1)) - std::max(x0, (double)i);
                            acc += (wy * wx) * (double)src[((size_t)j * w + i) * 3 + c];
                        }
                    }
                    small[((size_t)y * nw + x) * 3 + c] = (float)(acc / (sx * sy));
                }
            }
        }
        src.swap(small); w = nw; h = nh;
    }
    std::vector<float> lum((size_t)w * h), tmp((size_t)w * h);
    for (size_t i = 0; i < lum.size(); ++i) {                                  // :326-335, Rec.709 weights
        volatile float a = src[3 * i] * 0.2126f, b = src[3 * i + 1] * 0.7152f, c = src[3 * i + 2] * 0.0722f;
        volatile float ab = a + b;
        lum[i] = ab + c;
    }
    const float k0 = 0.25f, k1 = 0.5f, k2 = 0.25f;                             // :337-342, 3x3 gaussian
    for (unsigned int y = 0; y < h; ++y) for (unsigned int x = 0; x < w; ++x) {
        const float l = lum[(size_t)y * w + (x ? x - 1 : 0)], c = lum[(size_t)y * w + x], r = lum[(size_t)y * w + std::min(x + 1, w - 1)];
        volatile float a = l * k0, b = c * k1, d = r * k2; volatile float ab = a + b;
        tmp[(size_t)y * w + x] = ab + d;
    }
    result.resize((size_t)w * h);
    for (unsigned int y = 0; y < h; ++y) for (unsigned int x = 0; x < w; ++x) {
        const float u = tmp[(size_t)(y ? y - 1 : 0) * w + x], c = tmp[(size_t)y * w + x], d = tmp[(size_t)std::min(y + 1, h - 1) * w + x];
        volatile float a = u * k0, b = c * k1, e = d * k2; volatile float ab = a + b;
        result[(size_t)y * w + x] = ab + e;
    }
    outW = w; outH = h;
    return true;
}

// ---- image.cpp:349-389 -------------------------------------------------------------------------------------------------
static float calculateImageIntegral(const std::vector<float>& image, unsigned int imageWidth, unsigned int imageHeight, float* functionU)
{
    const float iW = (float)imageWidth, iH = (float)imageHeight, iA = iW * iH;
    float textureTimesSinSum = 0;
    for (unsigned int y = 0; y < imageHeight; ++y) {
        const float sinTheta = (float)sin(M_PI * ((float)y + 0.5f) / iH);
        for (unsigned int x = 0; x < imageWidth; ++x) {
            const float value = std::max(0.f, image[(size_t)y * imageWidth + x]);
            volatile float f = value * sinTheta;
            functionU[(size_t)y * imageWidth + x] = f;
            textureTimesSinSum += f;
        }
    }
    float integral = textureTimesSinSum / iA;
    integral *= 2.0f * M_PI * M_PI;          // float * double constants, as in the reference
    return integral;
}

// ---- image.cpp:68-283 ---------------------------------------------------------------------------------------------------
bool calculateCDF(const float* rgbPixels, unsigned int imageWidth, unsigned int imageHeight,
                  std::vector<float>& cdfUData, unsigned int& cdfUDataWidth, unsigned int& cdfUDataHeight,
                  std::vector<float>& cdfVData, float& environmentTextureIntegral)
{
    std::vector<float> filtered;
    if (!generateImageFunction(rgbPixels, imageWidth, imageHeight, filtered, imageWidth, imageHeight)) return false;
    cdfUDataWidth = imageWidth + 1;
    cdfUDataHeight = imageHeight;
    cdfUData.assign((size_t)cdfUDataWidth * cdfUDataHeight, 0.0f);
    cdfVData.assign(imageHeight + 1, 0.0f);
    std::vector<float> functionU((size_t)imageWidth * imageHeight), functionV(imageHeight + 1);
    environmentTextureIntegral = calculateImageIntegral(filtered, imageWidth, imageHeight, &functionU[0]);

    const unsigned int numStepsW = imageWidth;
    for (unsigned int y = 0; y < imageHeight; ++y) {
        float* row = &cdfUData[(size_t)y * cdfUDataWidth];
        row[0] = 0.0f;
        for (unsigned int x = 1; x <= imageWidth; ++x) row[x] = row[x - 1] + functionU[(size_t)y * imageWidth + x - 1] / numStepsW;
        const float rowIntegral = row[imageWidth];
        functionV[y] = rowIntegral;
        if (rowIntegral > 0.0f) for (unsigned int x = 1; x <= imageWidth; ++x) row[x] /= rowIntegral;
        else for (unsigned int x = 1; x <= imageWidth; ++x) row[x] = (float)x / numStepsW;     // black row: uniform
    }
    cdfVData[0] = 0.0f;
    for (unsigned int y = 1; y <= imageHeight; ++y) cdfVData[y] = cdfVData[y - 1] + functionV[y - 1] / imageHeight;
    const float imageIntegral = cdfVData[imageHeight];
    if (imageIntegral > 0.0f) for (unsigned int y = 1; y <= imageHeight; ++y) cdfVData[y] /= imageIntegral;
    else for (unsigned int y = 1; y <= imageHeight; ++y) cdfVData[y] = (float)y / (float)imageHeight;
    return true;
}
