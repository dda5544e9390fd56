// image.cpp -- environment-map ingest and the marginal / conditional CDF construction used to importance-sample
// it (PBRT-2 14.6.5). calculateCDF follows renderer/image.cpp:68-283 and :349-389 of the reference. The reference
// delegates file I/O, down-scaling, luminance and the 3x3 blur to OpenImageIO (un-vendored, un-pinned: SURVEY 8c);
// this build owns that chain: PFM / Radiance-HDR readers, box down-scale to <= 512, Rec.709 luminance, separable
// [1 2 1]/4 blur with clamped edges.
#include "vt_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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

// ---- image function: image.cpp:285-346 (generateImageFunction) -------------------------------------------------------
static bool generateImageFunction(const float* rgbPixels, unsigned int imageWidth, unsigned int imageHeight,
                                  std::vector<float>& result, unsigned int& outW, unsigned int& outH)
{
    std::vector<float> src(rgbPixels, rgbPixels + (size_t)imageWidth * imageHeight * 3);
    unsigned int w = imageWidth, h = imageHeight;
    const unsigned int m = std::max(w, h);
    if (m > MAX_CDF_SIZE) {                                                    // :309-321
        const unsigned int nw = (unsigned int)((float)w / m * MAX_CDF_SIZE);
        const unsigned int nh = (unsigned int)((float)h / m * MAX_CDF_SIZE);
        if (nw == 0 || nh == 0) return false;
        const unsigned int fx = w / nw, fy = h / nh;
        if (fx * nw != w || fy * nh != h) return false;                        // box filter needs integer factors
        std::vector<float> small((size_t)nw * nh * 3);
        for (unsigned int y = 0; y < nh; ++y) for (unsigned int x = 0; x < nw; ++x) for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (unsigned int j = 0; j < fy; ++j) for (unsigned int i = 0; i < fx; ++i)
                s += src[((size_t)(y * fy + j) * w + (x * fx + i)) * 3 + c];
            small[((size_t)y * nw + x) * 3 + c] = (float)(s / (double)(fx * fy));
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
