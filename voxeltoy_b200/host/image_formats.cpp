// image_formats.cpp -- OpenEXR and PNG files for the environment map and for Renderer::saveImage.
//
// The reference hands every image file to OpenImageIO (renderer/image.cpp:28-59 reads "any format" into float RGB,
// renderer.cpp:1108-1140 writes whatever the extension says). OpenImageIO and OpenEXR are not part of this build; the two
// formats that matter in practice for this path are restated here on top of zlib alone:
//   * OpenEXR, the usual container of HDR environment maps: single-part scanline files, HALF / FLOAT / UINT channels,
//     compression NONE, RLE, ZIPS and ZIP (file layout: openexr.com "OpenEXR File Layout"); written as FLOAT + ZIP.
//     Tiled, deep and multi-part files and the PIZ / PXR24 / B44 / DWA codecs are refused with a message.
//   * PNG (RFC 2083): 8 / 16-bit grey, grey + alpha, RGB, RGBA and 1..8-bit palette images, non-interlaced -- the reference's own
//     resources/*.png are of that kind. Samples map to floats as OpenImageIO does: v * (1 / (2^bits - 1)), no transfer function.
#include "vt_host.h"

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>

namespace {

bool slurp(const std::string& path, std::vector<unsigned char>& out)
{
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END);
    const long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (n < 0) { fclose(fp); return false; }
    out.resize((size_t)n);
    const bool ok = n == 0 || fread(&out[0], 1, (size_t)n, fp) == (size_t)n;
    fclose(fp);
    return ok;
}

struct Reader {
    const unsigned char* p; size_t n, at = 0; bool ok = true;
    Reader(const std::vector<unsigned char>& v) : p(v.empty() ? nullptr : &v[0]), n(v.size()) {}
    bool need(size_t k) { if (!ok || k > n - at) { ok = false; return false; } return true; }
    uint32_t u32() { if (!need(4)) return 0; uint32_t v; memcpy(&v, p + at, 4); at += 4; return v; }       // little endian hosts only (x86-64)
    uint64_t u64() { if (!need(8)) return 0; uint64_t v; memcpy(&v, p + at, 8); at += 8; return v; }
    int32_t i32() { return (int32_t)u32(); }
    unsigned char u8() { if (!need(1)) return 0; return p[at++]; }
    std::string cstr(size_t limit = 256)
    {
        std::string s;
        while (ok) { if (!need(1)) break; const char c = (char)p[at++]; if (!c) break; s.push_back(c); if (s.size() > limit) { ok = false; break; } }
        return s;
    }
};

float half_to_float(uint16_t h)
{
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu, bits;
    if (e == 0) {
        if (m == 0) bits = sign;
        else { e = 113; while (!(m & 0x400u)) { m <<= 1; --e; } bits = sign | (e << 23) | ((m & 0x3ffu) << 13); }     // subnormal half
    } else if (e == 31) bits = sign | 0x7f800000u | (m << 13);
    else bits = sign | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &bits, 4);
    return f;
}

// the byte predictor + even / odd interleave OpenEXR applies before its zlib and RLE codecs
void exr_unfilter(std::vector<unsigned char>& t, std::vector<unsigned char>& out)
{
    const size_t n = t.size();
    for (size_t i = 1; i < n; ++i) t[i] = (unsigned char)(t[i - 1] + t[i] - 128);
    out.resize(n);
    const size_t half = (n + 1) / 2;
    for (size_t i = 0; i < n; ++i) out[i] = (i & 1) ? t[half + i / 2] : t[i / 2];
}
void exr_filter(const unsigned char* raw, size_t n, std::vector<unsigned char>& t)
{
    t.resize(n);
    const size_t half = (n + 1) / 2;
    for (size_t i = 0; i < n; ++i) { if (i & 1) t[half + i / 2] = raw[i]; else t[i / 2] = raw[i]; }
    unsigned char prev = n ? t[0] : 0;
    for (size_t i = 1; i < n; ++i) { const unsigned char cur = t[i]; t[i] = (unsigned char)(cur - prev + 128); prev = cur; }
}

bool exr_unrle(const unsigned char* in, size_t n_in, std::vector<unsigned char>& out, size_t n_out)
{
    out.clear(); out.reserve(n_out);
    size_t i = 0;
    while (i < n_in) {
        const int c = (signed char)in[i++];
        if (c < 0) { const size_t k = (size_t)(-c); if (k > n_in - i) return false; out.insert(out.end(), in + i, in + i + k); i += k; }
        else { if (i >= n_in) return false; out.insert(out.end(), (size_t)c + 1, in[i++]); }
        if (out.size() > n_out) return false;
    }
    return out.size() == n_out;
}

struct ExrChannel { std::string name; int type; int xs, ys; };

} // namespace

bool readEXR(const std::string& path, unsigned int& outWidth, unsigned int& outHeight, std::vector<float>& rgb, std::string* why)
{
    auto fail = [&](const char* m) { if (why) *why = m; return false; };
    std::vector<unsigned char> file;
    if (!slurp(path, file)) return fail("cannot read the file");
    Reader r(file);
    if (r.u32() != 20000630u) return fail("not an OpenEXR file");
    const uint32_t version = r.u32();
    if ((version & 0xffu) != 2u) return fail("unsupported OpenEXR version");
    if (version & (0x200u | 0x800u | 0x1000u)) return fail("tiled, deep and multi-part OpenEXR files are not supported");
    std::vector<ExrChannel> channels;
    int compression = -1, line_order = 0;
    int dw[4] = { 0, 0, -1, -1 };
    for (;;) {
        const std::string name = r.cstr();
        if (!r.ok) return fail("truncated header");
        if (name.empty()) break;
        const std::string type = r.cstr();
        const int32_t size = r.i32();
        if (!r.ok || size < 0 || !r.need((size_t)size)) return fail("truncated header");
        const size_t end = r.at + (size_t)size;
        if (name == "channels" && type == "chlist") {
            while (r.ok && r.at < end) {
                ExrChannel c;
                c.name = r.cstr();
                if (c.name.empty()) break;
                c.type = r.i32(); r.u8(); r.u8(); r.u8(); r.u8(); c.xs = r.i32(); c.ys = r.i32();
                channels.push_back(c);
            }
        } else if (name == "compression") compression = r.u8();
        else if (name == "dataWindow") { for (int i = 0; i < 4; ++i) dw[i] = r.i32(); }
        else if (name == "lineOrder") line_order = r.u8();
        r.at = end;
    }
    (void)line_order;                                   // every chunk carries its y: the order in the file does not matter
    if (channels.empty() || compression < 0 || dw[2] < dw[0] || dw[3] < dw[1]) return fail("incomplete header");
    if (compression > 3) return fail("only NONE, RLE, ZIPS and ZIP compression are supported (not PIZ / PXR24 / B44 / DWA)");
    const int64_t W = (int64_t)dw[2] - dw[0] + 1, H = (int64_t)dw[3] - dw[1] + 1;
    if (W <= 0 || H <= 0 || W > 65536 || H > 65536 || W * H > ((int64_t)1 << 28)) return fail("unreasonable data window");
    size_t line_bytes = 0;
    std::vector<size_t> ch_off(channels.size());
    for (size_t i = 0; i < channels.size(); ++i) {
        if (channels[i].xs != 1 || channels[i].ys != 1) return fail("subsampled channels are not supported");
        if (channels[i].type < 0 || channels[i].type > 2) return fail("unknown pixel type");
        ch_off[i] = line_bytes;
        line_bytes += (size_t)W * (channels[i].type == 1 ? 2 : 4);
    }
    // which channel feeds R, G, B: exact names first, then the last component of layered names, then luminance Y
    int src[3] = { -1, -1, -1 };
    const char* want[3] = { "R", "G", "B" };
    for (int k = 0; k < 3; ++k) {
        for (size_t i = 0; i < channels.size(); ++i) if (channels[i].name == want[k]) src[k] = (int)i;
        if (src[k] < 0) for (size_t i = 0; i < channels.size(); ++i) {
            const std::string& n = channels[i].name;
            if (n.size() > 2 && n[n.size() - 2] == '.' && n[n.size() - 1] == want[k][0]) { src[k] = (int)i; break; }
        }
    }
    if (src[0] < 0 || src[1] < 0 || src[2] < 0) {
        int y = -1;
        for (size_t i = 0; i < channels.size(); ++i) if (channels[i].name == "Y") y = (int)i;
        if (y < 0) y = 0;                               // a single arbitrary channel: shown as grey
        src[0] = src[1] = src[2] = y;
    }
    const int lines_per_block = compression == 3 ? 16 : 1;
    const int64_t n_blocks = (H + lines_per_block - 1) / lines_per_block;
    std::vector<uint64_t> offsets((size_t)n_blocks);
    for (int64_t i = 0; i < n_blocks; ++i) offsets[(size_t)i] = r.u64();
    if (!r.ok) return fail("truncated offset table");
    rgb.assign((size_t)W * (size_t)H * 3, 0.0f);
    std::vector<unsigned char> tmp, raw;
    for (int64_t b = 0; b < n_blocks; ++b) {
        Reader c(file);
        c.at = (size_t)std::min<uint64_t>(offsets[(size_t)b], file.size());
        const int32_t y0 = c.i32(), size = c.i32();
        if (!c.ok || size < 0 || !c.need((size_t)size)) return fail("truncated pixel data");
        if (y0 < dw[1] || y0 > dw[3]) return fail("scanline outside the data window");
        const int lines = (int)std::min<int64_t>(lines_per_block, (int64_t)dw[3] - y0 + 1);
        const size_t want_bytes = line_bytes * (size_t)lines;
        const unsigned char* data = c.p + c.at;
        if ((size_t)size == want_bytes || compression == 0) {          // stored as is (also when compression would not have helped)
            if ((size_t)size != want_bytes) return fail("bad uncompressed block size");
            raw.assign(data, data + size);
        } else if (compression == 1) {
            if (!exr_unrle(data, (size_t)size, tmp, want_bytes)) return fail("bad RLE block");
            exr_unfilter(tmp, raw);
        } else {
            tmp.resize(want_bytes);
            uLongf got = (uLongf)want_bytes;
            if (uncompress(&tmp[0], &got, data, (uLong)size) != Z_OK || got != want_bytes) return fail("bad zlib block");
            exr_unfilter(tmp, raw);
        }
        for (int l = 0; l < lines; ++l) {
            const unsigned char* line = &raw[(size_t)l * line_bytes];
            float* dst = &rgb[(size_t)(y0 - dw[1] + l) * (size_t)W * 3];
            for (int k = 0; k < 3; ++k) {
                const ExrChannel& ch = channels[(size_t)src[k]];
                const unsigned char* s = line + ch_off[(size_t)src[k]];
                for (int64_t x = 0; x < W; ++x) {
                    float v;
                    if (ch.type == 1) { uint16_t h; memcpy(&h, s + 2 * x, 2); v = half_to_float(h); }
                    else if (ch.type == 2) memcpy(&v, s + 4 * x, 4);
                    else { uint32_t u; memcpy(&u, s + 4 * x, 4); v = (float)u; }
                    dst[3 * x + k] = v;
                }
            }
        }
    }
    outWidth = (unsigned int)W; outHeight = (unsigned int)H;
    return true;
}

// FLOAT channels (A), B, G, R; ZIP compression (blocks of 16 lines); rows top-down; `channels` = 3 or 4 interleaved floats per pixel
bool writeEXR(const std::string& path, const float* pixels, unsigned int w, unsigned int h, int channels)
{
    if (!pixels || w == 0 || h == 0 || (channels != 3 && channels != 4)) return false;
    std::vector<unsigned char> hd;
    auto put = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; hd.insert(hd.end(), b, b + n); };
    auto put32 = [&](int32_t v) { put(&v, 4); };
    auto puts0 = [&](const char* s) { put(s, strlen(s) + 1); };
    auto attr = [&](const char* name, const char* type, const std::vector<unsigned char>& v) { puts0(name); puts0(type); put32((int32_t)v.size()); put(v.data(), v.size()); };
    const uint32_t magic = 20000630u, version = 2u;
    put(&magic, 4); put(&version, 4);
    const char* names4[4] = { "A", "B", "G", "R" };
    const int n_ch = channels;
    {
        std::vector<unsigned char> v;
        for (int i = (channels == 4 ? 0 : 1); i < 4; ++i) {
            v.insert(v.end(), names4[i], names4[i] + 2);
            const int32_t type = 2, one = 1; const unsigned char lin[4] = { 0, 0, 0, 0 };
            v.insert(v.end(), (const unsigned char*)&type, (const unsigned char*)&type + 4); v.insert(v.end(), lin, lin + 4);
            v.insert(v.end(), (const unsigned char*)&one, (const unsigned char*)&one + 4); v.insert(v.end(), (const unsigned char*)&one, (const unsigned char*)&one + 4);
        }
        v.push_back(0);
        attr("channels", "chlist", v);
    }
    attr("compression", "compression", std::vector<unsigned char>(1, 3));
    {
        const int32_t box[4] = { 0, 0, (int32_t)w - 1, (int32_t)h - 1 };
        std::vector<unsigned char> v((const unsigned char*)box, (const unsigned char*)box + 16);
        attr("dataWindow", "box2i", v); attr("displayWindow", "box2i", v);
    }
    attr("lineOrder", "lineOrder", std::vector<unsigned char>(1, 0));
    { const float one = 1.0f; attr("pixelAspectRatio", "float", std::vector<unsigned char>((const unsigned char*)&one, (const unsigned char*)&one + 4)); }
    { const float z[2] = { 0.0f, 0.0f }; attr("screenWindowCenter", "v2f", std::vector<unsigned char>((const unsigned char*)z, (const unsigned char*)z + 8)); }
    { const float one = 1.0f; attr("screenWindowWidth", "float", std::vector<unsigned char>((const unsigned char*)&one, (const unsigned char*)&one + 4)); }
    hd.push_back(0);
    const unsigned int n_blocks = (h + 15) / 16;
    const size_t line_bytes = (size_t)w * 4 * (size_t)n_ch;
    std::vector<std::vector<unsigned char> > blocks(n_blocks);
    std::vector<unsigned char> raw, filtered;
    // source component of the file's channels in file order: (A) B G R
    const int comp4[4] = { 3, 2, 1, 0 };
    for (unsigned int b = 0; b < n_blocks; ++b) {
        const unsigned int y0 = b * 16, lines = std::min(16u, h - y0);
        raw.resize(line_bytes * lines);
        for (unsigned int l = 0; l < lines; ++l) {
            const float* src = pixels + (size_t)(y0 + l) * w * (size_t)channels;
            unsigned char* dst = &raw[(size_t)l * line_bytes];
            for (int c = 0; c < n_ch; ++c) {
                const int comp = comp4[c + (channels == 4 ? 0 : 1)];
                for (unsigned int x = 0; x < w; ++x) memcpy(dst + ((size_t)c * w + x) * 4, src + (size_t)x * channels + comp, 4);
            }
        }
        exr_filter(raw.data(), raw.size(), filtered);
        uLongf bound = compressBound((uLong)filtered.size());
        std::vector<unsigned char>& out = blocks[b];
        out.resize(bound);
        if (compress2(&out[0], &bound, filtered.data(), (uLong)filtered.size(), 6) != Z_OK) return false;
        if (bound >= raw.size()) out = raw; else out.resize(bound);          // a block that does not shrink is stored as is
    }
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    fwrite(hd.data(), 1, hd.size(), fp);
    uint64_t off = hd.size() + (uint64_t)n_blocks * 8;
    for (unsigned int b = 0; b < n_blocks; ++b) { fwrite(&off, 8, 1, fp); off += 8 + blocks[b].size(); }
    for (unsigned int b = 0; b < n_blocks; ++b) {
        const int32_t y0 = (int32_t)(b * 16), size = (int32_t)blocks[b].size();
        fwrite(&y0, 4, 1, fp); fwrite(&size, 4, 1, fp); fwrite(blocks[b].data(), 1, blocks[b].size(), fp);
    }
    const bool ok = ferror(fp) == 0;
    fclose(fp);
    return ok;
}

bool readPNG(const std::string& path, unsigned int& outWidth, unsigned int& outHeight, std::vector<float>& rgb, std::string* why)
{
    auto fail = [&](const char* m) { if (why) *why = m; return false; };
    std::vector<unsigned char> file;
    if (!slurp(path, file)) return fail("cannot read the file");
    static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    if (file.size() < 8 || memcmp(&file[0], sig, 8) != 0) return fail("not a PNG file");
    auto be32 = [&](size_t at) { return ((uint32_t)file[at] << 24) | ((uint32_t)file[at + 1] << 16) | ((uint32_t)file[at + 2] << 8) | (uint32_t)file[at + 3]; };
    uint32_t W = 0, H = 0; int depth = 0, ctype = -1, interlace = 0;
    std::vector<unsigned char> idat, plte;
    size_t at = 8;
    bool end = false;
    while (!end && at + 12 <= file.size()) {
        const uint32_t len = be32(at);
        if (len > file.size() - at - 12) return fail("truncated chunk");
        const char* type = (const char*)&file[at + 4];
        const unsigned char* data = &file[at + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) return fail("bad IHDR");
            W = be32(at + 8); H = be32(at + 12); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) end = true;
        at += 12 + (size_t)len;
    }
    if (W == 0 || H == 0 || W > 65536 || H > 65536 || (uint64_t)W * H > ((uint64_t)1 << 28)) return fail("bad image size");
    if (interlace != 0) return fail("interlaced PNG files are not supported");
    int comps;
    switch (ctype) { case 0: comps = 1; break; case 2: comps = 3; break; case 3: comps = 1; break; case 4: comps = 2; break; case 6: comps = 4; break; default: return fail("bad colour type"); }
    const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                          (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                          ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
    if (!depth_ok) return fail("bad bit depth");
    if (ctype == 3 && plte.size() < 3) return fail("palette image without PLTE");
    const size_t bpp = std::max<size_t>(1, (size_t)comps * depth / 8);              // bytes per complete pixel, for the filters
    const size_t stride = ((size_t)W * comps * depth + 7) / 8;
    std::vector<unsigned char> raw((stride + 1) * (size_t)H);
    uLongf got = (uLongf)raw.size();
    if (idat.empty() || uncompress(&raw[0], &got, idat.data(), (uLong)idat.size()) != Z_OK || got != raw.size()) return fail("bad image data");
    std::vector<unsigned char> prev(stride, 0), cur(stride);
    rgb.assign((size_t)W * H * 3, 0.0f);
    const float scale = 1.0f / (float)((1u << depth) - 1u);
    for (uint32_t y = 0; y < H; ++y) {
        const unsigned char* in = &raw[(stride + 1) * (size_t)y];
        const int filter = in[0];
        if (filter > 4) return fail("bad filter type");
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int pred = 0;
            switch (filter) {
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                default: break;
            }
            cur[i] = (unsigned char)(in[1 + i] + pred);
        }
        float* dst = &rgb[(size_t)y * W * 3];
        for (uint32_t x = 0; x < W; ++x) {
            unsigned int v[4] = { 0, 0, 0, 0 };
            for (int k = 0; k < comps; ++k) {
                const size_t s = (size_t)x * comps + k;
                if (depth == 16) v[k] = ((unsigned int)cur[2 * s] << 8) | cur[2 * s + 1];
                else if (depth == 8) v[k] = cur[s];
                else { const size_t bit = s * depth; v[k] = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u); }
            }
            if (ctype == 3) {
                const size_t e = (size_t)v[0] * 3;
                if (e + 2 < plte.size()) { const float s8 = 1.0f / 255.0f; dst[3 * x] = plte[e] * s8; dst[3 * x + 1] = plte[e + 1] * s8; dst[3 * x + 2] = plte[e + 2] * s8; }
            } else if (comps <= 2) { dst[3 * x] = dst[3 * x + 1] = dst[3 * x + 2] = (float)v[0] * scale; }
            else { dst[3 * x] = (float)v[0] * scale; dst[3 * x + 1] = (float)v[1] * scale; dst[3 * x + 2] = (float)v[2] * scale; }
        }
        prev.swap(cur);
    }
    outWidth = W; outHeight = H;
    return true;
}
