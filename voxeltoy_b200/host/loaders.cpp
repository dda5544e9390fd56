// loaders.cpp -- scene ingest on the host: material records, MagicaVoxel .vox, Wavefront .obj, Mesh.
// Semantics follow renderer/material/material.cpp, renderer/loaders/voxLoader.cpp,
// renderer/loaders/magicaVoxel.cpp, mesh/meshLoader.cpp, mesh/mesh.cpp and the vendored 2013
// tinyobjloader the reference links (thirdParty/tinyobjloader/tiny_obj_loader.cc) -- restated, not copied.
#include "vt_host.h"

#include <algorithm>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

using namespace vtm;

// ---- Material serialisation (material.cpp:8-81) -----------------------------------------------------------------
namespace Material {
static SerializedData leafFloat(const std::string& name, size_t offset, float value, float from, float to)
{
    SerializedData d;
    d.m_propertyName = name; d.m_dataOffset = offset; d.m_value = value;
    d.m_dataRangeFrom = from; d.m_dataRangeTo = to; d.m_propertyType = SerializedData::PROPERTY_TYPE_FLOAT;
    return d;
}
static SerializedData leafColor(const std::string& name, size_t offset, const float rgb[3])
{
    SerializedData d;
    d.m_propertyName = name; d.m_dataOffset = offset; d.m_propertyType = SerializedData::PROPERTY_TYPE_COLOR;
    d.m_dataRangeFrom = d.m_dataRangeTo = d.m_value = 0;
    static const char* ch[3] = { "red", "green", "blue" };
    for (int i = 0; i < 3; ++i) d.m_childProperties.push_back(leafFloat(ch[i], offset + i, rgb[i], 0, 1));
    return d;
}
static SerializedData node(const char* name, size_t offset)
{
    SerializedData d;
    d.m_propertyName = name; d.m_dataOffset = offset; d.m_propertyType = SerializedData::PROPERTY_TYPE_FLOAT;
    d.m_value = d.m_dataRangeFrom = d.m_dataRangeTo = 0;
    return d;
}
SerializedData serializeLambert(const LambertMaterialData& m, size_t off)
{
    SerializedData d = node("Lambert Material", off);
    d.m_childProperties.push_back(leafColor("emission", off + 1, m.emission));
    d.m_childProperties.push_back(leafColor("albedo", off + 4, m.albedo));
    return d;
}
SerializedData serializeMetal(const MetalMaterialData& m, size_t off)
{
    SerializedData d = node("Metal Material", off);
    d.m_childProperties.push_back(leafColor("emission", off + 1, m.emission));
    d.m_childProperties.push_back(leafColor("reflectance", off + 4, m.reflectance));
    d.m_childProperties.push_back(leafFloat("roughness", off + 7, m.roughness, 0, 1000));
    return d;
}
SerializedData serializePlastic(const PlasticMaterialData& m, size_t off)
{
    SerializedData d = node("Plastic Material", off);
    d.m_childProperties.push_back(leafColor("emission", off + 1, m.emission));
    d.m_childProperties.push_back(leafColor("diffuseAlbedo", off + 4, m.diffuseAlbedo));
    d.m_childProperties.push_back(leafFloat("roughness", off + 7, m.roughness, 0, 1000));
    return d;
}
} // namespace Material

// ---- VoxLoader material packers (voxLoader.cpp:8-91): record = [type][emission rgb][colour rgb]([roughness]) ----
static void appendRecord(std::vector<float>& out, Material::MaterialType type, V3f emission, V3f colour, const float* roughness)
{
    out.push_back((float)type);
    out.push_back(emission.x); out.push_back(emission.y); out.push_back(emission.z);
    out.push_back(colour.x); out.push_back(colour.y); out.push_back(colour.z);
    if (roughness) out.push_back(*roughness);
}
void VoxLoader::generateMaterialLambert(V3f emission, V3f albedo, std::vector<float>& materialData)
{
    appendRecord(materialData, Material::MT_LAMBERT, emission, albedo, NULL);
}
void VoxLoader::generateMaterialMetal(V3f emission, V3f reflectance, float roughness, std::vector<float>& materialData)
{
    appendRecord(materialData, Material::MT_METAL, emission, reflectance, &roughness);
}
void VoxLoader::generateMaterialPlastic(V3f emission, V3f diffuseAlbedo, float roughness, std::vector<float>& materialData)
{
    appendRecord(materialData, Material::MT_PLASTIC, emission, diffuseAlbedo, &roughness);
}
float VoxLoader::getMaterialEmisiveness(const float* rec)
{
    const int type = (int)rec[0];
    if (type < Material::MT_LAMBERT || type > Material::MT_PLASTIC) return 0.0f;
    return (rec[1] + rec[2] + rec[3]) / 3;
}

// ---- MagicaVoxel .vox v150 (magicaVoxel.cpp:122-321) ---------------------------------------------------------------
namespace {
struct ByteReader {
    const unsigned char* p; size_t n, pos;
    bool i32(int32_t& v) { if (pos + 4 > n) { v = 0; pos = n; return false; } memcpy(&v, p + pos, 4); pos += 4; return true; }
};
const int32_t kVOX = 0x20584f56, kMAIN = 0x4e49414d, kSIZE = 0x455a4953, kXYZI = 0x495a5958, kRGBA = 0x41424752;

// The MagicaVoxel default palette (magicaVoxel.cpp:245-263) is a closed form: index 0 is 0, then a 6x6x6 cube over
// the levels ff,cc,99,66,33,00 (blue fastest, then green, then red; its final all-zero entry is omitted), then
// 10-step ramps (ee,dd,bb,aa,88,77,55,44,22,11) of red, green, blue and grey. Byte order r,g,b,a.
void defaultPaletteRGBA(unsigned char out[256][4])
{
    static const unsigned char lv[6] = { 0xff, 0xcc, 0x99, 0x66, 0x33, 0x00 };
    static const unsigned char ramp[10] = { 0xee, 0xdd, 0xbb, 0xaa, 0x88, 0x77, 0x55, 0x44, 0x22, 0x11 };
    memset(out, 0, 256 * 4);
    int k = 1;
    for (int r = 0; r < 6; ++r) for (int g = 0; g < 6; ++g) for (int b = 0; b < 6; ++b) {
        if (r == 5 && g == 5 && b == 5) continue;
        out[k][0] = lv[r]; out[k][1] = lv[g]; out[k][2] = lv[b]; out[k][3] = 0xff; ++k;
    }
    for (int c = 0; c < 3; ++c) for (int i = 0; i < 10; ++i) { out[k][c] = ramp[i]; out[k][3] = 0xff; ++k; }
    for (int i = 0; i < 10; ++i) { out[k][0] = out[k][1] = out[k][2] = ramp[i]; out[k][3] = 0xff; ++k; }
}
} // namespace

bool MagicaVoxelLoader::loadFromMemory(const unsigned char* bytes, size_t n, std::vector<int32_t>& voxelMaterials,
                                       std::vector<float>& materialData, std::vector<int32_t>& emissiveVoxelIndices,
                                       V3i& voxelResolution)
{
    ByteReader rd = { bytes, n, 0 };
    int32_t magic, version, id, contentSize, childrenSize;
    rd.i32(magic);
    if (magic != kVOX) { m_error = "magic number does not match"; return false; }
    rd.i32(version);
    if (version != 150) { m_error = "version does not match"; return false; }
    rd.i32(id); rd.i32(contentSize); rd.i32(childrenSize);
    if (id != kMAIN) { m_error = "main chunk is not found"; return false; }
    // chunk sizes come from the file: negative or oversized values must not move the cursor backwards or past the data
    // (the reference trusts them, magicaVoxel.cpp:150-206)
    if (contentSize < 0 || childrenSize < 0 || (size_t)contentSize > n - rd.pos) { m_error = "corrupt chunk"; return false; }
    const size_t mainEnd = std::min(n, rd.pos + (size_t)contentSize + (size_t)childrenSize);
    rd.pos += contentSize;

    int sx = 0, sy = 0, sz = 0, numVoxels = 0;
    const unsigned char* voxels = NULL;
    unsigned char palette[256][4];
    bool customPalette = false;
    while (rd.pos < mainEnd && rd.pos < n) {
        if (rd.pos + 12 > mainEnd) break;                                   // trailing bytes that cannot hold a chunk header
        rd.i32(id); rd.i32(contentSize); rd.i32(childrenSize);
        if (contentSize < 0 || childrenSize < 0) { m_error = "corrupt chunk"; return false; }
        const size_t end = rd.pos + (size_t)contentSize + (size_t)childrenSize;
        if (end > mainEnd) { m_error = "corrupt chunk"; return false; }
        if (id == kSIZE) { int32_t a, b, c; rd.i32(a); rd.i32(b); rd.i32(c); sx = a; sy = b; sz = c; }
        else if (id == kXYZI) {
            int32_t cnt; rd.i32(cnt);
            if (cnt < 0) { m_error = "negative number of voxels"; return false; }
            if (rd.pos + (size_t)cnt * 4 > n) { m_error = "truncated XYZI chunk"; return false; }
            numVoxels = cnt; voxels = bytes + rd.pos;
        } else if (id == kRGBA) {
            if (rd.pos + 256 * 4 > n) { m_error = "truncated RGBA chunk"; return false; }
            customPalette = true;
            memset(palette[0], 0, 4);
            memcpy(palette[1], bytes + rd.pos, 255 * 4);      // file entry i is colour index i+1; the last one is reserved
        }
        rd.pos = end;
    }
    if (!customPalette) defaultPaletteRGBA(palette);

    // y <-> z axis conversion (magicaVoxel.cpp:276-278, 293-295)
    voxelResolution = V3i(sx, sz, sy);
    if (sx <= 0 || sy <= 0 || sz <= 0) { m_error = "empty model"; return false; }
    voxelMaterials.assign((size_t)sx * sy * sz, -1);
    materialData.clear();
    int32_t offsetOfColour[256];
    for (int i = 0; i < 256; ++i) offsetOfColour[i] = -1;
    for (int i = 0; i < numVoxels; ++i) {
        const unsigned char* v = voxels + 4 * (size_t)i;
        const int ci = v[3];
        if (ci == 254) continue;                                // "empty voxel" colour index (:297)
        if (v[0] >= sx || v[1] >= sy || v[2] >= sz) continue;   // malformed record: the reference would write out of bounds
        const size_t cell = (size_t)v[0] + (size_t)v[2] * voxelResolution.x + (size_t)v[1] * voxelResolution.x * voxelResolution.y;
        if (offsetOfColour[ci] < 0) {                           // first use of this colour: one Lambert record (:300-318)
            offsetOfColour[ci] = (int32_t)materialData.size();
            const V3f colour((float)palette[ci][0] / 255, (float)palette[ci][1] / 255, (float)palette[ci][2] / 255);
            const PaletteRule* rule = NULL;
            for (size_t r = 0; r < m_rules.size(); ++r) if (m_rules[r].colorIndex == ci) rule = &m_rules[r];
            if (!rule) generateMaterialLambert(V3f(0, 0, 0), colour, materialData);
            else if (rule->type == Material::MT_METAL) generateMaterialMetal(rule->emission, colour, rule->roughness, materialData);
            else if (rule->type == Material::MT_PLASTIC) generateMaterialPlastic(rule->emission, colour, rule->roughness, materialData);
            else generateMaterialLambert(rule->emission, colour, materialData);
        }
        voxelMaterials[cell] = offsetOfColour[ci];
    }
    // declared contract (voxLoader.h:23-24): every voxel whose material emits light, in index order
    emissiveVoxelIndices.clear();
    for (size_t i = 0; i < voxelMaterials.size(); ++i)
        if (voxelMaterials[i] >= 0 && getMaterialEmisiveness(&materialData[voxelMaterials[i]]) > 0) emissiveVoxelIndices.push_back((int32_t)i);
    return true;
}

static bool readWholeFile(const std::string& path, std::vector<unsigned char>& out)
{
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END);
    const long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    out.resize(sz > 0 ? (size_t)sz : 0);
    const size_t got = out.empty() ? 0 : fread(&out[0], 1, out.size(), fp);
    fclose(fp);
    return got == out.size();
}

bool MagicaVoxelLoader::load(const std::string& filePath, std::vector<int32_t>& voxelMaterials, std::vector<float>& materialData,
                             std::vector<int32_t>& emissiveVoxelIndices, V3i& voxelResolution)
{
    std::vector<unsigned char> bytes;
    if (!readWholeFile(filePath, bytes)) { m_error = "failed to open file"; return false; }
    return loadFromMemory(bytes.empty() ? NULL : &bytes[0], bytes.size(), voxelMaterials, materialData, emissiveVoxelIndices, voxelResolution);
}

// ---- Mesh + OBJ ----------------------------------------------------------------------------------------------------------
Box3f computeBounds(const float* vertices, size_t numVertices)                 // mesh.cpp:61-75
{
    Box3f b;
    for (size_t v = 0; v < numVertices; ++v) b.extendBy(V3f(vertices[3 * v], vertices[3 * v + 1], vertices[3 * v + 2]));
    return b;
}
Mesh::Mesh(const float* vertices, size_t numVertices, const unsigned int* indices, size_t numIndices)
    : m_vertices(vertices, vertices + 3 * numVertices), m_indices(indices, indices + numIndices)
{
    m_bounds = computeBounds(vertices, numVertices);
}

namespace {
struct Corner {
    int v, vt, vn;
    bool operator<(const Corner& o) const
    {
        if (v != o.v) return v < o.v;
        if (vn != o.vn) return vn < o.vn;
        return vt < o.vt;
    }
};
inline int fixIndex(int idx, int n) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : n + idx); }
inline const char* skipBlank(const char* p) { while (*p == ' ' || *p == '\t') ++p; return p; }
inline const char* skipToken(const char* p) { while (*p && *p != ' ' && *p != '\t' && *p != '\r') ++p; return p; }
inline float parseFloatTok(const char*& p) { p = skipBlank(p); const float f = (float)atof(p); p = skipToken(p); return f; }
Corner parseCorner(const char*& p, int nv, int nvn, int nvt)
{
    Corner c; c.v = c.vt = c.vn = -1;
    c.v = fixIndex(atoi(p), nv);
    p += strcspn(p, "/ \t\r");
    if (*p != '/') return c;
    ++p;
    if (*p == '/') { ++p; c.vn = fixIndex(atoi(p), nvn); p += strcspn(p, "/ \t\r"); return c; }   // i//k
    c.vt = fixIndex(atoi(p), nvt);
    p += strcspn(p, "/ \t\r");
    if (*p != '/') return c;
    ++p;
    c.vn = fixIndex(atoi(p), nvn);
    p += strcspn(p, "/ \t\r");
    return c;
}
// one "shape" = the faces gathered since the last g / o statement; vertices are emitted in first-use order
// and polygons become triangle fans (tiny_obj_loader.cc:226-266)
void flushShape(const std::vector<std::vector<Corner> >& faces, const std::vector<float>& pos,
                std::vector<float>& outVerts, std::vector<unsigned int>& outIdx)
{
    if (faces.empty()) return;
    std::map<Corner, unsigned int> cache;
    const unsigned int base = (unsigned int)(outVerts.size() / 3);
    unsigned int next = 0;
    for (size_t f = 0; f < faces.size(); ++f) {
        const std::vector<Corner>& face = faces[f];
        for (size_t k = 2; k < face.size(); ++k) {
            const Corner tri[3] = { face[0], face[k - 1], face[k] };
            for (int j = 0; j < 3; ++j) {
                std::map<Corner, unsigned int>::iterator it = cache.find(tri[j]);
                unsigned int id;
                if (it != cache.end()) id = it->second;
                else {
                    id = next++;
                    cache[tri[j]] = id;
                    const size_t s = 3 * (size_t)tri[j].v;
                    if (s + 2 < pos.size()) { outVerts.push_back(pos[s]); outVerts.push_back(pos[s + 1]); outVerts.push_back(pos[s + 2]); }
                    else { outVerts.push_back(0); outVerts.push_back(0); outVerts.push_back(0); }
                }
                outIdx.push_back(base + id);
            }
        }
    }
}
} // namespace

bool MeshLoader::loadFromOBJMemory(const char* text, size_t n, std::vector<float>& vertices, std::vector<unsigned int>& indices)
{
    vertices.clear(); indices.clear();
    std::vector<float> pos; int nvn = 0, nvt = 0;
    std::vector<std::vector<Corner> > faces;
    size_t i = 0;
    std::string line;
    while (i < n) {
        size_t e = i;
        while (e < n && text[e] != '\n') ++e;
        line.assign(text + i, e - i);
        i = e + 1;
        if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
        const char* p = skipBlank(line.c_str());
        if (*p == '\0' || *p == '#') continue;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            p += 2;
            const float x = parseFloatTok(p), y = parseFloatTok(p), z = parseFloatTok(p);
            pos.push_back(x); pos.push_back(y); pos.push_back(z);
        } else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) ++nvn;
        else if (p[0] == 'v' && p[1] == 't' && (p[2] == ' ' || p[2] == '\t')) ++nvt;
        else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            p = skipBlank(p + 2);
            std::vector<Corner> face;
            while (*p && *p != '\r' && *p != '\n') {
                face.push_back(parseCorner(p, (int)(pos.size() / 3), nvn, nvt));
                p += strspn(p, " \t\r");
            }
            faces.push_back(face);
        } else if ((p[0] == 'g' || p[0] == 'o') && (p[1] == ' ' || p[1] == '\t')) {
            flushShape(faces, pos, vertices, indices);
            faces.clear();
        } else if (strncmp(p, "usemtl", 6) == 0 && (p[6] == ' ' || p[6] == '\t')) {
            faces.clear();          // the 2013 loader drops the faces gathered so far here (tiny_obj_loader.cc "usemtl" branch)
        }
    }
    flushShape(faces, pos, vertices, indices);
    return !indices.empty();
}

void MeshLoader::loadFromOBJ(const char* filePath, std::vector<float>& vertices, std::vector<unsigned int>& indices)
{
    std::vector<unsigned char> bytes;
    vertices.clear(); indices.clear();
    if (!readWholeFile(filePath, bytes)) { fprintf(stderr, "Cannot open file [%s]\n", filePath); return; }
    loadFromOBJMemory(bytes.empty() ? "" : (const char*)&bytes[0], bytes.size(), vertices, indices);
}
Mesh* MeshLoader::loadFromOBJ(const char* filePath)
{
    std::vector<float> vertices; std::vector<unsigned int> indices;
    loadFromOBJ(filePath, vertices, indices);
    if (indices.empty()) return NULL;
    return new Mesh(&vertices[0], vertices.size() / 3, &indices[0], indices.size());
}
