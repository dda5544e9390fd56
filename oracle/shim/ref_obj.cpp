// TEST INFRASTRUCTURE -- oracle/_ref/libvt_ref_obj.so: the reference's OWN OBJ reader. thirdParty/tinyobjloader/tiny_obj_loader.cc
// is compiled where it lies under $(REF) (nothing is copied); this file only adds the merge of the shapes into one index
// space that MeshLoader::loadFromOBJ performs (mesh/meshLoader.cpp:27-64 -- that file itself needs Imath through mesh/mesh.h
// and cannot be compiled here), restated below, and a C entry point for ctypes.
#include "thirdParty/tinyobjloader/tiny_obj_loader.h"

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern "C" {

// returns 0 on success; *verts (3 floats per vertex) and *idx are malloc'ed, free with vtref_obj_free
int vtref_load_obj(const char* path, float** verts, size_t* n_floats, unsigned int** idx, size_t* n_idx)
{
    std::vector<tinyobj::shape_t> shapes;
    std::vector<tinyobj::material_t> materials;
    const std::string err = tinyobj::LoadObj(shapes, materials, path);        // meshLoader.cpp:19
    if (!err.empty()) return -1;
    size_t total_v = 0, total_i = 0;
    for (size_t s = 0; s < shapes.size(); ++s) { total_i += shapes[s].mesh.indices.size(); total_v += shapes[s].mesh.positions.size(); }   // :27-38
    *verts = (float*)malloc(sizeof(float) * (total_v ? total_v : 1));
    *idx = (unsigned int*)malloc(sizeof(unsigned int) * (total_i ? total_i : 1));
    size_t vo = 0, io = 0;
    for (size_t s = 0; s < shapes.size(); ++s) {                                // :40-63
        const tinyobj::shape_t& shape = shapes[s];
        const size_t nv = shape.mesh.positions.size() / 3, ni = shape.mesh.indices.size();
        if (nv) memcpy(*verts + 3 * vo, &shape.mesh.positions[0], nv * 3 * sizeof(float));
        for (size_t i = 0; i < ni; ++i) (*idx)[i + io] = shape.mesh.indices[i] + (unsigned int)vo;
        vo += nv; io += ni;
    }
    *n_floats = total_v; *n_idx = total_i;
    return 0;
}
void vtref_obj_free(void* p) { free(p); }

} // extern "C"
