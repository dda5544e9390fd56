// ref_glsl.cpp -- TEST INFRASTRUCTURE: harness that runs the reference's OWN GLSL programs on the CPU.
// The shader bodies are generated at build time from /root/reference/src/shaders by glsl2cpp.py (see Makefile) and
// included below; this file only plays the role of src/renderer/renderer.cpp + the rasteriser: it binds the
// uniforms / textures / SSBOs (renderer.cpp:410-444, :854-902, :1071-1101; servicePicking.cpp:31-44) from a
// vto_scene and invokes main() once per fragment / vertex / triangle.
#include "glsl_emu.h"
#include "../vto.h"
#include <omp.h>
#include <vector>

namespace glsl {
#include "gen_PathTracerFS.inc"
#include "gen_EditModeFS.inc"
#include "gen_SelectVoxelVS.inc"
#include "gen_FocalDistanceVS.inc"
#include "gen_AddVoxelVS.inc"
#include "gen_RemoveVoxelVS.inc"
#include "gen_VoxelizeVS.inc"
#include "gen_VoxelizeGS.inc"
#include "gen_VoxelizeGSFat.inc"

static mat4 from_row_major(const float* m)      // glUniformMatrix4fv(..., GL_TRUE, m): M[c][r] = m[4r + c]
{
    mat4 r;
    for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) r.c[c].d[k] = m[4 * k + c];
    return r;
}

// uniforms shared by the integrators and the picking services
template <class Sh> static void bind_camera_volume(Sh& sh, const vto_scene* s, const float* noise, int nw, int nh)
{
    sh.materialOffsetTexture.p = s->grid; sh.materialOffsetTexture.X = s->X; sh.materialOffsetTexture.Y = s->Y; sh.materialOffsetTexture.Z = s->Z;
    sh.noiseTexture.p = noise; sh.noiseTexture.w = nw; sh.noiseTexture.h = nh; sh.noiseTexture.ch = 4;
    sh.voxelResolution = ivec3(s->X, s->Y, s->Z);
    sh.volumeBoundsMin = vec3(s->bmin[0], s->bmin[1], s->bmin[2]);
    sh.volumeBoundsMax = vec3(s->bmax[0], s->bmax[1], s->bmax[2]);
    sh.wsVoxelSize = vec3(s->voxel_size[0], s->voxel_size[1], s->voxel_size[2]);
    sh.viewport = vec4(0.f, 0.f, (float)s->W, (float)s->H);
    sh.cameraProj = from_row_major(s->proj);
    sh.cameraInverseProj = from_row_major(s->inv_proj);
    sh.cameraInverseModelView = from_row_major(s->inv_modelview);
    sh.cameraLensRadius = s->lens_radius;
    sh.cameraLensModel = s->lens_model;
}

template <class Sh> static void bind_integrator(Sh& sh, const vto_scene* s, int sample_count)
{
    bind_camera_volume(sh, s, s->noise, s->noise_w, s->noise_h);
    sh.materialDataTexture.p = s->materials; sh.materialDataTexture.n = s->n_materials;
    sh.backgroundColorTop = vec3(s->bg_top[0], s->bg_top[1], s->bg_top[2]);
    sh.backgroundColorBottom = vec3(s->bg_bottom[0], s->bg_bottom[1], s->bg_bottom[2]);
    sh.backgroundUseImage = s->use_image;
    sh.backgroundTexture.p = s->env_rgb; sh.backgroundTexture.w = s->env_w; sh.backgroundTexture.h = s->env_h; sh.backgroundTexture.ch = 3;
    sh.backgroundCDFUTexture.p = s->cdf_u; sh.backgroundCDFUTexture.w = s->cdf_u_w; sh.backgroundCDFUTexture.h = s->cdf_u_h; sh.backgroundCDFUTexture.ch = 1;
    sh.backgroundCDFVTexture.p = s->cdf_v; sh.backgroundCDFVTexture.n = s->cdf_v_n;
    sh.backgroundIntegral = s->env_integral;
    sh.backgroundRotationRadians = s->env_rotation;
    sh.sampleCount = sample_count;
    sh.wireframeOpacity = s->wire_opacity;
    sh.wireframeThickness = s->wire_thickness;
    sh.FocalDistanceData.focalDistance = s->focal_distance;
    sh.SelectVoxelData.index = ivec4(s->sel_index[0], s->sel_index[1], s->sel_index[2], 0);
}

template <class Sh> static void run_fullscreen(const Sh& proto, const vto_scene* s, float* out_rgba, int n_threads)
{
    // drawFullscreenQuad (renderer.cpp:647-656): one fragment per pixel, gl_FragCoord = (x + .5, y + .5, 0.55, 1)
    #pragma omp parallel num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
    {
        Sh sh = proto;
        #pragma omp for schedule(dynamic, 4)
        for (int y = 0; y < s->H; ++y)
            for (int x = 0; x < s->W; ++x) {
                sh.gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.55f, 1.0f);
                sh.main();
                float* o = out_rgba + 4 * ((size_t)x + (size_t)y * s->W);
                o[0] = sh.outColor.x; o[1] = sh.outColor.y; o[2] = sh.outColor.z; o[3] = sh.outColor.w;
            }
    }
}
} // namespace glsl

using namespace glsl;

extern "C" {

const char* vtref_describe(void)
{
    return "reference GLSL (integrator/pathTracer.fs, integrator/editMode.fs, editVoxels/*.vs, focalDistance/focalDistance.vs, "
           "shared/voxelize.{vs,gs} + spliced headers) compiled for the CPU through oracle/shim/glsl_emu.h";
}

// K1: integrator/pathTracer.fs with "#define PINHOLE\n#define THINLENS\n" (renderer.cpp:227)
void vtref_render_pass(const vto_scene* s, int sample_count, float* out_rgba, int n_threads)
{
    PathTracerFS sh;
    bind_integrator(sh, s, sample_count);
    sh.materialDataTexture.p = s->materials;
    sh.emissiveVoxelIndicesTexture.p = s->emissive; sh.emissiveVoxelIndicesTexture.n = s->n_emissive;
    sh.pathtracerMaxNumBounces = s->max_bounces;
    run_fullscreen(sh, s, out_rgba, n_threads);
}

// K1 on a list of pixels (x, y pairs): the same shader invocation as the full-screen quad would run at those fragments
void vtref_render_pixels(const vto_scene* s, int sample_count, const int32_t* xy, size_t n, float* out_rgba, int n_threads)
{
    PathTracerFS proto;
    bind_integrator(proto, s, sample_count);
    proto.materialDataTexture.p = s->materials;
    proto.emissiveVoxelIndicesTexture.p = s->emissive; proto.emissiveVoxelIndicesTexture.n = s->n_emissive;
    proto.pathtracerMaxNumBounces = s->max_bounces;
    #pragma omp parallel num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
    {
        PathTracerFS sh = proto;
        #pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < (long)n; ++i) {
            sh.gl_FragCoord = vec4((float)xy[2 * i] + 0.5f, (float)xy[2 * i + 1] + 0.5f, 0.55f, 1.0f);
            sh.main();
            float* o = out_rgba + 4 * (size_t)i;
            o[0] = sh.outColor.x; o[1] = sh.outColor.y; o[2] = sh.outColor.z; o[3] = sh.outColor.w;
        }
    }
}

// K4: integrator/editMode.fs
void vtref_preview_pass(const vto_scene* s, int sample_count, float* out_rgba, int n_threads)
{
    EditModeFS sh;
    bind_integrator(sh, s, sample_count);
    run_fullscreen(sh, s, out_rgba, n_threads);
}

// The services never bind noiseTexture / cameraLensModel (servicePicking.cpp:31-44). The contract both the
// oracle and the product implement is the INTENDED un-jittered pinhole ray: a 1x1 noise texel of 0.5 gives jitter 0.
static const float kHalfNoise[4] = { 0.5f, 0.5f, 0.5f, 0.5f };

// K6: editVoxels/selectVoxel.vs with "#define PINHOLE\n" (servicePicking.cpp:21)
void vtref_pick(const vto_scene* s, float near_z, float px, float py, const float prev_normal[4], int32_t index[4], float normal[4])
{
    SelectVoxelVS sh;
    bind_camera_volume(sh, s, kHalfNoise, 1, 1);
    sh.cameraLensModel = 0;
    sh.cameraNear = near_z;
    sh.sampledFragment = vec2(px, py);
    sh.SelectVoxelData.normal = vec4(prev_normal[0], prev_normal[1], prev_normal[2], prev_normal[3]);
    sh.main();
    for (int i = 0; i < 4; ++i) { index[i] = sh.SelectVoxelData.index.d[i]; normal[i] = sh.SelectVoxelData.normal.d[i]; }
}

// K9: focalDistance/focalDistance.vs
float vtref_pick_focal(const vto_scene* s, float px, float py)
{
    FocalDistanceVS sh;
    bind_camera_volume(sh, s, kHalfNoise, 1, 1);
    sh.cameraLensModel = 0;
    sh.sampledFragment = vec2(px, py);
    sh.main();
    return sh.FocalDistanceData.focalDistance;
}

// K7: editVoxels/addVoxel.vs. Returns the coordinate the shader stores to (the value stored is stale).
void vtref_add_voxel(const vto_scene* s, const int32_t sel_index[4], const float sel_normal[4], float mx, float my, int32_t coord[3])
{
    AddVoxelVS sh;
    sh.cameraInverseModelView = from_row_major(s->inv_modelview);
    sh.screenSpaceMotion = vec2(mx, my);
    sh.SelectVoxelData.index = ivec4(sel_index[0], sel_index[1], sel_index[2], sel_index[3]);
    sh.SelectVoxelData.normal = vec4(sel_normal[0], sel_normal[1], sel_normal[2], sel_normal[3]);
    sh.voxelOccupancy.X = s->X; sh.voxelOccupancy.Y = s->Y; sh.voxelOccupancy.Z = s->Z;
    sh.main();
    coord[0] = sh.voxelOccupancy.last.x; coord[1] = sh.voxelOccupancy.last.y; coord[2] = sh.voxelOccupancy.last.z;
}

// K8: editVoxels/removeVoxel.vs
void vtref_remove_voxel(const int32_t sel_index[4], int32_t coord[3])
{
    RemoveVoxelVS sh;
    sh.SelectVoxelData.index = ivec4(sel_index[0], sel_index[1], sel_index[2], sel_index[3]);
    sh.voxelOccupancy.X = sh.voxelOccupancy.Y = sh.voxelOccupancy.Z = 1 << 30;
    sh.main();
    coord[0] = sh.voxelOccupancy.last.x; coord[1] = sh.voxelOccupancy.last.y; coord[2] = sh.voxelOccupancy.last.z;
}

// K5: shared/voxelize.vs per vertex + shared/voxelize.gs per triangle (Mesh::draw = glDrawElements(GL_TRIANGLES), mesh.cpp:55-59)
} // extern "C"
template <class GS>
static void run_voxelizer(const float* xyz, size_t n_verts, const uint32_t* idx, size_t n_idx, const float M[16],
                          int X, int Y, int Z, uint8_t* occupancy, int n_threads)
{
    std::vector<vec3> vs(n_verts);
    {
        VoxelizeVS v;
        v.voxelResolution = ivec3(X, Y, Z);
        v.modelTransform = from_row_major(M);
        for (size_t i = 0; i < n_verts; ++i) {
            v.in_vertexPos = vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
            v.main();
            vs[i] = v.Out.vsVertexPos;
        }
    }
    const long n_tris = (long)(n_idx / 3);
    #pragma omp parallel num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
    {
        GS g;
        g.voxelResolution = ivec3(X, Y, Z);
        g.voxelOccupancy.occ = occupancy; g.voxelOccupancy.X = X; g.voxelOccupancy.Y = Y; g.voxelOccupancy.Z = Z;
        #pragma omp for schedule(dynamic, 16)
        for (long t = 0; t < n_tris; ++t) {
            for (int k = 0; k < 3; ++k) g.In[k].vsVertexPos = vs[idx[3 * t + k]];
            g.main();
        }
    }
}
extern "C" {
void vtref_voxelize(const float* xyz, size_t n_verts, const uint32_t* idx, size_t n_idx, const float M[16],
                    int X, int Y, int Z, uint8_t* occupancy, int n_threads)
{
    run_voxelizer<VoxelizeGS>(xyz, n_verts, idx, n_idx, M, X, Y, Z, occupancy, n_threads);
}
// the FAT variant of the same shader text (voxelize.gs:15-19 with `#define THICKNESS FAT`)
void vtref_voxelize_fat(const float* xyz, size_t n_verts, const uint32_t* idx, size_t n_idx, const float M[16],
                        int X, int Y, int Z, uint8_t* occupancy, int n_threads)
{
    run_voxelizer<VoxelizeGSFat>(xyz, n_verts, idx, n_idx, M, X, Y, Z, occupancy, n_threads);
}

} // extern "C"
