/*
 * vto.h -- TEST INFRASTRUCTURE: public interface of the CPU oracle.
 *
 * The oracle is a plain-C restatement of the reference's device programs
 * (GLSL under /root/reference/src/shaders) and of the few host routines that
 * produce their inputs. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it. The product
 * (voxeltoy_b200/) never does.
 *
 * PARITY PINNING: the reference ships no tests, golden images or KATs
 * (SURVEY.md section 4). The oracle is pinned by
 *   (1) oracle/_ref/libvt_ref.so -- the reference's OWN sources compiled here:
 *       src/voxelize/cpuVoxelizer.cpp, src/thirdParty/tinyobjloader,
 *       src/renderer/noise.cpp, src/camera/ (Imath/Qt/boost replaced by
 *       the API shims in oracle/shim/), and the reference's GLSL shaders
 *       themselves, run on the CPU through oracle/shim/glsl_emu.h;
 *   (2) the self-derived KATs of SURVEY.md section 8(c) (tests/test_oracle_kat.py).
 * What stays unpinned: GL-driver behaviour (transcendental precision, bilinear
 * filter arithmetic, out-of-range texelFetch) and OpenImageIO's resize/blur --
 * see DESIGN.md "Parity status".
 */
#ifndef VTO_H
#define VTO_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Everything the integrator shaders read (pathTracer.fs:6-44 uniforms + textures). */
typedef struct vto_scene {
    /* volume: materialOffsetTexture (R32I, x fastest), renderer.cpp:854-872 */
    int32_t X, Y, Z;
    const int32_t* grid;
    /* materialDataTexture (R32F 1-D), renderer.cpp:874-887 */
    const float* materials;
    int32_t n_materials;            /* number of floats */
    /* emissiveVoxelIndicesTexture, renderer.cpp:889-902 */
    const int32_t* emissive;
    int32_t n_emissive;
    /* noiseTexture RGBA32F, renderer.cpp:740-760 */
    const float* noise;
    int32_t noise_w, noise_h;
    /* environment, renderer.cpp:947-1055 */
    int32_t use_image;
    const float* env_rgb;           /* 3 floats per texel, row 0 first */
    int32_t env_w, env_h;
    const float* cdf_u;             /* cdf_u_w x cdf_u_h */
    int32_t cdf_u_w, cdf_u_h;
    const float* cdf_v;
    int32_t cdf_v_n;
    float env_integral;
    float env_rotation;             /* radians */
    float bg_top[3], bg_bottom[3];
    /* camera: row-major host matrices (uploaded with transpose=GL_TRUE, renderer.cpp:420-433) */
    float inv_modelview[16];
    float proj[16];
    float inv_proj[16];
    float lens_radius;
    int32_t lens_model;             /* 0 pinhole, 1 thin lens, 2 orthographic */
    float focal_distance;           /* FocalDistanceData SSBO */
    /* frame */
    int32_t W, H;
    int32_t max_bounces;
    float wire_opacity, wire_thickness;
    int32_t sel_index[3];           /* SelectVoxelData.index.xyz */
    /* world bounds of the volume (derived by vto_volume_bounds) */
    float bmin[3], bmax[3], voxel_size[3];
} vto_scene;

/* algorithmic work counters of SURVEY.md 8(d) */
typedef struct vto_counters {
    uint64_t S;   /* DDA iterations that fetched a voxel */
    uint64_t R;   /* rand() calls */
    uint64_t Hm;  /* material record evaluations */
    uint64_t E;   /* CDF texel loads */
    uint64_t Q;   /* environment-map lookups */
    uint64_t paths;
} vto_counters;

/* renderer.cpp:845-850 + :926-929 */
void vto_volume_bounds(int X, int Y, int Z, float bmin[3], float bmax[3], float voxel_size[3]);

/* K1: one pass of integrator/pathTracer.fs over all pixels. out_rgba: W*H*4 (row 0 = bottom row).
 * primary_hit (optional, W*H): linear voxel index of the primary hit, -1 miss, -2 ground.
 * steps (optional, W*H): DDA iterations spent by this pixel's whole path. */
void vto_render_pass(const vto_scene* s, int sample_count, float* out_rgba,
                     int32_t* primary_hit, int32_t* steps, vto_counters* counters, int n_threads);

/* K1 on a list of n pixels (xy: n pairs); out_rgba: 4 floats per pixel, primary_hit optional (n) */
void vto_render_pixels(const vto_scene* s, int sample_count, const int32_t* xy, size_t n, float* out_rgba,
                       int32_t* primary_hit, int n_threads);

/* K4: integrator/editMode.fs */
void vto_preview_pass(const vto_scene* s, int sample_count, float* out_rgba, int n_threads);

/* K2: shared/accumulation.fs: avg' = (s + avg*n)/(n+1) over count floats */
void vto_accumulate(float* avg, const float* sample, int n, size_t count);

/* K5: shared/voxelize.{vs,gs} == voxelize/cpuVoxelizer.cpp. occupancy: X*Y*Z bytes, set to 1 where written. */
void vto_voxelize(const float* xyz, size_t n_verts, const uint32_t* idx, size_t n_idx,
                  const float M[16], int X, int Y, int Z, uint8_t* occupancy, int n_threads);

void vto_voxelize_fat(const float* xyz, size_t n_verts, const uint32_t* idx, size_t n_idx,
                      const float M[16], int X, int Y, int Z, uint8_t* occupancy, int n_threads);

/* K6: editVoxels/selectVoxel.vs; viewport = (x,y,w,h). index[4], normal[4] out. */
void vto_pick(const vto_scene* s, const float viewport[4], float near_z, float px, float py,
              int32_t index[4], float normal[4]);
/* K9: focalDistance/focalDistance.vs */
float vto_pick_focal(const vto_scene* s, const float viewport[4], float px, float py);
/* K7: editVoxels/addVoxel.vs; returns 1 and the written coordinate/offset, 0 if out of bounds */
int vto_add_voxel(const vto_scene* s, const int32_t sel_index[4], const float sel_normal[4],
                  float motion_x, float motion_y, int32_t* grid_rw, int32_t coord[3]);
/* K8: editVoxels/removeVoxel.vs */
int vto_remove_voxel(int X, int Y, int Z, const int32_t sel_index[4], int32_t* grid_rw);

/* the DDA alone on explicit rays (6 floats each: origin, direction); out: 4 floats per ray */
void vto_trace_rays(const vto_scene* s, const float* rays, size_t n, float* out);

/* host-side producers of kernel inputs */
void vto_noise_table(float* out, size_t n_floats);                    /* renderer.cpp:741-744 */
uint32_t vto_hash(uint32_t seed);                                     /* random.h:3-11 */
void vto_rng_offset(int px, int py, int sequence, int rw, int rh, int out[2]); /* random.h:13-18 */
int vto_dda_step_cap(int X, int Y, int Z);                            /* dda.h:98 */

/* image.cpp:309-321 as this build defines it: area-weighted box reduction of an RGB float image to nw x nh */
void vto_resize_box(const float* rgb, int w, int h, int nw, int nh, float* out);
/* image.cpp:68-283,349-389 on an already filtered single-channel image */
void vto_build_cdf(const float* lum, int w, int h, float* cdf_u /*(w+1)*h*/, float* cdf_v /*h+1*/,
                   float* integral);

/* scalar math exposed for the accuracy tests */
float vto_m_sin(float), vto_m_cos(float), vto_m_acos(float), vto_m_atan2(float, float),
      vto_m_pow(float, float), vto_m_exp2(float), vto_m_log2(float);

#ifdef __cplusplus
}
#endif
#endif
