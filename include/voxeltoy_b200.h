/*
 * voxeltoy_b200.h -- C ABI of the B200-native voxelToy hot path.
 *
 * This is the device boundary of the reference: every entry point below replaces a
 * group of OpenGL calls the reference's C++ classes (src/renderer, src/voxelize,
 * src/renderer/services) issue directly. The reference has no FFI of its own; the
 * file:line after each declaration names the GL call site it replaces
 * (paths relative to /root/reference/src).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no C++/torch types cross the boundary.
 *  - every function returns 0 (VT_OK) or a negative vt_status; the message of the
 *    last failure on a context is vt_last_error(ctx).
 *  - host pointers are borrowed for the duration of the call and copied.
 *  - device memory is owned by the context; one context per GPU; a context is
 *    externally synchronised (one host thread at a time -- the reference's single
 *    GL thread).
 *  - work is enqueued on the context's CUDA stream; only vt_read_* / vt_sync /
 *    vt_get_* block.
 *  - matrices are row-major float[16] in the column-vector convention, exactly the
 *    arrays the reference hands to glUniformMatrix4fv(..., GL_TRUE, &m.x[0][0]).
 *  - images are W*H RGBA float32, row 0 = BOTTOM row (GL window origin); the host
 *    Renderer::saveImage flips (renderer/renderer.cpp:1132-1136).
 *  - there is NO CPU fallback: without a CUDA device vt_create fails.
 */
#ifndef VOXELTOY_B200_H
#define VOXELTOY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vt_ctx vt_ctx;

typedef enum vt_status {
    VT_OK = 0,
    VT_ERR_INVALID = -1,      /* bad argument / call order                      */
    VT_ERR_CUDA = -2,         /* a CUDA runtime call failed (see vt_last_error) */
    VT_ERR_NO_DEVICE = -3,    /* no usable CUDA device: there is no CPU path    */
    VT_ERR_STATE = -4         /* scene / frame not set up yet                   */
} vt_status;

/* cameraLensModel, camera/cameraParameters.h:23-28 */
enum { VT_LENS_PINHOLE = 0, VT_LENS_THIN = 1, VT_LENS_ORTHO = 2 };
/* Renderer::Integrator, renderer/renderer.h:113-118 */
enum { VT_INTEGRATOR_PATHTRACER = 0, VT_INTEGRATOR_EDIT_MODE = 1 };
/* work partition of one frame across `world` contexts */
enum { VT_PART_NONE = 0, VT_PART_TILES = 1, VT_PART_SAMPLES = 2 };

/* uniforms uploaded by Renderer::updateCamera, renderer/renderer.cpp:410-444 */
typedef struct vt_camera {
    float inv_modelview[16];   /* cameraInverseModelView  :420-423 */
    float proj[16];            /* cameraProj              :425-428 */
    float inv_proj[16];        /* cameraInverseProj       :430-433 */
    float near_z, far_z;       /* cameraNear / cameraFar  :417-418 (near feeds the pick ray, selectVoxel.vs:41) */
    float lens_radius;         /* cameraLensRadius        :439 */
    int32_t lens_model;        /* cameraLensModel         :440 */
} vt_camera;

/* uniforms uploaded by Renderer::updateRenderSettings, renderer/renderer.cpp:1071-1101,
 * plus the frame size of resizeFrame (:448-554) and the integrator switch (:676-680) */
typedef struct vt_settings {
    int32_t width, height;             /* viewport = (0,0,width,height) :468-472 */
    int32_t max_bounces;               /* pathtracerMaxNumBounces */
    int32_t integrator;                /* VT_INTEGRATOR_* */
    float bg_top[3], bg_bottom[3];     /* backgroundColorTop / Bottom */
    int32_t use_env_image;             /* backgroundUseImage */
    float env_rotation_rad;            /* backgroundRotationRadians */
    float wireframe_opacity, wireframe_thickness;
} vt_settings;

/* algorithmic work counters: bytes_per_sample = 4S + 16R + 36H + 4E + 64Q + 32 */
typedef struct vt_counters {
    uint64_t dda_steps;     /* S: DDA iterations that fetched a voxel */
    uint64_t rand_calls;    /* R */
    uint64_t material_evals;/* H */
    uint64_t cdf_loads;     /* E */
    uint64_t env_lookups;   /* Q */
    uint64_t paths;         /* samples (pixel-passes) */
    uint64_t kernel_launches; /* launches of this library's kernels since vt_create */
} vt_counters;

typedef void (*vt_log_fn)(const char* msg, void* user);   /* log/logger.h:7-11 */

/* ---- lifetime ------------------------------------------------------------------ */
int vt_create(int device, vt_ctx** out);               /* Renderer::initialize, renderer.cpp:76-147 (glewInit + resource creation) */
void vt_destroy(vt_ctx* ctx);
const char* vt_last_error(const vt_ctx* ctx);          /* Renderer::getStatus, renderer.h:95 */
int vt_set_logger(vt_ctx* ctx, vt_log_fn fn, void* user); /* Renderer::setLogger, renderer.cpp:72-75 */
int vt_set_stream(vt_ctx* ctx, void* cuda_stream);     /* interop: run on a caller-owned cudaStream_t (NULL = the context's own) */
int vt_sync(vt_ctx* ctx);                              /* glFinish */

/* ---- scene upload: Renderer::createVoxelDataTexture, renderer.cpp:833-940 -------- */
/* R32I material-offset grid, x fastest, -1 = empty (glTexImage3D :863-872). NULL = all empty.
 * Derives world bounds / voxel size (:845-850) and the bit-packed occupancy + its mip. */
int vt_volume_upload(vt_ctx* ctx, const int32_t* mat_offsets, int X, int Y, int Z);
int vt_materials_upload(vt_ctx* ctx, const float* data, size_t n_floats);            /* glTexImage1D :879-887 */
int vt_material_update(vt_ctx* ctx, uint32_t offset, const float* values, int n);    /* updateMaterialColor/Value, renderer.cpp:1205-1235 */
int vt_emissive_upload(vt_ctx* ctx, const int32_t* voxel_indices, size_t n);         /* glTexImage1D :894-902 */
int vt_read_volume(vt_ctx* ctx, int32_t* mat_offsets_out);                           /* glGetTexImage of the offset texture */
int vt_read_materials(vt_ctx* ctx, float* out, size_t n_floats);                     /* getMaterials, renderer.cpp:1142-1160 */
int vt_get_volume_info(vt_ctx* ctx, int32_t res[3], float bounds_min[3], float bounds_max[3], float voxel_size[3]);

/* noise texture, renderer.cpp:740-760. rgba == NULL: the reference's own table (first w*h*4 outputs
 * of glibc rand()/RAND_MAX with the default seed), generated inside the library. */
int vt_noise_upload(vt_ctx* ctx, const float* rgba, int w, int h);

/* environment map + CDFs: Renderer::loadBackgroundImage, renderer.cpp:983-1046 */
int vt_env_upload(vt_ctx* ctx, const float* rgb, int w, int h,
                  const float* cdf_u, int cdf_u_w, int cdf_u_h,
                  const float* cdf_v, int cdf_v_n, float integral);
int vt_env_clear(vt_ctx* ctx);
/* The processing half of Renderer::loadBackgroundImage on the device (renderer.cpp:947-1055 + renderer/image.cpp:68-389:
 * RGBA conversion, <= 512 box reduction, luminance, 3x3 gaussian, sin(theta) weights, integral, CDF-U / CDF-V):
 * the caller passes the decoded RGB float image (row 0 = first scanline of the file, as OIIO delivers it) and gets the
 * state vt_env_upload would have set from the host-built arrays, bit for bit. Fails (VT_ERR_INVALID) like
 * calculateCDF when the image cannot be box-reduced by integer factors. */
int vt_env_build(vt_ctx* ctx, const float* rgb, int w, int h);
/* dims[6] = {env w, env h, CDF-U width, CDF-U height, CDF-V length, guide tables in use}; *integral = the uniform
 * backgroundIntegral (renderer.cpp:1048-1052); *build_ms = device time of the last vt_env_build. Any pointer may be NULL. */
int vt_get_env_info(vt_ctx* ctx, int32_t* dims, float* integral, float* build_ms);
/* read back the CDF textures (sizes from vt_get_env_info); either pointer may be NULL */
int vt_read_env_cdf(vt_ctx* ctx, float* cdf_u, float* cdf_v);

/* ---- per-frame state -------------------------------------------------------------- */
int vt_set_camera(vt_ctx* ctx, const vt_camera* cam);       /* updateCamera, renderer.cpp:410-444 */
int vt_set_settings(vt_ctx* ctx, const vt_settings* st);    /* updateRenderSettings :1057-1106, resizeFrame :448-554 */
int vt_set_focal_distance(vt_ctx* ctx, float d);            /* FocalDistanceData SSBO, renderer.cpp:712-720 */
int vt_get_focal_distance(vt_ctx* ctx, float* d);
int vt_set_selection(vt_ctx* ctx, const int32_t index[4], const float normal[4]);  /* SelectVoxelData SSBO, renderer.cpp:723-737 */
int vt_get_selection(vt_ctx* ctx, int32_t index[4], float normal[4]);

/* ---- the hot path: one or more progressive passes ------------------------------------ */
int vt_reset_accumulation(vt_ctx* ctx);                     /* resetRender, renderer.cpp:941-945 */
/* K1+K2 fused: passes first_sample .. first_sample+n_passes-1 of the current integrator, each folded into
 * the running average with n = number of passes accumulated so far (renderer.cpp:594-611, accumulation.fs:10-18).
 * first_sample is the `sampleCount` uniform of the first pass. */
int vt_render(vt_ctx* ctx, int first_sample, int n_passes);
int vt_get_num_samples(vt_ctx* ctx, int* n);                /* m_numberSamples */
int vt_read_average(vt_ctx* ctx, float* rgba_out);          /* glGetTexImage(average), renderer.cpp:1119-1127 */
/* the display blit (shared/textureMap.fs, renderer.cpp:613-637) as a read-out: W*H RGBA8, float -> UNORM8 as a GL
 * framebuffer write does it; flip_vertical != 0 gives top-down rows (saveImage's order, renderer.cpp:1131-1136) */
int vt_read_display(vt_ctx* ctx, uint8_t* rgba8_out, int flip_vertical);
int vt_read_primary_hits(vt_ctx* ctx, int32_t* out);        /* per pixel: linear voxel index, -1 miss, -2 ground (last pass) */
int vt_enable_primary_hits(vt_ctx* ctx, int enable);
/* multi-GPU: restrict this context to its share of the frame (tiles: 64x64 tiles dealt round-robin, result bit-identical
 * to one GPU) or of the samples (rank r of `world` renders sampleCount = (first_sample + p) * world + r and keeps a SUM). */
int vt_set_partition(vt_ctx* ctx, int mode, int rank, int world);
/* sample-partition mode keeps a running SUM instead of an average; expose the device buffer so the caller's
 * collective (NCCL through torch.distributed) can reduce it in place. W*H float4. */
void* vt_accum_device_ptr(vt_ctx* ctx);
/* render kernel variant: 0 = one-thread-per-pixel megakernel, 2 = wavefront with per-material shade queues and
 * self-refilling trace warps (default); both produce identical bits (1 was round 1's per-lane state machine: removed,
 * measured slower) */
int vt_set_kernel_variant(vt_ctx* ctx, int variant);
/* wavefront variant: upper bound on the paths (pixel-passes) in flight per batch; the passes of one vt_render call are
 * split into batches of floor(max_paths / pixels) passes (at least 1). Tuning knob, no effect on results. */
int vt_set_wavefront_max_paths(vt_ctx* ctx, size_t max_paths);
/* wavefront variant: number of batches in flight on internal streams (1..4, default 1). With 2 the ramp-down tail of one
 * batch's kernels is filled by the next batch (+2-4 % throughput); accumulation stays in pass order. No effect on results. */
int vt_set_wavefront_lanes(vt_ctx* ctx, int lanes);
/* exact empty-space skip of the wavefront DDA (csrc/vt_device.cuh dda_skip): 0 off, 1 auto (default: volumes whose every
 * side is >= 64 voxels), 2 always. Bit-identical results in every mode (the skipped steps' float additions are still done). */
int vt_set_empty_skip(vt_ctx* ctx, int mode);
/* device time of the wavefront kernels by kind, measured with cudaEvent pairs around every launch on the context's
 * stream while enabled; vt_get_kernel_times synchronises, returns the sums since the last call and clears them.
 * (the reference's only timer is the per-frame GL_TIMESTAMP pair of timer/gpuTimer.cpp:33-69) */
enum { VT_K_GENERATE = 0, VT_K_TRACE = 1, VT_K_OTHER = 2, VT_K_SHADE = 3, VT_K_ACCUMULATE = 4, VT_K_COUNT = 5 };
typedef struct vt_kernel_times { float ms[VT_K_COUNT]; uint32_t launches[VT_K_COUNT]; } vt_kernel_times;
int vt_kernel_timing_enable(vt_ctx* ctx, int enable);
int vt_get_kernel_times(vt_ctx* ctx, vt_kernel_times* out);
int vt_counters_enable(vt_ctx* ctx, int enable);
int vt_get_counters(vt_ctx* ctx, vt_counters* out);
int vt_reset_counters(vt_ctx* ctx);

/* ---- voxelizer: GPUVoxelizer::voxelizeMesh, voxelize/gpuVoxelizer.cpp:46-72 (+ Mesh::draw, mesh/mesh.cpp:55-59) ----
 * Clears the grid to `empty`, then writes fill_offset into every voxel the THIN surface voxelization
 * touches (voxelize.gs:118-251). Replaces the volume of the context (resolution X,Y,Z). */
int vt_voxelize(vt_ctx* ctx, const float* xyz, size_t n_verts, const uint32_t* indices, size_t n_indices,
                const float model_transform[16], int X, int Y, int Z, int32_t fill_offset);
/* replace material offsets of solid voxels by rule(x,y,z): new-build extension used by BASELINE config 3. */
int vt_volume_assign_materials(vt_ctx* ctx, const int32_t* offsets_table, int n_table, int rule);
/* elapsed GPU milliseconds of the last vt_voxelize (clear + scatter + derive), cudaEvent-timed */
/* THICKNESS of voxelize.gs:15-19: THIN (default, what the reference compiles: adjacent voxels connected at least by vertices)
 * or FAT (the conservative variant in the same shader text: adjacent voxels share at least a face) */
enum { VT_VOXELIZE_THIN = 0, VT_VOXELIZE_FAT = 1 };
int vt_set_voxelize_thickness(vt_ctx* ctx, int thickness);
int vt_get_last_voxelize_ms(vt_ctx* ctx, float* ms);
/* the same call's device time including what the pipeline needs next: the material-id grid made valid (cleared to "empty", the
 * solid voxels filled) and the renderer's empty-space distance field */
int vt_get_last_voxelize_full_ms(vt_ctx* ctx, float* ms);

/* ---- services: renderer/services/*.cpp (1-vertex "compute" draws, service.cpp:10-20) ---- */
int vt_pick(vt_ctx* ctx, float px, float py);                /* RendererServiceSelectActiveVoxel -> selectVoxel.vs */
int vt_pick_focal(vt_ctx* ctx, float px, float py);          /* RendererServiceSetFocalDistance  -> focalDistance.vs */
int vt_add_voxel(vt_ctx* ctx, float motion_x, float motion_y); /* RendererServiceAddVoxel -> addVoxel.vs (motion already y-flipped, serviceAddVoxel.cpp:74) */
int vt_remove_voxel(vt_ctx* ctx);                            /* RendererServiceRemoveVoxel -> removeVoxel.vs */

/* ---- diagnostics -------------------------------------------------------------------- */
int vt_device_count(void);
const char* vt_version(void);
/* test hook: trace n rays (origin xyz, dir xyz interleaved, 6 floats each) through the current volume with
 * the DDA of dda.h:63-100; out_hit[4n] = (x,y,z, code) with code 1 voxel hit, 2 ground, 0 miss. */
int vt_debug_trace_rays(vt_ctx* ctx, const float* rays, size_t n, float* out_hit);
/* measurement hook: L2 read bandwidth in GB/s (16-byte ld.global.cg over a `bytes`-sized buffer that fits in L2, `reps`
 * sweeps, all SMs) -- the denominator for the L2 roofline the north star asks for; MEASURED_PEAKS.json only has HBM. */
int vt_measure_l2_bandwidth(vt_ctx* ctx, size_t bytes, int reps, float* gb_per_s);
/* test hook: for each i, d[i] <- fl(d[i] + e[i]) while d[i] <= tau[i], at most nmax[i] times; k_out = additions done.
 * literal = 1 runs the plain loop, 0 the closed form used by the empty-space skip (advance_until). Operands must be > 0. */
int vt_debug_advance(vt_ctx* ctx, const float* d, const float* e, const float* tau, const int32_t* nmax, size_t n,
                     float* d_out, int32_t* k_out, int literal);

/* test hook: the 3-instruction division by a compile-time constant (csrc/vt_math.cuh, gdiv_by) against div.rn for ALL 2^32
 * numerators on the device; which = 0: PI, 1: 2 PI (the constants of shaders/shared/constants.h:1-3 the path divides by);
 * which = 2: the direction clamp of dda.h:29 as one select (gclamp_dir) against the literal mix / step, for all 2^32 inputs.
 * mismatches = how many numerators give different bits (NaN == NaN), first_bad = the smallest such bit pattern. */
int vt_debug_div_const(vt_ctx* ctx, int which, uint64_t* mismatches, uint32_t* first_bad);

/* ---- render groups: one frame sharded over several GPUs (new-build surface; the reference renders on a single GL context,
 * renderer/renderer.cpp:556-645). Paths are independent and the scene is replicated (every context of the group receives
 * the same uploads through the vt_* calls above), so rendering needs no collective; the group owns the partition
 * (vt_set_partition), the combination of the per-rank accumulators and the replication of edits.
 *   VT_PART_TILES    64x64 tiles dealt round-robin; the root gathers W*H*16/world bytes per rank; bit-identical to one GPU.
 *   VT_PART_SAMPLES  rank r renders sampleCount = p*world + r and keeps a float4 SUM; SUM-reduce to the root + one division
 *                    (fp summation order differs from the running average: equal within 1e-5 relative).
 * Exchange: NCCL bound at run time (ncclCommInitAll in one process, ncclCommInitRank across processes), or this library's own
 * kernel over peer memory (one process only; the only choice when two contexts share a device). The exchange runs on side
 * streams from a snapshot of the accumulators: rendering continues while it is in flight. */
typedef struct vt_group vt_group;
enum { VT_EXCHANGE_NCCL = 0, VT_EXCHANGE_PEER = 1 };
/* one process: creates (and owns) one context per listed device / adopts existing contexts (rank = position) */
int vt_group_create(int n, const int* devices, int mode, vt_group** out);
int vt_group_adopt(int n, vt_ctx* const* contexts, int mode, vt_group** out);
/* one process per GPU: rank 0 obtains the 128-byte NCCL id, the launcher hands it to every rank, every rank joins */
int vt_group_unique_id(void* out128);
int vt_group_join(vt_ctx* ctx, const void* id128, int rank, int world, int mode, vt_group** out);
void vt_group_destroy(vt_group* group);
const char* vt_group_last_error(const vt_group* group);
int vt_group_size(const vt_group* group);                       /* world size */
int vt_group_local_size(const vt_group* group);                 /* contexts held by this process */
vt_ctx* vt_group_context(vt_group* group, int local_index);     /* for the scene / camera / settings uploads */
int vt_group_rank(const vt_group* group, int local_index);
int vt_group_set_exchange(vt_group* group, int exchange);
int vt_group_get_exchange(const vt_group* group);
int vt_nccl_version(void);                                      /* 0 when libnccl.so.2 could not be loaded */
/* Renderer::render on every local context (each renders its share); asynchronous like vt_render */
int vt_group_render(vt_group* group, int first_sample, int n_passes);
int vt_group_reset_accumulation(vt_group* group);
int vt_group_sync(vt_group* group);
/* combination of the accumulators: begin returns at once (snapshot + exchange on side streams), end waits and, in the process
 * that holds rank 0, copies the finished W*H RGBA float32 frame to rgba_out (may be NULL); read_average = begin + end */
int vt_group_begin_combine(vt_group* group);
int vt_group_wait_combine(vt_group* group);                     /* stream-level: the contexts' streams wait for the exchange, the host does not */
int vt_group_end_combine(vt_group* group, float* rgba_out);
int vt_group_read_average(vt_group* group, float* rgba_out);
void* vt_group_result_device_ptr(vt_group* group);
int vt_group_last_exchange_ms(vt_group* group, float* ms);      /* device time of the last exchange on the root's side stream */
size_t vt_group_exchange_bytes(const vt_group* group);          /* bytes the root receives per combination */
/* edits on every replica (renderer/actions.cpp:20-52; volume edits reset the accumulation) */
int vt_group_pick(vt_group* group, float px, float py);
int vt_group_pick_focal(vt_group* group, float px, float py);
int vt_group_add_voxel(vt_group* group, float motion_x, float motion_y);
int vt_group_remove_voxel(vt_group* group);
/* up to 256 bytes from the process holding rank `root` to all ranks (the 32-byte action record of an edit issued on one rank) */
int vt_group_broadcast(vt_group* group, void* host_buf, size_t bytes, int root);

#ifdef __cplusplus
}
#endif
#endif /* VOXELTOY_B200_H */
