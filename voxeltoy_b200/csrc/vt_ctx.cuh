// vt_ctx.cuh -- the context object behind the C ABI (include/voxeltoy_b200.h), shared by the translation units that
// implement it (vt_api.cu: scene, frame, render, voxelizer, services; vt_group.cu: multi-GPU groups). Internal.
#pragma once
#include "../../include/voxeltoy_b200.h"
#include "vt_wavefront.cuh"

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

using namespace vt;

struct vt_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    vt_log_fn log_fn = nullptr;
    void* log_user = nullptr;

    // volume
    int X = 0, Y = 0, Z = 0, BX = 0, BY = 0, BZ = 0;     // voxel and brick counts; the brick array is padded by one brick per side
    int PBX = 0, PBY = 0, PBZ = 0;                         // padded brick counts (strides)
    // empty-space distance field over 8^3 cells (two buffers: the relaxation ping-pongs), see dda_skip
    unsigned char* d_dist[2] = {nullptr, nullptr}; int CX = 0, CY = 0, CZ = 0; int dist_cur = 0; bool dist_valid = false;   // dist_valid false: rebuilt before the next render that uses it
    unsigned char* d_dist4[2] = {nullptr, nullptr};        // near field over 4^3 bricks (ping-pong), capped at 3
    unsigned char* d_skip = nullptr;                       // what the DDA reads: one byte per brick, both levels (Volume::skip)
    int skip_mode = 1;                                     // 0 off, 1 auto (volumes with every side >= 64 voxels), 2 always
    // material ids: one byte per voxel (0xff = empty; two bytes, 0xffff = empty, when the volume has more than 255 distinct
    // material records) + the table id -> offset; the reference's R32I offsets exist only at the boundary (vt_volume_upload /
    // vt_read_volume). The record at offset 0 always has an id (`zero_id`): it is the material of the ground and of
    // voxels added next to an empty selection (addVoxel.vs:37-40).
    void* d_ids = nullptr; int id_bytes = 1; size_t ids_capacity = 0;
    int32_t* d_id_offset = nullptr; std::vector<int32_t> h_id_offset; int zero_id = 0;
    unsigned long long* d_bricks_alloc = nullptr;         // padded array
    unsigned long long* d_bricks = nullptr;               // brick (0,0,0): d_bricks_alloc + 1 + PBX + PBX*PBY
    unsigned long long* d_bricks_empty = nullptr;         // template of the empty grid (sentinel shell only): clearing is one D2D copy
    float bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0}, vsize[3] = {0, 0, 0};
    // scene arrays
    float* d_materials = nullptr; size_t n_materials = 0;
    int32_t* d_emissive = nullptr; size_t n_emissive = 0;
    float4* d_noise = nullptr; int noise_w = 0, noise_h = 0;
    float4* d_env = nullptr; int env_w = 0, env_h = 0;
    float* d_cdf_u = nullptr; int cdf_u_w = 0, cdf_u_h = 0;
    float* d_cdf_v = nullptr; int cdf_v_n = 0;
    unsigned short* d_guide_v = nullptr; unsigned short* d_guide_u = nullptr; int guide_k = 0;
    float env_integral = 0.f;
    // frame
    vt_camera cam{};
    vt_settings st{};
    bool have_cam = false, have_settings = false;
    float4* d_accum = nullptr; size_t accum_pixels = 0;
    uchar4* d_display = nullptr;                          // RGBA8 display image (vt_read_display), allocated on first use
    int32_t* d_primary = nullptr; bool primary_enabled = false;
    int num_samples = 0;
    Shared* d_shared = nullptr;
    int* d_result = nullptr;
    // partition
    int part_mode = VT_PART_NONE, part_rank = 0, part_world = 1;
    // render kernel variant: 0 = one-thread-per-pixel megakernel, 2 = wavefront (vt_wavefront.cuh, default for the path tracer)
    int variant = 2;
    // wavefront state (variant 2): SoA path state + queues, sized for wf_capacity paths
    // Batches run on up to kWfLanes internal streams ("lanes"), each with its own state pool, so the ramp-down tail of one
    // batch's kernels is filled by the other's CTAs; accumulation stays on the caller's stream, in pass order.
    static constexpr int kWfLanes = 4;
    int wf_lanes = 1;      // default 1: kernels of one batch at a time (clean per-kernel timing); 2 overlaps batches, +2-4 %
    WfState wf[kWfLanes]{}; void* d_wf_pool[kWfLanes] = {nullptr, nullptr, nullptr, nullptr}; size_t wf_capacity[kWfLanes] = {0, 0, 0, 0};
    WfCounts* d_wf_counts[kWfLanes] = {nullptr, nullptr, nullptr, nullptr}; int wf_counts_cap[kWfLanes] = {0, 0, 0, 0};
    cudaStream_t wf_stream[kWfLanes] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t wf_fork = nullptr, wf_done[kWfLanes] = {nullptr, nullptr, nullptr, nullptr}, wf_acc[kWfLanes] = {nullptr, nullptr, nullptr, nullptr};
    size_t wf_max_paths = (size_t)128 << 20;   // paths in flight per batch: 324 B each -> <= 43.5 GB of the 180 GB (C2, round 1: 32 -> 64 -> 128 Mi = 2 542 -> 2 582 -> 2 607 Msamples/s)
    size_t wf_queue_slack = 0;                 // queue entries beyond the path count: the chunked reservations of wf_trace and wf_generate
    size_t wf_slack_alloc[kWfLanes] = {0, 0, 0, 0};   // the slack each lane's pool was allocated with
    int wf_shade_blocks[2] = {0, 0}, wf_trace_blocks[2] = {0, 0}, wf_sms = 0;
    // per-kernel device timing (vt_kernel_timing_enable): event pairs around every wavefront launch
    bool timing = false;
    struct Timed { int kind; cudaEvent_t a, b; };
    std::vector<Timed> timed; std::vector<cudaEvent_t> ev_pool;
    // counters
    Counters* d_counters = nullptr; bool count_enabled = false;
    uint64_t paths = 0, launches = 0;
    // voxelizer timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr; float last_voxelize_ms = 0.f, last_voxelize_full_ms = 0.f, last_env_build_ms = 0.f;
    bool voxelize_fat = false;                             // voxelize.gs:15-19 THICKNESS (the reference compiles THIN)
    float* d_mesh_xyz = nullptr; unsigned int* d_mesh_idx = nullptr; float* d_mesh_M = nullptr; size_t mesh_verts_cap = 0, mesh_idx_cap = 0;
};

static inline int fail(vt_ctx* c, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) { c->err = buf; if (c->log_fn) c->log_fn(buf, c->log_user); }
    return code;
}
#define VT_CUDA(c, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail((c), VT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define VT_REQ(c, cond, msg) do { if (!(cond)) return fail((c), VT_ERR_INVALID, "%s", (msg)); } while (0)
#define VT_BIND(c) VT_CUDA(c, cudaSetDevice((c)->device))

static inline int grid_for(size_t n, int block) { size_t g = (n + block - 1) / block; const size_t cap = 148 * 32; return (int)(g < 1 ? 1 : (g > cap ? cap : g)); }

