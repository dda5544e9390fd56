// vt_api.cu -- implementation of the C ABI declared in include/voxeltoy_b200.h.
// Host side of the device boundary: owns device memory, the stream, and launches the
// kernels of vt_kernels.cuh. No OpenGL, no CPU fallback: every compute entry point
// needs a live CUDA context and fails with VT_ERR_CUDA / VT_ERR_NO_DEVICE otherwise.
#include "vt_ctx.cuh"
#include "vt_env.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

using namespace vt;

// splits [0, n) over the host's threads
template <class F> static void parallel_ranges(size_t n, F f)
{
    unsigned nt = std::thread::hardware_concurrency();
    nt = std::max(1u, std::min(nt, 32u));
    if (n < ((size_t)1 << 22)) nt = 1;
    std::vector<std::thread> th;
    const size_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const size_t a = std::min(n, (size_t)t * per), b = std::min(n, a + per);
        if (t + 1 == nt) f(t, a, b); else th.emplace_back(f, t, a, b);
    }
    for (auto& x : th) x.join();
}

// buckets of the CDF guide tables (cdf_search_guided): a power of two, so that s * K and the bucket index are exact. 256 -> 1024:
// twice the entries of the longer CDF (512), so most buckets bracket the answer without a probe (wf_shade 76.2 -> 74.9 ms per C2 step)
#ifndef VT_GUIDE_K
#define VT_GUIDE_K 1024
#endif

extern "C" {

int vt_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
const char* vt_version(void) { return "voxeltoy_b200 0.1 (sm_100a)"; }

// glibc rand() (TYPE_3 additive feedback, default seed 1), renderer/renderer.cpp:741-744
static void default_noise(std::vector<float>& out, size_t n)
{
    std::vector<uint32_t> r(n + 344);
    int32_t s[34];
    s[0] = 1;
    for (int i = 1; i < 31; ++i) {
        const int64_t hi = s[i - 1] / 127773, lo = s[i - 1] % 127773;
        int64_t w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        s[i] = (int32_t)w;
    }
    for (int i = 0; i < 31; ++i) r[i] = (uint32_t)s[i];
    for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
    for (size_t i = 34; i < 344; ++i) r[i] = r[i - 31] + r[i - 3];
    out.resize(n);
    for (size_t i = 344; i < 344 + n; ++i) {
        r[i] = r[i - 31] + r[i - 3];
        out[i - 344] = (float)(int32_t)(r[i] >> 1) / (float)2147483647;   // (float)rand() / RAND_MAX
    }
}

int vt_create(int device, vt_ctx** out)
{
    if (!out) return VT_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return VT_ERR_NO_DEVICE;
    vt_ctx* c = new vt_ctx();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c; return VT_ERR_CUDA;
    }
    c->stream = c->own_stream;
    if (const char* e = getenv("VT_WF_MAX_PATHS")) { const long long v = atoll(e); if (v > 0) c->wf_max_paths = (size_t)v; }   // tuning knob
    if (const char* e = getenv("VT_EMPTY_SKIP")) { const int v = atoi(e); if (v >= 0 && v <= 2) c->skip_mode = v; }
    if (const char* e = getenv("VT_WF_LANES")) { const int v = atoi(e); if (v >= 1 && v <= vt_ctx::kWfLanes) c->wf_lanes = v; }
    if (const char* e = getenv("VT_KERNEL_VARIANT")) { const int v = atoi(e); if (v == 0 || v == 2) c->variant = v; }
    Shared sh;
    sh.focal_distance = 99999999.0f;                                  // renderer.cpp:712-720
    sh.sel_index[0] = sh.sel_index[1] = sh.sel_index[2] = sh.sel_index[3] = 0;   // :723-737
    sh.sel_normal[0] = 1.f; sh.sel_normal[1] = sh.sel_normal[2] = sh.sel_normal[3] = 0.f;
    if (cudaMalloc(&c->d_shared, sizeof(Shared)) != cudaSuccess || cudaMalloc(&c->d_result, 4 * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&c->d_counters, sizeof(Counters)) != cudaSuccess ||
        cudaMemcpy(c->d_shared, &sh, sizeof sh, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(c->d_counters, 0, sizeof(Counters)) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess || cudaEventCreate(&c->ev2) != cudaSuccess) {
        vt_destroy(c); return VT_ERR_CUDA;
    }
    // defaults of the Renderer constructor (renderer.cpp:41-65) + pathTracer.fs:25-26,40-41
    c->st.width = 512; c->st.height = 512; c->st.max_bounces = 1; c->st.integrator = VT_INTEGRATOR_PATHTRACER;
    c->st.bg_top[0] = 153.0f / 255 * 2; c->st.bg_top[1] = 187.0f / 255 * 2; c->st.bg_top[2] = 201.0f / 255 * 2;
    c->st.bg_bottom[0] = 77.0f / 255; c->st.bg_bottom[1] = 64.0f / 255; c->st.bg_bottom[2] = 50.0f / 255;
    c->st.wireframe_opacity = 0.f; c->st.wireframe_thickness = 0.01f;
    *out = c;
    // noise texture + 16^3 empty volume, as Renderer::initialize does (renderer.cpp:94-98)
    int rc = vt_noise_upload(c, nullptr, 1024, 1024);
    if (rc == VT_OK) rc = vt_volume_upload(c, nullptr, 16, 16, 16);
    if (rc != VT_OK) { vt_destroy(c); *out = nullptr; return rc; }
    return VT_OK;
}

void vt_destroy(vt_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_ids); cudaFree(c->d_id_offset); cudaFree(c->d_bricks_alloc); cudaFree(c->d_bricks_empty); cudaFree(c->d_dist[0]); cudaFree(c->d_dist[1]); cudaFree(c->d_dist4[0]); cudaFree(c->d_dist4[1]); cudaFree(c->d_skip);
    cudaFree(c->d_materials); cudaFree(c->d_emissive);
    cudaFree(c->d_noise); cudaFree(c->d_env); cudaFree(c->d_cdf_u); cudaFree(c->d_cdf_v); cudaFree(c->d_accum); cudaFree(c->d_display);
    cudaFree(c->d_guide_v); cudaFree(c->d_guide_u);
    for (int k = 0; k < vt_ctx::kWfLanes; ++k) {
        cudaFree(c->d_wf_pool[k]); cudaFree(c->d_wf_counts[k]);
        if (c->wf_stream[k]) cudaStreamDestroy(c->wf_stream[k]);
        if (c->wf_done[k]) cudaEventDestroy(c->wf_done[k]);
        if (c->wf_acc[k]) cudaEventDestroy(c->wf_acc[k]);
    }
    if (c->wf_fork) cudaEventDestroy(c->wf_fork);
    for (auto& t : c->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    cudaFree(c->d_primary); cudaFree(c->d_shared); cudaFree(c->d_result); cudaFree(c->d_counters);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    cudaFree(c->d_mesh_xyz); cudaFree(c->d_mesh_idx); cudaFree(c->d_mesh_M);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* vt_last_error(const vt_ctx* c) { return c ? c->err.c_str() : "null context"; }
int vt_set_logger(vt_ctx* c, vt_log_fn fn, void* user) { if (!c) return VT_ERR_INVALID; c->log_fn = fn; c->log_user = user; return VT_OK; }
int vt_set_stream(vt_ctx* c, void* s) { if (!c) return VT_ERR_INVALID; c->stream = s ? (cudaStream_t)s : c->own_stream; return VT_OK; }
int vt_sync(vt_ctx* c) { if (!c) return VT_ERR_INVALID; VT_BIND(c); VT_CUDA(c, cudaStreamSynchronize(c->stream)); return VT_OK; }

// ---- volume ------------------------------------------------------------------------------------
static void volume_bounds(vt_ctx* c)
{
    // renderer.cpp:845-850 and :926-929
    const int m = std::max(c->X, std::max(c->Y, c->Z));
    const float voxel = 1000.0f / (float)m;
    const int r[3] = { c->X, c->Y, c->Z };
    for (int i = 0; i < 3; ++i) {
        const float sz = voxel * (float)r[i];
        c->bmin[i] = -sz * 0.5f; c->bmax[i] = sz * 0.5f;
        c->vsize[i] = (c->bmax[i] - c->bmin[i]) / (float)r[i];
    }
}

static constexpr int kMaxIds = 65535;

// (re)allocates the id grid for `id_bytes` per voxel
static int alloc_ids(vt_ctx* c, int id_bytes)
{
    const size_t need = (size_t)c->X * c->Y * c->Z * (size_t)id_bytes;
    if (need > c->ids_capacity || !c->d_ids) {
        cudaFree(c->d_ids); c->d_ids = nullptr; c->ids_capacity = 0;
        VT_CUDA(c, cudaMalloc(&c->d_ids, need));
        c->ids_capacity = need;
    }
    c->id_bytes = id_bytes;
    if (!c->d_id_offset) VT_CUDA(c, cudaMalloc(&c->d_id_offset, sizeof(int32_t) * (kMaxIds + 1)));
    return VT_OK;
}
// installs the table id -> offset (host mirror + device copy); guarantees an id for offset 0
static int set_id_table(vt_ctx* c, std::vector<int32_t> table)
{
    int zero = -1;
    for (size_t i = 0; i < table.size(); ++i) if (table[i] == 0) { zero = (int)i; break; }
    if (zero < 0) { zero = (int)table.size(); table.push_back(0); }
    if ((int)table.size() > kMaxIds) return fail(c, VT_ERR_INVALID, "more than %d distinct material offsets in the volume", kMaxIds - 1);
    if (!c->d_id_offset) VT_CUDA(c, cudaMalloc(&c->d_id_offset, sizeof(int32_t) * (kMaxIds + 1)));
    c->h_id_offset = table; c->zero_id = zero;
    VT_CUDA(c, cudaMemcpyAsync(c->d_id_offset, c->h_id_offset.data(), sizeof(int32_t) * table.size(), cudaMemcpyHostToDevice, c->stream));
    VT_CUDA(c, cudaStreamSynchronize(c->stream));       // h_id_offset may be replaced before the copy would otherwise run
    return VT_OK;
}

static int alloc_volume(vt_ctx* c, int X, int Y, int Z)
{
    VT_REQ(c, X > 0 && Y > 0 && Z > 0 && X <= 2048 && Y <= 2048 && Z <= 2048, "volume resolution must be in [1, 2048]^3");
    if (X != c->X || Y != c->Y || Z != c->Z || !c->d_bricks_alloc) {
        // allocate into locals first: the context changes only when every allocation has succeeded
        const int BX = (X + 3) / 4, BY = (Y + 3) / 4, BZ = (Z + 3) / 4, PBX = BX + 2, PBY = BY + 2, PBZ = BZ + 2;
        const int CX = (X + 7) / 8, CY = (Y + 7) / 8, CZ = (Z + 7) / 8;
        const size_t npb = (size_t)PBX * PBY * PBZ, nc = (size_t)CX * CY * CZ;
        unsigned long long *bricks = nullptr, *empty = nullptr; unsigned char *d0 = nullptr, *d1 = nullptr, *f0 = nullptr, *f1 = nullptr, *sk = nullptr;
        const size_t nb = (size_t)BX * BY * BZ;
        cudaError_t e = cudaMalloc(&bricks, npb * 8);
        if (e == cudaSuccess) e = cudaMalloc(&empty, npb * 8);
        if (e == cudaSuccess) e = cudaMalloc(&d0, nc);
        if (e == cudaSuccess) e = cudaMalloc(&d1, nc);
        if (e == cudaSuccess) e = cudaMalloc(&f0, nb);
        if (e == cudaSuccess) e = cudaMalloc(&f1, nb);
        if (e == cudaSuccess) e = cudaMalloc(&sk, nb);
        if (e != cudaSuccess) {
            cudaFree(bricks); cudaFree(empty); cudaFree(d0); cudaFree(d1); cudaFree(f0); cudaFree(f1); cudaFree(sk); cudaGetLastError();
            return fail(c, VT_ERR_CUDA, "volume %dx%dx%d: %s", X, Y, Z, cudaGetErrorString(e));
        }
        VT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_ids); cudaFree(c->d_bricks_alloc); cudaFree(c->d_bricks_empty); cudaFree(c->d_dist[0]); cudaFree(c->d_dist[1]);
        cudaFree(c->d_dist4[0]); cudaFree(c->d_dist4[1]); cudaFree(c->d_skip);
        c->d_ids = nullptr; c->ids_capacity = 0;
        c->d_bricks_alloc = bricks; c->d_bricks_empty = empty; c->d_dist[0] = d0; c->d_dist[1] = d1; c->dist_valid = false;
        c->d_dist4[0] = f0; c->d_dist4[1] = f1; c->d_skip = sk;
        c->X = X; c->Y = Y; c->Z = Z;
        c->BX = BX; c->BY = BY; c->BZ = BZ; c->PBX = PBX; c->PBY = PBY; c->PBZ = PBZ; c->CX = CX; c->CY = CY; c->CZ = CZ;
        c->d_bricks = c->d_bricks_alloc + (1 + (size_t)PBX + (size_t)PBX * PBY);
        // the empty template: zero inside the volume, the sentinel shell set (see dda_step); built once per resolution
        VT_CUDA(c, cudaMemsetAsync(c->d_bricks_empty, 0, npb * 8, c->stream));
        vt_sentinel_kernel<<<grid_for(npb, 256), 256, 0, c->stream>>>(c->d_bricks_empty, X, Y, Z, PBX, PBY, PBZ);
        c->launches += 1;
        VT_CUDA(c, cudaGetLastError());
    }
    volume_bounds(c);
    return VT_OK;
}

// empties the occupancy grid: zero everywhere inside the volume, the sentinel shell (every voxel of the padded brick array
// that lies outside the volume) set -- see dda_step
static int clear_occupancy(vt_ctx* c)
{
    const size_t npb = (size_t)c->PBX * c->PBY * c->PBZ;
    VT_CUDA(c, cudaMemcpyAsync(c->d_bricks_alloc, c->d_bricks_empty, npb * 8, cudaMemcpyDeviceToDevice, c->stream));
    return VT_OK;
}

// (re)builds the empty-space skip field from the bricks: the far field over 8^3 cells (cap 16), the near field over 4^3 bricks
// (cap 3) and the byte per brick the DDA reads. Cheap (one byte per 64 voxels): run after every change that can make an empty
// cell solid (upload, voxelize, add voxel). Removing voxels leaves it valid (lower bounds).
#ifndef VT_DIST_CAP
#define VT_DIST_CAP 16
#endif
static constexpr int kDistCap = VT_DIST_CAP, kDist4Cap = 3;      // kDistCap <= 63 (6 bits of the skip byte)
static int rebuild_dist(vt_ctx* c)
{
    const size_t nc = (size_t)c->CX * c->CY * c->CZ, nb = (size_t)c->BX * c->BY * c->BZ;
    vt_dist_init_kernel<<<grid_for(nc, 256), 256, 0, c->stream>>>(c->d_bricks, c->d_dist[0], c->X, c->Y, c->Z, c->PBX, c->PBX * c->PBY,
                                                                 c->CX, c->CY, c->CZ, kDistCap);
    int cur = 0;
    for (int axis = 0; axis < 3; ++axis, cur ^= 1)
        vt_dist_pass_kernel<<<grid_for(nc, 256), 256, 0, c->stream>>>(c->d_dist[cur], c->d_dist[cur ^ 1], c->CX, c->CY, c->CZ, axis, kDistCap);
    vt_dist4_init_kernel<<<grid_for(nb, 256), 256, 0, c->stream>>>(c->d_bricks, c->d_dist4[0], c->X, c->Y, c->Z, c->PBX, c->PBX * c->PBY,
                                                                  c->BX, c->BY, c->BZ, kDist4Cap);
    int cur4 = 0;
    for (int axis = 0; axis < 3; ++axis, cur4 ^= 1)
        vt_dist_pass_kernel<<<grid_for(nb, 256), 256, 0, c->stream>>>(c->d_dist4[cur4], c->d_dist4[cur4 ^ 1], c->BX, c->BY, c->BZ, axis, kDist4Cap);
    vt_skip_combine_kernel<<<grid_for(nb, 256), 256, 0, c->stream>>>(c->d_dist[cur], c->d_dist4[cur4], c->d_skip, c->BX, c->BY, c->BZ, c->CX, c->CY);
    c->dist_cur = cur; c->dist_valid = true;
    c->launches += 9;
    VT_CUDA(c, cudaGetLastError());
    return VT_OK;
}

static int rebuild_occupancy(vt_ctx* c)
{
    int rc = clear_occupancy(c);
    if (rc != VT_OK) return rc;
    const size_t rows = (size_t)c->BX * c->Y * c->Z;
    vt_build_bricks_kernel<<<grid_for(rows, 256), 256, 0, c->stream>>>(c->d_ids, c->id_bytes, c->d_bricks, c->X, c->Y, c->Z, c->BX, c->PBX, c->PBX * c->PBY);
    c->launches += 1;
    c->dist_valid = false;                             // the renderer's distance field is rebuilt on its next use
    VT_CUDA(c, cudaGetLastError());
    return VT_OK;
}

// R32I offsets (renderer.cpp:863-872: x fastest, -1 empty) -> id table + id grid on the device
int vt_volume_upload(vt_ctx* c, const int32_t* mat, int X, int Y, int Z)
{
    if (!c) return VT_ERR_INVALID;
    VT_BIND(c);
    int rc = alloc_volume(c, X, Y, Z);
    if (rc != VT_OK) return rc;
    const size_t n = (size_t)X * Y * Z;
    if (!mat) {
        rc = alloc_ids(c, 1); if (rc != VT_OK) return rc;
        rc = set_id_table(c, std::vector<int32_t>()); if (rc != VT_OK) return rc;
        VT_CUDA(c, cudaMemsetAsync(c->d_ids, 0xff, n, c->stream));
    } else {
        // pass 1: the distinct offsets (every negative value means empty)
        std::vector<std::vector<int32_t>> found(64);
        parallel_ranges(n, [&](unsigned t, size_t a, size_t b) {
            std::vector<int32_t>& v = found[t];
            int32_t last = -1;
            for (size_t i = a; i < b; ++i) {
                const int32_t o = mat[i];
                if (o < 0 || o == last) continue;
                last = o;
                if (std::find(v.begin(), v.end(), o) == v.end()) { v.push_back(o); if (v.size() > (size_t)kMaxIds) return; }
            }
        });
        std::vector<int32_t> table;
        for (auto& v : found) table.insert(table.end(), v.begin(), v.end());
        std::sort(table.begin(), table.end());
        table.erase(std::unique(table.begin(), table.end()), table.end());
        const size_t n_sorted = table.size();
        rc = set_id_table(c, table); if (rc != VT_OK) return rc;
        const int id_bytes = c->h_id_offset.size() <= 255 ? 1 : 2;
        rc = alloc_ids(c, id_bytes); if (rc != VT_OK) return rc;
        // pass 2: offsets -> ids (binary search in the sorted part of the table; a one-entry cache catches the runs)
        const std::vector<int32_t>& T = c->h_id_offset;
        std::vector<unsigned char> ids(n * (size_t)id_bytes);
        parallel_ranges(n, [&](unsigned, size_t a, size_t b) {
            int32_t last = -1; int last_id = -1;
            for (size_t i = a; i < b; ++i) {
                const int32_t o = mat[i];
                int id = -1;
                if (o >= 0) {
                    if (o == last) id = last_id;
                    else { id = (int)(std::lower_bound(T.begin(), T.begin() + n_sorted, o) - T.begin()); last = o; last_id = id; }
                }
                if (id_bytes == 1) ids[i] = (unsigned char)(id < 0 ? 0xff : id);
                else reinterpret_cast<uint16_t*>(ids.data())[i] = (uint16_t)(id < 0 ? 0xffff : id);
            }
        });
        VT_CUDA(c, cudaMemcpyAsync(c->d_ids, ids.data(), ids.size(), cudaMemcpyHostToDevice, c->stream));
        VT_CUDA(c, cudaStreamSynchronize(c->stream));   // `ids` goes out of scope
    }
    rc = rebuild_occupancy(c);
    if (rc != VT_OK) return rc;
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    return VT_OK;
}

static int upload_array(vt_ctx* c, void** dptr, const void* src, size_t bytes)
{
    cudaFree(*dptr); *dptr = nullptr;
    if (bytes == 0) return VT_OK;
    VT_CUDA(c, cudaMalloc(dptr, bytes));
    VT_CUDA(c, cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice));
    return VT_OK;
}

int vt_materials_upload(vt_ctx* c, const float* data, size_t n)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, data || n == 0, "null material data");
    VT_BIND(c);
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    c->n_materials = n;
    return upload_array(c, (void**)&c->d_materials, data, n * sizeof(float));
}

int vt_material_update(vt_ctx* c, uint32_t offset, const float* v, int n)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, v && n > 0 && (size_t)offset + (size_t)n <= c->n_materials, "material update out of range");
    VT_BIND(c);
    VT_CUDA(c, cudaMemcpyAsync(c->d_materials + offset, v, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    return VT_OK;
}

int vt_emissive_upload(vt_ctx* c, const int32_t* idx, size_t n)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, idx || n == 0, "null emissive list");
    VT_BIND(c);
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    c->n_emissive = n;
    return upload_array(c, (void**)&c->d_emissive, idx, n * sizeof(int32_t));
}

int vt_read_volume(vt_ctx* c, int32_t* out)            // the R32I view of the id grid (renderer.cpp:863-872)
{
    if (!c || !out) return VT_ERR_INVALID;
    VT_BIND(c);
    const size_t n = (size_t)c->X * c->Y * c->Z;
    std::vector<unsigned char> ids(n * (size_t)c->id_bytes);
    VT_CUDA(c, cudaMemcpyAsync(ids.data(), c->d_ids, ids.size(), cudaMemcpyDeviceToHost, c->stream));
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    const std::vector<int32_t>& T = c->h_id_offset;
    const int id_bytes = c->id_bytes;
    parallel_ranges(n, [&](unsigned, size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
            if (id_bytes == 1) { const unsigned v = ids[i]; out[i] = (v == 0xffu || v >= T.size()) ? -1 : T[v]; }
            else { const unsigned v = reinterpret_cast<const uint16_t*>(ids.data())[i]; out[i] = (v == 0xffffu || v >= T.size()) ? -1 : T[v]; }
        }
    });
    return VT_OK;
}

int vt_read_materials(vt_ctx* c, float* out, size_t n)
{
    if (!c || !out) return VT_ERR_INVALID;
    VT_REQ(c, n <= c->n_materials, "read past the material array");
    VT_BIND(c);
    VT_CUDA(c, cudaMemcpyAsync(out, c->d_materials, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    return VT_OK;
}

int vt_get_volume_info(vt_ctx* c, int32_t res[3], float bmin[3], float bmax[3], float vs[3])
{
    if (!c) return VT_ERR_INVALID;
    if (res) { res[0] = c->X; res[1] = c->Y; res[2] = c->Z; }
    for (int i = 0; i < 3; ++i) { if (bmin) bmin[i] = c->bmin[i]; if (bmax) bmax[i] = c->bmax[i]; if (vs) vs[i] = c->vsize[i]; }
    return VT_OK;
}

int vt_noise_upload(vt_ctx* c, const float* rgba, int w, int h)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, w > 0 && h > 0 && (size_t)w * (size_t)h <= ((size_t)1 << 29), "bad noise size (the table may hold at most 2^29 texels)");
    VT_BIND(c);
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<float> gen;
    if (!rgba) { default_noise(gen, (size_t)w * h * 4); rgba = gen.data(); }
    c->noise_w = w; c->noise_h = h;
    return upload_array(c, (void**)&c->d_noise, rgba, sizeof(float) * 4 * (size_t)w * h);
}

// guide table of one CDF search (csrc/vt_device.cuh cdf_search_guided): guide[j] = max({0} U {m in [1, n-2] : cdf[m] <= j/K}),
// guide[K] = n - 2. Returns false when cdf[1..n-2] is not sorted (then the device keeps the literal bisection).
static bool build_guide(const float* cdf, int n, int K, unsigned short* guide)
{
    if (n < 2 || n - 2 > 65535) return false;
    for (int m = 2; m <= n - 2; ++m) if (!(cdf[m] >= cdf[m - 1])) return false;
    if (n - 2 >= 1 && cdf[1] != cdf[1]) return false;
    int r = 0;
    for (int j = 0; j < K; ++j) {
        const float t = (float)j / (float)K;
        while (r + 1 <= n - 2 && cdf[r + 1] <= t) ++r;
        guide[j] = (unsigned short)r;
    }
    guide[K] = (unsigned short)std::max(0, n - 2);
    return true;
}

int vt_env_upload(vt_ctx* c, const float* rgb, int w, int h, const float* cdf_u, int cuw, int cuh,
                  const float* cdf_v, int cvn, float integral)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, rgb && cdf_u && cdf_v && w > 0 && h > 0 && cuw > 1 && cuh > 0 && cvn > 1, "bad environment arrays");
    VT_BIND(c);
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<float> rgba((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; ++i) {                        // GL_RGB -> GL_RGBA32F, alpha 1 (renderer.cpp:994-1002)
        rgba[4 * i] = rgb[3 * i]; rgba[4 * i + 1] = rgb[3 * i + 1]; rgba[4 * i + 2] = rgb[3 * i + 2]; rgba[4 * i + 3] = 1.0f;
    }
    int rc = upload_array(c, (void**)&c->d_env, rgba.data(), rgba.size() * sizeof(float));
    if (rc == VT_OK) rc = upload_array(c, (void**)&c->d_cdf_u, cdf_u, sizeof(float) * (size_t)cuw * cuh);
    if (rc == VT_OK) rc = upload_array(c, (void**)&c->d_cdf_v, cdf_v, sizeof(float) * (size_t)cvn);
    if (rc != VT_OK) return rc;
    c->env_w = w; c->env_h = h; c->cdf_u_w = cuw; c->cdf_u_h = cuh; c->cdf_v_n = cvn; c->env_integral = integral;
    // guide tables for the two CDF searches (only when every row the search can reach exists and is sorted)
    cudaFree(c->d_guide_v); cudaFree(c->d_guide_u); c->d_guide_v = nullptr; c->d_guide_u = nullptr; c->guide_k = 0;
    const int K = VT_GUIDE_K, rows = cvn - 1;
    if (rows >= 1 && cuh >= rows) {
        std::vector<unsigned short> gv(K + 1), gu((size_t)rows * (K + 1));
        bool ok = build_guide(cdf_v, cvn, K, gv.data());
        for (int r = 0; ok && r < rows; ++r) ok = build_guide(cdf_u + (size_t)r * cuw, cuw, K, gu.data() + (size_t)r * (K + 1));
        if (ok) {
            rc = upload_array(c, (void**)&c->d_guide_v, gv.data(), gv.size() * sizeof(unsigned short));
            if (rc == VT_OK) rc = upload_array(c, (void**)&c->d_guide_u, gu.data(), gu.size() * sizeof(unsigned short));
            if (rc != VT_OK) return rc;
            c->guide_k = K;
        }
    }
    return VT_OK;
}

int vt_env_clear(vt_ctx* c)
{
    if (!c) return VT_ERR_INVALID;
    VT_BIND(c);
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_env); cudaFree(c->d_cdf_u); cudaFree(c->d_cdf_v); cudaFree(c->d_guide_v); cudaFree(c->d_guide_u);
    c->d_guide_v = nullptr; c->d_guide_u = nullptr; c->guide_k = 0;
    c->d_env = nullptr; c->d_cdf_u = nullptr; c->d_cdf_v = nullptr;
    c->env_w = c->env_h = c->cdf_u_w = c->cdf_u_h = c->cdf_v_n = 0; c->env_integral = 0.f;
    return VT_OK;
}


// Renderer::loadBackgroundImage's processing half on the device (kernels in vt_env.cuh): the
// caller hands over the decoded RGB float image only. Same results, bit for bit, as calculateCDF of
// voxeltoy_b200/host/image.cpp (= image.cpp:68-389 of the reference) followed by vt_env_upload.
int vt_env_build(vt_ctx* c, const float* rgb, int w, int h)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, rgb && w > 0 && h > 0 && w <= 16384 && h <= 16384, "bad environment image");
    int nw = w, nh = h;
    const int m = std::max(w, h);
    if (m > VT_ENV_MAX_CDF_SIZE) {                                     // image.cpp:309-321 (area-weighted box filter, vt_env.cuh)
        nw = (int)((float)w / m * VT_ENV_MAX_CDF_SIZE); nh = (int)((float)h / m * VT_ENV_MAX_CDF_SIZE);
        VT_REQ(c, nw > 0 && nh > 0, "environment image too elongated for the CDF size limit");
    }
    VT_BIND(c);
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    vt_env_clear(c);
    const size_t npx = (size_t)w * h, ncdf = (size_t)nw * nh;
    std::vector<float> sin_row(nh);
    for (int y = 0; y < nh; ++y) sin_row[y] = (float)sin(M_PI * ((float)y + 0.5f) / (float)nh);     // image.cpp:366
    float *d_rgb = nullptr, *d_a = nullptr, *d_b = nullptr, *d_sin = nullptr, *d_fv = nullptr, *d_out = nullptr;
    int* d_sorted = nullptr;
    const int K = VT_GUIDE_K;
    auto cleanup = [&]() { cudaFree(d_rgb); cudaFree(d_a); cudaFree(d_b); cudaFree(d_sin); cudaFree(d_fv); cudaFree(d_out); cudaFree(d_sorted); };
#define VT_ENV_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); vt_env_clear(c); \
        return fail(c, VT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)
    VT_ENV_CUDA(cudaMalloc(&d_rgb, npx * 3 * sizeof(float)));
    VT_ENV_CUDA(cudaMalloc(&d_a, ncdf * sizeof(float)));
    VT_ENV_CUDA(cudaMalloc(&d_b, ncdf * sizeof(float)));
    VT_ENV_CUDA(cudaMalloc(&d_sin, nh * sizeof(float)));
    VT_ENV_CUDA(cudaMalloc(&d_fv, nh * sizeof(float)));
    VT_ENV_CUDA(cudaMalloc(&d_out, sizeof(float)));
    VT_ENV_CUDA(cudaMalloc(&d_sorted, sizeof(int)));
    VT_ENV_CUDA(cudaMalloc(&c->d_env, npx * sizeof(float4)));
    VT_ENV_CUDA(cudaMalloc(&c->d_cdf_u, (size_t)(nw + 1) * nh * sizeof(float)));
    VT_ENV_CUDA(cudaMalloc(&c->d_cdf_v, (size_t)(nh + 1) * sizeof(float)));
    VT_ENV_CUDA(cudaMalloc(&c->d_guide_v, (size_t)(K + 1) * sizeof(unsigned short)));
    VT_ENV_CUDA(cudaMalloc(&c->d_guide_u, (size_t)nh * (K + 1) * sizeof(unsigned short)));
    cudaStream_t st = c->stream;
    const int one = 1;
    VT_ENV_CUDA(cudaMemcpyAsync(d_rgb, rgb, npx * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    VT_ENV_CUDA(cudaMemcpyAsync(d_sin, sin_row.data(), nh * sizeof(float), cudaMemcpyHostToDevice, st));
    VT_ENV_CUDA(cudaMemcpyAsync(d_sorted, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    VT_ENV_CUDA(cudaEventRecord(c->ev0, st));
    const dim3 g2((nw + 127) / 128, nh);
    vt_env_rgba_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(d_rgb, c->d_env, npx);
    vt_env_luminance_kernel<<<g2, 128, 0, st>>>(d_rgb, w, h, nw, nh, d_a);
    vt_env_blur_kernel<<<g2, 128, 0, st>>>(d_a, d_b, nw, nh, 0);
    vt_env_blur_kernel<<<g2, 128, 0, st>>>(d_b, d_a, nw, nh, 1);
    vt_env_function_kernel<<<g2, 128, 0, st>>>(d_a, d_sin, nw, nh, d_b);
    vt_env_sum_kernel<<<1, 256, 0, st>>>(d_b, ncdf, d_out);
    vt_env_cdf_rows_kernel<<<(nh + 3) / 4, 128, 0, st>>>(d_b, nw, nh, c->d_cdf_u, d_fv);
    vt_env_cdf_v_kernel<<<1, 32, 0, st>>>(d_fv, nh, c->d_cdf_v);
    vt_env_guide_kernel<<<1, 128, 0, st>>>(c->d_cdf_v, nh + 1, 0, 1, K, c->d_guide_v, d_sorted);
    vt_env_guide_kernel<<<nh, 128, 0, st>>>(c->d_cdf_u, nw + 1, nw + 1, nh, K, c->d_guide_u, d_sorted);
    VT_ENV_CUDA(cudaGetLastError());
    VT_ENV_CUDA(cudaEventRecord(c->ev1, st));
    float sum = 0.f; int sorted = 0;
    VT_ENV_CUDA(cudaMemcpyAsync(&sum, d_out, sizeof(float), cudaMemcpyDeviceToHost, st));
    VT_ENV_CUDA(cudaMemcpyAsync(&sorted, d_sorted, sizeof(int), cudaMemcpyDeviceToHost, st));
    VT_ENV_CUDA(cudaStreamSynchronize(st));
    VT_ENV_CUDA(cudaEventElapsedTime(&c->last_env_build_ms, c->ev0, c->ev1));
#undef VT_ENV_CUDA
    cleanup();
    float integral = sum / ((float)nw * (float)nh);                      // image.cpp:381-386 (float * double constants)
    integral = (float)((double)integral * (2.0f * M_PI * M_PI));
    c->env_w = w; c->env_h = h; c->cdf_u_w = nw + 1; c->cdf_u_h = nh; c->cdf_v_n = nh + 1; c->env_integral = integral;
    if (sorted && nw - 1 <= 65535 && nh - 1 <= 65535) c->guide_k = K;
    else { cudaFree(c->d_guide_v); cudaFree(c->d_guide_u); c->d_guide_v = nullptr; c->d_guide_u = nullptr; c->guide_k = 0; }
    return VT_OK;
}

int vt_get_env_info(vt_ctx* c, int32_t* dims /* w, h, cdf_u_w, cdf_u_h, cdf_v_n, guided */, float* integral, float* build_ms)
{
    if (!c) return VT_ERR_INVALID;
    if (build_ms) *build_ms = c->last_env_build_ms;
    if (dims) { dims[0] = c->env_w; dims[1] = c->env_h; dims[2] = c->cdf_u_w; dims[3] = c->cdf_u_h; dims[4] = c->cdf_v_n; dims[5] = c->guide_k > 0; }
    if (integral) *integral = c->env_integral;
    return VT_OK;
}

int vt_read_env_cdf(vt_ctx* c, float* cdf_u, float* cdf_v)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, c->d_cdf_u && c->d_cdf_v, "no environment map loaded");
    VT_BIND(c);
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    if (cdf_u) VT_CUDA(c, cudaMemcpy(cdf_u, c->d_cdf_u, sizeof(float) * (size_t)c->cdf_u_w * c->cdf_u_h, cudaMemcpyDeviceToHost));
    if (cdf_v) VT_CUDA(c, cudaMemcpy(cdf_v, c->d_cdf_v, sizeof(float) * (size_t)c->cdf_v_n, cudaMemcpyDeviceToHost));
    return VT_OK;
}

// ---- per-frame state ----------------------------------------------------------------------------
int vt_set_camera(vt_ctx* c, const vt_camera* cam) { if (!c || !cam) return VT_ERR_INVALID; c->cam = *cam; c->have_cam = true; return VT_OK; }

int vt_set_settings(vt_ctx* c, const vt_settings* st)
{
    if (!c || !st) return VT_ERR_INVALID;
    VT_REQ(c, st->width > 0 && st->height > 0 && st->width <= 16384 && st->height <= 16384, "bad frame size");
    VT_REQ(c, st->max_bounces >= 0, "negative bounce count");
    VT_BIND(c);
    const size_t px = (size_t)st->width * st->height;
    if (px != c->accum_pixels || st->width != c->st.width || !c->d_accum) {
        VT_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_accum); cudaFree(c->d_primary); cudaFree(c->d_display); c->d_accum = nullptr; c->d_primary = nullptr; c->d_display = nullptr;
        VT_CUDA(c, cudaMalloc(&c->d_accum, px * sizeof(float4)));
        VT_CUDA(c, cudaMemsetAsync(c->d_accum, 0, px * sizeof(float4), c->stream));
        c->accum_pixels = px;
        c->num_samples = 0;
    }
    c->st = *st;
    c->have_settings = true;
    return VT_OK;
}

static int shared_rw(vt_ctx* c, Shared* host, bool write)
{
    VT_BIND(c);
    if (write) VT_CUDA(c, cudaMemcpyAsync(c->d_shared, host, sizeof(Shared), cudaMemcpyHostToDevice, c->stream));
    else VT_CUDA(c, cudaMemcpyAsync(host, c->d_shared, sizeof(Shared), cudaMemcpyDeviceToHost, c->stream));
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    return VT_OK;
}
int vt_set_focal_distance(vt_ctx* c, float d)
{
    if (!c) return VT_ERR_INVALID;
    Shared s; int rc = shared_rw(c, &s, false); if (rc) return rc;
    s.focal_distance = d; return shared_rw(c, &s, true);
}
int vt_get_focal_distance(vt_ctx* c, float* d)
{
    if (!c || !d) return VT_ERR_INVALID;
    Shared s; int rc = shared_rw(c, &s, false); if (rc) return rc;
    *d = s.focal_distance; return VT_OK;
}
int vt_set_selection(vt_ctx* c, const int32_t index[4], const float normal[4])
{
    if (!c || !index) return VT_ERR_INVALID;
    Shared s; int rc = shared_rw(c, &s, false); if (rc) return rc;
    for (int i = 0; i < 4; ++i) { s.sel_index[i] = index[i]; if (normal) s.sel_normal[i] = normal[i]; }
    return shared_rw(c, &s, true);
}
int vt_get_selection(vt_ctx* c, int32_t index[4], float normal[4])
{
    if (!c) return VT_ERR_INVALID;
    Shared s; int rc = shared_rw(c, &s, false); if (rc) return rc;
    for (int i = 0; i < 4; ++i) { if (index) index[i] = s.sel_index[i]; if (normal) normal[i] = s.sel_normal[i]; }
    return VT_OK;
}

// ---- launch parameter blocks ----------------------------------------------------------------------
static bool skip_wanted(const vt_ctx* c)
{
    return c->skip_mode == 2 || (c->skip_mode == 1 && std::min(c->X, std::min(c->Y, c->Z)) >= 64);
}

static Volume make_volume(const vt_ctx* c)
{
    Volume V;
    V.ids8 = c->id_bytes == 1 ? (const uint8_t*)c->d_ids : nullptr; V.ids16 = c->id_bytes == 2 ? (const uint16_t*)c->d_ids : nullptr;
    V.id_offset = c->d_id_offset; V.bricks = c->d_bricks; V.bricks_top = c->d_bricks_alloc;
    V.X = c->X; V.Y = c->Y; V.Z = c->Z;
    V.BX = c->PBX; V.BXY = c->PBX * c->PBY;               // strides of the padded brick array
    V.nBX = -V.BX; V.nBXY = -V.BXY;
    const bool skip = c->dist_valid && skip_wanted(c);
    V.skip = skip ? c->d_skip : nullptr; V.SX = c->BX; V.SXY = c->BX * c->BY;
    V.bmin.x = c->bmin[0]; V.bmin.y = c->bmin[1]; V.bmin.z = c->bmin[2];
    V.bmax.x = c->bmax[0]; V.bmax.y = c->bmax[1]; V.bmax.z = c->bmax[2];
    V.vsize.x = c->vsize[0]; V.vsize.y = c->vsize[1]; V.vsize.z = c->vsize[2];
    V.resf.x = (float)c->X; V.resf.y = (float)c->Y; V.resf.z = (float)c->Z;
    {   // dda.h:16 voxelExtent = 1.0 / (boundsMax - boundsMin): binary32 subtraction + division, identical on host and device
        volatile float ex = c->bmax[0] - c->bmin[0], ey = c->bmax[1] - c->bmin[1], ez = c->bmax[2] - c->bmin[2];
        V.inv_extent.x = 1.0f / ex; V.inv_extent.y = 1.0f / ey; V.inv_extent.z = 1.0f / ez;
    }
    // dda.h:98  int(2 * ceil(length(vec3(voxelResolution)))), binary32, unfused
    volatile float xx = V.resf.x * V.resf.x, yy = V.resf.y * V.resf.y, zz = V.resf.z * V.resf.z;
    volatile float s = xx + yy; s = s + zz;
    V.max_steps = (int)(2.0f * std::ceil(std::sqrt((float)s)));
    return V;
}

static Frame make_frame(const vt_ctx* c)
{
    Frame F;
    memcpy(F.inv_mv, c->cam.inv_modelview, sizeof F.inv_mv);
    memcpy(F.proj, c->cam.proj, sizeof F.proj);
    memcpy(F.inv_proj, c->cam.inv_proj, sizeof F.inv_proj);
    F.near_z = c->cam.near_z; F.lens_radius = c->cam.lens_radius; F.lens_model = c->cam.lens_model;
    F.W = c->st.width; F.H = c->st.height; F.max_bounces = c->st.max_bounces;
    F.bg_top.x = c->st.bg_top[0]; F.bg_top.y = c->st.bg_top[1]; F.bg_top.z = c->st.bg_top[2];
    F.bg_bottom.x = c->st.bg_bottom[0]; F.bg_bottom.y = c->st.bg_bottom[1]; F.bg_bottom.z = c->st.bg_bottom[2];
    F.use_image = (c->st.use_env_image && c->d_env) ? 1 : 0;
    F.env_rotation = c->st.env_rotation_rad; F.env_integral = c->env_integral;
    F.wire_opacity = c->st.wireframe_opacity; F.wire_thickness = c->st.wireframe_thickness;
    F.noise = c->d_noise; F.noise_w = c->noise_w; F.noise_h = c->noise_h;
    F.materials = c->d_materials; F.n_materials = (int)c->n_materials;
    F.emissive = c->d_emissive; F.n_emissive = (int)c->n_emissive;
    F.env = c->d_env; F.env_w = c->env_w; F.env_h = c->env_h;
    F.cdf_u = c->d_cdf_u; F.cdf_u_w = c->cdf_u_w; F.cdf_u_h = c->cdf_u_h;
    F.cdf_v = c->d_cdf_v; F.cdf_v_n = c->cdf_v_n;
    F.guide_v = c->d_guide_v; F.guide_u = c->d_guide_u; F.guide_k = c->guide_k;
    F.shared = c->d_shared;
    return F;
}

// ---- rendering ---------------------------------------------------------------------------------------
int vt_reset_accumulation(vt_ctx* c)
{
    if (!c) return VT_ERR_INVALID;
    VT_BIND(c);
    if (c->d_accum) VT_CUDA(c, cudaMemsetAsync(c->d_accum, 0, c->accum_pixels * sizeof(float4), c->stream));
    c->num_samples = 0;
    return VT_OK;
}

int vt_set_partition(vt_ctx* c, int mode, int rank, int world)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, mode >= VT_PART_NONE && mode <= VT_PART_SAMPLES && world >= 1 && rank >= 0 && rank < world, "bad partition");
    c->part_mode = (world == 1) ? VT_PART_NONE : mode; c->part_rank = rank; c->part_world = world;
    return VT_OK;
}

int vt_set_kernel_variant(vt_ctx* c, int variant)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, variant == 0 || variant == 2, "kernel variant must be 0 (megakernel) or 2 (wavefront)");
    c->variant = variant;
    return VT_OK;
}
int vt_set_wavefront_max_paths(vt_ctx* c, size_t n)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, n > 0, "path budget must be positive");
    c->wf_max_paths = n;
    return VT_OK;
}
int vt_kernel_timing_enable(vt_ctx* c, int enable)
{
    if (!c) return VT_ERR_INVALID;
    c->timing = enable != 0;
    return VT_OK;
}
int vt_get_kernel_times(vt_ctx* c, vt_kernel_times* out)
{
    if (!c || !out) return VT_ERR_INVALID;
    VT_BIND(c);
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    memset(out, 0, sizeof *out);
    for (auto& t : c->timed) {
        float ms = 0.f;
        VT_CUDA(c, cudaEventElapsedTime(&ms, t.a, t.b));
        out->ms[t.kind] += ms; out->launches[t.kind] += 1;
        c->ev_pool.push_back(t.a); c->ev_pool.push_back(t.b);
    }
    c->timed.clear();
    return VT_OK;
}
int vt_set_empty_skip(vt_ctx* c, int mode)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, mode >= 0 && mode <= 2, "empty-skip mode must be 0 (off), 1 (auto) or 2 (on)");
    c->skip_mode = mode;
    return VT_OK;
}
int vt_set_wavefront_lanes(vt_ctx* c, int lanes)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, lanes >= 1 && lanes <= vt_ctx::kWfLanes, "lanes must be in [1, 4]");
    c->wf_lanes = lanes;
    return VT_OK;
}
int vt_enable_primary_hits(vt_ctx* c, int enable) { if (!c) return VT_ERR_INVALID; c->primary_enabled = enable != 0; return VT_OK; }
int vt_counters_enable(vt_ctx* c, int enable) { if (!c) return VT_ERR_INVALID; c->count_enabled = enable != 0; return VT_OK; }

} // extern "C" (the wavefront driver is a template)

// ---- wavefront driver (variant 2, vt_wavefront.cuh) -------------------------------------------------------
// bytes of wavefront storage per path in flight: 2 generations of 64-byte path records, 2 regions of 40-byte ray records, the
// sample, 5 shade queues of 20-byte entries, 1 visibility bit
static constexpr size_t kWfBytesPerPath = 2 * 64 + 2 * 40 + 16 + kWfQueues * 20;

static int wf_reserve(vt_ctx* c, int lane, size_t n_paths, int n_iters)
{
    if (!c->wf_stream[lane]) {
        VT_CUDA(c, cudaStreamCreateWithFlags(&c->wf_stream[lane], cudaStreamNonBlocking));
        VT_CUDA(c, cudaEventCreateWithFlags(&c->wf_done[lane], cudaEventDisableTiming));
        VT_CUDA(c, cudaEventCreateWithFlags(&c->wf_acc[lane], cudaEventDisableTiming));
        if (!c->wf_fork) VT_CUDA(c, cudaEventCreateWithFlags(&c->wf_fork, cudaEventDisableTiming));
    }
    if (n_paths > c->wf_capacity[lane] || c->wf_queue_slack > c->wf_slack_alloc[lane]) {
        n_paths = std::max(n_paths, c->wf_capacity[lane]);
        VT_CUDA(c, cudaDeviceSynchronize());
        cudaFree(c->d_wf_pool[lane]); c->d_wf_pool[lane] = nullptr; c->wf_capacity[lane] = 0;
        const size_t n_pad = (n_paths + 63) & ~(size_t)63;                  // keeps every sub-array 256-byte aligned
        const size_t n_queue = n_pad + c->wf_queue_slack;
        const size_t bytes = n_pad * (2 * 64 + 2 * 40 + 16) + n_queue * kWfQueues * 20 + ((n_pad + 2047) / 2048) * 256 + 4096;
        const cudaError_t e = cudaMalloc(&c->d_wf_pool[lane], bytes);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(c, VT_ERR_CUDA, "wavefront pool of %zu bytes: %s", bytes, cudaGetErrorString(e)); }
        char* p = (char*)c->d_wf_pool[lane];
        auto take = [&](size_t b) { char* r = p; p += (b + 255) & ~(size_t)255; return (void*)r; };
        WfState& W = c->wf[lane];
        for (int g = 0; g < 2; ++g) W.state[g] = (float4*)take(n_pad * 64);
        W.rq_a = (int4*)take(2 * n_pad * 16); W.rq_b = (float4*)take(2 * n_pad * 16); W.rq_c = (float2*)take(2 * n_pad * 8); W.rq_cap = (unsigned int)n_pad;
        W.samples = (float4*)take(n_pad * 16);
        for (int k = 0; k < kWfQueues; ++k) { W.sq[k] = (int4*)take(n_queue * 16); W.sq_rng[k] = (int*)take(n_queue * 4); }
        W.vis = (unsigned int*)take(((n_pad + 2047) / 2048) * 256);
        c->wf_capacity[lane] = n_paths; c->wf_slack_alloc[lane] = c->wf_queue_slack;
    }
    if (n_iters > c->wf_counts_cap[lane]) {
        VT_CUDA(c, cudaDeviceSynchronize());
        cudaFree(c->d_wf_counts[lane]); c->d_wf_counts[lane] = nullptr;
        VT_CUDA(c, cudaMalloc(&c->d_wf_counts[lane], sizeof(WfCounts) * (size_t)n_iters));
        c->wf_counts_cap[lane] = n_iters;
    }
    return VT_OK;
}

struct WfTimer {      // RAII event pair around one launch when timing is on
    vt_ctx* c; int kind; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t get(vt_ctx* c) { if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; } cudaEvent_t e; cudaEventCreate(&e); return e; }
    WfTimer(vt_ctx* c_, int kind_, cudaStream_t st_) : c(c_), kind(kind_), st(st_) { if (c->timing) { a = get(c); b = get(c); cudaEventRecord(a, st); } }
    ~WfTimer() { if (a) { cudaEventRecord(b, st); c->timed.push_back({kind, a, b}); } }
};

template <bool COUNT>
static int wf_render(vt_ctx* c, const Volume& V, const Frame& F, const RenderLaunch& L, int my_tiles, int* prim)
{
    const int ci = COUNT ? 1 : 0;
    if (c->wf_shade_blocks[ci] == 0) {
        int per_sm_s = 0, per_sm_t = 0, sms = 0;
        VT_CUDA(c, wf_shade_blocks_per_sm(COUNT, &per_sm_s));
        VT_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_t, wf_trace_kernel<COUNT, true>, kWfTraceThreads, 0));
        VT_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        if (const char* e = getenv("VT_WF_SHADE_CTAS")) { const int v = atoi(e); if (v >= 1) per_sm_s = std::min(per_sm_s, v); }   // tuning knobs
        if (const char* e = getenv("VT_WF_TRACE_CTAS")) { const int v = atoi(e); if (v >= 1) per_sm_t = std::min(per_sm_t, v); }
        c->wf_shade_blocks[ci] = std::max(1, per_sm_s) * std::max(1, sms);
        c->wf_trace_blocks[ci] = std::max(1, per_sm_t) * std::max(1, sms);
        c->wf_sms = std::max(1, sms);
    }
    // every trace warp can leave one partly filled chunk per queue behind (COUNT and plain builds may differ in occupancy)
    const int n_items = my_tiles * kTile * kTile;
    // every warp that appends to the shade queues can leave one partly filled chunk per queue behind: the warps of wf_trace (COUNT and
    // plain builds may differ in occupancy) and those of wf_generate (one per 32 pixels and thread row)
    const int gen_x = n_items / 256, gen_rows_max = std::max(1, (std::max(1, c->wf_sms) * 8 + gen_x - 1) / gen_x);
    c->wf_queue_slack = std::max(c->wf_queue_slack, (size_t)c->wf_trace_blocks[ci] * (kWfTraceThreads / 32) * kWfQueueChunk + 4096);
    c->wf_queue_slack = std::max(c->wf_queue_slack, (size_t)gen_x * gen_rows_max * (256 / 32) * kWfQueueChunk + 4096);
    // split the passes of this call into batches: at most wf_max_paths paths in flight over all lanes, and at least
    // `lanes` batches when there are enough passes, so that two batches always overlap
    const int lanes = std::max(1, std::min(c->wf_lanes, L.n_passes));
    int batch_max;
    {
        const int budget = (int)std::max<size_t>(1, c->wf_max_paths / (size_t)n_items / (size_t)lanes);
        batch_max = std::max(1, std::min(budget, (L.n_passes + lanes - 1) / lanes));
        bool grow = false;
        for (int k = 0; k < lanes; ++k) grow = grow || (size_t)n_items * (size_t)batch_max > c->wf_capacity[k];
        if (grow) {
            // the pools never take more than half of the memory that is free right now (another context, torch or NCCL may share
            // the GPU). Asked only when a pool has to grow: cudaMemGetInfo is a slow driver call when several processes share a box.
            size_t free_b = 0, total_b = 0, held = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                for (int k = 0; k < vt_ctx::kWfLanes; ++k) held += c->wf_capacity[k] * kWfBytesPerPath;
                const size_t cap_paths = std::max<size_t>((free_b + held) / 2 / kWfBytesPerPath, (size_t)n_items * lanes);
                const int budget2 = (int)std::max<size_t>(1, std::min(c->wf_max_paths, cap_paths) / (size_t)n_items / (size_t)lanes);
                batch_max = std::max(1, std::min(budget2, (L.n_passes + lanes - 1) / lanes));
            }
        }
    }
    const int n_iters = F.max_bounces + 3;
    for (;;) {                                        // a smaller batch gives the same bits: halve it when the allocation fails
        int rc = VT_OK;
        for (int k = 0; k < lanes && rc == VT_OK; ++k) rc = wf_reserve(c, k, (size_t)n_items * (size_t)batch_max, n_iters);
        if (rc == VT_OK) break;
        if (batch_max == 1) return rc;
        batch_max = (batch_max + 1) / 2;
    }
    VT_CUDA(c, cudaEventRecord(c->wf_fork, c->stream));
    for (int k = 0; k < lanes; ++k) VT_CUDA(c, cudaStreamWaitEvent(c->wf_stream[k], c->wf_fork, 0));
    int batch = 0;
    for (int pass0 = 0; pass0 < L.n_passes; pass0 += batch_max, ++batch) {
        const int lane = batch % lanes;
        cudaStream_t st = c->wf_stream[lane];
        WfState S = c->wf[lane]; S.n_items = n_items;
        WfCounts* cn = c->d_wf_counts[lane];
        const int nb = std::min(batch_max, L.n_passes - pass0);
        const size_t vis_bytes = (((size_t)n_items * nb + 2047) / 2048) * 256;
        if (batch >= lanes) VT_CUDA(c, cudaStreamWaitEvent(st, c->wf_acc[lane], 0));      // the lane's samples were folded in
        VT_CUDA(c, cudaMemsetAsync(cn, 0, sizeof(WfCounts) * (size_t)n_iters, st));
        { WfTimer t(c, VT_K_GENERATE, st);
          const int gen_rows = std::min(nb, gen_rows_max);   // enough CTAs for every SM, else 1 row
          const dim3 gg((unsigned)gen_x, (unsigned)gen_rows);
          if (V.skip != nullptr && !COUNT) wf_generate_kernel<COUNT, true><<<gg, 256, 0, st>>>(V, F, L, S, pass0, nb, cn, prim, c->d_counters);
          else wf_generate_kernel<COUNT, false><<<gg, 256, 0, st>>>(V, F, L, S, pass0, nb, cn, prim, c->d_counters); }
        c->launches += 1;
        for (int it = 0; ; ++it) {
            // it == 0: the primary rays were traced (and routed) by wf_generate; it >= 1: the shadow + bounce rays emitted by
            // wf_shade(it - 1). Generation `it` of the paths lives in state[it & 1]; wf_shade compacts its survivors into
            // state[(it + 1) & 1]. Counts block `it` holds the size of generation `it`, block it + 1 the shade queues.
            if (it > 0) {
                { WfTimer t(c, VT_K_OTHER, st); VT_CUDA(c, cudaMemsetAsync(S.vis, 0, vis_bytes, st)); }
                { WfTimer t(c, VT_K_TRACE, st);
                  if (V.skip != nullptr && !COUNT) wf_trace_kernel<COUNT, true><<<c->wf_trace_blocks[ci], kWfTraceThreads, 0, st>>>(V, F, S, cn + it, cn + it + 1, c->d_counters);
                  else wf_trace_kernel<COUNT, false><<<c->wf_trace_blocks[ci], kWfTraceThreads, 0, st>>>(V, F, S, cn + it, cn + it + 1, c->d_counters); }
                c->launches += 1;
            }
            if (it == 0 && prim != nullptr && pass0 + nb == L.n_passes) {
                WfTimer t(c, VT_K_OTHER, st);
                wf_primary_kernel<<<dim3((unsigned)(c->wf_sms * 4), kWfQueues), 256, 0, st>>>(V, F, L, S, pass0, cn + 1, prim);
                c->launches += 1;
            }
            { WfTimer t(c, VT_K_SHADE, st); wf_shade_launch(COUNT, (unsigned)c->wf_shade_blocks[ci], st, V, F, S, it & 1, cn + it + 1, c->d_counters); }
            c->launches += 1;
            if (it == F.max_bounces) break;
        }
        // accumulation.fs in pass order: on the caller's stream, after this batch's last shade
        VT_CUDA(c, cudaEventRecord(c->wf_done[lane], st));
        VT_CUDA(c, cudaStreamWaitEvent(c->stream, c->wf_done[lane], 0));
        { WfTimer t(c, VT_K_ACCUMULATE, c->stream); wf_accumulate_kernel<<<(unsigned)((n_items + 255) / 256), 256, 0, c->stream>>>(F, L, S, pass0, nb, c->d_accum); }
        VT_CUDA(c, cudaEventRecord(c->wf_acc[lane], c->stream));
        c->launches += 1;
    }
    return VT_OK;
}

extern "C" {

int vt_render(vt_ctx* c, int first_sample, int n_passes)
{
    if (!c) return VT_ERR_INVALID;
    if (!c->have_cam || !c->have_settings || !c->d_accum) return fail(c, VT_ERR_STATE, "vt_render before vt_set_camera / vt_set_settings");
    VT_REQ(c, n_passes >= 0, "negative pass count");
    if (n_passes == 0) return VT_OK;
    VT_BIND(c);
    const int W = c->st.width, H = c->st.height;
    if (c->primary_enabled && !c->d_primary) VT_CUDA(c, cudaMalloc(&c->d_primary, c->accum_pixels * sizeof(int32_t)));
    RenderLaunch L;
    L.n_passes = n_passes; L.n_prev = c->num_samples; L.integrator = c->st.integrator;
    L.tiles_x = (W + kTile - 1) / kTile; L.tiles_y = (H + kTile - 1) / kTile;
    L.tile_rank = 0; L.tile_world = 1; L.sum_mode = 0; L.first_sample = first_sample; L.sample_stride = 1;
    if (c->part_mode == VT_PART_TILES) { L.tile_rank = c->part_rank; L.tile_world = c->part_world; }
    // sample partition: first_sample counts this rank's passes; global sampleCount = (first + p) * world + rank
    if (c->part_mode == VT_PART_SAMPLES) { L.sum_mode = 1; L.first_sample = first_sample * c->part_world + c->part_rank; L.sample_stride = c->part_world; }
    const int tiles = L.tiles_x * L.tiles_y;
    const int my_tiles = (tiles - L.tile_rank + L.tile_world - 1) / L.tile_world;
    if (my_tiles > 0) {
        if (!c->dist_valid && skip_wanted(c) && c->variant == 2) { const int rc = rebuild_dist(c); if (rc != VT_OK) return rc; }
        const Volume V = make_volume(c); const Frame F = make_frame(c);
        int* prim = c->primary_enabled ? c->d_primary : nullptr;
        if (c->variant == 2 && L.integrator == VT_INTEGRATOR_PATHTRACER && c->st.max_bounces <= kWfMaxBounces) {
            const int rc = c->count_enabled ? wf_render<true>(c, V, F, L, my_tiles, prim) : wf_render<false>(c, V, F, L, my_tiles, prim);
            if (rc != VT_OK) return rc;
            c->launches -= 1;          // the common epilogue below counts one launch
        } else {
            const dim3 grid((unsigned)(my_tiles * kCtasPerTile));
            if (c->count_enabled) vt_render_kernel<true><<<grid, 128, 0, c->stream>>>(V, F, L, c->d_accum, prim, c->d_counters);
            else vt_render_kernel<false><<<grid, 128, 0, c->stream>>>(V, F, L, c->d_accum, prim, c->d_counters);
        }
        VT_CUDA(c, cudaGetLastError());
        c->launches += 1;
        if (c->count_enabled) c->paths += (uint64_t)n_passes * (uint64_t)W * H / (c->part_mode == VT_PART_TILES ? c->part_world : 1);
    }
    c->num_samples += n_passes;
    return VT_OK;
}

int vt_get_num_samples(vt_ctx* c, int* n) { if (!c || !n) return VT_ERR_INVALID; *n = c->num_samples; return VT_OK; }

int vt_read_average(vt_ctx* c, float* out)
{
    if (!c || !out) return VT_ERR_INVALID;
    if (!c->d_accum) return fail(c, VT_ERR_STATE, "no frame");
    VT_BIND(c);
    VT_CUDA(c, cudaMemcpyAsync(out, c->d_accum, c->accum_pixels * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    return VT_OK;
}

int vt_read_primary_hits(vt_ctx* c, int32_t* out)
{
    if (!c || !out) return VT_ERR_INVALID;
    if (!c->d_primary) return fail(c, VT_ERR_STATE, "primary hits not enabled before the last vt_render");
    VT_BIND(c);
    VT_CUDA(c, cudaMemcpyAsync(out, c->d_primary, c->accum_pixels * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    return VT_OK;
}

// K3 (shared/textureMap.fs + the blit of renderer.cpp:613-637) as a read-out: the running average converted the way GL
// converts a float colour written to an 8-bit UNORM framebuffer (clamp to [0, 1], scale by 255, round to nearest; NaN -> 0),
// optionally flipped to top-down row order for image files (saveImage's negative stride, renderer.cpp:1131-1136).
int vt_read_display(vt_ctx* c, uint8_t* rgba8_out, int flip_vertical)
{
    if (!c || !rgba8_out) return VT_ERR_INVALID;
    if (!c->d_accum) return fail(c, VT_ERR_STATE, "no frame");
    VT_BIND(c);
    if (!c->d_display) VT_CUDA(c, cudaMalloc(&c->d_display, c->accum_pixels * sizeof(uchar4)));
    const int W = c->st.width, H = c->st.height;
    vt_display_kernel<<<(unsigned)((c->accum_pixels + 255) / 256), 256, 0, c->stream>>>(c->d_accum, c->d_display, W, H, flip_vertical ? 1 : 0);
    VT_CUDA(c, cudaGetLastError());
    VT_CUDA(c, cudaMemcpyAsync(rgba8_out, c->d_display, c->accum_pixels * sizeof(uchar4), cudaMemcpyDeviceToHost, c->stream));
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    return VT_OK;
}

void* vt_accum_device_ptr(vt_ctx* c) { return c ? (void*)c->d_accum : nullptr; }

int vt_get_counters(vt_ctx* c, vt_counters* out)
{
    if (!c || !out) return VT_ERR_INVALID;
    VT_BIND(c);
    Counters h;
    VT_CUDA(c, cudaMemcpyAsync(&h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    out->dda_steps = h.S; out->rand_calls = h.R; out->material_evals = h.H; out->cdf_loads = h.E; out->env_lookups = h.Q;
    out->paths = c->paths; out->kernel_launches = c->launches;
    return VT_OK;
}
int vt_reset_counters(vt_ctx* c)
{
    if (!c) return VT_ERR_INVALID;
    VT_BIND(c);
    VT_CUDA(c, cudaMemsetAsync(c->d_counters, 0, sizeof(Counters), c->stream));
    c->paths = 0;
    return VT_OK;
}

// ---- voxelizer ----------------------------------------------------------------------------------------
// mesh staging buffers of vt_voxelize, kept between calls (a 512^3 voxelization takes ~0.1 ms on the device: three cudaMalloc /
// cudaFree pairs per call would cost more than the kernels)
static int mesh_reserve(vt_ctx* c, size_t n_verts, size_t n_indices)
{
    if (n_verts > c->mesh_verts_cap || !c->d_mesh_xyz) {
        cudaFree(c->d_mesh_xyz); c->d_mesh_xyz = nullptr; c->mesh_verts_cap = 0;
        VT_CUDA(c, cudaMalloc(&c->d_mesh_xyz, std::max<size_t>(1, n_verts) * 3 * sizeof(float)));
        c->mesh_verts_cap = n_verts;
    }
    if (n_indices > c->mesh_idx_cap || !c->d_mesh_idx) {
        cudaFree(c->d_mesh_idx); c->d_mesh_idx = nullptr; c->mesh_idx_cap = 0;
        VT_CUDA(c, cudaMalloc(&c->d_mesh_idx, std::max<size_t>(1, n_indices) * sizeof(unsigned int)));
        c->mesh_idx_cap = n_indices;
    }
    if (!c->d_mesh_M) VT_CUDA(c, cudaMalloc(&c->d_mesh_M, 16 * sizeof(float)));
    return VT_OK;
}

int vt_voxelize(vt_ctx* c, const float* xyz, size_t n_verts, const uint32_t* indices, size_t n_indices,
                const float M[16], int X, int Y, int Z, int32_t fill)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, xyz && indices && M && n_indices % 3 == 0, "bad mesh arrays");
    VT_REQ(c, fill >= 0, "fill offset must be >= 0");
    for (size_t i = 0; i < n_indices; ++i) VT_REQ(c, indices[i] < n_verts, "vertex index out of range");
    VT_BIND(c);
    int rc = alloc_volume(c, X, Y, Z);
    if (rc == VT_OK) rc = alloc_ids(c, 1);
    if (rc == VT_OK) rc = set_id_table(c, std::vector<int32_t>(1, fill));     // id 0 = the fill record (+ an id for offset 0, see set_id_table)
    if (rc == VT_OK) rc = mesh_reserve(c, n_verts, n_indices);
    if (rc != VT_OK) return rc;
    VT_CUDA(c, cudaMemcpyAsync(c->d_mesh_xyz, xyz, n_verts * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    VT_CUDA(c, cudaMemcpyAsync(c->d_mesh_idx, indices, n_indices * sizeof(unsigned int), cudaMemcpyHostToDevice, c->stream));
    VT_CUDA(c, cudaMemcpyAsync(c->d_mesh_M, M, 16 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    const int n_tris = (int)(n_indices / 3);
    const size_t n = (size_t)X * Y * Z;
    // timed region A (ev0..ev1), SURVEY 8d "kernel time incl. grid clear, excl. OBJ parse and H2D": clear of the bit grid + scatter.
    // timed region B (ev0..ev2): + the id grid made valid (cleared to empty, solid voxels filled) + the renderer's distance field.
    VT_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    rc = clear_occupancy(c);
    if (rc != VT_OK) return rc;
    if (n_tris > 0) {
        const int ctas = std::max(1, std::min((n_tris + 3) / 4, 148 * 16));
        if (c->voxelize_fat) vt_voxelize_kernel<true><<<ctas, 128, 0, c->stream>>>(c->d_mesh_xyz, c->d_mesh_idx, n_tris, c->d_mesh_M, X, Y, Z, c->PBX, c->PBX * c->PBY, c->d_bricks);
        else vt_voxelize_kernel<false><<<ctas, 128, 0, c->stream>>>(c->d_mesh_xyz, c->d_mesh_idx, n_tris, c->d_mesh_M, X, Y, Z, c->PBX, c->PBX * c->PBY, c->d_bricks);
        c->launches += 1;
    }
    VT_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    VT_CUDA(c, cudaMemsetAsync(c->d_ids, 0xff, n, c->stream));
    vt_fill_solid_kernel<<<grid_for((size_t)c->BX * c->BY * c->BZ, 256), 256, 0, c->stream>>>(c->d_bricks, c->d_ids, 1, X, Y, Z, c->BX, c->BY, c->BZ, c->PBX, c->PBX * c->PBY, 0);
    c->launches += 1;
    c->n_emissive = 0;                                  // the volume was replaced: the old emissive list refers to nothing
    rc = rebuild_dist(c);
    if (rc != VT_OK) return rc;
    VT_CUDA(c, cudaEventRecord(c->ev2, c->stream));
    VT_CUDA(c, cudaGetLastError());
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    VT_CUDA(c, cudaEventElapsedTime(&c->last_voxelize_ms, c->ev0, c->ev1));
    VT_CUDA(c, cudaEventElapsedTime(&c->last_voxelize_full_ms, c->ev0, c->ev2));
    return VT_OK;
}

int vt_set_voxelize_thickness(vt_ctx* c, int thickness)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, thickness == VT_VOXELIZE_THIN || thickness == VT_VOXELIZE_FAT, "thickness must be VT_VOXELIZE_THIN or VT_VOXELIZE_FAT");
    c->voxelize_fat = thickness == VT_VOXELIZE_FAT;
    return VT_OK;
}
int vt_get_last_voxelize_ms(vt_ctx* c, float* ms) { if (!c || !ms) return VT_ERR_INVALID; *ms = c->last_voxelize_ms; return VT_OK; }
int vt_get_last_voxelize_full_ms(vt_ctx* c, float* ms) { if (!c || !ms) return VT_ERR_INVALID; *ms = c->last_voxelize_full_ms; return VT_OK; }

int vt_volume_assign_materials(vt_ctx* c, const int32_t* table, int n_table, int rule)
{
    if (!c) return VT_ERR_INVALID;
    VT_REQ(c, table && n_table > 0 && rule == 1, "bad material rule");
    VT_REQ(c, n_table <= 254, "rule-based assignment takes at most 254 material records");
    VT_REQ(c, c->d_ids != nullptr, "no volume");
    for (int i = 0; i < n_table; ++i) VT_REQ(c, table[i] >= 0, "material offsets must be >= 0");
    VT_BIND(c);
    if (c->id_bytes != 1) {                               // a volume uploaded with > 255 records: narrow it first (every solid voxel gets a new id anyway)
        std::vector<int32_t> grid((size_t)c->X * c->Y * c->Z);
        int rc = vt_read_volume(c, grid.data());
        if (rc != VT_OK) return rc;
        for (auto& v : grid) if (v >= 0) v = table[0];
        rc = vt_volume_upload(c, grid.data(), c->X, c->Y, c->Z);
        if (rc != VT_OK) return rc;
    }
    // the ids become indices into the caller's table
    int rc = set_id_table(c, std::vector<int32_t>(table, table + n_table));
    if (rc != VT_OK) return rc;
    const size_t n = (size_t)c->X * c->Y * c->Z;
    vt_assign_materials_kernel<<<grid_for(n, 256), 256, 0, c->stream>>>(c->d_ids, c->id_bytes, c->X, c->Y, c->Z, n_table, rule);
    c->launches += 1;
    VT_CUDA(c, cudaGetLastError());
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    return VT_OK;
}

// ---- services -------------------------------------------------------------------------------------------
static int need_frame(vt_ctx* c)
{
    if (!c->have_cam || !c->have_settings) return fail(c, VT_ERR_STATE, "service called before vt_set_camera / vt_set_settings");
    return VT_OK;
}
int vt_pick(vt_ctx* c, float px, float py)
{
    if (!c) return VT_ERR_INVALID;
    int rc = need_frame(c); if (rc) return rc;
    VT_BIND(c);
    vt_pick_kernel<<<1, 32, 0, c->stream>>>(make_volume(c), make_frame(c), px, py, c->d_shared);
    c->launches += 1;
    VT_CUDA(c, cudaGetLastError());
    return VT_OK;
}
int vt_pick_focal(vt_ctx* c, float px, float py)
{
    if (!c) return VT_ERR_INVALID;
    int rc = need_frame(c); if (rc) return rc;
    VT_BIND(c);
    vt_pick_focal_kernel<<<1, 32, 0, c->stream>>>(make_volume(c), make_frame(c), px, py, c->d_shared);
    c->launches += 1;
    VT_CUDA(c, cudaGetLastError());
    return VT_OK;
}
int vt_add_voxel(vt_ctx* c, float mx, float my)
{
    if (!c) return VT_ERR_INVALID;
    int rc = need_frame(c); if (rc) return rc;
    VT_BIND(c);
    vt_add_voxel_kernel<<<1, 32, 0, c->stream>>>(make_volume(c), make_frame(c), mx, my, c->d_shared, c->d_ids, c->id_bytes, c->zero_id, c->d_bricks, c->d_result);
    c->launches += 1;
    VT_CUDA(c, cudaGetLastError());
    c->dist_valid = false;                                  // an empty cell may have become solid
    return VT_OK;
}
int vt_remove_voxel(vt_ctx* c)
{
    if (!c) return VT_ERR_INVALID;
    VT_BIND(c);
    vt_remove_voxel_kernel<<<1, 32, 0, c->stream>>>(make_volume(c), c->d_shared, c->d_ids, c->id_bytes, c->d_bricks, c->d_result);
    c->launches += 1;
    VT_CUDA(c, cudaGetLastError());
    return VT_OK;
}

int vt_measure_l2_bandwidth(vt_ctx* c, size_t bytes, int reps, float* gbs)
{
    if (!c || !gbs || bytes < 4096 || reps < 1) return VT_ERR_INVALID;
    VT_BIND(c);
    uint4* buf = nullptr; unsigned int* sink = nullptr;
    VT_CUDA(c, cudaMalloc(&buf, bytes)); VT_CUDA(c, cudaMalloc(&sink, 4));
    VT_CUDA(c, cudaMemsetAsync(buf, 1, bytes, c->stream));
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const size_t n_vec = bytes / 16;
    vt_l2_read_kernel<<<sms * 8, 256, 0, c->stream>>>(buf, n_vec, 2, sink);          // warm: pull the buffer into L2
    VT_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    vt_l2_read_kernel<<<sms * 8, 256, 0, c->stream>>>(buf, n_vec, reps, sink);
    VT_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    c->launches += 2;
    VT_CUDA(c, cudaGetLastError());
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    VT_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *gbs = (float)((double)n_vec * 16.0 * reps / (ms * 1e-3) / 1e9);
    cudaFree(buf); cudaFree(sink);
    return VT_OK;
}

int vt_debug_advance(vt_ctx* c, const float* d, const float* e, const float* tau, const int32_t* nmax, size_t n, float* d_out, int32_t* k_out, int literal)
{
    if (!c || !d || !e || !tau || !nmax || !d_out || !k_out) return VT_ERR_INVALID;
    if (n == 0) return VT_OK;
    VT_BIND(c);
    float *dd = nullptr, *de = nullptr, *dt = nullptr, *dout = nullptr; int *dn = nullptr, *dk = nullptr;
    VT_CUDA(c, cudaMalloc(&dd, n * 4)); VT_CUDA(c, cudaMalloc(&de, n * 4)); VT_CUDA(c, cudaMalloc(&dt, n * 4));
    VT_CUDA(c, cudaMalloc(&dout, n * 4)); VT_CUDA(c, cudaMalloc(&dn, n * 4)); VT_CUDA(c, cudaMalloc(&dk, n * 4));
    VT_CUDA(c, cudaMemcpyAsync(dd, d, n * 4, cudaMemcpyHostToDevice, c->stream));
    VT_CUDA(c, cudaMemcpyAsync(de, e, n * 4, cudaMemcpyHostToDevice, c->stream));
    VT_CUDA(c, cudaMemcpyAsync(dt, tau, n * 4, cudaMemcpyHostToDevice, c->stream));
    VT_CUDA(c, cudaMemcpyAsync(dn, nmax, n * 4, cudaMemcpyHostToDevice, c->stream));
    vt_advance_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(dd, de, dt, dn, n, dout, dk, literal);
    c->launches += 1;
    VT_CUDA(c, cudaGetLastError());
    VT_CUDA(c, cudaMemcpyAsync(d_out, dout, n * 4, cudaMemcpyDeviceToHost, c->stream));
    VT_CUDA(c, cudaMemcpyAsync(k_out, dk, n * 4, cudaMemcpyDeviceToHost, c->stream));
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(dd); cudaFree(de); cudaFree(dt); cudaFree(dout); cudaFree(dn); cudaFree(dk);
    return VT_OK;
}

int vt_debug_div_const(vt_ctx* c, int which, uint64_t* mismatches, uint32_t* first_bad)
{
    if (!c || !mismatches || !first_bad || which < 0 || which > 2) return VT_ERR_INVALID;
    VT_BIND(c);
    unsigned long long* d_m = nullptr; unsigned int* d_f = nullptr;
    VT_CUDA(c, cudaMalloc(&d_m, 8)); VT_CUDA(c, cudaMalloc(&d_f, 4));
    VT_CUDA(c, cudaMemsetAsync(d_m, 0, 8, c->stream)); VT_CUDA(c, cudaMemsetAsync(d_f, 0xff, 4, c->stream));
    vt_div_const_kernel<<<148 * 8, 256, 0, c->stream>>>(which, d_m, d_f);
    c->launches += 1;
    VT_CUDA(c, cudaGetLastError());
    unsigned long long m = 0; unsigned int f = 0;
    VT_CUDA(c, cudaMemcpyAsync(&m, d_m, 8, cudaMemcpyDeviceToHost, c->stream));
    VT_CUDA(c, cudaMemcpyAsync(&f, d_f, 4, cudaMemcpyDeviceToHost, c->stream));
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_m); cudaFree(d_f);
    *mismatches = m; *first_bad = f;
    return VT_OK;
}

int vt_debug_trace_rays(vt_ctx* c, const float* rays, size_t n, float* out)
{
    if (!c || !rays || !out) return VT_ERR_INVALID;
    if (n == 0) return VT_OK;
    VT_BIND(c);
    float *d_r = nullptr, *d_o = nullptr;
    VT_CUDA(c, cudaMalloc(&d_r, n * 6 * sizeof(float)));
    VT_CUDA(c, cudaMalloc(&d_o, n * 4 * sizeof(float)));
    VT_CUDA(c, cudaMemcpyAsync(d_r, rays, n * 6 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    vt_trace_rays_kernel<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(make_volume(c), d_r, n, d_o);
    c->launches += 1;
    VT_CUDA(c, cudaGetLastError());
    VT_CUDA(c, cudaMemcpyAsync(out, d_o, n * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    VT_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_r); cudaFree(d_o);
    return VT_OK;
}

} // extern "C"

#include "vt_group.inl"
