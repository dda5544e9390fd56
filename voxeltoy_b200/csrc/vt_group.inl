// vt_group.inl -- multi-GPU render groups behind the C ABI (include/voxeltoy_b200.h, "render groups"); part of vt_api.cu.
//
// The reference renders on one GL context (renderer/renderer.cpp:556-645); sharding one frame over the 8 GPUs of a box is
// this build's addition. Paths are independent and the scene is replicated, so rendering needs no collective.
// The only exchanges are
//   * the combination of the per-rank accumulators at read-out:
//       TILES    every rank owns the 64x64 tiles t with t % world == rank; it packs them into a compact buffer
//                (vt_pack_tiles_kernel) and the root gathers exactly W*H*16/world bytes from each rank (ncclSend / ncclRecv in
//                one group), then scatters the tiles into the frame (vt_unpack_tiles_kernel). Bit-identical to one GPU.
//       SAMPLES  every rank holds a float4 SUM over its own sample indices; ncclReduce(sum) to the root, one division.
//     Both run on a side stream from a snapshot of the accumulator, so the next vt_group_render overlaps the exchange;
//   * a 32-byte edit record (vt_group_broadcast) when the edit is issued on one rank only.
// A second exchange path does the same combination with this library's own kernel over peer memory (VT_EXCHANGE_PEER: the
// root reads every rank's snapshot through NVLink P2P loads -- one pass, no staging, fixed summation order); it is the
// only path for contexts that share a device (NCCL refuses duplicate devices in one communicator) and needs all contexts
// in one process.
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch already loaded when running under Python, the system
// one otherwise), so the library itself has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclReduce) Reduce = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    bool ok = false;
};

NcclApi& nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    // a copy that is already in the process (PyTorch's, when running under Python) wins; VT_NCCL_LIBRARY names one explicitly
    if (const char* e = getenv("VT_NCCL_LIBRARY")) api.handle = dlopen(e, RTLD_NOW | RTLD_LOCAL);
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* nm : names) { if (!api.handle) api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD); }
    for (const char* nm : names) { if (!api.handle) api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL); }
    if (!api.handle) return api;
#define VT_NCCL_SYM(field, sym) api.field = (decltype(api.field))dlsym(api.handle, #sym)
    VT_NCCL_SYM(GetUniqueId, ncclGetUniqueId); VT_NCCL_SYM(CommInitRank, ncclCommInitRank); VT_NCCL_SYM(CommInitAll, ncclCommInitAll);
    VT_NCCL_SYM(CommDestroy, ncclCommDestroy); VT_NCCL_SYM(Reduce, ncclReduce); VT_NCCL_SYM(Broadcast, ncclBroadcast);
    VT_NCCL_SYM(Send, ncclSend); VT_NCCL_SYM(Recv, ncclRecv); VT_NCCL_SYM(GroupStart, ncclGroupStart); VT_NCCL_SYM(GroupEnd, ncclGroupEnd);
    VT_NCCL_SYM(GetErrorString, ncclGetErrorString); VT_NCCL_SYM(GetVersion, ncclGetVersion);
#undef VT_NCCL_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.Reduce && api.Broadcast && api.Send &&
             api.Recv && api.GroupStart && api.GroupEnd && api.GetErrorString;
    return api;
}

// ---- kernels of the exchange -------------------------------------------------------------------------------------
// owned tiles of `rank` -> compact buffer: local tile l (global tile rank + l * world) at packed[l * 4096 ...], row-major
// inside the tile; pixels beyond the frame edge are written as zero.
__global__ void __launch_bounds__(256)
vt_pack_tiles_kernel(const float4* __restrict__ frame, float4* __restrict__ packed, int W, int H, int tiles_x, int n_local, int rank, int world)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_local * (kTile * kTile)) return;
    const int l = (int)(i >> 12), in = (int)(i & 4095);
    const int tile = rank + l * world, tx = tile % tiles_x, ty = tile / tiles_x;
    const int px = tx * kTile + (in & 63), py = ty * kTile + (in >> 6);
    packed[i] = (px < W && py < H) ? frame[(size_t)px + (size_t)py * W] : make_float4(0.f, 0.f, 0.f, 0.f);
}
// gathered[r * chunk + l * 4096 + in] -> frame, for every rank r (chunk = max tiles per rank * 4096)
__global__ void __launch_bounds__(256)
vt_unpack_tiles_kernel(const float4* __restrict__ gathered, float4* __restrict__ frame, int W, int H, int tiles_x, int n_tiles, int world, size_t chunk)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_tiles * (kTile * kTile)) return;
    const int tile = (int)(i >> 12), in = (int)(i & 4095);
    const int r = tile % world, l = tile / world, tx = tile % tiles_x, ty = tile / tiles_x;
    const int px = tx * kTile + (in & 63), py = ty * kTile + (in >> 6);
    if (px < W && py < H) frame[(size_t)px + (size_t)py * W] = gathered[(size_t)r * chunk + ((size_t)l << 12) + in];
}
__global__ void __launch_bounds__(256)
vt_scale_kernel(float4* __restrict__ frame, size_t n, float divisor)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = frame[i];
    v.x = v.x / divisor; v.y = v.y / divisor; v.z = v.z / divisor; v.w = v.w / divisor;
    frame[i] = v;
}
// VT_EXCHANGE_PEER: one pass on the root over the snapshots of all ranks (peer pointers: NVLink P2P loads, or plain loads
// when the ranks share the device). TILES: each pixel is copied from its owner; SAMPLES: summed in rank order, divided once.
struct PeerPtrs { const float4* p[16]; };
__global__ void __launch_bounds__(256)
vt_peer_combine_kernel(PeerPtrs src, float4* __restrict__ out, int W, int H, int tiles_x, int world, int mode, float divisor)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)W * H) return;
    if (mode == VT_PART_TILES) {
        const int px = (int)(i % W), py = (int)(i / W);
        const int tile = (px >> 6) + (py >> 6) * tiles_x;
        out[i] = src.p[tile % world][i];
    } else {
        float4 s = src.p[0][i];
        for (int r = 1; r < world; ++r) { const float4 v = src.p[r][i]; s.x = s.x + v.x; s.y = s.y + v.y; s.z = s.z + v.z; s.w = s.w + v.w; }
        s.x = s.x / divisor; s.y = s.y / divisor; s.z = s.z / divisor; s.w = s.w / divisor;
        out[i] = s;
    }
}

} // namespace

struct vt_group {
    int mode = VT_PART_SAMPLES, world = 1, exchange = VT_EXCHANGE_NCCL;
    bool owns_ctx = false, in_process = true;
    std::vector<vt_ctx*> ctx;                  // local contexts
    std::vector<int> rank;                     // their global ranks
    std::vector<ncclComm_t> comm;              // one communicator per local context (empty: peer exchange only)
    std::vector<cudaStream_t> side;            // exchange streams
    std::vector<cudaEvent_t> ev_snap, ev_done;
    std::vector<float4*> snap;                 // snapshot of the accumulator the exchange reads
    std::vector<float4*> packed;               // TILES: this rank's tiles, compact
    float4* gathered = nullptr;                // root: world * chunk (TILES) or the reduced frame (SAMPLES) / combined frame
    float4* result = nullptr;                  // root: the finished frame
    size_t frame_px = 0, chunk = 0;
    int W = 0, H = 0;
    bool pending = false;
    float last_exchange_ms = 0.f; cudaEvent_t t0 = nullptr, t1 = nullptr;
    std::string err;
};

static int gfail(vt_group* g, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (g) g->err = buf;
    return code;
}
#define VTG_CUDA(g, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return gfail((g), VT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define VTG_NCCL(g, call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
    return gfail((g), VT_ERR_CUDA, "%s failed: %s", #call, nccl_api().GetErrorString(r_)); } while (0)

static int group_root_local(const vt_group* g) { for (size_t i = 0; i < g->rank.size(); ++i) if (g->rank[i] == 0) return (int)i; return -1; }

static void group_free_buffers(vt_group* g)
{
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        cudaSetDevice(g->ctx[i]->device);
        if (i < g->snap.size()) { cudaFree(g->snap[i]); g->snap[i] = nullptr; }
        if (i < g->packed.size()) { cudaFree(g->packed[i]); g->packed[i] = nullptr; }
    }
    const int rl = group_root_local(g);
    if (rl >= 0) { cudaSetDevice(g->ctx[rl]->device); cudaFree(g->gathered); cudaFree(g->result); }
    g->gathered = nullptr; g->result = nullptr; g->frame_px = 0;
}

// (re)allocates the exchange buffers for the current frame size of the contexts
static int group_prepare(vt_group* g)
{
    const int W = g->ctx[0]->st.width, H = g->ctx[0]->st.height;
    for (vt_ctx* c : g->ctx) {
        if (!c->d_accum || c->st.width != W || c->st.height != H) return gfail(g, VT_ERR_STATE, "the contexts of a group must share one frame size (vt_set_settings)");
    }
    const size_t px = (size_t)W * H;
    if (px == g->frame_px && W == g->W) return VT_OK;
    group_free_buffers(g);
    g->W = W; g->H = H;
    const int tiles = ((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
    g->chunk = (size_t)((tiles + g->world - 1) / g->world) * kTile * kTile;
    g->snap.assign(g->ctx.size(), nullptr); g->packed.assign(g->ctx.size(), nullptr);
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        VTG_CUDA(g, cudaSetDevice(g->ctx[i]->device));
        VTG_CUDA(g, cudaMalloc(&g->snap[i], px * sizeof(float4)));
        if (g->mode == VT_PART_TILES) VTG_CUDA(g, cudaMalloc(&g->packed[i], g->chunk * sizeof(float4)));
    }
    const int rl = group_root_local(g);
    if (rl >= 0) {
        VTG_CUDA(g, cudaSetDevice(g->ctx[rl]->device));
        VTG_CUDA(g, cudaMalloc(&g->result, px * sizeof(float4)));
        if (g->mode == VT_PART_TILES) VTG_CUDA(g, cudaMalloc(&g->gathered, g->chunk * (size_t)g->world * sizeof(float4)));
    }
    g->frame_px = px;
    return VT_OK;
}

static int group_finish_setup(vt_group* g)
{
    const size_t n = g->ctx.size();
    g->side.assign(n, nullptr); g->ev_snap.assign(n, nullptr); g->ev_done.assign(n, nullptr);
    for (size_t i = 0; i < n; ++i) {
        VTG_CUDA(g, cudaSetDevice(g->ctx[i]->device));
        VTG_CUDA(g, cudaStreamCreateWithFlags(&g->side[i], cudaStreamNonBlocking));
        VTG_CUDA(g, cudaEventCreateWithFlags(&g->ev_snap[i], cudaEventDisableTiming));
        VTG_CUDA(g, cudaEventCreateWithFlags(&g->ev_done[i], cudaEventDisableTiming));
        const int rc = vt_set_partition(g->ctx[i], g->mode, g->rank[i], g->world);
        if (rc != VT_OK) return gfail(g, rc, "vt_set_partition: %s", vt_last_error(g->ctx[i]));
    }
    const int rl = group_root_local(g);
    if (rl >= 0) {
        VTG_CUDA(g, cudaSetDevice(g->ctx[rl]->device));
        VTG_CUDA(g, cudaEventCreate(&g->t0)); VTG_CUDA(g, cudaEventCreate(&g->t1));
    }
    return VT_OK;
}

extern "C" {

const char* vt_group_last_error(const vt_group* g) { return g ? g->err.c_str() : "null group"; }
int vt_group_size(const vt_group* g) { return g ? g->world : 0; }
int vt_group_local_size(const vt_group* g) { return g ? (int)g->ctx.size() : 0; }
vt_ctx* vt_group_context(vt_group* g, int local_index) { return (g && local_index >= 0 && (size_t)local_index < g->ctx.size()) ? g->ctx[local_index] : nullptr; }
int vt_group_rank(const vt_group* g, int local_index) { return (g && local_index >= 0 && (size_t)local_index < g->rank.size()) ? g->rank[local_index] : -1; }

int vt_nccl_version(void) { int v = 0; NcclApi& a = nccl_api(); if (a.ok && a.GetVersion) a.GetVersion(&v); return v; }

void vt_group_destroy(vt_group* g)
{
    if (!g) return;
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        cudaSetDevice(g->ctx[i]->device);
        cudaStreamSynchronize(g->ctx[i]->stream);
        if (i < g->side.size() && g->side[i]) { cudaStreamSynchronize(g->side[i]); }
    }
    for (ncclComm_t c : g->comm) if (c) nccl_api().CommDestroy(c);
    group_free_buffers(g);
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        cudaSetDevice(g->ctx[i]->device);
        if (i < g->side.size() && g->side[i]) cudaStreamDestroy(g->side[i]);
        if (i < g->ev_snap.size() && g->ev_snap[i]) cudaEventDestroy(g->ev_snap[i]);
        if (i < g->ev_done.size() && g->ev_done[i]) cudaEventDestroy(g->ev_done[i]);
        vt_set_partition(g->ctx[i], VT_PART_NONE, 0, 1);
        if (g->owns_ctx) vt_destroy(g->ctx[i]);
    }
    if (g->t0) cudaEventDestroy(g->t0);
    if (g->t1) cudaEventDestroy(g->t1);
    delete g;
}

// In-process group over existing contexts (one per rank, rank = position). Contexts on distinct devices exchange through
// NCCL (ncclCommInitAll); when two contexts share a device only the peer-memory exchange is available.
int vt_group_adopt(int n, vt_ctx* const* ctxs, int mode, vt_group** out)
{
    if (!out) return VT_ERR_INVALID;
    *out = nullptr;
    if (n < 1 || n > 16 || !ctxs || (mode != VT_PART_TILES && mode != VT_PART_SAMPLES)) return VT_ERR_INVALID;
    for (int i = 0; i < n; ++i) if (!ctxs[i]) return VT_ERR_INVALID;
    vt_group* g = new vt_group();
    g->mode = mode; g->world = n; g->in_process = true;
    bool distinct = true;
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) { g->ctx.push_back(ctxs[i]); g->rank.push_back(i); devs[i] = ctxs[i]->device; for (int j = 0; j < i; ++j) if (devs[j] == devs[i]) distinct = false; }
    g->exchange = VT_EXCHANGE_PEER;
    if (n > 1 && distinct && nccl_api().ok) {
        g->comm.assign(n, nullptr);
        const ncclResult_t r = nccl_api().CommInitAll(g->comm.data(), n, devs.data());
        if (r == ncclSuccess) g->exchange = VT_EXCHANGE_NCCL;
        else g->comm.clear();
    }
    // peer access from the root to every other device (the peer exchange reads their snapshots directly)
    for (int i = 1; i < n; ++i) {
        if (devs[i] == devs[0]) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devs[0], devs[i]);
        if (can) { cudaSetDevice(devs[0]); if (cudaDeviceEnablePeerAccess(devs[i], 0) != cudaSuccess) cudaGetLastError(); }   // already enabled is fine
        else if (g->exchange == VT_EXCHANGE_PEER) { delete g; return VT_ERR_STATE; }      // neither NCCL nor peer access: no way to combine
    }
    const int rc = group_finish_setup(g);
    if (rc != VT_OK) { vt_group_destroy(g); return rc; }
    *out = g;
    return VT_OK;
}

// In-process group that creates (and owns) one context per listed device.
int vt_group_create(int n, const int* devices, int mode, vt_group** out)
{
    if (!out) return VT_ERR_INVALID;
    *out = nullptr;
    if (n < 1 || n > 16 || !devices) return VT_ERR_INVALID;
    std::vector<vt_ctx*> cs(n, nullptr);
    for (int i = 0; i < n; ++i) {
        const int rc = vt_create(devices[i], &cs[i]);
        if (rc != VT_OK) { for (int j = 0; j < i; ++j) vt_destroy(cs[j]); return rc; }
    }
    const int rc = vt_group_adopt(n, cs.data(), mode, out);
    if (rc != VT_OK) { for (int i = 0; i < n; ++i) vt_destroy(cs[i]); return rc; }
    (*out)->owns_ctx = true;
    return VT_OK;
}

// One-process-per-GPU flavour (torchrun, MPI ...): rank 0 calls vt_group_unique_id, the launcher's own plumbing hands the
// 128 bytes to every rank, every rank calls vt_group_join with its context.
int vt_group_unique_id(void* out128)
{
    if (!out128) return VT_ERR_INVALID;
    NcclApi& a = nccl_api();
    if (!a.ok) return VT_ERR_STATE;
    ncclUniqueId id;
    if (a.GetUniqueId(&id) != ncclSuccess) return VT_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return VT_OK;
}
int vt_group_join(vt_ctx* ctx, const void* id128, int rank, int world, int mode, vt_group** out)
{
    if (!out) return VT_ERR_INVALID;
    *out = nullptr;
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world || (mode != VT_PART_TILES && mode != VT_PART_SAMPLES)) return VT_ERR_INVALID;
    NcclApi& a = nccl_api();
    if (!a.ok) return fail(ctx, VT_ERR_STATE, "NCCL is not available (libnccl.so.2 could not be loaded)");
    vt_group* g = new vt_group();
    g->mode = mode; g->world = world; g->in_process = false; g->exchange = VT_EXCHANGE_NCCL;
    g->ctx.push_back(ctx); g->rank.push_back(rank);
    ncclUniqueId id; memcpy(&id, id128, 128);
    if (cudaSetDevice(ctx->device) != cudaSuccess) { delete g; return VT_ERR_CUDA; }
    g->comm.assign(1, nullptr);
    const ncclResult_t r = a.CommInitRank(&g->comm[0], world, id, rank);
    if (r != ncclSuccess) { fail(ctx, VT_ERR_CUDA, "ncclCommInitRank: %s", a.GetErrorString(r)); g->comm.clear(); delete g; return VT_ERR_CUDA; }
    const int rc = group_finish_setup(g);
    if (rc != VT_OK) { fail(ctx, rc, "%s", g->err.c_str()); vt_group_destroy(g); return rc; }
    *out = g;
    return VT_OK;
}

int vt_group_set_exchange(vt_group* g, int exchange)
{
    if (!g) return VT_ERR_INVALID;
    if (exchange == VT_EXCHANGE_NCCL && g->comm.empty()) return gfail(g, VT_ERR_STATE, "this group has no NCCL communicator (contexts share a device, or NCCL is missing)");
    if (exchange == VT_EXCHANGE_PEER && !g->in_process) return gfail(g, VT_ERR_STATE, "the peer-memory exchange needs all contexts in one process");
    if (exchange != VT_EXCHANGE_NCCL && exchange != VT_EXCHANGE_PEER) return gfail(g, VT_ERR_INVALID, "unknown exchange");
    g->exchange = exchange;
    return VT_OK;
}
int vt_group_get_exchange(const vt_group* g) { return g ? g->exchange : -1; }

int vt_group_render(vt_group* g, int first_sample, int n_passes)
{
    if (!g) return VT_ERR_INVALID;
    for (vt_ctx* c : g->ctx) { const int rc = vt_render(c, first_sample, n_passes); if (rc != VT_OK) return gfail(g, rc, "vt_render (device %d): %s", c->device, vt_last_error(c)); }
    return VT_OK;
}
int vt_group_reset_accumulation(vt_group* g)
{
    if (!g) return VT_ERR_INVALID;
    for (vt_ctx* c : g->ctx) { const int rc = vt_reset_accumulation(c); if (rc != VT_OK) return gfail(g, rc, "%s", vt_last_error(c)); }
    return VT_OK;
}
int vt_group_sync(vt_group* g)
{
    if (!g) return VT_ERR_INVALID;
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        VTG_CUDA(g, cudaSetDevice(g->ctx[i]->device));
        VTG_CUDA(g, cudaStreamSynchronize(g->ctx[i]->stream));
        VTG_CUDA(g, cudaStreamSynchronize(g->side[i]));
    }
    return VT_OK;
}

// Edits (renderer/actions.cpp:20-52 on every replica of the scene): the same service call on every local context; an edit
// that changes the volume resets the accumulation, as Action::m_invalidatesRender does. In the one-process-per-GPU flavour
// every rank makes the same call (vt_group_broadcast carries the record when only one rank knows it).
int vt_group_pick(vt_group* g, float px, float py)
{
    if (!g) return VT_ERR_INVALID;
    for (vt_ctx* c : g->ctx) { const int rc = vt_pick(c, px, py); if (rc != VT_OK) return gfail(g, rc, "%s", vt_last_error(c)); }
    return VT_OK;
}
int vt_group_pick_focal(vt_group* g, float px, float py)
{
    if (!g) return VT_ERR_INVALID;
    for (vt_ctx* c : g->ctx) { int rc = vt_pick_focal(c, px, py); if (rc == VT_OK) rc = vt_reset_accumulation(c); if (rc != VT_OK) return gfail(g, rc, "%s", vt_last_error(c)); }
    return VT_OK;
}
int vt_group_add_voxel(vt_group* g, float mx, float my)
{
    if (!g) return VT_ERR_INVALID;
    for (vt_ctx* c : g->ctx) { int rc = vt_add_voxel(c, mx, my); if (rc == VT_OK) rc = vt_reset_accumulation(c); if (rc != VT_OK) return gfail(g, rc, "%s", vt_last_error(c)); }
    return VT_OK;
}
int vt_group_remove_voxel(vt_group* g)
{
    if (!g) return VT_ERR_INVALID;
    for (vt_ctx* c : g->ctx) { int rc = vt_remove_voxel(c); if (rc == VT_OK) rc = vt_reset_accumulation(c); if (rc != VT_OK) return gfail(g, rc, "%s", vt_last_error(c)); }
    return VT_OK;
}

// `bytes` (<= 256) of host memory from the process that holds rank `root` to every rank, over NCCL on the contexts' streams.
int vt_group_broadcast(vt_group* g, void* host_buf, size_t bytes, int root)
{
    if (!g || !host_buf || bytes == 0 || bytes > 256 || root < 0 || root >= g->world) return VT_ERR_INVALID;
    if (g->comm.empty()) return VT_OK;                 // one process, every context sees the caller's buffer already
    NcclApi& a = nccl_api();
    std::vector<void*> dbuf(g->ctx.size(), nullptr);
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        VTG_CUDA(g, cudaSetDevice(g->ctx[i]->device));
        VTG_CUDA(g, cudaMalloc(&dbuf[i], 256));
        if (g->rank[i] == root) VTG_CUDA(g, cudaMemcpyAsync(dbuf[i], host_buf, bytes, cudaMemcpyHostToDevice, g->ctx[i]->stream));
    }
    VTG_NCCL(g, a.GroupStart());
    for (size_t i = 0; i < g->ctx.size(); ++i) VTG_NCCL(g, a.Broadcast(dbuf[i], dbuf[i], bytes, ncclChar, root, g->comm[i], g->ctx[i]->stream));
    VTG_NCCL(g, a.GroupEnd());
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        VTG_CUDA(g, cudaSetDevice(g->ctx[i]->device));
        if (i == 0) VTG_CUDA(g, cudaMemcpyAsync(host_buf, dbuf[i], bytes, cudaMemcpyDeviceToHost, g->ctx[i]->stream));
        VTG_CUDA(g, cudaStreamSynchronize(g->ctx[i]->stream));
        cudaFree(dbuf[i]);
    }
    return VT_OK;
}

// Starts the combination of the accumulators: snapshot on every context's stream, exchange on the side streams. Returns at
// once; rendering may continue (the exchange reads the snapshots).
int vt_group_begin_combine(vt_group* g)
{
    if (!g) return VT_ERR_INVALID;
    int rc = group_prepare(g);
    if (rc != VT_OK) return rc;
    NcclApi& a = nccl_api();
    const size_t px = g->frame_px;
    const int W = g->W, H = g->H, tiles_x = (W + kTile - 1) / kTile, tiles = tiles_x * ((H + kTile - 1) / kTile);
    const int rl = group_root_local(g);
    // total passes folded into the sums (SAMPLES): every rank has rendered the same number
    const float total = (float)((long long)g->ctx[0]->num_samples * (long long)g->world);
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        vt_ctx* c = g->ctx[i];
        VTG_CUDA(g, cudaSetDevice(c->device));
        VTG_CUDA(g, cudaMemcpyAsync(g->snap[i], c->d_accum, px * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
        VTG_CUDA(g, cudaEventRecord(g->ev_snap[i], c->stream));
        VTG_CUDA(g, cudaStreamWaitEvent(g->side[i], g->ev_snap[i], 0));
    }
    if (rl >= 0) { VTG_CUDA(g, cudaSetDevice(g->ctx[rl]->device)); VTG_CUDA(g, cudaEventRecord(g->t0, g->side[rl])); }
    if (g->world == 1) {
        VTG_CUDA(g, cudaMemcpyAsync(g->result, g->snap[0], px * sizeof(float4), cudaMemcpyDeviceToDevice, g->side[0]));
    } else if (g->exchange == VT_EXCHANGE_PEER) {
        // the root's kernel must not start before every snapshot exists
        PeerPtrs pp; for (int r = 0; r < 16; ++r) pp.p[r] = nullptr;
        for (size_t i = 0; i < g->ctx.size(); ++i) { pp.p[g->rank[i]] = g->snap[i]; if ((int)i != rl) VTG_CUDA(g, cudaStreamWaitEvent(g->side[rl], g->ev_snap[i], 0)); }
        VTG_CUDA(g, cudaSetDevice(g->ctx[rl]->device));
        vt_peer_combine_kernel<<<(unsigned)((px + 255) / 256), 256, 0, g->side[rl]>>>(pp, g->result, W, H, tiles_x, g->world, g->mode, total);
        VTG_CUDA(g, cudaGetLastError());
    } else if (g->mode == VT_PART_SAMPLES) {
        VTG_NCCL(g, a.GroupStart());
        for (size_t i = 0; i < g->ctx.size(); ++i)
            VTG_NCCL(g, a.Reduce(g->snap[i], g->rank[i] == 0 ? (void*)g->result : (void*)g->snap[i], px * 4, ncclFloat, ncclSum, 0, g->comm[i], g->side[i]));
        VTG_NCCL(g, a.GroupEnd());
        if (rl >= 0) {
            VTG_CUDA(g, cudaSetDevice(g->ctx[rl]->device));
            vt_scale_kernel<<<(unsigned)((px + 255) / 256), 256, 0, g->side[rl]>>>(g->result, px, total);
            VTG_CUDA(g, cudaGetLastError());
        }
    } else {
        for (size_t i = 0; i < g->ctx.size(); ++i) {
            const int r = g->rank[i];
            const int n_local = (tiles - r + g->world - 1) / g->world;
            VTG_CUDA(g, cudaSetDevice(g->ctx[i]->device));
            if (n_local > 0) vt_pack_tiles_kernel<<<(unsigned)(((size_t)n_local * 4096 + 255) / 256), 256, 0, g->side[i]>>>(g->snap[i], g->packed[i], W, H, tiles_x, n_local, r, g->world);
            VTG_CUDA(g, cudaGetLastError());
        }
        VTG_NCCL(g, a.GroupStart());
        for (size_t i = 0; i < g->ctx.size(); ++i) {
            const int r = g->rank[i];
            if (r == 0) {
                for (int p = 1; p < g->world; ++p) VTG_NCCL(g, a.Recv(g->gathered + (size_t)p * g->chunk, g->chunk * 4, ncclFloat, p, g->comm[i], g->side[i]));
            } else VTG_NCCL(g, a.Send(g->packed[i], g->chunk * 4, ncclFloat, 0, g->comm[i], g->side[i]));
        }
        VTG_NCCL(g, a.GroupEnd());
        if (rl >= 0) {
            VTG_CUDA(g, cudaSetDevice(g->ctx[rl]->device));
            VTG_CUDA(g, cudaMemcpyAsync(g->gathered, g->packed[rl], g->chunk * sizeof(float4), cudaMemcpyDeviceToDevice, g->side[rl]));
            vt_unpack_tiles_kernel<<<(unsigned)(((size_t)tiles * 4096 + 255) / 256), 256, 0, g->side[rl]>>>(g->gathered, g->result, W, H, tiles_x, tiles, g->world, g->chunk);
            VTG_CUDA(g, cudaGetLastError());
        }
    }
    if (rl >= 0) { VTG_CUDA(g, cudaSetDevice(g->ctx[rl]->device)); VTG_CUDA(g, cudaEventRecord(g->t1, g->side[rl])); }
    for (size_t i = 0; i < g->ctx.size(); ++i) { VTG_CUDA(g, cudaSetDevice(g->ctx[i]->device)); VTG_CUDA(g, cudaEventRecord(g->ev_done[i], g->side[i])); }
    g->pending = true;
    return VT_OK;
}

// Orders every local context's stream after the exchange in flight (no host wait): work enqueued afterwards -- and CUDA events
// recorded on those streams -- come after the combination. For device-side timing and for callers that consume
// vt_group_result_device_ptr on the context's stream.
int vt_group_wait_combine(vt_group* g)
{
    if (!g) return VT_ERR_INVALID;
    if (!g->pending) return VT_OK;
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        VTG_CUDA(g, cudaSetDevice(g->ctx[i]->device));
        VTG_CUDA(g, cudaStreamWaitEvent(g->ctx[i]->stream, g->ev_done[i], 0));
    }
    return VT_OK;
}

// Waits for the exchange; on the process that holds rank 0 copies the finished frame (W*H RGBA float32) to `rgba_out` (may be
// null: the frame stays on the device, vt_group_result_device_ptr). Other processes only wait for their part.
int vt_group_end_combine(vt_group* g, float* rgba_out)
{
    if (!g) return VT_ERR_INVALID;
    if (!g->pending) return gfail(g, VT_ERR_STATE, "vt_group_end_combine without vt_group_begin_combine");
    const int rl = group_root_local(g);
    for (size_t i = 0; i < g->ctx.size(); ++i) { VTG_CUDA(g, cudaSetDevice(g->ctx[i]->device)); VTG_CUDA(g, cudaStreamSynchronize(g->side[i])); }
    if (rl >= 0) {
        VTG_CUDA(g, cudaSetDevice(g->ctx[rl]->device));
        if (rgba_out) VTG_CUDA(g, cudaMemcpy(rgba_out, g->result, g->frame_px * sizeof(float4), cudaMemcpyDeviceToHost));
    }
    g->pending = false;
    return VT_OK;
}
int vt_group_read_average(vt_group* g, float* rgba_out)
{
    const int rc = vt_group_begin_combine(g);
    return rc != VT_OK ? rc : vt_group_end_combine(g, rgba_out);
}
void* vt_group_result_device_ptr(vt_group* g) { return g ? (void*)g->result : nullptr; }
// device time of the most recent exchange on the root's side stream (waits for it); 0 in processes that do not hold rank 0
int vt_group_last_exchange_ms(vt_group* g, float* ms)
{
    if (!g || !ms) return VT_ERR_INVALID;
    const int rl = group_root_local(g);
    if (rl >= 0 && g->t0 && g->frame_px != 0) {
        VTG_CUDA(g, cudaSetDevice(g->ctx[rl]->device));
        VTG_CUDA(g, cudaEventSynchronize(g->t1));
        VTG_CUDA(g, cudaEventElapsedTime(&g->last_exchange_ms, g->t0, g->t1));
    }
    *ms = g->last_exchange_ms;
    return VT_OK;
}
// bytes the root receives from the other ranks per combination (what crosses NVLink)
size_t vt_group_exchange_bytes(const vt_group* g)
{
    if (!g || g->world <= 1) return 0;
    const size_t frame = g->frame_px * sizeof(float4);
    if (g->mode == VT_PART_SAMPLES) return frame * (size_t)(g->world - 1);
    return g->exchange == VT_EXCHANGE_PEER ? frame - frame / (size_t)g->world : g->chunk * sizeof(float4) * (size_t)(g->world - 1);
}

} // extern "C"
