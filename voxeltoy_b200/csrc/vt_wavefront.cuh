// vt_wavefront.cuh -- wavefront path tracer (K1 + K2), render kernel variant 2 (default).
//
// Why wavefront: the one-thread-per-pixel megakernel (vt_render_kernel) issues with 9.6 of 32 lanes active (ncu,
// profiles/r01_v1_*): DDA trip counts differ per ray, paths end at different bounces and every lane waits for the slowest
// lane of its warp. Memory is idle there (L2 < 3 %), so the lever is lane utilisation.
//
// One progressive batch (P passes over this context's share of the frame) is a set of independent paths; the integrator
// loop of pathTracer.fs:214-292 becomes a sequence of kernels over compacted queues:
//
//   wf_generate   pathTracer.fs:172-196   RNG offset, camera ray, slab test, DDA set-up (dda.h:16-34) of the primary ray
//   repeat max_bounces + 1 times:
//     wf_trace    dda.h:38-57             the DDA loop alone, for primary, shadow and bounce rays: persistent warps whose
//                                         lanes refill themselves from the ray records as soon as their ray ends. A finished
//                                         primary / bounce ray is ROUTED here (:202-208, :214, :282-291): surface hit with
//                                         bounces left -> the shade queue of its material type (Lambert / metal / plastic /
//                                         other), everything else -> the finish queue. A finished shadow ray sets one bit.
//     wf_shade    :216-279 (+ :248 / :282-289 of the previous iteration) resolve the previous shadow ray, add the
//                                         environment on a miss, build the hit frame, sample the light and the BSDF;
//                                         emits the DDA set-up of one shadow and one bounce ray per surviving path
//   wf_accumulate accumulation.fs:10-18   folds the P samples of every pixel into the running average in pass order
//
// Round-2 data flow (profiles/r02_*). Round 1 kept the path state as SoA float4 arrays gathered through material-sorted
// slot queues: every 16-byte gather fetched a 32-byte sector (DRAM traffic 2x the useful bytes in wf_shade), wf_trace
// scattered 16-byte hit and 4-byte visibility words, and a separate wf_classify pass re-read them. Now:
//   * path record = 64 B AoS (exactly two sectors: origin, direction, BSDF pdf, radiance, throughput, pending light),
//     written sequentially by wf_shade into the next generation's compact slots and gathered by slot one generation later
//     with two 256-bit loads: every sector that moves is fully used whatever the queue order;
//   * a path's two rays live at ITS slot in a shadow region and a bounce region (40-byte records): no ray counters, and a
//     trace warp's 128-ray range holds rays of one kind (shadow rays towards the sky run long, bounce rays short);
//   * the per-path words that do not fit the record (path id, rng offset + NaN flags) ride in the bounce ray's record and
//     come back in the 20-byte queue entry wf_trace appends at retire together with the hit voxel: shade reads its queue
//     sequentially and never reads a hit array;
//   * queue appends from wf_trace are reserved per warp in chunks of kWfQueueChunk entries (one global atomic per 64
//     entries, no CTA barrier in the persistent loop); the unused tail of a warp's last chunk is marked invalid;
//   * shadow-ray results are one bit per slot (atomicOr into an L2-resident bitmap) instead of a 4-byte scatter;
//   * streams bypass L1 and are evict-first in L2, the gather tables (noise, environment, CDFs) evict-last (vt_mem.cuh).
//
// Every path performs exactly the arithmetic of trace_pixel() (vt_device.cuh) in the same order, and the per-pixel order
// of the running average is unchanged: results are bit-identical to the megakernel and to the reference shaders.
#pragma once
#include "vt_kernels.cuh"

namespace vt {

constexpr int kWfQueues = 5;          // 0 = finish, 1..3 = material type 0..2, 4 = other material types
#ifndef VT_WF_LIVE_MIN
#define VT_WF_LIVE_MIN 26
#endif
#ifndef VT_WF_SHADE_THREADS
#define VT_WF_SHADE_THREADS 128      // threads per wf_shade CTA (64 / 128 / 256 measured in round 1: 84.2 / 84.2 / 84.6 ms per step)
#endif
#ifndef VT_WF_SHADE_MIN_BLOCKS
#define VT_WF_SHADE_MIN_BLOCKS (1024 / VT_WF_SHADE_THREADS)
#endif
// wf_trace steps its rays in chunks of kWfStepChunk DDA iterations and keeps stepping, chunk after chunk, while at least
// kWfLiveMin lanes still run a ray; only then does it pay for a retire + refill round (about as many instructions as a
// chunk). Round 1 retired and refilled after every 16-iteration chunk: half the lanes had finished by then (rays average
// 17 iterations on C2) and 21 of 32 lanes did work. Re-tuned once the primary rays had left this kernel (they were its longest rays):
// 16 / 20 / 24 / 26 / 28 / 30 lanes -> 65.0 / 62.6 / 61.3 / 60.7 / 60.8 / 60.9 ms per C2 step; C3 trace 68.2 -> 63.1 ms at 26.
constexpr int kWfLiveMin = VT_WF_LIVE_MIN;
#ifndef VT_WF_STEP_CHUNK
#define VT_WF_STEP_CHUNK 16
#endif
constexpr int kWfStepChunk = VT_WF_STEP_CHUNK;
#ifndef VT_WF_STEP_CHUNK_SKIP
#define VT_WF_STEP_CHUNK_SKIP 16
#endif
constexpr int kWfStepChunkSkip = VT_WF_STEP_CHUNK_SKIP;   // chunk of the kernel instance with the empty-space skip (one skip attempt per chunk)
#ifndef VT_WF_GRAB
#define VT_WF_GRAB 256
#endif
#ifndef VT_WF_SKIP_MIN_LANES
#define VT_WF_SKIP_MIN_LANES 16
#endif
constexpr int kWfGrab = VT_WF_GRAB;                       // rays a warp reserves per atomic on the hand-out counter
constexpr int kWfSkipMinLanes = VT_WF_SKIP_MIN_LANES;     // lanes that must want an empty-space skip (or half of the running ones) before the warp pays for one
constexpr int kWfQueueChunk = 64;                         // shade-queue entries a trace warp reserves per atomic
constexpr int kWfTraceThreads = 256;
constexpr unsigned int kWfPark = 64;                      // ring entries per trace warp (a power of two >= 2 * 32 - 1)
constexpr unsigned int kWfInvalid = 0xffffffffu;          // slot word of an unused queue entry
constexpr int kWfMaxBounces = 511;                        // the bounce count travels in 9 bits of the ray record

enum { WF_RAY_SHADOW = 0, WF_RAY_BOUNCE = 1, WF_RAY_PRIMARY = 2 };
// flags of a queue entry (upper half of its third word): hit kind, primary, NaN mask of the hit position, bounce count
enum { WF_HIT_VOXEL = 1, WF_HIT_GROUND = 2, WF_HIT_PRIMARY = 4, WF_HIT_NAN_SHIFT = 3, WF_HIT_BOUNCE_SHIFT = 6 };

// counts block of one iteration (device memory, zeroed once per batch)
struct WfCounts {
    unsigned int tq;                  // paths of this generation = slots in use = ray records per region
    unsigned int sq[kWfQueues];       // entries (valid or not) in the shade queues holding this generation's traced paths
    unsigned int work;                // ray hand-out counter of wf_trace
    unsigned int pad;
};

// All per-batch storage. Generation g of the paths lives in the arrays of parity g & 1.
struct WfState {
    // path record, 64 B = 4 float4 per slot: (origin xyz, dir x) (dir y, dir z, bsdf pdf, radiance x)
    // (radiance y z, throughput x y) (throughput z, pending light xyz). A primary path has only the first two.
    float4* __restrict__ state[2];
    // ray records of a generation, indexed by the path's slot: region 0 = shadow rays, region 1 = bounce / primary rays.
    //   a: (ix & 0xffff) | iy << 16,  (iz & 0xffff) | kind << 16 | status << 18 | NaN mask << 20 | bounces << 23,
    //      shadow: light target (linear voxel index, -1 = environment) / else: path id,  rng word
    //   b: dis xyz, 1/d x      c: 1/d y z          (the DDA state after dda.h:16-34; the step direction is the sign of 1/d)
    // status = DDA_RUNNING for a ray to be traced; a ray that dda_begin already resolved (start voxel outside the grid or NaN)
    // carries its final status and position: wf_trace routes it without stepping, so every queue append happens in one place.
    int4* __restrict__ rq_a;                  // 2 * rq_cap records each: region r starts at r * rq_cap
    float4* __restrict__ rq_b;
    float2* __restrict__ rq_c;
    unsigned int rq_cap;
    unsigned int* __restrict__ vis;           // shadow-ray results of the generation being traced: bit slot & 31 of word slot >> 5, 1 = light visible (zeroed before every trace)
    // shade queues: (slot, hit ix | iy << 16, hit iz | flags << 16, path id) + the rng word (rng linear offset | pending NaN mask << 29)
    int4* __restrict__ sq[kWfQueues];
    int* __restrict__ sq_rng[kWfQueues];
    float4* __restrict__ samples;             // tone-mapped sample per path id (P * n_items)
    int n_items;                              // paths per pass (tiles * 4096)
};

VT_DEV float i2f(int i) { return __int_as_float(i); }
VT_DEV int f2bits(float f) { return __float_as_int(f); }
VT_DEV int wf_pack16(int lo, int hi) { return (lo & 0xffff) | (hi << 16); }
VT_DEV int wf_lo16(int w) { return (w << 16) >> 16; }            // sign-extending
VT_DEV int wf_hi16(int w) { return w >> 16; }

// item -> pixel: 64x64 tiles dealt round-robin over ranks, 8x4 pixel blocks inside a tile
VT_DEV bool wf_item_pixel(const Frame& F, const RenderLaunch& L, int item, int& px, int& py)
{
    const int local_tile = item >> 12, in_tile = item & 4095;
    const int tile = L.tile_rank + local_tile * L.tile_world;
    const int tx = tile % L.tiles_x, ty = tile / L.tiles_x;
    const int blk = in_tile >> 5, within = in_tile & 31;
    px = tx * kTile + (blk & 7) * 8 + (within & 7);
    py = ty * kTile + (blk >> 3) * 4 + (within >> 3);
    return px < F.W && py < F.H;
}

VT_DEV f3 wf_hit_pos(int ix, int iy, int iz, int flags)
{
    const float qn = __int_as_float(0x7fc00000);
    const int nm = (flags >> WF_HIT_NAN_SHIFT) & 7;
    return mk3((nm & 1) ? qn : (float)ix, (nm & 2) ? qn : (float)iy, (nm & 4) ? qn : (float)iz);
}

// rng offset <-> linear index into the noise table (rand() walks the table row by row, random.h:20-27)
VT_DEV int wf_rng_pack(const Frame& F, int2 off) { return off.x + off.y * F.noise_w; }
VT_DEV int2 wf_rng_unpack(const Frame& F, int idx)
{
    const int w = F.noise_w;
    if ((w & (w - 1)) == 0) return make_int2(idx & (w - 1), idx >> (__ffs(w) - 1));
    const int y = idx / w;
    return make_int2(idx - y * w, y);
}

// flags of a finished traversal (dda.h:63-79)
VT_DEV int wf_hit_flags(int status, const Dda& s)
{
    const bool ground = (status != DDA_HIT) && !(s.nanmask & 2) && (s.iy() < 0);     // dda.h:75-78
    return (status == DDA_HIT ? WF_HIT_VOXEL : (ground ? WF_HIT_GROUND : 0)) | (s.nanmask << WF_HIT_NAN_SHIFT);
}
// pathTracer.fs:134-153: is the sampled light visible, given the end of the shadow traversal
VT_DEV int wf_light_visible(const Volume& V, int target, int status, const Dda& s)
{
    const int flags = wf_hit_flags(status, s);
    if (target < 0) return (flags & 3) == 0;                                            // environment: nothing in the way
    // emissive voxel: the traversal must end on exactly that voxel (a ground or NaN position never equals it)
    return status == DDA_HIT && (flags >> WF_HIT_NAN_SHIFT) == 0 && (s.ix() + s.iy() * V.X + s.iz() * V.X * V.Y) == target;
}

// Where does a path go whose primary / bounce ray has just ended (pathTracer.fs:202-208, :214, :282-291)? Surface hit with
// bounces left: the shade queue of the hit voxel's material type; everything else: the finish queue. `bounces` is the count
// BEFORE the increment of :291 (-1 for a primary ray: the loop of :214 has not been entered yet).
VT_DEV int wf_route(const Volume& V, const Frame& F, int flags, int bounces, int ix, int iy, int iz)
{
    if ((flags & 3) == 0 || !(bounces + 1 < F.max_bounces)) return 0;
    const int type = f2i(fetch_mat(F, fetch_offset(V, ix, iy, iz)));
    return (type >= 0 && type <= 2) ? 1 + type : 4;
}
VT_DEV int4 wf_entry(unsigned int slot, const Dda& s, int flags, unsigned int pid)
{
    // a NaN component holds cvt.rzi(NaN) = 0; finite components stay within [-1, res] (a step moves one voxel): 16 bits each
    return make_int4((int)slot, wf_pack16(s.ix(), s.iy()), wf_pack16(s.iz(), flags), (int)pid);
}

VT_DEV void wf_store_ray(const WfState& S, int region, unsigned int slot, int status, const Dda& s, int kind, int bounces, int aux, int rngw)
{
    const size_t at = (size_t)slot + (region ? (size_t)S.rq_cap : 0);
    st_stream16(S.rq_a + at, make_int4(wf_pack16(s.ix(), s.iy()), (s.iz() & 0xffff) | (kind << 16) | (status << 18) | (s.nanmask << 20) | (bounces << 23), aux, rngw));
    st_stream16(S.rq_b + at, make_float4(s.dx, s.dy, s.dz, s.pos_x() ? s.ex : -s.ex));     // |1/d| > 0 (dda_begin): the sign bit is free
    st_stream8(S.rq_c + at, make_float2(s.pos_y() ? s.ey : -s.ey, s.pos_z() ? s.ez : -s.ez));
}

// Slot reservation in the next generation, aggregated lanes -> warp (ballot) -> CTA (shared-memory atomic) -> one global
// atomic per CTA: same-address global atomics serialise in L2 at about one per clock. ALL threads of the CTA must call.
struct WfBlockCounter { unsigned int cnt, base; };
VT_DEV void wf_reserve_init(WfBlockCounter& sm)
{
    if (threadIdx.x == 0) sm.cnt = 0u;
    __syncthreads();
}
VT_DEV unsigned int wf_reserve_slot(WfCounts* __restrict__ cnt, bool want, WfBlockCounter& sm)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    // sm.cnt is zero on entry: wf_reserve_init at kernel start, then thread 0 re-zeroes it below (two barriers per call)
    const unsigned mt = __ballot_sync(full, want);
    unsigned int wt = 0;
    if (mt != 0u) {
        if (lane == 0) wt = atomicAdd(&sm.cnt, (unsigned)__popc(mt));
        wt = __shfl_sync(full, wt, 0);
    }
    __syncthreads();
    if (threadIdx.x == 0 && sm.cnt != 0u) {
        sm.base = atomicAdd(&cnt->tq, sm.cnt);
        sm.cnt = 0u;                                                        // ready for the next call (visible after the barrier below)
    }
    __syncthreads();
    return sm.base + wt + (unsigned)__popc(mt & lt);
}

// Appends one queue entry per lane with q >= 0 to shade queue q. Positions are reserved per WARP in chunks of kWfQueueChunk
// entries (one global atomic per chunk, no CTA barrier): cur_next / cur_end = this warp's [next, end) per queue (shared memory,
// warp-uniform, touched by lane 0); the old chunk is filled up first, the rest opens a new one. ALL lanes of the warp must call.
VT_DEV void wf_append_entries(const WfState& S, WfCounts* __restrict__ cq, unsigned int* cur_next, unsigned int* cur_end,
                              int q, int4 entry, int rngw)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    unsigned pending = __ballot_sync(full, q >= 0);
    while (pending != 0u) {                                     // one round per queue present: usually one or two
        const int k = __shfl_sync(full, q, __ffs(pending) - 1);
        const unsigned m = __ballot_sync(full, q == k);
        pending &= ~m;
        const unsigned cnt_k = (unsigned)__popc(m);
        const unsigned nx = cur_next[k], avail = cur_end[k] - nx;
        unsigned base2 = 0;
        if (avail < cnt_k) {
            if (lane == 0) base2 = atomicAdd(&cq->sq[k], (unsigned)kWfQueueChunk);
            base2 = __shfl_sync(full, base2, 0);
        }
        if (q == k) {
            const unsigned rank = (unsigned)__popc(m & lt);
            const unsigned dst = rank < avail ? nx + rank : base2 + (rank - avail);
            st_stream16(S.sq[k] + dst, entry); st_stream4(S.sq_rng[k] + dst, rngw);
        }
        __syncwarp();
        if (lane == 0) {
            if (avail < cnt_k) { cur_next[k] = base2 + (cnt_k - avail); cur_end[k] = base2 + (unsigned)kWfQueueChunk; }
            else cur_next[k] = nx + cnt_k;
        }
        __syncwarp();
    }
    __syncwarp();
}
// the unused tail of the warp's chunks: marked invalid (wf_shade skips such entries)
VT_DEV void wf_invalidate_tails(const WfState& S, const unsigned int* cur_next, const unsigned int* cur_end)
{
    const int lane = threadIdx.x & 31;
    #pragma unroll
    for (int k = 0; k < kWfQueues; ++k)
        for (unsigned i = cur_next[k] + (unsigned)lane; i < cur_end[k]; i += 32u)
            S.sq[k][i] = make_int4((int)kWfInvalid, 0, 0, 0);
}

template <bool COUNT>
VT_DEV void wf_flush_tally(const Tally<COUNT>& tl, Counters* __restrict__ counters)
{
    if (COUNT) {
        unsigned long long v[5] = { tl.S, tl.R, tl.H, tl.E, tl.Q };
        #pragma unroll
        for (int i = 0; i < 5; ++i) {
            unsigned long long x = v[i];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            v[i] = x;
        }
        if ((threadIdx.x & 31) == 0) {
            if (v[0]) atomicAdd(&counters->S, v[0]);
            if (v[1]) atomicAdd(&counters->R, v[1]);
            if (v[2]) atomicAdd(&counters->H, v[2]);
            if (v[3]) atomicAdd(&counters->E, v[3]);
            if (v[4]) atomicAdd(&counters->Q, v[4]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// wf_generate: pathTracer.fs:172-208 -- RNG offset, camera ray, slab test, and the WHOLE traversal of the primary ray (dda.h:16-57).
// Primary rays of neighbouring pixels are coherent (the threads of a warp cover an 8x4 pixel block of one pass), so their DDA
// runs here in lockstep at nearly full width; this spares a 40-byte ray record, its trip through wf_trace's refill / retire
// machinery and the queue hand-over for a third of all rays (round 2: generate + first trace 10.2 -> see DESIGN.md per 64-pass
// batch). The finished primary path is routed to its shade queue right away (:202-208, :214).
// grid = (n_items / 256, rows): a thread generates every rows-th pass of its pixel; rows = 1 whenever the frame alone fills
// the machine. Pixels whose ray misses the volume's box are finished here; every other pixel gets a slot of generation 0.
// ---------------------------------------------------------------------------------------------------------
#ifndef VT_WF_GENERATE_MIN_BLOCKS
#define VT_WF_GENERATE_MIN_BLOCKS 3
#endif
template <bool COUNT, bool SKIP>
VT_GLOBAL void __launch_bounds__(256, VT_WF_GENERATE_MIN_BLOCKS)
wf_generate_kernel(const Volume V, const Frame F, const RenderLaunch L, const WfState S, int pass0, int n_batch,
                   WfCounts* __restrict__ cnt, int* __restrict__ primary, Counters* __restrict__ counters)
{
    // per-warp cursors into the shade queues, as in wf_trace
    __shared__ unsigned int cur_next[256 / 32][kWfQueues], cur_end[256 / 32][kWfQueues];
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) < kWfQueues) { cur_next[warp][threadIdx.x & 31] = 0u; cur_end[warp][threadIdx.x & 31] = 0u; }
    __syncwarp();
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    Tally<COUNT> tl; tl.clear();
    int px = 0, py = 0;
    const bool mine = item < S.n_items && wf_item_pixel(F, L, item, px, py);
    const f3 frag = mk3((float)px + 0.5f, (float)py + 0.5f, 0.55f);
    // thin lens: what does not depend on the sample is computed once per pixel for all the passes this thread generates
    f4 lens_fp = mk4(0.f, 0.f, 0.f, 0.f);
    if (mine && F.lens_model == 1) lens_fp = thin_lens_focal_point(F, frag);
    const int pixel_hash = rng_pixel_hash(px, py, F.noise_w);                      // random.h:15, the same in every pass
    // the passes of the batch are dealt to gridDim.y thread rows; every thread of a CTA runs the same number of iterations
    for (int pass_local = blockIdx.y; pass_local < n_batch; pass_local += gridDim.y) {
        const unsigned int pid = (unsigned)pass_local * (unsigned)S.n_items + (unsigned)item;
        int status = DDA_NOHIT;
        bool valid = false;
        f3 ro = mk3(0.f), rd = mk3(0.f);
        int2 rng = make_int2(0, 0);
        Dda s;
        s.set_pos(0, 0, 0); s.nanmask = 0; s.steps = 0; s.bkey = kNoBrick; s.brick = 0ull;
        s.set_sign(1, 1, 1); s.dx = s.dy = s.dz = s.ex = s.ey = s.ez = 0.f;
        if (mine) {
            valid = true;
            const int sample = L.first_sample + (pass0 + pass_local) * L.sample_stride;
            rng = rng_offset_from(pixel_hash, sample, F.noise_w, F.noise_h);           // :174
            if (F.lens_model == 1) {                                                   // :179, generateRay.h:30-101
                const f4 u = rng_next<COUNT>(F, rng, tl);
                thin_lens_ray(F, lens_fp, u.x, u.y, ro, rd);
            } else generate_ray<COUNT>(F, frag, rng, ro, rd, tl);
            const float t = ray_aabb(ro, rd, V.bmin, V.bmax);                          // :183
            if (!(t < 0.0f)) {
                status = dda_begin<COUNT>(V, ro + t * rd, rd, s, tl);                  // :196-202
            } else {
                // :187-194 the ray misses the volume's box: finished here. Such pixels come in whole 8x4 blocks (the sky), so the
                // warp does not diverge, and the path never costs a slot, a queue entry or a pass through wf_shade.
                const f3 c = tonemap(background_color<COUNT>(F, rd, tl));
                st_stream16(S.samples + pid, make_float4(c.x, c.y, c.z, 1.0f));
                valid = false;
            }
            if (primary != nullptr && pass0 + pass_local == L.n_passes - 1) primary[(size_t)px + (size_t)py * (size_t)F.W] = -1;
        }
        // ---- the primary traversal, in lockstep (dda.h:38-57) ----------------------------------------------------
        {
            int guard = (V.X + V.Y + V.Z) / 16 + 8;                                    // belt and braces: see dda_begin on why rays always leave
            for (;;) {
                const bool running = valid && status == DDA_RUNNING;
                if (__ballot_sync(full, running) == 0u) break;
                if (running) {
                    #pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        status = dda_step<COUNT>(V, s, tl);
                        if (status != DDA_RUNNING) break;
                    }
                    if (--guard < 0 && status == DDA_RUNNING) status = DDA_NOHIT;
                }
                if (SKIP && !COUNT) {                                                  // exact empty-space skip, as in wf_trace
                    const bool run2 = valid && status == DDA_RUNNING;
                    const int radius = run2 ? dda_skip_radius(V, s) : 0;
                    const unsigned m_running = __ballot_sync(full, run2), m_want = __ballot_sync(full, dda_skip_wanted(radius));
                    if (m_want != 0u && (__popc(m_want) >= kWfSkipMinLanes || 2 * __popc(m_want) >= __popc(m_running))) {
                        if (dda_skip_wanted(radius)) (void)dda_skip<true>(V, s, radius);
                    }
                }
            }
        }
        // ---- routing (:202-208, :214): one shade-queue entry per path. The slot of a generation-0 path is its path id: no
        // reservation, no CTA barrier (the barriers of a CTA-wide reservation were 12 % of this kernel's stall samples,
        // profiles/r02_v21_wf_generate_*), and the records of a warp are written side by side. Nothing walks generation 0 by
        // slot -- wf_shade reaches its records through the queue entries -- so the slots of sky pixels simply stay unused.
        int q = -1, flags = 0;
        if (valid) {
            flags = wf_hit_flags(status, s) | WF_HIT_PRIMARY;
            q = wf_route(V, F, flags, -1, s.ix(), s.iy(), s.iz());
            // a primary path has radiance 0, throughput 1, nothing pending, bounce 0, no BSDF pdf: wf_shade synthesises all of that
            // from the PRIMARY flag, and only the first sector of the record is written
            st_stream32(S.state[0] + 4 * (size_t)pid, make_float4(ro.x, ro.y, ro.z, rd.x), make_float4(rd.y, rd.z, 0.0f, 0.0f));
        }
        wf_append_entries(S, cnt + 1, cur_next[warp], cur_end[warp], q, wf_entry(pid, s, flags, pid), wf_rng_pack(F, rng));
    }
    wf_invalidate_tails(S, cur_next[warp], cur_end[warp]);
    wf_flush_tally<COUNT>(tl, counters);
}

// ---------------------------------------------------------------------------------------------------------
// wf_trace: the loop of dda.h:38-57 for the ray records of one generation, then the routing of the finished path.
// Persistent warps; a warp reserves up to kWfGrab rays per atomic (a static round-robin deal of the blocks was 20 % slower: every
// launch then waits for its unluckiest warp) and its lanes refill from that range whenever fewer than kWfLiveMin of them
// still run a ray. Ray index w in [0, 2n): w < n = shadow ray of slot w, else bounce / primary ray of slot w - n.
// Finished bounce / primary rays are not routed by the few lanes that happen to retire in a round: they park (slot, hit
// voxel, status) in a per-warp shared-memory ring, and whenever 32 are parked the whole warp routes and appends them at
// full width (wf_trace_flush).
// ---------------------------------------------------------------------------------------------------------
#ifndef VT_WF_TRACE_MIN_BLOCKS
#define VT_WF_TRACE_MIN_BLOCKS 6
#endif
template <bool COUNT, bool SKIP>
VT_GLOBAL void __launch_bounds__(kWfTraceThreads, VT_WF_TRACE_MIN_BLOCKS)
wf_trace_kernel(const Volume V, const Frame F, const WfState S,
                WfCounts* __restrict__ cnt, WfCounts* __restrict__ cnext, Counters* __restrict__ counters)
{
    // per-warp cursors into the shade queues: [next, end) of the chunk the warp is filling (warp-uniform, touched by lane 0)
    __shared__ unsigned int cur_next[kWfTraceThreads / 32][kWfQueues], cur_end[kWfTraceThreads / 32][kWfQueues];
    // per-warp ring of finished bounce / primary rays waiting to be routed: slot, hit x | y << 16, hit z | status << 16
    __shared__ unsigned int park_slot[kWfTraceThreads / 32][kWfPark];
    __shared__ int park_xy[kWfTraceThreads / 32][kWfPark], park_zs[kWfTraceThreads / 32][kWfPark];
    unsigned int park_head = 0, park_cnt = 0;       // warp-uniform
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    if (lane < kWfQueues) { cur_next[warp][lane] = 0u; cur_end[warp][lane] = 0u; }
    __syncwarp();
    const unsigned int n = cnt->tq;                                   // slots of this generation
    const unsigned int w_total = 2u * n;
    // rays per hand-out: kWfGrab when the launch is large (fewer same-address atomics), less when it is small, so that every warp of
    // the grid still gets about four ranges (a 1-pass frame has 0.5 M rays for 7 104 warps)
    const unsigned int grab = max(32u, min((unsigned)kWfGrab, (w_total / (gridDim.x * (kWfTraceThreads / 32) * 4u)) & ~31u));
    unsigned int* __restrict__ vis = S.vis;
    Tally<COUNT> tl; tl.clear();

#ifdef VT_SKIP_STATS
    unsigned long long dbg_calls = 0, dbg_ok = 0, dbg_steps = 0, dbg_plain = 0, dbg_long = 0;
#endif
    bool have = false, exhausted = false;
    unsigned int range_next = 0, range_end = 0;     // warp-uniform
    unsigned int w = 0;                             // ray index of the lane's ray
    int status = DDA_NOHIT, guard = 0;
    constexpr int kChunk = (SKIP && !COUNT) ? kWfStepChunkSkip : kWfStepChunk;
    const int chunk_guard = (V.X + V.Y + V.Z) / kChunk + 8;         // belt and braces: see dda_begin on why rays always leave
    Dda s;
    s.set_pos(0, 0, 0); s.nanmask = 0; s.steps = 0; s.bkey = kNoBrick; s.brick = 0ull;
    s.dx = s.dy = s.dz = 0.f; s.ex = s.ey = s.ez = 0.f; s.set_sign(1, 1, 1);

    // routes up to 32 parked paths at full width: looks at the ray record again (kind, bounce count, path id, rng word), finds
    // the queue (pathTracer.fs:202-208, :214, :282-291) and appends the entries; slots are reserved per warp in chunks
    auto route_parked = [&]() {
        const unsigned nf = min(park_cnt, 32u);
        int q = -1, rngw = 0;
        int4 entry = make_int4(0, 0, 0, 0);
        if ((unsigned)lane < nf) {
            const unsigned at = (park_head + (unsigned)lane) & (kWfPark - 1);
            const unsigned slot = park_slot[warp][at];
            const int xy = park_xy[warp][at], zs = park_zs[warp][at];
            const int4 a = S.rq_a[(size_t)slot + (size_t)S.rq_cap];   // kind, NaN mask, bounce count, path id, rng word
            const int st = zs >> 16, hy = wf_hi16(xy), nanmask = (a.y >> 20) & 7;
            const int kind = (a.y >> 16) & 3, bounces = (int)((unsigned)a.y >> 23);
            const bool ground = (st != DDA_HIT) && !(nanmask & 2) && (hy < 0);                     // dda.h:75-78
            const int flags = (st == DDA_HIT ? WF_HIT_VOXEL : (ground ? WF_HIT_GROUND : 0)) | (nanmask << WF_HIT_NAN_SHIFT)
                              | (kind == WF_RAY_PRIMARY ? WF_HIT_PRIMARY : 0) | (bounces << WF_HIT_BOUNCE_SHIFT);
            q = wf_route(V, F, flags, kind == WF_RAY_PRIMARY ? -1 : bounces, wf_lo16(xy), hy, wf_lo16(zs));
            entry = make_int4((int)slot, xy, (zs & 0xffff) | (flags << 16), a.z);
            rngw = a.w;
        }
        park_head = (park_head + nf) & (kWfPark - 1); park_cnt -= nf;
        // routed paths -> shade queues, slots reserved per warp in chunks
        wf_append_entries(S, cnext, cur_next[warp], cur_end[warp], q, entry, rngw);
    };

    for (;;) {
        // ---- refill idle lanes ---------------------------------------------------------------------------
        const unsigned need = __ballot_sync(full, !have);
        if (!exhausted && need != 0u) {
            if (range_next >= range_end) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&cnt->work, grab);
                base = __shfl_sync(full, base, 0);
                range_next = base;
                range_end = min(base + grab, w_total);
                if (base >= w_total) { exhausted = true; range_end = range_next = 0; }
            }
            if (!have) {
                const unsigned r = range_next + (unsigned)__popc(need & lt);
                if (r < range_end) {
                    const size_t at = r < n ? (size_t)r : (size_t)(r - n) + (size_t)S.rq_cap;
                    const int4 a = S.rq_a[at];                           // re-read at retire: default caching
                    const float4 b = ld_stream16(S.rq_b + at); const float2 c = ld_stream8(S.rq_c + at);
                    s.set_pos(wf_lo16(a.x), wf_hi16(a.x), wf_lo16(a.y));
                    s.dx = b.x; s.dy = b.y; s.dz = b.z; s.ex = gabs(b.w); s.ey = gabs(c.x); s.ez = gabs(c.y);
                    s.set_sign(f2bits(b.w) < 0 ? -1 : 1, f2bits(c.x) < 0 ? -1 : 1, f2bits(c.y) < 0 ? -1 : 1);
                    s.steps = 0; s.bkey = kNoBrick;
                    status = (a.y >> 18) & 3;                            // DDA_RUNNING, or the final status dda_begin found
                    guard = chunk_guard;
                    w = r;
                    have = true;
                }
            }
            range_next = min(range_next + (unsigned)__popc(need), range_end);
        }
        if (__ballot_sync(full, have) == 0u) { if (exhausted) break; else continue; }
        // ---- the hot loop ----------------------------------------------------------------------------------
        for (;;) {
            if (have && status == DDA_RUNNING) {   // lanes leave the chunk through `break`: one reconvergence point per chunk, not per step
                #pragma unroll
                for (int k = 0; k < kChunk; ++k) {
                    status = dda_step<COUNT>(V, s, tl);
#ifdef VT_SKIP_STATS
                    dbg_plain += 1;
#endif
                    if (status != DDA_RUNNING) break;
                }
                if (--guard < 0 && status == DDA_RUNNING) status = DDA_NOHIT;
            }
            if (SKIP && !COUNT) {                 // counting builds step every voxel so that S stays the algorithmic count
                // the cheap part (one byte per lane) runs converged; the skip itself only when enough lanes want it, so that its
                // set-up is not paid for one or two lanes while the rest of the warp idles
                const bool running = have && status == DDA_RUNNING;
                const int radius = running ? dda_skip_radius(V, s) : 0;
                const unsigned m_running = __ballot_sync(full, running), m_want = __ballot_sync(full, dda_skip_wanted(radius));
                if (m_want != 0u && (__popc(m_want) >= kWfSkipMinLanes || 2 * __popc(m_want) >= __popc(m_running))) {
                    if (dda_skip_wanted(radius)) {
                        const int skipped = dda_skip<false>(V, s, radius);
#ifdef VT_SKIP_STATS
                        dbg_calls += 1; dbg_ok += skipped > 0; dbg_steps += skipped; dbg_long += skipped >= 32;
#else
                        (void)skipped;
#endif
                    }
                }
            }
            // keep stepping while enough lanes still run a ray; once all rays are handed out (no refill can follow), until none does
            const unsigned m_run = __ballot_sync(full, have && status == DDA_RUNNING);
            if (exhausted ? (m_run == 0u) : (__popc(m_run) < kWfLiveMin)) break;
        }
        // ---- retire finished rays --------------------------------------------------------------------------
        bool park = false;
        unsigned int p_slot = 0; int p_xy = 0, p_zs = 0;
        if (have && status != DDA_RUNNING) {
            if (w < n) {                               // shadow ray of slot w
                if (F.n_emissive == 0) {
                    // towards the environment (the only light there is): visible iff nothing was hit, the virtual ground included
                    // (a NaN height is stored as 0, so the ground test needs no NaN mask). No second look at the record.
                    if (status != DDA_HIT && !(s.iy() < 0)) atomicOr(vis + (w >> 5), 1u << (w & 31));
                } else {
                    const int4 a = S.rq_a[w];                           // light target; NaN mask of a ray resolved by dda_begin
                    s.nanmask = (a.y >> 20) & 7;
                    if (wf_light_visible(V, a.z, status, s)) atomicOr(vis + (w >> 5), 1u << (w & 31));
                    s.nanmask = 0;
                }
            } else {                                   // bounce / primary ray: parked, routed later at full width
                park = true;
                p_slot = w - n; p_xy = wf_pack16(s.ix(), s.iy()); p_zs = (s.iz() & 0xffff) | (status << 16);
            }
            have = false;
        }
        const unsigned m_park = __ballot_sync(full, park);
        if (m_park != 0u) {
            if (park) {
                const unsigned at = (park_head + park_cnt + (unsigned)__popc(m_park & lt)) & (kWfPark - 1);
                park_slot[warp][at] = p_slot; park_xy[warp][at] = p_xy; park_zs[warp][at] = p_zs;
            }
            park_cnt += (unsigned)__popc(m_park);
            __syncwarp();
        }
        // ---- route 32 parked paths: every lane takes one (and once more after the loop for what is left) -------
        if (park_cnt >= 32u) route_parked();
    }
    if (park_cnt != 0u) route_parked();
    wf_invalidate_tails(S, cur_next[warp], cur_end[warp]);
#ifdef VT_SKIP_STATS
    if (SKIP) { atomicAdd(&counters->E, dbg_calls); atomicAdd(&counters->Q, dbg_ok); atomicAdd(&counters->H, dbg_steps); atomicAdd(&counters->S, dbg_plain); atomicAdd(&counters->R, dbg_long); }
#endif
    wf_flush_tally<COUNT>(tl, counters);
}

// ---------------------------------------------------------------------------------------------------------
// wf_primary: primary-hit codes of the frame's last pass (vt_read_primary_hits; off in production renders). Reads the
// shade queues filled by the first trace: every primary path that entered the volume's box has exactly one entry.
// ---------------------------------------------------------------------------------------------------------
VT_GLOBAL void __launch_bounds__(256)
wf_primary_kernel(const Volume V, const Frame F, const RenderLaunch L, const WfState S, int pass0,
                  const WfCounts* __restrict__ cq, int* __restrict__ primary)
{
    const int k = blockIdx.y;
    const unsigned int n = cq->sq[k];
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 e = S.sq[k][i];
        if ((unsigned)e.x == kWfInvalid) continue;
        const int flags = (int)((unsigned)e.z >> 16);
        if (!(flags & WF_HIT_PRIMARY) || (flags & 3) == 0) continue;
        const unsigned int pid = (unsigned)e.w;
        const int pass_local = (int)(pid / (unsigned)S.n_items), item = (int)(pid - (unsigned)pass_local * (unsigned)S.n_items);
        int px, py;
        if (pass0 + pass_local == L.n_passes - 1 && wf_item_pixel(F, L, item, px, py))
            primary[(size_t)px + (size_t)py * (size_t)F.W] =
                hit_code(V, wf_hit_pos(wf_lo16(e.y), wf_hi16(e.y), wf_lo16(e.z), flags), (flags & WF_HIT_GROUND) != 0);
    }
}

// ---------------------------------------------------------------------------------------------------------
// wf_shade: one loop iteration of pathTracer.fs:214-292 up to (not including) the two traversals, preceded by
// the tail of the previous iteration (shadow-ray result :134-164, environment on a miss :282-289, bounces++).
// Queue 0 holds the paths that end here; queues 1..4 the surface hits sorted by material type.
// cnt = the counts block whose queues are consumed and whose `tq` counts the next generation.
// ---------------------------------------------------------------------------------------------------------
template <bool COUNT>
VT_GLOBAL void __launch_bounds__(VT_WF_SHADE_THREADS, VT_WF_SHADE_MIN_BLOCKS)
wf_shade_kernel(const Volume V, const Frame F, const WfState S, const int gen, WfCounts* __restrict__ cnt, Counters* __restrict__ counters)
{
    __shared__ WfBlockCounter sm;
    wf_reserve_init(sm);
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const float4* __restrict__ st_in = S.state[gen];
    float4* __restrict__ st_out = S.state[gen ^ 1];
    const unsigned int* __restrict__ vis_in = S.vis;
    Tally<COUNT> tl; tl.clear();
    // chunks of 32 entries, queue after queue; a CTA takes warps_per_cta consecutive chunks per iteration
    unsigned int n_q[kWfQueues], chunks_before[kWfQueues + 1];
    chunks_before[0] = 0;
    #pragma unroll
    for (int k = 0; k < kWfQueues; ++k) { n_q[k] = cnt->sq[k]; chunks_before[k + 1] = chunks_before[k] + ((n_q[k] + 31u) >> 5); }
    const int sel_x = F.shared->sel_index[0], sel_y = F.shared->sel_index[1], sel_z = F.shared->sel_index[2];

    for (unsigned int chunk0 = blockIdx.x * warps_per_cta; chunk0 < chunks_before[kWfQueues]; chunk0 += gridDim.x * warps_per_cta) {
        const unsigned int chunk = chunk0 + (threadIdx.x >> 5);
        unsigned int cb = 0, nq = n_q[0];
        const int4* __restrict__ qp = S.sq[0];
        const int* __restrict__ qr = S.sq_rng[0];
        #pragma unroll
        for (int j = 1; j < kWfQueues; ++j) if (chunk >= chunks_before[j]) { cb = chunks_before[j]; nq = n_q[j]; qp = S.sq[j]; qr = S.sq_rng[j]; }
        const unsigned int idx = ((chunk - cb) << 5) + (unsigned)lane;
        bool valid = chunk < chunks_before[kWfQueues] && idx < nq;
        int4 e = make_int4((int)kWfInvalid, 0, 0, 0);
        int rngw = 0;
        if (valid) { e = ld_stream16(qp + idx); rngw = ld_stream4(qr + idx); }
        valid = valid && (unsigned)e.x != kWfInvalid;
        bool continues = false;
        unsigned int pid = 0;
        int st_a = DDA_NOHIT, st_b = DDA_NOHIT, target = -1, bounces_out = 0, rngw_out = 0;
        float4 o0, o1, o2, o3;
        o0 = o1 = o2 = o3 = make_float4(0.f, 0.f, 0.f, 0.f);
        Dda sa, sb;
        sa.set_pos(0, 0, 0); sa.nanmask = 0; sa.set_sign(1, 1, 1); sa.dx = sa.dy = sa.dz = sa.ex = sa.ey = sa.ez = 0.f;
        sb = sa;
        if (valid) {
            const unsigned int slot = (unsigned)e.x;
            pid = (unsigned)e.w;
            const int hx = wf_lo16(e.y), hy = wf_hi16(e.y), hz = wf_lo16(e.z);
            const int flags = (int)((unsigned)e.z >> 16);
            const bool is_primary = (flags & WF_HIT_PRIMARY) != 0;
            const f8 r01 = ld_stream32(st_in + 4 * (size_t)slot);
            f8 r23; r23.lo = r23.hi = make_float4(0.f, 0.f, 0.f, 0.f);
            int vis = 0;
            if (!is_primary) {
                r23 = ld_stream32(st_in + 4 * (size_t)slot + 2);
                vis = (int)((vis_in[slot >> 5] >> (slot & 31)) & 1u);
            }
            int2 rng = wf_rng_unpack(F, rngw & 0x1fffffff);
            const int pn = (int)((unsigned)rngw >> 29);                            // NaN mask of throughput * 0 (the hidden light's contribution)
            {   // gathers whose addresses are known now but whose values are needed hundreds of instructions later
                if ((unsigned)rng.x < (unsigned)F.noise_w && (unsigned)rng.y < (unsigned)F.noise_h) prefetch_keep(F.noise + ((size_t)rng.x + (size_t)rng.y * (size_t)F.noise_w));
                if ((unsigned)hx < (unsigned)V.X && (unsigned)hy < (unsigned)V.Y && (unsigned)hz < (unsigned)V.Z)
                    prefetch_id(V, (size_t)hx + (size_t)hy * (size_t)V.X + (size_t)hz * (size_t)V.X * (size_t)V.Y);
            }
            const f3 ro = mk3(r01.lo.x, r01.lo.y, r01.lo.z), rd = mk3(r01.lo.w, r01.hi.x, r01.hi.y);
            const float bsdf_pdf = r01.hi.z;
            f3 radiance = is_primary ? mk3(0.0f) : mk3(r01.hi.w, r23.lo.x, r23.lo.y);
            f3 throughput = is_primary ? mk3(1.0f) : mk3(r23.lo.z, r23.lo.w, r23.hi.x);
            int bounces = is_primary ? 0 : (flags >> WF_HIT_BOUNCE_SHIFT);
            const bool surface = (flags & 3) != 0;
            bool finished = false;
            // the environment seen by a ray that left the scene, primary (:187-208) or bounce (:282-289): ONE expansion of the
            // lat-long mapping + bilinear lookup for both uses (this kernel is bound by instruction fetch, see vt_math.cuh)
            f3 bg = mk3(0.0f);
            if (!surface) bg = background_color<COUNT>(F, rd, tl);
            if (is_primary) {
                if (!surface) {                                                    // :187-194, :202-208
                    radiance = bg;
                    finished = true;
                } else if (!(0 < F.max_bounces)) finished = true;                  // :214 never entered
            } else {
                // pathTracer.fs:248 with the shadow-ray result of the previous iteration
                if (vis != 0) {
                    radiance = radiance + mk3(r23.hi.y, r23.hi.z, r23.hi.w);
                    VT_TALLY(H, 1);                                                // the BSDF evaluation of :161
                } else {
                    const float qn = __int_as_float(0x7fc00000);                   // radiance + throughput * vec3(0)
                    if (pn & 1) radiance.x = qn;
                    if (pn & 2) radiance.y = qn;
                    if (pn & 4) radiance.z = qn;
                }
                if (!surface) {                                                    // :282-289 the bounce ray left the scene
                    const f4 Lp = env_with_pdf(F, bg);                                // lights.h:20-33
                    const float mis = power_heuristic(bsdf_pdf, Lp.w);
                    radiance = radiance + (throughput * xyz(Lp)) * mis;
                    finished = true;
                } else {
                    bounces++;                                                     // :291
                    if (!(bounces < F.max_bounces)) finished = true;               // :214
                }
            }
            if (!finished) {
                const f3 hit = wf_hit_pos(hx, hy, hz, flags);
                Basis hb;
                voxel_to_world(V, hit, ro, rd, hb);                                // :221-223
                const int mat_off = fetch_offset(V, hx, hy, hz);                   // :225-226
                if (hx == sel_x && hy == sel_y && hz == sel_z) {                   // :228-233
                    radiance = radiance + mk3(1.0f, 0.0f, 0.0f);
                    finished = true;
                } else {
                    const f3 wo = -rd;                                             // :237
                    const f3 lsWo = world_to_local(wo, hb);
                    if (bounces == 0) radiance = radiance + throughput * emission_material<COUNT>(F, mat_off, tl);   // :241-245
                    const LightSample ls = sample_light<COUNT>(V, F, hb, rng, tl); // :248 -> :69-124
                    Tally<false> untallied; untallied.clear();    // the reference evaluates the BSDF only for visible lights (:155-161)
                    const f3 pending = throughput * light_contribution<false>(F, mat_off, hb, wo, ls, untallied);
                    const f3 tz = throughput * 0.0f;
                    const int pending_nan = (tz.x != tz.x ? 1 : 0) | (tz.y != tz.y ? 2 : 0) | (tz.z != tz.z ? 4 : 0);
                    f4 bf;
                    const f3 lsWi = sample_material<COUNT>(F, mat_off, lsWo, rng, bf, tl);   // :255
                    if (F.wire_opacity > 0.0f) {                                   // :260-270
                        const float w = wireframe_factor(V, F, hb, hit);
                        bf.x *= w; bf.y *= w; bf.z *= w;
                    }
                    const f3 wi = local_to_world(lsWi, hb);                        // :273
                    throughput = throughput * ((xyz(bf) * gabs(dot(wi, hb.normal))) / bf.w);   // :276
                    o0 = make_float4(hb.position.x, hb.position.y, hb.position.z, wi.x);       // :278-279
                    o1 = make_float4(wi.y, wi.z, bf.w, radiance.x);
                    o2 = make_float4(radiance.y, radiance.z, throughput.x, throughput.y);
                    o3 = make_float4(throughput.z, pending.x, pending.y, pending.z);
                    bounces_out = bounces;
                    rngw_out = wf_rng_pack(F, rng) | (pending_nan << 29);
                    // dda.h:16-34 of the shadow ray (:133) and of the bounce ray (:282)
                    target = ls.target;
                    st_a = dda_begin<COUNT>(V, hb.position, xyz(ls.wl), sa, tl);
                    st_b = dda_begin<COUNT>(V, hb.position, wi, sb, tl);
                    continues = true;
                }
            }
            if (finished) {
                const f3 c = tonemap(radiance);                                    // :294-295
                S.samples[pid] = make_float4(c.x, c.y, c.z, 1.0f);               // default policy: the neighbouring pixel completes the sector soon
            }
        }
        // surviving paths: one slot of the next generation = one path record + the shadow and the bounce ray record at that slot
        const unsigned int tslot = wf_reserve_slot(cnt, continues, sm);
        if (continues) {
            st_stream32(st_out + 4 * (size_t)tslot, o0, o1);
            st_stream32(st_out + 4 * (size_t)tslot + 2, o2, o3);
            wf_store_ray(S, 0, tslot, st_a, sa, WF_RAY_SHADOW, 0, target, 0);
            wf_store_ray(S, 1, tslot, st_b, sb, WF_RAY_BOUNCE, bounces_out, (int)pid, rngw_out);
        }
    }
    wf_flush_tally<COUNT>(tl, counters);
}

// ---------------------------------------------------------------------------------------------------------
// wf_accumulate: accumulation.fs:10-18 over the batch's passes, in pass order. One thread per item.
// ---------------------------------------------------------------------------------------------------------
VT_GLOBAL void __launch_bounds__(256)
wf_accumulate_kernel(const Frame F, const RenderLaunch L, const WfState S, int pass0, int n_batch, float4* __restrict__ accum)
{
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    int px, py;
    if (item >= S.n_items || !wf_item_pixel(F, L, item, px, py)) return;
    const size_t pix = (size_t)px + (size_t)py * (size_t)F.W;
    float4 avg = accum[pix];
    for (int p = 0; p < n_batch; ++p) {
        const float4 s = ld_stream16(S.samples + ((size_t)p * (size_t)S.n_items + (size_t)item));
        if (L.sum_mode) {
            avg.x = avg.x + s.x; avg.y = avg.y + s.y; avg.z = avg.z + s.z; avg.w = avg.w + s.w;
        } else {
            const float n = (float)(L.n_prev + pass0 + p), n1 = (float)(L.n_prev + pass0 + p + 1);
            avg.x = (s.x + avg.x * n) / n1; avg.y = (s.y + avg.y * n) / n1;
            avg.z = (s.z + avg.z * n) / n1; avg.w = (s.w + avg.w * n) / n1;
        }
    }
    accum[pix] = avg;
}

// wf_shade_kernel is instantiated in its own translation unit (vt_shade.cu); these are its host-side entry points
cudaError_t wf_shade_blocks_per_sm(bool count, int* blocks);
void wf_shade_launch(bool count, unsigned int blocks, cudaStream_t st, const Volume& V, const Frame& F, const WfState& S, int gen,
                     WfCounts* cnt, Counters* counters);

} // namespace vt
