// vt_wavefront.cuh -- wavefront path tracer (K1 + K2), render kernel variant 2.
//
// Why: the one-thread-per-pixel megakernel (vt_render_kernel) issues with 9.6 of 32 lanes active (ncu,
// profiles/r01_v1_*): DDA trip counts differ per ray, paths end at different bounces and every lane waits for
// the slowest lane of its warp. The per-lane state machine (vt_pathstate.cuh) fixed the DDA loop but left the
// shading phases at ~4/32 lanes. Memory is idle in both (L2 < 3 %), so the lever is lane utilisation.
//
// How: one progressive batch (P passes over this context's share of the frame) is a set of independent paths
// whose state lives in HBM/L2 as SoA float4 arrays; the integrator loop of pathTracer.fs:214-292 becomes a
// sequence of kernels, each running full warps over a compacted queue of path ids:
//
//   wf_generate   pathTracer.fs:172-196   RNG offset, camera ray, slab test, DDA set-up (dda.h:16-34) of the primary ray
//   wf_trace      dda.h:38-57             the DDA loop alone, for primary, shadow and bounce rays alike: persistent warps
//                                         whose lanes refill themselves from a compact queue of 48-byte ray records as
//                                         soon as their ray ends, so the loop stays >= kWfLiveMin/32 lanes wide although
//                                         ray lengths differ by 100x
//   wf_classify   :202-208, :214, :282-291 routes every traced path: surface hit with bounces left -> the shade queue of
//                                         its material type (Lambert / metal / plastic / other: wf_shade does not diverge
//                                         on the material switch), everything else -> the finish queue
//   repeat max_bounces times:
//     wf_shade    :216-279 (+ :248/:282-289 of the previous iteration) resolve the previous shadow ray, add the
//                                         environment on a miss, build the hit frame, sample the light and the
//                                         BSDF; emits the DDA set-up of one shadow and one bounce ray per surviving path
//     wf_trace, wf_classify
//   wf_shade      (last)                  only the finish queue is populated: resolve, tone-map, write the sample
//   wf_accumulate accumulation.fs:10-18   folds the P samples of every pixel into the running average in pass order
//
// Every path performs exactly the arithmetic of trace_pixel() (vt_device.cuh) in the same order, and the
// per-pixel order of the running average is unchanged: results are bit-identical to variants 0 and 1.
#pragma once
#include "vt_kernels.cuh"

namespace vt {

constexpr int kWfQueues = 5;          // 0 = finish, 1..3 = material type 0..2, 4 = other material types
#ifndef VT_WF_LIVE_MIN
#define VT_WF_LIVE_MIN 28
#endif
#ifndef VT_WF_SHADE_THREADS
#define VT_WF_SHADE_THREADS 128      // threads per wf_shade CTA (64 / 128 / 256 measured: 84.2 / 84.2 / 84.6 ms per step)
#endif
#ifndef VT_WF_SHADE_MIN_BLOCKS
#define VT_WF_SHADE_MIN_BLOCKS (1024 / VT_WF_SHADE_THREADS)
#endif
constexpr int kWfLiveMin = VT_WF_LIVE_MIN;        // refill the warp when fewer lanes than this hold a ray
#ifndef VT_WF_STEP_CHUNK
#define VT_WF_STEP_CHUNK 16
#endif
// DDA iterations between two refill checks. C2 trace ms per 256-spp step with the final 28-instruction step:
// 6 / 8 / 10 / 12 / 14 / 16 / 20 / 24 iterations -> 88.3 / 84.4 / 82.5 / 81.2 / 82.1 / 81.0 / 84.0 / 84.6 (C3 and C4 also prefer 16)
constexpr int kWfStepChunk = VT_WF_STEP_CHUNK;
#ifndef VT_WF_GRAB
#define VT_WF_GRAB 128
#endif
#ifndef VT_WF_SKIP_MIN_LANES
#define VT_WF_SKIP_MIN_LANES 16
#endif
constexpr int kWfGrab = VT_WF_GRAB;                       // rays a warp reserves per atomic on the hand-out counter
constexpr int kWfSkipMinLanes = VT_WF_SKIP_MIN_LANES;     // lanes that must want an empty-space skip (or half of the running ones) before the warp pays for one; C3 trace: 4 / 8 / 12 / 16+ -> 109.5 / 108.3 / 103.0 / 100.3 ms

enum { WF_RAY_SHADOW = 0, WF_RAY_BOUNCE = 1, WF_RAY_PRIMARY = 2 };
enum { WF_HIT_VOXEL = 1, WF_HIT_GROUND = 2, WF_HIT_PRIMARY = 16 };    // hit.w flags (+ nanmask << 8, + the path's bounce count << 16:
                                                                      // wf_classify then needs no other word of the path state)

// counts block of one iteration (device memory, zeroed once per batch)
struct WfCounts {
    unsigned long long tq_rq;         // low 32 bits: paths in the trace list, high 32 bits: rays in the ray queue
    unsigned int sq[kWfQueues];       // entries in the shade queues
    unsigned int work;                // ray hand-out counter of wf_trace
};

// Path state of one generation, SoA, indexed by a COMPACT slot: wf_shade reads generation g by slot and writes the surviving
// paths to consecutive slots of generation g+1 (stream compaction of the state itself, not only of ids), so state traffic
// is coalesced and every fetched sector is fully used. wf_trace writes hit / vis of the generation being traced.
struct WfBuf {
    float4* __restrict__ ray0;        // ray origin xyz, dir x      (the ray whose hit is shaded next)
    float4* __restrict__ ray1;        // dir y, dir z, bsdf pdf, bounces (int bits)
    float4* __restrict__ rad0;        // radiance xyz, throughput x
    float4* __restrict__ rad1;        // throughput y z, pending x y
    float4* __restrict__ rad2;        // pending z, pending_nan (int bits), rng offset x y (int bits)
    int4* __restrict__ hit;           // hit voxel ix iy iz, flags WF_HIT_* | nanmask << 8
    int* __restrict__ vis;            // shadow ray result: 1 = light visible
    unsigned int* __restrict__ pid;   // path id = pass_local * n_items + item (pixel)
};

struct WfState {
    WfBuf buf[2];                     // generations alternate between the two
    float4* __restrict__ samples;     // tone-mapped sample per path id (P * n_items)
    // ray queue: the DDA state after dda.h:16-34, 48 bytes per ray, 2 rays per path at most
    int4* __restrict__ rq0;           // voxel ix iy iz, slot of the path in the generation being traced
    float4* __restrict__ rq1;         // dis xyz, aux (int bits): light target (shadow) / unused
    float4* __restrict__ rq2;         // |1/d| xyz, (int bits) sign bits 0..2 (1 = negative) | ray type << 4
    unsigned int* __restrict__ sq[kWfQueues];   // shade queues: slots of the generation just traced
    int n_items;                      // paths per pass (tiles * 4096)
};

VT_DEV float i2f(int i) { return __int_as_float(i); }
VT_DEV int f2bits(float f) { return __float_as_int(f); }

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

VT_DEV f3 wf_hit_pos(int4 h)
{
    const float qn = __int_as_float(0x7fc00000);
    const int nm = (h.w >> 8) & 7;
    return mk3((nm & 1) ? qn : (float)h.x, (nm & 2) ? qn : (float)h.y, (nm & 4) ? qn : (float)h.z);
}

// queue of the path whose primary / bounce ray ended on a surface
VT_DEV int wf_material_queue(const Volume& V, const Frame& F, int ix, int iy, int iz)
{
    const int off = fetch_offset(V, ix, iy, iz);
    const int type = f2i(fetch_mat(F, off));
    return (type >= 0 && type <= 2) ? 1 + type : 4;
}

// flags of a finished traversal (dda.h:63-79)
VT_DEV int wf_hit_flags(int status, const Dda& s)
{
    const bool ground = (status != DDA_HIT) && !(s.nanmask & 2) && (s.iy < 0);       // dda.h:75-78
    return (status == DDA_HIT ? WF_HIT_VOXEL : (ground ? WF_HIT_GROUND : 0)) | (s.nanmask << 8);
}
// pathTracer.fs:134-153: is the sampled light visible, given the end of the shadow traversal
VT_DEV int wf_light_visible(const Volume& V, int target, int status, const Dda& s)
{
    const int flags = wf_hit_flags(status, s);
    if (target < 0) return (flags & 3) == 0;                                            // environment: nothing in the way
    // emissive voxel: the traversal must end on exactly that voxel (a ground or NaN position never equals it)
    return status == DDA_HIT && (flags >> 8) == 0 && (s.ix + s.iy * V.X + s.iz * V.X * V.Y) == target;
}

VT_DEV void wf_store_ray(const WfState& S, unsigned int rslot, const Dda& s, unsigned int path_slot, int aux, int type)
{
    const unsigned int slot = rslot;
    S.rq0[slot] = make_int4(s.ix, s.iy, s.iz, (int)path_slot);
    S.rq1[slot] = make_float4(s.dx, s.dy, s.dz, i2f(aux));
    S.rq2[slot] = make_float4(s.ex, s.ey, s.ez, i2f((s.sx < 0 ? 1 : 0) | (s.sy < 0 ? 2 : 0) | (s.sz < 0 ? 4 : 0) | (type << 4)));
}

// Block-aggregated queue appends. Same-address global atomics serialise in L2 at roughly one per clock, and a wavefront
// step appends millions of entries, so every append is aggregated twice: lanes -> warp (ballot), warps -> CTA (shared-memory
// atomics), and one global atomicAdd per CTA and counter. ALL threads of the CTA must call these (uniform trip counts).
struct WfBlockCounters { unsigned int cnt[kWfQueues + 2]; unsigned int base[kWfQueues + 2]; };   // [0..4] shade queues, [5] trace list, [6] rays

// append path slot `slot` to shade queue q (q < 0: nothing)
VT_DEV void wf_enqueue(const WfState& S, WfCounts* __restrict__ cnt, int q, unsigned int slot, WfBlockCounters& sm)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    if (threadIdx.x < kWfQueues) sm.cnt[threadIdx.x] = 0u;
    __syncthreads();
    unsigned int woff = 0, rank = 0;
    #pragma unroll
    for (int k = 0; k < kWfQueues; ++k) {
        const unsigned m = __ballot_sync(full, q == k);
        if (m == 0u) continue;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&sm.cnt[k], (unsigned)__popc(m));
        base = __shfl_sync(full, base, 0);
        if (q == k) { woff = base; rank = (unsigned)__popc(m & lt); }
    }
    __syncthreads();
    if (threadIdx.x < kWfQueues && sm.cnt[threadIdx.x] != 0u) sm.base[threadIdx.x] = atomicAdd(&cnt->sq[threadIdx.x], sm.cnt[threadIdx.x]);
    __syncthreads();
    #pragma unroll
    for (int k = 0; k < kWfQueues; ++k)
        if (q == k) S.sq[k][sm.base[k] + woff + rank] = slot;
}

// reserve one trace-list entry per thread with `traced` and one ray-queue slot per set predicate.
// Returns the trace-list slot; slot_a / slot_b are valid where want_a / want_b.
// once per kernel, before the first wf_reserve_rays
VT_DEV void wf_reserve_init(WfBlockCounters& sm)
{
    if (threadIdx.x < 2) sm.cnt[kWfQueues + threadIdx.x] = 0u;
    __syncthreads();
}
VT_DEV unsigned int wf_reserve_rays(WfCounts* __restrict__ cnt, bool traced, bool want_a, bool want_b, unsigned int& slot_a, unsigned int& slot_b,
                                    WfBlockCounters& sm)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    // sm.cnt[kWfQueues .. +1] are zero on entry: wf_reserve_init at kernel start, then thread 0 re-zeroes them below (two barriers
    // per call instead of three)
    const unsigned mt = __ballot_sync(full, traced), ma = __ballot_sync(full, want_a), mb = __ballot_sync(full, want_b);
    unsigned int wt = 0, wr = 0;
    if (mt != 0u) {
        if (lane == 0) { wt = atomicAdd(&sm.cnt[kWfQueues], (unsigned)__popc(mt)); wr = atomicAdd(&sm.cnt[kWfQueues + 1], (unsigned)(__popc(ma) + __popc(mb))); }
        wt = __shfl_sync(full, wt, 0); wr = __shfl_sync(full, wr, 0);
    }
    __syncthreads();
    if (threadIdx.x == 0 && sm.cnt[kWfQueues] != 0u) {
        const unsigned long long b = atomicAdd(&cnt->tq_rq, (unsigned long long)sm.cnt[kWfQueues] | ((unsigned long long)sm.cnt[kWfQueues + 1] << 32));
        sm.base[kWfQueues] = (unsigned int)b; sm.base[kWfQueues + 1] = (unsigned int)(b >> 32);
        sm.cnt[kWfQueues] = 0u; sm.cnt[kWfQueues + 1] = 0u;                 // ready for the next call (visible after the barrier below)
    }
    __syncthreads();
    const unsigned int rbase = sm.base[kWfQueues + 1] + wr;
    slot_a = rbase + (unsigned)__popc(ma & lt);
    slot_b = rbase + (unsigned)__popc(ma) + (unsigned)__popc(mb & lt);
    return sm.base[kWfQueues] + wt + (unsigned)__popc(mt & lt);
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
// wf_generate: pathTracer.fs:172-196 + dda.h:16-34 of the primary ray. grid = (n_items / 256, rows): a thread generates every rows-th pass of its pixel;
// rows = 1 whenever the frame alone fills the machine (1 / 2 / 4 / 8 / 16 rows measured at 1080p: 12.4 / 12.5 / 12.7 / 13.3 / 14.3 ms).
// Every pixel of the frame gets a slot of generation 0; wf_classify routes the ones that miss the volume's box
// to the finish queue.
// ---------------------------------------------------------------------------------------------------------
template <bool COUNT>
VT_GLOBAL void __launch_bounds__(256)
wf_generate_kernel(const Volume V, const Frame F, const RenderLaunch L, const WfState S, const WfBuf out, int pass0, int n_batch,
                   WfCounts* __restrict__ cnt, int* __restrict__ primary, Counters* __restrict__ counters)
{
    __shared__ WfBlockCounters sm;
    wf_reserve_init(sm);
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
        int flags = WF_HIT_PRIMARY;
        Dda s;
        s.ix = s.iy = s.iz = 0; s.nanmask = 0; s.sx = s.sy = s.sz = 1; s.dx = s.dy = s.dz = s.ex = s.ey = s.ez = 0.f;
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
                if (status != DDA_RUNNING) flags = wf_hit_flags(status, s) | WF_HIT_PRIMARY;
            } else {
                // :187-194 the ray misses the volume's box: finished here. Such pixels come in whole 8x4 blocks (the sky), so the
                // warp does not diverge, and the path never costs a slot, a queue entry or a pass through wf_shade.
                const f3 c = tonemap(background_color<COUNT>(F, rd, tl));
                S.samples[pid] = make_float4(c.x, c.y, c.z, 1.0f);
                valid = false;
            }
            if (primary != nullptr && pass0 + pass_local == L.n_passes - 1) primary[(size_t)px + (size_t)py * (size_t)F.W] = -1;
        }
        unsigned int slot_a, slot_b;
        const bool want = valid && status == DDA_RUNNING;
        const unsigned int slot = wf_reserve_rays(cnt, valid, want, false, slot_a, slot_b, sm);
        if (valid) {
            out.ray0[slot] = make_float4(ro.x, ro.y, ro.z, rd.x);
            // a primary path has radiance 0, throughput 1, nothing pending, bounce 0, no BSDF pdf: wf_shade synthesises all of that, and
            // the only state there is -- the rng offset -- rides in the two unused words of ray1 (no rad0 / rad1 / rad2 traffic at all)
            out.ray1[slot] = make_float4(rd.y, rd.z, i2f(rng.x), i2f(rng.y));
            out.pid[slot] = pid;
            if (!want) out.hit[slot] = make_int4(s.ix, s.iy, s.iz, flags);
        }
        if (want) wf_store_ray(S, slot_a, s, slot, 0, WF_RAY_PRIMARY);
    }
    wf_flush_tally<COUNT>(tl, counters);
}

// ---------------------------------------------------------------------------------------------------------
// wf_trace: the loop of dda.h:38-57 for every ray record of the queue. Persistent warps; a warp reserves kWfGrab
// rays per atomic and its lanes refill from that range whenever fewer than kWfLiveMin of them hold a ray.
// ---------------------------------------------------------------------------------------------------------
#ifndef VT_WF_TRACE_MIN_BLOCKS
#define VT_WF_TRACE_MIN_BLOCKS 6
#endif
template <bool COUNT, bool SKIP>
VT_GLOBAL void __launch_bounds__(256, VT_WF_TRACE_MIN_BLOCKS)
wf_trace_kernel(const Volume V, const WfState S, const WfBuf out, WfCounts* __restrict__ cnt, Counters* __restrict__ counters)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned int n_rays = (unsigned int)(cnt->tq_rq >> 32);
    Tally<COUNT> tl; tl.clear();

#ifdef VT_SKIP_STATS
    unsigned long long dbg_calls = 0, dbg_ok = 0, dbg_steps = 0;
#endif
    bool have = false, exhausted = false;
    unsigned int range_next = 0, range_end = 0;     // warp-uniform
    unsigned int pid = 0;                           // slot of the ray's path in `out`
    int type = 0, status = DDA_NOHIT, aux = 0, chunks = 0;
    const int chunk_guard = (V.X + V.Y + V.Z) / kWfStepChunk + 8;   // belt and braces: see dda_begin on why rays always leave
    Dda s;
    s.ix = s.iy = s.iz = 0; s.nanmask = 0; s.steps = 0; s.bkey = -1; s.brick = 0ull;
    s.dx = s.dy = s.dz = 0.f; s.ex = s.ey = s.ez = 0.f; s.sx = s.sy = s.sz = 1;

    for (;;) {
        // ---- refill idle lanes ---------------------------------------------------------------------------
        const unsigned need = __ballot_sync(full, !have);
        if (!exhausted && __popc(need) > 32 - kWfLiveMin) {
            if (range_next >= range_end) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&cnt->work, (unsigned)kWfGrab);
                base = __shfl_sync(full, base, 0);
                range_next = base;
                range_end = min(base + (unsigned)kWfGrab, n_rays);
                if (base >= n_rays) { exhausted = true; range_end = range_next = 0; }
            }
            if (!have) {
                const unsigned r = range_next + (unsigned)__popc(need & lt);
                if (r < range_end) {
                    const int4 a = S.rq0[r]; const float4 b = S.rq1[r], c = S.rq2[r];
                    s.ix = a.x; s.iy = a.y; s.iz = a.z; pid = (unsigned)a.w;
                    s.dx = b.x; s.dy = b.y; s.dz = b.z; aux = f2bits(b.w);
                    s.ex = c.x; s.ey = c.y; s.ez = c.z;
                    const int bits = f2bits(c.w);
                    s.sx = (bits & 1) ? -1 : 1; s.sy = (bits & 2) ? -1 : 1; s.sz = (bits & 4) ? -1 : 1;
                    type = bits >> 4;
                    s.steps = 0; s.bkey = -1;
                    status = DDA_RUNNING;
                    chunks = 0;
                    have = true;
                }
            }
            range_next = min(range_next + (unsigned)__popc(need), range_end);
        }
        if (__ballot_sync(full, have) == 0u) { if (exhausted) break; else continue; }
        // ---- the hot loop ----------------------------------------------------------------------------------
        if (have && status == DDA_RUNNING) {       // lanes leave the chunk through `break`: one reconvergence point per chunk, not per step
            #pragma unroll
            for (int k = 0; k < kWfStepChunk; ++k) {
                status = dda_step<COUNT>(V, s, tl);
                if (status != DDA_RUNNING) break;
            }
        }
        if (++chunks > chunk_guard && status == DDA_RUNNING) status = DDA_NOHIT;
        if (SKIP && !COUNT) {                     // counting builds step every voxel so that S stays the algorithmic count
            // the cheap part (one byte per lane) runs converged; the skip itself only when enough lanes want it, so that its
            // divergent set-up is not paid for one or two lanes while the rest of the warp idles
            const bool running = have && status == DDA_RUNNING;
            const int radius = running ? dda_skip_radius(V, s) : 0;
            const unsigned m_run = __ballot_sync(full, running), m_want = __ballot_sync(full, radius >= 2);
            if (m_want != 0u && (__popc(m_want) >= kWfSkipMinLanes || 2 * __popc(m_want) >= __popc(m_run))) {
                if (radius >= 2) {
                    const int skipped = dda_skip(V, s, radius);
#ifdef VT_SKIP_STATS
                    dbg_calls += 1; dbg_ok += skipped > 0; dbg_steps += skipped;
#else
                    (void)skipped;
#endif
                }
            }
        }
        // ---- retire finished rays --------------------------------------------------------------------------
        if (have && status != DDA_RUNNING) {
            const int kind = type & 3;                        // type >> 2: bounce count of the path, handed through to wf_classify
            if (kind == WF_RAY_SHADOW) out.vis[pid] = wf_light_visible(V, aux, status, s);
            else out.hit[pid] = make_int4(s.ix, s.iy, s.iz, wf_hit_flags(status, s) | (kind == WF_RAY_PRIMARY ? WF_HIT_PRIMARY : 0) | ((type >> 2) << 16));
            have = false;
        }
    }
#ifdef VT_SKIP_STATS
    if (SKIP) { atomicAdd(&counters->E, dbg_calls); atomicAdd(&counters->Q, dbg_ok); atomicAdd(&counters->H, dbg_steps); }
#endif
    wf_flush_tally<COUNT>(tl, counters);
}

// ---------------------------------------------------------------------------------------------------------
// wf_classify: routes the paths of the generation just traced (pathTracer.fs:202-208, :214, :282-291).
// Sequential over the generation's slots: coalesced reads, one shade-queue entry (the slot) per path.
// ---------------------------------------------------------------------------------------------------------
VT_GLOBAL void __launch_bounds__(256)
wf_classify_kernel(const Volume V, const Frame F, const RenderLaunch L, const WfState S, const WfBuf out, int pass0,
                   const WfCounts* __restrict__ cin, WfCounts* __restrict__ cnext, int* __restrict__ primary)
{
    __shared__ WfBlockCounters sm;
    const unsigned int n = (unsigned int)cin->tq_rq;
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int n_round = (n + 255u) & ~255u;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        int q = -1;
        if (i < n) {
            const int4 h = out.hit[i];
            const bool surface = (h.w & 3) != 0;
            const bool is_primary = (h.w & WF_HIT_PRIMARY) != 0;
            const int bounces = is_primary ? -1 : (h.w >> 16);            // written next to the hit by wf_trace / wf_shade
            q = (surface && bounces + 1 < F.max_bounces) ? wf_material_queue(V, F, h.x, h.y, h.z) : 0;
            if (is_primary && surface && primary != nullptr) {
                const unsigned int pid = out.pid[i];
                const int pass_local = (int)(pid / (unsigned)S.n_items), item = (int)(pid - (unsigned)pass_local * (unsigned)S.n_items);
                int px, py;
                if (pass0 + pass_local == L.n_passes - 1 && wf_item_pixel(F, L, item, px, py))
                    primary[(size_t)px + (size_t)py * (size_t)F.W] = hit_code(V, wf_hit_pos(h), (h.w & WF_HIT_GROUND) != 0);
            }
        }
        wf_enqueue(S, cnext, q, i, sm);
    }
}

// ---------------------------------------------------------------------------------------------------------
// wf_shade: one loop iteration of pathTracer.fs:214-292 up to (not including) the two traversals, preceded by
// the tail of the previous iteration (shadow-ray result :134-164, environment on a miss :282-289, bounces++).
// Queue 0 holds the paths that end here; queues 1..4 the surface hits sorted by material type.
// ---------------------------------------------------------------------------------------------------------
template <bool COUNT>
VT_GLOBAL void __launch_bounds__(VT_WF_SHADE_THREADS, VT_WF_SHADE_MIN_BLOCKS)
wf_shade_kernel(const Volume V, const Frame F, const WfState S, const WfBuf in, const WfBuf out, WfCounts* __restrict__ cnt,
                Counters* __restrict__ counters)
{
    __shared__ WfBlockCounters sm;
    wf_reserve_init(sm);
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
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
        const unsigned int* __restrict__ qp = S.sq[0];
        #pragma unroll
        for (int j = 1; j < kWfQueues; ++j) if (chunk >= chunks_before[j]) { cb = chunks_before[j]; nq = n_q[j]; qp = S.sq[j]; }
        const unsigned int idx = ((chunk - cb) << 5) + (unsigned)lane;
        const bool valid = chunk < chunks_before[kWfQueues] && idx < nq;
        bool continues = false;
        unsigned int pid = 0;
        int st_a = DDA_NOHIT, st_b = DDA_NOHIT, target = -1;
        float4 o_ray0, o_ray1, o_rad0, o_rad1, o_rad2;
        o_ray0 = o_ray1 = o_rad0 = o_rad1 = o_rad2 = make_float4(0.f, 0.f, 0.f, 0.f);
        Dda sa, sb;
        sa.ix = sa.iy = sa.iz = 0; sa.nanmask = 0; sa.sx = sa.sy = sa.sz = 1; sa.dx = sa.dy = sa.dz = sa.ex = sa.ey = sa.ez = 0.f;
        sb = sa;
        if (valid) {
            const unsigned int slot = qp[idx];
            pid = in.pid[slot];
            const int4 h = in.hit[slot];
            const float4 r0 = in.ray0[slot], r1 = in.ray1[slot];
            const float4 a0 = (h.w & WF_HIT_PRIMARY) ? make_float4(0.f, 0.f, 0.f, 1.0f) : in.rad0[slot];     // see wf_generate
            {   // gathers whose addresses are known now but whose values are needed hundreds of instructions later
                float2 rs = make_float2(r1.z, r1.w);                                // rng offset: in ray1 for primary paths (wf_generate)
                if (!(h.w & WF_HIT_PRIMARY)) { const float4 a2p = in.rad2[slot]; rs = make_float2(a2p.z, a2p.w); }
                const int nx = f2bits(rs.x), ny = f2bits(rs.y);
                if ((unsigned)nx < (unsigned)F.noise_w && (unsigned)ny < (unsigned)F.noise_h) prefetch_l1(F.noise + ((size_t)nx + (size_t)ny * (size_t)F.noise_w));
                if ((unsigned)h.x < (unsigned)V.X && (unsigned)h.y < (unsigned)V.Y && (unsigned)h.z < (unsigned)V.Z)
                    prefetch_l1(V.mat + ((size_t)h.x + (size_t)h.y * (size_t)V.X + (size_t)h.z * (size_t)V.X * (size_t)V.Y));
            }
            const f3 ro = mk3(r0.x, r0.y, r0.z), rd = mk3(r0.w, r1.x, r1.y);
            f3 radiance = mk3(a0.x, a0.y, a0.z);
            int bounces = (h.w & WF_HIT_PRIMARY) ? 0 : f2bits(r1.w);
            const bool surface = (h.w & 3) != 0;
            bool finished = false;
            float4 a1 = make_float4(1.f, 1.f, 0.f, 0.f), a2 = make_float4(0.f, 0.f, 0.f, 0.f);
            // the environment seen by a ray that left the scene, primary (:187-208) or bounce (:282-289): ONE expansion of the
            // lat-long mapping + bilinear lookup for both uses (this kernel is bound by instruction fetch, see vt_math.cuh)
            f3 bg = mk3(0.0f);
            if (!surface) bg = background_color<COUNT>(F, rd, tl);
            if (h.w & WF_HIT_PRIMARY) {
                if (!surface) {                                                    // :187-194, :202-208
                    radiance = bg;
                    finished = true;
                } else if (!(0 < F.max_bounces)) finished = true;                  // :214 never entered
                a2 = make_float4(0.f, i2f(0), r1.z, r1.w);                         // a1 = throughput (1, 1), no pending light: the initialiser above
            } else {
                a1 = in.rad1[slot]; a2 = in.rad2[slot];
                // pathTracer.fs:248 with the shadow-ray result of the previous iteration
                if (in.vis[slot] != 0) {
                    radiance = radiance + mk3(a1.z, a1.w, a2.x);
                    VT_TALLY(H, 1);                                                // the BSDF evaluation of :161
                } else {
                    const float qn = __int_as_float(0x7fc00000);                   // radiance + throughput * vec3(0)
                    const int pn = f2bits(a2.y);
                    if (pn & 1) radiance.x = qn;
                    if (pn & 2) radiance.y = qn;
                    if (pn & 4) radiance.z = qn;
                }
                if (!surface) {                                                    // :282-289 the bounce ray left the scene
                    const f3 throughput = mk3(a0.w, a1.x, a1.y);
                    const f4 Lp = env_with_pdf(F, bg);                                // lights.h:20-33
                    const float mis = power_heuristic(r1.z, Lp.w);
                    radiance = radiance + (throughput * xyz(Lp)) * mis;
                    finished = true;
                } else {
                    bounces++;                                                     // :291
                    if (!(bounces < F.max_bounces)) finished = true;               // :214
                }
            }
            if (!finished) {
                f3 throughput = mk3(a0.w, a1.x, a1.y);
                int2 rng = make_int2(f2bits(a2.z), f2bits(a2.w));
                const f3 hit = wf_hit_pos(h);
                Basis hb;
                voxel_to_world(V, hit, ro, rd, hb);                                // :221-223
                const int mat_off = fetch_offset(V, h.x, h.y, h.z);                // :225-226
                if (h.x == sel_x && h.y == sel_y && h.z == sel_z) {                // :228-233
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
                    o_ray0 = make_float4(hb.position.x, hb.position.y, hb.position.z, wi.x);   // :278-279
                    o_ray1 = make_float4(wi.y, wi.z, bf.w, i2f(bounces));
                    o_rad0 = make_float4(radiance.x, radiance.y, radiance.z, throughput.x);
                    o_rad1 = make_float4(throughput.y, throughput.z, pending.x, pending.y);
                    o_rad2 = make_float4(pending.z, i2f(pending_nan), i2f(rng.x), i2f(rng.y));
                    // dda.h:16-34 of the shadow ray (:133) and of the bounce ray (:282)
                    target = ls.target;
                    st_a = dda_begin<COUNT>(V, hb.position, xyz(ls.wl), sa, tl);
                    st_b = dda_begin<COUNT>(V, hb.position, wi, sb, tl);
                    continues = true;
                }
            }
            if (finished) {
                const f3 c = tonemap(radiance);                                    // :294-295
                S.samples[pid] = make_float4(c.x, c.y, c.z, 1.0f);
            }
        }
        // surviving paths: one trace-list entry, up to two ray records
        unsigned int slot_a, slot_b;
        const bool want_a = continues && st_a == DDA_RUNNING, want_b = continues && st_b == DDA_RUNNING;
        const unsigned int tslot = wf_reserve_rays(cnt, continues, want_a, want_b, slot_a, slot_b, sm);
        if (continues) {                     // the surviving path moves to slot `tslot` of the next generation
            out.ray0[tslot] = o_ray0; out.ray1[tslot] = o_ray1; out.rad0[tslot] = o_rad0; out.rad1[tslot] = o_rad1; out.rad2[tslot] = o_rad2;
            out.pid[tslot] = pid;
            if (st_a != DDA_RUNNING) out.vis[tslot] = wf_light_visible(V, target, st_a, sa);
            if (st_b != DDA_RUNNING) out.hit[tslot] = make_int4(sb.ix, sb.iy, sb.iz, wf_hit_flags(st_b, sb) | (f2bits(o_ray1.w) << 16));
        }
        if (want_a) wf_store_ray(S, slot_a, sa, tslot, target, WF_RAY_SHADOW);
        if (want_b) wf_store_ray(S, slot_b, sb, tslot, 0, WF_RAY_BOUNCE | (f2bits(o_ray1.w) << 2));
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
        const float4 s = S.samples[(size_t)p * (size_t)S.n_items + (size_t)item];
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
void wf_shade_launch(bool count, unsigned int blocks, cudaStream_t st, const Volume& V, const Frame& F, const WfState& S, const WfBuf& in,
                     const WfBuf& out, WfCounts* cnt, Counters* counters);

} // namespace vt
