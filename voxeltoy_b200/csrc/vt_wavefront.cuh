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
//   wf_generate   pathTracer.fs:172-208   RNG offset, camera ray, slab test, primary DDA (coherent, traced inline);
//                                         misses write their tone-mapped background sample, hits are queued
//   repeat max_bounces times:
//     wf_shade    :216-279 (+ :248/:282-289 of the previous iteration) resolve the previous shadow ray, add the
//                                         environment on a miss, build the hit frame, sample the light and the
//                                         BSDF; emits one shadow ray and one bounce ray per surviving path
//     wf_trace    dda.h:63-100            persistent warps, every lane refills itself from a global ray counter as
//                                         soon as its ray ends, so the DDA loop stays >= kTraceMin/32 lanes wide;
//                                         the epilogue sorts surviving paths into one queue per material type
//                                         (Lambert / metal / plastic / other) so wf_shade does not diverge on the
//                                         material switch, and finished paths into a "finish" queue
//   wf_shade      (last)                  only the finish queue is populated: resolve, tone-map, write the sample
//   wf_accumulate accumulation.fs:10-18   folds the P samples of every pixel into the running average in pass order
//
// Every path performs exactly the arithmetic of trace_pixel() (vt_device.cuh) in the same order, and the
// per-pixel order of the running average is unchanged: results are bit-identical to variants 0 and 1.
#pragma once
#include "vt_kernels.cuh"

namespace vt {

constexpr int kWfQueues = 5;          // 0 = finish, 1..3 = material type 0..2, 4 = other material types
constexpr int kWfTraceMin = 24;       // leave the stepping loop when fewer lanes than this are tracing
constexpr int kWfStepChunk = 8;       // DDA iterations between two votes

// counts block of one iteration (device memory, zeroed once per batch)
struct WfCounts {
    unsigned int sq[kWfQueues];       // entries in the shade queues
    unsigned int tq;                  // entries in the trace queue (paths; 2 rays each)
    unsigned int work;                // ray hand-out counter of wf_trace
    unsigned int pad;
};

struct WfState {
    float4* __restrict__ ray0;        // ray origin xyz, dir x      (the ray whose hit is shaded next)
    float4* __restrict__ ray1;        // dir y, dir z, bsdf pdf, bounces (int bits)
    float4* __restrict__ shadow;      // shadow dir xyz, light target (int bits, -1 = environment)
    float4* __restrict__ rad0;        // radiance xyz, throughput x
    float4* __restrict__ rad1;        // throughput y z, pending x y
    float4* __restrict__ rad2;        // pending z, pending_nan (int bits), rng offset x y (int bits)
    int4* __restrict__ hit;           // hit voxel ix iy iz, code: bit0 hit voxel, bit1 ground, bits 8..10 nanmask
    int* __restrict__ vis;            // shadow ray result: 1 = light visible
    float4* __restrict__ samples;     // tone-mapped sample per path (P * n_items)
    unsigned int* __restrict__ sq[kWfQueues];
    unsigned int* __restrict__ tq;
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

// warp-aggregated append of `pid` to queue q (q < 0: nothing). All 32 lanes must call.
VT_DEV void wf_enqueue(const WfState& S, WfCounts* __restrict__ cnt, int q, unsigned int pid)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    #pragma unroll
    for (int k = 0; k < kWfQueues; ++k) {
        const unsigned m = __ballot_sync(full, q == k);
        if (m == 0u) continue;
        const int leader = __ffs(m) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(&cnt->sq[k], (unsigned)__popc(m));
        base = __shfl_sync(full, base, leader);
        if (q == k) S.sq[k][base + (unsigned)__popc(m & ((1u << lane) - 1u))] = pid;
    }
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
// wf_generate: pathTracer.fs:172-208. grid = (n_items / 128, n_passes)
// ---------------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(128)
wf_generate_kernel(const Volume V, const Frame F, const RenderLaunch L, const WfState S, int pass0,
                   WfCounts* __restrict__ cnt, int* __restrict__ primary, Counters* __restrict__ counters)
{
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    const int pass_local = blockIdx.y;
    const unsigned int pid = (unsigned)pass_local * (unsigned)S.n_items + (unsigned)item;
    Tally<COUNT> tl; tl.clear();
    int px, py, q = -1;
    if (item < S.n_items && wf_item_pixel(F, L, item, px, py)) {
        const int sample = L.first_sample + (pass0 + pass_local) * L.sample_stride;
        const f3 frag = mk3((float)px + 0.5f, (float)py + 0.5f, 0.55f);
        int2 rng = rng_offset(px, py, sample, F.noise_w, F.noise_h);               // :174
        f3 ro, rd, hit;
        generate_ray<COUNT>(F, frag, rng, ro, rd, tl);                             // :179
        const float t = ray_aabb(ro, rd, V.bmin, V.bmax);                          // :183
        int prim = -1;
        bool hit_ground = false;
        bool surface = false;
        if (!(t < 0.0f)) surface = traverse<COUNT>(V, ro + t * rd, rd, hit, hit_ground, tl);   // :196-202
        if (!surface) {
            const f3 c = tonemap(background_color<COUNT>(F, rd, tl));              // :187-194, :202-208
            S.samples[pid] = make_float4(c.x, c.y, c.z, 1.0f);
        } else {
            prim = hit_code(V, hit, hit_ground);
            if (!(0 < F.max_bounces)) {                                            // :214 never entered
                const f3 c = tonemap(mk3(0.0f));
                S.samples[pid] = make_float4(c.x, c.y, c.z, 1.0f);
            } else {
                const int nm = (hit.x != hit.x ? 1 : 0) | (hit.y != hit.y ? 2 : 0) | (hit.z != hit.z ? 4 : 0);
                const int ix = f2i(hit.x), iy = f2i(hit.y), iz = f2i(hit.z);
                S.ray0[pid] = make_float4(ro.x, ro.y, ro.z, rd.x);
                S.ray1[pid] = make_float4(rd.y, rd.z, 0.0f, i2f(0));
                S.rad0[pid] = make_float4(0.f, 0.f, 0.f, 1.0f);
                S.rad1[pid] = make_float4(1.0f, 1.0f, 0.f, 0.f);
                S.rad2[pid] = make_float4(0.f, i2f(0), i2f(rng.x), i2f(rng.y));
                S.hit[pid] = make_int4(ix, iy, iz, (hit_ground ? 2 : 1) | (nm << 8));
                q = wf_material_queue(V, F, ix, iy, iz);
            }
        }
        if (primary != nullptr && pass0 + pass_local == L.n_passes - 1) primary[(size_t)px + (size_t)py * (size_t)F.W] = prim;
    }
    wf_enqueue(S, cnt, q, pid);
    wf_flush_tally<COUNT>(tl, counters);
}

// ---------------------------------------------------------------------------------------------------------
// wf_shade: one loop iteration of pathTracer.fs:214-292 up to (not including) the two traversals, preceded by
// the tail of the previous iteration (shadow-ray result :134-164, environment on a miss :282-289, bounces++).
// first = 1: the paths come from wf_generate (no pending light sample, primary hit).
// ---------------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(128)
wf_shade_kernel(const Volume V, const Frame F, const WfState S, int first,
                const WfCounts* __restrict__ cin, WfCounts* __restrict__ cout, Counters* __restrict__ counters)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    Tally<COUNT> tl; tl.clear();
    // chunks of 32 entries, queue after queue
    unsigned int n_q[kWfQueues], chunks_before[kWfQueues + 1];
    chunks_before[0] = 0;
    #pragma unroll
    for (int k = 0; k < kWfQueues; ++k) { n_q[k] = cin->sq[k]; chunks_before[k + 1] = chunks_before[k] + ((n_q[k] + 31u) >> 5); }
    const int sel_x = F.shared->sel_index[0], sel_y = F.shared->sel_index[1], sel_z = F.shared->sel_index[2];

    for (unsigned int chunk = (unsigned)warp_global; chunk < chunks_before[kWfQueues]; chunk += (unsigned)n_warps) {
        unsigned int cb = 0, nq = n_q[0];
        const unsigned int* __restrict__ qp = S.sq[0];
        #pragma unroll
        for (int j = 1; j < kWfQueues; ++j) if (chunk >= chunks_before[j]) { cb = chunks_before[j]; nq = n_q[j]; qp = S.sq[j]; }
        const unsigned int idx = ((chunk - cb) << 5) + (unsigned)lane;
        const bool valid = idx < nq;
        bool continues = false;
        unsigned int pid = 0;
        if (valid) {
            pid = qp[idx];
            const float4 r0 = S.ray0[pid], r1 = S.ray1[pid], a0 = S.rad0[pid], a1 = S.rad1[pid], a2 = S.rad2[pid];
            const int4 h = S.hit[pid];
            f3 ro = mk3(r0.x, r0.y, r0.z), rd = mk3(r0.w, r1.x, r1.y);
            f3 radiance = mk3(a0.x, a0.y, a0.z), throughput = mk3(a0.w, a1.x, a1.y);
            int bounces = f2bits(r1.w);
            int2 rng = make_int2(f2bits(a2.z), f2bits(a2.w));
            bool finished = false;
            if (!first) {
                // pathTracer.fs:248 with the shadow-ray result of the previous iteration
                if (S.vis[pid] != 0) {
                    radiance = radiance + mk3(a1.z, a1.w, a2.x);
                    VT_TALLY(H, 1);                                                // the BSDF evaluation of :161
                } else {
                    const float qn = __int_as_float(0x7fc00000);                   // radiance + throughput * vec3(0)
                    const int pn = f2bits(a2.y);
                    if (pn & 1) radiance.x = qn;
                    if (pn & 2) radiance.y = qn;
                    if (pn & 4) radiance.z = qn;
                }
                if ((h.w & 3) == 0) {                                              // :282-289 the bounce ray left the scene
                    const f4 Lp = evaluate_env<COUNT>(F, rd, tl);
                    const float mis = power_heuristic(r1.z, Lp.w);
                    radiance = radiance + (throughput * xyz(Lp)) * mis;
                    finished = true;
                } else {
                    bounces++;                                                     // :291
                    if (!(bounces < F.max_bounces)) finished = true;               // :214
                }
            }
            if (!finished) {
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
                    S.ray0[pid] = make_float4(hb.position.x, hb.position.y, hb.position.z, wi.x);   // :278-279
                    S.ray1[pid] = make_float4(wi.y, wi.z, bf.w, i2f(bounces));
                    S.shadow[pid] = make_float4(ls.wl.x, ls.wl.y, ls.wl.z, i2f(ls.target));
                    S.rad0[pid] = make_float4(radiance.x, radiance.y, radiance.z, throughput.x);
                    S.rad1[pid] = make_float4(throughput.y, throughput.z, pending.x, pending.y);
                    S.rad2[pid] = make_float4(pending.z, i2f(pending_nan), i2f(rng.x), i2f(rng.y));
                    continues = true;
                }
            }
            if (finished) {
                const f3 c = tonemap(radiance);                                    // :294-295
                S.samples[pid] = make_float4(c.x, c.y, c.z, 1.0f);
            }
        }
        // surviving paths go to the trace queue (two rays each)
        const unsigned m = __ballot_sync(full, continues);
        if (m != 0u) {
            const int leader = __ffs(m) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(&cout->tq, (unsigned)__popc(m));
            base = __shfl_sync(full, base, leader);
            if (continues) S.tq[base + (unsigned)__popc(m & ((1u << lane) - 1u))] = pid;
        }
    }
    wf_flush_tally<COUNT>(tl, counters);
}

// ---------------------------------------------------------------------------------------------------------
// wf_trace: the shadow ray (:133) and the bounce ray (:282) of every path in the trace queue.
// ray r = 2 * slot + type (0 shadow, 1 bounce). Persistent warps; lanes refill from cnt->work.
// ---------------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(256)
wf_trace_kernel(const Volume V, const Frame F, const WfState S, WfCounts* __restrict__ cnt, WfCounts* __restrict__ cnext,
                Counters* __restrict__ counters)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned int n_rays = 2u * cnt->tq;
    Tally<COUNT> tl; tl.clear();

    bool have = false, exhausted = false;
    unsigned int pid = 0;
    int type = 0, status = DDA_NOHIT, aux = 0;     // aux: light target (shadow) or bounce count (bounce)
    Dda s;
    s.ix = s.iy = s.iz = 0; s.nanmask = 0; s.steps = 0; s.bkey = -1; s.brick = 0ull;
    s.dx = s.dy = s.dz = 0.f; s.ex = s.ey = s.ez = 0.f; s.sx = s.sy = s.sz = 0;

    for (;;) {
        // ---- refill idle lanes -------------------------------------------------------------------
        if (!exhausted) {
            const unsigned need = __ballot_sync(full, !have);
            if (need != 0u) {
                const int leader = __ffs(need) - 1;
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(&cnt->work, (unsigned)__popc(need));
                base = __shfl_sync(full, base, leader);
                if (base + (unsigned)__popc(need) >= n_rays) exhausted = true;      // warp-uniform
                if (!have) {
                    const unsigned r = base + (unsigned)__popc(need & lt);
                    if (r < n_rays) {
                        pid = S.tq[r >> 1];
                        type = (int)(r & 1u);
                        const float4 r0 = S.ray0[pid];
                        f3 d;
                        if (type == 0) { const float4 sh = S.shadow[pid]; d = mk3(sh.x, sh.y, sh.z); aux = f2bits(sh.w); }
                        else { const float4 r1 = S.ray1[pid]; d = mk3(r0.w, r1.x, r1.y); aux = f2bits(r1.w); }
                        status = dda_begin<COUNT>(V, mk3(r0.x, r0.y, r0.z), d, s, tl);
                        have = true;
                    }
                }
            }
        }
        // ---- the hot loop: step every tracing lane ---------------------------------------------------
        {
            unsigned live = __ballot_sync(full, have && status == DDA_RUNNING);
            const int thresh = exhausted ? 1 : kWfTraceMin;
            while (__popc(live) >= thresh) {
                #pragma unroll 1
                for (int k = 0; k < kWfStepChunk; ++k)
                    if (have && status == DDA_RUNNING) status = dda_step<COUNT>(V, s, tl);
                live = __ballot_sync(full, have && status == DDA_RUNNING);
            }
        }
        // ---- retire finished rays -------------------------------------------------------------------------
        int q = -1;
        if (have && status != DDA_RUNNING) {
            const bool ground = (status != DDA_HIT) && !(s.nanmask & 2) && (s.iy < 0);     // dda.h:75-78
            const bool surface = (status == DDA_HIT) || ground;
            if (type == 0) {
                S.vis[pid] = light_occluded(V, aux, surface, dda_position(s)) ? 0 : 1;      // pathTracer.fs:134-153
            } else {
                S.hit[pid] = make_int4(s.ix, s.iy, s.iz, (status == DDA_HIT ? 1 : (ground ? 2 : 0)) | (s.nanmask << 8));
                if (!surface || !(aux + 1 < F.max_bounces)) q = 0;                           // :282-289 / :214 -> finish
                else q = wf_material_queue(V, F, s.ix, s.iy, s.iz);
            }
            have = false;
        }
        wf_enqueue(S, cnext, q, pid);
        if (exhausted && __ballot_sync(full, have) == 0u) break;
    }
    wf_flush_tally<COUNT>(tl, counters);
}

// ---------------------------------------------------------------------------------------------------------
// wf_accumulate: accumulation.fs:10-18 over the batch's passes, in pass order. One thread per item.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
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

} // namespace vt
