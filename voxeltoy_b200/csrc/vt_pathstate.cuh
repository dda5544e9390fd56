// vt_pathstate.cuh -- persistent path-state-machine render kernel (K1 + K2 fused).
//
// Why: the one-thread-per-pixel megakernel (vt_render_kernel, vt_kernels.cuh) runs with 9.6 of 32 lanes
// active (ncu, profiles/r01_v1_*): DDA trip counts differ per ray, paths die at different bounces, and every
// lane waits for the slowest one of its warp. Memory is idle (L2 2.7 %, DRAM 0.03 %), so the fix is lane
// utilisation, not traffic.
//
// How: each lane owns one path at a time and is in one of four phases
//     NEW      needs a new sample (next pass of its pixel, or a new pixel from the global work counter)
//     TRACE    a DDA segment is in flight (primary, shadow or bounce ray); its state lives in registers
//     VERTEX   a primary/bounce segment ended on a surface: build the hit frame, sample the light (-> shadow
//              segment) and the BSDF (-> the bounce segment that follows the shadow segment)
//     RESOLVE  a segment ended and needs its cheap follow-up (shadow: add the direct light if visible and
//              start the bounce segment; primary/bounce miss: add the environment, finish the sample)
// The warp alternates between (a) stepping all TRACE lanes together -- the hot loop, identical arithmetic to
// dda.h for every lane -- while at least kTraceMin lanes are tracing, and (b) running the heavier phases for
// all lanes that queued up for them. A lane that finishes a sample immediately starts the next one, so no lane
// idles until its warp's longest path ends. Per-sample arithmetic and the per-pixel order of the running
// average are exactly those of the megakernel (and of pathTracer.fs / accumulation.fs): results are bit-identical.
#pragma once
#include "vt_kernels.cuh"

namespace vt {

enum { PH_NEW = 0, PH_TRACE = 1, PH_VERTEX = 2, PH_RESOLVE = 3, PH_EXIT = 4 };
enum { SEG_PRIMARY = 0, SEG_SHADOW = 1, SEG_BOUNCE = 2 };

constexpr int kTraceMin = 20;      // leave the stepping loop when fewer lanes than this are tracing
constexpr int kStepChunk = 8;      // DDA iterations between two votes

struct PathLane {
    // pixel / sample bookkeeping
    int px, py, pass, prim;
    bool has_pixel;
    float4 avg;
    int2 rng;
    // path
    f3 radiance, throughput, ro, rd;   // ro/rd: the ray whose hit is being shaded (primary: camera origin)
    int bounces;
    // segment in flight
    int seg, status;
    Dda dda;
    // produced by VERTEX, consumed when the shadow segment ends
    f3 pending;        // throughput * direct light, added if the light is visible
    int pending_nan;   // components of the (pre-update) throughput that are not finite: throughput * 0 is NaN there
    int light_target;  // emissive voxel index or -1
    f3 next_o, next_d; // the bounce ray
    float bsdf_pdf;
};

// segment finished: surfaces of primary / bounce rays go to VERTEX, everything else to RESOLVE
VT_DEV int phase_after_trace(const PathLane& p)
{
    if (p.seg == SEG_SHADOW) return PH_RESOLVE;
    const bool ground = (p.status != DDA_HIT) && !(p.dda.nanmask & 2) && (p.dda.iy < 0);      // dda.h:75-78
    return (p.status == DDA_HIT || ground) ? PH_VERTEX : PH_RESOLVE;
}

template <bool COUNT>
VT_DEV void begin_segment(const Volume& V, PathLane& p, int seg, f3 o, f3 d, int& phase, Tally<COUNT>& tl)
{
    p.seg = seg;
    p.status = dda_begin<COUNT>(V, o, d, p.dda, tl);
    phase = (p.status == DDA_RUNNING) ? PH_TRACE : phase_after_trace(p);
}

template <bool COUNT>
VT_GLOBAL void __launch_bounds__(128)
vt_render_ps_kernel(const Volume V, const Frame F, const RenderLaunch L, int n_items,
                    float4* __restrict__ accum, int* __restrict__ primary, Counters* __restrict__ counters,
                    unsigned int* __restrict__ work_counter)
{
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    Tally<COUNT> tl; tl.clear();
    PathLane p;
    p.pass = L.n_passes;           // "pixel finished": the first NEW fetches a pixel
    p.px = p.py = 0; p.prim = -1; p.has_pixel = false;
    int phase = PH_NEW;
    const int sel_x = F.shared->sel_index[0], sel_y = F.shared->sel_index[1], sel_z = F.shared->sel_index[2];

    for (;;) {
        const unsigned m_trace = __ballot_sync(full, phase == PH_TRACE);
        const unsigned m_other = __ballot_sync(full, phase != PH_TRACE && phase != PH_EXIT);
        if (m_trace == 0u && m_other == 0u) break;

        // ---------------- (a) the hot loop: step every tracing lane ---------------------------------------
        if (m_trace != 0u && (__popc(m_trace) >= kTraceMin || m_other == 0u)) {
            unsigned live = m_trace;
            do {
                #pragma unroll 1
                for (int k = 0; k < kStepChunk; ++k) {
                    if (phase == PH_TRACE) {
                        p.status = dda_step<COUNT>(V, p.dda, tl);
                        if (p.status != DDA_RUNNING) phase = phase_after_trace(p);
                    }
                }
                live = __ballot_sync(full, phase == PH_TRACE);
            } while (__popc(live) >= kTraceMin || (live != 0u && m_other == 0u && __popc(live) * 2 >= __popc(m_trace)));
            continue;
        }

        // ---------------- (b) RESOLVE: cheap follow-ups ------------------------------------------------------
        if (phase == PH_RESOLVE) {
            if (p.seg == SEG_SHADOW) {
                const bool ground = (p.status != DDA_HIT) && !(p.dda.nanmask & 2) && (p.dda.iy < 0);
                const bool hit_something = (p.status == DDA_HIT) || ground;
                if (!light_occluded(V, p.light_target, hit_something, dda_position(p.dda))) {
                    p.radiance = p.radiance + p.pending;                                   // pathTracer.fs:248
                    VT_TALLY(H, 1);
                } else {
                    const float qn = __int_as_float(0x7fc00000);                           // radiance + throughput * vec3(0)
                    if (p.pending_nan & 1) p.radiance.x = qn;
                    if (p.pending_nan & 2) p.radiance.y = qn;
                    if (p.pending_nan & 4) p.radiance.z = qn;
                }
                p.ro = p.next_o; p.rd = p.next_d;                                          // :278-279
                begin_segment<COUNT>(V, p, SEG_BOUNCE, p.ro, p.rd, phase, tl);             // :282
            } else {
                // primary or bounce ray that left the scene
                f4 out;
                if (p.seg == SEG_PRIMARY) {
                    out = mk4(tonemap(background_color<COUNT>(F, p.rd, tl)), 1.0f);         // :202-208
                } else {
                    const f4 Lp = evaluate_env<COUNT>(F, p.rd, tl);                         // :285-288
                    const float mis = power_heuristic(p.bsdf_pdf, Lp.w);
                    p.radiance = p.radiance + (p.throughput * xyz(Lp)) * mis;
                    out = mk4(tonemap(p.radiance), 1.0f);                                   // :294-295
                }
                // fold into the running average (accumulation.fs:17) or the per-rank sum
                if (L.sum_mode) {
                    p.avg.x += out.x; p.avg.y += out.y; p.avg.z += out.z; p.avg.w += out.w;
                } else {
                    const float n = (float)(L.n_prev + p.pass), n1 = (float)(L.n_prev + p.pass + 1);
                    p.avg.x = (out.x + p.avg.x * n) / n1; p.avg.y = (out.y + p.avg.y * n) / n1;
                    p.avg.z = (out.z + p.avg.z * n) / n1; p.avg.w = (out.w + p.avg.w * n) / n1;
                }
                p.pass++;
                phase = PH_NEW;
            }
        }

        // ---------------- (c) VERTEX: hit frame, light sample, BSDF sample -------------------------------------
        if (phase == PH_VERTEX) {
            const f3 hit = dda_position(p.dda);
            if (p.seg == SEG_PRIMARY) p.prim = hit_code(V, hit, p.status != DDA_HIT);
            bool finished = false;
            if (p.seg == SEG_BOUNCE) {                                                     // :291, loop test :214
                p.bounces++;
                if (!(p.bounces < F.max_bounces)) finished = true;
            } else if (!(0 < F.max_bounces)) finished = true;
            if (!finished) {
                Basis hb;
                voxel_to_world(V, hit, p.ro, p.rd, hb);                                    // :221-223
                const int ix = f2i(hit.x), iy = f2i(hit.y), iz = f2i(hit.z);               // :225
                const int mat_off = fetch_offset(V, ix, iy, iz);                           // :226
                if (ix == sel_x && iy == sel_y && iz == sel_z) {                           // :228-233
                    p.radiance = p.radiance + mk3(1.0f, 0.0f, 0.0f);
                    finished = true;
                } else {
                    const f3 wo = -p.rd;                                                   // :237
                    const f3 lsWo = world_to_local(wo, hb);
                    if (p.bounces == 0) p.radiance = p.radiance + p.throughput * emission_material<COUNT>(F, mat_off, tl);   // :241-245
                    const LightSample ls = sample_light<COUNT>(V, F, hb, p.rng, tl);       // :248 -> :69-124
                    p.light_target = ls.target;
                    Tally<false> untallied; untallied.clear();          // the reference evaluates the BSDF only for visible lights (:155-161)
                    p.pending = p.throughput * light_contribution<false>(F, mat_off, hb, wo, ls, untallied);
                    const f3 tz = p.throughput * 0.0f;
                    p.pending_nan = (tz.x != tz.x ? 1 : 0) | (tz.y != tz.y ? 2 : 0) | (tz.z != tz.z ? 4 : 0);
                    f4 bf;
                    const f3 lsWi = sample_material<COUNT>(F, mat_off, lsWo, p.rng, bf, tl);   // :255
                    if (F.wire_opacity > 0.0f) {                                           // :260-270
                        const float w = wireframe_factor(V, F, hb, hit);
                        bf.x *= w; bf.y *= w; bf.z *= w;
                    }
                    const f3 wi = local_to_world(lsWi, hb);                                // :273
                    p.throughput = p.throughput * ((xyz(bf) * gabs(dot(wi, hb.normal))) / bf.w);   // :276
                    p.bsdf_pdf = bf.w;
                    p.next_o = hb.position; p.next_d = wi;
                    begin_segment<COUNT>(V, p, SEG_SHADOW, hb.position, xyz(ls.wl), phase, tl);    // :133
                }
            }
            if (finished) {
                const f4 out = mk4(tonemap(p.radiance), 1.0f);                             // :294-295
                if (L.sum_mode) {
                    p.avg.x += out.x; p.avg.y += out.y; p.avg.z += out.z; p.avg.w += out.w;
                } else {
                    const float n = (float)(L.n_prev + p.pass), n1 = (float)(L.n_prev + p.pass + 1);
                    p.avg.x = (out.x + p.avg.x * n) / n1; p.avg.y = (out.y + p.avg.y * n) / n1;
                    p.avg.z = (out.z + p.avg.z * n) / n1; p.avg.w = (out.w + p.avg.w * n) / n1;
                }
                p.pass++;
                phase = PH_NEW;
            }
        }

        // ---------------- (d) NEW: next pass of the pixel, or a new pixel -----------------------------------------
        {
            bool need_pixel = (phase == PH_NEW) && (p.pass >= L.n_passes);
            if (need_pixel && p.has_pixel) {                       // all passes of this pixel are folded in: write it back
                const size_t pix = (size_t)p.px + (size_t)p.py * (size_t)F.W;
                accum[pix] = p.avg;
                if (primary != nullptr) primary[pix] = p.prim;
                p.has_pixel = false;
            }
            // fetch pixels: one atomic per warp
            for (;;) {
                const unsigned m_need = __ballot_sync(full, need_pixel);
                if (m_need == 0u) break;
                unsigned base = 0;
                const int leader = __ffs(m_need) - 1;
                if (lane == leader) base = atomicAdd(work_counter, (unsigned)__popc(m_need));
                base = __shfl_sync(full, base, leader);
                if (need_pixel) {
                    const unsigned item = base + (unsigned)__popc(m_need & ((1u << lane) - 1u));
                    if (item >= (unsigned)n_items) { phase = PH_EXIT; need_pixel = false; }
                    else {
                        // item -> tile (round-robin over ranks) -> 8x4 pixel blocks inside the tile
                        const int local_tile = (int)(item >> 12), in_tile = (int)(item & 4095u);
                        const int tile = L.tile_rank + local_tile * L.tile_world;
                        const int tx = tile % L.tiles_x, ty = tile / L.tiles_x;
                        const int blk = in_tile >> 5, within = in_tile & 31;
                        p.px = tx * kTile + (blk & 7) * 8 + (within & 7);
                        p.py = ty * kTile + (blk >> 3) * 4 + (within >> 3);
                        if (p.px < F.W && p.py < F.H) {
                            p.avg = accum[(size_t)p.px + (size_t)p.py * (size_t)F.W];
                            p.pass = 0; p.prim = -1; p.has_pixel = true;
                            need_pixel = false;
                        }
                    }
                }
            }
            if (phase == PH_NEW) {
                // pathTracer.fs:172-208: a new sample of pixel (px, py)
                const int sample = L.first_sample + p.pass * L.sample_stride;
                const f3 frag = mk3((float)p.px + 0.5f, (float)p.py + 0.5f, 0.55f);
                p.rng = rng_offset(p.px, p.py, sample, F.noise_w, F.noise_h);              // :174
                p.radiance = mk3(0.0f); p.throughput = mk3(1.0f); p.bounces = 0;
                if (L.integrator != 0) {
                    // edit-mode preview is short and coherent: run it inline (editMode.fs:62-142)
                    int prim = -1;
                    const f4 out = preview_pixel<COUNT>(V, F, p.px, p.py, sample, &prim, tl);
                    p.prim = prim;
                    if (L.sum_mode) { p.avg.x += out.x; p.avg.y += out.y; p.avg.z += out.z; p.avg.w += out.w; }
                    else {
                        const float n = (float)(L.n_prev + p.pass), n1 = (float)(L.n_prev + p.pass + 1);
                        p.avg.x = (out.x + p.avg.x * n) / n1; p.avg.y = (out.y + p.avg.y * n) / n1;
                        p.avg.z = (out.z + p.avg.z * n) / n1; p.avg.w = (out.w + p.avg.w * n) / n1;
                    }
                    p.pass++;
                } else {
                    generate_ray<COUNT>(F, frag, p.rng, p.ro, p.rd, tl);                   // :179
                    const float t = ray_aabb(p.ro, p.rd, V.bmin, V.bmax);                  // :183
                    p.prim = -1;
                    if (t < 0.0f) {                                                        // :187-194
                        p.seg = SEG_PRIMARY; p.status = DDA_NOHIT; p.dda.ix = p.dda.iy = p.dda.iz = 0; p.dda.nanmask = 0;
                        phase = PH_RESOLVE;
                    } else {
                        begin_segment<COUNT>(V, p, SEG_PRIMARY, p.ro + t * p.rd, p.rd, phase, tl);   // :196-202
                    }
                }
            }
        }
    }

    if (COUNT) {
        unsigned long long v[5] = { tl.S, tl.R, tl.H, tl.E, tl.Q };
        #pragma unroll
        for (int i = 0; i < 5; ++i) {
            unsigned long long x = v[i];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(full, x, o);
            v[i] = x;
        }
        if (lane == 0) {
            atomicAdd(&counters->S, v[0]); atomicAdd(&counters->R, v[1]); atomicAdd(&counters->H, v[2]);
            atomicAdd(&counters->E, v[3]); atomicAdd(&counters->Q, v[4]);
        }
    }
}

} // namespace vt
