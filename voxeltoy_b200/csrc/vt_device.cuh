// vt_device.cuh -- device-side data layout and the shared device functions of the
// path tracer: ray generation, slab test, DDA traversal, hit frame, BSDFs, lights.
// Each function cites the reference shader it implements (paths relative to
// /root/reference/src/shaders). Arithmetic contract: vt_math.cuh.
#pragma once
#include "vt_math.cuh"
#include "vt_mem.cuh"

namespace vt {

// ----------------------------------------------------------------------------------
// HBM layout of one scene (all pointers are device pointers owned by vt_ctx)
// ----------------------------------------------------------------------------------
// Occupancy is bit-packed in 4x4x4 bricks: one uint64 per brick, bit (x&3)|(y&3)<<2|(z&3)<<4, bricks x-fastest.
// The brick array is padded by one brick on every side and `bricks` points at brick (0,0,0), so brick coordinates -1 and
// BX.. are addressable; every voxel of the padded array that lies outside the volume has its bit SET (sentinel shell,
// vt_sentinel_kernel). A DDA step changes each coordinate by at most one, so a ray that leaves the volume "hits" the
// shell in its next iteration: the stepping loop carries no bounds test, and the exit is told from a real hit after
// the loop (dda_step). The DDA keeps a brick word in registers and touches memory only when the voxel moves to
// another brick; the per-step arithmetic is identical to dda.h.
// The reference's material-offset grid (R32I, x fastest, -1 empty: renderer.cpp:863-872) is stored as ONE BYTE per voxel: the
// id of the voxel's material record (0xff = empty) + a table id -> offset (vt_volume_upload builds both; volumes with more
// than 255 distinct records use 16-bit ids, 0xffff = empty). 128 MiB instead of 512 MiB at 512^3, 1 GiB instead of 4 GiB at
// 1024^3; it is read once per surface hit (fetch_offset) and the table is a few cache lines.
struct Volume {
    const uint8_t* __restrict__ ids8;      // X*Y*Z material ids, x fastest (nullptr when the volume uses 16-bit ids)
    const uint16_t* __restrict__ ids16;
    const int32_t* __restrict__ id_offset; // id -> offset into the material array
    const unsigned long long* __restrict__ bricks;   // brick (0,0,0) of the padded array
    const unsigned long long* __restrict__ bricks_top;   // brick (-1,-1,-1) = the first word of the padded array; dda_step indexes downwards from here with its tag
    int X, Y, Z;
    int BX, BXY;                           // strides of the padded brick array
    int nBX, nBXY;                         // -BX, -BXY (dda_step)
    // empty-space skip field (dda_skip), one byte per 4^3 brick (unpadded, x fastest): bits 2..7 = k8, the Chebyshev distance
    // field over 8^3-voxel cells (0 = the cell holds a solid voxel, k >= 1 = every cell within distance k-1 is empty, capped);
    // bits 0..1 = k4, the same over 4^3 bricks capped at 3 (the near field: thin shells leave rays within two 8^3 cells of a
    // surface for long stretches). nullptr = skipping disabled for this volume.
    const unsigned char* __restrict__ skip; int SX, SXY;
    f3 bmin, bmax, vsize, resf;            // world bounds, wsVoxelSize, vec3(voxelResolution)
    f3 inv_extent;                         // 1.0 / (bmax - bmin)  (dda.h:16, hoisted: same IEEE divisions on the host)
    int max_steps;                         // dda.h:98
};

// state the reference keeps in two SSBOs (focalDistanceDevice.h, selectVoxelDevice.h)
struct Shared {
    float focal_distance;
    int sel_index[4];
    float sel_normal[4];
};

struct Frame {
    float inv_mv[16], proj[16], inv_proj[16];
    float near_z;
    float lens_radius;
    int lens_model;
    int W, H;
    int max_bounces;
    f3 bg_top, bg_bottom;
    int use_image;
    float env_rotation, env_integral;
    float wire_opacity, wire_thickness;
    const float4* __restrict__ noise; int noise_w, noise_h;
    const float* __restrict__ materials; int n_materials;
    const int32_t* __restrict__ emissive; int n_emissive;
    const float4* __restrict__ env; int env_w, env_h;       // rgb padded to float4
    const float* __restrict__ cdf_u; int cdf_u_w, cdf_u_h;
    const float* __restrict__ cdf_v; int cdf_v_n;
    // guide tables of the two CDF searches (built by vt_env_upload when the CDFs are sorted; guide_k = 0: absent)
    const unsigned short* __restrict__ guide_v;   // guide_k + 1 entries
    const unsigned short* __restrict__ guide_u;   // (cdf_v_n - 1) rows x (guide_k + 1) entries
    int guide_k;
    const Shared* __restrict__ shared;
};

struct Basis { f3 position, tangent, normal, binormal; };   // coordinates.h:35-41

struct Counters { unsigned long long S, R, H, E, Q; };

template <bool COUNT> struct Tally {
    unsigned int S, R, H, E, Q;
    VT_DEV void clear() { S = R = H = E = Q = 0; }
};

#define VT_TALLY(field, n) do { if (COUNT) tl.field += (n); } while (0)

// ----------------------------------------------------------------------------------
// random.h
// ----------------------------------------------------------------------------------
VT_DEV int wang_hash(int seed)                                   // random.h:3-11
{
    seed = (seed ^ 61) ^ (seed >> 16);
    seed = (int)((unsigned)seed * 9u);
    seed = seed ^ (seed >> 4);
    seed = (int)((unsigned)seed * 0x27d4eb2du);
    seed = seed ^ (seed >> 15);
    return seed;
}
// C / GLSL integer division and remainder (truncation toward zero, remainder with the sign of the dividend) by 2^k
VT_DEV int div_pow2(int a, int k) { const unsigned m = (unsigned)(a >> 31); const unsigned q = (((unsigned)a ^ m) - m) >> k; return (int)((q ^ m) - m); }
VT_DEV int rem_pow2(int a, int mask) { const unsigned m = (unsigned)(a >> 31); const unsigned r = (((unsigned)a ^ m) - m) & (unsigned)mask; return (int)((r ^ m) - m); }
VT_DEV int2 rng_offset_from(int pixel_hash, int sequence, int rw, int rh)   // random.h:16-17 given hash(x + y * w)
{
    const int offset = pixel_hash ^ wang_hash(sequence);
    // the noise table is 1024 x 1024 unless the caller uploaded another one: shifts and masks instead of three ~25-instruction
    // signed divisions (same values for every int, negative offsets included)
    if (rw > 0 && rh > 0 && (rw & (rw - 1)) == 0 && (rh & (rh - 1)) == 0)
        return make_int2(rem_pow2(offset, rw - 1), rem_pow2(div_pow2(offset, __ffs(rw) - 1), rh - 1));
    return make_int2(offset % rw, (offset / rw) % rh);
}
VT_DEV int rng_pixel_hash(int px, int py, int rw) { return wang_hash((int)((unsigned)px + (unsigned)py * (unsigned)rw)); }   // random.h:15
VT_DEV int2 rng_offset(int px, int py, int sequence, int rw, int rh)   // random.h:13-18
{
    return rng_offset_from(rng_pixel_hash(px, py, rw), sequence, rw, rh);
}
template <bool COUNT>
VT_DEV f4 rng_next(const Frame& F, int2& off, Tally<COUNT>& tl)        // random.h:20-27
{
    f4 r = mk4(0.f, 0.f, 0.f, 0.f);
    if (off.x >= 0 && off.y >= 0 && off.x < F.noise_w && off.y < F.noise_h) {
        const float4 t = ldg_keep(F.noise + ((size_t)off.x + (size_t)off.y * (size_t)F.noise_w));
        r = mk4(t.x, t.y, t.z, t.w);
    }
    // (x + 1) % w and (y + 1) % h for 0 <= x < w, 0 <= y < h (rng_offset and this function keep them there)
    off.x = (off.x + 1 == F.noise_w) ? 0 : off.x + 1;
    if (off.x == 0) off.y = (off.y + 1 == F.noise_h) ? 0 : off.y + 1;
    VT_TALLY(R, 1);
    return r;
}

// hint: bring the line holding p towards L1 (random 4-16 B gathers whose address is known long before the value is needed)
VT_DEV void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }

// ----------------------------------------------------------------------------------
// aabb.h:1-32
// ----------------------------------------------------------------------------------
VT_DEV float ray_aabb(f3 o, f3 d, f3 bmin, f3 bmax)
{
    const f3 inv = mk3(1.0f) / d;
    const f3 t0 = (bmin - o) * inv;
    const f3 t1 = (bmax - o) * inv;
    const f3 tmin = gmin(t0, t1);
    const float tminf = gmax(tmin.x, gmax(tmin.y, tmin.z));
    const f3 tmax = gmax(t0, t1);
    const float tmaxf = gmin(tmax.x, gmin(tmax.y, tmax.z));
    if (tmaxf < 0.0f) return -1.0f;
    if (tminf > tmaxf) return -1.0f;
    return gmax(0.0f, tminf);
}

// ----------------------------------------------------------------------------------
// dda.h
// ----------------------------------------------------------------------------------
VT_DEV bool out_of_grid(f3 p, f3 resf)
{
    return (p.x < 0.0f) || (p.y < 0.0f) || (p.z < 0.0f) || (p.x >= resf.x) || (p.y >= resf.y) || (p.z >= resf.z);
}
VT_DEV int fetch_id(const Volume& V, size_t i)                    // material id of voxel i, -1 = empty
{
    if (V.ids8 != nullptr) { const int v = (int)__ldg(V.ids8 + i); return v == 0xff ? -1 : v; }
    const int v = (int)__ldg(V.ids16 + i);
    return v == 0xffff ? -1 : v;
}
VT_DEV int fetch_offset(const Volume& V, int x, int y, int z)     // texelFetch(materialOffsetTexture); out of range -> 0
{
    if ((unsigned)x >= (unsigned)V.X || (unsigned)y >= (unsigned)V.Y || (unsigned)z >= (unsigned)V.Z) return 0;
    const int id = fetch_id(V, (size_t)x + (size_t)y * (size_t)V.X + (size_t)z * (size_t)V.X * (size_t)V.Y);
    return id < 0 ? -1 : __ldg(V.id_offset + id);
}
VT_DEV void prefetch_id(const Volume& V, size_t i)
{
    prefetch_l1(V.ids8 != nullptr ? (const void*)(V.ids8 + i) : (const void*)(V.ids16 + i));
}
VT_DEV void store_id(void* ids, int id_bytes, size_t i, int id)   // id < 0: empty
{
    if (id_bytes == 1) reinterpret_cast<uint8_t*>(ids)[i] = (uint8_t)(id < 0 ? 0xff : id);
    else reinterpret_cast<uint16_t*>(ids)[i] = (uint16_t)(id < 0 ? 0xffff : id);
}

// Literal float-state restatement of dda.h:38-57, used only when the start voxel is NaN
// (the integer-state loop below cannot represent it). Never taken by finite rays.
template <bool COUNT>
__device__ __noinline__ bool raymarch_slow(const Volume& V, f3 vp, f3 dis, f3 sg, f3 inc, f3& hit_pos, Tally<COUNT>& tl)
{
    bool isect = false;
    for (int steps = 0; steps < V.max_steps; ++steps) {
        if (out_of_grid(vp, V.resf)) break;
        VT_TALLY(S, 1);
        if (fetch_offset(V, f2i(vp.x), f2i(vp.y), f2i(vp.z)) >= 0) { isect = true; break; }
        const f3 mask = mk3(gstep(dis.x, dis.y) * gstep(dis.x, dis.z),
                            gstep(dis.y, dis.x) * gstep(dis.y, dis.z),
                            gstep(dis.z, dis.y) * gstep(dis.z, dis.x));
        dis = dis + (mask * sg) * inc;
        vp = vp + mask * sg;
    }
    hit_pos = vp;
    return isect;
}

// ---- DDA state machine ------------------------------------------------------------------------
// The traversal of dda.h:7-61 split into begin / step so that the persistent kernel can interleave the
// steps of 32 independent rays (vt_pathstate.cuh) while the simple kernels run it to completion.
// Integer voxel coordinates are exact (0 <= vp < res once the start voxel passed the bounds test);
// (mask*sign)*inc of dda.h:52 is +-inc = |1/d| for a stepping axis and +-0 otherwise, so the float
// state `dis` is advanced by exactly the additions the shader performs.
struct Dda {
    // voxel position (hit position when done), kept the way the stepping loop wants it: complemented and pre-shifted into the
    // fields of the in-brick bit index -- cx = ~ix, cy = ~iy << 2, cz = ~iz << 4 -- so that the index of the voxel's bit, counted
    // from the sign position, is three masked ORs: (cx & 3) | (cy & 0xc) | (cz & 0x30) = 63 - ((ix&3) | (iy&3)<<2 | (iz&3)<<4)
    int cx, cy, cz;
    float dx, dy, dz;          // dis
    float ex, ey, ez;          // sign*inc per axis
    int nsx, nsy, nsz;         // what a step adds to cx, cy, cz: -sign, -4 sign, -16 sign
    int steps;
    int bkey;                  // cached brick tag (-1 none)
    unsigned long long brick;  // cached 4x4x4 occupancy word
    int nanmask;               // bit i: component i of the (float) position is NaN (only via the slow path)
    VT_DEV int ix() const { return ~cx; }
    VT_DEV int iy() const { return ~(cy >> 2); }
    VT_DEV int iz() const { return ~(cz >> 4); }
    VT_DEV void set_pos(int x, int y, int z) { cx = ~x; cy = (~y) << 2; cz = (~z) << 4; }
    VT_DEV void set_sign(int sx, int sy, int sz) { nsx = -sx; nsy = -4 * sy; nsz = -16 * sz; }
    VT_DEV bool pos_x() const { return nsx < 0; }     // the ray moves towards +x
    VT_DEV bool pos_y() const { return nsy < 0; }
    VT_DEV bool pos_z() const { return nsz < 0; }
    VT_DEV void advance(int kx, int ky, int kz) { cx += nsx * kx; cy += nsy * ky; cz += nsz * kz; }   // kx steps along x, ...
};
enum { DDA_RUNNING = 0, DDA_HIT = 1, DDA_NOHIT = 2 };
constexpr int kNoBrick = -1;             // Dda::bkey when no brick word is cached (brick indices are >= 0)

VT_DEV f3 dda_position(const Dda& s)
{
    const float qn = __int_as_float(0x7fc00000);
    return mk3((s.nanmask & 1) ? qn : (float)s.ix(), (s.nanmask & 2) ? qn : (float)s.iy(), (s.nanmask & 4) ? qn : (float)s.iz());
}

// dda.h:16-34. Returns DDA_RUNNING when stepping must follow, else the final status (position in s).
template <bool COUNT>
VT_DEV int dda_begin(const Volume& V, f3 o, f3 d, Dda& s, Tally<COUNT>& tl)
{
    s.set_pos(0, 0, 0); s.nanmask = 0; s.steps = 0; s.bkey = kNoBrick; s.brick = 0ull;   // hit_pos = 0 on the early-out path (contract U1)
    o = o + gsign(d) * 0.001f;                                    // :19
    const f3 vo = ((o - V.bmin) * V.inv_extent) * V.resf;         // :16,:20
    const f3 vp = gfloor(vo);                                     // :22
    if (out_of_grid(vp, V.resf)) return DDA_NOHIT;                // :24-25
    // :29  mix(d, 1e-5, step(abs(d), 1e-5))
    d = mk3(gclamp_dir(d.x), gclamp_dir(d.y), gclamp_dir(d.z));
    const f3 inc = mk3(1.0f) / d;                                 // :31
    const f3 sg = gsign(d);                                       // :32
    const f3 dis = (((vp - vo) + 0.5f) + sg * 0.5f) * inc;        // :34
    // The stepping loop below drops the step cap of :38/:98: every iteration moves at least one coordinate by one voxel
    // towards the outside (sign(d) = +-1 after the clamp of :29), so a ray leaves the grid within X+Y+Z iterations, and
    // X+Y+Z <= sqrt(3) |res| < 2 ceil(|res|) = the cap. The exceptions -- a NaN start voxel (passes the bounds test of :24)
    // or a NaN direction component (sign = 0, the ray may not move) -- take the literal loop with the cap.
    if (vp.x != vp.x || vp.y != vp.y || vp.z != vp.z || d.x != d.x || d.y != d.y || d.z != d.z) {
        f3 hp;
        const bool isect = raymarch_slow<COUNT>(V, vp, dis, sg, inc, hp, tl);
        s.nanmask = (hp.x != hp.x ? 1 : 0) | (hp.y != hp.y ? 2 : 0) | (hp.z != hp.z ? 4 : 0);
        s.set_pos(f2i(hp.x), f2i(hp.y), f2i(hp.z));
        return isect ? DDA_HIT : DDA_NOHIT;
    }
    s.set_pos(f2i(vp.x), f2i(vp.y), f2i(vp.z));
    s.set_sign(f2i(sg.x), f2i(sg.y), f2i(sg.z));
    s.ex = sg.x * inc.x; s.ey = sg.y * inc.y; s.ez = sg.z * inc.z;
    s.dx = dis.x; s.dy = dis.y; s.dz = dis.z;
    return DDA_RUNNING;
}

// one iteration of the while loop of dda.h:38-57. Branch-free apart from the two exits: the 4x4x4 occupancy word is
// re-loaded under a predicate when the voxel moved to another brick (about 4 steps in 10), so a warp whose lanes are at
// different places in their bricks still issues one common instruction stream per step.
template <bool COUNT>
VT_DEV int dda_step(const Volume& V, Dda& s, Tally<COUNT>& tl)
{
    // :41-42 the bounds test is implied: outside the volume the occupancy bit is the sentinel (see Volume)
    // index of the brick word, counted from brick (-1,-1,-1) = the first word of the padded array, from the complemented coordinates:
    // (ix>>2) + 1 = -(cx>>2), so index = -(cx>>2) - (cy>>4) BX - (cz>>6) BXY (nBX, nBXY are the negated strides)
    const int key = ((s.cy >> 4) * V.nBX - (s.cx >> 2)) + (s.cz >> 6) * V.nBXY;
    if (key != s.bkey) { s.bkey = key; s.brick = __ldg(V.bricks_top + key); }
    // the voxel's bit is moved to the SIGN position (shift left by 63 - bit), so the test is one compare instead of AND + compare;
    // the shift count is three masked ORs of the position words (see Dda). Kept opaque so that nvcc does not rewrite it.
    int occ;
    asm("{\n\t"
        ".reg .b64 t;\n\t"
        ".reg .b32 lo, n;\n\t"
        "lop3.b32 n, %2, 3, 0, 0xc0;\n\t"          // cx & 3
        "lop3.b32 n, %3, 0xc, n, 0xea;\n\t"        // (cy & 0xc) | n
        "lop3.b32 n, %4, 0x30, n, 0xea;\n\t"       // (cz & 0x30) | n
        "shl.b64 t, %1, n;\n\t"
        "mov.b64 {lo, %0}, t;\n\t"
        "}" : "=r"(occ) : "l"(s.brick), "r"(s.cx), "r"(s.cy), "r"(s.cz));
    if (occ < 0) {                                                // :44-50, or the voxel is outside: :41-42
        const bool inside = (unsigned)s.ix() < (unsigned)V.X && (unsigned)s.iy() < (unsigned)V.Y && (unsigned)s.iz() < (unsigned)V.Z;
        if (inside) VT_TALLY(S, 1);
        return inside ? DDA_HIT : DDA_NOHIT;
    }
    VT_TALLY(S, 1);
    // :51 mask = step(dis.xyz, dis.yxy) * step(dis.xyz, dis.zzx)   (ties step several axes at once)
    // Same arithmetic, instruction selection pinned: wf_trace is bound by instruction issue (80 % of the slots, profiles/r02_v19_*),
    // so the axis masks are one 3-way minimum + three equality tests (an axis steps iff its dis IS the minimum: dis holds no NaN
    // here, dda_begin) and the position updates are predicated multiply-adds, which issue on the FMA pipe.
    asm volatile("{\n\t"
                 ".reg .pred px, py, pz;\n\t"
                 ".reg .f32 m;\n\t"
                 "min.f32 m, %3, %4;\n\t"
                 "min.f32 m, m, %5;\n\t"
                 "setp.eq.f32 px, %3, m;\n\t"
                 "setp.eq.f32 py, %4, m;\n\t"
                 "setp.eq.f32 pz, %5, m;\n\t"
                 "@px add.rn.f32 %3, %3, %6;\n\t"
                 "@py add.rn.f32 %4, %4, %7;\n\t"
                 "@pz add.rn.f32 %5, %5, %8;\n\t"
                 "@px mad.lo.s32 %0, %9, 1, %0;\n\t"
                 "@py mad.lo.s32 %1, %10, 1, %1;\n\t"
                 "@pz mad.lo.s32 %2, %11, 1, %2;\n\t"
                 "}"
                 : "+r"(s.cx), "+r"(s.cy), "+r"(s.cz), "+f"(s.dx), "+f"(s.dy), "+f"(s.dz)
                 : "f"(s.ex), "f"(s.ey), "f"(s.ez), "r"(s.nsx), "r"(s.nsy), "r"(s.nsz));
    return DDA_RUNNING;                                           // :38,:55 the cap cannot be reached here, see dda_begin
}

// ---- exact empty-space skip ----------------------------------------------------------------------------------
// "Skipping must not change hit results", and dis advances by repeated ROUNDED additions (dda.h:52), so no closed form may
// replace the steps. What can be dropped is everything else a step does. While the ray is inside a box of voxels known to
// be empty, the iteration order of dda.h:51 is a merge of three increasing sequences D_a(j+1) = fl(D_a(j) + e_a): each
// iteration advances every axis whose D equals the current minimum. The state reached after all values <= tau have been
// processed is therefore, for ANY threshold tau, (i_a + s_a k_a, D_a(k_a)) with k_a = #{j : D_a(j) <= tau} -- computable
// per axis with the same float additions and nothing else. tau is placed just below the first value at which some axis would
// leave the box; the per-axis step limits verify that no axis left (otherwise the skip is abandoned with the state
// untouched), and the ordinary loop takes over for the last few steps.
// The box comes from the distance fields of Volume: a cube of (2k-1) cells around the ray's cell, clipped to the volume.
//
// advance_binade: d <- fl(d + e) while d <= tau, at most nmax times in total (k counts the additions), for a threshold tau that
// does not lie beyond the binade of d: every addition that is made then has its operand in that one binade. Round 2 first
// performed the additions one by one (2 instructions each): on the 512^3 bunny a skip replaces 35 DDA iterations on average,
// the lanes of a warp need very different numbers of them, and the addition loops ran at 6 of 32 lanes -- 40 % of wf_trace's
// instructions (profiles/r02_c3_v18_*). Now a run of additions is ONE integer multiply-add on the bit pattern, in straight-line
// code that every lane of the warp executes together:
//   d = M u with u the unit in the last place of the binade and M the 24-bit significand; e = (q + f) u, 0 <= f < 1. While
//   the sum stays in the binade, fl(d + e) = (M + q + [f > 1/2]) u; on a tie (f = 1/2) the sum goes to the EVEN neighbour.
//   After one addition made inside the binade the significand is even in the tie case, and from an even M a tie adds
//   q + (q & 1), leaving M even: from then on the bit pattern moves by a constant `inc` per addition, tie or no tie.
// So: three real additions (operands d, d1, d2 inside the binade, hence d2 settled), inc = bits(d3) - bits(d2), then
// floor((bits(tau) - bits(d3)) / inc) further additions in one step (each starts at or below tau and ends at or below tau, i.e.
// inside the binade), then the one real addition that carries d beyond tau -- possibly into the next binade, which is why it is
// a real one. A degenerate increment (e below half an ulp of d: d would never move) leaves the run unfinished and the caller
// abandons the skip. Checked against the literal loop on the device (vt_debug_advance, tests/test_gpu_configs.py) on 400 k realistic
// and adversarial operand sets, and on the CPU by tests/test_skip_closed_form.py (the same integer algebra in numpy).
VT_DEV void advance_binade(float& d, int& k, const float e, const float tau, const int nmax)
{
    if (d <= tau && k < nmax) { d = d + e; k += 1; }
    if (d <= tau && k < nmax) { d = d + e; k += 1; }
    const int b2 = __float_as_int(d);
    const bool third = d <= tau && k < nmax;
    if (third) { d = d + e; k += 1; }
    const int b3 = __float_as_int(d);
    const int inc = b3 - b2;
    // qd = floor(N / inc), 0 <= N < 2^23: approximate quotient (one MUFU), exact after one correction each way; a quotient too
    // large for that (>= 2^20) is clamped to nmax - k below whatever its last digits are
    const int N = __float_as_int(tau) - b3, ic = max(inc, 1);
    float rcp;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"((float)ic));
    int qd = __float2int_rz((float)N * rcp);
    const int rem = N - qd * ic;
    if (rem < 0) qd -= 1;
    if (rem >= ic) qd += 1;
    const int J = min(qd, nmax - k);
    if (third && d <= tau && inc > 0 && J > 0) { d = __int_as_float(b3 + J * inc); k += J; }
    if (d <= tau && k < nmax) { d = d + e; k += 1; }
}
// The additions themselves, four per trip (the values increase, so four may be taken at once iff the third result is still <= tau):
// 2 instructions per addition. Right for COHERENT rays -- the primary rays wf_generate traces in lockstep, whose lanes need about
// the same number of additions -- and free of the closed form's one-binade limit, which costs rays that start at t = 0 several calls.
VT_DEV int advance_additions(float& d, float e, float tau, int nmax)
{
    int k = 0;
    while (k + 4 <= nmax) {
        const float a1 = d + e, a2 = a1 + e, a3 = a2 + e;
        if (!(a3 <= tau)) break;
        d = a3 + e; k += 4;
    }
    while (d <= tau && k < nmax) { d = d + e; ++k; }
    return k;
}
// last float of the binade of a non-negative t (of the denormal range for t < 2^-126: the bit patterns are linear there too)
VT_DEV float binade_end(float t) { return __int_as_float(__float_as_int(t) | 0x7fffff); }

// the test hook's single-axis form: binade after binade
VT_DEV int advance_until(float& d, float e, float tau, int nmax)
{
    int k = 0;
    while (d <= tau && k < nmax) {
        const int k0 = k;
        if (__float_as_int(d) >= 0) advance_binade(d, k, e, gmin(tau, binade_end(d)), nmax);
        if (k == k0) { d = d + e; k += 1; }                       // -0 (its "binade" lies below it): one real addition
    }
    return k;
}

// skip-field byte of the ray's brick (0 when the ray is outside the volume); a skip is worth trying when either level is >= 2
VT_DEV int dda_skip_radius(const Volume& V, const Dda& s)
{
    const int ix = s.ix(), iy = s.iy(), iz = s.iz();
    if ((unsigned)ix >= (unsigned)V.X || (unsigned)iy >= (unsigned)V.Y || (unsigned)iz >= (unsigned)V.Z) return 0;
    return (int)__ldg(V.skip + ((ix >> 2) + (iy >> 2) * V.SX + (iz >> 2) * V.SXY));
}
#ifndef VT_SKIP_MAX_ADD
#define VT_SKIP_MAX_ADD 128
#endif
#ifndef VT_SKIP_PASSES
#define VT_SKIP_PASSES 1
#endif
VT_DEV bool dda_skip_wanted(int v) { return v >= (2 << 2) || (v & 3) >= 2; }
// v = dda_skip_radius with dda_skip_wanted(v); returns the number of steps skipped
template <bool COHERENT>
VT_DEV int dda_skip(const Volume& V, Dda& s, int v)
{
    // the far field (8^3 cells) when it allows a skip at all, else the near field (4^3 bricks)
    const int k8 = v >> 2;
    const int shift = (k8 >= 2) ? 3 : 2, k = (k8 >= 2) ? k8 : (v & 3);
    const int cell = (1 << shift) - 1;
    const int ix = s.ix(), iy = s.iy(), iz = s.iz();
    const int cx = ix >> shift, cy = iy >> shift, cz = iz >> shift;
    const int r = (k - 1) << shift;
    // steps that stay inside, per axis; at most VT_SKIP_MAX_ADD of them are taken in one go (any smaller box is as valid: the
    // additions of the whole warp run in lockstep, and one lane with a 127-step run would keep the others waiting)
    const int nx = min(VT_SKIP_MAX_ADD, s.pos_x() ? min((cx << shift) + cell + r, V.X - 1) - ix : ix - max((cx << shift) - r, 0));
    const int ny = min(VT_SKIP_MAX_ADD, s.pos_y() ? min((cy << shift) + cell + r, V.Y - 1) - iy : iy - max((cy << shift) - r, 0));
    const int nz = min(VT_SKIP_MAX_ADD, s.pos_z() ? min((cz << shift) + cell + r, V.Z - 1) - iz : iz - max((cz << shift) - r, 0));
    if (min(nx, min(ny, nz)) < 3) return 0;
    // first value at which axis a would step OUT of the box: D_a(n_a) ~ d_a + n_a e_a; stay 0.1 % below the smallest
    const float tau = gmin(s.dx + (float)nx * s.ex, gmin(s.dy + (float)ny * s.ey, s.dz + (float)nz * s.ez)) * 0.999f;
    if (!(tau > gmin(s.dx, gmin(s.dy, s.dz))) || !(tau < 3.0e38f)) return 0;
    // Incoherent rays (wf_trace): one pass = the run up to the end of the binade of the smallest dis (or to tau), straight-line
    // code for the whole warp. More passes per call (VT_SKIP_PASSES) or calls in a row were slower on the 512^3 bunny and the 256^3
    // terrain alike (trace 68.9 -> 81.6 ms with two passes): a ray that stopped at a binade boundary continues at its next skip.
    // A pass that leaves an axis unfinished (the box ended before tau, or a degenerate increment) is dropped.
    float dx = s.dx, dy = s.dy, dz = s.dz;
    int kx = 0, ky = 0, kz = 0, done = 0;
    if (COHERENT) {
        kx = advance_additions(dx, s.ex, tau, nx);
        ky = advance_additions(dy, s.ey, tau, ny);
        kz = advance_additions(dz, s.ez, tau, nz);
        if (dx <= tau || dy <= tau || dz <= tau) return 0;        // an axis ran out of box before tau: abandon, state untouched
        done = 1;
    } else
    #pragma unroll 1
    for (int pass = 0; pass < VT_SKIP_PASSES; ++pass) {
        const float tmin = gmin(dx, gmin(dy, dz));
        const float tc = gmin(tau, binade_end(tmin));
        if (!(tc > tmin) || __float_as_int(tmin) < 0) break;
        float ax = dx, ay = dy, az = dz;
        int jx = kx, jy = ky, jz = kz;
        advance_binade(ax, jx, s.ex, tc, nx);
        advance_binade(ay, jy, s.ey, tc, ny);
        advance_binade(az, jz, s.ez, tc, nz);
        if (ax <= tc || ay <= tc || az <= tc) break;
        dx = ax; dy = ay; dz = az; kx = jx; ky = jy; kz = jz; done = 1;
        if (tc == tau) break;
    }
    if (!done) return 0;
    s.dx = dx; s.dy = dy; s.dz = dz;
    s.advance(kx, ky, kz);
    return kx + ky + kz;
}

// dda.h:7-61 run to completion
template <bool COUNT>
VT_DEV bool raymarch(const Volume& V, f3 o, f3 d, f3& hit_pos, Tally<COUNT>& tl)
{
    Dda s;
    int st = dda_begin<COUNT>(V, o, d, s, tl);
    int guard = V.X + V.Y + V.Z + 8;                              // never reached (dda_begin): belt and braces against a hang
    while (st == DDA_RUNNING && --guard >= 0) st = dda_step<COUNT>(V, s, tl);
    if (st == DDA_RUNNING) st = DDA_NOHIT;
    hit_pos = dda_position(s);
    return st == DDA_HIT;
}

// dda.h:63-100
template <bool COUNT>
VT_DEV bool traverse(const Volume& V, f3 o, f3 d, f3& hit_pos, bool& hit_ground, Tally<COUNT>& tl)
{
    if (raymarch<COUNT>(V, o, d, hit_pos, tl)) { hit_ground = false; return true; }
    hit_ground = (hit_pos.y < 0.0f);
    return hit_ground;
}

// ----------------------------------------------------------------------------------
// coordinates.h
// ----------------------------------------------------------------------------------
VT_DEV f4 screen_to_eye_persp(const Frame& F, f3 ws)              // coordinates.h:12-27
{
    const float vx = 0.0f, vy = 0.0f, vz = (float)F.W, vw = (float)F.H;
    f3 ndc;
    ndc.x = ((2.0f * ws.x) - (2.0f * vx)) / vz - 1.0f;
    ndc.y = ((2.0f * ws.y) - (2.0f * vy)) / vw - 1.0f;
    ndc.z = (2.0f * ws.z - 0.0f - 1.0f) / (1.0f - 0.0f);
    // GLSL cameraProj[3][2] = host pm.x[2][3], [2][2] = pm.x[2][2], [2][3] = pm.x[3][2]
    const float cw = F.proj[11] / (ndc.z - (F.proj[10] / F.proj[14]));
    return mul44(F.inv_proj, ndc.x * cw, ndc.y * cw, ndc.z * cw, cw);
}
VT_DEV f4 screen_to_eye_ortho(const Frame& F, f3 ws)              // coordinates.h:1-10
{
    f3 ndc;
    ndc.x = (ws.x / (float)F.W) * 2.0f - 1.0f;
    ndc.y = (ws.y / (float)F.H) * 2.0f - 1.0f;
    ndc.z = (2.0f * ws.z - 0.0f - 1.0f) / (1.0f - 0.0f);
    return mul44(F.inv_proj, ndc.x, ndc.y, ndc.z, 1.0f);
}
VT_DEV f3 local_to_world(f3 v, const Basis& b)                    // coordinates.h:43-49
{
    return (v.x * b.tangent + v.y * b.normal) + v.z * b.binormal;
}
VT_DEV f3 world_to_local(f3 v, const Basis& b)                    // coordinates.h:51-57
{
    return mk3(dot(v, b.tangent), dot(v, b.normal), dot(v, b.binormal));
}
VT_DEV void voxel_to_world(const Volume& V, f3 vsP, f3 o, f3 d, Basis& b)   // coordinates.h:59-82
{
    const f3 vmin = vsP * V.vsize + V.bmin;
    const f3 vmax = vmin + V.vsize;
    const float t = ray_aabb(o, d, vmin, vmax);
    b.position = o + d * t;
    const f3 center = vmin + V.vsize * 0.5f;
    const f3 h = b.position - center;
    const f3 a = gabs(h);
    const f3 mask = mk3(gstep(a.y, a.x) * gstep(a.z, a.x),
                        gstep(a.x, a.y) * gstep(a.z, a.y),
                        gstep(a.x, a.z) * gstep(a.y, a.z));
    b.normal = mask * gsign(h);
    const f3 tan0 = cross(b.normal, mk3(0.57735026919f));
    b.binormal = normalize(cross(tan0, b.normal));
    b.tangent = normalize(cross(b.binormal, b.normal));
}
VT_DEV f3 spherical(float phi, float cosT, float sinT)            // coordinates.h:91-95
{
    float s, c; gsincos(phi, s, c);
    return mk3(sinT * c, cosT, sinT * s);
}
VT_DEV f2 uv_from_vector(f3 v, float rot)                         // coordinates.h:97-109
{
    const float theta = gacos(v.y);
    const float phi = gatan2(v.z, v.x) + VT_PI + rot;
    return mk2(gmod(gdiv_two_pi(phi), 1.0f), gdiv_pi(theta));
}
VT_DEV f3 direction_from_uv(f2 uv, float rot)                     // coordinates.h:111-116 + :85-89
{
    const float phi = uv.x * VT_TWO_PI - VT_PI - rot;
    const float theta = uv.y * VT_PI;
    float st, ct; gsincos(theta, st, ct);
    float sp, cp; gsincos(phi, sp, cp);
    return mk3(st * cp, ct, st * sp);
}

// ----------------------------------------------------------------------------------
// sampling.h
// ----------------------------------------------------------------------------------
VT_DEV f2 sample_disk(float ux, float uy)                         // sampling.h:2-7
{
    const float r = sqrtf(ux);
    const float theta = 2.0f * VT_PI * uy;
    float s, c; gsincos(theta, s, c);
    return mk2(r * c, r * s);
}
VT_DEV f4 cosine_hemisphere(float ux, float uy)                   // sampling.h:13-22
{
    const f2 p = sample_disk(ux, uy);
    const float y = sqrtf(gmax(0.0f, 1.0f - p.x * p.x - p.y * p.y));
    return mk4(p.x, y, p.y, gdiv_pi(y));
}
VT_DEV f4 uniform_hemisphere(float ux, float uy)                  // sampling.h:27-36
{
    const float y = ux;
    const float r = sqrtf(gmax(0.0f, 1.0f - y * y));
    const float phi = uy * 2.0f * VT_PI;
    float s, c; gsincos(phi, s, c);
    return mk4(r * c, y, r * s, 1.0f / (2.0f * VT_PI));
}
VT_DEV float power_heuristic(float f, float g) { return f * f / (f * f + g * g); }   // sampling.h:40-43

// ----------------------------------------------------------------------------------
// generateRay.h
// ----------------------------------------------------------------------------------
VT_DEV void pinhole_ray(const Frame& F, f3 frag, float ux, float uy, f3& ro, f3& rd)   // generateRay.h:10-26
{
    const f3 jitter = mk3(ux - 0.5f, uy - 0.5f, 0.0f);
    const f4 es = screen_to_eye_persp(F, frag + jitter);
    const f4 o = mul44(F.inv_mv, 0.0f, 0.0f, 0.0f, 1.0f);
    const f3 n = normalize(xyz(es));
    const f4 d = mul44(F.inv_mv, n.x, n.y, n.z, 0.0f);
    const float len = sqrtf(((d.x * d.x + d.y * d.y) + d.z * d.z) + d.w * d.w);     // normalize(vec4)
    ro = xyz(o);
    rd = mk3(d.x / len, d.y / len, d.z / len);
}
// generateRay.h:30-101 (thin lens), split at its only sample-dependent input, the point on the lens: there is no pixel jitter in
// this model, so the pixel's focal point (:60-84) is the same in every pass and wf_generate computes it once per pixel.
VT_DEV f4 thin_lens_focal_point(const Frame& F, f3 frag)
{
    const float fd = F.shared->focal_distance;
    const f3 es = xyz(screen_to_eye_persp(F, frag));
    const f3 esd = normalize(es);
    const float t = fd / -(esd.z);
    const f3 focal = es + t * esd;
    return mul44(F.inv_mv, focal.x, focal.y, focal.z, 1.0f);
}
VT_DEV void thin_lens_ray(const Frame& F, f4 fp, float ux, float uy, f3& ro, f3& rd)
{
    const f2 disk = sample_disk(ux, uy);
    const f4 o = mul44(F.inv_mv, disk.x * F.lens_radius, disk.y * F.lens_radius, 0.0f, 1.0f);
    ro = xyz(o);
    rd = normalize(xyz(fp) - ro);
}
template <bool COUNT>
VT_DEV void generate_ray(const Frame& F, f3 frag, int2& rng, f3& ro, f3& rd, Tally<COUNT>& tl)   // generateRay.h:104-123
{
    if (F.lens_model == 0) {
        const f4 u = rng_next<COUNT>(F, rng, tl);
        pinhole_ray(F, frag, u.x, u.y, ro, rd);
    } else if (F.lens_model == 1) {                               // generateRay.h:30-101
        const f4 u = rng_next<COUNT>(F, rng, tl);
        thin_lens_ray(F, thin_lens_focal_point(F, frag), u.x, u.y, ro, rd);
    } else {                                                      // generateRay.h:1-7
        const f4 e = screen_to_eye_ortho(F, frag);
        ro = xyz(mul44(F.inv_mv, e.x, e.y, e.z, e.w));
        rd = xyz(mul44(F.inv_mv, 0.0f, 0.0f, -1.0f, 0.0f));
    }
}

// ----------------------------------------------------------------------------------
// lights.h, color.h, envLight/envMapSample.h
// ----------------------------------------------------------------------------------
VT_DEV float luminance(f3 c) { return dot(c, mk3(0.2126f, 0.7152f, 0.0722f)); }      // color.h:1-8

// texture(backgroundTexture, uv).rgb: GL_LINEAR, clamp-to-edge (renderer.cpp:987-1002)
template <bool COUNT>
VT_DEV f3 env_lookup(const Frame& F, f2 uv, Tally<COUNT>& tl)
{
    const int w = F.env_w, h = F.env_h;
    const float x = uv.x * (float)w - 0.5f, y = uv.y * (float)h - 0.5f;
    const float x0f = floorf(x), y0f = floorf(y);
    const float a = x - x0f, b = y - y0f;
    int x0 = f2i(x0f), y0 = f2i(y0f);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = max(0, min(x0, w - 1)); x1 = max(0, min(x1, w - 1));
    y0 = max(0, min(y0, h - 1)); y1 = max(0, min(y1, h - 1));
    const float4 p00 = ldg_keep(F.env + ((size_t)x0 + (size_t)y0 * w));
    const float4 p10 = ldg_keep(F.env + ((size_t)x1 + (size_t)y0 * w));
    const float4 p01 = ldg_keep(F.env + ((size_t)x0 + (size_t)y1 * w));
    const float4 p11 = ldg_keep(F.env + ((size_t)x1 + (size_t)y1 * w));
    const f3 top = mk3(gmix(p00.x, p10.x, a), gmix(p00.y, p10.y, a), gmix(p00.z, p10.z, a));
    const f3 bot = mk3(gmix(p01.x, p11.x, a), gmix(p01.y, p11.y, a), gmix(p01.z, p11.z, a));
    VT_TALLY(Q, 1);
    return mk3(gmix(top.x, bot.x, b), gmix(top.y, bot.y, b), gmix(top.z, bot.z, b));
}
template <bool COUNT>
VT_DEV f3 background_color(const Frame& F, f3 v, Tally<COUNT>& tl)                   // lights.h:4-17
{
    if (F.use_image != 0) return env_lookup<COUNT>(F, uv_from_vector(v, F.env_rotation), tl);
    const float bias = gmax(0.0f, v.y);
    return F.bg_bottom * (1.0f - bias) + F.bg_top * bias;
}
VT_DEV f4 env_with_pdf(const Frame& F, f3 L)                                         // lights.h:27-32, L = getBackgroundColor(wi)
{
    const float pdf = (F.use_image != 0) ? luminance(L) / F.env_integral : 1.0f / (4.0f * VT_PI);
    return mk4(L, pdf);
}
template <bool COUNT>
VT_DEV f4 evaluate_env(const Frame& F, f3 wi, Tally<COUNT>& tl)                      // lights.h:20-33
{
    return env_with_pdf(F, background_color<COUNT>(F, wi, tl));
}
template <bool COUNT>
VT_DEV float cdf_u_at(const Frame& F, int x, int y, Tally<COUNT>& tl)
{
    VT_TALLY(E, 1);
    if ((unsigned)x >= (unsigned)F.cdf_u_w || (unsigned)y >= (unsigned)F.cdf_u_h) return 0.0f;
    return ldg_keep(F.cdf_u + ((size_t)x + (size_t)y * F.cdf_u_w));
}
template <bool COUNT>
VT_DEV float cdf_v_at(const Frame& F, int i, Tally<COUNT>& tl)
{
    VT_TALLY(E, 1);
    if ((unsigned)i >= (unsigned)F.cdf_v_n) return 0.0f;
    return ldg_keep(F.cdf_v + i);
}
// The searches of envMapSample.h:70-123 return R(s) = max({0} U {m in [1, n-2] : cdf[m] <= s}) whenever the CDF is sorted
// (the probe order is then irrelevant). The guide table stores R(j / K) for j = 0..K-1 and n-2 for j = K, so for
// s in [j/K, (j+1)/K) the answer lies in [guide[j], guide[j+1]] and the same bisection, started on that bracket, finds
// it after 1-3 probes instead of 8-9 dependent L2-latency loads. K is a power of two: s * K and the bucket are exact.
VT_DEV int cdf_search_guided(const float* __restrict__ cdf, const unsigned short* __restrict__ guide, int K, float s)
{
    const int j = min(K - 1, f2i(s * (float)K));
    int lo = (int)ldg_keep(guide + j), hi = (int)ldg_keep(guide + j + 1) + 1;
    while (lo != hi - 1) {
        const int m = (lo + hi) >> 1;
        if (s < ldg_keep(cdf + m)) hi = m; else lo = m;
    }
    return lo;
}
template <bool COUNT>
VT_DEV f2 sample_env_texture(const Frame& F, float su, float sv, Tally<COUNT>& tl)   // envMapSample.h:23-123
{
    const int sizeU = F.cdf_u_w, sizeV = F.cdf_v_n;
    int row, col;
    // counting builds run the literal searches so that E equals the oracle's load count
    if (!COUNT && F.guide_k > 0 && su >= 0.0f && su <= 1.0f && sv >= 0.0f && sv <= 1.0f) {
        row = cdf_search_guided(F.cdf_v, F.guide_v, F.guide_k, sv);
        col = cdf_search_guided(F.cdf_u + (size_t)row * F.cdf_u_w, F.guide_u + (size_t)row * (F.guide_k + 1), F.guide_k, su);
    } else {
        int lo = 0, hi = sizeV - 1;
        while (lo != hi - 1) {                                        // :70-94
            const int m = (lo + hi) / 2;
            if (sv < cdf_v_at<COUNT>(F, m, tl)) hi = m; else lo = m;
        }
        row = lo;
        lo = 0; hi = sizeU - 1;
        while (lo != hi - 1) {                                        // :98-123
            const int m = (lo + hi) / 2;
            if (su < cdf_u_at<COUNT>(F, m, row, tl)) hi = m; else lo = m;
        }
        col = lo;
    }
    float cl = cdf_u_at<COUNT>(F, col, row, tl), cu = cdf_u_at<COUNT>(F, col + 1, row, tl);   // :52-54
    const float du = (su - cl) / (cu - cl);
    cl = cdf_v_at<COUNT>(F, row, tl); cu = cdf_v_at<COUNT>(F, row + 1, tl);                    // :56-58
    const float dv = (sv - cl) / (cu - cl);
    return mk2(((float)col + du) / (float)(sizeU - 1), ((float)row + dv) / (float)(sizeV - 1)); // :61-62
}
template <bool COUNT>
VT_DEV f3 sample_env(const Frame& F, const Basis& b, float ux, float uy, f4& w_pdf, Tally<COUNT>& tl)   // lights.h:36-55
{
    if (F.use_image != 0) {
        const f2 uv = sample_env_texture<COUNT>(F, ux, uy, tl);
        const f3 w = direction_from_uv(uv, F.env_rotation);
        const f3 L = env_lookup<COUNT>(F, uv, tl);
        w_pdf = mk4(w, luminance(L) / F.env_integral);
        return L;
    }
    const f4 l = uniform_hemisphere(ux, uy);
    w_pdf = mk4(local_to_world(xyz(l), b), l.w);
    return background_color<COUNT>(F, xyz(w_pdf), tl);
}

// ----------------------------------------------------------------------------------
// bsdf/microfacet.h, bsdf/lambertian.h, materials/*.h
// ----------------------------------------------------------------------------------
#define VT_IOR 7.3f                                               // microfacet.h:2

VT_DEV float mf_G(f3 wo, f3 wi, f3 wh)                            // microfacet.h:5-15
{
    const float NdotWh = wh.y, NdotWo = wo.y, NdotWi = wi.y;
    const float WoDotWh = gabs(dot(wo, wh));
    return gmin(1.0f, gmin((2.0f * NdotWh * NdotWo / WoDotWh), (2.0f * NdotWh * NdotWi / WoDotWh)));
}
VT_DEV f4 mf_D(f3 refl, float e, f3 wo, f3 wh)                    // microfacet.h:19-37
{
    const float powCos = gpow(wh.y, e);
    const float f = (e + 2.0f) * VT_INV_TWOPI * powCos;
    const float woDotWh = dot(wo, wh);
    const float pdf = (woDotWh <= 0.0f) ? 0.0f : ((e + 1.0f) * powCos) / (VT_TWO_PI * 4.0f * woDotWh);
    return mk4(refl * f, pdf);
}
VT_DEV float mf_F(float cosH)                                     // microfacet.h:64-69
{
    const float sqrtR0 = (1.0f - VT_IOR) / (1.0f + VT_IOR);
    const float r0 = sqrtR0 * sqrtR0;
    return r0 + (1.0f - r0) * gpow(1.0f - cosH, 5.0f);
}
VT_DEV f4 eval_microfacet(f3 refl, float e, f3 wo, f3 wi)         // microfacet.h:73-91
{
    const float cosWi = wi.y, cosWo = wo.y;
    const f3 wh = normalize(wi + wo);
    const float cosH = dot(wi, wh);
    f4 fp = mf_D(refl, e, wo, wh);
    const float k = mf_G(wo, wi, wh) * mf_F(cosH) / (4.0f * cosWo * cosWi);
    fp.x *= k; fp.y *= k; fp.z *= k;
    return fp;
}
VT_DEV f3 sample_microfacet(f3 refl, float e, f3 wo, float ux, float uy, f4& f_pdf)   // microfacet.h:102-123 (+ sampleD :41-61)
{
    const float cosT = gpow(ux, 1.0f / (e + 1.0f));
    const float sinT = sqrtf(gmax(0.0f, 1.0f - cosT * cosT));
    const float phi = uy * 2.0f * VT_PI;
    const f3 wh0 = spherical(phi, cosT, sinT);
    const f3 wi = -wo + (2.0f * dot(wo, wh0)) * wh0;
    f4 fp = mf_D(refl, e, wo, wh0);
    const float cosWi = wi.y, cosWo = wo.y;
    const f3 wh = normalize(wi + wo);
    const float cosH = dot(wi, wh);
    const float g = mf_G(wo, wi, wh), f = mf_F(cosH), den = 4.0f * cosWo * cosWi;
    fp.x = refl.x * fp.x * g * f / den;                           // :116-120
    fp.y = refl.y * fp.y * g * f / den;
    fp.z = refl.z * fp.z * g * f / den;
    f_pdf = fp;
    return wi;
}

VT_DEV float fetch_mat(const Frame& F, int i)                     // texelFetch(materialDataTexture); out of range -> 0
{
    if ((unsigned)i >= (unsigned)F.n_materials) return 0.0f;
    return ldg_keep(F.materials + i);
}
VT_DEV f3 mat_vec(const Frame& F, int off) { return mk3(fetch_mat(F, off), fetch_mat(F, off + 1), fetch_mat(F, off + 2)); }

template <bool COUNT>
VT_DEV f4 evaluate_material(const Frame& F, int off, f3 wo, f3 wi, Tally<COUNT>& tl)  // materials.h:7-19
{
    const int type = f2i(fetch_mat(F, off));
    VT_TALLY(H, 1);
    off += 1;
    switch (type) {
    case 0: { const f3 a = mat_vec(F, off + 3); return mk4(mk3(gdiv_pi(a.x), gdiv_pi(a.y), gdiv_pi(a.z)), gdiv_pi(wi.y)); }               // matte.h:1-10, lambertian.h:23-29
    case 1:                                                                                            // metal.h:1-17
    case 2: {                                                                                          // plastic.h:1-17 (reflectance 1)
        const f3 refl = (type == 1) ? mat_vec(F, off + 3) : mk3(1.0f);
        return eval_microfacet(refl, fetch_mat(F, off + 6), wo, wi);
    }
    default: return mk4(0.f, 0.f, 0.f, 0.f);
    }
}
template <bool COUNT>
VT_DEV f3 sample_material(const Frame& F, int off, f3 wo, int2& rng, f4& f_pdf, Tally<COUNT>& tl)     // materials.h:30-43
{
    const int type = f2i(fetch_mat(F, off));
    VT_TALLY(H, 1);
    off += 1;
    if ((unsigned)type > 2u) { f_pdf = mk4(0.f, 0.f, 0.f, 0.f); return mk3(0.0f); }    // unknown type: no rand() is consumed
    const f4 u = rng_next<COUNT>(F, rng, tl);                     // one expansion for the three material types
    if (type == 0) {                                              // matte.h:12-22, lambertian.h:10-19
        const f3 a = mat_vec(F, off + 3);
        const f4 l = cosine_hemisphere(u.x, u.y);
        f_pdf = mk4(mk3(gdiv_pi(a.x), gdiv_pi(a.y), gdiv_pi(a.z)), l.w);
        return xyz(l);
    }
    const f3 refl = (type == 1) ? mat_vec(F, off + 3) : mk3(1.0f);   // metal.h:19-36; plastic.h:19-28 (reflectance 1)
    return sample_microfacet(refl, fetch_mat(F, off + 6), wo, u.x, u.y, f_pdf);
}
template <bool COUNT>
VT_DEV f3 emission_material(const Frame& F, int off, Tally<COUNT>& tl)                 // materials.h:46-56
{
    const int type = f2i(fetch_mat(F, off));
    VT_TALLY(H, 1);
    if (type == 0 || type == 1 || type == 2) return mat_vec(F, off + 1);
    return mk3(0.0f);
}

// ----------------------------------------------------------------------------------
// integrator pieces shared by pathTracer.fs and editMode.fs
// ----------------------------------------------------------------------------------

VT_DEV void voxel_index_to_pos(int idx, int X, int Y, int& x, int& y, int& z)          // coordinates.h:118-128
{
    const int dz = X * Y, dy = X;
    z = idx / dz; idx -= z * dz;
    y = idx / dy; idx -= y * dy;
    x = idx;
}

// pathTracer.fs:64-170 split at the shadow ray, so that the persistent kernel can trace it as a queued segment.
struct LightSample {
    f4 wl;          // direction to the light + pdf (already divided by the number of lights, :124)
    f3 L;           // radiance arriving from the sampled light
    int target;     // linear index of the sampled emissive voxel, -1 = environment
};
template <bool COUNT>
VT_DEV LightSample sample_light(const Volume& V, const Frame& F, const Basis& hb, int2& rng, Tally<COUNT>& tl)   // :69-124
{
    LightSample ls;
    ls.wl = mk4(0.f, 0.f, 0.f, 0.f);
    ls.target = -1;
    const f4 u = rng_next<COUNT>(F, rng, tl);                     // :77
    const int num_lights = F.n_emissive + 1;                      // :78
    const int light_index = f2i(u.x * (float)num_lights);         // :79
    if (light_index < num_lights - 1) {                           // :81
        // :85 texelFetch: an index outside the list (negative u.x: only from a caller-supplied noise table) reads 0
        const int eidx = ((unsigned)light_index < (unsigned)F.n_emissive) ? __ldg(F.emissive + light_index) : 0;
        int ex, ey, ez;
        voxel_index_to_pos(eidx, V.X, V.Y, ex, ey, ez);           // :86
        ls.target = ex + ey * V.X + ez * V.X * V.Y;
        const int eoff = fetch_offset(V, ex, ey, ez);             // :87
        ls.L = mk3(10.0f) * emission_material<COUNT>(F, eoff, tl);   // :88
        const f3 vse = mk3((float)ex, (float)ey, (float)ez);
        f3 ep = (vse / V.resf) * (V.bmax - V.bmin) + V.bmin;      // :90
        ep = ep + mk3(u.y, u.z, u.w) * V.vsize;                   // :92
        const f3 toL = ep - hb.position;                          // :94
        const float r = length(toL);
        const f3 w = toL / r;                                     // :96
        Basis lb;
        voxel_to_world(V, vse, hb.position, w, lb);               // :103-106
        const float area = 6.0f * V.vsize.x * V.vsize.y;          // :113
        const float jac = (r * r) / gabs(dot(-w, lb.normal));     // :114
        ls.wl = mk4(w, jac / area);                               // :115
    } else {
        ls.L = sample_env<COUNT>(F, hb, u.y, u.z, ls.wl, tl);     // :120
    }
    ls.wl.w = ls.wl.w / (float)num_lights;                        // :124
    return ls;
}
// :133-153: is the sampled light hidden, given the result of the shadow traversal
VT_DEV bool light_occluded(const Volume& V, int target, bool shadow_hit_something, f3 shadow_hit)
{
    if (target >= 0) {
        int ex, ey, ez;
        voxel_index_to_pos(target, V.X, V.Y, ex, ey, ez);
        return !shadow_hit_something || shadow_hit.x != (float)ex || shadow_hit.y != (float)ey || shadow_hit.z != (float)ez;
    }
    return shadow_hit_something;
}
// :155-164: contribution of a visible light sample
template <bool COUNT>
VT_DEV f3 light_contribution(const Frame& F, int mat_off, const Basis& hb, f3 wo, const LightSample& ls, Tally<COUNT>& tl)
{
    const f3 wdir = xyz(ls.wl);
    const f3 lsWo = world_to_local(wo, hb);                       // :159
    const f3 lsWi = world_to_local(wdir, hb);
    const f4 bf = evaluate_material<COUNT>(F, mat_off, lsWo, lsWi, tl);     // :161
    const float mis = power_heuristic(ls.wl.w, bf.w);             // :163
    return xyz(bf) * ls.L * gabs(dot(wdir, hb.normal)) * mis / ls.wl.w;   // :164
}

template <bool COUNT>
VT_DEV f3 direct_lighting(const Volume& V, const Frame& F, int mat_off, const Basis& hb, f3 wo, int2& rng, Tally<COUNT>& tl)   // pathTracer.fs:64-170
{
    const LightSample ls = sample_light<COUNT>(V, F, hb, rng, tl);
    f3 shadow_hit; bool hit_ground;
    const bool hit_something = traverse<COUNT>(V, hb.position, xyz(ls.wl), shadow_hit, hit_ground, tl);   // :133
    if (light_occluded(V, ls.target, hit_something, shadow_hit)) return mk3(0.0f);                       // :134-153
    return light_contribution<COUNT>(F, mat_off, hb, wo, ls, tl);
}

VT_DEV f3 tonemap(f3 rad)                                         // pathTracer.fs:294 (exposure 1, gamma 2.2: :43-44)
{
    const float inv_gamma = 1.0f / 2.2f;
    const f3 r = rad * 1.0f;
    return mk3(gpow(r.x, inv_gamma), gpow(r.y, inv_gamma), gpow(r.z, inv_gamma));
}

VT_DEV float wireframe_factor(const Volume& V, const Frame& F, const Basis& hb, f3 vs_hit)   // pathTracer.fs:260-270
{
    const f3 ctr = ((hb.position - V.bmin) / (V.bmax - V.bmin)) * V.resf;
    const f3 uvw = vs_hit - ctr;
    const f3 n = hb.normal;
    const float ux = gabs(dot(mk3(n.y, n.z, n.x), uvw));
    const float uy = gabs(dot(mk3(n.z, n.x, n.y), uvw));
    const float th = F.wire_thickness;
    const float w = gstep(th, ux) * gstep(ux, 1.0f - th) * gstep(th, uy) * gstep(uy, 1.0f - th);
    return (1.0f - F.wire_opacity) + F.wire_opacity * w;
}

VT_DEV int hit_code(const Volume& V, f3 p, bool ground)
{
    if (ground) return -2;
    return f2i(p.x) + f2i(p.y) * V.X + f2i(p.z) * V.X * V.Y;
}

// pathTracer.fs:172-296 for one fragment; returns the tone-mapped sample
template <bool COUNT>
VT_DEV f4 trace_pixel(const Volume& V, const Frame& F, int px, int py, int sample_count, int* primary, Tally<COUNT>& tl)
{
    const f3 frag = mk3((float)px + 0.5f, (float)py + 0.5f, 0.55f);   // gl_FragCoord (z: quad drawn at z = near = 0.1, identity MVP)
    int2 rng = rng_offset(px, py, sample_count, F.noise_w, F.noise_h);   // :174
    f3 radiance = mk3(0.0f), ro, rd, hit;
    generate_ray<COUNT>(F, frag, rng, ro, rd, tl);                // :179
    const float t = ray_aabb(ro, rd, V.bmin, V.bmax);             // :183
    if (primary) *primary = -1;
    if (t < 0.0f) return mk4(tonemap(background_color<COUNT>(F, rd, tl)), 1.0f);     // :187-194
    const f3 entry = ro + t * rd;                                 // :196
    f3 throughput = mk3(1.0f);
    bool hit_ground;
    if (!traverse<COUNT>(V, entry, rd, hit, hit_ground, tl))      // :202-208
        return mk4(tonemap(background_color<COUNT>(F, rd, tl)), 1.0f);
    if (primary) *primary = hit_code(V, hit, hit_ground);

    const int sel_x = F.shared->sel_index[0], sel_y = F.shared->sel_index[1], sel_z = F.shared->sel_index[2];
    int bounces = 0;
    while (bounces < F.max_bounces) {                             // :214
        Basis hb;
        voxel_to_world(V, hit, ro, rd, hb);                       // :221-223
        const int ix = f2i(hit.x), iy = f2i(hit.y), iz = f2i(hit.z);   // :225
        const int mat_off = fetch_offset(V, ix, iy, iz);          // :226
        if (ix == sel_x && iy == sel_y && iz == sel_z) { radiance = radiance + mk3(1.0f, 0.0f, 0.0f); break; }   // :228-233
        const f3 wo = -rd;                                        // :237
        const f3 lsWo = world_to_local(wo, hb);
        if (bounces == 0) radiance = radiance + throughput * emission_material<COUNT>(F, mat_off, tl);          // :241-245
        radiance = radiance + throughput * direct_lighting<COUNT>(V, F, mat_off, hb, wo, rng, tl);              // :248
        f4 bf;
        const f3 lsWi = sample_material<COUNT>(F, mat_off, lsWo, rng, bf, tl);                                  // :255
        if (F.wire_opacity > 0.0f) {                              // :260-270
            const float w = wireframe_factor(V, F, hb, hit);
            bf.x *= w; bf.y *= w; bf.z *= w;
        }
        const f3 wi = local_to_world(lsWi, hb);                   // :273
        throughput = throughput * ((xyz(bf) * gabs(dot(wi, hb.normal))) / bf.w);                                // :276
        ro = hb.position; rd = wi;                                // :278-279
        if (!traverse<COUNT>(V, ro, rd, hit, hit_ground, tl)) {   // :282-289
            const f4 Lp = evaluate_env<COUNT>(F, rd, tl);
            const float mis = power_heuristic(bf.w, Lp.w);
            radiance = radiance + (throughput * xyz(Lp)) * mis;
            break;
        }
        bounces++;
    }
    return mk4(tonemap(radiance), 1.0f);                          // :294-295
}

// integrator/editMode.fs:62-142
template <bool COUNT>
VT_DEV f4 preview_pixel(const Volume& V, const Frame& F, int px, int py, int sample_count, int* primary, Tally<COUNT>& tl)
{
    const f3 frag = mk3((float)px + 0.5f, (float)py + 0.5f, 0.55f);
    int2 rng = rng_offset(px, py, sample_count, F.noise_w, F.noise_h);
    f3 ro, rd, hit;
    generate_ray<COUNT>(F, frag, rng, ro, rd, tl);
    const float t = ray_aabb(ro, rd, V.bmin, V.bmax);
    if (primary) *primary = -1;
    if (t < 0.0f) return mk4(background_color<COUNT>(F, rd, tl), 1.0f);
    const f3 entry = ro + t * rd;
    bool hit_ground;
    if (!traverse<COUNT>(V, entry, rd, hit, hit_ground, tl)) return mk4(background_color<COUNT>(F, rd, tl), 1.0f);
    if (primary) *primary = hit_code(V, hit, hit_ground);
    Basis hb;
    voxel_to_world(V, hit, ro, rd, hb);
    if (f2i(hit.x) == F.shared->sel_index[0] && f2i(hit.y) == F.shared->sel_index[1] && f2i(hit.z) == F.shared->sel_index[2])
        return mk4(1.0f, 0.0f, 0.0f, 1.0f);                       // :107-112
    f3 albedo = mk3(1.0f);
    if (F.wire_opacity > 0.0f) albedo = albedo * wireframe_factor(V, F, hb, hit);
    float lighting = gmax(0.0f, dot(-rd, hb.normal));             // :130
    lighting = sqrtf(lighting);
    const f3 light_dir = mk3(1.0f, -1.0f, -1.0f);                 // :41 (never set by the host: SURVEY appendix B)
    f3 sh; bool g2;
    if (traverse<COUNT>(V, hb.position, -light_dir, sh, g2, tl)) lighting *= 0.5f;   // :134-137, ambientLight :42
    return mk4(albedo * lighting, 1.0f);
}

} // namespace vt
