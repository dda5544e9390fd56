/*
 * vto.c -- TEST INFRASTRUCTURE: CPU oracle for the voxelToy hot path.
 *
 * Plain-C restatement of the reference's GLSL device programs. Each function
 * cites the reference file:line it follows (paths relative to
 * /root/reference/src/shaders unless they start with renderer/ or voxelize/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this. The product never does.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp (oracle/Makefile).
 * Arithmetic contract: see vto_math.h and DESIGN.md.
 *
 * Parity pinning: see vto.h header comment.
 */
#include "vto.h"
#include "vto_math.h"

#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* image.cpp:309-321: reduction of the environment image before the CDFs are built. The reference calls OpenImageIO's resize
 * (unpinned, absent here: SURVEY 8c); the contract of this build is an area-weighted box filter: output pixel (x, y) averages the
 * source rectangle [x sx, (x+1) sx) x [y sy, (y+1) sy), edge pixels weighted by their overlap, double accumulator, rows outer. */
void vto_resize_box(const float* rgb, int w, int h, int nw, int nh, float* out)
{
    const double sx = (double)w / (double)nw, sy = (double)h / (double)nh;
    int x, y, c, i, j;
    for (y = 0; y < nh; y++) {
        const double y0 = (double)y * sy, y1 = (double)(y + 1) * sy;
        int j0 = (int)floor(y0), j1 = (int)ceil(y1);
        if (j1 > h) j1 = h;
        for (x = 0; x < nw; x++) {
            const double x0 = (double)x * sx, x1 = (double)(x + 1) * sx;
            int i0 = (int)floor(x0), i1 = (int)ceil(x1);
            if (i1 > w) i1 = w;
            for (c = 0; c < 3; c++) {
                double acc = 0.0;
                for (j = j0; j < j1; j++) {
                    const double wy = (y1 < (double)(j + 1) ? y1 : (double)(j + 1)) - (y0 > (double)j ? y0 : (double)j);
                    for (i = i0; i < i1; i++) {
                        const double wx = (x1 < (double)(i + 1) ? x1 : (double)(i + 1)) - (x0 > (double)i ? x0 : (double)i);
                        acc += (wy * wx) * (double)rgb[((size_t)j * w + i) * 3 + c];
                    }
                }
                out[((size_t)y * nw + x) * 3 + c] = (float)(acc / (sx * sy));
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* small helpers                                                              */
/* ------------------------------------------------------------------------- */

typedef struct { int x, y; } iv2;
typedef struct { v3 position, tangent, normal, binormal; } basis_t; /* coordinates.h:35-41 */

typedef struct {
    const vto_scene* s;
    v3 bmin, bmax, vsize;
    v3 resf;            /* vec3(voxelResolution) */
    int max_steps;      /* dda.h:98 */
    /* per-thread counters */
    uint64_t S, R, Hm, E, Q;
} ctx_t;

static inline v3 v3from(const float* p) { return V3(p[0], p[1], p[2]); }

/* texelFetch(materialOffsetTexture, p, 0).r ; out-of-range -> 0 (contract U2) */
static inline int32_t fetch_offset(const vto_scene* s, int x, int y, int z)
{
    if (x < 0 || y < 0 || z < 0 || x >= s->X || y >= s->Y || z >= s->Z) return 0;
    return s->grid[(size_t)x + (size_t)y * (size_t)s->X + (size_t)z * (size_t)s->X * (size_t)s->Y];
}
/* texelFetch(materialDataTexture, i, 0).r ; out-of-range -> 0 */
static inline float fetch_mat(const vto_scene* s, int i)
{
    if (i < 0 || i >= s->n_materials) return 0.0f;
    return s->materials[i];
}

/* ------------------------------------------------------------------------- */
/* random.h                                                                   */
/* ------------------------------------------------------------------------- */

/* random.h:3-11 -- int arithmetic wraps, >> is arithmetic on a signed int */
static inline int32_t hash_i(int32_t seed)
{
    uint32_t u;
    seed = (seed ^ 61) ^ (seed >> 16);
    u = (uint32_t)seed * 9u; seed = (int32_t)u;
    seed = seed ^ (seed >> 4);
    u = (uint32_t)seed * 0x27d4eb2du; seed = (int32_t)u;
    seed = seed ^ (seed >> 15);
    return seed;
}
uint32_t vto_hash(uint32_t seed) { return (uint32_t)hash_i((int32_t)seed); }

/* random.h:13-18 */
static inline iv2 rng_offset(int px, int py, int sequence, int rw, int rh)
{
    int32_t a = (int32_t)((uint32_t)px + (uint32_t)py * (uint32_t)rw);
    int32_t offset = hash_i(a) ^ hash_i(sequence);
    iv2 r; r.x = offset % rw; r.y = (offset / rw) % rh;
    return r;
}
void vto_rng_offset(int px, int py, int sequence, int rw, int rh, int out[2])
{
    iv2 r = rng_offset(px, py, sequence, rw, rh);
    out[0] = r.x; out[1] = r.y;
}

/* random.h:20-27 */
static inline v4 rng_next(ctx_t* c, iv2* off)
{
    const vto_scene* s = c->s;
    v4 r = { 0, 0, 0, 0 };
    if (off->x >= 0 && off->y >= 0 && off->x < s->noise_w && off->y < s->noise_h) {
        const float* p = s->noise + 4 * ((size_t)off->x + (size_t)off->y * (size_t)s->noise_w);
        r.x = p[0]; r.y = p[1]; r.z = p[2]; r.w = p[3];
    }
    off->x = (off->x + 1) % s->noise_w;
    if (off->x == 0) off->y = (off->y + 1) % s->noise_h;
    c->R++;
    return r;
}

/* renderer/renderer.cpp:741-744 : (float)rand()/RAND_MAX, glibc TYPE_3 additive
 * feedback generator with the default seed 1 (no srand anywhere in the reference).
 * Restated so that the table does not depend on the process-wide rand() state. */
void vto_noise_table(float* out, size_t n)
{
    /* glibc random_r.c, TYPE_3: degree 31, separation 3 */
    int32_t r[34];
    size_t i, k = 0;
    uint32_t* st;
    r[0] = 1;
    for (i = 1; i < 31; i++) {
        int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        int64_t w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    st = (uint32_t*)malloc(sizeof(uint32_t) * (n + 344 + 34));
    for (i = 0; i < 31; i++) st[i] = (uint32_t)r[i];
    for (i = 31; i < 34; i++) st[i] = st[i - 31];
    for (i = 34; i < 344; i++) st[i] = st[i - 31] + st[i - 3];
    for (i = 344; i < 344 + n; i++) {
        st[i] = st[i - 31] + st[i - 3];
        out[k++] = (float)(int32_t)(st[i] >> 1) / 2147483647.0f; /* (float)rand() / RAND_MAX: int->float, RAND_MAX->float */
    }
    free(st);
}

/* ------------------------------------------------------------------------- */
/* aabb.h:1-32                                                                */
/* ------------------------------------------------------------------------- */
static inline float ray_aabb(v3 o, v3 d, v3 bmin, v3 bmax)
{
    v3 inv = v3div(v3s(1.0f), d);
    v3 t0 = v3mul(v3sub(bmin, o), inv);
    v3 t1 = v3mul(v3sub(bmax, o), inv);
    v3 tmin = V3(g_min(t0.x, t1.x), g_min(t0.y, t1.y), g_min(t0.z, t1.z));
    float tminf = g_max(tmin.x, g_max(tmin.y, tmin.z));
    v3 tmax = V3(g_max(t0.x, t1.x), g_max(t0.y, t1.y), g_max(t0.z, t1.z));
    float tmaxf = g_min(tmax.x, g_min(tmax.y, tmax.z));
    if (tmaxf < 0.0f) return -1.0f;
    if (tminf > tmaxf) return -1.0f;
    return g_max(0.0f, tminf);
}

/* ------------------------------------------------------------------------- */
/* dda.h                                                                      */
/* ------------------------------------------------------------------------- */

int vto_dda_step_cap(int X, int Y, int Z)
{
    /* dda.h:98  int(2 * ceil(length(vec3(voxelResolution)))) */
    v3 r = V3((float)X, (float)Y, (float)Z);
    return g_f2i(2.0f * ceilf(v3length(r)));
}

static inline int out_of_grid(v3 p, v3 resf)
{
    /* any(lessThan(p, 0)) || any(greaterThanEqual(p, res)) ; NaN compares false */
    return (p.x < 0.0f) || (p.y < 0.0f) || (p.z < 0.0f) ||
           (p.x >= resf.x) || (p.y >= resf.y) || (p.z >= resf.z);
}

/* dda.h:7-61. *hit_pos is zero-initialised on entry (contract U1). */
static int raymarch(ctx_t* c, v3 o, v3 d, v3* hit_pos)
{
    const vto_scene* s = c->s;
    int isect = 0, steps = 0;
    v3 ext, vo, vp, inc, sg, dis, mask;
    *hit_pos = V3(0.0f, 0.0f, 0.0f);

    ext = v3div(v3s(1.0f), v3sub(c->bmax, c->bmin));                 /* :16 */
    o = v3add(o, v3scale(v3sign(d), 0.001f));                        /* :19 */
    vo = v3mul(v3mul(v3sub(o, c->bmin), ext), c->resf);              /* :20 */
    vp = v3floor(vo);                                                /* :22 */
    if (out_of_grid(vp, c->resf)) return 0;                          /* :24-25 */

    {   /* :29 mix(d, 1e-5, step(abs(d), 1e-5)) */
        v3 t = V3(g_step(g_abs(d.x), 1e-5f), g_step(g_abs(d.y), 1e-5f), g_step(g_abs(d.z), 1e-5f));
        d = V3(g_mix(d.x, 1e-5f, t.x), g_mix(d.y, 1e-5f, t.y), g_mix(d.z, 1e-5f, t.z));
    }
    inc = v3div(v3s(1.0f), d);                                       /* :31 */
    sg = v3sign(d);                                                  /* :32 */
    /* :34 (voxelPos-voxelOrigin + 0.5 + sign*0.5) * inc */
    dis = v3mul(v3add(v3add(v3sub(vp, vo), v3s(0.5f)), v3scale(sg, 0.5f)), inc);

    while (steps < c->max_steps) {                                   /* :38 */
        if (out_of_grid(vp, c->resf)) break;                         /* :41-42 */
        c->S++;
        if (fetch_offset(s, g_f2i(vp.x), g_f2i(vp.y), g_f2i(vp.z)) >= 0) { isect = 1; break; } /* :44-50 */
        /* :51 mask = step(dis.xyz, dis.yxy) * step(dis.xyz, dis.zzx) */
        mask = V3(g_step(dis.x, dis.y) * g_step(dis.x, dis.z),
                  g_step(dis.y, dis.x) * g_step(dis.y, dis.z),
                  g_step(dis.z, dis.y) * g_step(dis.z, dis.x));
        dis = v3add(dis, v3mul(v3mul(mask, sg), inc));               /* :52 */
        vp = v3add(vp, v3mul(mask, sg));                             /* :53 */
        steps++;
    }
    *hit_pos = vp;                                                   /* :59 */
    return isect;
}

/* dda.h:63-100 */
static int traverse(ctx_t* c, v3 o, v3 d, v3* hit_pos, int* hit_ground)
{
    if (raymarch(c, o, d, hit_pos)) { *hit_ground = 0; return 1; }
    *hit_ground = (hit_pos->y < 0.0f);
    return *hit_ground;
}

/* ------------------------------------------------------------------------- */
/* coordinates.h                                                              */
/* ------------------------------------------------------------------------- */

/* coordinates.h:12-27 */
static v4 screen_to_eye_persp(const vto_scene* s, v3 ws)
{
    const float vx = 0.0f, vy = 0.0f, vz = (float)s->W, vw = (float)s->H; /* viewport = (0,0,W,H) */
    v3 ndc; v4 clip;
    ndc.x = ((2.0f * ws.x) - (2.0f * vx)) / vz - 1.0f;
    ndc.y = ((2.0f * ws.y) - (2.0f * vy)) / vw - 1.0f;
    ndc.z = (2.0f * ws.z - 0.0f - 1.0f) / (1.0f - 0.0f);             /* gl_DepthRange = (0,1) */
    /* GLSL cameraProj[3][2] = host pm.x[2][3]; [2][2] = pm.x[2][2]; [2][3] = pm.x[3][2] */
    clip.w = s->proj[2 * 4 + 3] / (ndc.z - (s->proj[2 * 4 + 2] / s->proj[3 * 4 + 2]));
    clip.x = ndc.x * clip.w; clip.y = ndc.y * clip.w; clip.z = ndc.z * clip.w;
    return m4mulv(s->inv_proj, clip.x, clip.y, clip.z, clip.w);
}

/* coordinates.h:1-10 */
static v4 screen_to_eye_ortho(const vto_scene* s, v3 ws)
{
    v3 ndc;
    ndc.x = (ws.x / (float)s->W) * 2.0f - 1.0f;
    ndc.y = (ws.y / (float)s->H) * 2.0f - 1.0f;
    ndc.z = (2.0f * ws.z - 0.0f - 1.0f) / (1.0f - 0.0f);
    return m4mulv(s->inv_proj, ndc.x, ndc.y, ndc.z, 1.0f);
}

/* coordinates.h:43-57 */
static inline v3 local_to_world(v3 v, const basis_t* b)
{
    return v3add(v3add(v3scale(b->tangent, v.x), v3scale(b->normal, v.y)), v3scale(b->binormal, v.z));
}
static inline v3 world_to_local(v3 v, const basis_t* b)
{
    return V3(v3dot(v, b->tangent), v3dot(v, b->normal), v3dot(v, b->binormal));
}

/* coordinates.h:59-82 */
static void voxel_to_world(const ctx_t* c, v3 vsP, v3 o, v3 d, basis_t* b)
{
    v3 vmin = v3add(v3mul(vsP, c->vsize), c->bmin);
    v3 vmax = v3add(vmin, c->vsize);
    float t = ray_aabb(o, d, vmin, vmax);
    v3 center, h, a, mask, tan0;
    b->position = v3add(o, v3scale(d, t));
    center = v3add(vmin, v3scale(c->vsize, 0.5f));
    h = v3sub(b->position, center);
    a = v3abs(h);
    mask = V3(g_step(a.y, a.x) * g_step(a.z, a.x),
              g_step(a.x, a.y) * g_step(a.z, a.y),
              g_step(a.x, a.z) * g_step(a.y, a.z));
    b->normal = v3mul(mask, v3sign(h));
    tan0 = v3cross(b->normal, v3s(0.57735026919f));
    b->binormal = v3normalize(v3cross(tan0, b->normal));
    b->tangent = v3normalize(v3cross(b->binormal, b->normal));
}

/* coordinates.h:91-95 */
static inline v3 spherical3(float phi, float cosT, float sinT)
{
    return V3(sinT * g_cos(phi), cosT, sinT * g_sin(phi));
}
/* coordinates.h:85-89 */
static inline v3 spherical2(float phi, float theta)
{
    float sinT = g_sin(theta);
    return V3(sinT * g_cos(phi), g_cos(theta), sinT * g_sin(phi));
}
/* coordinates.h:97-109 */
static inline void uv_from_vector(v3 v, float rot, float* u, float* vv)
{
    float theta = g_acos(v.y);
    float phi = g_atan2(v.z, v.x) + VTO_PI + rot;
    *u = g_mod(phi / VTO_TWO_PI, 1.0f);
    *vv = theta / VTO_PI;
}
/* coordinates.h:111-116 */
static inline v3 direction_from_uv(float u, float v, float rot)
{
    float phi = u * VTO_TWO_PI - VTO_PI - rot;
    float theta = v * VTO_PI;
    return spherical2(phi, theta);
}
/* coordinates.h:118-128 */
static inline void voxel_index_to_pos(int idx, int X, int Y, int* x, int* y, int* z)
{
    int dz = X * Y, dy = X;
    *z = idx / dz; idx -= *z * dz;
    *y = idx / dy; idx -= *y * dy;
    *x = idx;
}

/* ------------------------------------------------------------------------- */
/* sampling.h                                                                 */
/* ------------------------------------------------------------------------- */
static inline void sample_disk(float ux, float uy, float* px, float* py)   /* sampling.h:2-7 */
{
    float r = sqrtf(ux);
    float theta = 2.0f * VTO_PI * uy;
    *px = r * g_cos(theta); *py = r * g_sin(theta);
}
static inline v4 cosine_hemisphere(float ux, float uy)                      /* sampling.h:13-22 */
{
    float px, py, y; v4 r;
    sample_disk(ux, uy, &px, &py);
    y = sqrtf(g_max(0.0f, 1.0f - px * px - py * py));
    r.x = px; r.y = y; r.z = py; r.w = y / VTO_PI;
    return r;
}
static inline v4 uniform_hemisphere(float ux, float uy)                     /* sampling.h:27-36 */
{
    float y = ux;
    float r = sqrtf(g_max(0.0f, 1.0f - y * y));
    float phi = uy * 2.0f * VTO_PI;
    v4 o; o.x = r * g_cos(phi); o.y = y; o.z = r * g_sin(phi); o.w = 1.0f / (2.0f * VTO_PI);
    return o;
}
static inline float power_heuristic(float f, float g) { return f * f / (f * f + g * g); } /* sampling.h:40-43 */

/* ------------------------------------------------------------------------- */
/* generateRay.h                                                              */
/* ------------------------------------------------------------------------- */
static void generate_ray(ctx_t* c, v3 frag, iv2* rng, v3* ro, v3* rd)
{
    const vto_scene* s = c->s;
    const float* im = s->inv_modelview;
    if (s->lens_model == 0) {
        /* generateRay.h:10-26 */
        v4 u = rng_next(c, rng);
        v3 jitter = V3(u.x - 0.5f, u.y - 0.5f, 0.0f);
        v4 es = screen_to_eye_persp(s, v3add(frag, jitter));
        v4 o = m4mulv(im, 0.0f, 0.0f, 0.0f, 1.0f);
        v3 n = v3normalize(V3(es.x, es.y, es.z));
        v4 d = m4mulv(im, n.x, n.y, n.z, 0.0f);
        float len = sqrtf(((d.x * d.x + d.y * d.y) + d.z * d.z) + d.w * d.w); /* normalize(vec4) */
        *ro = V3(o.x, o.y, o.z);
        *rd = V3(d.x / len, d.y / len, d.z / len);
    } else if (s->lens_model == 1) {
        /* generateRay.h:30-101 */
        float fd = s->focal_distance;
        v4 u = rng_next(c, rng);
        float dx, dy, t;
        v4 es4, fp, o;
        v3 es, esd, focal;
        sample_disk(u.x, u.y, &dx, &dy);
        es4 = screen_to_eye_persp(s, frag);
        es = V3(es4.x, es4.y, es4.z);
        esd = v3normalize(es);
        t = fd / -(esd.z);
        focal = v3add(es, v3scale(esd, t));
        fp = m4mulv(im, focal.x, focal.y, focal.z, 1.0f);
        o = m4mulv(im, dx * s->lens_radius, dy * s->lens_radius, 0.0f, 1.0f);
        *ro = V3(o.x, o.y, o.z);
        *rd = v3normalize(v3sub(V3(fp.x, fp.y, fp.z), *ro));
    } else {
        /* generateRay.h:1-7 */
        v4 e = screen_to_eye_ortho(s, frag);
        v4 o = m4mulv(im, e.x, e.y, e.z, e.w);
        v4 d = m4mulv(im, 0.0f, 0.0f, -1.0f, 0.0f);
        *ro = V3(o.x, o.y, o.z);
        *rd = V3(d.x, d.y, d.z);
    }
}

/* ------------------------------------------------------------------------- */
/* lights.h, color.h, envLight/envMapSample.h                                 */
/* ------------------------------------------------------------------------- */
static inline float luminance(v3 c) { return v3dot(c, V3(0.2126f, 0.7152f, 0.0722f)); } /* color.h:1-8 */

/* texture(backgroundTexture, uv).rgb : GL_LINEAR, clamp-to-edge, texel centres at (i+0.5)/size.
 * renderer/renderer.cpp:987-1002. Filter arithmetic fixed by the contract. */
static v3 env_lookup(ctx_t* c, float u, float v)
{
    const vto_scene* s = c->s;
    int w = s->env_w, h = s->env_h;
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float a = x - x0f, b = y - y0f;
    int x0 = g_f2i(x0f), y0 = g_f2i(y0f), x1, y1;
    const float *p00, *p10, *p01, *p11;
    v3 top, bot;
    x1 = x0 + 1; y1 = y0 + 1;
    if (x0 < 0) x0 = 0; if (x0 > w - 1) x0 = w - 1;
    if (x1 < 0) x1 = 0; if (x1 > w - 1) x1 = w - 1;
    if (y0 < 0) y0 = 0; if (y0 > h - 1) y0 = h - 1;
    if (y1 < 0) y1 = 0; if (y1 > h - 1) y1 = h - 1;
    p00 = s->env_rgb + 3 * ((size_t)x0 + (size_t)y0 * w);
    p10 = s->env_rgb + 3 * ((size_t)x1 + (size_t)y0 * w);
    p01 = s->env_rgb + 3 * ((size_t)x0 + (size_t)y1 * w);
    p11 = s->env_rgb + 3 * ((size_t)x1 + (size_t)y1 * w);
    top = V3(g_mix(p00[0], p10[0], a), g_mix(p00[1], p10[1], a), g_mix(p00[2], p10[2], a));
    bot = V3(g_mix(p01[0], p11[0], a), g_mix(p01[1], p11[1], a), g_mix(p01[2], p11[2], a));
    c->Q++;
    return V3(g_mix(top.x, bot.x, b), g_mix(top.y, bot.y, b), g_mix(top.z, bot.z, b));
}

/* lights.h:4-17 */
static v3 background_color(ctx_t* c, v3 v)
{
    const vto_scene* s = c->s;
    if (s->use_image != 0) {
        float uu, vv;
        uv_from_vector(v, s->env_rotation, &uu, &vv);
        return env_lookup(c, uu, vv);
    } else {
        float bias = g_max(0.0f, v.y);
        return v3add(v3scale(v3from(s->bg_bottom), 1.0f - bias), v3scale(v3from(s->bg_top), bias));
    }
}

/* lights.h:20-33 */
static v4 evaluate_env(ctx_t* c, v3 wi)
{
    const vto_scene* s = c->s;
    v4 r; v3 L;
    if (s->use_image != 0) {
        L = background_color(c, wi);
        r.w = luminance(L) / s->env_integral;
    } else {
        r.w = 1.0f / (4.0f * VTO_PI);
        L = background_color(c, wi);
    }
    r.x = L.x; r.y = L.y; r.z = L.z;
    return r;
}

static inline float cdf_u_at(ctx_t* c, int x, int y)
{
    const vto_scene* s = c->s;
    c->E++;
    if (x < 0 || y < 0 || x >= s->cdf_u_w || y >= s->cdf_u_h) return 0.0f;
    return s->cdf_u[(size_t)x + (size_t)y * s->cdf_u_w];
}
static inline float cdf_v_at(ctx_t* c, int i)
{
    const vto_scene* s = c->s;
    c->E++;
    if (i < 0 || i >= s->cdf_v_n) return 0.0f;
    return s->cdf_v[i];
}

/* envMapSample.h:23-123 */
static void sample_env_texture(ctx_t* c, float su, float sv, float* ou, float* ov)
{
    const vto_scene* s = c->s;
    int sizeU = s->cdf_u_w, sizeV = s->cdf_v_n;
    int lo, hi, row, col;
    float cl, cu, du, dv;
    lo = 0; hi = sizeV - 1;                                          /* :70-94 */
    while (lo != hi - 1) {
        int m = (lo + hi) / 2;
        float cdf = cdf_v_at(c, m);
        if (sv < cdf) hi = m; else lo = m;
    }
    row = lo;
    lo = 0; hi = sizeU - 1;                                          /* :98-123 */
    while (lo != hi - 1) {
        int m = (lo + hi) / 2;
        float cdf = cdf_u_at(c, m, row);
        if (su < cdf) hi = m; else lo = m;
    }
    col = lo;
    cl = cdf_u_at(c, col, row); cu = cdf_u_at(c, col + 1, row);      /* :52-54 */
    du = (su - cl) / (cu - cl);
    cl = cdf_v_at(c, row); cu = cdf_v_at(c, row + 1);                /* :56-58 */
    dv = (sv - cl) / (cu - cl);
    *ou = ((float)col + du) / (float)(sizeU - 1);                    /* :61-62 */
    *ov = ((float)row + dv) / (float)(sizeV - 1);
}

/* lights.h:36-55 ; returns radiance, writes wsW_pdf */
static v3 sample_env(ctx_t* c, const basis_t* b, float ux, float uy, v4* w_pdf)
{
    const vto_scene* s = c->s;
    if (s->use_image != 0) {
        float u, v; v3 w, L;
        sample_env_texture(c, ux, uy, &u, &v);
        w = direction_from_uv(u, v, s->env_rotation);
        L = env_lookup(c, u, v);
        w_pdf->x = w.x; w_pdf->y = w.y; w_pdf->z = w.z;
        w_pdf->w = luminance(L) / s->env_integral;
        return L;
    } else {
        v4 l = uniform_hemisphere(ux, uy);
        v3 w = local_to_world(V3(l.x, l.y, l.z), b);
        w_pdf->x = w.x; w_pdf->y = w.y; w_pdf->z = w.z; w_pdf->w = l.w;
        return background_color(c, w);
    }
}

/* ------------------------------------------------------------------------- */
/* bsdf/lambertian.h, bsdf/microfacet.h                                       */
/* ------------------------------------------------------------------------- */
#define IOR 7.3f   /* microfacet.h:2 */

static inline float mf_G(v3 wo, v3 wi, v3 wh)                               /* microfacet.h:5-15 */
{
    float NdotWh = wh.y, NdotWo = wo.y, NdotWi = wi.y;
    float WoDotWh = g_abs(v3dot(wo, wh));
    return g_min(1.0f, g_min((2.0f * NdotWh * NdotWo / WoDotWh), (2.0f * NdotWh * NdotWi / WoDotWh)));
}
static inline v4 mf_D(v3 refl, float e, v3 wo, v3 wh)                        /* microfacet.h:19-37 */
{
    float powCos = g_pow(wh.y, e);
    float f = (e + 2.0f) * VTO_INV_TWOPI * powCos;
    float woDotWh = v3dot(wo, wh);
    float pdf = (woDotWh <= 0.0f) ? 0.0f : ((e + 1.0f) * powCos) / (VTO_TWO_PI * 4.0f * woDotWh);
    v4 r; r.x = refl.x * f; r.y = refl.y * f; r.z = refl.z * f; r.w = pdf;
    return r;
}
static inline float mf_F(float cosH)                                         /* microfacet.h:64-69 */
{
    float sqrtR0 = (1.0f - IOR) / (1.0f + IOR);
    float r0 = sqrtR0 * sqrtR0;
    return r0 + (1.0f - r0) * g_pow(1.0f - cosH, 5.0f);
}
static v4 eval_microfacet(v3 refl, float e, v3 wo, v3 wi)                     /* microfacet.h:73-91 */
{
    float cosWi = wi.y, cosWo = wo.y;
    v3 wh = v3normalize(v3add(wi, wo));
    float cosH = v3dot(wi, wh);
    v4 fp = mf_D(refl, e, wo, wh);
    float k = mf_G(wo, wi, wh) * mf_F(cosH) / (4.0f * cosWo * cosWi);
    fp.x *= k; fp.y *= k; fp.z *= k;
    return fp;
}
static v3 sample_microfacet(ctx_t* c, v3 refl, float e, v3 wo, iv2* rng, v4* f_pdf) /* microfacet.h:102-123 */
{
    v4 u = rng_next(c, rng);
    /* sampleD, microfacet.h:41-61 */
    float cosT = g_pow(u.x, 1.0f / (e + 1.0f));
    float sinT = sqrtf(g_max(0.0f, 1.0f - cosT * cosT));
    float phi = u.y * 2.0f * VTO_PI;
    v3 wh0 = spherical3(phi, cosT, sinT);
    v3 wi = v3add(v3neg(wo), v3scale(wh0, 2.0f * v3dot(wo, wh0)));
    v4 fp = mf_D(refl, e, wo, wh0);
    float cosWi = wi.y, cosWo = wo.y;
    v3 wh = v3normalize(v3add(wi, wo));
    float cosH = v3dot(wi, wh);
    float g = mf_G(wo, wi, wh), f = mf_F(cosH), den = 4.0f * cosWo * cosWi;
    /* :116-120 reflectance * f_pdf.xyz * G * F / den */
    fp.x = refl.x * fp.x * g * f / den;
    fp.y = refl.y * fp.y * g * f / den;
    fp.z = refl.z * fp.z * g * f / den;
    *f_pdf = fp;
    return wi;
}

/* ------------------------------------------------------------------------- */
/* materials/ (.h)                                                            */
/* ------------------------------------------------------------------------- */
static inline v3 mat_vec(const vto_scene* s, int off) { return V3(fetch_mat(s, off), fetch_mat(s, off + 1), fetch_mat(s, off + 2)); }

/* materials.h:7-19 */
static v4 evaluate_material(ctx_t* c, int off, v3 wo, v3 wi)
{
    const vto_scene* s = c->s;
    int type = g_f2i(fetch_mat(s, off));
    v4 r = { 0, 0, 0, 0 };
    c->Hm++;
    off += 1;
    switch (type) {
    case 0: { v3 a = mat_vec(s, off + 3);                                     /* matte.h:1-10, lambertian.h:23-29 */
              r.x = a.x / VTO_PI; r.y = a.y / VTO_PI; r.z = a.z / VTO_PI; r.w = wi.y / VTO_PI; return r; }
    case 1: return eval_microfacet(mat_vec(s, off + 3), fetch_mat(s, off + 6), wo, wi);   /* metal.h:1-17 */
    case 2: return eval_microfacet(v3s(1.0f), fetch_mat(s, off + 6), wo, wi);             /* plastic.h:1-17 */
    default: return r;
    }
}
/* materials.h:30-43 */
static v3 sample_material(ctx_t* c, int off, v3 wo, iv2* rng, v4* f_pdf)
{
    const vto_scene* s = c->s;
    int type = g_f2i(fetch_mat(s, off));
    c->Hm++;
    off += 1;
    switch (type) {
    case 0: {                                                                  /* matte.h:12-22, lambertian.h:10-19 */
        v3 a = mat_vec(s, off + 3);
        v4 u = rng_next(c, rng);
        v4 l = cosine_hemisphere(u.x, u.y);
        f_pdf->x = a.x / VTO_PI; f_pdf->y = a.y / VTO_PI; f_pdf->z = a.z / VTO_PI; f_pdf->w = l.w;
        return V3(l.x, l.y, l.z);
    }
    case 1: return sample_microfacet(c, mat_vec(s, off + 3), fetch_mat(s, off + 6), wo, rng, f_pdf); /* metal.h:19-36 */
    case 2: return sample_microfacet(c, v3s(1.0f), fetch_mat(s, off + 6), wo, rng, f_pdf);           /* plastic.h:19-28 */
    default:
        /* materials.h:41 returns vec3(0) and leaves f_pdf unwritten: contract = zeros (out params zero-initialised) */
        f_pdf->x = f_pdf->y = f_pdf->z = f_pdf->w = 0.0f;
        return v3s(0.0f);
    }
}
/* materials.h:46-56 */
static v3 emission_material(ctx_t* c, int off)
{
    const vto_scene* s = c->s;
    int type = g_f2i(fetch_mat(s, off));
    c->Hm++;
    if (type == 0 || type == 1 || type == 2) return mat_vec(s, off + 1);
    return v3s(0.0f);
}

/* ------------------------------------------------------------------------- */
/* integrator/pathTracer.fs                                                   */
/* ------------------------------------------------------------------------- */

/* pathTracer.fs:64-170 */
static v3 direct_lighting(ctx_t* c, int mat_off, const basis_t* hb, v3 wo, iv2* rng)
{
    const vto_scene* s = c->s;
    v4 wl = { 0, 0, 0, 0 };
    v3 L, shadow_hit, lsWo, lsWi, wdir, res;
    int ex = 0, ey = 0, ez = 0, hit_ground, missed;
    v4 u = rng_next(c, rng);                                                   /* :77 */
    int num_lights = s->n_emissive + 1;                                        /* :78 */
    int light_index = g_f2i(u.x * (float)num_lights);                          /* :79 */
    int sampling_voxel = light_index < num_lights - 1;                         /* :81 */
    v4 bf; float mis, adot;

    if (sampling_voxel) {
        int eidx = ((unsigned)light_index < (unsigned)s->n_emissive) ? s->emissive[light_index] : 0;   /* :85 texelFetch, out of range -> 0 */
        int eoff; v3 ep, toL, vse; float r, area, jac; basis_t lb;
        voxel_index_to_pos(eidx, s->X, s->Y, &ex, &ey, &ez);                   /* :86 */
        eoff = fetch_offset(s, ex, ey, ez);                                    /* :87 */
        L = v3mul(v3s(10.0f), emission_material(c, eoff));                     /* :88 */
        vse = V3((float)ex, (float)ey, (float)ez);
        ep = v3add(v3mul(v3div(vse, c->resf), v3sub(c->bmax, c->bmin)), c->bmin);   /* :90 */
        ep = v3add(ep, v3mul(V3(u.y, u.z, u.w), c->vsize));                    /* :92 */
        toL = v3sub(ep, hb->position);                                         /* :94 */
        r = v3length(toL);
        wl.x = toL.x / r; wl.y = toL.y / r; wl.z = toL.z / r;                  /* :96 */
        voxel_to_world(c, vse, hb->position, V3(wl.x, wl.y, wl.z), &lb);       /* :103-106 */
        area = 6.0f * c->vsize.x * c->vsize.y;                                 /* :113 */
        jac = (r * r) / g_abs(v3dot(V3(-wl.x, -wl.y, -wl.z), lb.normal));      /* :114 */
        wl.w = jac / area;                                                     /* :115 */
    } else {
        L = sample_env(c, hb, u.y, u.z, &wl);                                  /* :120 */
    }
    wl.w = wl.w / (float)num_lights;                                           /* :124 */
    wdir = V3(wl.x, wl.y, wl.z);

    missed = !traverse(c, hb->position, wdir, &shadow_hit, &hit_ground);       /* :133 */
    if (sampling_voxel) {
        if (missed || (shadow_hit.x != (float)ex) || (shadow_hit.y != (float)ey) || (shadow_hit.z != (float)ez))
            return v3s(0.0f);                                                  /* :138-142 */
    } else {
        if (!missed) return v3s(0.0f);                                         /* :148-152 */
    }
    lsWo = world_to_local(wo, hb);                                             /* :159 */
    lsWi = world_to_local(wdir, hb);
    bf = evaluate_material(c, mat_off, lsWo, lsWi);                            /* :161 */
    mis = power_heuristic(wl.w, bf.w);                                         /* :163 */
    adot = g_abs(v3dot(wdir, hb->normal));
    /* :164  f * L * abs(dot) * mis / pdf */
    res = v3mul(V3(bf.x, bf.y, bf.z), L);
    res = v3scale(res, adot);
    res = v3scale(res, mis);
    res = v3divs(res, wl.w);
    return res;
}

static inline v3 tonemap(v3 rad)
{
    /* pathTracer.fs:294  pow(radiance * exposure, vec3(1.0 / gamma)), exposure 1, gamma 2.2 */
    const float inv_gamma = 1.0f / 2.2f;
    v3 r = v3scale(rad, 1.0f);
    return V3(g_pow(r.x, inv_gamma), g_pow(r.y, inv_gamma), g_pow(r.z, inv_gamma));
}

/* wireframe factor, pathTracer.fs:260-270 == editMode.fs:116-126 */
static float wireframe_factor(const ctx_t* c, const basis_t* hb, v3 vs_hit)
{
    const vto_scene* s = c->s;
    v3 ctr = v3mul(v3div(v3sub(hb->position, c->bmin), v3sub(c->bmax, c->bmin)), c->resf);
    v3 uvw = v3sub(vs_hit, ctr);
    v3 n = hb->normal;
    float ux = g_abs(v3dot(V3(n.y, n.z, n.x), uvw));
    float uy = g_abs(v3dot(V3(n.z, n.x, n.y), uvw));
    float th = s->wire_thickness;
    float w = g_step(th, ux) * g_step(ux, 1.0f - th) * g_step(th, uy) * g_step(uy, 1.0f - th);
    return (1.0f - s->wire_opacity) + s->wire_opacity * w;
}

static void ctx_init(ctx_t* c, const vto_scene* s)
{
    c->s = s;
    c->bmin = v3from(s->bmin); c->bmax = v3from(s->bmax); c->vsize = v3from(s->voxel_size);
    c->resf = V3((float)s->X, (float)s->Y, (float)s->Z);
    c->max_steps = vto_dda_step_cap(s->X, s->Y, s->Z);
    c->S = c->R = c->Hm = c->E = c->Q = 0;
}

static inline int32_t hit_code(const vto_scene* s, v3 p, int ground)
{
    if (ground) return -2;
    return (int32_t)(g_f2i(p.x) + g_f2i(p.y) * s->X + g_f2i(p.z) * s->X * s->Y);
}

/* pathTracer.fs:172-296 for one fragment */
static v4 trace_pixel(ctx_t* c, int px, int py, int sample_count, int32_t* primary)
{
    const vto_scene* s = c->s;
    v3 frag = V3((float)px + 0.5f, (float)py + 0.5f, 0.55f);  /* gl_FragCoord; z: quad at z=near=0.1, identity MVP */
    iv2 rng = rng_offset(px, py, sample_count, s->noise_w, s->noise_h);      /* :174 */
    v3 radiance = v3s(0.0f), ro, rd, entry, hit, throughput;
    float t; int hit_ground, bounces = 0;
    v4 out;

    generate_ray(c, frag, &rng, &ro, &rd);                                    /* :179 */
    t = ray_aabb(ro, rd, c->bmin, c->bmax);                                   /* :183 */
    if (primary) *primary = -1;
    if (t < 0.0f) {                                                           /* :187-194 */
        v3 tm = tonemap(background_color(c, rd));
        out.x = tm.x; out.y = tm.y; out.z = tm.z; out.w = 1.0f; return out;
    }
    entry = v3add(ro, v3scale(rd, t));                                        /* :196 */
    throughput = v3s(1.0f);
    if (!traverse(c, entry, rd, &hit, &hit_ground)) {                         /* :202-208 */
        v3 tm = tonemap(background_color(c, rd));
        out.x = tm.x; out.y = tm.y; out.z = tm.z; out.w = 1.0f; return out;
    }
    if (primary) *primary = hit_code(s, hit, hit_ground);

    while (bounces < s->max_bounces) {                                        /* :214 */
        basis_t hb; int ix, iy, iz, mat_off; v3 wo, lsWo, lsWi, wi; v4 bf; float k;
        voxel_to_world(c, hit, ro, rd, &hb);                                  /* :221-223 */
        ix = g_f2i(hit.x); iy = g_f2i(hit.y); iz = g_f2i(hit.z);              /* :225 */
        mat_off = fetch_offset(s, ix, iy, iz);                                /* :226 */
        if (ix == s->sel_index[0] && iy == s->sel_index[1] && iz == s->sel_index[2]) { /* :228-233 */
            radiance = v3add(radiance, V3(1.0f, 0.0f, 0.0f));
            break;
        }
        wo = v3neg(rd);                                                       /* :237 */
        lsWo = world_to_local(wo, &hb);
        if (bounces == 0)                                                     /* :241-245 */
            radiance = v3add(radiance, v3mul(throughput, emission_material(c, mat_off)));
        radiance = v3add(radiance, v3mul(throughput, direct_lighting(c, mat_off, &hb, wo, &rng))); /* :248 */
        lsWi = sample_material(c, mat_off, lsWo, &rng, &bf);                  /* :255 */
        if (s->wire_opacity > 0.0f) {                                         /* :260-270 */
            float w = wireframe_factor(c, &hb, hit);
            bf.x *= w; bf.y *= w; bf.z *= w;
        }
        wi = local_to_world(lsWi, &hb);                                       /* :273 */
        k = g_abs(v3dot(wi, hb.normal));                                      /* :276 */
        throughput = v3mul(throughput, v3divs(v3scale(V3(bf.x, bf.y, bf.z), k), bf.w));
        ro = hb.position; rd = wi;                                            /* :278-279 */
        if (!traverse(c, ro, rd, &hit, &hit_ground)) {                        /* :282-289 */
            v4 Lp = evaluate_env(c, rd);
            float mis = power_heuristic(bf.w, Lp.w);
            radiance = v3add(radiance, v3scale(v3mul(throughput, V3(Lp.x, Lp.y, Lp.z)), mis));
            break;
        }
        bounces++;
    }
    {
        v3 tm = tonemap(radiance);                                            /* :294-295 */
        out.x = tm.x; out.y = tm.y; out.z = tm.z; out.w = 1.0f;
    }
    return out;
}

void vto_render_pass(const vto_scene* s, int sample_count, float* out_rgba,
                     int32_t* primary_hit, int32_t* steps, vto_counters* counters, int n_threads)
{
    uint64_t S = 0, R = 0, Hm = 0, E = 0, Q = 0;
    int py;
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads) reduction(+:S,R,Hm,E,Q)
    for (py = 0; py < s->H; py++) {
        ctx_t c; int px;
        ctx_init(&c, s);
        for (px = 0; px < s->W; px++) {
            size_t i = (size_t)px + (size_t)py * (size_t)s->W;
            uint64_t s0 = c.S;
            v4 o = trace_pixel(&c, px, py, sample_count, primary_hit ? &primary_hit[i] : NULL);
            out_rgba[4 * i + 0] = o.x; out_rgba[4 * i + 1] = o.y; out_rgba[4 * i + 2] = o.z; out_rgba[4 * i + 3] = o.w;
            if (steps) steps[i] = (int32_t)(c.S - s0);
        }
        S += c.S; R += c.R; Hm += c.Hm; E += c.E; Q += c.Q;
    }
    if (counters) {
        counters->S += S; counters->R += R; counters->Hm += Hm; counters->E += E; counters->Q += Q;
        counters->paths += (uint64_t)s->W * (uint64_t)s->H;
    }
}

/* K1 on an explicit list of pixels (x, y pairs): the same trace_pixel as vto_render_pass, for comparisons of crops and
 * strided subsets of frames whose full size the CPU cannot render in test time. out_rgba: 4 floats per listed pixel. */
void vto_render_pixels(const vto_scene* s, int sample_count, const int32_t* xy, size_t n, float* out_rgba,
                       int32_t* primary_hit, int n_threads)
{
    long i;
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
    {
        ctx_t c;
        ctx_init(&c, s);
#pragma omp for schedule(dynamic, 64)
        for (i = 0; i < (long)n; i++) {
            v4 o = trace_pixel(&c, xy[2 * i], xy[2 * i + 1], sample_count, primary_hit ? &primary_hit[i] : NULL);
            out_rgba[4 * i + 0] = o.x; out_rgba[4 * i + 1] = o.y; out_rgba[4 * i + 2] = o.z; out_rgba[4 * i + 3] = o.w;
        }
    }
}

/* integrator/editMode.fs:62-142 */
static v4 preview_pixel(ctx_t* c, int px, int py, int sample_count)
{
    const vto_scene* s = c->s;
    v3 frag = V3((float)px + 0.5f, (float)py + 0.5f, 0.55f);
    iv2 rng = rng_offset(px, py, sample_count, s->noise_w, s->noise_h);
    v3 ro, rd, entry, hit, albedo, shadow_hit, bg;
    float t, lighting; int hit_ground; basis_t hb; v4 out;
    const v3 light_dir = { 1.0f, -1.0f, -1.0f };                              /* :41 */
    const float ambient = 0.5f;                                               /* :42 */
    out.w = 1.0f;
    generate_ray(c, frag, &rng, &ro, &rd);
    t = ray_aabb(ro, rd, c->bmin, c->bmax);
    if (t < 0.0f) { bg = background_color(c, rd); out.x = bg.x; out.y = bg.y; out.z = bg.z; return out; }
    entry = v3add(ro, v3scale(rd, t));
    if (!traverse(c, entry, rd, &hit, &hit_ground)) {
        bg = background_color(c, rd); out.x = bg.x; out.y = bg.y; out.z = bg.z; return out;
    }
    voxel_to_world(c, hit, ro, rd, &hb);
    if (g_f2i(hit.x) == s->sel_index[0] && g_f2i(hit.y) == s->sel_index[1] && g_f2i(hit.z) == s->sel_index[2]) {
        out.x = 1.0f; out.y = 0.0f; out.z = 0.0f; return out;                 /* :107-112 */
    }
    albedo = v3s(1.0f);
    if (s->wire_opacity > 0.0f) albedo = v3scale(albedo, wireframe_factor(c, &hb, hit));
    lighting = g_max(0.0f, v3dot(v3neg(rd), hb.normal));                      /* :130 */
    lighting = sqrtf(lighting);
    if (traverse(c, hb.position, v3neg(light_dir), &shadow_hit, &hit_ground)) lighting *= ambient; /* :134-137 */
    out.x = albedo.x * lighting; out.y = albedo.y * lighting; out.z = albedo.z * lighting;
    return out;
}

void vto_preview_pass(const vto_scene* s, int sample_count, float* out_rgba, int n_threads)
{
    int py;
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads)
    for (py = 0; py < s->H; py++) {
        ctx_t c; int px;
        ctx_init(&c, s);
        for (px = 0; px < s->W; px++) {
            size_t i = (size_t)px + (size_t)py * (size_t)s->W;
            v4 o = preview_pixel(&c, px, py, sample_count);
            out_rgba[4 * i + 0] = o.x; out_rgba[4 * i + 1] = o.y; out_rgba[4 * i + 2] = o.z; out_rgba[4 * i + 3] = o.w;
        }
    }
}

/* shared/accumulation.fs:10-18 */
void vto_accumulate(float* avg, const float* sample, int n, size_t count)
{
    size_t i;
    float nf = (float)n, n1 = (float)(n + 1);
    for (i = 0; i < count; i++) avg[i] = (sample[i] + avg[i] * nf) / n1;
}

/* renderer/renderer.cpp:845-850, :926-929 */
void vto_volume_bounds(int X, int Y, int Z, float bmin[3], float bmax[3], float voxel_size[3])
{
    int m = X > Y ? X : Y;
    float vs, sz[3];
    int r[3], i;
    r[0] = X; r[1] = Y; r[2] = Z;
    if (Z > m) m = Z;
    vs = 1000.0f / (float)m;
    for (i = 0; i < 3; i++) {
        sz[i] = vs * (float)r[i];
        bmin[i] = -sz[i] * 0.5f; bmax[i] = sz[i] * 0.5f;
        voxel_size[i] = (bmax[i] - bmin[i]) / (float)r[i];
    }
}

/* ------------------------------------------------------------------------- */
/* services: selectVoxel.vs, focalDistance.vs, addVoxel.vs, removeVoxel.vs     */
/* ------------------------------------------------------------------------- */

/* The reference's picking programs are compiled with PINHOLE only and never bind
 * the noise texture (servicePicking.cpp:21,31-44): the contract is the
 * un-jittered pinhole ray through (px,py) -- generateRay_Pinhole with u = (0.5, 0.5). */
static void pick_ray(const vto_scene* s, float px, float py, float z, v3* ro, v3* rd)
{
    const float* im = s->inv_modelview;
    v3 frag = V3(px, py, z);
    v3 jitter = V3(0.5f - 0.5f, 0.5f - 0.5f, 0.0f);
    v4 es = screen_to_eye_persp(s, v3add(frag, jitter));
    v4 o = m4mulv(im, 0.0f, 0.0f, 0.0f, 1.0f);
    v3 n = v3normalize(V3(es.x, es.y, es.z));
    v4 d = m4mulv(im, n.x, n.y, n.z, 0.0f);
    float len = sqrtf(((d.x * d.x + d.y * d.y) + d.z * d.z) + d.w * d.w);
    *ro = V3(o.x, o.y, o.z);
    *rd = V3(d.x / len, d.y / len, d.z / len);
}

/* editVoxels/selectVoxel.vs:36-71 */
void vto_pick(const vto_scene* s, const float viewport[4], float near_z, float px, float py,
              int32_t index[4], float normal[4])
{
    ctx_t c; v3 ro, rd, p, hit; float t; int hit_ground; basis_t hb;
    (void)viewport;
    ctx_init(&c, s);
    pick_ray(s, px, py, near_z, &ro, &rd);
    t = ray_aabb(ro, rd, c.bmin, c.bmax);
    index[0] = index[1] = index[2] = index[3] = 0;                            /* :47 */
    if (t < 0.0f) return;
    p = v3add(ro, v3scale(rd, t));
    if (!traverse(&c, p, rd, &hit, &hit_ground)) return;
    voxel_to_world(&c, hit, ro, rd, &hb);
    index[0] = g_f2i(hit.x); index[1] = g_f2i(hit.y); index[2] = g_f2i(hit.z); index[3] = 0; /* :69 */
    normal[0] = hb.normal.x; normal[1] = hb.normal.y; normal[2] = hb.normal.z; normal[3] = 0.0f;
}

/* focalDistance/focalDistance.vs:45-81 */
float vto_pick_focal(const vto_scene* s, const float viewport[4], float px, float py)
{
    ctx_t c; v3 ro, rd, p, hit, vmin, vmax; float t; int hit_ground;
    (void)viewport;
    ctx_init(&c, s);
    pick_ray(s, px, py, 0.0f, &ro, &rd);                                      /* :51 */
    t = ray_aabb(ro, rd, c.bmin, c.bmax);
    if (t < 0.0f) return 99999999.0f;                                         /* :43,57 */
    p = v3add(ro, v3scale(rd, t));
    if (!traverse(&c, p, rd, &hit, &hit_ground)) return 99999999.0f;
    vmin = v3add(v3mul(hit, c.vsize), c.bmin);                                /* :76-78 */
    vmax = v3add(vmin, c.vsize);
    return ray_aabb(ro, rd, vmin, vmax);
}

/* editVoxels/addVoxel.vs:16-41 */
int vto_add_voxel(const vto_scene* s, const int32_t sel[4], const float sel_normal[4],
                  float mx, float my, int32_t* grid, int32_t coord[3])
{
    const float* im = s->inv_modelview;
    v4 r4 = m4mulv(im, 1.0f, 0.0f, 0.0f, 0.0f), u4 = m4mulv(im, 0.0f, 1.0f, 0.0f, 0.0f);
    v3 right = V3(r4.x, r4.y, r4.z), up = V3(u4.x, u4.y, u4.z);
    v3 ar = v3abs(right), au = v3abs(up), aar, aau, normal;
    float amx = g_abs(mx), amy = g_abs(my), kx, ky;
    int32_t off;
    aar = v3mul(v3mul(V3(g_step(ar.y, ar.x), g_step(ar.x, ar.y), g_step(ar.x, ar.z)),
                      V3(g_step(ar.z, ar.x), g_step(ar.z, ar.y), g_step(ar.y, ar.z))), v3sign(right)); /* :24 */
    aau = v3mul(v3mul(V3(g_step(au.y, au.x), g_step(au.x, au.y), g_step(au.x, au.z)),
                      V3(g_step(au.z, au.x), g_step(au.z, au.y), g_step(au.y, au.z))), v3sign(up));    /* :25 */
    kx = g_step(amy, amx) * g_sign(mx);                                       /* :28-29 */
    ky = g_step(amx, amy) * g_sign(my);
    if (amx != 0.0f || amy != 0.0f)                                           /* :30 any(bvec2(abs)) */
        normal = v3add(v3scale(aar, kx), v3scale(aau, ky));
    else
        normal = V3(sel_normal[0], sel_normal[1], sel_normal[2]);
    coord[0] = sel[0] + g_f2i(normal.x);                                      /* :35 */
    coord[1] = sel[1] + g_f2i(normal.y);
    coord[2] = sel[2] + g_f2i(normal.z);
    if (coord[0] < 0 || coord[1] < 0 || coord[2] < 0 || coord[0] >= s->X || coord[1] >= s->Y || coord[2] >= s->Z)
        return 0;                                                             /* imageStore outside the image is discarded */
    /* material of the new voxel: copy of the selected voxel's (addVoxel.vs:37-40); a selection that is
     * empty or outside the grid (ground) takes the material at offset 0 (contract, matches U2). */
    off = fetch_offset(s, sel[0], sel[1], sel[2]);
    if (off < 0) off = 0;
    grid[(size_t)coord[0] + (size_t)coord[1] * s->X + (size_t)coord[2] * s->X * s->Y] = off;
    return 1;
}

/* editVoxels/removeVoxel.vs:8-11 (contract N2: the voxel becomes empty) */
int vto_remove_voxel(int X, int Y, int Z, const int32_t sel[4], int32_t* grid)
{
    if (sel[0] < 0 || sel[1] < 0 || sel[2] < 0 || sel[0] >= X || sel[1] >= Y || sel[2] >= Z) return 0;
    grid[(size_t)sel[0] + (size_t)sel[1] * X + (size_t)sel[2] * X * Y] = -1;
    return 1;
}

/* ------------------------------------------------------------------------- */
/* voxelizer: shared/voxelize.vs:20-28, voxelize.gs:55-251                     */
/*            == voxelize/cpuVoxelizer.cpp:42-278                              */
/* ------------------------------------------------------------------------- */
typedef struct { float x, y; } v2;
static inline float v2dot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline v2 V2(float x, float y) { v2 r = { x, y }; return r; }
static inline v2 edge_n(float nc, float ea, float eb) { return (nc >= 0.0f) ? V2(-eb, ea) : V2(eb, -ea); }
/* voxelize.gs:140 : dot(n, .5 - v) + 0.5 * max(abs(n.x), abs(n.y)) */
static inline float edge_d(v2 n, float va, float vb)
{
    return v2dot(n, V2(0.5f - va, 0.5f - vb)) + 0.5f * g_max(g_abs(n.x), g_abs(n.y));
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* voxelize.gs:151-163 (FAT): -dot(n, v) + max(0, n.x) + max(0, n.y) */
static inline float edge_d_fat(v2 n, float va, float vb)
{
    return (-v2dot(n, V2(va, vb)) + g_max(0.0f, n.x)) + g_max(0.0f, n.y);
}

static void voxelize_tri(v3 v0, v3 v1, v3 v2_, int X, int Y, int Z, uint8_t* occ, int fat)
{
    /* swizzleTri, voxelize.gs:55-109 */
    v3 n = v3cross(v3sub(v1, v0), v3sub(v2_, v1));
    v3 an = v3abs(n);
    int axis;
    int lo[3], hi[3], px, py, pz;
    v3 e0, e1, e2, nP, mn, mx;
    v2 n0xy, n1xy, n2xy, n0yz, n1yz, n2yz, n0zx, n1zx, n2zx;
    float d0xy, d1xy, d2xy, d0yz, d1yz, d2yz, d0zx, d1zx, d2zx, dTri, dThin, dFatMin, dFatMax, nzInv;
    if (an.x >= an.y && an.x >= an.z) {
        axis = 0;
        v0 = V3(v0.y, v0.z, v0.x); v1 = V3(v1.y, v1.z, v1.x); v2_ = V3(v2_.y, v2_.z, v2_.x); n = V3(n.y, n.z, n.x);
    } else if (an.y >= an.x && an.y >= an.z) {
        axis = 1;
        v0 = V3(v0.z, v0.x, v0.y); v1 = V3(v1.z, v1.x, v1.y); v2_ = V3(v2_.z, v2_.x, v2_.y); n = V3(n.z, n.x, n.y);
    } else axis = 2;

    /* main, voxelize.gs:244-248: clamp against the un-permuted resolution */
    mn = V3(g_min(g_min(v0.x, v1.x), v2_.x), g_min(g_min(v0.y, v1.y), v2_.y), g_min(g_min(v0.z, v1.z), v2_.z));
    mx = V3(g_max(g_max(v0.x, v1.x), v2_.x), g_max(g_max(v0.y, v1.y), v2_.y), g_max(g_max(v0.z, v1.z), v2_.z));
    lo[0] = g_f2i(g_clamp(floorf(mn.x), 0.0f, (float)X)); hi[0] = g_f2i(g_clamp(ceilf(mx.x), 0.0f, (float)X));
    lo[1] = g_f2i(g_clamp(floorf(mn.y), 0.0f, (float)Y)); hi[1] = g_f2i(g_clamp(ceilf(mx.y), 0.0f, (float)Y));
    lo[2] = g_f2i(g_clamp(floorf(mn.z), 0.0f, (float)Z)); hi[2] = g_f2i(g_clamp(ceilf(mx.z), 0.0f, (float)Z));

    /* voxelizeTriPostSwizzle, voxelize.gs:118-232 */
    e0 = v3sub(v1, v0); e1 = v3sub(v2_, v1); e2 = v3sub(v0, v2_);
    n0xy = edge_n(n.z, e0.x, e0.y); n1xy = edge_n(n.z, e1.x, e1.y); n2xy = edge_n(n.z, e2.x, e2.y);
    n0yz = edge_n(n.x, e0.y, e0.z); n1yz = edge_n(n.x, e1.y, e1.z); n2yz = edge_n(n.x, e2.y, e2.z);
    n0zx = edge_n(n.y, e0.z, e0.x); n1zx = edge_n(n.y, e1.z, e1.x); n2zx = edge_n(n.y, e2.z, e2.x);
    d0xy = edge_d(n0xy, v0.x, v0.y); d1xy = edge_d(n1xy, v1.x, v1.y); d2xy = edge_d(n2xy, v2_.x, v2_.y);
    d0yz = edge_d(n0yz, v0.y, v0.z); d1yz = edge_d(n1yz, v1.y, v1.z); d2yz = edge_d(n2yz, v2_.y, v2_.z);
    d0zx = edge_d(n0zx, v0.z, v0.x); d1zx = edge_d(n1zx, v1.z, v1.x); d2zx = edge_d(n2zx, v2_.z, v2_.x);
    if (fat) {                                                              /* voxelize.gs:151-163 */
        d0xy = edge_d_fat(n0xy, v0.x, v0.y); d1xy = edge_d_fat(n1xy, v1.x, v1.y); d2xy = edge_d_fat(n2xy, v2_.x, v2_.y);
        d0yz = edge_d_fat(n0yz, v0.y, v0.z); d1yz = edge_d_fat(n1yz, v1.y, v1.z); d2yz = edge_d_fat(n2yz, v2_.y, v2_.z);
        d0zx = edge_d_fat(n0zx, v0.z, v0.x); d1zx = edge_d_fat(n1zx, v1.z, v1.x); d2zx = edge_d_fat(n2zx, v2_.z, v2_.x);
    }
    nP = (n.z < 0.0f) ? v3neg(n) : n;
    dTri = v3dot(nP, v0);
    dThin = dTri - v2dot(V2(nP.x, nP.y), V2(0.5f, 0.5f));
    dFatMin = (dTri - g_max(nP.x, 0.0f)) - g_max(nP.y, 0.0f);               /* :170-171 */
    dFatMax = (dTri - g_min(nP.x, 0.0f)) - g_min(nP.y, 0.0f);
    nzInv = 1.0f / nP.z;

    for (px = lo[0]; px < hi[0]; px++) {
        for (py = lo[1]; py < hi[1]; py++) {
            v2 pxy = V2((float)px, (float)py);
            float a0 = d0xy + v2dot(n0xy, pxy), a1 = d1xy + v2dot(n1xy, pxy), a2 = d2xy + v2dot(n2xy, pxy);
            if ((a0 >= 0.0f) && (a1 >= 0.0f) && (a2 >= 0.0f)) {
                float dot_n_p = v2dot(V2(nP.x, nP.y), pxy);
                float zMinInt = fat ? (-dot_n_p + dFatMin) * nzInv : (-dot_n_p + dThin) * nzInv;     /* :195-201 */
                float zMaxInt = fat ? (-dot_n_p + dFatMax) * nzInv : zMinInt;
                float zf = floorf(zMinInt), zc = ceilf(zMaxInt);
                int zMin = g_f2i(zf) - (zf == zMinInt ? 1 : 0);
                int zMax = g_f2i(zc) + (zc == zMaxInt ? 1 : 0);
                if (zMin < lo[2]) zMin = lo[2];
                if (zMax > hi[2]) zMax = hi[2];
                for (pz = zMin; pz < zMax; pz++) {
                    v2 pyz = V2((float)py, (float)pz), pzx = V2((float)pz, (float)px);
                    float b0 = d0yz + v2dot(n0yz, pyz), b1 = d1yz + v2dot(n1yz, pyz), b2 = d2yz + v2dot(n2yz, pyz);
                    float c0 = d0zx + v2dot(n0zx, pzx), c1 = d1zx + v2dot(n1zx, pzx), c2 = d2zx + v2dot(n2zx, pzx);
                    if ((b0 >= 0.0f) && (b1 >= 0.0f) && (b2 >= 0.0f) && (c0 >= 0.0f) && (c1 >= 0.0f) && (c2 >= 0.0f)) {
                        int ox, oy, oz;
                        /* unswizzle: voxelize.gs:40-48 */
                        if (axis == 0) { ox = pz; oy = px; oz = py; }
                        else if (axis == 1) { ox = py; oy = pz; oz = px; }
                        else { ox = px; oy = py; oz = pz; }
                        /* imageStore outside the image is discarded (non-cubic grids only) */
                        if (ox >= 0 && oy >= 0 && oz >= 0 && ox < X && oy < Y && oz < Z)
                            occ[(size_t)ox + (size_t)oy * X + (size_t)oz * X * Y] = 1;
                    }
                }
            }
        }
    }
    (void)clampi;
}

static void voxelize_mesh(const float* xyz, size_t n_verts, const uint32_t* idx, size_t n_idx,
                          const float M[16], int X, int Y, int Z, uint8_t* occ, int n_threads, int fat)
{
    long t, ntri = (long)(n_idx / 3);
    v3* vs = (v3*)malloc(sizeof(v3) * (n_verts ? n_verts : 1));
    size_t i;
    /* voxelize.vs:20-28 */
    for (i = 0; i < n_verts; i++) {
        v4 l = m4mulv(M, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 1.0f);
        vs[i] = V3(l.x * (float)X, l.y * (float)Y, l.z * (float)Z);
    }
    if (n_threads < 1) n_threads = 1;
    /* all writers store the same value: the benign race of cpuVoxelizer.cpp:99-108 */
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
    for (t = 0; t < ntri; t++)
        voxelize_tri(vs[idx[3 * t]], vs[idx[3 * t + 1]], vs[idx[3 * t + 2]], X, Y, Z, occ, fat);
    free(vs);
}
void vto_voxelize(const float* xyz, size_t n_verts, const uint32_t* idx, size_t n_idx,
                  const float M[16], int X, int Y, int Z, uint8_t* occ, int n_threads)
{
    voxelize_mesh(xyz, n_verts, idx, n_idx, M, X, Y, Z, occ, n_threads, 0);
}
/* the FAT variant (voxelize.gs:15-19, `#define THICKNESS FAT`: adjacent voxels share at least a face) */
void vto_voxelize_fat(const float* xyz, size_t n_verts, const uint32_t* idx, size_t n_idx,
                      const float M[16], int X, int Y, int Z, uint8_t* occ, int n_threads)
{
    voxelize_mesh(xyz, n_verts, idx, n_idx, M, X, Y, Z, occ, n_threads, 1);
}

/* ------------------------------------------------------------------------- */
/* renderer/image.cpp:68-283 (calculateCDF) + :349-389 (calculateImageIntegral) */
/* on an already filtered single-channel image                                 */
/* ------------------------------------------------------------------------- */
void vto_build_cdf(const float* lum, int w, int h, float* cdf_u, float* cdf_v, float* integral)
{
    float* fu = (float*)malloc(sizeof(float) * (size_t)w * h);
    float* fv = (float*)malloc(sizeof(float) * (size_t)(h + 1));
    const float iW = (float)w, iH = (float)h, iA = iW * iH;
    float sum = 0.0f, img;
    int x, y;
    const int cw = w + 1;
    for (y = 0; y < h; y++) {                                                  /* image.cpp:361-375 */
        float sinT = (float)sin(M_PI * ((float)y + 0.5f) / iH);
        for (x = 0; x < w; x++) {
            float v = lum[(size_t)y * w + x];
            v = (0.f < v) ? v : 0.f;                                           /* std::max(0.f, value) */
            fu[(size_t)y * w + x] = v * sinT;
            sum += v * sinT;
        }
    }
    {   /* image.cpp:381-386 : float * double constants */
        float e = sum / iA;
        e = (float)((double)e * (2.0f * M_PI * M_PI));
        *integral = e;
    }
    for (y = 0; y < h; y++) {                                                  /* image.cpp:212-247 */
        size_t row = (size_t)y * cw;
        float rowI;
        const unsigned int steps = (unsigned int)iW;
        cdf_u[row] = 0.0f;
        for (x = 1; x <= w; x++) {
            float f = fu[(size_t)y * w + x - 1] / steps;
            cdf_u[row + x] = cdf_u[row + x - 1] + f;
        }
        rowI = cdf_u[row + w];
        fv[y] = rowI;
        if (rowI > 0.0f) for (x = 1; x <= w; x++) cdf_u[row + x] /= rowI;
        else for (x = 1; x <= w; x++) cdf_u[row + x] = (float)x / steps;
    }
    cdf_v[0] = 0.0f;                                                           /* image.cpp:251-280 */
    for (y = 1; y <= h; y++) cdf_v[y] = cdf_v[y - 1] + fv[y - 1] / (unsigned int)h;
    img = cdf_v[h];
    if (img > 0.0f) for (y = 1; y <= h; y++) cdf_v[y] /= img;
    else for (y = 1; y <= h; y++) cdf_v[y] = (float)y / iH;
    free(fu); free(fv);
}

/* DDA alone (dda.h:63-100) on explicit rays: out[4i..] = (hit voxel xyz, code 1 voxel / 2 ground / 0 miss) */
void vto_trace_rays(const vto_scene* s, const float* rays, size_t n, float* out)
{
    ctx_t c; size_t i;
    ctx_init(&c, s);
    for (i = 0; i < n; i++) {
        v3 hit; int g;
        int h = traverse(&c, V3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]),
                         V3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]), &hit, &g);
        out[4 * i] = hit.x; out[4 * i + 1] = hit.y; out[4 * i + 2] = hit.z;
        out[4 * i + 3] = h ? (g ? 2.0f : 1.0f) : 0.0f;
    }
}

/* scalar math exposed for the accuracy tests */
float vto_m_sin(float x) { return g_sin(x); }
float vto_m_cos(float x) { return g_cos(x); }
float vto_m_acos(float x) { return g_acos(x); }
float vto_m_atan2(float y, float x) { return g_atan2(y, x); }
float vto_m_pow(float x, float y) { return g_pow(x, y); }
float vto_m_exp2(float x) { return vto_exp2(x); }
float vto_m_log2(float x) { return vto_log2(x); }
