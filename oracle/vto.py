"""TEST INFRASTRUCTURE: ctypes binding of the CPU oracle (oracle/libvto.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module. The product (voxeltoy_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)


class Scene(C.Structure):
    """Mirror of `vto_scene` (oracle/vto.h)."""
    _fields_ = [
        ("X", C.c_int32), ("Y", C.c_int32), ("Z", C.c_int32),
        ("grid", i32p),
        ("materials", f32p), ("n_materials", C.c_int32),
        ("emissive", i32p), ("n_emissive", C.c_int32),
        ("noise", f32p), ("noise_w", C.c_int32), ("noise_h", C.c_int32),
        ("use_image", C.c_int32),
        ("env_rgb", f32p), ("env_w", C.c_int32), ("env_h", C.c_int32),
        ("cdf_u", f32p), ("cdf_u_w", C.c_int32), ("cdf_u_h", C.c_int32),
        ("cdf_v", f32p), ("cdf_v_n", C.c_int32),
        ("env_integral", C.c_float), ("env_rotation", C.c_float),
        ("bg_top", C.c_float * 3), ("bg_bottom", C.c_float * 3),
        ("inv_modelview", C.c_float * 16), ("proj", C.c_float * 16), ("inv_proj", C.c_float * 16),
        ("lens_radius", C.c_float), ("lens_model", C.c_int32), ("focal_distance", C.c_float),
        ("W", C.c_int32), ("H", C.c_int32), ("max_bounces", C.c_int32),
        ("wire_opacity", C.c_float), ("wire_thickness", C.c_float),
        ("sel_index", C.c_int32 * 3),
        ("bmin", C.c_float * 3), ("bmax", C.c_float * 3), ("voxel_size", C.c_float * 3),
    ]


class Counters(C.Structure):
    _fields_ = [("S", C.c_uint64), ("R", C.c_uint64), ("Hm", C.c_uint64),
                ("E", C.c_uint64), ("Q", C.c_uint64), ("paths", C.c_uint64)]


def build():
    """Compile libvto.so (and oracle/_ref when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "libvto.so"])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "libvto.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.vto_volume_bounds.argtypes = [C.c_int, C.c_int, C.c_int, f32p, f32p, f32p]
    L.vto_render_pass.argtypes = [C.POINTER(Scene), C.c_int, f32p, i32p, i32p, C.POINTER(Counters), C.c_int]
    L.vto_render_pixels.argtypes = [C.POINTER(Scene), C.c_int, i32p, C.c_size_t, f32p, i32p, C.c_int]
    L.vto_preview_pass.argtypes = [C.POINTER(Scene), C.c_int, f32p, C.c_int]
    L.vto_accumulate.argtypes = [f32p, f32p, C.c_int, C.c_size_t]
    L.vto_voxelize.argtypes = [f32p, C.c_size_t, C.POINTER(C.c_uint32), C.c_size_t, f32p,
                               C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_int]
    L.vto_voxelize_fat.argtypes = L.vto_voxelize.argtypes
    L.vto_pick.argtypes = [C.POINTER(Scene), f32p, C.c_float, C.c_float, C.c_float, i32p, f32p]
    L.vto_pick_focal.argtypes = [C.POINTER(Scene), f32p, C.c_float, C.c_float]
    L.vto_pick_focal.restype = C.c_float
    L.vto_add_voxel.argtypes = [C.POINTER(Scene), i32p, f32p, C.c_float, C.c_float, i32p, i32p]
    L.vto_add_voxel.restype = C.c_int
    L.vto_remove_voxel.argtypes = [C.c_int, C.c_int, C.c_int, i32p, i32p]
    L.vto_remove_voxel.restype = C.c_int
    L.vto_trace_rays.argtypes = [C.POINTER(Scene), f32p, C.c_size_t, f32p]
    L.vto_noise_table.argtypes = [f32p, C.c_size_t]
    L.vto_hash.argtypes = [C.c_uint32]
    L.vto_hash.restype = C.c_uint32
    L.vto_rng_offset.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.vto_dda_step_cap.argtypes = [C.c_int, C.c_int, C.c_int]
    L.vto_dda_step_cap.restype = C.c_int
    L.vto_resize_box.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
    L.vto_build_cdf.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p, f32p]
    for name in ("sin", "cos", "acos", "exp2", "log2"):
        fn = getattr(L, "vto_m_" + name)
        fn.argtypes = [C.c_float]
        fn.restype = C.c_float
    for name in ("atan2", "pow"):
        fn = getattr(L, "vto_m_" + name)
        fn.argtypes = [C.c_float, C.c_float]
        fn.restype = C.c_float
    _LIB = L
    return L


def _fp(a):
    return a.ctypes.data_as(f32p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(i32p) if a is not None else None


_NOISE = {}


def noise_table(w=1024, h=1024):
    """renderer.cpp:741-758: the first w*h*4 outputs of glibc rand()/RAND_MAX, seed 1."""
    key = (w, h)
    if key not in _NOISE:
        a = np.empty(w * h * 4, dtype=np.float32)
        lib().vto_noise_table(_fp(a), a.size)
        _NOISE[key] = a.reshape(h, w, 4)
    return _NOISE[key]


def volume_bounds(X, Y, Z):
    bmin = np.zeros(3, np.float32); bmax = np.zeros(3, np.float32); vs = np.zeros(3, np.float32)
    lib().vto_volume_bounds(X, Y, Z, _fp(bmin), _fp(bmax), _fp(vs))
    return bmin, bmax, vs


def make_scene(d):
    """Build a `Scene` from the plain dict produced by oracle/scene.py (keeps arrays alive)."""
    s = Scene()
    keep = []

    def arr(key, dtype):
        a = d.get(key)
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dtype)
        keep.append(a)
        return a

    grid = arr("grid", np.int32)
    s.X, s.Y, s.Z = [int(v) for v in d["res"]]
    assert grid.size == s.X * s.Y * s.Z
    s.grid = _ip(grid)
    mats = arr("materials", np.float32)
    s.materials = _fp(mats); s.n_materials = mats.size
    em = arr("emissive", np.int32)
    s.emissive = _ip(em) if em is not None and em.size else None
    s.n_emissive = 0 if em is None else em.size
    noise = arr("noise", np.float32)
    if noise is None:
        noise = noise_table(); keep.append(noise)
    s.noise = _fp(noise); s.noise_h, s.noise_w = noise.shape[0], noise.shape[1]
    env = d.get("env")
    if env is not None:
        rgb = np.ascontiguousarray(env["rgb"], np.float32); keep.append(rgb)
        cu = np.ascontiguousarray(env["cdf_u"], np.float32); keep.append(cu)
        cv = np.ascontiguousarray(env["cdf_v"], np.float32); keep.append(cv)
        s.use_image = 1
        s.env_rgb = _fp(rgb); s.env_h, s.env_w = rgb.shape[0], rgb.shape[1]
        s.cdf_u = _fp(cu); s.cdf_u_h, s.cdf_u_w = cu.shape
        s.cdf_v = _fp(cv); s.cdf_v_n = cv.size
        s.env_integral = float(env["integral"])
        s.env_rotation = float(env.get("rotation", 0.0))
    else:
        s.use_image = 0
    for i in range(3):
        s.bg_top[i] = float(d["bg_top"][i]); s.bg_bottom[i] = float(d["bg_bottom"][i])
    for i in range(16):
        s.inv_modelview[i] = float(d["inv_modelview"].reshape(-1)[i])
        s.proj[i] = float(d["proj"].reshape(-1)[i])
        s.inv_proj[i] = float(d["inv_proj"].reshape(-1)[i])
    s.lens_radius = float(d.get("lens_radius", 0.0))
    s.lens_model = int(d.get("lens_model", 0))
    s.focal_distance = float(d.get("focal_distance", 99999999.0))
    s.W, s.H = int(d["W"]), int(d["H"])
    s.max_bounces = int(d.get("max_bounces", 1))
    s.wire_opacity = float(d.get("wire_opacity", 0.0))
    s.wire_thickness = float(d.get("wire_thickness", 0.01))
    sel = d.get("sel_index", (-1, -1, -1))
    for i in range(3):
        s.sel_index[i] = int(sel[i])
    bmin, bmax, vs = volume_bounds(s.X, s.Y, s.Z)
    for i in range(3):
        s.bmin[i], s.bmax[i], s.voxel_size[i] = float(bmin[i]), float(bmax[i]), float(vs[i])
    s._keep = keep
    return s


def render_pass(scene, sample_count, n_threads=None, want_hits=True, want_steps=False):
    n_threads = n_threads or os.cpu_count() or 1
    W, H = scene.W, scene.H
    out = np.empty((H, W, 4), np.float32)
    hits = np.empty((H, W), np.int32) if want_hits else None
    steps = np.empty((H, W), np.int32) if want_steps else None
    cnt = Counters()
    lib().vto_render_pass(C.byref(scene), sample_count, _fp(out), _ip(hits), _ip(steps), C.byref(cnt), n_threads)
    counters = {k: getattr(cnt, k) for k, _ in Counters._fields_}
    return out, hits, steps, counters


def render_pixels(scene, sample_count, xy, n_threads=None, want_hits=False):
    """trace_pixel on the listed (x, y) pixels only: (n, 4) float32 (+ (n,) primary hits)."""
    xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
    out = np.empty((xy.shape[0], 4), np.float32)
    hits = np.empty(xy.shape[0], np.int32) if want_hits else None
    lib().vto_render_pixels(C.byref(scene), int(sample_count), _ip(xy), xy.shape[0], _fp(out), _ip(hits), n_threads or os.cpu_count() or 1)
    return (out, hits) if want_hits else out


def preview_pass(scene, sample_count, n_threads=None):
    n_threads = n_threads or os.cpu_count() or 1
    out = np.empty((scene.H, scene.W, 4), np.float32)
    lib().vto_preview_pass(C.byref(scene), sample_count, _fp(out), n_threads)
    return out


def accumulate(avg, sample, n):
    assert avg.dtype == np.float32 and sample.dtype == np.float32 and avg.flags.c_contiguous
    lib().vto_accumulate(_fp(avg), _fp(np.ascontiguousarray(sample)), n, avg.size)
    return avg


def render_average(scene, n_passes, first=0, n_threads=None):
    """renderer.cpp:594-611: K1 then K2 for sampleCount = first .. first+n_passes-1 (counter starts at 0)."""
    avg = np.zeros((scene.H, scene.W, 4), np.float32)
    for n in range(n_passes):
        s, _, _, _ = render_pass(scene, first + n, n_threads, want_hits=False)
        accumulate(avg, s, n)
    return avg


def voxelize(verts, idx, M, res, n_threads=None, fat=False):
    n_threads = n_threads or os.cpu_count() or 1
    verts = np.ascontiguousarray(verts, np.float32); idx = np.ascontiguousarray(idx, np.uint32)
    M = np.ascontiguousarray(M, np.float32)
    X, Y, Z = [int(v) for v in res]
    occ = np.zeros(X * Y * Z, np.uint8)
    (lib().vto_voxelize_fat if fat else lib().vto_voxelize)(_fp(verts), verts.size // 3, idx.ctypes.data_as(C.POINTER(C.c_uint32)), idx.size,
                                                            _fp(M), X, Y, Z, occ.ctypes.data_as(C.POINTER(C.c_uint8)), n_threads)
    return occ


def pick(scene, px, py, near_z=0.1, prev_normal=(1.0, 0.0, 0.0, 0.0)):
    """selectVoxel.vs: index is reset to 0 on a miss, normal keeps the SSBO's previous value (:47-62)."""
    vp = np.array([0, 0, scene.W, scene.H], np.float32)
    index = np.zeros(4, np.int32); normal = np.array(prev_normal, np.float32)
    lib().vto_pick(C.byref(scene), _fp(vp), near_z, px, py, _ip(index), _fp(normal))
    return index, normal


def pick_focal(scene, px, py):
    vp = np.array([0, 0, scene.W, scene.H], np.float32)
    return float(lib().vto_pick_focal(C.byref(scene), _fp(vp), px, py))


def add_voxel(scene, grid, sel_index, sel_normal, mx, my):
    coord = np.zeros(3, np.int32)
    si = np.ascontiguousarray(sel_index, np.int32); sn = np.ascontiguousarray(sel_normal, np.float32)
    ok = lib().vto_add_voxel(C.byref(scene), _ip(si), _fp(sn), mx, my, _ip(grid), _ip(coord))
    return bool(ok), coord


def remove_voxel(grid, res, sel_index):
    si = np.ascontiguousarray(sel_index, np.int32)
    return bool(lib().vto_remove_voxel(int(res[0]), int(res[1]), int(res[2]), _ip(si), _ip(grid)))


def trace_rays(scene, rays):
    r = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    out = np.empty((r.shape[0], 4), np.float32)
    lib().vto_trace_rays(C.byref(scene), _fp(r), r.shape[0], _fp(out))
    return out


def resize_box(rgb, nw, nh):
    """Area-weighted box reduction of an (h, w, 3) float32 image (the contract that stands in for OIIO's resize)."""
    rgb = np.ascontiguousarray(rgb, np.float32)
    out = np.empty((nh, nw, 3), np.float32)
    lib().vto_resize_box(_fp(rgb), rgb.shape[1], rgb.shape[0], int(nw), int(nh), _fp(out))
    return out


def build_cdf(lum):
    lum = np.ascontiguousarray(lum, np.float32)
    h, w = lum.shape
    cu = np.empty((h, w + 1), np.float32); cv = np.empty(h + 1, np.float32)
    integral = C.c_float(0)
    lib().vto_build_cdf(_fp(lum), w, h, _fp(cu), _fp(cv), C.cast(C.byref(integral), f32p))
    return cu, cv, float(integral.value)


def hash32(x):
    return int(lib().vto_hash(x & 0xFFFFFFFF))


def rng_offset(px, py, seq, rw=1024, rh=1024):
    out = (C.c_int * 2)()
    lib().vto_rng_offset(px, py, seq, rw, rh, out)
    return out[0], out[1]


def dda_step_cap(X, Y, Z):
    return int(lib().vto_dda_step_cap(X, Y, Z))
