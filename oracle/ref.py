"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libvt_ref.so -- the reference's OWN GLSL programs
(/root/reference/src/shaders) compiled for the CPU by oracle/shim/Makefile (glsl2cpp.py + glsl_emu.h + ref_glsl.cpp).

Used (a) by tests/ to pin the C restatement oracle/vto.c against the real shader text, and (b) by bench.py's CPU arm
(`cpu_baseline.kind = "reference"`, `--impl reference`). The library is built in this container, where /root/reference
exists, and travels to the GPU box as a prebuilt file; nothing here reads /root/reference at run time.
The product (voxeltoy_b200/) never imports this module.
"""
import ctypes as C
import os

import numpy as np

from . import vto

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libvt_ref.so")
_LIB = None

f32p, i32p = vto.f32p, vto.i32p


def available():
    return os.path.exists(_PATH)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(_PATH)
        S = C.POINTER(vto.Scene)
        L.vtref_describe.restype = C.c_char_p
        L.vtref_render_pass.argtypes = [S, C.c_int, f32p, C.c_int]
        L.vtref_render_pixels.argtypes = [S, C.c_int, i32p, C.c_size_t, f32p, C.c_int]
        L.vtref_preview_pass.argtypes = [S, C.c_int, f32p, C.c_int]
        L.vtref_pick.argtypes = [S, C.c_float, C.c_float, C.c_float, f32p, i32p, f32p]
        L.vtref_pick_focal.argtypes = [S, C.c_float, C.c_float]; L.vtref_pick_focal.restype = C.c_float
        L.vtref_add_voxel.argtypes = [S, i32p, f32p, C.c_float, C.c_float, i32p]
        L.vtref_remove_voxel.argtypes = [i32p, i32p]
        L.vtref_voxelize.argtypes = [f32p, C.c_size_t, C.POINTER(C.c_uint32), C.c_size_t, f32p, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_uint8), C.c_int]
        L.vtref_voxelize_fat.argtypes = L.vtref_voxelize.argtypes
        _LIB = L
    return _LIB


def describe():
    return lib().vtref_describe().decode()


make_scene = vto.make_scene          # the same `vto_scene` struct feeds both implementations


def _fp(a):
    return a.ctypes.data_as(f32p)


def _ip(a):
    return a.ctypes.data_as(i32p)


def render_pass(scene, sample_count, n_threads=None):
    """integrator/pathTracer.fs over the whole frame: (H, W, 4) float32, row 0 = bottom row."""
    out = np.empty((scene.H, scene.W, 4), np.float32)
    lib().vtref_render_pass(C.byref(scene), int(sample_count), _fp(out), int(n_threads or os.cpu_count() or 1))
    return out


def render_pixels(scene, sample_count, xy, n_threads=None):
    """pathTracer.fs at the listed (x, y) fragments only: (n, 4) float32."""
    xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
    out = np.empty((xy.shape[0], 4), np.float32)
    lib().vtref_render_pixels(C.byref(scene), int(sample_count), _ip(xy), xy.shape[0], _fp(out), int(n_threads or os.cpu_count() or 1))
    return out


def preview_pass(scene, sample_count, n_threads=None):
    out = np.empty((scene.H, scene.W, 4), np.float32)
    lib().vtref_preview_pass(C.byref(scene), int(sample_count), _fp(out), int(n_threads or os.cpu_count() or 1))
    return out


def pick(scene, px, py, near_z=0.1, prev_normal=(1.0, 0.0, 0.0, 0.0)):
    index = np.zeros(4, np.int32); normal = np.zeros(4, np.float32)
    pn = np.array(prev_normal, np.float32)
    lib().vtref_pick(C.byref(scene), near_z, px, py, _fp(pn), _ip(index), _fp(normal))
    return index, normal


def pick_focal(scene, px, py):
    return float(lib().vtref_pick_focal(C.byref(scene), px, py))


def add_voxel(scene, sel_index, sel_normal, mx, my):
    coord = np.zeros(3, np.int32)
    si = np.ascontiguousarray(sel_index, np.int32); sn = np.ascontiguousarray(sel_normal, np.float32)
    lib().vtref_add_voxel(C.byref(scene), _ip(si), _fp(sn), mx, my, _ip(coord))
    return coord


def remove_voxel(sel_index):
    coord = np.zeros(3, np.int32)
    si = np.ascontiguousarray(sel_index, np.int32)
    lib().vtref_remove_voxel(_ip(si), _ip(coord))
    return coord


def voxelize(verts, idx, M, res, n_threads=None, fat=False):
    verts = np.ascontiguousarray(verts, np.float32); idx = np.ascontiguousarray(idx, np.uint32)
    M = np.ascontiguousarray(M, np.float32)
    X, Y, Z = [int(v) for v in res]
    occ = np.zeros(X * Y * Z, np.uint8)
    fn = lib().vtref_voxelize_fat if fat else lib().vtref_voxelize
    fn(_fp(verts), verts.size // 3, idx.ctypes.data_as(C.POINTER(C.c_uint32)), idx.size, _fp(M), X, Y, Z,
       occ.ctypes.data_as(C.POINTER(C.c_uint8)), int(n_threads or os.cpu_count() or 1))
    return occ


# ---- the reference's OBJ reader (thirdParty/tinyobjloader/tiny_obj_loader.cc + the merge of mesh/meshLoader.cpp:27-64) ----
_OBJ_PATH = os.path.join(_HERE, "_ref", "libvt_ref_obj.so")
_OBJ_LIB = None


def obj_available():
    return os.path.exists(_OBJ_PATH)


def load_obj(path):
    """MeshLoader::loadFromOBJ through the reference's own tinyobjloader: (verts (n, 3) float32, indices uint32)."""
    global _OBJ_LIB
    if _OBJ_LIB is None:
        L = C.CDLL(_OBJ_PATH)
        L.vtref_load_obj.argtypes = [C.c_char_p, C.POINTER(f32p), C.POINTER(C.c_size_t), C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_size_t)]
        L.vtref_obj_free.argtypes = [C.c_void_p]
        _OBJ_LIB = L
    v = f32p(); nv = C.c_size_t(); i = C.POINTER(C.c_uint32)(); ni = C.c_size_t()
    if _OBJ_LIB.vtref_load_obj(path.encode(), C.byref(v), C.byref(nv), C.byref(i), C.byref(ni)) != 0:
        raise IOError("tinyobj::LoadObj failed on " + path)
    verts = np.ctypeslib.as_array(v, shape=(max(nv.value, 1),))[:nv.value].copy().reshape(-1, 3)
    idx = np.ctypeslib.as_array(i, shape=(max(ni.value, 1),))[:ni.value].copy()
    _OBJ_LIB.vtref_obj_free(v); _OBJ_LIB.vtref_obj_free(i)
    return verts, idx
