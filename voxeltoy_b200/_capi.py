"""ctypes binding of the C ABI (include/voxeltoy_b200.h). There is no CPU path: a missing
library or a missing CUDA device raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VT_LIB_PATH") or os.path.join(_HERE, "libvoxeltoy_b200.so")   # VT_LIB_PATH: developer A/B builds

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)

VT_LENS_PINHOLE, VT_LENS_THIN, VT_LENS_ORTHO = 0, 1, 2
VT_INTEGRATOR_PATHTRACER, VT_INTEGRATOR_EDIT_MODE = 0, 1
VT_PART_NONE, VT_PART_TILES, VT_PART_SAMPLES = 0, 1, 2


class VtCamera(C.Structure):
    _fields_ = [("inv_modelview", C.c_float * 16), ("proj", C.c_float * 16), ("inv_proj", C.c_float * 16),
                ("near_z", C.c_float), ("far_z", C.c_float), ("lens_radius", C.c_float), ("lens_model", C.c_int32)]


class VtSettings(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("max_bounces", C.c_int32), ("integrator", C.c_int32),
                ("bg_top", C.c_float * 3), ("bg_bottom", C.c_float * 3), ("use_env_image", C.c_int32),
                ("env_rotation_rad", C.c_float), ("wireframe_opacity", C.c_float), ("wireframe_thickness", C.c_float)]


class VtCounters(C.Structure):
    _fields_ = [("dda_steps", C.c_uint64), ("rand_calls", C.c_uint64), ("material_evals", C.c_uint64),
                ("cdf_loads", C.c_uint64), ("env_lookups", C.c_uint64), ("paths", C.c_uint64),
                ("kernel_launches", C.c_uint64)]


# every symbol include/voxeltoy_b200.h declares: name -> (restype, argtypes)
P = C.c_void_p
SIGNATURES = {
    "vt_create": (C.c_int, [C.c_int, C.POINTER(P)]),
    "vt_destroy": (None, [P]),
    "vt_last_error": (C.c_char_p, [P]),
    "vt_set_logger": (C.c_int, [P, C.c_void_p, C.c_void_p]),
    "vt_set_stream": (C.c_int, [P, C.c_void_p]),
    "vt_sync": (C.c_int, [P]),
    "vt_volume_upload": (C.c_int, [P, i32p, C.c_int, C.c_int, C.c_int]),
    "vt_materials_upload": (C.c_int, [P, f32p, C.c_size_t]),
    "vt_material_update": (C.c_int, [P, C.c_uint32, f32p, C.c_int]),
    "vt_emissive_upload": (C.c_int, [P, i32p, C.c_size_t]),
    "vt_read_volume": (C.c_int, [P, i32p]),
    "vt_read_materials": (C.c_int, [P, f32p, C.c_size_t]),
    "vt_get_volume_info": (C.c_int, [P, i32p, f32p, f32p, f32p]),
    "vt_noise_upload": (C.c_int, [P, f32p, C.c_int, C.c_int]),
    "vt_env_upload": (C.c_int, [P, f32p, C.c_int, C.c_int, f32p, C.c_int, C.c_int, f32p, C.c_int, C.c_float]),
    "vt_env_clear": (C.c_int, [P]),
    "vt_env_build": (C.c_int, [P, f32p, C.c_int, C.c_int]),
    "vt_get_env_info": (C.c_int, [P, i32p, f32p, f32p]),
    "vt_read_env_cdf": (C.c_int, [P, f32p, f32p]),
    "vt_set_camera": (C.c_int, [P, C.POINTER(VtCamera)]),
    "vt_set_settings": (C.c_int, [P, C.POINTER(VtSettings)]),
    "vt_set_focal_distance": (C.c_int, [P, C.c_float]),
    "vt_get_focal_distance": (C.c_int, [P, f32p]),
    "vt_set_selection": (C.c_int, [P, i32p, f32p]),
    "vt_get_selection": (C.c_int, [P, i32p, f32p]),
    "vt_reset_accumulation": (C.c_int, [P]),
    "vt_render": (C.c_int, [P, C.c_int, C.c_int]),
    "vt_get_num_samples": (C.c_int, [P, C.POINTER(C.c_int)]),
    "vt_read_display": (C.c_int, [P, C.c_void_p, C.c_int]),
    "vt_read_average": (C.c_int, [P, f32p]),
    "vt_read_primary_hits": (C.c_int, [P, i32p]),
    "vt_enable_primary_hits": (C.c_int, [P, C.c_int]),
    "vt_set_partition": (C.c_int, [P, C.c_int, C.c_int, C.c_int]),
    "vt_accum_device_ptr": (C.c_void_p, [P]),
    "vt_set_kernel_variant": (C.c_int, [P, C.c_int]),
    "vt_kernel_timing_enable": (C.c_int, [P, C.c_int]),
    "vt_get_kernel_times": (C.c_int, [P, C.c_void_p]),
    "vt_set_wavefront_max_paths": (C.c_int, [P, C.c_size_t]),
    "vt_set_wavefront_lanes": (C.c_int, [P, C.c_int]),
    "vt_set_empty_skip": (C.c_int, [P, C.c_int]),
    "vt_measure_l2_bandwidth": (C.c_int, [P, C.c_size_t, C.c_int, f32p]),
    "vt_debug_advance": (C.c_int, [P, f32p, f32p, f32p, i32p, C.c_size_t, f32p, i32p, C.c_int]),
    "vt_debug_div_const": (C.c_int, [P, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "vt_counters_enable": (C.c_int, [P, C.c_int]),
    "vt_get_counters": (C.c_int, [P, C.POINTER(VtCounters)]),
    "vt_reset_counters": (C.c_int, [P]),
    "vt_voxelize": (C.c_int, [P, f32p, C.c_size_t, u32p, C.c_size_t, f32p, C.c_int, C.c_int, C.c_int, C.c_int32]),
    "vt_volume_assign_materials": (C.c_int, [P, i32p, C.c_int, C.c_int]),
    "vt_set_voxelize_thickness": (C.c_int, [P, C.c_int]),
    "vt_get_last_voxelize_ms": (C.c_int, [P, f32p]),
    "vt_get_last_voxelize_full_ms": (C.c_int, [P, f32p]),
    "vt_pick": (C.c_int, [P, C.c_float, C.c_float]),
    "vt_pick_focal": (C.c_int, [P, C.c_float, C.c_float]),
    "vt_add_voxel": (C.c_int, [P, C.c_float, C.c_float]),
    "vt_remove_voxel": (C.c_int, [P]),
    "vt_device_count": (C.c_int, []),
    "vt_version": (C.c_char_p, []),
    "vt_debug_trace_rays": (C.c_int, [P, f32p, C.c_size_t, f32p]),
    # render groups (multi-GPU): G = vt_group*
    "vt_group_create": (C.c_int, [C.c_int, i32p, C.c_int, C.POINTER(C.c_void_p)]),
    "vt_group_adopt": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]),
    "vt_group_unique_id": (C.c_int, [C.c_void_p]),
    "vt_group_join": (C.c_int, [P, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vt_group_destroy": (None, [C.c_void_p]),
    "vt_group_last_error": (C.c_char_p, [C.c_void_p]),
    "vt_group_size": (C.c_int, [C.c_void_p]),
    "vt_group_local_size": (C.c_int, [C.c_void_p]),
    "vt_group_context": (C.c_void_p, [C.c_void_p, C.c_int]),
    "vt_group_rank": (C.c_int, [C.c_void_p, C.c_int]),
    "vt_group_set_exchange": (C.c_int, [C.c_void_p, C.c_int]),
    "vt_group_get_exchange": (C.c_int, [C.c_void_p]),
    "vt_nccl_version": (C.c_int, []),
    "vt_group_render": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vt_group_reset_accumulation": (C.c_int, [C.c_void_p]),
    "vt_group_sync": (C.c_int, [C.c_void_p]),
    "vt_group_begin_combine": (C.c_int, [C.c_void_p]),
    "vt_group_wait_combine": (C.c_int, [C.c_void_p]),
    "vt_group_end_combine": (C.c_int, [C.c_void_p, f32p]),
    "vt_group_read_average": (C.c_int, [C.c_void_p, f32p]),
    "vt_group_result_device_ptr": (C.c_void_p, [C.c_void_p]),
    "vt_group_last_exchange_ms": (C.c_int, [C.c_void_p, f32p]),
    "vt_group_exchange_bytes": (C.c_size_t, [C.c_void_p]),
    "vt_group_pick": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "vt_group_pick_focal": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "vt_group_add_voxel": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "vt_group_remove_voxel": (C.c_int, [C.c_void_p]),
    "vt_group_broadcast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
}

_LIB = None


def load():
    """Load libvoxeltoy_b200.so; raises if it has not been built (python -m voxeltoy_b200.build)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("voxeltoy_b200: %s is missing -- build it with `python -m voxeltoy_b200.build` "
                               "(there is no CPU or OpenGL fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


class VtError(RuntimeError):
    pass


def _fp(a):
    return a.ctypes.data_as(f32p)


def _ip(a):
    return a.ctypes.data_as(i32p)


class Context:
    """One GPU context (vt_ctx). Thin, stateless-in-Python wrapper: every method is one C-ABI call."""

    def __init__(self, device=0):
        self.lib = load()
        h = P()
        rc = self.lib.vt_create(device, C.byref(h))
        if rc != 0:
            raise VtError("vt_create(device=%d) failed with status %d: no usable CUDA device "
                          "(voxeltoy_b200 has no CPU fallback)" % (device, rc))
        self.h = h
        self.width = self.height = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.vt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise VtError("status %d: %s" % (rc, self.lib.vt_last_error(self.h).decode()))

    # ---- scene
    def volume_upload(self, grid, res):
        X, Y, Z = [int(v) for v in res]
        if grid is None:
            self._ck(self.lib.vt_volume_upload(self.h, None, X, Y, Z))
        else:
            g = np.ascontiguousarray(grid, np.int32)
            assert g.size == X * Y * Z
            self._ck(self.lib.vt_volume_upload(self.h, _ip(g), X, Y, Z))

    def materials_upload(self, data):
        d = np.ascontiguousarray(data, np.float32)
        self._ck(self.lib.vt_materials_upload(self.h, _fp(d), d.size))

    def material_update(self, offset, values):
        v = np.ascontiguousarray(values, np.float32)
        self._ck(self.lib.vt_material_update(self.h, offset, _fp(v), v.size))

    def emissive_upload(self, idx):
        e = np.ascontiguousarray(idx if idx is not None else [], np.int32)
        self._ck(self.lib.vt_emissive_upload(self.h, _ip(e) if e.size else None, e.size))

    def noise_upload(self, rgba=None, w=1024, h=1024):
        if rgba is None:
            self._ck(self.lib.vt_noise_upload(self.h, None, w, h))
        else:
            a = np.ascontiguousarray(rgba, np.float32)
            self._ck(self.lib.vt_noise_upload(self.h, _fp(a), a.shape[1], a.shape[0]))

    def env_upload(self, rgb, cdf_u, cdf_v, integral):
        rgb = np.ascontiguousarray(rgb, np.float32); cu = np.ascontiguousarray(cdf_u, np.float32)
        cv = np.ascontiguousarray(cdf_v, np.float32)
        self._ck(self.lib.vt_env_upload(self.h, _fp(rgb), rgb.shape[1], rgb.shape[0], _fp(cu), cu.shape[1], cu.shape[0],
                                        _fp(cv), cv.size, float(integral)))

    def env_build(self, rgb):
        """vt_env_build: RGB float image (h, w, 3) -> env texture + CDFs + integral, all on the device"""
        rgb = np.ascontiguousarray(rgb, np.float32)
        assert rgb.ndim == 3 and rgb.shape[2] == 3
        self._ck(self.lib.vt_env_build(self.h, _fp(rgb), rgb.shape[1], rgb.shape[0]))

    def env_info(self):
        dims = np.zeros(6, np.int32); integral = C.c_float(0); ms = C.c_float(0)
        self._ck(self.lib.vt_get_env_info(self.h, _ip(dims), C.cast(C.byref(integral), f32p), C.cast(C.byref(ms), f32p)))
        return dict(w=int(dims[0]), h=int(dims[1]), cdf_u_w=int(dims[2]), cdf_u_h=int(dims[3]), cdf_v_n=int(dims[4]), guided=bool(dims[5]),
                    integral=float(integral.value), build_ms=float(ms.value))

    def read_env_cdf(self):
        i = self.env_info()
        cu = np.empty((i["cdf_u_h"], i["cdf_u_w"]), np.float32); cv = np.empty(i["cdf_v_n"], np.float32)
        self._ck(self.lib.vt_read_env_cdf(self.h, _fp(cu), _fp(cv)))
        return cu, cv

    def env_clear(self):
        self._ck(self.lib.vt_env_clear(self.h))

    def read_volume(self):
        res, _, _, _ = self.volume_info()
        out = np.empty(int(res[0]) * int(res[1]) * int(res[2]), np.int32)
        self._ck(self.lib.vt_read_volume(self.h, _ip(out)))
        return out

    def read_materials(self, n):
        out = np.empty(n, np.float32)
        self._ck(self.lib.vt_read_materials(self.h, _fp(out), n))
        return out

    def volume_info(self):
        res = np.zeros(3, np.int32); a = np.zeros(3, np.float32); b = np.zeros(3, np.float32); v = np.zeros(3, np.float32)
        self._ck(self.lib.vt_get_volume_info(self.h, _ip(res), _fp(a), _fp(b), _fp(v)))
        return res, a, b, v

    # ---- frame state
    def set_camera(self, inv_modelview, proj, inv_proj, near_z=0.1, far_z=10000.0, lens_radius=0.0, lens_model=0):
        cam = VtCamera()
        for i in range(16):
            cam.inv_modelview[i] = float(np.asarray(inv_modelview, np.float32).reshape(-1)[i])
            cam.proj[i] = float(np.asarray(proj, np.float32).reshape(-1)[i])
            cam.inv_proj[i] = float(np.asarray(inv_proj, np.float32).reshape(-1)[i])
        cam.near_z, cam.far_z, cam.lens_radius, cam.lens_model = near_z, far_z, float(lens_radius), int(lens_model)
        self._ck(self.lib.vt_set_camera(self.h, C.byref(cam)))

    def set_settings(self, width, height, max_bounces=1, integrator=0, bg_top=None, bg_bottom=None, use_env_image=0,
                     env_rotation_rad=0.0, wireframe_opacity=0.0, wireframe_thickness=0.01):
        st = VtSettings()
        st.width, st.height, st.max_bounces, st.integrator = int(width), int(height), int(max_bounces), int(integrator)
        bt = bg_top if bg_top is not None else (153.0 / 255 * 2, 187.0 / 255 * 2, 201.0 / 255 * 2)
        bb = bg_bottom if bg_bottom is not None else (77.0 / 255, 64.0 / 255, 50.0 / 255)
        for i in range(3):
            st.bg_top[i] = float(np.float32(bt[i])); st.bg_bottom[i] = float(np.float32(bb[i]))
        st.use_env_image = int(use_env_image); st.env_rotation_rad = float(env_rotation_rad)
        st.wireframe_opacity = float(wireframe_opacity); st.wireframe_thickness = float(wireframe_thickness)
        self._ck(self.lib.vt_set_settings(self.h, C.byref(st)))
        self.width, self.height = int(width), int(height)

    def set_focal_distance(self, d):
        self._ck(self.lib.vt_set_focal_distance(self.h, float(d)))

    def get_focal_distance(self):
        d = C.c_float()
        self._ck(self.lib.vt_get_focal_distance(self.h, C.cast(C.byref(d), f32p)))
        return float(d.value)

    def set_selection(self, index, normal=None):
        i = np.zeros(4, np.int32); i[:len(index)] = index
        n = np.zeros(4, np.float32)
        if normal is not None:
            n[:len(normal)] = normal
        self._ck(self.lib.vt_set_selection(self.h, _ip(i), _fp(n) if normal is not None else None))

    def get_selection(self):
        i = np.zeros(4, np.int32); n = np.zeros(4, np.float32)
        self._ck(self.lib.vt_get_selection(self.h, _ip(i), _fp(n)))
        return i, n

    # ---- hot path
    def reset_accumulation(self):
        self._ck(self.lib.vt_reset_accumulation(self.h))

    def render(self, first_sample, n_passes):
        self._ck(self.lib.vt_render(self.h, int(first_sample), int(n_passes)))

    def sync(self):
        self._ck(self.lib.vt_sync(self.h))

    def num_samples(self):
        n = C.c_int()
        self._ck(self.lib.vt_get_num_samples(self.h, C.byref(n)))
        return n.value

    def read_average(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        self._ck(self.lib.vt_read_average(self.h, _fp(out)))
        return out

    def read_display(self, flip_vertical=False, out=None):
        """vt_read_display: the average as RGBA8 (GL float -> UNORM8 conversion), optionally in top-down row order"""
        if out is None:
            out = np.empty((self.height, self.width, 4), np.uint8)
        self._ck(self.lib.vt_read_display(self.h, out.ctypes.data_as(C.c_void_p), int(bool(flip_vertical))))
        return out

    def enable_primary_hits(self, on=True):
        self._ck(self.lib.vt_enable_primary_hits(self.h, int(on)))

    def read_primary_hits(self):
        out = np.empty((self.height, self.width), np.int32)
        self._ck(self.lib.vt_read_primary_hits(self.h, _ip(out)))
        return out

    def set_partition(self, mode, rank, world):
        self._ck(self.lib.vt_set_partition(self.h, mode, rank, world))

    def set_stream(self, cuda_stream):
        self._ck(self.lib.vt_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def accum_device_ptr(self):
        return self.lib.vt_accum_device_ptr(self.h)

    def set_kernel_variant(self, variant):
        self._ck(self.lib.vt_set_kernel_variant(self.h, int(variant)))

    KERNEL_KINDS = ("generate", "trace", "other", "shade", "accumulate")

    def kernel_timing_enable(self, on):
        self._ck(self.lib.vt_kernel_timing_enable(self.h, 1 if on else 0))

    def kernel_times(self):
        """{kind: (ms, launches)} of the wavefront kernels since the last call (device time, cudaEvent pairs)."""
        class KT(C.Structure):
            _fields_ = [("ms", C.c_float * 5), ("launches", C.c_uint32 * 5)]
        kt = KT()
        self._ck(self.lib.vt_get_kernel_times(self.h, C.byref(kt)))
        return {k: (float(kt.ms[i]), int(kt.launches[i])) for i, k in enumerate(self.KERNEL_KINDS)}

    def debug_advance(self, d, e, tau, nmax, literal=False):
        d = np.ascontiguousarray(d, np.float32); e = np.ascontiguousarray(e, np.float32)
        tau = np.ascontiguousarray(tau, np.float32); nmax = np.ascontiguousarray(nmax, np.int32)
        out = np.empty_like(d); k = np.empty_like(nmax)
        self._ck(self.lib.vt_debug_advance(self.h, _fp(d), _fp(e), _fp(tau), _ip(nmax), d.size, _fp(out), _ip(k), 1 if literal else 0))
        return out, k

    def debug_div_const(self, which):
        """(mismatches, first bad bit pattern) of gdiv_by vs div.rn over all 2^32 numerators; which = 0: PI, 1: 2 PI"""
        m = C.c_uint64(); f = C.c_uint32()
        self._ck(self.lib.vt_debug_div_const(self.h, int(which), C.byref(m), C.byref(f)))
        return int(m.value), int(f.value)

    def measure_l2_bandwidth(self, nbytes=48 << 20, reps=20):
        g = C.c_float()
        self._ck(self.lib.vt_measure_l2_bandwidth(self.h, int(nbytes), int(reps), C.cast(C.byref(g), f32p)))
        return float(g.value)

    def set_empty_skip(self, mode):
        self._ck(self.lib.vt_set_empty_skip(self.h, int(mode)))

    def set_wavefront_lanes(self, n):
        self._ck(self.lib.vt_set_wavefront_lanes(self.h, int(n)))

    def set_wavefront_max_paths(self, n):
        self._ck(self.lib.vt_set_wavefront_max_paths(self.h, int(n)))

    def counters_enable(self, on=True):
        self._ck(self.lib.vt_counters_enable(self.h, int(on)))

    def reset_counters(self):
        self._ck(self.lib.vt_reset_counters(self.h))

    def counters(self):
        c = VtCounters()
        self._ck(self.lib.vt_get_counters(self.h, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in VtCounters._fields_}

    # ---- voxelizer + services
    def voxelize(self, verts, indices, M, res, fill_offset=0):
        v = np.ascontiguousarray(verts, np.float32); i = np.ascontiguousarray(indices, np.uint32)
        m = np.ascontiguousarray(M, np.float32)
        self._ck(self.lib.vt_voxelize(self.h, _fp(v), v.size // 3, i.ctypes.data_as(u32p), i.size, _fp(m),
                                      int(res[0]), int(res[1]), int(res[2]), int(fill_offset)))

    def last_voxelize_ms(self):
        ms = C.c_float()
        self._ck(self.lib.vt_get_last_voxelize_ms(self.h, C.cast(C.byref(ms), f32p)))
        return float(ms.value)

    def set_voxelize_thickness(self, fat):
        """voxelize.gs:15-19 THICKNESS: False = THIN (the reference's build), True = FAT (conservative)."""
        self._ck(self.lib.vt_set_voxelize_thickness(self.h, 1 if fat else 0))

    def last_voxelize_full_ms(self):
        """Device time of the last vt_voxelize including the id grid and the distance field (scatter only: last_voxelize_ms)."""
        ms = C.c_float()
        self._ck(self.lib.vt_get_last_voxelize_full_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def assign_materials(self, table, rule=1):
        t = np.ascontiguousarray(table, np.int32)
        self._ck(self.lib.vt_volume_assign_materials(self.h, _ip(t), t.size, rule))

    def pick(self, px, py):
        self._ck(self.lib.vt_pick(self.h, float(px), float(py)))

    def pick_focal(self, px, py):
        self._ck(self.lib.vt_pick_focal(self.h, float(px), float(py)))

    def add_voxel(self, mx, my):
        self._ck(self.lib.vt_add_voxel(self.h, float(mx), float(my)))

    def remove_voxel(self):
        self._ck(self.lib.vt_remove_voxel(self.h))

    def trace_rays(self, rays):
        r = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        out = np.empty((r.shape[0], 4), np.float32)
        self._ck(self.lib.vt_debug_trace_rays(self.h, _fp(r), r.shape[0], _fp(out)))
        return out
