"""Python face of the C++ host classes (voxeltoy_b200/host/vt_host.h): Renderer, Camera, loaders, tools --
same names and call pattern as the reference's classes (renderer/renderer.h:23-122 etc.), one ctypes call each.
Nothing here computes: the C++ host issues C-ABI calls, the kernels do the work."""
import ctypes as C
import gzip
import os
import tempfile

import numpy as np

from . import _capi

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
P = C.c_void_p

# Action::PICKING_ACTION (renderer/actions.h:7-13)
PA_SELECT_FOCAL_POINT, PA_SELECT_ACTIVE_VOXEL, PA_ADD_VOXEL, PA_REMOVE_VOXEL = 0, 1, 2, 3
# CameraParameters::CameraLensModel (camera/cameraParameters.h:23-28)
CLM_PINHOLE, CLM_THIN_LENS, CLM_ORTHOGRAPHIC = 0, 1, 2
RR_SAMPLES_PENDING, RR_FINISHED_RENDERING = 0, 1
INTEGRATOR_PATHTRACER, INTEGRATOR_EDIT_MODE = 0, 1
LeftButton, RightButton, MiddleButton = 1, 2, 4
ControlModifier = 0x04000000
Key_Space, Key_A, Key_D, Key_F, Key_S, Key_W = 0x20, 0x41, 0x44, 0x46, 0x53, 0x57

_SIG = {
    "vth_renderer_create": (P, []), "vth_renderer_destroy": (None, [P]),
    "vth_renderer_initialize": (C.c_int, [P, C.c_int]),
    "vth_renderer_resize_frame": (None, [P] + [C.c_int] * 6),
    "vth_renderer_render": (C.c_int, [P]), "vth_renderer_render_passes": (C.c_int, [P, C.c_int]),
    "vth_renderer_reload_shaders": (None, [P, C.c_char_p]),
    "vth_renderer_load_vox_file": (None, [P, C.c_char_p]), "vth_renderer_load_mesh": (None, [P, C.c_char_p, C.c_int]),
    "vth_renderer_set_voxel_data": (None, [P, C.c_int, C.c_int, C.c_int, i32p, f32p, C.c_size_t, i32p, C.c_size_t]),
    "vth_renderer_save_image": (None, [P, C.c_char_p]), "vth_renderer_read_average": (C.c_int, [P, C.c_void_p]),
    "vth_renderer_reset_render": (None, [P]),
    "vth_renderer_on_mouse_move": (C.c_int, [P, C.c_int, C.c_int, C.c_int]), "vth_renderer_on_key_press": (C.c_int, [P, C.c_int]),
    "vth_renderer_request_action": (None, [P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]),
    "vth_renderer_update_render_settings": (None, [P]),
    "vth_renderer_status": (C.c_char_p, [P]), "vth_renderer_log": (C.c_char_p, [P]),
    "vth_renderer_context": (P, [P]), "vth_renderer_number_samples": (C.c_int, [P]),
    "vth_renderer_set_integrator": (None, [P, C.c_int]), "vth_renderer_set_partition": (None, [P, C.c_int, C.c_int, C.c_int]),
    "vth_renderer_camera_matrices": (None, [P, f32p, f32p, f32p]),
    "vth_renderer_volume_info": (None, [P, C.POINTER(C.c_int), f32p, f32p]),
    "vth_settings_set": (None, [P, C.c_int, C.c_int, C.c_float, C.c_float, f32p, f32p, C.c_char_p, C.c_int]),
    "vth_camera_set_lens_model": (None, [P, C.c_int]), "vth_camera_set_fstop": (None, [P, C.c_float]),
    "vth_camera_set_focal_length": (None, [P, C.c_float]), "vth_camera_set_lens_radius": (None, [P, C.c_float]),
    "vth_camera_set_controller": (None, [P, C.c_int]),
    "vth_camera_look_at": (None, [P, C.c_float, C.c_float, C.c_float]), "vth_camera_set_distance": (None, [P, C.c_float]),
    "vth_camera_orbit": (None, [P, C.c_float, C.c_float]), "vth_camera_get": (None, [P, f32p, f32p, f32p]),
    "vth_renderer_get_materials": (C.c_int, [P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vth_renderer_update_material_color": (None, [P, C.c_uint, f32p]),
    "vth_renderer_update_material_value": (None, [P, C.c_uint, C.c_float]),
    "vth_tool_create": (P, [P, C.c_int]), "vth_tool_destroy": (None, [P]),
    "vth_tool_mouse": (C.c_int, [P] + [C.c_int] * 7),
    "vth_vox_load": (P, [C.c_char_p]), "vth_vox_error": (C.c_char_p, [P]),
    "vth_vox_load_rules": (P, [C.c_char_p, f32p, C.c_int]), "vth_renderer_set_vox_palette_rules": (None, [P, f32p, C.c_int]),
    "vth_vox_dims": (None, [P, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "vth_vox_copy": (None, [P, i32p, f32p, i32p]), "vth_vox_free": (None, [P]),
    "vth_obj_load": (P, [C.c_char_p]), "vth_obj_dims": (None, [P, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "vth_obj_copy": (None, [P, f32p, u32p]), "vth_obj_free": (None, [P]),
    "vth_compute_mesh_transform": (None, [f32p, f32p, C.POINTER(C.c_int), f32p]),
    "vth_prune_emissive": (None, [i32p, C.c_int, C.c_int, C.c_int, i32p, C.POINTER(C.c_size_t)]),
    "vth_cdf_build": (P, [f32p, C.c_uint, C.c_uint]),
    "vth_cdf_dims": (None, [P, C.POINTER(C.c_uint), C.POINTER(C.c_uint), f32p]),
    "vth_cdf_copy": (None, [P, f32p, f32p]), "vth_cdf_free": (None, [P]),
    "vth_write_pfm": (C.c_int, [C.c_char_p, f32p, C.c_uint, C.c_uint]),
    "vth_renderer_resolution": (None, [P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vth_widget_create": (P, [P]), "vth_widget_destroy": (None, [P]),
    "vth_widget_run_script": (C.c_int, [P, C.c_char_p, C.c_char_p, C.c_int]), "vth_widget_pump": (C.c_int, [P, C.c_int]),
    "vth_widget_update_pending": (C.c_int, [P]), "vth_widget_paints": (C.c_ulong, [P]),
    "vth_write_png": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint, C.c_uint]),
    "vth_write_exr": (C.c_int, [C.c_char_p, f32p, C.c_uint, C.c_uint, C.c_int]),
    "vth_write_hdr": (C.c_int, [C.c_char_p, f32p, C.c_uint, C.c_uint, C.c_int]),
    "vth_load_image_dims": (C.c_int, [C.c_char_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    "vth_load_image": (C.c_int, [C.c_char_p, f32p]),
}
_bound = False


def lib():
    global _bound
    L = _capi.load()
    if not _bound:
        for name, (res, args) in _SIG.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _bound = True
    return L


def _fp(a):
    return a.ctypes.data_as(f32p)


def _ip(a):
    return a.ctypes.data_as(i32p)


_tmpdir = None


def plain_path(path):
    """The C++ loaders read plain files; the repo ships its two input assets gzip'd (tests/golden/*.gz)."""
    global _tmpdir
    if not path.endswith(".gz"):
        return path
    if _tmpdir is None:
        _tmpdir = tempfile.mkdtemp(prefix="voxeltoy_b200_")
    out = os.path.join(_tmpdir, os.path.basename(path)[:-3])
    if not os.path.exists(out):
        with gzip.open(path, "rb") as f, open(out, "wb") as g:
            g.write(f.read())
    return out


# ---- loaders ------------------------------------------------------------------------------------------------
MT_LAMBERT, MT_METAL, MT_PLASTIC = 0, 1, 2


def _rules_array(rules):
    """[(colour index, material type, (er, eg, eb), roughness), ...] -> the flat float array of the C wrappers."""
    a = np.zeros((len(rules), 6), np.float32)
    for i, (ci, mt, em, rough) in enumerate(rules):
        a[i] = (ci, mt, em[0], em[1], em[2], rough)
    return a


def load_vox(path, palette_rules=None):
    """MagicaVoxelLoader::load (with the new-build palette rules when given). Returns dict(res, grid, materials, emissive)."""
    L = lib()
    if palette_rules:
        ra = _rules_array(palette_rules)
        h = L.vth_vox_load_rules(plain_path(path).encode(), _fp(ra), len(palette_rules))
    else:
        h = L.vth_vox_load(plain_path(path).encode())
    try:
        res = (C.c_int * 3)(); nm = C.c_size_t(); ne = C.c_size_t()
        L.vth_vox_dims(h, res, C.byref(nm), C.byref(ne))
        if res[0] == 0:
            raise IOError("MagicaVoxelLoader: " + L.vth_vox_error(h).decode())
        grid = np.empty(res[0] * res[1] * res[2], np.int32); mats = np.empty(nm.value, np.float32)
        em = np.empty(ne.value, np.int32)
        L.vth_vox_copy(h, _ip(grid), _fp(mats), _ip(em))
    finally:
        L.vth_vox_free(h)
    return dict(res=(res[0], res[1], res[2]), grid=grid, materials=mats, emissive=em)


def load_obj(path):
    """MeshLoader::loadFromOBJ(path, vertices, indices)."""
    L = lib()
    h = L.vth_obj_load(plain_path(path).encode())
    try:
        nv = C.c_size_t(); ni = C.c_size_t()
        L.vth_obj_dims(h, C.byref(nv), C.byref(ni))
        v = np.empty((nv.value, 3), np.float32); i = np.empty(ni.value, np.uint32)
        L.vth_obj_copy(h, _fp(v), i.ctypes.data_as(u32p))
    finally:
        L.vth_obj_free(h)
    return v, i


def compute_mesh_transform(bmin, bmax, res):
    out = np.empty((4, 4), np.float32)
    a = np.ascontiguousarray(bmin, np.float32); b = np.ascontiguousarray(bmax, np.float32)
    r = (C.c_int * 3)(*[int(x) for x in res])
    lib().vth_compute_mesh_transform(_fp(a), _fp(b), r, _fp(out))
    return out


def prune_interior_emissive(grid, res, emissive):
    g = np.ascontiguousarray(grid, np.int32); e = np.ascontiguousarray(emissive, np.int32).copy()
    n = C.c_size_t(e.size)
    lib().vth_prune_emissive(_ip(g), int(res[0]), int(res[1]), int(res[2]), _ip(e), C.byref(n))
    return e[:n.value]


def calculate_cdf(rgb):
    """calculateCDF (renderer/image.cpp:68-283): returns dict(rgb, cdf_u, cdf_v, integral)."""
    L = lib()
    rgb = np.ascontiguousarray(rgb, np.float32)
    h = L.vth_cdf_build(_fp(rgb), rgb.shape[1], rgb.shape[0])
    try:
        w = C.c_uint(); hh = C.c_uint(); integral = C.c_float()
        L.vth_cdf_dims(h, C.byref(w), C.byref(hh), C.cast(C.byref(integral), f32p))
        if w.value == 0:
            raise ValueError("calculateCDF failed")
        cu = np.empty((hh.value, w.value), np.float32); cv = np.empty(hh.value + 1, np.float32)
        L.vth_cdf_copy(h, _fp(cu), _fp(cv))
    finally:
        L.vth_cdf_free(h)
    return dict(rgb=rgb, cdf_u=cu, cdf_v=cv, integral=float(integral.value))


def write_png(path, rgba8):
    """writePNG (host/image.cpp): (h, w, 4) uint8, rows top-down"""
    rgba8 = np.ascontiguousarray(rgba8, np.uint8)
    assert rgba8.ndim == 3 and rgba8.shape[2] == 4
    if lib().vth_write_png(path.encode(), rgba8.ctypes.data_as(C.c_void_p), rgba8.shape[1], rgba8.shape[0]) != 0:
        raise IOError("cannot write " + path)


def write_pfm(path, rgb):
    rgb = np.ascontiguousarray(rgb, np.float32)
    if lib().vth_write_pfm(path.encode(), _fp(rgb), rgb.shape[1], rgb.shape[0]) != 0:
        raise IOError("cannot write " + path)


def write_exr(path, pixels):
    """writeEXR (host/image_formats.cpp): (h, w, 3 or 4) float32, rows top-down -> FLOAT channels, ZIP compression"""
    pixels = np.ascontiguousarray(pixels, np.float32)
    assert pixels.ndim == 3 and pixels.shape[2] in (3, 4)
    if lib().vth_write_exr(path.encode(), _fp(pixels), pixels.shape[1], pixels.shape[0], pixels.shape[2]) != 0:
        raise IOError("cannot write " + path)


def write_hdr(path, pixels):
    """writeHDR (host/image.cpp): (h, w, 3 or 4) float32, rows top-down -> Radiance RGBE"""
    pixels = np.ascontiguousarray(pixels, np.float32)
    assert pixels.ndim == 3 and pixels.shape[2] in (3, 4)
    if lib().vth_write_hdr(path.encode(), _fp(pixels), pixels.shape[1], pixels.shape[0], pixels.shape[2]) != 0:
        raise IOError("cannot write " + path)


def load_image(path):
    """loadImage (renderer/image.cpp:28-59): .pfm, .hdr (RGBE), .exr and .png -> (h, w, 3) float32, row 0 = top"""
    w = C.c_uint(); h = C.c_uint()
    if lib().vth_load_image_dims(path.encode(), C.byref(w), C.byref(h)) != 0:
        raise IOError("cannot read " + path)
    out = np.empty((h.value, w.value, 3), np.float32)
    lib().vth_load_image(path.encode(), _fp(out))
    return out


# ---- Renderer -------------------------------------------------------------------------------------------------
class _CameraController:
    def __init__(self, r):
        self._r = r

    def lookAt(self, target):
        self._r._L.vth_camera_look_at(self._r._h, float(target[0]), float(target[1]), float(target[2]))

    def setDistanceFromTarget(self, d):
        self._r._L.vth_camera_set_distance(self._r._h, float(d))

    def orbitAroundTarget(self, theta, phi):
        self._r._L.vth_camera_orbit(self._r._h, float(theta), float(phi))


class _Camera:
    def __init__(self, r):
        self._r = r
        self._c = _CameraController(r)

    def controller(self):
        return self._c

    def setLensModel(self, m):
        self._r._L.vth_camera_set_lens_model(self._r._h, int(m))

    def setFStop(self, f):
        self._r._L.vth_camera_set_fstop(self._r._h, float(f))

    def setFocalLength(self, f):
        self._r._L.vth_camera_set_focal_length(self._r._h, float(f))

    def setLensRadius(self, f):
        self._r._L.vth_camera_set_lens_radius(self._r._h, float(f))

    def setCameraController(self, mode):
        self._r._L.vth_camera_set_controller(self._r._h, int(mode))

    def parameters(self):
        eye = np.zeros(3, np.float32); tgt = np.zeros(3, np.float32); s = np.zeros(8, np.float32)
        self._r._L.vth_camera_get(self._r._h, _fp(eye), _fp(tgt), _fp(s))
        return dict(eye=eye, target=tgt, fovY=float(s[0]), focalLength=float(s[1]), lensRadius=float(s[2]), near=float(s[3]),
                    far=float(s[4]), filmSize=(float(s[5]), float(s[6])), lensModel=int(s[7]))


class Renderer:
    """Mirror of the reference's Renderer (renderer/renderer.h:23-122)."""

    def __init__(self):
        self._L = lib()
        self._h = self._L.vth_renderer_create()
        self._cam = _Camera(self)
        self.width = self.height = 512

    def close(self):
        if self._h:
            self._L.vth_renderer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initialize(self, shaderPath="", device=0):
        if self._L.vth_renderer_initialize(self._h, device) != 0:
            raise _capi.VtError("Renderer.initialize: " + self.getStatus())

    def resizeFrame(self, w, h, vx=0, vy=0, vw=None, vh=None):
        self._L.vth_renderer_resize_frame(self._h, w, h, vx, vy, vw or w, vh or h)
        self.width, self.height = w, h

    def render(self):
        return self._L.vth_renderer_render(self._h)

    def renderPasses(self, n):
        return self._L.vth_renderer_render_passes(self._h, int(n))

    def reloadShaders(self, path=""):
        self._L.vth_renderer_reload_shaders(self._h, path.encode())

    def setVoxPaletteRules(self, rules):
        ra = _rules_array(rules or [])
        self._L.vth_renderer_set_vox_palette_rules(self._h, _fp(ra) if len(ra) else None, len(ra))

    def loadVoxFile(self, path):
        self._L.vth_renderer_load_vox_file(self._h, plain_path(path).encode())

    def loadMesh(self, path, resolution=0):
        self._L.vth_renderer_load_mesh(self._h, plain_path(path).encode(), int(resolution))

    def setVoxelData(self, res, grid, materials, emissive=None):
        g = np.ascontiguousarray(grid, np.int32); m = np.ascontiguousarray(materials, np.float32)
        e = np.ascontiguousarray(emissive if emissive is not None else [], np.int32)
        self._L.vth_renderer_set_voxel_data(self._h, int(res[0]), int(res[1]), int(res[2]), _ip(g), _fp(m), m.size,
                                            _ip(e) if e.size else None, e.size)

    def saveImage(self, path):
        self._L.vth_renderer_save_image(self._h, path.encode())

    def readAverage(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        ptr = out.ctypes.data if isinstance(out, np.ndarray) else out.data_ptr()
        if self._L.vth_renderer_read_average(self._h, ptr) != 0:
            raise _capi.VtError(self.getStatus())
        return out

    def resetRender(self):
        self._L.vth_renderer_reset_render(self._h)

    def onMouseMove(self, dx, dy, buttons):
        return bool(self._L.vth_renderer_on_mouse_move(self._h, dx, dy, buttons))

    def onKeyPress(self, key):
        return bool(self._L.vth_renderer_on_key_press(self._h, key))

    def camera(self):
        return self._cam

    def setRenderSettings(self, maxBounces=-1, maxSamples=-1, wireframeOpacity=-1.0, wireframeThickness=-1.0, backgroundTop=None,
                          backgroundBottom=None, backgroundImage=None, backgroundRotationDegrees=0):
        """renderer.renderSettings().m_* = ...; renderer.updateRenderSettings()."""
        t = np.ascontiguousarray(backgroundTop, np.float32) if backgroundTop is not None else None
        b = np.ascontiguousarray(backgroundBottom, np.float32) if backgroundBottom is not None else None
        self._L.vth_settings_set(self._h, maxBounces, maxSamples, wireframeOpacity, wireframeThickness,
                                 _fp(t) if t is not None else None, _fp(b) if b is not None else None,
                                 backgroundImage.encode() if backgroundImage is not None else None, backgroundRotationDegrees)
        self.updateRenderSettings()

    def updateRenderSettings(self):
        self._L.vth_renderer_update_render_settings(self._h)

    def getStatus(self):
        return self._L.vth_renderer_status(self._h).decode()

    def getLog(self):
        return self._L.vth_renderer_log(self._h).decode()

    def requestAction(self, x, y, dx, dy, action, restartAccumulation=True):
        self._L.vth_renderer_request_action(self._h, x, y, dx, dy, action, int(restartAccumulation))

    def getMaterials(self):
        types = (C.c_int * 4096)(); offs = (C.c_int * 4096)()
        n = self._L.vth_renderer_get_materials(self._h, 4096, types, offs)
        return [(types[i], offs[i]) for i in range(min(n, 4096))]

    def updateMaterialColor(self, offset, rgb):
        c = np.ascontiguousarray(rgb, np.float32)
        self._L.vth_renderer_update_material_color(self._h, offset, _fp(c))

    def updateMaterialValue(self, offset, v):
        self._L.vth_renderer_update_material_value(self._h, offset, float(v))

    # new-build accessors
    def renderSettingsResolution(self):
        w = C.c_int(); h = C.c_int()
        self._L.vth_renderer_resolution(self._h, C.byref(w), C.byref(h))
        return w.value, h.value

    def numberSamples(self):
        return self._L.vth_renderer_number_samples(self._h)

    def setIntegrator(self, i):
        self._L.vth_renderer_set_integrator(self._h, int(i))

    def setPartition(self, mode, rank, world):
        self._L.vth_renderer_set_partition(self._h, mode, rank, world)

    def cameraMatrices(self):
        a = np.empty((4, 4), np.float32); b = np.empty((4, 4), np.float32); c = np.empty((4, 4), np.float32)
        self._L.vth_renderer_camera_matrices(self._h, _fp(a), _fp(b), _fp(c))
        return a, b, c

    def volumeInfo(self):
        res = (C.c_int * 3)(); a = np.zeros(3, np.float32); b = np.zeros(3, np.float32)
        self._L.vth_renderer_volume_info(self._h, res, _fp(a), _fp(b))
        return (res[0], res[1], res[2]), a, b

    def context(self):
        """A _capi.Context view over the renderer's vt_ctx (not owning)."""
        ctx = _capi.Context.__new__(_capi.Context)
        ctx.lib = self._L
        ctx.h = P(self._L.vth_renderer_context(self._h))
        ctx.width, ctx.height = self.width, self.height
        ctx.close = lambda: None
        return ctx


class Tool:
    """ToolAddRemoveVoxel (kind 0) / ToolFocalDistance (kind 1), tools/*.cpp."""

    def __init__(self, renderer, kind):
        self._L = lib(); self._h = self._L.vth_tool_create(renderer._h, kind)

    def mousePressEvent(self, x, y, buttons, modifiers, w, h):
        return bool(self._L.vth_tool_mouse(self._h, 1, x, y, buttons, modifiers, w, h))

    def mouseMoveEvent(self, x, y, buttons, modifiers, w, h):
        return bool(self._L.vth_tool_mouse(self._h, 0, x, y, buttons, modifiers, w, h))

    def __del__(self):
        try:
            self._L.vth_tool_destroy(self._h)
        except Exception:
            pass


class HeadlessWidget:
    """The reference's GLWidget (ui/glwidget.cpp) without Qt: event routing tool -> renderer, UI slots, repaint while samples
    are pending. `run(script)` plays one command per line (host/headless.cpp lists them)."""

    def __init__(self, renderer):
        self._L = lib(); self._r = renderer; self._h = self._L.vth_widget_create(renderer._h)

    def run(self, script):
        # the repository keeps its scene fixtures gzip-ed: `vox` / `mesh` arguments go through plain_path like Renderer.loadVoxFile
        lines = []
        for line in script.split("\n"):
            w = line.split("#")[0].split()
            if len(w) >= 2 and w[0] in ("vox", "mesh"):
                line = "%s %s" % (w[0], plain_path(w[1]))
            lines.append(line)
        script = "\n".join(lines)
        err = C.create_string_buffer(512)
        if self._L.vth_widget_run_script(self._h, script.encode(), err, 512) != 0:
            raise ValueError(err.value.decode())
        res = self._r.renderSettingsResolution()
        self._r.width, self._r.height = res

    def pump(self, max_paints=1000000):
        return self._L.vth_widget_pump(self._h, int(max_paints))

    def updatePending(self):
        return bool(self._L.vth_widget_update_pending(self._h))

    def paints(self):
        return int(self._L.vth_widget_paints(self._h))

    def __del__(self):
        try:
            self._L.vth_widget_destroy(self._h)
        except Exception:
            pass


# ---- torch interop (plumbing only) ----------------------------------------------------------------------------------
class _DevPtr:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def device_view(ptr, shape):
    """Zero-copy torch view of a float32 device buffer owned by the library (for NCCL collectives)."""
    import torch
    return torch.as_tensor(_DevPtr(ptr, shape), device="cuda")
