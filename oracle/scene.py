"""TEST INFRASTRUCTURE: numpy restatement of the reference's host-side producers of
kernel inputs (loaders, camera matrices, environment CDFs). Used by tests/ to feed
the CPU oracle independently of the product's C++ host classes, and to check those
classes. Never imported by the product.

Citations are relative to /root/reference/src.
"""
import gzip
import math
import os
import struct

import numpy as np

from . import vto

f32 = np.float32

# ---------------------------------------------------------------------------
# .vox loader  (renderer/loaders/magicaVoxel.cpp:122-321, voxLoader.cpp:8-66)
# ---------------------------------------------------------------------------

_DEFAULT_PALETTE = None


def default_palette():
    """magicaVoxel.cpp:245-263: the MagicaVoxel default palette. It is a closed-form
    table: a 6x6x6 colour cube (ff,cc,99,66,33,00) followed by three 10-step ramps
    of red, green and blue and a 10-step grey ramp; entry 0 is 0."""
    global _DEFAULT_PALETTE
    if _DEFAULT_PALETTE is not None:
        return _DEFAULT_PALETTE
    pal = [0x00000000]
    lv = [0xff, 0xcc, 0x99, 0x66, 0x33, 0x00]
    for r in lv:
        for g in lv:
            for b in lv:
                pal.append(0xff000000 | (b << 16) | (g << 8) | r)
    pal.pop()  # the cube's last entry (black) is not stored: 215 entries after index 0
    ramp = [0xee, 0xdd, 0xbb, 0xaa, 0x88, 0x77, 0x55, 0x44, 0x22, 0x11]
    for v in ramp:
        pal.append(0xff000000 | v)            # red ramp  (0xff0000ee ...)
    for v in ramp:
        pal.append(0xff000000 | (v << 8))     # green ramp
    for v in ramp:
        pal.append(0xff000000 | (v << 16))    # blue ramp
    for v in ramp:
        pal.append(0xff000000 | (v << 16) | (v << 8) | v)
    assert len(pal) == 256, len(pal)
    _DEFAULT_PALETTE = np.array(pal, dtype="<u4").view(np.uint8).reshape(256, 4).copy()
    return _DEFAULT_PALETTE


def read_bytes(path):
    if path.endswith(".gz"):
        with gzip.open(path, "rb") as f:
            return f.read()
    with open(path, "rb") as f:
        return f.read()


def parse_vox(data):
    """MV_Model::readModelFile, magicaVoxel.cpp:122-212. Returns (size xyz, voxels[n,4] u8, palette[256,4] u8 or None)."""
    magic, version = struct.unpack_from("<4si", data, 0)
    if magic != b"VOX ":
        raise ValueError("magic number does not match")
    if version != 150:
        raise ValueError("version does not match")
    cid, csize, chsize = struct.unpack_from("<4sii", data, 8)
    if cid != b"MAIN":
        raise ValueError("main chunk is not found")
    pos = 20 + csize
    end = 20 + csize + chsize
    size = (0, 0, 0)
    voxels = np.zeros((0, 4), np.uint8)
    palette = None
    while pos < end:
        sid, ssize, schild = struct.unpack_from("<4sii", data, pos)
        body = pos + 12
        send = body + ssize + schild
        if sid == b"SIZE":
            size = struct.unpack_from("<iii", data, body)
        elif sid == b"XYZI":
            n = struct.unpack_from("<i", data, body)[0]
            if n < 0:
                raise ValueError("negative number of voxels")
            voxels = np.frombuffer(data, np.uint8, n * 4, body + 4).reshape(n, 4).copy()
        elif sid == b"RGBA":
            palette = np.zeros((256, 4), np.uint8)
            palette[1:256] = np.frombuffer(data, np.uint8, 255 * 4, body).reshape(255, 4)
        pos = send
    return size, voxels, palette


def load_vox(path_or_bytes, palette_rules=None):
    """MagicaVoxelLoader::load (declared 5-arg contract).
    Returns dict(res=(X,Y,Z), grid int32[X*Y*Z] x-fastest, materials float32[], emissive int32[]).
    palette_rules (new-build extension, not in the reference): [(colour index, type 0/1/2, (er, eg, eb), roughness)] turns a
    palette entry into a Lambert / Metal / Plastic record with emission; records are [type][emission][colour]([roughness])."""
    rules = {int(r[0]): r for r in (palette_rules or [])}
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else read_bytes(path_or_bytes)
    (sx, sy, sz), voxels, palette = parse_vox(data)
    X, Y, Z = sx, sz, sy                                   # magicaVoxel.cpp:276-278 (y<->z)
    grid = np.full(X * Y * Z, -1, np.int32)
    pal = palette if palette is not None else default_palette()
    mat_off = [-1] * 256
    materials = []
    for x, y, z, ci in voxels.tolist():
        off = x + z * X + y * X * Y                        # :293-295
        if ci == 254:                                      # :297
            continue
        if mat_off[ci] < 0:                                # :300-318
            mat_off[ci] = len(materials)
            albedo = [f32(pal[ci, k]) / f32(255) for k in range(3)]
            if ci in rules:
                _, mt, em, rough = rules[ci]
                materials.extend([f32(mt)] + [f32(v) for v in em] + albedo + ([f32(rough)] if int(mt) in (1, 2) else []))
            else:
                materials.extend([f32(0), f32(0), f32(0), f32(0)] + albedo)   # [type=0][emission][albedo]
        grid[off] = mat_off[ci]
    materials = np.array(materials, np.float32)
    emissive = emissive_voxels(grid, materials)
    return dict(res=(X, Y, Z), grid=grid, materials=materials, emissive=emissive)


def emissive_voxels(grid, materials):
    """voxLoader.h:23-24 + voxLoader.cpp:68-91: voxels whose material has mean emission > 0, in index order."""
    solid = np.nonzero(grid >= 0)[0]
    if solid.size == 0 or materials.size == 0:
        return np.zeros(0, np.int32)
    offs = grid[solid]
    e = (materials[offs + 1] + materials[offs + 2] + materials[offs + 3]) / f32(3)
    return solid[e > 0].astype(np.int32)


def prune_interior_emissive(grid, res, emissive):
    """Renderer::pruneInteriorEmissiveVoxels, renderer/import.cpp:133-203 (swap-with-last removal)."""
    X, Y, Z = res
    em = list(int(v) for v in emissive)
    g = grid.reshape(Z, Y, X)

    def occupied(v, d):
        z = v // (X * Y); r = v - z * X * Y; y = r // X; x = r - y * X
        x += d[0]; y += d[1]; z += d[2]
        if x < 0 or x >= X or y < 0 or y >= Y or z < 0 or z >= Z:
            return False
        return g[z, y, x] >= 0

    nb = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    i = 0
    while i < len(em):
        if all(occupied(em[i], d) for d in nb):
            em[i] = em[-1]
            em.pop()
            continue
        i += 1
    return np.array(em, np.int32)


# ---------------------------------------------------------------------------
# OBJ loader (mesh/meshLoader.cpp:11-65 over tinyobjloader semantics:
# tiny_obj_loader.cc:100-106 float parse, :226-266 fan + first-use vertex order)
# ---------------------------------------------------------------------------

def load_obj(path_or_bytes):
    """MeshLoader::loadFromOBJ over tinyobj::LoadObj (tiny_obj_loader.cc:489-716, pinned against the reference's own copy by
    tests/test_oracle_vs_reference.py): faces are gathered into a face group; `g` / `o` (and the end of the file) export the
    group as one shape -- polygons as triangle fans, vertices in first-use order with a cache that lives for ONE shape
    (:226-275, the cache is passed by value) -- and `usemtl` DROPS the faces gathered so far (:617-623). meshLoader.cpp:40-63
    then concatenates the shapes, offsetting their indices."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else read_bytes(path_or_bytes)
    pos_in = []
    n_vt = n_vn = 0
    verts, idx = [], []
    group = []

    def fix(i, n):                                   # fixIndex: 1-based, negative = relative to the end, 0 -> 0
        return i - 1 if i > 0 else (0 if i == 0 else n + i)

    def export():
        cache = {}
        base = len(verts)
        for face in group:
            for k in range(2, len(face)):
                for key in (face[0], face[k - 1], face[k]):
                    if key not in cache:
                        cache[key] = len(verts) - base
                        verts.append(pos_in[key[0]])
                    idx.append(base + cache[key])
        del group[:]

    for line in data.decode("latin-1").splitlines():
        t = line.strip(" \t\r")
        if t.startswith("v ") or t.startswith("v\t"):
            p = t.split()
            pos_in.append((f32(float(p[1])), f32(float(p[2])), f32(float(p[3]))))
        elif t.startswith("vn ") or t.startswith("vn\t"):
            n_vn += 1
        elif t.startswith("vt ") or t.startswith("vt\t"):
            n_vt += 1
        elif t.startswith("f ") or t.startswith("f\t"):
            face = []
            for tok in t.split()[1:]:
                parts = tok.split("/")
                vi = fix(int(parts[0]), len(pos_in))
                vt = fix(int(parts[1]), n_vt) if len(parts) > 1 and parts[1] else -1
                vn = fix(int(parts[2]), n_vn) if len(parts) > 2 and parts[2] else -1
                face.append((vi, vt, vn))
            group.append(face)
        elif t.startswith("usemtl ") or t.startswith("usemtl\t"):
            del group[:]
        elif t[:2] in ("g ", "g\t", "o ", "o\t"):
            export()
    export()
    return np.array(verts, np.float32).reshape(-1, 3), np.array(idx, np.uint32)


def mesh_bounds(verts):
    return verts.min(axis=0).astype(np.float32), verts.max(axis=0).astype(np.float32)


def mesh_transform(bmin, bmax, res):
    """computeMeshTransform, renderer/import.cpp:46-64 (row-major, column-vector convention)."""
    res = [int(r) for r in res]
    margin = [f32(1.0) / f32(r) for r in res]
    size = [f32(bmax[i]) - f32(bmin[i]) for i in range(3)]
    major = 0                                             # Imath Box::majorAxis
    if size[1] > size[major]:
        major = 1
    if size[2] > size[major]:
        major = 2
    s = f32((1.0 - 2.0 * float(margin[major])) / float(size[major]))     # float - double*float -> double, / float -> float
    t = [f32(-f32(bmin[i]) + f32(margin[i] / s)) for i in range(3)]
    M = np.zeros((4, 4), np.float32)
    for i in range(3):
        M[i, i] = s
        M[i, 3] = f32(t[i] * s)
    M[3, 3] = 1
    return M


# ---------------------------------------------------------------------------
# camera (camera/cameraParameters.cpp, renderer/renderer.cpp:343-446)
# ---------------------------------------------------------------------------

def _norm(v):
    v = np.asarray(v, np.float32)
    l = f32(math.sqrt(float(f32(v[0] * v[0]) + f32(v[1] * v[1]) + f32(v[2] * v[2]))))
    return (v / l).astype(np.float32)


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], np.float32)


def invert4(m):
    """General 4x4 inverse in float64, rounded to float32 (matrices are kernel inputs; Imath's exact
    rounding is unpinned -- SURVEY 8c)."""
    return np.linalg.inv(np.asarray(m, np.float64)).astype(np.float32)


class Camera:
    """CameraParameters defaults (cameraParameters.cpp:7-17) + Renderer ctor (renderer.cpp:44-46)."""

    def __init__(self):
        self.target = np.zeros(3, np.float32)
        self.eye = np.array([0, 0, -1], np.float32)
        self.near = f32(0.1); self.far = f32(10000)
        self.focal_distance = f32(100)
        self.lens_radius = f32(0)
        self.film = np.array([36, 36], np.float32)
        self.fov_y = f32(0)
        self.lens_model = 0
        self.set_focal_length(50)

    # cameraParameters.cpp:135-175
    def focal_length(self):
        return f32(self.film[1] / f32(f32(2.0) * f32(math.tan(float(f32(0.5) * self.fov_y)))))

    def set_focal_length(self, fl):
        self.fov_y = f32(f32(math.atan2(float(self.film[1] * f32(0.5)), float(f32(fl)))) * f32(2.0))

    def set_film_size(self, w, h):
        self.film = np.array([max(0.0, w), max(0.0, h)], np.float32)
        fl = self.focal_length()
        self.fov_y = f32(f32(math.atan2(float(self.film[1] * f32(0.5)), float(fl))) * f32(2.0))

    def set_fstop(self, fstop):
        self.lens_radius = f32(f32(self.focal_length() / f32(max(1e-4, fstop))) * f32(0.5))

    def basis(self):
        fwd = _norm(self.target - self.eye)
        right = _norm(_cross(np.array([0, 1, 0], np.float32), fwd))
        up = _cross(fwd, right)
        return fwd, right, up

    def distance_to_target(self):
        d = self.target - self.eye
        return f32(math.sqrt(float(f32(d[0] * d[0]) + f32(d[1] * d[1]) + f32(d[2] * d[2]))))

    def set_distance_from_target(self, dist):
        fwd, _, _ = self.basis()
        self.eye = (self.target - fwd * f32(dist)).astype(np.float32)

    def orbit_around_target(self, theta, phi):
        r = self.distance_to_target()
        st = f32(math.sin(theta))
        d = np.array([st * f32(math.cos(phi)), f32(math.cos(theta)), st * f32(math.sin(phi))], np.float32)
        self.eye = (self.target - r * d).astype(np.float32)

    def matrices(self, W, H):
        """Renderer::render film-size rule (renderer.cpp:571-579) + updateCamera (:343-446).
        Returns row-major (mvm, inv_mvm, proj, inv_proj)."""
        a = f32(W) / f32(H)
        if a >= 1.0:
            self.set_film_size(36.0, float(f32(36.0) / a))
        else:
            self.set_film_size(float(f32(36.0) * a), 36.0)
        fwd, right, up = self.basis()
        eye = self.eye
        dot = lambda p, q: f32(f32(f32(p[0] * q[0]) + f32(p[1] * q[1])) + f32(p[2] * q[2]))
        mvm = np.eye(4, dtype=np.float32)
        mvm[0, :3] = right; mvm[0, 3] = -dot(eye, right)
        mvm[1, :3] = up; mvm[1, 3] = -dot(eye, up)
        mvm[2, :3] = -fwd; mvm[2, 3] = dot(eye, fwd)
        pm = np.zeros((4, 4), np.float32)
        if self.lens_model == 2:
            t = f32(math.tan(float(self.fov_y / f32(2))))
            left = f32(-t * self.distance_to_target()); rgt = -left
            bottom = f32(f32(-t * self.distance_to_target()) / a); top = -bottom
            near = -self.near; far = -self.far
            pm[0, 0] = f32(2.0) / (rgt - left); pm[0, 3] = -(rgt + left) / (rgt - left)
            pm[1, 1] = f32(2.0) / (top - bottom); pm[1, 3] = -(top + bottom) / (top - bottom)
            pm[2, 2] = f32(-2.0) / (far - near); pm[2, 3] = -(far + near) / (far - near)
            pm[3, 3] = 1
        else:
            n, f = self.near, self.far
            e = f32(f32(1.0) / f32(math.tan(float(self.fov_y / f32(2)))))
            pm[0, 0] = e / a; pm[1, 1] = e
            pm[2, 2] = (f + n) / (n - f); pm[2, 3] = f32(2.0) * f * n / (n - f)
            pm[3, 2] = -1
        return mvm, invert4(mvm), pm, invert4(pm)


# ---------------------------------------------------------------------------
# environment map -> CDFs (renderer/image.cpp:285-346 with the repo's own
# box-downscale / 3x3 gaussian standing in for OpenImageIO)
# ---------------------------------------------------------------------------
MAX_CDF_SIZE = 512


def env_function(rgb):
    """generateImageFunction: shrink to <=512 on the longest side (box filter), Rec709 luminance, 3x3 gaussian."""
    rgb = np.asarray(rgb, np.float32)
    h, w = rgb.shape[:2]
    m = max(w, h)
    if m > MAX_CDF_SIZE:
        nw = int(f32(w) / f32(m) * f32(MAX_CDF_SIZE)); nh = int(f32(h) / f32(m) * f32(MAX_CDF_SIZE))
        assert nw > 0 and nh > 0, "image too elongated for the CDF size limit"
        rgb = vto.resize_box(rgb, nw, nh)          # area-weighted box filter (integer factors: the plain box mean)
    lum = (rgb[..., 0] * f32(0.2126) + rgb[..., 1] * f32(0.7152)) + rgb[..., 2] * f32(0.0722)
    k1 = np.array([0.25, 0.5, 0.25], np.float32)
    p = np.pad(lum, 1, mode="edge")
    hp = (p[:, :-2] * k1[0] + p[:, 1:-1] * k1[1]) + p[:, 2:] * k1[2]
    out = (hp[:-2] * k1[0] + hp[1:-1] * k1[1]) + hp[2:] * k1[2]
    return np.ascontiguousarray(out, np.float32)


def build_env(rgb, rotation=0.0):
    lum = env_function(rgb)
    cu, cv, integral = vto.build_cdf(lum)
    return dict(rgb=np.ascontiguousarray(rgb, np.float32), cdf_u=cu, cdf_v=cv, integral=integral, rotation=rotation)
