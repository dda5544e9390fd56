"""Deterministic synthetic inputs for the BASELINE.json configurations that are not files of the
reference: the C2 HDR environment, the C3 material set, the C4 terrain and the C5 dense
noise grid. Pure numpy, no RNG state: every array is a closed-form function of its indices, so the
GPU path and the CPU oracle consume identical bytes.

Material records follow the reference layout (renderer/material/material.h:15-33,
loaders/voxLoader.cpp:54-66): [type][emission rgb][colour rgb]([roughness]).
"""
import numpy as np

f32 = np.float32

MT_LAMBERT, MT_METAL, MT_PLASTIC = 0, 1, 2


class MaterialTable:
    """Appends records the way VoxLoader::storeMaterialData does and remembers their offsets."""

    def __init__(self):
        self.data = []
        self.offsets = []

    def lambert(self, albedo, emission=(0, 0, 0)):
        self.offsets.append(len(self.data))
        self.data += [f32(MT_LAMBERT)] + [f32(v) for v in emission] + [f32(v) for v in albedo]
        return self.offsets[-1]

    def metal(self, reflectance, roughness, emission=(0, 0, 0)):
        self.offsets.append(len(self.data))
        self.data += [f32(MT_METAL)] + [f32(v) for v in emission] + [f32(v) for v in reflectance] + [f32(roughness)]
        return self.offsets[-1]

    def plastic(self, albedo, roughness, emission=(0, 0, 0)):
        self.offsets.append(len(self.data))
        self.data += [f32(MT_PLASTIC)] + [f32(v) for v in emission] + [f32(v) for v in albedo] + [f32(roughness)]
        return self.offsets[-1]

    def array(self):
        return np.array(self.data, np.float32)


def synthetic_env(w=1024, h=512):
    """C2 environment: lat-long RGB float map, row 0 = +Y pole,
    L = 0.2 + 0.6*max(0, d.y) sky plus a gaussian sun (sigma 3 deg, peak 500) at theta 45 deg, phi 60 deg.
    Direction convention of coordinates.h:97-116 (u = (atan(z,x)+pi)/2pi, v = acos(y)/pi)."""
    v = (np.arange(h, dtype=np.float64) + 0.5) / h
    u = (np.arange(w, dtype=np.float64) + 0.5) / w
    theta = v * np.pi
    phi = u * 2 * np.pi - np.pi
    st, ct = np.sin(theta)[:, None], np.cos(theta)[:, None]
    d = np.stack([st * np.cos(phi)[None, :], np.broadcast_to(ct, (h, w)), st * np.sin(phi)[None, :]], -1)
    sky = 0.2 + 0.6 * np.maximum(0.0, d[..., 1])
    ts, ps = np.radians(45.0), np.radians(60.0)
    sun_dir = np.array([np.sin(ts) * np.cos(ps), np.cos(ts), np.sin(ts) * np.sin(ps)])
    ang = np.arccos(np.clip(d @ sun_dir, -1, 1))
    sun = 500.0 * np.exp(-0.5 * (ang / np.radians(3.0)) ** 2)
    rgb = np.stack([sky * 0.9 + sun, sky * 1.0 + sun * 0.95, sky * 1.15 + sun * 0.8], -1)
    return np.ascontiguousarray(rgb, np.float32)


def c3_material_table():
    """C3: {Lambert albedo .8, Metal reflectance (.9,.7,.4) roughness 200, Lambert albedo .5 + emission (4,3,2)}."""
    t = MaterialTable()
    t.lambert((0.8, 0.8, 0.8))
    t.metal((0.9, 0.7, 0.4), 200.0)
    t.lambert((0.5, 0.5, 0.5), emission=(4.0, 3.0, 2.0))
    return t


def c3_assign(grid, res, table_offsets):
    """Material rule of C3: id = ((x>>5) ^ (y>>5) ^ (z>>5)) % 3 for solid voxels; the emissive material (id 2) is kept
    only on 1/64 of the 32^3 bricks (brick hash below), elsewhere it falls back to id 0."""
    X, Y, Z = res
    g = grid.reshape(Z, Y, X)
    z, y, x = np.nonzero(g >= 0)
    bid = (x >> 5) ^ (y >> 5) ^ (z >> 5)
    ids = bid % 3
    h = ((x >> 5) * 73856093) ^ ((y >> 5) * 19349663) ^ ((z >> 5) * 83492791)
    keep = (h & 63) == 0
    ids = np.where((ids == 2) & ~keep, 0, ids)
    out = grid.copy().reshape(Z, Y, X)
    out[z, y, x] = np.asarray(table_offsets, np.int32)[ids]
    return out.reshape(-1)


def emissive_list(grid, materials):
    """voxLoader.h:23-24: linear indices of the voxels whose material has mean emission > 0, ascending."""
    solid = np.nonzero(grid >= 0)[0]
    if solid.size == 0:
        return np.zeros(0, np.int32)
    offs = grid[solid]
    e = (materials[offs + 1] + materials[offs + 2] + materials[offs + 3]) / f32(3)
    return solid[e > 0].astype(np.int32)


def pcg_hash(v):
    v = (v.astype(np.uint64) * np.uint64(747796405) + np.uint64(2891336453)) & np.uint64(0xFFFFFFFF)
    w = (((v >> ((v >> np.uint64(28)) + np.uint64(4))) ^ v) * np.uint64(277803737)) & np.uint64(0xFFFFFFFF)
    return ((w >> np.uint64(22)) ^ w) & np.uint64(0xFFFFFFFF)


def dense_noise_grid(n, density=0.35, seed=1, n_materials=8, chunk=64):
    """C5: solid iff pcg_hash(x,y,z,seed) < density*2^32, material id = (hash>>8) & (n_materials-1).
    Returns int32 grid of material IDS (caller maps ids to offsets)."""
    out = np.empty((n, n, n), np.int32)
    x = np.arange(n, dtype=np.uint64)[None, None, :]
    y = np.arange(n, dtype=np.uint64)[None, :, None]
    thr = np.uint64(int(density * 2 ** 32))
    for z0 in range(0, n, chunk):
        z = np.arange(z0, min(n, z0 + chunk), dtype=np.uint64)[:, None, None]
        key = (x + y * np.uint64(n) + z * np.uint64(n * n) + np.uint64(seed) * np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF)
        hsh = pcg_hash(key)
        ids = ((hsh >> np.uint64(8)) & np.uint64(n_materials - 1)).astype(np.int32)
        out[z0:z0 + chunk] = np.where(hsh < thr, ids, -1)
    return out.reshape(-1)


def dense_noise_offsets_torch(n, offsets, density=0.35, seed=1, n_materials=8, chunk=64, device="cuda"):
    """dense_noise_grid + ids_to_offsets evaluated with torch on the GPU (the numpy generator needs 40 s at 1024^3); the same
    integers (checked against the numpy version by tests/test_gpu_fullsize.py). Returns the int32 R32I grid on the host."""
    import torch
    dev = torch.device(device)
    M = 0xFFFFFFFF
    out = torch.empty((n, n, n), dtype=torch.int32)
    x = torch.arange(n, dtype=torch.int64, device=dev)[None, None, :]
    y = torch.arange(n, dtype=torch.int64, device=dev)[None, :, None]
    offs = torch.as_tensor(np.asarray(offsets, np.int64), device=dev)
    thr = int(density * 2 ** 32)
    for z0 in range(0, n, chunk):
        z = torch.arange(z0, min(n, z0 + chunk), dtype=torch.int64, device=dev)[:, None, None]
        v = (x + y * n + z * (n * n) + seed * 0x9E3779B9) & M
        v = (v * 747796405 + 2891336453) & M                                   # pcg_hash
        w = ((torch.bitwise_right_shift(v, (v >> 28) + 4) ^ v) * 277803737) & M
        h = ((w >> 22) ^ w) & M
        ids = (h >> 8) & (n_materials - 1)
        out[z0:z0 + chunk] = torch.where(h < thr, offs[ids], torch.full_like(ids, -1)).to(torch.int32).cpu()
    return out.reshape(-1).numpy()


def c4_scene(n=256):
    """BASELINE config 4: procedural terrain, rock / metal band / emissive lava. dict(res, grid, materials, emissive_all)."""
    ids = terrain_grid(n)
    t = MaterialTable()
    t.lambert((0.55, 0.5, 0.45)); t.metal((0.8, 0.8, 0.85), 60.0); t.lambert((0.3, 0.1, 0.05), emission=(6.0, 2.0, 0.5))
    grid = ids_to_offsets(ids, t.offsets); mats = t.array()
    return dict(res=(n, n, n), grid=grid, materials=mats, emissive_all=emissive_list(grid, mats))


def c5_material_table():
    t = MaterialTable()
    for k in range(8):
        (t.metal((0.9, 0.6 + 0.04 * k, 0.3), 30.0 + 20 * k) if k % 3 == 2 else t.lambert((0.3 + 0.08 * k, 0.5, 0.9 - 0.08 * k)))
    return t


def _value_noise2(x, z, seed):
    """Smooth deterministic 2-D value noise in [-1, 1] (bicubic-smoothstep interpolation of lattice hashes)."""
    xi, zi = np.floor(x).astype(np.int64), np.floor(z).astype(np.int64)
    fx, fz = x - xi, z - zi
    sx, sz = fx * fx * (3 - 2 * fx), fz * fz * (3 - 2 * fz)

    def lat(ix, iz):
        k = (ix.astype(np.uint64) * np.uint64(0x9E3779B1) + iz.astype(np.uint64) * np.uint64(0x85EBCA77) + np.uint64(seed)) & np.uint64(0xFFFFFFFF)
        return pcg_hash(k).astype(np.float64) / 2 ** 31 - 1.0

    a = lat(xi, zi) * (1 - sx) + lat(xi + 1, zi) * sx
    b = lat(xi, zi + 1) * (1 - sx) + lat(xi + 1, zi + 1) * sx
    return a * (1 - sz) + b * sz


def terrain_grid(n=256, octaves=6, persistence=0.5, seed=7):
    """C4: heightfield terrain, solid iff y < 3n/8 + n/4 * fbm(x/n, z/n).
    Returns (ids int32 [n^3] with -1 empty, 0 rock, 1 metal band, 2 lava), fbm in [-1, 1]."""
    xs = np.arange(n, dtype=np.float64)
    X, Zc = np.meshgrid(xs, xs, indexing="xy")           # [z, x]
    fbm = np.zeros((n, n)); amp, freq, norm = 1.0, 4.0, 0.0
    for o in range(octaves):
        fbm += amp * _value_noise2(X / n * freq, Zc / n * freq, seed + o)
        norm += amp; amp *= persistence; freq *= 2.0
    fbm /= norm
    height = np.clip(np.floor(3 * n / 8 + n / 4 * fbm), 1, n - 1).astype(np.int32)   # [z, x]
    y = np.arange(n, dtype=np.int32)[None, :, None]
    h = height[:, None, :]
    solid = y < h
    ids = np.full((n, n, n), -1, np.int32)
    ids[solid] = 0
    band = solid & ((y // max(1, n // 16)) % 4 == 3)
    ids[band] = 1
    lava = solid & (y == h - 1) & (fbm[:, None, :] > 0.35)
    ids[lava] = 2
    return ids.reshape(-1)


def ids_to_offsets(ids, offsets):
    offs = np.asarray(offsets, np.int32)
    out = np.full(ids.shape, -1, np.int32)
    m = ids >= 0
    out[m] = offs[ids[m]]
    return out
