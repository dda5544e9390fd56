"""Shared scene builders for the tests: one plain dict feeds both the CPU oracle (oracle.vto.make_scene)
and the CUDA context (upload())."""
import os

import numpy as np

from oracle import scene as oscene
from oracle import vto

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENE_FALL = os.path.join(GOLDEN, "scene_fall.vox.gz")
BUNNY = os.path.join(GOLDEN, "bunny.obj.gz")

GRADIENT_TOP = [np.float32(153.0 / 255 * 2), np.float32(187.0 / 255 * 2), np.float32(201.0 / 255 * 2)]
GRADIENT_BOTTOM = [np.float32(77.0 / 255), np.float32(64.0 / 255), np.float32(50.0 / 255)]
GREY = [np.float32(192.0 / 255)] * 3

_cache = {}


def scene_fall_volume():
    if "fall" not in _cache:
        _cache["fall"] = oscene.load_vox(SCENE_FALL)
    return dict(_cache["fall"])


def camera_for(res, W, H, theta_deg=None, phi_deg=None, lens_model=0, fstop=16.0, distance=None):
    """Renderer ctor + loadVoxFile camera (renderer.cpp:44-46, import.cpp:41), optionally orbited."""
    bmin, bmax, _ = vto.volume_bounds(*res)
    cam = oscene.Camera()
    cam.set_distance_from_target(100.0)
    cam.lens_model = lens_model
    diag = np.sqrt(np.sum((bmax - bmin).astype(np.float64) ** 2))
    cam.set_distance_from_target(np.float32(diag) * np.float32(0.5) if distance is None else distance)
    if theta_deg is not None:
        cam.orbit_around_target(np.radians(theta_deg), np.radians(phi_deg))
    mvm, imv, pm, ipm = cam.matrices(W, H)      # sets the film size for this aspect ratio
    cam.set_fstop(fstop)
    return cam, imv, pm, ipm


def make_frame(vol, W, H, bounces=1, theta=None, phi=None, lens_model=0, fstop=16.0, env=None, bg="gradient",
               focal_distance=99999999.0, wire_opacity=0.0, sel=(-1, -1, -1), distance=None):
    d = dict(vol)
    cam, imv, pm, ipm = camera_for(vol["res"], W, H, theta, phi, lens_model, fstop, distance)
    d.update(W=W, H=H, inv_modelview=imv, proj=pm, inv_proj=ipm, max_bounces=bounces, lens_model=lens_model,
             lens_radius=float(cam.lens_radius), focal_distance=focal_distance, wire_opacity=wire_opacity,
             wire_thickness=0.01, sel_index=sel, near_z=float(cam.near))
    if bg == "gradient":
        d.update(bg_top=GRADIENT_TOP, bg_bottom=GRADIENT_BOTTOM)
    else:
        d.update(bg_top=GREY, bg_bottom=GREY)
    if env is not None:
        d["env"] = env
    return d


def upload(ctx, d, integrator=0):
    """Push a scene dict through the C ABI."""
    ctx.volume_upload(d["grid"], d["res"])
    ctx.materials_upload(d["materials"])
    ctx.emissive_upload(d.get("emissive"))
    env = d.get("env")
    if env is not None:
        ctx.env_upload(env["rgb"], env["cdf_u"], env["cdf_v"], env["integral"])
    else:
        ctx.env_clear()
    ctx.set_camera(d["inv_modelview"], d["proj"], d["inv_proj"], near_z=d.get("near_z", 0.1),
                   lens_radius=d.get("lens_radius", 0.0), lens_model=d.get("lens_model", 0))
    ctx.set_settings(d["W"], d["H"], max_bounces=d["max_bounces"], integrator=integrator,
                     bg_top=d["bg_top"], bg_bottom=d["bg_bottom"], use_env_image=1 if env is not None else 0,
                     env_rotation_rad=(env or {}).get("rotation", 0.0),
                     wireframe_opacity=d.get("wire_opacity", 0.0), wireframe_thickness=d.get("wire_thickness", 0.01))
    ctx.set_focal_distance(d.get("focal_distance", 99999999.0))
    sel = d.get("sel_index", (-1, -1, -1))
    ctx.set_selection([sel[0], sel[1], sel[2], 0], [1, 0, 0, 0])
    ctx.reset_accumulation()


def same_bits(a, b):
    """Bit-exact float comparison that treats any NaN as equal to any NaN."""
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    eq = (a.view(np.uint32) == b.view(np.uint32)) | (na & nb) | ((a == 0) & (b == 0))
    return eq


def mixed_scene(n=48, seed=3):
    """Small synthetic volume with all three material types + emissive voxels + a ground gap."""
    from voxeltoy_b200 import scenes
    rng = np.random.RandomState(seed)
    t = scenes.MaterialTable()
    t.lambert((0.8, 0.7, 0.6)); t.metal((0.9, 0.7, 0.4), 200.0); t.plastic((0.2, 0.4, 0.8), 40.0)
    t.lambert((0.5, 0.5, 0.5), emission=(4.0, 3.0, 2.0)); t.metal((0.95, 0.95, 0.95), 5.0)
    ids = np.full((n, n, n), -1, np.int32)          # [z, y, x]
    ids[:, : n // 6, :] = 0
    for _ in range(40):
        c = rng.randint(2, n - 6, size=3); s = rng.randint(2, 6, size=3)
        ids[c[0]:c[0] + s[0], c[1]:c[1] + s[1], c[2]:c[2] + s[2]] = rng.randint(0, 5)
    grid = scenes.ids_to_offsets(ids.reshape(-1), t.offsets)
    mats = t.array()
    em = scenes.emissive_list(grid, mats)
    em = oscene.prune_interior_emissive(grid, (n, n, n), em)
    return dict(res=(n, n, n), grid=grid, materials=mats, emissive=em)
