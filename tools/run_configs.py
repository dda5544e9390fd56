"""Runs BASELINE configs 3, 4, 5 at FULL size on one GPU: timing plus the size-independent properties the domain offers
(idempotence of re-voxelization, tile partition == full frame bit for bit, accumulation reset after an edit, occupancy
popcount == solid material offsets). Usage: python tools/run_configs.py [c3] [c4] [c5]   (prints one JSON line per config)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import voxeltoy_b200 as vt
from voxeltoy_b200 import host, scenes


def timed(r, ctx, n, reps=3):
    """best of `reps` after one untimed run of the same length (the first one sizes the path-state pools)"""
    r.renderPasses(n); ctx.sync()
    best = 1e30
    for _ in range(reps):
        r.resetRender(); ctx.sync(); t = time.perf_counter(); r.renderPasses(n); ctx.sync(); best = min(best, time.perf_counter() - t)
    return best


def camera(r, W, H, theta, phi):
    r.resizeFrame(W, H)
    r.camera().controller().orbitAroundTarget(np.radians(theta), np.radians(phi))
    r.resetRender()
    ctx = r.context()                      # a fresh view: Context caches the frame size
    ctx.set_selection([-1, -1, -1, 0], [1, 0, 0, 0])
    return ctx


def c3():
    r = host.Renderer(); r.initialize("", 0); ctx = r.context()
    bunny = os.path.join("tests", "golden", "bunny.obj.gz")
    r.loadMesh(bunny, 512); ms1 = ctx.last_voxelize_ms(); a = ctx.read_volume()
    r.loadMesh(bunny, 512); ms2 = ctx.last_voxelize_ms(); b = ctx.read_volume()
    assert np.array_equal(a, b), "re-voxelization is not idempotent"
    t = scenes.c3_material_table()
    mats = t.array()
    ctx.materials_upload(mats)
    ctx.assign_materials(np.asarray(t.offsets, np.int32), rule=1)
    grid = ctx.read_volume()
    em = scenes.emissive_list(grid, mats)[::64]                       # thinned emissive set
    ctx.emissive_upload(em)
    r.setRenderSettings(maxBounces=4)
    ctx = camera(r, 1920, 1080, 130, 25)
    dt = timed(r, ctx, 16)
    ctx.kernel_timing_enable(True); ctx.kernel_times()
    r.resetRender(); r.renderPasses(16); ctx.sync()
    kt = {k: round(v[0], 2) for k, v in ctx.kernel_times().items()}; ctx.kernel_timing_enable(False)
    ctx.counters_enable(True); ctx.reset_counters(); r.renderPasses(1); cn = ctx.counters(); ctx.counters_enable(False)
    img = ctx.read_average()
    out = dict(config="C3 bunny 512^3 -> 1080p, 4 bounces, Lambert/metal/emissive", voxelize_ms=[ms1, ms2], solid=int((grid >= 0).sum()),
               emissive=int(em.size), msamples_per_s=1920 * 1080 * 16 / dt / 1e6, kernel_ms=kt,
               dda_steps_per_sample=cn["dda_steps"] / (1920 * 1080), nan_pixels=int(np.isnan(img).any(axis=2).sum()))
    r.close(); return out


def c4():
    n = 256
    ids = scenes.terrain_grid(n)
    t = scenes.MaterialTable()
    t.lambert((0.55, 0.5, 0.45)); t.metal((0.8, 0.8, 0.85), 60.0); t.lambert((0.3, 0.1, 0.05), emission=(6.0, 2.0, 0.5))
    grid = scenes.ids_to_offsets(ids, t.offsets); mats = t.array()
    em = host.prune_interior_emissive(grid, (n, n, n), scenes.emissive_list(grid, mats))
    r = host.Renderer(); r.initialize("", 0); ctx = r.context()
    r.setVoxelData((n, n, n), grid, mats, em)
    r.setRenderSettings(maxBounces=8)
    ctx = camera(r, 3840, 2160, 140, 35)
    r.renderPasses(1); ctx.sync(); r.resetRender()
    dt = timed(r, ctx, 8)
    full = ctx.read_average()
    acc = np.zeros_like(full)
    for rank in range(8):                                             # 8 "ranks" on one GPU: tiles must tile the frame exactly
        r.setPartition(vt.VT_PART_TILES, rank, 8); r.resetRender(); r.renderPasses(8); acc += ctx.read_average()
    r.setPartition(vt.VT_PART_NONE, 0, 1)
    same = bool(((acc.view(np.uint32) == full.view(np.uint32)) | (np.isnan(acc) & np.isnan(full))).all())
    ctx.kernel_timing_enable(True); ctx.kernel_times(); r.resetRender(); r.renderPasses(8); ctx.sync()
    kt = {k: round(v[0], 2) for k, v in ctx.kernel_times().items()}; ctx.kernel_timing_enable(False)
    out = dict(config="C4 terrain 256^3, 4K, 8 bounces, tiles", emissive=int(len(em)), msamples_per_s=3840 * 2160 * 8 / dt / 1e6, kernel_ms=kt,
               tiles_equal_full_bit_exact=same)
    r.close(); return out


def c5():
    n = 1024
    t = scenes.MaterialTable()
    for k in range(8):
        (t.metal((0.9, 0.6 + 0.04 * k, 0.3), 30.0 + 20 * k) if k % 3 == 2 else t.lambert((0.3 + 0.08 * k, 0.5, 0.9 - 0.08 * k)))
    t0 = time.perf_counter()
    grid = scenes.ids_to_offsets(scenes.dense_noise_grid(n, density=0.35), t.offsets)
    gen_s = time.perf_counter() - t0
    r = host.Renderer(); r.initialize("", 0); ctx = r.context()
    r.setVoxelData((n, n, n), grid, t.array(), None)
    r.setRenderSettings(maxBounces=16)
    ctx = camera(r, 3840, 2160, 125, 40)
    r.renderPasses(1); ctx.sync(); r.resetRender()
    dt = timed(r, ctx, 4)
    # interleaved edit: pick -> add resets the accumulation and changes exactly one voxel
    before = int(ctx.num_samples())
    r.requestAction(0.5, 0.5, 0.0, 0.0, host.PA_SELECT_ACTIVE_VOXEL); r.requestAction(0.5, 0.5, 0.0, 0.0, host.PA_ADD_VOXEL)
    r.renderPasses(1); ctx.sync()
    after = int(r.numberSamples())
    out = dict(config="C5 dense noise 1024^3 (35 %), 4K, 16 bounces", host_grid_gen_s=gen_s, msamples_per_s=3840 * 2160 * 4 / dt / 1e6,
               samples_before_edit=before, samples_after_edit=after)
    r.close(); return out


if __name__ == "__main__":
    for name in sys.argv[1:] or ["c3", "c4", "c5"]:
        print(json.dumps({"c3": c3, "c4": c4, "c5": c5}[name]()), flush=True)
