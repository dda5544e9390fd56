"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Bar (SURVEY 8 / BASELINE north star): primary-hit indices and occupancy bit-exact; radiance is held to
the same bar here because the product fixes its arithmetic contract (csrc/vt_math.cuh) -- any
difference is reported as max |delta| and must be exactly 0 (NaN == NaN)."""
import numpy as np
import pytest

from oracle import scene as oscene
from oracle import vto
from tests import util

pytestmark = pytest.mark.gpu


def _compare(ctx, d, n_passes=1, first=0, integrator=0):
    s = vto.make_scene(d)
    util.upload(ctx, d, integrator=integrator)
    ctx.enable_primary_hits(True)
    ctx.render(first, n_passes)
    got = ctx.read_average()
    hits = ctx.read_primary_hits()
    if integrator == 0:
        ref = vto.render_average(s, n_passes, first=first)
        _, ref_hits, _, _ = vto.render_pass(s, first + n_passes - 1)
    else:
        ref = np.zeros((d["H"], d["W"], 4), np.float32)
        for n in range(n_passes):
            vto.accumulate(ref, vto.preview_pass(s, first + n), n)
        ref_hits = None
    # the wavefront renderer (variant 2, default; also with a path budget that splits the passes into several
    # batches and with several batches in flight) and the one-thread-per-pixel megakernel (0) must produce the same bits
    for variant, budget, lanes in ((0, 0, 1), (2, 2 * 4096, 1), (2, 4 * 4096, 2), (2, 0, 3)):
        ctx.set_kernel_variant(variant)
        ctx.set_wavefront_lanes(lanes)
        if budget:
            ctx.set_wavefront_max_paths(budget)
        ctx.reset_accumulation()
        ctx.render(first, n_passes)
        got_v = ctx.read_average()
        hits_v = ctx.read_primary_hits()
        ctx.set_kernel_variant(2)
        ctx.set_wavefront_lanes(1)
        ctx.set_wavefront_max_paths(32 << 20)
        assert util.same_bits(got, got_v).all() and np.array_equal(hits, hits_v), "kernel variant %d disagrees" % variant
    eq = util.same_bits(got, ref)
    bad = int((~eq).sum())
    if bad:
        diff = np.nanmax(np.abs(got - ref))
        raise AssertionError("%d / %d floats differ, max |delta| = %g" % (bad, eq.size, diff))
    if ref_hits is not None:
        assert np.array_equal(hits, ref_hits), "primary hit indices differ in %d pixels" % int((hits != ref_hits).sum())
    return got, hits


def test_c1_scene_fall_primary_hits_and_radiance(vt_ctx):
    """BASELINE config 1: scene_fall 512x512, 1 spp, 1 bounce, pinhole, constant environment."""
    d = util.make_frame(util.scene_fall_volume(), 512, 512, bounces=1, bg="grey")
    got, hits = _compare(vt_ctx, d)
    assert (hits >= 0).mean() > 0.3 and not np.isnan(got).any()


def test_scene_fall_orbit_4_bounces_multi_pass(vt_ctx):
    d = util.make_frame(util.scene_fall_volume(), 320, 180, bounces=4, theta=120, phi=30)
    _compare(vt_ctx, d, n_passes=5)


def test_scene_fall_ibl_thin_lens(vt_ctx):
    """C2's feature set at reduced size: importance-sampled IBL + thin-lens DOF + 4 bounces."""
    from voxeltoy_b200 import scenes
    env = oscene.build_env(scenes.synthetic_env(256, 128))
    vol = util.scene_fall_volume()
    d = util.make_frame(vol, 256, 144, bounces=4, theta=120, phi=30, lens_model=1, fstop=2.8, env=env)
    # focal distance through the service, on both sides
    s = vto.make_scene(d)
    util.upload(vt_ctx, d)
    vt_ctx.pick_focal(128.0, 72.0)
    fd = vt_ctx.get_focal_distance()
    assert fd == vto.pick_focal(s, 128.0, 72.0)
    d["focal_distance"] = fd
    _compare(vt_ctx, d, n_passes=3)


def test_mixed_materials_emissive(vt_ctx):
    """Lambert + metal + plastic + emissive voxels (light sampling of emissive voxels, MIS, microfacet NaN paths)."""
    d = util.make_frame(util.mixed_scene(), 200, 160, bounces=5, theta=115, phi=40)
    _compare(vt_ctx, d, n_passes=4, first=7)


def test_orthographic_and_wireframe_and_selection(vt_ctx):
    vol = util.scene_fall_volume()
    d = util.make_frame(vol, 160, 120, bounces=2, theta=125, phi=-50, lens_model=2, wire_opacity=0.7, sel=(60, 1, 60))
    _compare(vt_ctx, d, n_passes=2)
    d = util.make_frame(vol, 160, 120, bounces=2, theta=125, phi=-50, wire_opacity=0.5, sel=(0, 0, 0))
    _compare(vt_ctx, d, n_passes=2)


def test_edit_mode_preview(vt_ctx):
    d = util.make_frame(util.scene_fall_volume(), 256, 192, bounces=1, theta=120, phi=30, wire_opacity=0.3)
    _compare(vt_ctx, d, n_passes=2, integrator=1)


def test_camera_inside_volume_and_miss(vt_ctx):
    vol = util.scene_fall_volume()
    d = util.make_frame(vol, 128, 96, bounces=3, theta=100, phi=10, distance=120.0)     # eye inside the bounds
    _compare(vt_ctx, d)
    d = util.make_frame(vol, 128, 96, bounces=3, theta=60, phi=30)                       # from below: ground slab only
    _compare(vt_ctx, d)


def test_dda_random_rays_including_ties_and_nan(vt_ctx):
    """The DDA alone on random, axis-aligned, diagonal (tie) and non-finite rays."""
    rng = np.random.RandomState(11)
    vol = util.mixed_scene(40, seed=5)
    d = util.make_frame(vol, 8, 8)
    s = vto.make_scene(d)
    util.upload(vt_ctx, d)
    n = 20000
    bmin, bmax, vs = vto.volume_bounds(*vol["res"])
    o = rng.uniform(bmin * 1.3, bmax * 1.3, size=(n, 3)).astype(np.float32)
    dr = rng.normal(size=(n, 3)).astype(np.float32)
    dr /= np.linalg.norm(dr, axis=1, keepdims=True)
    # exact diagonals from voxel corners (dis ties), axis-aligned rays, zero components
    k = n // 5
    o[:k] = (bmin + vs * rng.randint(0, 40, size=(k, 3))).astype(np.float32)
    dr[:k] = rng.choice([-1.0, 1.0], size=(k, 3)).astype(np.float32) * np.float32(0.57735026919)
    dr[k:2 * k] = np.eye(3, dtype=np.float32)[rng.randint(0, 3, size=k)] * rng.choice([-1.0, 1.0], size=(k, 1)).astype(np.float32)
    dr[2 * k:2 * k + 50, 1] = np.nan
    o[2 * k + 50:2 * k + 100, 0] = np.nan
    dr[2 * k + 100:2 * k + 150, 2] = np.inf
    rays = np.concatenate([o, dr], axis=1)
    got = vt_ctx.trace_rays(rays)
    ref = vto.trace_rays(s, rays)
    assert util.same_bits(got, ref).all()
    assert (ref[:, 3] == 1).sum() > 1000 and (ref[:, 3] == 0).sum() > 1000


@pytest.mark.parametrize("res", [64, 128])
def test_voxelizer_bunny_occupancy_bit_exact(vt_ctx, res):
    verts, idx = oscene.load_obj(util.BUNNY)
    assert verts.shape[0] == 2503 and idx.size == 4968 * 3
    bmin, bmax = oscene.mesh_bounds(verts)
    M = oscene.mesh_transform(bmin, bmax, (res, res, res))
    ref = vto.voxelize(verts, idx, M, (res, res, res))
    vt_ctx.voxelize(verts, idx, M, (res, res, res), fill_offset=0)
    got = vt_ctx.read_volume()
    assert np.array_equal(got >= 0, ref > 0)
    assert set(np.unique(got)) <= {-1, 0}
    assert ref.sum() > 1000


def test_voxelizer_non_cubic_and_degenerate(vt_ctx):
    rng = np.random.RandomState(2)
    verts = rng.uniform(0.05, 0.95, size=(300, 3)).astype(np.float32)
    idx = rng.randint(0, 300, size=(400, 3)).astype(np.uint32)
    idx[:10, 1] = idx[:10, 0]                     # degenerate triangles (nzInv = inf, voxelize.gs:175)
    verts[:20, 2] = 0.5                           # axis-aligned faces
    M = np.eye(4, dtype=np.float32)
    for res in [(48, 32, 40), (33, 33, 33)]:
        ref = vto.voxelize(verts, idx, M, res)
        vt_ctx.voxelize(verts, idx, M, res, fill_offset=7)
        got = vt_ctx.read_volume()
        assert np.array_equal(got >= 0, ref > 0)


def test_services_pick_add_remove(vt_ctx):
    vol = util.scene_fall_volume()
    d = util.make_frame(vol, 320, 240, bounces=1, theta=120, phi=30)
    s = vto.make_scene(d)
    util.upload(vt_ctx, d)
    grid = vol["grid"].copy()
    rng = np.random.RandomState(4)
    rn = np.array([1, 0, 0, 0], np.float32)
    for i in range(24):
        px, py = float(rng.uniform(0, 320)), float(rng.uniform(0, 240))
        vt_ctx.pick(px, py)
        gi, gn = vt_ctx.get_selection()
        ri, rn = vto.pick(s, px, py, near_z=d["near_z"], prev_normal=rn)
        assert np.array_equal(gi, ri) and np.array_equal(gn, rn), (px, py, gi, ri, gn, rn)
        vt_ctx.pick_focal(px, py)
        assert vt_ctx.get_focal_distance() == vto.pick_focal(s, px, py)
        if i % 3 == 0:
            mx, my = (0.0, 0.0) if i % 2 else (float(rng.normal()), float(rng.normal()))
            vt_ctx.add_voxel(mx, my)
            s2 = vto.make_scene(dict(d, grid=grid))
            vto.add_voxel(s2, grid, ri, rn, mx, my)
        elif i % 3 == 1:
            vt_ctx.remove_voxel()
            vto.remove_voxel(grid, vol["res"], ri)
        d = dict(d, grid=grid)
        s = vto.make_scene(d)
    assert np.array_equal(vt_ctx.read_volume(), grid)
    # the edited volume renders identically (occupancy bits + mip were updated in place)
    ref, ref_hits, _, _ = vto.render_pass(s, 0)
    vt_ctx.reset_accumulation()
    vt_ctx.set_selection([-1, -1, -1, 0], [1, 0, 0, 0])
    vt_ctx.enable_primary_hits(True)
    vt_ctx.render(0, 1)
    assert util.same_bits(vt_ctx.read_average(), ref).all()
    assert np.array_equal(vt_ctx.read_primary_hits(), ref_hits)


def test_partition_modes_match_single_context(vt_ctx):
    """Tile partition is bit-identical to one GPU; sample partition sums the same samples."""
    import voxeltoy_b200 as vt
    d = util.make_frame(util.scene_fall_volume(), 200, 136, bounces=2, theta=120, phi=30)
    util.upload(vt_ctx, d)
    vt_ctx.render(0, 4)
    full = vt_ctx.read_average()
    world = 3
    tiles = np.zeros_like(full)
    for r in range(world):
        util.upload(vt_ctx, d)
        vt_ctx.set_partition(vt.VT_PART_TILES, r, world)
        vt_ctx.render(0, 4)
        part = vt_ctx.read_average()
        assert not (np.any(tiles != 0, axis=2) & np.any(part != 0, axis=2)).any()     # disjoint tiles
        tiles += part
    assert util.same_bits(tiles, full).all()
    total = np.zeros(full.shape, np.float64)
    for r in range(2):
        util.upload(vt_ctx, d)
        vt_ctx.set_partition(vt.VT_PART_SAMPLES, r, 2)
        vt_ctx.render(0, 2)
        total += vt_ctx.read_average()
    vt_ctx.set_partition(vt.VT_PART_NONE, 0, 1)
    s = vto.make_scene(d)
    ref = sum(vto.render_pass(s, k, want_hits=False)[0].astype(np.float64) for k in range(4))
    assert np.allclose(total, ref, rtol=1e-6, atol=1e-6)


def test_counters_match_oracle(vt_ctx):
    d = util.make_frame(util.scene_fall_volume(), 160, 90, bounces=3, theta=120, phi=30)
    s = vto.make_scene(d)
    util.upload(vt_ctx, d)
    S = R = H = 0
    for k in range(2):
        _, _, _, cnt = vto.render_pass(s, k)
        S += cnt["S"]; R += cnt["R"]; H += cnt["Hm"]
    for variant in (0, 2):
        vt_ctx.set_kernel_variant(variant)
        vt_ctx.reset_accumulation()
        vt_ctx.counters_enable(True)
        vt_ctx.reset_counters()
        vt_ctx.render(0, 2)
        c = vt_ctx.counters()
        vt_ctx.counters_enable(False)
        assert (c["dda_steps"], c["rand_calls"], c["material_evals"]) == (S, R, H), variant
        assert c["paths"] == 2 * 160 * 90
    vt_ctx.set_kernel_variant(2)
