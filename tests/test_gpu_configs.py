"""BASELINE configs 3-5 at reduced size, CUDA path (through the C ABI) vs the CPU oracle, bit for bit.
The full-size versions are run by tools/run_configs.py (size-independent properties + timing)."""
import numpy as np
import pytest

import voxeltoy_b200 as vt
from oracle import scene as oscene
from oracle import vto
from tests import util
from voxeltoy_b200 import scenes

pytestmark = pytest.mark.gpu


def _render_both(ctx, d, n_passes, first=0):
    util.upload(ctx, d)
    ctx.enable_primary_hits(True)
    ctx.render(first, n_passes)
    got, hits = ctx.read_average(), ctx.read_primary_hits()
    s = vto.make_scene(d)
    ref = vto.render_average(s, n_passes, first=first)
    ref_hits = vto.render_pass(s, first + n_passes - 1)[1]
    eq = util.same_bits(got, ref)
    assert eq.all(), "%d floats differ, max |delta| %g" % (int((~eq).sum()), float(np.nanmax(np.abs(got - ref))))
    assert np.array_equal(hits, ref_hits)
    return got


def test_c3_bunny_voxelize_assign_render(vt_ctx):
    """C3 reduced: bunny.obj -> GPU voxelizer at 96^3 -> material rule on the GPU -> path trace (Lambert / metal / emissive)."""
    res = (96, 96, 96)
    verts, idx = oscene.load_obj(util.BUNNY)
    bmin, bmax = oscene.mesh_bounds(verts)
    M = oscene.mesh_transform(bmin, bmax, res)
    t = scenes.c3_material_table()
    vt_ctx.voxelize(verts, idx, M, res, fill_offset=0)
    occ = vto.voxelize(verts, idx, M, res)
    grid0 = np.where(occ > 0, 0, -1).astype(np.int32)
    assert np.array_equal(vt_ctx.read_volume(), grid0)
    # rule 1 of vt_volume_assign_materials = the id part of scenes.c3_assign without the emissive thinning
    vt_ctx.assign_materials(np.asarray(t.offsets, np.int32), rule=1)
    Z, Y, X = res[2], res[1], res[0]
    zz, yy, xx = np.nonzero(grid0.reshape(Z, Y, X) >= 0)
    expect = grid0.copy().reshape(Z, Y, X)
    expect[zz, yy, xx] = np.asarray(t.offsets, np.int32)[((xx >> 5) ^ (yy >> 5) ^ (zz >> 5)) % 3]
    grid = expect.reshape(-1)
    assert np.array_equal(vt_ctx.read_volume(), grid)
    mats = t.array()
    em = oscene.prune_interior_emissive(grid, res, scenes.emissive_list(grid, mats))
    d = util.make_frame(dict(res=res, grid=grid, materials=mats, emissive=em), 192, 108, bounces=4, theta=130, phi=25)
    got = _render_both(vt_ctx, d, 2)
    assert em.size > 0 and np.isfinite(got[..., :3]).mean() > 0.9


def test_c4_terrain_tiles(vt_ctx):
    """C4 reduced: procedural terrain with a metal band and emissive lava, 8 bounces, tile partition over 4 ranks."""
    n = 64
    ids = scenes.terrain_grid(n)
    t = scenes.MaterialTable()
    t.lambert((0.55, 0.5, 0.45)); t.metal((0.8, 0.8, 0.85), 60.0); t.lambert((0.3, 0.1, 0.05), emission=(6.0, 2.0, 0.5))
    grid = scenes.ids_to_offsets(ids, t.offsets); mats = t.array()
    em = oscene.prune_interior_emissive(grid, (n, n, n), scenes.emissive_list(grid, mats))
    d = util.make_frame(dict(res=(n, n, n), grid=grid, materials=mats, emissive=em), 256, 144, bounces=8, theta=140, phi=35)
    full = _render_both(vt_ctx, d, 2)
    world, acc = 4, None
    for r in range(world):
        util.upload(vt_ctx, d)
        vt_ctx.set_partition(vt.VT_PART_TILES, r, world)
        vt_ctx.render(0, 2)
        part = vt_ctx.read_average()
        acc = part if acc is None else acc + part
    vt_ctx.set_partition(vt.VT_PART_NONE, 0, 1)
    assert util.same_bits(acc, full).all()


def test_c5_dense_noise_samples_and_edits(vt_ctx):
    """C5 reduced: dense noise grid (35 % solid, 8 materials), 16 bounces, sample partition over 2 ranks, and an
    add / remove edit in between that resets the accumulation."""
    n = 64
    ids = scenes.dense_noise_grid(n, density=0.35)
    t = scenes.MaterialTable()
    for k in range(8):
        (t.metal((0.9, 0.6 + 0.04 * k, 0.3), 30.0 + 20 * k) if k % 3 == 2 else t.lambert((0.3 + 0.08 * k, 0.5, 0.9 - 0.08 * k)))
    grid = scenes.ids_to_offsets(ids, t.offsets); mats = t.array()
    d = util.make_frame(dict(res=(n, n, n), grid=grid, materials=mats, emissive=np.zeros(0, np.int32)), 160, 120, bounces=16,
                        theta=125, phi=40)
    _render_both(vt_ctx, d, 2)
    # sample partition: two "ranks" render disjoint sample indices and keep sums
    total = np.zeros((120, 160, 4), np.float64)
    for r in range(2):
        util.upload(vt_ctx, d)
        vt_ctx.set_partition(vt.VT_PART_SAMPLES, r, 2)
        vt_ctx.render(0, 2)
        total += vt_ctx.read_average()
    vt_ctx.set_partition(vt.VT_PART_NONE, 0, 1)
    s = vto.make_scene(d)
    ref = sum(vto.render_pass(s, k, want_hits=False)[0].astype(np.float64) for k in range(4))
    assert np.allclose(total, ref, rtol=1e-6, atol=1e-6, equal_nan=True)
    # scripted edit: pick at a fixed pixel -> add -> render; pick again -> remove -> render; both against the oracle
    util.upload(vt_ctx, d)
    g = grid.copy()
    for step in range(2):
        vt_ctx.pick(80.0, 60.0)
        ri, rn = vto.pick(vto.make_scene(dict(d, grid=g)), 80.0, 60.0, near_z=d["near_z"])
        gi, gn = vt_ctx.get_selection()
        assert np.array_equal(gi, ri)
        if step == 0:
            vt_ctx.add_voxel(0.0, 0.0); vto.add_voxel(vto.make_scene(dict(d, grid=g)), g, ri, rn, 0.0, 0.0)
        else:
            vt_ctx.remove_voxel(); vto.remove_voxel(g, (n, n, n), ri)
        assert np.array_equal(vt_ctx.read_volume(), g)
        vt_ctx.set_selection([-1, -1, -1, 0], [1, 0, 0, 0])
        vt_ctx.reset_accumulation()
        vt_ctx.render(0, 1)
        ref1 = vto.render_pass(vto.make_scene(dict(d, grid=g)), 0, want_hits=False)[0]
        assert util.same_bits(vt_ctx.read_average(), ref1).all()


def test_empty_space_skip_is_bit_identical(vt_ctx):
    """The exact empty-space skip (dda_skip) forced on, on scenes with large empty regions, short rays, rays that graze
    solid cells, a nearly empty volume and a dense one: every mode must reproduce the oracle's bits and primary hits."""
    rng = np.random.RandomState(11)
    t = scenes.MaterialTable()
    t.lambert((0.7, 0.6, 0.5)); t.metal((0.9, 0.8, 0.6), 80.0); t.lambert((0.4, 0.4, 0.4), emission=(5.0, 4.0, 3.0))
    cases = []
    n = 96                                                    # a few small boxes floating in a big empty volume + a floor
    ids = np.full((n, n, n), -1, np.int32); ids[:, :2, :] = 0
    for _ in range(12):
        c = rng.randint(4, n - 10, size=3); s = rng.randint(1, 7, size=3)
        ids[c[0]:c[0] + s[0], c[1]:c[1] + s[1], c[2]:c[2] + s[2]] = rng.randint(0, 3)
    cases.append(((n, n, n), ids.reshape(-1)))
    ids = np.full((72, 130, 200), -1, np.int32); ids[5, 100, 17] = 1; ids[60, 3, 180] = 2      # [z, y, x]: non-cubic, almost empty
    cases.append(((200, 130, 72), ids.reshape(-1)))
    cases.append(((64, 64, 64), scenes.dense_noise_grid(64, density=0.02) % 3 * (scenes.dense_noise_grid(64, density=0.02) >= 0) - (scenes.dense_noise_grid(64, density=0.02) < 0)))
    for res, idg in cases:
        grid = scenes.ids_to_offsets(np.asarray(idg, np.int32), t.offsets); mats = t.array()
        em = oscene.prune_interior_emissive(grid, res, scenes.emissive_list(grid, mats))
        for kw in (dict(theta=130, phi=25), dict(theta=40, phi=70, lens_model=1, fstop=2.0, focal_distance=600.0), dict(theta=200, phi=5, distance=150.0)):
            d = util.make_frame(dict(res=res, grid=grid, materials=mats, emissive=em), 160, 96, bounces=3, **kw)
            s = vto.make_scene(d)
            ref = vto.render_average(s, 2); ref_hits = vto.render_pass(s, 1)[1]
            for mode in (2, 0, 1):
                vt_ctx.set_empty_skip(mode)
                util.upload(vt_ctx, d)
                vt_ctx.enable_primary_hits(True)
                vt_ctx.render(0, 2)
                got = vt_ctx.read_average()
                eq = util.same_bits(got, ref)
                assert eq.all(), "skip mode %d, res %r, %r: %d floats differ" % (mode, res, kw, int((~eq).sum()))
                assert np.array_equal(vt_ctx.read_primary_hits(), ref_hits)
    vt_ctx.set_empty_skip(1)


def test_advance_until_equals_literal_loop(vt_ctx):
    """advance_until (`while (d <= tau && k < nmax) d += e`, four additions per trip) against the literal loop on the device,
    on realistic DDA operands and on adversarial ones (exact ties, power-of-two increments, tiny/huge ratios, zero start)."""
    rng = np.random.RandomState(5)
    n = 400000
    e = (10.0 ** rng.uniform(-3, 5, n)).astype(np.float32)
    e[::11] = (e[::11].view(np.uint32) & np.uint32(0xFFFFF000)).view(np.float32)        # low bits zero: ties
    e[::13] = (e[::13].view(np.uint32) & np.uint32(0xFF800000)).view(np.float32)        # powers of two
    K = (10.0 ** rng.uniform(0, 3.3, n)).astype(np.float32)
    d = ((rng.uniform(0, 1, n).astype(np.float32) + np.floor(K)) * e).astype(np.float32)
    d[::7] = 0.0
    d[::17] = (e[::17] * np.float32(2.0 ** 20)).astype(np.float32)                        # d >> e
    nmax = rng.randint(1, 300, n).astype(np.int32)
    steps = rng.randint(0, 320, n).astype(np.float32)
    tau = ((d + steps * e) * rng.choice(np.float32([0.999, 1.0, 0.5, 1.001]), n)).astype(np.float32)
    ok = (e > 0) & (tau > 0) & np.isfinite(tau)
    d, e, tau, nmax = d[ok], e[ok], tau[ok], nmax[ok]
    a_d, a_k = vt_ctx.debug_advance(d, e, tau, nmax, literal=True)
    b_d, b_k = vt_ctx.debug_advance(d, e, tau, nmax, literal=False)
    assert np.array_equal(a_k, b_k), "%d counts differ" % int((a_k != b_k).sum())
    assert np.array_equal(a_d.view(np.uint32), b_d.view(np.uint32))
    assert (a_k > 50).sum() > 10000 and (a_k == nmax).sum() > 1000 and (a_k == 0).sum() > 100


def test_full_size_properties_c3(vt_ctx):
    """BASELINE config 3 at FULL size through size-independent properties (its bit-exact comparison with voxelize.gs is
    tests/test_gpu_fullsize.py): re-voxelizing bunny.obj at 512^3 is idempotent, the occupancy equals the set of written
    offsets, the shell is thin (every solid voxel has an empty 6-neighbour) and 8x the 64^3 surface area within 25 %."""
    verts, idx = oscene.load_obj(util.BUNNY)
    bmin, bmax = oscene.mesh_bounds(verts)
    res = (512, 512, 512)
    M = oscene.mesh_transform(bmin, bmax, res)
    vt_ctx.voxelize(verts, idx, M, res, fill_offset=3)
    a = vt_ctx.read_volume()
    vt_ctx.voxelize(verts, idx, M, res, fill_offset=3)
    b = vt_ctx.read_volume()
    assert np.array_equal(a, b) and set(np.unique(a)) == {-1, 3}
    solid = (a >= 0).reshape(512, 512, 512)
    n512 = int(solid.sum())
    inner = solid[1:-1, 1:-1, 1:-1] & solid[:-2, 1:-1, 1:-1] & solid[2:, 1:-1, 1:-1] & solid[1:-1, :-2, 1:-1] & solid[1:-1, 2:, 1:-1] \
        & solid[1:-1, 1:-1, :-2] & solid[1:-1, 1:-1, 2:]
    assert int(inner.sum()) < n512 // 50                      # a surface shell, not a filled solid
    M64 = oscene.mesh_transform(bmin, bmax, (64, 64, 64))
    n64 = int((vto.voxelize(verts, idx, M64, (64, 64, 64)) > 0).sum())
    assert 0.75 < n512 / (64.0 * n64) < 1.25                  # area scales with the square of the resolution
    vt_ctx.volume_upload(np.full(16 ** 3, -1, np.int32), (16, 16, 16))


def _dense_noise_offsets_torch(n, offsets, **kw):
    return scenes.dense_noise_offsets_torch(n, offsets, **kw)
