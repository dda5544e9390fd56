"""CPU tests: the oracle against the self-derived known-answer values of SURVEY.md 8(c) (the reference ships
no tests or golden vectors of its own) and against closed-form cases."""
import math
import os
import re

import numpy as np
import pytest

from oracle import scene as oscene
from oracle import vto
from tests import util


def test_wang_hash_kat():
    assert [vto.hash32(i) for i in range(4)] == [1062685034, 663891101, 965162357, 2016309182]
    rng = np.random.RandomState(0)
    for x in rng.randint(0, 2 ** 31 - 1, size=200):
        assert vto.hash32(int(x)) < 2 ** 31          # sign bit always clear


def test_rng_offset_kat():
    assert vto.rng_offset(0, 0, 0) == (0, 0)
    assert vto.rng_offset(0, 0, 1) == (503, 345)
    assert vto.rng_offset(1, 0, 0) == (503, 345)
    assert vto.rng_offset(0, 1, 0) == (558, 463)
    assert vto.rng_offset(511, 511, 0) == (669, 511)
    assert vto.rng_offset(1919, 1079, 0) == (981, 97)
    # Q2: stream aliasing for W > 1024
    assert vto.rng_offset(1024 + 5, 7, 3) == vto.rng_offset(5, 8, 3)


def test_noise_table_is_glibc_rand():
    n = vto.noise_table()
    assert n.shape == (1024, 1024, 4)
    first = [1804289383, 846930886, 1681692777, 1714636915]
    assert np.array_equal(n[0, 0], (np.array(first, np.float32) / np.float32(2147483647)).astype(np.float32))
    assert n.min() >= 0.0 and n.max() <= 1.0
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    vals = np.array([libc.rand() for _ in range(4096)], np.float32) / np.float32(2147483647)
    assert np.array_equal(n.reshape(-1)[:4096], vals.astype(np.float32))


def test_dda_step_cap_kat():
    cases = {(16, 16, 16): 56, (64, 64, 64): 222, (126, 20, 126): 360, (256, 256, 256): 888,
             (512, 512, 512): 1774, (1024, 1024, 1024): 3548}
    for r, v in cases.items():
        assert vto.dda_step_cap(*r) == v


def test_scene_fall_vox_kat():
    v = util.scene_fall_volume()
    assert tuple(v["res"]) == (126, 20, 126)
    g = v["grid"].reshape(126, 20, 126)
    assert int((g >= 0).sum()) == 29193
    assert v["materials"].size == 133 and len(set(v["grid"][v["grid"] >= 0].tolist())) == 19
    assert np.array_equal(v["materials"][:7], np.array([0, 0, 0, 0, 0, 100 / 255, 0], np.float32))
    assert int((g[:, 0, :] >= 0).sum()) == 126 * 126 and g[0, 0, 0] >= 0
    assert v["emissive"].size == 0
    bmin, bmax, vs = vto.volume_bounds(126, 20, 126)
    assert np.allclose(bmax, [500, 79.365, 500], atol=1e-3) and np.allclose(vs, 7.93651, atol=1e-4)


def test_scene_fall_first_seen_material_order():
    data = oscene.read_bytes(util.SCENE_FALL)
    _, voxels, _ = oscene.parse_vox(data)
    assert voxels.shape[0] == 29197 and int((voxels[:, 3] == 254).sum()) == 4
    seen = []
    for ci in voxels[:, 3].tolist():
        if ci != 254 and ci not in seen:
            seen.append(ci)
    assert seen == [204, 197, 112, 16, 9, 24, 88, 12, 36, 120, 246, 218, 247, 96, 252, 173, 15, 250, 95]


@pytest.mark.skipif(not os.path.exists("/root/reference/src/renderer/loaders/magicaVoxel.cpp"), reason="reference tree absent")
def test_default_palette_matches_reference_table():
    src = open("/root/reference/src/renderer/loaders/magicaVoxel.cpp").read()
    blk = src[src.index("defaultPalette[ 256 ]"):]
    blk = blk[:blk.index("};")]
    vals = [int(x, 16) for x in re.findall(r"0x[0-9a-f]{8}", blk)]
    ref = np.array(vals, dtype="<u4").view(np.uint8).reshape(256, 4)
    assert np.array_equal(ref, oscene.default_palette())


@pytest.mark.skipif(not os.path.exists("/root/reference/resources/scene_fall.vox"), reason="reference tree absent")
def test_golden_assets_are_the_reference_inputs():
    for name, path in (("scene_fall.vox", util.SCENE_FALL), ("bunny.obj", util.BUNNY)):
        assert oscene.read_bytes(path) == open("/root/reference/resources/" + name, "rb").read()


def test_bunny_obj_kat():
    verts, idx = oscene.load_obj(util.BUNNY)
    assert verts.shape == (2503, 3) and idx.size == 4968 * 3 and idx.max() == 2502


def test_default_camera_kat():
    cam, imv, pm, ipm = util.camera_for((126, 20, 126), 512, 512)
    assert np.allclose(cam.eye, [0, 0, -711.5], atol=0.1)
    assert abs(float(cam.fov_y) - 2 * math.atan2(18, 50)) < 1e-6
    assert abs(float(cam.lens_radius) - 1.5625) < 1e-5
    assert np.allclose(imv @ np.linalg.inv(imv.astype(np.float64)), np.eye(4), atol=1e-5)
    assert pm[3, 2] == -1 and abs(pm[2, 3] - 2 * 10000 * 0.1 / (0.1 - 10000)) < 1e-6


def _tiny_scene(grid, res, **kw):
    mats = np.array([0, 0, 0, 0, 0.5, 0.5, 0.5], np.float32)
    vol = dict(res=res, grid=np.asarray(grid, np.int32).reshape(-1), materials=mats, emissive=np.zeros(0, np.int32))
    return vto.make_scene(util.make_frame(vol, 8, 8, **kw))


def test_dda_hand_built_grid_and_tie_diagonal_step():
    g = np.full((4, 4, 4), -1, np.int32)           # [z, y, x]
    g[3, 3, 3] = 0
    g[1, 1, 2] = 0                                  # would be hit by a face-stepping DDA from (0,0,0) along the diagonal
    s = _tiny_scene(g, (4, 4, 4))
    bmin, bmax, vs = vto.volume_bounds(4, 4, 4)
    # exact body diagonal from the min corner: dis ties on all axes -> diagonal jumps (dda.h:51), reaches (3,3,3)
    o = bmin.copy()
    d = np.full(3, 0.57735026919, np.float32)
    out = vto.trace_rays(s, np.concatenate([o, d])[None])
    assert out[0].tolist() == [3, 3, 3, 1]
    # axis-aligned ray along +x through row y=1,z=1 hits (2,1,1)
    o = bmin + vs * np.array([0.0, 1.5, 1.5], np.float32)
    out = vto.trace_rays(s, np.concatenate([o, [1, 0, 0]]).astype(np.float32)[None])
    assert out[0].tolist() == [2, 1, 1, 1]
    # downward ray in an empty column leaves through the floor: ground hit at y = -1 (dda.h:75-78)
    o = bmin + vs * np.array([0.5, 3.5, 0.5], np.float32)
    out = vto.trace_rays(s, np.concatenate([o, [0, -1, 0]]).astype(np.float32)[None])
    assert out[0].tolist() == [0, -1, 0, 2]
    # upward ray misses; start outside the grid returns a miss with hit position 0 (contract U1)
    out = vto.trace_rays(s, np.concatenate([o, [0, 1, 0]]).astype(np.float32)[None])
    assert out[0, 3] == 0
    out = vto.trace_rays(s, np.concatenate([bmax * 2, [0, -1, 0]]).astype(np.float32)[None])
    assert out[0].tolist() == [0, 0, 0, 0]


def test_aabb_and_accumulate_closed_form():
    avg = np.zeros(8, np.float32)
    for n, v in enumerate([1.0, 3.0, 8.0]):
        vto.accumulate(avg, np.full(8, v, np.float32), n)
    assert np.allclose(avg, 4.0)


def test_cdf_build_and_inverse_sampling_4x2():
    lum = np.array([[1, 1, 0, 2], [0, 0, 0, 0]], np.float32)
    cu, cv, integral = vto.build_cdf(lum)
    s0 = math.sin(math.pi * 0.5 / 2)
    assert np.allclose(cu[0], [0, 0.25, 0.5, 0.5, 1.0])
    assert np.allclose(cu[1], [0, 0.25, 0.5, 0.75, 1.0])         # black row: uniform (image.cpp:240-246)
    assert np.allclose(cv, [0, 1, 1])
    assert abs(integral - (4 * s0 / 8) * 2 * math.pi ** 2) < 1e-4


def test_transcendentals_close_to_libm():
    L = vto.lib()
    rng = np.random.RandomState(1)
    xs = rng.uniform(-20, 20, 4000).astype(np.float32)
    for fn, ref in ((L.vto_m_sin, np.sin), (L.vto_m_cos, np.cos)):
        got = np.array([fn(float(x)) for x in xs], np.float32)
        assert np.max(np.abs(got - ref(xs.astype(np.float64)))) < 4e-7
    xs = rng.uniform(-1, 1, 4000).astype(np.float32)
    got = np.array([L.vto_m_acos(float(x)) for x in xs], np.float32)
    assert np.max(np.abs(got - np.arccos(xs.astype(np.float64)))) < 6e-7
    a, b = rng.normal(size=4000).astype(np.float32), rng.normal(size=4000).astype(np.float32)
    got = np.array([L.vto_m_atan2(float(y), float(x)) for y, x in zip(a, b)], np.float32)
    assert np.max(np.abs(got - np.arctan2(a.astype(np.float64), b.astype(np.float64)))) < 6e-7
    x = rng.uniform(1e-4, 4, 4000).astype(np.float32); y = rng.uniform(0.1, 6, 4000).astype(np.float32)
    got = np.array([L.vto_m_pow(float(p), float(q)) for p, q in zip(x, y)], np.float64)
    ref = np.power(x.astype(np.float64), y.astype(np.float64))
    assert np.max(np.abs(got - ref) / ref) < 3e-6
    assert L.vto_m_pow(0.0, 1 / 2.2) == 0.0 and math.isnan(L.vto_m_pow(-1.0, 2.0))


def test_voxelizer_single_triangle_closed_form():
    # axis-aligned triangle in the plane z = 2.5 (voxel space of an 8^3 grid): one voxel layer, z = 2
    verts = np.array([[1, 1, 2.5], [6, 1, 2.5], [1, 6, 2.5]], np.float32) / 8
    occ = vto.voxelize(verts, np.array([0, 1, 2], np.uint32), np.eye(4, dtype=np.float32), (8, 8, 8)).reshape(8, 8, 8)
    assert occ[2].sum() == occ.sum() and occ[2, 1, 1] == 1 and occ[2, 6, 6] == 0 and occ.sum() >= 15


def test_prune_interior_emissive():
    g = np.zeros((3, 3, 3), np.int32)               # all solid
    em = np.arange(27, dtype=np.int32)
    kept = oscene.prune_interior_emissive(g.reshape(-1), (3, 3, 3), em)
    assert 13 not in kept.tolist() and kept.size == 26
