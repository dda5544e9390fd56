"""Pins the CPU oracle (oracle/vto.c, a C restatement) against the reference's OWN GLSL programs compiled for the CPU
(oracle/_ref/libvt_ref.so, built by oracle/shim/Makefile from /root/reference/src/shaders): same scene struct, same
built-in arithmetic (oracle/vto_math.h), so every result must agree BIT FOR BIT. A mismatch means the restatement
departed from the shader text (control flow, operand order, RNG consumption, tie handling ...).
CPU only; the library is prebuilt where /root/reference is absent."""
import os

import numpy as np
import pytest

from oracle import ref
from oracle import scene as oscene
from oracle import vto
from tests import util

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libvt_ref.so not built (needs /root/reference at build time)")


def _same(a, b):
    eq = util.same_bits(a, b)
    assert eq.all(), "%d / %d floats differ, max |delta| = %g" % (int((~eq).sum()), eq.size, np.nanmax(np.abs(a - b)))


def test_path_tracer_scene_fall_pinhole_gradient():
    """BASELINE config 1 (reduced): scene_fall.vox, 1 bounce, pinhole, and 4 bounces from an orbited camera."""
    vol = util.scene_fall_volume()
    for d in (util.make_frame(vol, 128, 128, bounces=1, bg="grey"),
              util.make_frame(vol, 160, 90, bounces=4, theta=120, phi=30),
              util.make_frame(vol, 96, 64, bounces=0, theta=60, phi=20)):
        s = vto.make_scene(d)
        for k in (0, 1, 7):
            _same(vto.render_pass(s, k, want_hits=False)[0], ref.render_pass(s, k))


def test_path_tracer_ibl_thin_lens():
    """BASELINE config 2's feature set: importance-sampled IBL (CDF search, bilinear lookup, rotation) + thin lens."""
    from voxeltoy_b200 import scenes
    env = oscene.build_env(scenes.synthetic_env(128, 64))
    env["rotation"] = 0.7
    d = util.make_frame(util.scene_fall_volume(), 128, 72, bounces=4, theta=120, phi=30, lens_model=1, fstop=2.8, env=env,
                        focal_distance=650.0)
    s = vto.make_scene(d)
    for k in (0, 3):
        _same(vto.render_pass(s, k, want_hits=False)[0], ref.render_pass(s, k))


def test_path_tracer_mixed_materials_emissive_wireframe_selection():
    """Lambert + metal + plastic + unknown type, emissive-voxel light sampling, wireframe overlay, selected voxel,
    orthographic lens; NaNs from the unguarded microfacet terms must appear in the same pixels."""
    vol = util.mixed_scene()
    vol["materials"] = np.concatenate([vol["materials"], np.array([7.0, 0, 0, 0, 1, 1, 1, 0], np.float32)])   # a type-7 record
    solid = np.flatnonzero(vol["grid"] >= 0)
    vol["grid"] = vol["grid"].copy(); vol["grid"][solid[::97]] = len(vol["materials"]) - 8
    n = vol["res"][0]
    sel = (int(solid[5] % n), int((solid[5] // n) % n), int(solid[5] // (n * n)))
    for kw in (dict(bounces=5, theta=115, phi=40), dict(bounces=3, theta=30, phi=60, wire_opacity=0.6, sel=sel),
               dict(bounces=2, theta=200, phi=35, lens_model=2)):
        d = util.make_frame(vol, 120, 96, **kw)
        s = vto.make_scene(d)
        a = vto.render_pass(s, 2, want_hits=False)[0]; b = ref.render_pass(s, 2)
        _same(a, b)
        assert np.array_equal(np.isnan(a), np.isnan(b))


def test_edit_mode_preview():
    d = util.make_frame(util.scene_fall_volume(), 128, 80, bounces=1, theta=100, phi=35, wire_opacity=0.5)
    s = vto.make_scene(d)
    _same(vto.preview_pass(s, 0), ref.preview_pass(s, 0))


def test_services_pick_focal_add_remove():
    vol = util.scene_fall_volume()
    d = util.make_frame(vol, 200, 120, bounces=1, theta=120, phi=30)
    s = vto.make_scene(d)
    rng = np.random.RandomState(1)
    for _ in range(60):
        px, py = float(rng.uniform(0, 200)), float(rng.uniform(0, 120))
        i0, n0 = vto.pick(s, px, py, near_z=d["near_z"]); i1, n1 = ref.pick(s, px, py, near_z=d["near_z"])
        assert np.array_equal(i0, i1) and util.same_bits(n0, n1).all(), (px, py, i0, i1, n0, n1)
        f0, f1 = vto.pick_focal(s, px, py), ref.pick_focal(s, px, py)
        assert np.float32(f0).tobytes() == np.float32(f1).tobytes(), (px, py, f0, f1)
        for mx, my in ((0.0, 0.0), (0.01, -0.002), (-0.003, 0.004), (0.0, -0.02)):
            grid = vol["grid"].copy()
            ok, c0 = vto.add_voxel(s, grid, i0, n0, mx, my)
            c1 = ref.add_voxel(s, i1, n1, mx, my)
            assert np.array_equal(c0, c1), (mx, my, c0, c1)
        assert np.array_equal(ref.remove_voxel(i1), i1[:3])


@pytest.mark.parametrize("res", [(32, 32, 32), (64, 64, 64), (96, 96, 96)])
def test_voxelizer_bunny(res):
    """shared/voxelize.{vs,gs} run per vertex / per triangle vs the oracle: identical occupancy sets."""
    verts, idx = oscene.load_obj(util.BUNNY)
    bmin, bmax = oscene.mesh_bounds(verts)
    M = oscene.mesh_transform(bmin, bmax, res)
    a = vto.voxelize(verts, idx, M, res); b = ref.voxelize(verts, idx, M, res)
    assert a.sum() > 0 and np.array_equal(a, b), "%d voxels differ" % int((a != b).sum())


def test_voxelizer_random_and_degenerate_triangles():
    rng = np.random.RandomState(7)
    verts = rng.uniform(0.05, 0.95, size=(300, 3)).astype(np.float32)
    verts[10] = verts[11]                                  # degenerate triangles: repeated vertex, collinear points
    verts[20] = (verts[21] + verts[22]) * np.float32(0.5)
    verts[30:33] = np.float32([[0.25, 0.25, 0.5], [0.75, 0.25, 0.5], [0.25, 0.75, 0.5]])   # axis aligned, on voxel faces at 16^3
    idx = np.arange(300, dtype=np.uint32)
    M = np.eye(4, dtype=np.float32)
    for res in ((16, 16, 16), (40, 40, 40)):
        a = vto.voxelize(verts, idx, M, res); b = ref.voxelize(verts, idx, M, res)
        assert np.array_equal(a, b), "%d voxels differ" % int((a != b).sum())


def test_render_pixels_equals_full_frame_both_sides():
    """The pixel-list entry points used by the full-size GPU tests (C5 at 4K) are the same shader invocations as the
    full-screen pass: oracle == reference GLSL == the corresponding pixels of a full frame."""
    d = util.make_frame(util.mixed_scene(), 200, 120, bounces=5, theta=130, phi=25, lens_model=1, fstop=2.0, focal_distance=500.0)
    s = vto.make_scene(d)
    full, hits, _, _ = vto.render_pass(s, 3)
    ys, xs = np.mgrid[0:120:7, 0:200:5]
    xy = np.stack([xs.ravel(), ys.ravel()], 1)
    a, h = vto.render_pixels(s, 3, xy, want_hits=True)
    _same(a, full[xy[:, 1], xy[:, 0]])
    assert np.array_equal(h, hits[xy[:, 1], xy[:, 0]])
    _same(ref.render_pixels(s, 3, xy), a)


@pytest.mark.skipif(not ref.obj_available(), reason="oracle/_ref/libvt_ref_obj.so not built (needs /root/reference at build time)")
def test_obj_readers_match_the_reference_tinyobjloader(tmp_path):
    """The reference's own OBJ reader (thirdParty/tinyobjloader/tiny_obj_loader.cc, compiled unmodified by oracle/shim/Makefile,
    + the shape merge of mesh/meshLoader.cpp:27-64) pins BOTH restatements: the oracle's (oracle/scene.py load_obj) and the
    product's MeshLoader::loadFromOBJ (voxeltoy_b200/host/loaders.cpp). bunny.obj, and a file with quads and polygons (fans),
    negative indices, v/vt/vn and v//vn corners, several groups and objects, comments, exponents and odd white space."""
    import gzip
    import voxeltoy_b200 as vt
    bunny = str(tmp_path / "bunny.obj")
    with open(bunny, "wb") as f:
        f.write(gzip.open(util.BUNNY).read())
    tricky = str(tmp_path / "tricky.obj")
    with open(tricky, "w") as f:
        f.write("# comment\nmtllib none.mtl\no first\n"
                "v 0 0 0\nv 1.5 0 0\nv 1.5e0 1 0\nv 0 1 -2.25E-1\n v   .5\t.5  1 \nv -1 -1 -1\n"
                "vt 0 0\nvt 1 0\nvt 1 1\nvn 0 0 1\nvn 0 1 0\n"
                "g quad\nf 1 2 3 4\n"
                "g mixed\nusemtl m\nf 1/1/1 2/2/1 5/3/2\nf 2//1 3//1 5//2\n"
                "o second\nv 2 2 2\nv 3 2 2\nv 3 3 2\nv 2 3 2\nv 2.5 3.5 2\n"
                "f -5 -4 -3 -2 -1\ns off\nf 7 8 9\nf 1 7 10\n"
                "g dropped\nf 1 2 3\nusemtl other\nf 3 4 5\nf 1 2 3\n")        # usemtl drops the face gathered before it
    for path in (bunny, tricky):
        rv, ri = ref.load_obj(path)
        ov, oi = oscene.load_obj(path)
        hv, hi = vt.host.load_obj(path)
        assert ri.size > 0 and ri.size % 3 == 0
        for name, v, i in (("oracle", ov, oi), ("product", hv, hi)):
            v = np.asarray(v, np.float32).reshape(-1, 3)
            assert np.array_equal(np.asarray(i, np.uint32), ri), name + ": indices differ from tinyobjloader on " + os.path.basename(path)
            assert v.shape == rv.shape and np.array_equal(v.view(np.uint32), rv.view(np.uint32)), name + ": vertices differ on " + os.path.basename(path)


def test_fat_voxelizer_matches_reference_glsl():
    """The FAT variant of voxelize.gs (`#define THICKNESS FAT`, :15-19, :151-163, :170-171, :198-200 -- compiled out in the reference,
    built here by rewriting that one #define): oracle == shader text on the bunny and on random / degenerate triangles, and every
    THIN voxel is also a FAT voxel."""
    verts, idx = oscene.load_obj(util.BUNNY)
    bmin, bmax = oscene.mesh_bounds(verts)
    for res in [(32, 32, 32), (64, 64, 64), (96, 80, 48)]:
        M = oscene.mesh_transform(bmin, bmax, res)
        a, b = ref.voxelize(verts, idx, M, res, fat=True), vto.voxelize(verts, idx, M, res, fat=True)
        assert np.array_equal(a > 0, b > 0)
        thin = vto.voxelize(verts, idx, M, res)
        assert ((thin > 0) <= (b > 0)).all() and (b > 0).sum() > (thin > 0).sum()
    rng = np.random.RandomState(2)
    v = rng.uniform(0.05, 0.95, size=(300, 3)).astype(np.float32)
    v[::7] = np.round(v[::7] * 16) / 16                                  # vertices on voxel boundaries: the == comparisons of :203-204
    i = rng.randint(0, 300, size=(400, 3)).astype(np.uint32)
    i[5] = (1, 1, 2); i[6] = (3, 3, 3)                                   # degenerate
    M = np.eye(4, dtype=np.float32)
    for res in [(16, 16, 16), (40, 24, 56)]:
        assert np.array_equal(ref.voxelize(v, i.reshape(-1), M, res, fat=True) > 0, vto.voxelize(v, i.reshape(-1), M, res, fat=True) > 0)
