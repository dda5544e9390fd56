"""CUDA path against the reference's OWN GLSL programs (oracle/_ref/libvt_ref.so: the shader text of /root/reference/src/shaders
compiled for the CPU by oracle/shim) with no restatement in between: every float of the sample image, the occupancy of the
voxelizer and the results of the pick / focal / add / remove programs must have the reference's bits. The library is prebuilt
(it travels to the GPU box); nothing here reads /root/reference."""
import numpy as np
import pytest

import voxeltoy_b200 as vt
from oracle import ref
from oracle import scene as oscene
from oracle import vto          # scene struct marshalling only (make_scene); no oracle arithmetic is used as the expected value
from tests import util
from voxeltoy_b200 import scenes

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libvt_ref.so not built")]


def _one_pass(ctx, d, k):
    util.upload(ctx, d)
    ctx.render(k, 1)                       # after a reset the running average of one pass is the pass itself
    return ctx.read_average()


def test_path_tracer_matches_reference_glsl(vt_ctx):
    env = oscene.build_env(scenes.synthetic_env(128, 64)); env["rotation"] = 0.7
    vol = util.scene_fall_volume()
    frames = [util.make_frame(vol, 128, 128, bounces=1, bg="grey"),                                        # BASELINE config 1, reduced
              util.make_frame(vol, 128, 72, bounces=4, theta=120, phi=30, lens_model=1, fstop=2.8, env=env, focal_distance=650.0),   # config 2
              util.make_frame(util.mixed_scene(), 120, 96, bounces=5, theta=115, phi=40),                   # metal / plastic / emissive, NaN pixels
              util.make_frame(util.mixed_scene(), 96, 64, bounces=3, theta=30, phi=60, wire_opacity=0.6, lens_model=2)]
    for d in frames:
        s = vto.make_scene(d)
        for k in (0, 5):
            want = ref.render_pass(s, k)
            for variant in (2, 0):
                vt_ctx.set_kernel_variant(variant)
                got = _one_pass(vt_ctx, d, k)
                eq = util.same_bits(got, want)
                assert eq.all(), "variant %d sample %d: %d floats differ" % (variant, k, int((~eq).sum()))
    vt_ctx.set_kernel_variant(2)


def test_voxelizer_and_services_match_reference_glsl(vt_ctx):
    verts, idx = oscene.load_obj(util.BUNNY)
    bmin, bmax = oscene.mesh_bounds(verts)
    for n in (48, 96):
        res = (n, n, n)
        M = oscene.mesh_transform(bmin, bmax, res)
        vt_ctx.voxelize(verts, idx, M, res, fill_offset=0)
        assert np.array_equal(vt_ctx.read_volume() >= 0, ref.voxelize(verts, idx, M, res) > 0)
    d = util.make_frame(util.scene_fall_volume(), 160, 120, bounces=1, theta=120, phi=30)
    util.upload(vt_ctx, d)
    s = vto.make_scene(d)
    prev = vt_ctx.get_selection()[1]       # selectVoxel.vs leaves the normal of the previous pick in place on a miss
    for (px, py) in ((80.0, 60.0), (40.0, 70.0), (3.0, 117.0), (150.0, 20.0)):
        vt_ctx.pick(px, py)
        gi, gn = vt_ctx.get_selection()
        ri, rn = ref.pick(s, px, py, near_z=d["near_z"], prev_normal=tuple(float(x) for x in prev))
        prev = gn
        assert np.array_equal(gi, ri) and util.same_bits(gn, rn).all()
        vt_ctx.pick_focal(px, py)
        assert np.float32(vt_ctx.get_focal_distance()).view(np.uint32) == np.float32(ref.pick_focal(s, px, py)).view(np.uint32)


def test_fat_voxelizer_matches_reference_glsl(vt_ctx):
    """vt_set_voxelize_thickness(FAT): the CUDA scatter against voxelize.gs compiled with `#define THICKNESS FAT` (:15-19), bit for
    bit, at 64^3, on a non-cubic grid and at the metric's 512^3; THIN is restored afterwards."""
    verts, idx = oscene.load_obj(util.BUNNY)
    bmin, bmax = oscene.mesh_bounds(verts)
    vt_ctx.set_voxelize_thickness(True)
    try:
        for res in [(64, 64, 64), (96, 80, 48), (512, 512, 512)]:
            M = oscene.mesh_transform(bmin, bmax, res)
            vt_ctx.voxelize(verts, idx, M, res, fill_offset=0)
            got = np.packbits(vt_ctx.read_volume() >= 0)
            assert np.array_equal(got, np.packbits(ref.voxelize(verts, idx, M, res, fat=True) > 0)), res
            assert np.array_equal(got, np.packbits(vto.voxelize(verts, idx, M, res, fat=True) > 0)), res
    finally:
        vt_ctx.set_voxelize_thickness(False)
    M = oscene.mesh_transform(bmin, bmax, (64, 64, 64))
    vt_ctx.voxelize(verts, idx, M, (64, 64, 64), fill_offset=0)
    assert np.array_equal(vt_ctx.read_volume() >= 0, ref.voxelize(verts, idx, M, (64, 64, 64)) > 0)
    vt_ctx.volume_upload(np.full(16 ** 3, -1, np.int32), (16, 16, 16))
