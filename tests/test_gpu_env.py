"""Environment ingest on the device: vt_env_build must leave exactly the state that the host chain
calculateCDF (renderer/image.cpp:68-389) + vt_env_upload leaves -- CDF-U, CDF-V and the integral bit for bit -- and a
render with it must equal a render with the uploaded arrays."""
import numpy as np
import pytest

import voxeltoy_b200 as vt
from oracle import scene as oscene
from tests import util
from voxeltoy_b200 import scenes

pytestmark = pytest.mark.gpu


def _images():
    rng = np.random.default_rng(5)
    yield "sky 256x128", scenes.synthetic_env(256, 128)
    yield "sky 1024x512 (2x box reduction)", scenes.synthetic_env(1024, 512)
    yield "sky 2048x512 (4x1... non-square factors)", np.repeat(scenes.synthetic_env(1024, 512), 2, axis=1)
    yield "sky 1000x500 (non-integer reduction 1.953)", scenes.synthetic_env(1000, 500)
    yield "sky 1500x750 (non-integer reduction 2.93)", scenes.synthetic_env(1500, 750)
    yield "noise 777x1301 (portrait, non-integer)", rng.random((1301, 777, 3)).astype(np.float32)
    noise = rng.standard_normal((37, 53, 3)).astype(np.float32) * 3.0            # negative values are clamped (image.cpp:370)
    noise[5] = 0.0; noise[11] = -1.0                                              # black rows: uniform conditional CDF (:236-241)
    yield "noise 53x37, negatives, black rows", noise
    yield "all black 16x8", np.zeros((8, 16, 3), np.float32)                      # marginal falls back to y/H (:272-277)
    yield "one pixel", np.full((1, 1, 3), 2.5, np.float32)
    hot = np.zeros((64, 128, 3), np.float32); hot[20, 100] = (5e4, 4e4, 3e4)      # one texel carries everything
    yield "single hot texel", hot


def test_env_build_equals_host_chain(vt_ctx):
    for name, rgb in _images():
        ref = oscene.build_env(rgb)
        vt_ctx.env_build(rgb)
        info = vt_ctx.env_info()
        cu, cv = vt_ctx.read_env_cdf()
        assert (info["w"], info["h"]) == (rgb.shape[1], rgb.shape[0]), name
        assert cu.shape == ref["cdf_u"].shape and cv.shape == ref["cdf_v"].shape, name
        assert util.same_bits(cu, ref["cdf_u"]).all(), name
        assert util.same_bits(cv, ref["cdf_v"]).all(), name
        assert np.float32(info["integral"]).view(np.uint32) == np.float32(ref["integral"]).view(np.uint32), name
        assert info["guided"], name                                               # these CDFs are sorted: guide tables in use
    vt_ctx.env_clear()


def test_env_build_rejects_what_calculate_cdf_rejects(vt_ctx):
    with pytest.raises(vt.VtError):
        vt_ctx.env_build(np.ones((1, 2048, 3), np.float32))                       # reduces to zero rows
    assert vt_ctx.env_info()["w"] == 0                                            # nothing half-built is left behind


def test_render_with_built_env_equals_uploaded_env(vt_ctx):
    rgb = scenes.synthetic_env(1024, 512)
    env = oscene.build_env(rgb)
    d = util.make_frame(util.scene_fall_volume(), 160, 90, bounces=3, theta=120, phi=30, lens_model=1, fstop=2.8, env=env, focal_distance=850.0)
    util.upload(vt_ctx, d)
    vt_ctx.render(0, 3)
    a = vt_ctx.read_average()
    vt_ctx.env_build(rgb)
    vt_ctx.reset_accumulation()
    vt_ctx.render(0, 3)
    b = vt_ctx.read_average()
    assert util.same_bits(a, b).all()
    vt_ctx.env_clear()
