"""Edge cases of the C ABI on the GPU: empty and tiny inputs, ragged frames, zero bounces, frames wider than the RNG
stride, error codes instead of fallbacks."""
import numpy as np
import pytest

import voxeltoy_b200 as vt
from oracle import vto
from tests import util
from voxeltoy_b200 import scenes

pytestmark = pytest.mark.gpu


def _check(ctx, d, n_passes=1):
    util.upload(ctx, d)
    ctx.enable_primary_hits(True)
    ctx.render(0, n_passes)
    s = vto.make_scene(d)
    ref = vto.render_average(s, n_passes)
    got = ctx.read_average()
    assert util.same_bits(got, ref).all()
    assert np.array_equal(ctx.read_primary_hits(), vto.render_pass(s, n_passes - 1)[1])
    return got


def test_empty_volume_and_single_voxel(vt_ctx):
    t = scenes.MaterialTable(); t.lambert((0.6, 0.6, 0.6))
    for res, solid in (((16, 16, 16), []), ((5, 3, 9), [(2, 1, 4)]), ((1, 1, 1), [(0, 0, 0)]), ((130, 4, 7), [(129, 3, 6), (0, 0, 0)])):
        grid = np.full(res[0] * res[1] * res[2], -1, np.int32)
        for (x, y, z) in solid:
            grid[x + y * res[0] + z * res[0] * res[1]] = 0
        d = util.make_frame(dict(res=res, grid=grid, materials=t.array(), emissive=np.zeros(0, np.int32)), 96, 64, bounces=2, theta=120, phi=35)
        _check(vt_ctx, d, 2)


def test_zero_bounces_one_pixel_and_ragged_frames(vt_ctx):
    vol = util.scene_fall_volume()
    for W, H, b in ((1, 1, 2), (65, 63, 0), (129, 1, 1), (3, 200, 3)):
        _check(vt_ctx, util.make_frame(vol, W, H, bounces=b, theta=120, phi=30))


def test_frame_wider_than_the_noise_table(vt_ctx):
    """random.h:15-16 uses a stride of 1024 whatever the frame width: pixels (x, y) and (x - 1024, y + 1) share streams."""
    d = util.make_frame(util.scene_fall_volume(), 1100, 6, bounces=2, theta=120, phi=30)
    _check(vt_ctx, d)


def test_many_passes_continue_the_running_average(vt_ctx):
    d = util.make_frame(util.scene_fall_volume(), 64, 48, bounces=2, theta=120, phi=30)
    util.upload(vt_ctx, d)
    vt_ctx.render(0, 3); vt_ctx.render(3, 1); vt_ctx.render(4, 5)          # 9 passes in three calls
    assert vt_ctx.num_samples() == 9
    assert util.same_bits(vt_ctx.read_average(), vto.render_average(vto.make_scene(d), 9)).all()


def test_division_by_constants_is_exact(vt_ctx):
    """x / PI and x / (2 PI) are computed as RN(x * rc) corrected by one exact residual (csrc/vt_math.cuh, gdiv_by): checked against
    div.rn for EVERY binary32 numerator -- zeros, denormals, infinities and NaNs included -- on the device."""
    for which in (0, 1, 2):                                          # / PI, / 2 PI, and the direction clamp of dda.h:29 as one select
        bad, first = vt_ctx.debug_div_const(which)
        assert bad == 0, "constant %d: %d numerators differ, first 0x%08x" % (which, bad, first)


def test_pools_follow_the_frame_size():
    """The wavefront pools are sized by paths per batch PLUS a queue slack that grows with the number of warps appending to the shade
    queues (one partly filled chunk per warp and queue): a small frame with many passes, then a frame with 40x the pixels and the same
    number of paths per batch, must re-size the slack (wf_slack_alloc) instead of running past the queues. Checked against the
    megakernel, which has no queues, bit for bit; and back down again."""
    ctx = vt.Context(0)
    try:
        vol = util.scene_fall_volume()
        for W, H, passes in ((64, 64, 160), (704, 640, 2), (64, 64, 3), (1280, 704, 1)):
            d = util.make_frame(vol, W, H, bounces=3, theta=120, phi=30)
            util.upload(ctx, d)
            imgs = []
            for variant in (2, 0):
                ctx.set_kernel_variant(variant)
                ctx.reset_accumulation()
                ctx.render(0, passes)
                imgs.append(ctx.read_average())
            ctx.set_kernel_variant(2)
            assert util.same_bits(imgs[0], imgs[1]).all(), (W, H, passes)
    finally:
        ctx.close()


def test_errors_are_reported_not_papered_over():
    ctx = vt.Context(0)
    try:
        with pytest.raises(vt.VtError):
            ctx.render(0, 1)                                               # no camera / settings yet
        with pytest.raises(vt.VtError):
            ctx.set_settings(0, 10)
        with pytest.raises(vt.VtError):
            ctx.volume_upload(None, (4096, 1, 1))                           # resolution limit 2048
        with pytest.raises(vt.VtError):
            ctx.set_partition(vt.VT_PART_TILES, 3, 2)
        with pytest.raises(vt.VtError):
            ctx.set_kernel_variant(1)                                       # round 1's per-lane state machine: removed
        with pytest.raises(vt.VtError):
            ctx.voxelize(np.zeros((3, 3), np.float32), np.array([0, 1, 5], np.uint32), np.eye(4, dtype=np.float32), (8, 8, 8))
        t = scenes.MaterialTable(); t.lambert((0.5, 0.5, 0.5))
        ctx.materials_upload(t.array())
        with pytest.raises(vt.VtError):
            ctx.material_update(1000, [1.0])
        ctx.voxelize(np.zeros((0, 3), np.float32), np.zeros(0, np.uint32), np.eye(4, dtype=np.float32), (8, 8, 8))   # no triangles: empty grid
        assert (ctx.read_volume() == -1).all()
    finally:
        ctx.close()
    with pytest.raises(vt.VtError):
        vt.Context(99)                                                     # no such device: there is no CPU path


def test_custom_noise_tables(vt_ctx):
    """vt_noise_upload with tables that are not 1024 x 1024: random.h:16-17 divides by the table's size, the kernels take shifts and
    masks for powers of two and the literal signed division otherwise (negative offsets included: such samples read texel 0)."""
    rng = np.random.default_rng(9)
    vol = util.scene_fall_volume()
    try:
        for (h, w) in ((777, 1000), (32, 64), (1, 1), (2048, 512)):
            noise = rng.random((h, w, 4), dtype=np.float32)
            d = util.make_frame(vol, 96, 64, bounces=2, theta=120, phi=30)
            d["noise"] = noise
            vt_ctx.noise_upload(noise)
            for variant in (2, 0):
                vt_ctx.set_kernel_variant(variant)
                _check(vt_ctx, d, 2)
    finally:
        vt_ctx.set_kernel_variant(2)
        vt_ctx.noise_upload(None)
