"""GPU tests of the drop-in class API: the C++ Renderer (driven through its flat wrappers the way the reference's
UI drives it) against the CPU oracle fed with the oracle's own loaders and camera."""
import os

import numpy as np
import pytest

from oracle import scene as oscene
from oracle import vto
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture()
def renderer():
    import voxeltoy_b200 as vt
    r = vt.host.Renderer()
    r.initialize("", 0)
    yield r
    r.close()


def _oracle_frame(r, vol, bounces, env=None, lens_model=0, focal=99999999.0, sel=(0, 0, 0)):
    imv, pm, ipm = r.cameraMatrices()
    cp = r.camera().parameters()
    d = dict(vol)
    d.update(W=r.width, H=r.height, inv_modelview=imv, proj=pm, inv_proj=ipm, max_bounces=bounces, lens_model=lens_model,
             lens_radius=cp["lensRadius"], focal_distance=focal, near_z=cp["near"], sel_index=sel,
             bg_top=util.GRADIENT_TOP, bg_bottom=util.GRADIENT_BOTTOM)
    if env is not None:
        d["env"] = env
    return d


def test_renderer_defaults_and_load_vox(renderer):
    """Renderer ctor + loadVoxFile defaults (renderer.cpp:41-65, import.cpp:11-44): 512x512, 1 bounce, pinhole, f/16,
    eye at half the bounds diagonal; the default selection (0,0,0) paints that voxel red."""
    r = renderer
    r.loadVoxFile(util.SCENE_FALL)
    res, bmin, bmax = r.volumeInfo()
    assert res == (126, 20, 126)
    cp = r.camera().parameters()
    assert np.allclose(cp["eye"], [0, 0, -711.5], atol=0.1) and abs(cp["lensRadius"] - 1.5625) < 1e-5
    # the host camera agrees with the oracle's independent restatement to float rounding
    _, imv, pm, ipm = util.camera_for(res, 512, 512)
    got = r.cameraMatrices()
    assert np.allclose(got[0], imv, atol=1e-4) and np.allclose(got[1], pm, rtol=1e-6) and np.allclose(got[2], ipm, rtol=1e-5, atol=1e-7)
    assert r.render() == 0 and r.numberSamples() == 1
    r.renderPasses(2)
    got = r.readAverage()
    s = vto.make_scene(_oracle_frame(r, util.scene_fall_volume(), 1))
    assert util.same_bits(got, vto.render_average(s, 3)).all()


def test_renderer_c2_flow_env_thin_lens_autofocus(renderer, tmp_path):
    """BASELINE config 2 at reduced size, set up exactly like bench.py does."""
    import voxeltoy_b200 as vt
    from voxeltoy_b200 import scenes
    r = renderer
    r.resizeFrame(240, 136)
    r.loadVoxFile(util.SCENE_FALL)
    env_rgb = scenes.synthetic_env(256, 128)
    path = str(tmp_path / "env.pfm")
    vt.host.write_pfm(path, env_rgb)
    r.setRenderSettings(maxBounces=4, backgroundImage=path)
    cam = r.camera()
    cam.setLensModel(vt.host.CLM_THIN_LENS)
    cam.controller().orbitAroundTarget(np.radians(120), np.radians(30))
    cam.setFStop(2.8)
    r.resetRender()                      # the UI calls resetRender() after every camera change (ui/glwidget.cpp:225-243)
    ctx = r.context()
    ctx.set_selection([-1, -1, -1, 0], [1, 0, 0, 0])
    r.requestAction(0.5, 0.5, 0, 0, vt.host.PA_SELECT_FOCAL_POINT)
    r.renderPasses(3)
    assert r.numberSamples() == 3
    got = r.readAverage()
    focal = ctx.get_focal_distance()
    env = oscene.build_env(env_rgb)
    d = _oracle_frame(r, util.scene_fall_volume(), 4, env=env, lens_model=1, sel=(-1, -1, -1))
    s = vto.make_scene(d)
    assert focal == vto.pick_focal(s, 0.5 * 240, (1.0 - 0.5) * 136)        # actions.cpp:31 flips y
    d["focal_distance"] = focal
    s = vto.make_scene(d)
    assert util.same_bits(got, vto.render_average(s, 3)).all()
    # saveImage flips vertically (renderer.cpp:1132-1136); PFM stores bottom-up, so the file equals GL order
    out = str(tmp_path / "o.pfm")
    r.saveImage(out)
    assert util.same_bits(vt.host.load_image(out)[::-1], got[..., :3]).all()
    # 8-bit export = the display blit (textureMap.fs) as a read-out: GL float -> UNORM8 conversion on the device, top row first
    from tests.test_host_parity import _decode_png
    want = np.rint(np.clip(np.nan_to_num(got, nan=0.0), 0.0, 1.0) * np.float32(255.0)).astype(np.uint8)
    assert np.array_equal(r.context().read_display(), want)
    png = str(tmp_path / "o.png")
    r.saveImage(png)
    assert np.array_equal(_decode_png(png), want[::-1])
    ppm = str(tmp_path / "o.ppm")
    r.saveImage(ppm)
    raw = open(ppm, "rb").read()
    head = b"P6\n240 136\n255\n"
    assert raw.startswith(head) and np.array_equal(np.frombuffer(raw[len(head):], np.uint8).reshape(136, 240, 3), want[::-1, :, :3])
    # OpenEXR keeps the floats (FLOAT channels, ZIP), top row first; and the written file is itself a valid environment map
    exr = str(tmp_path / "o.exr")
    r.saveImage(exr)
    assert util.same_bits(vt.host.load_image(exr), got[::-1, :, :3]).all()
    pfm_env = str(tmp_path / "env2.pfm")                       # a new name: the renderer keeps an image it has already loaded (renderer.cpp:953)
    vt.host.write_pfm(pfm_env, np.nan_to_num(got[::-1, :, :3], nan=0.0))
    r.setRenderSettings(backgroundImage=pfm_env); r.resetRender(); r.renderPasses(2); from_pfm = r.context().read_average()
    exr_env = str(tmp_path / "env2.exr")
    vt.host.write_exr(exr_env, np.nan_to_num(got[::-1, :, :3], nan=0.0))
    r.setRenderSettings(backgroundImage=exr_env); r.resetRender(); r.renderPasses(2)
    assert util.same_bits(r.context().read_average(), from_pfm).all()             # same pixels in, same frame out


def test_renderer_mesh_tools_and_edit_flow(renderer):
    """loadMesh (64^3, import.cpp:75) + the add/remove tool + accumulation restart (actions.cpp:25-29)."""
    import voxeltoy_b200 as vt
    r = renderer
    r.resizeFrame(160, 120)
    r.loadMesh(util.BUNNY)
    res, _, _ = r.volumeInfo()
    assert res == (64, 64, 64)
    verts, idx = oscene.load_obj(util.BUNNY)
    bmin, bmax = oscene.mesh_bounds(verts)
    occ = vto.voxelize(verts, idx, oscene.mesh_transform(bmin, bmax, (64, 64, 64)), (64, 64, 64))
    ctx = r.context()
    grid = ctx.read_volume()
    assert np.array_equal(grid >= 0, occ > 0)
    assert r.getMaterials() == [(0, 0)]
    r.renderPasses(4)
    assert r.numberSamples() == 4
    tool = vt.host.Tool(r, 0)
    tool.mouseMoveEvent(80, 60, 0, 0, 160, 120)                       # select under the cursor
    tool.mousePressEvent(80, 60, vt.host.LeftButton, 0, 160, 120)     # add on the picked face
    r.render()
    assert r.numberSamples() == 1                                     # the actions restarted the accumulation
    sel, normal = ctx.get_selection()
    grid2 = ctx.read_volume()
    changed = np.nonzero(grid2 != grid)[0]
    assert changed.size == 1
    c = sel[:3] + normal[:3].astype(np.int32)
    assert changed[0] == c[0] + c[1] * 64 + c[2] * 64 * 64
    tool.mouseMoveEvent(80, 60, vt.host.LeftButton, vt.host.ControlModifier, 160, 120)   # select + remove
    r.render()
    grid3 = ctx.read_volume()
    assert (grid3 >= 0).sum() == (grid2 >= 0).sum() - 1
    # material edit round trip (updateMaterialColor, renderer.cpp:1205-1219)
    r.updateMaterialColor(4, [0.9, 0.1, 0.2])
    assert np.allclose(ctx.read_materials(7)[4:7], [0.9, 0.1, 0.2])
    # keyboard: space toggles the integrator, F focuses the bounds; both restart accumulation (renderer.cpp:670-692)
    r.renderPasses(2)
    assert r.onKeyPress(vt.host.Key_Space) and r.numberSamples() == 0
    r.render()
    assert r.onKeyPress(vt.host.Key_F) and r.numberSamples() == 0
    assert r.onMouseMove(10, 5, vt.host.RightButton)
    assert not r.onKeyPress(0x51)


def test_uninitialized_renderer_is_inert():
    import voxeltoy_b200 as vt
    r = vt.host.Renderer()
    assert r.render() == vt.host.RR_FINISHED_RENDERING      # renderer.cpp:558
    r.loadVoxFile(util.SCENE_FALL)
    r.resetRender()
    assert r.numberSamples() == 0
    r.close()


def test_material_authoring_round_trip(renderer, tmp_path):
    """.vox palette rules -> Metal / Plastic / emissive records -> getMaterials (renderer.cpp:1142-1203) ->
    updateMaterialColor / updateMaterialValue (:1205-1235) -> the render follows, bit for bit against the oracle fed with the same
    edited material array."""
    import struct
    import voxeltoy_b200 as vt
    rng = np.random.RandomState(6)
    n = 24
    vox = [(x, y, 0, 3) for x in range(n) for y in range(n)]                            # a floor (MagicaVoxel z is up)
    for _ in range(30):
        c = rng.randint(2, n - 4, size=2); h = int(rng.randint(2, 8)); ci = int(rng.choice([9, 17, 40]))
        vox += [(int(c[0]) + dx, int(c[1]) + dy, z, ci) for dx in range(2) for dy in range(2) for z in range(1, h)]
    xyzi = struct.pack("<i", len(vox)) + b"".join(struct.pack("<4B", *v) for v in vox)
    chunks = b"SIZE" + struct.pack("<ii", 12, 0) + struct.pack("<iii", n, n, 10) + b"XYZI" + struct.pack("<ii", len(xyzi), 0) + xyzi
    path = str(tmp_path / "authored.vox")
    with open(path, "wb") as f:
        f.write(b"VOX " + struct.pack("<i", 150) + b"MAIN" + struct.pack("<ii", 0, len(chunks)) + chunks)
    rules = [(9, vt.host.MT_METAL, (0.0, 0.0, 0.0), 120.0), (17, vt.host.MT_LAMBERT, (4.0, 3.0, 2.0), 0.0), (40, vt.host.MT_PLASTIC, (0.0, 0.0, 0.0), 35.0)]
    r = renderer
    r.resizeFrame(160, 120)
    r.setVoxPaletteRules(rules)
    r.loadVoxFile(path)
    r.setVoxPaletteRules([])
    r.setRenderSettings(maxBounces=3)
    ref = oscene.load_vox(path, palette_rules=rules)
    ctx = r.context()
    assert np.array_equal(ctx.read_volume(), ref["grid"])
    mats = r.getMaterials()                                                        # [(type, offset)] in storage order
    assert sorted(t for t, _ in mats) == [0, 0, 1, 2]
    assert [o for _, o in mats] == sorted(set(int(v) for v in ref["grid"][ref["grid"] >= 0]))
    metal = [o for t, o in mats if t == 1][0]; plastic = [o for t, o in mats if t == 2][0]; lamb = [o for t, o in mats if t == 0][0]
    # the edits the reference's material panel makes (ui/mainwindow.cpp:155-250): reflectance colour, roughness, emission
    r.updateMaterialColor(metal + 4, [0.95, 0.64, 0.54])
    r.updateMaterialValue(metal + 7, 30.0)
    r.updateMaterialValue(plastic + 7, 400.0)
    r.updateMaterialColor(lamb + 1, [0.0, 0.25, 0.0])
    m2 = ref["materials"].copy()
    m2[metal + 4: metal + 7] = [0.95, 0.64, 0.54]; m2[metal + 7] = 30.0; m2[plastic + 7] = 400.0; m2[lamb + 1: lamb + 4] = [0.0, 0.25, 0.0]
    assert np.array_equal(ctx.read_materials(m2.size), m2)
    r.resetRender()
    ctx.set_selection([-1, -1, -1, 0], [1, 0, 0, 0])
    r.renderPasses(2)
    em = oscene.prune_interior_emissive(ref["grid"], ref["res"], ref["emissive"])
    d = _oracle_frame(r, dict(res=ref["res"], grid=ref["grid"], materials=m2, emissive=em), 3, sel=(-1, -1, -1))
    want = vto.render_average(vto.make_scene(d), 2)
    got = r.readAverage()
    assert util.same_bits(got, want).all(), "%d floats differ" % int((~util.same_bits(got, want)).sum())
