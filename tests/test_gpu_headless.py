"""The headless front-end: event traffic played through HeadlessWidget -- the reference's GLWidget routing
(ui/glwidget.cpp:142-223: tool first, renderer second; repaint while samples are pending) -- against the C++ Renderer, with
the final frame checked against the CPU oracle fed with the camera the events produced."""
import numpy as np
import pytest

import voxeltoy_b200 as vt
from oracle import vto
from tests import util
from tests.test_gpu_renderer import _oracle_frame
from voxeltoy_b200 import host

pytestmark = pytest.mark.gpu


def _widget():
    r = host.Renderer(); r.initialize("", 0); r.updateRenderSettings()          # GLWidget::initializeGL
    return r, host.HeadlessWidget(r)


SCRIPT = """
window 200 120                    # resizeGL: RM_MATCH_WINDOW -> the frame follows the window
vox {vox}
max_samples 50
max_bounces 2
paint 4
move 100 60 none                  # no tool: the renderer's camera controller sees the deltas
move 130 70 left                  # orbit drag
move 130 90 right                 # dolly
key f                             # focus on the bounds (renderer.cpp:670-692)
paint 3
"""


def test_script_drives_renderer_like_the_ui(tmp_path):
    r, w = _widget()
    w.run(SCRIPT.format(vox=host.plain_path(util.SCENE_FALL)))
    assert (r.width, r.height) == (200, 120)
    assert r.numberSamples() == 3 and w.updatePending()                        # the camera events restarted the accumulation
    assert w.pump(3) == 3 and r.numberSamples() == 6
    got = r.readAverage()
    # r.context() carries the default selection (0,0,0) -> SURVEY U3, reproduced by the oracle frame default
    s = vto.make_scene(_oracle_frame(r, util.scene_fall_volume(), 2))
    assert util.same_bits(got, vto.render_average(s, 6)).all()
    # the same traffic on a second widget gives the same bits
    r2, w2 = _widget()
    w2.run(SCRIPT.format(vox=host.plain_path(util.SCENE_FALL)) + "paint 3\n")
    assert util.same_bits(r2.readAverage(), got).all()
    # the repaint loop ends by itself: render() returns RR_FINISHED_RENDERING once `m_numberSamples++ < max` fails
    # (renderer.cpp:639-644), i.e. after max + 1 passes, the last one repeating sample index max - 1 (:594)
    w2.run("max_samples 5\npaint\n")
    assert r2.numberSamples() == 6 and not w2.updatePending()
    assert w2.pump() == 0                                                      # nothing pending: no paint, like an idle Qt loop
    r2.close(); r.close()


def test_tools_dialogs_and_slots(tmp_path):
    r, w = _widget()
    png = str(tmp_path / "frame.png")
    w.run("""
window 160 120
vox %s
max_samples 4
paint
tool edit
move 40 70 none                   # ToolAddRemoveVoxel: hover selects the voxel under the cursor
press 40 70 left                  # ... and a click adds one on the picked face
paint 1
""" % host.plain_path(util.SCENE_FALL))
    ctx = r.context()
    assert r.numberSamples() == 1                                              # the actions restarted the accumulation (actions.cpp:25-29)
    grid = util.scene_fall_volume()["grid"]
    assert int((ctx.read_volume() >= 0).sum()) == int((grid >= 0).sum()) + 1
    w.run("""
tool focal
press 80 100 left                 # ToolFocalDistance: autofocus at the cursor
paint 1
tool none
dialog begin
fstop 2.8
paint 5                           # a modal dialog pauses rendering (glwidget.cpp:152-153): paints happen, samples do not
""")
    assert ctx.get_focal_distance() < 9e7 and r.numberSamples() == 0
    w.run("dialog end\nlens 1\nbackground constant 0.5 0.5 0.5\nmaterial_color 4 0.9 0.1 0.2\npaint\nsave %s\n" % png)
    assert r.numberSamples() == 5                                              # max_samples + 1, see above
    assert np.allclose(ctx.read_materials(7)[4:7], [0.9, 0.1, 0.2])
    from tests.test_host_parity import _decode_png
    want = np.rint(np.clip(np.nan_to_num(r.readAverage(), nan=0.0), 0.0, 1.0) * np.float32(255.0)).astype(np.uint8)
    assert np.array_equal(_decode_png(png), want[::-1])
    with pytest.raises(ValueError, match="line 2"):
        w.run("paint 1\nfrobnicate 3\n")
    r.close()


def test_fixed_resolution_letterbox():
    r, w = _widget()
    w.run("window 300 100\nresolution fixed 64 64\n")
    assert (r.width, r.height) == (64, 64)                                     # RM_FIXED keeps the render size (glwidget.cpp:74-110)
    w.run("resolution longest 128\n")
    assert (r.width, r.height) == (128, 42)                                    # RM_LONGEST_AXIS: 128 / (300/100), truncated
    r.close()
