"""Sanity check of vt_group_last_exchange_ms on one GPU: a group of two contexts on the same device (peer exchange),
1080p frame; prints the library's figure next to a host-side wall clock of the same combination."""
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import voxeltoy_b200 as vt
from voxeltoy_b200 import group as vg, host


def main():
    rs = []
    for i in range(2):
        r = host.Renderer(); r.initialize("", 0); r.resizeFrame(1920, 1080)
        r.loadVoxFile("tests/golden/scene_fall.vox.gz"); r.setRenderSettings(maxBounces=2); r.resetRender()
        rs.append(r)
    g = vg.DeviceGroup.adopt([r.context() for r in rs], vg.PART_SAMPLES)
    for r in rs:
        r.renderPasses(2)
    for k in range(3):
        g.sync(); t = time.perf_counter(); g.begin_combine(); g.sync(); wall = (time.perf_counter() - t) * 1e3
        print("exchange %d: library %.4f ms, host wall clock around begin_combine + sync %.4f ms" % (k, g.last_exchange_ms(), wall))
    g.close()
    for r in rs:
        r.close()


if __name__ == "__main__":
    main()
