"""Headless front-end: plays an event script against Renderer through HeadlessWidget, the reference's
GLWidget without Qt (host/headless.cpp lists the commands).

    python -m voxeltoy_b200.headless script.txt [--device N]
"""
import argparse
import sys

from . import host


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("script")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    r = host.Renderer()
    r.initialize("", a.device)                       # GLWidget::initializeGL (glwidget.cpp:47-60)
    r.updateRenderSettings()
    w = host.HeadlessWidget(r)
    try:
        w.run(open(a.script).read())
    except ValueError as e:
        print("headless: %s" % e, file=sys.stderr)
        return 1
    log = r.getLog().strip()
    if log:
        print(log)
    print("paints=%d samples=%d resolution=%dx%d status=%s" % (w.paints(), r.numberSamples(), r.width, r.height, r.getStatus()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
