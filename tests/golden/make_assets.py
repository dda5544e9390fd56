"""Regenerates tests/golden/*.gz: gzip'd copies of the two benchmark INPUT DATA files BASELINE.json names
(resources/scene_fall.vox, resources/bunny.obj). They are data, not source; the GPU box has no /root/reference,
so the bench and the -m gpu tests read these. Run here (where /root/reference exists):
    python tests/golden/make_assets.py
"""
import gzip
import os
import shutil

REF = os.environ.get("VT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

for name in ("scene_fall.vox", "bunny.obj"):
    src = os.path.join(REF, "resources", name)
    dst = os.path.join(HERE, name + ".gz")
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", compresslevel=9, mtime=0) as g:
        shutil.copyfileobj(f, g)
    print(dst, os.path.getsize(dst))
