"""Developer timing script: voxelize bunny.obj at several resolutions (GPU ms = clear + scatter + derive, cudaEvents)
and compare occupancy with the CPU oracle at sizes it finishes quickly. Usage: python tests/devtools/voxelize_bench.py [res ...]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import voxeltoy_b200 as vt
from oracle import scene as oscene
from oracle import vto
from tests import util

ctx = vt.Context(0)
verts, idx = oscene.load_obj(util.BUNNY)
bmin, bmax = oscene.mesh_bounds(verts)
for res in [int(a) for a in sys.argv[1:]] or [64, 256, 512, 1024]:
    M = oscene.mesh_transform(bmin, bmax, (res,) * 3)
    ms = []
    for _ in range(5):
        ctx.voxelize(verts, idx, M, (res,) * 3)
        ms.append(ctx.last_voxelize_ms())
    msg = "%4d^3  gpu %.3f ms (min of 5; first %.3f)" % (res, min(ms), ms[0])
    if res <= 512:
        t = time.perf_counter(); ref = vto.voxelize(verts, idx, M, (res,) * 3); dt = time.perf_counter() - t
        got = ctx.read_volume() >= 0
        msg += "  cpu oracle %.1f ms  solid %d  equal %s" % (dt * 1e3, int(ref.sum()), bool(np.array_equal(got, ref > 0)))
    print(msg, flush=True)
