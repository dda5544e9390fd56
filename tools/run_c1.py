"""BASELINE config 1: scene_fall.vox at 512x512, 1 spp, 1 bounce Lambert, constant environment, pinhole, default camera.
Times the CPU arm (the reference's shaders compiled for the host, oracle/_ref; the C oracle when absent) on all host cores and, with
a GPU, the CUDA path on the same frame -- and checks that the two frames are bit-identical. Also the voxelizer's CPU baseline
(bunny.obj at 512^3). Prints one JSON line. TEST / MEASUREMENT TOOL: not part of the product."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import ref as oref
from oracle import scene as oscene
from oracle import vto
from tests import util


def main():
    cores = os.cpu_count() or 1
    d = util.make_frame(util.scene_fall_volume(), 512, 512, bounces=1, bg="grey")
    s = vto.make_scene(d)
    kind = "reference GLSL compiled for the CPU (oracle/_ref)" if oref.available() else "C oracle (oracle/vto.c)"
    render = (lambda k: oref.render_pass(s, k, cores)) if oref.available() else (lambda k: vto.render_pass(s, k, cores, want_hits=False)[0])
    render(0)
    ts = []
    for _ in range(5):
        t = time.perf_counter(); img = render(0); ts.append(time.perf_counter() - t)
    out = {"config": "C1 scene_fall 512x512, 1 spp, 1 bounce, constant environment, pinhole", "cpu_kind": kind, "cpu_cores": cores,
           "cpu_s_per_pass_median": float(np.median(ts)), "cpu_msamples_per_s": 0.262144 / float(np.median(ts))}
    verts, idx = oscene.load_obj(util.BUNNY)
    bmin, bmax = oscene.mesh_bounds(verts)
    M = oscene.mesh_transform(bmin, bmax, (512, 512, 512))
    vox = oref.voxelize if oref.available() else vto.voxelize
    vox(verts, idx, M, (512, 512, 512))
    tv = []
    for _ in range(5):
        t = time.perf_counter(); vox(verts, idx, M, (512, 512, 512)); tv.append(time.perf_counter() - t)
    out["cpu_voxelize_512_ms_median"] = float(np.median(tv)) * 1e3
    try:
        import voxeltoy_b200 as vt
        ctx = vt.Context(0)
        util.upload(ctx, d)
        ctx.render(0, 1); ctx.sync()
        same = bool(util.same_bits(ctx.read_average(), img).all())
        import torch
        tg = []
        for _ in range(20):
            ctx.reset_accumulation(); ctx.sync()
            t = time.perf_counter(); ctx.render(0, 1); ctx.sync(); tg.append(time.perf_counter() - t)
        # throughput of the same frame when the passes are not launched one by one: 256 spp in one call
        ctx.reset_accumulation(); ctx.render(0, 256); ctx.sync()
        t = time.perf_counter(); ctx.reset_accumulation(); ctx.render(0, 256); ctx.sync(); t256 = time.perf_counter() - t
        out.update({"gpu_s_per_pass_median": float(np.median(tg)), "gpu_msamples_per_s_1spp": 0.262144 / float(np.median(tg)),
                    "gpu_msamples_per_s_256spp": 0.262144 * 256 / t256, "gpu_frame_bit_identical_to_cpu": same})
        ctx.close()
    except Exception as e:          # no GPU here: the CPU half is still reported
        out["gpu"] = "unavailable: %r" % (e,)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
