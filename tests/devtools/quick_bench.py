"""Developer timing script (not the contract bench): times vt_render on a few configurations with wall clock
around vt_sync. Usage: python tests/devtools/quick_bench.py [config ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import voxeltoy_b200 as vt
from voxeltoy_b200 import scenes
from oracle import scene as oscene   # scene construction only (tools/ is developer tooling, not the product)
from tests import util


def run(ctx, name, d, passes, reps=3, counters=True):
    for variant in [int(v) for v in os.environ.get("VT_QB_VARIANTS", "0,2").split(",")]:
        ctx.set_kernel_variant(variant)
        _run(ctx, name + " v%d" % variant, d, passes, reps, counters and variant == 2)


def _run(ctx, name, d, passes, reps=3, counters=True):
    reps = int(os.environ.get("VT_QB_REPS", reps))
    util.upload(ctx, d)
    ctx.render(0, 1); ctx.sync()
    best = 1e9
    for _ in range(reps):
        ctx.reset_accumulation(); ctx.sync()
        t = time.perf_counter(); ctx.render(0, passes); ctx.sync(); best = min(best, time.perf_counter() - t)
    n = d["W"] * d["H"] * passes
    msg = "%-28s %4dx%-4d b=%d passes=%d  %.2f ms/pass  %.1f Msamples/s" % (name, d["W"], d["H"], d["max_bounces"], passes, best / passes * 1e3, n / best / 1e6)
    if counters:
        ctx.counters_enable(True); ctx.reset_counters(); ctx.reset_accumulation(); ctx.render(0, 1); c = ctx.counters(); ctx.counters_enable(False)
        px = d["W"] * d["H"]
        bps = (4 * c["dda_steps"] + 16 * c["rand_calls"] + 36 * c["material_evals"] + 4 * c["cdf_loads"] + 64 * c["env_lookups"]) / px + 32
        msg += "  S=%.1f R=%.2f H=%.2f E=%.1f Q=%.2f  B/sample=%.0f  -> %.1f GB/s" % (c["dda_steps"] / px, c["rand_calls"] / px, c["material_evals"] / px, c["cdf_loads"] / px, c["env_lookups"] / px, bps, bps * n / best / 1e9)
    print(msg, flush=True)


def main():
    which = sys.argv[1:] or ["c1", "c2", "c2s", "mixed"]
    ctx = vt.Context(0)
    vol = util.scene_fall_volume()
    if "c1" in which:
        run(ctx, "C1 fall 512 pinhole", util.make_frame(vol, 512, 512, bounces=1, bg="grey"), 64)
    if "c2s" in which:
        run(ctx, "fall 1080p sky pinhole", util.make_frame(vol, 1920, 1080, bounces=4, theta=120, phi=30), 16)
    if "c2" in which:
        env = oscene.build_env(scenes.synthetic_env(1024, 512))
        d = util.make_frame(vol, 1920, 1080, bounces=4, theta=120, phi=30, lens_model=1, fstop=2.8, env=env, focal_distance=850.0)
        run(ctx, "C2 fall 1080p IBL thinlens", d, 16)
    if "mixed" in which:
        run(ctx, "mixed 48^3 1080p b=5", util.make_frame(util.mixed_scene(), 1920, 1080, bounces=5, theta=115, phi=40), 8)
    if "noise" in which:
        n = 512
        ids = scenes.dense_noise_grid(n, density=0.02)
        t = scenes.MaterialTable()
        for k in range(8):
            t.lambert((0.3 + 0.08 * k, 0.5, 0.9 - 0.08 * k))
        voln = dict(res=(n, n, n), grid=scenes.ids_to_offsets(ids, t.offsets), materials=t.array(), emissive=np.zeros(0, np.int32))
        run(ctx, "noise 512^3 2% 1080p b=4", util.make_frame(voln, 1920, 1080, bounces=4, theta=115, phi=40), 4)


if __name__ == "__main__":
    main()
