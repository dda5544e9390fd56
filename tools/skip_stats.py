"""C3 with a -DVT_SKIP_STATS build (VT_LIB_PATH=voxeltoy_b200/variants/stats.so): how many empty-space skips wf_trace attempts,
how many succeed, how many DDA iterations they replace and how many iterations are still stepped one by one."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from voxeltoy_b200 import host, scenes
sys.path.insert(0, "tools")
import run_configs as rc


def main():
    r = host.Renderer(); r.initialize("", 0); ctx = r.context()
    r.loadMesh(os.path.join("tests", "golden", "bunny.obj.gz"), 512)
    t = scenes.c3_material_table(); mats = t.array()
    ctx.materials_upload(mats); ctx.assign_materials(np.asarray(t.offsets, np.int32), rule=1)
    grid = ctx.read_volume(); ctx.emissive_upload(scenes.emissive_list(grid, mats)[::64])
    r.setRenderSettings(maxBounces=4)
    ctx = rc.camera(r, 1920, 1080, 130, 25)
    r.renderPasses(4); ctx.sync()
    ctx.reset_counters(); r.resetRender(); r.renderPasses(4); ctx.sync()
    c = ctx.counters(); n = 1920 * 1080 * 4.0
    print(json.dumps({"per_sample": {"skip_calls": c["cdf_loads"] / n, "skip_ok": c["env_lookups"] / n, "steps_skipped": c["material_evals"] / n,
                                     "plain_steps_in_trace": c["dda_steps"] / n, "skips_ge_32": c["rand_calls"] / n},
                      "avg_skip_len": c["material_evals"] / max(1, c["env_lookups"])}))
    r.close()


if __name__ == "__main__":
    main()
