import sys, os, numpy as np
sys.path.insert(0, ".")
import voxeltoy_b200 as vt
from voxeltoy_b200 import host, scenes
r = host.Renderer(); r.initialize("", 0)
r.loadMesh(os.path.join("tests", "golden", "bunny.obj.gz"), 512)
ctx = r.context()
t = scenes.c3_material_table(); mats = t.array()
ctx.materials_upload(mats); ctx.assign_materials(np.asarray(t.offsets, np.int32), rule=1)
r.setRenderSettings(maxBounces=4)
r.resizeFrame(1920, 1080); r.camera().controller().orbitAroundTarget(np.radians(130), np.radians(25)); r.resetRender()
ctx = r.context(); ctx.set_selection([-1, -1, -1, 0], [1, 0, 0, 0])
r.renderPasses(1); ctx.sync(); ctx.reset_counters()
r.renderPasses(1); ctx.sync()
c = ctx.counters()
print("skip calls", c["cdf_loads"], "successes", c["env_lookups"], "steps skipped", c["material_evals"])
