"""TEST INFRASTRUCTURE / CPU baseline arm: times the reference's algorithm for BASELINE config 2 on the host cores.

kind "reference": oracle/_ref/libvt_ref.so -- the reference's own GLSL shaders compiled for the CPU (oracle/shim) --
when it has been built; kind "port": the C oracle (oracle/vto.c). Used only by bench.py (--impl reference and the
cpu_baseline leg). One step = one full-frame pass of the C2 frame (W*H samples), all host threads.
"""
import os
import time

import numpy as np

from . import scene as oscene
from . import vto

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c2_scene(W, H, bounces, theta, phi, fstop, env_size=(1024, 512)):
    """The same C2 set-up bench.py builds through the product's Renderer, restated with the oracle's own host code."""
    from voxeltoy_b200 import scenes          # closed-form synthetic environment (an input, not an algorithm under test)
    vol = oscene.load_vox(os.path.join(ROOT, "tests", "golden", "scene_fall.vox.gz"))
    env = oscene.build_env(scenes.synthetic_env(*env_size))
    bmin, bmax, _ = vto.volume_bounds(*vol["res"])
    cam = oscene.Camera()
    cam.set_distance_from_target(100.0)
    cam.set_fstop(16.0)
    cam.matrices(W, H)                        # resizeFrame: film size for this aspect
    diag = np.float32(np.sqrt(np.sum((bmax - bmin).astype(np.float32) ** 2, dtype=np.float32)))
    cam.set_distance_from_target(diag * np.float32(0.5))
    cam.lens_model = 1
    cam.orbit_around_target(np.radians(theta), np.radians(phi))
    cam.set_fstop(fstop)
    _, imv, pm, ipm = cam.matrices(W, H)
    d = dict(vol)
    d.update(W=W, H=H, inv_modelview=imv, proj=pm, inv_proj=ipm, max_bounces=bounces, lens_model=1,
             lens_radius=float(cam.lens_radius), env=env, near_z=float(cam.near),
             bg_top=[153.0 / 255 * 2, 187.0 / 255 * 2, 201.0 / 255 * 2], bg_bottom=[77.0 / 255, 64.0 / 255, 50.0 / 255],
             sel_index=(-1, -1, -1))
    s = vto.make_scene(d)
    d["focal_distance"] = vto.pick_focal(s, W * 0.5, H * 0.5)
    return d


def time_c2(W, H, bounces, theta, phi, fstop, steps=2, warmup=1, rows=None):
    d = c2_scene(W, H, bounces, theta, phi, fstop)
    cores = os.cpu_count() or 1
    kind = "port"
    render = None
    try:
        from . import ref as oref
        if oref.available():
            rs = oref.make_scene(d)
            render = lambda k: oref.render_pass(rs, k, cores)
            kind = "reference"
    except Exception:
        render = None
    if render is None:
        s = vto.make_scene(d)
        render = lambda k: vto.render_pass(s, k, cores, want_hits=False)
    for k in range(warmup):
        render(k)
    t = time.perf_counter()
    for k in range(steps):
        render(warmup + k)
    dt = time.perf_counter() - t
    n = float(W) * H * steps
    return dict(value=n / dt / 1e6, ms_per_step=dt / steps * 1e3, steps=steps, warmup=warmup, cores=cores, kind=kind,
                sample="%d full-frame passes of the C2 frame (%dx%d, %d bounces, IBL + thin lens), 1 pass per step, %d host threads"
                       % (steps, W, H, bounces, cores))
