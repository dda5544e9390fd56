#!/usr/bin/env python
"""bench.py -- headline benchmark of voxeltoy_b200 (contract: see the task statement / DESIGN.md "Measurement").

Metric (BASELINE.json): path-traced Msamples/s @1080p, 4 bounces. Default workload at every N = BASELINE config 2:
`resources/scene_fall.vox` at 1920x1080, 4 bounces, importance-sampled IBL + thin-lens DOF (synthetic HDR environment,
SURVEY 8d). One STEP = the whole job of that config: 256 progressive passes (256 spp) through Renderer::renderPasses ->
vt_render (batches of 64 passes = 128 Mi paths in flight), path trace fused with the running accumulation.

With N GPUs (one process per GPU under torchrun) the ranks form a render group behind the C ABI (vt_group_join: NCCL bound by
the library itself; torch.distributed only carries the 128-byte NCCL id and the max-over-ranks of the timings):

  --config c2 --scaling weak    (default) sample partition, every rank renders its own 256 sample indices per step; the step
                                ends with the NCCL SUM-reduce of the float4 accumulators to rank 0 inside the timed region
  --config c2 --scaling strong  the fixed 256-spp job split over the ranks (256 / N passes each) + the same reduce
  --config c4                   BASELINE config 4: 256^3 terrain, 3840x2160, 8 bounces, 32 passes, 64x64 tiles round-robin
                                over the ranks (strong scaling by construction), compact tile gather to rank 0
  --config c5                   BASELINE config 5: dense 1024^3 noise grid, 4K, 16 bounces, 16 passes per GPU and step, sample
                                partition, SUM-reduce every 4 passes, every 8 passes a scripted edit (pick -> add / remove)
                                issued on rank 0, broadcast to all ranks over NCCL, accumulation reset

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c4|c5] [--scaling weak|strong]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c2": dict(W=1920, H=1080, bounces=4, passes=256, theta=120.0, phi=30.0, fstop=2.8, partition="samples",
               workload="C2: scene_fall.vox 1920x1080, 4 bounces, IBL + thin-lens DOF, %d spp per step (batches of <= 128 Mi paths = 64 passes)",
               lens="thin f/2.8", env="synthetic HDR 1024x512 + CDF 512x256"),
    "c4": dict(W=3840, H=2160, bounces=8, passes=32, theta=140.0, phi=35.0, fstop=16.0, partition="tiles",
               workload="C4: procedural terrain 256^3 (rock / metal band / emissive lava), 3840x2160, 8 bounces, %d passes per step, 64x64 tiles round-robin",
               lens="pinhole", env="gradient sky"),
    "c5": dict(W=3840, H=2160, bounces=16, passes=16, theta=125.0, phi=40.0, fstop=16.0, partition="samples",
               workload="C5: dense noise 1024^3 (35 %% solid, 8 materials), 3840x2160, 16 bounces, %d passes per GPU and step, reduce every 4 passes, "
                        "scripted edit (pick -> add / remove) every 8 passes",
               lens="pinhole", env="gradient sky"),
}
METRIC = "path-traced Msamples/s @1080p, 4 bounces"


def config_dict(name, passes, scaling, world):
    """The `config` object of the JSON line: identical in the GPU arm and in the CPU (reference) arm."""
    c = CONFIGS[name]
    return {"workload": c["workload"] % passes, "config": name, "width": c["W"], "height": c["H"], "bounces": c["bounces"],
            "passes_per_step": passes, "lens": c["lens"], "env": c["env"],
            "partition": c["partition"] if world > 1 else "none", "scaling": scaling,
            "l2": "flushed between timed steps (256 MiB fill)"}


# ---------------------------------------------------------------------------------------------------
# scene set-up through the product's host classes (the reference's Renderer API; no oracle code on this path)
# ---------------------------------------------------------------------------------------------------
def scene_arrays(name, device):
    """Host arrays of the config's scene (res, grid, materials, emissive) for Renderer::setVoxelData; None for file scenes."""
    from voxeltoy_b200 import host, scenes
    if name == "c4":
        s = scenes.c4_scene(256)
        em = host.prune_interior_emissive(s["grid"], s["res"], s["emissive_all"])
        return dict(res=s["res"], grid=s["grid"], materials=s["materials"], emissive=em)
    if name == "c5":
        t = scenes.c5_material_table()
        grid = scenes.dense_noise_offsets_torch(1024, t.offsets, device="cuda:%d" % device)
        return dict(res=(1024, 1024, 1024), grid=grid, materials=t.array(), emissive=np.zeros(0, np.int32))
    return None


def setup_renderer(name, device, arrays=None):
    """The config driven the way the reference's UI drives its Renderer."""
    import tempfile
    from voxeltoy_b200 import host, scenes
    c = CONFIGS[name]
    r = host.Renderer()
    r.initialize("", device)
    r.resizeFrame(c["W"], c["H"])
    if name == "c2":
        r.loadVoxFile(os.path.join(ROOT, "tests", "golden", "scene_fall.vox.gz"))
        env_path = os.path.join(tempfile.gettempdir(), "voxeltoy_b200_c2_env_%d.pfm" % os.getpid())
        host.write_pfm(env_path, scenes.synthetic_env(1024, 512))
        r.setRenderSettings(maxBounces=c["bounces"], backgroundImage=env_path)
    else:
        r.setVoxelData(arrays["res"], arrays["grid"], arrays["materials"], arrays["emissive"])
        r.setRenderSettings(maxBounces=c["bounces"])
    cam = r.camera()
    if name == "c2":
        cam.setLensModel(host.CLM_THIN_LENS)
    cam.controller().orbitAroundTarget(np.radians(c["theta"]), np.radians(c["phi"]))
    cam.setFStop(c["fstop"])
    r.resetRender()                                                    # as the UI does after camera changes (ui/glwidget.cpp:225-243)
    ctx = r.context()
    ctx.set_selection([-1, -1, -1, 0], [1, 0, 0, 0])                   # no highlighted voxel
    if name == "c2":
        r.requestAction(0.5, 0.5, 0.0, 0.0, host.PA_SELECT_FOCAL_POINT)   # autofocus on the image centre, runs before the next pass
    r.renderPasses(1)
    r.resetRender()
    return r, ctx


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def bytes_per_sample(c, n_samples):
    """4S + 16R + 36H + 4E + 64Q + 32 algorithmic bytes per sample."""
    return (4 * c["dda_steps"] + 16 * c["rand_calls"] + 36 * c["material_evals"] + 4 * c["cdf_loads"]
            + 64 * c["env_lookups"]) / float(n_samples) + 32.0


# algorithmic bytes of one kernel's share of SURVEY 8(d), from the work counters of a counted replay
KERNEL_BYTES = {
    "trace": ("wf_trace_kernel", "4 B per DDA iteration (one 32-bit occupancy / offset word per step)", lambda c: 4.0 * c["dda_steps"]),
    "shade": ("wf_shade_kernel", "16 R + 36 H + 4 E + 64 Q (noise texels, material records, CDF texels, environment lookups)",
              lambda c: 16.0 * c["rand_calls"] + 36.0 * c["material_evals"] + 4.0 * c["cdf_loads"] + 64.0 * c["env_lookups"]),
    "generate": ("wf_generate_kernel", "16 B noise texel per primary path", lambda c: 16.0 * c["paths"]),
    "accumulate": ("wf_accumulate_kernel", "32 B accumulator read-modify-write per sample", lambda c: 32.0 * c["paths"]),
}


# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_rows=None):
    """The CPU arm: the reference's shaders compiled for the host (oracle/_ref) when present, else the C oracle port,
    on all host cores, over a bounded sample of the SAME workload (full-width rows of the C2 frame, 1 pass)."""
    from oracle import refrun
    c = CONFIGS["c2"]
    return refrun.time_c2(c["W"], c["H"], c["bounces"], c["theta"], c["phi"], c["fstop"], steps=steps, warmup=warmup, rows=sample_rows)


def run_config(args, name, scaling, passes, steps, warmup, rank, world, local_rank, dist, full=True):
    """Times `steps` steps of one config on this process's GPU (as rank `rank` of `world`). Returns the result dict on rank 0."""
    import torch
    import voxeltoy_b200 as vt
    from voxeltoy_b200 import group as vtgroup
    from voxeltoy_b200 import host as vthost

    c = CONFIGS[name]
    W, H = c["W"], c["H"]
    npx = W * H
    arrays = scene_arrays(name, local_rank)
    r, ctx = setup_renderer(name, local_rank, arrays)
    stream = torch.cuda.Stream()                  # a real (non-default) stream: handle 0 would mean "the context's own"
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    mode = vtgroup.PART_TILES if c["partition"] == "tiles" else vtgroup.PART_SAMPLES
    grp = None
    if world > 1:
        # the 128-byte NCCL id travels over torch.distributed (plumbing); the communicator and every collective on the data path
        # belong to libvoxeltoy_b200.so (csrc/vt_group.inl)
        box = [vtgroup.DeviceGroup.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        grp = vtgroup.DeviceGroup.join(ctx, box[0], rank, world, mode)
    # passes this rank renders per step
    if name == "c2" and scaling == "strong" and world > 1:
        local_passes = max(1, passes // world)
        samples_per_step = float(npx) * local_passes * world
    elif mode == vtgroup.PART_TILES:
        local_passes = passes
        samples_per_step = float(npx) * passes                       # the frame is split, not the samples
    else:
        local_passes = passes
        samples_per_step = float(npx) * passes * world
    reduce_every = 4 if name == "c5" else local_passes
    edit_every = 8 if name == "c5" else 0
    edits = {"n": 0}

    split = {"render": [], "combine": []}       # (start, end) event pairs of the timed steps

    def combine(record=True):
        if grp is not None:
            if record:
                a = torch.cuda.Event(enable_timing=True); a.record(stream)
            grp.begin_combine()
            grp.wait_combine()                    # stream-level: the context's stream (and the timing events) come after the exchange
            if record:
                b = torch.cuda.Event(enable_timing=True); b.record(stream)
                split["combine"].append((a, b))

    def scripted_edit():
        """pick at a fixed pixel -> add on the picked face (even edits) / remove the picked voxel (odd edits); issued on rank 0,
        the 32-byte action record reaches every rank through vt_group_broadcast, every replica runs it and resets its accumulation."""
        k = edits["n"]; edits["n"] += 1
        rec = np.array([0.5, 0.5, 0.0, 0.0, float(vthost.PA_ADD_VOXEL if k % 2 == 0 else vthost.PA_REMOVE_VOXEL), 1.0, 0.0, 0.0], np.float32)
        if rank != 0:
            rec[:] = 0
        if grp is not None:
            rec = grp.broadcast(rec, root=0)
        r.requestAction(float(rec[0]), float(rec[1]), 0.0, 0.0, vthost.PA_SELECT_ACTIVE_VOXEL)
        r.requestAction(float(rec[0]), float(rec[1]), float(rec[2]), float(rec[3]), int(rec[4]))

    def step(record=True):
        done = 0
        while done < local_passes:
            if edit_every and done > 0 and done % edit_every == 0:
                scripted_edit()
            n = min(reduce_every, local_passes - done)
            if record:
                a = torch.cuda.Event(enable_timing=True); a.record(stream)
            r.renderPasses(n)                                         # progressive: continues the running average / sum
            if record:
                b = torch.cuda.Event(enable_timing=True); b.record(stream)
                split["render"].append((a, b))
            done += n
            combine(record)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2
    pinned_out = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True)
    # ---- multi-GPU check (before any scripted edit touches the volume): the combined frame of a short render against one
    # context rendering the same sample indices alone
    check = None
    if world > 1:
        pc = 4
        r.resetRender()
        r.renderPasses(pc)
        grp.begin_combine()
        frame = grp.end_combine(want=(rank == 0))
        if rank == 0:
            r1, c1 = setup_renderer(name, local_rank, arrays)
            r1.renderPasses(pc if mode == vtgroup.PART_TILES else pc * world)
            want = c1.read_average()
            if mode == vtgroup.PART_TILES:
                same = bool(((frame.view(np.uint32) == want.view(np.uint32)) | (np.isnan(frame) & np.isnan(want))).all())
                check = "ok (tiles: %d-rank frame bit-identical to 1 context)" % world if same else "MISMATCH (tiles differ from 1 context)"
            else:
                fin = np.isfinite(want) & np.isfinite(frame)
                err = float(np.max(np.abs(frame[fin] - want[fin]) / (np.abs(want[fin]) + 1e-3))) if fin.any() else 0.0
                ok = bool(np.allclose(frame, want, rtol=1e-5, atol=1e-6, equal_nan=True))
                check = ("ok" if ok else "MISMATCH") + " (samples: %d x %d passes vs 1 context x %d passes, max rel err %.2e, rtol 1e-5)" % (world, pc, pc * world, err)
            r1.close()
        barrier()
    for i in range(warmup):
        step(record=False)
    barrier()

    del arrays

    r.resetRender()
    launches0 = ctx.counters()["kernel_launches"]
    ctx.kernel_timing_enable(True); ctx.kernel_times()             # cudaEvent pairs around every launch of the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_wall = time.perf_counter()
    ex_ms = []
    for i in range(steps):
        flush.fill_(i & 0xff)                                              # L2 flush between timed iterations (untimed)
        e0, e1 = evs[i]
        e0.record(stream)
        step()
        e1.record(stream)
        if grp is not None:                                                # device time of the step's last exchange (root's side stream)
            stream.synchronize(); grp.sync()
            ex_ms.append(grp.last_exchange_ms())
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.counters()["kernel_launches"] - launches0
    ktimes = ctx.kernel_times(); ctx.kernel_timing_enable(False)
    ms_steps = sum(a.elapsed_time(b) for a, b in evs)
    kernel_ms_total = sum(v[0] for v in ktimes.values())
    t = torch.tensor([ms_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = samples_per_step * steps / (ms_total * 1e-3) / 1e6
    ms_render = sum(x.elapsed_time(y) for x, y in split["render"]) / steps
    ms_combine = sum(x.elapsed_time(y) for x, y in split["combine"]) / steps

    per_rank = None
    if world > 1:
        # where a step's time goes on every rank: its own kernels, its render calls (kernels + launch gaps), the exchange + the wait
        # for the slowest rank, and the step as a whole
        mine = {"rank": rank, "ms_step": ms_steps / steps, "ms_render": ms_render, "ms_combine_incl_wait": ms_combine, "ms_kernels": kernel_ms_total / steps}
        box = [None] * world
        dist.all_gather_object(box, mine)
        per_rank = box
    out = {"value": value, "ms_per_step": ms_total / steps, "ms_render_per_step": ms_render, "ms_combine_per_step_incl_wait": ms_combine, "per_rank": per_rank, "wall_s": t_wall, "clocks": clocks, "gpu_launches": int(launches),
           "kernel_ms_per_step": {k: v[0] / steps for k, v in ktimes.items()}, "multi_gpu_check": check,
           "local_passes_per_step": local_passes, "edits_per_step": (edits["n"] / float(steps + warmup)) if edit_every else 0}
    if grp is not None:
        per_step = (local_passes + reduce_every - 1) // reduce_every
        out["collective_ms"] = {"per_exchange_device": float(np.mean(ex_ms)) if ex_ms else None, "per_exchange_device_each_step": [float(x) for x in ex_ms], "exchanges_per_step": per_step,
                                "bytes_to_root_per_exchange": grp.exchange_bytes(), "exchange": "nccl" if grp.exchange() == 0 else "peer",
                                "how": "cudaEvents on the root's side stream around snapshot -> %s -> normalise" %
                                       ("pack tiles + ncclSend/ncclRecv gather + unpack" if mode == vtgroup.PART_TILES else "ncclReduce(sum)"),
                                "share_of_step": (float(np.mean(ex_ms)) * per_step / (ms_total / steps)) if ex_ms else None}
    if not full:
        if grp is not None:
            grp.close()
        r.close()
        return out

    # ---- end to end through the host-facing API: every step hands the scene arrays to the Renderer from host memory
    # (setVoxelData = createVoxelDataTexture: H2D of grid + materials, occupancy rebuild), renders the step and reads the
    # combined frame back into pinned host memory.
    vol = vt.host.load_vox(os.path.join(ROOT, "tests", "golden", "scene_fall.vox.gz")) if name == "c2" else None
    e2e = None
    if vol is not None:
        e2e_steps = max(3, min(steps, 10))
        h2d = vol["grid"].nbytes + vol["materials"].nbytes + vol["emissive"].nbytes + 3 * 64 + 64
        d2h = npx * 16
        barrier()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record(stream)
        for i in range(e2e_steps):
            r.setVoxelData(vol["res"], vol["grid"], vol["materials"], vol["emissive"])
            r.renderPasses(local_passes)
            if grp is not None:
                grp.begin_combine()
                grp.end_combine(out=pinned_out.numpy() if rank == 0 else None, want=(rank == 0))
            else:
                r.readAverage(pinned_out)
            _ = float(pinned_out[H // 2, W // 2, 0])
        ee1.record(stream)
        barrier()
        te = torch.tensor([ee0.elapsed_time(ee1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": samples_per_step * e2e_steps / (float(te.item()) * 1e-3) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps}
    out["e2e"] = e2e

    # ---- algorithmic bytes of the timed launches (counted replay, untimed): same sample indices => identical work
    r.resetRender()
    ctx.counters_enable(True); ctx.reset_counters()
    n_count = min(steps, 2)
    for i in range(n_count):
        r.renderPasses(local_passes)
    ctx.sync()
    cnt = ctx.counters()
    # the same sample indices again with no bounce: what is left is the primary traversal, which wf_generate performs; the
    # difference is the work of wf_trace alone
    primary_steps = 0
    try:
        r.setRenderSettings(maxBounces=0)
        r.resetRender(); ctx.reset_counters()
        for i in range(n_count):
            r.renderPasses(local_passes)
        ctx.sync()
        primary_steps = ctx.counters()["dda_steps"]
    finally:
        ctx.counters_enable(False)
        r.setRenderSettings(maxBounces=c["bounces"]); r.resetRender()
    share = (1.0 / world) if mode == vtgroup.PART_TILES and world > 1 else 1.0
    n_samp = float(npx) * local_passes * n_count * share               # samples THIS rank rendered in the replay
    cnt = dict(cnt); cnt["paths"] = n_samp
    bps = bytes_per_sample(cnt, n_samp)
    peak, peak_src = measured_peak()
    # the dominant kernel is whichever took the most device time in the timed region (cudaEvent pairs around every launch)
    dom = max((k for k in ktimes if k in KERNEL_BYTES), key=lambda k: ktimes[k][0])
    order = sorted((k for k in ktimes if k in KERNEL_BYTES), key=lambda k: -ktimes[k][0])

    def kernel_roofline(k):
        kname, what, fn = KERNEL_BYTES[k]
        ms, n_launch = ktimes[k]
        bytes_step = fn(cnt) / n_count                                  # algorithmic bytes of one bench step on this rank
        per_launch = bytes_step * steps / max(1, n_launch)
        ach = per_launch / (ms / max(1, n_launch) * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "algorithmic_bytes": what, "algorithmic_bytes_per_launch": per_launch, "ms_per_launch": ms / max(1, n_launch),
                "launches_per_step": n_launch / float(steps), "share_of_step": ms / kernel_ms_total if kernel_ms_total else None}

    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp))
        except Exception:
            traffic = {}
    roof = kernel_roofline(dom)
    roof["traffic"] = traffic.get(KERNEL_BYTES[dom][0] + "_dram_bytes_per_launch")
    roof["traffic_source"] = traffic.get("source")
    roof["peak_source"] = peak_src
    roof["kernel_ms_per_step"] = out["kernel_ms_per_step"]
    step_achieved = bps * n_samp / n_count / (kernel_ms_total / steps * 1e-3) / 1e9
    roof["step"] = {"achieved": step_achieved, "frac": step_achieved / peak, "algorithmic_bytes_per_sample": bps,
                    "ms_per_step_kernels": kernel_ms_total / steps}
    roof["per_sample"] = {"S": cnt["dda_steps"] / n_samp, "R": cnt["rand_calls"] / n_samp, "H": cnt["material_evals"] / n_samp,
                          "E": cnt["cdf_loads"] / n_samp, "Q": cnt["env_lookups"] / n_samp}
    if len(order) > 1:
        ru = kernel_roofline(order[1])
        ru["traffic"] = traffic.get(KERNEL_BYTES[order[1]][0] + "_dram_bytes_per_launch")
        roof["runner_up"] = ru
    # L2 read bandwidth of this GPU, measured now (the north star quotes this path against the L2 roofline)
    try:
        l2_gbs = max(ctx.measure_l2_bandwidth(48 << 20, 20) for _ in range(3))
    except Exception:
        l2_gbs = None
    roof["l2"] = {"peak": l2_gbs, "unit": "GB/s", "how": "measured live: 48 MiB buffer, ld.global.cg 16 B, 148x8 CTAs, best of 3",
                  "kernel_frac": (roof["achieved"] / l2_gbs) if l2_gbs else None, "step_frac": (step_achieved / l2_gbs) if l2_gbs else None}
    # wf_trace is bound by instruction issue: one DDA iteration is 25 SASS instructions (cuobjdump), an SM
    # issues 4 warp instructions per clock
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)
    issue_peak = sms * 4 * sm_hz / 25.0 * 32.0

    def issue(kernel, steps_counted, ms, how):
        ach = steps_counted / n_count * steps / (ms * 1e-3) if ms > 0 else 0.0
        return {"kernel": kernel, "achieved": ach / 1e9, "peak": issue_peak / 1e9, "unit": "G DDA iterations/s", "frac": ach / issue_peak, "how": how}
    roof["issue"] = issue("wf_trace_kernel", cnt["dda_steps"] - primary_steps, ktimes["trace"][0],
                          "DDA iterations of the shadow + bounce rays (counted replay minus a counted 0-bounce replay of the same samples) / device time "
                          "of wf_trace vs SMs x 4 issue slots x SM clock / 25 instructions per iteration x 32 lanes")
    roof["issue_primary"] = issue("wf_generate_kernel", primary_steps, ktimes.get("generate", (0.0, 0))[0],
                                  "DDA iterations of the primary rays / device time of wf_generate (which also sets the rays up: conservative), same ceiling")
    roof["note"] = "both big kernels are issue / latency bound, not bandwidth bound: see profiles/ (issue slots busy, lanes per instruction)"
    out["roofline"] = roof
    if grp is not None:
        grp.close()
    r.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--passes", type=int, default=0, help="progressive passes (spp) per step; 0 = the config's own (C2: 256)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N > 1: skip the short strong-scaling C2 and tiled C4 runs appended to the default line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(3, args.warmup)
    name = args.config
    passes = args.passes if args.passes > 0 else CONFIGS[name]["passes"]
    scaling = "strong" if (name == "c4" or (name == "c2" and args.scaling == "strong")) else "weak"
    cfg = config_dict(name, passes, scaling, max(world, args.gpus))

    if args.impl == "reference":
        if rank != 0:
            return 0
        if name != "c2":
            print(json.dumps({"impl": "reference", "unavailable": "the CPU arm times BASELINE config 2 (the metric's config) only"}))
            return 0
        r = cpu_reference_run(max(1, args.steps), max(1, args.warmup))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Msamples/s",
                "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg,
                "cpu_baseline": {"value": r["value"], "unit": "Msamples/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- voxeltoy_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    res = run_config(args, name, scaling, passes, args.steps, warmup, rank, world, local_rank, dist, full=True)
    extras = {}
    if world > 1 and name == "c2" and scaling == "weak" and not args.no_extras:
        # the demanding multi-GPU cases next to the default (weak) line: the fixed 256-spp job split over the ranks, and config 4's
        # tile partition at 4K. Short runs; the default line above stays comparable with the 1-GPU bench.
        for key, (n2, s2, p2, k2) in {"c2_strong": ("c2", "strong", CONFIGS["c2"]["passes"], 5), "c4_tiles": ("c4", "strong", CONFIGS["c4"]["passes"], 3)}.items():
            try:
                x = run_config(args, n2, s2, p2, k2, 3, rank, world, local_rank, dist, full=False)
                extras[key] = {"value": x["value"], "unit": "Msamples/s", "ms_per_step": x["ms_per_step"], "steps": k2, "scaling": "strong",
                               "config": config_dict(n2, p2, s2, world), "collective_ms": x.get("collective_ms"),
                               "multi_gpu_check": x["multi_gpu_check"], "kernel_ms_per_step": x["kernel_ms_per_step"],
                               "ms_render_per_step": x["ms_render_per_step"], "ms_combine_per_step_incl_wait": x["ms_combine_per_step_incl_wait"],
                               "per_rank": x.get("per_rank")}
            except Exception as e:                                   # never lose the main line
                extras[key] = {"value": None, "error": repr(e)}

    if rank == 0:
        import voxeltoy_b200 as vt
        roof = res.get("roofline")
        line = {
            "metric": METRIC, "value": res["value"], "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "wall_s": res["wall_s"],
            "clocks": res["clocks"],
            "e2e": res.get("e2e"),
            "gpu_launches": res["gpu_launches"],
            "roofline": roof,
        }
        if world > 1:
            line["collective_ms"] = res.get("collective_ms")
            line["ms_render_per_step"] = res["ms_render_per_step"]
            line["ms_combine_per_step_incl_wait"] = res["ms_combine_per_step_incl_wait"]    # includes waiting for the slowest rank to arrive
            line["multi_gpu_check"] = res["multi_gpu_check"]
            line["per_rank"] = res.get("per_rank")
            if extras:
                line["extras"] = extras
        if name == "c5":
            line["edits_per_step"] = res["edits_per_step"]
        peak = roof["peak"] if roof else measured_peak()[0]
        if name == "c2":
            # second half of BASELINE's metric: voxelize ms @512^3 (bunny.obj through the host MeshLoader + GPUVoxelizer path)
            try:
                r2 = vt.host.Renderer(); r2.initialize("", local_rank)
                ms, ms_full = [], []
                for _ in range(5):
                    r2.loadMesh(os.path.join(ROOT, "tests", "golden", "bunny.obj.gz"), 512)
                    ms.append(r2.context().last_voxelize_ms()); ms_full.append(r2.context().last_voxelize_full_ms())
                vbytes = 512 ** 3 / 8 * 4
                line["voxelize"] = {"metric": "voxelize ms @512^3", "value": min(ms), "unit": "ms", "mesh": "bunny.obj (4968 triangles)",
                                    "runs_ms": ms, "algorithmic_bytes": vbytes, "achieved_gbs": vbytes / (min(ms) * 1e-3) / 1e9,
                                    "frac_of_hbm_peak": vbytes / (min(ms) * 1e-3) / 1e9 / peak,
                                    "timed": "cudaEvents inside vt_voxelize: clear of the bit grid (copy of the sentinel template) + triangle scatter (atomicOr)",
                                    "value_full": min(ms_full), "runs_full_ms": ms_full,
                                    "timed_full": "the same + the material-id grid made valid (128 MiB cleared, solid voxels filled) + the renderer's distance fields"}
                r2.close()
            except Exception as e:
                line["voxelize"] = {"metric": "voxelize ms @512^3", "value": None, "error": repr(e)}
            # environment ingest on the device vs the same chain on one host core
            try:
                c3 = vt.Context(local_rank)
                env_rgb = vt.scenes.synthetic_env(1024, 512)
                ems = []
                for _ in range(4):
                    c3.env_build(env_rgb); ems.append(c3.env_info()["build_ms"])
                t0 = time.perf_counter(); vt.host.calculate_cdf(env_rgb); host_ms = (time.perf_counter() - t0) * 1e3
                line["env_build"] = {"metric": "env ingest ms (1024x512 RGB -> RGBA + 512x256 CDFs)", "value": min(ems), "unit": "ms",
                                     "runs_ms": ems, "host_calculateCDF_ms_1_core": host_ms}
                c3.close()
            except Exception as e:
                line["env_build"] = {"value": None, "error": repr(e)}
        if world == 1 and name == "c2" and not args.no_cpu_baseline:
            try:
                rr = cpu_reference_run(steps=2, warmup=1)
                line["cpu_baseline"] = {"value": rr["value"], "unit": "Msamples/s", "cores": rr["cores"], "kind": rr["kind"], "sample": rr["sample"]}
            except Exception as e:   # the CPU arm is a reported baseline, never a reason to lose the GPU line
                line["cpu_baseline"] = {"value": None, "unit": "Msamples/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
