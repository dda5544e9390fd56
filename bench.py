#!/usr/bin/env python
"""bench.py -- headline benchmark of voxeltoy_b200 (contract: see the task statement / DESIGN.md "Measurement").

Metric (BASELINE.json): path-traced Msamples/s @1080p, 4 bounces. Workload at every N: BASELINE config 2,
`resources/scene_fall.vox` at 1920x1080, 4 bounces, importance-sampled IBL + thin-lens DOF (synthetic HDR
environment). One STEP = the whole job of that config: 256 progressive passes (256 spp) over the frame through
Renderer::renderPasses -> vt_render, which runs them as batches of 64 passes (128 Mi paths in flight), path trace fused with
the running accumulation (--passes changes the step size). With N GPUs the samples are partitioned
(rank r renders sampleCount = p*N + r): per-GPU work is fixed (weak scaling) and each step ends
with an NCCL reduce of the float4 accumulators to rank 0 inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
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

W, H, BOUNCES, PASSES = 1920, 1080, 4, 256
THETA, PHI, FSTOP = 120.0, 30.0, 2.8
WORKLOAD_FMT = "C2: scene_fall.vox 1920x1080, 4 bounces, IBL + thin-lens DOF, %d spp per step (batches of <= 128 Mi paths = 64 passes)"
WORKLOAD = WORKLOAD_FMT % PASSES


# ---------------------------------------------------------------------------------------------------
# scene set-up through the product's host classes (the reference's Renderer API; no oracle code on this path)
# ---------------------------------------------------------------------------------------------------
def setup_renderer(device):
    """Config 2 driven the way the reference's UI drives its Renderer."""
    import tempfile
    from voxeltoy_b200 import host, scenes
    r = host.Renderer()
    r.initialize("", device)
    r.resizeFrame(W, H)
    r.loadVoxFile(os.path.join(ROOT, "tests", "golden", "scene_fall.vox.gz"))
    env_path = os.path.join(tempfile.gettempdir(), "voxeltoy_b200_c2_env_%d.pfm" % os.getpid())
    host.write_pfm(env_path, scenes.synthetic_env(1024, 512))
    r.setRenderSettings(maxBounces=BOUNCES, backgroundImage=env_path)
    cam = r.camera()
    cam.setLensModel(host.CLM_THIN_LENS)
    cam.controller().orbitAroundTarget(np.radians(THETA), np.radians(PHI))
    cam.setFStop(FSTOP)
    r.resetRender()                                                    # as the UI does after camera changes (ui/glwidget.cpp:225-243)
    ctx = r.context()
    ctx.set_selection([-1, -1, -1, 0], [1, 0, 0, 0])                   # no highlighted voxel
    r.requestAction(0.5, 0.5, 0.0, 0.0, host.PA_SELECT_FOCAL_POINT)    # autofocus on the image centre, runs before the next pass
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


# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_rows=None):
    """The CPU arm: the reference's shaders compiled for the host (oracle/_ref) when present, else the C oracle port,
    on all host cores, over a bounded sample of the SAME workload (full-width rows of the C2 frame, 1 pass)."""
    from oracle import refrun
    return refrun.time_c2(W, H, BOUNCES, THETA, PHI, FSTOP, steps=steps, warmup=warmup, rows=sample_rows)


def main():
    global PASSES, WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--passes", type=int, default=PASSES, help="progressive passes (spp) per step; BASELINE config 2 is 256")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    PASSES = max(1, args.passes); WORKLOAD = WORKLOAD_FMT % PASSES
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(3, args.warmup)

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(max(1, args.steps), max(1, args.warmup))
        line = {"impl": "reference", "metric": "path-traced Msamples/s @1080p, 4 bounces", "value": r["value"], "unit": "Msamples/s",
                "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "sample": r["sample"]},
                "cpu_baseline": {"value": r["value"], "unit": "Msamples/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import voxeltoy_b200 as vt

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- voxeltoy_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    r, ctx = setup_renderer(local_rank)
    stream = torch.cuda.Stream()                  # a real (non-default) stream: handle 0 would mean "the context's own"
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    from voxeltoy_b200 import group as vtgroup
    grp = vtgroup.RenderGroup(vtgroup.PART_SAMPLES, rank, world)        # sample partition: per-GPU work fixed (weak scaling)
    grp.apply(r)
    vol = vt.host.load_vox(os.path.join(ROOT, "tests", "golden", "scene_fall.vox.gz"))     # host arrays for the e2e leg

    npx = W * H
    accum = vt.host.device_view(ctx.accum_device_ptr(), (H, W, 4))          # zero-copy torch view of the accumulator
    reduce_buf = torch.empty_like(accum) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2
    pinned_out = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True)

    def reduce_step():
        if world > 1:      # NCCL SUM-reduce of the float4 accumulators to rank 0 + normalisation (voxeltoy_b200/group.py)
            grp.combine(accum, max(1, ctx.num_samples()), dst=0, out=reduce_buf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        r.renderPasses(PASSES)
        reduce_step()
    barrier()
    r.resetRender()
    launches0 = ctx.counters()["kernel_launches"]
    ctx.kernel_timing_enable(True); ctx.kernel_times()             # cudaEvent pairs around every launch of the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xff)                                              # L2 flush between timed iterations (untimed)
        e0, e1 = evs[i]
        k0, k1 = kevs[i]
        e0.record(stream)
        k0.record(stream)
        r.renderPasses(PASSES)                                             # progressive: continues the running average
        k1.record(stream)
        reduce_step()
        e1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.counters()["kernel_launches"] - launches0
    ktimes = ctx.kernel_times(); ctx.kernel_timing_enable(False)
    ms_steps = sum(a.elapsed_time(b) for a, b in evs)
    ms_kernel = sum(a.elapsed_time(b) for a, b in kevs) / args.steps
    t = torch.tensor([ms_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    samples_total = float(npx) * PASSES * args.steps * world
    value = samples_total / (ms_total * 1e-3) / 1e6

    # ---- end to end through the host-facing API: every step hands the scene arrays to the Renderer from host memory
    # (setVoxelData = createVoxelDataTexture: H2D of grid + materials, occupancy rebuild), renders PASSES passes and reads
    # the frame back into pinned host memory.
    e2e_steps = max(3, min(args.steps, 10))
    h2d = vol["grid"].nbytes + vol["materials"].nbytes + vol["emissive"].nbytes + 3 * 64 + 64
    d2h = npx * 16
    barrier()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record(stream)
    for i in range(e2e_steps):
        r.setVoxelData(vol["res"], vol["grid"], vol["materials"], vol["emissive"])
        r.renderPasses(PASSES)
        if world > 1:
            reduce_step()
            if rank == 0:                                                  # the combined frame lives on rank 0 (reduce, dst=0)
                pinned_out.copy_(reduce_buf, non_blocking=True)
            stream.synchronize()
        else:
            r.readAverage(pinned_out)
        _ = float(pinned_out[H // 2, W // 2, 0])
    ee1.record(stream)
    barrier()
    te = torch.tensor([ee0.elapsed_time(ee1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = float(npx) * PASSES * e2e_steps * world / (float(te.item()) * 1e-3) / 1e6

    # ---- algorithmic bytes of the timed launches (counted replay, untimed): same sample indices => identical work
    r.resetRender()
    ctx.counters_enable(True); ctx.reset_counters()
    n_count = min(args.steps, 4)
    for i in range(n_count):
        r.renderPasses(PASSES)
    ctx.sync()
    cnt = ctx.counters(); ctx.counters_enable(False)
    n_samp = float(npx * PASSES * n_count)
    bps = bytes_per_sample(cnt, n_samp)
    peak, peak_src = measured_peak()
    # dominant kernel: wf_trace (the DDA loop of dda.h:38-57). Algorithmic bytes = 4 B per DDA iteration (one
    # 32-bit occupancy/offset word per step), counted by the kernel itself; duration = its launches inside the timed region.
    trace_ms, trace_launches = ktimes["trace"]
    steps_per_step = cnt["dda_steps"] / n_count                       # DDA iterations of one bench step (PASSES passes)
    trace_bytes_per_launch = 4.0 * steps_per_step * args.steps / max(1, trace_launches)
    trace_ms_per_launch = trace_ms / max(1, trace_launches)
    achieved = trace_bytes_per_launch / (trace_ms_per_launch * 1e-3) / 1e9
    step_achieved = bps * npx * PASSES / (ms_kernel * 1e-3) / 1e9     # every kernel of the step, all algorithmic bytes
    # wf_trace is bound by instruction issue, so the ceiling that explains it is the issue rate: one DDA iteration is 27 SASS
    # instructions (cuobjdump of wf_trace_kernel), an SM issues 4 warp instructions per clock
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)
    issue_peak = sms * 4 * sm_hz / 27.0 * 32.0                         # DDA iterations per second with every lane stepping
    issue_achieved = steps_per_step * args.steps / (trace_ms * 1e-3) if trace_ms > 0 else 0.0
    kernel_ms_total = sum(v[0] for v in ktimes.values())
    # L2 read bandwidth of this GPU, measured now (the north star's roofline for this path is L2, not HBM: the grid, the noise
    # table and the CDFs are L2-resident). 48 MiB buffer, 16-byte ld.global.cg, all SMs.
    try:
        l2_gbs = max(ctx.measure_l2_bandwidth(48 << 20, 20) for _ in range(3))
    except Exception:
        l2_gbs = None
    traffic = None
    shade_traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("wf_trace_kernel_dram_bytes_per_launch")
            shade_traffic = tj.get("wf_shade_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    # the runner-up (a near tie with wf_trace): wf_shade, whose algorithmic bytes are everything of SURVEY 8(d) except the
    # DDA words and the accumulator: 16 R + 36 H + 4 E + 64 Q
    shade_ms, shade_launches = ktimes["shade"]
    shade_bytes_step = (16 * cnt["rand_calls"] + 36 * cnt["material_evals"] + 4 * cnt["cdf_loads"] + 64 * cnt["env_lookups"]) / float(n_count)
    shade_bytes_per_launch = shade_bytes_step * args.steps / max(1, shade_launches)
    shade_achieved = shade_bytes_per_launch / (shade_ms / max(1, shade_launches) * 1e-3) / 1e9 if shade_ms > 0 else 0.0

    if rank == 0:
        line = {
            "metric": "path-traced Msamples/s @1080p, 4 bounces", "value": value, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "width": W, "height": H, "bounces": BOUNCES, "passes_per_step": PASSES,
                       "lens": "thin f/2.8", "env": "synthetic HDR 1024x512 + CDF 512x256", "partition": "samples" if world > 1 else "none",
                       "l2": "flushed between timed steps (256 MiB fill)", "wall_s": t_wall},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "wf_trace_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": trace_bytes_per_launch, "ms_per_launch": trace_ms_per_launch,
                         "launches_per_step": trace_launches / float(args.steps),
                         "share_of_step": trace_ms / kernel_ms_total if kernel_ms_total else None,
                         "kernel_ms_per_step": {k: v[0] / args.steps for k, v in ktimes.items()},
                         "step": {"achieved": step_achieved, "frac": step_achieved / peak, "algorithmic_bytes_per_sample": bps,
                                  "ms_per_step_device": ms_kernel},
                         "per_sample": {"S": cnt["dda_steps"] / n_samp, "R": cnt["rand_calls"] / n_samp,
                                        "H": cnt["material_evals"] / n_samp, "E": cnt["cdf_loads"] / n_samp,
                                        "Q": cnt["env_lookups"] / n_samp},
                         "l2": {"peak": l2_gbs, "unit": "GB/s", "how": "measured live: 48 MiB buffer, ld.global.cg 16 B, 148x8 CTAs, best of 3",
                                "trace_frac": (achieved / l2_gbs) if l2_gbs else None, "step_frac": (step_achieved / l2_gbs) if l2_gbs else None},
                         "runner_up": {"kernel": "wf_shade_kernel", "bound": "hbm", "achieved": shade_achieved, "peak": peak, "unit": "GB/s",
                                       "frac": shade_achieved / peak, "traffic": shade_traffic,
                                       "algorithmic_bytes_per_launch": shade_bytes_per_launch,
                                       "ms_per_launch": shade_ms / max(1, shade_launches),
                                       "share_of_step": shade_ms / kernel_ms_total if kernel_ms_total else None},
                         "issue": {"achieved": issue_achieved / 1e9, "peak": issue_peak / 1e9, "unit": "G DDA iterations/s",
                                   "frac": issue_achieved / issue_peak,
                                   "how": "wf_trace: counted DDA iterations / its device time vs SMs x 4 issue slots x SM clock / 27 "
                                          "instructions per iteration x 32 lanes"},
                         "note": "issue-bound, not bandwidth-bound: see profiles/ (issue slots busy, lanes per instruction)"},
        }
        # second half of BASELINE's metric: voxelize ms @512^3 (bunny.obj through the host MeshLoader + GPUVoxelizer path;
        # device time of clear + triangle scatter + derive of the R32I offset grid, cudaEvents inside vt_voxelize)
        try:
            r2 = vt.host.Renderer(); r2.initialize("", local_rank)
            ms = []
            for _ in range(5):
                r2.loadMesh(os.path.join(ROOT, "tests", "golden", "bunny.obj.gz"), 512)
                ms.append(r2.context().last_voxelize_ms())
            # bit grid: template copy (read + write), scatter, read by the sparse offset patch and by the distance-field pass;
            # the int32 entries of EMPTY voxels are cleared lazily, on read-back (the reference never clears them: SURVEY U5)
            vbytes = 512 ** 3 / 8 * 4
            line["voxelize"] = {"metric": "voxelize ms @512^3", "value": min(ms), "unit": "ms", "mesh": "bunny.obj (4968 triangles)",
                                "runs_ms": ms, "algorithmic_bytes": vbytes, "achieved_gbs": vbytes / (min(ms) * 1e-3) / 1e9,
                                "frac_of_hbm_peak": vbytes / (min(ms) * 1e-3) / 1e9 / peak}
            r2.close()
        except Exception as e:
            line["voxelize"] = {"metric": "voxelize ms @512^3", "value": None, "error": repr(e)}
        # environment ingest (RGBA conversion, importance function, CDFs, integral) on the device vs the
        # same chain on one host core (host/image.cpp calculateCDF, the reference's renderer/image.cpp:68-389)
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
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = cpu_reference_run(steps=2, warmup=1)
                line["cpu_baseline"] = {"value": r["value"], "unit": "Msamples/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            except Exception as e:   # the CPU arm is a reported baseline, never a reason to lose the GPU line
                line["cpu_baseline"] = {"value": None, "unit": "Msamples/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
