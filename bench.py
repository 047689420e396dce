#!/usr/bin/env python
"""bench.py — throughput of the CUDA path tracer on the reference's headline workload.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the configuration the north-star target is quoted on):
AccPathTracer semantics on the Stanford bunny (5k triangles) inside the Cornell box, 1920x1080,
1024 spp, depth 20, aspect 16/9 — SURVEY.md §8(d) cfg3.  A "step" renders that whole frame.
With N GPUs the 1024 samples of every pixel are split into N sample slices (strong scaling), the
partial linear frames are combined with one NCCL reduce, rank 0 resolves (÷spp, sqrt) the frame
(nrenderer_b200/multigpu.py).

One JSON line on rank 0:
  value      Mpath-samples/s, whole job, scene resident in HBM, device-timed (CUDA events, max over ranks)
  e2e        the same metric THROUGH THE DROP-IN: wall time around RenderComponent::exec of the registered plugin
             (nr_headless -> ComponentFactory::createComponent -> exec -> NRCuda::Adapter::render -> nrcu_* -> Screen::set),
             i.e. Scene flattening, H2D, BVH build, render, D2H into the adapter's pageable RGBA buffer and the
             Screen::set copy, every step; with N GPUs the plugin runs with NRCU_DEVICES=N (nrcu_render_multi)
  roofline   the binding roofline of the step: SM issue slots.  achieved = warp instructions per path sample (ncu launch
             list of this very code, profiles/) x paths of the timed region / its CUDA-event time; peak = SMs x 4 x the SM
             clock sampled during the run.  roofline_hbm: the same step against HBM with this implementation's own bytes
  cpu_baseline  the reference's own AccPathTracer (oracle/_ref, unmodified sources) on the host cores,
             on a bounded sample (same frame, few spp, two spp values to show the rate is spp-independent)
  other_workloads  BASELINE.json's other configs (cfg1, cfg2, cfg4-i/ii/iii; cfg5 at every N), one or two steps each
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (scene fixture, mode, width, height, spp, depth, aspect, reference component, edit applied to the fixture)
    "cfg1_raycast_cornell_500x500_1spp": ("ray_cast_cornel", 0, 500, 500, 1, 20, 1.0, "RayCast", None),
    "cfg2_simple_cornell_1024x1024_2048spp": ("path_tracing_cornel", 1, 1024, 1024, 2048, 20, 1.0, "SimplePathTracer", None),
    "cfg3_acc_bunny5k_1920x1080_1024spp": ("bunny5k_cornel", 2, 1920, 1080, 1024, 20, 16.0 / 9.0, "AccPathTracer", None),
    "cfg4_acc_gold_1920x1080_4096spp": ("pt_glass", 2, 1920, 1080, 4096, 20, 16.0 / 9.0, "AccPathTracer", None),
    # cfg4-ii / cfg4-iii are harness-defined: the sphere as Glass{ior 1.5}; Box/Pyramid as type-3 (microfacet) materials
    "cfg4ii_acc_glass_1920x1080_4096spp": ("pt_glass", 2, 1920, 1080, 4096, 20, 16.0 / 9.0, "AccPathTracer", "glass"),
    "cfg4iii_acc_microfacet_1920x1080_4096spp": ("pt_glass_conductors", 2, 1920, 1080, 4096, 20, 16.0 / 9.0, "AccPathTracer", "microfacet"),
    # cfg5: no reference semantics; synthetic 64x32 lat-long map
    "cfg5_acc_envmap_3840x2160_4096spp": ("env_map_spheres", 2, 3840, 2160, 4096, 20, 16.0 / 9.0, "AccPathTracer", "envmap"),
}
DEFAULT_WORKLOAD = "cfg3_acc_bunny5k_1920x1080_1024spp"
OTHER_WORKLOADS = ["cfg1_raycast_cornell_500x500_1spp", "cfg2_simple_cornell_1024x1024_2048spp", "cfg4_acc_gold_1920x1080_4096spp",
                   "cfg4ii_acc_glass_1920x1080_4096spp", "cfg4iii_acc_microfacet_1920x1080_4096spp"]
CFG5 = "cfg5_acc_envmap_3840x2160_4096spp"
# This implementation's own algorithmic HBM bytes: per ray 24 B read by stage 1 + 8 B hit written, 40 + 8 B read and
# 40 B written by the shading kernel; per path 32 B of radiance accumulation.  The scene itself is L1/L2 resident.
ALGO_BYTES_PER_RAY, ALGO_BYTES_PER_PATH = 120.0, 32.0
# SURVEY 8(d)'s figure for the REFERENCE's never-pruned binary tree (53.5 box tests/ray on the bunny): kept as a side key only
REFERENCE_TRAVERSAL_BYTES_PER_RAY = {"bunny5k_cornel": 2039.0}


def load_workload(name, spp_override=None):
    import numpy as np
    from nrenderer_b200.flatscene import FlatScene
    scene, mode, w, h, spp, depth, aspect, comp, edit = WORKLOADS[name]
    fs = FlatScene.load(os.path.join(REPO, "tests", "golden", scene + ".nrsc"))
    fs.width, fs.height, fs.samples_per_pixel, fs.depth, fs.cam_aspect = w, h, spp_override or spp, depth, aspect
    if edit == "glass":
        fs.sphere_material[:] = fs.add_material(2, ior=1.5, absorbed=[1, 1, 1])
    elif edit == "microfacet":        # conductors.scn's type-3 materials on the pyramid triangles and the box quads
        fs.triangle_material[:] = 4 + 6
        fs.plane_material[5:] = 4 + 0
    elif edit == "envmap":            # ambient = environment map (extension); deterministic synthetic texture
        g = np.random.default_rng(1)
        fs.ambient_type, fs.ambient_environment_map = 1, fs.add_texture(g.uniform(0.0, 2.0, (32, 64, 4)).astype(np.float32))
    return fs, mode, comp


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def csrc_hash():
    """Hash of the kernel sources: profiles/*_summary.json records it, so that instruction counts taken from an ncu capture
    of OTHER code are flagged (`profile_stale`)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(REPO, "nrenderer_b200", "csrc")
    for f in sorted(os.listdir(d)) + ["../../include/nrcu.h"]:
        with open(os.path.join(d, f), "rb") as fh:
            h.update(f.encode()); h.update(fh.read())
    return h.hexdigest()[:16]


def profile_summary():
    """Numbers derived from the committed ncu launch list of the bench workload (profiles/r2_summary.json, written by
    tools/make_profile_summary.py): warp instructions and DRAM bytes per path sample.  Not measured live - ncu serialises
    and slows the kernels - which is why the summary carries the hash of the code it was taken from."""
    for name in ("r2_summary.json", "r1_summary.json"):
        try:
            d = json.load(open(os.path.join(REPO, "profiles", name)))
            d["_file"] = "profiles/" + name
            return d
        except Exception:
            continue
    return {}


def sm_count():
    import torch
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


def rooflines(workload, paths, rays, ms, clocks, world, scheduler):
    """roofline (issue slots, the binding one) and roofline_hbm for the timed region of the default workload."""
    prof = profile_summary()
    issue = prof.get("issue") or {}
    per_sched = (prof.get("per_scheduler") or {}).get(scheduler) or {}
    wipp = per_sched.get("warp_inst_per_path_sample") or issue.get("warp_inst_per_path_sample")
    dram_pp = per_sched.get("dram_bytes_per_path_sample") or issue.get("dram_bytes_per_path_sample")
    stale = prof.get("csrc_sha") != csrc_hash()
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    sms = sm_count()
    hbm_peak, peak_src = peaks()
    secs = ms * 1e-3
    roof = {"bound": "issue", "achieved": None, "peak": sms * 4 * mhz * world, "unit": "warp-inst/us", "frac": None, "traffic": None,
            "kernel": "whole step (stage 1 + BVH traversal + shading; all concurrent streams)",
            "peak_source": f"{sms} SMs x 4 schedulers x {mhz:.0f} MHz sampled during the timed region x {world} GPU(s)",
            "warp_inst_per_path_sample": wipp, "threads_per_inst": per_sched.get("threads_per_inst") or issue.get("threads_per_inst_step"),
            "profile": prof.get("_file"), "profile_csrc_sha": prof.get("csrc_sha"), "csrc_sha": csrc_hash(), "profile_stale": stale,
            "note": "achieved = warp instructions per path sample (ncu launch list of the same command, committed under profiles/) x the path samples of "
                    "the timed region / its CUDA-event time; the scene is L1/L2 resident, so issue slots, not HBM, bound the step (roofline_hbm)"}
    hbm = {"bound": "hbm", "peak": hbm_peak * world, "unit": "GB/s", "peak_source": peak_src,
           "algorithmic_bytes_per_path_sample": None, "measured_dram_bytes_per_path_sample": dram_pp, "achieved": None, "frac": None, "achieved_measured_traffic": None}
    if workload == DEFAULT_WORKLOAD and secs > 0 and paths > 0:
        if wipp:
            roof["achieved"] = wipp * paths / (ms * 1e3)
            roof["frac"] = roof["achieved"] / roof["peak"]
        algo = ALGO_BYTES_PER_RAY * rays / paths + ALGO_BYTES_PER_PATH
        hbm["algorithmic_bytes_per_path_sample"] = algo
        hbm["achieved"] = algo * paths / secs * 1e-9
        hbm["frac"] = hbm["achieved"] / hbm["peak"]
        if dram_pp:
            roof["traffic"] = dram_pp * paths / max(1, 1)      # DRAM bytes of the timed region (ncu bytes per path sample x paths)
            hbm["achieved_measured_traffic"] = dram_pp * paths / secs * 1e-9
            hbm["traffic_over_algorithmic"] = dram_pp / algo
    return roof, hbm


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU component through its plugin API
# ---------------------------------------------------------------------------------------------
def run_reference_sample(fs, component, spp):
    from oracle import pyoracle as po
    f2 = fs.copy()
    f2.samples_per_pixel = spp
    _, info = po.run_reference(f2, component, timeout=1800)
    return info["seconds"], fs.width * fs.height * spp


def calibrated_reference_spp(fs, component, target_seconds):
    sec, paths = run_reference_sample(fs, component, 1)
    spp = max(1, min(64, int(round(target_seconds / max(sec, 1e-3)))))
    return spp, paths / sec * 1e-6


def reference_linearity(fs, component, spp):
    """the CPU rate is quoted from a few-spp sample of the frame, so show that it does not depend on spp."""
    a = max(1, spp // 2)
    sa, pa = run_reference_sample(fs, component, a)
    sb, pb = run_reference_sample(fs, component, 2 * a)
    return {"spp_a": a, "value_a": pa / sa * 1e-6, "spp_b": 2 * a, "value_b": pb / sb * 1e-6, "unit": "Mpath-samples/s",
            "note": "same frame at two sample counts: the rate is spp-independent, so the few-spp sample stands for the full-spp render"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    fs, mode, comp = load_workload(args.workload)
    linearity = None
    if not po.ref_available():   # the oracle port is the other CPU implementation of the path
        kind, cores = "port", os.cpu_count()
        osc = po.OracleScene(fs, mode)
        n_px = 4096

        def sample():
            import numpy as np
            px = np.random.default_rng(0).choice(fs.width * fs.height, n_px, replace=False).astype("uint32")
            t0 = time.perf_counter()
            osc.render_pt_accum(seed=0, s0=0, s1=4, pixels=px)
            return time.perf_counter() - t0, n_px * 4
        desc = f"oracle port, {n_px} random pixels x 4 spp of the frame"
    else:
        kind, cores = "reference", min(16, os.cpu_count() or 1)
        spp, _ = calibrated_reference_spp(fs, comp, args.ref_step_seconds)

        def sample():
            return run_reference_sample(fs, comp, spp)
        desc = f"{comp} (oracle/_ref, unmodified reference sources, 16 render threads) on the full {fs.width}x{fs.height} frame at {spp} spp, depth {fs.depth}"
        linearity = reference_linearity(fs, comp, spp)
    for _ in range(args.warmup):
        sample()
    secs, paths = 0.0, 0
    for _ in range(args.steps):
        s, p = sample()
        secs += s; paths += p
    value = paths / secs * 1e-6
    line = {"impl": "reference", "metric": "Mpath-samples/s", "value": value, "unit": "Mpath-samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "sample": desc},
            "cpu_baseline": {"value": value, "unit": "Mpath-samples/s", "cores": cores, "kind": kind, "sample": desc, "linearity": linearity},
            "e2e": {"value": value, "unit": "Mpath-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------
def scene_bytes(fs):
    return int(sum(getattr(fs, n).nbytes for n in fs._ARRAYS))


PLUGIN_COMPONENT = {0: "CudaRayCast", 1: "CudaSimplePathTracer", 2: "CudaAccPathTracer"}


def plugin_e2e(fs, mode, seed, steps, n_devices, device0, want_frame=False):
    """Wall time around RenderComponent::exec of the registered CUDA plugin, hosted by nr_headless (the reference's own
    ComponentFactory / Screen from libNRServer.so).  One warm-up exec (context creation, buffer allocation), then `steps` timed ones.
    Returns (seconds per exec, info, frame or None); raises when the harness or the plugin is not built."""
    import numpy as np
    import tempfile
    from nrenderer_b200 import build
    ref_dir = os.path.join(REPO, "oracle", "_ref")
    exe, so = os.path.join(ref_dir, "nr_headless"), build.plugin_path(mode)
    if not (os.path.exists(exe) and os.path.exists(so)):
        raise RuntimeError("nr_headless / plugin adapter not built (python __graft_entry__.py where /root/reference is mounted)")
    with tempfile.TemporaryDirectory() as td:
        sp, out = os.path.join(td, "scene.nrsc"), os.path.join(td, "frame.f32")
        fs.save(sp)
        env = dict(os.environ, NRCU_SEED=str(seed), NRCU_DEVICE=str(device0))
        env["LD_LIBRARY_PATH"] = ref_dir + os.pathsep + env.get("LD_LIBRARY_PATH", "")
        if n_devices > 1:
            env["NRCU_DEVICES"] = str(n_devices)
        cmd = [exe, "--flat", sp, "--plugin", so, "--component", PLUGIN_COMPONENT[mode], "--warmup", "1", "--repeat", str(steps)]
        if want_frame:
            cmd += ["--out", out]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1800)
        if r.returncode != 0:
            raise RuntimeError("nr_headless failed: " + r.stderr[-800:])
        info = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        if info.get("errors"):
            raise RuntimeError("plugin reported: " + info.get("last_error", ""))
        frame = np.fromfile(out, np.float32).reshape(info["height"], info["width"], 4) if want_frame else None
    return info["seconds_mean"], info, frame


def measure_workload(ctx, name, steps, seed, flush, spp_override=None):
    """One of the other BASELINE configs on this GPU: device-timed (CUDA events), scene resident, L2 flushed between steps."""
    import torch
    fs, mode, comp = load_workload(name, spp_override)
    w, h, spp = fs.width, fs.height, fs.samples_per_pixel
    ctx.upload(fs, mode)
    out = {"mode": ["RayCast", "SimplePathTracer", "AccPathTracer"][mode], "width": w, "height": h, "spp": spp, "depth": fs.depth, "steps": steps}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if mode == 0:    # RayCast: one deterministic pass; nrcu_render includes the D2H of the frame, the kernel time is in the stats
        import numpy as np
        host = np.empty((h, w, 4), np.float32)
        ctx.render(out=host)
        ms, rays = 0.0, 0
        for _ in range(steps):
            flush.fill_(1)
            _, st = ctx.render(out=host)
            ms += st["ms_total"]; rays += st["rays"]
        out.update({"ms_per_step": ms / steps, "value": w * h * steps / (ms * 1e-3) * 1e-6, "unit": "Mpath-samples/s (1 path = 1 pixel: primary + shadow ray)",
                    "mrays_per_s": rays / (ms * 1e-3) * 1e-6, "rays_per_path": rays / (w * h * steps), "timed": "k_raycast (CUDA events inside nrcu_render)"})
        return out
    accum = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    ctx.render_accumulate(accum.data_ptr(), s0=0, s1=min(spp, 64), seed=seed, want_stats=False)   # warm-up: allocations
    torch.cuda.synchronize()
    rays = paths = 0
    ev0.record()
    for _ in range(steps):
        flush.fill_(1)
        accum.zero_()
        st = ctx.render_accumulate(accum.data_ptr(), seed=seed, want_stats=True)
        rays += st["rays"]; paths += st["paths"]
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    out.update({"ms_per_step": ms / steps, "value": paths / (ms * 1e-3) * 1e-6, "unit": "Mpath-samples/s", "mrays_per_s": rays / (ms * 1e-3) * 1e-6,
                "rays_per_path": rays / max(paths, 1), "scheduler": {1: "waves", 2: "regen"}.get(st["scheduler"], "?")})
    del accum
    return out


def cuda_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from nrenderer_b200 import Context

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # Libraries print to fd 1 (NCCL's version banner at communicator creation): keep stdout for the ONE JSON line.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    else:
        torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    fs, mode, comp = load_workload(args.workload, args.spp)
    w, h, spp = fs.width, fs.height, fs.samples_per_pixel

    ctx = Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)   # our kernels, NCCL and the timing events share one stream
    ctx.upload(fs, mode)
    accum = torch.zeros(h, w, 4, dtype=torch.float32, device=dev)
    rgba = torch.empty(h, w, 4, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from nrenderer_b200 import multigpu

    def resolve(acc, out):
        ctx.resolve(acc.data_ptr(), out.data_ptr())

    def step(want_stats, coll_ev=None):
        flush.fill_(1)
        return multigpu.render_frame(
            lambda acc, a, b: ctx.render_accumulate(acc.data_ptr(), s0=a, s1=b, seed=args.seed, want_stats=want_stats, flags=4 if want_stats else 0),   # 4 = NRCU_FLAG_KERNEL_TIMES
            resolve, accum, rgba, spp, rank, world, collective_events=coll_ev)

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    coll_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)] if world > 1 else []
    # The timed steps are enqueued WITHOUT per-kernel statistics: reading back ~1 000 event spans per step would stall the
    # host - and with it the reduce - for about a millisecond after every frame.  Rays, launches and kernel times are taken
    # from one more, untimed step of the same frame (same seed: every step traces exactly the same paths).
    ev0.record()
    for k in range(args.steps):
        step(False, coll_events[k] if coll_events else None)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    ctx.synchronize()      # reports a traversal-stack overflow of the asynchronous steps, if any
    coll_ms = sum(a.elapsed_time(b) for a, b in coll_events)
    st = step(True) or {}
    barrier()
    agg = {k: st.get(k, 0) * args.steps for k in ("rays", "paths", "kernel_launches", "ms_trace", "ms_shade", "ms_stage2", "iterations")}
    # camera rays of the dead pixels (nrcu_stats.dead_pixels): closest-hit queries answered by the film rectangles, no ray traced
    agg["untraced"] = st.get("dead_pixels", 0) * (st.get("paths", 0) // max(1, w * h)) * args.steps
    sched = st.get("scheduler", 0)
    coll_iso = 0.0
    if world > 1:   # the exchange step on its own: all ranks start together, 10 repetitions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(12):
            if rep == 2:
                barrier(); e0.record()
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                resolve(accum, rgba)
        e1.record()
        torch.cuda.synchronize()
        coll_iso = e0.elapsed_time(e1) / 10
        step(False)          # leave the frame of the workload in rgba again
        barrier()
    t = torch.tensor([ms, agg["ms_trace"], agg["ms_shade"], agg["ms_stage2"], coll_ms, coll_iso], dtype=torch.float64, device=dev)
    sums = torch.tensor([agg["rays"], agg["paths"], agg["kernel_launches"], sched, agg["untraced"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mx = sums.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        sums[3] = mx[3]
    ms, ms_trace, ms_shade, ms_stage2, coll_ms, coll_iso = t.tolist()
    rays, paths, launches, sched, untraced = sums.tolist()
    launches += args.steps * (1 if rank == 0 else 0)   # resolve
    value = paths / (ms * 1e-3) * 1e-6
    scheduler = {1: "waves", 2: "regen"}.get(int(sched), "?")

    # ---- cfg5 (4K, 4096 spp: the configuration the north star's scaling claim is quoted on), 3 steps at every N --------------------
    cfg5 = None
    if args.workload == DEFAULT_WORKLOAD and not args.no_other_workloads:
        f5, m5, _ = load_workload(CFG5)
        ctx.upload(f5, m5)
        acc5 = torch.zeros(f5.height, f5.width, 4, dtype=torch.float32, device=dev)
        rgba5 = torch.empty_like(acc5)

        def step5(stats):
            flush.fill_(1)
            return multigpu.render_frame(lambda acc, a, b: ctx.render_accumulate(acc.data_ptr(), s0=a, s1=b, seed=args.seed, want_stats=stats),
                                         resolve, acc5, rgba5, f5.samples_per_pixel, rank, world)
        st = step5(True)            # warm-up (allocations) and the step that is counted: rays / paths of one frame
        r5, p5 = (3 * st["rays"], 3 * st["paths"]) if st else (0, 0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            step5(False)
        e1.record()
        barrier()
        t5 = torch.tensor([e0.elapsed_time(e1), r5, p5], dtype=torch.float64, device=dev)
        if world > 1:
            mx = t5.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(t5, op=dist.ReduceOp.SUM)
            t5[0] = mx[0]
        ms5, r5, p5 = t5.tolist()
        cfg5 = {"workload": CFG5, "width": f5.width, "height": f5.height, "spp": f5.samples_per_pixel, "depth": f5.depth, "steps": 3, "n_gpus": world,
                "ms_per_step": ms5 / 3, "value": p5 / (ms5 * 1e-3) * 1e-6, "unit": "Mpath-samples/s", "mrays_per_s": r5 / (ms5 * 1e-3) * 1e-6,
                "rays_per_path": r5 / max(p5, 1), "partition": f"sample slices x{world}, NCCL reduce of the 133 MB linear frame" if world > 1 else "single GPU",
                "note": "environment-map lighting is an extension (the reference renders this scene black, SURVEY A18); synthetic 64x32 lat-long map"}
        del acc5, rgba5
        ctx.upload(fs, mode)

    # ---- the other BASELINE configs on one GPU ----------------------------------------------------------------------------------
    others = None
    if world == 1 and args.workload == DEFAULT_WORKLOAD and not args.no_other_workloads:
        others = {}
        for name in OTHER_WORKLOADS:
            try:
                others[name] = measure_workload(ctx, name, 1 if "4096" in name else 2, args.seed, flush)
            except Exception as ex:
                others[name] = {"failed": str(ex)[:300]}
        ctx.upload(fs, mode)

    frame_nccl = rgba.cpu().numpy() if rank == 0 else None   # the frame of the last timed step (NCCL path at N > 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        ctx.close()
        return        # rank 0 goes on alone: the plugin opens its own contexts on all N devices

    # ---- end to end through the drop-in: RenderComponent::exec of the registered plugin, NRCU_DEVICES = N ----------------------------
    del accum, rgba
    ctx.close()
    torch.cuda.empty_cache()
    e2e_steps = max(5, min(args.steps, args.e2e_steps))
    h2d, d2h = scene_bytes(fs) * world, w * h * 16
    sampler2 = ClockSampler(local)
    sampler2.start()
    e2e = {"value": None, "unit": "Mpath-samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps}
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        sec, info, frame = plugin_e2e(fs, mode, args.seed, e2e_steps, world, local, want_frame=True)
        e2e.update({"value": w * h * spp / sec * 1e-6, "seconds_per_exec": sec, "seconds_best": info["seconds"],
                    "path": "nr_headless -> ComponentFactory::createComponent<RenderComponent> -> exec() -> NRCuda::Adapter::render (flatten, nrcu_upload_scene, "
                            + ("nrcu_render_multi over %d devices: peer-read reduce fused with the resolve" % world if world > 1 else "nrcu_render")
                            + ", the adapter's frame buffer) -> Screen::set; wall clock around exec(), through RenderComponent::exec",
                    "plugin_log": info.get("last_log"),
                    "max_abs_diff_vs_device_timed_frame": float(np.abs(np.clip(frame_nccl, 0, 1) - frame).max())})
    except Exception as ex:
        e2e["unavailable"] = str(ex)[:400]
    e2e["clocks"] = sampler2.stop()

    roof, roof_hbm = rooflines(args.workload, paths, rays - untraced, ms, clocks, world, scheduler)   # algorithmic bytes: traced rays only
    line = {
        "metric": "Mpath-samples/s", "value": value, "unit": "Mpath-samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "scene": "bunny_5k_faces.obj + path_tracing_cornel.scn (reference importers)" if "bunny" in args.workload else args.workload,
                   "width": w, "height": h, "spp": spp, "depth": fs.depth, "partition": f"sample slices x{world}, NCCL reduce" if world > 1 else "single GPU",
                   "l2": "256 MB flush buffer written between steps; the path state of a step (~5 GB) exceeds the 126 MB L2",
                   "glass_mode": "stochastic", "seed": args.seed, "scheduler": scheduler},
        "mrays_per_s": rays / (ms * 1e-3) * 1e-6, "rays_per_path": rays / max(paths, 1),
        "mrays_traced_per_s": (rays - untraced) / (ms * 1e-3) * 1e-6, "rays_traced_per_path": (rays - untraced) / max(paths, 1),
        "rays_note": "rays = closest-hit queries answered, as the reference counts them (one per trace() call); rays_traced leaves out the camera rays of "
                     "the pixels whose film footprint misses every primitive's bounds and every light: those are answered without generating a ray",
        "kernel_ms": {"closest_hit": ms_trace / args.steps, "of_which_bvh_traversal": ms_stage2 / args.steps, "shade": ms_shade / args.steps,
                      "note": "per-kernel CUDA-event spans summed over the concurrent streams: they overlap, so they add up to more than ms_per_step"},
        "iterations_per_step": agg["iterations"] / args.steps,
        "counters_from": "one untimed step of the same frame after the timed region (same seed, same paths); the timed steps run without per-kernel event read-backs",
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": roof,
        "roofline_hbm": roof_hbm,
        "clocks": clocks,
    }
    if world > 1:
        line["collective"] = {"what": "NCCL reduce(sum, fp32) of the linear frame to rank 0 + k_resolve", "bytes": w * h * 16,
                              "ms_per_step_in_the_timed_region": coll_ms / args.steps, "ms_isolated": coll_iso,
                              "note": "in the timed region the span starts when a rank has finished its slice, so it contains the wait for the slowest rank; "
                                      "isolated = the same reduce + resolve with all ranks starting together (max over ranks)"}
    ref_bpr = REFERENCE_TRAVERSAL_BYTES_PER_RAY.get(WORKLOADS[args.workload][0])
    if ref_bpr and ms_trace > 0:
        line["reference_traversal_equivalent"] = {"gbytes_per_s": rays * ref_bpr / (ms * 1e-3) * 1e-9, "bytes_per_ray": ref_bpr,
                                                  "note": "NOT a roofline: what the reference's never-pruned binary tree would have read per second (SURVEY 8d)"}
    if cfg5:
        line["cfg5"] = cfg5
    if others is not None:
        line["other_workloads"] = others
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import pyoracle as po
            if po.ref_available():
                spp_ref, _ = calibrated_reference_spp(fs, comp, args.cpu_baseline_seconds)
                sec, p = run_reference_sample(fs, comp, spp_ref)
                line["cpu_baseline"] = {"value": p / sec * 1e-6, "unit": "Mpath-samples/s", "cores": min(16, os.cpu_count() or 1), "kind": "reference",
                                        "sample": f"{comp} from oracle/_ref (unmodified reference sources, 16 render threads hard-coded, host has {os.cpu_count()} cpus) "
                                                  f"on the full {w}x{h} frame at {spp_ref} spp, depth {fs.depth}: {sec:.2f} s",
                                        "linearity": reference_linearity(fs, comp, max(2, spp_ref // 2))}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "Mpath-samples/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
        except Exception as ex:   # the GPU number stands on its own
            line["cpu_baseline"] = {"value": None, "unit": "Mpath-samples/s", "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=None, help="override samples per pixel (parity/debug only; the headline uses the workload's)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the plugin end-to-end leg (tuning / profiling runs)")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip cfg1/cfg2/cfg4/cfg5 (tuning runs)")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--ref-step-seconds", type=float, default=8.0)
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()
