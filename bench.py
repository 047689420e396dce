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
  e2e        the same metric through the reference-facing call sequence with HOST buffers: upload of
             the Scene arrays (H2D), render, D2H of the RGBA frame that Screen::set would receive
  roofline   the closest-hit kernels (k_raygen incl. fused stage 1, k_big, k_trace2): algorithmic bytes/ray x rays /
             CUDA-event time of their launches inside the timed region
  cpu_baseline  the reference's own AccPathTracer (oracle/_ref, unmodified sources) on the host cores,
             on a bounded sample (same frame, few spp)
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
    # name: (scene fixture, mode, width, height, spp, depth, aspect, reference component, algorithmic bytes per ray)
    "cfg2_simple_cornell_1024x1024_2048spp": ("path_tracing_cornel", 1, 1024, 1024, 2048, 20, 1.0, "SimplePathTracer", 866.0),
    "cfg3_acc_bunny5k_1920x1080_1024spp": ("bunny5k_cornel", 2, 1920, 1080, 1024, 20, 16.0 / 9.0, "AccPathTracer", 2039.0),
    "cfg4_acc_gold_1920x1080_4096spp": ("pt_glass", 2, 1920, 1080, 4096, 20, 16.0 / 9.0, "AccPathTracer", 866.0),
    # cfg5: no reference semantics; synthetic 64x32 lat-long map; meant for --gpus 8 (weak point: 34 G paths)
    "cfg5_acc_envmap_3840x2160_4096spp": ("env_map_spheres", 2, 3840, 2160, 4096, 20, 16.0 / 9.0, "AccPathTracer", 866.0),
}
DEFAULT_WORKLOAD = "cfg3_acc_bunny5k_1920x1080_1024spp"


def load_workload(name, spp_override=None):
    from nrenderer_b200.flatscene import FlatScene
    scene, mode, w, h, spp, depth, aspect, comp, bpr = WORKLOADS[name]
    fs = FlatScene.load(os.path.join(REPO, "tests", "golden", scene + ".nrsc"))
    fs.width, fs.height, fs.samples_per_pixel, fs.depth, fs.cam_aspect = w, h, spp_override or spp, depth, aspect
    if name.startswith("cfg5"):   # ambient = environment map (extension); deterministic synthetic texture
        import numpy as np
        g = np.random.default_rng(1)
        fs.ambient_type, fs.ambient_environment_map = 1, fs.add_texture(g.uniform(0.0, 2.0, (32, 64, 4)).astype(np.float32))
    return fs, mode, comp, bpr


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_summary():
    """Numbers derived from the committed ncu captures (profiles/r1_summary.json); not measured live."""
    p = os.path.join(REPO, "profiles", "r1_summary.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def issue_roofline(workload, paths, ms, clocks, world=1):
    """The binding roofline: warp instructions per path sample (committed ncu launch list of this workload) x the paths of
    the timed region / its live CUDA-event time, against 148 SMs x 4 schedulers x the SM clock sampled during the run."""
    d = dict(profile_summary().get("issue") or {})
    wipp = d.get("warp_inst_per_path_sample")
    if wipp and workload == DEFAULT_WORKLOAD and ms > 0:
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        peak = 148 * 4 * mhz * world              # warp instructions per microsecond, all GPUs
        d["step"] = {"bound": "issue", "achieved": wipp * paths / (ms * 1e3), "peak": peak, "unit": "warp-inst/us", "frac": wipp * paths / (ms * 1e3) / peak,
                     "sm_mhz": mhz, "note": "whole step, all kernels, both concurrent waves"}
    return d


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


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    fs, mode, comp, _ = load_workload(args.workload)
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
            "cpu_baseline": {"value": value, "unit": "Mpath-samples/s", "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": value, "unit": "Mpath-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------
def scene_bytes(fs):
    return int(sum(getattr(fs, n).nbytes for n in fs._ARRAYS))


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
    fs, mode, comp, bytes_per_ray = load_workload(args.workload, args.spp)
    w, h, spp = fs.width, fs.height, fs.samples_per_pixel

    ctx = Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)   # our kernels, NCCL and the timing events share one stream
    ctx.upload(fs, mode)
    accum = torch.zeros(h, w, 4, dtype=torch.float32, device=dev)
    rgba = torch.empty(h, w, 4, dtype=torch.float32, device=dev)
    host_rgba = torch.empty(h, w, 4, dtype=torch.float32, pin_memory=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from nrenderer_b200 import multigpu

    def resolve(acc, out):
        ctx.resolve(acc.data_ptr(), out.data_ptr())

    def step(want_stats):
        flush.fill_(1)
        return multigpu.render_frame(
            lambda acc, a, b: ctx.render_accumulate(acc.data_ptr(), s0=a, s1=b, seed=args.seed, want_stats=want_stats),
            resolve, accum, rgba, spp, rank, world)

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = {"rays": 0, "paths": 0, "kernel_launches": 0, "ms_trace": 0.0, "ms_shade": 0.0, "ms_stage2": 0.0}
    ev0.record()
    for _ in range(args.steps):
        st = step(True)
        for k in agg:
            agg[k] += st[k] if st else 0
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, agg["ms_trace"], agg["ms_shade"], agg["ms_stage2"]], dtype=torch.float64, device=dev)
    sums = torch.tensor([agg["rays"], agg["paths"], agg["kernel_launches"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms, ms_trace, ms_shade, ms_stage2 = t.tolist()
    ms_closest = ms_trace               # k_raygen (fused stage 1) + k_big_balanced + k_trace2; ms_stage2 = the k_trace2 part
    rays, paths, launches = sums.tolist()
    launches += args.steps * (1 if rank == 0 else 0)   # resolve
    value = paths / (ms * 1e-3) * 1e-6

    # ---- end to end: Scene arrays from host memory in, RGBA frame in host memory out ------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h2d, d2h = scene_bytes(fs), w * h * 16

    host_np = host_rgba.numpy()

    def e2e_step():
        ctx.upload(fs, mode)                                    # H2D of the scene + device-side flattening + BVH build
        if world == 1:                                          # exactly the plugin adapter's call sequence (NRCudaAdapter.cpp):
            ctx.render(seed=args.seed, out=host_np)             # nrcu_upload_scene + nrcu_render into HOST memory (pinned)
            return
        multigpu.render_frame(
            lambda acc, a, b: ctx.render_accumulate(acc.data_ptr(), s0=a, s1=b, seed=args.seed, want_stats=False),
            resolve, accum, rgba, spp, rank, world)
        if rank == 0:
            host_rgba.copy_(rgba, non_blocking=True)            # D2H of what Screen::set receives
        torch.cuda.synchronize()

    e2e_step()
    barrier()
    sampler2 = ClockSampler(local) if rank == 0 else None
    if sampler2:
        sampler2.start()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_clocks = sampler2.stop() if sampler2 else None
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = (w * h * spp * e2e_steps) / te.item() * 1e-6

    waves = -(-spp // max(1, (64 << 20) // (w * h))) if world == 1 else None   # 2 concurrent waves of 64 Mi path slots
    closest_hit_launches = args.steps * waves * (1 + (fs.depth - 1) + fs.depth) if waves else 0
    if rank == 0:
        peak, peak_src = peaks()
        achieved = rays * bytes_per_ray / (ms_closest * 1e-3) * 1e-9 if ms_closest > 0 else None
        line = {
            "metric": "Mpath-samples/s", "value": value, "unit": "Mpath-samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "scene": "bunny_5k_faces.obj + path_tracing_cornel.scn (reference importers)" if "bunny" in args.workload else args.workload,
                       "width": w, "height": h, "spp": spp, "depth": fs.depth, "partition": f"sample slices x{world}, NCCL reduce" if world > 1 else "single GPU",
                       "l2": "256 MB flush buffer written between steps; per-wave ray/path state (~7 GB) exceeds the 126 MB L2",
                       "glass_mode": "stochastic", "seed": args.seed},
            "mrays_per_s": rays / (ms * 1e-3) * 1e-6, "rays_per_path": rays / max(paths, 1),
            "kernel_ms": {"closest_hit": ms_closest / args.steps, "of_which_bvh_traversal": ms_stage2 / args.steps, "shade": ms_shade / args.steps},
            "e2e": {"value": e2e_value, "unit": "Mpath-samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps, "clocks": e2e_clocks},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                         "traffic": profile_summary().get("closest_hit", {}).get("dram_bytes_per_launch"),
                         "algorithmic_bytes_per_launch": (rays * bytes_per_ray / max(closest_hit_launches, 1)) if closest_hit_launches else None,
                         "kernel": "closest hit = k_raygen (camera rays + fused stage 1) + k_big_balanced + k_trace2", "algorithmic_bytes_per_ray": bytes_per_ray,
                         "peak_source": peak_src, "share_of_step": ms_closest / ms if ms else None, "concurrent_waves": 2,
                         "note": "two waves run side by side on two streams, so kernel times (summed per launch) overlap and their share of the step can exceed 1; "
                                 "algorithmic bytes are those of the REFERENCE's traversal (SURVEY 8d); the scene is L1/L2 resident, so the "
                                 "fraction can exceed 1 and the binding limit is issue slots x warp efficiency: see 'issue' (from the committed "
                                 "ncu launch list, profiles/) and DESIGN.md section 5",
                         "issue": issue_roofline(args.workload, paths, ms, clocks, world)},
            # the same kernels against the HBM roofline with THIS implementation's algorithmic bytes per ray
            "roofline_kernels": [
                {"kernel": "k_shade", "bound": "hbm", "algorithmic_bytes_per_ray": 88.0, "achieved": rays * 88.0 / (ms_shade * 1e-3) * 1e-9 if ms_shade > 0 else None,
                 "peak": peak, "unit": "GB/s", "frac": rays * 88.0 / (ms_shade * 1e-3) * 1e-9 / peak if ms_shade > 0 else None,
                 "traffic": profile_summary().get("kernels", {}).get("k_shade<1, 0>", {}).get("dram_bytes_per_launch")},
                {"kernel": "k_raygen + k_big_balanced (stage 1)", "bound": "issue", "algorithmic_bytes_per_ray": 36.0,
                 "achieved": rays * 36.0 / ((ms_trace - ms_stage2) * 1e-3) * 1e-9 if ms_trace > ms_stage2 else None, "peak": peak, "unit": "GB/s",
                 "frac": rays * 36.0 / ((ms_trace - ms_stage2) * 1e-3) * 1e-9 / peak if ms_trace > ms_stage2 else None,
                 "traffic": profile_summary().get("kernels", {}).get("k_big_balanced<1>", {}).get("dram_bytes_per_launch"),
                 "issue_active_pct": profile_summary().get("kernels", {}).get("k_big_balanced<1>", {}).get("issue_active_pct")},
            ],
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import pyoracle as po
                if po.ref_available():
                    spp_ref, _ = calibrated_reference_spp(fs, comp, args.cpu_baseline_seconds)
                    sec, p = run_reference_sample(fs, comp, spp_ref)
                    line["cpu_baseline"] = {"value": p / sec * 1e-6, "unit": "Mpath-samples/s", "cores": min(16, os.cpu_count() or 1), "kind": "reference",
                                            "sample": f"{comp} from oracle/_ref (unmodified reference sources, 16 render threads hard-coded, host has {os.cpu_count()} cpus) "
                                                      f"on the full {w}x{h} frame at {spp_ref} spp, depth {fs.depth}: {sec:.2f} s"}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "Mpath-samples/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
            except Exception as ex:   # the GPU number stands on its own
                line["cpu_baseline"] = {"value": None, "unit": "Mpath-samples/s", "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=None, help="override samples per pixel (parity/debug only; the headline uses the workload's)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=15.0)
    ap.add_argument("--ref-step-seconds", type=float, default=8.0)
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        cuda_arm(args)


if __name__ == "__main__":
    main()
