#!/usr/bin/env python
"""Turn the ncu captures of a round into the committed summaries under profiles/.

  python tools/make_profile_summary.py <round> <launch_list.csv> <hot_kernels_raw.csv>

launch_list.csv      `ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,... --clock-control none --csv`
                     of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (first N launches)
hot_kernels_raw.csv  `ncu -i <rep> --page raw --csv` of an `ncu --set full` capture of k_big / k_trace2 / k_shade

Writes profiles/<round>_launches_bench.csv (per-launch lines, compacted), profiles/<round>_hot_kernels_raw.csv (selected
metrics), profiles/<round>_summary.json (read by bench.py for the `traffic` / `issue` fields) and profiles/<round>_summary.md.
"""
import collections
import csv
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
SMS, SCHEDULERS = 148, 4
PATHS_PER_WAVE = 128 * 1920 * 1080   # path samples of one FRAME of the capture command (bench.py default workload cfg3 at --spp 128)
KEEP = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors.sum", "SM_B.TriageCompute.l1tex__t_sectors.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("nrcu::", "").strip()


def read_launch_list(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    iK, iM, iV, iID, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID"), hdr.index("Metric Unit")
    L = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[iV].replace(",", ""))
        if r[iM] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iU], 1e-3)   # -> us
        if r[iM].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[iU], 1.0)
        L.setdefault(r[iID], {"kernel": short(r[iK])})[r[iM]] = v
    return list(L.values())


def main():
    rnd, launch_csv, raw_csv = sys.argv[1:4]
    out_dir = os.path.join(REPO, "profiles")
    launches = read_launch_list(launch_csv)
    with open(os.path.join(out_dir, f"{rnd}_launches_bench.csv"), "w") as f:
        f.write("launch,kernel,time_us,warp_inst,threads_per_inst,issue_active_pct,dram_read_bytes,dram_write_bytes\n")
        for i, d in enumerate(launches):
            f.write(f"{i},{d['kernel']},{d.get('gpu__time_duration.sum', 0):.2f},{d.get('smsp__inst_executed.sum', 0):.0f},"
                    f"{d.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0):.2f},{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.1f},"
                    f"{d.get('dram__bytes_read.sum', 0):.0f},{d.get('dram__bytes_write.sum', 0):.0f}\n")
    ours = [d for d in launches if d["kernel"].startswith("k_")]
    render = [d for d in ours if not d["kernel"].startswith(("k_bvh", "k_build", "k_mesh", "k_big_rects", "k_env", "k_live", "k_scene_rects"))]
    T = sum(d["gpu__time_duration.sum"] for d in render)
    agg = collections.OrderedDict()
    for d in render:
        a = agg.setdefault(d["kernel"], dict(launches=0, time_us=0.0, warp_inst=0.0, lane_inst=0.0, dram=0.0, issue_w=0.0))
        t = d["gpu__time_duration.sum"]
        a["launches"] += 1; a["time_us"] += t; a["warp_inst"] += d.get("smsp__inst_executed.sum", 0)
        a["lane_inst"] += d.get("smsp__inst_executed.sum", 0) * d.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0)
        a["dram"] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        a["issue_w"] += d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0) * t
    kernels = {}
    for k, a in agg.items():
        kernels[k] = dict(launches=a["launches"], time_us=round(a["time_us"], 1), share=round(a["time_us"] / T, 4),
                          warp_inst=a["warp_inst"], threads_per_inst=round(a["lane_inst"] / max(a["warp_inst"], 1), 2),
                          issue_active_pct=round(a["issue_w"] / max(a["time_us"], 1e-9), 1),
                          dram_bytes_per_launch=round(a["dram"] / a["launches"]), dram_gbs=round(a["dram"] / (a["time_us"] * 1e-6) * 1e-9, 1))
    # warp instructions of the complete waves in the capture (everything up to the last k_accumulate), per path sample:
    # a wave of the bench workload is 32 samples of 1920 x 1080 pixels
    # complete frames of the capture: everything up to the last k_resolve; a frame of the capture command is 128 spp of 1920 x 1080
    # (the number of waves per frame depends on the live-pixel count, so the unit of work is the frame, not the wave)
    last_acc = max((i for i, d in enumerate(render) if d["kernel"].startswith("k_resolve")), default=-1)
    n_waves = sum(1 for d in render if d["kernel"].startswith("k_resolve"))
    inst_complete = sum(d.get("smsp__inst_executed.sum", 0) for d in render[:last_acc + 1])
    warp_inst_per_path = inst_complete / (n_waves * PATHS_PER_WAVE) if n_waves else None
    dram_complete = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in render[:last_acc + 1])
    dram_per_path = dram_complete / (n_waves * PATHS_PER_WAVE) if n_waves else None
    lane_complete = sum(d.get("smsp__inst_executed.sum", 0) * d.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0) for d in render[:last_acc + 1])
    import bench   # the hash of the kernel sources this capture was taken from (bench.py prints profile_stale when it differs)
    csrc_sha = bench.csrc_hash()
    ch = [k for k in kernels if k.startswith(("k_raygen", "k_big", "k_trace"))]
    ch_time = sum(kernels[k]["time_us"] for k in ch)
    ch_inst = sum(kernels[k]["warp_inst"] for k in ch)
    ch_dram = sum(kernels[k]["dram_bytes_per_launch"] * kernels[k]["launches"] for k in ch)
    # raw metrics of the full captures
    hot = []
    rows = list(csv.reader(open(raw_csv)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(out_dir, f"{rnd}_hot_kernels_raw.csv"), "w") as f:
        cols = ["Kernel Name"] + [k for k in KEEP if k in hdr]
        w = csv.writer(f); w.writerow(cols); w.writerow([""] + [units[hdr.index(k)] for k in cols[1:]])
        for r in rows[2:]:
            w.writerow([short(r[hdr.index("Kernel Name")])] + [r[hdr.index(k)] for k in cols[1:]])
            hot.append({"kernel": short(r[hdr.index("Kernel Name")]), **{k: r[hdr.index(k)] for k in cols[1:]}})
    summary = {
        "round": rnd,
        "csrc_sha": csrc_sha,
        "source": {"launch_list": os.path.basename(launch_csv), "captured_launches": len(launches), "full_capture": os.path.basename(raw_csv)},
        "render_kernel_time_us": round(T, 1),
        "kernels": kernels,
        "closest_hit": {"kernels": ch, "share_of_render_kernels": round(ch_time / T, 4)},
        # closest-hit DRAM traffic per captured unit of work, and the issue-slot utilisation of the closest-hit kernels:
        # warp instructions issued / (elapsed cycles x 148 SMs x 4 schedulers), clock from the capture (1.965 GHz locked by ncu --clock-control none = application clocks)
        "closest_hit_dram_bytes_per_step_equiv": None,
        "issue": {"warp_inst_per_path_sample": warp_inst_per_path, "dram_bytes_per_path_sample": dram_per_path,
                  "threads_per_inst_step": round(lane_complete / max(inst_complete, 1), 2),
                  "complete_waves_in_capture": n_waves, "paths_per_wave": PATHS_PER_WAVE,
                  "closest_hit_warp_inst_per_us": round(ch_inst / ch_time, 1),
                  "peak_warp_inst_per_us_at_1965MHz": SMS * SCHEDULERS * 1965.0,
                  "frac_of_issue_peak": round(ch_inst / ch_time / (SMS * SCHEDULERS * 1965.0), 4),
                  "threads_per_inst": {k: kernels[k]["threads_per_inst"] for k in ch},
                  "note": "from the committed ncu launch list (cold-cache, serialised launches); shares, not absolute times, carry over to the bench"},
        "closest_hit_dram_bytes_per_launch": {k: kernels[k]["dram_bytes_per_launch"] for k in ch},
        "hot_kernels_full_capture": hot,
    }
    summary["closest_hit"]["dram_bytes_per_launch"] = round(ch_dram / max(sum(kernels[k]["launches"] for k in ch), 1))
    summary["closest_hit"]["launches_per_wave"] = round(sum(kernels[k]["launches"] for k in ch) / max(sum(kernels[k]["launches"] for k in ch if k.startswith("k_raygen")), 1), 2)
    summary["closest_hit_dram_bytes_per_step_equiv"] = round(ch_dram / max(sum(kernels[k]["launches"] for k in ch if k.startswith("k_raygen")), 1))   # per wave
    json.dump(summary, open(os.path.join(out_dir, f"{rnd}_summary.json"), "w"), indent=1)
    with open(os.path.join(out_dir, f"{rnd}_summary.md"), "w") as f:
        f.write(f"# {rnd} profile summary (ncu, B200, `NRCU_WAVE_MSLOTS=128 bench.py --steps 1 --warmup 1 --spp 128 --no-cpu-baseline --no-e2e --no-other-workloads`, first {len(launches)} launches; kernel sources {csrc_sha})\n\n")
        f.write(f"Whole step: **{warp_inst_per_path:.1f} warp instructions and {dram_per_path:.0f} DRAM bytes per path sample**, {lane_complete / max(inst_complete, 1):.1f} active lanes per instruction "
                f"(complete frames of the capture: {n_waves} x {PATHS_PER_WAVE} path samples).\n\n")
        f.write("| kernel | launches | time (us) | share | warp-inst | lanes/inst | issue-active % | DRAM B/launch | DRAM GB/s |\n|---|---|---|---|---|---|---|---|---|\n")
        for k, a in sorted(kernels.items(), key=lambda kv: -kv[1]["time_us"]):
            f.write(f"| `{k}` | {a['launches']} | {a['time_us']:.0f} | {a['share']*100:.1f} % | {a['warp_inst']/1e6:.0f} M | {a['threads_per_inst']} | {a['issue_active_pct']} | {a['dram_bytes_per_launch']/1e6:.1f} M | {a['dram_gbs']} |\n")
        f.write(f"\nClosest-hit kernels ({', '.join(ch)}): {ch_time/T*100:.1f} % of the render-kernel time, "
                f"{ch_inst/ch_time/(SMS*SCHEDULERS*1965.0)*100:.1f} % of the issue-slot peak (148 SMs x 4 x 1.965 GHz), "
                f"DRAM traffic per wave {summary['closest_hit_dram_bytes_per_step_equiv']/1e9:.2f} GB.\n")
        f.write("\nncu serialises kernel launches: the two waves that run side by side in the real run (NRCU_WAVES=2, +25 %) appear back to back here, "
                "so the shares above are per-kernel work shares, not wall-clock shares of the overlapped run.\n")
        f.write("\nFull captures (`ncu --set full`, one launch each, early bounces of the first waves):\n\n| kernel | time us | regs | lanes/inst | issue % | warps active % | L1 hit % | L2 hit % | L1 GB/s (% of peak) | L2 GB/s (% of peak) | DRAM R MB | DRAM W MB | DRAM GB/s (of 6549.8) | long-scoreboard stall |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        U = {k: units[hdr.index(k)] for k in KEEP if k in hdr}
        def mb(h, k):
            try: v = float(str(h.get(k, 0) or 0).replace(',', ''))
            except ValueError: v = 0.0
            return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(U.get(k, "byte"), 1e-6)
        def us(h, k):
            try: v = float(str(h.get(k, 0) or 0).replace(',', ''))
            except ValueError: v = 0.0
            return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(U.get(k, "us"), 1.0)
        for h in hot:
            g = lambda k: h.get(k, "")
            def fl(x):
                try: return float(str(x).replace(',', ''))
                except ValueError: return 0.0
            f.write(f"| `{h['kernel']}` | {us(h, 'gpu__time_duration.sum'):.1f} | {g('launch__registers_per_thread')} | {g('smsp__thread_inst_executed_per_inst_executed.ratio')} | "
                    f"{fl(g('smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f} | {fl(g('sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | "
                    f"{fl(g('l1tex__t_sector_hit_rate.pct')):.1f} | {fl(g('lts__t_sector_hit_rate.pct')):.1f} | {fl(g('SM_B.TriageCompute.l1tex__t_sectors.sum')) * 32e-9 / max(us(h, 'gpu__time_duration.sum') * 1e-6, 1e-12):.0f} ({fl(g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed')):.0f} %) | {fl(g('lts__t_sectors.sum')) * 32e-9 / max(us(h, 'gpu__time_duration.sum') * 1e-6, 1e-12):.0f} ({fl(g('lts__throughput.avg.pct_of_peak_sustained_elapsed')):.0f} %) | {mb(h, 'dram__bytes_read.sum'):.1f} | {mb(h, 'dram__bytes_write.sum'):.1f} | {(mb(h, 'dram__bytes_read.sum') + mb(h, 'dram__bytes_write.sum')) / max(us(h, 'gpu__time_duration.sum'), 1e-9) * 1e3:.0f} ({(mb(h, 'dram__bytes_read.sum') + mb(h, 'dram__bytes_write.sum')) / max(us(h, 'gpu__time_duration.sum'), 1e-9) * 1e3 / 6549.8 * 100:.0f} %) | "
                    f"{fl(g('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio')):.2f} |\n")
    print(open(os.path.join(out_dir, f"{rnd}_summary.md")).read())


if __name__ == "__main__":
    main()
