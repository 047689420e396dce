#!/usr/bin/env python
"""A/B the traversal-kernel knobs on a GPU box: runs bench.py at reduced spp once per setting
(the knobs are read from the environment when libnrcuda.so first launches a trace kernel)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SETTINGS = [
    {"NRCU_TRACE_VARIANT": "2", "NRCU_TRACE_REFILL": "8"},
    {"NRCU_TRACE_VARIANT": "3", "NRCU_TRACE_REFILL": "8"},
    {"NRCU_TRACE_VARIANT": "3", "NRCU_TRACE_REFILL": "1"},
    {"NRCU_TRACE_VARIANT": "3", "NRCU_TRACE_REFILL": "4"},
    {"NRCU_TRACE_VARIANT": "3", "NRCU_TRACE_REFILL": "16"},
    {"NRCU_TRACE_VARIANT": "3", "NRCU_TRACE_REFILL": "8", "NRCU_TRACE_WNODE": "1", "NRCU_TRACE_WPRIM": "2"},
    {"NRCU_TRACE_VARIANT": "3", "NRCU_TRACE_REFILL": "8", "NRCU_TRACE_WNODE": "2", "NRCU_TRACE_WPRIM": "1"},
    {"NRCU_TRACE_VARIANT": "3", "NRCU_TRACE_REFILL": "8", "NRCU_TRACE_WNODE": "2", "NRCU_TRACE_WPRIM": "3"},
    {"NRCU_TRACE_VARIANT": "3", "NRCU_TRACE_REFILL": "8", "NRCU_TRACE_WNODE": "3", "NRCU_TRACE_WPRIM": "2"},
    {"NRCU_TRACE_VARIANT": "3", "NRCU_TRACE_REFILL": "8", "NRCU_TRACE_BLOCKS": "9"},
    {"NRCU_TRACE_VARIANT": "3", "NRCU_TRACE_REFILL": "8", "NRCU_TRACE_BLOCKS": "6"},
]


def main():
    spp = sys.argv[1] if len(sys.argv) > 1 else "64"
    extra = sys.argv[2:]
    settings = SETTINGS
    if os.environ.get("NRCU_TUNE_SETTINGS"):   # JSON list of env dicts
        settings = json.loads(os.environ["NRCU_TUNE_SETTINGS"])
    for st in settings:
        env = dict(os.environ, **st)
        r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "2", "--warmup", "1", "--spp", spp,
                            "--no-cpu-baseline", "--no-e2e", "--no-other-workloads"] + extra, capture_output=True, text=True, env=env)
        try:
            d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
            print(f"{st}: {d['value']:.1f} Mpath/s  {d['mrays_per_s']:.1f} Mrays/s  closest-hit {d['kernel_ms']['closest_hit']:.2f} ms  shade {d['kernel_ms']['shade']:.2f} ms  step {d['ms_per_step']:.2f} ms", flush=True)
        except Exception:
            print(st, "FAILED", r.stdout[-500:], r.stderr[-1500:], flush=True)


if __name__ == "__main__":
    main()
