"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the
keys the driver reads (and nothing else on stdout), the workloads name BASELINE.json's configs, the CUDA arm refuses to
run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

from conftest import REPO
from oracle import pyoracle as po


def test_workloads_cover_the_baseline_configs():
    sys.path.insert(0, REPO)
    import bench
    base = json.load(open(os.path.join(REPO, "BASELINE.json")))
    assert bench.DEFAULT_WORKLOAD.startswith("cfg3") and "1920x1080_1024spp" in bench.DEFAULT_WORKLOAD   # the config the target is quoted on
    for tag in ("cfg2", "cfg3", "cfg4", "cfg5"):
        assert any(w.startswith(tag) for w in bench.WORKLOADS), tag
    assert len(base["configs"]) == 5
    for name in bench.WORKLOADS:
        fs, mode, comp, bytes_per_ray = bench.load_workload(name, spp_override=1)
        assert fs.width * fs.height > 0 and mode in (1, 2) and comp in po.REF_PLUGINS and bytes_per_ray > 0


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built")
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-step-seconds", "0.5"], capture_output=True, text=True, timeout=300, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpath-samples/s" and d["unit"] == "Mpath-samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("cfg3")


def test_cuda_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=REPO)
    assert r.returncode != 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
