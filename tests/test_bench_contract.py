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
    for tag in ("cfg1", "cfg2", "cfg3", "cfg4_", "cfg4ii", "cfg4iii", "cfg5"):
        assert any(w.startswith(tag) for w in bench.WORKLOADS), tag
    assert len(base["configs"]) == 5
    assert set(bench.OTHER_WORKLOADS) | {bench.DEFAULT_WORKLOAD, bench.CFG5} == set(bench.WORKLOADS)   # every config is measured somewhere in the default line
    for name in bench.WORKLOADS:
        fs, mode, comp = bench.load_workload(name, spp_override=1)
        assert fs.width * fs.height > 0 and mode in (0, 1, 2) and comp in po.REF_PLUGINS
    glass, _, _ = bench.load_workload("cfg4ii_acc_glass_1920x1080_4096spp")
    assert int(glass.material_type_present[int(glass.sphere_material[0]), 0]) == 2            # the sphere is the Glass of env_map_spheres.scn
    mf, _, _ = bench.load_workload("cfg4iii_acc_microfacet_1920x1080_4096spp")
    assert int(mf.material_type_present[int(mf.triangle_material[0]), 0]) == 3                # conductors.scn's type 3 -> Microfacet


def test_profile_summary_carries_the_hash_of_the_code_it_was_taken_from():
    sys.path.insert(0, REPO)
    import bench
    prof = bench.profile_summary()
    assert len(bench.csrc_hash()) == 16
    if prof.get("_file", "").endswith("r2_summary.json"):
        assert "csrc_sha" in prof and prof["issue"]["warp_inst_per_path_sample"] > 0


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
    lin = d["cpu_baseline"]["linearity"]                    # the CPU rate at two spp values
    assert lin["spp_b"] == 2 * lin["spp_a"] and lin["value_a"] > 0 and lin["value_b"] > 0


def test_cuda_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=REPO)
    assert r.returncode != 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
