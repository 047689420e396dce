#!/usr/bin/env python
"""Generate the committed fixtures under tests/golden/ from the REAL reference.

Run in the build container (needs /root/reference and oracle/_ref built by oracle/build_ref.py):
    python tests/golden/make_fixtures.py [--scenes] [--raycast] [--pt]

  *.nrsc                       flat scenes written by nr_headless --dump-flat, i.e. parsed by the
                               reference's own ScnImporter / ObjImporter + SceneBuilder
  ray_cast_cornel_500_ref.npz  the reference RayCast frame (500x500, deterministic) + its md5
  pt_ref_<case>.npz            reference path-traced frames: R independent runs of the reference
                               component, converted to LINEAR space (the reference publishes
                               sqrt(mean), AccPathTracer.cpp:32-33), per-pixel mean and standard
                               error over the runs, and a validity mask (pixels that saturate in
                               Screen::set's clamp to [0,1] cannot be inverted)
The GPU box has no /root/reference: tests there read only these files.
"""
import argparse
import hashlib
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from nrenderer_b200.flatscene import FlatScene  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

RES = "/root/reference/resource"
HEADLESS = os.path.join(REPO, "oracle", "_ref", "nr_headless")

SCENES = {
    "ray_cast_cornel": ["--scn", f"{RES}/ray_cast_cornel.scn"],
    "path_tracing_cornel": ["--scn", f"{RES}/path_tracing_cornel.scn"],
    "bunny5k_cornel": ["--obj", f"{RES}/obj/bunny_5k_faces.obj", "--scn", f"{RES}/path_tracing_cornel.scn", "--mesh-material", "0"],
    "bunny200_cornel": ["--obj", f"{RES}/obj/bunny_200_faces.obj", "--scn", f"{RES}/path_tracing_cornel.scn", "--mesh-material", "0"],
    "pt_glass": ["--scn", f"{RES}/pt_glass.scn"],
    "pt_glass_conductors": ["--scn", f"{RES}/pt_glass.scn", "--scn", f"{RES}/conductors.scn"],
    "env_map_spheres": ["--scn", f"{RES}/env_map_spheres.scn"],
}


def glassify(fs):
    """cfg4-ii: the pt_glass sphere with the Glass material of env_map_spheres.scn:7-10."""
    fs.sphere_material[:] = fs.add_material(2, ior=1.5, absorbed=[1, 1, 1])


def microfacet(fs):
    """cfg4-iii: conductors.scn materials (type 3 -> Microfacet) on the box and pyramid of pt_glass.scn."""
    fs.triangle_material[:] = 4 + 6   # Copper
    fs.plane_material[5:] = 4 + 0     # Mirror on the box planes (planes 0-4 are the walls)


# name: (scene, component, mode, w, h, spp per run, depth, runs, edit)
PT_CASES = {
    "simple_cornell_d4": ("path_tracing_cornel", "SimplePathTracer", 1, 48, 48, 2048, 4, 8, None),
    "acc_cornell_d20": ("path_tracing_cornel", "AccPathTracer", 2, 48, 48, 1024, 20, 8, None),
    "acc_bunny5k_d20": ("bunny5k_cornel", "AccPathTracer", 2, 48, 48, 512, 20, 8, None),
    "acc_gold_d20": ("pt_glass", "AccPathTracer", 2, 48, 48, 1024, 20, 8, None),
    "acc_glass_d6": ("pt_glass", "AccPathTracer", 2, 48, 48, 512, 6, 8, glassify),
    "acc_microfacet_d8": ("pt_glass_conductors", "AccPathTracer", 2, 48, 48, 1024, 8, 8, microfacet),
}


def make_scenes():
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(REPO, "oracle", "_ref"))
    for name, args in SCENES.items():
        out = os.path.join(HERE, name + ".nrsc")
        subprocess.run([HEADLESS] + args + ["--dump-flat", out], check=True, env=env)
        print("wrote", out)


def make_raycast():
    fs = FlatScene.load(os.path.join(HERE, "ray_cast_cornel.nrsc"))
    img, info = po.run_reference(fs, "RayCast")
    md5 = hashlib.md5(img.tobytes()).hexdigest()
    np.savez_compressed(os.path.join(HERE, "ray_cast_cornel_500_ref.npz"), rgb=img[..., :3].copy(), md5=md5, seconds=info["seconds"])
    print("raycast md5", md5)


def make_pt():
    for name, (scene, comp, mode, w, h, spp, depth, runs, edit) in PT_CASES.items():
        fs = FlatScene.load(os.path.join(HERE, scene + ".nrsc"))
        fs.width, fs.height, fs.samples_per_pixel, fs.depth = w, h, spp, depth
        if edit:
            edit(fs)
        # Validity mask: Screen::set clamps the published sqrt(mean) to [0,1] (Screen.cpp:62-64), so bright pixels
        # cannot be inverted.  The mask must not depend on the reference's own noise (selecting pixels whose
        # reference runs happened to stay below 1 biases its mean low): it comes from an independent
        # high-spp estimate by the oracle (linear mean of every channel < 0.5).
        osc = po.OracleScene(fs, mode)
        oacc, _ = osc.render_pt_accum(seed=99, s0=0, s1=2048)
        valid = ((oacc[..., :3] / oacc[..., 3:4]) < 0.5).all(-1)
        lin, secs = [], []
        for r in range(runs):
            t0 = time.time()
            img, info = po.run_reference(fs, comp)
            secs.append(info["seconds"])
            rgb = img[..., :3].astype(np.float64)
            valid &= np.isfinite(rgb).all(-1)
            lin.append(rgb ** 2)
            # the reference seeds its samplers with time(0): make sure the next run sees another second
            time.sleep(max(0.0, 1.1 - (time.time() - t0)))
        lin = np.stack(lin)
        mean = lin.mean(0)
        sem = lin.std(0, ddof=1) / np.sqrt(runs)
        np.savez_compressed(os.path.join(HERE, f"pt_ref_{name}.npz"), mean=mean.astype(np.float32), sem=sem.astype(np.float32),
                            valid=valid, runs=runs, spp_per_run=spp, depth=depth, width=w, height=h, mode=mode,
                            scene=scene, component=comp, seconds=np.array(secs))
        print(f"{name}: mean {mean[valid].mean():.5f} sem/mean {sem[valid].mean() / mean[valid].mean():.4f} valid {valid.mean():.3f} {np.mean(secs):.2f}s/run")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", action="store_true"); ap.add_argument("--raycast", action="store_true"); ap.add_argument("--pt", action="store_true")
    a = ap.parse_args()
    every = not (a.scenes or a.raycast or a.pt)
    if not po.ref_available():
        raise SystemExit("oracle/_ref is not built")
    if a.scenes or every:
        make_scenes()
    if a.raycast or every:
        make_raycast()
    if a.pt or every:
        make_pt()
