#!/usr/bin/env python
"""Small renders through every kernel family, meant to run under compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
RayCast, SimplePathTracer and AccPathTracer modes, stochastic and branching glass, env map, trace_batch, multi-wave."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from conftest import env_texture, glassify, load_scene, microfacet, random_rays  # noqa: E402
from nrenderer_b200 import Context  # noqa: E402


def main():
    ctx = Context(0)
    fs = load_scene("ray_cast_cornel", width=64, height=48)
    ctx.upload(fs, 0); img, st = ctx.render(); print("raycast", st["rays"])
    for name, mode, edit, glass in [("path_tracing_cornel", 1, None, 0), ("bunny5k_cornel", 2, None, 0), ("pt_glass", 2, glassify, 0),
                                    ("pt_glass", 2, glassify, 1), ("pt_glass_conductors", 2, microfacet, 0), ("env_map_spheres", 2, env_texture, 0)]:
        fs = load_scene(name, width=48, height=32, samples_per_pixel=6, depth=8)
        if edit:
            edit(fs)
        ctx.upload(fs, mode)
        img, st = ctx.render(seed=1, glass_mode=glass, samples_per_wave=4)
        assert np.isfinite(img).all()
        pid, t = ctx.trace_batch(random_rays(5000, seed=2))
        print(name, mode, glass, st["rays"], int((pid >= 0).sum()))
    out, st = ctx.render_progressive(lambda f, d, t: False, samples_per_update=2)
    ctx.close()
    print("sanitize smoke ok")


if __name__ == "__main__":
    main()
