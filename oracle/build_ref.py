#!/usr/bin/env python
"""Build the REAL reference (civilizwa/nrenderer, /root/reference) for Linux into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing on the product path (nrenderer_b200/, libnrcuda.so) may
import, link or execute anything under oracle/.

What it builds (outputs only under oracle/_ref/, which is git-ignored but travels to the GPU box):
  libNRServer.so          code/server/**               (scene types, plugin registry, Screen, Logger)
  libRayCast.so           code/components/ray_cast
  libSimplePathTracer.so  code/components/simple_path_tracing
  libAccPathTracing.so    code/components/acc_path_tracing   (registers "AccPathTracer")
  nr_headless             nrenderer_b200/harness/nr_headless.cpp + the reference's own
                          ScnImporter/ObjImporter/SceneBuilder/ImageLoader sources
  include_patched.tar     NOT produced - the patched headers stay in the scratch overlay only.

The reference is MSVC-only code, so it is compiled from a scratch COPY (default
/tmp/nrref_overlay, never inside this repo) to which the mechanical patches of SURVEY.md
§8(c) are applied.  No reference source is copied into the repository.  The reference's own
CMake build is not used (Windows/OpenGL/GLFW binaries); this script is the whole recipe.
The CUDA plugin adapters (nrenderer_b200/plugin/) are compiled against the same overlay headers.
"""
import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
from tools.reference_overlay import OVERLAY, REF, include_flags, make_overlay  # noqa: E402

OUT = os.path.join(HERE, "_ref")
CXXFLAGS = ["-std=c++20", "-O2", "-fPIC", "-w", "-ffp-contract=off"]


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit(f"build_ref: command failed ({r.returncode})")


def sources(d):
    out = []
    for root, _, files in os.walk(d):
        out += [os.path.join(root, f) for f in sorted(files) if f.endswith(".cpp")]
    return sorted(out)


def build(jobs=8):
    os.makedirs(OUT, exist_ok=True)
    make_overlay()
    inc = include_flags()
    cxx = ["g++"] + CXXFLAGS
    run(cxx + ["-shared"] + inc + sources(os.path.join(OVERLAY, "server")) + ["-o", os.path.join(OUT, "libNRServer.so")])
    link = [f"-L{OUT}", "-lNRServer", "-Wl,-rpath,$ORIGIN"]

    def plugin(item):
        d, name = item
        src = sources(os.path.join(OVERLAY, "components", d, "src"))
        run(cxx + ["-shared"] + inc + [f"-I{OVERLAY}/components/{d}/include"] + src + link + ["-lpthread", "-o", os.path.join(OUT, f"lib{name}.so")])

    def harness(_):
        dep = os.path.join(REF, "code", "dependences")
        run(["gcc", "-O1", "-w", "-c", f"-I{dep}/glad/include", os.path.join(dep, "glad", "src", "glad.c"), "-o", os.path.join(OVERLAY, "glad.o")])
        src = [os.path.join(REPO, "nrenderer_b200", "harness", "nr_headless.cpp"),
               os.path.join(OVERLAY, "app/src/importer/ScnImporter.cpp"), os.path.join(OVERLAY, "app/src/importer/ObjImporter.cpp"),
               os.path.join(OVERLAY, "app/src/asset/SceneBuilder.cpp"), os.path.join(OVERLAY, "app/src/utilities/ImageLoader.cpp"),
               os.path.join(OVERLAY, "glad.o")]
        run(cxx + inc + src + link + ["-ldl", "-lpthread", "-o", os.path.join(OUT, "nr_headless")])

    items = [("ray_cast", "RayCast"), ("simple_path_tracing", "SimplePathTracer"), ("acc_path_tracing", "AccPathTracing")]
    with ThreadPoolExecutor(jobs) as ex:
        futs = [ex.submit(plugin, it) for it in items] + [ex.submit(harness, None)]
        for f in futs:
            f.result()
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--overlay-only", action="store_true", help="only (re)create the patched scratch overlay")
    a = ap.parse_args()
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: the reference can only be built where it is mounted")
    if a.overlay_only:
        print(make_overlay())
    else:
        print(build())
