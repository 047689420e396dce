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
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("NR_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
OVERLAY = os.environ.get("NR_OVERLAY", "/tmp/nrref_overlay")
CXXFLAGS = ["-std=c++20", "-O2", "-fPIC", "-w", "-ffp-contract=off"]


def sub_file(path, pattern, repl, count=0, must=True, flags=0):
    with open(path, encoding="utf-8", errors="surrogateescape") as f:
        s = f.read()
    s2, n = re.subn(pattern, repl, s, count=count, flags=flags)
    if must and n == 0:
        raise RuntimeError(f"patch did not apply: {path}: {pattern}")
    with open(path, "w", encoding="utf-8", errors="surrogateescape") as f:
        f.write(s2)


def make_overlay():
    code = os.path.join(REF, "code")
    if os.path.exists(OVERLAY):
        shutil.rmtree(OVERLAY)
    os.makedirs(OVERLAY)
    shutil.copytree(os.path.join(code, "include"), os.path.join(OVERLAY, "include"))
    shutil.copytree(os.path.join(code, "server"), os.path.join(OVERLAY, "server"))
    for c in ("ray_cast", "simple_path_tracing", "acc_path_tracing"):
        shutil.copytree(os.path.join(code, "components", c), os.path.join(OVERLAY, "components", c))
    shutil.copytree(os.path.join(code, "app", "include"), os.path.join(OVERLAY, "app", "include"))
    for rel in ("app/src/importer/ScnImporter.cpp", "app/src/importer/ObjImporter.cpp",
                "app/src/asset/SceneBuilder.cpp", "app/src/utilities/ImageLoader.cpp"):
        dst = os.path.join(OVERLAY, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copy(os.path.join(code, rel), dst)
    for root, _, files in os.walk(OVERLAY):
        os.chmod(root, 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)

    ov = lambda *p: os.path.join(OVERLAY, *p)
    # 1. case-sensitive include of "Server/..." (server/server/{Screen,Logger}.cpp:1)
    os.symlink("server", ov("include", "Server"))
    # 2. "HemiSphere.hpp" vs Hemisphere.hpp (components/*/include/samplers/SamplerInstance.hpp:5)
    for c in ("simple_path_tracing", "acc_path_tracing"):
        os.symlink("Hemisphere.hpp", ov("components", c, "include", "samplers", "HemiSphere.hpp"))
    # 3. Material.hpp:26-27: nested template Base with a default member initialiser is used by the
    #    variant before the enclosing class is complete -> hoist it out of the nested class.
    mat = ov("include", "scene", "Material.hpp")
    sub_file(mat, r"template<typename T>\s*struct Base \{ T value = \{\}; \};", "", count=1)
    sub_file(mat, r"(\n\s*struct Property\s*\{)",
             r"\n    template<typename T> struct PropertyBase { T value = {}; };\1\n        template<typename T> using Base = PropertyBase<T>;",
             count=1)
    sub_file(mat, r"class Wrapper\s*\{\s*private:", "class Wrapper\n        {\n        private:\n            template<typename T> using Base = PropertyBase<T>;", count=1)
    # 4. Model.hpp:32-39: Vec3 members inside an anonymous struct inside a union are rejected by GCC.
    mdl = ov("include", "scene", "Model.hpp")
    sub_file(mdl, r"union \{\s*struct \{\s*Vec3 v1;\s*Vec3 v2;\s*Vec3 v3;\s*\};\s*Vec3 v\[3\];\s*\};",
             "Vec3 v1; Vec3 v2; Vec3 v3;\n        Vec3& vertex(int i) { return i == 0 ? v1 : (i == 1 ? v2 : v3); }", count=1)
    for c in ("ray_cast", "simple_path_tracing", "acc_path_tracing"):
        sub_file(ov("components", c, "src", "VertexTransformer.cpp"), r"\.v\[i\]", ".vertex(i)")
    sub_file(ov("app", "src", "importer", "ScnImporter.cpp"), r"->v\[0\]", "->v1", must=False)
    sub_file(ov("app", "src", "importer", "ScnImporter.cpp"), r"->v\[1\]", "->v2", must=False)
    sub_file(ov("app", "src", "importer", "ScnImporter.cpp"), r"->v\[2\]", "->v3", must=False)
    # 5. Timer.hpp: high_resolution_clock::now() assigned to a steady_clock::time_point
    for c in ("simple_path_tracing", "acc_path_tracing"):
        sub_file(ov("components", c, "include", "Timer.hpp"), r"high_resolution_clock", "steady_clock")
    # 6. <thread> is not included transitively under libstdc++
    sub_file(ov("components", "simple_path_tracing", "src", "SimplePathTracer.cpp"), r'(#include "server/Server.hpp")', r"#include <thread>\n\1", count=1)
    sub_file(ov("components", "acc_path_tracing", "src", "AccPathTracer.cpp"), r'(#include "server/Server.hpp")', r"#include <thread>\n\1", count=1)
    # 7. ObjImporter.cpp:302: std::exception(const char*) is an MSVC extension
    sub_file(ov("app", "src", "importer", "ObjImporter.cpp"), r"exception e\((\".*?\")\);", r"std::runtime_error e(\1);", count=1)
    sub_file(ov("app", "src", "importer", "ObjImporter.cpp"), r'(#include "importer/ObjImporter.hpp")', r"#include <stdexcept>\n\1", count=1)
    return OVERLAY


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit(f"build_ref: command failed ({r.returncode})")


def include_flags():
    dep = os.path.join(REF, "code", "dependences")
    return [f"-I{OVERLAY}/include", f"-I{dep}/glm", f"-I{dep}/glad/include", f"-I{dep}/stb_image/include",
            f"-I{OVERLAY}/app/include", f"-I{REPO}/include"]


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
