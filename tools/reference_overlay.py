#!/usr/bin/env python
"""Patched scratch copy ("overlay") of the reference headers/sources needed to compile against
NRenderer's plugin API on Linux/GCC.

The reference (civilizwa/nrenderer, mounted read-only at /root/reference) is MSVC-only code.  This
module copies the handful of directories a plugin or the headless harness needs into a scratch
directory OUTSIDE this repository (default /tmp/nrref_overlay) and applies the mechanical patches
listed in SURVEY.md §8(c).  Nothing from the reference is copied into the repo.  Used by
oracle/build_ref.py (reference CPU components + harness) and nrenderer_b200/build.py (the CUDA
plugin adapters, which must be compiled against the same headers as libNRServer.so).
"""
import os
import re
import shutil

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NR_REFERENCE", "/root/reference")
OVERLAY = os.environ.get("NR_OVERLAY", "/tmp/nrref_overlay")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF, "code", "include"))


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


def include_flags():
    dep = os.path.join(REF, "code", "dependences")
    return [f"-I{OVERLAY}/include", f"-I{dep}/glm", f"-I{dep}/glad/include", f"-I{dep}/stb_image/include",
            f"-I{OVERLAY}/app/include", f"-I{REPO}/include"]


def ensure_overlay():
    if not os.path.exists(os.path.join(OVERLAY, "include", "scene", "Scene.hpp")):
        make_overlay()
    return OVERLAY
