"""In-tree builds: libnrcuda.so (nvcc, sm_100a) and the NRenderer plugin adapters (g++).

`build_cuda()` cross-compiles on a machine without a GPU.  `build_plugins()` needs the reference's
headers and libNRServer.so, i.e. /root/reference plus oracle/_ref — on the GPU box the prebuilt
.so files that travelled with the snapshot are used as they are.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libnrcuda.so")
PLUGIN_DIR = os.path.join(PKG, "plugin")
PLUGINS = {0: "CudaRayCast", 1: "CudaSimplePathTracer", 2: "CudaAccPathTracer", 3: "CudaMetropolisLightTransport"}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def cuda_sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(REPO, "include", "nrcu.h")]


def build_cuda(force=False, verbose=False, variant=None, defines=()) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> nrenderer_b200/libnrcuda.so

    `variant` + `defines` build an experiment copy, libnrcuda.<variant>.so, with extra -D flags (same-call A/B
    measurements: api.py loads the library named by NRCU_LIBRARY instead of the default one)."""
    lib = LIB if not variant else os.path.join(PKG, f"libnrcuda.{variant}.so")
    if not force and not variant and not _newer(lib, cuda_sources()):
        return lib
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + list(defines) + [f"-I{REPO}/include", f"-I{CSRC}", os.path.join(CSRC, "nrcu_api.cu"), "-o", lib]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return lib


def plugin_path(mode: int) -> str:
    return os.path.join(PLUGIN_DIR, f"libNR{PLUGINS[mode]}.so")


def build_plugins(force=False):
    """g++ the RenderComponent adapters against the patched reference headers (needs /root/reference)."""
    sys.path.insert(0, REPO)
    from tools.reference_overlay import ensure_overlay, include_flags, reference_available
    ref_dir = os.path.join(REPO, "oracle", "_ref")
    if not reference_available() or not os.path.exists(os.path.join(ref_dir, "libNRServer.so")):
        return [p for p in (plugin_path(m) for m in PLUGINS) if os.path.exists(p)]
    ensure_overlay()
    src = os.path.join(PLUGIN_DIR, "NRCudaAdapter.cpp")
    deps = [src, os.path.join(PLUGIN_DIR, "scene_bridge.hpp"), os.path.join(PKG, "host", "flat_scene.hpp"),
            os.path.join(REPO, "include", "nrcu.h")]
    out = []
    for mode in PLUGINS:
        so = plugin_path(mode)
        if force or _newer(so, deps):
            cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-w", "-shared", f"-DNRCU_PLUGIN_MODE={mode}"] + include_flags() + \
                  [f"-I{PLUGIN_DIR}", src, f"-L{PKG}", "-lnrcuda", f"-L{ref_dir}", "-lNRServer",
                   "-Wl,-Bsymbolic", "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath,$ORIGIN/../../oracle/_ref", "-o", so]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("plugin build failed:\n" + r.stdout + r.stderr)
        out.append(so)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:   # python -m nrenderer_b200.build --variant TAG -DFOO=1 ...
        tag = sys.argv[sys.argv.index("--variant") + 1]
        print(build_cuda(force=True, verbose="-v" in sys.argv, variant=tag, defines=[a for a in sys.argv if a.startswith("-D")]))
        sys.exit(0)
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_plugins(force="--force" in sys.argv))
