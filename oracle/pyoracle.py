"""ctypes binding of the CPU restatement oracle (oracle/libnroracle.so) and runner for oracle/_ref.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by nrenderer_b200/.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "libnroracle.so")
    src = [os.path.join(HERE, "nr_oracle.c"), os.path.join(HERE, "nr_oracle.h"), os.path.join(HERE, "..", "include", "nrcu.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s", "libnroracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.nro_prepare.restype = C.c_void_p
        L.nro_prepare.argtypes = [C.c_void_p, C.c_int]
        L.nro_free.argtypes = [C.c_void_p]
        L.nro_primitive_count.restype = C.c_uint32
        L.nro_primitive_count.argtypes = [C.c_void_p]
        L.nro_get_primitives.argtypes = [C.c_void_p] * 4
        L.nro_get_bounds.argtypes = [C.c_void_p] * 2
        L.nro_get_camera.argtypes = [C.c_void_p] * 3
        L.nro_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nro_bounds_intersectp.restype = C.c_int
        L.nro_bounds_intersectp.argtypes = [C.c_void_p] * 3
        L.nro_render_raycast.argtypes = [C.c_void_p, C.c_void_p]
        L.nro_render_pt.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.nro_render_pt_pixels.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.nro_render_pt_pixels_flags.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.nro_resolve.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.nro_camera_ray.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        L.nro_philox4x32.argtypes = [C.c_void_p] * 3
        L.nro_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


class OracleScene:
    """Prepared scene for one mode (0 RayCast, 1 SimplePathTracer, 2 AccPathTracer)."""

    def __init__(self, flat, mode: int):
        self.flat, self.mode = flat, mode
        view, keep = flat.c_view()
        self._h = lib().nro_prepare(C.addressof(view), mode)
        del keep
        self.width, self.height = flat.width, flat.height

    def __del__(self):
        if getattr(self, "_h", None):
            lib().nro_free(self._h)
            self._h = None

    @property
    def n_primitives(self) -> int:
        return int(lib().nro_primitive_count(self._h))

    def primitives(self):
        n = self.n_primitives
        kind, data, mat = np.zeros(n, np.uint32), np.zeros((n, 16), np.float32), np.zeros(n, np.int32)
        lib().nro_get_primitives(self._h, kind.ctypes.data, data.ctypes.data, mat.ctypes.data)
        return kind, data, mat

    def bounds(self):
        b = np.zeros((self.n_primitives, 6), np.float32)
        lib().nro_get_bounds(self._h, b.ctypes.data)
        return b

    def camera(self):
        c, lr = np.zeros(18, np.float32), C.c_float(0)
        lib().nro_get_camera(self._h, c.ctypes.data, C.addressof(lr))
        return c.reshape(6, 3), lr.value

    def trace_batch(self, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = len(rays)
        pid, t, tie = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
        lib().nro_trace_batch(self._h, rays.ctypes.data, n, pid.ctypes.data, t.ctypes.data, tie.ctypes.data)
        return pid, t, tie.astype(bool)

    def render_raycast(self) -> np.ndarray:
        out = np.zeros((self.height, self.width, 4), np.float32)
        lib().nro_render_raycast(self._h, out.ctypes.data)
        return out

    def render_pt_accum(self, seed=0, s0=0, s1=0, glass_mode=0, pixels=None, flags=0):
        """Linear sums (rgb) + sample count (a). Returns (accum, rays).  flags: nrcu_render_flags (1 = NEE extension)."""
        rays = C.c_uint64(0)
        if pixels is None:
            acc = np.zeros((self.height, self.width, 4), np.float32)
            lib().nro_render_pt_pixels_flags(self._h, seed, s0, s1, glass_mode, flags, None, 0, acc.ctypes.data, C.addressof(rays))
        else:
            pixels = np.ascontiguousarray(pixels, np.uint32)
            acc = np.zeros((len(pixels), 4), np.float32)
            lib().nro_render_pt_pixels_flags(self._h, seed, s0, s1, glass_mode, flags, pixels.ctypes.data, len(pixels), acc.ctypes.data, C.addressof(rays))
        return acc, rays.value

    def camera_ray(self, seed, pixel, sample):
        o = np.zeros(6, np.float32)
        lib().nro_camera_ray(self._h, seed, pixel, sample, o.ctypes.data)
        return o


def resolve(accum: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(accum, np.float32)
    out = np.zeros_like(a)
    lib().nro_resolve(a.ctypes.data, a.size // 4, out.ctypes.data)
    return out


def philox4x32(counter, key):
    c, k, o = np.asarray(counter, np.uint32), np.asarray(key, np.uint32), np.zeros(4, np.uint32)
    lib().nro_philox4x32(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def bounds_intersectp(box6, origin, direction) -> bool:
    b, o, d = (np.ascontiguousarray(x, np.float32) for x in (box6, origin, direction))
    return bool(lib().nro_bounds_intersectp(b.ctypes.data, o.ctypes.data, d.ctypes.data))


# ---------------------------------------------------------------------------------------------
# The real reference, compiled into oracle/_ref by oracle/build_ref.py
# ---------------------------------------------------------------------------------------------
REF_PLUGINS = {"RayCast": "libRayCast.so", "SimplePathTracer": "libSimplePathTracer.so", "AccPathTracer": "libAccPathTracing.so"}


def ref_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "nr_headless")) and os.path.exists(os.path.join(REF_DIR, "libNRServer.so"))


def run_reference(flat, component: str, *, repeat: int = 1, warmup: int = 0, extra_plugins=(), plugin_dirs=(), manager=False, env=None, timeout=3600):
    """Run a registered render component through the reference's plugin API on a flat scene
    (nr_headless: SceneBuilder-equivalent Scene -> ComponentFactory::createComponent -> RenderComponent::exec -> Screen).
    `manager`: go through ComponentManager::exec on a detached thread like the GUI.  `env`: extra environment (NRCU_*).
    Returns (rgba[h,w,4], info dict with wall seconds)."""
    if not ref_available():
        raise RuntimeError("oracle/_ref is not built (run oracle/build_ref.py where /root/reference is mounted)")
    with tempfile.TemporaryDirectory() as td:
        scene_path, out_path = os.path.join(td, "scene.nrsc"), os.path.join(td, "frame.f32")
        flat.save(scene_path)
        cmd = [os.path.join(REF_DIR, "nr_headless"), "--flat", scene_path]
        plugins = list(extra_plugins)
        if component in REF_PLUGINS:
            plugins.append(os.path.join(REF_DIR, REF_PLUGINS[component]))
        for p in plugins:
            cmd += ["--plugin", p]
        for d in plugin_dirs:
            cmd += ["--plugin-dir", d]
        if manager:
            cmd += ["--manager"]
        cmd += ["--component", component, "--out", out_path, "--repeat", str(repeat), "--warmup", str(warmup)]
        full_env = dict(os.environ)
        full_env.update(env or {})
        full_env["LD_LIBRARY_PATH"] = REF_DIR + os.pathsep + full_env.get("LD_LIBRARY_PATH", "")
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=full_env)
        if r.returncode != 0:
            raise RuntimeError(f"nr_headless failed: {r.stderr[-2000:]}")
        info = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        img = np.fromfile(out_path, np.float32).reshape(info["height"], info["width"], 4)
    return img, info
