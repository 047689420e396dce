"""ctypes binding of tests/host_emu/libnrcu_emu.so (CPU emulation of the device code; test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
_LIB = None


def build(force=False) -> str:
    so = os.path.join(HERE, "libnrcu_emu.so")
    csrc = os.path.join(REPO, "nrenderer_b200", "csrc")
    srcs = [os.path.join(HERE, "nrcu_emu.cpp"), os.path.join(REPO, "include", "nrcu.h")] + \
           [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".hpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DNRCU_HOST_EMU=1",
                        f"-I{REPO}/include", f"-I{csrc}", "-x", "c++", os.path.join(HERE, "nrcu_emu.cpp"), "-o", so],
                       check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.c_void_p, C.c_int]
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_last_error.restype = C.c_char_p
        L.emu_last_error.argtypes = [C.c_void_p]
        L.emu_primitive_count.restype = C.c_uint32
        L.emu_primitive_count.argtypes = [C.c_void_p]
        L.emu_primitives.argtypes = [C.c_void_p] * 5
        L.emu_camera.argtypes = [C.c_void_p] * 3
        L.emu_bvh_stats.argtypes = [C.c_void_p] * 2
        L.emu_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.emu_render_raycast.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_render_pt.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
        L.emu_camera_ray.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        L.emu_philox4x32.argtypes = [C.c_void_p] * 3
        _LIB = L
    return _LIB


class EmuScene:
    def __init__(self, flat, mode):
        view, keep = flat.c_view()
        self._h = lib().emu_create(C.addressof(view), mode)
        err = lib().emu_last_error(self._h).decode()
        if err:
            raise ValueError(err)
        self.width, self.height = flat.width, flat.height

    def __del__(self):
        if getattr(self, "_h", None):
            lib().emu_destroy(self._h)
            self._h = None

    @property
    def n_primitives(self):
        return int(lib().emu_primitive_count(self._h))

    def primitives(self):
        n = self.n_primitives
        kind, data, mat, box = np.zeros(n, np.uint32), np.zeros((n, 16), np.float32), np.zeros(n, np.int32), np.zeros((n, 6), np.float32)
        lib().emu_primitives(self._h, kind.ctypes.data, data.ctypes.data, mat.ctypes.data, box.ctypes.data)
        return kind, data, mat, box

    def camera(self):
        c, lr = np.zeros(18, np.float32), C.c_float(0)
        lib().emu_camera(self._h, c.ctypes.data, C.addressof(lr))
        return c.reshape(6, 3), lr.value

    def bvh_stats(self):
        o = np.zeros(8, np.int32)
        lib().emu_bvh_stats(self._h, o.ctypes.data)
        return dict(binary_nodes=int(o[0]), wide_nodes=int(o[1]), levels=int(o[2]), max_leaf=int(o[3]), leaf_slots=int(o[4]), root_ref=int(o[5]), n_big=int(o[6]))

    def trace_batch(self, rays, linear=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = len(rays)
        pid, t = np.zeros(n, np.int32), np.zeros(n, np.float32)
        lib().emu_trace_batch(self._h, rays.ctypes.data, n, pid.ctypes.data, t.ctypes.data, int(linear))
        return pid, t

    def render_raycast(self):
        out = np.zeros((self.height, self.width, 4), np.float32)
        lib().emu_render_raycast(self._h, out.ctypes.data)
        return out

    def render_pt_accum(self, seed=0, s0=0, s1=0, glass_mode=0, pixels=None, flags=0):
        rays = C.c_uint64(0)
        if pixels is None:
            acc = np.zeros((self.height, self.width, 4), np.float32)
            lib().emu_render_pt(self._h, seed, s0, s1, glass_mode, None, 0, acc.ctypes.data, C.addressof(rays), flags)
        else:
            pixels = np.ascontiguousarray(pixels, np.uint32)
            acc = np.zeros((len(pixels), 4), np.float32)
            lib().emu_render_pt(self._h, seed, s0, s1, glass_mode, pixels.ctypes.data, len(pixels), acc.ctypes.data, C.addressof(rays), flags)
        return acc, rays.value

    def camera_ray(self, seed, pixel, sample):
        o = np.zeros(6, np.float32)
        lib().emu_camera_ray(self._h, seed, pixel, sample, o.ctypes.data)
        return o


def philox4x32(counter, key):
    c, k, o = np.asarray(counter, np.uint32), np.asarray(key, np.uint32), np.zeros(4, np.uint32)
    lib().emu_philox4x32(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o
