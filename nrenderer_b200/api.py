"""ctypes binding of libnrcuda.so (the C ABI in include/nrcu.h).

Host plumbing for tests and bench.py: it owns no rendering arithmetic and has NO fallback — if the
CUDA library is missing, or no CUDA device is present, every entry point raises `NrcuError`.
The product's host side proper is the C++ RenderComponent adapter in nrenderer_b200/plugin/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .flatscene import FlatScene, MODE_ACC, MODE_RAYCAST, MODE_SIMPLE  # noqa: F401

FLAG_NEE = 1   # nrcu_render_flags.NRCU_FLAG_NEE
FLAG_ENV_IS = 2   # nrcu_render_flags.NRCU_FLAG_ENV_IS
FLAG_KERNEL_TIMES = 4   # nrcu_render_flags.NRCU_FLAG_KERNEL_TIMES: per-kernel CUDA-event spans in the stats
SCHED_AUTO, SCHED_WAVES, SCHED_REGEN = 0, 1, 2   # nrcu_scheduler
ERR_OVERFLOW = 6

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libnrcuda.so")

ABI_SYMBOLS = [
    "nrcu_abi_version", "nrcu_device_count", "nrcu_create", "nrcu_destroy", "nrcu_last_error", "nrcu_upload_scene",
    "nrcu_primitive_count", "nrcu_download_primitives", "nrcu_render", "nrcu_render_accumulate", "nrcu_resolve",
    "nrcu_render_multi", "nrcu_render_progressive", "nrcu_render_mlt", "nrcu_trace_batch", "nrcu_set_stream", "nrcu_synchronize", "nrcu_philox4x32",
    "nrcu_host_alloc", "nrcu_host_free",
]


class NrcuError(RuntimeError):
    pass


class NrcuRenderParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("sample_begin", C.c_uint32), ("sample_end", C.c_uint32),
                ("glass_mode", C.c_uint32), ("samples_per_wave", C.c_uint32), ("flags", C.c_uint32), ("scheduler", C.c_uint32)]


class NrcuStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("ms_total", C.c_float), ("ms_trace", C.c_float), ("ms_shade", C.c_float), ("ms_setup", C.c_float),
                ("bvh_nodes", C.c_uint32), ("n_primitives", C.c_uint32), ("max_queue", C.c_uint32), ("ms_stage2", C.c_float),
                ("scheduler", C.c_uint32), ("iterations", C.c_uint32), ("wave_retries", C.c_uint32), ("dead_pixels", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class NrcuMltParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("mutations_per_pixel", C.c_uint32), ("chains", C.c_uint32), ("n_init", C.c_uint32),
                ("large_step_prob", C.c_float), ("tone_map", C.c_uint32), ("reserved", C.c_uint32)]


MLT_TONE_SQRT, MLT_TONE_REFERENCE, MLT_TONE_LINEAR = 0, 1, 2
UPDATE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32)
_LIB = None


def load_library() -> C.CDLL:
    """Load libnrcuda.so from the package directory; raise if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("NRCU_LIBRARY") or LIB_PATH   # NRCU_LIBRARY: an experiment build (build.py --variant)
    if not os.path.exists(path):
        raise NrcuError(f"{path} is missing: build it with `python -m nrenderer_b200.build` "
                        "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
    L.nrcu_abi_version.restype = i32
    L.nrcu_device_count.restype = i32
    L.nrcu_create.argtypes = [i32, C.POINTER(vp)]
    L.nrcu_destroy.argtypes = [vp]
    L.nrcu_last_error.restype = C.c_char_p
    L.nrcu_last_error.argtypes = [vp]
    L.nrcu_upload_scene.argtypes = [vp, vp, i32]
    L.nrcu_primitive_count.argtypes = [vp, C.POINTER(u32)]
    L.nrcu_download_primitives.argtypes = [vp, vp, vp, vp]
    L.nrcu_render.argtypes = [vp, vp, vp, vp]
    L.nrcu_render_accumulate.argtypes = [vp, vp, vp, vp]
    L.nrcu_render_multi.argtypes = [vp, i32, vp, vp, vp]
    L.nrcu_render_progressive.argtypes = [vp, vp, u32, vp, UPDATE_FN, vp, vp]
    L.nrcu_render_mlt.argtypes = [vp, vp, vp, vp]
    L.nrcu_resolve.argtypes = [vp, vp, vp]
    L.nrcu_trace_batch.argtypes = [vp, vp, u32, vp, vp]
    L.nrcu_host_alloc.restype = vp
    L.nrcu_host_alloc.argtypes = [C.c_size_t]
    L.nrcu_host_free.argtypes = [vp]
    L.nrcu_host_free.restype = None
    L.nrcu_set_stream.argtypes = [vp, vp]
    L.nrcu_synchronize.argtypes = [vp]
    L.nrcu_philox4x32.argtypes = [vp, vp, vp]
    L.nrcu_philox4x32.restype = None
    _LIB = L
    return L


def device_count() -> int:
    return int(load_library().nrcu_device_count())


def philox4x32(counter, key) -> np.ndarray:
    c, k, o = np.asarray(counter, np.uint32), np.asarray(key, np.uint32), np.zeros(4, np.uint32)
    load_library().nrcu_philox4x32(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


class Context:
    """One nrcu_ctx: a CUDA device + stream + the uploaded scene."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.nrcu_create(device, C.byref(h))
        if rc != 0:
            raise NrcuError(f"nrcu_create failed ({rc}): {self._lib.nrcu_last_error(None).decode()}")
        self._h = h
        self.device = device
        self.width = self.height = 0
        self.mode = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nrcu_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            err = NrcuError(f"{what} failed ({rc}): {self._lib.nrcu_last_error(self._h).decode()}")
            err.status = rc
            raise err

    def set_stream(self, cuda_stream_ptr: int):
        self._check(self._lib.nrcu_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "nrcu_set_stream")

    def synchronize(self):
        self._check(self._lib.nrcu_synchronize(self._h), "nrcu_synchronize")

    def upload(self, flat: FlatScene, mode: int):
        view, keep = flat.c_view()
        self._check(self._lib.nrcu_upload_scene(self._h, C.addressof(view), mode), "nrcu_upload_scene")
        del keep
        self.width, self.height, self.mode = flat.width, flat.height, mode
        self.spp = flat.samples_per_pixel

    @property
    def n_primitives(self) -> int:
        n = C.c_uint32(0)
        self._check(self._lib.nrcu_primitive_count(self._h, C.byref(n)), "nrcu_primitive_count")
        return int(n.value)

    def primitives(self):
        n = self.n_primitives
        kind, data, mat = np.zeros(n, np.uint32), np.zeros((n, 16), np.float32), np.zeros(n, np.int32)
        self._check(self._lib.nrcu_download_primitives(self._h, kind.ctypes.data, data.ctypes.data, mat.ctypes.data),
                    "nrcu_download_primitives")
        return kind, data, mat

    @staticmethod
    def _params(seed, s0, s1, glass_mode, samples_per_wave, flags=0, scheduler=0):
        return NrcuRenderParams(seed=seed, sample_begin=s0, sample_end=s1, glass_mode=glass_mode,
                                samples_per_wave=samples_per_wave, flags=flags, scheduler=scheduler)

    def render(self, seed=0, glass_mode=0, samples_per_wave=0, out: np.ndarray | None = None, flags=0, scheduler=0):
        """Whole frame into HOST memory (what Screen::set takes). Returns (rgba[h,w,4], stats dict)."""
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        assert out.dtype == np.float32 and out.size == self.width * self.height * 4 and out.flags.c_contiguous
        p, st = self._params(seed, 0, 0, glass_mode, samples_per_wave, flags, scheduler), NrcuStats()
        self._check(self._lib.nrcu_render(self._h, C.addressof(p), out.ctypes.data, C.addressof(st)), "nrcu_render")
        return out, st.as_dict()

    def render_mlt(self, seed=0, mutations_per_pixel=0, chains=0, n_init=0, large_step_prob=0.0, tone_map=MLT_TONE_SQRT, out: np.ndarray | None = None):
        """nrcu_render_mlt: Metropolis light transport over this backend's path sampler. Returns (rgba[h,w,4], stats dict)."""
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        p, st = NrcuMltParams(seed=seed, mutations_per_pixel=mutations_per_pixel, chains=chains, n_init=n_init,
                              large_step_prob=large_step_prob, tone_map=tone_map), NrcuStats()
        self._check(self._lib.nrcu_render_mlt(self._h, C.addressof(p), out.ctypes.data, C.addressof(st)), "nrcu_render_mlt")
        return out, st.as_dict()

    def render_progressive(self, on_update, samples_per_update=0, seed=0, glass_mode=0, out: np.ndarray | None = None):
        """nrcu_render_progressive: on_update(frame[h,w,4], samples_done, samples_total) -> truthy to stop early."""
        if out is None:
            out = np.empty((self.height, self.width, 4), np.float32)
        p, st = self._params(seed, 0, 0, glass_mode, 0), NrcuStats()
        cb = UPDATE_FN(lambda user, rgba, done, total: int(bool(on_update(out, done, total))))
        self._check(self._lib.nrcu_render_progressive(self._h, C.addressof(p), samples_per_update, out.ctypes.data, cb, None, C.addressof(st)),
                    "nrcu_render_progressive")
        return out, st.as_dict()

    def render_accumulate(self, d_accum_ptr: int, s0=0, s1=0, seed=0, glass_mode=0, samples_per_wave=0, want_stats=True, flags=0, scheduler=0):
        """Add linear sums of samples [s0,s1) into the DEVICE buffer at d_accum_ptr (w*h*4 floats)."""
        p, st = self._params(seed, s0, s1, glass_mode, samples_per_wave, flags, scheduler), NrcuStats()
        self._check(self._lib.nrcu_render_accumulate(self._h, C.addressof(p), C.c_void_p(d_accum_ptr),
                                                     C.addressof(st) if want_stats else None), "nrcu_render_accumulate")
        return st.as_dict() if want_stats else None

    def resolve(self, d_accum_ptr: int, d_rgba_ptr: int):
        self._check(self._lib.nrcu_resolve(self._h, C.c_void_p(d_accum_ptr), C.c_void_p(d_rgba_ptr)), "nrcu_resolve")

    def trace_batch(self, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = len(rays)
        pid, t = np.zeros(n, np.int32), np.zeros(n, np.float32)
        self._check(self._lib.nrcu_trace_batch(self._h, rays.ctypes.data, n, pid.ctypes.data, t.ctypes.data), "nrcu_trace_batch")
        return pid, t


def render_multi(contexts, seed=0, glass_mode=0, samples_per_wave=0, out: np.ndarray | None = None, scheduler=0):
    """One frame on several devices of the box (nrcu_render_multi): contexts[g] must hold the same scene."""
    c0 = contexts[0]
    if out is None:
        out = np.empty((c0.height, c0.width, 4), np.float32)
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    p, st = Context._params(seed, 0, 0, glass_mode, samples_per_wave, 0, scheduler), NrcuStats()
    c0._check(c0._lib.nrcu_render_multi(arr, len(contexts), C.addressof(p), out.ctypes.data, C.addressof(st)), "nrcu_render_multi")
    return out, st.as_dict()
