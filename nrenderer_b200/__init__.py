"""nrenderer_b200 — B200-native (sm_100a) path-tracing backend for NRenderer.

What ships: `csrc/` (hand-written CUDA kernels + the C ABI of include/nrcu.h -> libnrcuda.so),
`plugin/` (the C++ `RenderComponent` adapters registered with REGISTER_RENDERER), `harness/`
(headless driver) and this thin ctypes plumbing used by tests and bench.py.  Nothing here falls
back to the CPU: without libnrcuda.so and a CUDA device every call raises.
"""
from .flatscene import (FlatScene, GLASS_BRANCH, GLASS_STOCHASTIC, MODE_ACC, MODE_RAYCAST, MODE_SIMPLE)  # noqa: F401
from .api import Context, ERR_OVERFLOW, FLAG_ENV_IS, FLAG_KERNEL_TIMES, FLAG_NEE, NrcuError, SCHED_AUTO, SCHED_REGEN, SCHED_WAVES, device_count, load_library, render_multi  # noqa: F401
